// vlg_pass2.cuh -- pass 2: deterministic source gradient (the transpose of the bilinear gather).
//
// torch's grid_sampler backward scatters with float atomics (ATen/native/cuda/GridSampler.cuh
// :250-256), so its d_src changes from run to run.  Here every SOURCE pixel pulls, in a fixed
// row-major order, the d_out of the output pixels whose bilinear footprint covers it:
//   near path: output pixels displaced by < VLG_NEAR_RADIUS px are found by scanning the
//              (2r+1)^2 window around the source pixel in shared memory -- no atomics at all;
//   far path : the (rare) output pixels displaced further are accumulated with 64-bit
//              fixed-point integer atomics, which are associative and therefore order-free.
#pragma once
#include "vlg_device.cuh"
#include "vlg_pass1.cuh"  // source_xy / taps_from_xy

namespace vlg {

constexpr int kRMax = VLG_NEAR_RADIUS;
constexpr int kQW = kTW + 2 * kRMax, kQH = kTH + 2 * kRMax, kQN = kQW * kQH;

struct Pass2Params {
    CoordCfg cc;
    int N, tiles_x, tiles_y;
    const float *coords;
    const float *d_out_rgb;   // fp32 staging written by pass 1
    const float *d_out_lay;
    void *d_src_rgb;          // type T, nullable
    void *d_src_lay;
    long long *far_acc;       // [P][3+K] fixed point, nullable (VLG_FLAG_NO_FAR_PATH)
    const float *tile_disp;   // [n_blocks] per-tile max NEAR displacement written by pass 1
    WsHeader *hdr;
    int64_t HW;
};

// Gather radius of one source tile: output pixels that can reach it lie within kRMax pixels, i.e.
// inside the 3x3 neighbourhood of tiles (kRMax < kTH); far pixels (disp >= kRMax) are excluded from
// the per-tile maxima and travel through the fixed-point path instead.
__device__ __forceinline__ int near_radius(const Pass2Params &p, int n, int tyi, int txi) {
    float m = 0.f;
    for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
            const int yy = tyi + dy, xx = txi + dx;
            if (yy >= 0 && yy < p.tiles_y && xx >= 0 && xx < p.tiles_x)
                m = fmaxf(m, __ldg(p.tile_disp + ((int64_t)n * p.tiles_y + yy) * p.tiles_x + xx));
        }
    return min(kRMax, (int)floorf(m) + 1);
}

// 2^e such that (sum of <= H*W contributions of magnitude <= maxgrad) * 2^e < 2^62
__device__ __forceinline__ double far_scale(const WsHeader *hdr, int64_t HW) {
    const float g = __uint_as_float(hdr->maxgrad_bits);
    int eg = 0, ehw = 0;
    frexpf(fmaxf(g, 1e-37f), &eg);
    frexp((double)HW, &ehw);
    return ldexp(1.0, 61 - eg - ehw);
}

template <int K>
constexpr size_t pass2_smem_bytes() {
    return (size_t)kQN * (sizeof(float2) + sizeof(float) * (3 + K));
}

template <typename T, int K>
__global__ void __launch_bounds__(kThreads) pass2_kernel(const Pass2Params p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2 *s_xy = reinterpret_cast<float2 *>(smem_raw);               // [kQN] source coords (NaN = skip)
    float *s_lay = reinterpret_cast<float *>(s_xy + kQN);             // [kQN][K]
    float *s_rgb = s_lay + (size_t)kQN * K;                            // [kQN][3]

    const CoordCfg &cc = p.cc;
    const int H = cc.H, W = cc.W;
    const int tid = threadIdx.x;
    const int bt = blockIdx.x;
    const int n = bt / (p.tiles_x * p.tiles_y);
    const int trem = bt - n * (p.tiles_x * p.tiles_y);
    const int ty0 = (trem / p.tiles_x) * kTH, tx0 = (trem % p.tiles_x) * kTW;
    const int64_t img_px = (int64_t)n * H * W;

    const bool has_far = __uint_as_float(p.hdr->maxdisp_bits) >= (float)kRMax;
    const int r = near_radius(p, n, trem / p.tiles_x, trem % p.tiles_x);
    const int qw = kTW + 2 * r, qh = kTH + 2 * r, qn = qw * qh;
    const bool want_rgb = p.d_src_rgb != nullptr && p.d_out_rgb != nullptr;
    const bool want_lay = p.d_src_lay != nullptr && p.d_out_lay != nullptr;

    // ---- stage coords and d_out of the candidate region ----
    __shared__ float s_bx[kQW], s_by[kQH];   // base-grid table: one IEEE division per row / column
    if (tid < qw) s_bx[tid] = base_coord(tx0 - r + tid, cc.Wm1);
    else if (tid >= 64 && tid < 64 + qh) s_by[tid - 64] = base_coord(ty0 - r + tid - 64, cc.Hm1);
    __syncthreads();
    const float2 *coords = reinterpret_cast<const float2 *>(p.coords) + img_px;
    const float qnan = __int_as_float(0x7fc00000);
    for (int q = tid; q < qn; q += kThreads) {
        const int ry = q / qw, rx = q - ry * qw;
        const int y = ty0 - r + ry, x = tx0 - r + rx;
        float2 xy = make_float2(qnan, qnan);
        if (y >= 0 && y < H && x >= 0 && x < W) {
            float mx, my;
            const float2 s = source_xy(cc, __ldg(coords + (int64_t)y * W + x), s_bx[rx], s_by[ry], mx, my);
            const float fx0 = floorf(s.x), fy0 = floorf(s.y);
            const bool dead = fx0 < -1.0f || fx0 >= (float)W || fy0 < -1.0f || fy0 >= (float)H;  // no tap inside
            const bool far = !dead && fmaxf(fabsf(s.x - (float)x), fabsf(s.y - (float)y)) >= (float)kRMax;
            if (!far && !dead) xy = s;
        }
        s_xy[q] = xy;
    }
    if (want_lay) {
        constexpr int V = K % 4 == 0 ? 4 : 1;  // floats per staged vector
        constexpr int VPP = K / V;             // vectors per pixel
        for (int i = tid; i < qn * VPP; i += kThreads) {
            const int q = i / VPP, v = i - q * VPP;
            const int ry = q / qw, rx = q - ry * qw;
            const int y = ty0 - r + ry, x = tx0 - r + rx;
            if (y >= 0 && y < H && x >= 0 && x < W) {
                const float *g = p.d_out_lay + (img_px + (int64_t)y * W + x) * K + v * V;
                if constexpr (V == 4)
                    *reinterpret_cast<float4 *>(s_lay + (size_t)q * K + v * 4) = __ldg(reinterpret_cast<const float4 *>(g));
                else
                    s_lay[(size_t)q * K + v] = __ldg(g);
            }
        }
    }
    if (want_rgb) {
        for (int i = tid; i < qn * 3; i += kThreads) {
            const int q = i / 3, c = i - q * 3;
            const int ry = q / qw, rx = q - ry * qw;
            const int y = ty0 - r + ry, x = tx0 - r + rx;
            if (y >= 0 && y < H && x >= 0 && x < W)
                s_rgb[q * 3 + c] = __ldg(p.d_out_rgb + (img_px + (int64_t)y * W + x) * 3 + c);
        }
    }
    __syncthreads();

    // ---- one thread per source pixel: fixed-order gather ----
    const int ty = tid / kTW, tx = tid - ty * kTW;
    const int sy = ty0 + ty, sx = tx0 + tx;
    if (sy >= H || sx >= W) return;
    float acc_l[K], acc_r[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < K; ++k) acc_l[k] = 0.f;
    const float fsx = (float)sx, fsy = (float)sy;
    for (int dy = -r; dy <= r; ++dy) {
        const int qrow = (ty + r + dy) * qw + tx + r;
        for (int dx = -r; dx <= r; ++dx) {
            const int q = qrow + dx;
            const float2 xy = s_xy[q];
            const float fx0 = floorf(xy.x), fy0 = floorf(xy.y);
            const float ax = fsx - fx0, ay = fsy - fy0;   // 0 -> west/north tap, 1 -> east/south tap
            const bool hit = (ax == 0.0f || ax == 1.0f) && (ay == 0.0f || ay == 1.0f);
            if (!hit) continue;
            const float wx = ax == 0.0f ? __fsub_rn(__fadd_rn(fx0, 1.0f), xy.x) : __fsub_rn(xy.x, fx0);
            const float wy = ay == 0.0f ? __fsub_rn(__fadd_rn(fy0, 1.0f), xy.y) : __fsub_rn(xy.y, fy0);
            const float w = __fmul_rn(wx, wy);
            if (want_lay) {
                const float *d = s_lay + (size_t)q * K;
                if constexpr (K % 4 == 0) {
#pragma unroll
                    for (int v = 0; v < K / 4; ++v) {
                        const float4 dv = *reinterpret_cast<const float4 *>(d + 4 * v);
                        acc_l[4 * v + 0] = fmaf(w, dv.x, acc_l[4 * v + 0]);
                        acc_l[4 * v + 1] = fmaf(w, dv.y, acc_l[4 * v + 1]);
                        acc_l[4 * v + 2] = fmaf(w, dv.z, acc_l[4 * v + 2]);
                        acc_l[4 * v + 3] = fmaf(w, dv.w, acc_l[4 * v + 3]);
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < K; ++k) acc_l[k] = fmaf(w, d[k], acc_l[k]);
                }
            }
            if (want_rgb) {
#pragma unroll
                for (int c = 0; c < 3; ++c) acc_r[c] = fmaf(w, s_rgb[q * 3 + c], acc_r[c]);
            }
        }
    }
    const int64_t so = img_px + (int64_t)sy * W + sx;
    if (has_far && p.far_acc) {
        const double inv = 1.0 / far_scale(p.hdr, p.HW);
        const long long *fa = p.far_acc + so * (3 + K);
#pragma unroll
        for (int c = 0; c < 3; ++c) acc_r[c] += (float)((double)fa[c] * inv);
#pragma unroll
        for (int k = 0; k < K; ++k) acc_l[k] += (float)((double)fa[3 + k] * inv);
    }
    if (want_rgb) store_px<T, 3>(reinterpret_cast<T *>(p.d_src_rgb) + so * 3, acc_r);
    if (want_lay) store_px<T, K>(reinterpret_cast<T *>(p.d_src_lay) + so * K, acc_l);
}

// ---- far path: zero the fixed-point accumulators (only when far pixels exist) ----
__global__ void far_zero_kernel(long long *acc, int64_t n_words, const WsHeader *hdr, uint32_t flags,
                                WsHeader *hdr_rw) {
    const float md = __uint_as_float(hdr->maxdisp_bits);
    if (md < (float)kRMax) return;
    if (flags & VLG_FLAG_NO_FAR_PATH) {
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(&hdr_rw->status, VLG_STATUS_FAR_TAPS);
        return;
    }
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    longlong2 *a2 = reinterpret_cast<longlong2 *>(acc);
    const int64_t n2 = n_words / 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride)
        a2[i] = make_longlong2(0, 0);
    if (blockIdx.x == 0 && threadIdx.x == 0 && (n_words & 1)) acc[n_words - 1] = 0;
}

// ---- far path: fixed-point scatter of the far output pixels ----
template <int K>
__global__ void far_scatter_kernel(const Pass2Params p, int64_t P) {
    const float md = __uint_as_float(p.hdr->maxdisp_bits);
    if (md < (float)kRMax || p.far_acc == nullptr) return;
    const CoordCfg &cc = p.cc;
    const int H = cc.H, W = cc.W;
    const double scale = far_scale(p.hdr, p.HW);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const float2 *coords = reinterpret_cast<const float2 *>(p.coords);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += stride) {
        const int64_t n = i / p.HW;
        const int64_t rem = i - n * p.HW;
        const int y = (int)(rem / W), x = (int)(rem - (int64_t)y * W);
        const Taps t = make_taps(cc, __ldg(coords + i), y, x);
        if (!(tap_displacement(cc, t, y, x) >= (float)kRMax)) continue;
        atomicAdd(&p.hdr->far_count, 1u);
        const int xs[4] = {t.x0, t.x0 + 1, t.x0, t.x0 + 1};
        const int ys[4] = {t.y0, t.y0, t.y0 + 1, t.y0 + 1};
        const float ws[4] = {t.nw, t.ne, t.sw, t.se};
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {
            if (xs[k4] < 0 || xs[k4] >= W || ys[k4] < 0 || ys[k4] >= H) continue;
            unsigned long long *dst = reinterpret_cast<unsigned long long *>(
                p.far_acc + (n * p.HW + (int64_t)ys[k4] * W + xs[k4]) * (3 + K));
            if (p.d_out_rgb)
                for (int c = 0; c < 3; ++c) {
                    const long long v = __double2ll_rn((double)__fmul_rn(ws[k4], __ldg(p.d_out_rgb + i * 3 + c)) * scale);
                    atomicAdd(dst + c, (unsigned long long)v);
                }
            if (p.d_out_lay)
                for (int c = 0; c < K; ++c) {
                    const long long v = __double2ll_rn((double)__fmul_rn(ws[k4], __ldg(p.d_out_lay + i * K + c)) * scale);
                    atomicAdd(dst + 3 + c, (unsigned long long)v);
                }
        }
    }
}

}  // namespace vlg
