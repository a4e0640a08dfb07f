// vlg_pass2.cuh -- pass 2: deterministic source gradient (the transpose of the bilinear gather).
//
// torch's grid_sampler backward scatters with float atomics (ATen/native/cuda/GridSampler.cuh
// :250-256), so its d_src changes from run to run.  Here every SOURCE pixel pulls, in a fixed
// row-major order, the d_out of the output pixels whose bilinear footprint covers it:
//   near path: output pixels displaced by < VLG_NEAR_RADIUS px are found by scanning the
//              (2r+1)^2 window around the source pixel in shared memory -- no atomics at all;
//              r is chosen per tile from the displacement maxima pass 1 recorded;
//   far path : the (rare) output pixels displaced further were queued by pass 1; they are
//              accumulated with 64-bit fixed-point integer atomics (associative, hence
//              order-free) into the source tiles pass 1 flagged, and only those tiles are
//              zeroed and read back.
#pragma once
#include "vlg_device.cuh"
#include "vlg_pass1.cuh"  // source_xy, cp_async16

namespace vlg {

constexpr int kRMax = VLG_NEAR_RADIUS;
constexpr int kQW = kTW + 2 * kRMax, kQH = kTH + 2 * kRMax, kQN = kQW * kQH;

struct Pass2Params {
    CoordCfg cc;
    int N, tiles_x, tiles_y;
    const float *coords;
    const float *d_out_rgb;   // fp32 staging written by pass 1: [N][H][pitch][3]
    const float *d_out_lay;   // [N][H][W][K]
    const uint32_t *rec_code; // [N][H][pitch] tap records written by pass 1 (tap_cell_code)
    const float2 *rec_frac;   // [N][H][pitch]
    int pitch;                // row pitch (pixels) of d_out_rgb / rec_code / rec_frac
    void *d_src_rgb;          // type T, nullable
    void *d_src_lay;
    long long *far_acc;       // [P][3+K] fixed point, nullable (VLG_FLAG_NO_FAR_PATH)
    const int4 *far_list;     // [far_count] far output pixels queued by pass 1: {pixel index, (x0 + 8) | (y0 + 8) << 16, bits of ix - x0, bits of iy - y0}
    const uint32_t *tile_flags;  // [n_blocks] source tiles that receive far contributions
    const int *flagged_list;  // [n_flagged] ids of those tiles
    const float *tile_disp;   // [n_blocks] per-tile max NEAR displacement written by pass 1
    WsHeader *hdr;
    int64_t HW;
    int use_tma;              // d_out_lay window staged by one cp.async.bulk.tensor per CTA
};

// Gather radius of one source tile: output pixels that can reach it lie within kRMax pixels, i.e.
// inside the 3x3 neighbourhood of tiles (kRMax < kTH); far pixels (disp >= kRMax) are excluded from
// the per-tile maxima and travel through the fixed-point path instead.
__device__ __forceinline__ int near_radius(const Pass2Params &p, int n, int tyi, int txi) {
    // nine independent loads; neighbours outside the image are clamped onto an existing tile so that
    // no load is predicated (the maximum is unaffected) and all nine are in flight together
    float v[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) {
        const int yy = min(max(tyi + j / 3 - 1, 0), p.tiles_y - 1), xx = min(max(txi + j % 3 - 1, 0), p.tiles_x - 1);
        v[j] = __ldg(p.tile_disp + ((int64_t)n * p.tiles_y + yy) * p.tiles_x + xx);
    }
    float m = 0.f;
#pragma unroll
    for (int j = 0; j < 9; ++j) m = fmaxf(m, v[j]);
    return min(kRMax, (int)floorf(m) + 1);
}

// 2^e such that (sum of <= H*W contributions of magnitude <= maxgrad) * 2^e < 2^62
__device__ __forceinline__ int far_scale_exp(const WsHeader *hdr, int64_t HW) {
    const float g = __uint_as_float(hdr->maxgrad_bits);
    int eg = 0, ehw = 0;
    frexpf(fmaxf(g, 1e-37f), &eg);
    frexp((double)HW, &ehw);
    return min(61 - eg - ehw, 96);   // capped so that 2^e is an fp32 number too (gradients below 2^-96 are noise anyway)
}
__device__ __forceinline__ double far_scale(const WsHeader *hdr, int64_t HW) { return ldexp(1.0, far_scale_exp(hdr, HW)); }

template <int K>
constexpr size_t pass2_smem_bytes() {
    return (size_t)kQN * (sizeof(float2) + sizeof(uint32_t) + sizeof(float) * (3 + K));
}

template <typename T, int K>
__global__ void __launch_bounds__(kThreads, 4) pass2_kernel(const Pass2Params p, const __grid_constant__ CUtensorMap dout_map) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *s_lay = reinterpret_cast<float *>(smem_raw);                 // [kQN][K]   (16-byte aligned rows)
    float2 *s_xy = reinterpret_cast<float2 *>(s_lay + (size_t)kQN * K); // [kQN] source coords (NaN = skip)
    float *s_rgb = reinterpret_cast<float *>(s_xy + kQN);              // [kQN][3]
    __shared__ float s_bx[kQW], s_by[kQH];   // base-grid table: one IEEE division per row / column
    __shared__ alignas(8) uint64_t s_bar;    // mbarrier of the TMA window load

    const CoordCfg &cc = p.cc;
    const int H = cc.H, W = cc.W;
    const int tid = threadIdx.x;
    const int lane = tid & 31, wid = tid >> 5;
    const int n = blockIdx.z;
    const int bt = (n * p.tiles_y + blockIdx.y) * p.tiles_x + blockIdx.x;
    const int ty0 = blockIdx.y * kTH, tx0 = blockIdx.x * kTW;
    const int64_t img_px = (int64_t)n * H * W;

    // TMA first: the fixed window does not depend on anything loaded from memory
    if (p.use_tma && p.d_src_lay != nullptr && p.d_out_lay != nullptr && tid == 0) {
        mbar_init(&s_bar, 1);
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        mbar_expect_tx(&s_bar, (unsigned)(kQN * K * sizeof(float)));
        tma_load_4d(s_lay, &dout_map, &s_bar, 0, tx0 - kRMax, ty0 - kRMax, n);
    }
    const uint32_t tile_far = p.far_acc ? __ldg(p.tile_flags + bt) : 0u;   // consumed at the very end: issue early
    const int r = near_radius(p, n, blockIdx.y, blockIdx.x);
    const int qw = kTW + 2 * r, qh = kTH + 2 * r;
    const bool want_rgb = p.d_src_rgb != nullptr && p.d_out_rgb != nullptr;
    const bool want_lay = p.d_src_lay != nullptr && p.d_out_lay != nullptr;
    const int xa = max(tx0 - r, 0), xb = min(tx0 - r + qw, W);   // in-image columns of the candidate region

    // ---- stage d_out of the candidate region ----
    // The layout part always lives in a FIXED window of kQW x kQH pixels anchored at
    // (tx0 - kRMax, ty0 - kRMax): with TMA it is one cp.async.bulk.tensor issued before anything else
    // (no dependence on the per-tile radius), otherwise cp.async row copies of the needed sub-window.
    const int lay_off = (kRMax - r) * kQW + (kRMax - r);   // window cell of region pixel (0, 0)
    for (int ry = wid; ry < qh; ry += kThreads / 32) {
        const int y = ty0 - r + ry;
        if (y < 0 || y >= H || xb <= xa) continue;
        const int64_t g0 = img_px + (int64_t)y * W + xa;
        const int q0 = ry * qw + (xa - (tx0 - r));
        if (want_lay && !p.use_tma) {
            const int c0 = lay_off + ry * kQW + (xa - (tx0 - r));
            const char *src = reinterpret_cast<const char *>(p.d_out_lay + g0 * K);
            char *dst = reinterpret_cast<char *>(s_lay + (size_t)c0 * K);
            const int nbytes = (xb - xa) * K * 4;
            if constexpr (K % 4 == 0) {
                for (int i = lane * 16; i < nbytes; i += 32 * 16) cp_async16(dst + i, src + i);
            } else {
                for (int i = lane; i < (xb - xa) * K; i += 32) s_lay[(size_t)c0 * K + i] = __ldg(p.d_out_lay + g0 * K + i);
            }
        }
        if (want_rgb) {   // 12-byte pixels: 4-byte cp.async (no register round trip, completes with the group)
            const float *rsrc = p.d_out_rgb + (((int64_t)n * H + y) * p.pitch + xa) * 3;
            for (int i = lane; i < (xb - xa) * 3; i += 32) cp_async4(s_rgb + q0 * 3 + i, rsrc + i);
        }
    }
    // ---- sampling coordinates of the candidate output pixels ----
    // flat index over the region, addresses clamped into the image: the (up to three) coords loads of
    // a thread are unconditional and issued before the barrier, together with the cp.async above
    const float2 *coords = reinterpret_cast<const float2 *>(p.coords) + img_px;
    const int qn = qw * qh;
    const float inv_qw = 1.0f / (float)qw;
    constexpr int kQIt = (kQN + kThreads - 1) / kThreads;
    float2 cf[kQIt];
    int cry[kQIt], crx[kQIt];
#pragma unroll
    for (int it = 0; it < kQIt; ++it) {
        const int q = min(tid + it * kThreads, qn - 1);
        const int ry = (int)(((float)q + 0.5f) * inv_qw), rx = q - ry * qw;   // exact for these small integers
        cry[it] = ry; crx[it] = rx;
        const int y = min(max(ty0 - r + ry, 0), H - 1), x = min(max(tx0 - r + rx, 0), W - 1);
        cf[it] = __ldg(coords + (int64_t)y * W + x);
    }
    if (tid < qw) s_bx[tid] = base_coord(tx0 - r + tid, cc.Wm1);
    else if (tid >= 64 && tid < 64 + qh) s_by[tid - 64] = base_coord(ty0 - r + tid - 64, cc.Hm1);
    __syncthreads();
    // Each candidate output pixel is reduced to a packed integer cell code (x0 | y0 << 16, relative
    // to the tile, biased by 8) plus its two fractional weights, so that the per-candidate test in
    // the gather below is three integer instructions.
    uint32_t *s_code = reinterpret_cast<uint32_t *>(s_rgb + (size_t)kQN * 3);   // [kQN]
#pragma unroll
    for (int it = 0; it < kQIt; ++it) {
        const int q = tid + it * kThreads;
        if (q < qn) {
            const int ry = cry[it], rx = crx[it];
            const int y = ty0 - r + ry, x = tx0 - r + rx;
            float mx, my;
            const float2 sxy = source_xy(cc, cf[it], s_bx[rx], s_by[ry], mx, my);
            const float fx0 = floorf(sxy.x), fy0 = floorf(sxy.y);
            const bool in_img = y >= 0 && y < H && x >= 0 && x < W;
            const bool dead = fx0 < -1.0f || fx0 >= (float)W || fy0 < -1.0f || fy0 >= (float)H;  // no tap inside
            const bool far = fmaxf(fabsf(sxy.x - (float)x), fabsf(sxy.y - (float)y)) >= (float)kRMax;
            const bool ok = in_img && !far && !dead;
            // near => |x0 - x| <= kRMax, so the biased fields stay within [0, 2^15); 0xFFFFFFFF never matches
            s_code[q] = ok ? ((uint32_t)((int)fx0 - tx0 + 8) | ((uint32_t)((int)fy0 - ty0 + 8) << 16)) : 0xFFFFFFFFu;
            s_xy[q] = ok ? make_float2(__fsub_rn(sxy.x, fx0), __fsub_rn(sxy.y, fy0)) : make_float2(0.f, 0.f);
        }
    }
    cp_async_commit_wait_all();
    if (want_lay && p.use_tma) mbar_wait(&s_bar, 0);
    __syncthreads();

    // ---- one thread per source pixel: fixed-order gather ----
    const int ty = tid / kTW, tx = tid - ty * kTW;
    const int sy = ty0 + ty, sx = tx0 + tx;
    const bool live = sy < H && sx < W;
    float acc_l[K], acc_r[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < K; ++k) acc_l[k] = 0.f;
    const uint32_t mycode = (uint32_t)(tx + 8) | ((uint32_t)(ty + 8) << 16);
    for (int dy = live ? -r : r + 1; dy <= r; ++dy) {
        const int qrow = (ty + r + dy) * qw + tx;   // candidate dx = -r sits at qrow, dx = +r at qrow + 2r
        // e = (sx - x0) + ((sy - y0) << 16): a hit iff both differences are 0 (west/north tap)
        // or 1 (east/south tap); any other difference (incl. borrows) leaves a bit outside {0,16}.
        // All codes of the row are fetched first (independent LDS) -- the kernel is latency-bound.
        uint32_t ev[2 * kRMax + 1];
#pragma unroll
        for (int j = 0; j < 2 * kRMax + 1; ++j) ev[j] = (j <= 2 * r) ? mycode - s_code[qrow + j] : 0xFFFFFFFFu;
#pragma unroll
        for (int j = 0; j < 2 * kRMax + 1; ++j) {
            const uint32_t e = ev[j];
            if (e & 0xFFFEFFFEu) continue;
            const int q = qrow + j;
            const float2 f = s_xy[q];
            // tap weights: east/south = frac, west/north = 1 - frac, which is bit-identical to the
            // forward's (x0 + 1) - ix for every in-image tap (both are exact for x0 >= 1, and the
            // same expression for x0 == 0)
            const float wx = (e & 1u) ? f.x : __fsub_rn(1.0f, f.x);
            const float wy = (e >> 16) ? f.y : __fsub_rn(1.0f, f.y);
            const float w = __fmul_rn(wx, wy);
            if (want_lay) {
                float v[K];
                load_px_smem<float, K>(s_lay + (size_t)(lay_off + (ty + r + dy) * kQW + tx + j) * K, v);
                fma2_bcast<K>(acc_l, v, w);
            }
            if (want_rgb) {
#pragma unroll
                for (int c = 0; c < 3; ++c) acc_r[c] = fmaf(w, s_rgb[q * 3 + c], acc_r[c]);
            }
        }
    }
    const int64_t so = img_px + (int64_t)sy * W + sx;
    if (live && tile_far) {
        const double inv = 1.0 / far_scale(p.hdr, p.HW);
        const long long *fa = p.far_acc + so * (3 + K);
#pragma unroll
        for (int c = 0; c < 3; ++c) acc_r[c] += (float)((double)fa[c] * inv);
#pragma unroll
        for (int k = 0; k < K; ++k) acc_l[k] += (float)((double)fa[3 + k] * inv);
    }
    if (live && want_rgb) store_px<T, 3>(reinterpret_cast<T *>(p.d_src_rgb) + so * 3, acc_r);
    if (want_lay) {
        // A thread-per-pixel store of 80-byte pixels costs 20 L1 wavefronts per 128-bit store
        // instruction; transposing through shared memory lets each warp write its tile row
        // (32 px x K, contiguous in HBM) as consecutive 16-byte words: 4 wavefronts each.
        __syncthreads();                                   // all gathers done: the d_out staging can be reused
        T *s_out = reinterpret_cast<T *>(s_lay);           // [kThreads][K]
        if (live) store_px<T, K>(s_out + (size_t)tid * K, acc_l);
        __syncwarp();
        const int sy_w = ty0 + wid;                        // warp w owns tile row w (kTW == 32)
        if (sy_w < H) {
            const int npx = min(kTW, W - tx0);
            T *grow = reinterpret_cast<T *>(p.d_src_lay) + (img_px + (int64_t)sy_w * W + tx0) * K;
            const T *srow = s_out + (size_t)wid * kTW * K;
            if constexpr ((K * sizeof(T)) % 16 == 0) {
                const int nvec = npx * (int)(K * sizeof(T) / 16);
                for (int i = lane; i < nvec; i += 32)
                    reinterpret_cast<uint4 *>(grow)[i] = reinterpret_cast<const uint4 *>(srow)[i];
            } else {
                for (int i = lane; i < npx * K; i += 32) grow[i] = srow[i];
            }
        }
    }
}


// ---- pass 2 from tap records: everything the gather needs arrives by TMA ----
// pass2_kernel above re-derives, per CTA, the tap cell and weights of all 532 candidate output pixels
// from the coordinates (source_xy -> floor -> code) and copies the rgb part of d_out row by row: 28 %
// of its instructions and two of its three barriers.  Here pass 1 has already written a 12-byte record
// per output pixel (cell code + the two fractional weights, tap_cell_code), rows padded to a pitch of
// 4 pixels, so the CTA's whole input is four TMA tensor loads on one mbarrier:
//   d_out_lay window  [kQH][kQW][K]      (as before)
//   d_out_rgb window  [kQH][kQW2 * 3]    fp32
//   fraction window   [kQH][kQW2]        float2
//   code window       [kQH][kQW2]        uint32 -- out-of-image cells arrive as 0 = "no contribution"
// the first anchored at (tx0 - kRMax, ty0 - kRMax), the narrow-pixel ones at (tx0 - kQX2, ty0 - kRMax): a TMA box must
// start on a 16-byte boundary, i.e. on a column that is a multiple of 4 for the 12-, 8- and 4-byte pixels.  The scan / hit loop is the same as pass2_kernel's, in the
// same order, so both kernels produce bit-identical gradients.
constexpr int kQW2 = 40;   // window width of the 12-, 8- and 4-byte-pixel arrays: rows of 480 / 320 / 160 bytes
constexpr int kQX2 = 4;    // their left margin (>= kRMax, multiple of 4; kTW is a multiple of 4 too)
static_assert(kQX2 >= kRMax && kQX2 % 4 == 0 && kTW % 4 == 0 && kQX2 + kTW + kRMax <= kQW2, "narrow-pixel window geometry");

template <int K>
struct Pass2RecSmem {
    alignas(128) float lay[kQN * K];
    alignas(128) float rgb[kQH * kQW2 * 3];
    alignas(128) float2 frac[kQH * kQW2];
    alignas(128) uint32_t code[kQH * kQW2];
    alignas(8) uint64_t bar;
};

template <typename T, int K>
__global__ void __launch_bounds__(kThreads, 4) pass2_rec_kernel(const Pass2Params p, const __grid_constant__ CUtensorMap lay_map,
                                                                const __grid_constant__ CUtensorMap rgb_map,
                                                                const __grid_constant__ CUtensorMap frac_map,
                                                                const __grid_constant__ CUtensorMap code_map) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Pass2RecSmem<K> &sm = *reinterpret_cast<Pass2RecSmem<K> *>(smem_raw);
    const int H = p.cc.H, W = p.cc.W;
    const int tid = threadIdx.x;
    const int lane = tid & 31, wid = tid >> 5;
    const int n = blockIdx.z;
    const int bt = (n * p.tiles_y + blockIdx.y) * p.tiles_x + blockIdx.x;
    const int ty0 = blockIdx.y * kTH, tx0 = blockIdx.x * kTW;
    const int64_t img_px = (int64_t)n * H * W;
    const bool want_rgb = p.d_src_rgb != nullptr && p.d_out_rgb != nullptr;
    const bool want_lay = p.d_src_lay != nullptr && p.d_out_lay != nullptr;

    if (tid == 0) {
        tma_prefetch_desc(&code_map); tma_prefetch_desc(&frac_map); tma_prefetch_desc(&lay_map); tma_prefetch_desc(&rgb_map);
        mbar_init(&sm.bar, 1);
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        unsigned bytes = (unsigned)(sizeof(sm.frac) + sizeof(sm.code));
        if (want_lay) bytes += (unsigned)sizeof(sm.lay);
        if (want_rgb) bytes += (unsigned)sizeof(sm.rgb);
        mbar_expect_tx(&sm.bar, bytes);
        tma_load_3d(sm.code, &code_map, &sm.bar, tx0 - kQX2, ty0 - kRMax, n);
        tma_load_3d(sm.frac, &frac_map, &sm.bar, (tx0 - kQX2) * 2, ty0 - kRMax, n);
        if (want_lay) tma_load_4d(sm.lay, &lay_map, &sm.bar, 0, tx0 - kRMax, ty0 - kRMax, n);
        if (want_rgb) tma_load_3d(sm.rgb, &rgb_map, &sm.bar, (tx0 - kQX2) * 3, ty0 - kRMax, n);
    }
    const uint32_t tile_far = p.far_acc ? __ldg(p.tile_flags + bt) : 0u;   // consumed at the very end: issue early
    const int r = near_radius(p, n, blockIdx.y, blockIdx.x);
    __syncthreads();          // the barrier object is initialised
    mbar_wait(&sm.bar, 0);    // the four windows have landed

    // ---- one thread per source pixel: fixed-order gather ----
    // A warp owns a compact 8 x 4 PATCH of the tile, not a 32-pixel row: the hit body below runs once per distinct
    // hit offset among the warp's pixels, and the flow varies less across a patch than along a row (fewer, fuller
    // bodies).  A quarter-warp is still 8 consecutive pixels of one row, so the 80-byte-pixel fetches stay
    // conflict-free.
    const int tx = (wid & 3) * 8 + (lane & 7), ty = (wid >> 2) * 4 + (lane >> 3);
    static_assert(kTW == 32 && kTH == 8, "patch mapping");
    const int sy = ty0 + ty, sx = tx0 + tx;
    const bool live = sy < H && sx < W;
    float acc_l[K], acc_r[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < K; ++k) acc_l[k] = 0.f;
    for (int dy = live ? -r : r + 1; dy <= r; ++dy) {
        // candidate output pixel (sx + ddx, sy + dy), ddx = j - r, sits at window cell (ty + kRMax + dy, tx + kQX2 + ddx).
        // With its record code = (x0 - x + 8) | (y0 - y + 8) << 16,
        //   e = ((8 - ddx) | (8 - dy) << 16) - code = (sx - x0) | (sy - y0) << 16:
        // a hit iff both differences are 0 (west / north tap) or 1 (east / south tap); any other difference
        // (incl. borrows, and code 0) leaves a bit outside {0, 16}.
        const int crow = (ty + kRMax + dy) * kQW2 + tx + kQX2 - r;
        const uint32_t c0 = (uint32_t)(8 + r) | ((uint32_t)(8 - dy) << 16);
        uint32_t ev[2 * kRMax + 1];
#pragma unroll
        for (int j = 0; j < 2 * kRMax + 1; ++j) ev[j] = (j <= 2 * r) ? (c0 - (uint32_t)j) - sm.code[crow + j] : 0xFFFFFFFFu;
#pragma unroll
        for (int j = 0; j < 2 * kRMax + 1; ++j) {
            const uint32_t e = ev[j];
            if (e & 0xFFFEFFFEu) continue;
            const int cell = crow + j;
            const float2 f = sm.frac[cell];
            // east/south = frac, west/north = 1 - frac: bit-identical to the forward's (x0 + 1) - ix for every
            // in-image tap (both are exact for x0 >= 1, and the same expression for x0 == 0)
            const float wx = (e & 1u) ? f.x : __fsub_rn(1.0f, f.x);
            const float wy = (e >> 16) ? f.y : __fsub_rn(1.0f, f.y);
            const float w = __fmul_rn(wx, wy);
            if (want_lay) {
                float v[K];
                load_px_smem<float, K>(sm.lay + (size_t)((ty + kRMax + dy) * kQW + tx + kRMax - r + j) * K, v);
                fma2_bcast<K>(acc_l, v, w);
            }
            if (want_rgb) {
#pragma unroll
                for (int c = 0; c < 3; ++c) acc_r[c] = fmaf(w, sm.rgb[cell * 3 + c], acc_r[c]);
            }
        }
    }
    const int64_t so = img_px + (int64_t)sy * W + sx;
    if (live && tile_far) {
        const double inv = 1.0 / far_scale(p.hdr, p.HW);
        const long long *fa = p.far_acc + so * (3 + K);
#pragma unroll
        for (int c = 0; c < 3; ++c) acc_r[c] += (float)((double)fa[c] * inv);
#pragma unroll
        for (int k = 0; k < K; ++k) acc_l[k] += (float)((double)fa[3 + k] * inv);
    }
    if (live && want_rgb) store_px<T, 3>(reinterpret_cast<T *>(p.d_src_rgb) + so * 3, acc_r);
    if (want_lay) {
        // transpose through shared memory: each warp writes its tile row (32 px x K, contiguous in HBM)
        // as consecutive 16-byte words
        __syncthreads();                                   // all gathers done: the d_out staging can be reused
        T *s_out = reinterpret_cast<T *>(sm.lay);          // [kThreads][K]
        if (live) store_px<T, K>(s_out + (size_t)(ty * kTW + tx) * K, acc_l);
        fence_async_smem();                                // the rows leave through the async proxy
        __syncthreads();                                   // a row holds pixels of four warps (patch mapping)
        const int sy_w = ty0 + wid;                        // warp w owns tile row w (kTW == 32)
        if (sy_w < H) {
            const int npx = min(kTW, W - tx0);
            T *grow = reinterpret_cast<T *>(p.d_src_lay) + (img_px + (int64_t)sy_w * W + tx0) * K;
            const T *srow = s_out + (size_t)wid * kTW * K;
            const unsigned row_bytes = (unsigned)(npx * K * (int)sizeof(T));
            if (row_bytes % 16 == 0 && (reinterpret_cast<uintptr_t>(grow) & 15) == 0) {
                // one bulk shared -> global copy per row instead of 32 lanes x (LDS.128 + STG.128)
                if (lane == 0) { bulk_store(grow, srow, row_bytes); bulk_store_wait_read(); }
            } else if constexpr ((K * sizeof(T)) % 16 == 0) {
                const int nvec = npx * (int)(K * sizeof(T) / 16);
                for (int i = lane; i < nvec; i += 32)
                    reinterpret_cast<uint4 *>(grow)[i] = reinterpret_cast<const uint4 *>(srow)[i];
            } else {
                for (int i = lane; i < npx * K; i += 32) grow[i] = srow[i];
            }
        }
    }
}

// Tap records for the pass-1 organisations that do not write them themselves (everything except
// lay_tile_kernel): one thread per output pixel.
__global__ void __launch_bounds__(256) tap_records_kernel(CoordCfg cc, int64_t P, int64_t HW, int pitch,
                                                          const float2 *__restrict__ coords, uint32_t *rec_code, float2 *rec_frac) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const int64_t n = i / HW, rem = i - n * HW;
    const int y = (int)(rem / cc.W), x = (int)(rem - (int64_t)y * cc.W);
    const Taps t = make_taps(cc, __ldg(coords + i), y, x);
    const int64_t ro = (n * cc.H + y) * pitch + x;
    rec_code[ro] = tap_cell_code(cc, t, y, x);
    rec_frac[ro] = make_float2(t.ix - t.fx0, t.iy - t.fy0);
}

// ---- far path: zero the fixed-point accumulators of the flagged source tiles only ----
// One CTA per flagged tile (grid-stride), one warp per tile row: a row of a tile is npx * (3 + K) contiguous
// 8-byte words, zeroed with fully coalesced stores.
template <int K>
__global__ void __launch_bounds__(kThreads) far_zero_kernel(const Pass2Params p) {
    const int n_flagged = (int)p.hdr->n_flagged;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int i = blockIdx.x; i < n_flagged; i += gridDim.x) {
        const int t = p.flagged_list[i];
        const int n = t / (p.tiles_y * p.tiles_x), rem = t - n * (p.tiles_y * p.tiles_x);
        const int y = (rem / p.tiles_x) * kTH + wid, x0 = (rem % p.tiles_x) * kTW;
        if (y >= p.cc.H) continue;
        const int words = min(kTW, p.cc.W - x0) * (3 + K);
        long long *a = p.far_acc + ((int64_t)n * p.HW + (int64_t)y * p.cc.W + x0) * (3 + K);
        for (int q = lane; q < words; q += 32) a[q] = 0;
    }
}

// ---- far path: fixed-point scatter of the queued far output pixels ----
// One WARP per far pixel, lanes = the 3 + K channels of d_out (each lane loads its channel once), a warp-uniform
// loop over the four taps: consecutive lanes hit consecutive 8-byte accumulators of one source pixel, so every
// warp instruction is one coalesced run of integer atomics.  The integer atomics are associative, so neither
// the queue order nor the thread schedule can change the sums.  The queue entry carries the tap cell and the
// fractional weights pass 1 computed, so nothing is re-derived from the coordinates here.
// (First version: one THREAD per contribution, ~400 instructions of index arithmetic each: 11.6 ms for the 6 M
// far pixels of BASELINE config 5.  Second: one warp per pixel, lanes = (tap, channel) pairs, taps re-derived by
// make_taps, fp64 scaling: 273 instructions per pixel, 2.56 ms.)
template <int K>
__global__ void __launch_bounds__(kThreads) far_scatter_kernel(const Pass2Params p) {
    const unsigned n_far = p.hdr->far_count;
    if (n_far == 0) return;
    const int H = p.cc.H, W = p.cc.W;
    constexpr int CH = 3 + K;
    static_assert(CH <= 32 || CH <= 64, "channels are walked in at most two lane rounds");
    const float scale = ldexpf(1.0f, far_scale_exp(p.hdr, p.HW));    // a power of two: v * scale is exact in fp32
    const unsigned lane = threadIdx.x & 31;
    const unsigned gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    const unsigned HWu = (unsigned)p.HW;
    for (unsigned j = gw; j < n_far; j += nw) {
        const int4 e = __ldg(p.far_list + j);
        const unsigned i = (unsigned)e.x;                    // pixel index < 2^31 (check_problem)
        const unsigned n = i / HWu, rem = i - n * HWu;
        const int y = (int)(rem / (unsigned)W), x = (int)(rem - (unsigned)y * (unsigned)W);
        const int x0 = (int)((unsigned)e.y & 0xFFFFu) - 8, y0 = (int)((unsigned)e.y >> 16) - 8;
        const float fx = __int_as_float(e.z), fy = __int_as_float(e.w);
        // east/south = frac, west/north = 1 - frac: the weights of the near path (pass2_rec_kernel), bit-identical to the
        // forward's for every in-image tap
        const float wx[2] = {__fsub_rn(1.0f, fx), fx}, wy[2] = {__fsub_rn(1.0f, fy), fy};
        const float *drgb = p.d_out_rgb ? p.d_out_rgb + (((int64_t)n * H + y) * p.pitch + x) * 3 : nullptr;
        const float *dlay = p.d_out_lay ? p.d_out_lay + (int64_t)i * K : nullptr;
        long long *img_acc = p.far_acc + (int64_t)n * p.HW * CH;
#pragma unroll
        for (int c0 = 0; c0 < CH; c0 += 32) {
            const int c = c0 + (int)lane;
            const float *src = c < 3 ? drgb : dlay;
            float d = 0.f;
            const bool have = c < CH && src != nullptr;
            if (have) d = __ldg(src + (c < 3 ? c : c - 3));
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
                const int xs = x0 + (k4 & 1), ys = y0 + (k4 >> 1);
                if (xs < 0 || xs >= W || ys < 0 || ys >= H || !have) continue;      // warp-uniform but for `have`
                const float wt = __fmul_rn(wx[k4 & 1], wy[k4 >> 1]);
                unsigned long long *dst = reinterpret_cast<unsigned long long *>(img_acc + ((int64_t)ys * W + xs) * CH + c);
                atomicAdd(dst, (unsigned long long)__float2ll_rn(__fmul_rn(__fmul_rn(wt, d), scale)));
            }
        }
    }
}

}  // namespace vlg
