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
    const uint32_t *seg_cnt;     // [n_blocks][kTH] far output pixels with a tap in that row of that source tile (pass 1 counted them)
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
    const float g = __uint_as_float(max(hdr->maxgrad_rgb_bits, hdr->maxgrad_lay_bits));   // non-negative floats order like their bits
    int eg = 0, ehw = 0;
    frexpf(fmaxf(g, 1e-37f), &eg);
    frexp((double)HW, &ehw);
    return min(61 - eg - ehw, 96);   // capped so that 2^e is an fp32 number too (gradients below 2^-96 are noise anyway)
}
__device__ __forceinline__ double far_scale(const WsHeader *hdr, int64_t HW) { return ldexp(1.0, far_scale_exp(hdr, HW)); }

// Packed far accumulators.  The scatter is bound by the 32-byte sectors the L2 atomic units process (measured: 286 G
// lane-operations/s for runs of 64-bit REDs, whatever the SMs do), so the unit that halves its time is the sector:
// two channels share one 64-bit word, V = (a1 << 32) + a0 with a = round(v * 2^e) as SIGNED 32-bit integers, added as one
// signed 64-bit integer.  Sums of V are exact; the low lane is read back as a signed 32-bit integer and subtracted
// before the high lane is taken, which undoes every borrow.  What makes 32 bits enough is a bound on the NUMBER of
// contributions: pass 1 counts, per row of every source tile (a "segment": 32 source pixels), the far output pixels
// that have a tap in it (`seg_cnt`) -- an output pixel adds at most one tap to a given source pixel, so no source pixel
// of the segment receives more than cnt contributions.  With h = ceil(log2 cnt), eg = the exponent of the largest
// |d_out| of the channel's group (rgb and layout gradients differ by orders of magnitude: two maxima) and
// e = 30 - eg - h, no lane can leave (-2^31, 2^31), and every contribution is rounded at 2^-(30-h) of the group's
// largest gradient: h <= 6 for any flow that is not compressive, i.e. 24 bits -- fp32's own resolution.
// Segments with more than kFarPackedMaxCnt pixels (strongly compressive flow) keep one 64-bit accumulator per
// channel, as does any K whose channel pairs do not fit 32 lanes.  The mode is a function of the integer count only,
// and integer adds are associative: the result does not depend on the order of the atomics in either mode.
// Storage: a segment owns npx * (3 + K) words of `far_acc` either way; the packed mode uses the first npx * PW of them.
constexpr unsigned kFarPackedMaxCnt = 512;
template <int K> struct FarPack {
    static constexpr int CH = 3 + K, PW = (CH + 1) / 2;
    static constexpr bool can = 2 * PW <= 32;
};
__device__ __forceinline__ int float_exp(uint32_t bits) {
    int eg = 0;
    frexpf(fmaxf(__uint_as_float(bits), 1e-37f), &eg);
    return eg;
}
__device__ __forceinline__ int ceil_log2(unsigned cnt) { return cnt > 1u ? 32 - __clz((int)(cnt - 1u)) : 0; }
__device__ __forceinline__ float pow2f(int e) { return __int_as_float((min(max(e, -126), 127) + 127) << 23); }
// exponent of the packed fixed point of a group (eg) in a segment that receives cnt far pixels
__device__ __forceinline__ int far_packed_exp(int eg, unsigned cnt) { return min(30 - eg, 96) - ceil_log2(cnt); }
template <int K> __device__ __forceinline__ bool far_seg_packed(unsigned cnt) { return FarPack<K>::can && cnt <= kFarPackedMaxCnt; }
// far pixels counted in segment `seg` (no counts -- VLG_FLAG_FAR_WIDE -- reads as "too many": the wide mode)
__device__ __forceinline__ unsigned far_seg_count(const uint32_t *seg_cnt, int64_t seg) { return seg_cnt ? __ldg(seg_cnt + seg) : 0xFFFFFFFFu; }

// adds the far sums of source pixel (n, sy, sx) of tile bt (whose first column is tx0) to acc_r / acc_l
template <int K>
__device__ __forceinline__ void far_read(const Pass2Params &p, int n, int bt, int sy, int sx, int tx0, float (&acc_r)[3], float (&acc_l)[K]) {
    constexpr int CH = FarPack<K>::CH, PW = FarPack<K>::PW;
    const unsigned cnt = far_seg_count(p.seg_cnt, (int64_t)bt * kTH + (sy & (kTH - 1)));
    const long long *seg = p.far_acc + ((int64_t)n * p.HW + (int64_t)sy * p.cc.W + tx0) * CH;
    if (far_seg_packed<K>(cnt)) {
        const float inv_r = pow2f(-far_packed_exp(float_exp(p.hdr->maxgrad_rgb_bits), cnt));
        const float inv_l = pow2f(-far_packed_exp(float_exp(p.hdr->maxgrad_lay_bits), cnt));
        const long long *fa = seg + (sx - tx0) * PW;
#pragma unroll
        for (int w = 0; w < PW; ++w) {
            const long long S = fa[w];
            const int lo = (int)(unsigned)((unsigned long long)S & 0xFFFFFFFFull);
            const int hi = (int)((S - (long long)lo) >> 32);
            const int c0 = 2 * w, c1 = 2 * w + 1;
            if (c0 < 3) acc_r[c0] += (float)lo * inv_r; else acc_l[c0 - 3] += (float)lo * inv_l;
            if (c1 < CH) { if (c1 < 3) acc_r[c1] += (float)hi * inv_r; else acc_l[c1 - 3] += (float)hi * inv_l; }
        }
    } else {
        const double inv = 1.0 / far_scale(p.hdr, p.HW);
        const long long *fa = seg + (sx - tx0) * CH;
#pragma unroll
        for (int c = 0; c < 3; ++c) acc_r[c] += (float)((double)fa[c] * inv);
#pragma unroll
        for (int k = 0; k < K; ++k) acc_l[k] += (float)((double)fa[3 + k] * inv);
    }
}

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
    if (live && tile_far) far_read<K>(p, n, bt, sy, sx, tx0, acc_r, acc_l);
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
    if (live && tile_far) far_read<K>(p, n, bt, sy, sx, tx0, acc_r, acc_l);
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
    // N*H*W < 2^31 (check_problem): 32-bit divisions (a 64-bit division by a run-time value costs ~60 instructions)
    const unsigned nu = (unsigned)i / (unsigned)HW, rem = (unsigned)i - nu * (unsigned)HW;
    const int64_t n = nu;
    const int y = (int)(rem / (unsigned)cc.W), x = (int)(rem - (unsigned)y * (unsigned)cc.W);
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
        const int wpp = far_seg_packed<K>(far_seg_count(p.seg_cnt, (int64_t)t * kTH + wid)) ? FarPack<K>::PW : 3 + K;   // words per pixel
        const int words = min(kTW, p.cc.W - x0) * wpp;
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
// One far pixel's four taps, by the whole warp.  `modes`: 5 bits per (row, side) segment -- bit 0 = packed, bits 1-4 = h.
template <int K>
__device__ __forceinline__ void far_scatter_one(const Pass2Params &p, unsigned lane, int wd, int side_of_lane, int eg0, int eg1, float scale,
                                                unsigned n, int x0, int y0, float fx, float fy, unsigned modes,
                                                const float (&d)[(3 + K + 31) / 32], float d0, float d1) {
    constexpr int CH = FarPack<K>::CH, PW = FarPack<K>::PW;
    const int H = p.cc.H, W = p.cc.W;
    // east/south = frac, west/north = 1 - frac: the weights of the near path (pass2_rec_kernel), bit-identical to the
    // forward's for every in-image tap
    const float wx[2] = {__fsub_rn(1.0f, fx), fx}, wy[2] = {__fsub_rn(1.0f, fy), fy};
    long long *img_acc = p.far_acc + (int64_t)n * p.HW * CH;
    const bool okw = x0 >= 0 && x0 < W, oke = x0 + 1 >= 0 && x0 + 1 < W;
    const int txw = x0 >> 5, txe = (x0 + 1) >> 5;                           // kTW == 32
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int ys = y0 + r;
        if (ys < 0 || ys >= H) continue;                                    // warp-uniform
        const unsigned mw = (modes >> (10 * r)) & 31u, me = (modes >> (10 * r + 5)) & 31u;
        long long *row_acc = img_acc + (int64_t)ys * W * CH;
        if (okw && oke && txw == txe && (mw & 1u)) {
            // both taps of the row in one packed segment: one run of 2 PW consecutive words
            if ((int)lane < 2 * PW) {
                const int h = (int)(mw >> 1);
                const float wt = __fmul_rn(wx[side_of_lane], wy[r]);
                const long long a0 = (long long)__float2int_rn(__fmul_rn(__fmul_rn(wt, d0), pow2f(min(30 - eg0, 96) - h)));
                const long long a1 = (long long)__float2int_rn(__fmul_rn(__fmul_rn(wt, d1), pow2f(min(30 - eg1, 96) - h)));
                unsigned long long *dst = reinterpret_cast<unsigned long long *>(row_acc + (int64_t)(txw * kTW) * CH + (x0 - txw * kTW) * PW + lane);
                atomicAdd(dst, (unsigned long long)(a1 * 4294967296ll + a0));
            }
            continue;
        }
#pragma unroll
        for (int sd = 0; sd < 2; ++sd) {
            if (!(sd ? oke : okw)) continue;
            const unsigned m = sd ? me : mw;
            const int xs = x0 + sd, tx0s = (sd ? txe : txw) * kTW;
            const float wt = __fmul_rn(wx[sd], wy[r]);
            long long *seg = row_acc + (int64_t)tx0s * CH;
            if (m & 1u) {
                if ((int)lane < PW) {
                    const int h = (int)(m >> 1);
                    const long long a0 = (long long)__float2int_rn(__fmul_rn(__fmul_rn(wt, d0), pow2f(min(30 - eg0, 96) - h)));
                    const long long a1 = (long long)__float2int_rn(__fmul_rn(__fmul_rn(wt, d1), pow2f(min(30 - eg1, 96) - h)));
                    atomicAdd(reinterpret_cast<unsigned long long *>(seg + (xs - tx0s) * PW + lane), (unsigned long long)(a1 * 4294967296ll + a0));
                }
            } else {
#pragma unroll
                for (int c0 = 0; c0 < CH; c0 += 32) {
                    const int c = c0 + (int)lane;
                    if (c < CH)
                        atomicAdd(reinterpret_cast<unsigned long long *>(seg + (xs - tx0s) * CH + c),
                                  (unsigned long long)__float2ll_rn(__fmul_rn(__fmul_rn(wt, d[c0 / 32]), scale)));
                }
            }
        }
    }
}

// The kernel is a chain of dependent memory round trips per far pixel (queue entry -> segment counts -> d_out), not a
// stream: with one pixel per warp iteration it ran at the speed of that chain (2.4 ms for the 6 M far pixels of BASELINE
// config 5, whatever the atomics cost).  So a warp takes 32 queue entries at a time, ONE PER LANE (entry, index
// arithmetic and the four segment counts are loaded / computed lane-parallel: 32 chains in flight), and then walks them
// VLG_FAR_U at a time: their d_out loads are issued before the first of them is consumed.
#ifndef VLG_FAR_U
#define VLG_FAR_U 2         // pixels whose d_out loads are in flight together.  Measured on BASELINE config 5 (U / CTAs per SM):
#endif                      // 2/4 1.33 ms, 2/5 1.42 (spills), 4/3 1.52, 4/4 1.56 (spills), 8/3 1.55, 8/2 2.11 -- warps beat depth
#ifndef VLG_FAR_CTAS
#define VLG_FAR_CTAS 4      // 62 registers
#endif
template <int K>
__global__ void __launch_bounds__(kThreads, VLG_FAR_CTAS) far_scatter_kernel(const Pass2Params p) {
    const unsigned n_far = p.hdr->far_count;
    if (n_far == 0) return;
    const int H = p.cc.H, W = p.cc.W;
    constexpr int CH = FarPack<K>::CH, PW = FarPack<K>::PW, U = VLG_FAR_U, ND = (CH + 31) / 32;
    static_assert(CH <= 64, "channels are walked in at most two lane rounds");
    const float scale = ldexpf(1.0f, far_scale_exp(p.hdr, p.HW));    // a power of two: v * scale is exact in fp32
    const unsigned lane = threadIdx.x & 31;
    // packed segments: lanes 0 .. 2 PW - 1 = (west | east tap, word); a lane adds channels 2 wd and 2 wd + 1 with one atomic
    const int wd = (int)lane % PW, side_of_lane = (int)lane / PW;
    const int eg0 = float_exp(2 * wd < 3 ? p.hdr->maxgrad_rgb_bits : p.hdr->maxgrad_lay_bits);
    const int eg1 = float_exp(2 * wd + 1 < 3 ? p.hdr->maxgrad_rgb_bits : p.hdr->maxgrad_lay_bits);
    const unsigned gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    const unsigned HWu = (unsigned)p.HW;
    const float S0 = pow2f(min(30 - eg0, 96)), S1 = pow2f(min(30 - eg1, 96));     // this lane's two channels: 2^(30 - eg) ...
    // where this lane finds its two channels of a pixel's d_out: base + pixel index * stride (rgb staging: pitched index q,
    // 12-byte pixels; layout staging: index i, K floats)
    const int c0 = 2 * wd, c1 = 2 * wd + 1;
    const bool has0 = (int)lane < 2 * PW && (c0 >= 3 || p.d_out_rgb != nullptr), has1 = (int)lane < 2 * PW && c1 < CH && (c1 >= 3 || p.d_out_rgb != nullptr);
    const float *base0 = c0 < 3 ? p.d_out_rgb + c0 : p.d_out_lay + (c0 - 3), *base1 = c1 < 3 ? p.d_out_rgb + c1 : p.d_out_lay + (c1 - 3);
    const unsigned stride0 = c0 < 3 ? 3u : (unsigned)K, stride1 = c1 < 3 ? 3u : (unsigned)K;
    const long long row_words = (long long)W * CH;
    // a short queue (a handful of far pixels in an otherwise smooth flow) is spread over all warps instead of being
    // walked by a few of them: `per` entries per warp and round
    const unsigned per = min(32u, max(1u, (n_far + nw - 1u) / nw));
    for (unsigned base = gw * per; base < n_far; base += nw * per) {
        // ---- lane-parallel: one queue entry per lane ----
        const unsigned j = lane < per ? base + lane : n_far;
        const int4 e = j < n_far ? __ldg(p.far_list + j) : make_int4(0, 8 | (8 << 16), 0, 0);
        const unsigned my_i = (unsigned)e.x;                 // pixel index < 2^31 (check_problem)
        const unsigned my_n = my_i / HWu, rem = my_i - my_n * HWu;
        const unsigned my_y = rem / (unsigned)W, my_x = rem - my_y * (unsigned)W;
        const unsigned my_q = (my_n * (unsigned)H + my_y) * (unsigned)p.pitch + my_x;     // pitched pixel index (< 2^32)
        const int my_x0 = (int)((unsigned)e.y & 0xFFFFu) - 8, my_y0 = (int)((unsigned)e.y >> 16) - 8;
        const float my_fx = __int_as_float(e.z), my_fy = __int_as_float(e.w);
        unsigned my_modes = 0u;
        const bool okw = my_x0 >= 0 && my_x0 < W, oke = my_x0 + 1 >= 0 && my_x0 + 1 < W;
        const int txw = my_x0 >> 5, txe = (my_x0 + 1) >> 5;                           // kTW == 32
        {
            unsigned cnt[4] = {0u, 0u, 0u, 0u};
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int ys = my_y0 + r;
                if (ys < 0 || ys >= H || j >= n_far) continue;
                const int64_t seg_row = ((int64_t)my_n * p.tiles_y + ys / kTH) * p.tiles_x * kTH + (ys & (kTH - 1));
                if (okw) cnt[2 * r] = far_seg_count(p.seg_cnt, seg_row + (int64_t)txw * kTH);
                if (oke) cnt[2 * r + 1] = (okw && txe == txw) ? cnt[2 * r] : far_seg_count(p.seg_cnt, seg_row + (int64_t)txe * kTH);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (far_seg_packed<K>(cnt[q])) my_modes |= (1u | ((unsigned)ceil_log2(cnt[q]) << 1)) << (5 * q);
        }
        // The common case, prepared here once per entry instead of once per entry AND lane: both taps of a row lie in one
        // packed segment (so a row is one run of 2 PW consecutive words), for every row inside the image.  Then the
        // headroom 2^-h goes into the row weight (a power of two: the products round exactly as without it), and the run
        // of row 0 starts at word  west * CH - xin * (CH - PW)  (west = pixel index of the west tap, xin = its column in the
        // tile); row 1 lies W * CH words further.
        const bool v0 = my_y0 >= 0 && my_y0 < H, v1 = my_y0 + 1 >= 0 && my_y0 + 1 < H;
        const bool pairable = okw && oke && txw == txe;
        const unsigned mr0 = my_modes & 31u, mr1 = (my_modes >> 10) & 31u;
        const bool fast = (!v0 || (pairable && (mr0 & 1u))) && (!v1 || (pairable && (mr1 & 1u)));
        const float my_wyh0 = __fmul_rn(__fsub_rn(1.0f, my_fy), pow2f(-(int)(mr0 >> 1))), my_wyh1 = __fmul_rn(my_fy, pow2f(-(int)(mr1 >> 1)));
        const long long my_run = ((long long)(my_n * HWu) + (long long)my_y0 * W + my_x0) * CH - (long long)(my_x0 & 31) * (CH - PW);
        // bits 0 / 1: row 0 / 1 receives something, bit 2: the fast path applies.  0 for the lanes past the end of the queue.
        const unsigned my_flags = j < n_far ? ((v0 ? 1u : 0u) | (v1 ? 2u : 0u) | (fast ? 4u : 0u)) : 0u;
        const int nb = (int)min(per, n_far - base);
        for (int k0 = 0; k0 < nb; k0 += U) {                  // k0 + u <= 31: U divides 32
            unsigned fl[U]; long long run[U]; float ffx[U], wyh0[U], wyh1[U], d0[U], d1[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int sl = k0 + u;
                const unsigned i = __shfl_sync(0xffffffffu, my_i, sl), q = __shfl_sync(0xffffffffu, my_q, sl);
                fl[u] = __shfl_sync(0xffffffffu, my_flags, sl);
                run[u] = __shfl_sync(0xffffffffu, my_run, sl);
                ffx[u] = __shfl_sync(0xffffffffu, my_fx, sl);
                wyh0[u] = __shfl_sync(0xffffffffu, my_wyh0, sl); wyh1[u] = __shfl_sync(0xffffffffu, my_wyh1, sl);
                // (lanes past the end of the queue point at pixel 0: a harmless load)
                d0[u] = has0 ? __ldg(base0 + (size_t)(c0 < 3 ? q : i) * stride0) : 0.f;
                d1[u] = has1 ? __ldg(base1 + (size_t)(c1 < 3 ? q : i) * stride1) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (fl[u] & 4u) {                                                     // warp-uniform
                    if ((int)lane < 2 * PW) {
                        const float wxs = side_of_lane ? ffx[u] : __fsub_rn(1.0f, ffx[u]);
                        const float dd0 = __fmul_rn(d0[u], S0), dd1 = __fmul_rn(d1[u], S1);     // ... exact, |.| < 2^30
                        long long *dst = p.far_acc + (run[u] + lane);
                        if (fl[u] & 1u) {
                            const float wt = __fmul_rn(wxs, wyh0[u]);
                            const long long a0 = (long long)__float2int_rn(__fmul_rn(wt, dd0)), a1 = (long long)__float2int_rn(__fmul_rn(wt, dd1));
                            atomicAdd(reinterpret_cast<unsigned long long *>(dst), (unsigned long long)((a1 << 32) + a0));
                        }
                        if (fl[u] & 2u) {
                            const float wt = __fmul_rn(wxs, wyh1[u]);
                            const long long a0 = (long long)__float2int_rn(__fmul_rn(wt, dd0)), a1 = (long long)__float2int_rn(__fmul_rn(wt, dd1));
                            atomicAdd(reinterpret_cast<unsigned long long *>(dst + row_words), (unsigned long long)((a1 << 32) + a0));
                        }
                    }
                } else if (fl[u] & 3u) {                      // warp-uniform, rare: straddled tile borders, wide segments, other K
                    const int sl = k0 + u;
                    const unsigned i = __shfl_sync(0xffffffffu, my_i, sl), q = __shfl_sync(0xffffffffu, my_q, sl);
                    float d[ND];
#pragma unroll
                    for (int cc0 = 0; cc0 < CH; cc0 += 32) {
                        const int c = cc0 + (int)lane;
                        const float *src = c < 3 ? (p.d_out_rgb ? p.d_out_rgb + (size_t)q * 3 + c : nullptr) : p.d_out_lay + (size_t)i * K + (c - 3);
                        d[cc0 / 32] = (c < CH && src != nullptr) ? __ldg(src) : 0.f;
                    }
                    far_scatter_one<K>(p, lane, wd, side_of_lane, eg0, eg1, scale, __shfl_sync(0xffffffffu, my_n, sl),
                                       __shfl_sync(0xffffffffu, my_x0, sl), __shfl_sync(0xffffffffu, my_y0, sl), ffx[u],
                                       __shfl_sync(0xffffffffu, my_fy, sl), __shfl_sync(0xffffffffu, my_modes, sl), d, d0[u], d1[u]);
                }
            }
        }
    }
}

}  // namespace vlg
