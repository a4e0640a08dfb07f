// vlg_lay.cuh -- the layout half of pass 1 as a per-warp strip pipeline fed by a TMA row ring.
//
// What it computes (reference gongaa/video-layout-generation):
//   warp of the K-channel layout  (absent upstream) F.grid_sample(bilinear, align_corners=True) on
//                                 the src/models/modules.py:69 grid, bit-exact FMA chain (App. A.6)
//   argmax layouts                src/trainer.py:342,423,467 (first maximal index)
//   CE                            src/trainer.py:124,250: mean over labels != ignore_index of -log_softmax[label]
//   TV on the flow                (absent upstream) stencils of src/loss.py:22,24
// plus d(loss)/d(warped layout) (fp32 staging for pass 2), the layout + TV part of d(loss)/d(coords)
// (added to the rgb part the strip kernel of vlg_rgb.cuh wrote), the far-pixel bookkeeping pass 2
// needs, and -- in the last CTA to finish -- the final fixed-order reduction of every partial sum.
//
// Organisation.  The tile kernel exposed two dependent global round trips per CTA (flow -> bounding
// box -> window) and three barriers; with 24 warps per SM it issued one instruction every third
// cycle.  Here ONE WARP owns a strip of 32 output columns and walks down the rows:
//   * the source layout rows it samples live in a per-warp RING of shared-memory row buffers; each
//     new source row is ONE cp.async.bulk.tensor (TMA) issued by lane 0 two output rows ahead of its
//     first use, completing on a per-slot mbarrier; the hardware zero-fills out-of-image pixels;
//   * a row's position in the ring and its column origin follow the flow (bounding box of the
//     row's taps, outliers excluded), so coherent large motion still samples from shared memory;
//     lanes whose taps are not resident fall back to the global 4-tap gather (same FMA chain);
//   * flow, labels and the rgb part of d_coords are loaded two to three rows ahead into registers;
//   * d(loss)/d(warped layout) leaves through shared memory as one bulk store per row (the direct
//     80-byte-stride stores cost 20 L1 wavefronts per instruction);
//   * no __syncthreads in the row loop: warps are independent, work is an equal contiguous run of
//     (image, strip, row) per resident warp.
#pragma once
#include <cuda.h>

#include "vlg_device.cuh"
#include "vlg_pass1.cuh"   // mbarrier / TMA helpers, source_xy, taps_from_xy

namespace vlg {

#ifdef VLG_LAY_PROFILE   // tuning builds only: per-section cycle counters (serialise the sections)
#define VLG_PROF_DECL long long prof_acc[16] = {0}; long long prof_t = clock64();
#define VLG_PROF_START prof_t = clock64();
#define VLG_PROF(i) { const long long now_ = clock64(); prof_acc[i] += now_ - prof_t; prof_t = now_; }
#define VLG_PROF_FLUSH if (lane == 0) { for (int i_ = 0; i_ < 16; ++i_) atomicAdd(&p.hdr->hist[16 + i_], (unsigned long long)prof_acc[i_]); }
#else
#define VLG_PROF_DECL
#define VLG_PROF_START
#define VLG_PROF(i)
#define VLG_PROF_FLUSH
#endif

#ifndef VLG_ABL
#define VLG_ABL 0   // tuning builds: ablation bits (results are wrong when non-zero)
#endif

constexpr int kLW = 32;            // output columns per strip (== kTW: pass 2's tile width)
constexpr int kLayThreads = 128;   // 4 independent warps per CTA
constexpr int kLayWarps = kLayThreads / 32;
constexpr int kLR = 5;             // ring slots (source rows resident per warp)
constexpr int kLD = 2;             // output rows between a row's TMA issue and its use
constexpr int kLBW = 40;           // staged pixels per source row
constexpr int kLayMaxWarps = 8192; // partial rows reserved in the workspace
#ifndef VLG_LAY_MIN_BLOCKS
#define VLG_LAY_MIN_BLOCKS 3
#endif

struct LayParams {
    CoordCfg cc;
    int N, strips;                 // strips = ceil(W / kLW)
    int tiles_y;                   // ceil(H / kTH) (tile geometry of pass 2)
    int64_t total_rows, chunk;
    const void *src_layout;
    const float *coords;
    const int64_t *label;
    int64_t ignore_index;
    const float *class_weight;
    int weighted_denom;
    float w_ce_over_scale;
    float c_tvh, c_tvw;
    int do_tv;
    int accum_dcoords;             // d_coords already holds the rgb part (the rgb strip kernel ran before): add to it
    float *d_coords;               // nullable (validation)
    float *d_out_lay;              // [P][K] fp32 staging, nullable
    uint32_t *rec_code;            // [N][H][pitch] tap records for pass 2 (tap_cell_code), nullable: lay_tile_kernel only
    float2 *rec_frac;              // [N][H][pitch] fractional tap weights (ix - x0, iy - y0)
    int pitch;
    int64_t *out_argmax;           // nullable
    float *partials;               // [n_warps][4]: ce, tv_h, tv_w, -
    float *tile_disp;              // [n_tiles] max NEAR displacement per 32x8 tile (zero-initialised, atomicMax)
    int *far_list;
    uint32_t *tile_flags;
    int *flagged_list;
    ReduceParams red;
    WsHeader *hdr;
};

template <typename T, int K>
struct LayWarpSmem {
    static constexpr int kSlotBytes = (kLBW * K * (int)sizeof(T) + 127) / 128 * 128;
    alignas(128) unsigned char ring[kLR][kSlotBytes];
    alignas(128) float obuf[kLW * K];      // one output row of d(loss)/d(warped layout)
    alignas(8) uint64_t bar[kLR];
    int ox[kLR];                           // column origin of the row held by each slot
};

template <typename T, int K, bool GRAD>
__global__ void __launch_bounds__(kLayThreads, VLG_LAY_MIN_BLOCKS) lay_strip_kernel(const LayParams p,
                                                                                   const __grid_constant__ CUtensorMap lay_map) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    using WS = LayWarpSmem<T, K>;
    const CoordCfg &cc = p.cc;
    const int H = cc.H, W = cc.W;
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int wib = __shfl_sync(FULL, (int)(threadIdx.x >> 5), 0);   // warp-uniform for the compiler
    const int gw = blockIdx.x * kLayWarps + wib;
    WS &sm = *reinterpret_cast<WS *>(smem_raw + (size_t)wib * sizeof(WS));
    constexpr int PXB = K * (int)sizeof(T);
    constexpr unsigned kRowBytes = (unsigned)(kLBW * PXB);

    if (lane < kLR) mbar_init(&sm.bar[lane], 1);
    __syncwarp();
    VLG_PROF_DECL
    // Ring protocol (all warp-uniform): source rows are issued in increasing order, `top` is the next row
    // to issue and lives in slot `slot_top`; rows below `ready` have been waited for.  Every load is
    // waited for exactly once, in issue order (when a row needs it, before its slot is re-armed, or at
    // the end of the segment), so an mbarrier never has two loads outstanding in one phase and no TMA
    // write is in flight when shared memory is reused or released.
    unsigned issue_par = 0u;  // bit s: parity the NEXT load issued on slot s will complete

    int64_t rho = (int64_t)gw * p.chunk;
    const int64_t rho_end = min(rho + p.chunk, p.total_rows);
    float s_ce = 0.f, s_tvh = 0.f, s_tvw = 0.f;
    float m_disp = 0.f, m_grad = 0.f;
    const float wreg = (p.class_weight && lane < K) ? __ldg(p.class_weight + lane) : 1.0f;   // class weights need K <= 32
    const float denom = GRAD ? (p.weighted_denom ? (float)__ldcg(&p.hdr->ce_denom) : (float)__ldcg(&p.hdr->n_valid)) : 1.0f;
    const float ce_unit = GRAD ? p.w_ce_over_scale / denom : 0.f;   // one division per thread, not one per pixel
    const T *src_all = reinterpret_cast<const T *>(p.src_layout);
    const float mx_c = (cc.coord_mode == VLG_COORD_FLOW) ? __fmul_rn(cc.Wm1, 0.5f) * cc.sx : __fmul_rn(cc.Wm1, 0.5f);
    const float my_c = (cc.coord_mode == VLG_COORD_FLOW) ? __fmul_rn(cc.Hm1, 0.5f) * cc.sy : __fmul_rn(cc.Hm1, 0.5f);
    const bool border = cc.padding == VLG_PAD_BORDER;

    while (rho < rho_end) {
        // ---- one segment: rows [ya, yb) of strip s of image n ----
        const int colid = (int)(rho / H);
        const int ya = (int)(rho - (int64_t)colid * H);
        const int yb = (int)min((int64_t)H, (int64_t)ya + (rho_end - rho));
        rho += yb - ya;
        const int n = colid / p.strips, s = colid - n * p.strips;
        const int x = s * kLW + lane;
        const bool col_ok = x < W;
        const int xc = min(x, W - 1);
        const float bxv = base_coord(xc, cc.Wm1);
        const int64_t img = (int64_t)n * H * W;
        const T *src_lay = src_all + img * K;
        const float2 *coords = reinterpret_cast<const float2 *>(p.coords) + img;
        const int64_t *labels = p.label + img;
        const int npx = min(kLW, W - s * kLW);
        const int ref_a = min(10, npx - 1), ref_b = min(21, npx - 1);   // the two lanes the ring follows

        // lane-constant pieces of the TV stencil: coefficient = 0 where the neighbour does not exist
        const float cR = (col_ok && x + 1 < W) ? p.c_tvw : 0.f, cL = (col_ok && x >= 1) ? p.c_tvw : 0.f;
        const float mRt = (col_ok && x + 1 < W) ? 1.f : 0.f, mCol = col_ok ? 1.f : 0.f;
        const float cV = col_ok ? p.c_tvh : 0.f;
        const int xe = lane == 0 ? x - 1 : x + 1;                       // column of the strip's outer neighbour
        const bool edge = p.do_tv && ((lane == 0 && x >= 1) || (lane == 31 && x + 1 < W));

        // ring state
        int top = 0, slot_top = 0, ready = 0, lo = 0;   // set by the first plan
        bool ring_started = false;
        int ref_y = 0, ref_dx = 0;                      // reference-lane tap row / column displacement of the last planned row

        // base-grid y coordinates of 32 consecutive rows, one IEEE division per lane per 32 rows
        int by_row0 = ya;
        float by_tab = base_coord(by_row0 + lane, cc.Hm1);

        auto load_flow = [&](int r) -> float2 {
            return (r >= 0 && r < H) ? __ldg(coords + (r * W + xc)) : make_float2(0.f, 0.f);
        };
        auto load_edge = [&](int r) -> float2 {   // flow of the column just outside the strip (lanes 0 and 31)
            return (edge && r < H) ? __ldg(coords + (r * W + xe)) : make_float2(0.f, 0.f);
        };
        auto load_label = [&](int r) -> int64_t { return (r < yb) ? __ldg(labels + (r * W + xc)) : (int64_t)0; };
        auto wait_next = [&]() {       // wait for the oldest issued row that nobody has waited for yet
            int sl = slot_top - (top - ready);
            sl = sl < 0 ? sl + kLR : sl;
            mbar_wait(&sm.bar[sl], ((issue_par >> sl) & 1u) ^ 1u);
            ++ready;
        };
        // sampling position of output row u + the ring loads it needs (stage B); returns the exclusive upper
        // bound of the source rows the row samples, as far as the reference lanes tell
        auto plan_row = [&](int u, float2 fl, float2 &xy_o, int &r_lo_o, int low_pending) -> int {
            r_lo_o = INT_MAX;
            if (u >= yb) return INT_MIN;
            if (u - by_row0 >= 32) { by_row0 += 32; by_tab = base_coord(by_row0 + lane, cc.Hm1); }
            float mx, my;
            const float2 xy = source_xy(cc, fl, bxv, __shfl_sync(FULL, by_tab, u - by_row0), mx, my);
            xy_o = xy;
            const int x0 = (int)fminf(fmaxf(floorf(xy.x), -4.0f), (float)W + 4.0f);
            const int y0 = (int)fminf(fmaxf(floorf(xy.y), -4.0f), (float)H + 4.0f);
            // Follow two reference lanes instead of a bounding box (no warp reductions): of the two, the
            // one closer to where the previous row was wins, so a single outlier does not derail the ring.
            const int ya_ = __shfl_sync(FULL, y0, ref_a), yb_ = __shfl_sync(FULL, y0, ref_b);
            const int xa_ = __shfl_sync(FULL, x0 - x, ref_a), xb_ = __shfl_sync(FULL, x0 - x, ref_b);
            const int cont = ref_y + 1;
            const bool pick_a = !ring_started || abs(ya_ - cont) <= abs(yb_ - cont);
            const int ry = pick_a ? ya_ : yb_, rdx = pick_a ? xa_ : xb_;
            const int r_lo = min(ry, ring_started ? min(ya_, yb_) : ry) - 1;   // one row of slack above
            const int want_top = max(ya_, yb_) + 2;                            // rows < want_top hold the reference lanes' taps
            r_lo_o = r_lo;
            const int low_needed = min(r_lo, low_pending);                     // rows the not yet finished output rows sample
            const int ox_new = s * kLW + rdx - (kLBW - kLW - 2) / 2;
            ref_y = ry; ref_dx = rdx;
            if (!ring_started || r_lo > top || r_lo < lo - 1) {
                // first row of the segment, or the flow jumped: drain what is in flight and restart the ring
                while (ready < top) wait_next();
                top = ready = lo = r_lo;
                slot_top = 0;
                ring_started = true;
            }
            if (top < want_top && top - kLR < low_needed) {
                __syncwarp();
                do {
                    if (ready <= top - kLR) wait_next();          // the slot's previous load must have been waited for
                    if (lane == 0) {
                        sm.ox[slot_top] = ox_new;
                        mbar_expect_tx(&sm.bar[slot_top], kRowBytes);
                        tma_load_4d(sm.ring[slot_top], &lay_map, &sm.bar[slot_top], 0, ox_new, top, n);
                    }
                    issue_par ^= 1u << slot_top;
                    ++top;
                    slot_top = slot_top + 1 == kLR ? 0 : slot_top + 1;
                } while (top < want_top && top - kLR < low_needed);   // never overwrite rows a pending output row samples
                __syncwarp();
            }
            return want_top;
        };

        // ---- prologue ----
        // Loaded values are consumed at the TOP of the next iteration and the register is re-loaded right
        // after: a value is never moved in the iteration that loads it (a move placed right behind its load
        // exposes the whole memory latency -- measured: 18 % of all stall samples sat on one such MOV).
        int t = ya;
        float2 fq[kLD + 1], xyq[kLD + 1];          // flows / sampling positions of rows t .. t+kLD
        int wtq[kLD + 1], rlq[kLD + 1];            // their want_top / lowest sampled row (as far as the reference lanes tell)
#pragma unroll
        for (int i = 0; i <= kLD; ++i) { fq[i] = make_float2(0.f, 0.f); xyq[i] = make_float2(0.f, 0.f); wtq[i] = INT_MIN; rlq[i] = INT_MAX; }
#pragma unroll
        for (int i = 0; i < kLD; ++i) {
            fq[i] = load_flow(t + i);
            wtq[i] = plan_row(t + i, fq[i], xyq[i], rlq[i], i ? rlq[0] : INT_MAX);
        }
        float2 pend_fl = load_flow(t + kLD);       // row t+kLD: planned at the top of iteration t
        int64_t pend_lab = load_label(t);          // row t
        float2 pend_edge = load_edge(t);           // row t
        auto load_dc = [&](int r) -> float2 {      // rgb part of d(loss)/d(coords) of row r
            return (GRAD && p.accum_dcoords && col_ok && r < yb)
                       ? __ldcg(reinterpret_cast<const float2 *>(p.d_coords) + img + (r * W + x)) : make_float2(0.f, 0.f);
        };
        float2 pend_dc = load_dc(t);               // row t
        float2 fl_prev = load_flow(t - 1);         // TV stencil
        float m_near_lane = 0.f, m_disp_lane = 0.f;
        auto flush_tile = [&](int tile_row) {
            const unsigned nmax = __reduce_max_sync(FULL, __float_as_uint(m_near_lane));
            if (lane == 0 && nmax != 0u)
                atomicMax(reinterpret_cast<unsigned *>(p.tile_disp) + ((int64_t)n * p.tiles_y + tile_row) * p.strips + s, nmax);
            m_near_lane = 0.f;
        };

#pragma unroll 1
        for (; t < yb; ++t) {
            VLG_PROF_START
            // ---- stage A: take over last iteration's loads, issue this iteration's ----
            fq[kLD] = pend_fl;
            const int64_t lb = pend_lab;
            const float2 fedge = pend_edge;
            const float2 dc = pend_dc;
            pend_dc = load_dc(t + 1);
            pend_fl = load_flow(t + kLD + 1);
            pend_lab = load_label(t + 1);
            pend_edge = load_edge(t + 1);
            VLG_PROF(7)
            // ---- stage B: plan row t+kLD ----
            wtq[kLD] = plan_row(t + kLD, fq[kLD], xyq[kLD], rlq[kLD], min(rlq[0], rlq[1]));
            VLG_PROF(0)

            // ---- stage C: output row t ----
            const float2 f = fq[0];
            const int y = t;
            const float2 xy = xyq[0];
            float mx = mx_c, my = my_c;
            if (border) {   // the clipped position tells whether the border clip was active
                mx = (xy.x <= 0.0f || xy.x >= cc.Wm1) ? 0.0f : mx;
                my = (xy.y <= 0.0f || xy.y >= cc.Hm1) ? 0.0f : my;
            }
            const Taps tp = taps_from_xy(cc, xy, mx, my);

            // displacement bookkeeping for pass 2 (near radius per tile, far queue)
            if ((y & (kTH - 1)) == 0 && y != ya) flush_tile(y / kTH - 1);
            const float disp = col_ok ? tap_displacement(cc, tp, y, x) : 0.f;
            const bool is_far = disp >= (float)VLG_NEAR_RADIUS;
            m_disp_lane = fmaxf(m_disp_lane, disp);
            m_near_lane = fmaxf(m_near_lane, is_far ? 0.f : disp);
            if (is_far && p.d_out_lay != nullptr && !(VLG_ABL & 8)) {
                if (p.far_list) {
                    p.far_list[atomicAdd(&p.hdr->far_count, 1u)] = (int)(img + (int64_t)y * W + x);
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4) {
                        const int xx = tp.x0 + (k4 & 1), yy = tp.y0 + (k4 >> 1);
                        if (xx >= 0 && xx < W && yy >= 0 && yy < H) {
                            const int tl = (n * p.tiles_y + yy / kTH) * p.strips + xx / kTW;
                            if (atomicOr(&p.tile_flags[tl], 1u) == 0u) p.flagged_list[atomicAdd(&p.hdr->n_flagged, 1u)] = tl;
                        }
                    }
                } else {
                    atomicOr(&p.hdr->status, VLG_STATUS_FAR_TAPS);
                }
            }
            VLG_PROF(1)

            // wait for the ring rows this output row samples (normally exactly one: the row issued kLD iterations ago)
            {
                // rows below want_top were issued >= kLD iterations ago; the next one (issued one iteration ago) is
                // waited for only when some lane samples a row below the reference lanes
                const bool deeper = __any_sync(FULL, col_ok && tp.y0 + 2 > wtq[0]);
                const int need = min(wtq[0] + (deeper ? 1 : 0), top);
                while (ready < need) wait_next();
            }
            VLG_PROF(2)

            const bool lab_ok = lb >= 0 && lb < K;
            if (col_ok && !lab_ok && lb != p.ignore_index) atomicOr(&p.hdr->status, VLG_STATUS_BAD_LABEL);
            const int il = lab_ok ? (int)lb : 0;
            // class weight of this pixel's label: lane k keeps weight k in a register (loaded once), so the loop has no
            // dependent global load -- a predicated-off LDG here still cost a long-scoreboard round trip (14 % of stalls)
            const float wl = __shfl_sync(FULL, wreg, il);

            float z[K], v[K];
            float vl[4];
            float zl;
            const T *st0 = nullptr, *st1 = nullptr;   // smem addresses of the nw / sw taps when resident
            {
                // rows [max(lo, top - kLR), ready) are resident and complete
                const int d_top = top - tp.y0;        // slot(y0) = slot_top - d_top (mod kLR)
                if (col_ok && tp.y0 >= lo && d_top <= kLR && tp.y0 + 1 < ready) {
                    int sl0 = slot_top - d_top;
                    sl0 = sl0 < 0 ? sl0 + kLR : sl0;
                    const int sl1 = sl0 + 1 == kLR ? 0 : sl0 + 1;
                    const unsigned rx0 = (unsigned)(tp.x0 - sm.ox[sl0]), rx1 = (unsigned)(tp.x0 - sm.ox[sl1]);
                    if (rx0 < (unsigned)(kLBW - 1) && rx1 < (unsigned)(kLBW - 1)) {
                        st0 = reinterpret_cast<const T *>(sm.ring[sl0]) + rx0 * K;
                        st1 = reinterpret_cast<const T *>(sm.ring[sl1]) + rx1 * K;
                    }
                }
            }
            VLG_PROF(9)
#ifdef VLG_LAY_PROFILE
            prof_acc[13] += __popc(__ballot_sync(FULL, col_ok && st0 == nullptr)); prof_acc[14] += __any_sync(FULL, col_ok && st0 == nullptr); prof_acc[15] += 1;
#endif
            if (st0) {
                load_px_smem<T, K>(st0, v);       mul2_bcast<K>(z, v, tp.nw);
                load_px_smem<T, K>(st0 + K, v);   fma2_bcast<K>(z, v, tp.ne);
                load_px_smem<T, K>(st1, v);       fma2_bcast<K>(z, v, tp.sw);
                load_px_smem<T, K>(st1 + K, v);   fma2_bcast<K>(z, v, tp.se);
                vl[0] = to_f<T>(st0[il]); vl[1] = to_f<T>(st0[K + il]);
                vl[2] = to_f<T>(st1[il]); vl[3] = to_f<T>(st1[K + il]);
            } else {
                gather_px<T, K>(src_lay, cc, tp, z);
                vl[0] = tap_global<T>(src_lay, K, il, tp.y0, tp.x0, H, W);
                vl[1] = tap_global<T>(src_lay, K, il, tp.y0, tp.x0 + 1, H, W);
                vl[2] = tap_global<T>(src_lay, K, il, tp.y0 + 1, tp.x0, H, W);
                vl[3] = tap_global<T>(src_lay, K, il, tp.y0 + 1, tp.x0 + 1, H, W);
            }
            // same FMA chain as z[il], so zl == z[il] bit for bit without indexing registers
            zl = __fmaf_rn(vl[3], tp.se, __fmaf_rn(vl[2], tp.sw, __fmaf_rn(vl[1], tp.ne, __fmul_rn(vl[0], tp.nw))));
            VLG_PROF(10)

            float m = z[0];
#pragma unroll
            for (int k = 1; k < K; ++k) m = fmaxf(m, z[k]);
            if (p.out_argmax && col_ok) {
                int best = K - 1;
#pragma unroll
                for (int k = K - 2; k >= 0; --k) best = (z[k] == m) ? k : best;   // first maximal index (src/trainer.py:342)
                p.out_argmax[img + (int64_t)y * W + x] = best;
            }
            const float L2E = 1.4426950408889634f;
            const float ml2 = m * L2E;
            float se4[4] = {0.f, 0.f, 0.f, 0.f};   // four partial sums: a 20-long dependent FADD chain is a stall
#pragma unroll
            for (int k = 0; k < K; ++k) {
                z[k] = ex2_approx(fmaf(z[k], L2E, -ml2));
                se4[k & 3] += z[k];
            }
            const float se = (se4[0] + se4[1]) + (se4[2] + se4[3]);
            if (lab_ok && col_ok) s_ce += wl * (fmaf(lg2_approx(se), 0.6931471805599453f, m) - zl);
            VLG_PROF(3)

            float gix = 0.f, giy = 0.f;
            if (GRAD) {
                const float cce = (lab_ok && col_ok) ? ce_unit * wl : 0.0f;
                const float inv = cce * rcp_approx(se);
                mul2_bcast<K>(z, z, inv);
                const float gl = fmaf(ex2_approx(fmaf(zl, L2E, -ml2)), inv, -cce);
                m_grad = fmaxf(m_grad, cce);
                float dnw, dne, dsw, dse;
                if (VLG_ABL & 4) { dnw = dne = dsw = dse = 0.f; gix = z[3]; giy = z[7]; } else
                if (st0) {
                    load_px_smem<T, K>(st0, v);       dnw = dot2<K>(z, v);
                    load_px_smem<T, K>(st0 + K, v);   dne = dot2<K>(z, v);
                    load_px_smem<T, K>(st1, v);       dsw = dot2<K>(z, v);
                    load_px_smem<T, K>(st1 + K, v);   dse = dot2<K>(z, v);
                    dnw = fmaf(-cce, vl[0], dnw); dne = fmaf(-cce, vl[1], dne);
                    dsw = fmaf(-cce, vl[2], dsw); dse = fmaf(-cce, vl[3], dse);
                    const float wx1 = tp.ix - tp.fx0, wx0 = (tp.fx0 + 1.0f) - tp.ix;
                    const float wy1 = tp.iy - tp.fy0, wy0 = (tp.fy0 + 1.0f) - tp.iy;
                    gix = (dne - dnw) * wy0 + (dse - dsw) * wy1;
                    giy = (dsw - dnw) * wx0 + (dse - dne) * wx1;
                } else {
                    float gfull[K];
#pragma unroll
                    for (int k = 0; k < K; ++k) gfull[k] = z[k] - (k == il ? cce : 0.0f);
                    coord_grad_px<T, K>(src_lay, cc, tp, gfull, gix, giy);
                }
                VLG_PROF(4)
                if (p.d_out_lay && !(VLG_ABL & 1)) {
                    // one contiguous row of d(loss)/d(warped layout) leaves through shared memory
                    if (lane == 0) bulk_store_wait_read();        // the previous row's bulk store has read obuf
                    __syncwarp();
                    VLG_PROF(12)
                    float *ob = sm.obuf + lane * K;
#pragma unroll
                    for (int k = 0; k < K; k += 4) *reinterpret_cast<float4 *>(ob + k) = make_float4(z[k], z[k + 1], z[k + 2], z[k + 3]);
                    if (lab_ok) ob[il] = gl;
                    VLG_PROF(11)
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0)
                        bulk_store(p.d_out_lay + (img + (int64_t)y * W + (int64_t)s * kLW) * K, sm.obuf, (unsigned)(npx * K * 4));
                }
            }
            VLG_PROF(5)

            // ---- coordinate gradient: rgb part (prefetched) + layout part + TV ----
            float gx = fmaf(tp.mx, gix, dc.x), gy = fmaf(tp.my, giy, dc.y);
            if (p.do_tv && !(VLG_ABL & 2)) {
                // horizontal neighbours by shuffle; the strip's edge lanes carry theirs in the prefetch pipeline
                float2 fr, flf;
                fr.x = __shfl_down_sync(FULL, f.x, 1); fr.y = __shfl_down_sync(FULL, f.y, 1);
                flf.x = __shfl_up_sync(FULL, f.x, 1);  flf.y = __shfl_up_sync(FULL, f.y, 1);
                if (lane == 31) fr = fedge;
                if (lane == 0) flf = fedge;
                const bool has_dn = y + 1 < H, has_up = y >= 1;      // warp-uniform
                const float cD = has_dn ? cV : 0.f, cU = has_up ? cV : 0.f;
                const float2 dd = make_float2(fq[1].x - f.x, fq[1].y - f.y), du = make_float2(f.x - fl_prev.x, f.y - fl_prev.y);
                const float2 dr = make_float2(fr.x - f.x, fr.y - f.y), dl = make_float2(f.x - flf.x, f.y - flf.y);
                s_tvh = fmaf(fabsf(dd.x) + fabsf(dd.y), has_dn ? mCol : 0.f, s_tvh);
                s_tvw = fmaf(fabsf(dr.x) + fabsf(dr.y), mRt, s_tvw);
                gx += (signed_c1(cU, du.x) - signed_c1(cD, dd.x)) + (signed_c1(cL, dl.x) - signed_c1(cR, dr.x));
                gy += (signed_c1(cU, du.y) - signed_c1(cD, dd.y)) + (signed_c1(cL, dl.y) - signed_c1(cR, dr.y));
            }
            if (GRAD && p.d_coords && col_ok && !(VLG_ABL & 16))
                reinterpret_cast<float2 *>(p.d_coords)[img + (int64_t)y * W + x] = make_float2(gx, gy);
            VLG_PROF(6)

            // ---- slide the pipelines of COMPUTED values (moves of landed data are harmless) ----
            fl_prev = f;
#pragma unroll
            for (int i = 0; i < kLD; ++i) { fq[i] = fq[i + 1]; xyq[i] = xyq[i + 1]; wtq[i] = wtq[i + 1]; rlq[i] = rlq[i + 1]; }
        }
        flush_tile((yb - 1) / kTH);
        m_disp = fmaxf(m_disp, m_disp_lane);
        while (ready < top) wait_next();      // nothing may be in flight when the ring is restarted or released
    }
    if (GRAD && p.d_out_lay && lane == 0) bulk_store_wait_read();

    VLG_PROF_FLUSH
    // ---- per-warp partial sums ----
    s_ce = warp_sum(s_ce);
    s_tvh = warp_sum(s_tvh);
    s_tvw = warp_sum(s_tvw);
    m_grad = warp_max(m_grad);
    m_disp = warp_max(m_disp);
    if (lane == 0) {
        reinterpret_cast<float4 *>(p.partials)[gw] = make_float4(s_ce, s_tvh, s_tvw, 0.f);
        if (m_disp > 0.f) atomicMax(&p.hdr->maxdisp_bits, __float_as_uint(m_disp));
        if (m_grad > 0.f) atomicMax(&p.hdr->maxgrad_bits, __float_as_uint(m_grad));
        if (gw == 0) p.hdr->n_lay = gridDim.x * kLayWarps;
        __threadfence();
    }

    // ---- the last CTA to finish reduces every partial row (fixed order) ----
    if (p.red.out != nullptr) {
        __shared__ int s_last;
        __syncthreads();
        if (threadIdx.x == 0) s_last = atomicAdd(&p.hdr->blocks_done, 1u) == gridDim.x - 1;
        __syncthreads();
        if (s_last) {
            __threadfence();
            reduce_partials_block<kLayThreads>(p.red, reinterpret_cast<double *>(smem_raw));
        }
    }
}

}  // namespace vlg
