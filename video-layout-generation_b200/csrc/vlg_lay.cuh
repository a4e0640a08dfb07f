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
    int accum_dcoords;             // d_coords already holds the rgb part
    float *d_coords;               // nullable (validation)
    float *d_out_lay;              // [P][K] fp32 staging, nullable
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

__device__ __forceinline__ void bulk_store(void *gdst, const void *ssrc, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(gdst),
                 "r"((unsigned)__cvta_generic_to_shared(ssrc)), "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
}
__device__ __forceinline__ void bulk_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

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
    constexpr int OFF = 64;   // keeps (row + OFF) non-negative: tap rows are clamped to >= -4

    if (lane < kLR) mbar_init(&sm.bar[lane], 1);
    __syncwarp();
    unsigned issue_par = 0u;  // bit s: parity the NEXT load issued on slot s will complete
    unsigned in_flight = 0u;  // bit s: a load was issued on slot s and nobody has waited for it yet
    // every issued load is waited for exactly once (before its slot is re-armed, when a row samples it,
    // or at the end of the segment), so an mbarrier never has two loads outstanding in one phase and no
    // TMA write is in flight when the CTA's shared memory is reused or released
    auto wait_slot = [&](int sl) {
        mbar_wait(&sm.bar[sl], ((issue_par >> sl) & 1u) ^ 1u);
        in_flight &= ~(1u << sl);
    };

    int64_t rho = (int64_t)gw * p.chunk;
    const int64_t rho_end = min(rho + p.chunk, p.total_rows);
    float s_ce = 0.f, s_tvh = 0.f, s_tvw = 0.f;
    float m_disp = 0.f, m_grad = 0.f;
    const float denom = GRAD ? (p.weighted_denom ? (float)__ldcg(&p.hdr->ce_denom) : (float)__ldcg(&p.hdr->n_valid)) : 1.0f;
    const T *src_all = reinterpret_cast<const T *>(p.src_layout);

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
        const int npx = min(kLW, W - s * kLW);

        // ring state (warp-uniform)
        int top = INT_MIN, lo = INT_MIN;        // source rows [max(lo, top - kLR), top) are resident / in flight
        int pend_ymin[kLD + 1];                 // lowest source row needed by output rows t .. t+kLD
#pragma unroll
        for (int i = 0; i <= kLD; ++i) pend_ymin[i] = INT_MAX;
        int row_ymin[kLD + 1], row_ymax[kLD + 1];   // bounding rows of output rows t .. t+kLD
#pragma unroll
        for (int i = 0; i <= kLD; ++i) { row_ymin[i] = INT_MAX; row_ymax[i] = INT_MIN; }

        auto load_flow = [&](int t) -> float2 {
            return (t >= 0 && t < H) ? __ldg(coords + (t * W + xc)) : make_float2(0.f, 0.f);
        };
        // plan + issue the ring loads output row u needs (stage B)
        auto plan_row = [&](int u, float2 fl, int &ymin_o, int &ymax_o) {
            ymin_o = INT_MAX; ymax_o = INT_MIN;
            if (u < ya || u >= yb) return;
            float mx, my;
            const float2 xy = source_xy(cc, fl, bxv, base_coord(u, cc.Hm1), mx, my);
            const int x0 = (int)fminf(fmaxf(floorf(xy.x), -4.0f), (float)W + 4.0f);
            const int y0 = (int)fminf(fmaxf(floorf(xy.y), -4.0f), (float)H + 4.0f);
            int xmin = __reduce_min_sync(FULL, col_ok ? x0 : INT_MAX), xmax = __reduce_max_sync(FULL, col_ok ? x0 : INT_MIN);
            int ymin = __reduce_min_sync(FULL, col_ok ? y0 : INT_MAX), ymax = __reduce_max_sync(FULL, col_ok ? y0 : INT_MIN);
            if (ymax - ymin > kLR - 2 || xmax - xmin > kLBW - 2) {
                // outliers: keep the lanes that move with the strip's centre lane
                const int rx = __shfl_sync(FULL, x0 - x, min(15, npx - 1)), ry = __shfl_sync(FULL, y0, min(15, npx - 1));
                const bool in = col_ok && abs(y0 - ry) <= (kLR - 2) / 2 && abs((x0 - x) - rx) <= (kLBW - kLW - 2) / 2;
                xmin = __reduce_min_sync(FULL, in ? x0 : INT_MAX); xmax = __reduce_max_sync(FULL, in ? x0 : INT_MIN);
                ymin = __reduce_min_sync(FULL, in ? y0 : INT_MAX); ymax = __reduce_max_sync(FULL, in ? y0 : INT_MIN);
            }
            ymin_o = ymin; ymax_o = ymax;
            // rows [ymin, ymax + 1], columns [xmin, xmax + 1]; slack split evenly on both sides
            const int ox_new = xmin - (kLBW - (xmax - xmin + 2)) / 2;
            if (top == INT_MIN || ymin > top || ymin < lo) {   // first use, or the flow jumped: restart the resident range
                top = ymin; lo = ymin;
            }
            int low_needed = ymin;
#pragma unroll
            for (int i = 0; i <= kLD; ++i) low_needed = min(low_needed, pend_ymin[i]);
            __syncwarp();
            while (top <= ymax + 1 && top - kLR < low_needed) {
                const int sl = (top + OFF) % kLR;
                if ((in_flight >> sl) & 1u) wait_slot(sl);
                if (lane == 0) {
                    sm.ox[sl] = ox_new;
                    mbar_expect_tx(&sm.bar[sl], kRowBytes);
                    tma_load_4d(sm.ring[sl], &lay_map, &sm.bar[sl], 0, ox_new, top, n);
                }
                issue_par ^= 1u << sl;
                in_flight |= 1u << sl;
                ++top;
            }
            __syncwarp();
        };

        // ---- prologue: flows of rows ya .. ya+kLD+1, plans of rows ya .. ya+kLD-1 ----
        int t = ya;
        float2 fl[kLD + 2];
#pragma unroll
        for (int i = 0; i < kLD + 2; ++i) fl[i] = load_flow(t + i);
        int64_t lab[kLD + 1];
        float2 dcp[kLD + 1];
        auto load_aux = [&](int u, int64_t &l, float2 &d) {
            l = 0; d = make_float2(0.f, 0.f);
            if (u >= ya && u < yb) {
                l = __ldg(p.label + img + (u * W + xc));
                if (GRAD && p.accum_dcoords) d = __ldcg(reinterpret_cast<const float2 *>(p.d_coords) + img + (u * W + xc));
            }
        };
#pragma unroll
        for (int i = 0; i <= kLD; ++i) load_aux(t + i, lab[i], dcp[i]);
#pragma unroll
        for (int i = 0; i < kLD; ++i) {
            plan_row(t + i, fl[i], row_ymin[i], row_ymax[i]);
            pend_ymin[i] = row_ymin[i];
        }
        float2 fl_prev = load_flow(t - 1);      // TV stencil
        int tile_row = -1;                      // 8-row tile the running maximum below belongs to
        float m_near_tile = 0.f;

#pragma unroll 1
        for (; t < yb; ++t) {
            // ---- stage A/B: plan row t+kLD, load flow of row t+kLD+2, aux of row t+kLD+1 ----
            plan_row(t + kLD, fl[kLD], row_ymin[kLD], row_ymax[kLD]);
            pend_ymin[kLD] = row_ymin[kLD];
            const float2 fl_new = load_flow(t + kLD + 2);
            int64_t lab_new; float2 dc_new;
            load_aux(t + kLD + 1, lab_new, dc_new);

            // ---- stage C: output row t ----
            const float2 f = fl[0];
            float mx, my;
            const float2 xy = source_xy(cc, f, bxv, base_coord(t, cc.Hm1), mx, my);
            const Taps tp = taps_from_xy(cc, xy, mx, my);
            const int y = t;

            // displacement bookkeeping for pass 2 (near radius per tile, far queue)
            float disp = col_ok ? tap_displacement(cc, tp, y, x) : 0.f;
            const bool is_far = disp >= (float)VLG_NEAR_RADIUS;
            {
                const unsigned dmax = __reduce_max_sync(FULL, __float_as_uint(disp));
                m_disp = fmaxf(m_disp, __uint_as_float(dmax));
                const unsigned nmax = __reduce_max_sync(FULL, __float_as_uint(is_far ? 0.f : disp));
                const int tr = y / kTH;
                if (tr != tile_row) {
                    if (tile_row >= 0 && lane == 0 && m_near_tile > 0.f)
                        atomicMax(reinterpret_cast<unsigned *>(p.tile_disp) + ((int64_t)n * p.tiles_y + tile_row) * p.strips + s,
                                  __float_as_uint(m_near_tile));
                    tile_row = tr; m_near_tile = 0.f;
                }
                m_near_tile = fmaxf(m_near_tile, __uint_as_float(nmax));
            }
            if (is_far && p.d_out_lay != nullptr) {
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

            // wait for the ring rows this output row samples
            {
                const int r0 = max(row_ymin[0], max(lo, top - kLR)), r1 = min(row_ymax[0] + 1, top - 1);
                for (int r = r0; r <= r1; ++r) wait_slot((r + OFF) % kLR);
            }

            const int64_t lb = lab[0];
            const bool lab_ok = lb >= 0 && lb < K;
            if (col_ok && !lab_ok && lb != p.ignore_index) atomicOr(&p.hdr->status, VLG_STATUS_BAD_LABEL);
            const int il = lab_ok ? (int)lb : 0;
            float wl = 1.0f;
            if (p.class_weight && lab_ok) wl = __ldg(p.class_weight + il);

            float z[K], v[K];
            float vl[4];
            float zl;
            const T *st0 = nullptr, *st1 = nullptr;   // smem addresses of the nw / sw taps when resident
            {
                const int res_lo = max(lo, top - kLR);
                const int sl0 = (tp.y0 + OFF) % kLR, sl1 = (tp.y0 + 1 + OFF) % kLR;
                const int rx0 = tp.x0 - sm.ox[sl0], rx1 = tp.x0 - sm.ox[sl1];
                if (tp.y0 >= res_lo && tp.y0 + 1 < top && rx0 >= 0 && rx0 + 1 < kLBW && rx1 >= 0 && rx1 + 1 < kLBW) {
                    st0 = reinterpret_cast<const T *>(sm.ring[sl0]) + rx0 * K;
                    st1 = reinterpret_cast<const T *>(sm.ring[sl1]) + rx1 * K;
                }
            }
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
            float se = 0.f;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                z[k] = ex2_approx(fmaf(z[k], L2E, -ml2));
                se += z[k];
            }
            if (lab_ok && col_ok) s_ce += wl * (fmaf(lg2_approx(se), 0.6931471805599453f, m) - zl);

            float gix = 0.f, giy = 0.f;
            if (GRAD) {
                const float cce = (lab_ok && col_ok) ? p.w_ce_over_scale * wl / denom : 0.0f;
                const float inv = cce / se;
                mul2_bcast<K>(z, z, inv);
                const float gl = fmaf(ex2_approx(fmaf(zl, L2E, -ml2)), inv, -cce);
                m_grad = fmaxf(m_grad, cce);
                float dnw, dne, dsw, dse;
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
                if (p.d_out_lay) {
                    // one contiguous row of d(loss)/d(warped layout) leaves through shared memory
                    if (lane == 0) bulk_store_wait_read();        // the previous row's bulk store has read obuf
                    __syncwarp();
                    float *ob = sm.obuf + lane * K;
#pragma unroll
                    for (int k = 0; k < K; k += 4) *reinterpret_cast<float4 *>(ob + k) = make_float4(z[k], z[k + 1], z[k + 2], z[k + 3]);
                    if (lab_ok) ob[il] = gl;
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0)
                        bulk_store(p.d_out_lay + (img + (int64_t)y * W + (int64_t)s * kLW) * K, sm.obuf, (unsigned)(npx * K * 4));
                }
            }

            // ---- coordinate gradient: rgb part (prefetched) + layout part + TV ----
            float gx = fmaf(tp.mx, gix, dcp[0].x), gy = fmaf(tp.my, giy, dcp[0].y);
            if (p.do_tv) {
                // horizontal neighbours by shuffle; the strip's edge lanes read them from memory
                float2 fr, flf;
                fr.x = __shfl_down_sync(FULL, f.x, 1); fr.y = __shfl_down_sync(FULL, f.y, 1);
                flf.x = __shfl_up_sync(FULL, f.x, 1);  flf.y = __shfl_up_sync(FULL, f.y, 1);
                if (lane == 31 && x + 1 < W) fr = __ldg(coords + (y * W + x + 1));
                if (lane == 0 && x >= 1) flf = __ldg(coords + (y * W + x - 1));
                if (col_ok) {
                    if (y + 1 < H) {
                        const float2 df = make_float2(fl[1].x - f.x, fl[1].y - f.y);
                        s_tvh += fabsf(df.x) + fabsf(df.y);
                        gx -= signed_c1(p.c_tvh, df.x);
                        gy -= signed_c1(p.c_tvh, df.y);
                    }
                    if (y >= 1) {
                        gx += signed_c1(p.c_tvh, f.x - fl_prev.x);
                        gy += signed_c1(p.c_tvh, f.y - fl_prev.y);
                    }
                    if (x + 1 < W) {
                        const float2 df = make_float2(fr.x - f.x, fr.y - f.y);
                        s_tvw += fabsf(df.x) + fabsf(df.y);
                        gx -= signed_c1(p.c_tvw, df.x);
                        gy -= signed_c1(p.c_tvw, df.y);
                    }
                    if (x >= 1) {
                        gx += signed_c1(p.c_tvw, f.x - flf.x);
                        gy += signed_c1(p.c_tvw, f.y - flf.y);
                    }
                }
            }
            if (GRAD && p.d_coords && col_ok)
                reinterpret_cast<float2 *>(p.d_coords)[img + (int64_t)y * W + x] = make_float2(gx, gy);

            // ---- slide the pipeline ----
            fl_prev = f;
#pragma unroll
            for (int i = 0; i < kLD + 1; ++i) fl[i] = fl[i + 1];
            fl[kLD + 1] = fl_new;
#pragma unroll
            for (int i = 0; i < kLD; ++i) {
                lab[i] = lab[i + 1]; dcp[i] = dcp[i + 1];
                row_ymin[i] = row_ymin[i + 1]; row_ymax[i] = row_ymax[i + 1]; pend_ymin[i] = pend_ymin[i + 1];
            }
            lab[kLD] = lab_new; dcp[kLD] = dc_new;
        }
        if (tile_row >= 0 && lane == 0 && m_near_tile > 0.f)
            atomicMax(reinterpret_cast<unsigned *>(p.tile_disp) + ((int64_t)n * p.tiles_y + tile_row) * p.strips + s,
                      __float_as_uint(m_near_tile));
#pragma unroll
        for (int sl = 0; sl < kLR; ++sl)
            if ((in_flight >> sl) & 1u) wait_slot(sl);
    }
    if (GRAD && p.d_out_lay && lane == 0) bulk_store_wait_read();

    // ---- per-warp partial sums ----
    s_ce = warp_sum(s_ce);
    s_tvh = warp_sum(s_tvh);
    s_tvw = warp_sum(s_tvw);
    m_grad = warp_max(m_grad);
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
