// vlg_pass1.cuh -- pass 1 of the fused op: warp + every per-pixel loss term + d(loss)/d(warped)
// + d(loss)/d(coords) in ONE kernel, one CTA per 32x8 output tile.
//
// Loss definitions followed (reference gongaa/video-layout-generation):
//   L1          src/trainer.py:130,248      mean |a-b| over N*3*H*W, sign(0)=0
//   GD          src/loss.py:20-25           sum||da|-|db|| along H and along W, / (N*C*H*W)
//   SSIM        src/loss.py:68-91           3x3 valid avg-pool stats, C1=1e-4, C2=9e-4,
//                                           clamp((1-SSIM)/2,0,1).mean() summed over channels
//   CE          src/trainer.py:124,250      mean over labels != ignore_index of -log_softmax[label]
//   TV          (absent upstream)           mean|d_H flow| + mean|d_W flow|, stencils of loss.py:22,24
//   composition src/trainer.py:248-251      w_l1*L1 + w_style*(GD+SSIM) + w_ce*CE (+ w_tv*TV)
#pragma once
#include "vlg_device.cuh"

namespace vlg {

struct Pass1Params {
    CoordCfg cc;
    int N;
    int tiles_x, tiles_y;
    const void *src_rgb, *src_layout;  // WARP: sources.  !WARP: `output` rgb / logits themselves
    const float *coords;
    const void *tgt_rgb;
    const int64_t *label;
    int64_t ignore_index;
    // gradient scales (upstream grad 1.0), host-computed in double
    float c_l1, c_gd, c_ssim, c_tvh, c_tvw;
    float w_ce_over_scale;  // w_ce * (N_local/N_global); divided by n_valid on the device
    int need_grad;          // compute gradients at all
    int do_tv;
    uint32_t terms;         // VLG_TERM_* bits to evaluate
    // outputs
    float *d_coords;        // WARP && need_grad
    void *d_out_rgb;        // WARP: fp32 staging (nullable). !WARP: user tensor of type T (nullable)
    void *d_out_lay;
    int64_t *out_argmax;    // nullable
    float *partials;        // [n_blocks][kPartialSlots]
    WsHeader *hdr;
    uint32_t flags;
};

struct __align__(16) Pass1Smem {
    float a[3][kRN];        // warped (or given) rgb over the tile + halo 2
    float b[3][kRN];        // target rgb
    float k[3][3][kWN];     // per-window SSIM adjoint coefficients (A,B,C) x channel
    float red[kThreads / 32][kPartialSlots];
    float redmax[kThreads / 32][2];
};

template <typename T, int K, bool WARP>
__global__ void __launch_bounds__(kThreads) pass1_kernel(const Pass1Params p) {
    __shared__ Pass1Smem sm;
    const CoordCfg &cc = p.cc;
    const int H = cc.H, W = cc.W;
    const int tid = threadIdx.x;
    const int bt = blockIdx.x;
    const int n = bt / (p.tiles_x * p.tiles_y);
    const int trem = bt - n * (p.tiles_x * p.tiles_y);
    const int ty0 = (trem / p.tiles_x) * kTH, tx0 = (trem % p.tiles_x) * kTW;
    const int64_t img_px = (int64_t)n * H * W;

    const bool has_rgb = p.src_rgb != nullptr && p.tgt_rgb != nullptr;
    const bool has_lay = p.src_layout != nullptr && p.label != nullptr;
    const T *src_rgb = has_rgb ? reinterpret_cast<const T *>(p.src_rgb) + img_px * 3 : nullptr;
    const T *tgt_rgb = has_rgb ? reinterpret_cast<const T *>(p.tgt_rgb) + img_px * 3 : nullptr;
    const T *src_lay = has_lay ? reinterpret_cast<const T *>(p.src_layout) + img_px * K : nullptr;
    const float2 *coords = WARP ? reinterpret_cast<const float2 *>(p.coords) + img_px : nullptr;

    float s_l1 = 0.f, s_gd = 0.f, s_ssim = 0.f, s_ce = 0.f, s_tvh = 0.f, s_tvw = 0.f;
    float m_disp = 0.f, m_grad = 0.f;

    // ---------------- phase 0: rgb over the tile + halo ----------------
    if (has_rgb) {
        for (int q = tid; q < kRN; q += kThreads) {
            const int ry = q / kRW, rx = q - ry * kRW;
            const int y = ty0 - kHalo + ry, x = tx0 - kHalo + rx;
            float a[3] = {0.f, 0.f, 0.f}, b[3] = {0.f, 0.f, 0.f};
            if (y >= 0 && y < H && x >= 0 && x < W) {
                const int64_t o = (int64_t)y * W + x;
                if (WARP) {
                    const Taps t = make_taps(cc, __ldg(coords + o), y, x);
                    gather_px<T, 3>(src_rgb, cc, t, a);
                } else {
                    load_px<T, 3>(src_rgb + o * 3, a);
                }
                load_px<T, 3>(tgt_rgb + o * 3, b);
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                sm.a[c][q] = a[c];
                sm.b[c][q] = b[c];
            }
        }
        __syncthreads();

        // ---------------- phase 1: SSIM windows (top-left anchored) ----------------
        if (p.terms & VLG_TERM_SSIM)
        for (int w = tid; w < kWN; w += kThreads) {
            const int wy = w / kWW, wx = w - wy * kWW;
            const int i = ty0 - kHalo + wy, j = tx0 - kHalo + wx;  // top-left pixel of the window
            const bool valid = i >= 0 && j >= 0 && i + 2 < H && j + 2 < W;
            const bool own = valid && wy >= kHalo && wx >= kHalo;  // top-left inside this tile
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float kA = 0.f, kB = 0.f, kC = 0.f;
                if (valid) {
                    float sx = 0.f, sy = 0.f, sxx = 0.f, syy = 0.f, sxy = 0.f;
#pragma unroll
                    for (int di = 0; di < 3; ++di)
#pragma unroll
                        for (int dj = 0; dj < 3; ++dj) {
                            const float xa = sm.a[c][(wy + di) * kRW + wx + dj];
                            const float yb = sm.b[c][(wy + di) * kRW + wx + dj];
                            sx += xa;
                            sy += yb;
                            sxx = fmaf(xa, xa, sxx);
                            syy = fmaf(yb, yb, syy);
                            sxy = fmaf(xa, yb, sxy);
                        }
                    const float C1 = 1e-4f, C2 = 9e-4f, inv9 = 1.0f / 9.0f;
                    const float mx = sx * inv9, my = sy * inv9;
                    const float vx = sxx * inv9 - mx * mx, vy = syy * inv9 - my * my;
                    const float vxy = sxy * inv9 - mx * my;
                    const float n1 = 2.f * mx * my + C1, n2 = 2.f * vxy + C2;
                    const float d1 = mx * mx + my * my + C1, d2 = vx + vy + C2;
                    const float inv_d1 = 1.0f / d1, inv_d2 = 1.0f / d2;
                    const float S = (n1 * n2) * (inv_d1 * inv_d2);
                    const float v = (1.0f - S) * 0.5f;
                    if (own) s_ssim += fminf(1.0f, fmaxf(0.0f, v));
                    if (p.need_grad && v >= 0.0f && v <= 1.0f) {
                        // dS/dx_p = A + B x_p + C y_p; loss = (1-S)/2 / M
                        const float r = inv_d1 * inv_d2;
                        const float kk = -p.c_ssim * (2.0f / 9.0f);
                        kB = kk * (-S * inv_d2);
                        kC = kk * (n1 * r);
                        kA = kk * (my * n2 * r - my * n1 * r - S * mx * inv_d1 + S * mx * inv_d2);
                    }
                }
                sm.k[c][0][w] = kA;
                sm.k[c][1][w] = kB;
                sm.k[c][2][w] = kC;
            }
        }
        __syncthreads();
    }

    // ---------------- phase 2: own pixel ----------------
    const int ty = tid / kTW, tx = tid - ty * kTW;
    const int y = ty0 + ty, x = tx0 + tx;
    const bool inside = y < H && x < W;
    if (inside) {
        const int64_t o = (int64_t)y * W + x;
        Taps t;
        if (WARP) {
            t = make_taps(cc, __ldg(coords + o), y, x);
            m_disp = tap_displacement(cc, t, y, x);
        }
        float gix = 0.f, giy = 0.f;

        if (has_rgb) {
            const int q0 = (ty + kHalo) * kRW + tx + kHalo;
            float dr[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float a = sm.a[c][q0], b = sm.b[c][q0];
                const float d = a - b;
                float g = 0.f;
                if (p.terms & VLG_TERM_L1) {
                    s_l1 += fabsf(d);
                    g = p.c_l1 * sgn(d);
                }
                if (p.terms & VLG_TERM_GD) {
                // vertical pairs (reference `xloss`, src/loss.py:21-22): (y -> y+1) owned here
                if (y + 1 < H) {
                    const float da = sm.a[c][q0 + kRW] - a, db = sm.b[c][q0 + kRW] - b;
                    const float tt = fabsf(da) - fabsf(db);
                    s_gd += fabsf(tt);
                    g -= p.c_gd * sgn(tt) * sgn(da);
                }
                if (y >= 1) {
                    const float da = a - sm.a[c][q0 - kRW], db = b - sm.b[c][q0 - kRW];
                    g += p.c_gd * sgn(fabsf(da) - fabsf(db)) * sgn(da);
                }
                // horizontal pairs (reference `yloss`, src/loss.py:23-24)
                if (x + 1 < W) {
                    const float da = sm.a[c][q0 + 1] - a, db = sm.b[c][q0 + 1] - b;
                    const float tt = fabsf(da) - fabsf(db);
                    s_gd += fabsf(tt);
                    g -= p.c_gd * sgn(tt) * sgn(da);
                }
                if (x >= 1) {
                    const float da = a - sm.a[c][q0 - 1], db = b - sm.b[c][q0 - 1];
                    g += p.c_gd * sgn(fabsf(da) - fabsf(db)) * sgn(da);
                }
                }
                if (p.need_grad && (p.terms & VLG_TERM_SSIM)) {
                    // SSIM adjoint: the <=9 windows whose footprint contains this pixel
                    float sA = 0.f, sB = 0.f, sC = 0.f;
#pragma unroll
                    for (int di = 0; di < 3; ++di)
#pragma unroll
                        for (int dj = 0; dj < 3; ++dj) {
                            const int w = (ty + di) * kWW + tx + dj;
                            sA += sm.k[c][0][w];
                            sB += sm.k[c][1][w];
                            sC += sm.k[c][2][w];
                        }
                    g += sA + sB * a + sC * b;
                }
                dr[c] = g;
                m_grad = fmaxf(m_grad, fabsf(g));
            }
            if (p.need_grad) {
                if (WARP) {
                    coord_grad_px<T, 3>(src_rgb, cc, t, dr, gix, giy);
                    if (p.d_out_rgb) store_px<float, 3>(reinterpret_cast<float *>(p.d_out_rgb) + (img_px + o) * 3, dr);
                } else if (p.d_out_rgb) {
                    store_px<T, 3>(reinterpret_cast<T *>(p.d_out_rgb) + (img_px + o) * 3, dr);
                }
            }
        }

        if (has_lay) {
            float z[K];
            if (WARP) gather_px<T, K>(src_lay, cc, t, z);
            else load_px<T, K>(src_lay + o * K, z);
            float m = z[0];
            int best = 0;
#pragma unroll
            for (int k = 1; k < K; ++k)
                if (z[k] > m) { m = z[k]; best = k; }   // first maximal index (src/trainer.py:342)
            if (p.out_argmax) p.out_argmax[img_px + o] = best;
            const int64_t lab = __ldg(p.label + img_px + o);
            const bool lab_ok = lab >= 0 && lab < K;
            if (!lab_ok && lab != p.ignore_index) atomicOr(&p.hdr->status, VLG_STATUS_BAD_LABEL);
            float e[K], se = 0.f, zl = 0.f;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                e[k] = expf(z[k] - m);
                se += e[k];
                if (k == (int)lab) zl = z[k];
            }
            if (lab_ok) s_ce += (logf(se) + m) - zl;
            if (p.need_grad) {
                const float nv = (float)p.hdr->n_valid;
                const float cce = lab_ok ? p.w_ce_over_scale / nv : 0.0f;
                const float inv = cce / se;
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    e[k] = e[k] * inv - (k == (int)lab ? cce : 0.0f);
                    m_grad = fmaxf(m_grad, fabsf(e[k]));
                }
                if (WARP) {
                    coord_grad_px<T, K>(src_lay, cc, t, e, gix, giy);
                    if (p.d_out_lay) store_px<float, K>(reinterpret_cast<float *>(p.d_out_lay) + (img_px + o) * K, e);
                } else if (p.d_out_lay) {
                    store_px<T, K>(reinterpret_cast<T *>(p.d_out_lay) + (img_px + o) * K, e);
                }
            }
        }

        if (WARP) {
            float gx = t.mx * gix, gy = t.my * giy;
            if (p.do_tv) {
                const float2 f = __ldg(coords + o);
                if (y + 1 < H) {
                    const float2 f1 = __ldg(coords + o + W);
                    const float dx = f1.x - f.x, dy = f1.y - f.y;
                    s_tvh += fabsf(dx) + fabsf(dy);
                    gx -= p.c_tvh * sgn(dx);
                    gy -= p.c_tvh * sgn(dy);
                }
                if (y >= 1) {
                    const float2 f1 = __ldg(coords + o - W);
                    gx += p.c_tvh * sgn(f.x - f1.x);
                    gy += p.c_tvh * sgn(f.y - f1.y);
                }
                if (x + 1 < W) {
                    const float2 f1 = __ldg(coords + o + 1);
                    const float dx = f1.x - f.x, dy = f1.y - f.y;
                    s_tvw += fabsf(dx) + fabsf(dy);
                    gx -= p.c_tvw * sgn(dx);
                    gy -= p.c_tvw * sgn(dy);
                }
                if (x >= 1) {
                    const float2 f1 = __ldg(coords + o - 1);
                    gx += p.c_tvw * sgn(f.x - f1.x);
                    gy += p.c_tvw * sgn(f.y - f1.y);
                }
            }
            if (p.need_grad && p.d_coords)
                reinterpret_cast<float2 *>(p.d_coords)[img_px + o] = make_float2(gx, gy);
        }
    }

    // ---------------- phase 3: block reduction -> one partial row per CTA ----------------
    float vals[6] = {s_l1, s_gd, s_ssim, s_ce, s_tvh, s_tvw};
    const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const float r = warp_sum(vals[i]);
        if (lane == 0) sm.red[wid][i] = r;
    }
    m_disp = warp_max(m_disp);
    m_grad = warp_max(m_grad);
    if (lane == 0) {
        sm.redmax[wid][0] = m_disp;
        sm.redmax[wid][1] = m_grad;
    }
    __syncthreads();
    if (tid < kPartialSlots) {
        float r = 0.f;
        if (tid < 6)
#pragma unroll
            for (int w = 0; w < kThreads / 32; ++w) r += sm.red[w][tid];
        p.partials[(int64_t)bt * kPartialSlots + tid] = r;
    }
    if (tid == 32) {
        float d = 0.f, g = 0.f;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) {
            d = fmaxf(d, sm.redmax[w][0]);
            g = fmaxf(g, sm.redmax[w][1]);
        }
        if (d > 0.f) atomicMax(&p.hdr->maxdisp_bits, __float_as_uint(d));
        if (g > 0.f) atomicMax(&p.hdr->maxgrad_bits, __float_as_uint(g));
    }
}

}  // namespace vlg
