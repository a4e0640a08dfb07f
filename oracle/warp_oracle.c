/* ORACLE (test infrastructure, NOT product code) -- strict-fp32 C restatement of the warp path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library.
 *
 * The reference (gongaa/video-layout-generation) contains no warp; its arithmetic for this path
 * lives in PyTorch (version unpinned upstream; pinned here to torch 2.11.0).  This file restates,
 * one rounding at a time, what `F.grid_sample(mode='bilinear', align_corners=True)` computes on
 * a grid built with the reference's own convention, so that "sampling indices bit-exact" has a
 * second witness that does not depend on torch being importable:
 *
 *   base grid          <- reference src/models/modules.py:69-70  (arange(S)/(S-1)*2-1)
 *   unnormalise        <- torch ATen/native/GridSampler.h:26-36   ((g+1)/2*(S-1))
 *   border clip        <- GridSampler.h:58-60
 *   zeros padding      <- GridSampler.h:205-207 (taps outside [0,S) contribute 0)
 *   tap weights        <- SURVEY.md Appendix A.5 (each product rounded once)
 *   accumulation       <- SURVEY.md Appendix A.6 (fma chain se<-sw<-ne<-nw)
 *   argmax             <- reference src/trainer.py:342 (first maximal index)
 *   backward           <- GridSampler.h:42-54,65-83 (d ix/d g = (S-1)/2; border clip zeroes it)
 *   loss terms         <- reference src/loss.py:16-25 (GD), :64-91 (SSIM);
 *                         src/trainer.py:124,130 (CE, L1)   [fp32 per pixel, fp64 accumulation]
 *
 * Parity status: warp / TV "parity unpinned" by the reference (absent there, no golden vectors);
 * pinned to torch 2.11.0 CPU outputs in tests/golden/ (tests/test_oracle.py checks bitwise).
 *
 * Build: gcc -O2 -fPIC -shared -ffp-contract=off -fno-fast-math -o libvlg_oracle.so warp_oracle.c -lm
 * Layout: NCHW contiguous fp32 (the reference's layout, src/trainer.py:197), grid [N,H,W,2].
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PAD_ZEROS 0
#define PAD_BORDER 1

/* reference src/models/modules.py:69: (arange(S).float() / (S-1)) * 2 - 1 */
static inline float base_coord(int i, int S) {
    float t = (float)i / (float)(S - 1);
    t = t * 2.0f;
    return t - 1.0f;
}

void vlgo_base_grid(int N, int H, int W, float *grid) {
    for (int n = 0; n < N; ++n)
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                float *g = grid + (((size_t)n * H + y) * W + x) * 2;
                g[0] = base_coord(x, W);
                g[1] = base_coord(y, H);
            }
}

/* grid = base + flow * scale, scale = fp32(2/(S-1)) rounded from double once */
void vlgo_flow_to_grid(const float *flow, int N, int H, int W, float *grid) {
    const float sx = (float)(2.0 / (double)(W - 1));
    const float sy = (float)(2.0 / (double)(H - 1));
    for (int n = 0; n < N; ++n)
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                size_t o = (((size_t)n * H + y) * W + x) * 2;
                float fu = flow[o] * sx;
                float fv = flow[o + 1] * sy;
                grid[o] = base_coord(x, W) + fu;
                grid[o + 1] = base_coord(y, H) + fv;
            }
}

/* GridSampler.h:31 then :58-60 */
static inline float source_index(float g, int S, int padding) {
    float c = ((g + 1.0f) / 2.0f) * (float)(S - 1);
    if (padding == PAD_BORDER) c = fminf((float)(S - 1), fmaxf(c, 0.0f));
    return c;
}

typedef struct {
    float ix, iy;
    int x0, y0;
    float nw, ne, sw, se;
} taps_t;

static inline taps_t make_taps(const float *g, int H, int W, int padding) {
    taps_t t;
    t.ix = source_index(g[0], W, padding);
    t.iy = source_index(g[1], H, padding);
    float fx0 = floorf(t.ix), fy0 = floorf(t.iy);
    float fx1 = fx0 + 1.0f, fy1 = fy0 + 1.0f;
    t.x0 = (int)fx0;
    t.y0 = (int)fy0;
    t.nw = (fx1 - t.ix) * (fy1 - t.iy);
    t.ne = (t.ix - fx0) * (fy1 - t.iy);
    t.sw = (fx1 - t.ix) * (t.iy - fy0);
    t.se = (t.ix - fx0) * (t.iy - fy0);
    return t;
}

/* Debug witness for the bit-exact index test: per output pixel (ix, iy) as fp32 bit patterns,
 * floor indices and the four weights. */
void vlgo_sample_coords(const float *grid, int N, int H, int W, int padding,
                        float *ixy /*[P,2]*/, int32_t *x0y0 /*[P,2]*/, float *w4 /*[P,4]*/) {
    size_t P = (size_t)N * H * W;
    for (size_t p = 0; p < P; ++p) {
        taps_t t = make_taps(grid + 2 * p, H, W, padding);
        ixy[2 * p] = t.ix; ixy[2 * p + 1] = t.iy;
        x0y0[2 * p] = t.x0; x0y0[2 * p + 1] = t.y0;
        w4[4 * p] = t.nw; w4[4 * p + 1] = t.ne; w4[4 * p + 2] = t.sw; w4[4 * p + 3] = t.se;
    }
}

static inline float tap_value(const float *plane, int y, int x, int H, int W) {
    return (y >= 0 && y < H && x >= 0 && x < W) ? plane[(size_t)y * W + x] : 0.0f;
}

void vlgo_warp_fwd(const float *src, const float *grid, int N, int C, int H, int W, int padding,
                   float *out) {
    for (int n = 0; n < N; ++n)
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                size_t p = ((size_t)n * H + y) * W + x;
                taps_t t = make_taps(grid + 2 * p, H, W, padding);
                for (int c = 0; c < C; ++c) {
                    const float *pl = src + ((size_t)n * C + c) * H * W;
                    float vnw = tap_value(pl, t.y0, t.x0, H, W);
                    float vne = tap_value(pl, t.y0, t.x0 + 1, H, W);
                    float vsw = tap_value(pl, t.y0 + 1, t.x0, H, W);
                    float vse = tap_value(pl, t.y0 + 1, t.x0 + 1, H, W);
                    float acc = vnw * t.nw;
                    acc = fmaf(vne, t.ne, acc);
                    acc = fmaf(vsw, t.sw, acc);
                    acc = fmaf(vse, t.se, acc);
                    out[((size_t)n * C + c) * H * W + (size_t)y * W + x] = acc;
                }
            }
}

/* reference src/trainer.py:342 */
void vlgo_argmax(const float *x, int N, int C, int H, int W, int64_t *out) {
    size_t HW = (size_t)H * W;
    for (int n = 0; n < N; ++n)
        for (size_t q = 0; q < HW; ++q) {
            const float *b = x + (size_t)n * C * HW + q;
            int best = 0;
            float bv = b[0];
            for (int c = 1; c < C; ++c)
                if (b[c * HW] > bv) { bv = b[c * HW]; best = c; }
            out[(size_t)n * HW + q] = best;
        }
}

/* grid_sample backward.  d_src accumulated in output-pixel order (p ascending, taps nw,ne,sw,se),
 * which is the fixed order the product's deterministic gather must be numerically close to.
 * d_grid follows GridSampler.h:42-54,65-83. */
void vlgo_warp_bwd(const float *src, const float *grid, const float *gout, int N, int C, int H, int W,
                   int padding, float *dsrc /*zeroed by callee*/, float *dgrid) {
    size_t HW = (size_t)H * W;
    memset(dsrc, 0, sizeof(float) * N * C * HW);
    const float gmx = (float)(W - 1) / 2.0f, gmy = (float)(H - 1) / 2.0f;
    for (int n = 0; n < N; ++n)
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                size_t p = ((size_t)n * H + y) * W + x;
                const float *g = grid + 2 * p;
                float ux = ((g[0] + 1.0f) / 2.0f) * (float)(W - 1);
                float uy = ((g[1] + 1.0f) / 2.0f) * (float)(H - 1);
                float mx = gmx, my = gmy;
                if (padding == PAD_BORDER) {
                    if (ux <= 0.0f || ux >= (float)(W - 1)) mx = 0.0f;
                    if (uy <= 0.0f || uy >= (float)(H - 1)) my = 0.0f;
                }
                taps_t t = make_taps(g, H, W, padding);
                float fx0 = (float)t.x0, fy0 = (float)t.y0, fx1 = fx0 + 1.0f, fy1 = fy0 + 1.0f;
                float gix = 0.0f, giy = 0.0f;
                for (int c = 0; c < C; ++c) {
                    const float *pl = src + ((size_t)n * C + c) * HW;
                    float *dp = dsrc + ((size_t)n * C + c) * HW;
                    float go = gout[((size_t)n * C + c) * HW + (size_t)y * W + x];
                    int xs[4] = {t.x0, t.x0 + 1, t.x0, t.x0 + 1};
                    int ys[4] = {t.y0, t.y0, t.y0 + 1, t.y0 + 1};
                    float ws[4] = {t.nw, t.ne, t.sw, t.se};
                    for (int k = 0; k < 4; ++k)
                        if (ys[k] >= 0 && ys[k] < H && xs[k] >= 0 && xs[k] < W)
                            dp[(size_t)ys[k] * W + xs[k]] += ws[k] * go;
                    float vnw = tap_value(pl, t.y0, t.x0, H, W);
                    float vne = tap_value(pl, t.y0, t.x0 + 1, H, W);
                    float vsw = tap_value(pl, t.y0 + 1, t.x0, H, W);
                    float vse = tap_value(pl, t.y0 + 1, t.x0 + 1, H, W);
                    gix -= vnw * (fy1 - t.iy) * go;  giy -= vnw * (fx1 - t.ix) * go;
                    gix += vne * (fy1 - t.iy) * go;  giy -= vne * (t.ix - fx0) * go;
                    gix -= vsw * (t.iy - fy0) * go;  giy += vsw * (fx1 - t.ix) * go;
                    gix += vse * (t.iy - fy0) * go;  giy += vse * (t.ix - fx0) * go;
                }
                dgrid[2 * p] = mx * gix;
                dgrid[2 * p + 1] = my * giy;
            }
}

/* ---- loss terms: fp32 per-pixel arithmetic, fp64 accumulation ---- */

/* reference src/trainer.py:130 */
double vlgo_l1(const float *a, const float *b, size_t numel) {
    double s = 0.0;
    for (size_t i = 0; i < numel; ++i) s += (double)fabsf(a[i] - b[i]);
    return s / (double)numel;
}

/* reference src/loss.py:20-25 */
double vlgo_gd(const float *a, const float *b, int N, int C, int H, int W) {
    double s = 0.0;
    for (int nc = 0; nc < N * C; ++nc) {
        const float *pa = a + (size_t)nc * H * W, *pb = b + (size_t)nc * H * W;
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                size_t o = (size_t)y * W + x;
                if (y + 1 < H) s += (double)fabsf(fabsf(pa[o + W] - pa[o]) - fabsf(pb[o + W] - pb[o]));
                if (x + 1 < W) s += (double)fabsf(fabsf(pa[o + 1] - pa[o]) - fabsf(pb[o + 1] - pb[o]));
            }
    }
    return s / ((double)N * C * H * W);
}

/* reference src/loss.py:68-91; window sums taken row-major like avg_pool2d, then /9 */
double vlgo_ssim(const float *x, const float *y, int N, int C, int H, int W) {
    const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
    double tot = 0.0;
    if (H < 3 || W < 3) return NAN;
    for (int c = 0; c < C; ++c) {
        double s = 0.0;
        for (int n = 0; n < N; ++n) {
            const float *px = x + ((size_t)n * C + c) * H * W, *py = y + ((size_t)n * C + c) * H * W;
            for (int i = 0; i + 2 < H; ++i)
                for (int j = 0; j + 2 < W; ++j) {
                    float sx = 0, sy = 0, sxx = 0, syy = 0, sxy = 0;
                    for (int di = 0; di < 3; ++di)
                        for (int dj = 0; dj < 3; ++dj) {
                            float a = px[(size_t)(i + di) * W + j + dj], b = py[(size_t)(i + di) * W + j + dj];
                            sx += a; sy += b; sxx += a * a; syy += b * b; sxy += a * b;
                        }
                    float mx = sx / 9.0f, my = sy / 9.0f;
                    float vx = sxx / 9.0f - mx * mx, vy = syy / 9.0f - my * my, vxy = sxy / 9.0f - mx * my;
                    float nn = (2.0f * mx * my + C1) * (2.0f * vxy + C2);
                    float dd = (mx * mx + my * my + C1) * (vx + vy + C2);
                    float v = (1.0f - nn / dd) / 2.0f;
                    v = fminf(1.0f, fmaxf(0.0f, v));
                    s += (double)v;
                }
        }
        tot += s / ((double)N * (H - 2) * (W - 2));
    }
    return tot;
}

/* reference src/trainer.py:124: mean over non-ignored pixels of -log_softmax[label] */
double vlgo_ce(const float *logits, const int64_t *label, int N, int C, int H, int W, int64_t ignore_index) {
    size_t HW = (size_t)H * W;
    double s = 0.0;
    size_t cnt = 0;
    for (int n = 0; n < N; ++n)
        for (size_t q = 0; q < HW; ++q) {
            int64_t l = label[(size_t)n * HW + q];
            if (l == ignore_index) continue;
            const float *b = logits + (size_t)n * C * HW + q;
            float m = b[0];
            for (int c = 1; c < C; ++c) m = fmaxf(m, b[c * HW]);
            double se = 0.0;
            for (int c = 0; c < C; ++c) se += exp((double)(b[c * HW] - m));
            s += log(se) - (double)(b[l * HW] - m);
            ++cnt;
        }
    return s / (double)cnt;
}

/* flow TV (absent upstream; stencils of src/loss.py:22,24 on the [N,H,W,2] flow) */
double vlgo_tv(const float *flow, int N, int H, int W) {
    double sh = 0.0, sw = 0.0;
    for (int n = 0; n < N; ++n)
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x)
                for (int k = 0; k < 2; ++k) {
                    size_t o = ((((size_t)n * H + y) * W + x) * 2) + k;
                    if (y + 1 < H) sh += (double)fabsf(flow[o + (size_t)W * 2] - flow[o]);
                    if (x + 1 < W) sw += (double)fabsf(flow[o + 2] - flow[o]);
                }
    return sh / ((double)N * (H - 1) * W * 2) + sw / ((double)N * H * (W - 1) * 2);
}
