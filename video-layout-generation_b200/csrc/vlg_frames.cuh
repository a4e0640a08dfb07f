// vlg_frames.cuh -- boundary fusions around the path: the per-channel affine
// renormalisation of rgb frames, the horizontal-flip augmentation and the NCHW -> NHWC re-layout in ONE pass.
//
// Reference (gongaa/video-layout-generation):
//   src/trainer.py:122-123   img_std_arr / img_mean_arr ([None,:,None,None] broadcasts)
//   src/trainer.py:193-195   frame = (frame - img_mean_arr) / img_std_arr          (mode NORMALIZE)
//   src/trainer.py:212,324   img = (img - mean_arr) / std_arr                      (mode NORMALIZE)
//   src/trainer.py:215       g_img = img * img_std_arr + img_mean_arr              (mode DENORMALIZE)
//   src/trainer.py:200-206   torch.flip(frame, [3]); torch.flip(seg3, [2])         (flip_w)
// In torch each of these is a full-tensor elementwise launch (plus a contiguous() copy for channels_last);
// here a frame is read once (12 B/px fp32) and written once (12 or 6 B/px): HBM-bound by construction.
// Arithmetic is the reference's, one IEEE rounding per op (sub then div; mul then add -- no FMA), so the
// result is bit-identical to torch's.
#pragma once
#include "vlg_device.cuh"

namespace vlg {

struct FrameAffine {
    float a[3], b[3];   // NORMALIZE: (x - a) / b     DENORMALIZE: x * b + a
    int denorm;
};

__device__ __forceinline__ float frame_affine1(const FrameAffine &fa, int c, float x) {
    return fa.denorm ? __fadd_rn(__fmul_rn(x, fa.b[c]), fa.a[c]) : __fdiv_rn(__fsub_rn(x, fa.a[c]), fa.b[c]);
}

// One thread per group of four consecutive pixels of a row (W % 4 == 0, 16-byte aligned bases): three
// 128-bit plane loads (NCHW) or three 128-bit pixel-group loads (NHWC), three 128-bit (fp32) stores.
template <typename T, bool IN_NCHW>
__global__ void __launch_bounds__(256) frame_affine_vec4_kernel(FrameAffine fa, int64_t groups, int H, int W, int flip,
                                                                const float *__restrict__ in, T *__restrict__ out) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= groups) return;
    const int gpr = W >> 2;                                   // groups per row
    const int64_t row = g / gpr;                              // n * H + y
    const int x = (int)(g - row * gpr) << 2;
    float v[4][3];                                            // [pixel][channel]
    if (IN_NCHW) {
        const int64_t n = row / H;
        const int y = (int)(row - n * H);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float4 q = __ldg(reinterpret_cast<const float4 *>(in + ((n * 3 + c) * H + y) * (int64_t)W + x));
            v[0][c] = q.x; v[1][c] = q.y; v[2][c] = q.z; v[3][c] = q.w;
        }
    } else {
        const float4 *q = reinterpret_cast<const float4 *>(in + (row * W + x) * 3);
        const float4 q0 = __ldg(q), q1 = __ldg(q + 1), q2 = __ldg(q + 2);
        v[0][0] = q0.x; v[0][1] = q0.y; v[0][2] = q0.z; v[1][0] = q0.w;
        v[1][1] = q1.x; v[1][2] = q1.y; v[2][0] = q1.z; v[2][1] = q1.w;
        v[2][2] = q2.x; v[3][0] = q2.y; v[3][1] = q2.z; v[3][2] = q2.w;
    }
    float o[12];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < 3; ++c) o[(flip ? 3 - i : i) * 3 + c] = frame_affine1(fa, c, v[i][c]);
    const int xo = flip ? W - 4 - x : x;
    store_px<T, 12>(out + (row * W + xo) * 3, o);
}

// Any width / alignment: one thread per pixel.
template <typename T, bool IN_NCHW>
__global__ void __launch_bounds__(256) frame_affine_px_kernel(FrameAffine fa, int64_t P, int H, int W, int flip,
                                                              const float *__restrict__ in, T *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const int64_t row = i / W;
    const int x = (int)(i - row * W);
    const int64_t n = row / H;
    const int y = (int)(row - n * H);
    const int xo = flip ? W - 1 - x : x;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float s = IN_NCHW ? __ldg(in + ((n * 3 + c) * H + y) * (int64_t)W + x) : __ldg(in + i * 3 + c);
        out[(row * W + xo) * 3 + c] = from_f<T>(frame_affine1(fa, c, s));
    }
}

// torch.flip(seg3, [2]) on [N,H,W] int64 labels (src/trainer.py:206)
__global__ void __launch_bounds__(256) flip_labels_kernel(int64_t P, int W, const int64_t *__restrict__ in, int64_t *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const int64_t row = i / W;
    const int x = (int)(i - row * W);
    out[row * W + (W - 1 - x)] = __ldg(in + i);
}

}  // namespace vlg
