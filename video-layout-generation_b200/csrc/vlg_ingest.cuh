// vlg_ingest.cuh -- what the dataset holds -> what the path consumes, in one pass per tensor.
//
// Reference (gongaa/video-layout-generation):
//   src/folder.py:122-127    cv2.imread + BGR2RGB              frames are uint8 [H,W,3]
//   src/data.py:33-35        transforms.ToTensor()             float32 CHW = uint8 / 255  (one IEEE division)
//   src/trainer.py:193-195   (frame - img_mean_arr) / img_std_arr   (sub, then div: one rounding each)
//   src/trainer.py:200-206   torch.flip(frame, [3]); torch.flip(seg3, [2])
//   src/folder.py:95-100     class maps are uint8 [H,W]; seg1/seg2 = .float().unsqueeze(0), seg3 = .long()
//   src/models/net_utils.py:14-24  transform_seg_one_hot: torch.eye(K)[seg.long()].permute(0,3,1,2)
// The reference uploads the float32 / int64 tensors (12 + 12 + 4 + 8 B/px per frame pair) and then runs four
// elementwise launches; here the host uploads the uint8 data (3 + 1 B/px) and the device expands it once:
// bit-identical to the torch expressions (every op is a single IEEE round-to-nearest, no FMA contraction).
#pragma once
#include "vlg_device.cuh"

namespace vlg {

struct IngestNorm {
    float mean[3], std[3];
    int normalize;          // 0: out = u8 / 255 (ToTensor only)
};

__device__ __forceinline__ float ingest1(const IngestNorm &nm, int c, unsigned u) {
    const float t = __fdiv_rn((float)u, 255.0f);                                   // ToTensor
    return nm.normalize ? __fdiv_rn(__fsub_rn(t, nm.mean[c]), nm.std[c]) : t;      // src/trainer.py:193-195
}

// uint8 [N,H,W,3] -> T [N,H,W,3]; one thread per group of four pixels of a row (W % 4 == 0): three 32-bit
// loads, three 128-bit (fp32) / 64-bit (bf16) stores.
template <typename T>
__global__ void __launch_bounds__(256) ingest_frames_vec4_kernel(IngestNorm nm, int64_t groups, int W, int flip,
                                                                 const uint8_t *__restrict__ in, T *__restrict__ out) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= groups) return;
    const int gpr = W >> 2;
    const int64_t row = g / gpr;
    const int x = (int)(g - row * gpr) << 2;
    const uint32_t *q = reinterpret_cast<const uint32_t *>(in + (row * W + x) * 3);
    const uint32_t w0 = __ldg(q), w1 = __ldg(q + 1), w2 = __ldg(q + 2);
    unsigned b[12];
#pragma unroll
    for (int i = 0; i < 4; ++i) { b[i] = (w0 >> (8 * i)) & 255u; b[4 + i] = (w1 >> (8 * i)) & 255u; b[8 + i] = (w2 >> (8 * i)) & 255u; }
    float o[12];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < 3; ++c) o[(flip ? 3 - i : i) * 3 + c] = ingest1(nm, c, b[i * 3 + c]);
    const int xo = flip ? W - 4 - x : x;
    store_px<T, 12>(out + (row * W + xo) * 3, o);
}

template <typename T>
__global__ void __launch_bounds__(256) ingest_frames_px_kernel(IngestNorm nm, int64_t P, int W, int flip, const uint8_t *__restrict__ in,
                                                               T *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const int64_t row = i / W;
    const int x = (int)(i - row * W);
    const int xo = flip ? W - 1 - x : x;
#pragma unroll
    for (int c = 0; c < 3; ++c) out[(row * W + xo) * 3 + c] = from_f<T>(ingest1(nm, c, __ldg(in + i * 3 + c)));
}

// uint8 class map [N,H,W] -> any of: int64 labels (seg3.long()), float32 class ids (seg1.float()), one-hot layout
// [N,H,W,K] of type T.  VEC (K * sizeof(T) a multiple of 16): warp-per-32-pixels, see below (the first version -- one
// thread per 16-byte chunk, 64-bit indices divided by run-time values -- was issue-bound: 62 us for the 168 MB of a C2
// one-hot layout).  Otherwise: one thread per pixel.
template <typename T, int K, bool VEC>
__global__ void __launch_bounds__(256) ingest_seg_kernel(int64_t P, int W, int flip, const uint8_t *__restrict__ seg, int64_t *__restrict__ out_label,
                                                         float *__restrict__ out_f32, T *__restrict__ out_onehot, WsHeader *hdr_or_null) {
    if constexpr (VEC) {
        // One WARP per 32 consecutive output pixels: every lane reads ONE class id, then the warp writes the 32 * CPP
        // 16-byte chunks of those pixels in CPP fully coalesced rounds; chunk e of the run belongs to pixel e / CPP, whose
        // class id comes from that lane by shuffle (all divisions by compile-time constants on values < 32 * CPP).
        constexpr int EPC = 16 / (int)sizeof(T), CPP = K / EPC;
        static_assert(K % EPC == 0, "vector path needs whole 16-byte chunks");
        const unsigned lane = threadIdx.x & 31;
        const int64_t px0 = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32;
        if (px0 >= P) return;                                             // warp-uniform
        const unsigned px = (unsigned)px0 + lane;                         // N*H*W < 2^31
        int l = 0;
        if (px < (uint64_t)P) {
            unsigned sp = px;
            if (flip) {
                const unsigned row = px / (unsigned)W, x = px - row * (unsigned)W;
                sp = row * (unsigned)W + ((unsigned)W - 1u - x);
            }
            l = (int)__ldg(seg + sp);
            if (out_label) out_label[px] = (int64_t)l;
            if (out_f32) out_f32[px] = (float)l;
            if (l >= K && hdr_or_null) atomicOr(&hdr_or_null->status, VLG_STATUS_BAD_LABEL);
        }
        const int npx = (int)min((int64_t)32, P - px0);
        uint4 *dst = reinterpret_cast<uint4 *>(out_onehot) + px0 * CPP;
#pragma unroll
        for (int it = 0; it < CPP; ++it) {
            const int e = it * 32 + (int)lane, q = e / CPP, chunk = e - q * CPP;
            const int lq = __shfl_sync(0xffffffffu, l, q);
            if (q < npx) {
                uint4 r = make_uint4(0u, 0u, 0u, 0u);
                T *v = reinterpret_cast<T *>(&r);
                const int rel = lq - chunk * EPC;
                if (rel >= 0 && rel < EPC) v[rel] = from_f<T>(1.0f);
                dst[e] = r;
            }
        }
    } else {
        const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if (i >= P) return;
        const unsigned px = (unsigned)i;
        unsigned sp = px;
        if (flip) {
            const unsigned row = px / (unsigned)W, x = px - row * (unsigned)W;
            sp = row * (unsigned)W + ((unsigned)W - 1u - x);
        }
        const int l = (int)__ldg(seg + sp);
        if (out_label) out_label[px] = (int64_t)l;
        if (out_f32) out_f32[px] = (float)l;
        if (l >= K && out_onehot && hdr_or_null) atomicOr(&hdr_or_null->status, VLG_STATUS_BAD_LABEL);
        if (out_onehot) {
#pragma unroll
            for (int k = 0; k < K; ++k) out_onehot[(int64_t)px * K + k] = from_f<T>(l == k ? 1.0f : 0.0f);
        }
    }
}

}  // namespace vlg
