// Microbenchmark: how fast does ONE SM pull a ~67 KB window of an NHWC [H][W][20] fp32 tensor through TMA,
// as a function of the box shape?  (A) 4-D box {20, 38, 22}: 836 inner rows of 80 bytes;
// (B) 3-D view {W*20, H}: five boxes {160, 22}: 110 rows of 640 bytes;  (C) {240, 22} x 4 (12 px, 960-byte rows).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/tma_bench tools/diag/tma_bench.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <vector>

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("{\n.reg .pred p;\nWAIT_LOOP:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra WAIT_DONE;\nbra WAIT_LOOP;\nWAIT_DONE:\n}\n" ::"r"(a), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma4(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n" ::"r"(
                     (unsigned)__cvta_generic_to_shared(dst)), "l"(map), "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma3(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n" ::"r"(
                     (unsigned)__cvta_generic_to_shared(dst)), "l"(map), "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// mode 0: 4-D box {20, bw, bh}; mode 1: nchunk 3-D boxes {cw*20, bh}.  depth = loads in flight (1 or 2).
__global__ void __launch_bounds__(32) bench(const __grid_constant__ CUtensorMap map, int mode, int bw, int bh, int cw, int nchunk, int depth,
                                            int iters, int H, int W, int N, long long *cycles) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ alignas(8) uint64_t bar[2];
    const unsigned stage = (unsigned)(mode == 0 ? bw * bh * 80 : nchunk * cw * bh * 80);
    const unsigned stage_al = (stage + 127) / 128 * 128;
    if (threadIdx.x == 0) {
        mbar_init(&bar[0], 1); mbar_init(&bar[1], 1);
        unsigned rng = blockIdx.x * 2654435761u + 12345u;
        auto issue = [&](int b) {
            rng = rng * 1664525u + 1013904223u;
            const int x = (int)((rng >> 8) % (unsigned)(W - bw)), y = (int)((rng >> 20) % (unsigned)(H - bh)), n = (int)((rng >> 4) % (unsigned)N);
            mbar_expect_tx(&bar[b], stage);
            unsigned char *dst = smem + (size_t)b * stage_al;
            if (mode == 0) tma4(dst, &map, &bar[b], 0, x, y, n);
            else for (int c = 0; c < nchunk; ++c) tma3(dst + (size_t)c * cw * bh * 80, &map, &bar[b], (x + c * cw) * 20, y, n);
        };
        for (int d = 0; d < depth; ++d) issue(d);
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const int b = i % depth;
            mbar_wait(&bar[b], (unsigned)((i / depth) & 1));
            if (i + depth < iters) issue(b);
        }
        cycles[blockIdx.x] = clock64() - t0;
    }
}

int main() {
    const int N = 8, H = 256, W = 512, K = 20;     // 84 MB: L2-resident after the first touches (126 MB L2)
    float *d; cudaMalloc(&d, (size_t)N * H * W * K * 4); cudaMemset(d, 0, (size_t)N * H * W * K * 4);
    long long *cyc; cudaMalloc(&cyc, 256 * 8);
    void *f = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)f;
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const int bw = 40, bh = 22;
    struct Cfg { const char *name; int mode, cw, nchunk, depth; };
    const Cfg cfgs[] = {{"4D {20,40,22} 80B rows, 1 in flight", 0, 0, 0, 1}, {"4D {20,40,22} 80B rows, 2 in flight", 0, 0, 0, 2},
                        {"3D 5x{160,22} 640B rows, 1 in flight", 1, 8, 5, 1}, {"3D 5x{160,22} 640B rows, 2 in flight", 1, 8, 5, 2},
                        {"3D 4x{200,22} 800B rows, 1 in flight", 1, 10, 4, 1}, {"3D 4x{200,22} 800B rows, 2 in flight", 1, 10, 4, 2},
                        {"3D 20x{40,22} 160B rows, 2 in flight", 1, 2, 20, 2}};
    for (const Cfg &c : cfgs) {
        CUtensorMap map; memset(&map, 0, sizeof(map));
        CUresult r;
        if (c.mode == 0) {
            const cuuint64_t gdim[4] = {(cuuint64_t)K, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
            const cuuint64_t gstr[3] = {(cuuint64_t)K * 4, (cuuint64_t)W * K * 4, (cuuint64_t)H * W * K * 4};
            const cuuint32_t box[4] = {(cuuint32_t)K, (cuuint32_t)bw, (cuuint32_t)bh, 1u}, es[4] = {1, 1, 1, 1};
            r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        } else {
            const cuuint64_t gdim[3] = {(cuuint64_t)W * K, (cuuint64_t)H, (cuuint64_t)N};
            const cuuint64_t gstr[2] = {(cuuint64_t)W * K * 4, (cuuint64_t)H * W * K * 4};
            const cuuint32_t box[3] = {(cuuint32_t)(c.cw * K), (cuuint32_t)bh, 1u}, es[3] = {1, 1, 1};
            r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        }
        if (r != CUDA_SUCCESS) { printf("%s: encode failed %d\n", c.name, (int)r); continue; }
        const size_t stage = (size_t)bw * bh * 80, smem = 2 * ((stage + 127) / 128 * 128);
        cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        const int iters = 200;
        for (int rep = 0; rep < 2; ++rep) bench<<<sms, 32, smem, 0>>>(map, c.mode, bw, bh, c.cw, c.nchunk, c.depth, iters, H, W, N, cyc);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<long long> h(sms);
        cudaMemcpy(h.data(), cyc, sms * 8, cudaMemcpyDeviceToHost);
        double avg = 0; for (long long v : h) avg += (double)v / sms;
        printf("%-44s %s  %.0f cycles/window  %.1f B/clk/SM  (all %d SMs busy: %.2f TB/s aggregate at %.2f GHz)\n", c.name, cudaGetErrorString(e),
               avg / iters, stage / (avg / iters), sms, stage / (avg / iters) * sms * khz * 1e3 / 1e12, khz / 1e6);
    }
    return 0;
}
