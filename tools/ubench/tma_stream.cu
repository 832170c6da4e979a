// Micro-benchmark: how fast can one persistent CTA per SM stream the increment-Gram chunk buffer into shared memory?
// Decides the global layout / copy instruction of the recursion kernel's feed.  No arithmetic: consumers only wait for
// a stage, read 16 B per lane and hand the stage back.
//   mode 0: 5-D tensor-map TMA, 2 KB box, layout [i][s][j][P]   (row stride = nj * P * 4 bytes, what v1/v2 use)
//   mode 1: same box, layout [i][jg][s][G][P]                    (a warp's rows are contiguous in memory)
//   mode 2: cp.async.bulk 1-D copies of STAGE bytes, layout as mode 1
// usage: tma_stream mode stage_bytes stages_per_warp consumer_warps [self_producer]
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t par) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(par) : "memory");
    }
}
__device__ __forceinline__ void tma5(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst), "l"(m), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void bulk1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

struct P {
    const float* buf;
    int mode, stage_bytes, S, rows, njg, ni;  // items = ni * njg, each `rows` rows of 2 KB
    long long nitems;
    float* sink;
};

__global__ void __maxnreg__(128) stream_kernel(const __grid_constant__ CUtensorMap tmap, const P p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int ncw = (blockDim.x >> 5) - 1, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, S = p.S;
    const uint32_t s0 = smem_u32(smem), full0 = s0 + ncw * S * p.stage_bytes, empty0 = full0 + ncw * S * 8;
    if (threadIdx.x == 0) {
        for (int k = 0; k < ncw * S; ++k) { mbar_init(full0 + 8 * k, 1); mbar_init(empty0 + 8 * k, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long NW = (long long)gridDim.x * ncw;
    const int rps = p.stage_bytes / 2048;           // rows per stage
    const int spi = (p.rows + rps - 1) / rps;       // stages per item
    if (warp == 0) {
        if (lane >= ncw) return;
        const long long wg = (long long)blockIdx.x * ncw + lane;
        const long long nloc = wg < p.nitems ? (p.nitems - wg + NW - 1) / NW : 0, total = nloc * spi;
        const uint32_t ring = s0 + lane * S * p.stage_bytes, fb = full0 + lane * S * 8, eb = empty0 + lane * S * 8;
        long long item = wg;
        int r = 0, stage = 0, round = 0;
        for (long long n = 0; n < total; ++n) {
            if (round > 0) mbar_wait(eb + 8 * stage, (round + 1) & 1);
            const int i = (int)(item / p.njg), jg = (int)(item % p.njg);
            const uint32_t bar = fb + 8 * stage;
            int nr = p.rows - r * rps; if (nr > rps) nr = rps;
            mbar_expect(bar, p.mode == 2 ? nr * 2048 : p.stage_bytes);  // tensor boxes count their zero-filled rows too
            if (p.mode == 0) tma5(ring + stage * p.stage_bytes, &tmap, bar, 0, 0, jg * 4, r * rps, i);
            else if (p.mode == 1) tma5(ring + stage * p.stage_bytes, &tmap, bar, 0, 0, 0, r * rps, (int)item);
            else bulk1d(ring + stage * p.stage_bytes, p.buf + ((long long)item * p.rows + (long long)r * rps) * 512, nr * 2048, bar);
            if (++r == spi) { r = 0; item += NW; }
            if (++stage == S) { stage = 0; ++round; }
        }
        return;
    }
    const int cw = warp - 1;
    const long long wg = (long long)blockIdx.x * ncw + cw;
    const long long nloc = wg < p.nitems ? (p.nitems - wg + NW - 1) / NW : 0, total = nloc * spi;
    const uint32_t ring = s0 + cw * S * p.stage_bytes, fb = full0 + cw * S * 8, eb = empty0 + cw * S * 8;
    int stage = 0; uint32_t ph = 0; float acc = 0.f;
    for (long long n = 0; n < total; ++n) {
        mbar_wait(fb + 8 * stage, ph);
        float4 v;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(ring + stage * p.stage_bytes + lane * 16));
        acc += v.x + v.y + v.z + v.w;
        __syncwarp();
        if (lane == 0) mbar_arrive(eb + 8 * stage);
        if (++stage == S) { stage = 0; ph ^= 1; }
    }
    if (acc == 123.456f) p.sink[0] = acc;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
    int mode = argc > 1 ? atoi(argv[1]) : 0, stage_bytes = argc > 2 ? atoi(argv[2]) : 2048, S = argc > 3 ? atoi(argv[3]) : 14;
    int ncw = argc > 4 ? atoi(argv[4]) : 8;
    int promo = argc > 5 ? atoi(argv[5]) : 2;
    const int ni = 30, nj = 4096, rows = 127, Pp = 128, G = 4;
    const size_t elems = (size_t)ni * rows * nj * Pp;
    float* buf; CK(cudaMalloc(&buf, elems * 4)); CK(cudaMemset(buf, 0, elems * 4));
    float* sink; CK(cudaMalloc(&sink, 4));
    void* sym = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q));
    EncodeFn enc = (EncodeFn)sym;
    CUtensorMap tmap;
    cuuint64_t dims[5], strides[4]; cuuint32_t box[5], es[5] = {1, 1, 1, 1, 1};
    if (mode == 0) {  // [i][s][j][P]: dims (32, 4, nj, rows, ni)
        dims[0] = 32; dims[1] = 4; dims[2] = nj; dims[3] = rows; dims[4] = ni;
        strides[0] = 128; strides[1] = 512; strides[2] = (cuuint64_t)nj * 512; strides[3] = (cuuint64_t)rows * nj * 512;
        box[0] = 32; box[1] = 4; box[2] = G; box[3] = stage_bytes / 2048; box[4] = 1;
    } else {          // [item][s][G][P]: dims (32, 4, G, rows, items)
        dims[0] = 32; dims[1] = 4; dims[2] = G; dims[3] = rows; dims[4] = (cuuint64_t)ni * nj / G;
        strides[0] = 128; strides[1] = 512; strides[2] = 2048; strides[3] = (cuuint64_t)rows * 2048;
        box[0] = 32; box[1] = 4; box[2] = G; box[3] = stage_bytes / 2048; box[4] = 1;
    }
    CUtensorMapL2promotion pr = promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
    CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, pr, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    P p; p.buf = buf; p.mode = mode; p.stage_bytes = stage_bytes; p.S = S; p.rows = rows; p.njg = nj / G; p.ni = ni;
    p.nitems = (long long)ni * nj / G; p.sink = sink;
    size_t smem = (size_t)ncw * S * (stage_bytes + 16);
    CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int it = 0; it < 5; ++it) {
        cudaEventRecord(e0);
        stream_kernel<<<148, (ncw + 1) * 32, smem>>>(tmap, p);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    printf("mode %d stage %d B  S %d  consumer warps %d  promo %d  smem %zu: %.3f ms  %.1f GB/s\n", mode, stage_bytes, S, ncw, promo, smem, best, elems * 4 / best / 1e6);
    return 0;
}
