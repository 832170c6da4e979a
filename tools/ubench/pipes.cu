// Micro-benchmarks that decide the inner-loop shape of the signature recursion kernel:
// per-SM issue throughput of FFMA (3 distinct regs), FADD, packed FFMA2 (fma.rn.f32x2),
// SHFL.UP, LDS.128 and a mix, on B200.  Prints cycles per warp-instruction per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, long long* cyc, float seed) {
    __shared__ float4 sm[256 * 4];
    float a[16], b = seed * 1.0001f, c = seed * 0.5f;
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = seed + i + threadIdx.x;
    sm[threadIdx.x] = make_float4(seed, seed, seed, seed);
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
        if (MODE == 0) {  // FFMA 3 distinct sources, 16 independent chains
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], b, c);
        } else if (MODE == 1) {  // FADD
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = a[i] + b;
        } else if (MODE == 2) {  // FFMA2 packed: 8 packed ops = 16 fmas
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                unsigned long long A, B, C;
                asm volatile("mov.b64 %0, {%1, %2};" : "=l"(A) : "f"(a[i]), "f"(a[i + 1]));
                asm volatile("mov.b64 %0, {%1, %2};" : "=l"(B) : "f"(b), "f"(b));
                asm volatile("mov.b64 %0, {%1, %2};" : "=l"(C) : "f"(c), "f"(c));
                asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(A) : "l"(A), "l"(B), "l"(C));
                asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(a[i]), "=f"(a[i + 1]) : "l"(A));
            }
        } else if (MODE == 3) {  // SHFL.UP
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = __shfl_up_sync(0xffffffffu, a[i], 1);
        } else if (MODE == 4) {  // LDS.128 conflict-free, 4 per iter -> 16 floats
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float4 v = sm[(threadIdx.x + i * 32 + (it & 1)) & 1023];
                a[4 * i] += v.x; a[4 * i + 1] += v.y; a[4 * i + 2] += v.z; a[4 * i + 3] += v.w;
            }
        } else if (MODE == 5) {  // mix: 12 FFMA + 4 SHFL
#pragma unroll
            for (int i = 0; i < 12; ++i) a[i] = fmaf(a[i], b, c);
#pragma unroll
            for (int i = 12; i < 16; ++i) a[i] = __shfl_up_sync(0xffffffffu, a[i], 1);
        } else if (MODE == 6) {  // FFMA with shared operand in same slot (reuse-friendly): a[i] = d*a[i]+a[i^1]
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fmaf(b, a[i], a[(i + 1) & 15]);
        } else if (MODE == 7) {  // FMUL
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = a[i] * b;
        } else if (MODE == 9) {  // FADD2 packed: 8 packed ops = 16 adds
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                unsigned long long A, B;
                asm volatile("mov.b64 %0, {%1, %2};" : "=l"(A) : "f"(a[i]), "f"(a[i + 1]));
                asm volatile("mov.b64 %0, {%1, %2};" : "=l"(B) : "f"(b), "f"(c));
                asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(A) : "l"(A), "l"(B));
                asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(a[i]), "=f"(a[i + 1]) : "l"(A));
            }
        } else if (MODE == 10) {  // dependent-chain mix like the recursion: per j: t=A+O (FADD2), A+=p (FADD2), p=fma(d,t,p) (FFMA2)
            unsigned long long P, O, D;
            asm volatile("mov.b64 %0, {%1, %2};" : "=l"(P) : "f"(b), "f"(c));
            asm volatile("mov.b64 %0, {%1, %2};" : "=l"(O) : "f"(c), "f"(b));
            asm volatile("mov.b64 %0, {%1, %2};" : "=l"(D) : "f"(b), "f"(b));
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                unsigned long long A, T;
                asm volatile("mov.b64 %0, {%1, %2};" : "=l"(A) : "f"(a[i]), "f"(a[i + 1]));
                asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(T) : "l"(A), "l"(O));
                asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(A) : "l"(A), "l"(P));
                asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(P) : "l"(D), "l"(T), "l"(P));
                asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(a[i]), "=f"(a[i + 1]) : "l"(A));
            }
            float p0, p1;
            asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(p0), "=f"(p1) : "l"(P));
            b = p0 * 1e-30f + 1.0001f; c = p1 * 1e-30f + 0.5f;
        } else if (MODE == 11) {  // same chain, scalar: 16 x (FADD, FADD, FFMA)
            float p = b, o = c;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                float t = a[i] + o;
                a[i] += p;
                p = fmaf(b, t, p);
            }
            c = p * 1e-30f + 0.5f;
        } else if (MODE == 8) {  // FSEL-like select
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = (it & (1 << (i & 7))) ? a[i] : c;
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int opsPerIter, int warpsPerSM) {
    int nsm = 148;
    float* out; long long* cyc;
    int threads = 256, blocks = nsm * (warpsPerSM * 32 / threads);
    cudaMalloc(&out, sizeof(float) * threads * blocks);
    cudaMalloc(&cyc, sizeof(long long) * blocks);
    k<MODE><<<blocks, threads>>>(out, cyc, 1.0f);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(out, cyc, 1.0f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[8]; cudaMemcpy(h, cyc, sizeof(long long) * 8, cudaMemcpyDeviceToHost);
    double warpInstr = (double)ITERS * opsPerIter * warpsPerSM;  // per SM
    printf("%-28s warps/SM=%2d  cycles=%lld  cyc/warp-instr/SM=%.3f  (ms=%.3f, err=%d)\n", name, warpsPerSM, h[0],
           (double)h[0] / warpInstr, ms, (int)cudaGetLastError());
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int w : {8, 16, 32}) {
        run<0>("FFMA 3src", 16, w);
        run<6>("FFMA shared-slot", 16, w);
        run<1>("FADD", 16, w);
        run<7>("FMUL", 16, w);
        run<2>("FFMA2 (8 packed =16 fma)", 8, w);
        run<3>("SHFL.UP", 16, w);
        run<4>("LDS.128 (+4 FADD each)", 4, w);
        run<5>("mix 12 FFMA + 4 SHFL", 16, w);
        run<8>("FSEL", 16, w);
        run<9>("FADD2 (8 packed = 16 add)", 8, w);
        run<10>("chain packed 24 ops(48 flop-lanes)", 24, w);
        run<11>("chain scalar 48 ops", 48, w);
    }
    return 0;
}
