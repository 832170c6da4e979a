// pipes.cu -- issue-rate microbenchmark for the FP32 instructions the fused kernels are made of (sm_100a):
// scalar FFMA / FADD, packed fma/add/mul.f32x2, MUFU.EX2, SHFL, and mixes.  Prints warp-instructions per clock per SM.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/ubench/pipes tools/ubench/pipes.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 2048, ILP = 8;

template <int OP>
__global__ void k(float* out, long long* cyc, float b, float c) {
    float a[ILP];
    unsigned long long p[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { a[i] = threadIdx.x * 1e-3f + i; p[i] = ((unsigned long long)__float_as_uint(a[i]) << 32) | __float_as_uint(a[i] + 1.f); }
    unsigned long long b2 = ((unsigned long long)__float_as_uint(b) << 32) | __float_as_uint(b);
    unsigned long long c2 = ((unsigned long long)__float_as_uint(c) << 32) | __float_as_uint(c);
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (OP == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
            if (OP == 1) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
            if (OP == 2) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(b2), "l"(c2));
            if (OP == 3) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(b2));
            if (OP == 4) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(b2));
            if (OP == 5) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (OP == 6) a[i] = __shfl_up_sync(0xffffffffu, a[i], 1);
            if (OP == 7) {  // 1 FFMA2 + 1 FADD
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(b2), "l"(c2));
                asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
            }
            if (OP == 8) {  // 1 FFMA2 + 1 FADD2
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(b2), "l"(c2));
                unsigned long long q = p[(i + 1) % ILP];
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(q));
            }
            if (OP == 9) {  // FFMA with three distinct non-invariant registers
                asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(a[i]) : "f"(a[(i + 1) % ILP]), "f"(a[(i + 2) % ILP]));
            }
            if (OP == 10) {  // FFMA2, distinct registers
                asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p[i]) : "l"(p[(i + 1) % ILP]), "l"(p[(i + 2) % ILP]));
            }
            if (OP == 11) {  // the fused RBF kernel's mix per entry: 5 FFMA2 + 12 scalar + 1 MUFU  (scaled to 1 + 2.4 + 0.2)
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(b2), "l"(c2));
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
                asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(c));
                if ((i & 3) == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            }
            if (OP == 12) {  // 8 FFMA : 1 MUFU
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
                if (i == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            }
        }
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += a[i] + __uint_as_float((unsigned)(p[i] >> 32)) + __uint_as_float((unsigned)p[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, double instr_per_slot, float* out, long long* cyc) {
    for (int threads : {128, 256, 384, 512, 1024}) {
        k<OP><<<148, threads>>>(out, cyc, 1.0001f, 1e-7f);
        cudaDeviceSynchronize();
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        k<OP><<<148, threads>>>(out, cyc, 1.0001f, 1e-7f);
        cudaEventRecord(e1);
        cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        long long h[148];
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
        const double winstr = (double)(threads / 32) * ITERS * ILP * instr_per_slot;
        printf("%-28s warps/SM %2d  cycles %9.0f  warp-instr/clk/SM %.3f  (%.3f ms)\n", name, threads / 32, avg, winstr / avg, ms);
    }
}

int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    run<0>("FFMA (a=a*b+c)", 1, out, cyc);
    run<9>("FFMA 3 distinct regs", 1, out, cyc);
    run<1>("FADD", 1, out, cyc);
    run<2>("FFMA2 (f32x2)", 1, out, cyc);
    run<10>("FFMA2 3 distinct regs", 1, out, cyc);
    run<3>("FADD2 (f32x2)", 1, out, cyc);
    run<4>("FMUL2 (f32x2)", 1, out, cyc);
    run<5>("MUFU.EX2", 1, out, cyc);
    run<6>("SHFL.UP", 1, out, cyc);
    run<7>("FFMA2 + FADD", 2, out, cyc);
    run<8>("FFMA2 + FADD2", 2, out, cyc);
    run<11>("mix FFMA2+FFMA+FADD+.25EX2", 3.25, out, cyc);
    run<12>("8 FFMA + 1 EX2", 1.125, out, cyc);
    cudaError_t e = cudaGetLastError();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
