// gram.cu -- static-kernel Gram stage of the covariance path (kernels.py:225-230, :198, :275-277, :328-333).
//
//  * prep kernels: lengthscale scaling (kernels.py:358), time increments, squared norms -- O(N L d), negligible.
//  * delta_producer: writes a row block of the INCREMENT Gram Delta[i, s, j, t] (what signature_algs.py:26 derives
//    from the raw Gram) straight into the chunk buffer the recursion kernel streams with TMA.  Layout
//    [ib][rows][nj][P] fp32, P = 16 * lanes-per-pair (multiple of 32, columns >= ncols stay zero).  The kernel is
//    store-bound by design: each thread keeps its 4 points y_{j,t..t+3} in registers for the whole row block and emits
//    one 16-byte store per (i, s); a warp writes 512 contiguous bytes.
//  * gram_kernel: plain (rows1 x rows2) Gram for the operator-level API (compute_base_kern_symm, K_tens, ...).
#include <string.h>

#include "internal.cuh"

namespace gpsig {

// ---- static kernels as functions of (dot, |x|^2, |y|^2, |x-y|^2) ---------------------------------------------------
template <int KIND> struct KernTraits;
template <> struct KernTraits<GPSIG_KERN_LINEAR>   { static constexpr bool dot = true,  sq = false, norms = false; };
template <> struct KernTraits<GPSIG_KERN_RBF>      { static constexpr bool dot = false, sq = true,  norms = false; };
template <> struct KernTraits<GPSIG_KERN_COSINE>   { static constexpr bool dot = true,  sq = false, norms = true;  };
template <> struct KernTraits<GPSIG_KERN_POLY>     { static constexpr bool dot = true,  sq = false, norms = false; };
template <> struct KernTraits<GPSIG_KERN_MIX>      { static constexpr bool dot = true,  sq = true,  norms = false; };
template <> struct KernTraits<GPSIG_KERN_MATERN12> { static constexpr bool dot = false, sq = true,  norms = false; };
template <> struct KernTraits<GPSIG_KERN_MATERN32> { static constexpr bool dot = false, sq = true,  norms = false; };
template <> struct KernTraits<GPSIG_KERN_MATERN52> { static constexpr bool dot = false, sq = true,  norms = false; };
template <> struct KernTraits<GPSIG_KERN_SPECTRAL> { static constexpr bool dot = false, sq = false, norms = false; };

template <int KIND>
__device__ __forceinline__ float kern_eval(float dot, float sq, float xx, float yy, const KernParams& kp) {
    if (KIND == GPSIG_KERN_LINEAR) return dot;                                   // kernels.py:799-806
    if (KIND == GPSIG_KERN_RBF) return __expf(-0.5f * sq);                       // kernels.py:862-864
    if (KIND == GPSIG_KERN_COSINE) return dot / (sqrtf(xx) * sqrtf(yy));         // kernels.py:820-828
    if (KIND == GPSIG_KERN_POLY) return powf(dot + kp.a, kp.b);                  // kernels.py:844-848
    if (KIND == GPSIG_KERN_MIX) return kp.a * __expf(-0.5f * sq) + (1.f - kp.a) * dot;  // kernels.py:881-892
    const float r = sqrtf(fmaxf(sq, 1e-40f));                                    // kernels.py:779-781
    if (KIND == GPSIG_KERN_MATERN12) return __expf(-r);                          // kernels.py:955-958
    if (KIND == GPSIG_KERN_MATERN32) {                                           // kernels.py:974-977
        const float t = 1.7320508075688772f * r;
        return (1.f + t) * __expf(-t);
    }
    const float t = 2.23606797749979f * r;                                       // kernels.py:991-993
    return (1.f + t + (5.f / 3.f) * r * r) * __expf(-t);
}

// ---- prep: out[n, r, 0..DP) = scaled points (mode 0) or scaled time increments (mode 1), zero padded to DP --------
//      mode 3 (RBF fast paths): points times sqrt(log2(e) / 2) / lengthscale, so that log2 k(x, y) = -|x' - y'|^2
//      (kernels.py:765-776, :862-864): the readers evaluate the squared distance from DIFFERENCES of nearby points
//      (warpfused.cu, tens.cu) or directly (the chunk producer below) -- never through |x|^2 + |y|^2 - 2<x, y>.
__global__ void prep_points_kernel(const float* __restrict__ X, long long n, int L, int d, const float* __restrict__ inv_ls,
                                   int mode, int DP, float* __restrict__ out, float* __restrict__ norms) {
    const int increments = mode == 1;
    const int Lo = increments ? L - 1 : L;
    const long long total = n * (long long)Lo;
    const float rs = mode == 3 ? 0.8493218002880191f : 1.f;  // sqrt(log2(e) / 2)
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long seq = idx / Lo;
        const int r = (int)(idx - seq * Lo);
        const float* x0 = X + (seq * L + r) * d;
        float nn = 0.f;
        for (int c = 0; c < DP; ++c) {
            float v = 0.f;
            if (c < d) {
                const float s = (inv_ls ? inv_ls[c] : 1.f) * rs;
                v = increments ? (x0[d + c] - x0[c]) * s : x0[c] * s;
            }
            out[idx * DP + c] = v;
            nn = fmaf(v, v, nn);
        }
        if (norms) norms[idx] = nn;
    }
}

// ---- the chunk producer (ProdParams: internal.cuh) ----------------------------------------------------------------
// One thread owns 4 consecutive columns t0..t0+3 of one pair column j and walks down the rows of the row block.
// DIFF2D: out = (f[s+1][t+1] - f[s+1][t]) - (f[s][t+1] - f[s][t]); the value at t0+4 comes from the next lane by
// shuffle (a thread whose right-hand neighbour is in another warp or another pair evaluates it itself, and only when that column is a real increment).
// Skip decisions are warp-uniform (a warp never straddles two consumer pair groups), so full-mask shuffles are safe.
template <int KIND, int DP, bool DIFF2D>
__global__ void __launch_bounds__(256) delta_producer_kernel(const ProdParams p) {
    using KT = KernTraits<KIND>;
    extern __shared__ float sA[];  // [rowsA][DP] (+ [rowsA] norms) of the current row sequence
    const int tpp = p.P >> 2;      // threads per pair row
    const int ppb = blockDim.x / tpp;
    const int jl = blockIdx.x * ppb + threadIdx.x / tpp;  // local pair column
    const int t0 = (threadIdx.x % tpp) * 4;
    const bool active = jl < p.nj;
    const int j = p.j0 + (active ? jl : p.nj - 1);
    const bool direct5 = DIFF2D && ((threadIdx.x % tpp == tpp - 1) || (threadIdx.x & 31) == 31) && (t0 + 4 < p.rowsB) &&
                         (t0 + 3 < p.ncols);  // the right-hand neighbour column is not held by the next lane of this warp
    constexpr int NPT = DIFF2D ? 5 : 4;
    float y[NPT][DP];
    float yn[NPT];
#pragma unroll
    for (int u = 0; u < NPT; ++u) {
        const int t = t0 + u;
        const bool ok = t < p.rowsB && (u < 4 || direct5);
#pragma unroll
        for (int c = 0; c < DP; ++c) y[u][c] = ok ? p.B[((long long)j * p.rowsB + t) * DP + c] : 0.f;
        yn[u] = (ok && p.Bn) ? p.Bn[(long long)j * p.rowsB + t] : 0.f;
    }
    float* sAn = sA + p.rowsA * DP;
    const int group_last_j = (j / p.G) * p.G + p.G - 1;
    const int ii_begin = p.diag ? jl : blockIdx.y, ii_end = p.diag ? jl + 1 : p.ni, ii_step = p.diag ? 1 : gridDim.y;
    for (int ii = ii_begin; ii < ii_end; ii += ii_step) {
        const int i = p.i0 + ii;
        __syncthreads();
        for (int e = threadIdx.x; e < p.rowsA * DP; e += blockDim.x) sA[e] = p.A[(long long)i * p.rowsA * DP + e];
        if (KT::norms)
            for (int e = threadIdx.x; e < p.rowsA; e += blockDim.x) sAn[e] = p.An ? p.An[(long long)i * p.rowsA + e] : 0.f;
        __syncthreads();
        if (!p.diag && p.upper_only && group_last_j < i) continue;
        float* orow = p.diag ? p.out + (long long)jl * p.P + t0
                             : p.out + (((long long)ii * p.out_rows) * p.nj + jl) * p.P + t0;
        long long row_stride = (long long)p.nj * p.P;
        if (p.stream) {  // consumer-ready layout: stream u % NW, position u / NW, skewed row s + strip, swizzled chunk
            const int jgl = jl / p.G, q = jl - jgl * p.G;
            const long long u = p.diag ? (long long)jgl
                                       : items_before(ii, p.njg, p.G, p.upper_only, p.i0, p.j0) + jgl -
                                             first_group(ii, p.G, p.upper_only, p.i0, p.j0);
            const long long w = u % p.NW, n = u / p.NW;
            const uint32_t sw = swizzle_in_row((uint32_t)(q * p.P + t0) * 4u);
            orow = p.out + (w * p.SR + n * p.out_rows + (t0 >> 4)) * kSkewRowFloats + (sw >> 2);
            row_stride = kSkewRowFloats;
        }
        float fprev[NPT];
#pragma unroll
        for (int u = 0; u < NPT; ++u) fprev[u] = 0.f;
        for (int s = 0; s < p.rowsA; ++s) {
            float f[NPT];
            const float xn = (KT::norms) ? sAn[s] : 0.f;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (KIND == GPSIG_KERN_SPECTRAL) {
                    f[u] = spectral_eval(sA + s * DP, y[u], DP, p.kp);
                    continue;
                }
                float dot = 0.f, sq = 0.f;
#pragma unroll
                for (int c = 0; c < DP; ++c) {
                    const float xv = sA[s * DP + c];
                    if (KT::dot) dot = fmaf(xv, y[u][c], dot);
                    if (KT::sq) { const float df = xv - y[u][c]; sq = fmaf(df, df, sq); }
                }
                f[u] = kern_eval<KIND>(dot, sq, xn, yn[u], p.kp);
            }
            if (DIFF2D) {
                f[4] = __shfl_down_sync(0xffffffffu, f[0], 1);
                if (direct5) {
                    float dot = 0.f, sq = 0.f;
#pragma unroll
                    for (int c = 0; c < DP; ++c) {
                        const float xv = sA[s * DP + c];
                        if (KT::dot) dot = fmaf(xv, y[NPT - 1][c], dot);
                        if (KT::sq) { const float df = xv - y[NPT - 1][c]; sq = fmaf(df, df, sq); }
                    }
                    f[4] = KIND == GPSIG_KERN_SPECTRAL ? spectral_eval(sA + s * DP, y[NPT - 1], DP, p.kp)
                                                       : kern_eval<KIND>(dot, sq, xn, yn[NPT - 1], p.kp);
                }
                if (s > 0 && active) {
                    float4 o;
                    o.x = (t0 + 0 < p.ncols) ? (f[1] - f[0]) - (fprev[1] - fprev[0]) : 0.f;
                    o.y = (t0 + 1 < p.ncols) ? (f[2] - f[1]) - (fprev[2] - fprev[1]) : 0.f;
                    o.z = (t0 + 2 < p.ncols) ? (f[3] - f[2]) - (fprev[3] - fprev[2]) : 0.f;
                    o.w = (t0 + 3 < p.ncols) ? (f[4] - f[3]) - (fprev[4] - fprev[3]) : 0.f;
                    *reinterpret_cast<float4*>(orow + (long long)(s - 1) * row_stride) = o;
                }
#pragma unroll
                for (int u = 0; u < NPT; ++u) fprev[u] = f[u];
            } else if (active) {
                float4 o;
                o.x = (t0 + 0 < p.ncols) ? f[0] : 0.f;
                o.y = (t0 + 1 < p.ncols) ? f[1] : 0.f;
                o.z = (t0 + 2 < p.ncols) ? f[2] : 0.f;
                o.w = (t0 + 3 < p.ncols) ? f[3] : 0.f;
                *reinterpret_cast<float4*>(orow + (long long)s * row_stride) = o;
            }
        }
    }
}

// ---- fast producer for the two headline static kernels -------------------------------------------------------------
// LINEAR (RBF = false): A / B are time increments, out = <dx_s, dy_t>.          DPA = padded d.
// RBF    (RBF = true) : A / B are scaled points (prep mode 3), f = 2^(-|x' - y'|^2) from the differences directly
//                       (robust wherever the data sits), out = 2-D increment of f.
// Same thread mapping, skip logic and output addressing as delta_producer_kernel; the arithmetic runs on packed
// add / fma.rn.f32x2 (two features per instruction, sm_100), the exponential is one ex2.approx.
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <bool RBF, int DPA>
__global__ void __launch_bounds__(256) delta_producer_fast_kernel(const ProdParams p) {
    extern __shared__ __align__(16) float sA[];  // [rowsA][DPA] of the current row sequence
    constexpr int H = DPA / 2;
    constexpr int NPT = RBF ? 5 : 4;
    const int tpp = p.P >> 2;      // threads per pair row
    const int ppb = blockDim.x / tpp;
    const int jl = blockIdx.x * ppb + threadIdx.x / tpp;  // local pair column
    const int t0 = (threadIdx.x % tpp) * 4;
    const bool active = jl < p.nj;
    const int j = p.j0 + (active ? jl : p.nj - 1);
    // RBF: the column to the right (t0 + 4) comes from the next lane by shuffle unless that lane is in another warp or
    // another pair (`bnd`); then the thread evaluates it itself (direct5) or, past the last point, reuses its own column.
    // Columns past the sequence need no masking: RBF clamps them to the last point (equal values difference to exactly
    // zero), LINEAR gives them zero increments.
    const bool bnd = (threadIdx.x % tpp == tpp - 1) || (threadIdx.x & 31) == 31;
    const bool direct5 = RBF && bnd && (t0 + 4 < p.rowsB);
    const bool any5 = RBF && __any_sync(0xffffffffu, direct5);  // warp-uniform (never true when P == rowsB)
    float2 y[NPT][H];
#pragma unroll
    for (int u = 0; u < NPT; ++u) {
        const int t = t0 + u;
        const bool ok = RBF || t < p.rowsB;
        const int tc = t < p.rowsB ? t : p.rowsB - 1;
        const float2* src = reinterpret_cast<const float2*>(p.B + ((long long)j * p.rowsB + tc) * DPA);
#pragma unroll
        for (int h = 0; h < H; ++h) y[u][h] = ok ? src[h] : make_float2(0.f, 0.f);
        if (RBF) {  // the evaluation adds x to the NEGATED point
#pragma unroll
            for (int h = 0; h < H; ++h) y[u][h] = make_float2(-y[u][h].x, -y[u][h].y);
        }
    }
    const int group_last_j = (j / p.G) * p.G + p.G - 1;
    const int ii_begin = p.diag ? jl : blockIdx.y, ii_end = p.diag ? jl + 1 : p.ni, ii_step = p.diag ? 1 : gridDim.y;
    for (int ii = ii_begin; ii < ii_end; ii += ii_step) {
        const int i = p.i0 + ii;
        __syncthreads();
        {
            const float4* srcA = reinterpret_cast<const float4*>(p.A + (long long)i * p.rowsA * DPA);
            float4* dstA = reinterpret_cast<float4*>(sA);
            for (int e = threadIdx.x; e < p.rowsA * (DPA / 4); e += blockDim.x) dstA[e] = srcA[e];
        }
        __syncthreads();
        if (!p.diag && p.upper_only && group_last_j < i) continue;
        float* orow = p.diag ? p.out + (long long)jl * p.P + t0
                             : p.out + (((long long)ii * p.out_rows) * p.nj + jl) * p.P + t0;
        long long row_stride = (long long)p.nj * p.P;
        if (p.stream) {
            const int jgl = jl / p.G, q = jl - jgl * p.G;
            const long long u = p.diag ? (long long)jgl
                                       : items_before(ii, p.njg, p.G, p.upper_only, p.i0, p.j0) + jgl -
                                             first_group(ii, p.G, p.upper_only, p.i0, p.j0);
            const long long w = u % p.NW, n = u / p.NW;
            const uint32_t sw = swizzle_in_row((uint32_t)(q * p.P + t0) * 4u);
            orow = p.out + (w * p.SR + n * p.out_rows + (t0 >> 4)) * kSkewRowFloats + (sw >> 2);
            row_stride = kSkewRowFloats;
        }
        // f[0..3] (and the halo f[4] for RBF) of row s
        auto eval_row = [&](int s, float (&f)[NPT]) {
            float2 x[H];
            const float2* xs = reinterpret_cast<const float2*>(sA + s * DPA);
#pragma unroll
            for (int h = 0; h < H; ++h) x[h] = xs[h];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                float2 acc = make_float2(0.f, 0.f);
#pragma unroll
                for (int h = 0; h < H; ++h) {
                    if (RBF) {
                        const float2 df = __fadd2_rn(x[h], y[u][h]);
                        acc = __ffma2_rn(df, df, acc);
                    } else {
                        acc = __ffma2_rn(x[h], y[u][h], acc);
                    }
                }
                const float v = acc.x + acc.y;
                f[u] = RBF ? ex2_approx(-v) : v;
            }
            if (RBF) {
                const float nb = __shfl_down_sync(0xffffffffu, f[0], 1);
                f[NPT - 1] = bnd ? f[3] : nb;
                if (any5) {  // uniform branch
                    float2 acc = make_float2(0.f, 0.f);
#pragma unroll
                    for (int h = 0; h < H; ++h) {
                        const float2 df = __fadd2_rn(x[h], y[NPT - 1][h]);
                        acc = __ffma2_rn(df, df, acc);
                    }
                    const float v = ex2_approx(-(acc.x + acc.y));
                    if (direct5) f[NPT - 1] = v;
                }
            }
        };
        if (RBF) {
            float fprev[NPT], f[NPT];
            eval_row(0, fprev);
            float* o = orow;
            for (int s = 1; s < p.rowsA; ++s, o += row_stride) {
                eval_row(s, f);
                float4 v;
                v.x = (f[1] - f[0]) - (fprev[1] - fprev[0]);
                v.y = (f[2] - f[1]) - (fprev[2] - fprev[1]);
                v.z = (f[3] - f[2]) - (fprev[3] - fprev[2]);
                v.w = (f[NPT - 1] - f[3]) - (fprev[NPT - 1] - fprev[3]);
                if (active) *reinterpret_cast<float4*>(o) = v;
#pragma unroll
                for (int u = 0; u < NPT; ++u) fprev[u] = f[u];
            }
        } else {
            float f[NPT];
            float* o = orow;
            for (int s = 0; s < p.rowsA; ++s, o += row_stride) {
                eval_row(s, f);
                if (active) *reinterpret_cast<float4*>(o) = make_float4(f[0], f[1], f[2], f[3]);
            }
        }
    }
}

template <bool RBF, int DPA>
static int launch_producer_fast_dpa(const ProdParams& p, cudaStream_t st) {
    const int tpp = p.P >> 2;
    const int threads = p.diag ? tpp : 256;
    const int ppb = threads / tpp;
    const int gx = (p.nj + ppb - 1) / ppb;
    int gy = p.diag ? 1 : p.ni;
    const int target = num_sms() * 8;
    if ((long long)gx * gy > target) gy = (target + gx - 1) / gx;
    if (gy < 1) gy = 1;
    if (gy > p.ni) gy = p.ni;
    const size_t smem = (size_t)p.rowsA * DPA * sizeof(float);
    auto k = delta_producer_fast_kernel<RBF, DPA>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<<<dim3(gx, gy), threads, smem, st>>>(p);
    return check_launch();
}

// DPA: floats per prepared point.  Returns GPSIG_E_UNSUPPORTED when the shape has no fast instantiation.
int launch_delta_producer_fast(bool rbf, const ProdParams& p, int DPA, cudaStream_t st) {
    ProfScope prof(GPSIG_PROF_PRODUCER, st, p.diag ? (double)p.nj : (double)p.ni * p.nj);
    if (rbf) {
        switch (DPA) {
            case 4: return launch_producer_fast_dpa<true, 4>(p, st);
            case 8: return launch_producer_fast_dpa<true, 8>(p, st);
            case 12: return launch_producer_fast_dpa<true, 12>(p, st);
            case 16: return launch_producer_fast_dpa<true, 16>(p, st);
        }
    } else {
        switch (DPA) {
            case 4: return launch_producer_fast_dpa<false, 4>(p, st);
            case 8: return launch_producer_fast_dpa<false, 8>(p, st);
            case 12: return launch_producer_fast_dpa<false, 12>(p, st);
            case 16: return launch_producer_fast_dpa<false, 16>(p, st);
        }
    }
    return fail(GPSIG_E_UNSUPPORTED, "no fast producer for padded dimension %d", DPA);
}

template <int KIND, int DP>
static int launch_producer_dp(const ProdParams& p, bool diff2d, cudaStream_t st) {
    const int tpp = p.P >> 2;          // P <= 512 -> at most 128 threads per pair row
    const int threads = p.diag ? tpp : 256;
    const int ppb = threads / tpp;
    const int gx = (p.nj + ppb - 1) / ppb;
    int gy = p.diag ? 1 : p.ni;
    const int target = num_sms() * 8;
    if ((long long)gx * gy > target) gy = (target + gx - 1) / gx;
    if (gy < 1) gy = 1;
    if (gy > p.ni) gy = p.ni;
    const size_t smem = (size_t)p.rowsA * (DP + 1) * sizeof(float);
    dim3 grid(gx, gy);
    if (diff2d) {
        auto k = delta_producer_kernel<KIND, DP, true>;
        if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k<<<grid, threads, smem, st>>>(p);
    } else {
        auto k = delta_producer_kernel<KIND, DP, false>;
        if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k<<<grid, threads, smem, st>>>(p);
    }
    return check_launch();
}

template <int KIND>
static int launch_producer_kind(const ProdParams& p, int DP, bool diff2d, cudaStream_t st) {
    switch (DP) {
        case 4: return launch_producer_dp<KIND, 4>(p, diff2d, st);
        case 8: return launch_producer_dp<KIND, 8>(p, diff2d, st);
        case 12: return launch_producer_dp<KIND, 12>(p, diff2d, st);
        case 16: return launch_producer_dp<KIND, 16>(p, diff2d, st);
    }
    return fail(GPSIG_E_UNSUPPORTED, "state-space dimension > 16 not supported by the chunk producer yet (DP=%d)", DP);
}

// kind LINEAR with difference: the caller passes time INCREMENTS as A/B and diff2d = false (Delta = <dx_s, dy_t>,
// kernels.py:226 + signature_algs.py:26 by bilinearity); every other kind passes scaled points and diff2d = difference.
int launch_delta_producer(int kind, const ProdParams& p, int DP, bool diff2d, cudaStream_t st) {
    ProfScope prof(GPSIG_PROF_PRODUCER, st, p.diag ? (double)p.nj : (double)p.ni * p.nj);
    switch (kind) {
        case GPSIG_KERN_LINEAR: return launch_producer_kind<GPSIG_KERN_LINEAR>(p, DP, diff2d, st);
        case GPSIG_KERN_RBF: return launch_producer_kind<GPSIG_KERN_RBF>(p, DP, diff2d, st);
        case GPSIG_KERN_COSINE: return launch_producer_kind<GPSIG_KERN_COSINE>(p, DP, diff2d, st);
        case GPSIG_KERN_POLY: return launch_producer_kind<GPSIG_KERN_POLY>(p, DP, diff2d, st);
        case GPSIG_KERN_MIX: return launch_producer_kind<GPSIG_KERN_MIX>(p, DP, diff2d, st);
        case GPSIG_KERN_MATERN12: return launch_producer_kind<GPSIG_KERN_MATERN12>(p, DP, diff2d, st);
        case GPSIG_KERN_MATERN32: return launch_producer_kind<GPSIG_KERN_MATERN32>(p, DP, diff2d, st);
        case GPSIG_KERN_MATERN52: return launch_producer_kind<GPSIG_KERN_MATERN52>(p, DP, diff2d, st);
        case GPSIG_KERN_SPECTRAL: return launch_producer_kind<GPSIG_KERN_SPECTRAL>(p, DP, diff2d, st);
    }
    return fail(GPSIG_E_BADARG, "unknown static kernel kind %d", kind);
}

int launch_prep_points(const float* X, long long n, int L, int d, const float* inv_ls, int mode, int DP, float* out,
                       float* norms, cudaStream_t st) {
    const long long total = n * (long long)(mode == 1 ? L - 1 : L);
    if (total <= 0) return GPSIG_OK;
    ProfScope prof(GPSIG_PROF_PREP, st, (double)n);
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)num_sms() * 16;
    prep_points_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, st>>>(X, n, L, d, inv_ls, mode, DP, out, norms);
    return check_launch();
}

// ---- plain Gram (operator-level API) -------------------------------------------------------------------------------
template <int KIND>
__global__ void gram_kernel(const float* __restrict__ X, long long rows1, const float* __restrict__ X2, long long rows2,
                            int d, KernParams kp, float* __restrict__ out, long long ld) {
    // rows on grid.x (up to 2^31 - 1 blocks of 8 rows: N L of the low-rank mode's Nystrom map runs into the millions), column
    // tiles on grid.y with a stride loop (grid.y is capped at 65535)
    const long long r = (long long)blockIdx.x * blockDim.y + threadIdx.y;
    if (r >= rows1) return;
    const float* x = X + r * d;
    for (long long c = (long long)blockIdx.y * blockDim.x + threadIdx.x; c < rows2; c += (long long)gridDim.y * blockDim.x) {
        const float* y = X2 + c * d;
        float dot = 0.f, sq = 0.f, xx = 0.f, yy = 0.f;
        for (int k = 0; k < d; ++k) {
            const float a = x[k], b = y[k];
            dot = fmaf(a, b, dot);
            const float df = a - b;
            sq = fmaf(df, df, sq);
            xx = fmaf(a, a, xx);
            yy = fmaf(b, b, yy);
        }
        out[r * ld + c] = KIND == GPSIG_KERN_SPECTRAL ? spectral_eval(x, y, d, kp) : kern_eval<KIND>(dot, sq, xx, yy, kp);
    }
}

template <int KIND>
static int launch_gram_kind(const float* X, long long r1, const float* X2, long long r2, int d, KernParams kp, float* out,
                            long long ld, cudaStream_t st) {
    const long long ctiles = (r2 + 31) / 32;
    dim3 block(32, 8), grid((unsigned)((r1 + 7) / 8), (unsigned)(ctiles < 65535 ? ctiles : 65535));
    gram_kernel<KIND><<<grid, block, 0, st>>>(X, r1, X2, r2, d, kp, out, ld);
    return check_launch();
}

KernParams make_kern_params(int kind, const float* params) {
    KernParams kp;
    memset(&kp, 0, sizeof(kp));
    if (kind == GPSIG_KERN_SPECTRAL && params) {  // {family, Q, d, alpha[Q], omega[Q*d], gamma[Q*d]}; sizes clamped
        kp.fam = (int)params[0];
        int Q = (int)params[1], d = (int)params[2];
        const int Qc = Q < kSpecMaxQ ? Q : kSpecMaxQ, dc = d < kSpecMaxD ? d : kSpecMaxD;
        kp.Q = Qc;
        const float* al = params + 3;
        const float* om = al + Q;
        const float* ga = om + (size_t)Q * d;
        for (int q = 0; q < Qc; ++q) {
            kp.alpha[q] = al[q];
            for (int c = 0; c < dc; ++c) {
                kp.omega[q * kSpecMaxD + c] = om[q * d + c];
                kp.gamma[q * kSpecMaxD + c] = ga[q * d + c];
            }
        }
    }
    if (kind == GPSIG_KERN_POLY) { kp.a = params ? params[0] : 1.f; kp.b = params ? params[1] : 3.f; }
    if (kind == GPSIG_KERN_MIX) kp.a = params ? params[0] : 0.5f;
    return kp;
}

__global__ void scale_features_kernel(const float* __restrict__ X, long long total, int d, const float* __restrict__ inv_ls,
                                      int nf, float* __restrict__ out) {
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(idx % d);
        out[idx] = inv_ls ? X[idx] * inv_ls[c % nf] : X[idx];
    }
}

// lags.py:7-63.  time grid l / (L - 1); query max(time - lag, 0); left knot = last grid point <= query (+ jitter, lags.py:23)
__global__ void add_lags_kernel(const float* __restrict__ X, long long n, int L, int d, const float* __restrict__ lags,
                                int num_lags, float* __restrict__ out) {
    const int P1 = num_lags + 1;
    const long long total = n * (long long)L * P1 * d;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(idx % d);
        const int pl = (int)((idx / d) % P1);
        const int l = (int)((idx / ((long long)d * P1)) % L);
        const long long seq = idx / ((long long)d * P1 * L);
        const float* xs = X + seq * (long long)L * d;
        if (pl == 0) { out[idx] = xs[(long long)l * d + c]; continue; }
        const double Lm1 = (double)(L - 1);
        double tq = (double)l / Lm1 - (double)lags[pl - 1];
        if (tq < 0.0) tq = 0.0;
        int left = (int)floor((tq + 1e-6) * Lm1 + 1e-9);  // settings.jitter = 1e-6
        if (left > L - 2) left = L - 2;
        if (left < 0) left = 0;
        const double tl = (double)left / Lm1, tr = (double)(left + 1) / Lm1;
        const double xl = xs[(long long)left * d + c], xr = xs[(long long)(left + 1) * d + c];
        out[idx] = (float)(xl + (tq - tl) * (xr - xl) / (tr - tl));
    }
}

}  // namespace gpsig

using namespace gpsig;

extern "C" int gpsig_add_lags(const float* X, long n, int L, int d, const float* lags, int num_lags, float* out, void* stream) {
    if (!X || !out || !lags || n < 0 || L < 2 || d < 1 || num_lags < 1) return fail(GPSIG_E_BADARG, "add_lags: bad arguments");
    const long long total = (long long)n * L * (num_lags + 1) * d;
    if (total == 0) return GPSIG_OK;
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)num_sms() * 16;
    ProfScope prof(GPSIG_PROF_PREP, (cudaStream_t)stream, (double)n);
    add_lags_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>(X, n, L, d, lags, num_lags, out);
    return check_launch();
}

extern "C" int gpsig_scale_features(const float* X, long rows, int d, const float* inv_lengthscales, int num_features,
                                    float* out, void* stream) {
    if (!X || !out || rows < 0 || d < 1 || num_features < 1) return fail(GPSIG_E_BADARG, "scale_features: bad arguments");
    const long long total = (long long)rows * d;
    if (total == 0) return GPSIG_OK;
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)num_sms() * 16;
    scale_features_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>(X, total, d, inv_lengthscales,
                                                                                                 num_features, out);
    return check_launch();
}

extern "C" int gpsig_gram(int kind, const float* X, long rows1, const float* X2, long rows2, int d, const float* params,
                          float* out, long ld, void* stream) {
    if (!X || !out || rows1 < 1 || d < 1) return fail(GPSIG_E_BADARG, "gram: bad arguments");
    if (!X2) { X2 = X; rows2 = rows1; }
    if (rows2 < 1 || ld < rows2) return fail(GPSIG_E_BADARG, "gram: bad rows2/ld");
    cudaStream_t st = (cudaStream_t)stream;
    KernParams kp = make_kern_params(kind, params);
    switch (kind) {
        case GPSIG_KERN_LINEAR: return launch_gram_kind<GPSIG_KERN_LINEAR>(X, rows1, X2, rows2, d, kp, out, ld, st);
        case GPSIG_KERN_RBF: return launch_gram_kind<GPSIG_KERN_RBF>(X, rows1, X2, rows2, d, kp, out, ld, st);
        case GPSIG_KERN_COSINE: return launch_gram_kind<GPSIG_KERN_COSINE>(X, rows1, X2, rows2, d, kp, out, ld, st);
        case GPSIG_KERN_POLY: return launch_gram_kind<GPSIG_KERN_POLY>(X, rows1, X2, rows2, d, kp, out, ld, st);
        case GPSIG_KERN_MIX: return launch_gram_kind<GPSIG_KERN_MIX>(X, rows1, X2, rows2, d, kp, out, ld, st);
        case GPSIG_KERN_MATERN12: return launch_gram_kind<GPSIG_KERN_MATERN12>(X, rows1, X2, rows2, d, kp, out, ld, st);
        case GPSIG_KERN_MATERN32: return launch_gram_kind<GPSIG_KERN_MATERN32>(X, rows1, X2, rows2, d, kp, out, ld, st);
        case GPSIG_KERN_MATERN52: return launch_gram_kind<GPSIG_KERN_MATERN52>(X, rows1, X2, rows2, d, kp, out, ld, st);
        case GPSIG_KERN_SPECTRAL:
            if (d > kSpecMaxD) return fail(GPSIG_E_UNSUPPORTED, "spectral kernel supports at most %d features", kSpecMaxD);
            return launch_gram_kind<GPSIG_KERN_SPECTRAL>(X, rows1, X2, rows2, d, kp, out, ld, st);
    }
    return fail(GPSIG_E_BADARG, "unknown static kernel kind %d", kind);
}
