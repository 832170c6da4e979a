// tens.cu -- inducing-tensor side of the covariance path.
//   tensor_kern            signature_algs.py:76-99   + the 2x2 increment of kernels.py:275-277
//   tens_vs_seq            signature_algs.py:101-127 (first order) / :129-160 (higher order) + kernels.py:329-330
//   tens_seq_kern (fused)  kernels.py:313-340 (_K_tens_vs_seq) without the (T, nz, n, L) Gram ever touching HBM
#include "internal.cuh"

namespace gpsig {

constexpr int kMaxLevelsTens = 10;                                       // T = M(M+1)/2 <= 55
constexpr int kMaxT = kMaxLevelsTens * (kMaxLevelsTens + 1) / 2;

// ---- a9 ------------------------------------------------------------------------------------------------------------
__global__ void tensor_kern_kernel(const float* __restrict__ M, int nlev, long long nz, long long nz2, int increments,
                                   float* __restrict__ out) {
    const long long per = nz * nz2;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < per; idx += (long long)gridDim.x * blockDim.x) {
        const long long z = idx / nz2, z2 = idx - z * nz2;
        out[idx] = 1.f;
        int k = 0;
        for (int m = 1; m <= nlev; ++m) {
            float r = 1.f;
            for (int c = 0; c < m; ++c, ++k) {
                float v;
                if (increments) {  // raw Gram (T, nz, 2, nz2, 2): M11 + M00 - M10 - M01 (kernels.py:277)
                    const float* b = M + (((long long)k * nz + z) * 2) * nz2 * 2;
                    const float m00 = b[z2 * 2], m01 = b[z2 * 2 + 1], m10 = b[nz2 * 2 + z2 * 2], m11 = b[nz2 * 2 + z2 * 2 + 1];
                    v = m11 + m00 - m10 - m01;
                } else {
                    v = M[(long long)k * per + idx];
                }
                r *= v;
            }
            out[(long long)m * per + idx] = r;
        }
    }
}

// ---- a10/a11 on a materialised Gram -------------------------------------------------------------------------------
// one thread per (z, n); time is swept serially with the T running prefixes c[k] in registers/local memory.
//   first order : r_p[t] = H_{k(m,p)}[t] * c[k(m,p-1)] (exclusive), c[k(m,p)] += r_p[t];  K_m = c[k(m,m-1)] at the end
//   higher order: a vector cur[l] per level carries the same-time repeats (signature_algs.py:151-156)
struct TvsParams {
    const float* M;
    int nlev, order, difference, increments;
    long long nz, n;
    int L;
    float* out;
};

__device__ __forceinline__ void tvs_step(const float* h, int nlev, int order, float* c) {
    // h[k]: increment of component k at this time step; c[k]: running exclusive prefixes (updated in place)
    int k0 = 0;
    for (int m = 1; m <= nlev; ++m) {
        if (order == 1) {
            float prev_excl = 0.f;
            for (int pi = 0; pi < m; ++pi) {
                const float val = pi == 0 ? h[k0] : h[k0 + pi] * prev_excl;
                prev_excl = c[k0 + pi];
                c[k0 + pi] += val;
            }
        } else {
            float cur[kMaxLevelsTens], nxt[kMaxLevelsTens];
            int dcur = 1;
            cur[0] = h[k0];
            for (int pi = 1; pi < m; ++pi) {
                const int dn = (pi + 1 < order) ? pi + 1 : order;
                float s = 0.f;
                for (int l = 0; l < dcur; ++l) s += cur[l];
                nxt[0] = h[k0 + pi] * c[k0 + pi - 1];
                for (int l = 1; l < dn; ++l) nxt[l] = h[k0 + pi] * cur[l - 1] / (float)(l + 1);
                c[k0 + pi - 1] += s;
                for (int l = 0; l < dn; ++l) cur[l] = nxt[l];
                dcur = dn;
            }
            float s = 0.f;
            for (int l = 0; l < dcur; ++l) s += cur[l];
            c[k0 + m - 1] += s;
        }
        k0 += m;
    }
}

__global__ void tens_vs_seq_kernel(const TvsParams p) {
    const long long per = p.nz * p.n;
    const int T = p.nlev * (p.nlev + 1) / 2;
    const int nt = p.difference ? p.L - 1 : p.L;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < per; idx += (long long)gridDim.x * blockDim.x) {
        const long long z = idx / p.n, nn = idx - z * p.n;
        float c[kMaxT], hprev[kMaxT], h[kMaxT];
        for (int k = 0; k < T; ++k) { c[k] = 0.f; hprev[k] = 0.f; }
        for (int t = 0; t < p.L; ++t) {
            for (int k = 0; k < T; ++k) {
                float v;
                if (p.increments) {  // (T, nz, 2, n, L): M[:, :, 1] - M[:, :, 0] (kernels.py:330)
                    const float* b = p.M + ((((long long)k * p.nz + z) * 2) * p.n + nn) * p.L + t;
                    v = b[p.n * p.L] - b[0];
                } else {
                    v = p.M[(((long long)k * p.nz + z) * p.n + nn) * p.L + t];
                }
                h[k] = p.difference ? v - hprev[k] : v;
                hprev[k] = v;
            }
            if (p.difference && t == 0) continue;
            tvs_step(h, p.nlev, p.order, c);
        }
        (void)nt;
        p.out[idx] = 1.f;
        int k0 = 0;
        for (int m = 1; m <= p.nlev; ++m) {
            p.out[(long long)m * per + idx] = c[k0 + m - 1];
            k0 += m;
        }
    }
}

// ---- fused: static kernel evaluated on the fly ----------------------------------------------------------------------
struct TsfParams {
    const float* Z;   // scaled tensors (T, nz, [2,] DP)
    const float* Zn;  // squared norms, same leading shape
    const float* X;   // scaled points (n, L, DP)
    const float* Xn;  // squared norms (n, L)
    int nlev, order, difference, increments;
    long long nz, n;
    int L, DP;
    KernParams kp;
    float* out;
};

template <int KIND>
__device__ __forceinline__ float kern_eval_t(float dot, float sq, float xx, float yy, const KernParams& kp) {
    if (KIND == GPSIG_KERN_LINEAR) return dot;
    if (KIND == GPSIG_KERN_RBF) return __expf(-0.5f * sq);
    if (KIND == GPSIG_KERN_COSINE) return dot / (sqrtf(xx) * sqrtf(yy));
    if (KIND == GPSIG_KERN_POLY) return powf(dot + kp.a, kp.b);
    if (KIND == GPSIG_KERN_MIX) return kp.a * __expf(-0.5f * sq) + (1.f - kp.a) * dot;
    const float r = sqrtf(fmaxf(sq, 1e-40f));
    if (KIND == GPSIG_KERN_MATERN12) return __expf(-r);
    if (KIND == GPSIG_KERN_MATERN32) { const float t = 1.7320508075688772f * r; return (1.f + t) * __expf(-t); }
    const float t = 2.23606797749979f * r;
    return (1.f + t + (5.f / 3.f) * r * r) * __expf(-t);
}

// block = one inducing tensor z (its T [x2] component points staged in shared memory) x 128 sequences
template <int KIND>
__global__ void __launch_bounds__(128) tens_seq_fused_kernel(const TsfParams p) {
    extern __shared__ float sz[];  // [T * ninc][DP] then norms [T * ninc]
    const int T = p.nlev * (p.nlev + 1) / 2, ninc = p.increments ? 2 : 1, DP = p.DP;
    const long long z = blockIdx.y;
    const int npts = T * ninc;
    for (int e = threadIdx.x; e < npts * DP; e += blockDim.x) {
        const int pt = e / DP, cidx = e - pt * DP;
        const int k = pt / ninc, w = pt - k * ninc;
        sz[e] = p.Z[(((long long)k * p.nz + z) * ninc + w) * DP + cidx];
    }
    float* szn = sz + npts * DP;
    for (int e = threadIdx.x; e < npts; e += blockDim.x) {
        const int k = e / ninc, w = e - k * ninc;
        szn[e] = p.Zn[((long long)k * p.nz + z) * ninc + w];
    }
    __syncthreads();
    const long long nn = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (nn >= p.n) return;
    float c[kMaxT], hprev[kMaxT], h[kMaxT];
    for (int k = 0; k < T; ++k) { c[k] = 0.f; hprev[k] = 0.f; }
    for (int t = 0; t < p.L; ++t) {
        const float* x = p.X + (nn * p.L + t) * DP;
        const float xn = p.Xn[nn * p.L + t];
        for (int k = 0; k < T; ++k) {
            float v = 0.f;
            for (int w = 0; w < ninc; ++w) {
                const float* zp = sz + (k * ninc + w) * DP;
                float dot = 0.f, sq = 0.f;
                for (int cc = 0; cc < DP; ++cc) {
                    const float a = zp[cc], b = x[cc];
                    dot = fmaf(a, b, dot);
                    const float df = a - b;
                    sq = fmaf(df, df, sq);
                }
                const float f = KIND == GPSIG_KERN_SPECTRAL ? spectral_eval(zp, x, DP, p.kp)
                                                            : kern_eval_t<KIND>(dot, sq, szn[k * ninc + w], xn, p.kp);
                v = (ninc == 2) ? (w == 0 ? -f : v + f) : f;
            }
            h[k] = p.difference ? v - hprev[k] : v;
            hprev[k] = v;
        }
        if (p.difference && t == 0) continue;
        tvs_step(h, p.nlev, p.order, c);
    }
    const long long per = p.nz * p.n, idx = z * p.n + nn;
    p.out[idx] = 1.f;
    int k0 = 0;
    for (int m = 1; m <= p.nlev; ++m) {
        p.out[(long long)m * per + idx] = c[k0 + m - 1];
        k0 += m;
    }
}

template <int KIND>
static int launch_tsf(const TsfParams& p, cudaStream_t st) {
    const int T = p.nlev * (p.nlev + 1) / 2, ninc = p.increments ? 2 : 1;
    const size_t smem = (size_t)T * ninc * (p.DP + 1) * sizeof(float);
    dim3 grid((unsigned)((p.n + 127) / 128), (unsigned)p.nz);
    ProfScope prof(GPSIG_PROF_TENS, st, (double)p.nz * p.n);
    auto k = tens_seq_fused_kernel<KIND>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<<<grid, 128, smem, st>>>(p);
    return check_launch();
}


// ---- fast fused tensor-vs-sequence kernel (first order, LINEAR / RBF) ----------------------------------------------
// kernels.py:313-340 + signature_algs.py:101-127 without the (T, nz, n, L) Gram: one thread sweeps time for TWO
// sequences against ONE inducing tensor whose T component points sit in shared memory (broadcast reads); the T
// running prefixes of both sequences live in registers (everything is unrolled over the component index).
//   LINEAR: h_k(t) = <dz_k, dx_t>        (dz = z^1 - z^0 with increments, dx = time increment: bilinearity of :330, :114)
//   RBF   : points scaled so that log2 k = -|z - x|^2 (prep mode 3).  With w = x_t - z^0_k,
//             k(z^0_k, x_t) = 2^(-|w|^2)                                        (direct differences)
//             k(z^1_k, x_t) = 2^(-|w|^2 + 2 <w, dz_k> - |dz_k|^2),  dz_k = z^1_k - z^0_k   (anchored on z^0_k)
//           so every term is bounded by the distance of x_t to the tensor point and by the tensor's own increment -- no
//           global centre (round 1 expanded around X[0, 0, :]).  A tensor whose increment is long (|dz|^2 above
//           kWfJumpThreshold in scaled units) takes the direct form for z^1 too; the test is uniform over the warp.
//           v_k(t) = k(z^1) - k(z^0) (or k(z^0) without increments);  h_k(t) = v_k(t) - v_k(t-1).
// The arithmetic runs on packed add / fma.rn.f32x2.
__device__ __forceinline__ float tsf_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct TsfFastParams {
    const float* Z;   // prepared tensor points (T, nz, ninc, DPA)
    const float* X;   // prepared sequence points / increments (n, rowsX, DPA)
    long long nz, n;
    int rowsX;        // L (points) or L - 1 (increments)
    int increments, difference;
    float* out;       // (NLEV + 1, nz, n)
    const unsigned* tcflag;  // non-NULL: the tcgen05 kernel (tens_tc.cu) takes the call when the data is compact
};

constexpr int kTsfZPerBlock = 4;   // warps per block, one inducing tensor each
constexpr int kTsfSeqPerThread = 2;

template <bool RBF, int NLEV, int DPA>
__global__ void __launch_bounds__(kTsfZPerBlock * 32) tens_seq_fast_kernel(const TsfFastParams p) {
    constexpr int T = NLEV * (NLEV + 1) / 2, H = DPA / 2, NS = kTsfSeqPerThread;
    if (p.tcflag != nullptr && __uint_as_float(*p.tcflag) <= kTcRadius2) return;  // tens_tc.cu did this call
    // per warp: [T][nst][DPA] then [T] floats; LINEAR stores dz (nst = 1); RBF with increments stores z^0 and 2 dz, and
    // -|dz|^2 in the trailing array
    extern __shared__ __align__(16) float szf[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ninc = p.increments ? 2 : 1;
    const int nst = (RBF && p.increments) ? 2 : 1;  // vectors kept per component in shared memory
    const long long z = (long long)blockIdx.y * kTsfZPerBlock + warp;
    const bool zok = z < p.nz;
    const int per_warp = T * nst * DPA + ((T + 3) & ~3);
    float* sz = szf + (size_t)warp * per_warp;
    float* snd = sz + T * nst * DPA;
    float zmax = 0.f;
    if (zok) {
        for (int e = lane; e < T * nst * DPA; e += 32) {
            const int k = e / (nst * DPA), r = e - k * nst * DPA, w = r / DPA, c = r - w * DPA;
            const float* src = p.Z + (((long long)k * p.nz + z) * ninc) * DPA;
            float v;
            if (p.increments) {
                const float dz = src[DPA + c] - src[c];
                v = RBF ? (w == 0 ? src[c] : dz + dz) : dz;
            } else {
                v = src[c];
            }
            sz[e] = v;
        }
        if (RBF && p.increments) {
            for (int k = lane; k < T; k += 32) {
                const float* src = p.Z + (((long long)k * p.nz + z) * 2) * DPA;
                float acc = 0.f;
                for (int c = 0; c < DPA; ++c) { const float dz = src[DPA + c] - src[c]; acc = fmaf(dz, dz, acc); }
                snd[k] = -acc;
                zmax = fmaxf(zmax, acc);
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) zmax = fmaxf(zmax, __shfl_xor_sync(0xffffffffu, zmax, o));
    const bool zjump = zmax > kWfJumpThreshold;  // warp-uniform
    __syncwarp();
    long long nn[NS];
    bool nok[NS];
#pragma unroll
    for (int a = 0; a < NS; ++a) {
        nn[a] = ((long long)blockIdx.x * NS + a) * 32 + lane;
        nok[a] = nn[a] < p.n;
    }
    if (!zok) return;
    float c[NS][T], vprev[NS][T];
#pragma unroll
    for (int a = 0; a < NS; ++a)
#pragma unroll
        for (int k = 0; k < T; ++k) { c[a][k] = 0.f; vprev[a][k] = 0.f; }
    const bool tdiff = RBF && p.difference;  // LINEAR takes its time difference from the prepared increments
    // the point of time step t + 1 is fetched while step t is computed
    float2 xnext[NS][H];
    auto fetch = [&](int t, float2 (&dst)[NS][H]) {
#pragma unroll
        for (int a = 0; a < NS; ++a) {
            const float4* xs = reinterpret_cast<const float4*>(p.X + ((nok[a] ? nn[a] : 0) * p.rowsX + t) * DPA);
#pragma unroll
            for (int h4 = 0; h4 < DPA / 4; ++h4) {
                const float4 v = __ldg(xs + h4);
                dst[a][2 * h4] = make_float2(v.x, v.y);
                dst[a][2 * h4 + 1] = make_float2(v.z, v.w);
            }
        }
    };
    fetch(0, xnext);
    for (int t = 0; t < p.rowsX; ++t) {
        float2 x[NS][H];
#pragma unroll
        for (int a = 0; a < NS; ++a)
#pragma unroll
            for (int h = 0; h < H; ++h) x[a][h] = xnext[a][h];
        if (t + 1 < p.rowsX) fetch(t + 1, xnext);
        int k = 0;
#pragma unroll
        for (int m = 1; m <= NLEV; ++m) {
            float prev_excl[NS];
#pragma unroll
            for (int a = 0; a < NS; ++a) prev_excl[a] = 0.f;
#pragma unroll
            for (int pi = 0; pi < m; ++pi, ++k) {
                float v[NS];
                float2 z0[H];
                {
                    const float4* zp = reinterpret_cast<const float4*>(sz + (k * nst) * DPA);
#pragma unroll
                    for (int h4 = 0; h4 < DPA / 4; ++h4) {
                        const float4 u0 = zp[h4];
                        z0[2 * h4] = make_float2(u0.x, u0.y); z0[2 * h4 + 1] = make_float2(u0.z, u0.w);
                    }
                }
                if (RBF) {
                    float2 w[NS][H];
                    float nw[NS];
#pragma unroll
                    for (int a = 0; a < NS; ++a) {
#pragma unroll
                        for (int h = 0; h < H; ++h) w[a][h] = __fadd2_rn(x[a][h], make_float2(-z0[h].x, -z0[h].y));
                        float2 ww = __fmul2_rn(w[a][0], w[a][0]);
#pragma unroll
                        for (int h = 1; h < H; ++h) ww = __ffma2_rn(w[a][h], w[a][h], ww);
                        nw[a] = ww.x + ww.y;
                        v[a] = tsf_ex2(-nw[a]);
                    }
                    if (nst == 2) {
                        float2 d2[H];
                        const float4* zp = reinterpret_cast<const float4*>(sz + (k * 2 + 1) * DPA);
#pragma unroll
                        for (int h4 = 0; h4 < DPA / 4; ++h4) {
                            const float4 u1 = zp[h4];
                            d2[2 * h4] = make_float2(u1.x, u1.y); d2[2 * h4 + 1] = make_float2(u1.z, u1.w);
                        }
                        const float nd = snd[k];
                        if (!zjump) {
#pragma unroll
                            for (int a = 0; a < NS; ++a) {
                                float2 acc = make_float2(nd - nw[a], 0.f);
#pragma unroll
                                for (int h = 0; h < H; ++h) acc = __ffma2_rn(w[a][h], d2[h], acc);
                                v[a] = tsf_ex2(acc.x + acc.y) - v[a];
                            }
                        } else {
                            const float2 mh = make_float2(-0.5f, -0.5f);
#pragma unroll
                            for (int a = 0; a < NS; ++a) {
                                float2 w1 = __ffma2_rn(d2[0], mh, w[a][0]);
                                float2 ww = __fmul2_rn(w1, w1);
#pragma unroll
                                for (int h = 1; h < H; ++h) {
                                    w1 = __ffma2_rn(d2[h], mh, w[a][h]);
                                    ww = __ffma2_rn(w1, w1, ww);
                                }
                                v[a] = tsf_ex2(-(ww.x + ww.y)) - v[a];
                            }
                        }
                    }
                } else {
#pragma unroll
                    for (int a = 0; a < NS; ++a) {
                        float2 a0 = __fmul2_rn(z0[0], x[a][0]);
#pragma unroll
                        for (int h = 1; h < H; ++h) a0 = __ffma2_rn(z0[h], x[a][h], a0);
                        v[a] = a0.x + a0.y;
                    }
                }
#pragma unroll
                for (int a = 0; a < NS; ++a) {
                    const float h = tdiff ? v[a] - vprev[a][k] : v[a];
                    if (tdiff) vprev[a][k] = v[a];
                    const float val = (tdiff && t == 0) ? 0.f : (pi == 0 ? h : h * prev_excl[a]);
                    prev_excl[a] = c[a][k];
                    c[a][k] += val;
                }
            }
        }
    }
    const long long per = p.nz * p.n;
#pragma unroll
    for (int a = 0; a < NS; ++a) {
        if (!nok[a]) continue;
        const long long idx = z * p.n + nn[a];
        p.out[idx] = 1.f;
        int k0 = 0;
#pragma unroll
        for (int m = 1; m <= NLEV; ++m) {
            p.out[(long long)m * per + idx] = c[a][k0 + m - 1];
            k0 += m;
        }
    }
}

template <bool RBF, int NLEV, int DPA>
static int launch_tsf_fast_inst(const TsfFastParams& p, cudaStream_t st) {
    constexpr int T = NLEV * (NLEV + 1) / 2;
    const int nst = (RBF && p.increments) ? 2 : 1;
    const size_t smem = (size_t)kTsfZPerBlock * (T * nst * DPA + ((T + 3) & ~3)) * sizeof(float);
    dim3 grid((unsigned)((p.n + 32 * kTsfSeqPerThread - 1) / (32 * kTsfSeqPerThread)),
              (unsigned)((p.nz + kTsfZPerBlock - 1) / kTsfZPerBlock));
    // units are counted once per call: with a tensor-core launch in front (tens_tc.cu) that launch carries them
    ProfScope prof(GPSIG_PROF_TENS, st, p.tcflag ? 0.0 : (double)p.nz * p.n);
    auto k = tens_seq_fast_kernel<RBF, NLEV, DPA>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<<<grid, kTsfZPerBlock * 32, smem, st>>>(p);
    return check_launch();
}

template <bool RBF, int DPA>
static int launch_tsf_fast_lev(int nlev, const TsfFastParams& p, cudaStream_t st) {
    switch (nlev) {
        case 1: return launch_tsf_fast_inst<RBF, 1, DPA>(p, st);
        case 2: return launch_tsf_fast_inst<RBF, 2, DPA>(p, st);
        case 3: return launch_tsf_fast_inst<RBF, 3, DPA>(p, st);
        case 4: return launch_tsf_fast_inst<RBF, 4, DPA>(p, st);
        case 5: return launch_tsf_fast_inst<RBF, 5, DPA>(p, st);
        case 6: return launch_tsf_fast_inst<RBF, 6, DPA>(p, st);
    }
    return GPSIG_E_UNSUPPORTED;
}

// returns GPSIG_E_UNSUPPORTED (without an error detail) when there is no instantiation for the shape
static int launch_tsf_fast(bool rbf, int nlev, int DPA, const TsfFastParams& p, cudaStream_t st) {
    if (rbf) {
        switch (DPA) {
            case 4: return launch_tsf_fast_lev<true, 4>(nlev, p, st);
            case 8: return launch_tsf_fast_lev<true, 8>(nlev, p, st);
            case 12: return launch_tsf_fast_lev<true, 12>(nlev, p, st);
            case 16: return launch_tsf_fast_lev<true, 16>(nlev, p, st);
        }
    } else {
        switch (DPA) {
            case 4: return launch_tsf_fast_lev<false, 4>(nlev, p, st);
            case 8: return launch_tsf_fast_lev<false, 8>(nlev, p, st);
            case 12: return launch_tsf_fast_lev<false, 12>(nlev, p, st);
            case 16: return launch_tsf_fast_lev<false, 16>(nlev, p, st);
        }
    }
    return GPSIG_E_UNSUPPORTED;
}

}  // namespace gpsig

using namespace gpsig;

extern "C" int gpsig_tensor_kern_levels(const float* M, int num_levels, long nz, long nz2, int increments, float* out_levels,
                                        void* stream) {
    if (!M || !out_levels || num_levels < 1 || nz < 1 || nz2 < 1) return fail(GPSIG_E_BADARG, "tensor_kern: bad arguments");
    const long long per = (long long)nz * nz2;
    long long blocks = (per + 255) / 256;
    const long long cap = (long long)num_sms() * 16;
    tensor_kern_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>(M, num_levels, nz, nz2, increments,
                                                                                             out_levels);
    return check_launch();
}

extern "C" int gpsig_tens_vs_seq_levels(const float* M, int num_levels, long nz, long n, int L, int order, int difference,
                                        int increments, float* out_levels, void* stream) {
    if (!M || !out_levels || num_levels < 1 || nz < 1 || n < 1 || L < 1) return fail(GPSIG_E_BADARG, "tens_vs_seq: bad arguments");
    if (num_levels > kMaxLevelsTens) return fail(GPSIG_E_UNSUPPORTED, "tens_vs_seq supports num_levels <= %d", kMaxLevelsTens);
    if (order < 1 || order > num_levels) return fail(GPSIG_E_BADARG, "order must be in [1, num_levels]");
    TvsParams p;
    p.M = M; p.nlev = num_levels; p.order = order; p.difference = difference ? 1 : 0; p.increments = increments ? 1 : 0;
    p.nz = nz; p.n = n; p.L = L; p.out = out_levels;
    const long long per = (long long)nz * n;
    long long blocks = (per + 127) / 128;
    const long long cap = (long long)num_sms() * 16;
    ProfScope prof(GPSIG_PROF_TENS, (cudaStream_t)stream, (double)per);
    tens_vs_seq_kernel<<<(int)(blocks < cap ? blocks : cap), 128, 0, (cudaStream_t)stream>>>(p);
    return check_launch();
}

// keep stream-ordered scratch allocations cached in the device's default pool instead of returning them to the driver
// at every synchronisation (one attribute write per device, first call only)
static void keep_pool_cached() {
    static thread_local int done_dev = -1;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev == done_dev) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        unsigned long long thr = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    done_dev = dev;
}

// Z: (T, nz, d) or (T, nz, 2, d) raw; X (n, L, d) raw.  Scaling and zero-padding to DP happen in a workspace that is
// allocated stream-ordered (a few MB: (T nz ninc + n L) (DP + 1) floats) and released after the launch.
extern "C" int gpsig_tens_seq_kern_levels(int kind, const float* params, const float* Z, long nz, int increments,
                                          const float* X, long n, int L, int d, const float* inv_lengthscales, int num_levels,
                                          int order, int difference, float* out_levels, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!Z || !X || !out_levels || nz < 1 || n < 1 || L < 1 || d < 1 || num_levels < 1)
        return fail(GPSIG_E_BADARG, "tens_seq_kern: bad arguments");
    if (num_levels > kMaxLevelsTens) return fail(GPSIG_E_UNSUPPORTED, "tens_seq_kern supports num_levels <= %d", kMaxLevelsTens);
    if (order < 1 || order > num_levels) return fail(GPSIG_E_BADARG, "order must be in [1, num_levels]");
    const int T = num_levels * (num_levels + 1) / 2, ninc = increments ? 2 : 1, DP = (d + 3) / 4 * 4;
    const long long zpts = (long long)T * nz * ninc, xpts = (long long)n * L;
    keep_pool_cached();
    // fast path: first order, LINEAR / RBF, up to 6 levels
    const bool fast = order == 1 && num_levels <= 6 && DP <= 16 && (kind == GPSIG_KERN_LINEAR || kind == GPSIG_KERN_RBF) &&
                      !(kind == GPSIG_KERN_LINEAR && difference && L < 2);
    if (fast) {
        const bool rbf = kind == GPSIG_KERN_RBF;
        const int DPA = DP;
        const int xmode = rbf ? 3 : (difference ? 1 : 0);
        const int rowsX = xmode == 1 ? L - 1 : L;
        float* fb = nullptr;
        cudaError_t fe = cudaMallocAsync((void**)&fb, (size_t)(zpts + xpts) * DPA * sizeof(float) + 256, st);
        if (fe != cudaSuccess) return (int)fe;
        float* Zs = fb;
        float* Xs = Zs + zpts * DPA;
        unsigned* tcflag = nullptr;
        int frc = GPSIG_OK;
        // tensor-core path first (returns at once on the device when the data is too spread for it), then the CUDA-core
        // kernel, which returns at once when the tensor-core kernel did the work
        if (tens_tc_supported(kind, d, num_levels, order, increments, difference, L)) {
            tcflag = reinterpret_cast<unsigned*>(reinterpret_cast<uint8_t*>(fb) + ((size_t)(zpts + xpts) * DPA * sizeof(float) + 15) / 16 * 16);
            frc = launch_tens_seq_tc(Z, nz, X, n, L, d, inv_lengthscales, num_levels, out_levels, tcflag, st);
        }
        if (!frc) frc = launch_prep_points(Z, zpts, 1, d, inv_lengthscales, rbf ? 3 : 0, DPA, Zs, nullptr, st);
        if (!frc) frc = launch_prep_points(X, n, L, d, inv_lengthscales, xmode, DPA, Xs, nullptr, st);
        if (!frc) {
            TsfFastParams fp;
            fp.Z = Zs; fp.X = Xs; fp.nz = nz; fp.n = n; fp.rowsX = rowsX;
            fp.increments = increments ? 1 : 0; fp.difference = difference ? 1 : 0; fp.out = out_levels;
            fp.tcflag = tcflag;
            frc = launch_tsf_fast(rbf, num_levels, DPA, fp, st);
        }
        cudaFreeAsync(fb, st);
        if (frc != GPSIG_E_UNSUPPORTED) return frc;
    }
    float* buf = nullptr;
    cudaError_t e = cudaMallocAsync((void**)&buf, (size_t)(zpts + xpts) * (DP + 1) * sizeof(float), st);
    if (e != cudaSuccess) return (int)e;
    float* Zs = buf;
    float* Xs = Zs + zpts * DP;
    float* Zn = Xs + xpts * DP;
    float* Xn = Zn + zpts;
    int rc = launch_prep_points(Z, zpts, 1, d, inv_lengthscales, 0, DP, Zs, Zn, st);
    if (!rc) rc = launch_prep_points(X, n, L, d, inv_lengthscales, 0, DP, Xs, Xn, st);
    if (!rc) {
        TsfParams p;
        p.Z = Zs; p.Zn = Zn; p.X = Xs; p.Xn = Xn;
        p.nlev = num_levels; p.order = order; p.difference = difference ? 1 : 0; p.increments = increments ? 1 : 0;
        p.nz = nz; p.n = n; p.L = L; p.DP = DP;
        p.kp = make_kern_params(kind, params);
        p.out = out_levels;
        switch (kind) {
            case GPSIG_KERN_LINEAR: rc = launch_tsf<GPSIG_KERN_LINEAR>(p, st); break;
            case GPSIG_KERN_RBF: rc = launch_tsf<GPSIG_KERN_RBF>(p, st); break;
            case GPSIG_KERN_COSINE: rc = launch_tsf<GPSIG_KERN_COSINE>(p, st); break;
            case GPSIG_KERN_POLY: rc = launch_tsf<GPSIG_KERN_POLY>(p, st); break;
            case GPSIG_KERN_MIX: rc = launch_tsf<GPSIG_KERN_MIX>(p, st); break;
            case GPSIG_KERN_MATERN12: rc = launch_tsf<GPSIG_KERN_MATERN12>(p, st); break;
            case GPSIG_KERN_MATERN32: rc = launch_tsf<GPSIG_KERN_MATERN32>(p, st); break;
            case GPSIG_KERN_MATERN52: rc = launch_tsf<GPSIG_KERN_MATERN52>(p, st); break;
            case GPSIG_KERN_SPECTRAL: rc = launch_tsf<GPSIG_KERN_SPECTRAL>(p, st); break;
            default: rc = fail(GPSIG_E_BADARG, "unknown static kernel kind %d", kind);
        }
    }
    cudaFreeAsync(buf, st);
    return rc;
}
