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
                const float f = kern_eval_t<KIND>(dot, sq, szn[k * ninc + w], xn, p.kp);
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
            default: rc = fail(GPSIG_E_BADARG, "unknown static kernel kind %d", kind);
        }
    }
    cudaFreeAsync(buf, st);
    return rc;
}
