// vjp.cu -- reverse mode (vector-Jacobian products) of the two first-order recursions, SURVEY.md 8f rank 1.
//
// The reference trains through signature_algs.py:28-33 (sequence vs sequence) and :116-125 (tensor vs sequence) by
// TensorFlow autodiff, which keeps every intermediate R_m tensor of the forward graph alive.  Here each recursion gets a
// hand-written adjoint that stores NOTHING per entry: the forward state at the end of the sweep is run BACKWARDS
// (the recursions are additive: A_m[s, .] = A_m[s+1, .] - rowprefix(Delta[s, .] * A_{m-1}[s, .]), level by level), while
// the adjoint state runs the mirrored recursion from the far corner.  Cost: one forward sweep plus one backward sweep that
// does about twice the work of a forward row; traffic: Delta read twice, the adjoint written once.
//
//   sequence vs sequence, K_m = sum R_m,  R_1 = Delta,  R_{m+1} = Delta * A_m,  A_m = exclusive 2-D prefix of R_m:
//     with g_m = dL/dK_m and  B_M = g_M,  B_m = g_m + exclusive 2-D SUFFIX of (Delta * B_{m+1}),
//         dL/dDelta[s, t] = sum_{k=0}^{M-1} A_k[s, t] * B_{k+1}[s, t]                 (A_0 = 1).
//   tensor vs sequence, level m with components H_0 .. H_{m-1}:  r_0 = H_0,  r_p[t] = H_p[t] * c_{p-1}[t],
//     c_p[t] = sum_{t' < t} r_p[t'],  K_m = sum_t r_{m-1}[t]:
//         rbar_{m-1}[t] = g_m,   rbar_p[t] = sum_{t' > t} rbar_{p+1}[t'] * H_{p+1}[t']   (p < m-1),
//         dL/dH_p[t] = rbar_p[t] * c_{p-1}[t]   (p >= 1),   dL/dH_0[t] = rbar_0[t].
//
// Both operate on the INCREMENTS (what signature_algs.py:26 / :114 produce); the differencing, the static-kernel Gram
// and the scaling by lengthscales are differentiated by the host's autograd (plain tensor algebra).  First order only.
#include "internal.cuh"

namespace gpsig {

struct FoVjpParams {
    const float* D;      // increments Delta[i, s, j, t]: D[i * si + s * ss + j * sj + t]
    int n1, L1, n2, L2;  // pairs and increment-tile size
    long long si, ss, sj;
    const float* G;      // (NLEV + 1, n1, n2): dL/dK_levels (level 0 is a constant: ignored)
    float* Dbar;         // dense (n1, L1, n2, L2)
};

__device__ __forceinline__ float warp_excl_prefix(float tot, int lane) {
    float incl = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    const float ex = __shfl_up_sync(0xffffffffu, incl, 1);
    return lane == 0 ? 0.f : ex;
}
__device__ __forceinline__ float warp_excl_suffix(float tot, int lane) {
    float incl = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float v = __shfl_down_sync(0xffffffffu, incl, o);
        if (lane + o < 32) incl += v;
    }
    const float ex = __shfl_down_sync(0xffffffffu, incl, 1);
    return lane == 31 ? 0.f : ex;
}

// one warp per pair (i, j); lane l owns the columns [l * WC, (l + 1) * WC); rows are processed in lockstep (no skew: the
// backward sweep needs the row prefix from the left AND the row suffix from the right of the same row)
template <int NLEV, int WC>
__global__ void __launch_bounds__(128) sigkern_fo_vjp_kernel(const FoVjpParams p) {
    constexpr int NA = NLEV - 1;  // A_1 .. A_{NLEV-1}, Q_1 .. Q_{NLEV-1}
    const int lane = threadIdx.x & 31;
    const long long npairs = (long long)p.n1 * p.n2;
    const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
    const int c0 = lane * WC;
    for (long long pair = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); pair < npairs; pair += warps) {
        const int i = (int)(pair / p.n2), j = (int)(pair - (long long)i * p.n2);
        const float* src = p.D + i * p.si + j * p.sj;
        float* dst = p.Dbar + ((long long)i * p.L1 * p.n2 + j) * p.L2;
        const long long dss = (long long)p.n2 * p.L2;
        float g[NLEV + 1];
#pragma unroll
        for (int m = 1; m <= NLEV; ++m) g[m] = p.G[(long long)m * npairs + pair];
        float A[NA > 0 ? NA : 1][WC], Q[NA > 0 ? NA : 1][WC];
#pragma unroll
        for (int m = 0; m < NA; ++m)
#pragma unroll
            for (int c = 0; c < WC; ++c) { A[m][c] = 0.f; Q[m][c] = 0.f; }
        // ---- forward sweep: A_m after the last row ----
        for (int s = 0; s < p.L1; ++s) {
            float d[WC];
#pragma unroll
            for (int c = 0; c < WC; ++c) d[c] = (c0 + c < p.L2) ? src[s * p.ss + c0 + c] : 0.f;
#pragma unroll
            for (int m = NA; m >= 1; --m) {  // descending: A_{m-1} is still the value at row s
                float pre[WC], run = 0.f;
#pragma unroll
                for (int c = 0; c < WC; ++c) {
                    pre[c] = run;
                    run += m == 1 ? d[c] : d[c] * A[m - 2][c];
                }
                const float off = warp_excl_prefix(run, lane);
#pragma unroll
                for (int c = 0; c < WC; ++c) A[m - 1][c] += off + pre[c];
            }
        }
        // ---- backward sweep ----
        for (int s = p.L1 - 1; s >= 0; --s) {
            float d[WC];
#pragma unroll
            for (int c = 0; c < WC; ++c) d[c] = (c0 + c < p.L2) ? src[s * p.ss + c0 + c] : 0.f;
            // the forward state one row back (ascending: A_{m-1}[s] is needed for R_m[s])
#pragma unroll
            for (int m = 1; m <= NA; ++m) {
                float pre[WC], run = 0.f;
#pragma unroll
                for (int c = 0; c < WC; ++c) {
                    pre[c] = run;
                    run += m == 1 ? d[c] : d[c] * A[m - 2][c];
                }
                const float off = warp_excl_prefix(run, lane);
#pragma unroll
                for (int c = 0; c < WC; ++c) A[m - 1][c] -= off + pre[c];
            }
            // B_{k+1} = g_{k+1} + Q_{k+1} (Q_NLEV = 0);  dL/dDelta = sum_k A_k B_{k+1}
            float B[NLEV + 1][WC], out[WC];
#pragma unroll
            for (int c = 0; c < WC; ++c) {
#pragma unroll
                for (int m = 1; m <= NLEV; ++m) B[m][c] = g[m] + (m <= NA ? Q[m - 1][c] : 0.f);
                float acc = B[1][c];
#pragma unroll
                for (int k = 1; k <= NA; ++k) acc = fmaf(A[k - 1][c], B[k + 1][c], acc);
                out[c] = acc;
            }
#pragma unroll
            for (int c = 0; c < WC; ++c)
                if (c0 + c < p.L2) dst[s * dss + c0 + c] = out[c];
            // Q_m[s-1, t] = Q_m[s, t] + sum_{t' > t} Delta[s, t'] B_{m+1}[s, t']
#pragma unroll
            for (int m = 1; m <= NA; ++m) {
                float suf[WC], run = 0.f;
#pragma unroll
                for (int c = WC - 1; c >= 0; --c) {
                    suf[c] = run;
                    run += d[c] * B[m + 1][c];
                }
                const float off = warp_excl_suffix(run, lane);
#pragma unroll
                for (int c = 0; c < WC; ++c) Q[m - 1][c] += off + suf[c];
            }
        }
    }
}

template <int NLEV>
static int launch_fo_vjp_lev(const FoVjpParams& p, cudaStream_t st) {
    const long long npairs = (long long)p.n1 * p.n2;
    long long blocks = (npairs + 3) / 4;
    const long long cap = (long long)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    const int wc = (p.L2 + 31) / 32;
    ProfScope prof(GPSIG_PROF_VJP, st, (double)npairs);
    if (wc <= 1) sigkern_fo_vjp_kernel<NLEV, 1><<<(int)blocks, 128, 0, st>>>(p);
    else if (wc <= 2) sigkern_fo_vjp_kernel<NLEV, 2><<<(int)blocks, 128, 0, st>>>(p);
    else if (wc <= 4) sigkern_fo_vjp_kernel<NLEV, 4><<<(int)blocks, 128, 0, st>>>(p);
    else if (wc <= 8) sigkern_fo_vjp_kernel<NLEV, 8><<<(int)blocks, 128, 0, st>>>(p);
    else return fail(GPSIG_E_UNSUPPORTED, "sigkern_levels_vjp supports sequences of at most 257 points (got %d increments)", p.L2);
    return check_launch();
}

// ---- tensor vs sequence ---------------------------------------------------------------------------------------------
struct TvsVjpParams {
    const float* H;   // (T, nz, n, Lh) increments
    int nlev, Lh;
    long long nz, n;
    const float* G;   // (nlev + 1, nz, n)
    float* Hbar;      // (T, nz, n, Lh)
};

constexpr int kVjpMaxLevels = 10;

// one thread per (z, n), level by level (the components of different levels do not interact): forward sweep over time for
// the running prefixes, backward sweep that un-adds them while the adjoint suffix sums grow
__global__ void __launch_bounds__(128) tens_vs_seq_vjp_kernel(const TvsVjpParams p) {
    const long long per = p.nz * p.n;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < per; idx += (long long)gridDim.x * blockDim.x) {
        int k0 = 0;
        for (int m = 1; m <= p.nlev; k0 += m, ++m) {
            const float g = p.G[(long long)m * per + idx];
            const float* hp = p.H + ((long long)k0 * per + idx) * p.Lh;  // component k0 + q lives q * per * Lh further
            float* hb = p.Hbar + ((long long)k0 * per + idx) * p.Lh;
            const long long cs = per * p.Lh;
            float c[kVjpMaxLevels], e[kVjpMaxLevels];
            for (int q = 0; q < m; ++q) { c[q] = 0.f; e[q] = 0.f; }
            for (int t = 0; t < p.Lh; ++t) {
                for (int q = m - 1; q >= 1; --q) c[q] = fmaf(hp[q * cs + t], c[q - 1], c[q]);
                c[0] += hp[t];
            }
            for (int t = p.Lh - 1; t >= 0; --t) {
                float h[kVjpMaxLevels], rbar[kVjpMaxLevels];
                for (int q = 0; q < m; ++q) h[q] = hp[q * cs + t];
                c[0] -= h[0];  // back to the exclusive prefixes at time t (ascending: c_{q-1} first)
                for (int q = 1; q < m; ++q) c[q] = fmaf(-h[q], c[q - 1], c[q]);
                for (int q = 0; q < m; ++q) rbar[q] = q == m - 1 ? g : e[q];
                hb[t] = rbar[0];
                for (int q = 1; q < m; ++q) {
                    hb[q * cs + t] = rbar[q] * c[q - 1];
                    e[q - 1] = fmaf(rbar[q], h[q], e[q - 1]);
                }
            }
        }
    }
}

}  // namespace gpsig

using namespace gpsig;

extern "C" int gpsig_sigkern_levels_vjp(const float* Delta, int n1, int L1, int n2, int L2, long stride_i, long stride_s,
                                        long stride_j, int num_levels, const float* G, float* Delta_bar, void* stream) {
    if (!Delta || !G || !Delta_bar || n1 < 1 || n2 < 1 || L1 < 1 || L2 < 1 || num_levels < 1)
        return fail(GPSIG_E_BADARG, "sigkern_levels_vjp: bad arguments");
    FoVjpParams p;
    p.D = Delta; p.n1 = n1; p.L1 = L1; p.n2 = n2; p.L2 = L2;
    p.si = stride_i; p.ss = stride_s; p.sj = stride_j;
    p.G = G; p.Dbar = Delta_bar;
    cudaStream_t st = (cudaStream_t)stream;
    switch (num_levels) {
        case 1: return launch_fo_vjp_lev<1>(p, st);
        case 2: return launch_fo_vjp_lev<2>(p, st);
        case 3: return launch_fo_vjp_lev<3>(p, st);
        case 4: return launch_fo_vjp_lev<4>(p, st);
        case 5: return launch_fo_vjp_lev<5>(p, st);
        case 6: return launch_fo_vjp_lev<6>(p, st);
        case 7: return launch_fo_vjp_lev<7>(p, st);
        case 8: return launch_fo_vjp_lev<8>(p, st);
    }
    return fail(GPSIG_E_UNSUPPORTED, "sigkern_levels_vjp supports num_levels <= 8");
}

extern "C" int gpsig_tens_vs_seq_levels_vjp(const float* H, int num_levels, long nz, long n, int Lh, const float* G,
                                            float* H_bar, void* stream) {
    if (!H || !G || !H_bar || num_levels < 1 || nz < 1 || n < 1 || Lh < 1)
        return fail(GPSIG_E_BADARG, "tens_vs_seq_levels_vjp: bad arguments");
    if (num_levels > kVjpMaxLevels) return fail(GPSIG_E_UNSUPPORTED, "tens_vs_seq_levels_vjp supports num_levels <= %d", kVjpMaxLevels);
    TvsVjpParams p;
    p.H = H; p.nlev = num_levels; p.Lh = Lh; p.nz = nz; p.n = n; p.G = G; p.Hbar = H_bar;
    const long long per = (long long)nz * n;
    long long blocks = (per + 127) / 128;
    const long long cap = (long long)num_sms() * 16;
    ProfScope prof(GPSIG_PROF_VJP, (cudaStream_t)stream, (double)per);
    tens_vs_seq_vjp_kernel<<<(int)(blocks < cap ? blocks : cap), 128, 0, (cudaStream_t)stream>>>(p);
    return check_launch();
}
