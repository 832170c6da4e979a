// warpfused.cu -- K(X, X) / K(X, X2) level stacks with every warp computing AND consuming its own increment-Gram rows:
// no HBM intermediate, no shared-memory ring, no barriers between warps (second design; fused.cu
// holds the first, producer/consumer-split one).
//
// One persistent CTA per SM; each warp owns an independent stream of work items.  An item is G = 32 / LP neighbouring
// pairs (i, j0..j0+G-1); LP lanes cooperate on one pair and lane l owns the 8-COLUMN strip t in [8 l, 8 l + 8): its 8
// points of the column sequence y_j live in registers for the whole item (RBF: the strip of increment columns is shifted
// left by one so that the halo value comes from lane l - 1, which is one row AHEAD in the skew), A_m[s, t] of all levels
// too (32 registers at M = 5 -- half of the 16-column strips of sigstream.cu, which is what lets the Gram arithmetic
// fit beside the recursion).  Lanes run skewed by one row exactly as in sigstream.cu: at step T lane l evaluates the
// increments Delta[T - l, strip l] (packed fma.rn.f32x2 dot products against the row point x_i[T - l] read from the
// warp's shared-memory tile; RBF: ex2.approx of the augmented product, then the 2-D increment against the previous
// row's values) and immediately feeds them to
//         A_m[r+1, t] = A_m[r, t] + p_m ;   p_m += Delta[r, t] * A_{m-1}[r, t]
// with the running row prefix p_m arriving from the left strip by one shfl.up per level.  The row stream never stops at
// item boundaries.  Per item the warp stages two small tiles itself (coalesced loads, __syncwarp only): x_i (double
// buffered: strips cross the item boundary at different steps) and the y_j of its G pairs (each lane copies its
// points from there when ITS strip starts the item).  Arithmetic per entry is the same as gram.cu + sigstream.cu, so the
// results are bit-identical to the two-kernel path.
//
// Measured alternatives (round 1, K(X,X) N=4096 L=128 d=8 M=5): 4-column strips (half the state, 16 warps) 214 ms
// Linear / 267 ms RBF -- the 32 distinct x rows per load and twice the shuffles per entry cost more than the extra warps
// give; padding the x tile against the 2-way bank conflict of the row reads costs a warp of shared memory (165-171 ms);
// fewer warps: 10 -> 188 ms, 8 -> 190 ms.  Splitting the step loop into the LP "event" steps of an item period and
// Lrow - LP test-free steps left Linear at 153 ms (12 warps, FMA-pipe / latency bound) but took RBF from 231 to 188 ms
// (8 warps: the branches were what kept its two warps per scheduler from overlapping).
#include <stdlib.h>

#include <type_traits>

#include "internal.cuh"

namespace gpsig {

constexpr int kWfCols = 8;  // columns per lane strip

struct WfParams {
    const float* A;   // prepared row-side points / increments    (rows i, rowsA, DPA)
    const float* B;   // prepared column-side points / increments (rows j, rowsB, DPA)
    int rowsA, rowsB; // rowsA == stream rows per item (RBF: row 0 only primes the differencing)
    int P;            // padded columns per pair = 8 LP
    int LP, log2LP, G, njg;
    int n1, n2, upper_only, i_off, j_off;
    long long nitems;
    int NW;
    long long ldo;
    float* out;
    long long out_level_stride;
    int xfloats, yfloats;  // per-warp tile sizes (floats): x tile (one buffer), y tile
};

__device__ __forceinline__ float wf_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct WfTrack { int i, rel, cnt; };  // item -> row i, position of its group among the groups row i keeps

__device__ __forceinline__ void wf_track_init(const WfParams& p, WfTrack& t, long long u) {
    int i, jg;
    if (!p.upper_only) {
        i = (int)(u / p.njg);
        jg = (int)(u - (long long)i * p.njg);
    } else {
        int lo = 0, hi = p.n1 - 1;
        while (lo < hi) {
            int mid = (lo + hi + 1) >> 1;
            if (items_before(mid, p.njg, p.G, 1, p.i_off, p.j_off) <= u) lo = mid; else hi = mid - 1;
        }
        i = lo;
        jg = (int)(u - items_before(lo, p.njg, p.G, 1, p.i_off, p.j_off)) + first_group(lo, p.G, 1, p.i_off, p.j_off);
    }
    const int fg = first_group(i, p.G, p.upper_only, p.i_off, p.j_off);
    t.i = i; t.rel = jg - fg; t.cnt = p.njg - fg;
}
__device__ __forceinline__ void wf_track_next(const WfParams& p, WfTrack& t) {
    t.rel += p.NW;
    while (t.rel >= t.cnt && t.i + 1 < p.n1) {
        t.rel -= t.cnt;
        ++t.i;
        t.cnt = p.njg - first_group(t.i, p.G, p.upper_only, p.i_off, p.j_off);
    }
}
__device__ __forceinline__ int wf_track_jg(const WfParams& p, const WfTrack& t) {
    return t.rel + first_group(t.i, p.G, p.upper_only, p.i_off, p.j_off);
}

// DPA = floats per prepared point; HU = leading float2 pairs that carry data (DPA/2 LINEAR, DPA/2 - 1 RBF)
template <bool RBF, int NLEV, int DPA, int HU, int MAXW>
__global__ void __launch_bounds__(MAXW * 32, 1) sigkern_warpfused_kernel(const WfParams p) {
    extern __shared__ __align__(16) float wsm[];
    constexpr int NA = NLEV > 1 ? NLEV - 1 : 1;
    constexpr int W = kWfCols;
    constexpr int NPT = W;
    const int nwarps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Lrow = p.rowsA, LP = p.LP;
    const long long wg = (long long)blockIdx.x * nwarps + warp;
    const long long nloc = wg >= p.nitems ? 0 : (p.nitems - wg + p.NW - 1) / p.NW;
    const long long total = nloc * Lrow;
    if (total == 0) return;
    const long long nsteps = total + LP - 1;
    float* xt = wsm + (size_t)warp * (2 * p.xfloats + p.yfloats);  // x tile, two buffers
    float* yt = xt + 2 * p.xfloats;                                 // y tile of the item strip 0 entered last
    const int l = lane & (LP - 1), q = lane >> p.log2LP;
    const int t0 = l * W;

    float A[NA][W];
    float psum[NLEV], ksum[NLEV];
    float2 y[NPT][HU];
    float fprev[NPT];
    float f7 = 0.f, flprev = 0.f;  // RBF: this lane's last column value of the previous step / left-halo value of the previous row
#pragma unroll
    for (int m = 0; m < NLEV; ++m) { psum[m] = 0.f; ksum[m] = 0.f; }
#pragma unroll
    for (int m = 0; m < NA; ++m)
#pragma unroll
        for (int j = 0; j < W; ++j) A[m][j] = 0.f;
#pragma unroll
    for (int u = 0; u < NPT; ++u) {
        fprev[u] = 0.f;
#pragma unroll
        for (int h = 0; h < HU; ++h) y[u][h] = make_float2(0.f, 0.f);
    }

    WfTrack tx, ty;      // item whose tiles are staged next (warp-uniform) / item this lane works on
    wf_track_init(p, tx, wg);
    ty = tx;
    int s = -l;          // row of this lane's current item (negative: not started)
    int par = 0;         // x-tile buffer of this lane's current item
    int par0 = 0;        // x-tile buffer of strip 0's item (warp-uniform)

    // CHECK: lanes may be outside their stream (first LP - 1 and last LP - 1 steps of the warp).  EV: the step lies in the
    // first LP steps of an item period, where things happen -- strip 0 stages the tiles (`stage`), every strip copies its
    // column points and resets its state when it enters the item, the last strip writes the finished item.  The other
    // Lrow - LP steps of a period run the same arithmetic without any of those tests.
    auto step = [&](auto check_tag, auto ev_tag, long long T, bool stage) {
        constexpr bool CHECK = decltype(check_tag)::value;
        constexpr bool EV = decltype(ev_tag)::value;
        // ---- strip 0 enters a new item: the warp stages x_i (other buffer) and the y_j of the G pairs ----
        if (EV && stage) {
            const int i = tx.i, jg0 = wf_track_jg(p, tx) * p.G;
            wf_track_next(p, tx);
            const float4* srcx = reinterpret_cast<const float4*>(p.A + (long long)(p.i_off + i) * p.rowsA * DPA);
            float4* dstx = reinterpret_cast<float4*>(xt + par0 * p.xfloats);  // par0 flips after the fill
            for (int e = lane; e < p.rowsA * (DPA / 4); e += 32) dstx[e] = __ldg(srcx + e);
            float4* dsty = reinterpret_cast<float4*>(yt);
            const int per = p.rowsB * (DPA / 4);
            for (int g = 0; g < p.G; ++g) {
                int jl = jg0 + g;
                if (jl > p.n2 - 1) jl = p.n2 - 1;  // padding pair of a ragged last group
                const float4* srcy = reinterpret_cast<const float4*>(p.B + (long long)(p.j_off + jl) * p.rowsB * DPA);
                for (int e = lane; e < per; e += 32) dsty[g * per + e] = __ldg(srcy + e);
            }
            __syncwarp();
            par0 ^= 1;
        }
        const bool valid = CHECK ? (s >= 0 && T - l < total) : true;
        // ---- this strip enters the item: its column points move from the tile into registers ----
        if (EV && valid && s == 0) {
#pragma unroll
            for (int u = 0; u < NPT; ++u) {
                const int t = t0 + u;
                const bool ok = RBF || t < p.rowsB;
                const int tc = t < p.rowsB ? t : p.rowsB - 1;
                const float2* src = reinterpret_cast<const float2*>(yt + ((size_t)q * p.rowsB + tc) * DPA);
#pragma unroll
                for (int h = 0; h < HU; ++h) y[u][h] = ok ? src[h] : make_float2(0.f, 0.f);
                if (RBF) {  // column side of the augmented product: (..., 1, -|y|^2/2)
                    const float2 a = y[u][HU - 1];
                    y[u][HU - 1] = make_float2(a.y, a.x);
                }
            }
#pragma unroll
            for (int m = 0; m < NLEV; ++m) ksum[m] = 0.f;
#pragma unroll
            for (int m = 0; m < NA; ++m)
#pragma unroll
                for (int j = 0; j < W; ++j) A[m][j] = 0.f;
        }
        // running row prefixes arrive from the strip to the left (it finished this row one step ago)
        float pin[NLEV];
#pragma unroll
        for (int m = 0; m < NLEV; ++m) {
            pin[m] = __shfl_up_sync(0xffffffffu, psum[m], 1);
            if (l == 0) pin[m] = 0.f;
        }
        float fl = 0.f;
        if (RBF) fl = __shfl_up_sync(0xffffffffu, f7, 1);
        // ---- increments of row s of the strip ----
        float d[W];
#pragma unroll
        for (int u = 0; u < W; ++u) d[u] = 0.f;
        if (valid) {
            const float2* xs = reinterpret_cast<const float2*>(xt + par * p.xfloats + (size_t)s * DPA);
            float2 x[HU];
#pragma unroll
            for (int h = 0; h < HU; ++h) x[h] = xs[h];
            float f[NPT];
#pragma unroll
            for (int u = 0; u < NPT; ++u) {
                float2 acc = make_float2(0.f, 0.f);
#pragma unroll
                for (int h = 0; h < HU; ++h) acc = __ffma2_rn(x[h], y[u][h], acc);
                const float v = acc.x + acc.y;
                f[u] = RBF ? wf_ex2(v) : v;
            }
            if (RBF) {
                // lane l owns the increment columns 8 l - 1 .. 8 l + 6 (column -1 is a zero pad): the value to the LEFT of
                // its first point belongs to lane l - 1, which evaluated this very row one step ago
                if (!EV || s > 0) {
                    d[0] = l == 0 ? 0.f : (f[0] - fl) - (fprev[0] - flprev);
#pragma unroll
                    for (int u = 1; u < W; ++u) d[u] = (f[u] - f[u - 1]) - (fprev[u] - fprev[u - 1]);
                }
#pragma unroll
                for (int u = 0; u < NPT; ++u) fprev[u] = f[u];
                flprev = fl;
                f7 = f[W - 1];
            } else {
#pragma unroll
                for (int u = 0; u < W; ++u) d[u] = f[u];
            }
        }
        // ---- the recursion: 2 FP ops per entry per level ----
#pragma unroll
        for (int m = 0; m < NLEV; ++m) psum[m] = pin[m];
#pragma unroll
        for (int j = 0; j < W; ++j) {
            const float dj = d[j];
#pragma unroll
            for (int m = NLEV - 1; m >= 1; --m) {
                const float a_prev = A[m - 1][j];
                if (m < NLEV - 1) A[m][j] += psum[m];
                psum[m] = fmaf(dj, a_prev, psum[m]);
            }
            if (NLEV > 1) A[0][j] += psum[0];
            psum[0] += dj;
        }
        if (valid) {
#pragma unroll
            for (int m = 0; m < NLEV; ++m) ksum[m] += psum[m];
            if (EV && s == Lrow - 1 && l == LP - 1) {
                const int j = wf_track_jg(p, ty) * p.G + q;
                if (j < p.n2) {
                    float* o = p.out + (long long)(p.i_off + ty.i) * p.ldo + p.j_off + j;
                    o[0] = 1.f;
#pragma unroll
                    for (int m = 0; m < NLEV; ++m) o[(long long)(m + 1) * p.out_level_stride] = ksum[m];
                }
            }
        }
        // ---- advance the row counters (the lane's started at -l) ----
        if (++s == Lrow) { s = 0; par ^= 1; wf_track_next(p, ty); }  // strip 0 wraps on the last step of a period
    };

    const std::true_type yes{};
    const std::false_type no{};
    for (long long k = 0; k < nloc; ++k) {  // one period per item of strip 0
        const long long base = k * Lrow;
        if (k == 0) {
            for (int e = 0; e < LP; ++e) step(yes, yes, base + e, e == 0);
        } else {
            for (int e = 0; e < LP; ++e) step(no, yes, base + e, e == 0);
        }
        for (long long T = base + LP; T < base + Lrow; ++T) step(no, no, T, false);
    }
    for (long long T = total; T < nsteps; ++T) step(yes, yes, T, false);  // the other strips finish the last item
}

// ---------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------
static int wf_log2(int x) {
    int l = 0;
    while ((1 << l) < x) ++l;
    return l;
}

// lanes per pair for `ncols` increment columns with 8-column strips (>= 2)
int wf_lanes_per_pair(int ncols) {
    int need = (ncols + kWfCols - 1) / kWfCols;
    return 1 << wf_log2(need < 2 ? 2 : need);
}

// Default path of K(X, X) / K(X, X2) for the shapes it is instantiated for (measured on the headline shape: Linear 153 ms,
// RBF 188 ms per step against 190-210 / 202-217 ms for producer + stream recursion).  GPSIG_WARPFUSED=0 switches back to
// the two-kernel pipeline (bench.py does that for its "pipeline" pass).
bool warpfused_supported(bool rbf, int d, int nlev, int ncols, int rowsA) {
    (void)rbf;
    if (nlev < 2 || nlev > 5 || d > 8) return false;
    if (ncols + 1 > 32 * kWfCols || rowsA < 48) return false;
    const char* v = getenv("GPSIG_WARPFUSED");
    return !(v && *v == '0');
}

template <bool RBF, int NLEV, int DPA, int HU, int MAXW>
static int launch_wf_maxw(WfParams& p, cudaStream_t st) {
    auto kern = sigkern_warpfused_kernel<RBF, NLEV, DPA, HU, MAXW>;
    const size_t per_warp = (size_t)(2 * p.xfloats + p.yfloats) * sizeof(float);
    int nw = MAXW;
    {
        const char* v = getenv("GPSIG_WARPFUSED_WARPS");
        if (v && *v) { nw = atoi(v); if (nw < 1) nw = 1; if (nw > MAXW) nw = MAXW; }
    }
    while (nw > 1 && per_warp * nw > 232448) --nw;
    if (per_warp * nw > 232448) return GPSIG_E_UNSUPPORTED;
    const size_t smem = per_warp * nw;
    long long want = (p.nitems + nw - 1) / nw;
    const int grid = (int)(want < num_sms() ? want : num_sms());
    p.NW = grid * nw;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    ProfScope prof(GPSIG_PROF_FUSED, st, (double)p.nitems * p.G);
    kern<<<grid, nw * 32, smem, st>>>(p);
    return check_launch();
}

// register budget per thread: 12 warps -> 168, 8 warps -> 255 (allocation granule: 4 warps).  LINEAR fits 168; RBF needs
// ~230 (the exponent arguments of 8 points and the previous row's values), so it runs 8 warps per SM (a 12-warp build
// spills ~400 bytes per thread and was measured at 445 ms against 231 ms)
template <bool RBF, int NLEV, int DPA, int HU>
static int launch_wf_inst(WfParams& p, cudaStream_t st) {
    return launch_wf_maxw<RBF, NLEV, DPA, HU, RBF ? 8 : 12>(p, st);
}

template <bool RBF, int DPA, int HU>
static int launch_wf_lev(int nlev, WfParams& p, cudaStream_t st) {
    switch (nlev) {
        case 2: return launch_wf_inst<RBF, 2, DPA, HU>(p, st);
        case 3: return launch_wf_inst<RBF, 3, DPA, HU>(p, st);
        case 4: return launch_wf_inst<RBF, 4, DPA, HU>(p, st);
        case 5: return launch_wf_inst<RBF, 5, DPA, HU>(p, st);
    }
    return GPSIG_E_UNSUPPORTED;
}

// Level stacks of the pair block rows [i_off, i_off + n1) x cols [j_off, j_off + n2) from prepared points (gram.cu prep
// modes 1 / 2).  `ncols` = increment columns per pair.  Returns GPSIG_E_UNSUPPORTED when there is no instantiation.
// Rows [i_off, i_off + n1) against all n2_total columns (symmetric: only the groups right of the diagonal); output
// out[m * lvl_stride + i * ldo + j] with GLOBAL i, j.
int launch_sigkern_warpfused(bool rbf, const float* A, const float* B, int rowsA, int rowsB, int DPA, int ncols, int n1,
                             int n2_total, int nlev, int upper_only, int i_off, long long ldo, long long lvl_stride, float* out,
                             cudaStream_t st) {
    if (!A || !B || !out || n1 < 1 || n2_total < 1 || rowsA < 1 || rowsB < 1)
        return fail(GPSIG_E_BADARG, "sigkern_warpfused: bad sizes");
    WfParams p;
    p.A = A; p.B = B; p.rowsA = rowsA; p.rowsB = rowsB;
    p.LP = wf_lanes_per_pair(rbf ? ncols + 1 : ncols); p.log2LP = wf_log2(p.LP); p.G = 32 / p.LP; p.P = p.LP * kWfCols;
    const int j_off = upper_only ? (i_off / p.G) * p.G : 0;  // first column group any of these rows keeps
    const int n2 = n2_total - j_off;
    p.njg = (n2 + p.G - 1) / p.G;
    p.n1 = n1; p.n2 = n2; p.upper_only = upper_only ? 1 : 0; p.i_off = i_off; p.j_off = j_off;
    p.nitems = items_before(n1, p.njg, p.G, p.upper_only, i_off, j_off);
    p.ldo = ldo; p.out = out; p.out_level_stride = lvl_stride;
    p.xfloats = rowsA * DPA;
    p.yfloats = p.G * rowsB * DPA;
    if (p.nitems < 1) return GPSIG_OK;
    if (rbf) {
        if (DPA == 8) return launch_wf_lev<true, 8, 3>(nlev, p, st);
        if (DPA == 12) return launch_wf_lev<true, 12, 5>(nlev, p, st);
    } else {
        if (DPA == 4) return launch_wf_lev<false, 4, 2>(nlev, p, st);
        if (DPA == 8) return launch_wf_lev<false, 8, 4>(nlev, p, st);
    }
    return GPSIG_E_UNSUPPORTED;
}

}  // namespace gpsig
