// warpfused.cu -- K(X, X) / K(X, X2) / Kdiag level stacks with every warp computing AND consuming its own increment-Gram
// rows: no HBM intermediate, no shared-memory ring, no barriers between warps.  Replaces
// kernels.py:225-230 (static-kernel Gram) + signature_algs.py:26 (2-D increment) + :28-33 (all levels) in ONE launch.
//
// One persistent CTA per SM; each warp owns an independent stream of work items.  An item is G = 32 / LP neighbouring
// pairs (i, j0..j0+G-1); LP lanes cooperate on one pair and lane l owns the 8-COLUMN strip of points t in [8 l, 8 l + 8):
// everything it needs of the column sequence y_j lives in registers for the whole item, A_m[s, t] of all levels too.
// Lanes run skewed by one row: at step T lane l evaluates the increments Delta[T - l, strip l] against the row point
// x_i[T - l] (read from the warp's shared-memory tile) and immediately feeds them to
//         A_m[r+1, t] = A_m[r, t] + p_m ;   p_m += Delta[r, t] * A_{m-1}[r, t]
// with the running row prefix p_m arriving from the left strip by one shfl.up per level.  The row stream never stops at
// item boundaries.  Per item the warp stages two small tiles itself (coalesced loads, __syncwarp only): x_i (double
// buffered: strips cross the item boundary at different steps) and the y_j of its G pairs (each lane copies its points
// from there when ITS strip starts the item).
//
// Static kernels (MODE):
//   0 LINEAR  prepared data = scaled time increments; Delta[s, t] = <dx_s, dy_t> (bilinearity of kernels.py:802 + :26).
//   1 RBF, anchored form.  Points are scaled so that log2 k(x, y) = -|x - y|^2.  Every strip has a LOCAL anchor a (its
//     point number 3): with w = x_s - a (once per row) and u_t = y_t - a (once per item),
//         -|x_s - y_t|^2 = -|w|^2 + 2 <w, u_t> - |u_t|^2,
//     i.e. one packed dot product per entry like the textbook expansion, but every term is bounded by the distance of
//     the two points to something next to y_t: the fp32 error is ~1e-7 (|w| + |u|)^2 relative to k, and k underflows
//     long before |w| matters.  No global centre, no dependence on where the data sits (round 1 centred the expansion on
//     X[0, 0, :]: 4e-4 errors 30 lengthscales away).  The anchor column itself costs the plain |w|^2.
//   2 RBF, direct form sum_c (x_c - y_c)^2: used instead of mode 1 when some strip of the call has |u_t|^2 above
//     kWfJumpThreshold (a path that jumps several lengthscales within 4 steps), decided ON THE DEVICE: the prep kernel
//     leaves max |u|^2 in a flag word, both instantiations are launched and the one that does not apply returns at once.
// The recursion is identical in all modes (2 FP32 operations per entry per level).
//
// Rows are given as a LIST of row blocks (multi-GPU shards own several, parallel.py) processed by one launch.
#include <type_traits>

#include "internal.cuh"

namespace gpsig {

constexpr int kWfCols = 8;    // points per lane strip
constexpr int kWfAnchor = 3;  // strip-local anchor point (RBF mode 1)

struct WfParams {
    const float* A;   // prepared row-side data    (rows i, rowsA, D)
    const float* B;   // prepared column-side data (rows j, rowsB, D)
    // RBF anchored form of the column side, padded to LP whole strips per sequence (points past the end are copies of the
    // last one): row 8 s + 3 of Bu holds MINUS the anchor of strip s (its own u is zero and never read), the other rows
    // u_t = 2 (y_t - a);  Bnu = -|y_t - a|^2
    const float* Bu;   // (rows j, 8 LP, D)
    const float* Bnu;  // (rows j, 8 LP)
    const unsigned* flag;  // RBF: float bits of max |u|^2 over the column side (NULL = never jumpy)
    int rowsA, rowsB;      // rowsA == steps per item (RBF: row 0 only primes the differencing)
    int LP, log2LP, G;
    int NJG;               // column groups of the whole problem (global numbering)
    int n2, upper_only, diag;
    int nblk;
    int blk_begin[kWfMaxRowBlocks], blk_end[kWfMaxRowBlocks];   // global row ranges
    long long blk_out_row[kWfMaxRowBlocks];                      // output row of blk_begin
    long long blk_items0[kWfMaxRowBlocks + 1];                   // items before the block
    long long nitems;
    int NW;
    long long ldo;
    float* out;
    long long out_level_stride;
    int xfloats, yfloats;  // per-warp tile sizes (floats): x tile (one buffer), y tile
    int nufloats;          // RBF anchored: -|u|^2 tile, ONE buffer of G * 8 LP floats (double buffered like the x tile)
};

__device__ __forceinline__ float wf_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// item -> (row block, row i, position of its group among the groups row i keeps)
struct WfTrack { int blk, i, rel, cnt; };

__device__ __forceinline__ int wf_first_group(const WfParams& p, int i) { return p.upper_only ? i / p.G : 0; }
// items of rows [b, i) of a block starting at global row b
__device__ __forceinline__ long long wf_items_rows(const WfParams& p, int b, int i) {
    long long n = (long long)(i - b) * p.NJG;
    if (p.upper_only) n -= tri_floor(i, p.G) - tri_floor(b, p.G);
    return n;
}
__device__ __forceinline__ void wf_track_init(const WfParams& p, WfTrack& t, long long u) {
    int k = 0;
    while (k + 1 < p.nblk && p.blk_items0[k + 1] <= u) ++k;
    const long long v = u - p.blk_items0[k];
    int lo = p.blk_begin[k], hi = p.blk_end[k] - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (wf_items_rows(p, p.blk_begin[k], mid) <= v) lo = mid; else hi = mid - 1;
    }
    t.blk = k; t.i = lo;
    t.rel = (int)(v - wf_items_rows(p, p.blk_begin[k], lo));
    t.cnt = p.NJG - wf_first_group(p, lo);
}
__device__ __forceinline__ void wf_track_next(const WfParams& p, WfTrack& t) {
    t.rel += p.NW;
    while (t.rel >= t.cnt) {
        int ni = t.i + 1, nb = t.blk;
        if (ni >= p.blk_end[nb]) {
            if (nb + 1 >= p.nblk) break;  // past the last item: stays put (never staged, never written)
            ++nb;
            ni = p.blk_begin[nb];
        }
        t.rel -= t.cnt;
        t.i = ni; t.blk = nb;
        t.cnt = p.NJG - wf_first_group(p, ni);
    }
}
__device__ __forceinline__ int wf_track_jg(const WfParams& p, const WfTrack& t) { return t.rel + wf_first_group(p, t.i); }

// D = floats per prepared point (4 or 8)
//
// Issue slots, not only the FP32 pipe, bound this kernel (B200: packed f32x2 instructions issue once but occupy the FMA
// pipe for two cycles -- tools/ubench/pipes.cu), so everything that pairs naturally is packed: the level state is kept
// as float2 pairs (A_{2i}, A_{2i+1}) and (p_{2i}, p_{2i+1}), which turns the NLEV - 1 adds of A_m += p_m into half as many
// add.f32x2; the row differencing and the level sums likewise.  The p_m updates stay scalar FFMAs (their operands pair
// up with the OTHER parity).
template <int MODE, int NLEV, int D, int MAXW>
__global__ void __launch_bounds__(MAXW * 32, 1) sigkern_warpfused_kernel(const WfParams p) {
    constexpr bool RBF = MODE != 0;
    constexpr int NA = NLEV - 1;        // A_0 .. A_{NA-1}
    constexpr int NAP = NA / 2;         // pairs of A levels (+ one scalar level if NA is odd)
    constexpr int NPP = NLEV / 2;       // pairs of p levels (+ one scalar level if NLEV is odd)
    constexpr int W = kWfCols, H = D / 2, C4 = D / 4;
    static_assert(NLEV >= 2 && NLEV <= 6, "levels");
    if (RBF) {  // exactly one of the two RBF instantiations does the work of a call
        const bool jumpy = p.flag != nullptr && __uint_as_float(*p.flag) > kWfJumpThreshold;
        if (jumpy != (MODE == 2)) return;
    }
    extern __shared__ __align__(16) float wsm[];
    const int nwarps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Lrow = p.rowsA, LP = p.LP;
    const long long wg = (long long)blockIdx.x * nwarps + warp;
    const long long nloc = wg >= p.nitems ? 0 : (p.nitems - wg + p.NW - 1) / p.NW;
    const long long total = nloc * Lrow;
    if (total == 0) return;
    const long long nsteps = total + LP - 1;
    float* xt = wsm + (size_t)warp * (2 * p.xfloats + p.yfloats + 2 * p.nufloats);  // x tile, two buffers
    float* yt = xt + 2 * p.xfloats;                                 // y tile of the item strip 0 entered last
    float* nut = yt + p.yfloats;                                    // -|u|^2 of the column points, two buffers (read every step)
    const int l = lane & (LP - 1), q = lane >> p.log2LP;
    const int t0 = l * W;
    const int xq = p.diag ? q * Lrow * D : 0;  // diag: pair q reads its own row sequence
    const bool first = l == 0;

    float2 AP[NAP > 0 ? NAP : 1][W];
    float AS[W];                        // level NA - 1 when NA is odd
    float2 PP[NPP], KP[NPP];
    float PS = 0.f, KS = 0.f;           // level NLEV - 1 when NLEV is odd
    float2 y[W][H];      // LINEAR: dy_t;  RBF anchored: 2 (y_t - a);  RBF direct: -y_t
    float2 nanc[H];      // RBF anchored: minus the anchor a
    float2 gprev[W / 2]; // RBF: column differences f[s-1, t] - f[s-1, t-1] of the previous row
    float flast = 0.f;   // RBF: this lane's last column value of the previous step (the right neighbour's left halo)
#pragma unroll
    for (int i = 0; i < NPP; ++i) { PP[i] = make_float2(0.f, 0.f); KP[i] = make_float2(0.f, 0.f); }
#pragma unroll
    for (int j = 0; j < W; ++j) {
        AS[j] = 0.f;
#pragma unroll
        for (int i = 0; i < NAP; ++i) AP[i][j] = make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < W; ++u)
#pragma unroll
        for (int h = 0; h < H; ++h) y[u][h] = make_float2(0.f, 0.f);
#pragma unroll
    for (int u = 0; u < W / 2; ++u) gprev[u] = make_float2(0.f, 0.f);
#pragma unroll
    for (int h = 0; h < H; ++h) nanc[h] = make_float2(0.f, 0.f);

    // scalar views of the paired level state (indices are compile-time constants after unrolling)
    auto Pm = [&](int m) -> float& { return m < 2 * NPP ? ((m & 1) ? PP[m >> 1].y : PP[m >> 1].x) : PS; };
    auto Am = [&](int m, int j) -> float& { return m < 2 * NAP ? ((m & 1) ? AP[m >> 1][j].y : AP[m >> 1][j].x) : AS[j]; };

    WfTrack tx;          // item whose tiles are staged next (warp-uniform)
    wf_track_init(p, tx, wg);
    // where the results go, for the item staged last and the one before it (warp-uniform).  The last strip finishes an item
    // LP - 2 steps into the NEXT period, after the next item has been staged: it always writes the "previous" one
    int cur_j0 = 0, prev_j0 = 0;            // first column of the item's group
    int cur_row = 0, prev_row = 0;          // output row
    int s = -l;          // row of this lane's current item (negative: not started)
    int par = 0;         // x-tile buffer of this lane's current item
    int par0 = 0;        // x-tile buffer the next staging fills (warp-uniform)

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
            prev_j0 = cur_j0; prev_row = cur_row;
            cur_j0 = jg0;
            cur_row = p.diag ? 0 : (int)p.blk_out_row[tx.blk] + (i - p.blk_begin[tx.blk]);
            wf_track_next(p, tx);
            float4* dstx = reinterpret_cast<float4*>(xt + par0 * p.xfloats);
            const int perx = Lrow * C4;
            const int nx = p.diag ? p.G : 1;
            for (int g = 0; g < nx; ++g) {
                int src_seq = i;
                if (p.diag) { src_seq = jg0 + g; if (src_seq > p.n2 - 1) src_seq = p.n2 - 1; }
                const float4* srcx = reinterpret_cast<const float4*>(p.A + (long long)src_seq * Lrow * D);
                for (int e = lane; e < perx; e += 32) {
                    // D == 8: the two 16-byte halves of a row swap places every 4 rows, which makes the skewed row reads
                    // (lane l reads row T - l) conflict free
                    const int row = e / C4, c = e - row * C4;
                    const int pc = (C4 == 2) ? (c ^ ((row >> 2) & 1)) : c;
                    dstx[g * perx + row * C4 + pc] = __ldg(srcx + e);
                }
            }
            if (!p.diag || MODE == 1) {  // column side (diag, modes 0 / 2: the points are taken from the x tile)
                float4* dsty = reinterpret_cast<float4*>(yt);
                const int yrows = MODE == 1 ? LP * W : p.rowsB;  // anchored form: whole strips
                const int per = yrows * C4;
                const float* ysrc = MODE == 1 ? p.Bu : p.B;
                for (int g = 0; g < p.G; ++g) {
                    int jl = jg0 + g;
                    if (jl > p.n2 - 1) jl = p.n2 - 1;  // padding pair of a ragged last group
                    const float4* srcy = reinterpret_cast<const float4*>(ysrc + (long long)jl * yrows * D);
                    for (int e = lane; e < per; e += 32) dsty[g * per + e] = __ldg(srcy + e);
                    if (MODE == 1) {  // -|u|^2 of every strip point
                        float4* dn = reinterpret_cast<float4*>(nut + par0 * p.nufloats + g * LP * W);
                        const float4* sn = reinterpret_cast<const float4*>(p.Bnu + (long long)jl * LP * W);
                        for (int e = lane; e < LP * W / 4; e += 32) dn[e] = __ldg(sn + e);
                    }
                }
            }
            __syncwarp();
            par0 ^= 1;
        }
        // a strip that finished its item on the previous step moves on (always within the EV steps: strip l finishes at
        // period step l - 1, strip 0 on the last step of the period)
        if (EV && s == Lrow) { s = 0; par ^= 1; }
        const bool valid = CHECK ? (s >= 0 && T - l < total) : true;
        // ---- this strip enters the item: its column points move from the tile into registers ----
        if (EV && valid && s == 0) {
            if (MODE == 1) {
                // everything was put into the anchored form once per call (wf_anchor_prep_kernel), tails and strips past the
                // end included: plain copies; the anchor's own row of the tile holds -a
                const float4* src = reinterpret_cast<const float4*>(yt + ((size_t)q * LP * W + t0) * D);
#pragma unroll
                for (int u = 0; u < W; ++u) {
#pragma unroll
                    for (int c = 0; c < C4; ++c) {
                        const float4 v = src[u * C4 + c];
                        if (u == kWfAnchor) {
                            nanc[2 * c] = make_float2(v.x, v.y);
                            nanc[2 * c + 1] = make_float2(v.z, v.w);
                        } else {
                            y[u][2 * c] = make_float2(v.x, v.y);
                            y[u][2 * c + 1] = make_float2(v.z, v.w);
                        }
                    }
                }
            } else {
#pragma unroll
                for (int u = 0; u < W; ++u) {
                    const int t = t0 + u;
                    const bool ok = RBF || t < p.rowsB;          // LINEAR: increments past the end are zero
                    const int tc = t < p.rowsB ? t : p.rowsB - 1;  // RBF: clamp (equal points difference to exactly zero)
                    float2 raw[H];
                    if (p.diag) {
                        const float4* src = reinterpret_cast<const float4*>(xt + par * p.xfloats + xq) + tc * C4;
                        const int sw = (C4 == 2) ? ((tc >> 2) & 1) : 0;
#pragma unroll
                        for (int c = 0; c < C4; ++c) {
                            const float4 v = src[c ^ sw];
                            raw[2 * c] = make_float2(v.x, v.y);
                            raw[2 * c + 1] = make_float2(v.z, v.w);
                        }
                    } else {
                        const float4* src = reinterpret_cast<const float4*>(yt + ((size_t)q * p.rowsB + tc) * D);
#pragma unroll
                        for (int c = 0; c < C4; ++c) {
                            const float4 v = src[c];
                            raw[2 * c] = make_float2(v.x, v.y);
                            raw[2 * c + 1] = make_float2(v.z, v.w);
                        }
                    }
                    if (MODE == 2) {
#pragma unroll
                        for (int h = 0; h < H; ++h) y[u][h] = make_float2(-raw[h].x, -raw[h].y);
                    } else {
#pragma unroll
                        for (int h = 0; h < H; ++h) y[u][h] = ok ? raw[h] : make_float2(0.f, 0.f);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < NPP; ++i) KP[i] = make_float2(0.f, 0.f);
            KS = 0.f;
#pragma unroll
            for (int j = 0; j < W; ++j) {
                AS[j] = 0.f;
#pragma unroll
                for (int i = 0; i < NAP; ++i) AP[i][j] = make_float2(0.f, 0.f);
            }
        }
        if (EV) __syncwarp();  // the event blocks above are divergent: reconverge before the shuffles
        // running row prefixes arrive from the strip to the left (it finished this row one step ago)
#pragma unroll
        for (int m = 0; m < NLEV; ++m) {
            const float pin = __shfl_up_sync(0xffffffffu, Pm(m), 1);
            Pm(m) = first ? 0.f : pin;
        }
        float fl = 0.f;
        if (RBF) fl = __shfl_up_sync(0xffffffffu, flast, 1);
        // ---- row s of the strip, two columns at a time: evaluate -> difference -> recursion (nothing but the previous
        //      column's value and the running prefixes is carried from one pair of columns to the next) ----
        float2 x[H];
        if (valid) {
            const float4* xs = reinterpret_cast<const float4*>(xt + par * p.xfloats + xq) + s * C4;
            const int sw = (C4 == 2) ? ((s >> 2) & 1) : 0;
#pragma unroll
            for (int c = 0; c < C4; ++c) {
                const float4 v = xs[c ^ sw];
                x[2 * c] = make_float2(v.x, v.y);
                x[2 * c + 1] = make_float2(v.z, v.w);
            }
        } else {
#pragma unroll
            for (int h = 0; h < H; ++h) x[h] = make_float2(0.f, 0.f);
        }
        float nw = 0.f;      // MODE 1: |w|^2, and x[] holds w = x - a from here on
        float nu[W];         // MODE 1: -|y_t - a|^2 of the strip's points (shared memory: registers are the scarce resource)
        if (MODE == 1) {
            const float4* nup = reinterpret_cast<const float4*>(nut + par * p.nufloats + (q * LP + l) * W);
            const float4 n0 = nup[0], n1 = nup[1];
            nu[0] = n0.x; nu[1] = n0.y; nu[2] = n0.z; nu[3] = n0.w;
            nu[4] = n1.x; nu[5] = n1.y; nu[6] = n1.z; nu[7] = n1.w;
#pragma unroll
            for (int h = 0; h < H; ++h) x[h] = __fadd2_rn(x[h], nanc[h]);
            float2 ww = __fmul2_rn(x[0], x[0]);
#pragma unroll
            for (int h = 1; h < H; ++h) ww = __ffma2_rn(x[h], x[h], ww);
            nw = ww.x + ww.y;
        }
        // Gram values (RBF) / increments (LINEAR) of the 8 columns of this row, four dot products in flight at a time: the
        // accumulator chains of a column are dependent FFMA2s, interleaving four of them is what keeps the pipe fed with
        // three warps per scheduler
        float f[W];
#pragma unroll
        for (int half = 0; half < W / 4; ++half) {
            float2 acc[4];
#pragma unroll
            for (int u4 = 0; u4 < 4; ++u4) {
                const int u = 4 * half + u4;
                if (MODE == 0) acc[u4] = __fmul2_rn(x[0], y[u][0]);
                else if (MODE == 1) acc[u4] = make_float2(nu[u] - nw, 0.f);
                else { const float2 df = __fadd2_rn(x[0], y[u][0]); acc[u4] = __fmul2_rn(df, df); }
            }
#pragma unroll
            for (int h = (MODE == 1 ? 0 : 1); h < H; ++h)
#pragma unroll
                for (int u4 = 0; u4 < 4; ++u4) {
                    const int u = 4 * half + u4;
                    if (MODE == 1 && u == kWfAnchor) continue;
                    if (MODE == 2) { const float2 df = __fadd2_rn(x[h], y[u][h]); acc[u4] = __ffma2_rn(df, df, acc[u4]); }
                    else acc[u4] = __ffma2_rn(x[h], y[u][h], acc[u4]);
                }
#pragma unroll
            for (int u4 = 0; u4 < 4; ++u4) {
                const int u = 4 * half + u4;
                const float v = acc[u4].x + acc[u4].y;
                if (MODE == 0) f[u] = v;
                else if (MODE == 1) f[u] = u == kWfAnchor ? wf_ex2(-nw) : wf_ex2(v);
                else f[u] = wf_ex2(-v);
            }
        }
        float fcol = fl;     // RBF: value of the column to the left (lane l owns the increment columns 8 l - 1 .. 8 l + 6; the
                             // value left of its first point belongs to lane l - 1, which evaluated this row one step ago)
        const bool live = valid && (!RBF || !EV || s > 0);  // RBF: the first row of an item only primes the differencing
#pragma unroll
        for (int up = 0; up < W / 2; ++up) {
            const float f0 = f[2 * up], f1 = f[2 * up + 1];
            float2 dd;
            if (RBF) {
                const float2 g = make_float2(f0 - fcol, f1 - f0);
                fcol = f1;
                dd = __ffma2_rn(gprev[up], make_float2(-1.f, -1.f), g);
                if (valid) gprev[up] = g;
                if (up == 0 && first) dd.x = 0.f;  // column -1 is a zero pad
            } else {
                dd = make_float2(f0, f1);
            }
            if (!live) dd = make_float2(0.f, 0.f);
            // the recursion: 2 FP ops per entry per level
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int j = 2 * up + e;
                const float dj = e ? dd.y : dd.x;
                float pn[NLEV];  // p_m after this column (levels >= 1 read the OLD A_{m-1} and the OLD p_m)
#pragma unroll
                for (int m = 1; m < NLEV; ++m) pn[m] = fmaf(dj, Am(m - 1, j), Pm(m));
                pn[0] = Pm(0) + dj;
#pragma unroll
                for (int i = 0; i < NAP; ++i) AP[i][j] = __fadd2_rn(AP[i][j], PP[i]);
                if (NA & 1) AS[j] += Pm(NA - 1);
#pragma unroll
                for (int m = 0; m < NLEV; ++m) Pm(m) = pn[m];
            }
        }
        if (RBF && valid) flast = fcol;
        if (valid) {
#pragma unroll
            for (int i = 0; i < NPP; ++i) KP[i] = __fadd2_rn(KP[i], PP[i]);
            if (NLEV & 1) KS += PS;
            if (EV && s == Lrow - 1 && l == LP - 1) {
                const int j = prev_j0 + q;
                if (j < p.n2) {
                    float* o = p.out + (long long)prev_row * p.ldo + j;
                    o[0] = 1.f;
#pragma unroll
                    for (int m = 0; m < NLEV; ++m) {
                        const float kv = m < 2 * NPP ? ((m & 1) ? KP[m >> 1].y : KP[m >> 1].x) : KS;
                        o[(long long)(m + 1) * p.out_level_stride] = kv;
                    }
                }
            }
        }
        ++s;  // the lane's row counter started at -l
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
    prev_j0 = cur_j0; prev_row = cur_row;  // no staging any more: the item staged last is the one the tail finishes
    for (long long T = total; T < nsteps; ++T) step(yes, yes, T, false);  // the other strips finish the last item
}

// ---------------------------------------------------------------------------------------------------------------------
// column side of the anchored RBF form, once per call.  Every column sequence is padded to `rp` = 8 LP points (whole lane
// strips; points past the end are copies of the last one).  For the point y_t (scaled as in prep mode 3), with a = the
// anchor of its strip (point 8 (t / 8) + 3, clamped to the last point):
//     Bu[t] = 2 (y_t - a)   -- except the row 8 (t / 8) + 3 itself, which holds -a (the kernel never reads the anchor's u),
//     Bnu[t] = -|y_t - a|^2;   flag = max |y_t - a|^2 over the call (float bits)
// A strip entirely past the end is anchored AT the last point (u = 0 everywhere), so its anchor column evaluates the same
// value as its other columns.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void wf_anchor_prep_kernel(const float* __restrict__ B, long long n, int rows, int D, int rp,
                                      float* __restrict__ Bu, float* __restrict__ Bnu, unsigned* __restrict__ flag) {
    const long long total = n * rp;
    float worst = 0.f;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long seq = idx / rp;
        const int t = (int)(idx - seq * rp);
        const int slot = t / kWfCols * kWfCols + kWfAnchor;
        const int ta = slot < rows ? slot : rows - 1;
        const int tc = t < rows ? t : rows - 1;
        const float* yp = B + (seq * rows + tc) * D;
        const float* ap = B + (seq * rows + ta) * D;
        float acc = 0.f;
        for (int c = 0; c < D; ++c) {
            const float df = yp[c] - ap[c];
            acc = fmaf(df, df, acc);
            Bu[idx * D + c] = t == slot ? -ap[c] : df + df;
        }
        Bnu[idx] = -acc;
        worst = fmaxf(worst, acc);
    }
    for (int o = 16; o > 0; o >>= 1) worst = fmaxf(worst, __shfl_xor_sync(0xffffffffu, worst, o));
    if ((threadIdx.x & 31) == 0 && worst > 0.f) atomicMax(flag, __float_as_uint(worst));  // non-negative floats order as uints
}

int wf_lanes_per_pair(int npts);

size_t wf_anchor_bytes(long long n, int rows, int D) {
    const int rp = wf_lanes_per_pair(rows) * kWfCols;
    auto up = [](size_t x) { return (x + 255) / 256 * 256; };
    return up((size_t)n * rp * D * 4) + up((size_t)n * rp * 4);
}

// fills the two arrays (carved out of `buf`, wf_anchor_bytes() bytes) and the flag word
int launch_wf_anchor_prep(const float* B, long long n, int rows, int D, void* buf, unsigned* flag, WfAnchored* out,
                          cudaStream_t st) {
    const int rp = wf_lanes_per_pair(rows) * kWfCols;
    auto up = [](size_t x) { return (x + 255) / 256 * 256; };
    uint8_t* w = (uint8_t*)buf;
    out->Bu = (float*)w; w += up((size_t)n * rp * D * 4);
    out->Bnu = (float*)w;
    out->rows_padded = rp;
    cudaError_t e = cudaMemsetAsync(flag, 0, sizeof(unsigned), st);
    if (e != cudaSuccess) return (int)e;
    const long long total = n * rp;
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)num_sms() * 8;
    wf_anchor_prep_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, st>>>(B, n, rows, D, rp, out->Bu, out->Bnu, flag);
    return check_launch();
}

// ---------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------
static int wf_log2(int x) {
    int l = 0;
    while ((1 << l) < x) ++l;
    return l;
}

// lanes per pair for `npts` strip points with 8-point strips (>= 2)
int wf_lanes_per_pair(int npts) {
    int need = (npts + kWfCols - 1) / kWfCols;
    return 1 << wf_log2(need < 2 ? 2 : need);
}

// Default path of K(X, X) / K(X, X2) / Kdiag for LINEAR and RBF with d <= 8, 2 <= M <= 5 and sequences of 17..257 points.
// GPSIG_WARPFUSED=0 (read once, at load) switches back to the two-kernel pipeline (bench.py's "pipeline" pass).
bool warpfused_supported(bool rbf, int d, int nlev, int ncols, int rowsA) {
    if (nlev < 2 || nlev > 5 || d > 8) return false;
    const int npts = rbf ? ncols + 1 : ncols;
    if (npts > 32 * kWfCols) return false;
    if (rowsA < wf_lanes_per_pair(npts) || rowsA < 16) return false;  // an item period must cover the LP event steps
    return env_knobs().warpfused != 0;
}

template <int MODE, int NLEV, int D, int MAXW>
static int launch_wf_maxw(WfParams& p, cudaStream_t st) {
    auto kern = sigkern_warpfused_kernel<MODE, NLEV, D, MAXW>;
    const size_t per_warp = (size_t)(2 * p.xfloats + p.yfloats + 2 * p.nufloats) * sizeof(float);
    int nw = MAXW;
    if (env_knobs().warpfused_warps > 0 && env_knobs().warpfused_warps < nw) nw = env_knobs().warpfused_warps;
    while (nw > 1 && per_warp * nw > 232448) --nw;
    if (per_warp * nw > 232448) return GPSIG_E_UNSUPPORTED;
    const size_t smem = per_warp * nw;
    long long want = (p.nitems + nw - 1) / nw;
    const int grid = (int)(want < num_sms() ? want : num_sms());
    p.NW = grid * nw;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    kern<<<grid, nw * 32, smem, st>>>(p);
    return check_launch();
}

// register budget per thread: 12 warps -> 168, 8 warps -> 255 (allocation granule: 4 warps)
template <int MODE, int D>
static int launch_wf_lev(int nlev, WfParams& p, cudaStream_t st) {
    constexpr int MAXW = (MODE == 0 || MODE == 1) ? 12 : 8;
    if (nlev == 5 && MAXW == 12 && env_knobs().warpfused_warps == 8)  // experiment: 8 warps with the 255-register budget
        return launch_wf_maxw<MODE, 5, D, 8>(p, st);
    switch (nlev) {
        case 2: return launch_wf_maxw<MODE, 2, D, MAXW>(p, st);
        case 3: return launch_wf_maxw<MODE, 3, D, MAXW>(p, st);
        case 4: return launch_wf_maxw<MODE, 4, D, MAXW>(p, st);
        case 5: return launch_wf_maxw<MODE, 5, D, MAXW>(p, st);
    }
    return GPSIG_E_UNSUPPORTED;
}

// Level stacks of the listed row blocks against all n2 columns (symmetric: only the column groups right of the diagonal)
// from prepared data (gram.cu prep modes 1 / 3).  `npts` = strip points per pair (RBF: ncols + 1, LINEAR: ncols).
// out[m * lvl_stride + (blk_out_row[k] + i - blk_begin[k]) * ldo + j].  diag != 0: pairs (e, e), out[m * lvl_stride + e].
// Returns GPSIG_E_UNSUPPORTED when there is no instantiation.
int launch_sigkern_warpfused(bool rbf, const float* A, const float* B, const WfAnchored* anch, const unsigned* flag, int rowsA,
                             int rowsB, int D,
                             int npts, int n2, int nlev, int upper_only, int diag, int nblk, const int* blk_begin,
                             const int* blk_end, const long long* blk_out_row, long long ldo, long long lvl_stride, float* out,
                             cudaStream_t st) {
    if (!A || !B || !out || n2 < 1 || rowsA < 1 || rowsB < 1 || nblk < 1)
        return fail(GPSIG_E_BADARG, "sigkern_warpfused: bad sizes");
    if (nblk > kWfMaxRowBlocks) return fail(GPSIG_E_UNSUPPORTED, "at most %d row blocks per launch", kWfMaxRowBlocks);
    WfParams p;
    if (rbf && !anch) return fail(GPSIG_E_BADARG, "sigkern_warpfused: the RBF form needs the anchored column side");
    p.A = A; p.B = B; p.flag = flag; p.rowsA = rowsA; p.rowsB = rowsB;
    p.Bu = rbf ? anch->Bu : nullptr; p.Bnu = rbf ? anch->Bnu : nullptr;
    p.LP = wf_lanes_per_pair(npts); p.log2LP = wf_log2(p.LP); p.G = 32 / p.LP;
    p.n2 = n2; p.upper_only = upper_only ? 1 : 0; p.diag = diag ? 1 : 0;
    p.NJG = (n2 + p.G - 1) / p.G;
    p.nblk = diag ? 1 : nblk;
    long long items = 0;
    for (int k = 0; k < p.nblk; ++k) {
        const int b = diag ? 0 : blk_begin[k], e = diag ? 1 : blk_end[k];
        if (b < 0 || e <= b) return fail(GPSIG_E_BADARG, "sigkern_warpfused: empty row block");
        p.blk_begin[k] = b; p.blk_end[k] = e; p.blk_out_row[k] = diag ? 0 : blk_out_row[k];
        p.blk_items0[k] = items;
        long long n = (long long)(e - b) * p.NJG;
        if (p.upper_only) n -= tri_floor(e, p.G) - tri_floor(b, p.G);
        items += n;
    }
    p.blk_items0[p.nblk] = items;
    p.nitems = items;
    p.ldo = ldo; p.out = out; p.out_level_stride = lvl_stride;
    p.xfloats = (diag ? p.G : 1) * rowsA * D;
    if (p.nitems < 1) return GPSIG_OK;
    ProfScope prof(GPSIG_PROF_FUSED, st, (double)p.nitems * p.G);
    int rc = GPSIG_E_UNSUPPORTED;
    if (rbf) {
        // anchored instantiation: y tile + -|u|^2 tile + anchor tile;  direct instantiation: plain y tile (diag: none)
        if (anch->rows_padded != p.LP * kWfCols) return fail(GPSIG_E_BADARG, "sigkern_warpfused: anchored form padded for another shape");
        p.yfloats = p.G * p.LP * kWfCols * D;
        p.nufloats = p.G * p.LP * kWfCols;
        if (D == 4) rc = launch_wf_lev<1, 4>(nlev, p, st);
        if (D == 8) rc = launch_wf_lev<1, 8>(nlev, p, st);
        if (!rc && flag) {
            p.yfloats = diag ? 0 : p.G * rowsB * D;
            p.nufloats = 0;
            if (D == 4) rc = launch_wf_lev<2, 4>(nlev, p, st);
            if (D == 8) rc = launch_wf_lev<2, 8>(nlev, p, st);
        }
    } else {
        p.yfloats = diag ? 0 : p.G * rowsB * D;
        p.nufloats = 0;
        if (D == 4) rc = launch_wf_lev<0, 4>(nlev, p, st);
        if (D == 8) rc = launch_wf_lev<0, 8>(nlev, p, st);
    }
    return rc;
}

}  // namespace gpsig
