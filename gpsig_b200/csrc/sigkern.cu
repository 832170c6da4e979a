// sigkern.cu -- the level-m signature recursion over the pairwise (increment) Gram tensor.
//
// Replaces gpsig/signature_algs.py:8-35 (signature_kern_first_order): 4 slices + 3 binary ops for the 2-D increment
// (:26), then per level two exclusive cumsums + a multiply + a full reduction (:31-33) -- ~40 passes over an fp64
// tensor of n1*n2*L1*L2 entries -- by ONE pass: every Gram entry is read from HBM exactly once.
//
// Design (B200):
//   * persistent kernel, one CTA per SM, each WARP owns an independent stream of work items; an item is G = 32/LP
//     neighbouring pairs (i, j0..j0+G-1); LP lanes cooperate on one pair, lane l owns the 16-column strip
//     t in [16 l, 16 l + 16) of that pair's tile and keeps A_m[s, t] for ALL levels in registers.
//   * rows are swept once.  Lanes of a pair run SKEWED by one row (lane l works on row r while lane l-1 works on
//     row r+1), so the running row prefix p_m = sum_{t'<t} R_m[r, t'] is handed from strip to strip with a single
//     shfl.up per level per row instead of a log-step scan, and the update is 2 FP ops per entry per level:
//         A_m[r+1, t] = A_m[r, t] + p_m ;   p_m += Delta[r, t] * A_{m-1}[r, t]
//     (descending m inside a column, so A_{m-1} is still the row-r value).  K_m = sum_r p_m(end of row).
//   * the row stream never stops at item boundaries (each lane resets its own state when its row counter wraps),
//     so the skew costs LP-1 steps per kernel, not per pair.
//   * one Gram row of the G pairs (2 KB) is one TMA box (5-D tensor map, 128B swizzle -> conflict-free LDS.128 by
//     strips) landing in a per-warp ring of stages guarded by mbarriers; lane 0 re-arms a stage as soon as the last
//     strip lane has consumed it.  No CTA-wide barrier anywhere.
#include "internal.cuh"

namespace gpsig {

constexpr int kW = 16;             // columns per lane strip
constexpr int kStageBytes = 2048;  // G pairs x (16*LP) columns x 4 B, G*LP == 32
constexpr int kMaxWarps = 8;
constexpr int kMaxLevelsFast = 8;

struct FoParams {
    int n1, n2;
    int Lin;        // input rows per pair (Gram rows if DIFF else increment rows)
    int LP, log2LP; // lanes per pair
    int G;          // pairs per warp item
    int njg;        // ceil(n2 / G)
    long long nitems;
    int upper_only;
    int i_off, j_off;  // global (row, col) of local pair (0, 0); j_off is a multiple of G.  Used by upper_only and
                       // by the output address: out[(i_off + i) * ldo + j_off + j]
    long long ldo;
    int nstages;
    int log2R;      // Gram rows per TMA box (1 or 2): one bulk-tensor copy per box, one mbarrier per box
    int slot_j, slot_s, slot_i;  // tensor-map coordinate slots (2..4), dims sorted by stride
    float* out;
    long long out_level_stride;
};

// item index -> (i, jg).  With upper_only, row i only has the groups jg >= i / G.
__device__ __forceinline__ void decode_item(const FoParams& p, long long u, int& i, int& jg) {
    if (!p.upper_only) {
        i = (int)(u / p.njg);
        jg = (int)(u - (long long)i * p.njg);
        return;
    }
    int lo = 0, hi = p.n1 - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (items_before(mid, p.njg, p.G, 1, p.i_off, p.j_off) <= u) lo = mid; else hi = mid - 1;
    }
    i = lo;
    jg = (int)(u - items_before(lo, p.njg, p.G, 1, p.i_off, p.j_off)) + ((p.i_off + lo) / p.G - p.j_off / p.G);
}

template <int NLEV, bool DIFF>
__global__ void __launch_bounds__(kMaxWarps * 32, 1)
sigkern_fo_tma_kernel(const __grid_constant__ CUtensorMap tmap, const FoParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int nwarps = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = p.nstages;
    const uint32_t ring_u = smem_u32(smem) + (uint32_t)warp * S * kStageBytes;
    const uint32_t bars_u = smem_u32(smem) + (uint32_t)nwarps * S * kStageBytes + (uint32_t)warp * S * 8;

    if (lane == 0) {
        tma_prefetch_desc(&tmap);
        for (int s = 0; s < S; ++s) mbar_init(bars_u + 8 * s, 1);
        fence_mbar_init();
    }
    __syncwarp();

    const int LP = p.LP, l = lane & (LP - 1), q = lane >> p.log2LP;
    const int Lin = p.Lin;
    const long long wg = (long long)blockIdx.x * nwarps + warp, NW = (long long)gridDim.x * nwarps;
    const long long nloc = wg < p.nitems ? (p.nitems - wg + NW - 1) / NW : 0;
    const long long total = nloc * Lin;  // rows this warp streams
    if (total == 0) return;

    // swizzled byte offsets of this lane's four 16-byte chunks inside a stage (SWIZZLE_128B: chunk ^= line & 7)
    const int P4 = LP * kW * 4;  // bytes per pair row
    const uint32_t line = (uint32_t)(q * P4 + l * 64) >> 7;
    uint32_t off[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) off[k] = (line << 7) + ((((l & 1) * 4 + k) ^ (line & 7)) << 4);
    uint32_t off16 = 0;  // 17th column (first column of the next strip), DIFF only
    if (DIFF) {
        const uint32_t b = (uint32_t)(q * P4 + (l + 1) * 64);
        const uint32_t ln = b >> 7;
        off16 = (ln << 7) + (((((l + 1) & 1) * 4) ^ (ln & 7)) << 4);
    }

    // ---- producer state (meaningful on lane 0 only) ----
    long long prod_seq = 0;
    int prod_rho = 0, prod_stage = 0, prod_i = 0, prod_jg = 0;
    long long prod_item = wg;
    auto issue = [&]() {
        if (prod_rho == 0) decode_item(p, prod_item, prod_i, prod_jg);
        int c[5] = {0, 0, 0, 0, 0};
        c[p.slot_j] = prod_jg * p.G;
        c[p.slot_s] = prod_rho;
        c[p.slot_i] = prod_i;
        const uint32_t bar = bars_u + 8 * (prod_stage >> p.log2R);
        mbar_arrive_expect_tx(bar, kStageBytes << p.log2R);
        tma_load_5d(ring_u + prod_stage * kStageBytes, &tmap, bar, c[0], c[1], c[2], c[3], c[4]);
        ++prod_seq;
        prod_rho += 1 << p.log2R;  // Lin is a multiple of the box height
        if (prod_rho == Lin) { prod_rho = 0; prod_item += NW; }
        prod_stage += 1 << p.log2R;
        if (prod_stage == S) prod_stage = 0;
    };
    if (lane == 0) {
        const long long pre = total < S ? total : S;
        for (long long n = 0; n < pre; n += 1 << p.log2R) issue();
    }

    // ---- consumer state (per lane).  Issue slots bound this kernel (about 250 warp instructions per 2 KB row group), so
    // the level state is kept as float2 pairs (A_{2i}, A_{2i+1}), (p_{2i}, p_{2i+1}) as in warpfused.cu: the NLEV - 1 adds
    // A_m += p_m of a column become half as many add.f32x2, the second difference and the level sums likewise ----
    constexpr int NAL = NLEV - 1;       // A_0 .. A_{NLEV-2}
    constexpr int NAP = NAL / 2;        // pairs of A levels (+ one scalar level if NAL is odd)
    constexpr int NPP = NLEV / 2;       // pairs of p levels (+ one scalar level if NLEV is odd)
    float2 AP[NAP > 0 ? NAP : 1][kW];
    float AS[kW];
    float2 PP[NPP > 0 ? NPP : 1], KP[NPP > 0 ? NPP : 1];
    float PS = 0.f, KS = 0.f;
    float2 gdp[kW / 2];                 // DIFF: column differences g[t + 1] - g[t] of the previous row
#pragma unroll
    for (int i = 0; i < (NPP > 0 ? NPP : 1); ++i) { PP[i] = make_float2(0.f, 0.f); KP[i] = make_float2(0.f, 0.f); }
#pragma unroll
    for (int j = 0; j < kW; ++j) {
        AS[j] = 0.f;
#pragma unroll
        for (int i = 0; i < (NAP > 0 ? NAP : 1); ++i) AP[i][j] = make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int j = 0; j < kW / 2; ++j) gdp[j] = make_float2(0.f, 0.f);
    auto Pm = [&](int m) -> float& { return m < 2 * NPP ? ((m & 1) ? PP[m >> 1].y : PP[m >> 1].x) : PS; };
    auto Am = [&](int m, int j) -> float& { return m < 2 * NAP ? ((m & 1) ? AP[m >> 1][j].y : AP[m >> 1][j].x) : AS[j]; };

    long long n_l = -(long long)l;  // sequence number of the row this lane handles at the current step
    int rho = 0, stage = 0;
    uint32_t phase = 0;
    long long item = wg;

    const long long nsteps = total + LP - 1;
    for (long long T = 0; T < nsteps; ++T, ++n_l) {
        // running row prefixes arrive from the strip to the left (it finished this row one step ago)
        float pin[NLEV];
#pragma unroll
        for (int m = 0; m < NLEV; ++m) {
            pin[m] = __shfl_up_sync(0xffffffffu, Pm(m), 1);
            if (l == 0) pin[m] = 0.f;
        }
        const bool valid = (n_l >= 0) && (n_l < total);
        float d[kW];
#pragma unroll
        for (int j = 0; j < kW; ++j) d[j] = 0.f;
        bool first_row = false;
        if (valid) {
            mbar_wait(bars_u + 8 * (stage >> p.log2R), phase);
            const uint32_t base = ring_u + stage * kStageBytes;
            float g[kW + 1];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                             : "=f"(g[4 * k]), "=f"(g[4 * k + 1]), "=f"(g[4 * k + 2]), "=f"(g[4 * k + 3])
                             : "r"(base + off[k]));
            }
            if (DIFF) {
                if (l == LP - 1) g[kW] = g[kW - 1];
                else asm volatile("ld.shared.f32 %0, [%1];" : "=f"(g[kW]) : "r"(base + off16));
                first_row = (rho == 0);
#pragma unroll
                for (int j = 0; j < kW / 2; ++j) {
                    const float2 gd = make_float2(g[2 * j + 1] - g[2 * j], g[2 * j + 2] - g[2 * j + 1]);
                    if (!first_row) {
                        const float2 dd = __ffma2_rn(gdp[j], make_float2(-1.f, -1.f), gd);   // signature_algs.py:26
                        d[2 * j] = dd.x; d[2 * j + 1] = dd.y;
                    }
                    gdp[j] = gd;
                }
            } else {
                first_row = (rho == 0);
#pragma unroll
                for (int j = 0; j < kW; ++j) d[j] = g[j];
            }
        }
        __syncwarp();
        // the row consumed by the last strip lanes in this step is free in all pair groups: refill its stage
        if (lane == 0) {
            const long long freed = T - (LP - 1);
            if (p.log2R == 0) {
                if (freed >= 0 && freed + S < total) issue();
            } else if (freed >= 1 && (freed & 1) && freed - 1 + S < total) {
                issue();  // both rows of the box are free
            }
        }
        if (valid && first_row) {
#pragma unroll
            for (int i = 0; i < (NPP > 0 ? NPP : 1); ++i) KP[i] = make_float2(0.f, 0.f);
            KS = 0.f;
#pragma unroll
            for (int j = 0; j < kW; ++j) {
                AS[j] = 0.f;
#pragma unroll
                for (int i = 0; i < (NAP > 0 ? NAP : 1); ++i) AP[i][j] = make_float2(0.f, 0.f);
            }
        }
        // ---- the recursion: 2 FP ops per entry per level ----
#pragma unroll
        for (int m = 0; m < NLEV; ++m) Pm(m) = pin[m];
#pragma unroll
        for (int j = 0; j < kW; ++j) {
            const float dj = d[j];
            float pn[NLEV];  // p_m after this column (levels >= 1 read the OLD A_{m-1} and the OLD p_m)
#pragma unroll
            for (int m = 1; m < NLEV; ++m) pn[m] = fmaf(dj, Am(m - 1, j), Pm(m));
            pn[0] = Pm(0) + dj;
#pragma unroll
            for (int i = 0; i < NAP; ++i) AP[i][j] = __fadd2_rn(AP[i][j], PP[i]);
            if (NAL & 1) AS[j] += Pm(NAL - 1);
#pragma unroll
            for (int m = 0; m < NLEV; ++m) Pm(m) = pn[m];
        }
        if (valid) {
#pragma unroll
            for (int i = 0; i < NPP; ++i) KP[i] = __fadd2_rn(KP[i], PP[i]);
            if (NLEV & 1) KS += PS;
            if (rho == Lin - 1 && l == LP - 1) {
                int i, jg;
                decode_item(p, item, i, jg);
                const int j = jg * p.G + q;
                if (j < p.n2) {
                    float* o = p.out + (long long)(p.i_off + i) * p.ldo + p.j_off + j;
                    o[0] = 1.f;
#pragma unroll
                    for (int m = 0; m < NLEV; ++m)
                        o[(long long)(m + 1) * p.out_level_stride] = m < 2 * NPP ? ((m & 1) ? KP[m >> 1].y : KP[m >> 1].x) : KS;
                }
            }
            if (++rho == Lin) { rho = 0; item += NW; }
            if (++stage == S) { stage = 0; phase ^= 1u; }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Generic fallback (any L1/L2/strides, any num_levels <= 16): one warp per pair, state in shared memory, a warp scan
// per row and level.  Correct everywhere, fast nowhere; the TMA kernel above takes every shape the covariance
// pipeline produces itself.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kMaxLevelsGeneric = 16;

struct GenParams {
    const float* M;
    int n1, n2, L1, L2;
    long long si, ss, sj;
    int nlev, difference, upper_only;
    int nr, nc;  // increment rows / cols
    int cw;      // columns per lane
    int i_off, j_off;
    long long ldo;
    float* out;
    long long out_level_stride;
};

__device__ __forceinline__ float warp_excl_scan(float v, int lane, float& total) {
    float x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        float y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    total = __shfl_sync(0xffffffffu, x, 31);
    return x - v;
}

__global__ void sigkern_fo_generic_kernel(const GenParams p) {
    extern __shared__ float sm_state[];
    const int nwarps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nlev = p.nlev, nc = p.nc;
    float* A = sm_state + (size_t)warp * (nlev > 1 ? nlev - 1 : 1) * nc;  // A[m][c]
    const long long npairs = (long long)p.n1 * p.n2;
    for (long long pr = (long long)blockIdx.x * nwarps + warp; pr < npairs; pr += (long long)gridDim.x * nwarps) {
        const int i = (int)(pr / p.n2), j = (int)(pr % p.n2);
        if (p.upper_only && p.j_off + j < p.i_off + i) continue;
        const float* T = p.M + i * p.si + j * p.sj;
        for (int m = 0; m + 1 < nlev; ++m)
            for (int c = lane; c < nc; c += 32) A[m * nc + c] = 0.f;
        __syncwarp();
        float ksum[kMaxLevelsGeneric];
#pragma unroll
        for (int m = 0; m < kMaxLevelsGeneric; ++m) ksum[m] = 0.f;
        const int c0 = lane * p.cw, c1 = min(nc, c0 + p.cw);
        for (int r = 0; r < p.nr; ++r) {
            const float* row0 = T + (long long)r * p.ss;
            const float* row1 = row0 + p.ss;
            float tot[kMaxLevelsGeneric], offs[kMaxLevelsGeneric];
#pragma unroll
            for (int m = 0; m < kMaxLevelsGeneric; ++m) tot[m] = 0.f;
            // pass 1: lane totals of R_m[r, strip] = Delta * A_{m-1}
            for (int c = c0; c < c1; ++c) {
                const float dl = p.difference ? (row1[c + 1] - row1[c]) - (row0[c + 1] - row0[c]) : row0[c];
                tot[0] += dl;
#pragma unroll
                for (int m = 1; m < kMaxLevelsGeneric; ++m)
                    if (m < nlev) tot[m] = fmaf(dl, A[(m - 1) * nc + c], tot[m]);
            }
#pragma unroll
            for (int m = 0; m < kMaxLevelsGeneric; ++m) {
                if (m < nlev) {
                    float t;
                    offs[m] = warp_excl_scan(tot[m], lane, t);
                    ksum[m] += t;
                }
            }
            // pass 2: A_m += exclusive row prefix (descending m keeps A_{m-1} at its row-r value)
            for (int c = c0; c < c1; ++c) {
                const float dl = p.difference ? (row1[c + 1] - row1[c]) - (row0[c + 1] - row0[c]) : row0[c];
#pragma unroll
                for (int m = kMaxLevelsGeneric - 1; m >= 1; --m) {
                    if (m < nlev) {
                        const float a_prev = A[(m - 1) * nc + c];
                        if (m < nlev - 1) A[m * nc + c] += offs[m];
                        offs[m] = fmaf(dl, a_prev, offs[m]);
                    }
                }
                if (nlev > 1) A[c] += offs[0];
                offs[0] += dl;
            }
        }
        __syncwarp();
        if (lane == 0) {
            float* o = p.out + (long long)(p.i_off + i) * p.ldo + p.j_off + j;
            o[0] = 1.f;
#pragma unroll
            for (int m = 0; m < kMaxLevelsGeneric; ++m)
                if (m < nlev) o[(long long)(m + 1) * p.out_level_stride] = ksum[m];
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Higher-order recursion (signature_algs.py:37-74), correctness-first: one THREAD per pair sweeps its tile serially; the
// per-column state of every level (the 2-D prefix AA_m and the column prefixes V_m^k) lives in shared memory, the
// row-running prefixes (p_m, H_m^j) in thread-local arrays.  At a point (s, t), writing R_m[a][b] for the level-m grid:
//   R_{m+1}[1][1] = D * AA_m          AA_m[s,t] = sum_{s'<s,t'<t} sum_ab R_m[a][b]              (:64)
//   R_{m+1}[1][k] = D * V_m^{k-1} / k  V_m^{k'}[s,t] = sum_{s'<s} sum_a R_m[a][k'][s',t]          (:66)
//   R_{m+1}[j][1] = D * H_m^{j-1} / j  H_m^{j'}[s,t] = sum_{t'<t} sum_b R_m[j'][b][s,t']          (:67)
//   R_{m+1}[j][k] = D * R_m[j-1][k-1] / (j k)                                                    (:69)
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kMaxLevelsHO = 10;

struct HoParams {
    const float* M;
    int n1, n2;
    long long si, ss, sj;
    int nlev, order, difference, upper_only;
    int nr, nc;
    int i_off, j_off;
    long long ldo;
    float* out;
    long long out_level_stride;
};

__global__ void sigkern_ho_serial_kernel(const HoParams p) {
    extern __shared__ float hst[];  // [(q * nc + c) * ppb + tid], q = (m-1) * D + slot
    const int D = p.order, nlev = p.nlev, nc = p.nc, ppb = blockDim.x, tid = threadIdx.x;
    const long long npairs = (long long)p.n1 * p.n2;
    for (long long pr = (long long)blockIdx.x * ppb + tid; pr < npairs; pr += (long long)gridDim.x * ppb) {
        const int i = (int)(pr / p.n2), j = (int)(pr % p.n2);
        if (p.upper_only && p.j_off + j < p.i_off + i) continue;
        const float* T = p.M + i * p.si + j * p.sj;
        const int nq = (nlev - 1) * D;
        for (int q = 0; q < nq; ++q)
            for (int c = 0; c < nc; ++c) hst[((long long)q * nc + c) * ppb + tid] = 0.f;
        float K[kMaxLevelsHO];
        for (int m = 0; m < kMaxLevelsHO; ++m) K[m] = 0.f;
        for (int r = 0; r < p.nr; ++r) {
            const float* row0 = T + (long long)r * p.ss;
            const float* row1 = row0 + p.ss;
            float prow[kMaxLevelsHO];
            float H[kMaxLevelsHO][kMaxLevelsHO];
            for (int m = 0; m < nlev; ++m) {
                prow[m] = 0.f;
                for (int a = 0; a < D; ++a) H[m][a] = 0.f;
            }
            for (int c = 0; c < nc; ++c) {
                const float dl = p.difference ? (row1[c + 1] - row1[c]) - (row0[c + 1] - row0[c]) : row0[c];
                float Rc[kMaxLevelsHO][kMaxLevelsHO], Rn[kMaxLevelsHO][kMaxLevelsHO];
                Rc[0][0] = dl;
                int dc = 1;
                K[0] += dl;
                for (int m = 1; m < nlev; ++m) {  // source level m -> level m + 1
                    float* stq = hst + ((long long)(m - 1) * D * nc + c) * ppb + tid;  // slot s at stq[s * nc * ppb]
                    const long long slot = (long long)nc * ppb;
                    const int dn = (m + 1 < D) ? m + 1 : D;
                    Rn[0][0] = dl * stq[0];
                    for (int k = 2; k <= dn; ++k) Rn[0][k - 1] = dl * stq[(k - 1) * slot] / (float)k;
                    for (int jj = 2; jj <= dn; ++jj) Rn[jj - 1][0] = dl * H[m][jj - 1] / (float)jj;
                    for (int jj = 2; jj <= dn; ++jj)
                        for (int k = 2; k <= dn; ++k) Rn[jj - 1][k - 1] = dl * Rc[jj - 2][k - 2] / (float)(jj * k);
                    // fold the level-m values at this point into the level-m states
                    float tot = 0.f;
                    for (int a = 0; a < dc; ++a)
                        for (int b = 0; b < dc; ++b) tot += Rc[a][b];
                    stq[0] += prow[m];
                    prow[m] += tot;
                    const int dlim = dc < D - 1 ? dc : D - 1;
                    for (int k = 1; k <= dlim; ++k) {
                        float cs = 0.f, rs = 0.f;
                        for (int a = 0; a < dc; ++a) { cs += Rc[a][k - 1]; rs += Rc[k - 1][a]; }
                        stq[k * slot] += cs;
                        H[m][k] += rs;
                    }
                    float tn = 0.f;
                    for (int a = 0; a < dn; ++a)
                        for (int b = 0; b < dn; ++b) { tn += Rn[a][b]; Rc[a][b] = Rn[a][b]; }
                    K[m] += tn;
                    dc = dn;
                }
            }
        }
        float* o = p.out + (long long)(p.i_off + i) * p.ldo + p.j_off + j;
        o[0] = 1.f;
        for (int m = 0; m < nlev; ++m) o[(long long)(m + 1) * p.out_level_stride] = K[m];
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------------------------------
template <int NLEV, bool DIFF>
static int launch_tma_inst(const CUtensorMap& tmap, const FoParams& p, int nwarps, size_t smem, int grid, cudaStream_t st) {
    auto kern = sigkern_fo_tma_kernel<NLEV, DIFF>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    kern<<<grid, nwarps * 32, smem, st>>>(tmap, p);
    return check_launch();
}

template <bool DIFF>
static int launch_tma_lev(int nlev, const CUtensorMap& tmap, const FoParams& p, int nwarps, size_t smem, int grid,
                          cudaStream_t st) {
    switch (nlev) {
        case 1: return launch_tma_inst<1, DIFF>(tmap, p, nwarps, smem, grid, st);
        case 2: return launch_tma_inst<2, DIFF>(tmap, p, nwarps, smem, grid, st);
        case 3: return launch_tma_inst<3, DIFF>(tmap, p, nwarps, smem, grid, st);
        case 4: return launch_tma_inst<4, DIFF>(tmap, p, nwarps, smem, grid, st);
        case 5: return launch_tma_inst<5, DIFF>(tmap, p, nwarps, smem, grid, st);
        case 6: return launch_tma_inst<6, DIFF>(tmap, p, nwarps, smem, grid, st);
        case 7: return launch_tma_inst<7, DIFF>(tmap, p, nwarps, smem, grid, st);
        case 8: return launch_tma_inst<8, DIFF>(tmap, p, nwarps, smem, grid, st);
    }
    return fail(GPSIG_E_UNSUPPORTED, "fast path supports num_levels <= %d", kMaxLevelsFast);
}

static int ilog2_ceil_pow2(int x) {
    int l = 0;
    while ((1 << l) < x) ++l;
    return l;
}

// lanes-per-pair for `ncols` increment columns (>= 2 so that a pair row is at least one 128-byte swizzle line)
int fo_lanes_per_pair(int ncols) {
    int need = (ncols + kW - 1) / kW;
    int lg = ilog2_ceil_pow2(need < 2 ? 2 : need);
    return 1 << lg;
}

bool fo_tma_eligible(const float* M, int Lrows, int ncols, int pitch, long long si, long long ss, long long sj,
                     int nlev) {
    if (nlev < 1 || nlev > kMaxLevelsFast) return false;
    if (ncols < 1 || ncols > 512) return false;
    if (pitch != kW * fo_lanes_per_pair(ncols)) return false;
    if (((uintptr_t)M & 15u) || (si & 3) || (ss & 3) || (sj & 3)) return false;
    if (Lrows < 1) return false;
    return true;
}

// First-order recursion over M[n1, Lrows, n2, pitch] (element strides si, ss, sj; unit stride along t).
// `pitch` is the number of readable columns per pair row; columns >= ncols(+1 if diff) must be zero or absent.
// Output: out[m * out_level_stride + (i_off + i) * ldo + j_off + j]; upper_only compares GLOBAL indices.
int launch_sigkern_fo(const float* M, int n1, int Lrows, int n2, int ncols, int pitch, long long si, long long ss,
                      long long sj, int nlev, int difference, int upper_only, int i_off, int j_off, long long ldo,
                      long long lvl_stride, float* out, cudaStream_t st, int force_generic) {
    if (!M || !out || n1 < 1 || n2 < 1 || Lrows < 1 || nlev < 1) return fail(GPSIG_E_BADARG, "sigkern: bad sizes");
    if (!force_generic && fo_tma_eligible(M, Lrows, ncols, pitch, si, ss, sj, nlev)) {
        FoParams p;
        p.n1 = n1; p.n2 = n2; p.Lin = Lrows;
        p.LP = fo_lanes_per_pair(ncols);
        p.log2LP = ilog2_ceil_pow2(p.LP);
        p.G = 32 / p.LP;
        p.njg = (n2 + p.G - 1) / p.G;
        p.upper_only = upper_only ? 1 : 0;
        if (p.upper_only && (j_off % p.G) != 0) return fail(GPSIG_E_BADARG, "j_off must be a multiple of the pair group");
        p.i_off = i_off; p.j_off = j_off; p.ldo = ldo;
        p.nitems = items_before(n1, p.njg, p.G, p.upper_only, i_off, j_off);
        p.out = out; p.out_level_stride = lvl_stride;
        // warps per CTA and ring depth: depth = LP (skew) + prefetch, limited by 227 KB of shared memory
        const int max_smem = 232448;
        int nwarps = kMaxWarps;
        int pf = 6;
        while (nwarps > 1 && (size_t)nwarps * (p.LP + pf) * (kStageBytes + 8) > (size_t)max_smem) nwarps >>= 1;
        int S = p.LP + pf;
        // spend leftover shared memory on a deeper prefetch (bounded)
        while (S < p.LP + 12 && (size_t)nwarps * (S + 1) * (kStageBytes + 8) <= (size_t)max_smem && nwarps < kMaxWarps) ++S;
        // two Gram rows per TMA box when the rows of a pair group are adjacent in the box (stride_s > stride_j) and
        // items hold an even number of rows: halves the bulk-tensor copies, whose issue rate (~1 per 100 cycles per SM)
        // is what limits this kernel (tools/ubench/tma_stream.cu)
        p.log2R = (Lrows % 2 == 0 && ss > sj && n1 > 1) ? 1 : 0;
        if (p.log2R) {
            if (S & 1) ++S;
            while ((size_t)nwarps * S * (kStageBytes + 8) > (size_t)max_smem) S -= 2;
        }
        p.nstages = S;
        // tensor map: dims (32, pitch/32, .., .., ..) with (j, s, i) ordered by ascending stride
        struct Dim { long long stride; uint64_t size; uint32_t box; int which; };
        Dim dj{sj, (uint64_t)n2, (uint32_t)p.G, 0}, ds{ss, (uint64_t)Lrows, 1u << p.log2R, 1}, di{si, (uint64_t)n1, 1u, 2};
        if (n1 == 1) di.stride = (long long)1 << 38;  // never dereferenced beyond coordinate 0; keep it the largest
        Dim order[3] = {dj, ds, di};
        for (int a = 0; a < 3; ++a)
            for (int b = a + 1; b < 3; ++b)
                if (order[b].stride < order[a].stride) { Dim t = order[a]; order[a] = order[b]; order[b] = t; }
        uint64_t dims[5] = {32, (uint64_t)(pitch / 32), 0, 0, 0};
        uint64_t strides[4] = {128, 0, 0, 0};
        uint32_t box[5] = {32, (uint32_t)(pitch / 32), 1, 1, 1};
        for (int a = 0; a < 3; ++a) {
            dims[2 + a] = order[a].size;
            strides[1 + a] = (uint64_t)order[a].stride * 4ull;
            box[2 + a] = order[a].box;
            if (order[a].which == 0) p.slot_j = 2 + a;
            if (order[a].which == 1) p.slot_s = 2 + a;
            if (order[a].which == 2) p.slot_i = 2 + a;
        }
        if (n1 == 1) strides[p.slot_i - 1] = strides[p.slot_i - 2] * dims[p.slot_i - 1];  // plausible dense stride
        CUtensorMap tmap;
        int rc = encode_tensor_map_f32(&tmap, M, 5, dims, strides, box, 1);
        if (rc != GPSIG_OK) return rc;
        const size_t smem = (size_t)nwarps * S * (kStageBytes + 8);
        long long want = (p.nitems + nwarps - 1) / nwarps;
        int grid = (int)(want < num_sms() ? want : num_sms());
        if (grid < 1) grid = 1;
        ProfScope prof(GPSIG_PROF_RECURSION, st, (double)p.nitems * p.G);
        return difference ? launch_tma_lev<true>(nlev, tmap, p, nwarps, smem, grid, st)
                          : launch_tma_lev<false>(nlev, tmap, p, nwarps, smem, grid, st);
    }
    // generic path
    if (nlev > kMaxLevelsGeneric) return fail(GPSIG_E_UNSUPPORTED, "num_levels > %d", kMaxLevelsGeneric);
    GenParams g;
    g.M = M; g.n1 = n1; g.n2 = n2; g.L1 = Lrows; g.L2 = pitch;
    g.si = si; g.ss = ss; g.sj = sj;
    g.nlev = nlev; g.difference = difference ? 1 : 0; g.upper_only = upper_only ? 1 : 0;
    g.nr = difference ? Lrows - 1 : Lrows;
    g.nc = ncols;
    g.cw = (g.nc + 31) / 32;
    g.i_off = i_off; g.j_off = j_off; g.ldo = ldo;
    g.out = out; g.out_level_stride = lvl_stride;
    int nwarps = 4;
    size_t per_warp = (size_t)(nlev > 1 ? nlev - 1 : 1) * (g.nc > 0 ? g.nc : 1) * sizeof(float);
    while (nwarps > 1 && per_warp * nwarps > 200 * 1024) nwarps >>= 1;
    if (per_warp * nwarps > 227 * 1024) return fail(GPSIG_E_UNSUPPORTED, "sequence too long for the generic kernel");
    const size_t smem = per_warp * nwarps;
    cudaError_t e = cudaFuncSetAttribute(sigkern_fo_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    long long npairs = (long long)n1 * n2;
    long long blocks = (npairs + nwarps - 1) / nwarps;
    int grid = (int)(blocks < (long long)num_sms() * 8 ? blocks : (long long)num_sms() * 8);
    ProfScope prof(GPSIG_PROF_RECURSION_OTHER, st, (double)npairs);
    sigkern_fo_generic_kernel<<<grid, nwarps * 32, smem, st>>>(g);
    return check_launch();
}

// Higher-order recursion, one WARP per pair (the path with a number: notebooks/signature_kernel.ipynb runs order = M).
// Lane l owns the WC columns [l WC, (l + 1) WC); rows are swept in lockstep, level by level inside a row:
//   * the grid R_m[a][b] of the row lives in registers (D x D x WC values for the current and the next level);
//   * AA_m (2-D exclusive prefix of sum_ab R_m) and V_m^k (exclusive column prefix of sum_a R_m[a][k]) are per-column
//     registers carried down the rows;
//   * H_m^j (exclusive ROW prefix of sum_b R_m[j][b]) and the row prefix that feeds AA_m are warp scans of the row
//     (local prefix + 5-step shuffle scan of the lane totals): at most D scans per level per row.
// Instantiated for D * D * WC <= 64 (order <= 5 at 64 columns, order <= 4 at 128); other shapes take the serial kernel.
__device__ __forceinline__ float ho_warp_excl_prefix(float tot, int lane) {
    float incl = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    const float ex = __shfl_up_sync(0xffffffffu, incl, 1);
    return lane == 0 ? 0.f : ex;
}

template <int NLEV, int D, int WC>
__global__ void __launch_bounds__(128) sigkern_ho_warp_kernel(const HoParams p) {
    constexpr int NA = NLEV > 1 ? NLEV - 1 : 1;
    constexpr int DV = D > 1 ? D - 1 : 1;
    const int lane = threadIdx.x & 31;
    const long long npairs = (long long)p.n1 * p.n2;
    const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
    const int c0 = lane * WC;
    for (long long pr = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); pr < npairs; pr += warps) {
        const int i = (int)(pr / p.n2), j = (int)(pr % p.n2);
        if (p.upper_only && p.j_off + j < p.i_off + i) continue;
        const float* T = p.M + i * p.si + j * p.sj;
        float AA[NA][WC], V[NA][DV][WC], K[NLEV];
#pragma unroll
        for (int m = 0; m < NLEV; ++m) K[m] = 0.f;
#pragma unroll
        for (int m = 0; m < NA; ++m)
#pragma unroll
            for (int c = 0; c < WC; ++c) {
                AA[m][c] = 0.f;
#pragma unroll
                for (int k = 0; k < DV; ++k) V[m][k][c] = 0.f;
            }
        float raw0[WC + 1];  // difference: the previous raw Gram row (with the right-hand halo column)
#pragma unroll
        for (int c = 0; c <= WC; ++c) raw0[c] = (p.difference && c0 + c <= p.nc) ? T[c0 + c] : 0.f;
        for (int r = 0; r < p.nr; ++r) {
            float d[WC];
            if (p.difference) {
                const float* row1 = T + (long long)(r + 1) * p.ss;
                float raw1[WC + 1];
#pragma unroll
                for (int c = 0; c <= WC; ++c) raw1[c] = (c0 + c <= p.nc) ? row1[c0 + c] : 0.f;
#pragma unroll
                for (int c = 0; c < WC; ++c) d[c] = (c0 + c < p.nc) ? (raw1[c + 1] - raw1[c]) - (raw0[c + 1] - raw0[c]) : 0.f;
#pragma unroll
                for (int c = 0; c <= WC; ++c) raw0[c] = raw1[c];
            } else {
                const float* row0 = T + (long long)r * p.ss;
#pragma unroll
                for (int c = 0; c < WC; ++c) d[c] = (c0 + c < p.nc) ? row0[c0 + c] : 0.f;
            }
            float Rc[D][D][WC];
#pragma unroll
            for (int c = 0; c < WC; ++c) { Rc[0][0][c] = d[c]; K[0] += d[c]; }
#pragma unroll
            for (int m = 1; m < NLEV; ++m) {  // source level m (grid size dc) -> level m + 1 (grid size dn)
                constexpr int dummy = 0; (void)dummy;
                const int dc = m < D ? m : D, dn = (m + 1 < D) ? m + 1 : D;
                float Rn[D][D][WC];
                // [1][1]: Delta * AA_m, then AA_m += exclusive row prefix of the level-m totals
                float tot[WC], run = 0.f, pre[WC];
#pragma unroll
                for (int c = 0; c < WC; ++c) {
                    float t = 0.f;
#pragma unroll
                    for (int a = 0; a < D; ++a)
#pragma unroll
                        for (int b = 0; b < D; ++b)
                            if (a < dc && b < dc) t += Rc[a][b][c];
                    tot[c] = t;
                    pre[c] = run;
                    run += t;
                }
                {
                    const float off = ho_warp_excl_prefix(run, lane);
#pragma unroll
                    for (int c = 0; c < WC; ++c) {
                        Rn[0][0][c] = d[c] * AA[m - 1][c];
                        AA[m - 1][c] += off + pre[c];
                    }
                }
#pragma unroll
                for (int k = 2; k <= D; ++k) {
                    if (k > dn) continue;
                    // [1][k]: Delta * V_m^{k-1} / k, then V_m^{k-1} += sum_a R_m[a][k-1]
#pragma unroll
                    for (int c = 0; c < WC; ++c) {
                        Rn[0][k - 1][c] = d[c] * V[m - 1][k - 2][c] * (1.f / (float)k);
                        float cs = 0.f;
#pragma unroll
                        for (int a = 0; a < D; ++a)
                            if (a < dc) cs += Rc[a][k - 2][c];
                        V[m - 1][k - 2][c] += cs;
                    }
                    // [k][1]: Delta * H_m^{k-1} / k with H the exclusive row prefix of sum_b R_m[k-1][b]
                    float rs[WC], rrun = 0.f, rpre[WC];
#pragma unroll
                    for (int c = 0; c < WC; ++c) {
                        float t = 0.f;
#pragma unroll
                        for (int b = 0; b < D; ++b)
                            if (b < dc) t += Rc[k - 2][b][c];
                        rs[c] = t;
                        rpre[c] = rrun;
                        rrun += t;
                    }
                    const float roff = ho_warp_excl_prefix(rrun, lane);
#pragma unroll
                    for (int c = 0; c < WC; ++c) Rn[k - 1][0][c] = d[c] * (roff + rpre[c]) * (1.f / (float)k);
                    // [k][k']: Delta * R_m[k-1][k'-1] / (k k')
#pragma unroll
                    for (int k2 = 2; k2 <= D; ++k2) {
                        if (k2 > dn) continue;
#pragma unroll
                        for (int c = 0; c < WC; ++c) Rn[k - 1][k2 - 1][c] = d[c] * Rc[k - 2][k2 - 2][c] * (1.f / (float)(k * k2));
                    }
                }
#pragma unroll
                for (int a = 0; a < D; ++a)
#pragma unroll
                    for (int b = 0; b < D; ++b)
#pragma unroll
                        for (int c = 0; c < WC; ++c) {
                            Rc[a][b][c] = (a < dn && b < dn) ? Rn[a][b][c] : 0.f;
                            if (a < dn && b < dn) K[m] += Rn[a][b][c];
                        }
            }
        }
#pragma unroll
        for (int m = 0; m < NLEV; ++m)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) K[m] += __shfl_xor_sync(0xffffffffu, K[m], o);
        if (lane == 0) {
            float* o = p.out + (long long)(p.i_off + i) * p.ldo + p.j_off + j;
            o[0] = 1.f;
#pragma unroll
            for (int m = 0; m < NLEV; ++m) o[(long long)(m + 1) * p.out_level_stride] = K[m];
        }
    }
}

template <int NLEV, int D, int WC>
static int launch_ho_warp_inst(const HoParams& h, cudaStream_t st) {
    const long long npairs = (long long)h.n1 * h.n2;
    long long blocks = (npairs + 3) / 4;
    const long long cap = (long long)num_sms() * 8;
    sigkern_ho_warp_kernel<NLEV, D, WC><<<(int)(blocks < cap ? blocks : cap), 128, 0, st>>>(h);
    return check_launch();
}

template <int NLEV, int D>
static int launch_ho_warp_wc(const HoParams& h, cudaStream_t st) {
    const int wc = (h.nc + 31) / 32;
    if (wc <= 1) return launch_ho_warp_inst<NLEV, D, 1>(h, st);
    if constexpr (D * D * 2 <= 64) { if (wc <= 2) return launch_ho_warp_inst<NLEV, D, 2>(h, st); }
    if constexpr (D * D * 4 <= 64) { if (wc <= 4) return launch_ho_warp_inst<NLEV, D, 4>(h, st); }
    return GPSIG_E_UNSUPPORTED;
}

// (levels, order) pairs with order in [2, levels], levels <= 5
static int launch_ho_warp(const HoParams& h, cudaStream_t st) {
    switch (h.nlev * 10 + h.order) {
        case 22: return launch_ho_warp_wc<2, 2>(h, st);
        case 32: return launch_ho_warp_wc<3, 2>(h, st);
        case 33: return launch_ho_warp_wc<3, 3>(h, st);
        case 42: return launch_ho_warp_wc<4, 2>(h, st);
        case 43: return launch_ho_warp_wc<4, 3>(h, st);
        case 44: return launch_ho_warp_wc<4, 4>(h, st);
        case 52: return launch_ho_warp_wc<5, 2>(h, st);
        case 53: return launch_ho_warp_wc<5, 3>(h, st);
        case 54: return launch_ho_warp_wc<5, 4>(h, st);
        case 55: return launch_ho_warp_wc<5, 5>(h, st);
    }
    return GPSIG_E_UNSUPPORTED;
}

int launch_sigkern_ho(const float* M, int n1, int Lrows, int n2, int ncols, long long si, long long ss, long long sj,
                      int nlev, int order, int difference, int upper_only, int i_off, int j_off, long long ldo,
                      long long lvl_stride, float* out, cudaStream_t st) {
    if (!M || !out || n1 < 1 || n2 < 1 || Lrows < 1 || nlev < 1) return fail(GPSIG_E_BADARG, "sigkern_ho: bad sizes");
    if (nlev > kMaxLevelsHO) return fail(GPSIG_E_UNSUPPORTED, "higher-order path supports num_levels <= %d", kMaxLevelsHO);
    HoParams h;
    h.M = M; h.n1 = n1; h.n2 = n2; h.si = si; h.ss = ss; h.sj = sj;
    h.nlev = nlev; h.order = order; h.difference = difference ? 1 : 0; h.upper_only = upper_only ? 1 : 0;
    h.nr = difference ? Lrows - 1 : Lrows;
    h.nc = ncols;
    h.i_off = i_off; h.j_off = j_off; h.ldo = ldo;
    h.out = out; h.out_level_stride = lvl_stride;
    {
        ProfScope prof(GPSIG_PROF_RECURSION_OTHER, st, (double)n1 * n2);
        const int rc = launch_ho_warp(h, st);  // warp-per-pair kernel where instantiated
        if (rc != GPSIG_E_UNSUPPORTED) return rc;
    }
    const size_t per_pair = (size_t)(nlev > 1 ? nlev - 1 : 1) * order * (ncols > 0 ? ncols : 1) * sizeof(float);
    int ppb = 32;
    while (ppb > 1 && per_pair * ppb > 200 * 1024) ppb >>= 1;
    if (per_pair * ppb > 227 * 1024) return fail(GPSIG_E_UNSUPPORTED, "sequence too long for the higher-order kernel");
    const size_t smem = per_pair * ppb;
    cudaError_t e = cudaFuncSetAttribute(sigkern_ho_serial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const long long npairs = (long long)n1 * n2;
    long long blocks = (npairs + ppb - 1) / ppb;
    const long long cap = (long long)num_sms() * 8;
    ProfScope prof(GPSIG_PROF_RECURSION_OTHER, st, (double)npairs);
    sigkern_ho_serial_kernel<<<(int)(blocks < cap ? blocks : cap), ppb, smem, st>>>(h);
    return check_launch();
}

}  // namespace gpsig
