// seqkern.cu -- sequence-vs-sequence covariance pipeline: chunked increment-Gram producer -> TMA recursion kernel ->
// normalise / weight / sum epilogue.  Host orchestration of kernels.py:188-237 (_K_seq_diag, _K_seq) and :430-476.
#include <vector>

#include "internal.cuh"

namespace gpsig {

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// the stream geometry (consumer warps per CTA) depends on the number of levels: size workspaces for the larger one
static size_t stream_bytes_worst(long long nitems, int rows, int LP) {
    const size_t a = stream_bytes(stream_geometry(nitems, rows, LP, 5)), b = stream_bytes(stream_geometry(nitems, rows, LP, 8));
    return a > b ? a : b;
}

// ---- small helper kernels -----------------------------------------------------------------------------------------
__global__ void fill_levels_trivial_kernel(float* out, long long per_level, int nl) {
    // level 0 = 1, levels >= 1 = 0 (no increments at all: L == 1 with differencing)
    const long long total = per_level * nl;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x)
        out[idx] = idx < per_level ? 1.f : 0.f;
}

__global__ void mirror_upper_kernel(float* out, int n, int nl) {
    // out[m][i][j] = out[m][j][i] for i > j
    const long long per = (long long)n * n, total = per * nl;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long m = idx / per, r = idx - m * per;
        const int i = (int)(r / n), j = (int)(r - (long long)i * n);
        if (i > j) out[idx] = out[m * per + (long long)j * n + i];
    }
}

// multi-GPU assembly: rows arrive in gathered order with only j >= i valid.  32 x 32 tiles: a tile on or above the
// diagonal is copied from its row shard, a tile below it is the transpose of the mirrored tile (through shared memory,
// so both the reads and the writes are coalesced).
__global__ void assemble_symmetric_kernel(const float* __restrict__ rows, const int* __restrict__ row_src, int n,
                                          float* __restrict__ K) {
    __shared__ float tile[32][33];
    const int bi = blockIdx.y, bj = blockIdx.x;
    const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
    if (bj >= bi) {
        for (int r = ty; r < 32; r += 8) {
            const int i = bi * 32 + r, j = bj * 32 + tx;
            if (i < n && j < n) {
                const float up = rows[(long long)row_src[i] * n + j];
                K[(long long)i * n + j] = (j >= i) ? up : rows[(long long)row_src[j] * n + i];
            }
        }
    } else {
        for (int r = ty; r < 32; r += 8) {  // read the mirrored tile (bj, bi): rows j, columns i
            const int j = bj * 32 + r, i = bi * 32 + tx;
            tile[r][tx] = (j < n && i < n) ? rows[(long long)row_src[j] * n + i] : 0.f;
        }
        __syncthreads();
        for (int r = ty; r < 32; r += 8) {
            const int i = bi * 32 + r, j = bj * 32 + tx;
            if (i < n && j < n) K[(long long)i * n + j] = tile[tx][r];
        }
    }
}

// a7: normalise, weight, sum (kernels.py:430-433 / :455-469 / :471 / :473-476)
__global__ void normalize_weight_sum_kernel(const float* __restrict__ levels, int nl, long long n1, long long n2,
                                            const float* __restrict__ diag1, const float* __restrict__ diag2,
                                            const int* __restrict__ diag_cols, float jitter, int symmetric,
                                            const float* __restrict__ weights, float* __restrict__ levels_out,
                                            float* __restrict__ out_sum) {
    const long long per = n1 * n2;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < per; idx += (long long)gridDim.x * blockDim.x) {
        const long long i = idx / n2, j = idx - i * n2;
        float acc = 0.f;
        for (int m = 0; m < nl; ++m) {
            float v = levels[m * per + idx];
            if (symmetric) {
                const float di = levels[m * per + i * n2 + i] + jitter, dj = levels[m * per + j * n2 + j] + jitter;
                if (i == j) v += jitter;
                v = v / (sqrtf(di) * sqrtf(dj));
            } else if (diag_cols) {  // row shard of a symmetric problem: same arithmetic as the symmetric branch
                if (j == diag_cols[i]) v += jitter;
                v = v / (sqrtf(diag1[m * n1 + i] + jitter) * sqrtf(diag2[m * n2 + j] + jitter));
            } else {
                if (diag1) v = v / sqrtf(diag1[m * n1 + i] + jitter);
                if (diag2) v = v / sqrtf(diag2[m * n2 + j] + jitter);
            }
            v *= weights ? weights[m] : 1.f;
            if (levels_out) levels_out[m * per + idx] = v;
            acc += v;
        }
        if (out_sum) out_sum[idx] = acc;
    }
}

static int grid_for(long long total, int threads) {
    long long b = (total + threads - 1) / threads;
    const long long cap = (long long)num_sms() * 16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

// ---- shapes of one pipeline run -------------------------------------------------------------------------------------
struct SeqPlan {
    int DP;                 // floats per prepared point (padded state-space dimension, +4 for the augmented RBF form)
    int prep_mode;          // launch_prep_points mode
    bool fast_prod;         // packed-FMA producer (LINEAR with increments, RBF)
    bool lin_incr;          // linear + difference: produce <dx, dy> directly
    bool diff2d;            // producer differences the point Gram in both time directions
    int rowsA, rowsB;       // rows of the prepared A / B arrays per sequence
    int out_rows, ncols;    // increment tile size
    int LP, P, G;
    bool fast;              // stream-fed recursion eligible (P = 16 LP <= 512 columns)
    size_t bytesA, bytesB, bytesAn, bytesBn, bytesAnch, fixed;
};

static int make_plan(int kind, int L1, int L2, int d, int n1, int n2, bool symmetric, int difference, SeqPlan& pl) {
    pl.DP = (d + 3) / 4 * 4;
    pl.lin_incr = (kind == GPSIG_KERN_LINEAR) && difference;
    pl.diff2d = difference && !pl.lin_incr;
    pl.prep_mode = pl.lin_incr ? 1 : 0;
    pl.fast_prod = false;
    if (pl.lin_incr && pl.DP <= 16) pl.fast_prod = true;
    if (kind == GPSIG_KERN_RBF && difference && pl.DP <= 16) {
        pl.fast_prod = true;
        pl.prep_mode = 3;
    }
    pl.rowsA = pl.lin_incr ? L1 - 1 : L1;
    pl.rowsB = pl.lin_incr ? L2 - 1 : L2;
    pl.out_rows = difference ? L1 - 1 : L1;
    pl.ncols = difference ? L2 - 1 : L2;
    if (pl.ncols > 1024) return fail(GPSIG_E_UNSUPPORTED, "sequences longer than 1025 points are not supported yet");
    if (pl.ncols <= 512) {
        pl.LP = fo_lanes_per_pair(pl.ncols);
        pl.P = 16 * pl.LP;
        pl.G = 32 / pl.LP;
        pl.fast = true;
    } else {
        pl.LP = 32; pl.G = 1;
        pl.P = (pl.ncols + 3) / 4 * 4;
        pl.fast = false;
    }
    pl.bytesA = align_up((size_t)n1 * pl.rowsA * pl.DP * 4, 256);
    pl.bytesAn = align_up((size_t)n1 * pl.rowsA * 4, 256);
    pl.bytesB = symmetric ? 0 : align_up((size_t)n2 * pl.rowsB * pl.DP * 4, 256);
    pl.bytesBn = symmetric ? 0 : align_up((size_t)n2 * pl.rowsB * 4, 256);
    pl.bytesAnch = pl.prep_mode == 3 ? wf_anchor_bytes(symmetric ? n1 : n2, pl.rowsB, pl.DP) : 0;  // anchored column side
    pl.fixed = 256 /* jump flag */ + pl.bytesA + pl.bytesAn + pl.bytesB + pl.bytesBn + pl.bytesAnch + 1024;
    return GPSIG_OK;
}

}  // namespace gpsig

using namespace gpsig;

extern "C" size_t gpsig_seq_kern_workspace_bytes(int n1, int L1, int n2, int L2, int d, size_t budget_bytes) {
    if (n1 < 1 || n2 < 1 || L1 < 1 || L2 < 1 || d < 1) return 0;
    SeqPlan pl;
    // worst case over kinds/difference: points (L rows) -- a few MB, the chunk buffer dominates
    if (make_plan(GPSIG_KERN_RBF, L1, L2, d, n1, n2, false, 0, pl) != GPSIG_OK) return 0;
    pl.fixed += wf_anchor_bytes(n1 > n2 ? n1 : n2, L1 > L2 ? L1 : L2, pl.DP);
    size_t row_bytes = (size_t)pl.out_rows * n2 * pl.P * 4;  // one row block of i
    size_t all = row_bytes * (size_t)n1;
    if (pl.fast) {  // stream layout: whole streams of skewed rows
        const long long njg = (n2 + pl.G - 1) / pl.G;
        row_bytes = stream_bytes_worst(njg, pl.out_rows, pl.LP);
        all = stream_bytes_worst(njg * n1, pl.out_rows, pl.LP);
    }
    size_t chunk = budget_bytes > pl.fixed ? budget_bytes - pl.fixed : 0;
    if (chunk < row_bytes) chunk = row_bytes;
    if (chunk > all) chunk = all;
    return pl.fixed + align_up(chunk, 1024) + 1024;
}

// rows are given as a list of blocks [begin_k, end_k); block k is written to output rows out_row_k .. (compact layouts of
// GPU shards), every row at all n2 columns (symmetric: the entries j >= i only)
static int seq_kern_levels_impl(int kind, const float* params, const float* X, int n1, int L1, const float* X2, int n2, int L2,
                                int d, const float* inv_lengthscales, int num_levels, int order, int difference, int nblk,
                                const int* blk_begin, const int* blk_end, const long long* blk_out_row, float* out_levels,
                                long out_rows_total, int mirror, void* workspace, size_t workspace_bytes, cudaStream_t st) {
    const bool symmetric = (X2 == nullptr);
    if (symmetric) { n2 = n1; L2 = L1; }
    if (!X || !out_levels || !workspace || n1 < 1 || n2 < 1 || L1 < 1 || L2 < 1 || d < 1 || num_levels < 1)
        return fail(GPSIG_E_BADARG, "seq_kern_levels: bad arguments");
    if (order < 1 || order > num_levels) return fail(GPSIG_E_BADARG, "order must be in [1, num_levels]");
    if (nblk < 1 || !blk_begin || !blk_end || !blk_out_row) return fail(GPSIG_E_BADARG, "seq_kern_levels: no row blocks");
    long long rows_listed = 0;
    for (int k = 0; k < nblk; ++k) {
        if (blk_begin[k] < 0 || blk_end[k] > n1 || blk_begin[k] >= blk_end[k])
            return fail(GPSIG_E_BADARG, "bad row range [%d, %d)", blk_begin[k], blk_end[k]);
        if (blk_out_row[k] < 0 || blk_out_row[k] + (blk_end[k] - blk_begin[k]) > out_rows_total)
            return fail(GPSIG_E_BADARG, "output rows out of range");
        rows_listed += blk_end[k] - blk_begin[k];
    }
    const bool whole = nblk == 1 && blk_begin[0] == 0 && blk_end[0] == n1 && blk_out_row[0] == 0 && out_rows_total == n1;
    if (mirror && (!symmetric || !whole)) return fail(GPSIG_E_BADARG, "mirror needs the full symmetric problem");
    const int nl = num_levels + 1;
    const long long per_level = (long long)out_rows_total * n2;
    SeqPlan pl;
    int rc = make_plan(kind, L1, L2, d, n1, n2, symmetric, difference, pl);
    if (rc) return rc;
    if (pl.out_rows < 1 || pl.ncols < 1) {
        if (!whole) return fail(GPSIG_E_UNSUPPORTED, "row ranges need sequences with at least one increment");
        fill_levels_trivial_kernel<<<grid_for(per_level * nl, 256), 256, 0, st>>>(out_levels, per_level, nl);
        return check_launch();
    }
    if (((uintptr_t)workspace & 255u) != 0) return fail(GPSIG_E_ALIGN, "workspace must be 256-byte aligned");
    if (workspace_bytes < pl.fixed + (size_t)pl.out_rows * pl.P * 4 * (size_t)pl.G)
        return fail(GPSIG_E_WORKSPACE, "workspace smaller than one pair group");
    uint8_t* w = (uint8_t*)workspace;
    unsigned* flag = (unsigned*)w; w += 256;
    float* A = (float*)w; w += pl.bytesA;
    float* An = (float*)w; w += pl.bytesAn;
    float* B = A; float* Bn = An;
    if (!symmetric) { B = (float*)w; w += pl.bytesB; Bn = (float*)w; w += pl.bytesBn; }
    void* anch_buf = w; w += pl.bytesAnch;
    w = (uint8_t*)align_up((size_t)w, 1024);
    float* chunk = (float*)w;
    const size_t chunk_bytes = workspace_bytes - (size_t)(w - (uint8_t*)workspace);

    rc = launch_prep_points(X, n1, L1, d, inv_lengthscales, pl.prep_mode, pl.DP, A, An, st);
    if (rc) return rc;
    if (!symmetric) {
        rc = launch_prep_points(X2, n2, L2, d, inv_lengthscales, pl.prep_mode, pl.DP, B, Bn, st);
        if (rc) return rc;
    }
    const bool use_ho = order > 1;
    const bool upper = symmetric;
    const bool use_stream = pl.fast && !use_ho && num_levels <= 8;
    const bool rbf = kind == GPSIG_KERN_RBF;
    auto do_mirror = [&]() {
        if (!(mirror && n1 > 1)) return (int)GPSIG_OK;
        ProfScope prof(GPSIG_PROF_EPILOGUE, st, (double)per_level * nl);
        mirror_upper_kernel<<<grid_for(per_level * nl, 256), 256, 0, st>>>(out_levels, n1, nl);
        return check_launch();
    };
    // warp-fused path: every warp computes and consumes its own increment rows (no chunk buffer, no HBM intermediate);
    // ONE launch for all row blocks
    if (use_stream && pl.fast_prod && nblk <= kWfMaxRowBlocks && warpfused_supported(rbf, d, num_levels, pl.ncols, pl.rowsA)) {
        WfAnchored anch{};
        if (rbf) {
            rc = launch_wf_anchor_prep(B, n2, pl.rowsB, pl.DP, anch_buf, flag, &anch, st);
            if (rc) return rc;
        }
        rc = launch_sigkern_warpfused(rbf, A, B, rbf ? &anch : nullptr, rbf ? flag : nullptr, pl.rowsA, pl.rowsB, pl.DP,
                                      rbf ? pl.ncols + 1 : pl.ncols,
                                      n2, num_levels, upper ? 1 : 0, 0, nblk, blk_begin, blk_end, blk_out_row, n2, per_level,
                                      out_levels, st);
        if (rc != GPSIG_E_UNSUPPORTED) return rc ? rc : do_mirror();
    }
    for (int k = 0; k < nblk; ++k) {
        const int row_begin = blk_begin[k], row_end = blk_end[k];
        // rows are addressed by their GLOBAL index i: out[(i - row_begin + out_row0) * n2 + j]
        float* out_base = out_levels + ((long long)blk_out_row[k] - row_begin) * n2;
        int i0 = row_begin;
        while (i0 < row_end) {
            const int j_off = upper ? (i0 / pl.G) * pl.G : 0;
            const int nj = n2 - j_off;
            const int njg = (nj + pl.G - 1) / pl.G;
            long long ib = 0, nitems = 0;
            StreamGeom geom{};
            if (use_stream) {
                // largest row block whose stream buffer fits the workspace (items grow monotonically with the block)
                auto geom_for = [&](long long rows_i) {
                    return stream_geometry(items_before((int)rows_i, njg, pl.G, upper ? 1 : 0, i0, j_off), pl.out_rows, pl.LP, num_levels);
                };
                if (stream_bytes(geom_for(1)) > chunk_bytes)
                    return fail(GPSIG_E_WORKSPACE, "workspace too small for one row block: need %zu more bytes",
                                stream_bytes(geom_for(1)) - chunk_bytes);
                long long lo = 1, hi = row_end - i0;
                while (lo < hi) {
                    const long long mid = (lo + hi + 1) >> 1;
                    if (stream_bytes(geom_for(mid)) <= chunk_bytes) lo = mid; else hi = mid - 1;
                }
                ib = lo;
                nitems = items_before((int)ib, njg, pl.G, upper ? 1 : 0, i0, j_off);
                geom = geom_for(ib);
            } else {
                const size_t row_bytes = (size_t)pl.out_rows * nj * pl.P * 4;
                ib = (long long)(chunk_bytes / row_bytes);
                if (ib < 1) return fail(GPSIG_E_WORKSPACE, "workspace too small for one row block: need %zu more bytes",
                                        row_bytes - chunk_bytes);
                if (ib > row_end - i0) ib = row_end - i0;
            }
            ProdParams pp;
            pp.A = A; pp.B = B; pp.An = An; pp.Bn = Bn;
            pp.rowsA = pl.rowsA; pp.rowsB = pl.rowsB;
            pp.i0 = i0; pp.ni = (int)ib; pp.j0 = j_off; pp.nj = nj;
            pp.P = pl.P; pp.out_rows = pl.out_rows; pp.ncols = pl.ncols;
            pp.upper_only = upper ? 1 : 0; pp.G = pl.G; pp.diag = 0;
            pp.kp = make_kern_params(kind, params);
            pp.out = chunk;
            pp.stream = use_stream ? 1 : 0; pp.NW = geom.NW; pp.SR = geom.SR; pp.njg = njg;
            rc = pl.fast_prod ? launch_delta_producer_fast(rbf, pp, pl.DP, st) : launch_delta_producer(kind, pp, pl.DP, pl.diff2d, st);
            if (rc) return rc;
            const long long ss = (long long)nj * pl.P, sj = pl.P, si = (long long)pl.out_rows * ss;
            if (use_stream)
                rc = launch_sigkern_stream(chunk, geom, nitems, (int)ib, nj, pl.out_rows, pl.LP, num_levels, upper ? 1 : 0, i0, j_off,
                                           n2, per_level, out_base, st);
            else if (use_ho)
                rc = launch_sigkern_ho(chunk, (int)ib, pl.out_rows, nj, pl.ncols, si, ss, sj, num_levels, order, 0, upper ? 1 : 0, i0,
                                       j_off, n2, per_level, out_base, st);
            else
                rc = launch_sigkern_fo(chunk, (int)ib, pl.out_rows, nj, pl.ncols, pl.P, si, ss, sj, num_levels, 0, upper ? 1 : 0, i0,
                                       j_off, n2, per_level, out_base, st, 1);
            if (rc) return rc;
            i0 += (int)ib;
        }
    }
    return do_mirror();
}

// fixed part for ALL n sequences (prepared points + norms + flag) plus a chunk of up to `budget_bytes` (at least one
// diagonal pair group)
extern "C" size_t gpsig_seq_kern_diag_workspace_bytes(int n, int L, int d, size_t budget_bytes) {
    if (n < 1 || L < 1 || d < 1) return 0;
    SeqPlan pl;
    if (make_plan(GPSIG_KERN_RBF, L, L, d, n, n, true, 0, pl) != GPSIG_OK) return 0;
    pl.fixed += wf_anchor_bytes(n, L, pl.DP);
    size_t one = (size_t)pl.out_rows * pl.P * 4 * (size_t)pl.G, all = (size_t)pl.out_rows * pl.P * 4 * (size_t)n;
    if (pl.fast) {
        one = stream_bytes_worst(1, pl.out_rows, pl.LP);
        all = stream_bytes_worst(((long long)n + pl.G - 1) / pl.G, pl.out_rows, pl.LP);
    }
    size_t chunk = budget_bytes > pl.fixed ? budget_bytes - pl.fixed : 0;
    if (chunk < one) chunk = one;
    if (chunk > all) chunk = all;
    return pl.fixed + align_up(chunk, 1024) + 2048;
}

extern "C" int gpsig_seq_kern_levels(int kind, const float* params, const float* X, int n1, int L1, const float* X2, int n2,
                                     int L2, int d, const float* inv_lengthscales, int num_levels, int order, int difference,
                                     int row_begin, int row_end, float* out_levels, long out_row0, long out_rows_total,
                                     int mirror, void* workspace, size_t workspace_bytes, void* stream) {
    const long long orow = out_row0;
    return seq_kern_levels_impl(kind, params, X, n1, L1, X2, n2, L2, d, inv_lengthscales, num_levels, order, difference, 1,
                                &row_begin, &row_end, &orow, out_levels, out_rows_total, mirror, workspace, workspace_bytes,
                                (cudaStream_t)stream);
}

extern "C" int gpsig_seq_kern_levels_blocks(int kind, const float* params, const float* X, int n1, int L1, const float* X2,
                                            int n2, int L2, int d, const float* inv_lengthscales, int num_levels, int order,
                                            int difference, const int* row_blocks, int num_blocks, float* out_levels,
                                            long out_rows_total, void* workspace, size_t workspace_bytes, void* stream) {
    if (!row_blocks || num_blocks < 1) return fail(GPSIG_E_BADARG, "seq_kern_levels_blocks: no row blocks");
    std::vector<int> b(num_blocks), e(num_blocks);
    std::vector<long long> o(num_blocks);
    long long row = 0;
    for (int k = 0; k < num_blocks; ++k) {
        b[k] = row_blocks[2 * k]; e[k] = row_blocks[2 * k + 1];
        o[k] = row;
        row += e[k] - b[k];
    }
    return seq_kern_levels_impl(kind, params, X, n1, L1, X2, n2, L2, d, inv_lengthscales, num_levels, order, difference,
                                num_blocks, b.data(), e.data(), o.data(), out_levels, out_rows_total, 0, workspace, workspace_bytes,
                                (cudaStream_t)stream);
}

extern "C" int gpsig_seq_kern_diag_levels(int kind, const float* params, const float* X, int n, int L, int d,
                                          const float* inv_lengthscales, int num_levels, int order, int difference,
                                          float* out_levels, void* workspace, size_t workspace_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!X || !out_levels || !workspace || n < 1 || L < 1 || d < 1 || num_levels < 1)
        return fail(GPSIG_E_BADARG, "seq_kern_diag_levels: bad arguments");
    if (order < 1 || order > num_levels) return fail(GPSIG_E_BADARG, "order must be in [1, num_levels]");
    const int nl = num_levels + 1;
    SeqPlan pl;
    int rc = make_plan(kind, L, L, d, n, n, true, difference, pl);
    if (rc) return rc;
    if (pl.out_rows < 1 || pl.ncols < 1) {
        fill_levels_trivial_kernel<<<grid_for((long long)n * nl, 256), 256, 0, st>>>(out_levels, n, nl);
        return check_launch();
    }
    if (((uintptr_t)workspace & 255u) != 0) return fail(GPSIG_E_ALIGN, "workspace must be 256-byte aligned");
    uint8_t* w = (uint8_t*)workspace;
    unsigned* flag = (unsigned*)w; w += 256;
    float* A = (float*)w; w += pl.bytesA;
    float* An = (float*)w; w += pl.bytesAn;
    void* anch_buf = w; w += pl.bytesAnch;
    w = (uint8_t*)align_up((size_t)w, 1024);
    float* chunk = (float*)w;
    if (workspace_bytes < (size_t)(w - (uint8_t*)workspace))
        return fail(GPSIG_E_WORKSPACE, "workspace too small: gpsig_seq_kern_diag_workspace_bytes() gives the size");
    const size_t chunk_bytes = workspace_bytes - (size_t)(w - (uint8_t*)workspace);
    const bool use_stream = pl.fast && order == 1 && num_levels <= 8;
    const bool rbf = kind == GPSIG_KERN_RBF;
    // warp-fused path (diag mode: pairs (e, e)): no chunk buffer
    if (use_stream && pl.fast_prod && warpfused_supported(rbf, d, num_levels, pl.ncols, pl.rowsA)) {
        WfAnchored anch{};
        rc = launch_prep_points(X, n, L, d, inv_lengthscales, pl.prep_mode, pl.DP, A, An, st);
        if (!rc && rbf) rc = launch_wf_anchor_prep(A, n, pl.rowsA, pl.DP, anch_buf, flag, &anch, st);
        if (rc) return rc;
        const int b0 = 0, e0 = 1;
        const long long o0 = 0;
        rc = launch_sigkern_warpfused(rbf, A, A, rbf ? &anch : nullptr, rbf ? flag : nullptr, pl.rowsA, pl.rowsA, pl.DP,
                                      rbf ? pl.ncols + 1 : pl.ncols, n,
                                      num_levels, 0, 1, 1, &b0, &e0, &o0, n, n, out_levels, st);
        if (rc != GPSIG_E_UNSUPPORTED) return rc;
    }
    long long cap;
    if (use_stream) {
        // pairs per launch: whole groups of G, largest count whose stream buffer fits
        auto bytes_for = [&](long long pairs) { return stream_bytes(stream_geometry((pairs + pl.G - 1) / pl.G, pl.out_rows, pl.LP, num_levels)); };
        if (bytes_for(1) > chunk_bytes) return fail(GPSIG_E_WORKSPACE, "workspace too small for one diagonal tile");
        long long lo = 1, hi = n;
        while (lo < hi) {
            const long long mid = (lo + hi + 1) >> 1;
            if (bytes_for(mid) <= chunk_bytes) lo = mid; else hi = mid - 1;
        }
        cap = lo < n ? (lo / pl.G) * pl.G : lo;
        if (cap < 1) return fail(GPSIG_E_WORKSPACE, "workspace too small for one diagonal pair group");
    } else {
        const size_t pair_bytes = (size_t)pl.out_rows * pl.P * 4;
        cap = (long long)(chunk_bytes / pair_bytes);
        if (cap < 1) return fail(GPSIG_E_WORKSPACE, "workspace too small for one diagonal tile");
    }
    rc = launch_prep_points(X, n, L, d, inv_lengthscales, pl.prep_mode, pl.DP, A, An, st);
    if (rc) return rc;
    for (int e0 = 0; e0 < n; e0 += (int)cap) {
        const int ne = (int)((long long)(n - e0) < cap ? (n - e0) : cap);
        const long long nitems = (ne + pl.G - 1) / pl.G;
        const StreamGeom geom = stream_geometry(nitems, pl.out_rows, pl.LP, num_levels);
        ProdParams pp;
        pp.A = A; pp.B = A; pp.An = An; pp.Bn = An;
        pp.rowsA = pl.rowsA; pp.rowsB = pl.rowsB;
        pp.i0 = e0; pp.ni = ne; pp.j0 = e0; pp.nj = ne;
        pp.P = pl.P; pp.out_rows = pl.out_rows; pp.ncols = pl.ncols;
        pp.upper_only = 0; pp.G = pl.G; pp.diag = 1;
        pp.kp = make_kern_params(kind, params);
        pp.out = chunk;
        pp.stream = use_stream ? 1 : 0; pp.NW = geom.NW; pp.SR = geom.SR; pp.njg = (int)nitems;
        rc = pl.fast_prod ? launch_delta_producer_fast(kind == GPSIG_KERN_RBF, pp, pl.DP, st)
                          : launch_delta_producer(kind, pp, pl.DP, pl.diff2d, st);
        if (rc) return rc;
        // one "row" of pairs, pair e at column e
        const long long ss = (long long)ne * pl.P, sj = pl.P, si = (long long)pl.out_rows * ss;
        if (use_stream)
            rc = launch_sigkern_stream(chunk, geom, nitems, 1, ne, pl.out_rows, pl.LP, num_levels, 0, 0, e0, n, n, out_levels, st);
        else if (order > 1)
            rc = launch_sigkern_ho(chunk, 1, pl.out_rows, ne, pl.ncols, si, ss, sj, num_levels, order, 0, 0, 0, e0, n, n,
                                   out_levels, st);
        else
            rc = launch_sigkern_fo(chunk, 1, pl.out_rows, ne, pl.ncols, pl.P, si, ss, sj, num_levels, 0, 0, 0, e0, n, n,
                                   out_levels, st, 1);
        if (rc) return rc;
    }
    return GPSIG_OK;
}

extern "C" int gpsig_sigkern_levels(const float* M, int n1, int L1, int n2, int L2, long stride_i, long stride_s,
                                    long stride_j, int num_levels, int order, int difference, int upper_only,
                                    float* out_levels, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!M || !out_levels || n1 < 1 || n2 < 1 || L1 < 1 || L2 < 1 || num_levels < 1)
        return fail(GPSIG_E_BADARG, "sigkern_levels: bad arguments");
    if (order < 1 || order > num_levels) return fail(GPSIG_E_BADARG, "order must be in [1, num_levels]");
    if (upper_only && n1 != n2) return fail(GPSIG_E_BADARG, "upper_only needs n1 == n2");
    const long long per_level = (long long)n1 * n2;
    const int nl = num_levels + 1;
    const int ncols = difference ? L2 - 1 : L2, nrows = difference ? L1 - 1 : L1;
    if (ncols < 1 || nrows < 1) {
        fill_levels_trivial_kernel<<<grid_for(per_level * nl, 256), 256, 0, st>>>(out_levels, per_level, nl);
        return check_launch();
    }
    if (order > 1)
        return launch_sigkern_ho(M, n1, L1, n2, ncols, stride_i, stride_s, stride_j, num_levels, order, difference, upper_only,
                                 0, 0, n2, per_level, out_levels, st);
    return launch_sigkern_fo(M, n1, L1, n2, ncols, L2, stride_i, stride_s, stride_j, num_levels, difference, upper_only, 0, 0,
                             n2, per_level, out_levels, st, 0);
}

extern "C" int gpsig_assemble_symmetric(const float* rows, const int* row_src, int n, float* K, void* stream) {
    if (!rows || !row_src || !K || n < 1) return fail(GPSIG_E_BADARG, "assemble_symmetric: bad arguments");
    ProfScope prof(GPSIG_PROF_EPILOGUE, (cudaStream_t)stream, (double)n * n);
    const unsigned nb = (unsigned)((n + 31) / 32);
    assemble_symmetric_kernel<<<dim3(nb, nb), dim3(32, 8), 0, (cudaStream_t)stream>>>(rows, row_src, n, K);
    return check_launch();
}

extern "C" int gpsig_mirror_upper(float* levels, int nl, int n, void* stream) {
    if (!levels || nl < 1 || n < 1) return fail(GPSIG_E_BADARG, "mirror_upper: bad arguments");
    ProfScope prof(GPSIG_PROF_EPILOGUE, (cudaStream_t)stream, (double)n * n * nl);
    mirror_upper_kernel<<<grid_for((long long)n * n * nl, 256), 256, 0, (cudaStream_t)stream>>>(levels, n, nl);
    return check_launch();
}

extern "C" int gpsig_normalize_weight_sum(const float* levels, int nl, long n1, long n2, const float* diag1,
                                          const float* diag2, const int* diag_cols, float jitter, int symmetric,
                                          const float* weights, float* levels_out, float* out_sum, void* stream) {
    if (!levels || nl < 1 || n1 < 1 || n2 < 1) return fail(GPSIG_E_BADARG, "normalize_weight_sum: bad arguments");
    if (diag_cols && (!diag1 || !diag2)) return fail(GPSIG_E_BADARG, "diag_cols needs both diagonals");
    if (symmetric && n1 != n2) return fail(GPSIG_E_BADARG, "symmetric normalisation needs a square matrix");
    if (symmetric && levels_out == levels)
        return fail(GPSIG_E_BADARG, "symmetric normalisation reads the diagonal: levels_out must not alias levels");
    ProfScope prof(GPSIG_PROF_EPILOGUE, (cudaStream_t)stream, (double)n1 * n2);
    normalize_weight_sum_kernel<<<grid_for((long long)n1 * n2, 256), 256, 0, (cudaStream_t)stream>>>(
        levels, nl, n1, n2, diag1, diag2, diag_cols, jitter, symmetric, weights, levels_out, out_sum);
    return check_launch();
}
