// internal.cuh -- structures and launchers shared between translation units (not part of the C ABI).
#pragma once
#include "common.cuh"

namespace gpsig {

// tuning / experiment knobs from the environment, read ONCE (first use) -- see tools/README.md
struct EnvKnobs {
    int warpfused;        // GPSIG_WARPFUSED (default 1): 0 = two-kernel pipeline instead of the warp-fused kernel
    int warpfused_warps;  // GPSIG_WARPFUSED_WARPS (default 0 = as many as fit)
    int stream_ncw, stream_r, stream_s;  // GPSIG_STREAM_NCW / _R / _S (0 = default geometry)
    int tens_tc;          // GPSIG_TENS_TC (default 1): 0 = CUDA-core Kuf kernel even where the tcgen05 one applies
};
const EnvKnobs& env_knobs();

constexpr int kSpecMaxQ = 8, kSpecMaxD = 16;
struct KernParams {
    float a, b;  // poly: gamma, degree; mix: mixing
    // spectral (kernels.py:894-942): Q mixture components, per-feature frequencies omega and inverse scales gamma
    int Q, fam;  // fam 0: exp(-|g d|^2 / 2), 1: exp(-|g d| / 2)
    float alpha[kSpecMaxQ];
    float omega[kSpecMaxQ * kSpecMaxD];
    float gamma[kSpecMaxQ * kSpecMaxD];
};

#if defined(__CUDACC__)
// sum_q alpha_q * env(|gamma_q * (x - y)|) * cos(2 pi <omega_q, x - y>)   (kernels.py:921-942)
template <typename XA, typename YA>
__device__ __forceinline__ float spectral_eval(const XA& x, const YA& y, int n, const KernParams& kp) {
    float acc = 0.f;
    for (int q = 0; q < kp.Q; ++q) {
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int c = 0; c < kSpecMaxD; ++c) {
            if (c < n) {
                const float df = x[c] - y[c];
                const float g = df * kp.gamma[q * kSpecMaxD + c];
                s1 = fmaf(g, g, s1);
                s2 = fmaf(df, kp.omega[q * kSpecMaxD + c], s2);
            }
        }
        const float env = kp.fam == 1 ? __expf(-0.5f * sqrtf(s1)) : __expf(-0.5f * s1);
        acc = fmaf(kp.alpha[q] * env, cospif(2.f * s2), acc);
    }
    return acc;
}
#endif

// ---- work-item enumeration shared by the chunk producer and the recursion kernels ---------------------------------
// An item is a group of G neighbouring pairs (i, jg*G .. jg*G+G-1).  Items are numbered row by row.
// sum_{x < i} floor(x / G)
__host__ __device__ inline long long tri_floor(long long i, int G) {
    long long b = i / G, r = i % G;
    return (long long)G * b * (b - 1) / 2 + r * b;
}
// number of items in local rows < il.  With upper_only, local row il keeps the groups whose GLOBAL group index is
// >= floor((i_off + il) / G); the chunk starts at global group j_off / G.
__host__ __device__ inline long long items_before(int il, int njg, int G, int upper_only, int i_off, int j_off) {
    if (!upper_only) return (long long)il * njg;
    long long skipped = tri_floor((long long)i_off + il, G) - tri_floor(i_off, G) - (long long)il * (j_off / G);
    return (long long)il * njg - skipped;
}

// the chunk's local group index of the first group row il keeps
__host__ __device__ inline int first_group(int il, int G, int upper_only, int i_off, int j_off) {
    return upper_only ? (i_off + il) / G - j_off / G : 0;
}

// ---- consumer-ready stream layout of the increment-Gram chunk (written by gram.cu, read by sigstream.cu) -----------
// Item u belongs to stream u % NW (one stream per consumer warp of the recursion launch) at position u / NW.  A stream
// is a sequence of 2 KB "skewed rows": skewed row n*rows + s + l holds, for each of the G pairs of item n, the 16-column
// strip l of increment row s -- so at step T every lane of the consumer warp reads the SAME skewed row T, and rows are
// contiguous in memory (large 1-D bulk copies).  Inside a skewed row the 16-byte chunks are XOR-swizzled over 128-byte
// lines exactly like CU_TENSOR_MAP_SWIZZLE_128B, which makes the strip-wise LDS.128 reads bank-conflict free.
constexpr int kSkewRowFloats = 512;  // 2048 bytes
struct StreamGeom {
    int NW;         // streams (consumer warps in the grid)
    long long SR;   // skewed rows per stream (pitch)
    int R;          // skewed rows per bulk copy / smem stage
    int S;          // stages per consumer ring
    int ncw;        // consumer warps per CTA
    int grid;
};
StreamGeom stream_geometry(long long nitems, int rows, int LP, int nlev);
inline size_t stream_bytes(const StreamGeom& g) { return (size_t)g.NW * (size_t)g.SR * 2048; }
__host__ __device__ inline uint32_t swizzle_in_row(uint32_t byte_in_row) {
    return ((byte_in_row >> 7) << 7) | ((((byte_in_row >> 4) & 7u) ^ ((byte_in_row >> 7) & 7u)) << 4) | (byte_in_row & 15u);
}

// chunk producer arguments (gram.cu)
struct ProdParams {
    const float* A;   // (n1, rowsA, DP) points or increments of X   (rows i)
    const float* B;   // (n2, rowsB, DP) points or increments of X2  (cols j)
    const float* An;  // squared norms (n1, rowsA) or NULL
    const float* Bn;
    int rowsA, rowsB;  // points per sequence in A / B
    int i0, ni;        // row block [i0, i0 + ni)
    int j0, nj;        // column block [j0, j0 + nj)   (buffer has nj pairs per row)
    int P;             // column pitch of the chunk buffer
    int out_rows;      // rows per pair in the buffer (rowsA - 1 if diff2d else rowsA)
    int ncols;         // valid increment columns (rowsB - 1 if diff2d else rowsB)
    int upper_only;    // skip pair groups that lie entirely below the diagonal (global j < global i)
    int G;             // pair-group size of the consumer (a group is kept if its last j >= i)
    int diag;          // 1: buffer is [rows][n][P] holding only pairs (i, i), i in [i0, i0+ni)
    KernParams kp;
    float* out;
    // stream layout (stream != 0): see StreamGeom
    int stream, NW, njg;
    long long SR;
};

int launch_delta_producer(int kind, const ProdParams& p, int DP, bool diff2d, cudaStream_t st);
// fast producers (LINEAR on increments, RBF on augmented points); DPA = floats per prepared point
int launch_delta_producer_fast(bool rbf, const ProdParams& p, int DPA, cudaStream_t st);
// mode 0: scaled points, 1: scaled time increments, 3: points scaled so that log2 k_rbf(x, y) = -|x' - y'|^2
int launch_prep_points(const float* X, long long n, int L, int d, const float* inv_ls, int mode, int DP, float* out,
                       float* norms, cudaStream_t st);
KernParams make_kern_params(int kind, const float* params);

int fo_lanes_per_pair(int ncols);
// First-order recursion over M[n1, Lrows, n2, pitch] (element strides si, ss, sj; unit stride along t).
// Output: out[m * lvl_stride + (i_off + i) * ldo + j_off + j]; upper_only compares GLOBAL indices.
int launch_sigkern_fo(const float* M, int n1, int Lrows, int n2, int ncols, int pitch, long long si, long long ss,
                      long long sj, int nlev, int difference, int upper_only, int i_off, int j_off, long long ldo,
                      long long lvl_stride, float* out, cudaStream_t st, int force_generic);
// First-order recursion over a chunk in stream layout (sigstream.cu); items enumerate (i, jg) of an n1 x n2 pair block.
int launch_sigkern_stream(const float* buf, const StreamGeom& g, long long nitems, int n1, int n2, int rows, int LP, int nlev,
                          int upper_only, int i_off, int j_off, long long ldo, long long lvl_stride, float* out,
                          cudaStream_t st);
// Warp-autonomous fused Gram + recursion (warpfused.cu): every warp computes and consumes its own increment rows.
constexpr int kWfMaxRowBlocks = 32;
// RBF strips whose points lie further than this (squared, in the scaled units of prep mode 3: log2 k = -|x - y|^2) from
// the strip's anchor switch the call to the direct-difference instantiation (fp32 error of the anchored form is about
// 1.5e-7 |u|^2 relative to k)
constexpr float kWfJumpThreshold = 4.0f;
bool warpfused_supported(bool rbf, int d, int nlev, int ncols, int rowsA);
struct WfAnchored { float* Bu; float* Bnu; int rows_padded; };  // column side of the anchored RBF form (warpfused.cu)
size_t wf_anchor_bytes(long long n, int rows, int D);
int launch_wf_anchor_prep(const float* B, long long n, int rows, int D, void* buf, unsigned* flag, WfAnchored* out,
                          cudaStream_t st);
int launch_sigkern_warpfused(bool rbf, const float* A, const float* B, const WfAnchored* anch, const unsigned* flag, int rowsA,
                             int rowsB, int D,
                             int npts, int n2, int nlev, int upper_only, int diag, int nblk, const int* blk_begin,
                             const int* blk_end, const long long* blk_out_row, long long ldo, long long lvl_stride, float* out,
                             cudaStream_t st);
// tcgen05 Kuf (tens_tc.cu): split-TF32 Gram on the tensor cores; applies when the scaled data lies within kTcRadius2 of
// its mean (decided on the device: `flag` holds the float bits of max |x - c|^2).
constexpr float kTcRadius2 = 16.0f;
bool tens_tc_supported(int kind, int d, int nlev, int order, int increments, int difference, int L);
int launch_tens_seq_tc(const float* Z, long long nz, const float* X, long long n, int L, int d, const float* inv_ls, int nlev,
                       float* out, unsigned* flag, cudaStream_t st);
// Higher-order recursion (signature_algs.py:37-74), same addressing.
int launch_sigkern_ho(const float* M, int n1, int Lrows, int n2, int ncols, long long si, long long ss, long long sj,
                      int nlev, int order, int difference, int upper_only, int i_off, int j_off, long long ldo,
                      long long lvl_stride, float* out, cudaStream_t st);

}  // namespace gpsig
