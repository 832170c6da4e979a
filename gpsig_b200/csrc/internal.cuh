// internal.cuh -- structures and launchers shared between translation units (not part of the C ABI).
#pragma once
#include "common.cuh"

namespace gpsig {

struct KernParams {
    float a, b;  // poly: gamma, degree; mix: mixing
};

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
};

int launch_delta_producer(int kind, const ProdParams& p, int DP, bool diff2d, cudaStream_t st);
int launch_prep_points(const float* X, long long n, int L, int d, const float* inv_ls, int increments, int DP, float* out,
                       float* norms, cudaStream_t st);
KernParams make_kern_params(int kind, const float* params);

int fo_lanes_per_pair(int ncols);
// First-order recursion over M[n1, Lrows, n2, pitch] (element strides si, ss, sj; unit stride along t).
// Output: out[m * lvl_stride + (i_off + i) * ldo + j_off + j]; upper_only compares GLOBAL indices.
int launch_sigkern_fo(const float* M, int n1, int Lrows, int n2, int ncols, int pitch, long long si, long long ss,
                      long long sj, int nlev, int difference, int upper_only, int i_off, int j_off, long long ldo,
                      long long lvl_stride, float* out, cudaStream_t st, int force_generic);
// Higher-order recursion (signature_algs.py:37-74), same addressing.
int launch_sigkern_ho(const float* M, int n1, int Lrows, int n2, int ncols, long long si, long long ss, long long sj,
                      int nlev, int order, int difference, int upper_only, int i_off, int j_off, long long ldo,
                      long long lvl_stride, float* out, cudaStream_t st);

}  // namespace gpsig
