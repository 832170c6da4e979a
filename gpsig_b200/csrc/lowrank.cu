// lowrank.cu -- device side of the reference's low-rank mode (gpsig/low_rank_calculations.py, signature_algs.py:162-222).
//
// The reference compresses the outer product of two feature vectors (k1 * k2 entries) to `r` components with a very
// sparse random projection R (density 1/s), and does it as gather -> multiply -> DENSE matmul against the non-zero rows
// of R (low_rank_calculations.py:182-193).  Here the projection arrives in compressed-sparse-column form and every
// output component only touches its own non-zeros: per row  r * nnz_per_column  FMAs instead of  nnz_rows * r.
// The coordinate-subsampling variant (:104-127) is the same kernel with one non-zero (a Rademacher sign) per column.
//   lr_hadamard_csc_kernel : C[x, c] = scale * sum_{e in col c} A[x, ia[e]] * B[x, ib[e]] * val[e]
//   lr_seq_level_kernel    : one level of signature_algs.py:182-188 for one sequence per block, the exclusive cumsum
//                            over time (:183) fused: B is the running prefix of the previous level's factor.
#include "internal.cuh"

namespace gpsig {

__global__ void lr_hadamard_csc_kernel(const float* __restrict__ A, long long rows, int k1, const float* __restrict__ B, int k2,
                                       const int* __restrict__ colptr, const int* __restrict__ ia, const int* __restrict__ ib,
                                       const float* __restrict__ val, int r, float scale, float* __restrict__ out, int rb) {
    extern __shared__ float sh[];  // [rb][k1] then [rb][k2]
    float* sa = sh;
    float* sb = sh + (size_t)rb * k1;
    for (long long x0 = (long long)blockIdx.x * rb; x0 < rows; x0 += (long long)gridDim.x * rb) {
        const int nr = (int)((rows - x0) < rb ? (rows - x0) : rb);
        __syncthreads();
        for (int e = threadIdx.x; e < nr * k1; e += blockDim.x) sa[e] = A[x0 * k1 + e];
        for (int e = threadIdx.x; e < nr * k2; e += blockDim.x) sb[e] = B[x0 * k2 + e];
        __syncthreads();
        for (int idx = threadIdx.x; idx < nr * r; idx += blockDim.x) {
            const int xl = idx / r, c = idx - xl * r;
            const float* a = sa + xl * k1;
            const float* b = sb + xl * k2;
            float acc = 0.f;
            for (int e = colptr[c]; e < colptr[c + 1]; ++e) acc = fmaf(a[ia[e]] * b[ib[e]], val[e], acc);
            out[(x0 + xl) * r + c] = scale * acc;
        }
    }
}

// grid: one block per sequence.  U (n, Lr, k1), P_in (n, Lr, k2) -> P_out (n, Lr, r), phi (n, r) = sum_t P_out.
__global__ void lr_seq_level_kernel(const float* __restrict__ U, const float* __restrict__ P_in, int Lr, int k1, int k2,
                                    const int* __restrict__ colptr, const int* __restrict__ ia, const int* __restrict__ ib,
                                    const float* __restrict__ val, int r, float scale, float* __restrict__ P_out,
                                    float* __restrict__ phi) {
    extern __shared__ float sh[];  // u[k1], q[k2]
    float* su = sh;
    float* sq = sh + k1;
    const long long n = blockIdx.x;
    const float* Un = U + n * (long long)Lr * k1;
    const float* Pn = P_in + n * (long long)Lr * k2;
    float* On = P_out + n * (long long)Lr * r;
    for (int e = threadIdx.x; e < k2; e += blockDim.x) sq[e] = 0.f;
    float phis[4] = {0.f, 0.f, 0.f, 0.f};  // up to 4 output components per thread
    for (int t = 0; t < Lr; ++t) {
        __syncthreads();
        for (int e = threadIdx.x; e < k1; e += blockDim.x) su[e] = Un[(long long)t * k1 + e];
        __syncthreads();
        int slot = 0;
        for (int c = threadIdx.x; c < r; c += blockDim.x, ++slot) {
            float acc = 0.f;
            for (int e = colptr[c]; e < colptr[c + 1]; ++e) acc = fmaf(su[ia[e]] * sq[ib[e]], val[e], acc);
            acc *= scale;
            On[(long long)t * r + c] = acc;
            if (slot < 4) phis[slot] += acc;
        }
        __syncthreads();
        for (int e = threadIdx.x; e < k2; e += blockDim.x) sq[e] += Pn[(long long)t * k2 + e];  // exclusive prefix for t + 1
    }
    int slot = 0;
    for (int c = threadIdx.x; c < r; c += blockDim.x, ++slot)
        if (slot < 4) phi[n * r + c] = phis[slot];
}

}  // namespace gpsig

using namespace gpsig;

extern "C" int gpsig_lr_hadamard_csc(const float* A, long rows, int k1, const float* B, int k2, const int* colptr, const int* ia,
                                     const int* ib, const float* val, int r, float scale, float* out, void* stream) {
    if (!A || !B || !colptr || !out || rows < 0 || k1 < 1 || k2 < 1 || r < 1) return fail(GPSIG_E_BADARG, "lr_hadamard_csc: bad arguments");
    if (rows == 0) return GPSIG_OK;
    int rb = 8;
    while (rb > 1 && (size_t)rb * (k1 + k2) * sizeof(float) > 160 * 1024) rb >>= 1;
    const size_t smem = (size_t)rb * (k1 + k2) * sizeof(float);
    if (smem > 227 * 1024) return fail(GPSIG_E_UNSUPPORTED, "feature dimensions too large for lr_hadamard_csc (%d + %d)", k1, k2);
    if (smem > 48 * 1024) cudaFuncSetAttribute(lr_hadamard_csc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    long long blocks = (rows + rb - 1) / rb;
    const long long cap = (long long)num_sms() * 8;
    ProfScope prof(GPSIG_PROF_TENS, (cudaStream_t)stream, (double)rows);
    lr_hadamard_csc_kernel<<<(int)(blocks < cap ? blocks : cap), 256, smem, (cudaStream_t)stream>>>(A, rows, k1, B, k2, colptr, ia, ib,
                                                                                                    val, r, scale, out, rb);
    return check_launch();
}

extern "C" int gpsig_lr_seq_level(const float* U, const float* P_in, long n, int Lr, int k1, int k2, const int* colptr,
                                  const int* ia, const int* ib, const float* val, int r, float scale, float* P_out, float* phi,
                                  void* stream) {
    if (!U || !P_in || !colptr || !P_out || !phi || n < 0 || Lr < 1 || k1 < 1 || k2 < 1 || r < 1)
        return fail(GPSIG_E_BADARG, "lr_seq_level: bad arguments");
    if (P_in == P_out) return fail(GPSIG_E_BADARG, "lr_seq_level: P_out must not alias P_in");
    if (n == 0) return GPSIG_OK;
    int threads = 32;
    while (threads < r && threads < 256) threads <<= 1;
    if (r > 4 * threads) return fail(GPSIG_E_UNSUPPORTED, "rank_bound > 1024 is not supported by lr_seq_level");
    const size_t smem = (size_t)(k1 + k2) * sizeof(float);
    if (smem > 227 * 1024) return fail(GPSIG_E_UNSUPPORTED, "feature dimensions too large for lr_seq_level");
    if (smem > 48 * 1024) cudaFuncSetAttribute(lr_seq_level_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    ProfScope prof(GPSIG_PROF_TENS, (cudaStream_t)stream, (double)n);
    lr_seq_level_kernel<<<(unsigned)n, threads, smem, (cudaStream_t)stream>>>(U, P_in, Lr, k1, k2, colptr, ia, ib, val, r, scale, P_out,
                                                                             phi);
    return check_launch();
}
