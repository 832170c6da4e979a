// sigstream.cu -- the level-m signature recursion (signature_algs.py:8-35) over a chunk of the increment Gram tensor
// held in the consumer-ready STREAM layout that gram.cu writes (internal.cuh: StreamGeom).  This is the kernel of the
// K(X, X) / K(X, X2) / Kdiag pipeline; sigkern.cu keeps the tensor-map variant for caller-supplied Gram tensors.
//
// Why a second feed: with one 2 KB TMA box per Gram row the SM's TMA issue rate (about one bulk copy per 100 cycles,
// tools/ubench/tma_stream.cu) caps the stream at ~55 % of HBM bandwidth, and the row skew of the strip lanes keeps LP
// rows per warp resident in shared memory.  Here the producer kernel has already applied the skew and the 128-byte
// XOR swizzle in global memory, so
//   * a consumer warp's input is ONE contiguous byte stream: the TMA producer warp fetches it with 1-D bulk copies of
//     R = 4 skewed rows (8 KB) per instruction into a 2-stage ring per consumer (full/empty mbarriers); 11 consumer
//     warps + 1 producer warp per SM up to 5 levels (168 registers), 7 + 1 for 6..8 levels (255 registers);
//   * at step T all 32 lanes read skewed row T (conflict-free LDS.128), so a stage is dead after R steps and every
//     byte of shared memory is prefetch depth;
//   * the arithmetic is unchanged: lane l of a pair owns the 16-column strip l and keeps A_m[s, t] of all levels in
//     registers; rows are swept once; lane l works on row T - l, the running row prefix p_m moves to the next strip by
//     one shfl.up per level per row;  A_m[r+1, t] = A_m[r, t] + p_m ;  p_m += Delta[r, t] * A_{m-1}[r, t].
//   * row T + 1 is loaded into registers while row T is computed and the barrier of the next stage is probed one step
//     before it is needed, so neither LDS nor mbarrier latency sits on the critical path.
#include <stdlib.h>

#include "stream_consumer.cuh"

namespace gpsig {

constexpr int kStreamMaxLevels = 8;

struct StParams {
    const float* buf;
    long long SR;     // skewed rows per stream
    StreamItems it;
};

// 1-D bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

// MAXW = warps the register budget is sized for (registers are allocated in groups of 4 warps: 12 warps -> 168
// registers per thread, enough up to 5 levels; 8 warps -> 255 for 6..8 levels).  The launch may use fewer warps.
template <int NLEV, int MAXW>
__global__ void __launch_bounds__(MAXW * 32, 1) sigkern_fo_stream_kernel(const StParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int ncw = (blockDim.x >> 5) - 1;  // consumer warps; warp 0 is the bulk-copy producer
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = p.it.S, R = p.it.R;
    const uint32_t stage_bytes = (uint32_t)R * kRowBytes;
    const uint32_t smem0 = smem_u32(smem);
    const uint32_t full0 = smem0 + (uint32_t)ncw * S * stage_bytes;  // full[w][s]: bytes of the stage have landed
    const uint32_t empty0 = full0 + (uint32_t)ncw * S * 8;           // empty[w][s]: the consumer is done with the stage
    if (threadIdx.x == 0) {
        for (int k = 0; k < ncw * S; ++k) {
            mbar_init(full0 + 8 * k, 1);
            mbar_init(empty0 + 8 * k, 1);
        }
        fence_mbar_init();
    }
    __syncthreads();  // the only CTA-wide barrier

    const int my = warp == 0 ? lane : warp - 1;  // ring / stream this thread works for
    if (warp == 0 && lane >= ncw) return;
    const long long wg = (long long)blockIdx.x * ncw + my;
    const uint32_t ring = smem0 + (uint32_t)my * S * stage_bytes;
    const uint32_t fb = full0 + (uint32_t)my * S * 8, eb = empty0 + (uint32_t)my * S * 8;

    if (warp == 0) {
        // ===== producer: lane w streams the bytes of consumer w =====================================================
        const long long nloc = wg >= p.it.nitems ? 0 : (p.it.nitems - wg + p.it.NW - 1) / p.it.NW;
        const long long nsteps = nloc * p.it.Lin + p.it.LP - 1;  // skewed rows of the stream
        if (nloc == 0) return;
        const uint8_t* src = reinterpret_cast<const uint8_t*>(p.buf) + (size_t)wg * (size_t)p.SR * kRowBytes;
        const long long nstages = (nsteps + R - 1) / R;
        int stage = 0, round = 0;
        for (long long k = 0; k < nstages; ++k) {
            if (round > 0) mbar_wait(eb + 8 * stage, (uint32_t)(round + 1) & 1u);
            const long long left = nsteps - k * R;
            const uint32_t bytes = (uint32_t)(left < R ? left : R) * kRowBytes;
            const uint32_t bar = fb + 8 * stage;
            mbar_arrive_expect_tx(bar, bytes);
            bulk_load_1d(ring + stage * stage_bytes, src + (size_t)k * stage_bytes, bytes, bar);
            if (++stage == S) { stage = 0; ++round; }
        }
        return;
    }
    run_stream_consumer<NLEV>(p.it, ring, fb, eb, wg, lane);
}

// ---------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------
static int ilog2_exact(int x) {
    int l = 0;
    while ((1 << l) < x) ++l;
    return l;
}

// geometry of a launch over `nitems` items of `rows` increment rows each
StreamGeom stream_geometry(long long nitems, int rows, int LP, int nlev) {
    StreamGeom g;
    // consumer warps per CTA / rows per bulk copy / ring depth (GPSIG_STREAM_* are tuning knobs for experiments)
    const int maxw = nlev <= 5 ? 12 : 8;
    const EnvKnobs& ek = env_knobs();
    g.ncw = nlev <= 5 ? (ek.stream_ncw > 0 ? ek.stream_ncw : 11) : 7;
    if (g.ncw > maxw - 1) g.ncw = maxw - 1;
    if (g.ncw < 1) g.ncw = 1;
    g.R = ek.stream_r > 0 ? ek.stream_r : 4;
    g.S = ek.stream_s > 0 ? ek.stream_s : (nlev <= 5 ? 2 : 3);
    if (g.R < 1) g.R = 1;
    if (g.S < 2) g.S = 2;
    while ((size_t)g.ncw * g.S * ((size_t)g.R * 2048 + 16) > 232448 && g.S > 2) --g.S;
    while ((size_t)g.ncw * g.S * ((size_t)g.R * 2048 + 16) > 232448 && g.R > 1) --g.R;
    long long want = (nitems + g.ncw - 1) / g.ncw;
    g.grid = (int)(want < num_sms() ? want : num_sms());
    if (g.grid < 1) g.grid = 1;
    g.NW = g.grid * g.ncw;
    const long long per = (nitems + g.NW - 1) / g.NW;   // items of the longest stream
    const long long sr = per * rows + LP - 1;
    g.SR = (sr + g.R - 1) / g.R * g.R;
    return g;
}

template <int NLEV>
static int launch_stream_inst(const StParams& p, const StreamGeom& g, size_t smem, cudaStream_t st) {
    constexpr int MAXW = NLEV <= 5 ? 12 : 8;
    if (g.ncw + 1 > MAXW) return fail(GPSIG_E_BADARG, "stream geometry has too many warps for %d levels", NLEV);
    auto kern = sigkern_fo_stream_kernel<NLEV, MAXW>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    kern<<<g.grid, (g.ncw + 1) * 32, smem, st>>>(p);
    return check_launch();
}

int launch_sigkern_stream(const float* buf, const StreamGeom& g, long long nitems, int n1, int n2, int rows, int LP, int nlev,
                          int upper_only, int i_off, int j_off, long long ldo, long long lvl_stride, float* out,
                          cudaStream_t st) {
    if (!buf || !out || nitems < 1 || rows < 1 || nlev < 1) return fail(GPSIG_E_BADARG, "sigkern_stream: bad sizes");
    if (nlev > kStreamMaxLevels) return fail(GPSIG_E_UNSUPPORTED, "stream path supports num_levels <= %d", kStreamMaxLevels);
    if ((uintptr_t)buf & 127u) return fail(GPSIG_E_ALIGN, "stream buffer must be 128-byte aligned");
    StParams sp;
    sp.buf = buf; sp.SR = g.SR;
    StreamItems& p = sp.it;
    p.nitems = nitems; p.NW = g.NW; p.R = g.R; p.S = g.S;
    p.Lin = rows; p.LP = LP; p.log2LP = ilog2_exact(LP); p.G = 32 / LP;
    p.njg = (n2 + p.G - 1) / p.G;
    p.n1 = n1; p.n2 = n2;
    p.upper_only = upper_only ? 1 : 0; p.i_off = i_off; p.j_off = j_off;
    p.ldo = ldo; p.out = out; p.out_level_stride = lvl_stride;
    const size_t smem = (size_t)g.ncw * g.S * ((size_t)g.R * kRowBytes + 16);
    ProfScope prof(GPSIG_PROF_RECURSION, st, (double)nitems * p.G);
    switch (nlev) {
        case 1: return launch_stream_inst<1>(sp, g, smem, st);
        case 2: return launch_stream_inst<2>(sp, g, smem, st);
        case 3: return launch_stream_inst<3>(sp, g, smem, st);
        case 4: return launch_stream_inst<4>(sp, g, smem, st);
        case 5: return launch_stream_inst<5>(sp, g, smem, st);
        case 6: return launch_stream_inst<6>(sp, g, smem, st);
        case 7: return launch_stream_inst<7>(sp, g, smem, st);
        case 8: return launch_stream_inst<8>(sp, g, smem, st);
    }
    return fail(GPSIG_E_UNSUPPORTED, "stream path supports num_levels <= %d", kStreamMaxLevels);
}

}  // namespace gpsig
