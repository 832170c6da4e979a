// tens_tc.cu -- Kuf (kernels.py:313-340 + signature_algs.py:101-127) for SignatureRBF with the static-kernel Gram on the
// 5th-generation tensor cores (tcgen05.mma kind::tf32, accumulators in TMEM, operands staged by TMA).
//
// The Gram kappa(z, x_t) of an inducing-tensor point against every time step of every sequence is > 95 % of the arithmetic
// of Kuf and it IS a contraction (K = d): exactly what the north star keeps tensor cores for.  In scaled coordinates
// (log2 k = -|z - x|^2) the exponent is 2<z, x> - |z|^2 - |x|^2 -- one dot product of augmented vectors.  Plain TF32
// (10-bit mantissa) is useless for an exponent, so every factor is split into two TF32 pieces and the norms into three,
// all laid out along K:
//     A row (tensor point): [ 2z_hi | 2z_hi | 2z_lo | 1 1 1 | nz_1 nz_2 nz_3 ]       nz = -|z|^2
//     B row (time step)   : [  x_hi |  x_lo |  x_hi | nx_1 nx_2 nx_3 | 1 1 1 ]       nx = -|x|^2
// (3d + 6 K-slots, padded to a multiple of 32: one or two 128-byte swizzle atoms), accumulated in fp32 by the MMA:
// error ~ 3 * 2^-22 * |z| |x| in the exponent.  That is small only when the data sits within a few lengthscales of
// the centre the coordinates are taken from (the mean of the tensor points): the prep kernel leaves max |x - c|^2 in a flag word, this
// kernel returns at once when it exceeds kTcRadius2, and the CUDA-core kernel of tens.cu (anchored differences, good
// anywhere) returns at once when it does not -- both are launched, no host synchronisation.
// d = 9, 10 (3d + 6 = 33, 36 slots: one too many atoms) use the ZOUT layout instead: nx in TWO pieces and the z norm out
// of the MMA, applied as a per-row factor in the epilogue (v = 2^D1 * 2^-|z1|^2 - 2^D0 * 2^-|z0|^2): 3d + 2 <= 32 slots.
//
// One CTA per SM, persistent over work items (tile of 128 rows x chunk of sequences):
//   * rows: the (tensor, component) pairs.  A level's components form a CHAIN (r_p[t] = h_p[t] * sum_{t'<t} r_{p-1}[t']);
//     chains are packed into warps of 32 TMEM lanes without ever splitting one (row map built on the host), lane = row.
//   * warps 8 and 9, one elected lane each: the producers, ONE PER WARP SET (a single producer serving both sets in program
//     order made them run in lockstep).  TMA loads (A tiles once per item by producer 0, B tiles of 64 time steps through
//     the set's own ring) and the MMAs -- per tile 2 x (K/8) tcgen05.mma (M=128, N=64): D0 = exponents against the z^0
//     points, D1 against the z^1 points, side by side in TMEM (128 columns per buffer, 2 buffers for each of the 2 warp
//     sets = all 512 columns).
//   * warps 0-7: two sets of 4 (one warp per TMEM lane quarter); set s takes the sequences n = s (mod 2) of the chunk.  The
//     blocks of NB time steps of a set's sequences form one stream.  At step b every thread has its row's 2 x NB
//     exponents of block b in registers (tcgen05.ld 32x32b, double buffered: block b + 2 is requested when block b has
//     been consumed), v = 2^D1 - 2^D0, h = time increment of v, and parks h in a per-warp
//     shared-memory FIFO.  The chain runs SKEWED: the lane at chain position p works on block b - 1 - p (its h from the
//     FIFO, the exclusive prefix c_in of its predecessor from a double-buffered patch the predecessor filled one step
//     earlier): r = h * c_in, running prefix, hand the prefix on.  One round per block instead of one per level, and the
//     chain never waits for the exponentials of the same step.  TMEM addressing is warp-uniform in the column, which is
//     why the skew lives in shared memory.  The prefix carries over tiles; at the end of a sequence the last lane of each
//     chain owns K_m(z, x_n).
// MUFU.EX2 is the bound by design (2 per Gram entry; everything else is ~7 issue slots per entry).
#include <type_traits>
#include <vector>

#include "internal.cuh"

namespace gpsig {

constexpr int kTcRows = 128;     // rows (TMEM lanes) per tile
constexpr int kTcNT = 64;        // time steps per MMA tile
constexpr int kTcThreads = 320;  // 8 epilogue warps (2 sets x 4 TMEM lane quarters) + 2 producer warps

struct TcRow { int z, m, p, k; };  // tensor, level (0 = padding row), position in the chain, component index

struct TcParams {
    const TcRow* rows;       // (ntiles * 128)
    const unsigned* flag;    // float bits of max |x - c|^2 (scaled)
    int ntiles, nch, chunk;  // row tiles, sequence chunks, sequences per chunk
    long long nz, n;
    int L, Lp;               // sequence length, padded sequence length (multiple of 64)
    float* out;              // (NLEV + 1, nz, n)
    const float* zscale;     // ZOUT layout: 2^(-|z|^2) per A row, z^0 rows then z^1 rows (2 x ntiles x 128)
};

// ---- PTX wrappers --------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_ld(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tc_ld(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float tc_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// shared-memory matrix descriptor: K-major, 128-byte swizzle, 8-row groups 1024 bytes apart (cute::UMMA::SmemDescriptor:
// start address >> 4 in [0, 14), LBO in [16, 30) (= 1, unused for swizzled K-major), SBO in [32, 46), version 1 in
// [46, 48), layout type SWIZZLE_128B = 2 in [61, 64))
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t addr) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 (1 << 4), A = B = TF32 (2 << 7, 2 << 10), both K-major,
// N >> 3 in [17, 23), M >> 4 in [24, 29)
constexpr uint32_t kTcIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kTcNT >> 3) << 17) | ((uint32_t)(kTcRows >> 4) << 24);

// ---- the kernel ----------------------------------------------------------------------------------------------------
// KA = 128-byte K atoms per operand row (1: K = 32 slots, d <= 8, and d = 9, 10 in the ZOUT layout;  2: K = 64, d <= 19);
// S = stages of each B ring;
// NB = time steps per epilogue block (16, or 8 where shared memory is short)
// register fence: the values of a completed tcgen05.ld may only be read after tcgen05.wait::ld; the empty volatile asm
// keeps its place after the (volatile) wait and every use of v depends on it
template <int N>
__device__ __forceinline__ void tc_reg_fence(float (&v)[N]) {
#pragma unroll
    for (int i = 0; i < N; i += 8)
        asm volatile("" : "+f"(v[i]), "+f"(v[i + 1]), "+f"(v[i + 2]), "+f"(v[i + 3]), "+f"(v[i + 4]), "+f"(v[i + 5]), "+f"(v[i + 6]),
                          "+f"(v[i + 7]));
}

template <int NLEV, int KA, int S, int NB, bool ZOUT>
__global__ void __launch_bounds__(kTcThreads, 1)
tens_seq_tc_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapZ0,
                   const __grid_constant__ CUtensorMap mapZ1, const TcParams p) {
    if (__uint_as_float(*p.flag) > kTcRadius2) return;  // data too spread for the split-TF32 exponent: tens.cu does the call
    extern __shared__ __align__(1024) uint8_t tsm_raw[];
    constexpr uint32_t kABytes = KA * kTcRows * 128, kBBytes = KA * kTcNT * 128;
    // layout: A0 | A1 | B ring of set 0 | B ring of set 1 | FIFOs and prefix patches of the 8 epilogue warps | barriers | tmem
    // base   (1024-byte aligned: the 128-byte swizzle of TMA and of the MMA descriptors is a function of the absolute address)
    uint8_t* tsm = tsm_raw + ((1024u - (smem_u32(tsm_raw) & 1023u)) & 1023u);
    const uint32_t smem0 = smem_u32(tsm);
    const uint32_t sA0 = smem0, sA1 = sA0 + kABytes, sB = sA1 + kABytes;
    // per epilogue warp: FIFO of h blocks [F][32 lanes][NB] and the prefix patch [2][33 rows][NB] (row 32 = ones); 16-byte
    // chunks of a lane row are XOR-swizzled with the lane index, which makes the per-lane 16-byte accesses conflict free
    constexpr int F = NLEV + 1;
    constexpr uint32_t kWarpFloats = (F * 32 + 2 * 33) * NB;
    float* xpatch = reinterpret_cast<float*>(tsm + 2 * kABytes + 2 * S * kBBytes);
    constexpr uint32_t kPatchBytes = 8 * kWarpFloats * 4;
    const uint32_t bars = smem0 + 2 * kABytes + 2 * S * kBBytes + kPatchBytes;
    // b_full[set][S] | b_empty[set][S] | t_full[4] | t_empty[4] | a_full | mma_done[2]
    const uint32_t b_full = bars, b_empty = bars + 16 * S, t_full = bars + 32 * S, t_empty = t_full + 32, a_full = t_empty + 32,
                   mma_done = a_full + 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tsm + 2 * kABytes + 2 * S * kBBytes + kPatchBytes + 32 * S + 96);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2 * S; ++i) { mbar_init(b_full + 8 * i, 1); mbar_init(b_empty + 8 * i, 1); }
        for (int i = 0; i < 4; ++i) { mbar_init(t_full + 8 * i, 1); mbar_init(t_empty + 8 * i, 4); }
        mbar_init(a_full, 1);
        mbar_init(mma_done, 1);
        mbar_init(mma_done + 8, 1);
        fence_mbar_init();
    }
    if (warp == 8) {  // TMEM: all 512 columns (one CTA per SM)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (warp < 8) {  // the "ones" rows of the warp's prefix patch: what the first lane of a chain multiplies by
        float* pb = xpatch + warp * kWarpFloats + F * 32 * NB;
        if (lane < NB) { pb[32 * NB + lane] = 1.f; pb[33 * NB + 32 * NB + lane] = 1.f; }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int tiles_per_seq = p.Lp / kTcNT;
    const long long nitems = (long long)p.ntiles * p.nch;

    if (warp >= 8) {
        // ===== two producers (TMA + MMA issue, one lane each): warp 8 feeds warp set 0, warp 9 feeds set 1.  Each owns a ring
        // of B stages and the two TMEM buffers of its set, so a set that runs ahead never waits for the other; the A tiles
        // of an item are shared (loaded by producer 0 once BOTH producers' MMAs of the previous item have drained).
        // (plain 32-bit counters and incremental iterators: every 64-bit division in this loop would cost more than the
        //  MMAs of a tile)
        const int set = warp - 8;
        if (lane == 0) {
            tma_prefetch_desc(&mapX);
            if (set == 0) { tma_prefetch_desc(&mapZ0); tma_prefetch_desc(&mapZ1); }
            const uint32_t bf = b_full + 8 * S * set, be = b_empty + 8 * S * set, sBr = sB + set * S * kBBytes;
            int ld_stage = 0, mma_stage = 0;
            uint32_t ld_phase = 0, mma_phase = 0;   // ring phases (flip when the stage index wraps)
            uint32_t cnt = 0;                       // tiles issued for this set (TMEM buffer = cnt & 1)
            uint32_t item_no = 0;
            for (long long item = blockIdx.x; item < nitems; item += gridDim.x, ++item_no) {
                const int zt = (int)(item / p.nch), ch = (int)(item - (long long)zt * p.nch);
                const int n0 = ch * p.chunk;
                const int n1 = (long long)n0 + p.chunk < p.n ? n0 + p.chunk : (int)p.n;
                const int nseq_set = n0 + set < n1 ? (n1 - n0 - set + 1) / 2 : 0;   // sequences n0 + set, n0 + set + 2, ...
                const int ntile = nseq_set * tiles_per_seq;
                if (set == 0) {
                    if (item_no > 0) {
                        mbar_wait(mma_done, (item_no - 1) & 1u);
                        mbar_wait(mma_done + 8, (item_no - 1) & 1u);
                    }
                    mbar_arrive_expect_tx(a_full, 2 * kABytes);
#pragma unroll
                    for (int a = 0; a < KA; ++a) {
                        tma_load_2d(sA0 + a * kTcRows * 128, &mapZ0, a_full, 32 * a, zt * kTcRows);
                        tma_load_2d(sA1 + a * kTcRows * 128, &mapZ1, a_full, 32 * a, zt * kTcRows);
                    }
                }
                int l_na = n0 + set, l_tt = 0;   // next tile to load
                auto issue_load = [&]() {
                    mbar_wait(be + 8 * ld_stage, ld_phase ^ 1u);
                    mbar_arrive_expect_tx(bf + 8 * ld_stage, kBBytes);
                    const int row0 = l_na * p.Lp + l_tt * kTcNT;
#pragma unroll
                    for (int a = 0; a < KA; ++a)
                        tma_load_2d(sBr + ld_stage * kBBytes + a * kTcNT * 128, &mapX, bf + 8 * ld_stage, 32 * a, row0);
                    if (++ld_stage == S) { ld_stage = 0; ld_phase ^= 1u; }
                    if (++l_tt == tiles_per_seq) { l_tt = 0; l_na += 2; }
                };
                int loaded = 0;
                for (; loaded < ntile && loaded < S - 1; ++loaded) issue_load();
                // EVERY producer waits for the item's A tiles, also one without tiles of its own: the wait is what keeps it
                // from running an item ahead of producer 0 (where the parity of the A barrier would alias the item before)
                mbar_wait(a_full, item_no & 1u);
                for (int q = 0; q < ntile; ++q) {
                    if (loaded < ntile) { issue_load(); ++loaded; }
                    const uint32_t buf = cnt & 1u;
                    // the epilogue has drained the buffer (two tiles of slack: back off instead of spinning on the issue port)
                    while (!mbar_try_wait(t_empty + 8 * (set * 2 + buf), ((cnt >> 1) & 1u) ^ 1u)) __nanosleep(256);
                    mbar_wait(bf + 8 * mma_stage, mma_phase);                          // the time steps have landed
                    tc_fence_after();
                    const uint32_t d0 = tmem_base + (uint32_t)((set * 2 + buf) * 128), d1 = d0 + kTcNT;
                    const uint32_t sBs = sBr + mma_stage * kBBytes;
#pragma unroll
                    for (int ks = 0; ks < 4 * KA; ++ks)
                        tc_mma_tf32(d0, tc_smem_desc(sA0 + (ks >> 2) * kTcRows * 128 + (ks & 3) * 32),
                                    tc_smem_desc(sBs + (ks >> 2) * kTcNT * 128 + (ks & 3) * 32), kTcIdesc, ks > 0);
#pragma unroll
                    for (int ks = 0; ks < 4 * KA; ++ks)
                        tc_mma_tf32(d1, tc_smem_desc(sA1 + (ks >> 2) * kTcRows * 128 + (ks & 3) * 32),
                                    tc_smem_desc(sBs + (ks >> 2) * kTcNT * 128 + (ks & 3) * 32), kTcIdesc, ks > 0);
                    tc_commit(be + 8 * mma_stage);                  // the ring stage is free once these MMAs have read it
                    tc_commit(t_full + 8 * (set * 2 + buf));        // ... and the accumulators are complete
                    if (++mma_stage == S) { mma_stage = 0; mma_phase ^= 1u; }
                    ++cnt;
                }
                if (ntile > 0) tc_commit(mma_done + 8 * set);       // this set's reads of the A tiles are over
                else mbar_arrive(mma_done + 8 * set);
            }
        }
    } else {
        // ===== epilogue: warp set `set`, TMEM lane quarter `quarter` =================================================
        const int set = warp >> 2, quarter = warp & 3;
        float* fifo = xpatch + warp * kWarpFloats;     // [F][32][NB]
        float* patch = fifo + F * 32 * NB;             // [2][33][NB]
        const uint32_t tmem_set = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(set * 256);
        uint32_t tile = 0;  // tiles this set has opened so far (TMEM buffer = tile & 1)
        const long long per = p.nz * p.n;
        constexpr int kSB = kTcNT / NB;  // blocks per tile (even: the register buffer of a block is its parity)
        constexpr int C4 = NB / 4;       // 16-byte chunks per lane row
        static_assert(kSB % 2 == 0 && kSB >= 4, "two register buffers alternate over the blocks of a tile");
        const int swz = (lane >> 1) & (C4 - 1);
        for (long long item = blockIdx.x; item < nitems; item += gridDim.x) {
            const int zt = (int)(item / p.nch), ch = (int)(item - (long long)zt * p.nch);
            const int n0 = ch * p.chunk;
            const int n1 = (long long)n0 + p.chunk < p.n ? n0 + p.chunk : (int)p.n;
            const int nseq_set = n0 + set < n1 ? (n1 - n0 - set + 1) / 2 : 0;
            if (nseq_set == 0) continue;
            const TcRow row = p.rows[(long long)zt * kTcRows + quarter * 32 + lane];
            // ZOUT: the z-side norm is not in the MMA (no room for its slots): kappa = 2^D * 2^(-|z|^2); padding rows carry 0
            float zs0 = 1.f, zs1 = 1.f;
            if (ZOUT) {
                const long long ri = (long long)zt * kTcRows + quarter * 32 + lane;
                zs0 = p.zscale[ri];
                zs1 = p.zscale[(long long)p.ntiles * kTcRows + ri];
            }
            const bool pad = row.m == 0;
            const bool last = !pad && row.p == row.m - 1;
            const int lag = pad ? 1 : row.p + 1;                   // this lane works on block (step - lag)
            const int src = (pad || row.p == 0) ? 32 : lane - 1;    // whose prefix it multiplies by (32 = the ones row)
            const int src_swz = src == 32 ? 0 : ((src >> 1) & (C4 - 1));
            // blocks of the last tile that hold real time steps (rounded up to a pair: the register buffer of a block is its
            // parity); the padding behind them repeats the last point -- zero increments -- and is skipped
            int last_nb = (((p.L - (tiles_per_seq - 1) * kTcNT + NB - 1) / NB) + 1) & ~1;
            if (last_nb > kSB) last_nb = kSB;
            const int per_seq = (tiles_per_seq - 1) * kSB + last_nb;
            const int tiles_item = nseq_set * tiles_per_seq;
            float a[2][2][NB];         // [register buffer][z^0 / z^1 accumulator][time step]
            float vprev = 0.f, carry = 0.f;
            int bis = -lag, sq = 0;    // block within the sequence / sequence ordinal of THIS lane's (lagged) block
            int wslot = 0, par = 0;    // FIFO slot written this step, patch buffer written this step
            auto open_tile = [&](uint32_t t) {
                mbar_wait(t_full + 8 * (set * 2 + (t & 1u)), (t >> 1) & 1u);
                tc_fence_after();
            };
            auto load_block = [&](uint32_t t, int sb, float (&dst)[2][NB]) {
                const uint32_t d0 = tmem_set + (t & 1u) * 128 + sb * NB;
                tc_ld(d0, dst[0]);
                tc_ld(d0 + kTcNT, dst[1]);
            };
            // one step of the skewed chain; HAVE: a new block (in registers `blk`) enters the FIFO this step
            auto step = [&](auto have_c, float (&blk)[2][NB], bool first_block) {
                constexpr bool HAVE = decltype(have_c)::value;
                const bool active = bis >= 0 && sq < nseq_set;
                int rslot = wslot - lag;
                if (rslot < 0) rslot += F;
                float h[NB], cin[NB];
                {
                    const float4* h4 = reinterpret_cast<const float4*>(fifo + (rslot * 32 + lane) * NB);
                    const float4* c4 = reinterpret_cast<const float4*>(patch + ((par ^ 1) * 33 + src) * NB);
#pragma unroll
                    for (int q = 0; q < C4; ++q) {
                        const float4 hv = h4[q ^ swz], cv = c4[q ^ src_swz];
                        h[4 * q] = hv.x; h[4 * q + 1] = hv.y; h[4 * q + 2] = hv.z; h[4 * q + 3] = hv.w;
                        cin[4 * q] = cv.x; cin[4 * q + 1] = cv.y; cin[4 * q + 2] = cv.z; cin[4 * q + 3] = cv.w;
                    }
                }
                if constexpr (HAVE) {
                    // the increments h of the newest block go into the FIFO (padding rows of A are zero rows: both
                    // accumulators are equal there, so their h is exactly 0)
                    float hnew[NB];
#pragma unroll
                    for (int t = 0; t < NB; ++t) {
                        const float v = ZOUT ? fmaf(tc_ex2(blk[1][t]), zs1, -(tc_ex2(blk[0][t]) * zs0))   // kernels.py:330
                                             : tc_ex2(blk[1][t]) - tc_ex2(blk[0][t]);
                        if (t == 0) vprev = first_block ? v : vprev;            // first time step of a sequence: no increment yet
                        hnew[t] = v - vprev;                                    // signature_algs.py:114
                        vprev = v;
                    }
                    float4* w4 = reinterpret_cast<float4*>(fifo + (wslot * 32 + lane) * NB);
#pragma unroll
                    for (int c = 0; c < C4; ++c) w4[c ^ swz] = make_float4(hnew[4 * c], hnew[4 * c + 1], hnew[4 * c + 2], hnew[4 * c + 3]);
                }
                float run = bis == 0 ? 0.f : carry;
                float4* w4 = reinterpret_cast<float4*>(patch + (par * 33 + lane) * NB);
#pragma unroll
                for (int q = 0; q < C4; ++q) {
                    float4 cv;                                                   // exclusive prefix (signature_algs.py:123)
                    cv.x = run; run = fmaf(h[4 * q], cin[4 * q], run);
                    cv.y = run; run = fmaf(h[4 * q + 1], cin[4 * q + 1], run);
                    cv.z = run; run = fmaf(h[4 * q + 2], cin[4 * q + 2], run);
                    cv.w = run; run = fmaf(h[4 * q + 3], cin[4 * q + 3], run);
                    w4[q ^ swz] = cv;
                }
                carry = active ? run : carry;
                if (active && bis == per_seq - 1 && last) {                      // end of a sequence
                    const long long idx = (long long)row.z * p.n + (n0 + set + 2 * sq);
                    p.out[(long long)row.m * per + idx] = carry;               // signature_algs.py:125
                    if (row.m == 1) p.out[idx] = 1.f;
                }
                if (++bis == per_seq) { bis = 0; ++sq; }
                if (++wslot == F) wslot = 0;
                par ^= 1;
                __syncwarp();
            };
            // the first two blocks of the item
            open_tile(tile);
            load_block(tile, 0, a[0]);
            load_block(tile, 1, a[1]);
            tc_wait_ld();
            tc_reg_fence(a[0][0]); tc_reg_fence(a[0][1]); tc_reg_fence(a[1][0]); tc_reg_fence(a[1][1]);
            int tl = 0;   // tile of the item
            // FULL: every tile uses all its blocks (the sequence length is within a pair of blocks of a multiple of 64) --
            // the loop over blocks has compile-time bounds; otherwise the last tile of a sequence stops early
            auto run_tiles = [&](auto full_c) {
                constexpr bool FULL = decltype(full_c)::value;
                for (int s = 0; s < nseq_set; ++s)
                    for (int tt = 0; tt < tiles_per_seq; ++tt, ++tile, ++tl) {
                        const int nb = FULL ? kSB : (tt == tiles_per_seq - 1 ? last_nb : kSB);   // blocks of this tile (even)
#pragma unroll
                        for (int sb = 0; sb < kSB; ++sb) {
                            if (!FULL && sb >= nb) break;
                            step(std::true_type{}, a[sb & 1], tt == 0 && sb == 0);
                            // block (sb + 1) -- loaded a step ago -- is complete after this wait; block (sb + 2) takes the
                            // registers this step has just consumed
                            tc_wait_ld();
                            tc_reg_fence(a[(sb + 1) & 1][0]); tc_reg_fence(a[(sb + 1) & 1][1]);
                            if (sb == nb - 2) {    // every block of this tile that is used has left TMEM: the buffer is free
                                tc_fence_before();
                                __syncwarp();
                                if (lane == 0) mbar_arrive(t_empty + 8 * (set * 2 + (tile & 1u)));
                            }
                            if (sb + 2 < nb) {
                                load_block(tile, sb + 2, a[sb & 1]);
                            } else if (tl + 1 < tiles_item) {   // sb = nb - 2 (even) / nb - 1 (odd): blocks 0 / 1 of the next tile
                                if ((sb & 1) == 0) open_tile(tile + 1);
                                load_block(tile + 1, sb & 1, a[sb & 1]);
                            }
                        }
                    }
            };
            if (last_nb == kSB) run_tiles(std::true_type{});
            else run_tiles(std::false_type{});
            // drain: the deepest chain position consumes the last block NLEV steps after it entered
#pragma unroll 1
            for (int dr = 0; dr < NLEV; ++dr) step(std::false_type{}, a[0], false);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ---- operand preparation --------------------------------------------------------------------------------------------
__device__ __forceinline__ float tf32_rn(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ void tf32_split3(float v, float& p1, float& p2, float& p3) {
    p1 = tf32_rn(v);
    const float r = v - p1;
    p2 = tf32_rn(r);
    p3 = tf32_rn(r - p2);
}

// column sums of the scaled points (for the centre): sums[c] = sum over rows of X[r, c] * inv_ls[c].  ONE block, fixed
// summation order: the centre must be bit-identical from call to call (sharded == single-GPU, reproducible runs)
__global__ void __launch_bounds__(256) tc_centre_kernel(const float* __restrict__ X, long long rows, int d,
                                                         const float* __restrict__ inv_ls, float* __restrict__ sums) {
    __shared__ float part[256];
    for (int c = 0; c < d; ++c) {
        float acc = 0.f;
        for (long long r = threadIdx.x; r < rows; r += 256) acc += X[r * d + c] * (inv_ls ? inv_ls[c] : 1.f);
        part[threadIdx.x] = acc;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) {
            if ((int)threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
            __syncthreads();
        }
        if (threadIdx.x == 0) sums[c] = part[0];
        __syncthreads();
    }
}

// B operand: one row per (sequence, padded time step); rows past the end of a sequence repeat its last point
__global__ void tc_prep_x_kernel(const float* __restrict__ X, long long n, int L, int Lp, int d, const float* __restrict__ inv_ls,
                                 const float* __restrict__ sums, float inv_rows, int KW, int zout, float* __restrict__ out,
                                 unsigned* __restrict__ flag) {
    const float rs = 0.8493218002880191f;  // sqrt(log2(e) / 2): log2 k = -|x' - z'|^2
    float worst = 0.f;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n * Lp; idx += (long long)gridDim.x * blockDim.x) {
        const long long seq = idx / Lp;
        int t = (int)(idx - seq * Lp);
        if (t > L - 1) t = L - 1;
        const float* x0 = X + (seq * L + t) * d;
        float* o = out + idx * KW;
        float nn = 0.f;
        for (int c = 0; c < d; ++c) {
            const float s = inv_ls ? inv_ls[c] : 1.f;
            const float v = (x0[c] * s - sums[c] * inv_rows) * rs;
            nn = fmaf(v, v, nn);
            const float hi = tf32_rn(v), lo = tf32_rn(v - hi);
            o[c] = hi; o[d + c] = lo; o[2 * d + c] = hi;
        }
        float p1, p2, p3;
        tf32_split3(-nn, p1, p2, p3);
        if (zout) {  // 3 d + 2 slots: -|x|^2 in two pieces (2^-22 |x|^2 left over), the z norm is applied in the epilogue
            o[3 * d] = p1; o[3 * d + 1] = p2;
            for (int c = 3 * d + 2; c < KW; ++c) o[c] = 0.f;
        } else {
            o[3 * d] = p1; o[3 * d + 1] = p2; o[3 * d + 2] = p3;
            o[3 * d + 3] = 1.f; o[3 * d + 4] = 1.f; o[3 * d + 5] = 1.f;
            for (int c = 3 * d + 6; c < KW; ++c) o[c] = 0.f;
        }
        worst = fmaxf(worst, nn);
    }
    for (int o2 = 16; o2 > 0; o2 >>= 1) worst = fmaxf(worst, __shfl_xor_sync(0xffffffffu, worst, o2));
    if ((threadIdx.x & 31) == 0 && worst > 0.f) atomicMax(flag, __float_as_uint(worst));
}

// A operands: row r of the row map -> the z^0 (w = 0) and z^1 (w = 1) points of its (tensor, component); padding rows are zero
__global__ void tc_prep_z_kernel(const float* __restrict__ Z, long long nz, int d, const float* __restrict__ inv_ls,
                                 const float* __restrict__ sums, float inv_rows, const TcRow* __restrict__ rows, long long nrows,
                                 int KW, float* __restrict__ zscale, float* __restrict__ out0, float* __restrict__ out1) {
    const float rs = 0.8493218002880191f;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < 2 * nrows; idx += (long long)gridDim.x * blockDim.x) {
        const long long r = idx >> 1;
        const int w = (int)(idx & 1);
        float* o = (w ? out1 : out0) + r * KW;
        const TcRow row = rows[r];
        if (row.m == 0) {
            for (int c = 0; c < KW; ++c) o[c] = 0.f;
            if (zscale) zscale[(long long)w * nrows + r] = 0.f;
            continue;
        }
        const float* z0 = Z + (((long long)row.k * nz + row.z) * 2 + w) * d;
        float nn = 0.f;
        for (int c = 0; c < d; ++c) {
            const float s = inv_ls ? inv_ls[c] : 1.f;
            const float v = (z0[c] * s - sums[c] * inv_rows) * rs;
            nn = fmaf(v, v, nn);
            const float v2 = v + v;
            const float hi = tf32_rn(v2), lo = tf32_rn(v2 - hi);
            o[c] = hi; o[d + c] = hi; o[2 * d + c] = lo;
        }
        if (zscale) {
            o[3 * d] = 1.f; o[3 * d + 1] = 1.f;
            for (int c = 3 * d + 2; c < KW; ++c) o[c] = 0.f;
            zscale[(long long)w * nrows + r] = exp2f(-nn);
        } else {
            float p1, p2, p3;
            tf32_split3(-nn, p1, p2, p3);
            o[3 * d] = 1.f; o[3 * d + 1] = 1.f; o[3 * d + 2] = 1.f;
            o[3 * d + 3] = p1; o[3 * d + 4] = p2; o[3 * d + 5] = p3;
            for (int c = 3 * d + 6; c < KW; ++c) o[c] = 0.f;
        }
    }
}

// ---- host ----------------------------------------------------------------------------------------------------------
// chains (tensor, level) packed into warps of 32 rows, never split
static void tc_row_map(long long nz, int nlev, std::vector<TcRow>& rows) {
    rows.clear();
    int used = 0;  // rows used in the current warp
    auto pad_to_warp = [&]() { while (used % 32) { rows.push_back(TcRow{0, 0, 0, 0}); ++used; } };
    for (long long z = 0; z < nz; ++z)
        for (int m = nlev; m >= 1; --m) {
            if ((used % 32) + m > 32) pad_to_warp();
            const int k0 = m * (m - 1) / 2;
            for (int q = 0; q < m; ++q) { rows.push_back(TcRow{(int)z, m, q, k0 + q}); ++used; }
        }
    while (rows.size() % kTcRows) rows.push_back(TcRow{0, 0, 0, 0});
}

bool tens_tc_supported(int kind, int d, int nlev, int order, int increments, int difference, int L) {
    return kind == GPSIG_KERN_RBF && order == 1 && increments && difference && nlev >= 1 && nlev <= 6 && 3 * d + 6 <= 64 && L >= 2 &&
           env_knobs().tens_tc != 0;
}

template <int NLEV, int KA, bool ZOUT>
static int launch_tc_inst(const CUtensorMap& mx, const CUtensorMap& mz0, const CUtensorMap& mz1, const TcParams& p, cudaStream_t st) {
    constexpr int S = KA == 1 ? 3 : 2;   // stages of EACH set's B ring
    constexpr int NB = KA == 1 ? 16 : 8;
    auto kern = tens_seq_tc_kernel<NLEV, KA, S, NB, ZOUT>;
    const size_t smem = 2 * (size_t)KA * kTcRows * 128 + 2 * (size_t)S * KA * kTcNT * 128 + 8 * (size_t)((NLEV + 1) * 32 + 66) * NB * 4 +
                        32 * S + 96 + 16 + 1024;
    if (smem > 232448) return GPSIG_E_UNSUPPORTED;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const long long nitems = (long long)p.ntiles * p.nch;
    const int grid = (int)(nitems < num_sms() ? nitems : num_sms());
    kern<<<grid, kTcThreads, smem, st>>>(mx, mz0, mz1, p);
    return check_launch();
}

template <int KA, bool ZOUT>
static int launch_tc_lev(int nlev, const CUtensorMap& mx, const CUtensorMap& mz0, const CUtensorMap& mz1, const TcParams& p,
                         cudaStream_t st) {
    switch (nlev) {
        case 1: return launch_tc_inst<1, KA, ZOUT>(mx, mz0, mz1, p, st);
        case 2: return launch_tc_inst<2, KA, ZOUT>(mx, mz0, mz1, p, st);
        case 3: return launch_tc_inst<3, KA, ZOUT>(mx, mz0, mz1, p, st);
        case 4: return launch_tc_inst<4, KA, ZOUT>(mx, mz0, mz1, p, st);
        case 5: return launch_tc_inst<5, KA, ZOUT>(mx, mz0, mz1, p, st);
        case 6: return launch_tc_inst<6, KA, ZOUT>(mx, mz0, mz1, p, st);
    }
    return GPSIG_E_UNSUPPORTED;
}

// Z (T, nz, 2, d), X (n, L, d) RAW; inv_ls may be NULL; out (nlev + 1, nz, n); flag: one device word (written here)
int launch_tens_seq_tc(const float* Z, long long nz, const float* X, long long n, int L, int d, const float* inv_ls, int nlev,
                       float* out, unsigned* flag, cudaStream_t st) {
    // operand layouts: d <= 8: 3 d + 6 slots in one 32-slot atom;  d = 9, 10: 3 d + 2 slots in one atom with the z norm applied
    // in the epilogue (ZOUT);  d <= 19: two atoms
    const bool zout = 3 * d + 6 > 32 && 3 * d + 2 <= 32;
    const int KA = (3 * d + 6 <= 32 || zout) ? 1 : 2, KW = 32 * KA;
    const int Lp = (L + kTcNT - 1) / kTcNT * kTcNT;
    std::vector<TcRow> rows;
    tc_row_map(nz, nlev, rows);
    const long long nrows = (long long)rows.size();
    const int ntiles = (int)(nrows / kTcRows);
    // scratch: sums (d floats, zeroed) | row map | Xop | Z0op | Z1op | z scales (ZOUT)
    auto up = [](size_t x) { return (x + 1023) / 1024 * 1024; };
    const size_t b_sums = up(64 * 4), b_rows = up((size_t)nrows * sizeof(TcRow)), b_x = up((size_t)n * Lp * KW * 4),
                 b_z = up((size_t)nrows * KW * 4);
    uint8_t* buf = nullptr;
    const size_t b_zs = up((size_t)2 * nrows * 4);
    cudaError_t e = cudaMallocAsync((void**)&buf, b_sums + b_rows + b_x + 2 * b_z + b_zs + 1024, st);
    if (e != cudaSuccess) return (int)e;
    uint8_t* w = (uint8_t*)(((uintptr_t)buf + 1023) / 1024 * 1024);
    float* sums = (float*)w; w += b_sums;
    TcRow* drows = (TcRow*)w; w += b_rows;
    float* Xop = (float*)w; w += b_x;
    float* Z0 = (float*)w; w += b_z;
    float* Z1 = (float*)w; w += b_z;
    float* zscale = zout ? (float*)w : nullptr;
    int rc = GPSIG_OK;
    auto done = [&](int code) { cudaFreeAsync(buf, st); return code; };
    if ((e = cudaMemsetAsync(sums, 0, b_sums, st)) != cudaSuccess) return done((int)e);
    if ((e = cudaMemsetAsync(flag, 0, sizeof(unsigned), st)) != cudaSuccess) return done((int)e);
    // the row map is tiny (16 bytes per row) and a pure function of (nz, nlev): staged through the stream
    if ((e = cudaMemcpyAsync(drows, rows.data(), (size_t)nrows * sizeof(TcRow), cudaMemcpyHostToDevice, st)) != cudaSuccess)
        return done((int)e);
    // cudaMemcpyAsync from pageable memory returns after the source has been staged, so `rows` may go out of scope
    const int cap = num_sms() * 8;
    {
        // the centre: the mean of the inducing-tensor points -- the same for every shard of the sequence axis (parallel.py),
        // so sharded and single-GPU results agree bit for bit; tensors sit where the data sits (gpsig/utils.py:25-63), and if
        // they do not the flag sends the call to the CUDA-core kernel
        const int T = nlev * (nlev + 1) / 2;
        const long long zpts = (long long)T * nz * 2;
        tc_centre_kernel<<<1, 256, 0, st>>>(Z, zpts, d, inv_ls, sums);
        if ((rc = check_launch())) return done(rc);
        long long blocks = (n * Lp + 255) / 256;
        tc_prep_x_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, st>>>(X, n, L, Lp, d, inv_ls, sums, 1.f / (float)zpts, KW, zout ? 1 : 0, Xop,
                                                                             flag);
        if ((rc = check_launch())) return done(rc);
        blocks = (2 * nrows + 255) / 256;
        tc_prep_z_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, st>>>(Z, nz, d, inv_ls, sums, 1.f / (float)zpts, drows, nrows,
                                                                             KW, zscale, Z0, Z1);
        if ((rc = check_launch())) return done(rc);
    }
    CUtensorMap mx, mz0, mz1;
    {
        uint64_t dims[2] = {(uint64_t)KW, (uint64_t)(n * Lp)}, strides[1] = {(uint64_t)KW * 4};
        uint32_t box[2] = {32, (uint32_t)kTcNT};
        if ((rc = encode_tensor_map_f32(&mx, Xop, 2, dims, strides, box, 1))) return done(rc);
        uint64_t dz[2] = {(uint64_t)KW, (uint64_t)nrows};
        uint32_t bz[2] = {32, (uint32_t)kTcRows};
        if ((rc = encode_tensor_map_f32(&mz0, Z0, 2, dz, strides, bz, 1))) return done(rc);
        if ((rc = encode_tensor_map_f32(&mz1, Z1, 2, dz, strides, bz, 1))) return done(rc);
    }
    TcParams p;
    p.rows = drows; p.flag = flag; p.ntiles = ntiles; p.nz = nz; p.n = n; p.L = L; p.Lp = Lp; p.out = out; p.zscale = zscale;
    // chunks of sequences: enough items to balance 148 SMs (a few per SM), an even number of sequences per chunk
    long long chunk = 64;
    while (chunk > 8 && (long long)ntiles * ((n + chunk - 1) / chunk) < 4LL * num_sms()) chunk >>= 1;
    p.chunk = (int)chunk;
    p.nch = (int)((n + chunk - 1) / chunk);
    {
        ProfScope prof(GPSIG_PROF_TENS, st, (double)nz * n);
        rc = zout ? launch_tc_lev<1, true>(nlev, mx, mz0, mz1, p, st)
                  : (KA == 1 ? launch_tc_lev<1, false>(nlev, mx, mz0, mz1, p, st) : launch_tc_lev<2, false>(nlev, mx, mz0, mz1, p, st));
    }
    return done(rc);
}

}  // namespace gpsig
