// fused.cu -- K(X, X) / K(X, X2) level stacks with the increment Gram computed INSIDE the recursion kernel: the
// (N L)^2 tensor of kernels.py:226 never exists, not even chunk-wise in HBM.
//
// One persistent CTA per SM, 12 warps: 4 CONSUMER warps run the level recursion of stream_consumer.cuh on 2 KB skewed
// rows; each is fed by a PAIR of PRODUCER warps that compute those rows straight into the consumer's shared-memory ring
// (same stream layout as gram.cu writes to HBM for the two-kernel pipeline: row T holds strip l of item row T - l,
// 16-byte chunks XOR-swizzled).  A producer thread owns 8 consecutive columns of one pair of the current item -- its 8
// (9 for RBF: the halo column) points of the column sequence stay in registers for the whole item -- and walks the
// item's rows in the SAME skew as the consumer lanes, so every producer iteration completes one stream row.  The row
// sequence x_i of the item sits in a double-buffered shared-memory tile (strips roll over to the next item at different
// iterations).  Arithmetic is identical to gram.cu's fast producers (packed fma.rn.f32x2 dot products, ex2.approx on
// the augmented form for RBF) and to sigstream.cu's consumer, so the results are bit-identical to the two-kernel path.
//   full[c][s]  : count 2 -- both producer warps have stored their halves of the stage
//   empty[c][s] : count 1 -- the consumer has read the stage
//
// STATUS (round 1): correct and bit-identical to the two-kernel path (tests/test_gpu_large.py), but NOT the default:
// measured 222 ms (Linear) / 287 ms (RBF) for K(X,X) at N=4096, L=128, d=8, M=5 against 190-210 / 202-226 ms for
// producer + stream recursion (first version: 304 / 393 ms, profiles/r1m_fused_v1.md, r1p_fused_v2.md -- sleep-polling
// producers and a 64-bit-division item decode per strip cost 40 % of the issued instructions; both are gone).
// With the producers' arithmetic switched off (GPSIG_FUSED_DBG=1) the kernel still takes 181 / 208 ms: the bound is the
// consumer side -- 4 recursion warps per SM where the stream kernel runs 11 (a consumer needs 158 registers), each
// waiting on a two-warp handshake every 4 rows.  Next round: 8-column strips (A_m state 32 instead of 64 registers
// -> ~95 registers -> 10-12 consumers per SM) and prefetch of the next item's columns / x tile.
// Enable with GPSIG_FUSED=1.
#include <stdlib.h>

#include "stream_consumer.cuh"

namespace gpsig {

constexpr int kFusedConsumers = 4;
constexpr int kFusedWarps = 12;  // 4 consumers + 4 x 2 producers; 168 registers per thread
constexpr int kFusedCols = 8;    // columns per producer thread

struct FusedParams {
    const float* A;   // prepared row-side points / increments   (n1_total, rowsA, DPA)
    const float* B;   // prepared column-side points / increments (n2_total, rowsB, DPA)
    int rowsA, rowsB; // == stream rows per item (RBF: the first row only primes the differencing and emits zeros)
    int P;            // padded columns per pair (16 LP)
    int dbg;          // experiments only: 1 = producers skip the arithmetic (stores zeros)
    StreamItems it;   // it.Lin == rowsA
};

__device__ __forceinline__ float fused_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void pair_barrier(int id) {  // the two producer warps of one consumer
    asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory");
}

// DPA = floats per prepared point; HU = leading float2 pairs that carry data: DPA/2 for LINEAR, DPA/2 - 1 for RBF (padded
// features, then the two augmentation slots; the last pair of the RBF layout is padding).
template <bool RBF, int NLEV, int DPA, int HU>
__global__ void __launch_bounds__(kFusedWarps * 32, 1) sigkern_fused_kernel(const FusedParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const StreamItems& it = p.it;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = it.S, R = it.R, Lrow = it.Lin;
    const uint32_t stage_bytes = (uint32_t)R * kRowBytes;
    const uint32_t smem0 = smem_u32(smem);
    const uint32_t xbytes = (uint32_t)p.rowsA * DPA * 4;                       // one row sequence
    const uint32_t x0 = smem0 + (uint32_t)kFusedConsumers * S * stage_bytes;  // xbuf[c][2]
    const uint32_t full0 = x0 + (uint32_t)kFusedConsumers * 2 * xbytes;
    const uint32_t empty0 = full0 + (uint32_t)kFusedConsumers * S * 8;
    if (threadIdx.x == 0) {
        for (int k = 0; k < kFusedConsumers * S; ++k) {
            mbar_init(full0 + 8 * k, 2);
            mbar_init(empty0 + 8 * k, 1);
        }
        fence_mbar_init();
    }
    __syncthreads();

    // warps 0..7 produce (pair c = warp / 2), warps 8..11 consume: the SM's warp arbiter favours the higher warp ids, and
    // the consumers are the critical path (a producer that is ahead only polls its empty barrier)
    constexpr int kProducerWarps = kFusedWarps - kFusedConsumers;
    const int c = warp >= kProducerWarps ? warp - kProducerWarps : warp >> 1;  // consumer / stream inside the CTA
    const long long wg = (long long)blockIdx.x * kFusedConsumers + c;
    const uint32_t ring = smem0 + (uint32_t)c * S * stage_bytes;
    const uint32_t fb = full0 + (uint32_t)c * S * 8, eb = empty0 + (uint32_t)c * S * 8;
    if (warp >= kProducerWarps) {
        run_stream_consumer<NLEV>(it, ring, fb, eb, wg, lane);
        return;
    }

    // ===== producers ================================================================================================
    constexpr int NPT = RBF ? kFusedCols + 1 : kFusedCols;
    const long long nloc = wg >= it.nitems ? 0 : (it.nitems - wg + it.NW - 1) / it.NW;
    const long long total = nloc * Lrow;
    const long long nsteps = total + it.LP - 1;
    if (total == 0) return;
    const int half = warp & 1;
    const int pt = half * 32 + lane;             // 0..63 inside the pair of warps
    const int cb = pt * kFusedCols;              // column block inside the 512-column item row
    const int q = cb / p.P, t0 = cb - q * p.P;   // pair inside the item, first column inside the pair
    const int l = t0 >> 4;                       // strip: this thread works on item row (iteration - l)
    const uint32_t b0 = (uint32_t)(q * p.P + t0) * 4u;
    const uint32_t sw0 = swizzle_in_row(b0), sw1 = swizzle_in_row(b0 + 16u);
    const uint32_t xb = x0 + (uint32_t)c * 2 * xbytes;
    const float* xgen = reinterpret_cast<const float*>(smem) + ((xb - smem0) >> 2);
    const int barid = 1 + c;

    float2 y[NPT][HU];
    float fprev[NPT];
#pragma unroll
    for (int u = 0; u < NPT; ++u) {
        fprev[u] = 0.f;
#pragma unroll
        for (int h = 0; h < HU; ++h) y[u][h] = make_float2(0.f, 0.f);
    }
    // (i, jg) of an item of this stream, advanced item by item (u -> u + NW) without divisions: `rel` is the position of the
    // group inside the groups row i keeps (all of them, or those right of the diagonal for symmetric K)
    struct Track { int i, rel, cnt; };
    auto track_init = [&](Track& t, long long u) {
        int jg;
        st_decode_item(it, u, t.i, jg);
        const int fg = first_group(t.i, it.G, it.upper_only, it.i_off, it.j_off);
        t.rel = jg - fg;
        t.cnt = it.njg - fg;
    };
    auto track_next = [&](Track& t) {
        t.rel += it.NW;
        while (t.rel >= t.cnt && t.i + 1 < it.n1) {
            t.rel -= t.cnt;
            ++t.i;
            t.cnt = it.njg - first_group(t.i, it.G, it.upper_only, it.i_off, it.j_off);
        }
    };
    Track tx, ty;        // item whose row sequence is staged next / item this strip works on
    track_init(tx, wg);
    ty = tx;
    int s = -l;          // row of the current item (negative: the strip has not started yet)
    long long m = 0;     // index of the current item inside the stream
    int stage = 0, srow = 0, round = 0;
    int s0 = 0;          // row of strip 0 (= sg mod Lrow)
    long long mi = 0;    // item of strip 0 (= sg div Lrow)
    for (long long sg = 0; sg < nsteps; ++sg) {
        // ---- the pair stages the row sequence of item (sg / Lrow) when strip 0 reaches it ----
        if (s0 == 0 && sg < total) {
            const int i = tx.i;
            track_next(tx);
            const float4* src = reinterpret_cast<const float4*>(p.A + (long long)(it.i_off + i) * p.rowsA * DPA);
            const uint32_t dst = xb + (uint32_t)(mi & 1) * xbytes;
            for (int e = pt; e < p.rowsA * (DPA / 4); e += 64) {
                const float4 v = __ldg(src + e);
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst + e * 16), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                             : "memory");
            }
            pair_barrier(barid);
        }
        // ---- entering a ring stage: wait until the consumer has handed it back ----
        if (srow == 0 && round > 0) {
            if (lane == 0) mbar_wait(eb + 8 * stage, (uint32_t)(round + 1) & 1u);  // try_wait suspends in hardware
            __syncwarp();
        }
        const bool valid = s >= 0 && (sg - l) < total;
        // ---- this strip starts a new item: its column points move into registers ----
        if (valid && s == 0) {
            const int jg = ty.rel + first_group(ty.i, it.G, it.upper_only, it.i_off, it.j_off);
            track_next(ty);
            int jl = jg * it.G + q;
            if (jl > it.n2 - 1) jl = it.n2 - 1;  // padding pair of a ragged last group: any valid sequence will do
            const long long j = (long long)it.j_off + jl;
#pragma unroll
            for (int u = 0; u < NPT; ++u) {
                const int t = t0 + u;
                const bool ok = RBF || t < p.rowsB;
                const int tc = t < p.rowsB ? t : p.rowsB - 1;
                const float2* src = reinterpret_cast<const float2*>(p.B + (j * p.rowsB + tc) * DPA);
#pragma unroll
                for (int h = 0; h < HU; ++h) y[u][h] = ok ? __ldg(src + h) : make_float2(0.f, 0.f);
                if (RBF) {  // column side of the augmented product: (..., 1, -|y|^2/2)
                    const float2 a = y[u][HU - 1];
                    y[u][HU - 1] = make_float2(a.y, a.x);
                }
            }
        }
        float o[kFusedCols];
#pragma unroll
        for (int u = 0; u < kFusedCols; ++u) o[u] = 0.f;
        if (valid && p.dbg != 1) {
            const float* xs = xgen + (size_t)(m & 1) * (xbytes >> 2) + (size_t)s * DPA;
            float2 x[HU];
#pragma unroll
            for (int h = 0; h < HU; ++h) x[h] = *reinterpret_cast<const float2*>(xs + 2 * h);
            float f[NPT];
#pragma unroll
            for (int u = 0; u < NPT; ++u) {
                float2 acc = make_float2(0.f, 0.f);
#pragma unroll
                for (int h = 0; h < HU; ++h) acc = __ffma2_rn(x[h], y[u][h], acc);
                const float v = acc.x + acc.y;
                f[u] = RBF ? fused_ex2(v) : v;
            }
            if (RBF) {
                if (s > 0) {
#pragma unroll
                    for (int u = 0; u < kFusedCols; ++u) o[u] = (f[u + 1] - f[u]) - (fprev[u + 1] - fprev[u]);
                }
#pragma unroll
                for (int u = 0; u < NPT; ++u) fprev[u] = f[u];
            } else {
#pragma unroll
                for (int u = 0; u < kFusedCols; ++u) o[u] = f[u];
            }
        }
        // ---- two swizzled 16-byte chunks of stream row sg ----
        {
            const uint32_t row = ring + stage * stage_bytes + srow * kRowBytes;
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(row + sw0), "f"(o[0]), "f"(o[1]), "f"(o[2]), "f"(o[3])
                         : "memory");
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(row + sw1), "f"(o[4]), "f"(o[5]), "f"(o[6]), "f"(o[7])
                         : "memory");
        }
        if (++srow == R || sg == nsteps - 1) {
            __syncwarp();
            if (lane == 0) mbar_arrive(fb + 8 * stage);
            srow = 0;
            if (++stage == S) { stage = 0; ++round; }
        }
        if (++s == Lrow) { s = 0; ++m; }  // the strip's row counter started at -l
        if (++s0 == Lrow) { s0 = 0; ++mi; }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------
static int flog2(int x) {
    int l = 0;
    while ((1 << l) < x) ++l;
    return l;
}

template <bool RBF, int NLEV, int DPA, int HU>
static int launch_fused_inst(const FusedParams& p, int grid, size_t smem, cudaStream_t st) {
    auto kern = sigkern_fused_kernel<RBF, NLEV, DPA, HU>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    kern<<<grid, kFusedWarps * 32, smem, st>>>(p);
    return check_launch();
}

template <bool RBF, int DPA, int HU>
static int launch_fused_lev(int nlev, const FusedParams& p, int grid, size_t smem, cudaStream_t st) {
    switch (nlev) {
        case 2: return launch_fused_inst<RBF, 2, DPA, HU>(p, grid, smem, st);
        case 3: return launch_fused_inst<RBF, 3, DPA, HU>(p, grid, smem, st);
        case 4: return launch_fused_inst<RBF, 4, DPA, HU>(p, grid, smem, st);
        case 5: return launch_fused_inst<RBF, 5, DPA, HU>(p, grid, smem, st);
    }
    return GPSIG_E_UNSUPPORTED;
}

// instantiated: up to 8 features, 2..5 levels; short sequences stay on the two-kernel path (the x tile of an item is
// reused two items later: the producer warps must not be able to run a whole item ahead of each other)
bool fused_supported(bool rbf, int d, int nlev, int LP, int rowsA) {
    (void)rbf;
    if (nlev < 2 || nlev > 5 || d > 8) return false;
    if (LP < 2 || LP > 32 || rowsA < 48) return false;
    // opt-in (GPSIG_FUSED=1): round-1 measurements have it 1.5x SLOWER than the two-kernel pipeline -- see the header
    const char* v = getenv("GPSIG_FUSED");
    return v && *v == '1';
}

// Level stacks of the pair block rows [i_off, i_off + n1) x cols [j_off, j_off + n2) from prepared points (gram.cu prep
// modes 1 / 2).  Returns GPSIG_E_UNSUPPORTED (no detail) when there is no instantiation.
int launch_sigkern_fused(bool rbf, const float* A, const float* B, int rowsA, int rowsB, int d, int DPA, int P, int LP,
                         long long nitems, int n1, int n2, int nlev, int upper_only, int i_off, int j_off, long long ldo,
                         long long lvl_stride, float* out, cudaStream_t st) {
    if (!A || !B || !out || nitems < 1 || rowsA < 1 || rowsB < 1) return fail(GPSIG_E_BADARG, "sigkern_fused: bad sizes");
    FusedParams p;
    p.A = A; p.B = B; p.rowsA = rowsA; p.rowsB = rowsB; p.P = P;
    { const char* dv = getenv("GPSIG_FUSED_DBG"); p.dbg = (dv && *dv) ? atoi(dv) : 0; }
    StreamItems& it = p.it;
    it.nitems = nitems;
    it.R = 4; it.S = 3;
    {   // tuning knobs for experiments
        const char* r = getenv("GPSIG_FUSED_R");
        const char* ss = getenv("GPSIG_FUSED_S");
        if (r && *r) it.R = atoi(r);
        if (ss && *ss) it.S = atoi(ss);
        if (it.R < 1) it.R = 1;
        if (it.S < 2) it.S = 2;
    }
    it.Lin = rowsA; it.LP = LP; it.log2LP = flog2(LP); it.G = 32 / LP;
    it.njg = (n2 + it.G - 1) / it.G;
    it.n1 = n1; it.n2 = n2;
    it.upper_only = upper_only ? 1 : 0; it.i_off = i_off; it.j_off = j_off;
    it.ldo = ldo; it.out = out; it.out_level_stride = lvl_stride;
    long long want = (nitems + kFusedConsumers - 1) / kFusedConsumers;
    const int grid = (int)(want < num_sms() ? want : num_sms());
    it.NW = grid * kFusedConsumers;
    const size_t smem = (size_t)kFusedConsumers * it.S * ((size_t)it.R * kRowBytes + 16) + (size_t)kFusedConsumers * 2 * rowsA * DPA * 4;
    if (smem > 232448) return GPSIG_E_UNSUPPORTED;
    (void)d;
    ProfScope prof(GPSIG_PROF_FUSED, st, (double)nitems * it.G);
    if (rbf) {
        if (DPA == 8) return launch_fused_lev<true, 8, 3>(nlev, p, grid, smem, st);
        if (DPA == 12) return launch_fused_lev<true, 12, 5>(nlev, p, grid, smem, st);
    } else {
        if (DPA == 4) return launch_fused_lev<false, 4, 2>(nlev, p, grid, smem, st);
        if (DPA == 8) return launch_fused_lev<false, 8, 4>(nlev, p, grid, smem, st);
    }
    return GPSIG_E_UNSUPPORTED;
}

}  // namespace gpsig
