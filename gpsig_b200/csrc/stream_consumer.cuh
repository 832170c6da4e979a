// stream_consumer.cuh -- the consumer side of the level recursion (signature_algs.py:8-35), shared by the stream-fed
// kernel (sigstream.cu: rows arrive by bulk copies from the chunk buffer) and the fused kernel (fused.cu: rows are
// computed by producer warps of the same CTA).  A consumer warp reads 2 KB skewed rows from its ring of S stages of R
// rows, guarded by full[s] / empty[s] mbarriers, and owns an independent stream of work items.
#pragma once
#include <type_traits>

#include "internal.cuh"

namespace gpsig {

constexpr int kSW = 16;               // columns per lane strip
constexpr int kRowBytes = 2048;       // one skewed row: G pairs x (16 LP) columns x 4 B, G * LP == 32

// what a consumer needs to know about the work-item stream and the output
struct StreamItems {
    long long nitems;
    int NW;           // streams in the launch (consumer warps of the grid)
    int R, S;         // rows per stage, stages per ring
    int Lin;          // stream rows per item
    int LP, log2LP, G;
    int njg;          // pair groups per row of the pair block
    int n1, n2;
    int upper_only, i_off, j_off;
    long long ldo;
    float* out;
    long long out_level_stride;
};

__device__ __forceinline__ void st_decode_item(const StreamItems& p, long long u, int& i, int& jg) {
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
    jg = (int)(u - items_before(lo, p.njg, p.G, 1, p.i_off, p.j_off)) + first_group(lo, p.G, 1, p.i_off, p.j_off);
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// ring: shared address of this consumer's stages; fb / eb: its full / empty barrier arrays; wg: its stream index
template <int NLEV>
__device__ __forceinline__ void run_stream_consumer(const StreamItems& p, uint32_t ring, uint32_t fb, uint32_t eb, long long wg,
                                                    int lane) {
    constexpr int NA = NLEV > 1 ? NLEV - 1 : 1;
    const int S = p.S, R = p.R, Lin = p.Lin, LP = p.LP;
    const uint32_t stage_bytes = (uint32_t)R * kRowBytes;
    const long long nloc = wg >= p.nitems ? 0 : (p.nitems - wg + p.NW - 1) / p.NW;
    const long long total = nloc * Lin;           // item rows of the stream
    const long long nsteps = total + LP - 1;      // skewed rows of the stream
    if (total == 0) return;
    const int l = lane & (LP - 1), q = lane >> p.log2LP;
    // swizzled byte offsets of this lane's four 16-byte chunks inside a skewed row (same function as the writer)
    uint32_t off[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) off[k] = swizzle_in_row((uint32_t)(q * LP * kSW * 4 + l * 64 + k * 16));

    float A[NA][kSW];
    float psum[NLEV], ksum[NLEV];
    float g[kSW];  // increments of the current row, loaded one step ahead
#pragma unroll
    for (int m = 0; m < NLEV; ++m) { psum[m] = 0.f; ksum[m] = 0.f; }
#pragma unroll
    for (int m = 0; m < NA; ++m)
#pragma unroll
        for (int j = 0; j < kSW; ++j) A[m][j] = 0.f;

    int lstage = 0, lrow = 0;  // stage / row-in-stage of the next skewed row to load (warp-uniform)
    uint32_t lphase = 0;
    uint32_t okn = 0;          // the stage about to be entered was already seen complete
    int rho = 0;               // increment row of the current item this lane is at
    long long item = wg;

    // skewed row -> registers; on entering a stage wait for its bytes, on leaving it hand it back to the producer
    auto load_row = [&](float (&gn)[kSW]) {
        if (lrow == 0 && !okn) mbar_wait(fb + 8 * lstage, lphase);
        const uint32_t base = ring + lstage * stage_bytes + lrow * kRowBytes;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(gn[4 * k]), "=f"(gn[4 * k + 1]), "=f"(gn[4 * k + 2]), "=f"(gn[4 * k + 3])
                         : "r"(base + off[k]));
        }
        okn = 0;
        if (++lrow == R) {
            lrow = 0;
            __syncwarp();
            if (lane == 0) mbar_arrive(eb + 8 * lstage);
            if (++lstage == S) { lstage = 0; lphase ^= 1u; }
            okn = mbar_test_wait(fb + 8 * lstage, lphase);  // looked at one step later
        }
    };

    auto step = [&](auto check_tag, long long T) {
        constexpr bool CHECK = decltype(check_tag)::value;
        // running row prefixes arrive from the strip to the left (it finished this row one step ago)
        float pin[NLEV];
#pragma unroll
        for (int m = 0; m < NLEV; ++m) {
            pin[m] = __shfl_up_sync(0xffffffffu, psum[m], 1);
            if (l == 0) pin[m] = 0.f;
        }
        const bool valid = CHECK ? (T - l >= 0 && T - l < total) : true;
        float gn[kSW];
        if (!CHECK || T + 1 < nsteps) load_row(gn);
        float d[kSW];
#pragma unroll
        for (int j = 0; j < kSW; ++j) d[j] = valid ? g[j] : 0.f;
        if (valid && rho == 0) {  // first row of an item
#pragma unroll
            for (int m = 0; m < NLEV; ++m) ksum[m] = 0.f;
#pragma unroll
            for (int m = 0; m < NA; ++m)
#pragma unroll
                for (int j = 0; j < kSW; ++j) A[m][j] = 0.f;
        }
        // ---- the recursion: 2 FP ops per entry per level ----
#pragma unroll
        for (int m = 0; m < NLEV; ++m) psum[m] = pin[m];
#pragma unroll
        for (int j = 0; j < kSW; ++j) {
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
            if (rho == Lin - 1) {
                if (l == LP - 1) {
                    int i, jg;
                    st_decode_item(p, item, i, jg);
                    const int j = jg * p.G + q;
                    if (j < p.n2) {
                        float* o = p.out + (long long)(p.i_off + i) * p.ldo + p.j_off + j;
                        o[0] = 1.f;
#pragma unroll
                        for (int m = 0; m < NLEV; ++m) o[(long long)(m + 1) * p.out_level_stride] = ksum[m];
                    }
                }
                rho = 0;
                item += p.NW;
            } else {
                ++rho;
            }
        }
#pragma unroll
        for (int j = 0; j < kSW; ++j) g[j] = gn[j];
    };

    load_row(g);  // skewed row 0
    long long T = 0;
    const long long ramp = (LP - 1) < nsteps ? (LP - 1) : nsteps;
    for (; T < ramp; ++T) step(std::true_type{}, T);
    for (; T < total - 1; ++T) step(std::false_type{}, T);
    for (; T < nsteps; ++T) step(std::true_type{}, T);
}

}  // namespace gpsig
