// common.cuh -- error plumbing and sm_100a PTX wrappers (mbarrier, TMA bulk-tensor copies) shared by all kernels.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/gpsig_b200.h"

namespace gpsig {

constexpr int kNumSMsFallback = 148;

// ---- host-side error detail (thread local) -------------------------------------------------------------------------
void set_error_detail(const char* fmt, ...);
int fail(int code, const char* fmt, ...);
int num_sms();
void count_launch();  // every kernel launch of the library passes through check_launch(): gpsig_launch_count()
inline int check_launch() {
    count_launch();
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? GPSIG_OK : (int)e;
}

// ---- measurement hooks (gpsig_profile_*): CUDA events around one kernel launch on its own stream ------------------
// Inactive (two predictable branches) unless gpsig_profile_enable(1) was called.  `units` is the number of work units
// the launch processes (sequence pairs for the recursion / producer) so that bench.py can turn the measured
// durations into algorithmic bytes per second.
struct ProfScope {
    int cls;
    cudaStream_t st;
    void* rec;
    ProfScope(int cls, cudaStream_t st, double units);
    ~ProfScope();
};

// ---- device-side PTX helpers ---------------------------------------------------------------------------------------
#if defined(__CUDACC__)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok;
}
// non-blocking probe of a phase (try_wait may suspend the thread for a while; test_wait never does)
__device__ __forceinline__ uint32_t mbar_test_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

// 5-D tiled TMA load global -> shared, completion signalled on an mbarrier (SASS: UTMALDG)
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], "
        "[%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
#endif

// ---- host: tensor-map encode through the driver entry point (no -lcuda link dependency) ----------------------------
// dims/strides innermost first; strides_bytes has rank-1 entries (dimension 0 is contiguous).
int encode_tensor_map_f32(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                          const uint64_t* strides_bytes, const uint32_t* box, int swizzle128);

}  // namespace gpsig
