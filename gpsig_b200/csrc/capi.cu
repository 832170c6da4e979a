// capi.cu -- C-ABI plumbing shared by every entry point: error strings, device queries, tensor-map encoding.
#include <string.h>

#include <atomic>
#include <mutex>
#include <vector>

#include <stdlib.h>

#include "internal.cuh"

namespace gpsig {

static thread_local char g_detail[512] = "";

void set_error_detail(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_detail, sizeof(g_detail), fmt, ap);
    va_end(ap);
}

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_detail, sizeof(g_detail), fmt, ap);
    va_end(ap);
    return code;
}

// ---- launch counter and per-class event timing --------------------------------------------------------------------
static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

struct ProfRec { int cls; double units; cudaEvent_t e0, e1; };
static std::atomic<int> g_prof_on{0};
static std::mutex g_prof_mu;
static std::vector<ProfRec*> g_prof_recs;

ProfScope::ProfScope(int cls_, cudaStream_t st_, double units) : cls(cls_), st(st_), rec(nullptr) {
    if (!g_prof_on.load(std::memory_order_relaxed)) return;
    ProfRec* r = new ProfRec{cls_, units, nullptr, nullptr};
    if (cudaEventCreate(&r->e0) != cudaSuccess || cudaEventCreate(&r->e1) != cudaSuccess) { delete r; return; }
    cudaEventRecord(r->e0, st);
    rec = r;
}
ProfScope::~ProfScope() {
    if (!rec) return;
    ProfRec* r = (ProfRec*)rec;
    cudaEventRecord(r->e1, st);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof_recs.push_back(r);
}

int num_sms() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return kNumSMsFallback;
    if (dev != cached_dev) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = kNumSMsFallback;
        cached = v;
        cached_dev = dev;
    }
    return cached;
}

// ---- tuning knobs: environment read once, then only gpsig_set_knob() changes them ---------------------------------
static int env_int_once(const char* name, int dflt) {
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}
static EnvKnobs& knobs_mut() {
    static EnvKnobs k = {env_int_once("GPSIG_WARPFUSED", 1), env_int_once("GPSIG_WARPFUSED_WARPS", 0),
                         env_int_once("GPSIG_STREAM_NCW", 0), env_int_once("GPSIG_STREAM_R", 0),
                         env_int_once("GPSIG_STREAM_S", 0), env_int_once("GPSIG_TENS_TC", 1)};
    return k;
}
const EnvKnobs& env_knobs() { return knobs_mut(); }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int encode_tensor_map_f32(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                          const uint32_t* box, int swizzle128) {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !sym)
            return fail(GPSIG_E_DRIVER, "cuTensorMapEncodeTiled not available from the driver (cudaError %d)", (int)e);
        fn = (EncodeTiledFn)sym;
    }
    cuuint64_t gdims[5], gstrides[4];
    cuuint32_t gbox[5], estr[5];
    for (int i = 0; i < rank; ++i) { gdims[i] = dims[i]; gbox[i] = box[i]; estr[i] = 1; }
    for (int i = 0; i + 1 < rank; ++i) gstrides[i] = strides_bytes[i];
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), gdims, gstrides, gbox, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        return fail(GPSIG_E_DRIVER,
                    "cuTensorMapEncodeTiled failed (CUresult %d): dims=[%llu,%llu,%llu,%llu,%llu] strides=[%llu,%llu,%llu,%llu] "
                    "box=[%u,%u,%u,%u,%u]",
                    (int)r, (unsigned long long)gdims[0], (unsigned long long)gdims[1], (unsigned long long)gdims[2],
                    (unsigned long long)gdims[3], (unsigned long long)gdims[4], (unsigned long long)gstrides[0],
                    (unsigned long long)gstrides[1], (unsigned long long)gstrides[2], (unsigned long long)gstrides[3], gbox[0],
                    gbox[1], gbox[2], gbox[3], gbox[4]);
    }
    return GPSIG_OK;
}

}  // namespace gpsig

extern "C" long long gpsig_launch_count(void) { return gpsig::g_launches.load(); }

extern "C" int gpsig_profile_enable(int on) {
    gpsig::g_prof_on.store(on ? 1 : 0);
    return GPSIG_OK;
}

extern "C" int gpsig_profile_reset(void) {
    std::lock_guard<std::mutex> lk(gpsig::g_prof_mu);
    for (gpsig::ProfRec* r : gpsig::g_prof_recs) {
        cudaEventSynchronize(r->e1);
        cudaEventDestroy(r->e0);
        cudaEventDestroy(r->e1);
        delete r;
    }
    gpsig::g_prof_recs.clear();
    return GPSIG_OK;
}

extern "C" int gpsig_profile_read(int cls, double* total_ms, long long* launches, double* units) {
    if (cls < 0 || cls >= GPSIG_PROF_NUM_CLASSES) return gpsig::fail(GPSIG_E_BADARG, "profile_read: unknown class %d", cls);
    double ms = 0.0, un = 0.0;
    long long n = 0;
    std::lock_guard<std::mutex> lk(gpsig::g_prof_mu);
    for (gpsig::ProfRec* r : gpsig::g_prof_recs) {
        if (r->cls != cls) continue;
        cudaError_t e = cudaEventSynchronize(r->e1);
        if (e != cudaSuccess) return (int)e;
        float t = 0.f;
        e = cudaEventElapsedTime(&t, r->e0, r->e1);
        if (e != cudaSuccess) return (int)e;
        ms += t; un += r->units; ++n;
    }
    if (total_ms) *total_ms = ms;
    if (launches) *launches = n;
    if (units) *units = un;
    return GPSIG_OK;
}

extern "C" int gpsig_set_knob(const char* name, int value) {
    if (!name) return gpsig::fail(GPSIG_E_BADARG, "set_knob: null name");
    gpsig::EnvKnobs& k = gpsig::knobs_mut();
    if (!strcmp(name, "warpfused")) k.warpfused = value;
    else if (!strcmp(name, "warpfused_warps")) k.warpfused_warps = value;
    else if (!strcmp(name, "stream_ncw")) k.stream_ncw = value;
    else if (!strcmp(name, "stream_r")) k.stream_r = value;
    else if (!strcmp(name, "stream_s")) k.stream_s = value;
    else if (!strcmp(name, "tens_tc")) k.tens_tc = value;
    else return gpsig::fail(GPSIG_E_BADARG, "set_knob: unknown knob '%s'", name);
    return GPSIG_OK;
}

extern "C" int gpsig_version(void) { return GPSIG_B200_VERSION; }

extern "C" const char* gpsig_error_string(int code) {
    switch (code) {
        case GPSIG_OK: return "ok";
        case GPSIG_E_BADARG: return "invalid argument";
        case GPSIG_E_UNSUPPORTED: return "unsupported configuration";
        case GPSIG_E_WORKSPACE: return "workspace too small";
        case GPSIG_E_ALIGN: return "alignment requirement not met";
        case GPSIG_E_DRIVER: return "CUDA driver entry point unavailable or failed";
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "unknown error";
}

extern "C" const char* gpsig_last_error_detail(void) { return gpsig::g_detail; }
