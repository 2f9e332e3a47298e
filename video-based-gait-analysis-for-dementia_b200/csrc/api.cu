// Library-level entry points: version, error text, device check, launch counter.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>

#include "common.cuh"

namespace gait {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int device_sm_count() {
    static std::atomic<int> cache[64];
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64) {
        const int c = cache[dev].load(std::memory_order_relaxed);
        if (c > 0) return c;
    }
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) { cudaGetLastError(); sms = 1; }
    if (dev >= 0 && dev < 64) cache[dev].store(sms, std::memory_order_relaxed);
    return sms;
}

static std::atomic<int> g_pdl_mask{-1};
int pdl_mask() {
    int on = g_pdl_mask.load(std::memory_order_relaxed);
    if (on < 0) {
        const char* e = getenv("GAITB200_PDL");
        on = e ? atoi(e) : 3;
        g_pdl_mask.store(on, std::memory_order_relaxed);
    }
    return on;
}

}  // namespace gait

extern "C" {

int gait_abi_version(void) { return GAIT_ABI_VERSION; }

const char* gait_error_string(int code) {
    switch (code) {
        case GAIT_OK: return "ok";
        case GAIT_ERR_INVALID: return "invalid argument";
        case GAIT_ERR_CUDA: return "CUDA error";
        case GAIT_ERR_UNSUPPORTED: return "unsupported shape";
        case GAIT_ERR_WORKSPACE: return "workspace too small";
        default: return "unknown error";
    }
}

const char* gait_last_error(void) { return gait::g_err; }

int gait_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0;
    GAIT_CUDA(cudaGetDevice(&dev));
    int sms = 0, major = 0, minor = 0;
    GAIT_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    GAIT_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    GAIT_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
    if (sm_count) *sm_count = sms;
    if (cc_major) *cc_major = major;
    if (cc_minor) *cc_minor = minor;
    if (major != 10) {
        gait::set_error("device is sm_%d%d; this library is built for sm_100a only", major, minor);
        return GAIT_ERR_UNSUPPORTED;
    }
    return GAIT_OK;
}

int gait_debug_pdl_mask(int mask) {
    const int prev = gait::pdl_mask();
    gait::g_pdl_mask.store(mask, std::memory_order_relaxed);      // < 0: back to GAITB200_PDL / the default at the next launch
    return prev;
}

int64_t gait_launch_count(void) { return gait::g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
