// C-ABI plumbing shared by all kernels: error string, launch check, version, device probe.
#include "common.cuh"
#include <cstdarg>
#include <cstdio>

namespace lv {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// Launch-configuration errors surface here; asynchronous faults surface at the caller's next sync.
int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) return LV_OK;
    set_error("%s: %s", what, cudaGetErrorString(e));
    return int(e);
}

}  // namespace lv

extern "C" int lv_version(void) { return 100; }   // 0.1.0

extern "C" const char* lv_last_error(void) { return lv::g_err; }

// Fills SM count and compute capability of the current device; returns 0 or a cudaError_t.
extern "C" int lv_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess && sm_count) e = cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (e == cudaSuccess && cc_major) e = cudaDeviceGetAttribute(cc_major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e == cudaSuccess && cc_minor) e = cudaDeviceGetAttribute(cc_minor, cudaDevAttrComputeCapabilityMinor, dev);
    if (e != cudaSuccess) { lv::set_error("lv_device_info: %s", cudaGetErrorString(e)); return int(e); }
    return LV_OK;
}
