// api.cu — library-level entry points: version, device info and the per-thread error string.
#include <stdarg.h>
#include <string.h>

#include <atomic>

#include "common.cuh"

namespace mvs {

long long launches();

static thread_local char g_error[512] = "no error";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
long long launches() { return g_launches.load(std::memory_order_relaxed); }

}  // namespace mvs

extern "C" int mvs_version(void) { return 100; }   // 0.1.0

extern "C" long long mvs_launch_count(void) { return mvs::launches(); }

extern "C" const char* mvs_last_error_string(void) { return mvs::g_error; }

extern "C" int mvs_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0;
    MVS_CUDA_OK(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    MVS_CUDA_OK(cudaGetDeviceProperties(&prop, dev));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    return MVS_OK;
}
