// common.cuh — error plumbing and small device helpers shared by the mvs_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/mvs_b200.h"

namespace mvs {

void set_error(const char* fmt, ...);
void count_launch();   // bumps the counter read by mvs_launch_count()

#define MVS_REQUIRE(cond, ...)                      \
    do {                                            \
        if (!(cond)) {                              \
            ::mvs::set_error(__VA_ARGS__);          \
            return MVS_ERR_INVALID_ARGUMENT;        \
        }                                           \
    } while (0)

#define MVS_UNSUPPORTED(...)                        \
    do {                                            \
        ::mvs::set_error(__VA_ARGS__);              \
        return MVS_ERR_UNSUPPORTED;                 \
    } while (0)

#define MVS_CUDA_OK(expr)                                                                   \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            ::mvs::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return MVS_ERR_CUDA;                                                            \
        }                                                                                   \
    } while (0)

// After a kernel launch: catches launch-configuration errors without synchronising.
#define MVS_LAUNCH_OK(name)                                                                 \
    do {                                                                                    \
        cudaError_t _e = cudaGetLastError();                                                \
        if (_e != cudaSuccess) {                                                            \
            ::mvs::set_error("launch of %s failed: %s", name, cudaGetErrorString(_e));      \
            return MVS_ERR_CUDA;                                                            \
        }                                                                                   \
        ::mvs::count_launch();                                                              \
    } while (0)

static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// fp32 -> TF32 (round to nearest, ties away; low 13 mantissa bits cleared)
__device__ __forceinline__ float round_to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

}  // namespace mvs

#include "geometry.cuh"
