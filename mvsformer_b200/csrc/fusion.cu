// fusion.cu — CUDA build of the depth-map fusion kernels (misc/fusion.py:69-118 as driven by
// test.py:404-435): photometric threshold, reprojection of every source depth map into the reference view,
// geometric-consistency masks, averaged depth, world points.  Per-thread bodies in fusion_kernels.cuh, entry points
// in fusion_entry.inl, both shared with the test-suite's CPU emulation.  One thread per pixel; HBM-bound: per
// reference view (1 + V) depth maps in, (5 V + 2) maps out.  First version: written after the round's GPU minutes
// were spent, verified by emulation against the reference's own functions; GPU tests gated (tests/test_gpu_experimental.py).
#include <math.h>

#include "common.cuh"
#include "fusion_kernels.cuh"

namespace mvs {
namespace fusion {

template <class F>
__global__ void __launch_bounds__(256) flat_kernel(F f, int64_t nthreads) {
    const int64_t tid = (int64_t)blockIdx.x * 256 + threadIdx.x;
    f(tid, nthreads);
}

template <class F>
static int launch_flat(const F& f, int64_t nthreads, void* stream, const char* name) {
    const int64_t blocks = (nthreads + 255) / 256;
    MVS_REQUIRE(blocks >= 1 && blocks <= 0x7fffffffLL, "%s: %lld threads do not fit one grid", name, (long long)nthreads);
    flat_kernel<F><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(f, blocks * 256);
    MVS_LAUNCH_OK(name);
    return MVS_OK;
}

}  // namespace fusion
}  // namespace mvs

#include "fusion_entry.inl"
