// train.cu — CUDA build of the training-path kernels (SURVEY.md §8b "Autograd", BASELINE cfg 5):
// materialised per-view group correlation and its backward (feature gradients by fp32 atomics),
// visibility-weighted aggregation fwd/bwd, train-mode BatchNorm (batch statistics in fp64,
// running-stat update, fused normalise + ReLU + skip) fwd/bwd, convolution weight gradients,
// the thin few-channel convolutions with device-resident weights, sigmoid / softmax backward.
// Data gradients of the 3D convolutions reuse mvs_conv3d_cl / mvs_deconv3d_cl with re-packed
// weights (the host side does the re-packing; see mvsformer_b200/autograd.py).
//
// The per-thread bodies live in train_kernels.cuh and the entry points in train_entry.inl, both
// shared verbatim with the test-suite's CPU emulation (tests/emu/emu.cpp).  This first version is
// written for correctness (flat kernels, atomics); it is FP32 CUDA-core code and is not yet tuned.
#include <math.h>

#include "common.cuh"

#define MVS_ATOMIC_ADD_F(ptr, v) atomicAdd((ptr), (v))
#define MVS_ATOMIC_ADD_D(ptr, v) atomicAdd((ptr), (v))

#include "train_kernels.cuh"

namespace mvs {
namespace train {

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

}  // namespace train
}  // namespace mvs

#include "train_entry.inl"
