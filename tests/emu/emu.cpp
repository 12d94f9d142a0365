// emu.cpp — CPU emulation harness of the training-path kernels.  TEST INFRASTRUCTURE ONLY.
//
// Compiles mvsformer_b200/csrc/{geometry.cuh, train_kernels.cuh, train_entry.inl} — the very
// source the CUDA library is built from — with host definitions of the handful of CUDA names they
// use, and runs every "thread" in a loop on host memory.  It exports the same C symbols as
// libmvs_b200.so for the training path, plus plain-loop restatements of mvs_conv3d_cl /
// mvs_deconv3d_cl (whose CUDA kernels use shared memory and cannot be emulated this way), so the
// CPU test-suite can drive the package's autograd layer end to end without a GPU and compare
// it with torch autograd.  Nothing in mvsformer_b200/ loads this library.
#include <math.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/mvs_b200.h"

struct float4 { float x, y, z, w; };
#define __device__
#define __forceinline__ inline
#define __restrict__
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
#define MVS_ATOMIC_ADD_F(ptr, v) (*(ptr) += (v))
#define MVS_ATOMIC_ADD_D(ptr, v) (*(ptr) += (v))

static thread_local char g_error[512] = "no error";
namespace mvs {
static void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}
}  // namespace mvs
#define MVS_REQUIRE(cond, ...)               \
    do {                                     \
        if (!(cond)) {                       \
            ::mvs::set_error(__VA_ARGS__);   \
            return MVS_ERR_INVALID_ARGUMENT; \
        }                                    \
    } while (0)

#include "../../mvsformer_b200/csrc/geometry.cuh"
#include "../../mvsformer_b200/csrc/train_kernels.cuh"

static long long g_launches = 0;

namespace mvs {
namespace train {
template <class F>
static int launch_flat(const F& f, int64_t nthreads, void*, const char*) {
    const int64_t total = (nthreads + 255) / 256 * 256;
    for (int64_t tid = 0; tid < total; ++tid) f(tid, total);
    ++g_launches;
    return MVS_OK;
}
}  // namespace train
}  // namespace mvs

#include "../../mvsformer_b200/csrc/train_entry.inl"

#include "../../mvsformer_b200/csrc/fusion_kernels.cuh"
namespace mvs {
namespace fusion {
template <class F>
static int launch_flat(const F& f, int64_t nthreads, void*, const char*) {
    const int64_t total = (nthreads + 255) / 256 * 256;
    for (int64_t tid = 0; tid < total; ++tid) f(tid, total);
    ++g_launches;
    return MVS_OK;
}
}  // namespace fusion
}  // namespace mvs
#include "../../mvsformer_b200/csrc/fusion_entry.inl"

extern "C" const char* mvs_last_error_string(void) { return g_error; }
extern "C" long long mvs_launch_count(void) { return g_launches; }

// ---- restated semantics of mvs_homo_warp (warp.cu), on the shared geometry helpers ----------------------------
extern "C" int mvs_homo_warp(const float* src_fea, const float* relproj, const float* depth, int depth_is_map, float* warped,
                             uint8_t* mask, int B, int C, int D, int H, int W, void*) {
    using namespace mvs;
    MVS_REQUIRE(src_fea && relproj && depth && warped, "mvs_homo_warp: null pointer");
    const int64_t hw = (int64_t)H * W;
    for (int b = 0; b < B; ++b)
        for (int d = 0; d < D; ++d)
            for (int y = 0; y < H; ++y)
                for (int x = 0; x < W; ++x) {
                    const RelProj m = load_relproj(relproj + (int64_t)b * 12);
                    const PixelRay ray = pixel_ray(m, (float)x, (float)y);
                    const float dep = depth_is_map ? depth[((int64_t)b * D + d) * hw + (int64_t)y * W + x] : depth[(int64_t)b * D + d];
                    float gx, gy, qz;
                    const Taps t = make_taps(m, ray, dep, H, W, (float)((W - 1) / 2.0), (float)((H - 1) / 2.0), &gx, &gy, &qz);
                    if (mask) mask[((int64_t)b * D + d) * hw + (int64_t)y * W + x] =
                        ((gx > 1.0f) || (gx < -1.0f) || (gy > 1.0f) || (gy < -1.0f) || (qz <= 0.0f)) ? 1 : 0;
                    for (int c = 0; c < C; ++c)
                        warped[(((int64_t)b * C + c) * D + d) * hw + (int64_t)y * W + x] = sample4(src_fea + ((int64_t)b * C + c) * hw, t);
                }
    ++g_launches;
    return MVS_OK;
}

// ---- restated semantics of the two FP32 convolution entry points (include/mvs_b200.h) --------
extern "C" int mvs_conv3d_cl(const float* x, const float* w, const float* shift, const float* skip, float* y, int B, int D,
                             int H, int W, int Cin, int Cout, int kd, int sd, int sh, int sw, int relu, void*) {
    MVS_REQUIRE(x && w && y, "mvs_conv3d_cl: null pointer");
    MVS_REQUIRE(Cin % 4 == 0 && Cout % 8 == 0, "mvs_conv3d_cl: Cin %% 4 / Cout %% 8 (got %d, %d)", Cin, Cout);
    const int pd = kd / 2;
    const int Do = (D + 2 * pd - kd) / sd + 1, Ho = (H - 1) / sh + 1, Wo = (W - 1) / sw + 1;
    for (int b = 0; b < B; ++b)
        for (int oz = 0; oz < Do; ++oz)
            for (int oy = 0; oy < Ho; ++oy)
                for (int ox = 0; ox < Wo; ++ox) {
                    const int64_t o = ((((int64_t)b * Do + oz) * Ho + oy) * Wo + ox) * Cout;
                    for (int co = 0; co < Cout; ++co) {
                        float acc = 0.0f;
                        for (int kz = 0; kz < kd; ++kz)
                            for (int ky = 0; ky < 3; ++ky)
                                for (int kx = 0; kx < 3; ++kx) {
                                    const int iz = oz * sd - pd + kz, iy = oy * sh - 1 + ky, ix = ox * sw - 1 + kx;
                                    if (iz < 0 || iz >= D || iy < 0 || iy >= H || ix < 0 || ix >= W) continue;
                                    const float* px = x + ((((int64_t)b * D + iz) * H + iy) * W + ix) * Cin;
                                    const float* pw = w + (int64_t)((kz * 3 + ky) * 3 + kx) * Cin * Cout + co;
                                    for (int ci = 0; ci < Cin; ++ci) acc = fmaf(px[ci], pw[(int64_t)ci * Cout], acc);
                                }
                        if (shift) acc += shift[co];
                        if (relu) acc = fmaxf(acc, 0.0f);
                        if (skip) acc += skip[o + co];
                        y[o + co] = acc;
                    }
                }
    ++g_launches;
    return MVS_OK;
}

extern "C" int mvs_deconv3d_cl(const float* x, const float* w, const float* shift, const float* skip, float* y, int B,
                               int D, int H, int W, int Cin, int Cout, int kd, int sd, int relu, void*) {
    MVS_REQUIRE(x && w && y, "mvs_deconv3d_cl: null pointer");
    MVS_REQUIRE(Cin % 4 == 0 && Cout % 8 == 0, "mvs_deconv3d_cl: Cin %% 4 / Cout %% 8 (got %d, %d)", Cin, Cout);
    const int pd = kd / 2;
    const int Do = D * sd, Ho = H * 2, Wo = W * 2;
    const int64_t n = (int64_t)B * Do * Ho * Wo * Cout;
    for (int64_t i = 0; i < n; ++i) y[i] = 0.0f;
    // scatter form: o = i * s - pad + k
    for (int b = 0; b < B; ++b)
        for (int iz = 0; iz < D; ++iz)
            for (int iy = 0; iy < H; ++iy)
                for (int ix = 0; ix < W; ++ix) {
                    const float* px = x + ((((int64_t)b * D + iz) * H + iy) * W + ix) * Cin;
                    for (int kz = 0; kz < kd; ++kz)
                        for (int ky = 0; ky < 3; ++ky)
                            for (int kx = 0; kx < 3; ++kx) {
                                const int oz = iz * sd - pd + kz, oy = iy * 2 - 1 + ky, ox = ix * 2 - 1 + kx;
                                if (oz < 0 || oz >= Do || oy < 0 || oy >= Ho || ox < 0 || ox >= Wo) continue;
                                float* py = y + ((((int64_t)b * Do + oz) * Ho + oy) * Wo + ox) * Cout;
                                const float* pw = w + (int64_t)((kz * 3 + ky) * 3 + kx) * Cin * Cout;
                                for (int ci = 0; ci < Cin; ++ci)
                                    for (int co = 0; co < Cout; ++co) py[co] = fmaf(px[ci], pw[(int64_t)ci * Cout + co], py[co]);
                            }
                }
    for (int64_t i = 0; i < n; ++i) {
        float v = y[i];
        if (shift) v += shift[i % Cout];
        if (relu) v = fmaxf(v, 0.0f);
        if (skip) v += skip[i];
        y[i] = v;
    }
    ++g_launches;
    return MVS_OK;
}
