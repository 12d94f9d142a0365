// geometry.cuh — plane-sweep sampling geometry shared by the warp, cost-volume and training kernels.
//
// Portable on purpose: it uses only __device__ / __forceinline__ / __restrict__, the IEEE
// round-to-nearest intrinsics (__fadd_rn, __fmul_rn, __fdiv_rn), __ldg and <math.h>, so the
// CPU emulation harness of the test-suite (tests/emu/emu.cpp) can compile the very same source
// with one-line host definitions of those names.  Under nvcc it is included through common.cuh.
#pragma once

namespace mvs {

// ---------------------------------------------------------------------------------------------
// Plane-sweep sampling geometry shared by the warp and cost-volume kernels.
// Mirrors models/warping.py:84-96 followed by ATen's grid_sampler (bilinear, zeros,
// align_corners=True): the normalise / un-normalise round trip is kept in fp32 so that the
// sample position rounds exactly as the reference's does.
// ---------------------------------------------------------------------------------------------
struct RelProj {  // rows of [R|t]
    float r00, r01, r02, t0;
    float r10, r11, r12, t1;
    float r20, r21, r22, t2;
};

__device__ __forceinline__ RelProj load_relproj(const float* __restrict__ p) {
    RelProj m;
    m.r00 = __ldg(p + 0); m.r01 = __ldg(p + 1); m.r02 = __ldg(p + 2);  m.t0 = __ldg(p + 3);
    m.r10 = __ldg(p + 4); m.r11 = __ldg(p + 5); m.r12 = __ldg(p + 6);  m.t1 = __ldg(p + 7);
    m.r20 = __ldg(p + 8); m.r21 = __ldg(p + 9); m.r22 = __ldg(p + 10); m.t2 = __ldg(p + 11);
    return m;
}

struct PixelRay {  // R (x, y, 1)^T
    float x, y, z;
};

__device__ __forceinline__ PixelRay pixel_ray(const RelProj& m, float px, float py) {
    PixelRay r;
    // same accumulation order as a k = 0,1,2 dot product (warping.py:90)
    r.x = __fadd_rn(__fadd_rn(__fmul_rn(m.r00, px), __fmul_rn(m.r01, py)), m.r02);
    r.y = __fadd_rn(__fadd_rn(__fmul_rn(m.r10, px), __fmul_rn(m.r11, py)), m.r12);
    r.z = __fadd_rn(__fadd_rn(__fmul_rn(m.r20, px), __fmul_rn(m.r21, py)), m.r22);
    return r;
}

struct Taps {
    int o00, o01, o10, o11;    // clamped element offsets y*W + x of the four taps (nw, ne, sw, se)
    float w00, w01, w10, w11;  // bilinear weights, already zeroed for out-of-image taps
};

// Projects one (pixel, depth) into the source image and prepares the four bilinear taps.
// gx/gy (normalised coordinates) and qz are returned for the out-of-bounds mask (warping.py:99-103).
__device__ __forceinline__ Taps make_taps(const RelProj& m, const PixelRay& ray, float depth, int H, int W,
                                          float half_w, float half_h, float* gx_out = nullptr,
                                          float* gy_out = nullptr, float* qz_out = nullptr) {
    // warping.py:91-93 : separate multiply and add (the reference materialises rot_depth_xyz)
    const float qx = __fadd_rn(__fmul_rn(ray.x, depth), m.t0);
    const float qy = __fadd_rn(__fmul_rn(ray.y, depth), m.t1);
    const float qz = __fadd_rn(__fmul_rn(ray.z, depth), m.t2);
    const float den = __fadd_rn(qz, 1e-6f);
    const float px = __fdiv_rn(qx, den);
    const float py = __fdiv_rn(qy, den);
    // warping.py:94-95
    const float gx = __fadd_rn(__fdiv_rn(px, half_w), -1.0f);
    const float gy = __fadd_rn(__fdiv_rn(py, half_h), -1.0f);
    // grid_sampler_unnormalize(align_corners=True): ((g + 1) / 2) * (size - 1)
    const float ix = __fmul_rn(__fmul_rn(__fadd_rn(gx, 1.0f), 0.5f), (float)(W - 1));
    const float iy = __fmul_rn(__fmul_rn(__fadd_rn(gy, 1.0f), 0.5f), (float)(H - 1));
    if (gx_out) *gx_out = gx;
    if (gy_out) *gy_out = gy;
    if (qz_out) *qz_out = qz;

    const float x0 = floorf(ix), y0 = floorf(iy);
    const float x1 = x0 + 1.0f, y1 = y0 + 1.0f;
    const float wx0 = x1 - ix, wx1 = ix - x0;
    const float wy0 = y1 - iy, wy1 = iy - y0;
    const float wmax = (float)(W - 1), hmax = (float)(H - 1);
    // comparisons are false for NaN, so non-finite coordinates sample zeros like the reference
    const bool vx0 = (x0 >= 0.0f) && (x0 <= wmax);
    const bool vx1 = (x1 >= 0.0f) && (x1 <= wmax);
    const bool vy0 = (y0 >= 0.0f) && (y0 <= hmax);
    const bool vy1 = (y1 >= 0.0f) && (y1 <= hmax);
    const int xi0 = (int)fminf(fmaxf(x0, 0.0f), wmax);
    const int xi1 = (int)fminf(fmaxf(x1, 0.0f), wmax);
    const int yi0 = (int)fminf(fmaxf(y0, 0.0f), hmax);
    const int yi1 = (int)fminf(fmaxf(y1, 0.0f), hmax);
    Taps t;
    t.o00 = yi0 * W + xi0; t.o01 = yi0 * W + xi1;
    t.o10 = yi1 * W + xi0; t.o11 = yi1 * W + xi1;
    t.w00 = (vx0 && vy0) ? wx0 * wy0 : 0.0f;
    t.w01 = (vx1 && vy0) ? wx1 * wy0 : 0.0f;
    t.w10 = (vx0 && vy1) ? wx0 * wy1 : 0.0f;
    t.w11 = (vx1 && vy1) ? wx1 * wy1 : 0.0f;
    return t;
}

__device__ __forceinline__ float sample4(const float* __restrict__ plane, const Taps& t) {
    // accumulation order nw, ne, sw, se as in ATen's grid_sampler_2d
    float v = __ldg(plane + t.o00) * t.w00;
    v = fmaf(__ldg(plane + t.o01), t.w01, v);
    v = fmaf(__ldg(plane + t.o10), t.w10, v);
    v = fmaf(__ldg(plane + t.o11), t.w11, v);
    return v;
}

}  // namespace mvs
