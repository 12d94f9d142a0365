// fusion_kernels.cuh — per-thread bodies of the depth-map fusion kernels (the
// geometric-consistency filter and depth averaging of misc/fusion.py:79-118 as used by
// test.py:404-431, "the same reprojection / bilinear-gather primitive as the cost volume").
//
// Flat kernels like train_kernels.cuh (one thread = one pixel, no shared memory, no barriers), so the
// test-suite's CPU emulation (tests/emu/emu.cpp) compiles this very source and checks it against the
// reference's own functions without a GPU.  The includer provides __device__, __forceinline__,
// __restrict__, __ldg, <math.h>, <stdint.h>.
//
// Camera algebra follows misc/fusion.py operation by operation, including its conventions: pixel centres at
// (x + 0.5, y + 0.5) (:8-13), a "+ 1e-9" in every homogeneous division (:24-49), and project_img's
// normalisation by width / height (not width - 1) before grid_sample(align_corners=True) (:59-65).
// Matrices arrive per (batch, source view) as FUSION_MAT_FLOATS floats, inverted on the host in fp64:
//   ref: Kinv (9) Einv (16) E (16) K (9)   then   src: Kinv (9) Einv (16) E (16) K (9)
#pragma once

#define MVS_FUSION_MAT_FLOATS 100

namespace mvs {
namespace fusion {

struct V3 { float x, y, z; };
struct V4 { float x, y, z, w; };

__device__ __forceinline__ V3 mul3(const float* __restrict__ m, float x, float y, float z) {
    V3 r;
    r.x = m[0] * x + m[1] * y + m[2] * z;
    r.y = m[3] * x + m[4] * y + m[5] * z;
    r.z = m[6] * x + m[7] * y + m[8] * z;
    return r;
}

__device__ __forceinline__ V4 mul4(const float* __restrict__ m, const V4& v) {
    V4 r;
    r.x = m[0] * v.x + m[1] * v.y + m[2] * v.z + m[3] * v.w;
    r.y = m[4] * v.x + m[5] * v.y + m[6] * v.z + m[7] * v.w;
    r.z = m[8] * v.x + m[9] * v.y + m[10] * v.z + m[11] * v.w;
    r.w = m[12] * v.x + m[13] * v.y + m[14] * v.z + m[15] * v.w;
    return r;
}

// idx_img2cam (:24-29): K^-1 (u, v, 1), divided by (z + 1e-9), times depth; homogeneous 1
__device__ __forceinline__ V4 img2cam(const float* __restrict__ kinv, float u, float v, float depth) {
    const V3 p = mul3(kinv, u, v, 1.0f);
    const float s = p.z + 1e-9f;
    V4 r;
    r.x = p.x / s * depth; r.y = p.y / s * depth; r.z = p.z / s * depth; r.w = 1.0f;
    return r;
}

// idx_cam2world (:32-35) with E^-1, idx_world2cam (:38-41) with E: 4x4 product, divided by (w + 1e-9)
__device__ __forceinline__ V4 transform_h(const float* __restrict__ m, const V4& p) {
    V4 r = mul4(m, p);
    const float s = r.w + 1e-9f;
    r.x /= s; r.y /= s; r.z /= s; r.w /= s;
    return r;
}

// idx_cam2img (:44-48): (c.xyz / (c.w + 1e-9)) through K, divided by (z + 1e-9)
__device__ __forceinline__ V3 cam2img(const float* __restrict__ k, const V4& c) {
    const float s = c.w + 1e-9f;
    V3 i = mul3(k, c.x / s, c.y / s, c.z / s);
    const float t = i.z + 1e-9f;
    i.x /= t; i.y /= t; i.z /= t;
    return i;
}

// get_reproj (:79-98): for every reference pixel and source view, (x, y, depth) of the source depth map's
// surface re-expressed in the reference view, bilinearly gathered at the reference pixel's projection into
// the source view; plus project_img's in-range flag.  The intermediate per-source-pixel image srcs2ref_xyd of
// the reference is evaluated on the fly at the four taps.
//   ref_depth [n,h,w], src_depths [n,v,h,w], mats [n,v,100] -> reproj_xyd [n,v,3,h,w], in_range [n,v,h,w]
// one thread per (n, v, y, x)
__device__ __forceinline__ void reproject_thread(const float* __restrict__ ref_depth, const float* __restrict__ src_depths,
                                                 const float* __restrict__ mats, float* __restrict__ reproj_xyd,
                                                 float* __restrict__ in_range, int N, int V, int H, int W, int64_t tid) {
    const int64_t hw = (int64_t)H * W;
    if (tid >= (int64_t)N * V * hw) return;
    const int x = (int)(tid % W), y = (int)((tid / W) % H);
    const int64_t nv = tid / hw;
    const int64_t n = nv / V;
    const float* m = mats + nv * MVS_FUSION_MAT_FLOATS;
    const float *r_kinv = m, *r_einv = m + 9, *r_e = m + 25, *r_k = m + 41;
    const float *s_kinv = m + 50, *s_einv = m + 59, *s_e = m + 75, *s_k = m + 91;

    // reference pixel -> source image (project_img :51-58)
    const float dr = __ldg(ref_depth + n * hw + (int64_t)y * W + x);
    const V4 pr = img2cam(r_kinv, (float)x + 0.5f, (float)y + 0.5f, dr);
    const V3 is = cam2img(s_k, transform_h(s_e, transform_h(r_einv, pr)));
    float gx = is.x / (float)W * 2.0f - 1.0f, gy = is.y / (float)H * 2.0f - 1.0f;            // :59-61
    gx = fminf(fmaxf(gx, -1.1f), 1.1f);
    gy = fminf(fmaxf(gy, -1.1f), 1.1f);
    const bool inr = (-1.0f <= gx) && (gx <= 1.0f) && (-1.0f <= gy) && (gy <= 1.0f);         // :62-63 (false for NaN)
    // grid_sample(bilinear, zeros, align_corners=True)
    const float ix = (gx + 1.0f) * 0.5f * (float)(W - 1), iy = (gy + 1.0f) * 0.5f * (float)(H - 1);
    const float x0 = floorf(ix), y0 = floorf(iy);
    const float wx1 = ix - x0, wx0 = 1.0f - wx1, wy1 = iy - y0, wy0 = 1.0f - wy1;
    float ox = 0.0f, oy = 0.0f, od = 0.0f;
    const float* sd = src_depths + nv * hw;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const float fx = x0 + (float)(t & 1), fy = y0 + (float)(t >> 1);
        if (!(fx >= 0.0f && fx <= (float)(W - 1) && fy >= 0.0f && fy <= (float)(H - 1))) continue;   // zero padding / NaN
        const float wt = ((t & 1) ? wx1 : wx0) * ((t >> 1) ? wy1 : wy0);
        const int qx = (int)fx, qy = (int)fy;
        // srcs2ref_xyd at source pixel q (:87-91)
        const float ds = __ldg(sd + (int64_t)qy * W + qx);
        const V4 cr = transform_h(r_e, transform_h(s_einv, img2cam(s_kinv, (float)qx + 0.5f, (float)qy + 0.5f, ds)));
        const V3 ir = cam2img(r_k, cr);
        ox += ir.x * wt; oy += ir.y * wt; od += cr.z * wt;
    }
    float* out = reproj_xyd + nv * 3 * hw + (int64_t)y * W + x;
    out[0] = ox; out[hw] = oy; out[2 * hw] = od;
    in_range[tid] = inr ? 1.0f : 0.0f;
}

// vis_filter (:101-109) + ave_fusion (:112-114): one thread per (n, y, x)
//   masks [n,v,h,w] (1/0 as float, like the reference's), mask [n,h,w] (1/0), ave [n,h,w]
__device__ __forceinline__ void filter_thread(const float* __restrict__ ref_depth, const float* __restrict__ reproj_xyd,
                                              const float* __restrict__ in_range, float img_dist_thresh, float depth_thresh,
                                              float vthresh, float* __restrict__ masks, float* __restrict__ mask,
                                              float* __restrict__ ave, int N, int V, int H, int W, int64_t tid) {
    const int64_t hw = (int64_t)H * W;
    if (tid >= (int64_t)N * hw) return;
    const int64_t n = tid / hw, pix = tid % hw;
    const float px = (float)(pix % W) + 0.5f, py = (float)(pix / W) + 0.5f;
    const float dr = __ldg(ref_depth + tid);
    float count = 0.0f, dsum = 0.0f;
    for (int v = 0; v < V; ++v) {
        const int64_t nv = n * V + v;
        const float* r = reproj_xyd + nv * 3 * hw + pix;
        const float rx = __ldg(r), ry = __ldg(r + hw), rd = __ldg(r + 2 * hw);
        const float ex = rx - px, ey = ry - py;
        const bool dist_ok = sqrtf(ex * ex + ey * ey) < img_dist_thresh;
        const bool depth_ok = fabsf(dr - rd) < fmaxf(dr, rd) * depth_thresh;
        // bin_op_reduce(min) of three 0/1 images
        const float mk = fminf(fminf(__ldg(in_range + nv * hw + pix), dist_ok ? 1.0f : 0.0f), depth_ok ? 1.0f : 0.0f);
        masks[nv * hw + pix] = mk;
        count += mk;
        dsum += rd * mk;
    }
    mask[tid] = (count >= vthresh - 1.1f) ? 1.0f : 0.0f;
    ave[tid] = (dsum + dr) / (count + 1.0f);
}

// World points of a depth map (test.py:433-435: idx_img2cam + idx_cam2world): depth [n,h,w], mats [n,25] = Kinv (9)
// Einv (16) -> points [n,3,h,w]; one thread per (n, y, x)
__device__ __forceinline__ void points_thread(const float* __restrict__ depth, const float* __restrict__ mats,
                                              float* __restrict__ points, int N, int H, int W, int64_t tid) {
    const int64_t hw = (int64_t)H * W;
    if (tid >= (int64_t)N * hw) return;
    const int64_t n = tid / hw, pix = tid % hw;
    const float* m = mats + n * 25;
    const V4 p = transform_h(m + 9, img2cam(m, (float)(pix % W) + 0.5f, (float)(pix / W) + 0.5f, __ldg(depth + tid)));
    float* out = points + n * 3 * hw + pix;
    out[0] = p.x; out[hw] = p.y; out[2 * hw] = p.z;
}

// prob_filter (:69-76): AND over channels of prob[:, i] > thresh[i]; prob [n,c,h,w] -> mask [n,h,w] (1/0)
#define MVS_FUSION_MAX_PROB_CHANNELS 8
struct ProbThresh { float t[MVS_FUSION_MAX_PROB_CHANNELS]; };
__device__ __forceinline__ void prob_filter_thread(const float* __restrict__ prob, ProbThresh th, int nth,
                                                   float* __restrict__ mask, int N, int C, int64_t hw, int64_t tid) {
    if (tid >= (int64_t)N * hw) return;
    const int64_t n = tid / hw, pix = tid % hw;
    bool ok = true;
    for (int i = 0; i < nth; ++i) ok = ok && (__ldg(prob + (n * C + i) * hw + pix) > th.t[i]);
    mask[tid] = ok ? 1.0f : 0.0f;
}

// get_reproj_dynamic (misc/fusion.py:116-152): the reference pixel is projected into the source view, the source
// DEPTH is bilinearly sampled there (grid normalised with (w-1)/2, i.e. the projected coordinate itself is the sample
// index), and the sampled surface point is re-expressed in the reference view.  -> reproj_xyd [n,v,3,h,w]
// one thread per (n, v, y, x)
__device__ __forceinline__ void reproject_dynamic_thread(const float* __restrict__ ref_depth, const float* __restrict__ src_depths,
                                                         const float* __restrict__ mats, float* __restrict__ reproj_xyd, int N,
                                                         int V, int H, int W, int64_t tid) {
    const int64_t hw = (int64_t)H * W;
    if (tid >= (int64_t)N * V * hw) return;
    const int x = (int)(tid % W), y = (int)((tid / W) % H);
    const int64_t nv = tid / hw;
    const int64_t n = nv / V;
    const float* m = mats + nv * MVS_FUSION_MAT_FLOATS;
    const float *r_kinv = m, *r_einv = m + 9, *r_e = m + 25, *r_k = m + 41;
    const float *s_kinv = m + 50, *s_einv = m + 59, *s_e = m + 75, *s_k = m + 91;
    const float dr = __ldg(ref_depth + n * hw + (int64_t)y * W + x);
    const V3 is = cam2img(s_k, transform_h(s_e, transform_h(r_einv, img2cam(r_kinv, (float)x + 0.5f, (float)y + 0.5f, dr))));
    // :133-138  normalise with (size-1)/2, grid_sample(align_corners=True) un-normalises with the same factor
    const float gx = is.x / ((float)(W - 1) / 2.0f) - 1.0f, gy = is.y / ((float)(H - 1) / 2.0f) - 1.0f;
    const float ix = (gx + 1.0f) * 0.5f * (float)(W - 1), iy = (gy + 1.0f) * 0.5f * (float)(H - 1);
    const float x0 = floorf(ix), y0 = floorf(iy);
    const float wx1 = ix - x0, wx0 = 1.0f - wx1, wy1 = iy - y0, wy0 = 1.0f - wy1;
    const float* sd = src_depths + nv * hw;
    float ds = 0.0f;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const float fx = x0 + (float)(t & 1), fy = y0 + (float)(t >> 1);
        if (!(fx >= 0.0f && fx <= (float)(W - 1) && fy >= 0.0f && fy <= (float)(H - 1))) continue;
        ds += __ldg(sd + (int64_t)fy * W + (int64_t)fx) * (((t & 1) ? wx1 : wx0) * ((t >> 1) ? wy1 : wy0));
    }
    // :140-149  lift the projected coordinate with the sampled depth, back into the reference view
    const V4 cr = transform_h(r_e, transform_h(s_einv, img2cam(s_kinv, is.x, is.y, ds)));
    const V3 ir = cam2img(r_k, cr);
    float* out = reproj_xyd + nv * 3 * hw + (int64_t)y * W + x;
    out[0] = ir.x; out[hw] = ir.y; out[2 * hw] = cr.z;
}

// vis_filter_dynamic (:155-168) + the consistency vote and averaging of test.py:502-511: a view is consistent at level
// k (k = 2..v) if its reprojection error is < k / dist_base pixels and its relative depth error < k / rel_diff_base;
// a pixel is kept if, for some k, at least k views are consistent at level k; the depth is averaged over the views
// consistent at the loosest level (k = v).
//   vis_mask [n,v,h,w] (level v), geo_mask [n,h,w], ave [n,h,w], level_counts [n,v-1,h,w] (views consistent at each level)
// one thread per (n, y, x)
#define MVS_FUSION_MAX_VIEWS 16
__device__ __forceinline__ void filter_dynamic_thread(const float* __restrict__ ref_depth, const float* __restrict__ reproj_xyd,
                                                      float dist_base, float rel_diff_base, float* __restrict__ vis_mask,
                                                      float* __restrict__ geo_mask, float* __restrict__ ave,
                                                      float* __restrict__ level_counts, int N, int V, int H, int W, int64_t tid) {
    const int64_t hw = (int64_t)H * W;
    if (tid >= (int64_t)N * hw) return;
    const int64_t n = tid / hw, pix = tid % hw;
    const float px = (float)(pix % W) + 0.5f, py = (float)(pix / W) + 0.5f;
    const float dr = __ldg(ref_depth + tid);
    float cd[MVS_FUSION_MAX_VIEWS], dd[MVS_FUSION_MAX_VIEWS];
    float dsum = 0.0f, cnt_last = 0.0f;
    for (int v = 0; v < V; ++v) {
        const float* r = reproj_xyd + (n * V + v) * 3 * hw + pix;
        const float ex = __ldg(r) - px, ey = __ldg(r + hw) - py, rd = __ldg(r + 2 * hw);
        cd[v] = sqrtf(ex * ex + ey * ey);
        dd[v] = fabsf(dr - rd) / dr;
        const bool ok = (cd[v] < (float)V / dist_base) && (dd[v] < (float)V / rel_diff_base);     // level k = v
        vis_mask[(n * V + v) * hw + pix] = ok ? 1.0f : 0.0f;
        if (ok) { dsum += rd; cnt_last += 1.0f; }
    }
    bool keep = false;
    for (int k = 2; k <= V; ++k) {
        float cnt = 0.0f;
        for (int v = 0; v < V; ++v) cnt += ((cd[v] < (float)k / dist_base) && (dd[v] < (float)k / rel_diff_base)) ? 1.0f : 0.0f;
        if (level_counts) level_counts[(n * (V - 1) + (k - 2)) * hw + pix] = cnt;
        // test.py:509-511: for i in range(2, dy_range) with dy_range = v + 1
        keep = keep || (cnt >= (float)k);
    }
    geo_mask[tid] = keep ? 1.0f : 0.0f;
    ave[tid] = (dsum + dr) / (cnt_last + 1.0f);
}

struct Reproject {
    const float *ref_depth, *src_depths, *mats; float *reproj_xyd, *in_range; int N, V, H, W;
    __device__ __forceinline__ void operator()(int64_t tid, int64_t) const { reproject_thread(ref_depth, src_depths, mats, reproj_xyd, in_range, N, V, H, W, tid); }
};
struct Filter {
    const float *ref_depth, *reproj_xyd, *in_range; float img_dist_thresh, depth_thresh, vthresh; float *masks, *mask, *ave; int N, V, H, W;
    __device__ __forceinline__ void operator()(int64_t tid, int64_t) const { filter_thread(ref_depth, reproj_xyd, in_range, img_dist_thresh, depth_thresh, vthresh, masks, mask, ave, N, V, H, W, tid); }
};
struct ReprojectDynamic {
    const float *ref_depth, *src_depths, *mats; float* reproj_xyd; int N, V, H, W;
    __device__ __forceinline__ void operator()(int64_t tid, int64_t) const { reproject_dynamic_thread(ref_depth, src_depths, mats, reproj_xyd, N, V, H, W, tid); }
};
struct FilterDynamic {
    const float *ref_depth, *reproj_xyd; float dist_base, rel_diff_base; float *vis_mask, *geo_mask, *ave, *level_counts; int N, V, H, W;
    __device__ __forceinline__ void operator()(int64_t tid, int64_t) const { filter_dynamic_thread(ref_depth, reproj_xyd, dist_base, rel_diff_base, vis_mask, geo_mask, ave, level_counts, N, V, H, W, tid); }
};
struct Points {
    const float *depth, *mats; float* points; int N, H, W;
    __device__ __forceinline__ void operator()(int64_t tid, int64_t) const { points_thread(depth, mats, points, N, H, W, tid); }
};
struct ProbFilter {
    const float* prob; ProbThresh th; int nth; float* mask; int N, C; int64_t hw;
    __device__ __forceinline__ void operator()(int64_t tid, int64_t) const { prob_filter_thread(prob, th, nth, mask, N, C, hw, tid); }
};

}  // namespace fusion
}  // namespace mvs
