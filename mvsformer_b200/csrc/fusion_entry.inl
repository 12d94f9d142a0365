// fusion_entry.inl — C-ABI entry points of the depth-map fusion kernels (include/mvs_b200.h, "depth-map fusion").
// Included by fusion.cu (CUDA) and tests/emu/emu.cpp (CPU emulation, tests only); the includer provides
// MVS_REQUIRE, MVS_OK, fusion_kernels.cuh and mvs::fusion::launch_flat (see train_entry.inl).

extern "C" int mvs_fusion_reproject(const float* ref_depth, const float* src_depths, const float* mats, float* reproj_xyd,
                                    float* in_range, int N, int V, int H, int W, void* stream) {
    using namespace mvs::fusion;
    MVS_REQUIRE(ref_depth && src_depths && mats && reproj_xyd && in_range, "mvs_fusion_reproject: null pointer");
    MVS_REQUIRE(N >= 1 && V >= 1 && H >= 1 && W >= 1, "mvs_fusion_reproject: empty shape");
    Reproject f{ref_depth, src_depths, mats, reproj_xyd, in_range, N, V, H, W};
    return launch_flat(f, (int64_t)N * V * H * W, stream, "fusion_reproject");
}

extern "C" int mvs_fusion_filter(const float* ref_depth, const float* reproj_xyd, const float* in_range, float img_dist_thresh,
                                 float depth_thresh, float vthresh, float* masks, float* mask, float* ave, int N, int V, int H,
                                 int W, void* stream) {
    using namespace mvs::fusion;
    MVS_REQUIRE(ref_depth && reproj_xyd && in_range && masks && mask && ave, "mvs_fusion_filter: null pointer");
    MVS_REQUIRE(N >= 1 && V >= 1 && H >= 1 && W >= 1, "mvs_fusion_filter: empty shape");
    Filter f{ref_depth, reproj_xyd, in_range, img_dist_thresh, depth_thresh, vthresh, masks, mask, ave, N, V, H, W};
    return launch_flat(f, (int64_t)N * H * W, stream, "fusion_filter");
}

extern "C" int mvs_fusion_points(const float* depth, const float* mats, float* points, int N, int H, int W, void* stream) {
    using namespace mvs::fusion;
    MVS_REQUIRE(depth && mats && points, "mvs_fusion_points: null pointer");
    MVS_REQUIRE(N >= 1 && H >= 1 && W >= 1, "mvs_fusion_points: empty shape");
    Points f{depth, mats, points, N, H, W};
    return launch_flat(f, (int64_t)N * H * W, stream, "fusion_points");
}

extern "C" int mvs_fusion_prob_filter(const float* prob, const float* thresh_host, int nthresh, float* mask, int N, int C,
                                      int H, int W, void* stream) {
    using namespace mvs::fusion;
    MVS_REQUIRE(prob && thresh_host && mask, "mvs_fusion_prob_filter: null pointer");
    MVS_REQUIRE(N >= 1 && C >= 1 && H >= 1 && W >= 1, "mvs_fusion_prob_filter: empty shape");
    MVS_REQUIRE(nthresh >= 1 && nthresh <= C && nthresh <= MVS_FUSION_MAX_PROB_CHANNELS,
                "mvs_fusion_prob_filter: need 1 <= thresholds (%d) <= channels (%d) and <= %d", nthresh, C,
                MVS_FUSION_MAX_PROB_CHANNELS);
    ProbThresh th;
    for (int i = 0; i < MVS_FUSION_MAX_PROB_CHANNELS; ++i) th.t[i] = i < nthresh ? thresh_host[i] : 0.0f;
    ProbFilter f{prob, th, nthresh, mask, N, C, (int64_t)H * W};
    return launch_flat(f, (int64_t)N * H * W, stream, "fusion_prob_filter");
}

extern "C" int mvs_fusion_reproject_dynamic(const float* ref_depth, const float* src_depths, const float* mats, float* reproj_xyd,
                                            int N, int V, int H, int W, void* stream) {
    using namespace mvs::fusion;
    MVS_REQUIRE(ref_depth && src_depths && mats && reproj_xyd, "mvs_fusion_reproject_dynamic: null pointer");
    MVS_REQUIRE(N >= 1 && V >= 1 && H >= 1 && W >= 1, "mvs_fusion_reproject_dynamic: empty shape");
    ReprojectDynamic f{ref_depth, src_depths, mats, reproj_xyd, N, V, H, W};
    return launch_flat(f, (int64_t)N * V * H * W, stream, "fusion_reproject_dynamic");
}

extern "C" int mvs_fusion_filter_dynamic(const float* ref_depth, const float* reproj_xyd, float dist_base, float rel_diff_base,
                                         float* vis_mask, float* geo_mask, float* ave, float* level_counts, int N, int V, int H,
                                         int W, void* stream) {
    using namespace mvs::fusion;
    MVS_REQUIRE(ref_depth && reproj_xyd && vis_mask && geo_mask && ave, "mvs_fusion_filter_dynamic: null pointer");
    MVS_REQUIRE(N >= 1 && V >= 1 && V <= MVS_FUSION_MAX_VIEWS && H >= 1 && W >= 1,
                "mvs_fusion_filter_dynamic: need 1 <= views <= %d", MVS_FUSION_MAX_VIEWS);
    FilterDynamic f{ref_depth, reproj_xyd, dist_base, rel_diff_base, vis_mask, geo_mask, ave, level_counts, N, V, H, W};
    return launch_flat(f, (int64_t)N * H * W, stream, "fusion_filter_dynamic");
}
