// warp.cu — A2: materialising homography warp (homo_warping_3D / homo_warping_3D_with_mask,
// models/warping.py:69-109,155-189).  API-completeness kernel: StageNet never calls it (the
// cost-volume kernels sample on the fly and never write the warped tensor); it exists so the
// reference's warping functions remain available with identical results.
//
// One thread per (d, y, x); lanes run along x so the C-strided output stores and the depth /
// mask accesses are coalesced; the source taps of neighbouring lanes are neighbouring texels.
// Bound: HBM write of B*C*D*H*W*4 bytes.
#include "common.cuh"

namespace mvs {

__global__ void __launch_bounds__(256)
homo_warp_kernel(const float* __restrict__ src, const float* __restrict__ relproj, const float* __restrict__ depth,
                 int depth_is_map, float* __restrict__ warped, uint8_t* __restrict__ mask, int C, int D, int H, int W) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int b = blockIdx.z / D, d = blockIdx.z % D;
    if (x >= W || y >= H) return;
    const RelProj m = load_relproj(relproj + (int64_t)b * 12);
    const PixelRay ray = pixel_ray(m, (float)x, (float)y);
    const int64_t hw = (int64_t)H * W;
    const float dep = depth_is_map ? __ldg(depth + ((int64_t)b * D + d) * hw + (int64_t)y * W + x)
                                   : __ldg(depth + (int64_t)b * D + d);
    float gx, gy, qz;
    const Taps t = make_taps(m, ray, dep, H, W, (float)((W - 1) / 2.0), (float)((H - 1) / 2.0), &gx, &gy, &qz);
    if (mask) {
        // warping.py:99-103 (NaN compares false, as in torch)
        const bool oob = (gx > 1.0f) || (gx < -1.0f) || (gy > 1.0f) || (gy < -1.0f) || (qz <= 0.0f);
        mask[((int64_t)b * D + d) * hw + (int64_t)y * W + x] = oob ? 1 : 0;
    }
    const float* sp = src + (int64_t)b * C * hw;
    float* op = warped + (((int64_t)b * C) * D + d) * hw + (int64_t)y * W + x;
    for (int c = 0; c < C; ++c) {
        op[(int64_t)c * D * hw] = sample4(sp + (int64_t)c * hw, t);
    }
}

}  // namespace mvs

extern "C" int mvs_homo_warp(const float* src_fea, const float* relproj, const float* depth, int depth_is_map,
                             float* warped, uint8_t* mask, int B, int C, int D, int H, int W, void* stream) {
    MVS_REQUIRE(src_fea && relproj && depth && warped, "mvs_homo_warp: null pointer");
    MVS_REQUIRE(B >= 1 && C >= 1 && D >= 1 && H >= 1 && W >= 1, "mvs_homo_warp: empty shape B=%d C=%d D=%d H=%d W=%d", B, C, D, H, W);
    MVS_REQUIRE((int64_t)B * D <= 65535, "mvs_homo_warp: B*D = %lld exceeds 65535", (long long)B * D);
    dim3 block(32, 8);
    dim3 grid(mvs::cdiv(W, 32), mvs::cdiv(H, 8), B * D);
    mvs::homo_warp_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(src_fea, relproj, depth, depth_is_map, warped, mask,
                                                                    C, D, H, W);
    MVS_LAUNCH_OK("homo_warp_kernel");
    return MVS_OK;
}
