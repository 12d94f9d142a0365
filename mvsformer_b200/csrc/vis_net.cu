// vis_net.cu — A4: StageNet.vis, the per-view visibility network
// (models/mvsformer_model.py:37,91; ConvBnReLU = models/module.py:168-197):
//   entropy[1] -conv3x3+BN+ReLU-> 16 -conv3x3+BN+ReLU-> 16 -conv3x3+BN+ReLU-> 8 -conv1x1+bias-> 1 -sigmoid.
// One fused kernel: a CTA produces a 16x16 tile of weights; the 22x22 entropy tile (3-pixel
// halo) and the two 16-channel intermediate tiles stay in shared memory, so HBM sees one read
// and one write of an [M,H,W] map.  BN is folded by the caller.  The 3.6k folded weights travel
// as a kernel parameter (constant bank): every FMA takes its weight straight from c[0][..], no
// shared-memory or register traffic for weights; the kernel is FP32-FMA bound
// (3.6 kMAC per pixel + halo recompute).
#include "common.cuh"

namespace mvs {

struct VisParams {
    float w1[16][9];
    float b1[16];
    float w2[16][16][9];   // [co][ci][tap]
    float b2[16];
    float w3[8][16][9];
    float b3[8];
    float w4[8];
    float b4;
};
static_assert(sizeof(VisParams) == MVS_VIS_PARAM_FLOATS * sizeof(float), "VisParams packing");

constexpr int VT = 16;            // output tile
constexpr int VE = VT + 6;        // entropy tile 22
constexpr int V1 = VT + 4;        // layer-1 tile 20
constexpr int V2 = VT + 2;        // layer-2 tile 18
constexpr int VIS_THREADS = 352;  // >= V2*V2 = 324 so layer 2 (the heaviest) is a single round

__global__ void __launch_bounds__(VIS_THREADS)
vis_net_kernel(const float* __restrict__ entropy, float* __restrict__ weight, int H, int W, const __grid_constant__ VisParams P) {
    __shared__ float s_e[VE * VE];
    __shared__ float s_1[16][V1 * V1];
    __shared__ float s_2[16][V2 * V2];
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * VT, y0 = blockIdx.y * VT;
    const int64_t plane = (int64_t)blockIdx.z * H * W;

    for (int i = tid; i < VE * VE; i += VIS_THREADS) {
        const int gy = y0 - 3 + i / VE, gx = x0 - 3 + i % VE;
        s_e[i] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? __ldg(entropy + plane + (int64_t)gy * W + gx) : 0.0f;
    }
    __syncthreads();

    // layer 1: 1 -> 16 on the 20x20 region.  Positions outside the image must hold ZERO (they
    // are the zero padding seen by layer 2), not conv-of-padding.
    for (int i = tid; i < V1 * V1; i += VIS_THREADS) {
        const int ly = i / V1, lx = i % V1;
        const int gy = y0 - 2 + ly, gx = x0 - 2 + lx;
        const bool inside = (gy >= 0 && gy < H && gx >= 0 && gx < W);
        float in[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) in[t] = s_e[(ly + t / 3) * VE + lx + t % 3];
#pragma unroll
        for (int co = 0; co < 16; ++co) {
            float a = P.b1[co];
#pragma unroll
            for (int t = 0; t < 9; ++t) a = fmaf(in[t], P.w1[co][t], a);
            s_1[co][i] = inside ? fmaxf(a, 0.0f) : 0.0f;
        }
    }
    __syncthreads();

    // layer 2: 16 -> 16 on the 18x18 region
    if (tid < V2 * V2) {
        const int ly = tid / V2, lx = tid % V2;
        const int gy = y0 - 1 + ly, gx = x0 - 1 + lx;
        const bool inside = (gy >= 0 && gy < H && gx >= 0 && gx < W);
        float acc[16];
#pragma unroll
        for (int co = 0; co < 16; ++co) acc[co] = P.b2[co];
#pragma unroll
        for (int ci = 0; ci < 16; ++ci) {
            float in[9];
#pragma unroll
            for (int t = 0; t < 9; ++t) in[t] = s_1[ci][(ly + t / 3) * V1 + lx + t % 3];
#pragma unroll
            for (int co = 0; co < 16; ++co) {
#pragma unroll
                for (int t = 0; t < 9; ++t) acc[co] = fmaf(in[t], P.w2[co][ci][t], acc[co]);
            }
        }
#pragma unroll
        for (int co = 0; co < 16; ++co) s_2[co][tid] = inside ? fmaxf(acc[co], 0.0f) : 0.0f;
    }
    __syncthreads();

    // layer 3 (16 -> 8) + 1x1 (8 -> 1) + sigmoid on the 16x16 tile
    if (tid < VT * VT) {
        const int ly = tid / VT, lx = tid % VT;
        const int gy = y0 + ly, gx = x0 + lx;
        float acc[8];
#pragma unroll
        for (int co = 0; co < 8; ++co) acc[co] = P.b3[co];
#pragma unroll
        for (int ci = 0; ci < 16; ++ci) {
            float in[9];
#pragma unroll
            for (int t = 0; t < 9; ++t) in[t] = s_2[ci][(ly + t / 3) * V2 + lx + t % 3];
#pragma unroll
            for (int co = 0; co < 8; ++co) {
#pragma unroll
                for (int t = 0; t < 9; ++t) acc[co] = fmaf(in[t], P.w3[co][ci][t], acc[co]);
            }
        }
        float z = P.b4;
#pragma unroll
        for (int co = 0; co < 8; ++co) z = fmaf(fmaxf(acc[co], 0.0f), P.w4[co], z);
        if (gy < H && gx < W) weight[plane + (int64_t)gy * W + gx] = 1.0f / (1.0f + expf(-z));
    }
}

}  // namespace mvs

extern "C" int mvs_vis_weight(const float* entropy, const float* params_host, float* weight, int M, int H, int W,
                              void* stream) {
    MVS_REQUIRE(entropy && params_host && weight, "mvs_vis_weight: null pointer");
    MVS_REQUIRE(M >= 1 && H >= 1 && W >= 1, "mvs_vis_weight: empty shape M=%d H=%d W=%d", M, H, W);
    MVS_REQUIRE(M <= 65535, "mvs_vis_weight: at most 65535 maps per call (got %d)", M);
    mvs::VisParams P;
    memcpy(&P, params_host, sizeof(P));
    dim3 grid(mvs::cdiv(W, mvs::VT), mvs::cdiv(H, mvs::VT), M);
    mvs::vis_net_kernel<<<grid, mvs::VIS_THREADS, 0, (cudaStream_t)stream>>>(entropy, weight, H, W, P);
    MVS_LAUNCH_OK("vis_net_kernel");
    return MVS_OK;
}
