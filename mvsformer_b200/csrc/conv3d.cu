// conv3d.cu — A7: the 3D-CNN regulariser layers (models/module.py:83-159 blocks used by
// CostRegNet :469-505, CostRegNet2D :508-547, CostRegNet3D :550-594) on channels-last
// activations [B,D,H,W,C], FP32 CUDA-core path.
//
//   y = act(conv(x) + shift) (+ skip)      BN folded by the caller, skip added after the ReLU
//
// conv3d_cl_kernel:   direct convolution, kernel (kd,3,3), stride 1 or 2 per axis.  A thread owns
//   VPT output voxels x COUT_T output channels in registers; per tap it reads the Cin-contiguous
//   input vector with 128-bit loads (lanes = consecutive x -> fully coalesced) and streams the
//   [Cin][COUT_T] weight slice from shared memory with broadcast 128-bit loads.
// deconv3d_cl_kernel: ConvTranspose3d kernel (kd,3,3), stride (sd,2,2), pad k/2, output_padding
//   stride-1, in gather form: a thread owns the output pair (2j, 2j+1) of one output row, for
//   which the contributing taps are fixed (kw=1 | kw=2 and kw=0), so no work is predicated off.
//
// These kernels are FP32-FMA bound (AI of a 3x3x3 16->16 layer = 432 flop/B) and use the packed FFMA2 form (a three-register
// FFMA holds the FMA pipe for two issue cycles per warp on sm_100; fma.rn.f32x2 does two IEEE fmas in that slot); the tcgen05
// implicit-GEMM path (conv3d_tc.cu) supersedes them for the heavy layers.
#include "common.cuh"

namespace mvs {

struct ConvDims {
    int B, D, H, W;        // input
    int Do, Ho, Wo;        // output
    int Cin, Cout;
    int kd, sd, sh, sw;
    int relu;
};

constexpr int CONV_THREADS = 128;

template <int CIN_T, int COUT_T, int VPT>
__global__ void __launch_bounds__(CONV_THREADS)
conv3d_cl_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ shift,
                 const float* __restrict__ skip, float* __restrict__ y, ConvDims d) {
    extern __shared__ __align__(16) float s_w[];   // [ntaps][Cin][COUT_T]
    const int cin = CIN_T > 0 ? CIN_T : d.Cin;
    const int ntaps = d.kd * 9;
    const int co0 = blockIdx.y * COUT_T;
    for (int i = threadIdx.x; i < ntaps * cin * (COUT_T / 4); i += CONV_THREADS) {
        const int row = i / (COUT_T / 4), q = i % (COUT_T / 4);
        reinterpret_cast<float4*>(s_w)[i] = __ldg(reinterpret_cast<const float4*>(w + (int64_t)row * d.Cout + co0) + q);
    }
    __syncthreads();

    const int64_t total = (int64_t)d.B * d.Do * d.Ho * d.Wo;
    const int64_t base = (int64_t)blockIdx.x * (CONV_THREADS * VPT) + threadIdx.x;
    int ox[VPT], oy[VPT], oz[VPT], ob[VPT];
    bool live[VPT];
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
        int64_t idx = base + (int64_t)j * CONV_THREADS;
        live[j] = idx < total;
        if (!live[j]) idx = 0;
        ox[j] = (int)(idx % d.Wo); idx /= d.Wo;
        oy[j] = (int)(idx % d.Ho); idx /= d.Ho;
        oz[j] = (int)(idx % d.Do); ob[j] = (int)(idx / d.Do);
    }
    float acc[VPT][COUT_T];
#pragma unroll
    for (int j = 0; j < VPT; ++j)
#pragma unroll
        for (int c = 0; c < COUT_T; ++c) acc[j][c] = 0.0f;

    const int pd = d.kd / 2;
    for (int kz = 0; kz < d.kd; ++kz) {
        for (int ky = 0; ky < 3; ++ky) {
#pragma unroll 1
            for (int kx = 0; kx < 3; ++kx) {
                const float* wt = s_w + (size_t)((kz * 3 + ky) * 3 + kx) * cin * COUT_T;
                const float* ptr[VPT];
#pragma unroll
                for (int j = 0; j < VPT; ++j) {
                    const int iz = oz[j] * d.sd - pd + kz, iy = oy[j] * d.sh - 1 + ky, ix = ox[j] * d.sw - 1 + kx;
                    const bool ok = live[j] && iz >= 0 && iz < d.D && iy >= 0 && iy < d.H && ix >= 0 && ix < d.W;
                    ptr[j] = ok ? x + ((((int64_t)ob[j] * d.D + iz) * d.H + iy) * d.W + ix) * cin : nullptr;
                }
#pragma unroll 2
                for (int c4 = 0; c4 < cin / 4; ++c4) {
                    float4 v[VPT];
#pragma unroll
                    for (int j = 0; j < VPT; ++j)
                        v[j] = ptr[j] ? __ldg(reinterpret_cast<const float4*>(ptr[j]) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const float4* wrow = reinterpret_cast<const float4*>(wt + (size_t)(c4 * 4 + u) * COUT_T);
#pragma unroll
                        for (int q = 0; q < COUT_T / 4; ++q) {
                            const float4 ww = wrow[q];
#pragma unroll
                            for (int j = 0; j < VPT; ++j) {
                                const float a = u == 0 ? v[j].x : (u == 1 ? v[j].y : (u == 2 ? v[j].z : v[j].w));
                                fma4_bcast(&acc[j][q * 4], a, ww);          // two FFMA2 (packed FP32) instead of four FFMA
                            }
                        }
                    }
                }
            }
        }
    }

#pragma unroll
    for (int j = 0; j < VPT; ++j) {
        if (!live[j]) continue;
        const int64_t o = ((((int64_t)ob[j] * d.Do + oz[j]) * d.Ho + oy[j]) * d.Wo + ox[j]) * d.Cout + co0;
#pragma unroll
        for (int q = 0; q < COUT_T / 4; ++q) {
            float4 r = make_float4(acc[j][q * 4], acc[j][q * 4 + 1], acc[j][q * 4 + 2], acc[j][q * 4 + 3]);
            if (shift) {
                const float4 s = __ldg(reinterpret_cast<const float4*>(shift + co0) + q);
                r.x += s.x; r.y += s.y; r.z += s.z; r.w += s.w;
            }
            if (d.relu) { r.x = fmaxf(r.x, 0.f); r.y = fmaxf(r.y, 0.f); r.z = fmaxf(r.z, 0.f); r.w = fmaxf(r.w, 0.f); }
            if (skip) {
                const float4 s = __ldg(reinterpret_cast<const float4*>(skip + o) + q);
                r.x += s.x; r.y += s.y; r.z += s.z; r.w += s.w;
            }
            reinterpret_cast<float4*>(y + o)[q] = r;
        }
    }
}

// Which input index feeds output index o through kernel tap k of a transposed convolution
// (stride s, padding pad):  o = i*s - pad + k.
__device__ __forceinline__ bool deconv_src(int o, int k, int s, int pad, int n_in, int* i) {
    const int t = o + pad - k;
    if (t < 0 || (t % s) != 0) return false;
    *i = t / s;
    return *i < n_in;
}

template <int CIN_T, int COUT_T>
__global__ void __launch_bounds__(CONV_THREADS)
deconv3d_cl_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ shift,
                   const float* __restrict__ skip, float* __restrict__ y, ConvDims d) {
    extern __shared__ __align__(16) float s_w[];   // [kd*9][Cin][COUT_T]
    const int cin = CIN_T > 0 ? CIN_T : d.Cin;
    const int ntaps = d.kd * 9;
    const int co0 = blockIdx.y * COUT_T;
    for (int i = threadIdx.x; i < ntaps * cin * (COUT_T / 4); i += CONV_THREADS) {
        const int row = i / (COUT_T / 4), q = i % (COUT_T / 4);
        reinterpret_cast<float4*>(s_w)[i] = __ldg(reinterpret_cast<const float4*>(w + (int64_t)row * d.Cout + co0) + q);
    }
    __syncthreads();

    // one thread per (b, zo, yo, j): outputs x = 2j and 2j+1
    const int64_t total = (int64_t)d.B * d.Do * d.Ho * d.W;
    int64_t idx = (int64_t)blockIdx.x * CONV_THREADS + threadIdx.x;
    if (idx >= total) return;
    const int j = (int)(idx % d.W); idx /= d.W;
    const int yo = (int)(idx % d.Ho); idx /= d.Ho;
    const int zo = (int)(idx % d.Do);
    const int b = (int)(idx / d.Do);

    float acc0[COUT_T], acc1[COUT_T];
#pragma unroll
    for (int c = 0; c < COUT_T; ++c) { acc0[c] = 0.0f; acc1[c] = 0.0f; }
    const int pd = d.kd / 2;
    const bool has_next = (j + 1) < d.W;

    for (int kz = 0; kz < d.kd; ++kz) {
        int iz;
        if (!deconv_src(zo, kz, d.sd, pd, d.D, &iz)) continue;
        for (int ky = 0; ky < 3; ++ky) {
            int iy;
            if (!deconv_src(yo, ky, 2, 1, d.H, &iy)) continue;
            const float* pa = x + ((((int64_t)b * d.D + iz) * d.H + iy) * d.W + j) * cin;
            const float* w0 = s_w + (size_t)((kz * 3 + ky) * 3 + 0) * cin * COUT_T;   // kw = 0: in[j+1] -> out 2j+1
            const float* w1 = w0 + (size_t)cin * COUT_T;                               // kw = 1: in[j]   -> out 2j
            const float* w2 = w1 + (size_t)cin * COUT_T;                               // kw = 2: in[j]   -> out 2j+1
#pragma unroll 2
            for (int c4 = 0; c4 < cin / 4; ++c4) {
                const float4 va = __ldg(reinterpret_cast<const float4*>(pa) + c4);
                const float4 vb = has_next ? __ldg(reinterpret_cast<const float4*>(pa + cin) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float a = u == 0 ? va.x : (u == 1 ? va.y : (u == 2 ? va.z : va.w));
                    const float bn = u == 0 ? vb.x : (u == 1 ? vb.y : (u == 2 ? vb.z : vb.w));
                    const size_t ro = (size_t)(c4 * 4 + u) * COUT_T;
#pragma unroll
                    for (int q = 0; q < COUT_T / 4; ++q) {
                        const float4 k0 = reinterpret_cast<const float4*>(w0 + ro)[q];
                        const float4 k1 = reinterpret_cast<const float4*>(w1 + ro)[q];
                        const float4 k2 = reinterpret_cast<const float4*>(w2 + ro)[q];
                        fma4_bcast(&acc0[q * 4], a, k1);                    // packed FP32: two FFMA2 per four channels
                        fma4_bcast(&acc1[q * 4], bn, k0);                   // (same order as fmaf(a, k2, fmaf(bn, k0, acc1)))
                        fma4_bcast(&acc1[q * 4], a, k2);
                    }
                }
            }
        }
    }

    const int64_t o0 = ((((int64_t)b * d.Do + zo) * d.Ho + yo) * d.Wo + 2 * j) * d.Cout + co0;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int64_t o = o0 + (int64_t)half * d.Cout;
        const float* acc = half ? acc1 : acc0;
#pragma unroll
        for (int q = 0; q < COUT_T / 4; ++q) {
            float4 r = make_float4(acc[q * 4], acc[q * 4 + 1], acc[q * 4 + 2], acc[q * 4 + 3]);
            if (shift) {
                const float4 s = __ldg(reinterpret_cast<const float4*>(shift + co0) + q);
                r.x += s.x; r.y += s.y; r.z += s.z; r.w += s.w;
            }
            if (d.relu) { r.x = fmaxf(r.x, 0.f); r.y = fmaxf(r.y, 0.f); r.z = fmaxf(r.z, 0.f); r.w = fmaxf(r.w, 0.f); }
            if (skip) {
                const float4 s = __ldg(reinterpret_cast<const float4*>(skip + o) + q);
                r.x += s.x; r.y += s.y; r.z += s.z; r.w += s.w;
            }
            reinterpret_cast<float4*>(y + o)[q] = r;
        }
    }
}

// ------------------------------------------------------------------------------------------
// layout transforms at the module boundary
// ------------------------------------------------------------------------------------------
constexpr int TR_VOX = 64;

__global__ void __launch_bounds__(256)
ncdhw_to_cl_kernel(const float* __restrict__ x, float* __restrict__ y, int C, int64_t S, int round_tf32) {
    extern __shared__ float s_t[];   // [C][TR_VOX + 1]
    const int b = blockIdx.y;
    const int64_t s0 = (int64_t)blockIdx.x * TR_VOX;
    const int n = (int)min((int64_t)TR_VOX, S - s0);
    for (int i = threadIdx.x; i < C * TR_VOX; i += blockDim.x) {
        const int c = i / TR_VOX, s = i % TR_VOX;
        if (s < n) s_t[c * (TR_VOX + 1) + s] = __ldg(x + ((int64_t)b * C + c) * S + s0 + s);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n * C; i += blockDim.x) {
        const int s = i / C, c = i % C;
        float v = s_t[c * (TR_VOX + 1) + s];
        if (round_tf32) v = round_to_tf32(v);
        y[((int64_t)b * S + s0 + s) * C + c] = v;
    }
}

__global__ void __launch_bounds__(256)
cl_to_ncdhw_kernel(const float* __restrict__ x, float* __restrict__ y, int C, int64_t S) {
    extern __shared__ float s_t[];
    const int b = blockIdx.y;
    const int64_t s0 = (int64_t)blockIdx.x * TR_VOX;
    const int n = (int)min((int64_t)TR_VOX, S - s0);
    for (int i = threadIdx.x; i < n * C; i += blockDim.x) {
        const int s = i / C, c = i % C;
        s_t[c * (TR_VOX + 1) + s] = __ldg(x + ((int64_t)b * S + s0 + s) * C + c);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C * TR_VOX; i += blockDim.x) {
        const int c = i / TR_VOX, s = i % TR_VOX;
        if (s < n) y[((int64_t)b * C + c) * S + s0 + s] = s_t[c * (TR_VOX + 1) + s];
    }
}

template <int CIN_T, int COUT_T>
static int launch_conv(const float* x, const float* w, const float* shift, const float* skip, float* y, const ConvDims& d,
                       cudaStream_t st) {
    constexpr int VPT = 2;
    const size_t smem = (size_t)d.kd * 9 * d.Cin * COUT_T * sizeof(float);
    MVS_REQUIRE(smem <= 227 * 1024, "mvs_conv3d_cl: weight slice of %zu bytes exceeds shared memory", smem);
    auto kern = conv3d_cl_kernel<CIN_T, COUT_T, VPT>;
    MVS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t total = (int64_t)d.B * d.Do * d.Ho * d.Wo;
    dim3 grid(cdiv(total, CONV_THREADS * VPT), d.Cout / COUT_T);
    kern<<<grid, CONV_THREADS, smem, st>>>(x, w, shift, skip, y, d);
    MVS_LAUNCH_OK("conv3d_cl_kernel");
    return MVS_OK;
}

template <int CIN_T, int COUT_T>
static int launch_deconv(const float* x, const float* w, const float* shift, const float* skip, float* y,
                         const ConvDims& d, cudaStream_t st) {
    const size_t smem = (size_t)d.kd * 9 * d.Cin * COUT_T * sizeof(float);
    MVS_REQUIRE(smem <= 227 * 1024, "mvs_deconv3d_cl: weight slice of %zu bytes exceeds shared memory", smem);
    auto kern = deconv3d_cl_kernel<CIN_T, COUT_T>;
    MVS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t total = (int64_t)d.B * d.Do * d.Ho * d.W;
    dim3 grid(cdiv(total, CONV_THREADS), d.Cout / COUT_T);
    kern<<<grid, CONV_THREADS, smem, st>>>(x, w, shift, skip, y, d);
    MVS_LAUNCH_OK("deconv3d_cl_kernel");
    return MVS_OK;
}

#define MVS_DISPATCH_CIN(LAUNCH, COUT_T)                                            \
    switch (d.Cin) {                                                                \
        case 8: return LAUNCH<8, COUT_T>(x, w, shift, skip, y, d, st);              \
        case 16: return LAUNCH<16, COUT_T>(x, w, shift, skip, y, d, st);            \
        case 32: return LAUNCH<32, COUT_T>(x, w, shift, skip, y, d, st);            \
        case 64: return LAUNCH<64, COUT_T>(x, w, shift, skip, y, d, st);            \
        default: return LAUNCH<0, COUT_T>(x, w, shift, skip, y, d, st);             \
    }

static int check_conv_args(const char* fn, const float* x, const float* w, float* y, int B, int D, int H, int W, int Cin,
                           int Cout, int kd) {
    MVS_REQUIRE(x && w && y, "%s: null pointer", fn);
    MVS_REQUIRE(B >= 1 && D >= 1 && H >= 1 && W >= 1, "%s: empty shape B=%d D=%d H=%d W=%d", fn, B, D, H, W);
    MVS_REQUIRE(kd == 1 || kd == 3, "%s: depth kernel size must be 1 or 3 (got %d)", fn, kd);
    MVS_REQUIRE(Cin >= 4 && Cin % 4 == 0, "%s: Cin must be a positive multiple of 4 (got %d)", fn, Cin);
    MVS_REQUIRE(Cout >= 8 && Cout % 8 == 0, "%s: Cout must be a positive multiple of 8 (got %d)", fn, Cout);
    return MVS_OK;
}

}  // namespace mvs

extern "C" int mvs_conv3d_cl(const float* x, const float* w, const float* shift, const float* skip, float* y, int B,
                             int D, int H, int W, int Cin, int Cout, int kd, int sd, int sh, int sw, int relu,
                             void* stream) {
    using namespace mvs;
    int rc = check_conv_args("mvs_conv3d_cl", x, w, y, B, D, H, W, Cin, Cout, kd);
    if (rc) return rc;
    MVS_REQUIRE((sd == 1 || sd == 2) && (sh == 1 || sh == 2) && (sw == 1 || sw == 2),
                "mvs_conv3d_cl: strides must be 1 or 2 (got %d,%d,%d)", sd, sh, sw);
    const int pd = kd / 2;
    ConvDims d{B, D, H, W, (D + 2 * pd - kd) / sd + 1, (H + 2 - 3) / sh + 1, (W + 2 - 3) / sw + 1, Cin, Cout, kd, sd, sh, sw, relu};
    cudaStream_t st = (cudaStream_t)stream;
    if (Cout % 16 == 0) { MVS_DISPATCH_CIN(launch_conv, 16) }
    MVS_DISPATCH_CIN(launch_conv, 8)
}

extern "C" int mvs_deconv3d_cl(const float* x, const float* w, const float* shift, const float* skip, float* y, int B,
                               int D, int H, int W, int Cin, int Cout, int kd, int sd, int relu, void* stream) {
    using namespace mvs;
    int rc = check_conv_args("mvs_deconv3d_cl", x, w, y, B, D, H, W, Cin, Cout, kd);
    if (rc) return rc;
    MVS_REQUIRE(sd == 1 || sd == 2, "mvs_deconv3d_cl: depth stride must be 1 or 2 (got %d)", sd);
    MVS_REQUIRE(!(kd == 1 && sd != 1), "mvs_deconv3d_cl: kd = 1 requires sd = 1");
    ConvDims d{B, D, H, W, D * sd, H * 2, W * 2, Cin, Cout, kd, sd, 2, 2, relu};
    cudaStream_t st = (cudaStream_t)stream;
    if (Cout % 16 == 0) { MVS_DISPATCH_CIN(launch_deconv, 16) }
    MVS_DISPATCH_CIN(launch_deconv, 8)
}

static int ncdhw_to_cl_impl(const float* x, float* y, int B, int C, int D, int H, int W, int round_tf32, void* stream) {
    using namespace mvs;
    MVS_REQUIRE(x && y, "mvs_ncdhw_to_cl: null pointer");
    MVS_REQUIRE(B >= 1 && C >= 1 && D >= 1 && H >= 1 && W >= 1 && B <= 65535, "mvs_ncdhw_to_cl: bad shape");
    const int64_t S = (int64_t)D * H * W;
    const size_t smem = (size_t)C * (TR_VOX + 1) * sizeof(float);
    MVS_REQUIRE(smem <= 48 * 1024, "mvs_ncdhw_to_cl: C = %d too large", C);
    ncdhw_to_cl_kernel<<<dim3(cdiv(S, TR_VOX), B), 256, smem, (cudaStream_t)stream>>>(x, y, C, S, round_tf32);
    MVS_LAUNCH_OK("ncdhw_to_cl_kernel");
    return MVS_OK;
}

extern "C" int mvs_ncdhw_to_cl(const float* x, float* y, int B, int C, int D, int H, int W, void* stream) {
    return ncdhw_to_cl_impl(x, y, B, C, D, H, W, 0, stream);
}

extern "C" int mvs_ncdhw_to_cl_tf32(const float* x, float* y, int B, int C, int D, int H, int W, void* stream) {
    return ncdhw_to_cl_impl(x, y, B, C, D, H, W, 1, stream);
}

extern "C" int mvs_cl_to_ncdhw(const float* x, float* y, int B, int C, int D, int H, int W, void* stream) {
    using namespace mvs;
    MVS_REQUIRE(x && y, "mvs_cl_to_ncdhw: null pointer");
    MVS_REQUIRE(B >= 1 && C >= 1 && D >= 1 && H >= 1 && W >= 1 && B <= 65535, "mvs_cl_to_ncdhw: bad shape");
    const int64_t S = (int64_t)D * H * W;
    const size_t smem = (size_t)C * (TR_VOX + 1) * sizeof(float);
    MVS_REQUIRE(smem <= 48 * 1024, "mvs_cl_to_ncdhw: C = %d too large", C);
    cl_to_ncdhw_kernel<<<dim3(cdiv(S, TR_VOX), B), 256, smem, (cudaStream_t)stream>>>(x, y, C, S);
    MVS_LAUNCH_OK("cl_to_ncdhw_kernel");
    return MVS_OK;
}
