// conv3d_tc.cu — A7 on the 5th-gen tensor cores: implicit-GEMM 3D convolution of channels-last
// fp32 activations with tcgen05.mma (kind::tf32, fp32 accumulation in TMEM).
//
// GEMM view: M = 128 consecutive output voxels of one (b, z) plane in a *padded virtual index*
// j = y * PW + x (PW = Wo + 2 for stride 1, Wo + 1 for stride 2), N = Cout tile, K = taps x Cin.
// In that index space the input voxel of tap (kh, kw) is a constant shift of j, so for one
// (kz, kh) the 128 A rows of all three kw taps are windows of ONE slab of 130 consecutive
// virtual input voxels.  The slab is staged once per (kz, kh) in the canonical no-swizzle K-major
// operand layout [Cin/4][slot][4 floats] (a "plane" per 16-byte channel chunk, rows 16 B apart,
// SBO = 128 B, LBO = plane pitch), and the kw shift is just +16 B on the descriptor start
// address: no im2col copy ever exists, in HBM or in shared memory.  Stride-2 layers stage the
// even and odd input columns as two planes.  Weights for the three kw taps of the iteration are
// staged next to it, pre-packed on the host in operand order.
//
// Pipeline: 2 smem stages; all 128 threads stage iteration it+1 (LDG.128 -> cvt.rna.tf32 ->
// STS.128) while the tensor core runs iteration it; tcgen05.commit -> mbarrier frees a stage.
// Several CTAs per SM (23 KB smem at Cin = 16) overlap each other's staging, MMA and epilogue.
// Epilogue: tcgen05.ld (lane = voxel) -> + shift -> ReLU -> + skip -> 128-bit stores.
//
// Precision modes: TF32 (inputs rounded to nearest, like cuDNN's default conv math on GPU) and
// 3xTF32 (hi/lo split of both operands, three MMAs: error ~2^-21, fp32-grade).
#include "common.cuh"
#include "tcgen05.cuh"

namespace mvs {
namespace tc {

constexpr int TC_THREADS = 128;
constexpr int TC_SLOTS = 132;                 // 130 used (128 rows + 2 shifts), padded
constexpr int TC_SL = TC_SLOTS * 16;          // bytes per 16-byte-chunk plane

struct TcDims {
    int B, D, H, W, Do, Ho, Wo, Cin, Cout;
    int kd, sd, s2;          // s2 = 1: stride 2 in y and x
    int relu;
    int PW, tiles_per_plane;
};

// ------------------------------------------------------------------------------------------------
// Probe: runs `nk` MMAs (M = 128, N, K = 8 each) on caller-supplied shared-memory images with
// caller-supplied descriptor strides and dumps the 128 x N accumulator.  tests/ use it to pin the
// operand-layout conventions against a plain matrix product.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS)
tc_probe_kernel(const float* __restrict__ a_img, int a_bytes, const float* __restrict__ b_img, int b_bytes, uint32_t a_lbo,
                uint32_t a_sbo, uint32_t b_lbo, uint32_t b_sbo, int N, int nk, uint32_t a_kstep, uint32_t b_kstep,
                float* __restrict__ d_out) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    float* sa = reinterpret_cast<float*>(smem);
    float* sb = reinterpret_cast<float*>(smem + ((a_bytes + 127) / 128) * 128);
    for (int i = threadIdx.x; i < a_bytes / 4; i += TC_THREADS) sa[i] = a_img[i];
    for (int i = threadIdx.x; i < b_bytes / 4; i += TC_THREADS) sb[i] = b_img[i];
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (threadIdx.x < 32) tmem_alloc(&tmem_slot, 64);
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc_tf32(128, N);
        for (int k = 0; k < nk; ++k) {
            const uint64_t ad = make_smem_desc(smem_u32(sa) + k * a_kstep, a_lbo, a_sbo);
            const uint64_t bd = make_smem_desc(smem_u32(sb) + k * b_kstep, b_lbo, b_sbo);
            mma_tf32_ss(tmem, ad, bd, idesc, k > 0 ? 1u : 0u);
        }
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after_sync();
    const int warp = threadIdx.x >> 5;
    for (int c0 = 0; c0 < N; c0 += 16) {
        float v[16];
        tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        for (int i = 0; i < 16; ++i) d_out[(size_t)threadIdx.x * N + c0 + i] = v[i];
    }
    tc_fence_before_sync();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem, 64);
}

// Probe of the A-from-TMEM path: each K step's A slice is first copied smem -> TMEM with
// tcgen05.cp.128x256b (descriptor start shifted by `a_shift_bytes`, as the conv kernels do for the
// kw taps), then multiplied with tcgen05.mma (A in TMEM, B in smem).
__global__ void __launch_bounds__(TC_THREADS)
tc_probe_ts_kernel(const float* __restrict__ a_img, int a_bytes, const float* __restrict__ b_img, int b_bytes, uint32_t a_lbo,
                   uint32_t a_sbo, uint32_t b_lbo, uint32_t b_sbo, int N, int nk, uint32_t a_kstep, uint32_t b_kstep,
                   uint32_t a_shift_bytes, float* __restrict__ d_out) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    float* sa = reinterpret_cast<float*>(smem);
    float* sb = reinterpret_cast<float*>(smem + ((a_bytes + 127) / 128) * 128);
    for (int i = threadIdx.x; i < a_bytes / 4; i += TC_THREADS) sa[i] = a_img[i];
    for (int i = threadIdx.x; i < b_bytes / 4; i += TC_THREADS) sb[i] = b_img[i];
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (threadIdx.x < 32) tmem_alloc(&tmem_slot, 256);
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_slot;
    const uint32_t a_tmem = tmem + 64;                        // A slices live behind the accumulator columns
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc_tf32(128, N);
        for (int k = 0; k < nk; ++k) {
            const uint64_t ad = make_smem_desc(smem_u32(sa) + k * a_kstep + a_shift_bytes, a_lbo, a_sbo);
            tmem_cp_128x256b(a_tmem + k * 8, ad);
        }
        for (int k = 0; k < nk; ++k) {
            const uint64_t bd = make_smem_desc(smem_u32(sb) + k * b_kstep, b_lbo, b_sbo);
            mma_tf32_ts(tmem, a_tmem + k * 8, bd, idesc, k > 0 ? 1u : 0u);
        }
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after_sync();
    const int warp = threadIdx.x >> 5;
    for (int c0 = 0; c0 < N; c0 += 16) {
        float v[16];
        tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        for (int i = 0; i < 16; ++i) d_out[(size_t)threadIdx.x * N + c0 + i] = v[i];
    }
    tc_fence_before_sync();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem, 256);
}

// ------------------------------------------------------------------------------------------------
// convolution
// ------------------------------------------------------------------------------------------------
// CS = input channels staged per pipeline iteration (Cin is processed in Cin/CS slices so that
// wide layers still fit two stages in shared memory).
template <int CIN, int NT, bool X3>
struct TcSmem {
    static constexpr int NSPLIT = X3 ? 2 : 1;
    static constexpr int CH = CIN / 4;                       // 16-byte channel chunks
    static constexpr int A_PLANE = CH * TC_SL;               // one (split, plane) block
    static constexpr int B_TAP = CH * NT * 16;               // one kw tap
    static constexpr int B_BLOCK = 3 * B_TAP;                // one split
    static __host__ __device__ constexpr int a_bytes(int nplanes) { return 2 * NSPLIT * nplanes * A_PLANE; }
    static __host__ __device__ constexpr int b_bytes() { return 2 * NSPLIT * B_BLOCK; }
    static __host__ __device__ constexpr int total(int nplanes) { return a_bytes(nplanes) + b_bytes() + 128; }
};

template <int CIN, int NT, bool X3>
__global__ void __launch_bounds__(TC_THREADS)
conv3d_tc_kernel(const float* __restrict__ x, const float* __restrict__ w_hi, const float* __restrict__ w_lo,
                 const float* __restrict__ shift, const float* __restrict__ skip, float* __restrict__ y, TcDims d) {
    using L = TcSmem<CIN, NT, X3>;
    extern __shared__ __align__(128) uint8_t smem[];
    const int nplanes = d.s2 ? 2 : 1;
    uint8_t* sA = smem;
    uint8_t* sB = smem + L::a_bytes(nplanes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::a_bytes(nplanes) + L::b_bytes());
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
    constexpr uint32_t TMEM_COLS = NT <= 32 ? 32 : (NT <= 64 ? 64 : 128);

    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *tmem_slot;

    const int plane_idx = blockIdx.x / d.tiles_per_plane;          // b * Do + z
    const int tile = blockIdx.x - plane_idx * d.tiles_per_plane;
    const int b = plane_idx / d.Do, z = plane_idx - b * d.Do;
    const int j0 = tile * 128;
    const int co0 = blockIdx.y * NT;
    const int pd = d.kd / 2;
    const int nch = d.Cin / CIN;                 // channel slices per (kz, kh)
    const int nit = d.kd * 3 * nch;

    // The slots this thread stages are the same in every iteration: slot tid, and slot tid + 128
    // for the two threads that own the shifted tail.
    int sy[2], sa[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int jv = j0 + tid + u * 128;
        sy[u] = jv / d.PW;
        sa[u] = jv - sy[u] * d.PW;
    }
    const int nslot_iters = (tid + 128 < 130) ? 2 : 1;

    const float* wt_hi = w_hi + (size_t)blockIdx.y * nit * (L::B_BLOCK / 4);
    const float* wt_lo = X3 ? w_lo + (size_t)blockIdx.y * nit * (L::B_BLOCK / 4) : nullptr;

    for (int it = 0; it < nit; ++it) {
        const int buf = it & 1;
        const int kz = it / (3 * nch), rem = it - kz * 3 * nch;
        const int kh = rem / nch, ch = rem - kh * nch;
        if (it >= 2) mbar_wait(&bars[buf], ((it >> 1) - 1) & 1);     // MMAs that read this stage are done

        // ---- stage A: one virtual row-slab per plane ------------------------------------------
        const int zz = z * d.sd - pd + kz;
        const bool zok = zz >= 0 && zz < d.D;
        for (int p = 0; p < nplanes; ++p) {
            uint8_t* dst_hi = sA + (size_t)((buf * L::NSPLIT + 0) * nplanes + p) * L::A_PLANE;
            uint8_t* dst_lo = sA + (size_t)((buf * L::NSPLIT + (X3 ? 1 : 0)) * nplanes + p) * L::A_PLANE;
            for (int u = 0; u < nslot_iters; ++u) {
                const int slot = tid + u * 128;
                const int yy = d.s2 ? 2 * sy[u] + kh - 1 : sy[u] + kh - 1;
                const int xx = d.s2 ? (p == 0 ? 2 * sa[u] : 2 * sa[u] - 1) : sa[u] - 1;
                const bool ok = zok && yy >= 0 && yy < d.H && xx >= 0 && xx < d.W;
                const float4* src = reinterpret_cast<const float4*>(
                    x + ((((size_t)b * d.D + (ok ? zz : 0)) * d.H + (ok ? yy : 0)) * d.W + (ok ? xx : 0)) * d.Cin + ch * CIN);
#pragma unroll
                for (int q0 = 0; q0 < L::CH; q0 += 4) {
                    float4 v[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (q0 + q < L::CH) v[q] = ok ? __ldg(src + q0 + q) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        if (q0 + q >= L::CH) continue;
                        float4 hi = make_float4(to_tf32(v[q].x), to_tf32(v[q].y), to_tf32(v[q].z), to_tf32(v[q].w));
                        *reinterpret_cast<float4*>(dst_hi + (size_t)(q0 + q) * TC_SL + slot * 16) = hi;
                        if (X3) {
                            float4 lo = make_float4(to_tf32(v[q].x - hi.x), to_tf32(v[q].y - hi.y), to_tf32(v[q].z - hi.z),
                                                    to_tf32(v[q].w - hi.w));
                            *reinterpret_cast<float4*>(dst_lo + (size_t)(q0 + q) * TC_SL + slot * 16) = lo;
                        }
                    }
                }
            }
        }
        // ---- stage B: the three kw taps of this (kz, kh), already in operand order ---------------
        {
            float4* dst_hi = reinterpret_cast<float4*>(sB + (size_t)(buf * L::NSPLIT + 0) * L::B_BLOCK);
            const float4* src_hi = reinterpret_cast<const float4*>(wt_hi) + (size_t)it * (L::B_BLOCK / 16);
            for (int i = tid; i < L::B_BLOCK / 16; i += TC_THREADS) dst_hi[i] = __ldg(src_hi + i);
            if (X3) {
                float4* dst_lo = reinterpret_cast<float4*>(sB + (size_t)(buf * L::NSPLIT + 1) * L::B_BLOCK);
                const float4* src_lo = reinterpret_cast<const float4*>(wt_lo) + (size_t)it * (L::B_BLOCK / 16);
                for (int i = tid; i < L::B_BLOCK / 16; i += TC_THREADS) dst_lo[i] = __ldg(src_lo + i);
            }
        }
        fence_proxy_async_smem();
        __syncthreads();

        // ---- MMAs of this iteration (one thread) -------------------------------------------------
        if (tid == 0) {
            tc_fence_after_sync();
            constexpr uint32_t idesc = make_idesc_tf32(128, NT);
            const uint32_t a_hi = smem_u32(sA) + (uint32_t)((buf * L::NSPLIT + 0) * nplanes) * L::A_PLANE;
            const uint32_t a_lo = smem_u32(sA) + (uint32_t)((buf * L::NSPLIT + (X3 ? 1 : 0)) * nplanes) * L::A_PLANE;
            const uint32_t b_hi = smem_u32(sB) + (uint32_t)(buf * L::NSPLIT + 0) * L::B_BLOCK;
            const uint32_t b_lo = smem_u32(sB) + (uint32_t)(buf * L::NSPLIT + (X3 ? 1 : 0)) * L::B_BLOCK;
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                // which plane holds the input column of this tap, and by how many slots it is shifted
                const int p = d.s2 ? (kw == 1 ? 0 : 1) : 0;
                const int sh = d.s2 ? (kw == 2 ? 1 : 0) : kw;
                const uint32_t aoff = (uint32_t)p * L::A_PLANE + (uint32_t)sh * 16;
#pragma unroll
                for (int kk = 0; kk < CIN / 8; ++kk) {
                    const uint32_t ao = aoff + (uint32_t)(2 * kk) * TC_SL;
                    const uint32_t bo = (uint32_t)kw * L::B_TAP + (uint32_t)(2 * kk) * NT * 16;
                    const uint32_t first = (it == 0 && kw == 0 && kk == 0) ? 0u : 1u;
                    const uint64_t adh = make_smem_desc(a_hi + ao, TC_SL, 128);
                    const uint64_t bdh = make_smem_desc(b_hi + bo, NT * 16, 128);
                    if (X3) {
                        const uint64_t adl = make_smem_desc(a_lo + ao, TC_SL, 128);
                        const uint64_t bdl = make_smem_desc(b_lo + bo, NT * 16, 128);
                        mma_tf32_ss(tmem, adl, bdh, idesc, first);
                        mma_tf32_ss(tmem, adh, bdl, idesc, 1u);
                        mma_tf32_ss(tmem, adh, bdh, idesc, 1u);
                    } else {
                        mma_tf32_ss(tmem, adh, bdh, idesc, first);
                    }
                }
            }
            mma_commit(&bars[buf]);
        }
    }

    // ---- epilogue: TMEM lane = tile row = output voxel ------------------------------------------
    const int last = nit - 1;
    mbar_wait(&bars[last & 1], (last >> 1) & 1);
    tc_fence_after_sync();
    float acc[NT];
#pragma unroll
    for (int c0 = 0; c0 < NT; c0 += 16) tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, acc + c0);

    const int oy = sy[0], ox = sa[0];
    if (oy < d.Ho && ox < d.Wo) {
        const size_t o = ((((size_t)b * d.Do + z) * d.Ho + oy) * d.Wo + ox) * d.Cout + co0;
        const bool live_cols = true;
#pragma unroll
        for (int q = 0; q < NT / 4; ++q) {
            if (co0 + q * 4 >= d.Cout) break;            // N padded beyond Cout (Cout = 8 with NT = 16)
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
        (void)live_cols;
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, TMEM_COLS);
}

template <int CIN, int NT, bool X3>
static int launch_tc(const float* x, const float* w_hi, const float* w_lo, const float* shift, const float* skip, float* y,
                     const TcDims& d, cudaStream_t st) {
    using L = TcSmem<CIN, NT, X3>;
    const int nplanes = d.s2 ? 2 : 1;
    const size_t smem = L::total(nplanes);
    MVS_REQUIRE(smem <= 227 * 1024, "mvs_conv3d_tc: needs %zu bytes of shared memory", smem);
    auto kern = conv3d_tc_kernel<CIN, NT, X3>;
    MVS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((size_t)d.B * d.Do * d.tiles_per_plane), (unsigned)((d.Cout + NT - 1) / NT));
    kern<<<grid, TC_THREADS, smem, st>>>(x, w_hi, w_lo, shift, skip, y, d);
    MVS_LAUNCH_OK("conv3d_tc_kernel");
    return MVS_OK;
}

// ------------------------------------------------------------------------------------------------
// transposed convolution (ConvTranspose3d kernel (kd,3,3), stride (sd,2,2), pad k/2,
// output_padding stride-1) in gather form.
//
// A CTA owns 128 consecutive *input* positions of one output depth slice zo, in the virtual index
// j = y * (W + 1) + x, and produces the four output parity classes (2y+py, 2x+px) as four
// accumulators in TMEM.  Output row parity fixes the kh taps and the input row they read
// (py=0: kh=1 from row y; py=1: kh=2 from row y and kh=0 from row y+1), likewise in x
// (px=0: kw=1 from x; px=1: kw=2 from x and kw=0 from x+1), so every one of the 9 in-plane taps
// is exactly one MMA chain and nothing is predicated off.  Per (kz, dy, channel slice) one slab
// of 129 virtual input voxels is staged; the x+1 tap is the +16 B window of the same slab.
// ------------------------------------------------------------------------------------------------
template <int CIN, int NT, bool X3>
struct TcDeconvSmem {
    static constexpr int NSPLIT = X3 ? 2 : 1;
    static constexpr int CH = CIN / 4;
    static constexpr int A_PLANE = CH * TC_SL;
    static constexpr int B_TAP = CH * NT * 16;
    static constexpr int B_BLOCK = 6 * B_TAP;                  // up to six taps per iteration
    static __host__ __device__ constexpr int a_bytes() { return 2 * NSPLIT * A_PLANE; }
    static __host__ __device__ constexpr int b_bytes() { return 2 * NSPLIT * B_BLOCK; }
    static __host__ __device__ constexpr int total() { return a_bytes() + b_bytes() + 128; }
};

template <int CIN, int NT, bool X3>
__global__ void __launch_bounds__(TC_THREADS)
deconv3d_tc_kernel(const float* __restrict__ x, const float* __restrict__ w_hi, const float* __restrict__ w_lo,
                   const float* __restrict__ shift, const float* __restrict__ skip, float* __restrict__ y, TcDims d) {
    using L = TcDeconvSmem<CIN, NT, X3>;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* sA = smem;
    uint8_t* sB = smem + L::a_bytes();
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::a_bytes() + L::b_bytes());
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
    constexpr uint32_t TMEM_COLS = 4 * NT <= 32 ? 32 : (4 * NT <= 64 ? 64 : (4 * NT <= 128 ? 128 : 256));

    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *tmem_slot;

    const int plane_idx = blockIdx.x / d.tiles_per_plane;          // b * Do + zo
    const int tile = blockIdx.x - plane_idx * d.tiles_per_plane;
    const int b = plane_idx / d.Do, zo = plane_idx - b * d.Do;
    const int j0 = tile * 128;
    const int co0 = blockIdx.y * NT;
    const int pd = d.kd / 2;
    const int nch = d.Cin / CIN;

    // depth taps that feed this output slice: zo = iz * sd - pd + kz
    int kzv[3], izv[3], nkz = 0;
    for (int kz = 0; kz < d.kd; ++kz) {
        const int t = zo + pd - kz;
        if (t < 0 || (t % d.sd) != 0) continue;
        const int iz = t / d.sd;
        if (iz >= d.D) continue;
        kzv[nkz] = kz; izv[nkz] = iz; ++nkz;
    }
    const int nit = nkz * 2 * nch;

    int sy[2], sa[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int jv = j0 + tid + u * 128;
        sy[u] = jv / d.PW;
        sa[u] = jv - sy[u] * d.PW;
    }
    const int nslot_iters = (tid + 128 < 129) ? 2 : 1;
    const size_t wtile = (size_t)d.kd * 2 * nch * (L::B_BLOCK / 4);          // floats per Cout tile
    uint32_t started = 0;                                                    // per-class "accumulator written" bits (thread 0)

    for (int it = 0; it < nit; ++it) {
        const int buf = it & 1;
        const int kzi = it / (2 * nch), rem = it - kzi * 2 * nch;
        const int dy = rem / nch, ch = rem - dy * nch;
        const int kz = kzv[kzi], iz = izv[kzi];
        if (it >= 2) mbar_wait(&bars[buf], ((it >> 1) - 1) & 1);

        // ---- stage A: input row y + dy ---------------------------------------------------------
        {
            uint8_t* dst_hi = sA + (size_t)(buf * L::NSPLIT + 0) * L::A_PLANE;
            uint8_t* dst_lo = sA + (size_t)(buf * L::NSPLIT + (X3 ? 1 : 0)) * L::A_PLANE;
            for (int u = 0; u < nslot_iters; ++u) {
                const int slot = tid + u * 128;
                const int yy = sy[u] + dy, xx = sa[u];
                const bool ok = yy < d.H && xx < d.W;
                const float4* src = reinterpret_cast<const float4*>(
                    x + ((((size_t)b * d.D + iz) * d.H + (ok ? yy : 0)) * d.W + (ok ? xx : 0)) * d.Cin + ch * CIN);
#pragma unroll
                for (int q0 = 0; q0 < L::CH; q0 += 4) {
                    float4 v[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (q0 + q < L::CH) v[q] = ok ? __ldg(src + q0 + q) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        if (q0 + q >= L::CH) continue;
                        float4 hi = make_float4(to_tf32(v[q].x), to_tf32(v[q].y), to_tf32(v[q].z), to_tf32(v[q].w));
                        *reinterpret_cast<float4*>(dst_hi + (size_t)(q0 + q) * TC_SL + slot * 16) = hi;
                        if (X3) {
                            float4 lo = make_float4(to_tf32(v[q].x - hi.x), to_tf32(v[q].y - hi.y), to_tf32(v[q].z - hi.z),
                                                    to_tf32(v[q].w - hi.w));
                            *reinterpret_cast<float4*>(dst_lo + (size_t)(q0 + q) * TC_SL + slot * 16) = lo;
                        }
                    }
                }
            }
        }
        // ---- stage B: six (dy = 0) or three (dy = 1) taps, packed [kz][dy][ch][tap][CH][NT][4] -----
        const int ntaps = dy == 0 ? 6 : 3;
        {
            const size_t blk = ((size_t)kz * 2 + dy) * nch + ch;
            float4* dst_hi = reinterpret_cast<float4*>(sB + (size_t)(buf * L::NSPLIT + 0) * L::B_BLOCK);
            const float4* src_hi = reinterpret_cast<const float4*>(w_hi + (size_t)blockIdx.y * wtile) + blk * (L::B_BLOCK / 16);
            for (int i = tid; i < ntaps * (L::B_TAP / 16); i += TC_THREADS) dst_hi[i] = __ldg(src_hi + i);
            if (X3) {
                float4* dst_lo = reinterpret_cast<float4*>(sB + (size_t)(buf * L::NSPLIT + 1) * L::B_BLOCK);
                const float4* src_lo = reinterpret_cast<const float4*>(w_lo + (size_t)blockIdx.y * wtile) + blk * (L::B_BLOCK / 16);
                for (int i = tid; i < ntaps * (L::B_TAP / 16); i += TC_THREADS) dst_lo[i] = __ldg(src_lo + i);
            }
        }
        fence_proxy_async_smem();
        __syncthreads();

        if (tid == 0) {
            tc_fence_after_sync();
            constexpr uint32_t idesc = make_idesc_tf32(128, NT);
            const uint32_t a_hi = smem_u32(sA) + (uint32_t)(buf * L::NSPLIT + 0) * L::A_PLANE;
            const uint32_t a_lo = smem_u32(sA) + (uint32_t)(buf * L::NSPLIT + (X3 ? 1 : 0)) * L::A_PLANE;
            const uint32_t b_hi = smem_u32(sB) + (uint32_t)(buf * L::NSPLIT + 0) * L::B_BLOCK;
            const uint32_t b_lo = smem_u32(sB) + (uint32_t)(buf * L::NSPLIT + (X3 ? 1 : 0)) * L::B_BLOCK;
            for (int t = 0; t < ntaps; ++t) {
                // tap order: dy = 0 -> (kh=1,kw=0..2), (kh=2,kw=0..2); dy = 1 -> (kh=0,kw=0..2)
                const int kh = dy == 0 ? 1 + t / 3 : 0;
                const int kw = t % 3;
                const int py = (kh == 1) ? 0 : 1;
                const int px = (kw == 1) ? 0 : 1;
                const int sh = (kw == 0) ? 1 : 0;                   // kw = 0 reads input x + 1
                const int cls = py * 2 + px;
                const uint32_t dcol = tmem + (uint32_t)cls * NT;
#pragma unroll
                for (int kk = 0; kk < CIN / 8; ++kk) {
                    const uint32_t ao = (uint32_t)sh * 16 + (uint32_t)(2 * kk) * TC_SL;
                    const uint32_t bo = (uint32_t)t * L::B_TAP + (uint32_t)(2 * kk) * NT * 16;
                    const uint32_t acc = (started >> cls) & 1u;
                    started |= 1u << cls;
                    const uint64_t adh = make_smem_desc(a_hi + ao, TC_SL, 128);
                    const uint64_t bdh = make_smem_desc(b_hi + bo, NT * 16, 128);
                    if (X3) {
                        const uint64_t adl = make_smem_desc(a_lo + ao, TC_SL, 128);
                        const uint64_t bdl = make_smem_desc(b_lo + bo, NT * 16, 128);
                        mma_tf32_ss(dcol, adl, bdh, idesc, acc);
                        mma_tf32_ss(dcol, adh, bdl, idesc, 1u);
                        mma_tf32_ss(dcol, adh, bdh, idesc, 1u);
                    } else {
                        mma_tf32_ss(dcol, adh, bdh, idesc, acc);
                    }
                }
            }
            mma_commit(&bars[buf]);
        }
    }

    const int last = nit - 1;
    mbar_wait(&bars[last & 1], (last >> 1) & 1);
    tc_fence_after_sync();
    const int iy = sy[0], ix = sa[0];
    const bool live = iy < d.H && ix < d.W;
#pragma unroll
    for (int cls = 0; cls < 4; ++cls) {
        float acc[NT];
#pragma unroll
        for (int c0 = 0; c0 < NT; c0 += 16) tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + cls * NT + c0, acc + c0);
        if (!live) continue;
        const int oy = 2 * iy + (cls >> 1), ox = 2 * ix + (cls & 1);
        const size_t o = ((((size_t)b * d.Do + zo) * d.Ho + oy) * d.Wo + ox) * d.Cout + co0;
#pragma unroll
        for (int q = 0; q < NT / 4; ++q) {
            if (co0 + q * 4 >= d.Cout) break;
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
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, TMEM_COLS);
}

template <int CIN, int NT, bool X3>
static int launch_deconv_tc(const float* x, const float* w_hi, const float* w_lo, const float* shift, const float* skip,
                            float* y, const TcDims& d, cudaStream_t st) {
    using L = TcDeconvSmem<CIN, NT, X3>;
    const size_t smem = L::total();
    MVS_REQUIRE(smem <= 227 * 1024, "mvs_deconv3d_tc: needs %zu bytes of shared memory", smem);
    auto kern = deconv3d_tc_kernel<CIN, NT, X3>;
    MVS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((size_t)d.B * d.Do * d.tiles_per_plane), (unsigned)((d.Cout + NT - 1) / NT));
    kern<<<grid, TC_THREADS, smem, st>>>(x, w_hi, w_lo, shift, skip, y, d);
    MVS_LAUNCH_OK("deconv3d_tc_kernel");
    return MVS_OK;
}

}  // namespace tc
}  // namespace mvs

extern "C" int mvs_tc_probe(const float* a_img, int a_bytes, const float* b_img, int b_bytes, unsigned a_lbo,
                            unsigned a_sbo, unsigned b_lbo, unsigned b_sbo, int N, int nk, unsigned a_kstep,
                            unsigned b_kstep, float* d_out, void* stream) {
    using namespace mvs;
    MVS_REQUIRE(a_img && b_img && d_out, "mvs_tc_probe: null pointer");
    MVS_REQUIRE(N >= 16 && N <= 64 && N % 16 == 0 && nk >= 1, "mvs_tc_probe: bad N/nk");
    MVS_REQUIRE(a_bytes % 16 == 0 && b_bytes % 16 == 0 && a_bytes + b_bytes + 256 <= 200 * 1024, "mvs_tc_probe: bad image sizes");
    const size_t smem = (size_t)((a_bytes + 127) / 128) * 128 + b_bytes + 128;
    MVS_CUDA_OK(cudaFuncSetAttribute(tc::tc_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tc::tc_probe_kernel<<<1, tc::TC_THREADS, smem, (cudaStream_t)stream>>>(a_img, a_bytes, b_img, b_bytes, a_lbo, a_sbo, b_lbo,
                                                                           b_sbo, N, nk, a_kstep, b_kstep, d_out);
    MVS_LAUNCH_OK("tc_probe_kernel");
    return MVS_OK;
}

extern "C" int mvs_tc_probe_ts(const float* a_img, int a_bytes, const float* b_img, int b_bytes, unsigned a_lbo,
                               unsigned a_sbo, unsigned b_lbo, unsigned b_sbo, int N, int nk, unsigned a_kstep,
                               unsigned b_kstep, unsigned a_shift_bytes, float* d_out, void* stream) {
    using namespace mvs;
    MVS_REQUIRE(a_img && b_img && d_out, "mvs_tc_probe_ts: null pointer");
    MVS_REQUIRE(N >= 16 && N <= 64 && N % 16 == 0 && nk >= 1 && nk <= 16, "mvs_tc_probe_ts: bad N/nk");
    MVS_REQUIRE(a_bytes % 16 == 0 && b_bytes % 16 == 0 && a_bytes + b_bytes + 256 <= 200 * 1024, "mvs_tc_probe_ts: bad image sizes");
    const size_t smem = (size_t)((a_bytes + 127) / 128) * 128 + b_bytes + 128;
    MVS_CUDA_OK(cudaFuncSetAttribute(tc::tc_probe_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tc::tc_probe_ts_kernel<<<1, tc::TC_THREADS, smem, (cudaStream_t)stream>>>(a_img, a_bytes, b_img, b_bytes, a_lbo, a_sbo, b_lbo,
                                                                              b_sbo, N, nk, a_kstep, b_kstep, a_shift_bytes, d_out);
    MVS_LAUNCH_OK("tc_probe_ts_kernel");
    return MVS_OK;
}

extern "C" int mvs_conv3d_tc(const float* x, const float* w_hi, const float* w_lo, const float* shift, const float* skip,
                             float* y, int B, int D, int H, int W, int Cin, int Cout, int n_tile, int kd, int sd, int shw,
                             int relu, void* stream) {
    using namespace mvs;
    using namespace mvs::tc;
    MVS_REQUIRE(x && w_hi && y, "mvs_conv3d_tc: null pointer");
    MVS_REQUIRE(B >= 1 && D >= 1 && H >= 1 && W >= 1, "mvs_conv3d_tc: empty shape");
    MVS_REQUIRE(kd == 1 || kd == 3, "mvs_conv3d_tc: depth kernel size must be 1 or 3 (got %d)", kd);
    MVS_REQUIRE((sd == 1 || sd == 2) && (shw == 1 || shw == 2), "mvs_conv3d_tc: strides must be 1 or 2");
    MVS_REQUIRE(Cout % 8 == 0 && Cout >= 8, "mvs_conv3d_tc: Cout must be a multiple of 8 (got %d)", Cout);
    const int pd = kd / 2;
    TcDims d;
    d.B = B; d.D = D; d.H = H; d.W = W;
    d.Do = (D + 2 * pd - kd) / sd + 1; d.Ho = (H - 1) / shw + 1; d.Wo = (W - 1) / shw + 1;
    d.Cin = Cin; d.Cout = Cout; d.kd = kd; d.sd = sd; d.s2 = (shw == 2) ? 1 : 0; d.relu = relu;
    d.PW = d.s2 ? d.Wo + 1 : d.Wo + 2;
    d.tiles_per_plane = (int)(((int64_t)d.Ho * d.PW + 127) / 128);
    MVS_REQUIRE((int64_t)B * d.Do * d.tiles_per_plane < 2147483647LL, "mvs_conv3d_tc: grid too large");
    cudaStream_t st = (cudaStream_t)stream;
    const bool x3 = (w_lo != nullptr);
    const int cs = Cin >= 32 ? 32 : Cin;       // channels per pipeline stage
    MVS_REQUIRE(Cin % cs == 0 && (cs == 8 || cs == 16 || cs == 32), "mvs_conv3d_tc: Cin must be 8, 16 or a multiple of 32 (got %d)", Cin);
#define MVS_TC_CASE(CIN, NT)                                                                                   \
    if (cs == CIN && n_tile == NT)                                                                             \
        return x3 ? launch_tc<CIN, NT, true>(x, w_hi, w_lo, shift, skip, y, d, st)                             \
                  : launch_tc<CIN, NT, false>(x, w_hi, w_lo, shift, skip, y, d, st);
    MVS_TC_CASE(8, 16)
    MVS_TC_CASE(16, 16)
    MVS_TC_CASE(16, 32)
    MVS_TC_CASE(32, 16)
    MVS_TC_CASE(32, 32)
    MVS_TC_CASE(32, 64)
#undef MVS_TC_CASE
    MVS_UNSUPPORTED("mvs_conv3d_tc: no tensor-core instantiation for Cin=%d, N tile=%d", Cin, n_tile);
}

extern "C" int mvs_deconv3d_tc(const float* x, const float* w_hi, const float* w_lo, const float* shift, const float* skip,
                               float* y, int B, int D, int H, int W, int Cin, int Cout, int n_tile, int kd, int sd,
                               int relu, void* stream) {
    using namespace mvs;
    using namespace mvs::tc;
    MVS_REQUIRE(x && w_hi && y, "mvs_deconv3d_tc: null pointer");
    MVS_REQUIRE(B >= 1 && D >= 1 && H >= 1 && W >= 1, "mvs_deconv3d_tc: empty shape");
    MVS_REQUIRE(kd == 1 || kd == 3, "mvs_deconv3d_tc: depth kernel size must be 1 or 3 (got %d)", kd);
    MVS_REQUIRE(sd == 1 || sd == 2, "mvs_deconv3d_tc: depth stride must be 1 or 2");
    MVS_REQUIRE(!(kd == 1 && sd != 1), "mvs_deconv3d_tc: kd = 1 requires sd = 1");
    MVS_REQUIRE(Cout % 8 == 0 && Cout >= 8, "mvs_deconv3d_tc: Cout must be a multiple of 8 (got %d)", Cout);
    TcDims d;
    d.B = B; d.D = D; d.H = H; d.W = W;
    d.Do = D * sd; d.Ho = 2 * H; d.Wo = 2 * W;
    d.Cin = Cin; d.Cout = Cout; d.kd = kd; d.sd = sd; d.s2 = 1; d.relu = relu;
    d.PW = W + 1;
    d.tiles_per_plane = (int)(((int64_t)H * d.PW + 127) / 128);
    MVS_REQUIRE((int64_t)B * d.Do * d.tiles_per_plane < 2147483647LL, "mvs_deconv3d_tc: grid too large");
    cudaStream_t st = (cudaStream_t)stream;
    const bool x3 = (w_lo != nullptr);
    const int cs = Cin >= 32 ? 32 : Cin;
    MVS_REQUIRE(Cin % cs == 0 && (cs == 8 || cs == 16 || cs == 32), "mvs_deconv3d_tc: Cin must be 8, 16 or a multiple of 32 (got %d)", Cin);
#define MVS_TCD_CASE(CIN, NT)                                                                                  \
    if (cs == CIN && n_tile == NT)                                                                             \
        return x3 ? launch_deconv_tc<CIN, NT, true>(x, w_hi, w_lo, shift, skip, y, d, st)                      \
                  : launch_deconv_tc<CIN, NT, false>(x, w_hi, w_lo, shift, skip, y, d, st);
    MVS_TCD_CASE(16, 16)
    MVS_TCD_CASE(32, 16)
    MVS_TCD_CASE(32, 32)
#undef MVS_TCD_CASE
    MVS_UNSUPPORTED("mvs_deconv3d_tc: no tensor-core instantiation for Cin=%d, N tile=%d", Cin, n_tile);
}
