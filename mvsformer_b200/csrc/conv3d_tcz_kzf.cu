// conv3d_tcz_kzf.cu — OPT-IN "kz-fused N" variant of the depth-fused tcgen05 convolution of conv3d_tcz.cu
// (MVS_TCZ_KZF=1; compiled and reviewed, not yet executed on a B200 — see DESIGN.md §10 item 2).
//
// It lives in its own translation unit, as a modified copy of conv3d_tcz_kernel and its launch code, so that
// the shipped, GPU-verified kernels of conv3d_tcz.cu keep the exact binary they were verified with; once this
// variant has passed its parity tests on the device the two files merge again.
#include <stdlib.h>

#include "common.cuh"
#include "tcgen05.cuh"

namespace mvs {
namespace tc {
namespace kzf {

constexpr int TZ_THREADS = 128;
constexpr int TZ_SLOTS = 132;
constexpr int TZ_SL = TZ_SLOTS * 16;
constexpr int TZ_STAGES = 4;

struct TzDims {
    int B, D, H, W, Ho, Wo, Cin, Cout;
    int kd;                  // 1 or 3 (depth stride is 1)
    int s2;                  // conv: stride 2 in y and x
    int relu;
    int PW, tiles_per_plane;
    int zc;                  // output depth slices per CTA
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// TAPS = weight taps staged per (kh | dy, channel-slice) group and depth tap: conv 3 (kw), deconv 6.
template <int CS, int NT, int TAPS, int NPLANES>
struct TzSmem {
    static constexpr int CH = CS / 4;
    static constexpr int A_STAGE = NPLANES * CH * TZ_SL;
    static constexpr int B_TAP = CH * NT * 16;
    static __host__ __device__ constexpr int b_group(int kd) { return kd * TAPS * B_TAP; }
};

// ------------------------------------------------------------------------------------------------
// forward convolution, kernel (kd,3,3), stride (1, s, s)
// ------------------------------------------------------------------------------------------------
// KZF (opt-in, MVS_TCZ_KZF=1; not yet timed): "kz-fused N".  Every tcgen05.mma re-reads its 4 KB A tile from
// shared memory, i.e. 4/N bytes per MAC, and N = Cout tile is only 16..64 here — the wide layers sit at that
// operand-read bound.  The depth taps kz = 0,1,2 of an input slab iz feed the output slices
// iz+1, iz, iz-1; with the slice accumulators laid out in DEcreasing slice order they form one contiguous
// window of TMEM columns, so ONE MMA with the B rows [kz][n] (N = 3*NT, weights packed
// [Cout_tiles][kh][Cin/CS][kw][CS/4][kd][n_tile][4]) replaces three and reads A once.  Only the very first
// MMA of each slab in the first (kh, channel-slice) group stays per-slice: that is where accumulators are
// initialised (accumulate = 0) and the slices of a window are in different states.
template <int CS, int NT, bool S2, bool KZF = false>
__global__ void __launch_bounds__(TZ_THREADS)
conv3d_tcz_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ shift,
                  const float* __restrict__ skip, float* __restrict__ y, TzDims d) {
    constexpr int NPL = S2 ? 2 : 1;
    using L = TzSmem<CS, NT, 3, NPL>;
    extern __shared__ __align__(128) uint8_t smem[];
    const int bgroup = L::b_group(d.kd);
    const int br = d.zc >= 3 ? 2 : TZ_STAGES;                 // weight-buffer ring (see header)
    uint8_t* sA = smem;
    uint8_t* sB = smem + TZ_STAGES * L::A_STAGE;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB + (size_t)br * bgroup);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + TZ_STAGES);

    const int tid = threadIdx.x, warp = tid >> 5;
    const int ncols = d.zc * NT;
    const uint32_t tmem_cols = ncols <= 32 ? 32 : (ncols <= 64 ? 64 : (ncols <= 128 ? 128 : (ncols <= 256 ? 256 : 512)));
    if (tid == 0) {
        for (int s = 0; s < TZ_STAGES; ++s) mbar_init(&bars[s], 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, tmem_cols);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *tmem_slot;

    const int b = blockIdx.z;
    const int tile = blockIdx.x;
    const int z0 = blockIdx.y * d.zc;                          // first output slice of this CTA
    const int nz = min(d.zc, d.D - z0);
    const int j0 = (tile % d.tiles_per_plane) * 128;
    const int ct = tile / d.tiles_per_plane;                   // Cout tile
    const int co0 = ct * NT;
    const int pd = d.kd / 2;
    const int nch = d.Cin / CS;

    // input slices that feed this z-chunk
    const int iz_lo = max(z0 - pd, 0), iz_hi = min(z0 + nz - 1 + pd, d.D - 1);
    const int niz = iz_hi - iz_lo + 1;
    const int ngroups = 3 * nch;                               // (kh, channel slice)
    const int nit = ngroups * niz;
    const int glen = br == 2 ? niz : 1;                        // iterations sharing one weight buffer

    int sy[2], sa[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int jv = j0 + tid + u * 128;
        sy[u] = jv / d.PW;
        sa[u] = jv - sy[u] * d.PW;
    }
    const int nslot_iters = (tid + 128 < 130) ? 2 : 1;
    const float* wt = w + (size_t)ct * ngroups * (bgroup / 4);

    auto issue = [&](int it) {
        const int g = it / niz, iz = iz_lo + (it - g * niz);
        const int kh = g / nch, ch = g - kh * nch;
        const uint32_t a_base = smem_u32(sA) + (uint32_t)(it % TZ_STAGES) * L::A_STAGE;
#pragma unroll
        for (int p = 0; p < NPL; ++p) {
            for (int u = 0; u < nslot_iters; ++u) {
                const int slot = tid + u * 128;
                const int yy = S2 ? 2 * sy[u] + kh - 1 : sy[u] + kh - 1;
                const int xx = S2 ? (p == 0 ? 2 * sa[u] : 2 * sa[u] - 1) : sa[u] - 1;
                const bool ok = yy >= 0 && yy < d.H && xx >= 0 && xx < d.W;
                const float* src = x + ((((size_t)b * d.D + iz) * d.H + (ok ? yy : 0)) * d.W + (ok ? xx : 0)) * d.Cin + ch * CS;
                const uint32_t dst = a_base + (uint32_t)p * L::CH * TZ_SL + slot * 16;
#pragma unroll
                for (int q = 0; q < L::CH; ++q) cp_async16(dst + q * TZ_SL, src + q * 4, ok ? 16u : 0u);
            }
        }
        // weights: once per group (glen > 1) or with every iteration (glen == 1)
        if (glen == 1 || it == g * niz) {
            const int slot_b = glen == 1 ? it % TZ_STAGES : g % 2;
            const uint32_t b_base = smem_u32(sB) + (uint32_t)slot_b * bgroup;
            const float4* srcb = reinterpret_cast<const float4*>(wt) + (size_t)g * (bgroup / 16);
            for (int i = tid; i < bgroup / 16; i += TZ_THREADS) cp_async16(b_base + i * 16, srcb + i, 16u);
        }
    };

#pragma unroll
    for (int i = 0; i < TZ_STAGES - 1; ++i) {
        if (i < nit) issue(i);
        cp_async_commit();
    }

    uint32_t started = 0;                                      // per-slice "accumulator written" bits (thread 0)
    for (int it = 0; it < nit; ++it) {
        const int nx = it + TZ_STAGES - 1;
        if (nx < nit) {
            // stage of iteration nx was last read by the MMAs of iteration it-1
            if (it >= 1) mbar_wait(&bars[(it - 1) % TZ_STAGES], ((it - 1) / TZ_STAGES) & 1);
            issue(nx);
        }
        cp_async_commit();
        cp_async_wait<TZ_STAGES - 1>();                        // this thread's copies of iteration `it` have landed
        fence_proxy_async_smem();
        __syncthreads();

        if (tid == 0) {
            tc_fence_after_sync();
            constexpr uint32_t idesc = make_idesc_tf32(128, NT);
            const int g = it / niz, iz = iz_lo + (it - g * niz);
            const uint32_t a_base = smem_u32(sA) + (uint32_t)(it % TZ_STAGES) * L::A_STAGE;
            const uint32_t b_base = smem_u32(sB) + (uint32_t)(glen == 1 ? it % TZ_STAGES : g % 2) * bgroup;
            if (!KZF) {
            for (int kz = 0; kz < d.kd; ++kz) {
                const int oz = iz + pd - kz;                   // output slice fed through depth tap kz
                if (oz < z0 || oz >= z0 + nz) continue;
                const uint32_t dcol = tmem + (uint32_t)(oz - z0) * NT;
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const int p = S2 ? (kw == 1 ? 0 : 1) : 0;
                    const int sh = S2 ? (kw == 2 ? 1 : 0) : kw;
                    const uint32_t aoff = (uint32_t)p * L::CH * TZ_SL + (uint32_t)sh * 16;
#pragma unroll
                    for (int kk = 0; kk < CS / 8; ++kk) {
                        const uint32_t acc = (started >> (oz - z0)) & 1u;
                        started |= 1u << (oz - z0);
                        const uint64_t ad = make_smem_desc(a_base + aoff + (uint32_t)(2 * kk) * TZ_SL, TZ_SL, 128);
                        const uint64_t bd = make_smem_desc(b_base + (uint32_t)(kz * 3 + kw) * L::B_TAP + (uint32_t)(2 * kk) * NT * 16, NT * 16, 128);
                        mma_tf32_ss(dcol, ad, bd, idesc, acc);
                    }
                }
            }
            } else {
                // depth taps whose output slice lies in this CTA's chunk: one contiguous range
                const int kz_lo = max(0, iz + pd - (z0 + nz - 1)), kz_hi = min(d.kd - 1, iz + pd - z0);
                if (kz_lo <= kz_hi) {
                    const uint32_t plane = (uint32_t)(d.kd * NT) * 16;      // bytes between the K chunks of a tap: rows [kz][n]
                    const uint32_t btap = (uint32_t)L::CH * plane;          // bytes per kw
                    const int zi_top = iz + pd - kz_lo - z0;                // slice of tap kz_lo = first column block of the window
                    const uint32_t dwin = tmem + (uint32_t)(nz - 1 - zi_top) * NT;
                    const uint32_t idesc_w = make_idesc_tf32(128, (kz_hi - kz_lo + 1) * NT);
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) {
                        const int p = S2 ? (kw == 1 ? 0 : 1) : 0;
                        const int sh = S2 ? (kw == 2 ? 1 : 0) : kw;
                        const uint32_t aoff = (uint32_t)p * L::CH * TZ_SL + (uint32_t)sh * 16;
#pragma unroll
                        for (int kk = 0; kk < CS / 8; ++kk) {
                            const uint64_t ad = make_smem_desc(a_base + aoff + (uint32_t)(2 * kk) * TZ_SL, TZ_SL, 128);
                            const uint32_t bk = b_base + (uint32_t)kw * btap + (uint32_t)(2 * kk) * plane;
                            if (g == 0 && kw == 0 && kk == 0) {
                                // accumulator initialisation: per slice, each with its own accumulate flag
                                for (int kz = kz_lo; kz <= kz_hi; ++kz) {
                                    const int zi = iz + pd - kz - z0;
                                    const uint32_t acc = (started >> zi) & 1u;
                                    started |= 1u << zi;
                                    const uint64_t bd = make_smem_desc(bk + (uint32_t)(kz * NT) * 16, plane, 128);
                                    mma_tf32_ss(tmem + (uint32_t)(nz - 1 - zi) * NT, ad, bd, idesc, acc);
                                }
                            } else {
                                const uint64_t bd = make_smem_desc(bk + (uint32_t)(kz_lo * NT) * 16, plane, 128);
                                mma_tf32_ss(dwin, ad, bd, idesc_w, 1u);
                            }
                        }
                    }
                }
            }
            mma_commit(&bars[it % TZ_STAGES]);
        }
    }

    const int last = nit - 1;
    mbar_wait(&bars[last % TZ_STAGES], (last / TZ_STAGES) & 1);
    tc_fence_after_sync();
    const int oy = sy[0], ox = sa[0];
    const bool live = oy < d.Ho && ox < d.Wo;
    for (int zi = 0; zi < nz; ++zi) {
        float acc[NT];
#pragma unroll
        for (int c0 = 0; c0 < NT; c0 += 16)
            tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (KZF ? nz - 1 - zi : zi) * NT + c0, acc + c0);
        if (!live) continue;
        const size_t o = ((((size_t)b * d.D + z0 + zi) * d.Ho + oy) * d.Wo + ox) * d.Cout + co0;
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
            // consumers read this tensor as a TF32 operand: round once here
            r.x = to_tf32(r.x); r.y = to_tf32(r.y); r.z = to_tf32(r.z); r.w = to_tf32(r.w);
            reinterpret_cast<float4*>(y + o)[q] = r;
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, tmem_cols);
}

template <int CS, int NT, bool S2, bool KZF>
static int launch_conv_tcz(const float* x, const float* w, const float* shift, const float* skip, float* y, const TzDims& d,
                           cudaStream_t st) {
    using L = TzSmem<CS, NT, 3, S2 ? 2 : 1>;
    const int br = d.zc >= 3 ? 2 : TZ_STAGES;
    const size_t smem = (size_t)TZ_STAGES * L::A_STAGE + (size_t)br * L::b_group(d.kd) + 128;
    MVS_REQUIRE(smem <= 227 * 1024, "mvs_conv3d_tcz_kzf: needs %zu bytes of shared memory", smem);
    auto kern = conv3d_tcz_kernel<CS, NT, S2, KZF>;
    MVS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int ntiles = (d.Cout + NT - 1) / NT;
    dim3 grid((unsigned)(d.tiles_per_plane * ntiles), (unsigned)((d.D + d.zc - 1) / d.zc), (unsigned)d.B);
    kern<<<grid, TZ_THREADS, smem, st>>>(x, w, shift, skip, y, d);
    MVS_LAUNCH_OK("conv3d_tcz_kernel<KZF>");
    return MVS_OK;
}

// ------------------------------------------------------------------------------------------------
// transposed convolution, kernel (3,3,3), stride (1,2,2), kz-fused: the accumulators of a CTA are laid out
// [parity class][slice], so for a tap (-> class) the output slices iz-1, iz, iz+1 of the depth taps kz = 0,1,2
// are one contiguous window (ascending in kz) and one MMA of N = 3*NT feeds them.  Weights packed
// [Cout_tiles][2 dy][Cin/CS][6 taps][CS/4][kd][n_tile][4].  Modified copy of deconv3d_tcz_kernel.
// ------------------------------------------------------------------------------------------------
template <int CS, int NT>
__global__ void __launch_bounds__(TZ_THREADS)
deconv3d_tcz_kzf_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ shift,
                    const float* __restrict__ skip, float* __restrict__ y, TzDims d) {
    using L = TzSmem<CS, NT, 6, 1>;
    extern __shared__ __align__(128) uint8_t smem[];
    const int bgroup = L::b_group(d.kd);
    const int br = d.zc >= 3 ? 2 : TZ_STAGES;
    uint8_t* sA = smem;
    uint8_t* sB = smem + TZ_STAGES * L::A_STAGE;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB + (size_t)br * bgroup);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + TZ_STAGES);

    const int tid = threadIdx.x, warp = tid >> 5;
    const int ncols = d.zc * 4 * NT;
    const uint32_t tmem_cols = ncols <= 32 ? 32 : (ncols <= 64 ? 64 : (ncols <= 128 ? 128 : (ncols <= 256 ? 256 : 512)));
    if (tid == 0) {
        for (int s = 0; s < TZ_STAGES; ++s) mbar_init(&bars[s], 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, tmem_cols);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *tmem_slot;

    const int b = blockIdx.z;
    const int z0 = blockIdx.y * d.zc;
    const int nz = min(d.zc, d.D - z0);
    const int j0 = (blockIdx.x % d.tiles_per_plane) * 128;
    const int ct = blockIdx.x / d.tiles_per_plane;
    const int co0 = ct * NT;
    const int pd = d.kd / 2;
    const int nch = d.Cin / CS;
    const int iz_lo = max(z0 - pd, 0), iz_hi = min(z0 + nz - 1 + pd, d.D - 1);
    const int niz = iz_hi - iz_lo + 1;
    const int ngroups = 2 * nch;                               // (dy, channel slice)
    const int nit = ngroups * niz;
    const int glen = br == 2 ? niz : 1;

    int sy[2], sa[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int jv = j0 + tid + u * 128;
        sy[u] = jv / d.PW;
        sa[u] = jv - sy[u] * d.PW;
    }
    const int nslot_iters = (tid + 128 < 129) ? 2 : 1;
    const float* wt = w + (size_t)ct * ngroups * (bgroup / 4);

    auto issue = [&](int it) {
        const int g = it / niz, iz = iz_lo + (it - g * niz);
        const int dy = g / nch, ch = g - dy * nch;
        const uint32_t a_base = smem_u32(sA) + (uint32_t)(it % TZ_STAGES) * L::A_STAGE;
        for (int u = 0; u < nslot_iters; ++u) {
            const int slot = tid + u * 128;
            const int yy = sy[u] + dy, xx = sa[u];
            const bool ok = yy < d.H && xx < d.W;
            const float* src = x + ((((size_t)b * d.D + iz) * d.H + (ok ? yy : 0)) * d.W + (ok ? xx : 0)) * d.Cin + ch * CS;
            const uint32_t dst = a_base + slot * 16;
#pragma unroll
            for (int q = 0; q < L::CH; ++q) cp_async16(dst + q * TZ_SL, src + q * 4, ok ? 16u : 0u);
        }
        if (glen == 1 || it == g * niz) {
            const int slot_b = glen == 1 ? it % TZ_STAGES : g % 2;
            const uint32_t b_base = smem_u32(sB) + (uint32_t)slot_b * bgroup;
            const float4* srcb = reinterpret_cast<const float4*>(wt) + (size_t)g * (bgroup / 16);
            for (int i = tid; i < bgroup / 16; i += TZ_THREADS) cp_async16(b_base + i * 16, srcb + i, 16u);
        }
    };

#pragma unroll
    for (int i = 0; i < TZ_STAGES - 1; ++i) {
        if (i < nit) issue(i);
        cp_async_commit();
    }

    uint32_t started = 0;                                      // bit (class * nz + slice)
    for (int it = 0; it < nit; ++it) {
        const int nx = it + TZ_STAGES - 1;
        if (nx < nit) {
            if (it >= 1) mbar_wait(&bars[(it - 1) % TZ_STAGES], ((it - 1) / TZ_STAGES) & 1);
            issue(nx);
        }
        cp_async_commit();
        cp_async_wait<TZ_STAGES - 1>();
        fence_proxy_async_smem();
        __syncthreads();

        if (tid == 0) {
            tc_fence_after_sync();
            constexpr uint32_t idesc = make_idesc_tf32(128, NT);
            const int g = it / niz, iz = iz_lo + (it - g * niz);
            const int dy = g / nch;
            const uint32_t a_base = smem_u32(sA) + (uint32_t)(it % TZ_STAGES) * L::A_STAGE;
            const uint32_t b_base = smem_u32(sB) + (uint32_t)(glen == 1 ? it % TZ_STAGES : g % 2) * bgroup;
            const int ntaps = dy == 0 ? 6 : 3;
            // depth taps whose output slice (oz = iz - pd + kz) lies in this CTA's chunk: one contiguous range
            const int kz_lo = max(0, z0 - iz + pd), kz_hi = min(d.kd - 1, z0 + nz - 1 - iz + pd);
            if (kz_lo <= kz_hi) {
                const uint32_t plane = (uint32_t)(d.kd * NT) * 16;          // bytes between the K chunks of a tap: rows [kz][n]
                const uint32_t btap = (uint32_t)L::CH * plane;              // bytes per tap
                const int zi_lo = iz - pd + kz_lo - z0;                     // first slice of the window
                const uint32_t idesc_w = make_idesc_tf32(128, (kz_hi - kz_lo + 1) * NT);
                for (int t = 0; t < ntaps; ++t) {
                    const int kh = dy == 0 ? 1 + t / 3 : 0;
                    const int kw = t % 3;
                    const int cls = ((kh == 1) ? 0 : 2) + ((kw == 1) ? 0 : 1);
                    const int sh = (kw == 0) ? 1 : 0;
                    // accumulators are laid out [class][slice]: the window of a class is contiguous, ascending in kz
                    const uint32_t dwin = tmem + (uint32_t)(cls * nz + zi_lo) * NT;
                    // every class is first touched in group 0 (dy = 0, first channel slice) by these taps
                    const bool starter = g == 0 && (t == 0 || t == 1 || t == 3 || t == 4);
#pragma unroll
                    for (int kk = 0; kk < CS / 8; ++kk) {
                        const uint64_t ad = make_smem_desc(a_base + (uint32_t)sh * 16 + (uint32_t)(2 * kk) * TZ_SL, TZ_SL, 128);
                        const uint32_t bk = b_base + (uint32_t)t * btap + (uint32_t)(2 * kk) * plane;
                        if (starter && kk == 0) {
                            for (int kz = kz_lo; kz <= kz_hi; ++kz) {
                                const int slot_acc = cls * nz + (iz - pd + kz - z0);
                                const uint32_t acc = (started >> slot_acc) & 1u;
                                started |= 1u << slot_acc;
                                const uint64_t bd = make_smem_desc(bk + (uint32_t)(kz * NT) * 16, plane, 128);
                                mma_tf32_ss(tmem + (uint32_t)slot_acc * NT, ad, bd, idesc, acc);
                            }
                        } else {
                            const uint64_t bd = make_smem_desc(bk + (uint32_t)(kz_lo * NT) * 16, plane, 128);
                            mma_tf32_ss(dwin, ad, bd, idesc_w, 1u);
                        }
                    }
                }
            }
            mma_commit(&bars[it % TZ_STAGES]);
        }
    }

    const int last = nit - 1;
    mbar_wait(&bars[last % TZ_STAGES], (last / TZ_STAGES) & 1);
    tc_fence_after_sync();
    // Epilogue.  The skip tensor does not depend on the MMAs: its loads for the next group of
    // accumulators are issued before the current group is drained, so their latency is hidden.
    const int iy = sy[0], ix = sa[0];
    const bool live = iy < d.H && ix < d.W;
    constexpr int U = 64 / NT;                                  // accumulators per prefetch group (<= 16 float4)
    const int nacc = nz * 4;
    float4 skA[U][NT / 4], skB[U][NT / 4];                      // ping-pong register buffers (static indexing)
    auto out_index = [&](int a) -> size_t {
        const int zi = a >> 2, cls = a & 3;
        return ((((size_t)b * d.D + z0 + zi) * d.Ho + 2 * iy + (cls >> 1)) * d.Wo + 2 * ix + (cls & 1)) * d.Cout + co0;
    };
    auto prefetch = [&](int a0, float4 (&buf)[U][NT / 4]) {
        if (!skip || !live) return;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (a0 + u < nacc) {
                const float4* sp = reinterpret_cast<const float4*>(skip + out_index(a0 + u));
#pragma unroll
                for (int q = 0; q < NT / 4; ++q)
                    if (co0 + q * 4 < d.Cout) buf[u][q] = __ldg(sp + q);
            }
        }
    };
    auto drain = [&](int a0, float4 (&buf)[U][NT / 4]) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (a0 + u < nacc) {                                 // uniform across the CTA
                float acc[NT];
#pragma unroll
                for (int c0 = 0; c0 < NT; c0 += 16) tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (((a0 + u) & 3) * nz + ((a0 + u) >> 2)) * NT + c0, acc + c0);
                if (live) {
                    const size_t o = out_index(a0 + u);
#pragma unroll
                    for (int q = 0; q < NT / 4; ++q) {
                        if (co0 + q * 4 < d.Cout) {
                            float4 r = make_float4(acc[q * 4], acc[q * 4 + 1], acc[q * 4 + 2], acc[q * 4 + 3]);
                            if (shift) {
                                const float4 sft = __ldg(reinterpret_cast<const float4*>(shift + co0) + q);
                                r.x += sft.x; r.y += sft.y; r.z += sft.z; r.w += sft.w;
                            }
                            if (d.relu) { r.x = fmaxf(r.x, 0.f); r.y = fmaxf(r.y, 0.f); r.z = fmaxf(r.z, 0.f); r.w = fmaxf(r.w, 0.f); }
                            if (skip) { r.x += buf[u][q].x; r.y += buf[u][q].y; r.z += buf[u][q].z; r.w += buf[u][q].w; }
                            r.x = to_tf32(r.x); r.y = to_tf32(r.y); r.z = to_tf32(r.z); r.w = to_tf32(r.w);
                            reinterpret_cast<float4*>(y + o)[q] = r;
                        }
                    }
                }
            }
        }
    };
    prefetch(0, skA);
    for (int a0 = 0; a0 < nacc; a0 += 2 * U) {
        if (a0 + U < nacc) prefetch(a0 + U, skB);
        drain(a0, skA);
        if (a0 + 2 * U < nacc) prefetch(a0 + 2 * U, skA);
        if (a0 + U < nacc) drain(a0 + U, skB);
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, tmem_cols);
}

template <int CS, int NT>
static int launch_deconv_tcz_kzf(const float* x, const float* w, const float* shift, const float* skip, float* y, const TzDims& d,
                             cudaStream_t st) {
    using L = TzSmem<CS, NT, 6, 1>;
    const int br = d.zc >= 3 ? 2 : TZ_STAGES;
    const size_t smem = (size_t)TZ_STAGES * L::A_STAGE + (size_t)br * L::b_group(d.kd) + 128;
    MVS_REQUIRE(smem <= 227 * 1024, "mvs_deconv3d_tcz_kzf: needs %zu bytes of shared memory", smem);
    auto kern = deconv3d_tcz_kzf_kernel<CS, NT>;
    MVS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int ntiles = (d.Cout + NT - 1) / NT;
    dim3 grid((unsigned)(d.tiles_per_plane * ntiles), (unsigned)((d.D + d.zc - 1) / d.zc), (unsigned)d.B);
    kern<<<grid, TZ_THREADS, smem, st>>>(x, w, shift, skip, y, d);
    MVS_LAUNCH_OK("deconv3d_tcz_kzf_kernel");
    return MVS_OK;
}

// ------------------------------------------------------------------------------------------------
// row-tiled convolution, kh-fused (MVS_TCZ_KZF; the two tensor-core layers of the visibility net run here with
// kd = 1, R = 8 rows per CTA): an input row iy feeds the output rows iy+1, iy, iy-1 through kh = 0,1,2; with a
// slice's accumulators laid out in DEcreasing row order they are one contiguous window, so one MMA with the B rows
// [kh][n] (N = 3*NT; weights [Cout_tiles][kd][3 kw][Cin/4][3 kh][n_tile][4]) replaces three.  An MMA is fused
// whenever every accumulator of its window has been initialised, per-row otherwise.  Modified copy of conv3d_tcr_kernel.
// ------------------------------------------------------------------------------------------------
struct TrDims {
    int B, D, H, W, Cin, Cout;
    int kd, relu;
    int nxb;                 // 128-column blocks per row
    int R, zc;               // rows and depth slices per CTA
    int nrb, nzc;            // row blocks, z chunks
};

template <int CS, int NT>
__global__ void __launch_bounds__(TZ_THREADS)
conv3d_tcr_khf_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ shift,
                  const float* __restrict__ skip, float* __restrict__ y, TrDims d) {
    constexpr int CH = CS / 4;
    constexpr int A_STAGE = CH * TZ_SL;
    constexpr int B_TAP = CH * NT * 16;
    extern __shared__ __align__(128) uint8_t smem[];
    const int b_bytes = d.kd * 9 * B_TAP;
    uint8_t* sA = smem;
    uint8_t* sB = smem + TZ_STAGES * A_STAGE;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB + b_bytes);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + TZ_STAGES);

    const int tid = threadIdx.x, warp = tid >> 5;
    const int ncols = d.R * d.zc * NT;
    const uint32_t tmem_cols = ncols <= 32 ? 32 : (ncols <= 64 ? 64 : (ncols <= 128 ? 128 : (ncols <= 256 ? 256 : 512)));
    if (tid == 0) {
        for (int s = 0; s < TZ_STAGES; ++s) mbar_init(&bars[s], 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, tmem_cols);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *tmem_slot;

    const int b = blockIdx.z;
    const int xb = blockIdx.x % d.nxb, ct = blockIdx.x / d.nxb;
    const int rb = blockIdx.y % d.nrb, zcix = blockIdx.y / d.nrb;
    const int x0 = xb * 128, y0 = rb * d.R, z0 = zcix * d.zc;
    const int nr = min(d.R, d.H - y0), nz = min(d.zc, d.D - z0);
    const int co0 = ct * NT;
    const int pd = d.kd / 2;

    const int iz_lo = max(z0 - pd, 0), iz_hi = min(z0 + nz - 1 + pd, d.D - 1);
    const int iy_lo = max(y0 - 1, 0), iy_hi = min(y0 + nr, d.H - 1);
    const int niy = iy_hi - iy_lo + 1;
    const int nit = (iz_hi - iz_lo + 1) * niy;
    const int nslot_iters = (tid + 128 < 130) ? 2 : 1;

    auto issue = [&](int it) {
        const int iz = iz_lo + it / niy, iy = iy_lo + it % niy;
        const uint32_t a_base = smem_u32(sA) + (uint32_t)(it % TZ_STAGES) * A_STAGE;
        for (int u = 0; u < nslot_iters; ++u) {
            const int slot = tid + u * 128;
            const int xx = x0 - 1 + slot;
            const bool ok = xx >= 0 && xx < d.W;
            const float* src = x + ((((size_t)b * d.D + iz) * d.H + iy) * d.W + (ok ? xx : 0)) * d.Cin;
            const uint32_t dst = a_base + slot * 16;
#pragma unroll
            for (int q = 0; q < CH; ++q) cp_async16(dst + q * TZ_SL, src + q * 4, ok ? 16u : 0u);
        }
    };

    // resident weights ride in the first cp.async group
    {
        const float4* srcb = reinterpret_cast<const float4*>(w) + (size_t)ct * (b_bytes / 16);
        const uint32_t b_base = smem_u32(sB);
        for (int i = tid; i < b_bytes / 16; i += TZ_THREADS) cp_async16(b_base + i * 16, srcb + i, 16u);
    }
#pragma unroll
    for (int i = 0; i < TZ_STAGES - 1; ++i) {
        if (i < nit) issue(i);
        cp_async_commit();
    }

    uint32_t started = 0;
    for (int it = 0; it < nit; ++it) {
        const int nx = it + TZ_STAGES - 1;
        if (nx < nit) {
            if (it >= 1) mbar_wait(&bars[(it - 1) % TZ_STAGES], ((it - 1) / TZ_STAGES) & 1);
            issue(nx);
        }
        cp_async_commit();
        cp_async_wait<TZ_STAGES - 1>();
        fence_proxy_async_smem();
        __syncthreads();

        if (tid == 0) {
            tc_fence_after_sync();
            constexpr uint32_t idesc = make_idesc_tf32(128, NT);
            const int iz = iz_lo + it / niy, iy = iy_lo + it % niy;
            const uint32_t a_base = smem_u32(sA) + (uint32_t)(it % TZ_STAGES) * A_STAGE;
            const uint32_t b_base = smem_u32(sB);
            constexpr uint32_t plane = 3u * NT * 16;                       // bytes between the K chunks of a tap: rows [kh][n]
            constexpr uint32_t btap = (uint32_t)CH * plane;               // bytes per (kz, kw)
            // rows fed by this input row: oy = iy + 1 - kh, one contiguous range of kh
            const int kh_lo = max(0, iy + 1 - (y0 + nr - 1)), kh_hi = min(2, iy + 1 - y0);
            for (int kz = 0; kz < d.kd; ++kz) {
                const int oz = iz + pd - kz;
                if (oz < z0 || oz >= z0 + nz || kh_lo > kh_hi) continue;
                // accumulators of a slice are laid out in DEcreasing row order: the window ascends with kh
                const int nk = kh_hi - kh_lo + 1;
                const int slot0 = (oz - z0) * d.R + (nr - 1 - (iy + 1 - kh_lo - y0));
                const uint32_t wmask = ((1u << nk) - 1u) << slot0;
                const uint32_t idesc_w = make_idesc_tf32(128, nk * NT);
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
#pragma unroll
                    for (int kk = 0; kk < CS / 8; ++kk) {
                        const uint64_t ad = make_smem_desc(a_base + (uint32_t)kw * 16 + (uint32_t)(2 * kk) * TZ_SL, TZ_SL, 128);
                        const uint32_t bk = b_base + (uint32_t)(kz * 3 + kw) * btap + (uint32_t)(2 * kk) * plane;
                        if ((started & wmask) == wmask) {
                            // every accumulator of the window is initialised: ONE MMA of N = nk * NT reads A once
                            const uint64_t bd = make_smem_desc(bk + (uint32_t)(kh_lo * NT) * 16, plane, 128);
                            mma_tf32_ss(tmem + (uint32_t)slot0 * NT, ad, bd, idesc_w, 1u);
                        } else {
                            for (int kh = kh_lo; kh <= kh_hi; ++kh) {
                                const int slot_acc = slot0 + (kh - kh_lo);
                                const uint32_t acc = (started >> slot_acc) & 1u;
                                started |= 1u << slot_acc;
                                const uint64_t bd = make_smem_desc(bk + (uint32_t)(kh * NT) * 16, plane, 128);
                                mma_tf32_ss(tmem + (uint32_t)slot_acc * NT, ad, bd, idesc, acc);
                            }
                        }
                    }
                }
            }
            mma_commit(&bars[it % TZ_STAGES]);
        }
    }

    const int last = nit - 1;
    mbar_wait(&bars[last % TZ_STAGES], (last / TZ_STAGES) & 1);
    tc_fence_after_sync();
    const int ox = x0 + tid;
    const bool live = ox < d.W;
    for (int zi = 0; zi < nz; ++zi) {
        for (int ri = 0; ri < nr; ++ri) {
            float acc[NT];
#pragma unroll
            for (int c0 = 0; c0 < NT; c0 += 16) tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (zi * d.R + (nr - 1 - ri)) * NT + c0, acc + c0);
            if (!live) continue;
            const size_t o = ((((size_t)b * d.D + z0 + zi) * d.H + y0 + ri) * d.W + ox) * d.Cout + co0;
#pragma unroll
            for (int q = 0; q < NT / 4; ++q) {
                if (co0 + q * 4 >= d.Cout) break;
                float4 r = make_float4(acc[q * 4], acc[q * 4 + 1], acc[q * 4 + 2], acc[q * 4 + 3]);
                if (shift) {
                    const float4 s4 = __ldg(reinterpret_cast<const float4*>(shift + co0) + q);
                    r.x += s4.x; r.y += s4.y; r.z += s4.z; r.w += s4.w;
                }
                if (d.relu) { r.x = fmaxf(r.x, 0.f); r.y = fmaxf(r.y, 0.f); r.z = fmaxf(r.z, 0.f); r.w = fmaxf(r.w, 0.f); }
                if (skip) {
                    const float4 s4 = __ldg(reinterpret_cast<const float4*>(skip + o) + q);
                    r.x += s4.x; r.y += s4.y; r.z += s4.z; r.w += s4.w;
                }
                r.x = to_tf32(r.x); r.y = to_tf32(r.y); r.z = to_tf32(r.z); r.w = to_tf32(r.w);
                reinterpret_cast<float4*>(y + o)[q] = r;
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, tmem_cols);
}

template <int CS, int NT>
static int launch_conv_tcr_khf(const float* x, const float* w, const float* shift, const float* skip, float* y, const TrDims& d,
                           cudaStream_t st) {
    const size_t smem = (size_t)TZ_STAGES * (CS / 4) * TZ_SL + (size_t)d.kd * 9 * (CS / 4) * NT * 16 + 128;
    MVS_REQUIRE(smem <= 227 * 1024, "mvs_conv3d_tcr_khf: needs %zu bytes of shared memory", smem);
    auto kern = conv3d_tcr_khf_kernel<CS, NT>;
    MVS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int ntiles = (d.Cout + NT - 1) / NT;
    dim3 grid((unsigned)(d.nxb * ntiles), (unsigned)(d.nrb * d.nzc), (unsigned)d.B);
    MVS_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "mvs_conv3d_tcr_khf: grid too large");
    kern<<<grid, TZ_THREADS, smem, st>>>(x, w, shift, skip, y, d);
    MVS_LAUNCH_OK("conv3d_tcr_khf_kernel");
    return MVS_OK;
}

static int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

static int pick_zc(int D, int cols_per_slice, size_t a_bytes, size_t bgroup, int64_t ctas_per_slice) {
    const int max_cols = env_int("MVS_TCZ_MAX_COLS", 256);
    const int min_ctas = env_int("MVS_TCZ_MIN_CTAS", 296);
    int best_fit = 0, best_cols = 0, best_all = 0;
    for (int zc = D < 8 ? D : 8; zc >= 1; --zc) {
        if (D % zc) continue;
        if (zc * cols_per_slice > 512) continue;
        const int br = zc >= 3 ? 2 : TZ_STAGES;
        if (a_bytes + (size_t)br * bgroup + 128 > 227 * 1024) continue;
        if (!best_fit) best_fit = zc;
        const bool cols_ok = zc * cols_per_slice <= max_cols;
        if (cols_ok && !best_cols) best_cols = zc;
        if (cols_ok && ctas_per_slice * (D / zc) >= min_ctas && !best_all) best_all = zc;
    }
    return best_all ? best_all : (best_cols ? best_cols : best_fit);
}

}  // namespace kzf
}  // namespace tc
}  // namespace mvs

static int conv3d_tcz_kzf_impl(const float* x, const float* w, const float* shift, const float* skip, float* y, int B, int D,
                               int H, int W, int Cin, int Cout, int n_tile, int kd, int shw, int relu, void* stream) {
    using namespace mvs;
    using namespace mvs::tc;
    using namespace mvs::tc::kzf;
    MVS_REQUIRE(x && w && y, "mvs_conv3d_tcz_kzf: null pointer");
    MVS_REQUIRE(B >= 1 && D >= 1 && H >= 1 && W >= 1 && B <= 65535, "mvs_conv3d_tcz_kzf: bad shape");
    MVS_REQUIRE(kd == 1 || kd == 3, "mvs_conv3d_tcz_kzf: depth kernel size must be 1 or 3 (got %d)", kd);
    MVS_REQUIRE(shw == 1 || shw == 2, "mvs_conv3d_tcz_kzf: in-plane stride must be 1 or 2");
    MVS_REQUIRE(Cout % 8 == 0 && Cout >= 8, "mvs_conv3d_tcz_kzf: Cout must be a multiple of 8 (got %d)", Cout);
    const int cs = Cin >= 32 ? 32 : Cin;
    MVS_REQUIRE(Cin % cs == 0 && (cs == 8 || cs == 16 || cs == 32), "mvs_conv3d_tcz_kzf: Cin must be 8, 16 or a multiple of 32 (got %d)", Cin);
    TzDims d;
    d.B = B; d.D = D; d.H = H; d.W = W;
    d.Ho = (H - 1) / shw + 1; d.Wo = (W - 1) / shw + 1;
    d.Cin = Cin; d.Cout = Cout; d.kd = kd; d.s2 = shw == 2; d.relu = relu;
    d.PW = d.s2 ? d.Wo + 1 : d.Wo + 2;
    d.tiles_per_plane = (int)(((int64_t)d.Ho * d.PW + 127) / 128);
    const size_t a_bytes = (size_t)TZ_STAGES * (d.s2 ? 2 : 1) * (cs / 4) * TZ_SL;
    const size_t bgroup = (size_t)kd * 3 * (cs / 4) * n_tile * 16;
    d.zc = pick_zc(D, n_tile, a_bytes, bgroup, (int64_t)B * d.tiles_per_plane * ((Cout + n_tile - 1) / n_tile));
    if (d.zc == 0) MVS_UNSUPPORTED("mvs_conv3d_tcz_kzf: no depth chunking fits D=%d with N tile %d", D, n_tile);
    cudaStream_t st = (cudaStream_t)stream;
#define MVS_TZ_CASE(CS_, NT_)                                                                          \
    if (cs == CS_ && n_tile == NT_)                                                                    \
        return d.s2 ? launch_conv_tcz<CS_, NT_, true, true>(x, w, shift, skip, y, d, st)               \
                    : launch_conv_tcz<CS_, NT_, false, true>(x, w, shift, skip, y, d, st);
    MVS_TZ_CASE(8, 16)
    MVS_TZ_CASE(16, 16)
    MVS_TZ_CASE(16, 32)
    MVS_TZ_CASE(32, 16)
    MVS_TZ_CASE(32, 32)
    MVS_TZ_CASE(32, 64)
#undef MVS_TZ_CASE
    MVS_UNSUPPORTED("mvs_conv3d_tcz_kzf: no instantiation for Cin=%d, N tile=%d", Cin, n_tile);
}

// kz-fused variant (see conv3d_tcz_kernel, KZF): weights packed [Cout_tiles][3 kh][Cin/CS][3 kw][CS/4][kd][n_tile][4].
extern "C" int mvs_conv3d_tcz_kzf(const float* x, const float* w, const float* shift, const float* skip, float* y, int B,
                                  int D, int H, int W, int Cin, int Cout, int n_tile, int kd, int shw, int relu, void* stream) {
    MVS_REQUIRE(kd == 3, "mvs_conv3d_tcz_kzf: depth kernel size must be 3 (got %d)", kd);
    return conv3d_tcz_kzf_impl(x, w, shift, skip, y, B, D, H, W, Cin, Cout, n_tile, kd, shw, relu, stream);
}

// kz-fused transposed convolution (see deconv3d_tcz_kzf_kernel).
extern "C" int mvs_deconv3d_tcz_kzf(const float* x, const float* w, const float* shift, const float* skip, float* y, int B, int D,
                                int H, int W, int Cin, int Cout, int n_tile, int kd, int relu, void* stream) {
    using namespace mvs;
    using namespace mvs::tc;
    using namespace mvs::tc::kzf;
    MVS_REQUIRE(x && w && y, "mvs_deconv3d_tcz_kzf: null pointer");
    MVS_REQUIRE(B >= 1 && D >= 1 && H >= 1 && W >= 1 && B <= 65535, "mvs_deconv3d_tcz_kzf: bad shape");
    MVS_REQUIRE(kd == 3, "mvs_deconv3d_tcz_kzf: depth kernel size must be 3 (got %d)", kd);
    MVS_REQUIRE(Cout % 8 == 0 && Cout >= 8, "mvs_deconv3d_tcz_kzf: Cout must be a multiple of 8 (got %d)", Cout);
    const int cs = Cin >= 32 ? 32 : Cin;
    MVS_REQUIRE(Cin % cs == 0 && (cs == 16 || cs == 32), "mvs_deconv3d_tcz_kzf: Cin must be 16 or a multiple of 32 (got %d)", Cin);
    TzDims d;
    d.B = B; d.D = D; d.H = H; d.W = W;
    d.Ho = 2 * H; d.Wo = 2 * W;
    d.Cin = Cin; d.Cout = Cout; d.kd = kd; d.s2 = 1; d.relu = relu;
    d.PW = W + 1;
    d.tiles_per_plane = (int)(((int64_t)H * d.PW + 127) / 128);
    const size_t a_bytes = (size_t)TZ_STAGES * (cs / 4) * TZ_SL;
    const size_t bgroup = (size_t)kd * 6 * (cs / 4) * n_tile * 16;
    d.zc = pick_zc(D, 4 * n_tile, a_bytes, bgroup, (int64_t)B * d.tiles_per_plane * ((Cout + n_tile - 1) / n_tile));
    if (d.zc == 0) MVS_UNSUPPORTED("mvs_deconv3d_tcz_kzf: no depth chunking fits D=%d with N tile %d", D, n_tile);
    cudaStream_t st = (cudaStream_t)stream;
#define MVS_TZD_CASE(CS_, NT_) \
    if (cs == CS_ && n_tile == NT_) return launch_deconv_tcz_kzf<CS_, NT_>(x, w, shift, skip, y, d, st);
    MVS_TZD_CASE(16, 16)
    MVS_TZD_CASE(32, 16)
    MVS_TZD_CASE(32, 32)
#undef MVS_TZD_CASE
    MVS_UNSUPPORTED("mvs_deconv3d_tcz_kzf: no instantiation for Cin=%d, N tile=%d", Cin, n_tile);
}

// kh-fused row-tiled convolution (see conv3d_tcr_khf_kernel).
extern "C" int mvs_conv3d_tcr_khf(const float* x, const float* w, const float* shift, const float* skip, float* y, int B, int D,
                              int H, int W, int Cin, int Cout, int n_tile, int kd, int relu, void* stream) {
    using namespace mvs;
    using namespace mvs::tc;
    using namespace mvs::tc::kzf;
    MVS_REQUIRE(x && w && y, "mvs_conv3d_tcr_khf: null pointer");
    MVS_REQUIRE(B >= 1 && D >= 1 && H >= 1 && W >= 1 && B <= 65535, "mvs_conv3d_tcr_khf: bad shape");
    MVS_REQUIRE(kd == 1 || kd == 3, "mvs_conv3d_tcr_khf: depth kernel size must be 1 or 3 (got %d)", kd);
    MVS_REQUIRE(Cout % 8 == 0 && Cout >= 8, "mvs_conv3d_tcr_khf: Cout must be a multiple of 8 (got %d)", Cout);
    MVS_REQUIRE(Cin == 8 || Cin == 16 || Cin == 32, "mvs_conv3d_tcr_khf: Cin must be 8, 16 or 32 (got %d)", Cin);
    TrDims d;
    d.B = B; d.D = D; d.H = H; d.W = W; d.Cin = Cin; d.Cout = Cout; d.kd = kd; d.relu = relu;
    d.nxb = (W + 127) / 128;
    // rows x slices per CTA: accumulators within 256 TMEM columns (two CTAs can co-reside), at most 32 of them
    int zc = D < 4 ? D : 4;
    while (D % zc) --zc;
    int R = 256 / (zc * n_tile);
    if (R > 2) R = 2;            // measured: 2 rows (more, smaller CTAs) beats 4 on B200
    if (kd == 1) {               // 2D layers (visibility net): slices share nothing, spend the accumulators on rows
        zc = 1;
        R = 128 / n_tile;
        if (R > 8) R = 8;
    }
    if (R < 1) { R = 1; while (zc > 1 && zc * n_tile > 512) --zc; }
    R = env_int("MVS_TCR_ROWS", R);
    MVS_REQUIRE(R * zc * n_tile <= 512 && R * zc <= 32, "mvs_conv3d_tcr_khf: accumulators do not fit TMEM (R=%d zc=%d N=%d)", R, zc, n_tile);
    d.R = R; d.zc = zc;
    d.nrb = (H + R - 1) / R; d.nzc = (D + zc - 1) / zc;
    cudaStream_t st = (cudaStream_t)stream;
#define MVS_TR_CASE(CS_, NT_) \
    if (Cin == CS_ && n_tile == NT_) return launch_conv_tcr_khf<CS_, NT_>(x, w, shift, skip, y, d, st);
    MVS_TR_CASE(8, 16)
    MVS_TR_CASE(16, 16)
    MVS_TR_CASE(16, 32)
    MVS_TR_CASE(32, 16)
    MVS_TR_CASE(32, 32)
#undef MVS_TR_CASE
    MVS_UNSUPPORTED("mvs_conv3d_tcr_khf: no instantiation for Cin=%d, N tile=%d", Cin, n_tile);
}
