// conv3d_tma.cu — round-2 implicit-GEMM convolutions for the depth-unstrided layers (CostRegNet3D, models/module.py:550-594,
// = 93 % of the regulariser's FLOPs, and the two 3x3 tensor-core layers of the visibility net, models/mvsformer_model.py:37):
// persistent, warp-specialised, TMA-fed tcgen05 pipeline.
//
// Why (profiles/r01_conv_tensorcore_full.csv, profiles/r02c_launches.csv): the round-1 kernels are bulk-synchronous — all 128
// threads cp.async one slab, wait, __syncthreads, thread 0 issues the MMAs — one 128-thread CTA per SM, the TMEM drain never
// overlaps the next tile, every input voxel is staged three times (once per kh).  Tensor pipe 4-23 % active, 4-5x the HBM floor.
//
// Here a CTA lives for the whole launch (grid = #SMs) and runs three roles on private warps:
//   warp 0 (one lane)  TMA producer: the layer's whole weight tile once (bulk copy), then per work item one 5-D tiled TMA per
//                      input depth slab into a ring of shared-memory stages;
//   warp 1 (one lane)  MMA issuer: for every slab 9 taps x Cin/8 tcgen05.mma (kind::tf32, M = 128, N = (1..3) x NT: the depth
//                      taps of a slab feed a contiguous window of slice accumulators = one wide MMA), accumulators in TMEM,
//                      DOUBLE-BUFFERED so that item i+1 accumulates while item i is drained;
//   warps 2-9          epilogue: tcgen05.ld (one accumulator block ahead) -> shared-memory transpose -> + shift (folded BN) -> ReLU
//                      -> + skip -> TF32 round -> 128-bit stores (or, last layer, the 1x1x1 `prob` conv and one float per voxel).
// Three modes share the kernel: stride-1 conv, stride-(1,2,2) conv, and the (1,2,2) transposed conv in gather form.
// Work item = 8 x 16 output voxels (transposed: input voxels, each with its four output parity classes) x all depth slices
// x one Cout tile.  Stride 1: its input slab is the box [18 rows][10 columns] of one
// depth slice (zero-filled outside the volume = the convolution's padding), staged ONCE for all nine (kh, kw) taps exactly
// as it lies in HBM: channels-last voxels [row][column][C floats], written by TMA with the 32/64/128-byte swizzle that
// matches the voxel size.  That IS the canonical swizzled K-major operand layout (8-voxel tile rows = core groups, rows one
// voxel apart, SBO = tile-row pitch), and since the swizzle is a function of the shared-memory address for TMA and tensor
// core alike, a (kh, kw) shift is just +(kh*10 + kw) voxels on the descriptor start address — no im2col, no transposition,
// and TMA requests are whole voxels (64-128 B) instead of 16-byte pieces (the first version of this kernel staged
// [quad][column][4 floats] with 16-byte TMA rows and was request-bound: 1.4 TB/s, profiles/r02c_conv_tma_first.json).
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "tcgen05.cuh"

namespace mvs {
namespace tc {
namespace tma3 {

constexpr int THREADS = 320;            // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue (two per TMEM lane quarter)
constexpr int MODE_S1 = 0, MODE_S2 = 1, MODE_DECONV = 2;

struct Dims {
    int B, D, H, W;          // input volume
    int Ho, Wo, Cout;        // output plane size, output channels
    int relu;
    int tiles_x, tiles_y, nitems;
    int nbuf;                // TMEM accumulator buffers (2 = the epilogue of item i overlaps the MMAs of item i+1)
    float pw[8], pbias;      // PROB: the regulariser's 1x1x1 `prob` conv (8 -> 1, + bias) applied in the epilogue
};

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

// tcgen05.mma / tcgen05.commit executed by a CONVERGENT warp, one elected lane issuing: the operands stay in uniform
// registers (inside `if (lane == 0)` the compiler shuttles every descriptor through R2UR and wraps each MMA in an elect
// loop — ~12 dependent instructions, ~100 clk per MMA measured, profiles/r02c_conv_tma_s1_full.csv)
__device__ __forceinline__ void mma_tf32_ss_elect(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mma_commit_elect(uint64_t* bar) {
    asm volatile(
        "{\n\t"
        ".reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
        "}" ::"r"(smem_u32(bar)) : "memory");
}

// K-major swizzled operand: 8-row core groups `sbo` bytes apart, rows 32/64/128 B apart (implied by the layout type), the
// K chunks of a row contiguous; version 1 (sm_100), base offset 0.
__device__ __forceinline__ uint64_t make_smem_desc_sw(uint32_t saddr, uint32_t sbo_bytes, uint64_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;                                  // leading byte offset: unused for swizzled K-major
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= layout << 61;
    return d;
}

// Shared-memory plan.  A stage holds NPL planes (stride 2: input rows / columns split by parity, so that the taps of a
// parity are again unit shifts) of NSUB sub-slabs (voxel rows are at most 128 B: 64 channels = two halves), each a dense
// [PY rows][PX columns][CH floats] TMA box.
template <int MODE, int CIN, int NT, int STAGES, int NEPI = 8>
struct Smem {
    static constexpr int CQ = CIN / 4;
    static constexpr int CH = CIN >= 32 ? 32 : CIN;          // channels per sub-slab: voxel rows of 32 / 64 / 128 bytes
    static constexpr int NSUB = CIN / CH;
    static constexpr int NPL = MODE == MODE_S2 ? 4 : 1;
    static constexpr int PX = MODE == MODE_S1 ? 10 : 9;      // columns per plane row (8 + halo)
    static constexpr int PY = MODE == MODE_S1 ? 18 : 17;     // rows per plane (16 + halo)
    static constexpr int ROWB = CH * 4;                      // bytes between the voxels of a tile row (fixed by the swizzle mode)
    static constexpr int TROW = PX * ROWB;                   // bytes between tile rows (SBO: tile rows are the 8-voxel core groups)
    static constexpr int SUB_TX = PY * TROW;                 // bytes one TMA box delivers
    static constexpr int SUB = (SUB_TX + 1023) / 1024 * 1024;
    static constexpr int PLANE = NSUB * SUB;
    static constexpr int SLAB = NPL * PLANE;                 // bytes per stage
    static constexpr uint64_t LAYOUT = ROWB == 128 ? 2 : (ROWB == 64 ? 4 : 6);   // UMMA layout type: SWIZZLE_128B / 64B / 32B
    static constexpr int STG_PITCH = NT + 4;                 // floats per staged accumulator row (epilogue transpose)
    static constexpr int STG = NEPI * 32 * STG_PITCH * 4;    // bytes: one [32][NT+4] buffer per epilogue warp
    static __host__ __device__ constexpr int wbytes(int kd) { return 9 * CQ * kd * NT * 16; }
    static __host__ __device__ constexpr size_t total(int kd) { return (size_t)STAGES * SLAB + wbytes(kd) + STG + 256 + 1024; }
};

// x channels-last [B,D,H,W,CIN]; y [B,D,Ho,Wo,Cout]; w packed [Cout tiles][kh][kw][CIN/4][kd][NT][4] (TF32-rounded, BN folded)
// NEPI = epilogue warps (8, or 4 where the stages leave no room for eight staging buffers)
// PROB (Cout = 8, the regulariser's last layer): the 8-channel result never reaches HBM — the epilogue applies the 1x1x1
// `prob` conv + bias (models/module.py:582,593) to it and writes prob_volume_pre [B,D,Ho,Wo] through `y`: 1/8 of the stores,
// no second pass over the tensor.  Same arithmetic as the two-kernel route (TF32-rounded activations, the FMA chain of
// prob_conv1_kernel in channel order), so the results are bit-identical.
template <int MODE, int CIN, int NT, int STAGES, int KD, int NEPI, bool PROB = false>
__global__ void __launch_bounds__(THREADS, 1)
conv3d_tma_kernel(const __grid_constant__ CUtensorMap amap, const float* __restrict__ w, const float* __restrict__ shift,
                  const float* __restrict__ skip, float* __restrict__ y, Dims d) {
    using L = Smem<MODE, CIN, NT, STAGES, NEPI>;
    constexpr int NCLS = MODE == MODE_DECONV ? 4 : 1;                 // output parity classes per input voxel
    constexpr int pd = KD / 2;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // stages 1024-byte aligned (the swizzle pattern repeats every 8 rows); the offset is computed on the shared-space
    // address so that the pointers keep their address space
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sA = smem;
    uint8_t* sW = smem + STAGES * L::SLAB;
    float* sStage = reinterpret_cast<float*>(sW + L::wbytes(KD));
    uint64_t* full = reinterpret_cast<uint64_t*>(sW + L::wbytes(KD) + L::STG);
    uint64_t* empty = full + STAGES;
    uint64_t* wbar = empty + STAGES;
    uint64_t* acc_full = wbar + 1;
    uint64_t* acc_empty = acc_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    // warp index broadcast from lane 0: the compiler then knows it is warp-uniform, keeps the role branches and everything
    // derived inside them (descriptors, TMEM addresses) on the uniform datapath — the MMA issue loop is the pipeline's pace-maker
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
    const int ct = blockIdx.y, co0 = ct * NT;
    const int ncols = NCLS * d.D * NT;                                // accumulator columns of one buffer
    const int alloc = d.nbuf * ncols;
    const uint32_t tmem_cols = alloc <= 32 ? 32 : (alloc <= 64 ? 64 : (alloc <= 128 ? 128 : (alloc <= 256 ? 256 : 512)));

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(wbar, 1);
        mbar_init(&acc_full[0], 1); mbar_init(&acc_full[1], 1);
        mbar_init(&acc_empty[0], NEPI * 32); mbar_init(&acc_empty[1], NEPI * 32);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);

    if (warp == 0) {
        // ===================================================== producer ==============================================
        if (lane == 0) {
            mbar_expect_tx(wbar, L::wbytes(KD));
            bulk_load(smem_u32(sW), w + (size_t)ct * (L::wbytes(KD) / 4), L::wbytes(KD), wbar);
            int n = 0;                                                // slabs issued so far
            for (int item = blockIdx.x; item < d.nitems; item += gridDim.x) {
                const int tx = item % d.tiles_x, ty = (item / d.tiles_x) % d.tiles_y, b = item / (d.tiles_x * d.tiles_y);
                for (int iz = 0; iz < d.D; ++iz, ++n) {
                    const int s = n % STAGES;
                    if (n >= STAGES) mbar_wait(&empty[s], ((n / STAGES) - 1) & 1);
                    mbar_expect_tx(&full[s], L::NPL * L::NSUB * L::SUB_TX);
#pragma unroll
                    for (int pl = 0; pl < L::NPL; ++pl) {
                        // stride 2: plane = 2 * (row parity class) + (column parity class); class 0 = the odd input rows /
                        // columns 2t-1 (taps 0 and 2), class 1 = the even ones 2t (tap 1)
                        const int x0 = MODE == MODE_S1 ? tx * 8 - 1 : (MODE == MODE_S2 ? 2 * tx * 8 - 1 + (pl & 1) : tx * 8);
                        const int y0 = MODE == MODE_S1 ? ty * 16 - 1 : (MODE == MODE_S2 ? 2 * ty * 16 - 1 + (pl >> 1) : ty * 16);
#pragma unroll
                        for (int h = 0; h < L::NSUB; ++h)
                            tma_load_4d(smem_u32(sA) + s * L::SLAB + pl * L::PLANE + h * L::SUB, &amap, &full[s], h * L::CH, x0, y0,
                                        b * d.D + iz);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer (whole warp, convergent) =====================
        mbar_wait(wbar, 0);
        // The issuing thread is the pipeline's pace-maker.  Everything address-like is folded into compile-time offsets on
        // descriptors built once: a descriptor's start-address field is its low 14 bits in 16-byte units.
        constexpr uint32_t idesc1 = make_idesc_tf32(128, NT);
        constexpr uint32_t plane16 = (uint32_t)(KD * NT);             // 16-byte units between the K chunks of a tap: rows [kz][n]
        const uint64_t bdesc0 = make_smem_desc(smem_u32(sW), plane16 * 16, 128);
        const uint64_t adesc0 = make_smem_desc_sw(smem_u32(sA), L::TROW, L::LAYOUT);
        int n = 0, it = 0;
        for (int item = blockIdx.x; item < d.nitems; item += gridDim.x, ++it) {
            const int buf = d.nbuf == 2 ? (it & 1) : 0, use = d.nbuf == 2 ? (it >> 1) : it;
            if (use >= 1) mbar_wait(&acc_empty[buf], (use - 1) & 1);
            tc_fence_after_sync();
            const uint32_t tacc = tmem + (uint32_t)buf * ncols;
            uint32_t started = 0;                                     // per (class, slice) "accumulator written" bits
            for (int iz = 0; iz < d.D; ++iz, ++n) {
                const int s = n % STAGES;
                mbar_wait(&full[s], (n / STAGES) & 1);
                tc_fence_after_sync();
                const uint64_t adesc = adesc0 + (uint64_t)(s * (L::SLAB >> 4));
                // Depth taps whose output slice exists: one contiguous range, fed by ONE MMA of N = (taps) x NT.  conv:
                // oz = iz + pd - kz, slices laid out in DEcreasing order; transposed: oz = iz - pd + kz, INcreasing order —
                // either way the accumulator columns of the window ascend with kz.
                const int kz_lo = MODE == MODE_DECONV ? max(0, pd - iz) : max(0, iz + pd - (d.D - 1));
                const int kz_hi = MODE == MODE_DECONV ? min(KD - 1, d.D - 1 - iz + pd) : min(KD - 1, iz + pd);
                const int blk_lo = MODE == MODE_DECONV ? iz - pd + kz_lo : d.D - 1 - (iz + pd - kz_lo);   // column block of tap kz_lo
                const uint32_t idw = make_idesc_tf32(128, 0) | ((uint32_t)(((kz_hi - kz_lo + 1) * NT) >> 3) << 17);
                const uint64_t bdesc = bdesc0 + (uint64_t)(kz_lo * NT);
#pragma unroll
                for (int tap = 0; tap < 9; ++tap) {
                    constexpr int dummy = 0; (void)dummy;
                    const int kh = tap / 3, kw = tap % 3;
                    // where the tap's operand starts inside the stage, and which class of accumulators it feeds
                    const int pl = MODE == MODE_S2 ? 2 * (kh == 1) + (kw == 1) : 0;
                    const int rsh = MODE == MODE_S1 ? kh : (MODE == MODE_S2 ? (kh == 2) : (kh == 0));
                    const int csh = MODE == MODE_S1 ? kw : (MODE == MODE_S2 ? (kw == 2) : (kw == 0));
                    const int cls = MODE == MODE_DECONV ? 2 * (kh != 1) + (kw != 1) : 0;
                    // the first tap (in issue order) of a class initialises its accumulators
                    const bool starter = MODE == MODE_DECONV ? (tap == 0 || tap == 1 || tap == 3 || tap == 4) : tap == 0;
                    const uint32_t cbase = tacc + (uint32_t)(cls * d.D) * NT;
#pragma unroll
                    for (int kk = 0; kk < CIN / 8; ++kk) {
                        const uint32_t aoff = (uint32_t)(pl * L::PLANE + ((kk * 8) / L::CH) * L::SUB + (rsh * L::PX + csh) * L::ROWB + ((kk * 8) % L::CH) * 4);
                        const uint64_t ad = adesc + (uint64_t)(aoff >> 4);
                        const uint64_t bd = bdesc + (uint64_t)((uint32_t)(tap * L::CQ + 2 * kk) * plane16);
                        if (starter && kk == 0) {
                            // per slice, each with its own accumulate flag (slices this slab is the first to touch start here)
                            for (int kz = kz_lo; kz <= kz_hi; ++kz) {
                                const int blk = blk_lo + (kz - kz_lo);
                                const int bit = cls * d.D + blk;
                                const uint32_t acc = (started >> bit) & 1u;
                                started |= 1u << bit;
                                mma_tf32_ss_elect(cbase + (uint32_t)blk * NT, ad, bd + (uint64_t)((kz - kz_lo) * NT), idesc1, acc);
                            }
                        } else {
                            mma_tf32_ss_elect(cbase + (uint32_t)blk_lo * NT, ad, bd, idw, 1u);
                        }
                    }
                }
                mma_commit_elect(&empty[s]);                          // stage free once these MMAs have read it
            }
            mma_commit_elect(&acc_full[buf]);                         // accumulators of this item complete
        }
    } else if (warp < 2 + NEPI) {
        // ===================================================== epilogue (warps 2-9) ===================================
        // Two warps per TMEM lane quarter: they take the even / odd accumulator blocks (class, slice) of an item.
        // A warp owns the TMEM lane quarter q = 32 accumulator rows = 4 tile rows x 8 voxels.  tcgen05.ld hands every lane
        // one whole row; stored like that, a warp instruction would touch 32 half-filled sectors 64-128 B apart (measured
        // store-bound at n_tile 32).  The rows go through a per-warp shared-memory buffer instead, so that a store
        // instruction writes NT/4 lanes per voxel: 8 consecutive voxels of a tile row per instruction.
        constexpr int CPR = NT / 4;                                   // 16-byte chunks per accumulator row
        // chunks per row that exist in the output (a 16-wide MMA tile may carry only 8 real channels): lanes are spread
        // over the real chunks only, so that no store slot is wasted on padding
        const int cpr = min(NT, d.Cout - co0) / 4, rpi = 32 / cpr;    // cpr in {1, 2, 4, 8}; rows per store instruction
        const int q = warp & 3;                                       // TMEM lane quarter this warp may read
        const int half = (warp - 2) >> 2;                             // which blocks of the item this warp drains
        float* stg = sStage + (half * 4 + q) * 32 * L::STG_PITCH;
        const int chunk = lane % cpr, rsub = lane / cpr;
        const bool cvalid = co0 + chunk * 4 < d.Cout;
        const float4 sh = (shift && cvalid) ? __ldg(reinterpret_cast<const float4*>(shift + co0) + chunk) : make_float4(0.f, 0.f, 0.f, 0.f);
        // PROB (cpr = 2): the `prob` weights of this lane's four channels
        const float4 pw = (PROB && chunk == 1) ? make_float4(d.pw[4], d.pw[5], d.pw[6], d.pw[7]) : make_float4(d.pw[0], d.pw[1], d.pw[2], d.pw[3]);
        int it = 0;
        for (int item = blockIdx.x; item < d.nitems; item += gridDim.x, ++it) {
            const int buf = d.nbuf == 2 ? (it & 1) : 0, use = d.nbuf == 2 ? (it >> 1) : it;
            const int tx = item % d.tiles_x, ty = (item / d.tiles_x) % d.tiles_y, b = item / (d.tiles_x * d.tiles_y);
            // Output voxel this lane writes in store instruction j of accumulator block cz (class, slice) = block origin
            // (one 64-bit product per block) + the lane's pixel offset inside a slice (per item, 32-bit): the address math
            // of the drain loop is two adds per store.
            int pixoff[CPR];
            unsigned livemask = 0;
#pragma unroll
            for (int j = 0; j < CPR; ++j) {
                const int row = j * rpi + rsub;                       // accumulator row within the quarter
                const int ty_v = ty * 16 + q * 4 + (row >> 3), tx_v = tx * 8 + (row & 7);
                // transposed: tile voxel = input voxel, the class adds the output parity
                const bool live = MODE == MODE_DECONV ? (ty_v < d.H && tx_v < d.W) : (ty_v < d.Ho && tx_v < d.Wo);
                pixoff[j] = MODE == MODE_DECONV ? 2 * ty_v * d.Wo + 2 * tx_v : ty_v * d.Wo + tx_v;
                if (live && cvalid && j < cpr) livemask |= 1u << j;
            }
            const size_t plane = (size_t)d.Ho * d.Wo;
            auto block_base = [&](int cz) -> size_t {
                const int cls = cz / d.D, blk = cz - cls * d.D;
                const int z = MODE == MODE_DECONV ? blk : d.D - 1 - blk;
                return ((size_t)b * d.D + z) * plane + (MODE == MODE_DECONV ? (size_t)((cls >> 1) * d.Wo + (cls & 1)) : (size_t)0);
            };
            // The skip tensor does not depend on the MMAs: its loads run one accumulator block ahead (the first block's before
            // the wait on the accumulators), so their latency never sits between a TMEM drain and its stores.
            float4 skn[CPR];
            size_t base_n = 0;
            auto prefetch_skip = [&](int cz) {
                base_n = block_base(cz);
#pragma unroll
                for (int j = 0; j < CPR; ++j)
                    skn[j] = (skip && ((livemask >> j) & 1u)) ? __ldg(reinterpret_cast<const float4*>(skip + (base_n + pixoff[j]) * d.Cout + co0 + chunk * 4))
                                                              : make_float4(0.f, 0.f, 0.f, 0.f);
            };
            const int nblk = NCLS * d.D;
            if (half < nblk) prefetch_skip(half);
            mbar_wait(&acc_full[buf], use & 1);
            tc_fence_after_sync();
            const uint32_t tacc = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)buf * ncols;
            uint32_t accr[NT];
            auto load_block = [&](int cz) {
#pragma unroll
                for (int c0 = 0; c0 < NT; c0 += 16) tmem_ld16_async(tacc + (uint32_t)cz * NT + c0, accr + c0);
            };
            if (half < nblk) load_block(half);
#pragma unroll 1
            for (int cz = half; cz < nblk; cz += NEPI / 4) {
                float4 skc[CPR];
#pragma unroll
                for (int j = 0; j < CPR; ++j) skc[j] = skn[j];
                const size_t base_c = base_n;
                if (cz + NEPI / 4 < nblk) prefetch_skip(cz + NEPI / 4);
                // The block's accumulator rows were requested one block ago (or before the loop): the TMEM read latency runs
                // under the previous block's stores.  Once they are staged in shared memory the registers are free, and the
                // next block's read is issued before this block's stores.  (Measured neutral on B200 — the drain is bound by
                // its shared-memory round trip and stores, profiles/r02k_ab.json — kept because it is never slower.)
                tmem_ld_wait();
                __syncwarp();                                         // previous block's reads of the buffer are done
#pragma unroll
                for (int c4 = 0; c4 < CPR; ++c4)
                    *reinterpret_cast<uint4*>(stg + lane * L::STG_PITCH + c4 * 4) = make_uint4(accr[c4 * 4], accr[c4 * 4 + 1], accr[c4 * 4 + 2], accr[c4 * 4 + 3]);
                __syncwarp();
                if (cz + NEPI / 4 < nblk) load_block(cz + NEPI / 4);
#pragma unroll
                for (int j = 0; j < CPR; ++j) {
                    const bool ok = (livemask >> j) & 1u;
                    if (PROB ? j >= cpr : !ok) continue;              // PROB: warp-uniform (every lane takes part in the shuffle)
                    const size_t vox = base_c + pixoff[j], o = vox * d.Cout + co0 + chunk * 4;
                    float4 r = *reinterpret_cast<const float4*>(stg + (j * rpi + rsub) * L::STG_PITCH + chunk * 4);
                    r.x += sh.x; r.y += sh.y; r.z += sh.z; r.w += sh.w;
                    if (d.relu) { r.x = fmaxf(r.x, 0.f); r.y = fmaxf(r.y, 0.f); r.z = fmaxf(r.z, 0.f); r.w = fmaxf(r.w, 0.f); }
                    r.x += skc[j].x; r.y += skc[j].y; r.z += skc[j].z; r.w += skc[j].w;
                    // consumers read this tensor as a TF32 operand: round once here
                    r.x = to_tf32(r.x); r.y = to_tf32(r.y); r.z = to_tf32(r.z); r.w = to_tf32(r.w);
                    if (PROB) {
                        // the two lanes of a voxel (channels 0-3, 4-7) run ONE chain bias -> c0 .. c7: the even lane's partial
                        // sum is handed to the odd lane, which finishes and stores
                        float acc = d.pbias;
                        acc = fmaf(r.x, pw.x, acc); acc = fmaf(r.y, pw.y, acc); acc = fmaf(r.z, pw.z, acc); acc = fmaf(r.w, pw.w, acc);
                        float fin = __shfl_sync(0xffffffffu, acc, lane & ~1);
                        fin = fmaf(r.x, pw.x, fin); fin = fmaf(r.y, pw.y, fin); fin = fmaf(r.z, pw.z, fin); fin = fmaf(r.w, pw.w, fin);
                        if (ok && chunk == 1) y[vox] = fin;
                    } else {
                        *reinterpret_cast<float4*>(y + o) = r;
                    }
                }
            }
            tc_fence_before_sync();
            mbar_arrive(&acc_empty[buf]);                             // 256 arrivals: buffer may be overwritten
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, tmem_cols);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

// channels-last activation [BD][H][W][C] as (c, x, y, batch*depth): a box of `ch` channels x `nx` columns x `ny` rows (every
// `step`-th column / row) lands in shared memory as [row][column][ch floats] with the swizzle mode that matches ch*4 bytes
static int make_act_map(CUtensorMap* map, const float* x, int BD, int H, int W, int C, int ch, int nx, int ny, int step) {
    EncodeTiledFn fn = encode_fn();
    MVS_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
    const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)BD};
    const cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
    const cuuint32_t box[4] = {(cuuint32_t)ch, (cuuint32_t)((nx - 1) * step + 1), (cuuint32_t)((ny - 1) * step + 1), 1};
    const cuuint32_t estr[4] = {1, (cuuint32_t)step, (cuuint32_t)step, 1};
    const CUtensorMapSwizzle sw = ch * 4 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (ch * 4 == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    CUresult rc = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MVS_REQUIRE(rc == CUDA_SUCCESS, "cuTensorMapEncodeTiled (activation map) failed with code %d", (int)rc);
    return MVS_OK;
}

static int sm_count() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

template <int MODE, int CIN, int NT, int STAGES, int KD, int NEPI, bool PROB = false>
static int launch_k(const float* x, const float* w, const float* shift, const float* skip, float* y, int B, int D, int H, int W,
                    int Cout, int relu, cudaStream_t st, const float* prob_w = nullptr, float prob_bias = 0.0f) {
    using L = Smem<MODE, CIN, NT, STAGES, NEPI>;
    const size_t smem = L::total(KD);
    MVS_REQUIRE(smem <= 227 * 1024, "mvs_conv3d_tma: needs %zu bytes of shared memory", smem);
    CUtensorMap map;
    int rc = make_act_map(&map, x, B * D, H, W, CIN, L::CH, L::PX, L::PY, MODE == MODE_S2 ? 2 : 1);
    if (rc) return rc;
    Dims d = {};
    d.B = B; d.D = D; d.H = H; d.W = W; d.Cout = Cout; d.relu = relu;
    if (PROB) {
        for (int i = 0; i < 8; ++i) d.pw[i] = prob_w[i];
        d.pbias = prob_bias;
    }
    d.Ho = MODE == MODE_S2 ? (H + 1) / 2 : (MODE == MODE_DECONV ? 2 * H : H);
    d.Wo = MODE == MODE_S2 ? (W + 1) / 2 : (MODE == MODE_DECONV ? 2 * W : W);
    const int th = MODE == MODE_DECONV ? H : d.Ho, tw = MODE == MODE_DECONV ? W : d.Wo;      // the tiled plane
    d.tiles_x = cdiv(tw, 8); d.tiles_y = cdiv(th, 16);
    d.nitems = B * d.tiles_x * d.tiles_y;
    const int ncols = (MODE == MODE_DECONV ? 4 : 1) * D * NT;
    MVS_REQUIRE(ncols <= 512, "mvs_conv3d_tma: %d accumulator columns exceed the tensor memory", ncols);
    d.nbuf = 2 * ncols <= 512 ? 2 : 1;
    const int ntiles = cdiv(Cout, NT);
    auto kern = conv3d_tma_kernel<MODE, CIN, NT, STAGES, KD, NEPI, PROB>;
    MVS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int gx = sm_count() / ntiles;
    if (gx < 1) gx = 1;
    if (gx > d.nitems) gx = d.nitems;
    kern<<<dim3((unsigned)gx, (unsigned)ntiles), THREADS, smem, st>>>(map, w, shift, skip, y, d);
    MVS_LAUNCH_OK("conv3d_tma_kernel");
    return MVS_OK;
}

template <int MODE, int CIN, int NT, int STAGES, int NEPI>
static int launch(const float* x, const float* w, const float* shift, const float* skip, float* y, int B, int D, int H, int W,
                  int Cout, int kd, int relu, cudaStream_t st) {
    return kd == 1 ? launch_k<MODE, CIN, NT, STAGES, 1, NEPI>(x, w, shift, skip, y, B, D, H, W, Cout, relu, st)
                   : launch_k<MODE, CIN, NT, STAGES, 3, NEPI>(x, w, shift, skip, y, B, D, H, W, Cout, relu, st);
}

}  // namespace tma3
}  // namespace tc
}  // namespace mvs

// x [B,D,H,W,Cin] channels-last (TF32-rounded values), w packed by the host ([Cout tiles][kh][kw][Cin/4][kd][n_tile][4]),
// shift [Cout] or NULL, skip / y [B,D,Ho,Wo,Cout].  Depth stride 1, kernel (kd,3,3), padding (kd/2,1,1).
// mode 0: stride 1; mode 1: stride (1,2,2), Ho = ceil(H/2); mode 2: transposed, stride (1,2,2), output_padding (0,1,1), Ho = 2H.
extern "C" int mvs_conv3d_tma(const float* x, const float* w, const float* shift, const float* skip, float* y, int B, int D,
                              int H, int W, int Cin, int Cout, int n_tile, int kd, int mode, int relu, void* stream) {
    using namespace mvs::tc::tma3;
    MVS_REQUIRE(x && w && y, "mvs_conv3d_tma: null pointer");
    MVS_REQUIRE(B >= 1 && D >= 1 && H >= 1 && W >= 1, "mvs_conv3d_tma: empty shape");
    MVS_REQUIRE(kd == 1 || kd == 3, "mvs_conv3d_tma: kernel depth must be 1 or 3 (got %d)", kd);
    MVS_REQUIRE(mode >= 0 && mode <= 2, "mvs_conv3d_tma: mode must be 0 (stride 1), 1 (stride 2) or 2 (transposed), got %d", mode);
    MVS_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0 && ((uintptr_t)w & 15) == 0 && (!skip || ((uintptr_t)skip & 15) == 0),
                "mvs_conv3d_tma: tensors must be 16-byte aligned");
    MVS_REQUIRE(Cout % 4 == 0, "mvs_conv3d_tma: Cout must be a multiple of 4 (got %d)", Cout);
    if ((mode == 2 ? 4 : 1) * D * n_tile > 512)
        MVS_UNSUPPORTED("mvs_conv3d_tma: %d depth slices x n_tile %d exceed the tensor memory", D, n_tile);
    cudaStream_t st = (cudaStream_t)stream;
#define MVS_TMA_CASE(M, C, N, S, E) if (mode == M && Cin == C && n_tile == N) return launch<M, C, N, S, E>(x, w, shift, skip, y, B, D, H, W, Cout, kd, relu, st)
    //           mode         Cin  NT stages epilogue warps
    MVS_TMA_CASE(MODE_S1,     16, 16, 6, 8);
    MVS_TMA_CASE(MODE_S1,     32, 32, 3, 8);
    MVS_TMA_CASE(MODE_S1,     64, 16, 2, 8);
    MVS_TMA_CASE(MODE_S2,      8, 16, 6, 8);
    MVS_TMA_CASE(MODE_S2,     16, 32, 3, 8);
    MVS_TMA_CASE(MODE_S2,     32, 16, 2, 4);
    MVS_TMA_CASE(MODE_DECONV, 64, 16, 2, 8);
    MVS_TMA_CASE(MODE_DECONV, 32, 16, 4, 8);
    MVS_TMA_CASE(MODE_DECONV, 16, 16, 6, 8);
#undef MVS_TMA_CASE
    MVS_UNSUPPORTED("mvs_conv3d_tma: no instantiation for mode %d, Cin=%d, n_tile=%d", mode, Cin, n_tile);
}

// The regulariser's last two layers in one kernel (CostRegNet3D, models/module.py:575,582,592-593): transposed conv 16 -> 8
// (+ folded BN shift, ReLU, + skip) and the 1x1x1 `prob` conv 8 -> 1 (+ bias) applied to its result in the epilogue.
// x [B,D,H,W,16] channels-last; w packed as for mvs_conv3d_tma mode 2 (n_tile 16); skip [B,D,2H,2W,8] or NULL;
// prob_w_host [8] / prob_bias: host values; pre [B,D,2H,2W].  Bit-identical to mvs_conv3d_tma + mvs_prob_conv_cl.
extern "C" int mvs_conv3d_tma_prob(const float* x, const float* w, const float* shift, const float* skip, const float* prob_w_host,
                                   float prob_bias, float* pre, int B, int D, int H, int W, int Cin, int kd, int relu, void* stream) {
    using namespace mvs::tc::tma3;
    MVS_REQUIRE(x && w && pre && prob_w_host, "mvs_conv3d_tma_prob: null pointer");
    MVS_REQUIRE(B >= 1 && D >= 1 && H >= 1 && W >= 1, "mvs_conv3d_tma_prob: empty shape");
    MVS_REQUIRE(kd == 1 || kd == 3, "mvs_conv3d_tma_prob: kernel depth must be 1 or 3 (got %d)", kd);
    MVS_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)w & 15) == 0 && (!skip || ((uintptr_t)skip & 15) == 0),
                "mvs_conv3d_tma_prob: tensors must be 16-byte aligned");
    if (Cin != 16) MVS_UNSUPPORTED("mvs_conv3d_tma_prob: only Cin = 16 (-> 8 -> 1) is built (got %d)", Cin);
    if (4 * D * 16 > 512) MVS_UNSUPPORTED("mvs_conv3d_tma_prob: %d depth slices exceed the tensor memory", D);
    cudaStream_t st = (cudaStream_t)stream;
    return kd == 1 ? launch_k<MODE_DECONV, 16, 16, 6, 1, 8, true>(x, w, shift, skip, pre, B, D, H, W, 8, relu, st, prob_w_host, prob_bias)
                   : launch_k<MODE_DECONV, 16, 16, 6, 3, 8, true>(x, w, shift, skip, pre, B, D, H, W, 8, relu, st, prob_w_host, prob_bias);
}
