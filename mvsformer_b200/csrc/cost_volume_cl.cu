// cost_volume_cl.cu — round-2 production cost-volume kernels: fused homography warp + group-wise correlation
// + entropy / visibility-weighted aggregation (models/mvsformer_model.py:61-105, models/warping.py:84-107)
// over CHANNELS-LAST feature maps  [B*V][H][W][C]  (made once per feature set by nchw_to_cl_kernel below).
//
// Why a second generation (profiles/r01_cost_volume_tma_full.csv): the first TMA kernel sampled channel-planar
// tiles with 4 scalar LDS.32 per channel-tap at 1.5-1.8 shared-memory wavefronts each (lanes = neighbouring
// pixels, whose sample positions scatter), on a single-buffered tile with the TMA round trip exposed.  Here
//   * a tap is ONE LDS.128 for four channels (a "chunk") of a channels-last texel;
//   * a thread is one sample (pixel, hypothesis) and owns all channels, so every reduction (over c' for the
//     group correlation, over g for the eval-only cosine similarity) is thread-local — no shuffles;
//   * lanes read their chunks in a lane-dependent XOR order  q = j ^ x(lane):  the 8 lanes of an LDS.128 phase
//     hit 8 different 16-byte bank groups whatever their texels are (texel size >= 128 B, stages 1-2), or are
//     the hypotheses of ONE pixel, whose texels are neighbours along the epipolar line (stages 3-4) — bank
//     conflicts no longer depend on how noisy the depth map is.  Registers keep a static "local" chunk order;
//     the XOR only enters addresses and a 3-stage select network on the 8 per-group outputs;
//   * work is cut into items (source view, block of HB hypotheses, channel half): one hypothesis per thread
//     per item keeps the live state small, so the NEXT item's box is projected, bounded and fetched by TMA
//     into the second tile buffer while this item is sampled (one __syncthreads per item, no exposed TMA).
// Out-of-box samples (wild geometry) take a predicated global path with the reference's per-tap bounds
// checks, so results never depend on the box size.  The N x C x D x H x W warped tensor, the sampling grid
// and (at stage 4) the per-view correlation never exist in HBM.
#include <cuda.h>
#include <float.h>
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "tcgen05.cuh"

namespace mvs {
namespace k1cl {

using tc::mbar_fence_init;
using tc::mbar_init;
using tc::mbar_wait;
using tc::smem_u32;

struct Params {
    const float* feat;        // channels-last [B*V][H][W][C]
    const float* relproj;     // [B, N, 12]
    const float* depth;       // [B, D, H, W]
    int N, V, H, W;
    float* entropy;           // pass A out [B, N, H, W]
    float* sim_depth;         // pass A out [B, H, W] (SIM): hypothesis with the largest cosine similarity summed over views
    float* corr;              // pass A out [B, N, D, H, W, 8] (STORE)
    const float* vis_weight;  // pass B in  [B, N, H, W]
    float* volume;            // pass B out [B, D, H, W, 8]
    int round_tf32;
    int use_slots;            // the views of the (single) batch item are maps slot[0] (reference), slot[1..N] of `feat`
    int slot[17];
};

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

// a / b with a correctly rounded reciprocal rb = RN(1/b): q = RN(a*rb); r = a - b*q (exact, one FMA); RN(q + r*rb) —
// Markstein's sequence, bit-identical to __fdiv_rn for operands in the normal range (samples outside it fail the
// `sane` test and take the global path, which uses plain IEEE divisions in make_taps).
__device__ __forceinline__ float div_by(float a, float b, float rb) {
    const float q = __fmul_rn(a, rb);
    const float r = __fmaf_rn(-b, q, a);
    return __fmaf_rn(r, rb, q);
}

__device__ __forceinline__ RelProj load_relproj_smem(const float* q) {
    RelProj m;
    m.r00 = q[0]; m.r01 = q[1]; m.r02 = q[2];  m.t0 = q[3];
    m.r10 = q[4]; m.r11 = q[5]; m.r12 = q[6];  m.t1 = q[7];
    m.r20 = q[8]; m.r21 = q[9]; m.r22 = q[10]; m.t2 = q[11];
    return m;
}

struct Geo {   // per-kernel constants of the normalise / un-normalise round trip (warping.py:94-95 + grid_sample)
    float half_w, half_h, rcp_half_w, rcp_half_h, wm1, hm1;
};

// Sample position of (pixel ray, depth): the arithmetic of make_taps() in geometry.cuh, operation by operation.
__device__ __forceinline__ void project(const RelProj& m, const PixelRay& ray, float depth, const Geo& g, float* ix, float* iy) {
    const float qx = __fadd_rn(__fmul_rn(ray.x, depth), m.t0);
    const float qy = __fadd_rn(__fmul_rn(ray.y, depth), m.t1);
    const float qz = __fadd_rn(__fmul_rn(ray.z, depth), m.t2);
    const float den = __fadd_rn(qz, 1e-6f);
    const float rden = __frcp_rn(den);
    const float gx = __fadd_rn(div_by(div_by(qx, den, rden), g.half_w, g.rcp_half_w), -1.0f);
    const float gy = __fadd_rn(div_by(div_by(qy, den, rden), g.half_h, g.rcp_half_h), -1.0f);
    *ix = __fmul_rn(__fmul_rn(__fadd_rn(gx, 1.0f), 0.5f), g.wm1);
    *iy = __fmul_rn(__fmul_rn(__fadd_rn(gy, 1.0f), 0.5f), g.hm1);
}

// out[i ^ mask] = v[i] for the bits of `mask` below NB: a butterfly of conditional swaps (selects, no memory).
template <int NV, int NB>
__device__ __forceinline__ void xor_permute(float (&v)[NV], int mask) {
#pragma unroll
    for (int bit = 0; bit < NB; ++bit) {
        const bool sw = (mask >> bit) & 1;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            if ((i >> bit) & 1) continue;
            const float a = v[i], b = v[i | (1 << bit)];
            v[i] = sw ? b : a;
            v[i | (1 << bit)] = sw ? a : b;
        }
    }
}

// Per-stage tiling.  TW x TH pixel tile; HB threads per pixel with KPT hypotheses each = HB*KPT hypotheses per item
// (threads = TW*TH*HB = 256; more hypotheses per thread amortise the per-item work — ray, box, waits, address math — where a
// sample has few channels), BW x BH texel box per
// item (sized from measured coverage, scripts/box_coverage_cl.py), NCH channel halves (stage 1: the 64-channel texel
// is fetched as two 128-byte halves so that two tile buffers of two CTAs fit one SM).
template <int C_, int D_, int TW_, int TH_, int HB_, int KPT_, int BW_, int BH_, int NCH_, int MINB_>
struct Cfg {
    static constexpr int C = C_, D = D_, TW = TW_, TH = TH_, HB = HB_, KPT = KPT_, BW = BW_, BH = BH_, NCH = NCH_, MINB = MINB_;
    static constexpr int G = 8, CPG = C / G, TP = TW * TH, HPI = HB * KPT, NHB = D / HPI, CC = C / NCH, NQ = CC / 4, GPI = CC / CPG;
    static constexpr int TILE_F = BW * BH * CC;                        // floats per tile buffer
    static constexpr bool REF_SMEM = C > 32;                           // reference features: shared memory or registers
    // per-pixel stride of the [pixel][hypothesis] tables (keeps a warp's accesses on different banks)
    static constexpr int SD = KPT > 1 ? D + 1 : (D == 32 ? 40 : (D == 16 ? 24 : D));
    static_assert(TP * HB == 256 && D % HPI == 0 && C % (4 * NCH) == 0 && (TILE_F * 4) % 128 == 0, "tiling");
    static_assert(NQ == 8 || NQ == 4 || NQ == 2, "chunks per texel");
};

constexpr int MAXN = 16;          // source views whose relative projections are staged in shared memory

template <class CFG, int MODE, bool SIM>
constexpr size_t smem_bytes() {
    return (size_t)(2 * CFG::TILE_F + (CFG::REF_SMEM ? CFG::TP * CFG::C : 0) +
                    CFG::TP * CFG::SD * (1 + (MODE != 2 ? 1 : 0) + (SIM ? 1 : 0)) + MAXN * 12) * 4 + 64;
}

// shared-memory counter increment with acquire-release semantics: the warp that sees 7 knows the other seven warps'
// tile reads and box minima happened before (their release) and may overwrite the tile (its acquire)
__device__ __forceinline__ unsigned atom_inc_acq_rel(unsigned* ctr) {
    unsigned old;
    asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(smem_u32(ctr)) : "memory");
    return old;
}

// MODE 0: pass A (entropy [+ cosine similarity]);  1: pass A that also stores the per-view group correlation;
// MODE 2: pass B (weighted aggregation, writes the volume).
// 256 threads = 8 warps, thread = pixel x hypothesis of the item.  No CTA-wide barrier in the item loop: a warp that has
// consumed item i (and delivered its share of item i+2's box) bumps cnt[i & 1]; the LAST of the eight fetches item i+2
// into the buffer item i used (TMA).  Warps only ever wait for full[slot], so they drift apart by up to an item and
// the projection / sampling / epilogue phases of different warps overlap.
template <class CFG, int MODE, bool SIM>
__global__ void __launch_bounds__(256, CFG::MINB)
cost_volume_cl_kernel(const __grid_constant__ CUtensorMap tmap, Params p) {
    constexpr int G = 8, C = CFG::C, D = CFG::D, CPG = CFG::CPG, TW = CFG::TW, TP = CFG::TP, HB = CFG::HB, NHB = CFG::NHB;
    constexpr int KPT = CFG::KPT, HPI = CFG::HPI;
    constexpr int NCH = CFG::NCH, NQ = CFG::NQ, BW = CFG::BW, BH = CFG::BH, SD = CFG::SD, GPI = CFG::GPI, C4 = C / 4;
    constexpr bool STORE = MODE == 1, PASS_B = MODE == 2, REF_SMEM = CFG::REF_SMEM;
    constexpr bool GROUPS = PASS_B || STORE;                              // per-group sums are needed
    static_assert(KPT == 1 || NCH == 1, "several hypotheses per thread only with whole texels");
    constexpr int NCP = CPG >= 4 ? (CPG == 8 ? 8 : 4) : CPG;             // c' values a thread tracks for SIM
    static_assert(!(PASS_B && SIM) && !(PASS_B && (NCH != 1 || NHB != 1)), "pass B is the stage-4 kernel");

    extern __shared__ __align__(128) uint8_t smem_raw[];
    float* tile0 = reinterpret_cast<float*>(smem_raw);                    // [2][BH][BW][CC]
    float* s_ref = tile0 + 2 * CFG::TILE_F;                               // [TP][C] (REF_SMEM)
    float* s_dep = s_ref + (REF_SMEM ? TP * C : 0);                       // [TP][SD] hypotheses
    float* s_col = s_dep + TP * SD;                                       // [TP][SD] per-view correlation sums (pass A)
    float* s_cos = s_col + (PASS_B ? 0 : TP * SD);                        // [TP][SD] cosine similarity over views (SIM)
    float* s_rel = s_cos + (SIM ? TP * SD : 0);                           // [MAXN][12] relative projections
    int* s_box = reinterpret_cast<int*>(s_rel + MAXN * 12);               // [2][2] running minima of the next boxes
    int* s_org = s_box + 4;                                               // [2][2] origins of the boxes in flight
    uint64_t* full = reinterpret_cast<uint64_t*>(s_org + 4);              // [2] tile landed
    unsigned* s_cnt = reinterpret_cast<unsigned*>(full + 2);              // [2] warps that consumed the tile in the slot

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.z;
    const int64_t hw = (int64_t)p.H * p.W;
    const float4* feat4 = reinterpret_cast<const float4*>(p.feat);
    const int NI = p.N * NHB * NCH;

    if (tid == 0) {
        mbar_init(&full[0], 1); mbar_init(&full[1], 1); mbar_fence_init();
        s_box[0] = s_box[1] = s_box[2] = s_box[3] = INT_MAX;
        s_cnt[0] = s_cnt[1] = 0;
    }
    auto issue = [&](int it) {                                            // one thread; box of item `it` complete, its slot free
        const int slot = it & 1, ch = it % NCH, v = it / (NCH * NHB);
        const int bx0 = s_box[slot * 2] == INT_MAX ? 0 : s_box[slot * 2];
        const int by0 = s_box[slot * 2 + 1] == INT_MAX ? 0 : s_box[slot * 2 + 1];
        s_org[slot * 2] = bx0; s_org[slot * 2 + 1] = by0;
        s_box[slot * 2] = INT_MAX; s_box[slot * 2 + 1] = INT_MAX;        // re-armed for item it+2: ordered before the other
        s_cnt[slot] = 0;                                                  // warps' next atomics by the release of this arrive
        mbar_expect_tx(&full[slot], CFG::TILE_F * 4);                    // and the acquire of their wait on full[slot]
        tma_load_4d(tile0 + slot * CFG::TILE_F, &tmap, &full[slot], ch * CFG::CC, bx0, by0, p.use_slots ? p.slot[v + 1] : b * p.V + v + 1);
    };

    const int pix = tid / HB, kk = tid - pix * HB;
    const int x = blockIdx.x * TW + (pix % TW), y = blockIdx.y * CFG::TH + (pix / TW);
    const bool live = x < p.W && y < p.H;
    const int pixoff = live ? y * p.W + x : 0;
    const float4* ref4 = feat4 + ((int64_t)(p.use_slots ? p.slot[0] : b * p.V) * hw + pixoff) * C4;
    // chunk order of this lane: q = j ^ xq (see the header).  NQ = 8: the 8 lanes of an LDS.128 phase get 8 different
    // bank groups whatever they sample; NQ = 4 / 2: lanes that share a chunk are hypotheses of one pixel (neighbouring
    // texels along the epipolar line) or, with every hypothesis of a pixel in one thread, four neighbouring pixels.
    const int xq = NQ == 8 ? (lane & 7)
                 : NQ == 4 ? (HB == 8 ? ((kk >> 1) & 3) : (kk & 3))
                           : (HB == 1 ? ((lane >> 2) & 1) : (pix & 1));
    Geo geo;
    geo.half_w = (float)((p.W - 1) / 2.0); geo.half_h = (float)((p.H - 1) / 2.0);
    geo.rcp_half_w = __frcp_rn(geo.half_w); geo.rcp_half_h = __frcp_rn(geo.half_h);
    geo.wm1 = (float)(p.W - 1); geo.hm1 = (float)(p.H - 1);
    const float inv_cpg = 1.0f / (float)CPG;

    float4 rreg[REF_SMEM ? 1 : C4];                                       // reference features in LOCAL chunk order
    if (!REF_SMEM) {
#pragma unroll
        for (int j = 0; j < (REF_SMEM ? 1 : C4); ++j) rreg[j] = live ? __ldg(ref4 + (j ^ xq)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // (the reference-feature loads above are in flight while the tile's tables are filled)
    {   // hypotheses of the tile, coalesced along x; every load is issued before the first store (the loop used to expose one
        // global-load latency per iteration: 7-15 % of the stall samples of the stage-3/4 kernels, profiles/r02m_k1_source_stalls.json)
        constexpr int NLD = (D * TP + 255) / 256;
        float hv[NLD];
#pragma unroll
        for (int u = 0; u < NLD; ++u) {
            const int i = tid + u * 256, k = i / TP, q = i - k * TP;
            const int qx = blockIdx.x * TW + (q % TW), qy = blockIdx.y * CFG::TH + (q / TW);
            hv[u] = (i < D * TP && qx < p.W && qy < p.H) ? __ldg(p.depth + ((int64_t)b * D + k) * hw + (int64_t)qy * p.W + qx) : 1.0f;
        }
#pragma unroll
        for (int u = 0; u < NLD; ++u) {
            const int i = tid + u * 256, k = i / TP, q = i - k * TP;
            if (i < D * TP) s_dep[q * SD + k] = hv[u];
        }
    }
    for (int i = tid; i < p.N * 12; i += 256) s_rel[i] = __ldg(p.relproj + (int64_t)b * p.N * 12 + i);
    if (SIM)
        for (int i = tid; i < TP * SD; i += 256) s_cos[i] = 0.0f;
    if (REF_SMEM) {
        float4* s_ref4 = reinterpret_cast<float4*>(s_ref);
        constexpr int NRF = (TP * C4 + 255) / 256;                        // loads first, stores after (as above)
        float4 rv[REF_SMEM ? NRF : 1];
#pragma unroll
        for (int u = 0; u < (REF_SMEM ? NRF : 1); ++u) {
            const int i = tid + u * 256, q = i / C4, cq = i - q * C4;
            const int qx = blockIdx.x * TW + (q % TW), qy = blockIdx.y * CFG::TH + (q / TW);
            rv[u] = (i < TP * C4 && qx < p.W && qy < p.H)
                        ? __ldg(feat4 + ((int64_t)(p.use_slots ? p.slot[0] : b * p.V) * hw + (int64_t)qy * p.W + qx) * C4 + cq)
                        : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < (REF_SMEM ? NRF : 1); ++u)
            if (tid + u * 256 < TP * C4) s_ref4[tid + u * 256] = rv[u];
    }
    __syncthreads();
    // 1 / max(||ref[:, c']||, 1e-12): F.normalize of the reference features (SIM), indexed by LOCAL c'
    float rinv[SIM ? NCP : 1];
    if (SIM) {
#pragma unroll
        for (int cp = 0; cp < NCP; ++cp) rinv[cp] = 0.0f;
        if (REF_SMEM) {
#pragma unroll
            for (int cp = 0; cp < (SIM ? NCP : 1); ++cp) {
                const int actual = (((cp >> 2) ^ (xq & 1)) << 2) | (cp & 3);
#pragma unroll
                for (int g = 0; g < G; ++g) { const float r = s_ref[pix * C + g * CPG + actual]; rinv[cp] = fmaf(r, r, rinv[cp]); }
            }
        } else {
#pragma unroll
            for (int j = 0; j < (REF_SMEM ? 0 : C4); ++j) {
                const float4 r = rreg[REF_SMEM ? 0 : j];
                const float e[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) rinv[SIM ? (i % NCP) : 0] = fmaf(e[i], e[i], rinv[SIM ? (i % NCP) : 0]);
            }
        }
#pragma unroll
        for (int cp = 0; cp < NCP; ++cp) rinv[cp] = 1.0f / fmaxf(sqrtf(rinv[cp]), 1e-12f);
    }

    // sample positions of my KPT hypotheses in item `it`; the pixel ray is kept while the view stays the same
    int ray_view = -1;
    PixelRay ray;
    RelProj mt;
    mt.t0 = mt.t1 = mt.t2 = 0.f;
    auto item_pos = [&](int it, float (&px)[KPT], float (&py)[KPT]) {
        const int hb = (it / NCH) % NHB, v = it / (NCH * NHB);
        if (v != ray_view) {                                              // uniform
            const RelProj m = load_relproj_smem(s_rel + v * 12);
            ray = pixel_ray(m, (float)x, (float)y);
            mt.t0 = m.t0; mt.t1 = m.t1; mt.t2 = m.t2;
            ray_view = v;
        }
#pragma unroll
        for (int j = 0; j < KPT; ++j) project(mt, ray, s_dep[pix * SD + hb * HPI + kk * KPT + j], geo, &px[j], &py[j]);
    };
    auto box_contrib = [&](int slot, const float (&px)[KPT], const float (&py)[KPT]) {
        int mx = INT_MAX, my = INT_MAX;
#pragma unroll
        for (int j = 0; j < KPT; ++j) {
            const bool sane = live && fabsf(px[j]) < 1e7f && fabsf(py[j]) < 1e7f;   // false for NaN / inf too
            if (sane) { mx = min(mx, (int)floorf(px[j])); my = min(my, (int)floorf(py[j])); }
        }
        mx = __reduce_min_sync(0xffffffffu, mx);
        my = __reduce_min_sync(0xffffffffu, my);
        if (lane == 0) { atomicMin(&s_box[slot * 2], mx); atomicMin(&s_box[slot * 2 + 1], my); }
    };

    float ix[KPT], iy[KPT], n1x[KPT], n1y[KPT], n2x[KPT], n2y[KPT];
#pragma unroll
    for (int j = 0; j < KPT; ++j) n1x[j] = n1y[j] = n2x[j] = n2y[j] = 0.f;
    item_pos(0, ix, iy);
    box_contrib(0, ix, iy);
    if (NI > 1) {
        if (NCH == 2) { n1x[0] = ix[0]; n1y[0] = iy[0]; } else item_pos(1, n1x, n1y);
        box_contrib(1, n1x, n1y);
    }
    __syncthreads();                                                      // boxes of items 0 and 1 are complete
    if (tid == 0) {
        issue(0);
        if (NI > 1) issue(1);
    }

    float acc[PASS_B ? KPT : 1][PASS_B ? G : 1];                          // pass B: weighted sums, LOCAL group order
#pragma unroll
    for (int j = 0; j < (PASS_B ? KPT : 1); ++j)
#pragma unroll
        for (int g = 0; g < (PASS_B ? G : 1); ++g) acc[j][g] = 0.0f;
    float wsum = 0.0f;
    float wv_next = (PASS_B && live) ? __ldg(p.vis_weight + (int64_t)b * p.N * hw + pixoff) : 0.0f;     // view 0
    float a_c[SIM ? NCP : 1], wn[SIM ? NCP : 1];                          // SIM: per-c' dot products / squared norms
    float sview = 0.0f;

    for (int it = 0; it < NI; ++it) {
        const int slot = it & 1, ch = it % NCH, hb = (it / NCH) % NHB, v = it / (NCH * NHB);
        const bool have2 = it + 2 < NI;
        if (have2) {
            if (NCH == 2 && (it & 1)) { n2x[0] = n1x[0]; n2y[0] = n1y[0]; }   // items 2m, 2m+1: same samples, other channel half
            else item_pos(it + 2, n2x, n2y);
        }
        float wv = 0.0f;
        if (PASS_B) {                                                     // this item's weight was requested one item ago
            wv = wv_next;
            if (it + 1 < NI) wv_next = live ? __ldg(p.vis_weight + ((int64_t)b * p.N + (it + 1) / (NCH * NHB)) * hw + pixoff) : 0.0f;
        }
        mbar_wait(&full[slot], (it >> 1) & 1);
        if (have2) box_contrib(slot, n2x, n2y);                           // this slot's box counters were re-armed by issue(it)
        const int bx0 = s_org[slot * 2], by0 = s_org[slot * 2 + 1];
        const float4* tile4 = reinterpret_cast<const float4*>(tile0 + slot * CFG::TILE_F);

#pragma unroll
        for (int jh = 0; jh < KPT; ++jh) {
            const int k = hb * HPI + kk * KPT + jh;
            const bool sane = live && fabsf(ix[jh]) < 1e7f && fabsf(iy[jh]) < 1e7f;
            const float x0 = floorf(ix[jh]), y0 = floorf(iy[jh]);
            const float fx = ix[jh] - x0, fy = iy[jh] - y0;
            const int lx = sane ? (int)x0 - bx0 : -1, ly = sane ? (int)y0 - by0 : -1;
            const bool inbox = lx >= 0 && lx + 1 < BW && ly >= 0 && ly + 1 < BH;

            float ag[GROUPS ? GPI : 1];                                   // this sample's per-group sums, LOCAL group order
#pragma unroll
            for (int g = 0; g < (GROUPS ? GPI : 1); ++g) ag[g] = 0.0f;
            if (ch == 0) {
                sview = 0.0f;
#pragma unroll
                for (int cp = 0; cp < (SIM ? NCP : 1); ++cp) { a_c[cp] = 0.0f; wn[cp] = 0.0f; }
            }

            // one chunk: four channels of the warped sample (w) against the reference (r); all indices static.  Where two
            // neighbouring channels feed two DIFFERENT accumulators (one group per channel; the per-c' similarity sums) the
            // pair is one packed FFMA2 — each half is the scalar fmaf it replaces, in the same order (common.cuh)
            auto accumulate = [&](int j, const float4& r4, const float4& w4) {
                const float r[4] = {r4.x, r4.y, r4.z, r4.w}, w[4] = {w4.x, w4.y, w4.z, w4.w};
                if (GROUPS && CPG == 1) {
#pragma unroll
                    for (int i = 0; i < 4; i += 2) {
                        const int g = GROUPS ? 4 * j + i : 0;
                        const float2 o = unpack2(ffma2(pack2(r[i], r[i + 1]), pack2(w[i], w[i + 1]), pack2(ag[g], ag[GROUPS ? g + 1 : 0])));
                        ag[g] = o.x; ag[GROUPS ? g + 1 : 0] = o.y;
                    }
                }
                if (SIM && NCP >= 2) {
#pragma unroll
                    for (int i = 0; i < 4; i += 2) {
                        const int c0 = SIM ? (4 * j + i) % NCP : 0, c1 = SIM ? (4 * j + i + 1) % NCP : 0;
                        const f32x2 wp = pack2(w[i], w[i + 1]);
                        const float2 a = unpack2(ffma2(pack2(r[i], r[i + 1]), wp, pack2(a_c[c0], a_c[c1])));
                        const float2 n = unpack2(ffma2(wp, wp, pack2(wn[c0], wn[c1])));
                        a_c[c0] = a.x; a_c[c1] = a.y; wn[c0] = n.x; wn[c1] = n.y;
                    }
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int lc = 4 * j + i;                             // local channel of the item
                    if (GROUPS && CPG > 1) ag[GROUPS ? lc / CPG : 0] = fmaf(r[i], w[i], ag[GROUPS ? lc / CPG : 0]);
                    if (SIM && NCP < 2) {
                        a_c[SIM ? lc % NCP : 0] = fmaf(r[i], w[i], a_c[SIM ? lc % NCP : 0]);
                        wn[SIM ? lc % NCP : 0] = fmaf(w[i], w[i], wn[SIM ? lc % NCP : 0]);
                    }
                    if (!GROUPS && !SIM) sview = fmaf(r[i], w[i], sview);
                }
            };
            // bilinear blend of four texel chunks, two channels per packed operation: ((t00*w00 + t01*w01) + t10*w10) + t11*w11
            // with the roundings of the nested scalar fmaf chain
            auto blend = [&](const float4& t00, const float4& t01, const float4& t10, const float4& t11, float w00, float w01, float w10,
                             float w11) {
                const f32x2 b00 = pack2(w00, w00), b01 = pack2(w01, w01), b10 = pack2(w10, w10), b11 = pack2(w11, w11);
                const float2 lo = unpack2(ffma2(pack2(t11.x, t11.y), b11, ffma2(pack2(t10.x, t10.y), b10,
                                          ffma2(pack2(t01.x, t01.y), b01, fmul2(pack2(t00.x, t00.y), b00)))));
                const float2 hi = unpack2(ffma2(pack2(t11.z, t11.w), b11, ffma2(pack2(t10.z, t10.w), b10,
                                          ffma2(pack2(t01.z, t01.w), b01, fmul2(pack2(t00.z, t00.w), b00)))));
                return make_float4(lo.x, lo.y, hi.x, hi.y);
            };

            if (live && inbox) {
                const float4* t = tile4 + (ly * BW + lx) * NQ;
                const float w00 = (1.0f - fx) * (1.0f - fy), w01 = fx * (1.0f - fy), w10 = (1.0f - fx) * fy, w11 = fx * fy;
                // 64-byte texels (stage 3): an LDS.128 phase is the 4 + 4 hypothesis lanes of two pixels; a lane's bank group is
                // 4 * (texel x parity) + chunk, and the chunk rotation only separates the lanes of ONE pixel — on noisy depth maps
                // the two pixels' texel parities are random and 38 % of the load wavefronts were conflict replays
                // (l1tex__data_bank_conflicts..._op_ld 19.4 M of 50.8 M, profiles/r02m_k1_source_stalls.json).  The two x taps of
                // a row are neighbouring texels = opposite parities, so each lane takes FIRST the tap whose parity equals its
                // pixel's position in the phase: every load instruction is then conflict-free whatever is sampled.
                constexpr bool XSWAP = NQ == 4 && HB == 4;
                static_assert(!XSWAP || (BW * NQ) % 8 == 0, "tile rows keep the bank phase");
                const bool swp = XSWAP && (((lx ^ (lane >> 2)) & 1) != 0);
                const int oa = swp ? NQ : 0, ob = swp ? 0 : NQ;            // first / second x tap, in 16-byte units
                const float wa0 = swp ? w01 : w00, wb0 = swp ? w00 : w01, wa1 = swp ? w11 : w10, wb1 = swp ? w10 : w11;
#pragma unroll
                for (int j = 0; j < NQ; ++j) {
                    const int q = j ^ xq;
                    const float4 t00 = t[oa + q], t01 = t[ob + q], t10 = t[BW * NQ + oa + q], t11 = t[BW * NQ + ob + q];
                    const float4 w = blend(t00, t01, t10, t11, wa0, wb0, wa1, wb1);
                    const float4 r = REF_SMEM ? reinterpret_cast<const float4*>(s_ref)[pix * C4 + ch * NQ + q] : rreg[REF_SMEM ? 0 : j];
                    accumulate(j, r, w);
                }
            } else if (live) {
                // predicated global path: the reference's per-tap bounds checks, any geometry
                const RelProj m = load_relproj_smem(s_rel + v * 12);
                const PixelRay sray = pixel_ray(m, (float)x, (float)y);
                const Taps tp = make_taps(m, sray, s_dep[pix * SD + k], p.H, p.W, geo.half_w, geo.half_h);
                const float4* src = feat4 + ((int64_t)(p.use_slots ? p.slot[v + 1] : b * p.V + v + 1) * hw) * C4 + ch * NQ;
#pragma unroll
                for (int j = 0; j < NQ; ++j) {
                    const int q = j ^ xq;
                    const float4 t00 = __ldg(src + (int64_t)tp.o00 * C4 + q), t01 = __ldg(src + (int64_t)tp.o01 * C4 + q);
                    const float4 t10 = __ldg(src + (int64_t)tp.o10 * C4 + q), t11 = __ldg(src + (int64_t)tp.o11 * C4 + q);
                    const float4 w = blend(t00, t01, t10, t11, tp.w00, tp.w01, tp.w10, tp.w11);
                    const float4 r = REF_SMEM ? reinterpret_cast<const float4*>(s_ref)[pix * C4 + ch * NQ + q] : rreg[REF_SMEM ? 0 : j];
                    accumulate(j, r, w);
                }
            }

            // ---- per-sample epilogue ------------------------------------------------------------------------------
            if (GROUPS) {
#pragma unroll
                for (int g = 0; g < (GROUPS ? GPI : 1); ++g) sview += ag[g];
            } else if (SIM && ch == NCH - 1) {
#pragma unroll
                for (int cp = 0; cp < (SIM ? NCP : 1); ++cp) sview += a_c[cp];
            }
            if (STORE && live) {
                float o[GPI];
#pragma unroll
                for (int g = 0; g < GPI; ++g) o[g] = ag[STORE ? g : 0] * inv_cpg;
                // local group -> actual group: XOR by the chunk mask expressed in groups
                if (CPG == 8) xor_permute<GPI, 2>(o, xq >> 1);
                else if (CPG == 4) xor_permute<GPI, 3>(o, xq);
                else if (CPG == 2) xor_permute<GPI, 3>(o, xq << 1);
                else xor_permute<GPI, 3>(o, xq << 2);
                float* out = p.corr + ((((int64_t)b * p.N + v) * D + k) * hw + pixoff) * G + ch * GPI;
#pragma unroll
                for (int g = 0; g < GPI; g += 4) *reinterpret_cast<float4*>(out + g) = make_float4(o[g], o[g + 1], o[g + 2], o[g + 3]);
            }
            if (PASS_B) {
                const f32x2 icp = pack2(inv_cpg, inv_cpg), wv2 = pack2(wv, wv);
#pragma unroll
                for (int g = 0; g < (PASS_B ? G : 2); g += 2) {           // acc = fmaf(ag * inv_cpg, wv, acc), two groups per FFMA2
                    const int j0 = PASS_B ? jh : 0, g0 = PASS_B ? g : 0, g1 = PASS_B ? g + 1 : 0;
                    const float2 o = unpack2(ffma2(fmul2(pack2(ag[g0], ag[g1]), icp), wv2, pack2(acc[j0][g0], acc[j0][g1])));
                    acc[j0][g0] = o.x; acc[j0][g1] = o.y;
                }
            } else if (ch == NCH - 1) {
                s_col[pix * SD + k] = sview * inv_cpg;                    // read back by this thread only
                if (SIM) {
                    // sum_c' a / (max(|ref|, eps) * max(|warped|, eps)): reciprocal square root instead of sqrt + division
                    float cosv = 0.0f;
#pragma unroll
                    for (int cp = 0; cp < (SIM ? NCP : 1); ++cp) cosv = fmaf(a_c[cp] * rinv[cp], rsqrtf(fmaxf(wn[cp], 1e-24f)), cosv);
                    s_cos[pix * SD + k] += cosv;
                }
            }
        }
        if (PASS_B) wsum += wv;
        __syncwarp();
        if (lane == 0 && have2 && atom_inc_acq_rel(&s_cnt[slot]) == 7) issue(it + 2);   // last warp out fetches item it+2

        // ---- end of a view: entropy of the softmax over the pixel's depth column; its HB threads are lanes of one warp
        if (!PASS_B && ch == NCH - 1 && hb == NHB - 1) {
            float mx = -FLT_MAX;
#pragma unroll
            for (int h = 0; h < NHB * KPT; ++h) mx = fmaxf(mx, s_col[pix * SD + (h / KPT) * HPI + kk * KPT + (h % KPT)]);
#pragma unroll
            for (int o = HB / 2; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            float e[NHB * KPT], den = 0.0f;
#pragma unroll
            for (int h = 0; h < NHB * KPT; ++h) { e[h] = expf(s_col[pix * SD + (h / KPT) * HPI + kk * KPT + (h % KPT)] - mx); den += e[h]; }
#pragma unroll
            for (int o = HB / 2; o > 0; o >>= 1) den += __shfl_xor_sync(0xffffffffu, den, o);
            float ent = 0.0f;
#pragma unroll
            for (int h = 0; h < NHB * KPT; ++h) { const float pr = e[h] / den; ent -= pr * logf(pr + 1e-7f); }
#pragma unroll
            for (int o = HB / 2; o > 0; o >>= 1) ent += __shfl_xor_sync(0xffffffffu, ent, o);
            if (kk == 0 && live) p.entropy[((int64_t)b * p.N + v) * hw + pixoff] = ent;
        }
#pragma unroll
        for (int j = 0; j < KPT; ++j) { ix[j] = n1x[j]; iy[j] = n1y[j]; n1x[j] = n2x[j]; n1y[j] = n2y[j]; }
    }

    if (PASS_B && live) {
        const float inv = 1.0f / (wsum + 1e-6f);
#pragma unroll
        for (int jh = 0; jh < (PASS_B ? KPT : 1); ++jh) {
            float o[G];
#pragma unroll
            for (int g = 0; g < G; ++g) {
                o[g] = acc[PASS_B ? jh : 0][PASS_B ? g : 0] * inv;
                if (p.round_tf32) o[g] = tc::to_tf32(o[g]);
            }
            if (CPG == 8) xor_permute<G, 2>(o, xq >> 1);
            else if (CPG == 4) xor_permute<G, 3>(o, xq);
            else if (CPG == 2) xor_permute<G, 3>(o, xq << 1);
            else xor_permute<G, 3>(o, xq << 2);
            float* out = p.volume + (((int64_t)b * D + kk * KPT + jh) * hw + pixoff) * G;     // NHB == 1 in pass B
            *reinterpret_cast<float4*>(out) = make_float4(o[0], o[1], o[2], o[3]);
            *reinterpret_cast<float4*>(out + 4) = make_float4(o[4], o[5], o[6], o[7]);
        }
    }
    if (!PASS_B && SIM) {
        // sim_depth = depth[argmax_k sum_v cos_v[k]] (models/mvsformer_model.py:151-156; first maximum, like torch.argmax).
        // The pixel's HB threads are lanes of one warp; every lane of the warp takes part in the shuffles.
        float best = -FLT_MAX;
        int bk = 0;
#pragma unroll
        for (int h = 0; h < NHB * KPT; ++h) {
            const int k = (h / KPT) * HPI + kk * KPT + (h % KPT);
            const float c = s_cos[pix * SD + k];
            if (c > best || (c == best && k < bk)) { best = c; bk = k; }
        }
#pragma unroll
        for (int o = HB / 2; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int ok = __shfl_xor_sync(0xffffffffu, bk, o);
            if (ob > best || (ob == best && ok < bk)) { best = ob; bk = ok; }
        }
        if (kk == 0 && live) p.sim_depth[(int64_t)b * hw + pixoff] = s_dep[pix * SD + bk];
    }
}

// ------------------------------------------------------------------------------------------------
// NCHW -> channels-last.  in [M][C][HW] -> out [M][HW][C].  A CTA transposes a tile of P = 4096 / C pixels x C channels
// (16 KB) through shared memory: 128-bit loads along the pixels, 128-bit stores along the channels, every sector
// written whole.  One launch covers every stage of a feature set (up to 4 segments).  HBM-bound: reads + writes the
// features once (2 x 566 MB per reference view at cfg 2).
// ------------------------------------------------------------------------------------------------
struct ClSegment {
    const float* in;
    float* out;
    int C;
    int64_t hw;        // pixels per map
    int64_t tiles;     // M * ceil(hw / P)
};
struct ClParams {
    ClSegment seg[4];
    int nseg;
};

template <int C>
__device__ __forceinline__ void nchw_to_cl_tile(const ClSegment& sg, int64_t bid, float* tile) {
    constexpr int P = 4096 / C, PITCH = P + 4, CQ = C / 4, LQ = CQ < 4 ? CQ : 4;
    const int64_t tpm = (sg.hw + P - 1) / P;
    const int64_t m = bid / tpm, p0 = (bid - m * tpm) * P;
    const float* in = sg.in + m * C * sg.hw;
    float* out = sg.out + m * C * sg.hw;
    const bool vec = (sg.hw & 3) == 0 && ((reinterpret_cast<uintptr_t>(sg.in) | reinterpret_cast<uintptr_t>(sg.out)) & 15) == 0;
    // read: float4 = 4 consecutive pixels of one channel
    for (int idx = threadIdx.x; idx < C * (P / 4); idx += 256) {
        const int c = idx / (P / 4), p4 = idx - c * (P / 4);
        const int64_t px = p0 + 4 * p4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (vec && px + 3 < sg.hw) {
            v = __ldg(reinterpret_cast<const float4*>(in + (int64_t)c * sg.hw + px));
        } else {
            if (px < sg.hw) v.x = __ldg(in + (int64_t)c * sg.hw + px);
            if (px + 1 < sg.hw) v.y = __ldg(in + (int64_t)c * sg.hw + px + 1);
            if (px + 2 < sg.hw) v.z = __ldg(in + (int64_t)c * sg.hw + px + 2);
            if (px + 3 < sg.hw) v.w = __ldg(in + (int64_t)c * sg.hw + px + 3);
        }
        *reinterpret_cast<float4*>(tile + c * PITCH + 4 * p4) = v;
    }
    __syncthreads();
    // write: float4 = 4 consecutive channels of one pixel; a warp covers 32 / LQ pixels x LQ channel quads
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cl = lane % LQ, pl = lane / LQ;
    constexpr int PPW = 32 / LQ;                                          // pixels per warp per step
    for (int pb = warp * PPW; pb < P; pb += 8 * PPW) {
        const int px = pb + pl;
        if (p0 + px >= sg.hw) continue;
#pragma unroll
        for (int cb = 0; cb < CQ; cb += LQ) {
            const int c = 4 * (cb + cl);
            const float4 v = make_float4(tile[c * PITCH + px], tile[(c + 1) * PITCH + px], tile[(c + 2) * PITCH + px],
                                         tile[(c + 3) * PITCH + px]);
            float* dst = out + (p0 + px) * C + c;
            if (vec) *reinterpret_cast<float4*>(dst) = v;
            else { dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w; }
        }
    }
}

__global__ void __launch_bounds__(256) nchw_to_cl_kernel(ClParams p) {
    extern __shared__ __align__(16) float cl_tile[];
    int64_t bid = blockIdx.x;
    int s = 0;
    while (s + 1 < p.nseg && bid >= p.seg[s].tiles) { bid -= p.seg[s].tiles; ++s; }
    const ClSegment sg = p.seg[s];
    switch (sg.C) {
        case 64: nchw_to_cl_tile<64>(sg, bid, cl_tile); break;
        case 32: nchw_to_cl_tile<32>(sg, bid, cl_tile); break;
        case 16: nchw_to_cl_tile<16>(sg, bid, cl_tile); break;
        default: nchw_to_cl_tile<8>(sg, bid, cl_tile); break;
    }
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

// 4-D map over the channels-last feature tensor [B*V][H][W][C] (innermost first: c, x, y, view); a box is
// [CC channels][BW][BH] of one view, out-of-image texels zero-filled (= grid_sample's zero padding).
static int make_feature_map(CUtensorMap* map, const float* feat, int BV, int C, int H, int W, int cc, int bw, int bh) {
    EncodeTiledFn fn = encode_fn();
    MVS_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
    const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)BV};
    const cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
    const cuuint32_t box[4] = {(cuuint32_t)cc, (cuuint32_t)bw, (cuuint32_t)bh, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult rc = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(feat), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MVS_REQUIRE(rc == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with code %d", (int)rc);
    return MVS_OK;
}

template <class CFG, int MODE, bool SIM>
static int launch(const Params& p, int B, int nmaps, cudaStream_t st) {
    constexpr size_t smem = smem_bytes<CFG, MODE, SIM>();
    CUtensorMap map;
    int rc = make_feature_map(&map, p.feat, nmaps, CFG::C, p.H, p.W, CFG::CC, CFG::BW, CFG::BH);
    if (rc) return rc;
    auto kern = cost_volume_cl_kernel<CFG, MODE, SIM>;
    MVS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(cdiv(p.W, CFG::TW), cdiv(p.H, CFG::TH), B);
    kern<<<grid, 256, smem, st>>>(map, p);
    MVS_LAUNCH_OK("cost_volume_cl_kernel");
    return MVS_OK;
}

// Tilings (pixel tile, hypotheses per item, box) chosen from measured coverage of the bench workload's noisy depth maps
// (scripts/box_coverage.py): >= 99.9 % of the samples of every view take the shared-memory path.  Hypotheses per thread
// (KPT) from an A/B on the B200 (profiles/r02b_k1_tiling_ab.json): KPT = 1 costs +6 % at stage 3 and +25 % at stage 4; stage 4
// with KPT = 2 at 80 registers / 3 CTAs per SM is 5 % slower than KPT = 4 at 2 CTAs (profiles/r02l_k1_stage4_tiling_ab.json).
//             C   D  TW TH HB KPT BW  BH NCH MINB
using Stage1 = Cfg<64, 32, 32, 1, 8, 1, 56, 5, 2, 2>;
using Stage2 = Cfg<32, 16, 32, 1, 8, 1, 56, 5, 1, 2>;
using Stage3 = Cfg<16, 8, 32, 2, 4, 2, 80, 6, 1, 3>;       // 2 hypotheses per thread
using Stage4 = Cfg<8, 4, 32, 8, 1, 4, 80, 14, 1, 2>;       // all 4 hypotheses of a pixel in one thread
}  // namespace k1cl

// Pass A over channels-last features; corr != nullptr also stores the per-view group correlation (C/G >= 2 only).
// Returns 1 when the shape is not covered (the caller falls back to the NCHW kernels), 0 on success, negative on error.
static void set_slots(k1cl::Params& p, const int* view_slots, int V) {
    p.use_slots = view_slots != nullptr;
    for (int i = 0; i < 17; ++i) p.slot[i] = (view_slots && i < V) ? view_slots[i] : 0;
}

int cost_volume_cl_entropy(const float* feat_cl, int nmaps, const int* view_slots, const float* relproj, const float* depth,
                           float* entropy, float* sim_depth, float* corr, int B, int V, int C, int G, int D, int H, int W,
                           cudaStream_t st) {
    using namespace k1cl;
    if (G != 8 || ((uintptr_t)feat_cl & 15) || V - 1 > MAXN) return 1;
    Params p{feat_cl, relproj, depth, V - 1, V, H, W, entropy, sim_depth, corr, nullptr, nullptr, 0};
    set_slots(p, view_slots, V);
    const bool sim = sim_depth != nullptr;
    if (C == 64 && D == 32 && corr) return sim ? launch<Stage1, 1, true>(p, B, nmaps, st) : launch<Stage1, 1, false>(p, B, nmaps, st);
    if (C == 32 && D == 16 && corr) return sim ? launch<Stage2, 1, true>(p, B, nmaps, st) : launch<Stage2, 1, false>(p, B, nmaps, st);
    if (C == 16 && D == 8 && corr) return sim ? launch<Stage3, 1, true>(p, B, nmaps, st) : launch<Stage3, 1, false>(p, B, nmaps, st);
    if (C == 8 && D == 4 && !corr) return sim ? launch<Stage4, 0, true>(p, B, nmaps, st) : launch<Stage4, 0, false>(p, B, nmaps, st);
    return 1;
}

int cost_volume_cl_aggregate(const float* feat_cl, int nmaps, const int* view_slots, const float* relproj, const float* depth,
                             const float* vis_weight, float* volume, int B, int V, int C, int G, int D, int H, int W,
                             int round_tf32, cudaStream_t st) {
    using namespace k1cl;
    if (G != 8 || ((uintptr_t)feat_cl & 15) || V - 1 > MAXN) return 1;
    Params p{feat_cl, relproj, depth, V - 1, V, H, W, nullptr, nullptr, nullptr, vis_weight, volume, round_tf32};
    set_slots(p, view_slots, V);
    if (C == 8 && D == 4) return launch<Stage4, 2, false>(p, B, nmaps, st);
    return 1;
}

int features_to_cl(const float* const* in, float* const* out, const int* C, const int64_t* hw, const int64_t* maps, int nseg,
                   cudaStream_t st) {
    k1cl::ClParams p;
    p.nseg = nseg;
    int64_t total = 0;
    for (int s = 0; s < nseg; ++s) {
        const int64_t P = 4096 / C[s];
        p.seg[s] = {in[s], out[s], C[s], hw[s], maps[s] * ((hw[s] + P - 1) / P)};
        total += p.seg[s].tiles;
    }
    const size_t smem = 64 * (64 + 4) * 4;                               // C x (4096 / C + 4) floats, largest at C = 64
    k1cl::nchw_to_cl_kernel<<<(unsigned)total, 256, smem, st>>>(p);
    MVS_LAUNCH_OK("nchw_to_cl_kernel");
    return MVS_OK;
}

}  // namespace mvs

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" int mvs_features_to_cl(const float* const* in, float* const* out, const int* channels, const int64_t* hw,
                                  const int64_t* maps, int nseg, void* stream) {
    MVS_REQUIRE(in && out && channels && hw && maps, "mvs_features_to_cl: null pointer");
    MVS_REQUIRE(nseg >= 1 && nseg <= 4, "mvs_features_to_cl: 1..4 segments per call (got %d)", nseg);
    for (int s = 0; s < nseg; ++s) {
        MVS_REQUIRE(in[s] && out[s], "mvs_features_to_cl: null segment pointer");
        MVS_REQUIRE((channels[s] == 8 || channels[s] == 16 || channels[s] == 32 || channels[s] == 64) && hw[s] >= 1 && maps[s] >= 1,
                    "mvs_features_to_cl: segment %d has C=%d (8, 16, 32 or 64), hw=%lld, maps=%lld", s, channels[s],
                    (long long)hw[s], (long long)maps[s]);
    }
    return mvs::features_to_cl(in, out, channels, hw, maps, nseg, (cudaStream_t)stream);
}

static int check_slots(const char* fn, int nmaps, const int* view_slots, int B, int V) {
    if (!view_slots) {
        MVS_REQUIRE(nmaps == B * V, "%s: without view_slots the feature tensor must hold B*V = %d maps (got %d)", fn, B * V, nmaps);
        return MVS_OK;
    }
    MVS_REQUIRE(B == 1, "%s: view_slots address the views of ONE batch item (got B = %d)", fn, B);
    MVS_REQUIRE(V <= 17, "%s: at most 17 views with view_slots (got %d)", fn, V);
    for (int i = 0; i < V; ++i)
        MVS_REQUIRE(view_slots[i] >= 0 && view_slots[i] < nmaps, "%s: view_slots[%d] = %d is outside the %d maps", fn, i, view_slots[i], nmaps);
    return MVS_OK;
}

extern "C" int mvs_cost_volume_cl_entropy(const float* feat_cl, int nmaps, const int* view_slots, const float* relproj,
                                          const float* depth, float* entropy, float* sim_depth, float* corr, int B, int V, int C,
                                          int G, int D, int H, int W, void* stream) {
    MVS_REQUIRE(feat_cl && relproj && depth && entropy, "mvs_cost_volume_cl_entropy: null pointer");
    if (int rc = check_slots("mvs_cost_volume_cl_entropy", nmaps, view_slots, B, V)) return rc;
    MVS_REQUIRE(B >= 1 && V >= 2 && C >= 1 && G >= 1 && D >= 1 && H >= 1 && W >= 1, "mvs_cost_volume_cl_entropy: empty shape");
    MVS_REQUIRE(C % G == 0, "mvs_cost_volume_cl_entropy: %d channels do not split into %d groups", C, G);
    return mvs::cost_volume_cl_entropy(feat_cl, nmaps, view_slots, relproj, depth, entropy, sim_depth, corr, B, V, C, G, D, H, W,
                                       (cudaStream_t)stream);
}

extern "C" int mvs_cost_volume_cl_aggregate(const float* feat_cl, int nmaps, const int* view_slots, const float* relproj,
                                            const float* depth, const float* vis_weight, float* volume, int B, int V, int C,
                                            int G, int D, int H, int W, int round_tf32, void* stream) {
    MVS_REQUIRE(feat_cl && relproj && depth && vis_weight && volume, "mvs_cost_volume_cl_aggregate: null pointer");
    if (int rc = check_slots("mvs_cost_volume_cl_aggregate", nmaps, view_slots, B, V)) return rc;
    MVS_REQUIRE(B >= 1 && V >= 2 && C >= 1 && G >= 1 && D >= 1 && H >= 1 && W >= 1, "mvs_cost_volume_cl_aggregate: empty shape");
    MVS_REQUIRE(C % G == 0, "mvs_cost_volume_cl_aggregate: %d channels do not split into %d groups", C, G);
    return mvs::cost_volume_cl_aggregate(feat_cl, nmaps, view_slots, relproj, depth, vis_weight, volume, B, V, C, G, D, H, W,
                                         round_tf32, (cudaStream_t)stream);
}
