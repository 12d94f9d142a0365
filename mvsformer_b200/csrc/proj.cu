// proj.cu — A1: camera composition and relative projection, evaluated on the device in fp64.
// Replaces models/mvsformer_model.py:69-72 (K @ E[:3,:4]) and models/warping.py:80-82
// (src_proj @ inverse(ref_proj)): in the reference these are a cuSOLVER batched LU, two cuBLAS
// calls and several copies per source view per stage; here one tiny kernel on the caller's
// stream, no host round trip.
#include "common.cuh"

namespace mvs {

__device__ void compose_4x4(const float* __restrict__ pair, double P[4][4]) {
    // pair = [2][4][4]: extrinsic E, intrinsic K (upper-left 3x3 used)
    const float* E = pair;
    const float* K = pair + 16;
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 4; ++c) {
            double s = 0.0;
            for (int k = 0; k < 3; ++k) s += (double)K[r * 4 + k] * (double)E[k * 4 + c];
            P[r][c] = s;
        }
    for (int c = 0; c < 4; ++c) P[3][c] = (double)E[12 + c];
}

__device__ void load_4x4(const float* __restrict__ m, double P[4][4]) {
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) P[r][c] = (double)m[r * 4 + c];
}

// Gauss-Jordan with partial pivoting; returns false for a singular matrix.
__device__ bool invert_4x4(double A[4][4], double Inv[4][4]) {
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) Inv[r][c] = (r == c) ? 1.0 : 0.0;
    for (int col = 0; col < 4; ++col) {
        int piv = col;
        double best = fabs(A[col][col]);
        for (int r = col + 1; r < 4; ++r)
            if (fabs(A[r][col]) > best) { best = fabs(A[r][col]); piv = r; }
        if (best == 0.0) return false;
        if (piv != col)
            for (int c = 0; c < 4; ++c) {
                double t = A[col][c]; A[col][c] = A[piv][c]; A[piv][c] = t;
                t = Inv[col][c]; Inv[col][c] = Inv[piv][c]; Inv[piv][c] = t;
            }
        const double inv_p = 1.0 / A[col][col];
        for (int c = 0; c < 4; ++c) { A[col][c] *= inv_p; Inv[col][c] *= inv_p; }
        for (int r = 0; r < 4; ++r) {
            if (r == col) continue;
            const double f = A[r][col];
            if (f == 0.0) continue;
            for (int c = 0; c < 4; ++c) { A[r][c] -= f * A[col][c]; Inv[r][c] -= f * Inv[col][c]; }
        }
    }
    return true;
}

__device__ void write_rel(const double S[4][4], double R[4][4], float* __restrict__ out) {
    double Rinv[4][4];
    const bool ok = invert_4x4(R, Rinv);
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 4; ++c) {
            double s = 0.0;
            for (int k = 0; k < 4; ++k) s += S[r][k] * Rinv[k][c];
            // a singular reference projection yields NaNs (torch.inverse raises; we cannot raise
            // from the device without a sync, NaNs propagate to every output instead)
            out[r * 4 + c] = ok ? (float)s : __int_as_float(0x7fc00000);
        }
}

__global__ void relproj_from_pairs_kernel(const float* __restrict__ proj, int B, int V, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int N = V - 1;
    if (i >= B * N) return;
    const int b = i / N, v = i % N + 1;
    double R[4][4], S[4][4];
    compose_4x4(proj + ((int64_t)b * V + 0) * 32, R);
    compose_4x4(proj + ((int64_t)b * V + v) * 32, S);
    write_rel(S, R, out + (int64_t)i * 12);
}

__global__ void relproj_from_composed_kernel(const float* __restrict__ src, const float* __restrict__ ref, int B,
                                             float* __restrict__ out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double R[4][4], S[4][4];
    load_4x4(ref + (int64_t)b * 16, R);
    load_4x4(src + (int64_t)b * 16, S);
    write_rel(S, R, out + (int64_t)b * 12);
}

}  // namespace mvs

extern "C" int mvs_relative_projections(const float* proj_matrices, int B, int V, float* relproj, void* stream) {
    MVS_REQUIRE(proj_matrices && relproj, "mvs_relative_projections: null pointer");
    MVS_REQUIRE(B >= 1 && V >= 2, "mvs_relative_projections: need B >= 1 and V >= 2 (got B=%d V=%d)", B, V);
    MVS_REQUIRE(V - 1 <= MVS_MAX_SRC_VIEWS, "mvs_relative_projections: at most %d source views (got %d)",
                MVS_MAX_SRC_VIEWS, V - 1);
    const int n = B * (V - 1);
    mvs::relproj_from_pairs_kernel<<<mvs::cdiv(n, 64), 64, 0, (cudaStream_t)stream>>>(proj_matrices, B, V, relproj);
    MVS_LAUNCH_OK("relproj_from_pairs_kernel");
    return MVS_OK;
}

extern "C" int mvs_relative_projection_pair(const float* src_proj, const float* ref_proj, int B, float* relproj,
                                            void* stream) {
    MVS_REQUIRE(src_proj && ref_proj && relproj, "mvs_relative_projection_pair: null pointer");
    MVS_REQUIRE(B >= 1, "mvs_relative_projection_pair: B must be >= 1 (got %d)", B);
    mvs::relproj_from_composed_kernel<<<mvs::cdiv(B, 64), 64, 0, (cudaStream_t)stream>>>(src_proj, ref_proj, B, relproj);
    MVS_LAUNCH_OK("relproj_from_composed_kernel");
    return MVS_OK;
}
