// fp32 (FFMA) GEMM for the exact-precision mode of the logit / projection / affinity contractions and
// their adjoints.  C[M,N] = alpha * op(A) op(B) (+ C), with A(m,k) = A[m*sam + k*sak], B(k,n) = B[k*sbk + n*sbn].
// 128x128x16 tiles, 256 threads, 8x8 register micro-tiles.  The tensor-core (tcgen05) path lives in gemm_tc.cu;
// this one is the bit-stable fp32 comparator the parity tests pin the tensor path against.
//
// Roofline: fp32 FFMA pipe (not the tensor pipe) -- used where 1e-3-relative parity against the fp32 oracle is
// required without operand rounding.
#include "common.cuh"

namespace clipgp {

constexpr int GBM = 128, GBN = 128, GBK = 16, GTHREADS = 256;

// KMAJOR_A: k is the contiguous index of A (sak == 1); otherwise m is (sam == 1).  Same for B with n.
template <bool A_KMAJOR, bool B_KMAJOR>
__global__ void __launch_bounds__(GTHREADS) gemm_f32_kernel(const float* __restrict__ A, int64_t sam, int64_t sak,
                                                            const float* __restrict__ B, int64_t sbk, int64_t sbn,
                                                            float* __restrict__ Cm, int64_t ldc, int M, int N, int K,
                                                            float alpha, int accumulate, int k_per_split) {
    __shared__ __align__(16) float As[GBK][GBM + 4];
    __shared__ __align__(16) float Bs[GBK][GBN + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * GBM, n0 = blockIdx.x * GBN;
    const int tx = tid % 16, ty = tid / 16;          // 16 x 16 threads, each 8 x 8 outputs
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    const int k_begin = blockIdx.z * k_per_split;
    const int k_end = min(K, k_begin + k_per_split);
    for (int k0 = k_begin; k0 < k_end; k0 += GBK) {
        // ---- load A tile [GBM x GBK] -> As[k][m]
#pragma unroll
        for (int it = 0; it < (GBM * GBK) / GTHREADS; ++it) {
            const int idx = tid + it * GTHREADS;
            int m, k;
            if (A_KMAJOR) { k = idx % GBK; m = idx / GBK; } else { m = idx % GBM; k = idx / GBM; }
            const int gm = m0 + m, gk = k0 + k;
            As[k][m] = (gm < M && gk < k_end) ? __ldg(A + (int64_t)gm * sam + (int64_t)gk * sak) : 0.f;
        }
#pragma unroll
        for (int it = 0; it < (GBN * GBK) / GTHREADS; ++it) {
            const int idx = tid + it * GTHREADS;
            int n, k;
            if (B_KMAJOR) { k = idx % GBK; n = idx / GBK; } else { n = idx % GBN; k = idx / GBN; }
            const int gn = n0 + n, gk = k0 + k;
            Bs[k][n] = (gn < N && gk < k_end) ? __ldg(B + (int64_t)gk * sbk + (int64_t)gn * sbn) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < GBK; ++k) {
            float a[8], b[8];
            const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[k][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][64 + tx * 4]);
            a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
            b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int gm = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int gn = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
            if (gn < N) {
                float* p = Cm + (int64_t)gm * ldc + gn;
                const float v = alpha * acc[i][j];
                if (gridDim.z > 1) atomicAdd(p, v);            // split-K: C was zeroed (or holds the accumulate base)
                else *p = accumulate ? (*p + v) : v;
            }
        }
    }
}

}  // namespace clipgp

using namespace clipgp;

extern "C" int clipgp_gemm_f32(const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbk, int64_t sbn,
                               float* Cm, int64_t ldc, int64_t M, int64_t N, int64_t K, float alpha, int accumulate,
                               void* stream) {
    CLIPGP_REQUIRE(M >= 0 && N >= 0 && K >= 0, "gemm_f32: negative size");
    CLIPGP_REQUIRE(M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), "gemm_f32: size too large");
    if (M == 0 || N == 0) return CLIPGP_OK;
    CLIPGP_REQUIRE(A && B && Cm, "gemm_f32: NULL pointer");
    CLIPGP_REQUIRE(ldc >= N, "gemm_f32: ldc < N");
    CLIPGP_REQUIRE(sak == 1 || sam == 1, "gemm_f32: A needs a unit stride (sam=%lld sak=%lld)", (long long)sam, (long long)sak);
    CLIPGP_REQUIRE(sbk == 1 || sbn == 1, "gemm_f32: B needs a unit stride (sbk=%lld sbn=%lld)", (long long)sbk, (long long)sbn);
    dim3 grid((unsigned)((N + GBN - 1) / GBN), (unsigned)((M + GBM - 1) / GBM));
    CLIPGP_REQUIRE(grid.y <= 65535, "gemm_f32: M too large for one launch (chunk the rows)");
    cudaStream_t st = (cudaStream_t)stream;
    // split-K when the output grid cannot fill the SMs (skinny adjoint GEMMs: K = S*C or B is the long axis)
    int splits = 1;
    const int64_t tiles = (int64_t)grid.x * grid.y;
    const bool deterministic = (accumulate & 2) != 0;       // bit 1: no split-K (atomic accumulation order varies from run to run)
    accumulate &= 1;
    if (!deterministic && tiles < num_sms() && K >= 8 * GBK) {
        splits = (int)((2 * num_sms() + tiles - 1) / tiles);
        const int max_splits = (int)(K / (4 * GBK));
        if (splits > max_splits) splits = max_splits;
        if (splits > 64) splits = 64;
        if (splits < 1) splits = 1;
    }
    int k_per_split = (int)K;
    if (splits > 1) {
        k_per_split = (int)(((K + splits - 1) / splits + GBK - 1) / GBK * GBK);
        splits = (int)((K + k_per_split - 1) / k_per_split);
        grid.z = (unsigned)splits;
        if (!accumulate) CLIPGP_CUDA(cudaMemset2DAsync(Cm, sizeof(float) * ldc, 0, sizeof(float) * N, M, st));
    }
    const bool ak = (sak == 1), bk = (sbk == 1);
    if (ak && bk) gemm_f32_kernel<true, true><<<grid, GTHREADS, 0, st>>>(A, sam, sak, B, sbk, sbn, Cm, ldc, (int)M, (int)N, (int)K, alpha, accumulate, k_per_split);
    else if (ak && !bk) gemm_f32_kernel<true, false><<<grid, GTHREADS, 0, st>>>(A, sam, sak, B, sbk, sbn, Cm, ldc, (int)M, (int)N, (int)K, alpha, accumulate, k_per_split);
    else if (!ak && bk) gemm_f32_kernel<false, true><<<grid, GTHREADS, 0, st>>>(A, sam, sak, B, sbk, sbn, Cm, ldc, (int)M, (int)N, (int)K, alpha, accumulate, k_per_split);
    else gemm_f32_kernel<false, false><<<grid, GTHREADS, 0, st>>>(A, sam, sak, B, sbk, sbn, Cm, ldc, (int)M, (int)N, (int)K, alpha, accumulate, k_per_split);
    return check_launch("gemm_f32_kernel");
}
