// Tip-Adapter cache model (trainers/tip_adapter.py:43-80, 250-260, 281-290, 309-318, 331-333, 371-383):
//   affinity = f keys^T ;  cache_logits = exp(-(beta - beta*affinity)) @ one_hot(labels_tr) ;
//   tip_logits = clip_logits + alpha * cache_logits.
// The cache values are one-hot, so the second "GEMM" of the reference is a label-segmented sum: no [N_tr, C]
// matrix is ever formed.  These kernels are the exact-fp32 row passes over a materialised
// affinity block; the tensor-core path fuses the same reduction into the affinity GEMM's epilogue (gemm_tc.cu).
#include "common.cuh"

namespace clipgp {

// One CTA per image row b.  aff [B, N_tr] (row stride lda).  If store_e != 0 the affinity block is overwritten
// with e = exp(beta*(aff-1)) for the adjoint.  out[b,c] = clip_logits[b,c] + alpha * sum_{j: lab_j == c} e[b,j].
__global__ void __launch_bounds__(256) tip_forward_kernel(float* __restrict__ aff, int64_t lda, const int64_t* __restrict__ labels_tr,
                                                          int64_t N_tr, int C, float beta, float alpha,
                                                          const float* __restrict__ clip_logits, int64_t ldc,
                                                          float* __restrict__ out, int64_t ldo, int store_e) {
    extern __shared__ float acc[];   // [C]
    const int64_t b = blockIdx.x;
    for (int c = threadIdx.x; c < C; c += blockDim.x) acc[c] = 0.f;
    __syncthreads();
    float* row = aff + b * lda;
    for (int64_t j = threadIdx.x; j < N_tr; j += blockDim.x) {
        const float e = expf(-(beta - beta * row[j]));       // same association as the reference expression
        if (store_e) row[j] = e;
        atomicAdd(&acc[(int)labels_tr[j]], e);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x)
        out[b * ldo + c] = (clip_logits ? clip_logits[b * ldc + c] : 0.f) + acc[c] * alpha;
}

// G[b,j] = dout[b, lab_j] * alpha * beta * e[b,j]   (in place over e): d tip / d affinity.
__global__ void __launch_bounds__(256) tip_backward_kernel(float* __restrict__ e, int64_t lda, const int64_t* __restrict__ labels_tr,
                                                           int64_t N_tr, const float* __restrict__ dout, int64_t ldd, float beta,
                                                           float alpha) {
    const int64_t b = blockIdx.y;
    float* row = e + b * lda;
    const float* drow = dout + b * ldd;
    const float ab = alpha * beta;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < N_tr; j += (int64_t)gridDim.x * blockDim.x)
        row[j] = drow[(int)labels_tr[j]] * ab * row[j];
}

}  // namespace clipgp

using namespace clipgp;

extern "C" int clipgp_tip_forward(float* affinity, int64_t lda, const int64_t* labels_tr, int64_t B, int64_t N_tr, int64_t C,
                                  float beta, float alpha, const float* clip_logits, int64_t ldc, float* out, int64_t ldo,
                                  int store_e, void* stream) {
    CLIPGP_REQUIRE(B >= 0 && N_tr >= 0 && C >= 1 && C <= 12000, "tip_forward: bad shape (C must be <= 12000)");
    if (B == 0) return CLIPGP_OK;
    CLIPGP_REQUIRE(out && ldo >= C && (N_tr == 0 || (affinity && labels_tr && lda >= N_tr)), "tip_forward: bad pointers / strides");
    CLIPGP_REQUIRE(!clip_logits || ldc >= C, "tip_forward: ldc < C");
    CLIPGP_REQUIRE(B <= 2147483647ll, "tip_forward: B too large");
    tip_forward_kernel<<<(unsigned)B, 256, sizeof(float) * C, (cudaStream_t)stream>>>(affinity, lda, labels_tr, N_tr, (int)C, beta, alpha,
                                                                                   clip_logits, ldc, out, ldo, store_e);
    return check_launch("tip_forward_kernel");
}

extern "C" int clipgp_tip_backward(float* e, int64_t lda, const int64_t* labels_tr, int64_t B, int64_t N_tr, const float* dout,
                                   int64_t ldd, float beta, float alpha, void* stream) {
    CLIPGP_REQUIRE(B >= 0 && N_tr >= 0, "tip_backward: bad shape");
    if (B == 0 || N_tr == 0) return CLIPGP_OK;
    CLIPGP_REQUIRE(e && labels_tr && dout && lda >= N_tr, "tip_backward: bad pointers");
    CLIPGP_REQUIRE(B <= 65535, "tip_backward: B too large for one launch (chunk the rows)");
    int64_t bx = (N_tr + 255) / 256;
    if (bx > 64) bx = 64;
    dim3 grid((unsigned)bx, (unsigned)B);
    tip_backward_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(e, lda, labels_tr, N_tr, dout, ldd, beta, alpha);
    return check_launch("tip_backward_kernel");
}
