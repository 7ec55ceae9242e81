// One-time GP setup: the median-heuristic RBF length-scale of gp_template_weigher.py:103-107,
//     pdist = torch.cdist(flat_emb, flat_emb);  ls = pdist[pdist > 0].median()
// over ALL C*T unit-normalised reduced templates.  At the ImageNet shape that is a 32 000 x 32 000 fp32 matrix (4 GB) followed by a
// boolean mask and a sort of 10^9 values.  Here: an exact radix select of the k-th smallest pairwise distance over the 32-bit
// patterns of the (non-negative) distances, in three histogram passes (11 + 11 + 10 bits).  Every pass recomputes the distances
// tile by tile (64 x 64 pairs per CTA, fp32 FFMA, symmetric: only tiles on / above the diagonal, off-diagonal tiles counted twice) and
// never stores them: O(N d) memory traffic per tile row, O(1) extra memory.  Distances follow torch's matmul form of cdist,
// sqrt(max(|a|^2 + |b|^2 - 2 a.b, 0)); pairs (i, i) are excluded as the reference's `pdist > 0` intends.
// Roofline: fp32 FFMA pipe (N^2 d / 2 MACs per pass = 1.3e11 at N = 32 000, d = 256).
#include "common.cuh"

namespace clipgp {

constexpr int PT = 64;        // pairs tile
constexpr int PK = 16;        // feature chunk

// hist[b] += (number of ordered pairs (i != j) whose distance bits d satisfy (d >> prefix_shift) == prefix (when prefix_bits > 0)
//             and ((d >> shift) & (nbins - 1)) == b).  Also accumulates the number of ordered pairs with distance > 0 into positives[0]
//             (pass 0 only: count_pos != 0).
__global__ void __launch_bounds__(256) pairdist_hist_kernel(const float* __restrict__ X, const float* __restrict__ sq, int N, int d,
                                                            int prefix_shift, unsigned prefix, int use_prefix, int shift, int nbins,
                                                            unsigned long long* __restrict__ hist, unsigned long long* positives,
                                                            int count_pos) {
    extern __shared__ unsigned int s_hist[];                 // [nbins]
    __shared__ float sa[PK][PT + 1], sb[PK][PT + 1];
    __shared__ unsigned int s_pos;
    // map the linear block index onto the upper-triangular tile pairs (bi <= bj)
    const int nt = (N + PT - 1) / PT;
    int bi = 0, rem = blockIdx.x, len = nt;
    while (rem >= len) { rem -= len; --len; ++bi; }
    const int bj = bi + rem;
    const int weight = (bi == bj) ? 1 : 2;
    for (int i = threadIdx.x; i < nbins; i += blockDim.x) s_hist[i] = 0;
    if (threadIdx.x == 0) s_pos = 0;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 x 16 threads, 4 x 4 pairs each
    float acc[4][4];
#pragma unroll
    for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) acc[x][y] = 0.f;
    const int r0 = bi * PT, c0 = bj * PT;
    for (int k0 = 0; k0 < d; k0 += PK) {
        __syncthreads();
        for (int idx = threadIdx.x; idx < PT * PK; idx += blockDim.x) {
            const int r = idx / PK, k = idx - r * PK;
            sa[k][r] = (r0 + r < N && k0 + k < d) ? __ldg(X + (size_t)(r0 + r) * d + k0 + k) : 0.f;
            sb[k][r] = (c0 + r < N && k0 + k < d) ? __ldg(X + (size_t)(c0 + r) * d + k0 + k) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < PK; ++k) {
            float a[4], b[4];
#pragma unroll
            for (int x = 0; x < 4; ++x) { a[x] = sa[k][ty * 4 + x]; b[x] = sb[k][tx * 4 + x]; }
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y) acc[x][y] = fmaf(a[x], b[y], acc[x][y]);
        }
    }
    __syncthreads();
    unsigned int pos = 0;
#pragma unroll
    for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) {
            const int i = r0 + ty * 4 + x, j = c0 + tx * 4 + y;
            if (i < N && j < N && i != j) {
                const float d2 = fmaxf(sq[i] + sq[j] - 2.f * acc[x][y], 0.f);
                const float dist = sqrtf(d2);
                if (dist > 0.f) {
                    ++pos;
                    const unsigned bits = __float_as_uint(dist);
                    if (!use_prefix || (bits >> prefix_shift) == prefix) atomicAdd(&s_hist[(bits >> shift) & (unsigned)(nbins - 1)], 1u);
                }
            }
        }
    if (count_pos) atomicAdd(&s_pos, pos);
    __syncthreads();
    for (int i = threadIdx.x; i < nbins; i += blockDim.x)
        if (s_hist[i]) atomicAdd(&hist[i], (unsigned long long)s_hist[i] * weight);
    if (count_pos && threadIdx.x == 0 && s_pos) atomicAdd(positives, (unsigned long long)s_pos * weight);
}

__global__ void __launch_bounds__(256) row_sqnorm_kernel(const float* __restrict__ X, int N, int d, float* __restrict__ sq) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= N) return;
    float q = 0.f;
    for (int k = lane; k < d; k += 32) { const float v = X[(size_t)row * d + k]; q = fmaf(v, v, q); }
    q = warp_sum(q);
    if (lane == 0) sq[row] = q;
}

}  // namespace clipgp

using namespace clipgp;

extern "C" int clipgp_row_sqnorm(const float* X, int64_t N, int64_t d, float* sq, void* stream) {
    CLIPGP_REQUIRE(N >= 0 && d >= 1 && N < (1ll << 31) && d < (1ll << 31), "row_sqnorm: bad shape");
    if (N == 0) return CLIPGP_OK;
    CLIPGP_REQUIRE(X && sq, "row_sqnorm: NULL pointer");
    row_sqnorm_kernel<<<(unsigned)((N + 7) / 8), 256, 0, (cudaStream_t)stream>>>(X, (int)N, (int)d, sq);
    return check_launch("row_sqnorm_kernel");
}

extern "C" int clipgp_pairdist_radix_hist(const float* X, const float* sqnorm, int64_t N, int64_t d, int prefix_shift,
                                          uint32_t prefix, int use_prefix, int shift, int nbins, unsigned long long* hist,
                                          unsigned long long* positives, void* stream) {
    CLIPGP_REQUIRE(N >= 1 && d >= 1 && N <= 2000000 && d < (1ll << 31), "pairdist_radix_hist: bad shape");
    CLIPGP_REQUIRE(nbins >= 2 && nbins <= 4096 && (nbins & (nbins - 1)) == 0, "pairdist_radix_hist: nbins must be a power of two <= 4096");
    CLIPGP_REQUIRE(shift >= 0 && shift < 32 && prefix_shift >= 0 && prefix_shift <= 32, "pairdist_radix_hist: bad shifts");
    CLIPGP_REQUIRE(X && sqnorm && hist, "pairdist_radix_hist: NULL pointer");
    const int64_t nt = (N + PT - 1) / PT;
    const int64_t blocks = nt * (nt + 1) / 2;
    CLIPGP_REQUIRE(blocks < (1ll << 31), "pairdist_radix_hist: too many tiles");
    pairdist_hist_kernel<<<(unsigned)blocks, 256, sizeof(unsigned int) * nbins, (cudaStream_t)stream>>>(
        X, sqnorm, (int)N, (int)d, prefix_shift >= 32 ? 31 : prefix_shift, prefix, use_prefix, shift, nbins, hist, positives,
        positives != nullptr ? 1 : 0);
    return check_launch("pairdist_hist_kernel");
}
