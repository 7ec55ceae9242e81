// Warp-per-class, register-resident building blocks of the GP template weighter for T <= 32 templates
// (n = T + 1 inducing points).  Lane i owns row i (or column i) of every T x T matrix in registers; the one
// extra inducing row (the learnable class token) is carried as a border:
//
//     K_ZZ + jI = [ Kt + jI   k ]      L = [ Lt   0 ]      l = Lt^-1 k,   lam = sqrt(kappa + j - l.l)
//                 [ k^T  kappa+j ]         [ l^T lam ]
//
// All loops over register arrays are fully unrolled (static register indices); cross-lane traffic is warp shuffles.
// No shared-memory matrices, no block barriers: a CTA is just a bundle of independent warps, so many classes are in
// flight per SM and their dependent chains (Cholesky columns, substitutions) overlap.
#pragma once
#include "common.cuh"

namespace clipgp {
namespace gpw {

constexpr int TM = 32;            // lanes = max templates
constexpr unsigned FULL = 0xffffffffu;

template <typename T>
__device__ __forceinline__ T bcast(T v, int src) { return __shfl_sync(FULL, v, src); }

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

// Row-per-lane lower Cholesky, in place: lane i holds a[0..31] = row i of a symmetric PD matrix (only columns <= i are
// used / produced).  Right-looking, fully unrolled.  inv_diag = 1 / L[lane][lane].  Returns (uniformly) true on failure.
template <typename T>
__device__ __forceinline__ bool chol_rows(T (&a)[TM], T& inv_diag) {
    const int lane = threadIdx.x & 31;
    bool fail = false;
    inv_diag = (T)1;
#pragma unroll
    for (int j = 0; j < TM; ++j) {
        const T ajj = bcast(a[j], j);
        if (!(ajj > (T)0)) fail = true;
        const T inv = rsqrt(ajj);
        const T lij = a[j] * inv;                       // column j entry of this lane's row (valid for lane >= j)
        if (lane == j) inv_diag = inv;
        a[j] = (lane == j) ? ajj * inv : lij;
#pragma unroll
        for (int k = j + 1; k < TM; ++k) {
            const T lkj = bcast(lij, k);                // L[k][j]
            a[k] -= lij * lkj;                          // meaningful for lane >= k; upper part is never read
        }
    }
    return fail;
}

// x = L^-1 b for a vector with one element per lane (b_i in lane i); L row-per-lane, inv_diag per lane.
template <typename T>
__device__ __forceinline__ T fwd_subst_vec(const T (&l)[TM], T inv_diag, T b) {
    const int lane = threadIdx.x & 31;
    T acc = b, x = (T)0;
#pragma unroll
    for (int m = 0; m < TM; ++m) {
        if (lane == m) x = acc * inv_diag;              // finalise element m
        const T xm = bcast(x, m);
        if (lane > m) acc -= l[m] * xm;
    }
    return x;
}

// x = L^-T b (back substitution), vector with one element per lane.  Needs column access L[k][m] for k > m, i.e.
// element m of lane k's row: broadcast per (m): each lane accumulates sum_{k>m} L[k][m] x_k by reduction over lanes.
template <typename T>
__device__ __forceinline__ T bwd_subst_vec(const T (&l)[TM], T inv_diag, T b) {
    const int lane = threadIdx.x & 31;
    T x = (T)0;
    // x_m = (b_m - sum_{k>m} L[k][m] x_k) / L[m][m], m = 31 .. 0.  Lane k contributes l[m] * x_k once x_k is known.
    T acc = b;                                          // lane m accumulates its own right-hand side
#pragma unroll
    for (int k = TM - 1; k >= 0; --k) {
        if (lane == k) x = acc * inv_diag;
        const T xk = bcast(x, k);
        // lane m (< k) needs L[k][m] = element m of lane k's row: lane-dependent register index -> use a transposed copy
        // supplied by the caller instead (see bwd_subst_cols); this vector form is only used with lt = transposed rows.
        if (lane < k) acc -= l[k] * xk;                 // here l is the TRANSPOSED factor: l[k] = L[k][lane]
    }
    return x;
}

// 32 x 32 in-register transpose: lane i holds row i in a[0..31]; afterwards lane i holds column i.  5 butterfly stages.
template <typename T>
__device__ __forceinline__ void transpose_rows(T (&a)[TM]) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        const bool up = (lane & s) != 0;
#pragma unroll
        for (int j = 0; j < TM; ++j) {
            if ((j & s) == 0) {
                // exchange a[j + s] of the lower lane with a[j] of the upper lane
                const T send = up ? a[j] : a[j + s];
                const T recv = __shfl_xor_sync(FULL, send, s);
                if (up) a[j] = recv; else a[j + s] = recv;
            }
        }
    }
}

}  // namespace gpw
}  // namespace clipgp
