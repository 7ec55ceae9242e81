// Warp-per-class building blocks of the GP template weighter for T <= 32 templates (n = T + 1 inducing points).
//
// One warp owns one class; every matrix of the class (<= 33 x 33) lives in that warp's slice of shared memory and all
// loops are run-time loops over it (compact code: a fully unrolled register-resident variant of these kernels was
// instruction-fetch bound, ncu stall_no_inst ~ 40 %).  Conventions that keep shared-memory traffic conflict free:
//   * row stride LD = 33 (odd): "lane = row" walks A[lane*LD + k] hit 32 distinct banks,
//   * "lane = column" accesses A[k*LD + lane] are contiguous, the other operand of a product is a broadcast read.
// A CTA is ONE class with NW = 4 warps (31 KB of shared memory, 7 classes resident per SM, so the 1000 classes of the ImageNet
// shape are one wave on 148 SMs): the matrix products and staging copies are spread over the four warps, the inherently
// sequential factorisations and triangular sweeps run in warp 0 (+ warp 1 for the 33rd row / column).
#pragma once
#include "gp_common.cuh"
#include "gp_block.cuh"

namespace clipgp {
namespace gpw {

constexpr int LD = 33;                 // row stride (elements) of every per-class matrix
constexpr int NN = 33 * LD + 3;        // elements per matrix slot (3 pad elements: 4-wide tiles may read past the last row)
constexpr unsigned FULL = 0xffffffffu;

constexpr int NW = 4;                  // warps per class CTA
constexpr int NT = NW * 32;

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int warp_id() { return threadIdx.x >> 5; }

// idx / cols for 0 <= idx < 2^20 without an integer division: (idx + 0.5) / cols is at least 0.5 / cols away from any integer,
// far more than the fp32 rounding of the product.
__device__ __forceinline__ int fast_div(int idx, float inv_cols) { return __float2int_rz(((float)idx + 0.5f) * inv_cols); }

// Block-strided walk over a dense row-major [rows, cols] array: f(idx, i, j).
template <typename F>
__device__ __forceinline__ void each_block(int rows, int cols, F f) {
    const int total = rows * cols;
    const float inv = 1.f / (float)cols;
    for (int idx = threadIdx.x; idx < total; idx += NT) { const int i = fast_div(idx, inv); f(idx, i, idx - i * cols); }
}

// Staging copy with four global loads in flight per thread (block-strided).
template <typename V, typename LD_, typename ST_>
__device__ __forceinline__ void stage_block(int rows, int cols, LD_ ld, ST_ st) {
    const int total = rows * cols;
    const float inv = 1.f / (float)cols;
    for (int base = threadIdx.x; base < total; base += 4 * NT) {
        V v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { const int idx = base + NT * u; if (idx < total) v[u] = ld(idx); }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int idx = base + NT * u;
            if (idx < total) { const int i = fast_div(idx, inv); st(idx, i, idx - i * cols, v[u]); }
        }
    }
}

// init - sum_{k0 <= k < k1} a[k * SA] * b[k * SB].  The sequential sweeps below (Cholesky columns, triangular substitutions) are
// latency bound: one warp, a dependent chain per column.  Eight operand pairs are loaded before any is consumed and the products
// go to four accumulators, so a step costs about one shared-memory round trip per EIGHT terms instead of per two.
template <typename T, int SA, int SB>
__device__ __forceinline__ T dot_sub(T init, const T* a, const T* b, int k0, int k1) {
    T s0 = init, s1 = (T)0, s2 = (T)0, s3 = (T)0;
    int k = k0;
    for (; k + 8 <= k1; k += 8) {
        T x[8], y[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { x[u] = a[(k + u) * SA]; y[u] = b[(k + u) * SB]; }
        s0 -= x[0] * y[0]; s1 -= x[1] * y[1]; s2 -= x[2] * y[2]; s3 -= x[3] * y[3];
        s0 -= x[4] * y[4]; s1 -= x[5] * y[5]; s2 -= x[6] * y[6]; s3 -= x[7] * y[7];
    }
    if (k + 4 <= k1) {
        T x[4], y[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { x[u] = a[(k + u) * SA]; y[u] = b[(k + u) * SB]; }
        s0 -= x[0] * y[0]; s1 -= x[1] * y[1]; s2 -= x[2] * y[2]; s3 -= x[3] * y[3];
        k += 4;
    }
    if (k + 2 <= k1) {
        const T x0 = a[k * SA], y0 = b[k * SB], x1 = a[(k + 1) * SA], y1 = b[(k + 1) * SB];
        s0 -= x0 * y0; s1 -= x1 * y1;
        k += 2;
    }
    if (k < k1) s2 -= a[k * SA] * b[k * SB];
    return (s0 + s1) + (s2 + s3);
}

// Left-looking lower Cholesky of an n x n matrix (n <= 33) by one warp, in place on the lower triangle.  Lane i owns row i;
// when n == 33 the extra row 32 is carried by lane j-1 (idle in column j >= 1, because its own row is above the diagonal), so
// the warp makes ONE pass per column instead of two.  invd[j] = 1 / L[j][j].  Returns true (uniformly) on a non-positive pivot.
template <typename T>
__device__ __forceinline__ bool chol33(T* __restrict__ A, int n, T* __restrict__ invd) {
    const int lane = lane_id();
    bool fail = false;
    for (int j = 0; j < n; ++j) {
        int i = lane;
        if (n == 33 && lane == j - 1) i = 32;
        const bool active = (i >= j) && (i < n);
        const T* ri = A + (active ? i : j) * LD;
        const T acc = dot_sub<T, 1, 1>(ri[j], ri, A + j * LD, 0, j);
        // the pivot is the accumulator of row j: lane j for j < 32, lane 31 (carrying row 32) for j == 32
        const T sj = __shfl_sync(FULL, acc, j < 32 ? j : 31);
        if (!(sj > (T)0)) fail = true;
        const T inv = rsqrt(sj);
        __syncwarp();
        if (active && i > j) A[i * LD + j] = acc * inv;
        if (n == 33 && j == 0 && lane == 0) A[32 * LD] = A[32 * LD] * inv;      // column 0 has no idle lane: row 32 by lane 0
        if (lane == 0) { A[j * LD + j] = sj * inv; invd[j] = inv; }
        __syncwarp();
    }
    return fail;
}

// Right-looking Cholesky of an m x m fp32 matrix (m <= 32) by one warp with the matrix in REGISTERS: lane i holds row i, the
// column loop is fully unrolled, L[k][j] reaches the other rows by shuffle.  One column costs a pivot broadcast, one MUFU and
// 2 (31 - j) independent SHFL / FFMA (about 1.3 k instructions in total) instead of ~150 dependent instructions around a
// shared-memory dot product: measured 24 k -> 2 k cycles for a 32 x 32 factor (tools/gp_phase_ts.py).  Rows >= m are padded with
// the identity.  In place on the lower triangle of A (row stride LD); invd[j] = 1 / L[j][j]; returns true (uniformly) on a
// non-positive pivot.
__device__ __forceinline__ bool chol32_regs(float* __restrict__ A, int m, float* __restrict__ invd) {
    const int lane = lane_id();
    float a[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) a[k] = (lane < m && k <= lane) ? A[lane * LD + k] : (k == lane ? 1.f : 0.f);
    bool fail = false;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const float d = __shfl_sync(FULL, a[j], j);
        if (!(d > 0.f)) fail = true;
        const float inv = rsqrtf(d);
        const float lij = (lane == j) ? d * inv : a[j] * inv;          // L[i][j], meaningful for i >= j
        a[j] = lij;
        if (lane == j) invd[j] = inv;
#pragma unroll
        for (int k = j + 1; k < 32; ++k) a[k] = fmaf(-lij, __shfl_sync(FULL, lij, k), a[k]);   // A[i][k] -= L[i][j] L[k][j] (used for i >= k)
    }
    if (lane < m) {
#pragma unroll
        for (int k = 0; k < 32; ++k)
            if (k <= lane) A[lane * LD + k] = a[k];
    }
    __syncwarp();
    return fail;
}

// X <- L^-1 X (forward substitution), X is [n][ncol] with lane = column; two partial sums shorten the dependent chain.
template <typename T>
__device__ __forceinline__ void trsm_lower_cols(const T* __restrict__ L, const T* __restrict__ invd, T* __restrict__ X, int n, int ncol) {
    const int lane = lane_id();
    if (lane < ncol) {
        for (int i = 0; i < n; ++i) {
            X[i * LD + lane] = dot_sub<T, 1, LD>(X[i * LD + lane], L + i * LD, X + lane, 0, i) * invd[i];
        }
    }
    __syncwarp();
}

// X <- L^-1 X for n == 33, fully unrolled over rows and terms: every L[i][k] is a broadcast load with an immediate offset, every
// X[k][lane] stays in a register (the column of X a lane owns), no address arithmetic and no loop control on the dependent chain.
template <typename T>
__device__ __forceinline__ void trsm_lower_cols33(const T* __restrict__ L, const T* __restrict__ invd, T* __restrict__ X, int ncol) {
    const int lane = lane_id();
    if (lane < ncol) {
        T x[33];
#pragma unroll
        for (int i = 0; i < 33; ++i) x[i] = X[i * LD + lane];
#pragma unroll
        for (int i = 0; i < 33; ++i) {
            T s0 = x[i], s1 = (T)0;
#pragma unroll
            for (int k = 0; k < i; ++k) {
                if (k & 1) s1 -= L[i * LD + k] * x[k];
                else s0 -= L[i * LD + k] * x[k];
            }
            x[i] = (s0 + s1) * invd[i];
            X[i * LD + lane] = x[i];
        }
    }
    __syncwarp();
}

// Back substitution x_i = (x_i - sum_{k > i} L[k][i] x_k) / L[i][i], i descending, for a vector of N = 33 (or 32) entries, fully unrolled.
// `x` is addressed as base[k * SX]: a column of a row-major matrix (SX = LD, lane = column: L^-T X) or a row (SX = 1, lane = row:
// X L^-1).  Measured per 33 x 33 fp64 sweep: ~18 k -> ~13 k cycles (a register-resident x spills at the adjoint kernel's 72 registers).
template <typename T, int SX, int N = 33>
__device__ __forceinline__ void backsub33(const T* __restrict__ L, const T* __restrict__ invd, T* base) {
    // shared-memory resident, statically unrolled: immediate offsets, no loop control; only the newest x_k is on the dependent chain
#pragma unroll
    for (int i = N - 1; i >= 0; --i) {
        T s0 = base[i * SX], s1 = (T)0;
#pragma unroll
        for (int k = N - 1; k > i; --k) {
            const T xk = *reinterpret_cast<volatile T*>(base + k * SX);
            if ((k - i) & 1) s0 -= L[k * LD + i] * xk;
            else s1 -= L[k * LD + i] * xk;
        }
        *reinterpret_cast<volatile T*>(base + i * SX) = (s0 + s1) * invd[i];
    }
}

// X <- L^-T X (back substitution), lane = column.
template <typename T>
__device__ __forceinline__ void trsm_lowerT_cols(const T* __restrict__ L, const T* __restrict__ invd, T* __restrict__ X, int n, int ncol) {
    const int lane = lane_id();
    if (lane < ncol) {
        if (n == 33) backsub33<T, LD>(L, invd, X + lane);
        else
            for (int i = n - 1; i >= 0; --i) {
                X[i * LD + lane] = dot_sub<T, LD, LD>(X[i * LD + lane], L + i, X + lane, i + 1, n) * invd[i];
            }
    }
    __syncwarp();
}

// Sort-free sparsemax over the T <= 32 lanes (entmax SparsemaxFunction.forward: same support and threshold as the sorted rule).
__device__ __forceinline__ float sparsemax_lanes(float f, int T) {
    const int lane = lane_id();
    const bool valid = lane < T;
    const float fm = warp_max(valid ? f : -INFINITY);
    const float z = f - fm;
    bool sup = valid;
    int cnt = T;
    float tau;
    for (;;) {      // Michelot's fixed point: tau <- (sum_S z - 1) / |S|, S <- {z > tau}; ends at the support of the sorted rule
        tau = (warp_sum(sup ? z : 0.f) - 1.f) / (float)cnt;
        const bool nsup = sup && z > tau;
        const int c = __popc(__ballot_sync(FULL, nsup));
        if (c == cnt) break;
        sup = nsup; cnt = c;
    }
    return valid ? fmaxf(z - tau, 0.f) : 0.f;
}

// Adjoint of L = chol(A) in solve form, by the whole CTA (output convention of gp::warp_cholesky_rev: strict lower triangle = SUM of the
// sensitivities of A_ij and A_ji, diagonal = sensitivity of A_ii):
//     P = Phi(L^T dL)  (lower triangle, halved diagonal),   Y = L^-T P L^-1,   dA = (Y + Y^T) / 2.
// The product is spread over the warps (4-row register tiles); each of the two triangular sweeps runs with lane = column (rows)
// 0..31 in warp 0 and, for a 33 x 33 factor, column (row) 32 in warp 1.  About half the instructions of Murray's level-2 reverse
// sweep (gp::warp_cholesky_rev, which the general block kernel still uses).  L must be zero above its diagonal; P is an m x m scratch matrix (row stride LD); G holds dL on entry.
template <typename T>
__device__ __forceinline__ void chol_adj_block(const T* __restrict__ L, const T* __restrict__ invd, T* __restrict__ G, T* __restrict__ P, int m,
                                               T* __restrict__ scratch = nullptr) {
    const int lane = lane_id(), wid = warp_id();
    __syncthreads();
    {   // P = Phi(L^T G): P[i][j] = sum_{k >= i} L[k][i] G[k][j], i >= j   (lane = column j)
        const int lj = lane < m ? lane : m - 1;
        for (int i0 = 4 * wid; i0 < m; i0 += 4 * NW) {
            T acc0 = (T)0, acc1 = (T)0, acc2 = (T)0, acc3 = (T)0;
            for (int k = i0; k < m; ++k) {
                const T g = (k >= lj) ? G[k * LD + lj] : (T)0;         // dL is lower triangular: nothing above the diagonal
                const T* l = L + k * LD + i0;
                acc0 += l[0] * g; acc1 += l[1] * g; acc2 += l[2] * g; acc3 += l[3] * g;
            }
            const T accs[4] = {acc0, acc1, acc2, acc3};
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                const int i = i0 + x;
                if (i < m && lane < m) P[i * LD + lane] = (lane < i) ? accs[x] : (lane == i ? (T)0.5 * accs[x] : (T)0);
            }
        }
        if (m == 33 && wid == 1) {                                      // column 32: zeros above the diagonal, one entry on it
            P[lane * LD + 32] = (T)0;
            if (lane == 0) P[32 * LD + 32] = (T)0.5 * L[32 * LD + 32] * G[32 * LD + 32];
        }
    }
    __syncthreads();
    if (scratch != nullptr) {
        // whole-CTA blocked back substitutions (gp_block.cuh: 4 x 4 register tiles, one barrier per four rows) instead of the two
        // one-warp sweeps below: X = L^-T P, then (X L^-1)^T = L^-T X^T on the transposed matrix; the symmetrisation below does not
        // care which of Y, Y^T sits in P
        gp::blk4_trsm_lowerT_left<T, 33, 33>(L, LD, invd, P, LD, m, m, scratch);
        each_block(m, m, [&](int idx, int i, int j) {
            if (i < j) { const T t = P[i * LD + j]; P[i * LD + j] = P[j * LD + i]; P[j * LD + i] = t; }
        });
        gp::blk4_trsm_lowerT_left<T, 33, 33>(L, LD, invd, P, LD, m, m, scratch);
    } else {
    // X = L^-T P: back substitution down the rows, lane = column
    if (wid < 2) {
        const int col = (wid == 0) ? lane : 32;
        if (col < m && (wid == 0 || lane == 0)) {
            if (m == 33) backsub33<T, LD>(L, invd, P + col);
            else
                for (int i = m - 1; i >= 0; --i) {
                    P[i * LD + col] = dot_sub<T, LD, LD>(P[i * LD + col], L + i, P + col, i + 1, m) * invd[i];
                }
        }
    }
    __syncthreads();
    // Y = X L^-1: y_i = (x_i - sum_{k > i} y_k L[k][i]) / L[i][i], i descending, lane = row
    if (wid < 2) {
        const int row = (wid == 0) ? lane : 32;
        if (row < m && (wid == 0 || lane == 0)) {
            T* xr = P + row * LD;
            if (m == 33) backsub33<T, 1>(L, invd, xr);
            else
                for (int i = m - 1; i >= 0; --i) {
                    xr[i] = dot_sub<T, 1, LD>(xr[i], xr, L + i, i + 1, m) * invd[i];
                }
        }
    }
    }
    __syncthreads();
    each_block(m, m, [&](int idx, int i, int j) {
        if (i > j) G[i * LD + j] = P[i * LD + j] + P[j * LD + i];
        else if (i == j) G[i * LD + i] = P[i * LD + i];
    });
    __syncthreads();
}

// Kernel adjoint of ONE symmetric-input Gram block K(Z, Z) for the warp path (n <= 36 rows, learnable row n - 1): the arithmetic of
// gp::kernel_adjoint_block (q_k = sum_ij W_ij (u_ik - u_jk)^2 in expanded form with V = W U from 4 x 4 register tiles; gradient of the
// learnable row; amplitude gradient), re-laid out so that BOTH operands of the tile product are 16-byte shared-memory loads:
//   U chunk  [row][KV = 36]  (row stride a multiple of 4: u[j][4kq .. 4kq+3] is one LDS.128, and the loader writes with STS.128),
//   W^T      [j][KV]         (W_T[j][i0 .. i0+3] is one LDS.128; W is NOT symmetric here: the block carries dK_ZZ + [dK_ZX | 0] + dK_XX),
// i.e. 2 loads per 16 multiply-adds instead of 8 (kernel adjoint = the longest phase of the adjoint kernel, 86 k of 317 k cycles).
//   dK   : [n][LD] upstream gradient;  Kv: [n][LD] kernel values (both read only)
//   WT   : [n][KV] scratch;  tile: [pad4(n)][KV] scratch;  q, dzl: [d] accumulators (zeroed by the caller);  rs, cs: [>= n]
constexpr int KV = 36;
__device__ inline float kernel_adjoint_sym(const float* __restrict__ dK, const float* __restrict__ Kv, float* __restrict__ WT,
                                           const float* __restrict__ gZ, int n, int d, int kt, float amp, const float* __restrict__ invls,
                                           float* __restrict__ tile, float* __restrict__ q, float* __restrict__ dzl,
                                           float* __restrict__ rs, float* __restrict__ cs) {
    using gp::KC;
    float damp = 0.f;
    const float inv_amp = 1.f / amp;
    const int tid = threadIdx.x;
    each_block(n, n, [&](int idx, int i, int j) {
        const float kv = Kv[i * LD + j], g = dK[i * LD + j];
        float wv;
        if (kt == CLIPGP_KERNEL_RBF) { damp += g * kv * inv_amp; wv = -0.5f * g * kv; }
        else if (kt == CLIPGP_KERNEL_MATERN12) { const float rr = -logf(kv); wv = (kv < 1.f && rr > 0.f) ? (-0.5f * g * kv / rr) : 0.f; }
        else { damp += g * kv * inv_amp; wv = g * amp; }
        WT[j * KV + i] = wv;
    });
    if (tid < 3 * KV) {                                       // rows n .. n+2 of a padded 4-row tile read W_T[j][i >= n]: zero columns
        for (int j = 0; j < n; ++j) if (tid >= n && tid < KV) WT[j * KV + tid] = 0.f;
    }
    __syncthreads();
    const bool dot = (kt == CLIPGP_KERNEL_LINEAR);
    const int rowL = n - 1;
    if (!dot) {
        for (int i = tid; i < n; i += NT) { float t = 0.f; for (int j = 0; j < n; ++j) t += WT[j * KV + i]; rs[i] = t; }      // row sums of W
        for (int j = tid; j < n; j += NT) { float t = 0.f; for (int i = 0; i < n; ++i) t += WT[j * KV + i]; cs[j] = t; }      // column sums
    }
    const int pA = gp::pad4(n);
    constexpr int KQ = KC / 4;
    const int vtiles = (pA >> 2) * KQ;
    const bool vec = ((d & 3) == 0) && ((reinterpret_cast<uintptr_t>(gZ) & 15u) == 0);
    for (int k0 = 0; k0 < d; k0 += KC) {
        __syncthreads();
        // ---- chunk loader: columns [k0, k0 + KC) of Z (scaled by the inverse length-scales), rows >= n and columns >= d zero
        if (vec) {
            const int total = pA * KQ;
            for (int base = tid; base < total; base += 4 * NT) {
                float4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int idx = base + u * NT;
                    const int r = idx / KQ, k = k0 + (idx - r * KQ) * 4;
                    v[u] = (idx < total && r < n && k < d) ? __ldg(reinterpret_cast<const float4*>(gZ + (size_t)r * d + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int idx = base + u * NT;
                    if (idx < total) {
                        const int r = idx / KQ, qq = (idx - r * KQ) * 4, k = k0 + qq;
                        float4 x = v[u];
                        if (!dot && k < d) { const float4 sc = *reinterpret_cast<const float4*>(invls + k); x.x *= sc.x; x.y *= sc.y; x.z *= sc.z; x.w *= sc.w; }
                        *reinterpret_cast<float4*>(tile + r * KV + qq) = x;
                    }
                }
            }
        } else {
            for (int idx = tid; idx < pA * KC; idx += NT) {
                const int r = idx / KC, k = idx - r * KC;
                float v = 0.f;
                if (r < n && k0 + k < d) { v = __ldg(gZ + (size_t)r * d + k0 + k); if (!dot) v *= invls[k0 + k]; }
                tile[r * KV + k] = v;
            }
        }
        __syncthreads();
        if (!dot) {
            for (int tl0 = 0; tl0 < vtiles; tl0 += NT) {              // all threads iterate together: the reduction below shuffles across the warp
                const int tl = tl0 + tid;
                const bool live = tl < vtiles;
                const int it = live ? tl / KQ : 0, kq = tl & (KQ - 1), i0 = it * 4;
                float v[4][4];
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int y = 0; y < 4; ++y) v[x][y] = 0.f;
                const float* wp = WT + i0;
                const float* up = tile + kq * 4;
#pragma unroll 3
                for (int j = 0; j < (live ? n : 0); ++j) {
                    const float4 w4 = *reinterpret_cast<const float4*>(wp + j * KV);
                    const float4 u4 = *reinterpret_cast<const float4*>(up + j * KV);
                    const float w[4] = {w4.x, w4.y, w4.z, w4.w}, u[4] = {u4.x, u4.y, u4.z, u4.w};
#pragma unroll
                    for (int x = 0; x < 4; ++x)
#pragma unroll
                        for (int y = 0; y < 4; ++y) v[x][y] = fmaf(w[x], u[y], v[x][y]);
                }
#pragma unroll
                for (int y = 0; y < 4; ++y) {
                    const int k = k0 + kq * 4 + y;
                    float accq = 0.f;
                    if (live && k < d) {
#pragma unroll
                        for (int x = 0; x < 4; ++x) {
                            const int i = i0 + x;
                            if (i < n) { const float ua = tile[i * KV + kq * 4 + y]; accq += ua * (rs[i] * ua - 2.f * v[x][y]); }
                        }
                    }
                    // lanes l, l + 8, l + 16, l + 24 hold the same feature column (tile index = row tile * 8 + column quad): one shared-memory
                    // atomic per warp and column instead of one per tile (the atomics were 4 % of the adjoint kernel's stall samples)
                    accq += __shfl_xor_sync(FULL, accq, 8);
                    accq += __shfl_xor_sync(FULL, accq, 16);
                    if (lane_id() < KQ && k < d) atomicAdd(&q[k], accq);
                }
            }
            if (tid < KC && k0 + tid < d) {                                // column term sum_j c_j u_jk^2
                float accq = 0.f;
                for (int j = 0; j < n; ++j) { const float uj = tile[j * KV + tid]; accq = fmaf(cs[j] * uj, uj, accq); }
                atomicAdd(&q[k0 + tid], accq);
            }
        }
        if (tid >= NT - KC) {                                               // the learnable row (as an A row and as a B row)
            const int kk = tid - (NT - KC), k = k0 + kk;
            if (k < d) {
                float dz = 0.f;
                const float ul = tile[rowL * KV + kk];
                for (int j = 0; j < n; ++j) {
                    const float uj = tile[j * KV + kk];
                    const float wa = WT[j * KV + rowL];                     // W[rowL][j]
                    const float wb = WT[rowL * KV + j];                     // W[j][rowL]
                    dz = fmaf(wa, dot ? uj : (ul - uj), dz);
                    dz = fmaf(wb, dot ? uj : (ul - uj), dz);
                }
                if (!dot) dz *= 2.f * invls[k];
                dzl[k] += dz;
            }
        }
    }
    __syncthreads();
    return damp;
}

// Offsets of the per-class record in Ksave ([alias flag | K_ZZ n*n | K_ZX n*T | K_XX T*T]).  For aliased classes the
// K_ZX / K_XX part is unused by the forward pass; the adjoint uses it as scratch for d loss / d K_ZZ (n*n <= n*T + T*T for T >= 2).
__device__ __forceinline__ size_t ksave_stride(int n, int T) { return (size_t)1 + (size_t)n * n + (size_t)n * T + (size_t)T * T; }

}  // namespace gpw
}  // namespace clipgp
