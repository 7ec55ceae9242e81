// Per-class building blocks of the GP template weighter kernels (one CTA per class, everything in
// shared memory): streamed Gram accumulation, block Cholesky, triangular solves, Cholesky adjoint.
// Algorithm sheet: oracle/gp_manual.py (validated against autograd through oracle/gp.py).
#pragma once
#include "common.cuh"

namespace clipgp {

// Classes one launch covers: [c_begin, c_begin + c_count) of the C classes the tensors are laid out for (c_count == 0: all).
inline unsigned gp_grid(const clipgp_gp_args* a) { return (unsigned)(a->c_count > 0 ? a->c_count : a->C - a->c_begin); }

namespace gp {

constexpr int kThreads = 128;  // threads per class CTA
constexpr int kThreadsMax = 512;  // launch bound of the general block kernels (wide CTAs for n > 33, see general_threads)

// Threads per class CTA of the general block kernels.  n <= 33: 128 (several classes per SM).  n > 33: the shared-memory footprint
// allows one class per SM, so the only parallelism an SM sees is inside the CTA: every phase but the one-warp factorisations is
// written against blockDim.x and spreads over 16 warps.
inline int general_threads(long long n) {
    static const char* e = getenv("CLIPGP_GP_WIDE_THREADS");
    const int wide = e ? atoi(e) : 512;
    return n > 33 ? wide : kThreads;
}
constexpr int KC = 32;         // feature-dim chunk streamed through shared memory
constexpr int KCP = KC + 1;    // padded row stride of a chunk tile (conflict-free row access)
constexpr int kMaxTiles = 3;   // 4x4 register tiles per thread: 3*128 >= ceil(65/4)^2 = 289

__host__ __device__ __forceinline__ int pad4(int x) { return (x + 3) & ~3; }

// Copy columns [k0, k0+KC) of a row-major [rows, d] global matrix into a [rows_pad][KCP] tile, optionally
// scaled per column (inverse length-scales).  Rows >= rows and columns >= d are zero-filled.
// Loads are issued four-deep per thread (128-bit when d % 4 == 0) before any is consumed, so a thread keeps
// 64 bytes in flight instead of stalling on every element.
__device__ __forceinline__ void load_chunk(float* __restrict__ tile, const float* __restrict__ G, int rows,
                                           int rows_pad, int d, int k0, const float* __restrict__ col_scale) {
    constexpr int V = KC / 4;
    if (((d & 3) == 0) && ((reinterpret_cast<uintptr_t>(G) & 15u) == 0)) {
        const int total = rows_pad * V;
        for (int base = threadIdx.x; base < total; base += 4 * blockDim.x) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int idx = base + u * blockDim.x;
                const int r = idx / V, k = k0 + (idx - r * V) * 4;
                v[u] = (idx < total && r < rows && k < d) ? __ldg(reinterpret_cast<const float4*>(G + (size_t)r * d + k))
                                                          : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int idx = base + u * blockDim.x;
                if (idx < total) {
                    const int r = idx / V, q = (idx - r * V) * 4, k = k0 + q;
                    float4 x = v[u];
                    if (col_scale && k < d) { x.x *= col_scale[k]; x.y *= col_scale[k + 1]; x.z *= col_scale[k + 2]; x.w *= col_scale[k + 3]; }
                    float* t = tile + r * KCP + q;
                    t[0] = x.x; t[1] = x.y; t[2] = x.z; t[3] = x.w;
                }
            }
        }
        return;
    }
    for (int idx = threadIdx.x; idx < rows_pad * KC; idx += blockDim.x) {
        const int r = idx / KC, k = idx - r * KC;
        float v = 0.f;
        if (r < rows && k0 + k < d) {
            v = __ldg(G + (size_t)r * d + k0 + k);
            if (col_scale) v *= col_scale[k0 + k];
        }
        tile[r * KCP + k] = v;
    }
}

// All-threads test X[0 .. count) == Z[0 .. count) (bitwise as floats), with four 128-bit loads in flight per array.
__device__ __forceinline__ int rows_identical(const float* __restrict__ X, const float* __restrict__ Z, int count) {
    int eq = 1;
    if (((count & 3) == 0) && (((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(Z)) & 15u) == 0)) {
        const int c4 = count >> 2;
        const float4* x4 = reinterpret_cast<const float4*>(X);
        const float4* z4 = reinterpret_cast<const float4*>(Z);
        for (int base = threadIdx.x; base < c4; base += 4 * blockDim.x) {
            float4 a[4], b[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = base + u * blockDim.x;
                a[u] = (i < c4) ? __ldg(x4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
                b[u] = (i < c4) ? __ldg(z4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                eq &= (a[u].x == b[u].x) & (a[u].y == b[u].y) & (a[u].z == b[u].z) & (a[u].w == b[u].w);
        }
    } else {
        for (int i = threadIdx.x; i < count; i += blockDim.x) eq &= (__ldg(X + i) == __ldg(Z + i));
    }
    return __syncthreads_and(eq);
}

// acc[t] += sum_k op(a_ik, b_jk) over one chunk for this thread's 4x4 tiles.
// DOT: a*b ; otherwise (a-b)^2.
// Tile enumeration.  General blocks: row-major over (ti, tj).  Symmetric blocks (A rows == B rows): only the
// tiles with tj >= ti, enumerated row by row; the lower part is mirrored when the block is written out.
__device__ __forceinline__ int num_tiles(int tiles_m, int tiles_n, bool sym) {
    return sym ? (tiles_m * (tiles_m + 1)) / 2 : tiles_m * tiles_n;
}
__device__ __forceinline__ void tile_coords(int tile, int tiles_n, bool sym, int& ti, int& tj) {
    if (!sym) { ti = tile / tiles_n; tj = tile - ti * tiles_n; return; }
    int row = 0, left = tile, len = tiles_n;
    while (left >= len) { left -= len; --len; ++row; }
    ti = row; tj = row + left;
}

template <bool DOT, int MT>
__device__ __forceinline__ void gram_chunk(float (&acc)[MT][16], const float* __restrict__ tA,
                                           const float* __restrict__ tB, const int (&tis)[MT], const int (&tjs)[MT],
                                           const int kb = 0, const int ke = KC) {
#pragma unroll
    for (int t = 0; t < MT; ++t) {
        if (tis[t] >= 0) {
            const int ti = tis[t], tj = tjs[t];
            const float* pa = tA + (ti * 4) * KCP;
            const float* pb = tB + (tj * 4) * KCP;
#pragma unroll 4
            for (int k = kb; k < ke; ++k) {
                float a[4], b[4];
#pragma unroll
                for (int x = 0; x < 4; ++x) { a[x] = pa[x * KCP + k]; b[x] = pb[x * KCP + k]; }
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int y = 0; y < 4; ++y) {
                        if (DOT) acc[t][x * 4 + y] = fmaf(a[x], b[y], acc[t][x * 4 + y]);
                        else { const float df = a[x] - b[y]; acc[t][x * 4 + y] = fmaf(df, df, acc[t][x * 4 + y]); }
                    }
            }
        }
    }
}

__device__ __forceinline__ float kernel_value(int kernel_type, float acc, float amp) {
    if (kernel_type == CLIPGP_KERNEL_RBF) return amp * expf(-0.5f * acc);                     // os * exp(-r2/2)
    if (kernel_type == CLIPGP_KERNEL_MATERN12) return expf(-sqrtf(fmaxf(acc, 1e-30f)));        // exp(-r)
    return amp * acc;                                                                          // v * <a,b>
}

// Full Gram block K(A rows, B rows) -> out[i*ldo + j] (type OutT), i < nA, j < nB.
//   gA/gB: global row-major [nA,d] / [nB,d]; if gB == gA the second tile load is skipped.
//   raw_out (optional, float [nA][ldr]): the un-transformed accumulator (r^2 or dot), needed by the adjoint.
//   MT: 4x4 register tiles per thread (MT * blockDim.x must cover the tiles; MT = 1 keeps the register footprint small when the
//   block is called from the warp-path algebra kernel, n <= 33 -> 45 symmetric tiles).
//   KS = 2 (MT = 1 only, 2 * tiles <= blockDim.x): two adjacent lanes share one tile and take one half of every chunk's
//   feature columns each; their partial sums meet through one shuffle per accumulator after the last chunk.  With n <= 33 the
//   45 symmetric tiles then keep 90 of the 128 threads busy instead of 45 (the Gram phase of a class is latency bound).
template <typename OutT, int MT = kMaxTiles, int KS = 1>
__device__ void gram_block(OutT* __restrict__ out, int ldo, float* __restrict__ raw_out, int ldr,
                           const float* __restrict__ gA, int nA, const float* __restrict__ gB, int nB, int d,
                           int kernel_type, float amp, const float* __restrict__ inv_ls, float* tileA, float* tileB) {
    static_assert(KS == 1 || (KS == 2 && MT == 1), "k-split needs one tile per thread");
    const bool same = (gA == gB);
    const bool sym = same && (nA == nB);
    const int pA = pad4(nA), pB = pad4(nB);
    const int tiles_m = pA >> 2, tiles_n = pB >> 2;
    const bool dot = (kernel_type == CLIPGP_KERNEL_LINEAR);
    float acc[MT][16];
    int tis[MT], tjs[MT];
    const int ntiles = num_tiles(tiles_m, tiles_n, sym);
    const int khalf = (KS == 2) ? (int)(threadIdx.x & 1) : 0;
#pragma unroll
    for (int t = 0; t < MT; ++t) {
        const int tile = (KS == 2) ? (int)(threadIdx.x >> 1) : (int)(threadIdx.x + t * blockDim.x);
        tis[t] = -1; tjs[t] = 0;
        if (tile < ntiles) tile_coords(tile, tiles_n, sym, tis[t], tjs[t]);
#pragma unroll
        for (int x = 0; x < 16; ++x) acc[t][x] = 0.f;
    }
    for (int k0 = 0; k0 < d; k0 += KC) {
        __syncthreads();
        load_chunk(tileA, gA, nA, pA, d, k0, dot ? nullptr : inv_ls);
        if (!same) load_chunk(tileB, gB, nB, pB, d, k0, dot ? nullptr : inv_ls);
        __syncthreads();
        const int kb = (KS == 2) ? khalf * (KC / 2) : 0, ke = (KS == 2) ? kb + KC / 2 : KC;
        if (dot) gram_chunk<true, MT>(acc, tileA, same ? tileA : tileB, tis, tjs, kb, ke);
        else gram_chunk<false, MT>(acc, tileA, same ? tileA : tileB, tis, tjs, kb, ke);
    }
    if (KS == 2) {      // both lanes of a pair end up with the full sums; the even lane writes
#pragma unroll
        for (int x = 0; x < 16; ++x) acc[0][x] += __shfl_xor_sync(0xffffffffu, acc[0][x], 1);
        if (khalf) tis[0] = -1;
    }
#pragma unroll
    for (int t = 0; t < MT; ++t) {
        if (tis[t] >= 0) {
            const int ti = tis[t], tj = tjs[t];
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y) {
                    const int i = ti * 4 + x, j = tj * 4 + y;
                    if (i < nA && j < nB) {
                        const float a = acc[t][x * 4 + y];
                        const OutT kv = (OutT)kernel_value(kernel_type, a, amp);
                        if (raw_out) raw_out[i * ldr + j] = a;
                        if (out) out[i * ldo + j] = kv;
                        if (sym && tj > ti) {      // mirror (diagonal tiles hold both halves already)
                            if (raw_out) raw_out[j * ldr + i] = a;
                            if (out) out[j * ldo + i] = kv;
                        }
                    }
                }
        }
    }
    __syncthreads();
}

// Left-looking lower Cholesky executed by ONE warp (call from a single warp; lanes own rows lane, lane+32, lane+64),
// in place on the lower triangle, __syncwarp only.  invd[j] = 1 / L[j][j].  Returns true (warp-uniformly) when a
// pivot was not strictly positive.  n <= 96.
template <typename T>
__device__ bool warp_cholesky(T* __restrict__ A, int n, int ld, T* __restrict__ invd) {
    const int lane = threadIdx.x & 31;
    bool fail = false;
    for (int j = 0; j < n; ++j) {
        T s[3];
#pragma unroll
        for (int u = 0; u < 3; ++u) {
            const int i = lane + 32 * u;
            T acc = (T)0;
            if (i >= j && i < n) {
                acc = A[i * ld + j];
                const T* ri = A + i * ld;
                const T* rj = A + j * ld;
                for (int k = 0; k < j; ++k) acc -= ri[k] * rj[k];
            }
            s[u] = acc;
        }
        const int slot = j >> 5;
        const T sj = __shfl_sync(0xffffffffu, slot == 0 ? s[0] : (slot == 1 ? s[1] : s[2]), j & 31);
        if (!(sj > (T)0)) fail = true;
        const T inv = (T)1 / sqrt(sj);
        __syncwarp();
#pragma unroll
        for (int u = 0; u < 3; ++u) {
            const int i = lane + 32 * u;
            if (i > j && i < n) A[i * ld + j] = s[u] * inv;
        }
        if (lane == 0) { A[j * ld + j] = sj * inv; invd[j] = inv; }
        __syncwarp();
    }
    return fail;
}

// Adjoint of L = chol(A) by ONE warp: Murray's level-2 reverse sweep ("Differentiation of the Cholesky decomposition",
// 2016).  In: L (lower), invd, G = dL (lower triangle; destroyed).  Out (in G): the strict lower triangle holds the SUM of
// the sensitivities of A_ij and A_ji, the diagonal the sensitivity of A_ii; i.e. the symmetric gradient is
// dA = (tril(G,-1) + tril(G,-1)^T) / 2 + diag(G)   (use sym_from_rev below).
template <typename T>
__device__ void warp_cholesky_rev(const T* __restrict__ L, int ldl, const T* __restrict__ invd, T* __restrict__ G, int ld, int n) {
    const int lane = threadIdx.x & 31;
    for (int j = n - 1; j >= 0; --j) {
        const T inv = invd[j];
        T part = (T)0;
        for (int i = j + 1 + lane; i < n; i += 32) part += L[i * ldl + j] * G[i * ld + j];
        part = warp_sum(part);
        const T db = (G[j * ld + j] - part * inv) * inv;            // (d_bar - c^T c_bar / d) / d
        __syncwarp();
        for (int i = j + 1 + lane; i < n; i += 32) G[i * ld + j] *= inv;   // c_bar /= d
        __syncwarp();
        for (int k = lane; k < j; k += 32) {                            // r_bar -= d_bar r + c_bar^T B
            T sacc = G[j * ld + k] - db * L[j * ldl + k];
            for (int i = j + 1; i < n; ++i) sacc -= G[i * ld + j] * L[i * ldl + k];
            G[j * ld + k] = sacc;
        }
        for (int i = j + 1 + lane; i < n; i += 32) {                    // B_bar -= c_bar r
            const T cbi = G[i * ld + j];
            T* gi = G + i * ld;
            const T* rj = L + j * ldl;
            for (int k = 0; k < j; ++k) gi[k] -= cbi * rj[k];
        }
        if (lane == 0) G[j * ld + j] = db * (T)0.5;
        __syncwarp();
    }
}
template <typename T>
__device__ __forceinline__ T sym_from_rev(const T* __restrict__ G, int ld, int i, int j) {
    return i == j ? G[i * ld + i] : (T)0.5 * (i > j ? G[i * ld + j] : G[j * ld + i]);
}

// ---------------------------------------------------------------------------------------------------------------------------
// Whole-CTA (right-looking) factorisation, solves and Cholesky adjoint: one barrier per column / row, every warp busy.
// The one-warp routines above cost ~700-1000 cycles per column whatever the arithmetic type (a dependent chain of shared-memory
// dot products issued by one warp while the rest of the CTA waits); these spread each step's rank-1 update over the CTA
// (rows over warps, columns over lanes).  Used by the general block kernels (n up to 65, 128-512 threads).

// A = L L^T in place on the lower triangle of A [n][ld]; optionally B [n][ldb] (ncol columns) <- L^-1 B in the same sweep
// (forward elimination of the augmented matrix [A | B]).  Step j applies the update of the UNSCALED column j
// (a_ij a_kj / a_jj): nobody writes column j after step j-1, so the scaling l_ij = a_ij / sqrt(a_jj) is one parallel pass at the end.
// invd[j] = 1 / L[j][j].  Returns true (uniformly) when a pivot was not strictly positive.  flag: one shared int.
template <typename T>
__device__ bool block_cholesky_solve(T* __restrict__ A, int n, int ld, T* __restrict__ invd, T* __restrict__ B, int ldb, int ncol,
                                     int* __restrict__ flag) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __syncthreads();
    if (threadIdx.x == 0) *flag = 0;
    for (int j = 0; j + 1 < n; ++j) {
        __syncthreads();
        const T ip = (T)1 / A[j * ld + j];
        const T* cj = A + j;
        for (int i = j + 1 + warp; i < n; i += nw) {
            const T f = A[i * ld + j] * ip;
            T* ri = A + i * ld;
            for (int k = j + 1 + lane; k <= i; k += 32) ri[k] -= f * cj[k * ld];
            if (B != nullptr) {
                T* bi = B + i * ldb;
                const T* bj = B + j * ldb;
                for (int c = lane; c < ncol; c += 32) bi[c] -= f * bj[c];
            }
        }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        const T p = A[j * ld + j];
        if (!(p > (T)0)) *flag = 1;
        const T dd = sqrt(p);
        A[j * ld + j] = dd;
        invd[j] = (T)1 / dd;
    }
    __syncthreads();
    for (int i = 1 + warp; i < n; i += nw)
        for (int k = lane; k < i; k += 32) A[i * ld + k] *= invd[k];
    if (B != nullptr)
        for (int i = warp; i < n; i += nw) {
            const T s = invd[i];
            for (int c = lane; c < ncol; c += 32) B[i * ldb + c] *= s;
        }
    __syncthreads();
    return *flag != 0;
}

// Solve L^T X = B in place (B [n][ldb], ncol columns): right-looking back substitution, one barrier per row.
template <typename T>
__device__ void block_trsm_lowerT_left(const T* __restrict__ L, int ldl, const T* __restrict__ invd, T* __restrict__ B, int ldb,
                                       int n, int ncol) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int k = n - 1; k > 0; --k) {
        __syncthreads();
        const T ik = invd[k];
        const T* bk = B + k * ldb;
        const T* lk = L + k * ldl;
        for (int i = warp; i < k; i += nw) {
            const T f = lk[i] * ik;
            T* bi = B + i * ldb;
            for (int c = lane; c < ncol; c += 32) bi[c] -= f * bk[c];
        }
    }
    __syncthreads();
    for (int i = warp; i < n; i += nw) {
        const T s = invd[i];
        for (int c = lane; c < ncol; c += 32) B[i * ldb + c] *= s;
    }
    __syncthreads();
}

// Adjoint of L = chol(A) in solve form: dA = sym(L^-T Phi(L^T dL) L^-1), Phi = lower triangle with halved diagonal.
// In: L, invd, G = dL (lower triangle read).  Out: G = the SYMMETRIC gradient dA (full matrix).  W: scratch [n][ld].
template <typename T>
__device__ void block_cholesky_adjoint(const T* __restrict__ L, int ldl, const T* __restrict__ invd, T* __restrict__ G,
                                       T* __restrict__ W, int ld, int n) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __syncthreads();
    for (int i = warp; i < n; i += nw)
        for (int j = lane; j < n; j += 32) {
            T s = (T)0;
            if (j <= i) {
                for (int k = i; k < n; ++k) s += L[k * ldl + i] * G[k * ld + j];
                if (j == i) s *= (T)0.5;
            }
            W[i * ld + j] = s;
        }
    block_trsm_lowerT_left<T>(L, ldl, invd, W, ld, n, n);          // X = L^-T Phi
    for (int i = warp; i < n; i += nw)
        for (int j = lane; j < n; j += 32) G[i * ld + j] = W[j * ld + i];
    block_trsm_lowerT_left<T>(L, ldl, invd, G, ld, n, n);          // M^T = L^-T X^T
    for (int i = 1 + warp; i < n; i += nw)
        for (int j = lane; j < i; j += 32) {
            const T v = (T)0.5 * (G[i * ld + j] + G[j * ld + i]);
            G[i * ld + j] = v; G[j * ld + i] = v;
        }
    __syncthreads();
}

// Solve L X = B in place (B is [n][ncol], leading dim ldb); one thread per column.
template <typename T>
__device__ void trsm_lower_left(const T* __restrict__ L, int ldl, const T* __restrict__ invd, T* __restrict__ B,
                                int ldb, int n, int ncol) {
    for (int c = threadIdx.x; c < ncol; c += blockDim.x) {
        for (int i = 0; i < n; ++i) {
            T s = B[i * ldb + c];
            for (int k = 0; k < i; ++k) s -= L[i * ldl + k] * B[k * ldb + c];
            B[i * ldb + c] = s * invd[i];
        }
    }
    __syncthreads();
}

// Solve L^T X = B in place (back substitution); one thread per column.
template <typename T>
__device__ void trsm_lowerT_left(const T* __restrict__ L, int ldl, const T* __restrict__ invd, T* __restrict__ B,
                                 int ldb, int n, int ncol) {
    for (int c = threadIdx.x; c < ncol; c += blockDim.x) {
        for (int i = n - 1; i >= 0; --i) {
            T s = B[i * ldb + c];
            for (int k = i + 1; k < n; ++k) s -= L[k * ldl + i] * B[k * ldb + c];
            B[i * ldb + c] = s * invd[i];
        }
    }
    __syncthreads();
}

// One Gram block's contribution to the kernel adjoints, from the kernel values the forward pass saved.
//   dK   : [nA][ld] float (smem) upstream gradient of K(A rows, B rows)
//   Wm   : [nA][ld] float (smem) in: K values of the block; out: W = d loss / d raw  (raw = r^2 or <a,b>)
//   q    : [d] accumulates sum_ij W_ij (u_ik - u_jk)^2 (rbf / matern; u = z / lengthscale), evaluated in expanded form
//          sum_i r_i u_ik^2 + sum_j c_j u_jk^2 - 2 sum_i u_ik (W U_B)_ik  with 4x4 register tiles for W U_B
//   dzl  : [d] accumulates the gradient of one designated row (rowA as an A row and/or rowB as a B row; -1 = none)
//   rs/cs: [>= nA] / [>= nB] scratch for the row / column sums of W
// Returns this thread's partial of d loss / d amp (outputscale or variance).
__device__ inline float kernel_adjoint_block(const float* dK, int ld, float* Wm, const float* gA, int nA, const float* gB,
                                      int nB, int d, int kt, float amp, const float* invls, float* tileA, float* tileB,
                                      float* q, float* dzl, int rowA, int rowB, float* rs, float* cs) {
    float damp = 0.f;
    const float inv_amp = 1.f / amp;
    for (int idx = threadIdx.x; idx < nA * nB; idx += blockDim.x) {
        const int i = idx / nB, j = idx - i * nB;
        const float kv = Wm[i * ld + j];
        const float g = dK[i * ld + j];
        float wv;
        if (kt == CLIPGP_KERNEL_RBF) {
            damp += g * kv * inv_amp;              // dK/d os = exp(-r2/2) = K / os
            wv = -0.5f * g * kv;                   // dK/d r2 = -K/2
        } else if (kt == CLIPGP_KERNEL_MATERN12) {
            const float rr = -logf(kv);            // K = exp(-r)
            wv = (kv < 1.f && rr > 0.f) ? (-0.5f * g * kv / rr) : 0.f;   // r2 <= 1e-30 (K == 1): clamp kills the gradient
        } else {
            damp += g * kv * inv_amp;              // dK/d v = <a,b> = K / v
            wv = g * amp;                          // dK/d dot = v
        }
        Wm[i * ld + j] = wv;
    }
    __syncthreads();
    const bool same = (gA == gB);
    const bool dot = (kt == CLIPGP_KERNEL_LINEAR);
    if (!dot) {
        for (int i = threadIdx.x; i < nA; i += blockDim.x) { float t = 0.f; for (int j = 0; j < nB; ++j) t += Wm[i * ld + j]; rs[i] = t; }
        for (int j = threadIdx.x; j < nB; j += blockDim.x) { float t = 0.f; for (int i = 0; i < nA; ++i) t += Wm[i * ld + j]; cs[j] = t; }
    }
    const int pA = pad4(nA), pB = pad4(nB);
    constexpr int KQ = KC / 4;
    const int vtiles = (pA >> 2) * KQ;
    for (int k0 = 0; k0 < d; k0 += KC) {
        __syncthreads();
        load_chunk(tileA, gA, nA, pA, d, k0, dot ? nullptr : invls);
        if (!same) load_chunk(tileB, gB, nB, pB, d, k0, dot ? nullptr : invls);
        __syncthreads();
        const float* tB = same ? tileA : tileB;
        if (!dot) {
            for (int tile = threadIdx.x; tile < vtiles; tile += blockDim.x) {
                const int it = tile / KQ, kq = tile - it * KQ;
                const int i0 = it * 4;
                float v[4][4];
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int y = 0; y < 4; ++y) v[x][y] = 0.f;
                const float* w0 = Wm + (i0 + 0 < nA ? i0 + 0 : 0) * ld;
                const float* w1 = Wm + (i0 + 1 < nA ? i0 + 1 : 0) * ld;
                const float* w2 = Wm + (i0 + 2 < nA ? i0 + 2 : 0) * ld;
                const float* w3 = Wm + (i0 + 3 < nA ? i0 + 3 : 0) * ld;
                const float* ub = tB + kq * 4;
                for (int j = 0; j < nB; ++j) {
                    const float w[4] = {w0[j], w1[j], w2[j], w3[j]};
                    const float u[4] = {ub[j * KCP], ub[j * KCP + 1], ub[j * KCP + 2], ub[j * KCP + 3]};
#pragma unroll
                    for (int x = 0; x < 4; ++x)
#pragma unroll
                        for (int y = 0; y < 4; ++y) v[x][y] = fmaf(w[x], u[y], v[x][y]);
                }
#pragma unroll
                for (int y = 0; y < 4; ++y) {
                    const int k = k0 + kq * 4 + y;
                    if (k < d) {
                        float accq = 0.f;
#pragma unroll
                        for (int x = 0; x < 4; ++x) {
                            const int i = i0 + x;
                            if (i < nA) { const float ua = tileA[i * KCP + kq * 4 + y]; accq += ua * (rs[i] * ua - 2.f * v[x][y]); }
                        }
                        atomicAdd(&q[k], accq);
                    }
                }
            }
            if (threadIdx.x < KC && k0 + threadIdx.x < d) {            // column term sum_j c_j u_jk^2
                const int kk = threadIdx.x;
                float accq = 0.f;
                for (int j = 0; j < nB; ++j) { const float uj = tB[j * KCP + kk]; accq = fmaf(cs[j] * uj, uj, accq); }
                atomicAdd(&q[k0 + kk], accq);
            }
        }
        if ((rowA >= 0 || rowB >= 0) && threadIdx.x >= blockDim.x - KC) {   // the one learnable row: O(nA + nB) per column
            const int kk = threadIdx.x - (blockDim.x - KC);
            const int k = k0 + kk;
            if (k < d) {
                float dz = 0.f;
                if (rowA >= 0) {
                    const float ui = tileA[rowA * KCP + kk];
                    for (int j = 0; j < nB; ++j) {
                        const float uj = tB[j * KCP + kk];
                        dz = fmaf(Wm[rowA * ld + j], dot ? uj : (ui - uj), dz);
                    }
                }
                if (rowB >= 0) {
                    const float uj = tB[rowB * KCP + kk];
                    for (int i = 0; i < nA; ++i) {
                        const float ui = tileA[i * KCP + kk];
                        dz = fmaf(Wm[i * ld + rowB], dot ? ui : (uj - ui), dz);
                    }
                }
                if (!dot) dz *= 2.f * invls[k];
                dzl[k] += dz;
            }
        }
    }
    __syncthreads();
    return damp;
}

// Warp-level sparsemax over T <= 64 values (lane holds elements lane and lane+32).
// Sort-free (Michelot's fixed point on the support; same support and threshold as the sorted rule k z_(k) > cumsum_k - 1).
// Returns w0/w1 and the support size (entmax SparsemaxFunction.forward).
__device__ __forceinline__ void warp_sparsemax(float f0, float f1, int T, float& w0, float& w1, int& ksz) {
    const int lane = threadIdx.x & 31;
    const bool v0 = lane < T, v1 = lane + 32 < T;
    float mx = fmaxf(v0 ? f0 : -INFINITY, v1 ? f1 : -INFINITY);
    mx = warp_max(mx);
    const float z0 = f0 - mx, z1 = f1 - mx;
    // support by Michelot's fixed point: tau <- (sum_S z - 1) / |S|, S <- {z > tau}, until S stops shrinking.  It ends at the support of
    // the sort-based rule (k z_(k) > cumsum_k - 1) in a handful of warp reductions instead of a T-step rank loop.
    bool s0 = v0, s1 = v1;
    int cnt = T;
    float tau;
    for (;;) {
        const float ssum = warp_sum((s0 ? z0 : 0.f) + (s1 ? z1 : 0.f));
        tau = (ssum - 1.f) / (float)cnt;
        const bool n0 = s0 && z0 > tau, n1 = s1 && z1 > tau;
        const int c = __popc(__ballot_sync(0xffffffffu, n0)) + __popc(__ballot_sync(0xffffffffu, n1));
        if (c == cnt) break;
        s0 = n0; s1 = n1; cnt = c;
    }
    w0 = v0 ? fmaxf(z0 - tau, 0.f) : 0.f;
    w1 = v1 ? fmaxf(z1 - tau, 0.f) : 0.f;
    ksz = cnt;
}

}  // namespace gp
}  // namespace clipgp
