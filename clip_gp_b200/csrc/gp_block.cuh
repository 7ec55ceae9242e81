// Whole-CTA dense algebra for the general GP block kernels (n up to 65, T up to 64; gp_forward.cu / gp_backward.cu).
//
// The one-warp left-looking routines of gp_common.cuh cost 700-1400 cycles per column whatever the arithmetic type: a dependent
// chain of shared-memory dot products issued by one warp while the rest of the CTA waits (tools/gp_general_ts.py: 90 k cycles for the
// fp64 factor + solve of a 65 x 65 block, 160 k for its adjoint).  Here every step of a factorisation / substitution is ONE
// or two barriers per FOUR columns: each thread keeps one 4 x 4 tile of the matrix in REGISTERS for the whole sweep, the panel of the
// current block travels through shared memory, and a trailing tile's update is 32 loaded values feeding 64 independent FMAs.
// (Unblocked predecessors, measured on the same 65 x 65 fp64 factor + solve: one-warp left-looking 173 k cycles, whole-CTA
// right-looking on shared memory 91 k, register tiles with one barrier per column 81 k -- tools/micro/block_algebra.cu.)  Products are 4 x 4 register tiles with strided columns (odd leading
// dimensions: conflict-free scalar loads).
#pragma once
#include "gp_common.cuh"

namespace clipgp {
namespace gp {

#ifndef BLK4_TS
#define BLK4_TS(i) do { } while (0)       // tools/micro/block_algebra.cu defines it to record clock64() per phase
#endif

// Reciprocal of a pivot on the critical path of a sweep: hardware seed + Newton steps, no special-case branch (the library
// reciprocals carry a guarded slow path: +1 branch and a longer dependent chain per pivot).  Pivots are positive normal numbers or the
// factorisation is flagged as failed anyway; accuracy: <= 1 ulp (fp64, two steps from the 20-bit seed), <= 1 ulp (fp32, one step).
__device__ __forceinline__ double rcp_t(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(fma(-x, r, 1.0), r, r);
    return fma(fma(-x, r, 1.0), r, r);
}
__device__ __forceinline__ float rcp_t(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return fmaf(fmaf(-x, r, 1.f), r, r);
}

// 4 consecutive elements (16-byte aligned for float, 32 for double) as vector loads / stores
__device__ __forceinline__ void ld4(const float* p, float (&v)[4]) { const float4 t = *reinterpret_cast<const float4*>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
__device__ __forceinline__ void ld4(const double* p, double (&v)[4]) {
    const double2 a = *reinterpret_cast<const double2*>(p), b = *reinterpret_cast<const double2*>(p + 2);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
__device__ __forceinline__ void st4(float* p, const float (&v)[4]) { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
__device__ __forceinline__ void st4(double* p, const double (&v)[4]) {
    *reinterpret_cast<double2*>(p) = make_double2(v[0], v[1]); *reinterpret_cast<double2*>(p + 2) = make_double2(v[2], v[3]);
}

// Group g (4 elements) of a scratch row that holds `ngroups` groups and is only ever accessed through these two helpers.  float: the
// 16-byte vector at row + 4 g.  double: the two 16-byte halves live in two planes of the row (first halves of all groups, then the second
// halves), so the eight threads of a quarter warp that read consecutive groups touch eight consecutive vectors; a 32-byte-strided pair
// of double2 accesses is a 2-way bank conflict (profiles/r2_gp_bank_conflicts.txt, gp_block.cuh ld4 / st4 lines).
__device__ __forceinline__ void ld4g(const float* row, int ngroups, int g, float (&v)[4]) { (void)ngroups; ld4(row + 4 * g, v); }
__device__ __forceinline__ void st4g(float* row, int ngroups, int g, const float (&v)[4]) { (void)ngroups; st4(row + 4 * g, v); }
__device__ __forceinline__ void ld4g(const double* row, int ngroups, int g, double (&v)[4]) {
    const double2 a = *reinterpret_cast<const double2*>(row + 2 * g), b = *reinterpret_cast<const double2*>(row + 2 * ngroups + 2 * g);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
__device__ __forceinline__ void st4g(double* row, int ngroups, int g, const double (&v)[4]) {
    *reinterpret_cast<double2*>(row + 2 * g) = make_double2(v[0], v[1]);
    *reinterpret_cast<double2*>(row + 2 * ngroups + 2 * g) = make_double2(v[2], v[3]);
}

// rows of a [.][ld] matrix of T at p can be read / written as 16-byte vectors at column offsets that are multiples of 4
template <typename T>
__device__ __forceinline__ bool rows_vec16(const T* p, int ld) {
    return p != nullptr && (reinterpret_cast<uintptr_t>(p) & 15u) == 0 && ((ld * (int)sizeof(T)) & 15) == 0;
}

template <int MAXN, int MAXB>
struct Blk4 {
    static constexpr int NT = (MAXN + 3) / 4, NA = NT * (NT + 1) / 2, NB = MAXB / 4, NP = 4 * NT;
    static constexpr int kCholScratch = 36 + 8 * NP + 4 * (MAXB > 0 ? MAXB : 4);      // elements of T
    static constexpr int kCholThreads = NA + NT * NB;
};

// Blocked (block size 4 = tile size) right-looking factorisation A = L L^T in place on the lower triangle of A [n][ld] (n <= MAXN) and,
// in the same sweep, B [n][ldb] (ncol <= MAXB columns) <- L^-1 B (forward elimination of the augmented matrix [A | B]; MAXB = 0:
// the factor alone).  Every thread owns ONE 4 x 4 tile of [A | B] in registers (lower-triangular tiles of A first, then the tiles of B).
// Big step J = columns 4J .. 4J + 3, two barriers:
//   (a) the owner of the diagonal tile eliminates it in registers (LDL^T form: pivots d, u = unscaled entries, g = u / d) and publishes it;
//   (b) the panel tiles below it finish their four columns against the diagonal block and publish u and g per row; the tiles of row
//       block J of B finish their rows and publish them;
//   (c) every trailing tile takes the rank-4 update  m -= G_rows . U_cols^T  (32 loaded values feed 64 independent FMAs).
// The matrix keeps UNSCALED columns until the end (l_ij = u_ij / sqrt(d_j) is one parallel pass), as in the unblocked sweep.
// A single-column sweep costs 500-1200 cycles per column on this machine however little work a step holds (in-order issue of the
// step's ~60 instructions per warp + two shared-memory round trips + barrier: tools/micro/step_cost.cu); this form pays that per FOUR columns.
// Needs blockDim.x >= Blk4<MAXN, MAXB>::kCholThreads; scr: kCholScratch elements of T, 16-byte aligned.
// invd[j] = 1 / L[j][j].  Returns true (uniformly) when a pivot was not strictly positive.
template <typename T, int MAXN, int MAXB>
__device__ bool blk4_cholesky_solve(T* __restrict__ A, int n, int ld, T* __restrict__ invd, T* __restrict__ B, int ldb, int ncol,
                                    T* __restrict__ scr, int* __restrict__ flag) {
    using P = Blk4<MAXN, MAXB>;
    constexpr int NT = P::NT, NA = P::NA, NB = P::NB, NP = P::NP, MB = (MAXB > 0 ? MAXB : 4);
    static_assert(MAXB % 4 == 0, "B tiles are 4 columns wide");
    T* dU = scr;               // [4][4] diagonal block, unscaled (lower triangle)
    T* dG = scr + 16;          // [4][4] g = u / d (strict lower triangle)
    T* dR = scr + 32;          // [4]    1 / d
    T* pU = scr + 36;          // [4][NP] panel, unscaled; column-of-the-block major: the 4 rows of a tile are one vector, and the tiles
    T* pG = pU + 4 * NP;       // [4][NP] panel / pivots     of a warp (consecutive tile columns) read consecutive vectors -- no bank conflicts
    T* bR = pG + 4 * NP;       // [4][MB] rows 4J .. 4J + 3 of B, final (unscaled)
    const int tid = threadIdx.x;
    int ti, tk;                // tile row; tile column (A part: tk <= ti; B part: column block tk of B)
    bool isB = false, owner = true;
    if (tid < NA) {
        ti = (int)((sqrtf(8.f * (float)tid + 1.f) - 1.f) * 0.5f);
        while (ti * (ti + 1) / 2 > tid) --ti;
        while ((ti + 1) * (ti + 2) / 2 <= tid) ++ti;
        tk = tid - ti * (ti + 1) / 2;
    } else if (tid < NA + NT * NB) {
        isB = true; ti = (tid - NA) / (NB > 0 ? NB : 1); tk = (tid - NA) - ti * NB;
    } else { owner = false; ti = 0; tk = 0; }
    const int i0 = 4 * ti, c0 = 4 * tk;
    T m[4][4];
    __syncthreads();
    if (tid == 0) *flag = 0;
    // tile rows as one 16-byte-vector access where the four columns are all inside the triangle / the column range (a scalar walk of
    // 4-column tiles is a 4-way bank conflict: ncu, profiles/r2_gp_bank_conflicts.txt); ragged edges and diagonal tiles stay scalar
    const bool vA = rows_vec16(A, ld), vB = rows_vec16(B, ldb);
#pragma unroll
    for (int x = 0; x < 4; ++x) {
        const int i = i0 + x;
        T r[4] = {(T)0, (T)0, (T)0, (T)0};
        if (owner && i < n) {
            if (!isB) {
                if (vA && c0 + 3 <= i) ld4(A + i * ld + c0, r);
                else {
#pragma unroll
                    for (int y = 0; y < 4; ++y) if (c0 + y <= i) r[y] = A[i * ld + c0 + y];
                }
            } else {
                if (vB && c0 + 3 < ncol) ld4(B + i * ldb + c0, r);
                else {
#pragma unroll
                    for (int y = 0; y < 4; ++y) if (c0 + y < ncol) r[y] = B[i * ldb + c0 + y];
                }
            }
        }
#pragma unroll
        for (int y = 0; y < 4; ++y) m[x][y] = r[y];
    }
    const int nblk = (n + 3) >> 2;
    BLK4_TS(0);
    for (int J = 0; J < nblk; ++J) {
        const int j0 = 4 * J;
        if (J == 2) BLK4_TS(1);
        // ---- (a) diagonal tile
        if (owner && !isB && ti == J && tk == J) {
            T r[4], g[4][4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                r[k] = (j0 + k < n) ? rcp_t(m[k][k]) : (T)0;
#pragma unroll
                for (int x = 0; x < 4; ++x) g[x][k] = (x > k) ? m[x][k] * r[k] : (T)0;
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int y = 0; y < 4; ++y)
                        if (y > k && x >= y) m[x][y] -= g[x][k] * m[y][k];
            }
#pragma unroll
            for (int x = 0; x < 4; ++x) { st4(dU + 4 * x, m[x]); st4(dG + 4 * x, g[x]); }
            st4(dR, r);
        }
        if (J == 2) BLK4_TS(2);
        __syncthreads();
        if (J == 2) BLK4_TS(3);
        // ---- (b) panel below the diagonal tile; row block J of B
        if (owner && !isB && tk == J && ti > J) {
            T r[4], u[4][4];
            ld4(dR, r);
#pragma unroll
            for (int y = 0; y < 4; ++y) ld4(dU + 4 * y, u[y]);
            T g[4][4];
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                    g[x][k] = m[x][k] * r[k];
#pragma unroll
                    for (int y = 0; y < 4; ++y) if (y > k) m[x][y] -= g[x][k] * u[y][k];
                }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const T uc[4] = {m[0][k], m[1][k], m[2][k], m[3][k]}, gc[4] = {g[0][k], g[1][k], g[2][k], g[3][k]};
                st4g(pU + NP * k, NT, ti, uc); st4g(pG + NP * k, NT, ti, gc);
            }
        }
        if (MAXB > 0 && owner && isB && ti == J) {
            T g[4][4];
#pragma unroll
            for (int x = 0; x < 4; ++x) ld4(dG + 4 * x, g[x]);
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
                for (int x = 0; x < 4; ++x)
                    if (x > k) {
#pragma unroll
                        for (int y = 0; y < 4; ++y) m[x][y] -= g[x][k] * m[k][y];
                    }
#pragma unroll
            for (int x = 0; x < 4; ++x) st4g(bR + MB * x, MB / 4, tk, m[x]);
        }
        if (J + 1 == nblk) break;
        if (J == 2) BLK4_TS(4);
        __syncthreads();
        if (J == 2) BLK4_TS(5);
        // ---- (c) rank-4 update of the trailing tiles
        if (owner && ti > J && (isB || tk > J)) {
            T g[4][4], u[4][4];                                                          // g[k][x], u[k][y]
#pragma unroll
            for (int k = 0; k < 4; ++k) ld4g(pG + NP * k, NT, ti, g[k]);
            if (!isB) {
#pragma unroll
                for (int k = 0; k < 4; ++k) ld4g(pU + NP * k, NT, tk, u[k]);
#pragma unroll
                for (int k = 0; k < 4; ++k)
#pragma unroll
                    for (int x = 0; x < 4; ++x)
#pragma unroll
                        for (int y = 0; y < 4; ++y) m[x][y] -= g[k][x] * u[k][y];
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) ld4g(bR + MB * k, MB / 4, tk, u[k]);                  // u[k][y]
#pragma unroll
                for (int k = 0; k < 4; ++k)
#pragma unroll
                    for (int x = 0; x < 4; ++x)
#pragma unroll
                        for (int y = 0; y < 4; ++y) m[x][y] -= g[k][x] * u[k][y];
            }
        }
        if (J == 2) BLK4_TS(6);
    }
    BLK4_TS(7);
    if (owner) {
#pragma unroll
        for (int x = 0; x < 4; ++x) {
            const int i = i0 + x;
            if (i < n) {
                if (!isB) {
                    if (vA && c0 + 3 <= i) st4(A + i * ld + c0, m[x]);
                    else {
#pragma unroll
                        for (int y = 0; y < 4; ++y) if (c0 + y <= i) A[i * ld + c0 + y] = m[x][y];
                    }
                } else {
                    if (vB && c0 + 3 < ncol) st4(B + i * ldb + c0, m[x]);
                    else {
#pragma unroll
                        for (int y = 0; y < 4; ++y) if (c0 + y < ncol) B[i * ldb + c0 + y] = m[x][y];
                    }
                }
            }
        }
    }
    __syncthreads();
    for (int j = tid; j < n; j += blockDim.x) {
        const T p = A[j * ld + j];
        if (!(p > (T)0)) *flag = 1;
        const T dd = sqrt(p);
        A[j * ld + j] = dd;
        invd[j] = (T)1 / dd;
    }
    __syncthreads();
    {
        const int lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
        for (int i = 1 + warp; i < n; i += nw)
            for (int k = lane; k < i; k += 32) A[i * ld + k] *= invd[k];
        if (MAXB > 0 && B != nullptr)
            for (int i = warp; i < n; i += nw) {
                const T s = invd[i];
                for (int c = lane; c < ncol; c += 32) B[i * ldb + c] *= s;
            }
    }
    __syncthreads();
    BLK4_TS(8);
    return *flag != 0;
}

// Solve L^T X = B in place (B [n][ldb], ncol <= MAXC columns, n <= MAXN): blocked back substitution, one 4 x 4 tile of B per thread in
// registers.  Big step K (descending) = rows 4K .. 4K + 3, ONE barrier: the tiles of row block K finish their rows against the
// diagonal block of L and publish them (double buffered); every tile above takes the rank-4 update m -= L[4K.., rows]^T X.
// Needs blockDim.x >= ceil(MAXN / 4) * ceil(MAXC / 4).  scr: 2 * 4 * 4 ceil(MAXC / 4) elements of T, 16-byte aligned.
template <typename T, int MAXN, int MAXC>
__device__ void blk4_trsm_lowerT_left(const T* __restrict__ L, int ldl, const T* __restrict__ invd, T* __restrict__ B, int ldb, int n,
                                      int ncol, T* __restrict__ scr) {
    constexpr int NT = (MAXN + 3) / 4, NC = (MAXC + 3) / 4, W = 4 * NC;
    const int tid = threadIdx.x;
    const bool owner = tid < NT * NC;
    const int ti = owner ? tid / NC : 0, tc = owner ? tid - ti * NC : 0, i0 = 4 * ti, c0 = 4 * tc;
    T m[4][4];
    __syncthreads();
    const bool vB = rows_vec16(B, ldb) && c0 + 3 < ncol;       // whole tile rows as 16-byte vectors (see blk4_cholesky_solve)
#pragma unroll
    for (int x = 0; x < 4; ++x) {
        T r[4] = {(T)0, (T)0, (T)0, (T)0};
        if (owner && i0 + x < n) {
            if (vB) ld4(B + (i0 + x) * ldb + c0, r);
            else {
#pragma unroll
                for (int y = 0; y < 4; ++y) if (c0 + y < ncol) r[y] = B[(i0 + x) * ldb + c0 + y];
            }
        }
#pragma unroll
        for (int y = 0; y < 4; ++y) m[x][y] = r[y];
    }
    const int nblk = (n + 3) >> 2;
    int par = 0;
    for (int K = nblk - 1; K >= 0; --K, par ^= 1) {
        const int k0 = 4 * K;
        T* X = scr + par * 4 * W;
        if (owner && ti == K) {
            // rows k0 + 3 .. k0: x_a = (b_a - sum_{b > a} L[k0 + b][k0 + a] x_b) / L[k0 + a][k0 + a]
#pragma unroll
            for (int a = 3; a >= 0; --a) {
                const bool va = k0 + a < n;
                const T ia = va ? invd[k0 + a] : (T)0;
#pragma unroll
                for (int b = 3; b > a; --b) {
                    const T l = (k0 + b < n) ? L[(k0 + b) * ldl + k0 + a] : (T)0;
#pragma unroll
                    for (int y = 0; y < 4; ++y) m[a][y] -= l * m[b][y];
                }
#pragma unroll
                for (int y = 0; y < 4; ++y) m[a][y] *= ia;
            }
#pragma unroll
            for (int x = 0; x < 4; ++x) st4g(X + W * x, NC, tc, m[x]);
        }
        if (K == 0) break;
        __syncthreads();
        if (owner && ti < K) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {              // one row of the block at a time: 8 live operand values next to the 16 of the tile
                T xr[4], l[4];
                ld4g(X + W * k, NC, tc, xr);
#pragma unroll
                for (int x = 0; x < 4; ++x) l[x] = (k0 + k < n) ? L[(k0 + k) * ldl + i0 + x] : (T)0;
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int y = 0; y < 4; ++y) m[x][y] -= l[x] * xr[y];
            }
        }
    }
    if (owner) {
#pragma unroll
        for (int x = 0; x < 4; ++x) {
            if (i0 + x < n) {
                if (vB) st4(B + (i0 + x) * ldb + c0, m[x]);
                else {
#pragma unroll
                    for (int y = 0; y < 4; ++y) if (c0 + y < ncol) B[(i0 + x) * ldb + c0 + y] = m[x][y];
                }
            }
        }
    }
    __syncthreads();
}

// C(i, j) = sum_k a_at(i, k) * b_at(k, j), i < M, j < N; 4 x 4 register tiles, rows contiguous (4 ti + x), columns strided (tj + TN y).
// KMODE 0: k in [0, K); 1: k in [4 ti, K) (operand A zero for k < i); 2: k in [0, min(K, 4 ti + 4)) (operand A zero for k > i).
// Operand reads are clamped to valid rows / columns, results outside (M, N) are dropped.  No barrier inside.
template <typename T, int KMODE, typename FA, typename FB, typename FE>
__device__ __forceinline__ void block_gemm(int M, int N, int K, FA a_at, FB b_at, FE store) {
    const int TN = (N + 3) >> 2, ntiles = ((M + 3) >> 2) * TN;
    for (int tile = threadIdx.x; tile < ntiles; tile += blockDim.x) {
        const int ti = tile / TN, tj = tile - ti * TN, i0 = ti * 4;
        int ii[4], jj[4];
#pragma unroll
        for (int x = 0; x < 4; ++x) { ii[x] = min(i0 + x, M - 1); jj[x] = min(tj + TN * x, N - 1); }
        T acc[4][4];
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
            for (int y = 0; y < 4; ++y) acc[x][y] = (T)0;
        const int kb = (KMODE == 1) ? i0 : 0, ke = (KMODE == 2) ? min(K, i0 + 4) : K;
#pragma unroll 2
        for (int k = kb; k < ke; ++k) {
            T av[4], bv[4];
#pragma unroll
            for (int x = 0; x < 4; ++x) { av[x] = a_at(ii[x], k); bv[x] = b_at(k, jj[x]); }
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y) acc[x][y] += av[x] * bv[y];
        }
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
            for (int y = 0; y < 4; ++y) {
                const int i = i0 + x, j = tj + TN * y;
                if (i < M && j < N) store(i, j, acc[x][y]);
            }
    }
}

// Adjoint of L = chol(A) in solve form: dA = sym(L^-T Phi(L^T dL) L^-1), Phi = lower triangle with halved diagonal.
// In: L (lower triangle read), invd, G = dL (lower triangle read).  Out: G = the SYMMETRIC gradient dA (full matrix).
// Wk: scratch [n][ld]; line: the exchange scratch of blk4_trsm_lowerT_left.
template <typename T, int MAXN>
__device__ void tile_cholesky_adjoint(const T* __restrict__ L, int ldl, const T* __restrict__ invd, T* __restrict__ G, T* __restrict__ Wk,
                                     int ld, int n, T* __restrict__ line) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __syncthreads();
    // Wk = Phi(L^T G):  Wk[i][j] = sum_{k >= i} L[k][i] G[k][j] for j <= i (the sum may start at 4 ti: L[k][i] = 0 is NOT guaranteed
    // above the diagonal, so the tile's own rows are masked in the operand)
    block_gemm<T, 1>(n, n, n,
                     [&](int i, int k) { return k >= i ? L[k * ldl + i] : (T)0; },
                     [&](int k, int j) { return k >= j ? G[k * ld + j] : (T)0; },
                     [&](int i, int j, T v) { Wk[i * ld + j] = (j < i) ? v : (j == i ? (T)0.5 * v : (T)0); });
    blk4_trsm_lowerT_left<T, MAXN, MAXN>(L, ldl, invd, Wk, ld, n, n, line);          // X = L^-T Phi
    for (int i = warp; i < n; i += nw)
        for (int j = lane; j < n; j += 32) G[i * ld + j] = Wk[j * ld + i];
    blk4_trsm_lowerT_left<T, MAXN, MAXN>(L, ldl, invd, G, ld, n, n, line);           // M^T = L^-T X^T
    for (int i = 1 + warp; i < n; i += nw)
        for (int j = lane; j < i; j += 32) {
            const T v = (T)0.5 * (G[i * ld + j] + G[j * ld + i]);
            G[i * ld + j] = v; G[j * ld + i] = v;
        }
    __syncthreads();
}

// Dispatch on the CTA width the general kernels are launched with (general_threads: 512 for n > 33, else 128).
template <typename T, bool WITHB>
__device__ __forceinline__ bool cta_cholesky_solve(T* A, int n, int ld, T* invd, T* B, int ldb, int ncol, T* scratch, int* flag) {
    if (blockDim.x >= 512 && n <= 65 && ncol <= 64) return blk4_cholesky_solve<T, 65, WITHB ? 64 : 0>(A, n, ld, invd, B, ldb, ncol, scratch, flag);
    if (blockDim.x >= 128 && n <= 33 && ncol <= 32) return blk4_cholesky_solve<T, 33, WITHB ? 32 : 0>(A, n, ld, invd, B, ldb, ncol, scratch, flag);
    return block_cholesky_solve<T>(A, n, ld, invd, B, ldb, ncol, flag);
}
template <typename T>
__device__ __forceinline__ void cta_trsm_lowerT_left(const T* L, int ldl, const T* invd, T* B, int ldb, int n, int ncol, T* scratch) {
    if (blockDim.x >= 512 && n <= 65 && ncol <= 65) return blk4_trsm_lowerT_left<T, 65, 65>(L, ldl, invd, B, ldb, n, ncol, scratch);
    if (blockDim.x >= 128 && n <= 33 && ncol <= 33) return blk4_trsm_lowerT_left<T, 33, 33>(L, ldl, invd, B, ldb, n, ncol, scratch);
    block_trsm_lowerT_left<T>(L, ldl, invd, B, ldb, n, ncol);
}
template <typename T>
__device__ __forceinline__ void cta_cholesky_adjoint(const T* L, int ldl, const T* invd, T* G, T* Wk, int ld, int n, T* scratch) {
    if (blockDim.x >= 512 && n <= 65) return tile_cholesky_adjoint<T, 65>(L, ldl, invd, G, Wk, ld, n, scratch);
    if (blockDim.x >= 128 && n <= 33) return tile_cholesky_adjoint<T, 33>(L, ldl, invd, G, Wk, ld, n, scratch);
    block_cholesky_adjoint<T>(L, ldl, invd, G, Wk, ld, n);
}

}  // namespace gp
}  // namespace clipgp
