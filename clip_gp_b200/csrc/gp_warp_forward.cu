// GP template weighter, forward, warp-per-class register-resident fast path (T <= 32, n = T + 1, test inputs equal
// to the frozen inducing rows).  Same mathematics and outputs as gp_forward.cu (which remains the general path for
// T up to 64 and for un-aliased inputs); see gp_warp.cuh for the layout.  Saves only the kernel block K_ZZ for the
// adjoint (gp_warp_backward.cu recomputes the factorisations from it: ~10k instructions per class, cheaper than
// round-tripping L (fp64), A and R through HBM).
#include "gp_warp.cuh"

namespace clipgp {
namespace gpw {

constexpr int WPB = 2;                      // warps (= classes) per CTA
constexpr int TS = 36;                      // tile row stride in floats (16-byte aligned rows)
constexpr int TILE = 33 * TS;               // one [33][36] staging tile
constexpr int SMEM_PER_WARP = 2 * TILE;     // two tiles: Gram chunk / Lq staging, and the (Bm | A) exchange buffers

__device__ __forceinline__ float kval(int kt, float raw, float amp) {
    if (kt == CLIPGP_KERNEL_RBF) return amp * expf(-0.5f * raw);
    if (kt == CLIPGP_KERNEL_MATERN12) return expf(-sqrtf(fmaxf(raw, 1e-30f)));
    return amp * raw;
}

// Stream the n = T+1 inducing rows through the tile in 32-column chunks and accumulate, for this lane's row i < T,
// raw[j] = sum_k (z_ik - z_jk)^2 / l_k^2 (or sum_k z_ik z_jk) for j = 0..T (j = T is the token).  tt = raw(token, token).
__device__ __forceinline__ void gram_rows(const float* __restrict__ Zc, const float* __restrict__ raw_ls_c, int T, int d, int kt,
                                          float* __restrict__ tile, float* __restrict__ ils, float (&raw)[33], float& tt) {
    const int lane = threadIdx.x & 31;
    const bool dot = (kt == CLIPGP_KERNEL_LINEAR);
#pragma unroll
    for (int j = 0; j < 33; ++j) raw[j] = 0.f;
    float ttp = 0.f;
    for (int k0 = 0; k0 < d; k0 += 32) {
        __syncwarp();
        ils[lane] = (!dot && k0 + lane < d) ? 1.f / softplusf(__ldg(raw_ls_c + k0 + lane)) : 1.f;
        __syncwarp();
        float4 own[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int k = k0 + 4 * q;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (lane < T && k < d) v = __ldg(reinterpret_cast<const float4*>(Zc + (size_t)lane * d + k));
            const float4 s = *reinterpret_cast<const float4*>(ils + 4 * q);
            v.x *= s.x; v.y *= s.y; v.z *= s.z; v.w *= s.w;
            own[q] = v;
            if (lane < T) *reinterpret_cast<float4*>(tile + lane * TS + 4 * q) = v;
        }
        {   // the token row (index T) goes through lane q of the first 8 lanes, 16 bytes each
            if (lane < 8) {
                const int k = k0 + 4 * lane;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (k < d) v = __ldg(reinterpret_cast<const float4*>(Zc + (size_t)T * d + k));
                const float4 s = *reinterpret_cast<const float4*>(ils + 4 * lane);
                v.x *= s.x; v.y *= s.y; v.z *= s.z; v.w *= s.w;
                *reinterpret_cast<float4*>(tile + T * TS + 4 * lane) = v;
                if (dot) ttp += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
            }
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 33; ++j) {
            if (j <= T) {                                        // warp-uniform
                float s = raw[j];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float4 b = *reinterpret_cast<const float4*>(tile + j * TS + 4 * q);   // broadcast read
                    if (dot) {
                        s = fmaf(own[q].x, b.x, s); s = fmaf(own[q].y, b.y, s); s = fmaf(own[q].z, b.z, s); s = fmaf(own[q].w, b.w, s);
                    } else {
                        float t;
                        t = own[q].x - b.x; s = fmaf(t, t, s); t = own[q].y - b.y; s = fmaf(t, t, s);
                        t = own[q].z - b.z; s = fmaf(t, t, s); t = own[q].w - b.w; s = fmaf(t, t, s);
                    }
                }
                raw[j] = s;
            }
        }
    }
    tt = warp_sum(ttp);
    __syncwarp();
}

// Solve Lt X = B for 32 right-hand sides, lane j holding column j of B in x[] (in place).  Lt row-per-lane.
__device__ __forceinline__ void fwd_subst_cols(const double (&l)[TM], double inv_diag, double (&x)[TM]) {
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        double s = x[i];
#pragma unroll
        for (int m = 0; m < i; ++m) s -= bcast(l[m], i) * x[m];
        x[i] = s * bcast(inv_diag, i);
    }
}

// Everything from the kernel block to (mu, R) for one class.  Inputs: K row of this lane (kr[j] = Kt[lane][j], masked so
// that lanes / columns >= T form an identity block), kvec = K[lane][T], kappa.  Outputs in registers:
//   acol[0..31] column `lane` of A_T, acol[32] = A[T][lane];  bm[] likewise for Bm = Lq^T A;  mu;  r[] = row `lane` of
//   R = chol32(Sigma) (entries k > lane zeroed), sig[] = row of Sigma.  Returns status (0 ok, k retries, <0 failure).
struct Predictive {
    float acol[33], bm[33], sig[TM], r[TM], mu, inv_r;
    double lv, inv_lam;         // border of L: l = Lt^-1 k (one element per lane), 1/lambda
};

__device__ __forceinline__ int predictive(const clipgp_gp_args& a, int c, int T, const float (&kr)[TM], float kvec, float kappa,
                                          double (&l)[TM], double& invd, float* __restrict__ tileA, float* __restrict__ tileB,
                                          Predictive& P, float (&lq)[TM], float& qv, float& rho, float& m_lane, float& m_tok) {
    const int lane = threadIdx.x & 31;
    const int n = T + 1;
    // ---- L = chol64(Kt + 1e-4 I), border l = Lt^-1 k, lambda
#pragma unroll
    for (int j = 0; j < TM; ++j) l[j] = (double)(kr[j] + (j == lane && lane < T ? 1e-4f : 0.f));
    bool failL = chol_rows<double>(l, invd);
    P.lv = fwd_subst_vec<double>(l, invd, (double)kvec);
    const double lam2 = (double)(kappa + 1e-4f) - warp_sum_d(P.lv * P.lv);
    if (!(lam2 > 0.0)) failL = true;
    P.inv_lam = rsqrt(lam2);
    // ---- A = L^-1 K_ZX : columns of Kt (no jitter) are its rows
    {
        double x[TM];
#pragma unroll
        for (int i = 0; i < TM; ++i) x[i] = (double)kr[i];
        fwd_subst_cols(l, invd, x);
        double al = (double)kvec;
#pragma unroll
        for (int m = 0; m < TM; ++m) al -= bcast(P.lv, m) * x[m];
        al *= P.inv_lam;
#pragma unroll
        for (int i = 0; i < TM; ++i) P.acol[i] = (float)x[i];
        P.acol[32] = (lane < T) ? (float)al : 0.f;
    }
    // ---- variational parameters: m (one per lane + token), Lq rows via a coalesced staging copy
    m_lane = (lane < T) ? __ldg(a.var_mean + (size_t)c * n + lane) : 0.f;
    m_tok = __ldg(a.var_mean + (size_t)c * n + T);
    __syncwarp();
    for (int idx = lane; idx < n * n; idx += 32) {
        const int i = idx / n, j = idx - i * n;
        tileA[i * TS + j] = (j <= i) ? __ldg(a.chol_var + (size_t)c * n * n + idx) : 0.f;
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < TM; ++j) lq[j] = (lane < T && j <= lane) ? tileA[lane * TS + j] : 0.f;
    qv = (lane < T) ? tileA[T * TS + lane] : 0.f;            // Lq[T][lane]
    rho = tileA[T * TS + T];
    // ---- mu = A^T m + mean_x
    {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < TM; ++i) s = fmaf(P.acol[i], bcast(m_lane, i), s);
        s = fmaf(P.acol[32], m_tok, s);
        P.mu = s + ((a.mean_x && lane < T) ? __ldg(a.mean_x + (size_t)c * T + lane) : 0.f);
    }
    // ---- Bm = Lq^T A  (column per lane)
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        float s = bcast(qv, i) * P.acol[32];
#pragma unroll
        for (int k = i; k < TM; ++k) s = fmaf(bcast(lq[i], k), P.acol[k], s);
        P.bm[i] = s;
    }
    P.bm[32] = rho * P.acol[32];
    // ---- Sigma = Kt + 1e-4 I + Bm^T Bm - A^T A : exchange the columns through shared memory ([k][lane] layout)
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 33; ++k) { tileA[k * TS + lane] = P.bm[k]; tileB[k * TS + lane] = P.acol[k]; }
    __syncwarp();
#pragma unroll
    for (int i4 = 0; i4 < 8; ++i4) {
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
        for (int k = 0; k < 33; ++k) {
            const float4 b = *reinterpret_cast<const float4*>(tileA + k * TS + 4 * i4);    // Bm[k][i..i+3], broadcast
            const float4 aa = *reinterpret_cast<const float4*>(tileB + k * TS + 4 * i4);   // A[k][i..i+3]
            const float bk = P.bm[k], ak = P.acol[k];
            s0 += b.x * bk - aa.x * ak; s1 += b.y * bk - aa.y * ak; s2 += b.z * bk - aa.z * ak; s3 += b.w * bk - aa.w * ak;
        }
        P.sig[4 * i4 + 0] = s0; P.sig[4 * i4 + 1] = s1; P.sig[4 * i4 + 2] = s2; P.sig[4 * i4 + 3] = s3;
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const float base = (lane < T) ? kr[i] + ((i == lane) ? 1e-4f : 0.f) : ((i == lane) ? 1.f : 0.f);
        P.sig[i] = (lane < T && i < T) ? base + P.sig[i] : base;
    }
    // ---- R = chol32(Sigma), psd_safe_cholesky retries with total diagonal jitter 1e-6, 1e-5, 1e-4
    int retries = 0;
    bool failR = true;
#pragma unroll 1
    for (int attempt = 0; attempt < 4; ++attempt) {
        const float jit = attempt == 0 ? 0.f : (attempt == 1 ? 1e-6f : (attempt == 2 ? 1e-5f : 1e-4f));
#pragma unroll
        for (int i = 0; i < TM; ++i) P.r[i] = P.sig[i] + ((i == lane && lane < T) ? jit : 0.f);
        failR = chol_rows<float>(P.r, P.inv_r);
        if (!failR) break;
        ++retries;
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) P.r[i] = (i <= lane && lane < T) ? P.r[i] : 0.f;
    return failL ? -2 : (failR ? -1 : retries);
}

__global__ void __launch_bounds__(WPB * 32, 6) gp_forward_warp_kernel(const clipgp_gp_args a) {
    extern __shared__ __align__(16) float smw[];
    __shared__ __align__(16) float ils_all[WPB][32];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int c = blockIdx.x * WPB + wib;
    if (c >= a.C) return;
    float* tileA = smw + wib * SMEM_PER_WARP;
    float* tileB = tileA + TILE;
    float* ils = ils_all[wib];
    const int T = (int)a.T, n = T + 1, d = (int)a.d, S = (int)a.S, kt = a.kernel_type;
    const float* Zc = a.Z + (size_t)c * n * d;
    const float* ks = a.Ksave + (size_t)c * (1 + n * n + n * T + T * T);

    // ---- kernel block K_ZZ (no jitter) saved by the streamed Gram kernel; classes it flagged as un-aliased are already done
    if (ks[0] == 0.f) return;
    (void)Zc; (void)ils; (void)d;
    for (int idx = lane; idx < n * n; idx += 32) { const int i = idx / n, j = idx - i * n; tileA[i * TS + j] = __ldg(ks + 1 + idx); }
    __syncwarp();
    float kr[TM], kvec, kappa;
#pragma unroll
    for (int j = 0; j < TM; ++j) kr[j] = (lane < T && j < T) ? tileA[lane * TS + j] : ((lane >= T && j == lane) ? 1.f : 0.f);
    kvec = (lane < T) ? tileA[lane * TS + T] : 0.f;
    kappa = tileA[T * TS + T];
    __syncwarp();

    double l[TM], invd;
    float lq[TM], qv, rho, m_lane, m_tok;
    Predictive P;
    const int st = predictive(a, c, T, kr, kvec, kappa, l, invd, tileA, tileB, P, lq, qv, rho, m_lane, m_tok);
    if (lane == 0 && a.status) a.status[c] = st;

    // ---- optional saved factors (only needed when the general block adjoint consumes this forward pass)
    if (a.A) {
#pragma unroll
        for (int k = 0; k < TM; ++k) if (k < T && lane < T) a.A[(size_t)c * n * T + (size_t)k * T + lane] = P.acol[k];
        if (lane < T) a.A[(size_t)c * n * T + (size_t)T * T + lane] = P.acol[32];
    }
    if (a.L) {
        double* stg = reinterpret_cast<double*>(tileA);          // [33][33] doubles span both tiles
        __syncwarp();
#pragma unroll
        for (int j = 0; j < TM; ++j) if (lane < T) stg[lane * 33 + j] = (j <= lane) ? l[j] : 0.0;
        if (lane < T) { stg[T * 33 + lane] = P.lv; stg[lane * 33 + T] = 0.0; }
        if (lane == 0) stg[T * 33 + T] = 1.0 / P.inv_lam;
        __syncwarp();
        for (int idx = lane; idx < n * n; idx += 32) { const int i = idx / n, j = idx - i * n; a.L[(size_t)c * n * n + idx] = stg[i * 33 + j]; }
        __syncwarp();
    }
    if (a.R) {
        __syncwarp();
#pragma unroll
        for (int j = 0; j < TM; ++j) tileA[lane * TS + j] = P.r[j];
        __syncwarp();
        for (int idx = lane; idx < T * T; idx += 32) { const int i = idx / T, j = idx - i * T; a.R[(size_t)c * T * T + idx] = tileA[i * TS + j]; }
        __syncwarp();
    }

    // ---- KL(q(u) || N(0,I)) = 1/2 (|Lq|_F^2 + |m|^2 - n - sum log Lq_ii^2)
    if (a.kl) {
        float part = m_lane * m_lane + qv * qv;
#pragma unroll
        for (int j = 0; j < TM; ++j) part = fmaf(lq[j], lq[j], part);
        float ld = 0.f;
#pragma unroll
        for (int j = 0; j < TM; ++j) if (j == lane) ld = lq[j];
        if (lane < T) part -= logf(ld * ld);
        part = warp_sum(part);
        if (lane == 0) a.kl[c] = 0.5f * (part + m_tok * m_tok + rho * rho - logf(rho * rho) - (float)n);
    }

    // ---- f_s = mu + R eps_s ; w_s = sparsemax(f_s)
    uint64_t seed = 0, step = 0;
    if (a.eps == nullptr) { seed = a.rng_state[0]; step = a.rng_state[1]; }
    for (int s = 0; s < S; ++s) {
        float e = 0.f;
        if (lane < T) {
            if (a.eps) e = a.eps[(size_t)c * a.eps_sc + (size_t)lane * a.eps_st + (size_t)s * a.eps_ss];
            else e = philox_normal(seed, step, ((uint64_t)c * T + lane) * (uint64_t)a.S_total + (uint64_t)(a.s_offset + s));
        }
        float f = 0.f;
#pragma unroll
        for (int k = 0; k < TM; ++k) f = fmaf(P.r[k], bcast(e, k), f);
        f += P.mu;
        // sparsemax over the T lanes (sort-free: rank by value desc, index asc)
        const bool valid = lane < T;
        const float fm = warp_max(valid ? f : -INFINITY);
        const float z = f - fm;
        int kk = 0; float cs = 0.f;
#pragma unroll 8
        for (int j = 0; j < TM; ++j) {
            const float zj = bcast(z, j);
            const bool before = (j < T) && ((zj > z) || (zj == z && j <= lane));
            if (before) { ++kk; cs += zj; }
        }
        const bool sup = valid && ((float)kk * z > cs - 1.f);
        const int cnt = __popc(__ballot_sync(FULL, sup));
        const float tau = (warp_sum(sup ? z : 0.f) - 1.f) / (float)cnt;
        float wv = valid ? fmaxf(z - tau, 0.f) : 0.f;
        if (st < 0) wv = 0.f;
        if (valid) a.w[((size_t)s * a.C + c) * T + lane] = wv;
    }
}

}  // namespace gpw
}  // namespace clipgp

using namespace clipgp;

// Fast-path eligibility (the general kernel handles everything else).
extern "C" int clipgp_gp_warp_path_ok(int64_t T, int64_t n, int64_t d) {
    return (T >= 1 && T <= 32 && n == T + 1 && d >= 4 && (d % 4) == 0) ? 1 : 0;
}

int clipgp_gp_forward_warp_launch(const clipgp_gp_args* a, cudaStream_t st) {
    const size_t smem = sizeof(float) * gpw::WPB * gpw::SMEM_PER_WARP;
    const unsigned grid = (unsigned)((a->C + gpw::WPB - 1) / gpw::WPB);
    gpw::gp_forward_warp_kernel<<<grid, gpw::WPB * 32, smem, st>>>(*a);
    return check_launch("gp_forward_warp_kernel");
}
