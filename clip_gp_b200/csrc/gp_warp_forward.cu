// GP template weighter, forward, warp-per-class path (T <= 32, n = T + 1, test inputs equal to the frozen inducing
// rows).  Same mathematics and outputs as gp_forward.cu (which remains the general path for T up to 64 and for
// un-aliased inputs): starts from the kernel block K_ZZ that the streamed Gram kernel saved in Ksave, and performs
//   L = chol64(K_ZZ + 1e-4 I), A = L^-1 K_ZX, mu = A^T m + mean_x, Bm = Lq^T A, Sigma = K_XX + 1e-4 I + Bm^T Bm - A^T A,
//   R = chol32(Sigma) (psd_safe jitter retries), KL(q(u) || N(0,I)), f_s = mu + R eps_s, w_s = sparsemax(f_s)
// (gp_template_weigher.py:166-173,194-219 + gpytorch whitened VariationalStrategy.forward, rsample, entmax.sparsemax).
// One 4-warp CTA per class, matrices in shared memory, run-time loops (see gp_warp.cuh).
#include "gp_warp.cuh"

namespace clipgp {
namespace gpw {

struct FwdSmem {
    double Ld[NN];       // K_ZZ + jI -> L (fp64)
    double Ad[NN];       // K_ZX -> A (fp64); afterwards reused as float Sigma [T][LD]
    double invd[34];
    float Af[NN];        // A (fp32) [n][LD], lane = test point
    float Bm[NN];        // Lq^T A
    float Lq[NN];        // tril(chol_var); afterwards reused as R = chol32(Sigma)
    float mvec[36];
    float invdR[32];
    float mu[32];
    int flag[4];
};

// fuse_gram != 0 (the caller guarantees aliased test inputs, x_is_z_prefix == 2): the streamed Gram block K_ZZ is computed by this
// CTA itself (gp::gram_block, one 4x4 tile per thread) instead of being read from the Gram kernel's hand-over record.
__global__ void __launch_bounds__(NT, 7) gp_forward_warp_kernel(const clipgp_gp_args a, const int fuse_gram) {
    extern __shared__ __align__(16) unsigned char smw[];
    FwdSmem& s = *reinterpret_cast<FwdSmem*>(smw);
    const int lane = lane_id(), wid = warp_id(), tid = threadIdx.x, c = (int)a.c_begin + blockIdx.x;
    const int T = (int)a.T, n = T + 1, S = (int)a.S;
    float* ks = a.Ksave + (size_t)c * ksave_stride(n, T);
    if (!fuse_gram && ks[0] == 0.f) return;        // un-aliased class: finished by the block kernel (uniform per CTA)
    const float* K = ks + 1;
    const int lt = lane < T ? lane : T - 1;

    if (fuse_gram) {
        // ---- K_ZZ by this CTA: scratch in regions that are written only later (K0: Af, inverse length-scales: Bm, chunk tile: Ad)
        const int d = (int)a.d, kt = a.kernel_type;
        float* K0 = s.Af;
        float* invls = s.Bm;
        float* tile = reinterpret_cast<float*>(s.Ad);
        float amp = 1.f;
        if (kt == CLIPGP_KERNEL_RBF) amp = softplusf(a.raw_outputscale[c]);
        if (kt == CLIPGP_KERNEL_LINEAR) amp = softplusf(a.raw_variance[c]);
        if (kt != CLIPGP_KERNEL_LINEAR)
            for (int k = tid; k < d; k += NT) invls[k] = 1.f / softplusf(a.raw_lengthscale[(size_t)c * d + k]);
        __syncthreads();
        const float* Zc = a.Z + (size_t)c * n * d;
        gp::gram_block<float, 1>(K0, LD, nullptr, 0, Zc, n, Zc, n, d, kt, amp, invls, tile, tile);
        if (tid == 0) ks[0] = 1.f;
        // hand-over record for the adjoint (and the K_XX read of the Sigma stage below) + the fp64 operands
        each_block(n, n, [&](int idx, int i, int j) {
            const float v = K0[i * LD + j];
            ks[1 + idx] = v;
            s.Ld[i * LD + j] = (double)(v + (i == j ? 1e-4f : 0.f));
            if (j < T) s.Ad[i * LD + j] = (double)v;
        });
    } else {
        // ---- stage K_ZZ: jitter is added in fp32 before the cast, as gpytorch does (add_jitter, then .double())
        stage_block<float>(n, n, [&](int idx) { return __ldg(K + idx); }, [&](int idx, int i, int j, float v) {
            s.Ld[i * LD + j] = (double)(v + (i == j ? 1e-4f : 0.f));
            if (j < T) s.Ad[i * LD + j] = (double)v;
        });
    }
    {
        const float* cv = a.chol_var + (size_t)c * n * n;
        stage_block<float>(n, n, [&](int idx) { return __ldg(cv + idx); },
                           [&](int idx, int i, int j, float v) { s.Lq[i * LD + j] = (j <= i) ? v : 0.f; });
    }
    for (int i = tid; i < n; i += NT) s.mvec[i] = __ldg(a.var_mean + (size_t)c * n + i);
    if (tid < 3) s.Lq[33 * LD + tid] = 0.f;
    __syncthreads();

    // ---- warp 0: L = chol64(K_ZZ + 1e-4 I), A = L^-1 K_ZX;  warp 1 meanwhile: KL(q(u) || N(0,I))
    if (wid == 0) {
        const bool f = chol33<double>(s.Ld, n, s.invd);
        if (lane == 0) s.flag[0] = f ? 1 : 0;
        __syncwarp();
        trsm_lower_cols<double>(s.Ld, s.invd, s.Ad, n, T);
    } else if (wid == 1 && a.kl) {
        float part = 0.f;                          // 1/2 (|Lq|_F^2 + |m|^2 - n - sum log Lq_ii^2)
        for (int i = lane; i < n; i += 32) {
            const float* row = s.Lq + i * LD;
            float q = 0.f;
            for (int j = 0; j <= i; ++j) q = fmaf(row[j], row[j], q);
            part += q - logf(row[i] * row[i]) + s.mvec[i] * s.mvec[i];
        }
        part = warp_sum(part);
        if (lane == 0) a.kl[c] = 0.5f * (part - (float)n);
    }
    __syncthreads();
    const bool failL = s.flag[0] != 0;
    each_block(n, T, [&](int idx, int i, int j) { s.Af[i * LD + j] = (float)s.Ad[i * LD + j]; });
    __syncthreads();

    // ---- mu = A^T m + mean_x (warp 3, lane = test point)
    if (wid == 3 && lane < T) {
        float mu = 0.f;
        for (int i = 0; i < n; ++i) mu = fmaf(s.Af[i * LD + lane], s.mvec[i], mu);
        if (a.mean_x) mu += __ldg(a.mean_x + (size_t)c * T + lane);
        s.mu[lane] = mu;
    }
    // ---- Bm = Lq^T A : Bm[i][t] = sum_{k >= i} Lq[k][i] A[k][t]  (Lq is zero above the diagonal: no k >= i test needed)
    for (int i0 = 4 * wid; i0 < n; i0 += 4 * NW) {
        float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
        for (int k = i0; k < n; ++k) {
            const float av = s.Af[k * LD + lt];
            const float* lq = s.Lq + k * LD + i0;
            acc0 = fmaf(lq[0], av, acc0); acc1 = fmaf(lq[1], av, acc1); acc2 = fmaf(lq[2], av, acc2); acc3 = fmaf(lq[3], av, acc3);
        }
        if (lane < T) {
            s.Bm[i0 * LD + lane] = acc0;
            if (i0 + 1 < n) s.Bm[(i0 + 1) * LD + lane] = acc1;
            if (i0 + 2 < n) s.Bm[(i0 + 2) * LD + lane] = acc2;
            if (i0 + 3 < n) s.Bm[(i0 + 3) * LD + lane] = acc3;
        }
    }
    __syncthreads();

    // ---- Sigma = K_XX + 1e-4 I + Bm^T Bm - A^T A (lower triangle), lane = column u, four rows t per pass
    float* Sig = reinterpret_cast<float*>(s.Ad);
    for (int t0 = 4 * wid; t0 < T; t0 += 4 * NW) {
        float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
        for (int k = 0; k < n; ++k) {
            const float bu = s.Bm[k * LD + lt], au = s.Af[k * LD + lt];
            const float* bt = s.Bm + k * LD + t0;
            const float* at = s.Af + k * LD + t0;
            acc0 += bt[0] * bu - at[0] * au; acc1 += bt[1] * bu - at[1] * au;
            acc2 += bt[2] * bu - at[2] * au; acc3 += bt[3] * bu - at[3] * au;
        }
        const float accs[4] = {acc0, acc1, acc2, acc3};
#pragma unroll
        for (int x = 0; x < 4; ++x) {
            const int t = t0 + x;
            if (t < T && lane <= t) Sig[t * LD + lane] = (K[t * n + lane] + (t == lane ? 1e-4f : 0.f)) + accs[x];   // coherent load: see fuse_gram
        }
    }
    __syncthreads();

    // ---- R = chol32(Sigma), psd_safe_cholesky retries with total diagonal jitter 1e-6, 1e-5, 1e-4
    float* R = s.Lq;                                // Lq is dead (its zeros above the diagonal stay in place)
    int retries = 0;
    bool failR = true;
    for (int attempt = 0; attempt < 4; ++attempt) {
        const float jit = attempt == 0 ? 0.f : (attempt == 1 ? 1e-6f : (attempt == 2 ? 1e-5f : 1e-4f));
        each_block(T, T, [&](int idx, int t, int u) { if (u <= t) R[t * LD + u] = Sig[t * LD + u] + (t == u ? jit : 0.f); });
        __syncthreads();
        if (wid == 0) {
            const bool f = chol33<float>(R, T, s.invdR);
            if (lane == 0) s.flag[1] = f ? 1 : 0;
        }
        __syncthreads();
        failR = s.flag[1] != 0;
        if (!failR) break;
        ++retries;
        __syncthreads();                            // everyone has read the flag before the next attempt rewrites it
    }
    const int st = failL ? -2 : (failR ? -1 : retries);
    if (tid == 0 && a.status) a.status[c] = st;

    // ---- saved factors for the adjoint
    if (a.L) each_block(n, n, [&](int idx, int i, int j) { a.L[(size_t)c * n * n + idx] = (j <= i) ? s.Ld[i * LD + j] : 0.0; });
    if (a.A) each_block(n, T, [&](int idx, int i, int j) { a.A[(size_t)c * n * T + idx] = s.Af[i * LD + j]; });
    if (a.R) each_block(T, T, [&](int idx, int i, int j) { a.R[(size_t)c * T * T + idx] = (j <= i) ? R[i * LD + j] : 0.f; });

    // ---- f_s = mu + R eps_s ; w_s = sparsemax(f_s)   (one sample per warp at a time, lane = template)
    uint64_t seed = 0, step = 0;
    if (a.eps == nullptr) { seed = a.rng_state[0]; step = a.rng_state[1]; }
    const float* Rrow = R + lt * LD;
    const float mu = s.mu[lt];
    for (int sidx = wid; sidx < S; sidx += NW) {
        float e = 0.f;
        if (lane < T) {
            if (a.eps) e = a.eps[(size_t)c * a.eps_sc + (size_t)lane * a.eps_st + (size_t)sidx * a.eps_ss];
            else {
                e = philox_normal(seed, step, ((uint64_t)c * T + lane) * (uint64_t)a.S_total + (uint64_t)(a.s_offset + sidx));
                if (a.eps_save) a.eps_save[((size_t)sidx * a.C + c) * T + lane] = e;
            }
        }
        float f0 = 0.f, f1 = 0.f;
        int k = 0;
        for (; k + 1 < T; k += 2) {
            f0 = fmaf(Rrow[k], __shfl_sync(FULL, e, k), f0);
            f1 = fmaf(Rrow[k + 1], __shfl_sync(FULL, e, k + 1), f1);
        }
        if (k < T) f0 = fmaf(Rrow[k], __shfl_sync(FULL, e, k), f0);
        float wv = sparsemax_lanes(f0 + f1 + mu, T);
        if (st < 0) wv = 0.f;
        if (lane < T) a.w[((size_t)sidx * a.C + c) * T + lane] = wv;
    }
}

}  // namespace gpw
}  // namespace clipgp

using namespace clipgp;

// Fast-path eligibility (the general kernel handles everything else).
extern "C" int clipgp_gp_warp_path_ok(int64_t T, int64_t n, int64_t d) {
    return (T >= 2 && T <= 32 && n == T + 1 && d >= 4 && (d % 4) == 0) ? 1 : 0;
}

int clipgp_gp_forward_warp_launch(const clipgp_gp_args* a, cudaStream_t st, int fuse_gram) {
    static bool attr_set = false;
    if (!attr_set) {
        CLIPGP_CUDA(cudaFuncSetAttribute(gpw::gp_forward_warp_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                         cudaSharedmemCarveoutMaxShared));
        attr_set = true;
    }
    gpw::gp_forward_warp_kernel<<<gp_grid(a), gpw::NT, sizeof(gpw::FwdSmem), st>>>(*a, fuse_gram);
    return check_launch("gp_forward_warp_kernel");
}
