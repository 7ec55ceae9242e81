// GP template weighter, adjoint: one CTA per class.  Hand-derived reverse pass of gp_forward.cu; the
// formulas are the ones of oracle/gp_manual.py::backward (checked there against torch.autograd).
//
//   B1  sparsemax adjoint per sample, dmu = sum_s df_s, dR = tril(sum_s df_s eps_s^T), dSigma = chol32 adjoint
//   B3  Sigma = K_XX + jI + Bm^T Bm - A^T A, Bm = Lq^T A, mu = A^T m + mean_x  ->  dA, dLq, dm (+ KL terms)
//   B4  A = L^-1 K_ZX in fp64: dK_ZX = L^-T dA, dL = -tril(dK_ZX A^T), dK_ZZ = chol64 adjoint
//   B5  kernel adjoint: d lengthscale / outputscale / variance and the learnable inducing row Z[n-1]
#include <stdlib.h>

#include "gp_layout.cuh"
#include "gp_block.cuh"

namespace clipgp {
namespace gp {

// Kernel-adjoint stage of the warp path (gp_warp_backward.cu): d loss / d K block (scratch part of the class record in
// Ksave) -> gradients of the length-scales, the output-scale / variance and the learnable inducing row, by one streamed pass
// over Z with the 4x4 register tiles of kernel_adjoint_block.  One CTA per (aliased) class.
__global__ void __launch_bounds__(kThreads) gp_kernel_adjoint_kernel(const clipgp_gp_args a, const clipgp_gp_bwd_args b) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int c = (int)a.c_begin + blockIdx.x, tid = threadIdx.x;
    const int T = (int)a.T, n = (int)a.n, d = (int)a.d;
    const int ldn = n | 1, dp = (d + 3) & ~3;
    const float* ks = a.Ksave + (size_t)c * (1 + n * n + n * T + T * T);
    if (ks[0] == 0.f) return;
    float* dK = reinterpret_cast<float*>(smem);
    float* raw = dK + n * ldn;
    float* invls = raw + n * ldn;
    float* qls = invls + dp;
    float* dzl = qls + dp;
    float* rs = dzl + dp;
    float* cs = rs + ((n + 3) & ~3);
    float* tileA = cs + ((n + 3) & ~3);
    __shared__ float red[32];
    const float* Zc = a.Z + (size_t)c * n * d;
    const int kt = a.kernel_type;
    if (kt != CLIPGP_KERNEL_LINEAR)
        for (int k = tid; k < d; k += blockDim.x) invls[k] = 1.f / softplusf(a.raw_lengthscale[(size_t)c * d + k]);
    float amp = 1.f;
    if (kt == CLIPGP_KERNEL_RBF) amp = softplusf(a.raw_outputscale[c]);
    if (kt == CLIPGP_KERNEL_LINEAR) amp = softplusf(a.raw_variance[c]);
    for (int k = tid; k < d; k += blockDim.x) { qls[k] = 0.f; dzl[k] = 0.f; }
    const float* dKt = ks + 1 + n * n;
    for (int idx = tid; idx < n * n; idx += blockDim.x) {
        const int i = idx / n, j = idx - i * n;
        dK[i * ldn + j] = dKt[idx];
        raw[i * ldn + j] = ks[1 + idx];
    }
    __syncthreads();
    const float damp = kernel_adjoint_block(dK, ldn, raw, Zc, n, Zc, n, d, kt, amp, invls, tileA, tileA, qls, dzl, n - 1, n - 1, rs, cs);
    const float damp_tot = block_sum(damp, red);
    if (tid == 0) {
        if (kt == CLIPGP_KERNEL_RBF && b.draw_outputscale) b.draw_outputscale[c] = damp_tot * sigmoidf_(a.raw_outputscale[c]);
        if (kt == CLIPGP_KERNEL_LINEAR && b.draw_variance) b.draw_variance[c] = damp_tot * sigmoidf_(a.raw_variance[c]);
    }
    __syncthreads();
    for (int k = tid; k < d; k += blockDim.x) {
        if (kt != CLIPGP_KERNEL_LINEAR && b.draw_lengthscale)
            b.draw_lengthscale[(size_t)c * d + k] = -2.f * qls[k] * invls[k] * sigmoidf_(a.raw_lengthscale[(size_t)c * d + k]);
        if (b.dZ_last) b.dZ_last[(size_t)c * d + k] = dzl[k];
    }
}

#ifdef CLIPGP_PHASE_TS
__device__ long long g_gen_ts_bwd[32];
#define GENB_TS(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_gen_ts_bwd[i] = clock64(); } while (0)
#else
#define GENB_TS(i) do { } while (0)
#endif

// only_unaliased != 0: classes served by the warp path (alias flag set) are skipped.
__global__ void __launch_bounds__(kThreadsMax) gp_backward_kernel(const clipgp_gp_args a, const clipgp_gp_bwd_args b, const int only_unaliased) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int c = (int)a.c_begin + blockIdx.x, tid = threadIdx.x;
    const int T = (int)a.T, n = (int)a.n, d = (int)a.d, S = (int)a.S;
    if (only_unaliased && a.Ksave[(size_t)c * (1 + n * n + n * T + T * T)] != 0.f) return;
    const Dims D = make_dims(T, n, d);
    const BwdLayout Y = make_bwd_layout(D);
    const int ldn = D.ldn, ldt = D.ldt;
    double* Ld = reinterpret_cast<double*>(smem + Y.Ld);
    double* invd = reinterpret_cast<double*>(smem + Y.invd);
    double* dAd = reinterpret_cast<double*>(smem + Y.dAd);
    float* Af = reinterpret_cast<float*>(smem + Y.Af);
    float* dSig = reinterpret_cast<float*>(smem + Y.dSig);
    float* mvec = reinterpret_cast<float*>(smem + Y.mvec);
    float* dmu = reinterpret_cast<float*>(smem + Y.dmu);
    float* invls = reinterpret_cast<float*>(smem + Y.invls);
    float* invdR = reinterpret_cast<float*>(smem + Y.invdR);
    float* qls = reinterpret_cast<float*>(smem + Y.dls);
    float* dzl = reinterpret_cast<float*>(smem + Y.dzl);
    __shared__ float red[32];

    const float* Zc = a.Z + (size_t)c * n * d;
    const float* Xc = a.X + (size_t)c * T * d;
    const int kt = a.kernel_type;
    const float dkl = b.dkl ? b.dkl[c] : b.dkl_scalar;

    GENB_TS(0);
    // ---- load persistent state
    if (kt != CLIPGP_KERNEL_LINEAR)
        for (int k = tid; k < d; k += blockDim.x) invls[k] = 1.f / softplusf(a.raw_lengthscale[(size_t)c * d + k]);
    float amp = 1.f;
    if (kt == CLIPGP_KERNEL_RBF) amp = softplusf(a.raw_outputscale[c]);
    if (kt == CLIPGP_KERNEL_LINEAR) amp = softplusf(a.raw_variance[c]);
    for (int k = tid; k < d; k += blockDim.x) { qls[k] = 0.f; dzl[k] = 0.f; }
    for (int idx = tid; idx < n * n; idx += blockDim.x) {
        const int i = idx / n, j = idx - i * n;
        const double v = a.L[(size_t)c * n * n + idx];
        Ld[i * ldn + j] = v;
        if (i == j) invd[i] = 1.0 / v;
    }
    for (int idx = tid; idx < n * T; idx += blockDim.x) {
        const int i = idx / T, j = idx - i * T;
        Af[i * ldt + j] = a.A[(size_t)c * n * T + idx];
    }
    for (int i = tid; i < n; i += blockDim.x) mvec[i] = a.var_mean[(size_t)c * n + i];
    for (int j = tid; j < T; j += blockDim.x) dmu[j] = 0.f;

    GENB_TS(1);
    // =========================== B1 ===========================
    {
        float* R = reinterpret_cast<float*>(smem + Y.p_R);
        float* scrF = reinterpret_cast<float*>(smem + Y.p_scrF);
        float* dfb = reinterpret_cast<float*>(smem + Y.p_df);    // [SCH][ldt]
        float* ebuf = reinterpret_cast<float*>(smem + Y.p_eps);  // [T][SCH]
        for (int idx = tid; idx < T * T; idx += blockDim.x) {
            const int i = idx / T, j = idx - i * T;
            const float v = a.R[(size_t)c * T * T + idx];
            R[i * ldt + j] = v;
            if (i == j) invdR[i] = 1.f / v;
            dSig[i * ldt + j] = 0.f;
        }
        uint64_t seed = 0, step = 0;
        if (a.eps == nullptr) { seed = a.rng_state[0]; step = a.rng_state[1]; }
        const int lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
        for (int s0 = 0; s0 < S; s0 += SCH) {
            const int sc = min(SCH, S - s0);
            __syncthreads();
            for (int idx = tid; idx < T * sc; idx += blockDim.x) {
                const int t = idx / sc, ss = idx - t * sc;
                float e;
                if (a.eps) e = a.eps[(size_t)c * a.eps_sc + (size_t)t * a.eps_st + (size_t)(s0 + ss) * a.eps_ss];
                else e = philox_normal(seed, step, ((uint64_t)c * T + t) * (uint64_t)a.S_total + (uint64_t)(a.s_offset + s0 + ss));
                ebuf[t * SCH + ss] = e;
            }
            // sparsemax adjoint: df = [w>0] (dw - mean_{support} dw)
            for (int ss = warp; ss < sc; ss += nwarps) {
                const size_t off = ((size_t)(s0 + ss) * a.C + c) * T;
                const float w0 = lane < T ? a.w[off + lane] : 0.f;
                const float w1 = lane + 32 < T ? a.w[off + lane + 32] : 0.f;
                const float g0 = (w0 > 0.f) ? b.dw[off + lane] : 0.f;
                const float g1 = (w1 > 0.f) ? b.dw[off + lane + 32] : 0.f;
                const int cnt = __popc(__ballot_sync(0xffffffffu, w0 > 0.f)) + __popc(__ballot_sync(0xffffffffu, w1 > 0.f));
                const float vhat = warp_sum(g0 + g1) / (float)max(cnt, 1);
                if (lane < T) dfb[ss * ldt + lane] = (w0 > 0.f) ? g0 - vhat : 0.f;
                if (lane + 32 < T) dfb[ss * ldt + lane + 32] = (w1 > 0.f) ? g1 - vhat : 0.f;
            }
            __syncthreads();
            for (int j = tid; j < T; j += blockDim.x) {
                float s = 0.f;
                for (int ss = 0; ss < sc; ++ss) s += dfb[ss * ldt + j];
                dmu[j] += s;
            }
            block_gemm<float, 0>(T, T, sc, [&](int j, int ss) { return dfb[ss * ldt + j]; }, [&](int ss, int k) { return ebuf[k * SCH + ss]; },
                                 [&](int j, int k, float v) { if (k <= j) dSig[j * ldt + k] += v; });
        }
        __syncthreads();
        GENB_TS(2);
        cta_cholesky_adjoint<float>(R, ldt, invdR, dSig, scrF, ldt, T, reinterpret_cast<float*>(smem + Y.line));     // dSig <- dSigma (full, symmetric)
        GENB_TS(3);
    }

    // =========================== B3 ===========================
    {
        float* Lq = reinterpret_cast<float*>(smem + Y.p_Lq);
        float* Bm = reinterpret_cast<float*>(smem + Y.p_Bm);
        float* dBm = reinterpret_cast<float*>(smem + Y.p_dBm);
        float* dAf = reinterpret_cast<float*>(smem + Y.p_dAf);
        for (int idx = tid; idx < n * n; idx += blockDim.x) {
            const int i = idx / n, j = idx - i * n;
            Lq[i * ldn + j] = (j <= i) ? a.chol_var[(size_t)c * n * n + idx] : 0.f;
        }
        __syncthreads();
        // Bm = Lq^T A (Lq is stored with a zero upper triangle)
        block_gemm<float, 1>(n, T, n, [&](int i, int k) { return Lq[k * ldn + i]; }, [&](int k, int j) { return Af[k * ldt + j]; },
                             [&](int i, int j, float v) { Bm[i * ldt + j] = v; });
        __syncthreads();
        // dBm = 2 Bm dSigma
        block_gemm<float, 0>(n, T, T, [&](int i, int k) { return Bm[i * ldt + k]; }, [&](int k, int j) { return dSig[k * ldt + j]; },
                             [&](int i, int j, float v) { dBm[i * ldt + j] = 2.f * v; });
        // dA = -2 A dSigma + m dmu^T  (+ Lq dBm below)
        block_gemm<float, 0>(n, T, T, [&](int i, int k) { return Af[i * ldt + k]; }, [&](int k, int j) { return dSig[k * ldt + j]; },
                             [&](int i, int j, float v) { dAf[i * ldt + j] = -2.f * v + mvec[i] * dmu[j]; });
        __syncthreads();
        block_gemm<float, 2>(n, T, n, [&](int i, int k) { return Lq[i * ldn + k]; }, [&](int k, int j) { return dBm[k * ldt + j]; },
                             [&](int i, int j, float v) { const float t = dAf[i * ldt + j] + v; dAf[i * ldt + j] = t; dAd[i * ldt + j] = (double)t; });
        // dLq = tril(A dBm^T) + dkl (Lq - diag(1/Lq_ii))
        block_gemm<float, 0>(n, n, T, [&](int i, int t) { return Af[i * ldt + t]; }, [&](int t, int j) { return dBm[j * ldt + t]; },
                             [&](int i, int j, float v) {
                                 b.dchol_var[(size_t)c * n * n + (size_t)i * n + j] =
                                     (j <= i) ? v + dkl * (Lq[i * ldn + j] - (i == j ? 1.f / Lq[i * ldn + i] : 0.f)) : 0.f;
                             });
        for (int i = tid; i < n; i += blockDim.x) {                // dm = A dmu + dkl m
            float s = 0.f;
            for (int j = 0; j < T; ++j) s = fmaf(Af[i * ldt + j], dmu[j], s);
            b.dvar_mean[(size_t)c * n + i] = s + dkl * mvec[i];
        }
        if (b.dmean_x)
            for (int j = tid; j < T; j += blockDim.x) b.dmean_x[(size_t)c * T + j] = dmu[j];
        __syncthreads();
    }

    GENB_TS(4);
    // =========================== B4 (fp64) ===========================
    float* dKzz = reinterpret_cast<float*>(smem + Y.p_dKzz);
    {
        double* scrD = reinterpret_cast<double*>(smem + Y.p_scrD);
        double* dLd = reinterpret_cast<double*>(smem + Y.p_dLd);
        cta_trsm_lowerT_left<double>(Ld, ldn, invd, dAd, ldt, n, T, reinterpret_cast<double*>(smem + Y.line));   // dAd <- dK_ZX = L^-T dA
        GENB_TS(5);
        // dL = -tril(dK_ZX A^T)
        block_gemm<double, 0>(n, n, T, [&](int i, int t) { return dAd[i * ldt + t]; }, [&](int t, int j) { return (double)Af[j * ldt + t]; },
                              [&](int i, int j, double v) { dLd[i * ldn + j] = (j <= i) ? -v : 0.0; });
        __syncthreads();
        GENB_TS(6);
        cta_cholesky_adjoint<double>(Ld, ldn, invd, dLd, scrD, ldn, n, reinterpret_cast<double*>(smem + Y.line));    // dLd <- dK_ZZ (full, symmetric)
        GENB_TS(7);
        for (int idx = tid; idx < n * n; idx += blockDim.x) {
            const int i = idx / n, j = idx - i * n;
            dKzz[i * ldn + j] = (float)dLd[i * ldn + j];
        }
        __syncthreads();
    }

    GENB_TS(8);
    // =========================== B5 ===========================
    {
        float* dKzx = reinterpret_cast<float*>(smem + Y.p_dKzx);   // [n][ldt]
        float* raw = reinterpret_cast<float*>(smem + Y.p_raw);
        float* tileA = reinterpret_cast<float*>(smem + Y.p_tiles);
        float* tileB = tileA + (size_t)pad4(n) * KCP;
        const float* ks = a.Ksave + (size_t)c * (1 + n * n + n * T + T * T);
        const int alias = ks[0] != 0.f;
        float* rs = reinterpret_cast<float*>(smem + Y.rs);
        float* cs = reinterpret_cast<float*>(smem + Y.cs);
        float damp = 0.f;
        if (alias) {
            // one block: K(Z,Z) carries dK_ZZ + [dK_ZX | 0] + [[dK_XX, 0],[0, 0]]
            for (int idx = tid; idx < n * n; idx += blockDim.x) {
                const int i = idx / n, j = idx - i * n;
                float v = dKzz[i * ldn + j];
                if (j < T) v += (float)dAd[i * ldt + j];
                if (i < T && j < T) v += dSig[i * ldt + j];
                dKzz[i * ldn + j] = v;
                raw[i * ldn + j] = ks[1 + idx];
            }
            __syncthreads();
            damp += kernel_adjoint_block(dKzz, ldn, raw, Zc, n, Zc, n, d, kt, amp, invls, tileA, tileB, qls, dzl, n - 1, n - 1, rs, cs);
        } else {
            for (int idx = tid; idx < n * T; idx += blockDim.x) {
                const int i = idx / T, j = idx - i * T;
                dKzx[i * ldt + j] = (float)dAd[i * ldt + j];
            }
            for (int idx = tid; idx < n * n; idx += blockDim.x) { const int i = idx / n, j = idx - i * n; raw[i * ldn + j] = ks[1 + idx]; }
            __syncthreads();
            damp += kernel_adjoint_block(dKzz, ldn, raw, Zc, n, Zc, n, d, kt, amp, invls, tileA, tileB, qls, dzl, n - 1, n - 1, rs, cs);
            for (int idx = tid; idx < n * T; idx += blockDim.x) { const int i = idx / T, j = idx - i * T; raw[i * ldt + j] = ks[1 + n * n + idx]; }
            __syncthreads();
            damp += kernel_adjoint_block(dKzx, ldt, raw, Zc, n, Xc, T, d, kt, amp, invls, tileA, tileB, qls, dzl, n - 1, -1, rs, cs);
            for (int idx = tid; idx < T * T; idx += blockDim.x) { const int i = idx / T, j = idx - i * T; raw[i * ldt + j] = ks[1 + n * n + n * T + idx]; }
            __syncthreads();
            damp += kernel_adjoint_block(dSig, ldt, raw, Xc, T, Xc, T, d, kt, amp, invls, tileA, tileB, qls, dzl, -1, -1, rs, cs);
        }
        const float damp_tot = block_sum(damp, red);
        if (tid == 0) {
            if (kt == CLIPGP_KERNEL_RBF && b.draw_outputscale) b.draw_outputscale[c] = damp_tot * sigmoidf_(a.raw_outputscale[c]);
            if (kt == CLIPGP_KERNEL_LINEAR && b.draw_variance) b.draw_variance[c] = damp_tot * sigmoidf_(a.raw_variance[c]);
        }
        __syncthreads();
        for (int k = tid; k < d; k += blockDim.x) {
            if (kt != CLIPGP_KERNEL_LINEAR && b.draw_lengthscale) {
                // r2 = sum_k (z_ik - z_jk)^2 / l_k^2  ->  d r2 / d l_k = -2 (u_ik-u_jk)^2 / l_k ;  d l / d raw = sigmoid(raw)
                const float rawls = a.raw_lengthscale[(size_t)c * d + k];
                b.draw_lengthscale[(size_t)c * d + k] = -2.f * qls[k] * invls[k] * sigmoidf_(rawls);
            }
            if (b.dZ_last) b.dZ_last[(size_t)c * d + k] = dzl[k];
        }
    }
    GENB_TS(9);
}

}  // namespace gp
}  // namespace clipgp

using namespace clipgp;

#ifdef CLIPGP_PHASE_TS
extern "C" int clipgp_debug_general_ts_bwd(long long* out) { return (int)cudaMemcpyFromSymbol(out, gp::g_gen_ts_bwd, sizeof(long long) * 32); }
#endif

int clipgp_gp_backward_warp_launch(const clipgp_gp_args* a, const clipgp_gp_bwd_args* b, cudaStream_t st, int fuse);   // gp_warp_backward.cu
extern "C" int clipgp_gp_warp_fused_adjoint_ok(int64_t n, int64_t d);

extern "C" int clipgp_gp_backward(const clipgp_gp_args* a, const clipgp_gp_bwd_args* b, void* stream) {
    CLIPGP_REQUIRE(a && b, "gp_backward: NULL args");
    CLIPGP_REQUIRE(a->C >= 0 && a->T >= 1 && a->T <= CLIPGP_GP_MAX_T && a->n >= 1 && a->n <= CLIPGP_GP_MAX_T + 1 &&
                       a->d >= 1 && a->S >= 1, "gp_backward: unsupported shape");
    CLIPGP_REQUIRE(a->kernel_type >= 0 && a->kernel_type <= 2, "gp_backward: Unsupported kernel: %d", a->kernel_type);
    CLIPGP_REQUIRE(a->c_begin >= 0 && a->c_count >= 0 && a->c_begin + a->c_count <= a->C, "gp_backward: class shard outside [0, C)");
    if (a->C == 0 || gp_grid(a) == 0) return CLIPGP_OK;
    CLIPGP_REQUIRE(a->Z && a->X && a->var_mean && a->chol_var && a->w && a->L && a->A && a->R && a->Ksave,
                   "gp_backward: forward tensors missing (Z, X, var_mean, chol_var, w, L, A, R, Ksave)");
    CLIPGP_REQUIRE(a->eps || a->rng_state, "gp_backward: need eps or rng_state");
    CLIPGP_REQUIRE(b->dw && b->dvar_mean && b->dchol_var, "gp_backward: dw / dvar_mean / dchol_var is NULL");
    if (a->kernel_type != CLIPGP_KERNEL_LINEAR) CLIPGP_REQUIRE(a->raw_lengthscale, "gp_backward: raw_lengthscale is NULL");
    if (a->kernel_type == CLIPGP_KERNEL_RBF) CLIPGP_REQUIRE(a->raw_outputscale, "gp_backward: raw_outputscale is NULL");
    if (a->kernel_type == CLIPGP_KERNEL_LINEAR) CLIPGP_REQUIRE(a->raw_variance, "gp_backward: raw_variance is NULL");
    const size_t smem = (size_t)clipgp_gp_smem_bytes(a->T, a->n, a->d, 1);
    CLIPGP_REQUIRE(smem > 0 && smem <= 227 * 1024, "gp_backward: needs %zu bytes of shared memory (> 227 KB); reduce d", smem);
    static size_t smem_set = 0;
    if (smem > smem_set) {
        CLIPGP_CUDA(cudaFuncSetAttribute(gp::gp_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set = smem;
    }
    // warp path (T <= 32, aliased test inputs): dense algebra warp-per-class, then the streamed kernel adjoint; classes whose
    // inputs are not aliased (x_is_z_prefix == 1 and the forward pass's device check failed) go through the block kernel
    static const bool block_only = (getenv("CLIPGP_GP_BLOCK_ONLY") != nullptr);
    if (b->tl_Z != nullptr) {
        CLIPGP_REQUIRE(b->tl_dlT && b->proto_EEt && b->proto_norm && b->tl_B >= 1 && a->S >= 1 && a->S <= 12 && (b->tl_mode == 0 || b->tl_mode == 1) &&
                       b->tl_Z_ld >= a->C * a->T && b->tl_dlT_ld >= b->tl_B && clipgp_gp_warp_path_ok(a->T, a->n, a->d) && a->x_is_z_prefix != 0,
                       "gp_backward: template-logit adjoint needs the warp path, S <= 12, dlogits^T / EEt / norms and consistent strides");
    }
    const bool warp_path = !block_only && clipgp_gp_warp_path_ok(a->T, a->n, a->d) && a->x_is_z_prefix != 0 &&
                           ((reinterpret_cast<uintptr_t>(a->Z) & 15u) == 0);
    if (warp_path) {
        int rc;
        if (a->x_is_z_prefix != 2) {
            gp::gp_backward_kernel<<<gp_grid(a), gp::kThreads, smem, (cudaStream_t)stream>>>(*a, *b, 1);
            rc = check_launch("gp_backward_kernel(unaliased)");
            if (rc != CLIPGP_OK) return rc;
        }
        // the kernel adjoint runs inside the same CTA when its buffers fit the algebra kernel's shared memory (d <= ~1000)
        static const bool no_fuse = (getenv("CLIPGP_GP_NO_FUSED_ADJOINT") != nullptr);
        const int fuse = (!no_fuse && clipgp_gp_warp_fused_adjoint_ok(a->n, a->d)) ? 1 : 0;
        rc = clipgp_gp_backward_warp_launch(a, b, (cudaStream_t)stream, fuse);
        if (rc != CLIPGP_OK || fuse) return rc;
        const int n = (int)a->n, d = (int)a->d, ldn = n | 1, dp = (d + 3) & ~3, np = (n + 3) & ~3;
        const size_t sm2 = sizeof(float) * ((size_t)2 * n * ldn + 3 * dp + 2 * np + (size_t)gp::pad4(n) * gp::KCP) + 16;
        static size_t sm2_set = 48 * 1024;
        if (sm2 > sm2_set) {
            CLIPGP_CUDA(cudaFuncSetAttribute(gp::gp_kernel_adjoint_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));
            sm2_set = sm2;
        }
        gp::gp_kernel_adjoint_kernel<<<gp_grid(a), gp::kThreads, sm2, (cudaStream_t)stream>>>(*a, *b);
        return check_launch("gp_kernel_adjoint_kernel");
    }
    gp::gp_backward_kernel<<<gp_grid(a), gp::general_threads(a->n), smem, (cudaStream_t)stream>>>(*a, *b, 0);
    return check_launch("gp_backward_kernel");
}
