// GP template weighter, adjoint: one CTA per class.  Hand-derived reverse pass of gp_forward.cu; the
// formulas are the ones of oracle/gp_manual.py::backward (checked there against torch.autograd).
//
//   B1  sparsemax adjoint per sample, dmu = sum_s df_s, dR = tril(sum_s df_s eps_s^T), dSigma = chol32 adjoint
//   B3  Sigma = K_XX + jI + Bm^T Bm - A^T A, Bm = Lq^T A, mu = A^T m + mean_x  ->  dA, dLq, dm (+ KL terms)
//   B4  A = L^-1 K_ZX in fp64: dK_ZX = L^-T dA, dL = -tril(dK_ZX A^T), dK_ZZ = chol64 adjoint
//   B5  kernel adjoint: d lengthscale / outputscale / variance and the learnable inducing row Z[n-1]
#include "gp_layout.cuh"

namespace clipgp {
namespace gp {

// One Gram block's contribution to the kernel adjoints.
//   dK   : [nA][ld] float (smem) upstream gradient of K(A rows, B rows)
//   raw  : [nA][ld] float scratch (receives r^2 / dot, then W = d loss / d raw)
//   q    : [d] accumulates sum_ij W_ij (u_ik - u_jk)^2      (rbf / matern; u = z * invls)
//   dzl  : [d] accumulates the gradient of one designated row (rowA as A-row and/or rowB as B-row), or -1
// Returns this thread's partial of d loss / d amp (outputscale or variance).
__device__ float kernel_adjoint_block(const float* dK, int ld, float* raw, const float* gA, int nA, const float* gB,
                                      int nB, int d, int kt, float amp, const float* invls, float* tileA, float* tileB,
                                      float* q, float* dzl, int rowA, int rowB) {
    gram_block<float>(nullptr, 0, raw, ld, gA, nA, gB, nB, d, kt, amp, invls, tileA, tileB);
    float damp = 0.f;
    for (int idx = threadIdx.x; idx < nA * nB; idx += blockDim.x) {
        const int i = idx / nB, j = idx - i * nB;
        const float r = raw[i * ld + j];
        const float g = dK[i * ld + j];
        float wv;
        if (kt == CLIPGP_KERNEL_RBF) {
            const float e = expf(-0.5f * r);
            damp += g * e;                         // dK/d os = exp(-r2/2)
            wv = -0.5f * g * amp * e;              // dK/d r2 = -K/2
        } else if (kt == CLIPGP_KERNEL_MATERN12) {
            const float rr = sqrtf(fmaxf(r, 1e-30f));
            wv = (r > 1e-30f) ? (-0.5f * g * expf(-rr) / rr) : 0.f;   // clamp_min(1e-30) kills the gradient
        } else {
            damp += g * r;                         // dK/d v = <a,b>
            wv = g * amp;                          // dK/d dot = v
        }
        raw[i * ld + j] = wv;
    }
    __syncthreads();
    const bool same = (gA == gB);
    const bool dot = (kt == CLIPGP_KERNEL_LINEAR);
    const int pA = pad4(nA), pB = pad4(nB);
    // thread -> (column k of the chunk, slice of the A rows)
    const int kk = threadIdx.x % KC, part = threadIdx.x / KC, nparts = blockDim.x / KC;
    for (int k0 = 0; k0 < d; k0 += KC) {
        __syncthreads();
        load_chunk(tileA, gA, nA, pA, d, k0, dot ? nullptr : invls);
        if (!same) load_chunk(tileB, gB, nB, pB, d, k0, dot ? nullptr : invls);
        __syncthreads();
        const float* tB = same ? tileA : tileB;
        const int k = k0 + kk;
        if (k < d && part < nparts) {
            if (!dot) {
                float qk = 0.f;
                for (int i = part; i < nA; i += nparts) {
                    const float ui = tileA[i * KCP + kk];
                    const float* wrow = raw + i * ld;
                    for (int j = 0; j < nB; ++j) {
                        const float df = ui - tB[j * KCP + kk];
                        qk = fmaf(wrow[j] * df, df, qk);
                    }
                }
                atomicAdd(&q[k], qk);
            }
            if (part == 0 && (rowA >= 0 || rowB >= 0)) {       // the one learnable row: O(nA + nB) per column
                float dz = 0.f;
                if (rowA >= 0) {
                    const float ui = tileA[rowA * KCP + kk];
                    for (int j = 0; j < nB; ++j) {
                        const float uj = tB[j * KCP + kk];
                        dz = fmaf(raw[rowA * ld + j], dot ? uj : (ui - uj), dz);
                    }
                }
                if (rowB >= 0) {
                    const float uj = tB[rowB * KCP + kk];
                    for (int i = 0; i < nA; ++i) {
                        const float ui = tileA[i * KCP + kk];
                        dz = fmaf(raw[i * ld + rowB], dot ? ui : (uj - ui), dz);
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

__global__ void __launch_bounds__(kThreads) gp_backward_kernel(const clipgp_gp_args a, const clipgp_gp_bwd_args b) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int c = blockIdx.x, tid = threadIdx.x;
    const int T = (int)a.T, n = (int)a.n, d = (int)a.d, S = (int)a.S;
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
            for (int idx = tid; idx < T * T; idx += blockDim.x) {
                const int j = idx / T, k = idx - j * T;
                if (k <= j) {
                    float s = 0.f;
                    for (int ss = 0; ss < sc; ++ss) s = fmaf(dfb[ss * ldt + j], ebuf[k * SCH + ss], s);
                    dSig[j * ldt + k] += s;
                }
            }
        }
        __syncthreads();
        cholesky_adjoint<float>(R, ldt, invdR, dSig, dSig, ldt, T, scrF);   // dSig <- dSigma (full, symmetric)
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
        for (int idx = tid; idx < n * T; idx += blockDim.x) {
            const int i = idx / T, j = idx - i * T;
            float s = 0.f;
            for (int k = i; k < n; ++k) s = fmaf(Lq[k * ldn + i], Af[k * ldt + j], s);
            Bm[i * ldt + j] = s;
        }
        __syncthreads();
        for (int idx = tid; idx < n * T; idx += blockDim.x) {      // dBm = 2 Bm dSigma
            const int i = idx / T, j = idx - i * T;
            float s = 0.f;
            for (int k = 0; k < T; ++k) s = fmaf(Bm[i * ldt + k], dSig[k * ldt + j], s);
            dBm[i * ldt + j] = 2.f * s;
        }
        __syncthreads();
        for (int idx = tid; idx < n * T; idx += blockDim.x) {      // dA = -2 A dSigma + Lq dBm + m dmu^T
            const int i = idx / T, j = idx - i * T;
            float s = 0.f;
            for (int k = 0; k < T; ++k) s = fmaf(Af[i * ldt + k], dSig[k * ldt + j], s);
            float s2 = 0.f;
            for (int k = 0; k <= i; ++k) s2 = fmaf(Lq[i * ldn + k], dBm[k * ldt + j], s2);
            const float v = -2.f * s + s2 + mvec[i] * dmu[j];
            dAf[i * ldt + j] = v;
            dAd[i * ldt + j] = (double)v;
        }
        for (int idx = tid; idx < n * n; idx += blockDim.x) {      // dLq = tril(A dBm^T) + dkl (Lq - diag(1/Lq_ii))
            const int i = idx / n, j = idx - i * n;
            float v = 0.f;
            if (j <= i) {
                for (int t = 0; t < T; ++t) v = fmaf(Af[i * ldt + t], dBm[j * ldt + t], v);
                v += dkl * (Lq[i * ldn + j] - (i == j ? 1.f / Lq[i * ldn + i] : 0.f));
            }
            b.dchol_var[(size_t)c * n * n + idx] = v;
        }
        for (int i = tid; i < n; i += blockDim.x) {                // dm = A dmu + dkl m
            float s = 0.f;
            for (int j = 0; j < T; ++j) s = fmaf(Af[i * ldt + j], dmu[j], s);
            b.dvar_mean[(size_t)c * n + i] = s + dkl * mvec[i];
        }
        if (b.dmean_x)
            for (int j = tid; j < T; j += blockDim.x) b.dmean_x[(size_t)c * T + j] = dmu[j];
        __syncthreads();
    }

    // =========================== B4 (fp64) ===========================
    float* dKzz = reinterpret_cast<float*>(smem + Y.p_dKzz);
    {
        double* scrD = reinterpret_cast<double*>(smem + Y.p_scrD);
        double* dLd = reinterpret_cast<double*>(smem + Y.p_dLd);
        trsm_lowerT_left<double>(Ld, ldn, invd, dAd, ldt, n, T);   // dAd <- dK_ZX = L^-T dA
        for (int idx = tid; idx < n * n; idx += blockDim.x) {      // dL = -tril(dK_ZX A^T)
            const int i = idx / n, j = idx - i * n;
            double s = 0.0;
            if (j <= i)
                for (int t = 0; t < T; ++t) s += dAd[i * ldt + t] * (double)Af[j * ldt + t];
            dLd[i * ldn + j] = -s;
        }
        __syncthreads();
        cholesky_adjoint<double>(Ld, ldn, invd, dLd, dLd, ldn, n, scrD);
        for (int idx = tid; idx < n * n; idx += blockDim.x) {
            const int i = idx / n, j = idx - i * n;
            dKzz[i * ldn + j] = (float)dLd[i * ldn + j];
        }
        __syncthreads();
    }

    // =========================== B5 ===========================
    {
        float* dKzx = reinterpret_cast<float*>(smem + Y.p_dKzx);   // [n][ldt]
        float* raw = reinterpret_cast<float*>(smem + Y.p_raw);
        float* tileA = reinterpret_cast<float*>(smem + Y.p_tiles);
        float* tileB = tileA + (size_t)pad4(n) * KCP;
        int alias = 0;
        if (a.x_is_z_prefix == 2) alias = 1;
        else if (a.x_is_z_prefix == 1) alias = rows_identical(Xc, Zc, T * d);
        float damp = 0.f;
        if (alias) {
            // one block: K(Z,Z) carries dK_ZZ + [dK_ZX | 0] + [[dK_XX, 0],[0, 0]]
            for (int idx = tid; idx < n * n; idx += blockDim.x) {
                const int i = idx / n, j = idx - i * n;
                float v = dKzz[i * ldn + j];
                if (j < T) v += (float)dAd[i * ldt + j];
                if (i < T && j < T) v += dSig[i * ldt + j];
                dKzz[i * ldn + j] = v;
            }
            __syncthreads();
            damp += kernel_adjoint_block(dKzz, ldn, raw, Zc, n, Zc, n, d, kt, amp, invls, tileA, tileB, qls, dzl, n - 1, n - 1);
        } else {
            for (int idx = tid; idx < n * T; idx += blockDim.x) {
                const int i = idx / T, j = idx - i * T;
                dKzx[i * ldt + j] = (float)dAd[i * ldt + j];
            }
            __syncthreads();
            damp += kernel_adjoint_block(dKzz, ldn, raw, Zc, n, Zc, n, d, kt, amp, invls, tileA, tileB, qls, dzl, n - 1, n - 1);
            damp += kernel_adjoint_block(dKzx, ldt, raw, Zc, n, Xc, T, d, kt, amp, invls, tileA, tileB, qls, dzl, n - 1, -1);
            damp += kernel_adjoint_block(dSig, ldt, raw, Xc, T, Xc, T, d, kt, amp, invls, tileA, tileB, qls, dzl, -1, -1);
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
}

}  // namespace gp
}  // namespace clipgp

using namespace clipgp;

extern "C" int clipgp_gp_backward(const clipgp_gp_args* a, const clipgp_gp_bwd_args* b, void* stream) {
    CLIPGP_REQUIRE(a && b, "gp_backward: NULL args");
    CLIPGP_REQUIRE(a->C >= 0 && a->T >= 1 && a->T <= CLIPGP_GP_MAX_T && a->n >= 1 && a->n <= CLIPGP_GP_MAX_T + 1 &&
                       a->d >= 1 && a->S >= 1, "gp_backward: unsupported shape");
    CLIPGP_REQUIRE(a->kernel_type >= 0 && a->kernel_type <= 2, "gp_backward: Unsupported kernel: %d", a->kernel_type);
    if (a->C == 0) return CLIPGP_OK;
    CLIPGP_REQUIRE(a->Z && a->X && a->var_mean && a->chol_var && a->w && a->L && a->A && a->R,
                   "gp_backward: forward tensors missing (Z, X, var_mean, chol_var, w, L, A, R)");
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
    gp::gp_backward_kernel<<<(unsigned)a->C, gp::kThreads, smem, (cudaStream_t)stream>>>(*a, *b);
    return check_launch("gp_backward_kernel");
}
