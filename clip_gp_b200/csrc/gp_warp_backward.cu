// GP template weighter, adjoint, warp-per-class dense-algebra part (T <= 32, n = T + 1, aliased test inputs).
// Formulas: oracle/gp_manual.py::backward (validated there against torch.autograd); same results as the block kernel
// of gp_backward.cu, which remains the general path.  Per class, from the saved L (fp64), A, R and w / dw / eps:
//
//   P1  sparsemax adjoint per sample -> dmu, dR = tril(sum_s df_s eps_s^T) -> dSigma (chol32 adjoint, Murray reverse sweep)
//   P2  H = A dSigma, dBm = 2 Lq^T H, dA = -2 H + Lq dBm + m dmu^T, dLq = tril(A dBm^T) + KL term, dm = A dmu + KL term
//       (Bm = Lq^T A is never needed: dBm = 2 Bm dSigma = 2 Lq^T (A dSigma))
//   P3  fp64: dK_ZX = L^-T dA, dL = -tril(dK_ZX A^T), dK_ZZ = chol64 adjoint
//   out d loss / d K_ZZ-block = dK_ZZ + [dK_ZX | 0] + [[dSigma, 0], [0, 0]]  (aliased inputs: one kernel block carries all three)
//       written to the scratch part of the class record in Ksave; gp_kernel_adjoint_kernel (gp_backward.cu) turns it into the
//       length-scale / output-scale / variance / learnable-row gradients with one streamed pass over Z.
//
// One 4-warp CTA per class; shared memory: three fp64 [33][33] regions re-used across the phases + A (fp32) = 31 KB,
// 7 classes per SM.
#include "gp_warp.cuh"

extern "C" int clipgp_gp_warp_path_ok(int64_t T, int64_t n, int64_t d);

#ifndef KADJ_V
#define KADJ_V 1         // kernel adjoint with 16-byte operand loads (stride-36 chunk tile + transposed W); 0: gp::kernel_adjoint_block
#endif
#ifndef SPARSE_W
#define SPARSE_W 1       // skip a[s][t] = <dP_s, E_t> where the sparsemax weight w[s][t] is zero
#endif
#ifndef BLK4_ADJ
#define BLK4_ADJ 1      // blocked whole-CTA back substitutions inside the two Cholesky adjoints (0: the one-warp sweeps)
#endif

namespace clipgp {
namespace gpw {

#ifdef CLIPGP_PHASE_TS
__device__ long long g_phase_ts_bwd[64];
#define GPB_TS(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_phase_ts_bwd[i] = clock64(); } while (0)
#else
#define GPB_TS(i) do { } while (0)
#endif

struct BwdSmem {
    double RA[NN];       // P2..P3: dA -> dK_ZX (fp64)
    double RB[NN];       // P1: R | dR->dSigma (fp32 halves);  P2: dBm | dSigma;  P3: L (fp64)
    double RC[NN];       // P2: Lq | H (fp32 halves);  P3: dL -> chol adjoint (fp64)
    double invd[34];
    float Af[NN];
    float mvec[36];
    float vecT[32];      // P1: 1 / R_ii;  P2: dmu
};

__global__ void __launch_bounds__(NT, 7) gp_backward_warp_kernel(const clipgp_gp_args a, const clipgp_gp_bwd_args b,
                                                                 const int fuse_kernel_adjoint) {
    extern __shared__ __align__(16) unsigned char smw[];
    BwdSmem& s = *reinterpret_cast<BwdSmem*>(smw);
    const int lane = lane_id(), wid = warp_id(), tid = threadIdx.x, c = (int)a.c_begin + blockIdx.x;
    const int T = (int)a.T, n = T + 1, S = (int)a.S;
    float* ks = a.Ksave + (size_t)c * ksave_stride(n, T);
    if (ks[0] == 0.f) return;                      // un-aliased class: handled by the block kernel (uniform per CTA)
    float* dKt = ks + 1 + n * n;                   // scratch: d loss / d K block, [n][n]
    const float dkl = b.dkl ? b.dkl[c] : b.dkl_scalar;
    const int lt = lane < T ? lane : T - 1;        // clamped lane for row-per-lane walks

    GPB_TS(0);
    // =========================== P0 (optional): prototype adjoint -> dw of this class ===========================
    // dwsm [S][32] lives at the end of RC until P2 stages Lq there; it replaces the global dw read of P1.
    float* dwsm = nullptr;
    float* abuf = reinterpret_cast<float*>(s.RA) + 3 * 2 * NN - 2 * S * 32;   // [S][32]  a[s][t]  (the last 2 S 32 floats of RA | RB | RC)
    float* Gc = s.Af;                                                         // [T][T] = E[c] E[c]^T (Af is dead until P2)
    if (b.tl_Z != nullptr) {
        // ---- small-batch form: a[s][t] = scale * sum_b dlogits[b,s,c] Zt[b,c,t] from the per-template cosines (see clipgp.h)
        constexpr int BC = 128, SMAX = 12;
        float* dls = reinterpret_cast<float*>(s.RA);                  // [S][BC]   dlogits of this class, one batch chunk
        float* Zs = dls + SMAX * BC;                                  // [BC][32]  per-template cosines of the chunk
        dwsm = abuf + S * 32;
        for (int idx = tid; idx < T * T; idx += NT) Gc[idx] = __ldg(b.proto_EEt + (size_t)c * T * T + idx);
        const __nv_bfloat16* dlT = reinterpret_cast<const __nv_bfloat16*>(b.tl_dlT);
        const int Bt = (int)b.tl_B;
        float acc[3] = {0.f, 0.f, 0.f};                               // samples wid, wid + 4, wid + 8 (S <= 12), lane = template
        for (int b0 = 0; b0 < Bt; b0 += BC) {
            const int bc = min(BC, Bt - b0);
            __syncthreads();
            for (int idx = tid; idx < S * BC; idx += NT) {
                const int sidx = idx / BC, bb = idx - sidx * BC;
                float v = 0.f;
                if (bb < bc) {
                    const __nv_bfloat16* row = dlT + ((size_t)sidx * a.C + c) * b.tl_dlT_ld + b0 + bb;
                    v = __bfloat162float(row[0]);
                    if (b.tl_mode == 1) v += __bfloat162float(row[2 * b.tl_seg]);
                }
                dls[idx] = v;
            }
            for (int idx = tid; idx < BC * 32; idx += NT) {
                const int bb = idx >> 5, t = idx & 31;
                Zs[idx] = (bb < bc && t < T) ? __ldg(b.tl_Z + (size_t)(b0 + bb) * b.tl_Z_ld + (size_t)c * T + t) : 0.f;
            }
            __syncthreads();
#pragma unroll 2
            for (int bb = 0; bb < BC; bb += 4) {
                const float z0 = Zs[bb * 32 + lane], z1 = Zs[(bb + 1) * 32 + lane], z2 = Zs[(bb + 2) * 32 + lane], z3 = Zs[(bb + 3) * 32 + lane];
#pragma unroll
                for (int u = 0; u < 3; ++u) {
                    const int sidx = wid + 4 * u;
                    if (sidx < S) {
                        const float4 dv = *reinterpret_cast<const float4*>(dls + sidx * BC + bb);
                        acc[u] = fmaf(dv.x, z0, fmaf(dv.y, z1, fmaf(dv.z, z2, fmaf(dv.w, z3, acc[u]))));
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 3; ++u) {
            const int sidx = wid + 4 * u;
            if (sidx < S) abuf[sidx * 32 + lane] = b.tl_scale * acc[u];
        }
        __syncthreads();
    } else if (b.proto_dP != nullptr) {
        const int D = (int)b.proto_D, D4 = D >> 2;
        float* gbuf = reinterpret_cast<float*>(s.RA);                 // [S][D] upstream gradient rows (RA | RB | start of RC)
        dwsm = abuf + S * 32;                                         // [S][32]  (last S*32 floats of RC)
#if SPARSE_W
        float wpre[12];                                               // this lane's template weight in every sample: in flight while dP is staged
#pragma unroll
        for (int u = 0; u < 12; ++u) wpre[u] = (u < S && lane < T) ? __ldg(a.w + ((size_t)u * a.C + c) * T + lane) : 0.f;
#endif
        for (int sidx = 0; sidx < S; ++sidx) {
            const float4* src = reinterpret_cast<const float4*>(b.proto_dP + (size_t)sidx * b.proto_dP_stride_s + (size_t)c * D);
            for (int col = tid; col < D4; col += NT) {
                float4 v = __ldg(src + col);
                v.x *= b.proto_dP_scale; v.y *= b.proto_dP_scale; v.z *= b.proto_dP_scale; v.w *= b.proto_dP_scale;
                reinterpret_cast<float4*>(gbuf + (size_t)sidx * D)[col] = v;
            }
        }
        for (int idx = tid; idx < T * T; idx += NT) Gc[idx] = __ldg(b.proto_EEt + (size_t)c * T * T + idx);
        __syncthreads();
        GPB_TS(1);
        // a[s][t] = <g_s, E[c,t,:]>: warp = template row (the next row's loads in flight), lanes over the 16-byte column groups
        const float4* Ec = reinterpret_cast<const float4*>(b.proto_E + (size_t)c * T * D);
        constexpr int PSB = 12;
        // sparsemax: a[s][t] is consumed only where w[s][t] > 0 (q_s = <w_s, a_s>, and the sparsemax adjoint masks dw by the support), and
        // ~60 % of the weights are exact zeros: lane t keeps the S-bit mask of the samples that need a[., t]
        unsigned need = 0xFFFu;
#if SPARSE_W
        need = 0u;
#pragma unroll
        for (int u = 0; u < 12; ++u)
            if (wpre[u] > 0.f) need |= 1u << u;
#endif
        if (D4 <= 128) {
            // D <= 512: a lane owns at most four column groups of a row; the next row's loads are issued before this row is consumed
            float4 nxt[4];
            auto fetch = [&](int t) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int col = lane + 32 * u;
                    nxt[u] = (t < T && col < D4) ? __ldg(Ec + (size_t)t * D4 + col) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            };
            fetch(wid);
            for (int t = wid; t < T; t += NW) {
                float4 e[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) e[u] = nxt[u];
                fetch(t + NW);
                const unsigned nt = __shfl_sync(FULL, need, t);      // samples whose support holds template t (warp-uniform)
                float part[PSB];
#pragma unroll
                for (int u = 0; u < PSB; ++u) part[u] = 0.f;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int col = lane + 32 * q;
                    if (col < D4) {
#pragma unroll
                        for (int u = 0; u < PSB; ++u) {
                            if (u < S && ((nt >> u) & 1u)) {
                                const float4 v = reinterpret_cast<const float4*>(gbuf + (size_t)u * D)[col];
                                part[u] += e[q].x * v.x + e[q].y * v.y + e[q].z * v.z + e[q].w * v.w;
                            }
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < PSB; ++u) {
                    if (u < S) {
                        float tot = 0.f;
                        if ((nt >> u) & 1u) tot = warp_sum(part[u]);
                        if (lane == 0) abuf[u * 32 + t] = tot;
                    }
                }
            }
        } else {
            for (int t = wid; t < T; t += NW) {
                float part[PSB];
#pragma unroll
                for (int u = 0; u < PSB; ++u) part[u] = 0.f;
                for (int col = lane; col < D4; col += 32) {
                    const float4 e = __ldg(Ec + (size_t)t * D4 + col);
#pragma unroll
                    for (int u = 0; u < PSB; ++u) {
                        if (u < S) {
                            const float4 v = reinterpret_cast<const float4*>(gbuf + (size_t)u * D)[col];
                            part[u] += e.x * v.x + e.y * v.y + e.z * v.z + e.w * v.w;
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < PSB; ++u) {
                    if (u < S) {
                        const float tot = warp_sum(part[u]);
                        if (lane == 0) abuf[u * 32 + t] = tot;
                    }
                }
            }
        }
        __syncthreads();
    }
    if (dwsm != nullptr) {
        GPB_TS(2);
        // dw[s][t] = (a[s][t] - q_s (w_s G)[t] / |P_s|) / |P_s|,  q_s = <w_s, a_s> / |P_s|   (one sample per warp, lane = template)
        for (int sidx = wid; sidx < S; sidx += NW) {
            const size_t off = ((size_t)sidx * a.C + c) * T;
            const float wv = lane < T ? a.w[off + lane] : 0.f;
            const float av = lane < T ? abuf[sidx * 32 + lane] : 0.f;
            const float invn = 1.f / fmaxf(__ldg(b.proto_norm + (size_t)sidx * a.C + c), 1e-12f);
            const float q = warp_sum(wv * av) * invn;
            float wg = 0.f;
            for (int k = 0; k < T; ++k) wg = fmaf(__shfl_sync(FULL, wv, k), Gc[k * T + lt], wg);
            const float dwv = (av - q * wg * invn) * invn;
            if (lane < T) {
                dwsm[sidx * 32 + lane] = dwv;
                if (b.dw_out) b.dw_out[off + lane] = dwv;
            }
        }
        __syncthreads();
    }

    GPB_TS(3);
    // =========================== P1 ===========================
    float* R = reinterpret_cast<float*>(s.RB);
    float* G = R + NN;
    {
        const float* Rg = a.R + (size_t)c * T * T;
        stage_block<float>(T, T, [&](int idx) { return __ldg(Rg + idx); },
                           [&](int idx, int i, int j, float v) { R[i * LD + j] = v; G[i * LD + j] = 0.f; });
    }
    __syncthreads();
    if (tid < T) s.vecT[tid] = 1.f / R[tid * LD + tid];
    float dmu = 0.f;
    {
        // every warp walks all samples (the sparsemax adjoint is a few instructions) and accumulates its own eight columns of dR
        uint64_t seed = 0, step = 0;
        if (a.eps == nullptr) { seed = a.rng_state[0]; step = a.rng_state[1]; }
        float* Grow = G + lt * LD;
        const int kq0 = 8 * wid, kq1 = min(T, kq0 + 8);
        for (int sidx = 0; sidx < S; ++sidx) {
            const size_t off = ((size_t)sidx * a.C + c) * T;
            const float wv = lane < T ? a.w[off + lane] : 0.f;
            const bool sup = wv > 0.f;
            const float g = sup ? (dwsm ? dwsm[sidx * 32 + lane] : b.dw[off + lane]) : 0.f;
            const int cnt = __popc(__ballot_sync(FULL, sup));
            const float vhat = warp_sum(g) / (float)max(cnt, 1);
            const float df = sup ? g - vhat : 0.f;              // entmax sparsemax backward
            dmu += df;
            float e = 0.f;
            if (lane < T) {
                if (a.eps) e = a.eps[(size_t)c * a.eps_sc + (size_t)lane * a.eps_st + (size_t)sidx * a.eps_ss];
                else if (a.eps_save) e = a.eps_save[off + lane];           // drawn and stored by the forward kernel
                else e = philox_normal(seed, step, ((uint64_t)c * T + lane) * (uint64_t)a.S_total + (uint64_t)(a.s_offset + sidx));
            }
            if (cnt == 0) continue;                             // warp-uniform
            for (int k = kq0; k < kq1; ++k) {                   // dR[lane][k] += df_lane eps_k (upper part is never read)
                const float ek = __shfl_sync(FULL, e, k);
                if (lane < T) Grow[k] = fmaf(df, ek, Grow[k]);
            }
        }
    }
    GPB_TS(4);
    chol_adj_block<float>(R, s.vecT, G, reinterpret_cast<float*>(s.RC), T,       // RC is free until P2 stages Lq there; its second half
                          BLK4_ADJ ? reinterpret_cast<float*>(s.RC) + NN : nullptr);   // (16-byte aligned) is the exchange scratch
    each_block(T, T, [&](int idx, int i, int j) {               // G <- dSigma, full symmetric
        if (i > j) { const float v = 0.5f * G[i * LD + j]; G[i * LD + j] = v; G[j * LD + i] = v; }
    });
    __syncthreads();
    each_block(n, n, [&](int idx, int i, int j) { dKt[idx] = (i < T && j < T) ? G[i * LD + j] : 0.f; });
    if (wid == 0) {
        if (b.dmean_x && lane < T) b.dmean_x[(size_t)c * T + lane] = dmu;
        if (lane < T) s.vecT[lane] = dmu;                       // 1 / R_ii is dead
    }

    GPB_TS(5);
    // =========================== P2 ===========================
    float* Lq = reinterpret_cast<float*>(s.RC);
    float* H = Lq + NN;
    float* dBm = R;                                             // R is dead
    {
        const float* Ag = a.A + (size_t)c * n * T;
        const float* cv = a.chol_var + (size_t)c * n * n;
        stage_block<float>(n, T, [&](int idx) { return __ldg(Ag + idx); }, [&](int idx, int i, int j, float v) { s.Af[i * LD + j] = v; });
        stage_block<float>(n, n, [&](int idx) { return __ldg(cv + idx); }, [&](int idx, int i, int j, float v) { Lq[i * LD + j] = (j <= i) ? v : 0.f; });
    }
    for (int i = tid; i < n; i += NT) s.mvec[i] = __ldg(a.var_mean + (size_t)c * n + i);
    if (tid < 3) Lq[33 * LD + tid] = 0.f;
    __syncthreads();
    // H = A dSigma  (lane = column t)
    for (int i0 = 4 * wid; i0 < n; i0 += 4 * NW) {
        const float* a0 = s.Af + i0 * LD;
        const float* a1 = s.Af + min(i0 + 1, n - 1) * LD;
        const float* a2 = s.Af + min(i0 + 2, n - 1) * LD;
        const float* a3 = s.Af + min(i0 + 3, n - 1) * LD;
        float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
        for (int u = 0; u < T; ++u) {
            const float ds = G[u * LD + lt];
            acc0 = fmaf(a0[u], ds, acc0); acc1 = fmaf(a1[u], ds, acc1); acc2 = fmaf(a2[u], ds, acc2); acc3 = fmaf(a3[u], ds, acc3);
        }
        if (lane < T) {
            H[i0 * LD + lane] = acc0;
            if (i0 + 1 < n) H[(i0 + 1) * LD + lane] = acc1;
            if (i0 + 2 < n) H[(i0 + 2) * LD + lane] = acc2;
            if (i0 + 3 < n) H[(i0 + 3) * LD + lane] = acc3;
        }
    }
    __syncthreads();
    // dBm = 2 Lq^T H : dBm[i][t] = 2 sum_{k >= i} Lq[k][i] H[k][t]  (zeros above the diagonal of Lq make k >= i implicit)
    for (int i0 = 4 * wid; i0 < n; i0 += 4 * NW) {
        float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
        for (int k = i0; k < n; ++k) {
            const float hv = H[k * LD + lt];
            const float* lq = Lq + k * LD + i0;
            acc0 = fmaf(lq[0], hv, acc0); acc1 = fmaf(lq[1], hv, acc1); acc2 = fmaf(lq[2], hv, acc2); acc3 = fmaf(lq[3], hv, acc3);
        }
        if (lane < T) {
            dBm[i0 * LD + lane] = 2.f * acc0;
            if (i0 + 1 < n) dBm[(i0 + 1) * LD + lane] = 2.f * acc1;
            if (i0 + 2 < n) dBm[(i0 + 2) * LD + lane] = 2.f * acc2;
            if (i0 + 3 < n) dBm[(i0 + 3) * LD + lane] = 2.f * acc3;
        }
    }
    __syncthreads();
    // dA = -2 H + Lq dBm + m dmu^T  (lane = column t) -> fp64
    {
        const float dmu_t = s.vecT[lt];
        for (int i0 = 4 * wid; i0 < n; i0 += 4 * NW) {
            const float* l0 = Lq + i0 * LD;
            const float* l1 = Lq + min(i0 + 1, n - 1) * LD;
            const float* l2 = Lq + min(i0 + 2, n - 1) * LD;
            const float* l3 = Lq + min(i0 + 3, n - 1) * LD;
            const int kmax = min(i0 + 3, n - 1);
            float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
            for (int k = 0; k <= kmax; ++k) {
                const float bv = dBm[k * LD + lt];
                acc0 = fmaf(l0[k], bv, acc0); acc1 = fmaf(l1[k], bv, acc1); acc2 = fmaf(l2[k], bv, acc2); acc3 = fmaf(l3[k], bv, acc3);
            }
            const float accs[4] = {acc0, acc1, acc2, acc3};
            if (lane < T) {
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                    const int i = i0 + x;
                    if (i < n) s.RA[i * LD + lane] = (double)(-2.f * H[i * LD + lane] + accs[x] + s.mvec[i] * dmu_t);
                }
            }
        }
    }
    // dLq = tril(A dBm^T) + dkl (Lq - diag(1 / Lq_ii))  (lane = column j < 32: row j of dBm), written straight to HBM
    {
        const float* brow = dBm + min(lane, n - 1) * LD;
        float* out = b.dchol_var + (size_t)c * n * n;
        for (int i0 = 4 * wid; i0 < n; i0 += 4 * NW) {
            const float* a0 = s.Af + i0 * LD;
            const float* a1 = s.Af + min(i0 + 1, n - 1) * LD;
            const float* a2 = s.Af + min(i0 + 2, n - 1) * LD;
            const float* a3 = s.Af + min(i0 + 3, n - 1) * LD;
            float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
            for (int t = 0; t < T; ++t) {
                const float bv = brow[t];
                acc0 = fmaf(a0[t], bv, acc0); acc1 = fmaf(a1[t], bv, acc1); acc2 = fmaf(a2[t], bv, acc2); acc3 = fmaf(a3[t], bv, acc3);
            }
            const float accs[4] = {acc0, acc1, acc2, acc3};
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                const int i = i0 + x;
                if (i < n && lane < n) {
                    float v = 0.f;
                    if (lane <= i) v = accs[x] + dkl * (Lq[i * LD + lane] - (i == lane ? 1.f / Lq[i * LD + i] : 0.f));
                    out[(size_t)i * n + lane] = v;
                }
            }
        }
        if (n == 33 && wid == 1) {                              // column 32: zeros above the diagonal, one entry on it
            float v = lane < T ? s.Af[32 * LD + lane] * dBm[32 * LD + lane] : 0.f;
            v = warp_sum(v);
            out[(size_t)lane * n + 32] = 0.f;
            if (lane == 0) out[(size_t)32 * n + 32] = v + dkl * (Lq[32 * LD + 32] - 1.f / Lq[32 * LD + 32]);
        }
    }
    // dm = A dmu + dkl m  (one thread per row i)
    for (int i = tid; i < n; i += NT) {
        const float* arow = s.Af + i * LD;
        float acc = 0.f;
        for (int t = 0; t < T; ++t) acc = fmaf(arow[t], s.vecT[t], acc);
        b.dvar_mean[(size_t)c * n + i] = acc + dkl * s.mvec[i];
    }
    __syncthreads();

    GPB_TS(6);
    // =========================== P3 (fp64) ===========================
    double* Ld = s.RB;
    double* Gd = s.RC;
    {
        const double* Lg = a.L + (size_t)c * n * n;
        stage_block<double>(n, n, [&](int idx) { return Lg[idx]; }, [&](int idx, int i, int j, double v) { Ld[i * LD + j] = v; });
    }
    __syncthreads();
    for (int i = tid; i < n; i += NT) s.invd[i] = 1.0 / Ld[i * LD + i];
    __syncthreads();
    GPB_TS(7);
#if BLK4_ADJ
    gp::blk4_trsm_lowerT_left<double, 33, 33>(Ld, LD, s.invd, s.RA, LD, n, T, Gd);   // RA <- dK_ZX = L^-T dA (scratch: RC, written only below)
#else
    if (wid == 0) trsm_lowerT_cols<double>(Ld, s.invd, s.RA, n, T);   // RA <- dK_ZX = L^-T dA
    __syncthreads();
#endif
    GPB_TS(8);
    // dL = -tril(dK_ZX A^T)  (lane = column j < 32: row j of A)
    {
        const float* arow = s.Af + min(lane, n - 1) * LD;
        for (int i0 = 4 * wid; i0 < n; i0 += 4 * NW) {
            const double* d0 = s.RA + i0 * LD;
            const double* d1 = s.RA + min(i0 + 1, n - 1) * LD;
            const double* d2 = s.RA + min(i0 + 2, n - 1) * LD;
            const double* d3 = s.RA + min(i0 + 3, n - 1) * LD;
            double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
            for (int t = 0; t < T; ++t) {
                const double av = (double)arow[t];
                acc0 += d0[t] * av; acc1 += d1[t] * av; acc2 += d2[t] * av; acc3 += d3[t] * av;
            }
            const double accs[4] = {acc0, acc1, acc2, acc3};
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                const int i = i0 + x;
                if (i < n && lane <= i && lane < n) Gd[i * LD + lane] = -accs[x];
            }
        }
        if (n == 33 && wid == 1) {
            double v = lane < T ? s.RA[32 * LD + lane] * (double)s.Af[32 * LD + lane] : 0.0;
            v = warp_sum(v);
            if (lane == 0) Gd[32 * LD + 32] = -v;
        }
    }
    __syncthreads();
    each_block(n, n, [&](int idx, int i, int j) {              // dK_ZX leaves RA for the output block: RA becomes the adjoint's scratch
        if (j < T) dKt[idx] += (float)s.RA[i * LD + j];         // (the same thread wrote dKt[idx], the dSigma block, in P1)
    });
    GPB_TS(9);
    chol_adj_block<double>(Ld, s.invd, Gd, s.RA, n, BLK4_ADJ ? reinterpret_cast<double*>(s.Af) : nullptr);     // Af is dead after dL

    GPB_TS(10);
    // =========================== P4: kernel adjoint (same CTA; d < 0 skips it: the stand-alone kernel then runs) ===========================
    const int d = (int)a.d;
    if (!fuse_kernel_adjoint) {
        each_block(n, n, [&](int idx, int i, int j) { dKt[idx] += (float)gp::sym_from_rev<double>(Gd, LD, i, j); });
        return;
    }
    // d loss / d K block and the kernel values into RA (the adjoint's scratch is dead); then RB + RC (L and dL are dead, the two
    // regions are contiguous) hold the streamed chunk tile and the per-feature accumulators of gp::kernel_adjoint_block
    // (one Gram-shaped pass over Z with 4x4 register tiles)
    float* dK = reinterpret_cast<float*>(s.RA);
    float* Wm = dK + NN;
    each_block(n, n, [&](int idx, int i, int j) {
        dK[i * LD + j] = dKt[idx] + (float)gp::sym_from_rev<double>(Gd, LD, i, j);
        Wm[i * LD + j] = ks[1 + idx];
    });
    __syncthreads();
    const int dp = (d + 3) & ~3;
    float* invls = reinterpret_cast<float*>(s.RB);            // RB | RC: 2 * NN doubles = 17 KB
    float* qls = invls + dp;
    float* dzl = qls + dp;
    float* rs = dzl + dp;
    float* cs = rs + 36;
    float* tileA = cs + 36;
    const int kt = a.kernel_type;
    if (kt != CLIPGP_KERNEL_LINEAR)
        for (int k = tid; k < d; k += NT) invls[k] = 1.f / softplusf(a.raw_lengthscale[(size_t)c * d + k]);
    float amp = 1.f;
    if (kt == CLIPGP_KERNEL_RBF) amp = softplusf(a.raw_outputscale[c]);
    if (kt == CLIPGP_KERNEL_LINEAR) amp = softplusf(a.raw_variance[c]);
    for (int k = tid; k < d; k += NT) { qls[k] = 0.f; dzl[k] = 0.f; }
    __syncthreads();
    const float* Zc = a.Z + (size_t)c * n * d;
    GPB_TS(11);
#if KADJ_V
    const float damp = kernel_adjoint_sym(dK, Wm, tileA + gp::pad4(n) * KV, Zc, n, d, kt, amp, invls, tileA, qls, dzl, rs, cs);
#else
    const float damp = gp::kernel_adjoint_block(dK, LD, Wm, Zc, n, Zc, n, d, kt, amp, invls, tileA, tileA, qls, dzl, n - 1, n - 1, rs, cs);
#endif
    GPB_TS(12);
    __shared__ float red[32];
    const float damp_tot = block_sum(damp, red);
    if (tid == 0) {
        if (kt == CLIPGP_KERNEL_RBF && b.draw_outputscale) b.draw_outputscale[c] = damp_tot * sigmoidf_(a.raw_outputscale[c]);
        if (kt == CLIPGP_KERNEL_LINEAR && b.draw_variance) b.draw_variance[c] = damp_tot * sigmoidf_(a.raw_variance[c]);
    }
    __syncthreads();
    for (int k = tid; k < d; k += NT) {
        if (kt != CLIPGP_KERNEL_LINEAR && b.draw_lengthscale)
            b.draw_lengthscale[(size_t)c * d + k] = -2.f * qls[k] * invls[k] * sigmoidf_(a.raw_lengthscale[(size_t)c * d + k]);
        if (b.dZ_last) b.dZ_last[(size_t)c * d + k] = dzl[k];
    }
}

}  // namespace gpw
}  // namespace clipgp

using namespace clipgp;

#ifdef CLIPGP_PHASE_TS
extern "C" int clipgp_debug_phase_ts_bwd(long long* out) {
    return (int)cudaMemcpyFromSymbol(out, gpw::g_phase_ts_bwd, sizeof(long long) * 64);
}
#endif

// Fused prototype adjoint: S gradient rows of D floats plus two [S][32] tables must fit RA | RB | RC, S <= 12 register accumulators.
extern "C" int clipgp_gp_fused_proto_bwd_ok(int64_t T, int64_t n, int64_t d, int64_t D, int64_t S) {
    return (clipgp_gp_warp_path_ok(T, n, d) && D >= 4 && (D % 4) == 0 && S >= 1 && S <= 12 &&
            (size_t)S * (size_t)(D + 64) * sizeof(float) <= 3 * sizeof(double) * gpw::NN) ? 1 : 0;
}

// Shared memory the fused kernel-adjoint stage needs in RB + RC (floats): inverse length-scales, two per-feature accumulators,
// row / column sums and one [pad4(n)][KCP] chunk tile.
extern "C" int clipgp_gp_warp_fused_adjoint_ok(int64_t n, int64_t d) {
    const size_t need = 3 * (size_t)((d + 3) & ~3) + 72 + (size_t)gp::pad4((int)n) * gpw::KV + (size_t)n * gpw::KV;   // + transposed W
    return need * sizeof(float) <= 2 * sizeof(double) * gpw::NN ? 1 : 0;
}

int clipgp_gp_backward_warp_launch(const clipgp_gp_args* a, const clipgp_gp_bwd_args* b, cudaStream_t st, int fuse) {
    static bool attr_set = false;
    if (!attr_set) {
        CLIPGP_CUDA(cudaFuncSetAttribute(gpw::gp_backward_warp_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                         cudaSharedmemCarveoutMaxShared));
        attr_set = true;
    }
    gpw::gp_backward_warp_kernel<<<gp_grid(a), gpw::NT, sizeof(gpw::BwdSmem), st>>>(*a, *b, fuse);
    return check_launch("gp_backward_warp_kernel");
}
