// Weighted class prototypes P[s,c,:] = sum_t w[s,c,t] E[c,t,:]  (gp_template_weigher.py:221, the einsum
// "skm,kmd->skd"), fused with the row normalisation every head applies next (adapter.py:246/425,
// taskres.py:109-113, clip_adapter.py:94, tip_adapter.py:136) and with the MC reductions of the
// collapsed eval path (mean_s of unit prototypes) and of the prototype-init path (normalize(mean_s P_s),
// taskres.py:281-285).
//
// Roofline: HBM.  Algorithmic bytes per class: E[c] (T*D*4) read once per sample chunk + outputs.
#include "common.cuh"

namespace clipgp {

constexpr int kProtoThreads = 128;

// grid = (C, ceil(S/SB)); each thread owns column groups of 4 floats.
// mode bits: see clipgp.h CLIPGP_PROTO_*
template <int SB, int G /* float4 groups per thread */>
__global__ void __launch_bounds__(kProtoThreads) proto_forward_kernel(
    const float* __restrict__ w, const float* __restrict__ E, int64_t S, int64_t C, int T, int D,
    const float* __restrict__ residual, float alpha,
    float* __restrict__ P_raw, float* __restrict__ P_hat, float* __restrict__ norm, __nv_bfloat16* __restrict__ P_hat_bf16,
    float* __restrict__ mean_hat_accum, float* __restrict__ mean_raw_accum) {
    __shared__ float ws[SB][CLIPGP_GP_MAX_T];
    __shared__ float red[SB][kProtoThreads / 32];
    __shared__ float inv[SB], inv2[SB];
    const int c = blockIdx.x;
    const int64_t s0 = (int64_t)blockIdx.y * SB;
    const int sb = (int)min((int64_t)SB, S - s0);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int idx = tid; idx < SB * T; idx += blockDim.x) {
        const int s = idx / T, t = idx - s * T;
        ws[s][t] = (s < sb) ? w[((s0 + s) * C + c) * T + t] : 0.f;
    }
    __syncthreads();
    const int D4 = D >> 2;
    float4 acc[SB][G];
#pragma unroll
    for (int s = 0; s < SB; ++s)
#pragma unroll
        for (int g = 0; g < G; ++g) acc[s][g] = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4* Ec = reinterpret_cast<const float4*>(E + (size_t)c * T * D);
    // TB template rows are fetched before any is consumed: TB*G 16-byte loads in flight per thread hide the HBM latency
    constexpr int TB = (G == 1) ? 8 : (G == 2 ? 4 : 2);
    for (int t0 = 0; t0 < T; t0 += TB) {
        float4 e[TB][G];
#pragma unroll
        for (int u = 0; u < TB; ++u)
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const int col = tid + g * kProtoThreads;
                e[u][g] = (col < D4 && t0 + u < T) ? __ldg(Ec + (size_t)(t0 + u) * D4 + col) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
        for (int u = 0; u < TB; ++u) {
            if (t0 + u < T) {
#pragma unroll
                for (int s = 0; s < SB; ++s) {
                    const float wv = ws[s][t0 + u];
#pragma unroll
                    for (int g = 0; g < G; ++g) {
                        acc[s][g].x = fmaf(wv, e[u][g].x, acc[s][g].x); acc[s][g].y = fmaf(wv, e[u][g].y, acc[s][g].y);
                        acc[s][g].z = fmaf(wv, e[u][g].z, acc[s][g].z); acc[s][g].w = fmaf(wv, e[u][g].w, acc[s][g].w);
                    }
                }
            }
        }
    }
    // ---- row norms
#pragma unroll
    for (int s = 0; s < SB; ++s) {
        float q = 0.f;
#pragma unroll
        for (int g = 0; g < G; ++g) q += acc[s][g].x * acc[s][g].x + acc[s][g].y * acc[s][g].y + acc[s][g].z * acc[s][g].z + acc[s][g].w * acc[s][g].w;
        q = warp_sum(q);
        if (lane == 0) red[s][warp] = q;
    }
    __syncthreads();
    if (tid < SB) {
        float q = 0.f;
        for (int k = 0; k < kProtoThreads / 32; ++k) q += red[tid][k];
        const float nrm = sqrtf(q);
        inv[tid] = 1.f / fmaxf(nrm, 1e-12f);   // F.normalize eps
        if (norm && tid < sb) norm[(s0 + tid) * C + c] = nrm;
    }
    __syncthreads();
    // ---- optional TaskRes residual: t_s = normalize(p_hat_s + alpha x_c)   (taskres.py:111-113)
    float4 rs[G];
    if (residual) {
        const float4* Rc = reinterpret_cast<const float4*>(residual + (size_t)c * D);
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const int col = tid + g * kProtoThreads;
            rs[g] = (col < D4) ? __ldg(Rc + col) : make_float4(0.f, 0.f, 0.f, 0.f);
            rs[g].x *= alpha; rs[g].y *= alpha; rs[g].z *= alpha; rs[g].w *= alpha;
        }
        __syncthreads();
#pragma unroll
        for (int s = 0; s < SB; ++s) {
            float q = 0.f;
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const float iv = inv[s];
                const float x = fmaf(acc[s][g].x, iv, rs[g].x), y = fmaf(acc[s][g].y, iv, rs[g].y);
                const float z = fmaf(acc[s][g].z, iv, rs[g].z), u = fmaf(acc[s][g].w, iv, rs[g].w);
                q += x * x + y * y + z * z + u * u;
            }
            q = warp_sum(q);
            if (lane == 0) red[s][warp] = q;
        }
        __syncthreads();
        if (tid < SB) {
            float q = 0.f;
            for (int k = 0; k < kProtoThreads / 32; ++k) q += red[tid][k];
            inv2[tid] = 1.f / sqrtf(q);          // reference divides by the plain norm here (taskres.py:113)
        }
        __syncthreads();
    }
    // ---- outputs
    float4 msum[G], rsum[G];
#pragma unroll
    for (int g = 0; g < G; ++g) { msum[g] = make_float4(0.f, 0.f, 0.f, 0.f); rsum[g] = msum[g]; }
#pragma unroll
    for (int s = 0; s < SB; ++s) {
        if (s < sb) {
            const size_t row = ((size_t)(s0 + s) * C + c) * D;
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const int col = tid + g * kProtoThreads;
                if (col < D4) {
                    const float4 a = acc[s][g];
                    if (P_raw) reinterpret_cast<float4*>(P_raw + row)[col] = a;
                    float4 h = make_float4(a.x * inv[s], a.y * inv[s], a.z * inv[s], a.w * inv[s]);
                    if (residual) {
                        h.x = (h.x + rs[g].x) * inv2[s]; h.y = (h.y + rs[g].y) * inv2[s];
                        h.z = (h.z + rs[g].z) * inv2[s]; h.w = (h.w + rs[g].w) * inv2[s];
                    }
                    if (P_hat) reinterpret_cast<float4*>(P_hat + row)[col] = h;
                    if (P_hat_bf16) {
                        __nv_bfloat162 lo = __floats2bfloat162_rn(h.x, h.y), hi = __floats2bfloat162_rn(h.z, h.w);
                        uint2 pk; pk.x = *reinterpret_cast<uint32_t*>(&lo); pk.y = *reinterpret_cast<uint32_t*>(&hi);
                        reinterpret_cast<uint2*>(P_hat_bf16 + row)[col] = pk;
                    }
                    msum[g].x += h.x; msum[g].y += h.y; msum[g].z += h.z; msum[g].w += h.w;
                    rsum[g].x += a.x; rsum[g].y += a.y; rsum[g].z += a.z; rsum[g].w += a.w;
                }
            }
        }
    }
    // sums over the samples of this chunk; the host divides / normalises after the last chunk
    if (mean_hat_accum || mean_raw_accum) {
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const int col = tid + g * kProtoThreads;
            if (col < D4) {
                if (mean_hat_accum) {
                    float* p = mean_hat_accum + (size_t)c * D + col * 4;
                    if (gridDim.y == 1) { p[0] = msum[g].x; p[1] = msum[g].y; p[2] = msum[g].z; p[3] = msum[g].w; }
                    else { atomicAdd(p, msum[g].x); atomicAdd(p + 1, msum[g].y); atomicAdd(p + 2, msum[g].z); atomicAdd(p + 3, msum[g].w); }
                }
                if (mean_raw_accum) {
                    float* p = mean_raw_accum + (size_t)c * D + col * 4;
                    if (gridDim.y == 1) { p[0] = rsum[g].x; p[1] = rsum[g].y; p[2] = rsum[g].z; p[3] = rsum[g].w; }
                    else { atomicAdd(p, rsum[g].x); atomicAdd(p + 1, rsum[g].y); atomicAdd(p + 2, rsum[g].z); atomicAdd(p + 3, rsum[g].w); }
                }
            }
        }
    }
}

// rows [R, D] *= scale, then optionally L2-normalise (used to finish the MC means).
__global__ void __launch_bounds__(128) scale_normalize_rows_kernel(float* __restrict__ x, int64_t R, int D, float scale,
                                                                   int normalize, __nv_bfloat16* __restrict__ out_bf16) {
    __shared__ float red[32];
    const int64_t r = blockIdx.x;
    float* row = x + r * D;
    float q = 0.f;
    for (int k = threadIdx.x; k < D; k += blockDim.x) { const float v = row[k] * scale; q += v * v; }
    const float tot = block_sum(q, red);
    const float f = normalize ? scale / sqrtf(tot) : scale;
    for (int k = threadIdx.x; k < D; k += blockDim.x) {
        const float v = row[k] * f;
        row[k] = v;
        if (out_bf16) out_bf16[r * D + k] = __float2bfloat16_rn(v);
    }
}

// dw[s,c,t] = <dP[s,c,:], E[c,t,:]>  with dP either given directly (P_hat == NULL) or derived from the
// gradient of the normalised rows:  dP = (dP_hat - P_hat <P_hat, dP_hat>) / |P|.
template <int SB>
__global__ void __launch_bounds__(kProtoThreads) proto_backward_kernel(
    const float* __restrict__ dP, int64_t dP_stride_s, float dP_scale, const float* __restrict__ P_hat,
    const float* __restrict__ norm, const float* __restrict__ E, int64_t S, int64_t C, int T, int D, float* __restrict__ dw) {
    extern __shared__ __align__(16) float g[];   // [SB][D]
    __shared__ float red[SB][kProtoThreads / 32];
    __shared__ float dots[SB];
    const int c = blockIdx.x;
    const int64_t s0 = (int64_t)blockIdx.y * SB;
    const int sb = (int)min((int64_t)SB, S - s0);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = kProtoThreads / 32;
    const int D4 = D >> 2;
    for (int s = 0; s < SB; ++s) {
        float q = 0.f;
        if (s < sb) {
            const size_t row = ((size_t)(s0 + s) * C + c) * D;
            const size_t grow = (size_t)(s0 + s) * dP_stride_s + (size_t)c * D;
            for (int col = tid; col < D4; col += blockDim.x) {
                float4 v = __ldg(reinterpret_cast<const float4*>(dP + grow) + col);
                v.x *= dP_scale; v.y *= dP_scale; v.z *= dP_scale; v.w *= dP_scale;
                reinterpret_cast<float4*>(g + (size_t)s * D)[col] = v;
                if (P_hat) {
                    const float4 h = __ldg(reinterpret_cast<const float4*>(P_hat + row) + col);
                    q += v.x * h.x + v.y * h.y + v.z * h.z + v.w * h.w;
                }
            }
        } else {
            for (int col = tid; col < D4; col += blockDim.x) reinterpret_cast<float4*>(g + (size_t)s * D)[col] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        q = warp_sum(q);
        if (lane == 0) red[s][warp] = q;
    }
    __syncthreads();
    if (P_hat) {
        if (tid < SB) {
            float q = 0.f;
            for (int k = 0; k < nwarps; ++k) q += red[tid][k];
            dots[tid] = q;
        }
        __syncthreads();
        for (int s = 0; s < sb; ++s) {
            const size_t row = ((size_t)(s0 + s) * C + c) * D;
            const float invn = 1.f / fmaxf(norm[(s0 + s) * C + c], 1e-12f);
            const float dt = dots[s];
            for (int col = tid; col < D4; col += blockDim.x) {
                float4 v = reinterpret_cast<float4*>(g + (size_t)s * D)[col];
                const float4 h = __ldg(reinterpret_cast<const float4*>(P_hat + row) + col);
                v.x = (v.x - h.x * dt) * invn; v.y = (v.y - h.y * dt) * invn;
                v.z = (v.z - h.z * dt) * invn; v.w = (v.w - h.w * dt) * invn;
                reinterpret_cast<float4*>(g + (size_t)s * D)[col] = v;
            }
        }
        __syncthreads();
    }
    const float4* Ec = reinterpret_cast<const float4*>(E + (size_t)c * T * D);
    // each warp walks its template rows with the next row's loads (up to 4 x 16 bytes per lane) already in flight
    constexpr int EP = 4;
    float4 nxt[EP];
    auto fetch = [&](int t, int col0) {
#pragma unroll
        for (int u = 0; u < EP; ++u) {
            const int col = col0 + lane + 32 * u;
            nxt[u] = (t < T && col < D4) ? __ldg(Ec + (size_t)t * D4 + col) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    fetch(warp, 0);
    for (int t = warp; t < T; t += nwarps) {
        float part[SB];
#pragma unroll
        for (int s = 0; s < SB; ++s) part[s] = 0.f;
        for (int col0 = 0; col0 < D4; col0 += 32 * EP) {
            float4 e[EP];
#pragma unroll
            for (int u = 0; u < EP; ++u) e[u] = nxt[u];
            if (col0 + 32 * EP < D4) fetch(t, col0 + 32 * EP); else fetch(t + nwarps, 0);
#pragma unroll
            for (int u = 0; u < EP; ++u) {
                const int col = col0 + lane + 32 * u;
                if (col < D4) {
#pragma unroll
                    for (int s = 0; s < SB; ++s) {
                        const float4 v = reinterpret_cast<const float4*>(g + (size_t)s * D)[col];
                        part[s] += e[u].x * v.x + e[u].y * v.y + e[u].z * v.z + e[u].w * v.w;
                    }
                }
            }
        }
#pragma unroll
        for (int s = 0; s < SB; ++s) {
            const float tot = warp_sum(part[s]);
            if (lane == 0 && s < sb) dw[((s0 + s) * C + c) * T + t] = tot;
        }
    }
}

template <int SB>
static int launch_proto_fwd(const float* w, const float* E, int64_t S, int64_t C, int T, int D, const float* residual,
                            float alpha, float* P_raw, float* P_hat, float* norm, __nv_bfloat16* P_bf16, float* mh,
                            float* mr, cudaStream_t st) {
    dim3 grid((unsigned)C, (unsigned)((S + SB - 1) / SB));
    const int D4 = D >> 2;
    const int G = (D4 + kProtoThreads - 1) / kProtoThreads;
#define LAUNCH(GG)                                                                                                   \
    proto_forward_kernel<SB, GG><<<grid, kProtoThreads, 0, st>>>(w, E, S, C, T, D, residual, alpha, P_raw, P_hat, norm, \
                                                                 P_bf16, mh, mr)
    if (G <= 1) LAUNCH(1);
    else if (G == 2) LAUNCH(2);
    else if (G <= 4) LAUNCH(4);
    else { set_error("proto_forward: D = %d too large (max 2048)", D); return CLIPGP_ERR_INVALID; }
#undef LAUNCH
    return check_launch("proto_forward_kernel");
}

}  // namespace clipgp

using namespace clipgp;

extern "C" int clipgp_proto_forward(const float* w, const float* E, int64_t S, int64_t C, int64_t T, int64_t D,
                                    const float* residual, float alpha, float* P_raw, float* P_hat, float* norm,
                                    void* P_hat_bf16, float* mean_hat, float* mean_raw, int finish_mean, void* stream) {
    CLIPGP_REQUIRE(S >= 1 && C >= 0 && T >= 1 && T <= CLIPGP_GP_MAX_T, "proto_forward: bad shape S=%lld C=%lld T=%lld",
                   (long long)S, (long long)C, (long long)T);
    CLIPGP_REQUIRE(D >= 4 && D % 4 == 0 && D <= 2048, "proto_forward: D must be a multiple of 4 in [4,2048] (got %lld)", (long long)D);
    if (C == 0) return CLIPGP_OK;
    CLIPGP_REQUIRE(w && E, "proto_forward: NULL input");
    cudaStream_t st = (cudaStream_t)stream;
    // samples per CTA: bounded by the register budget of the accumulators (SB x G float4 per thread)
    const int G = (int)((D / 4 + kProtoThreads - 1) / kProtoThreads);
    int SB = (G >= 3) ? 4 : (G == 2 ? 8 : 16);
    if (S <= 4) SB = 4;
    else if (S <= 8 && SB > 8) SB = 8;
    const int64_t chunks = (S + SB - 1) / SB;
    if (chunks > 1) {   // several sample chunks accumulate the MC sums atomically
        if (mean_hat) CLIPGP_CUDA(cudaMemsetAsync(mean_hat, 0, sizeof(float) * C * D, st));
        if (mean_raw) CLIPGP_CUDA(cudaMemsetAsync(mean_raw, 0, sizeof(float) * C * D, st));
    }
    int rc;
    __nv_bfloat16* pb = reinterpret_cast<__nv_bfloat16*>(P_hat_bf16);
    if (SB == 4) rc = launch_proto_fwd<4>(w, E, S, C, (int)T, (int)D, residual, alpha, P_raw, P_hat, norm, pb, mean_hat, mean_raw, st);
    else if (SB == 8) rc = launch_proto_fwd<8>(w, E, S, C, (int)T, (int)D, residual, alpha, P_raw, P_hat, norm, pb, mean_hat, mean_raw, st);
    else rc = launch_proto_fwd<16>(w, E, S, C, (int)T, (int)D, residual, alpha, P_raw, P_hat, norm, pb, mean_hat, mean_raw, st);
    if (rc != CLIPGP_OK) return rc;
    if (finish_mean) {
        // mean_hat <- (1/S) sum_s p_hat_s            (collapsed logit-mean eval: mean_s scale f.p_hat_s = scale f.mean_s p_hat_s)
        // mean_raw <- normalize((1/S) sum_s P_s)      (taskres.py:281-285 / clip_adapter.py:284-288 / tip_adapter.py:152-156)
        if (mean_hat) {
            scale_normalize_rows_kernel<<<(unsigned)C, 128, 0, st>>>(mean_hat, C, (int)D, 1.f / (float)S, 0, nullptr);
            rc = check_launch("scale_normalize_rows_kernel");
            if (rc != CLIPGP_OK) return rc;
        }
        if (mean_raw) {
            scale_normalize_rows_kernel<<<(unsigned)C, 128, 0, st>>>(mean_raw, C, (int)D, 1.f / (float)S, 1, nullptr);
            rc = check_launch("scale_normalize_rows_kernel");
            if (rc != CLIPGP_OK) return rc;
        }
    }
    return CLIPGP_OK;
}

extern "C" int clipgp_proto_backward(const float* dP, int64_t dP_stride_s, float dP_scale, const float* P_hat, const float* norm,
                                     const float* E, int64_t S, int64_t C, int64_t T, int64_t D, float* dw, void* stream) {
    CLIPGP_REQUIRE(S >= 1 && C >= 0 && T >= 1 && T <= CLIPGP_GP_MAX_T, "proto_backward: bad shape");
    CLIPGP_REQUIRE(D >= 4 && D % 4 == 0 && D <= 2048, "proto_backward: D must be a multiple of 4 in [4,2048]");
    if (C == 0) return CLIPGP_OK;
    CLIPGP_REQUIRE(dP && E && dw, "proto_backward: NULL pointer");
    CLIPGP_REQUIRE((P_hat == nullptr) == (norm == nullptr), "proto_backward: P_hat and norm go together");
    CLIPGP_REQUIRE(dP_stride_s == 0 || dP_stride_s >= C * D, "proto_backward: dP_stride_s must be 0 (broadcast over samples) or >= C*D");
    // samples per CTA: E[c] is streamed once per chunk of SB samples and the FMA / shared-memory work grows with SB, so split S
    // into ceil(S/8) equal chunks (S = 10 -> 2 x 5 instead of 8 + 2)
    const int chunks = (int)((S + 7) / 8);
    const int SB = (int)((S + chunks - 1) / chunks);
    const size_t smem = sizeof(float) * 8 * (size_t)D;
    dim3 grid((unsigned)C, (unsigned)chunks);
#define LAUNCH_BWD(SBV)                                                                                                          \
    do {                                                                                                                         \
        static size_t smem_set = 0;                                                                                              \
        if (smem > smem_set) {                                                                                                   \
            CLIPGP_CUDA(cudaFuncSetAttribute(proto_backward_kernel<SBV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            smem_set = smem;                                                                                                     \
        }                                                                                                                        \
        proto_backward_kernel<SBV><<<grid, kProtoThreads, smem, (cudaStream_t)stream>>>(dP, dP_stride_s, dP_scale, P_hat, norm, E, S, C, \
                                                                                        (int)T, (int)D, dw);                     \
    } while (0)
    switch (SB) {
        case 1: case 2: case 3: case 4: LAUNCH_BWD(4); break;
        case 5: LAUNCH_BWD(5); break;
        case 6: LAUNCH_BWD(6); break;
        case 7: LAUNCH_BWD(7); break;
        default: LAUNCH_BWD(8); break;
    }
#undef LAUNCH_BWD
    return check_launch("proto_backward_kernel");
}
