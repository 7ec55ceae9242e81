// Row-wise kernels around the logit GEMMs: L2 row normalisation (forward / adjoint), softmax
// cross-entropy (forward + gradient in one pass), the ||W - I||_F^2 regulariser, and the fused AdamW
// update.  All HBM-bound streaming kernels (one warp per row / grid-stride, float4 where aligned).
#include <float.h>

#include "common.cuh"

namespace clipgp {

// y = x / max(|x|, eps) per row (F.normalize, adapter.py:240,246); inv_norm[r] = 1 / max(|x|, eps).
__global__ void __launch_bounds__(256) rownorm_fwd_kernel(const float* __restrict__ x, int64_t R, int D,
                                                          float* __restrict__ y, float* __restrict__ inv_norm,
                                                          __nv_bfloat16* __restrict__ y_bf16) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= R) return;
    const float* xr = x + row * D;
    float q = 0.f;
    for (int k = lane; k < D; k += 32) { const float v = xr[k]; q = fmaf(v, v, q); }
    q = warp_sum(q);
    const float inv = 1.f / fmaxf(sqrtf(q), 1e-12f);
    for (int k = lane; k < D; k += 32) {
        const float v = xr[k] * inv;
        if (y) y[row * D + k] = v;
        if (y_bf16) y_bf16[row * D + k] = __float2bfloat16_rn(v);
    }
    if (lane == 0 && inv_norm) inv_norm[row] = inv;
}

// F.normalize fused with the bf16 operand cast of the next GEMM: out[r, g*seg_stride + k] in the layouts of cast_bf16_kernel
// (mode 0 plain, 1 A-split, 2 B-split).  The fp32 unit rows are written only if y != NULL.  One warp per row, 16-byte reads.
__global__ void __launch_bounds__(256) rownorm_cast_kernel(const float* __restrict__ x, int64_t R, int D, float* __restrict__ y,
                                                           float* __restrict__ inv_norm, __nv_bfloat16* __restrict__ out, int64_t out_ld,
                                                           int64_t seg_stride, int mode) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= R) return;
    const float4* xr = reinterpret_cast<const float4*>(x + row * D);
    const int D4 = D >> 2;
    float q = 0.f;
    for (int k = lane; k < D4; k += 32) { const float4 v = __ldg(xr + k); q += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w; }
    q = warp_sum(q);
    const float inv = 1.f / fmaxf(sqrtf(q), 1e-12f);
    for (int k = lane; k < D4; k += 32) {
        float4 v = __ldg(xr + k);
        v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv;
        if (y) reinterpret_cast<float4*>(y + row * D)[k] = v;
        const __nv_bfloat162 h0 = __floats2bfloat162_rn(v.x, v.y), h1 = __floats2bfloat162_rn(v.z, v.w);
        uint2 hi; hi.x = *reinterpret_cast<const uint32_t*>(&h0); hi.y = *reinterpret_cast<const uint32_t*>(&h1);
        __nv_bfloat16* o = out + row * out_ld + 4 * k;
        *reinterpret_cast<uint2*>(o) = hi;
        if (mode != 0) {
            const __nv_bfloat162 l0 = __floats2bfloat162_rn(v.x - __low2float(h0), v.y - __high2float(h0));
            const __nv_bfloat162 l1 = __floats2bfloat162_rn(v.z - __low2float(h1), v.w - __high2float(h1));
            uint2 lo; lo.x = *reinterpret_cast<const uint32_t*>(&l0); lo.y = *reinterpret_cast<const uint32_t*>(&l1);
            *reinterpret_cast<uint2*>(o + seg_stride) = (mode == 1) ? hi : lo;
            *reinterpret_cast<uint2*>(o + 2 * seg_stride) = (mode == 1) ? lo : hi;
        }
    }
    if (lane == 0 && inv_norm) inv_norm[row] = inv;
}

// dx = (dy - y <y, dy>) * inv_norm
__global__ void __launch_bounds__(256) rownorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                                          const float* __restrict__ inv_norm, int64_t R, int D,
                                                          float* __restrict__ dx) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= R) return;
    float dt = 0.f;
    for (int k = lane; k < D; k += 32) dt = fmaf(y[row * D + k], dy[row * D + k], dt);
    dt = warp_sum(dt);
    const float inv = inv_norm[row];
    for (int k = lane; k < D; k += 32) dx[row * D + k] = (dy[row * D + k] - y[row * D + k] * dt) * inv;
}

// One warp per logits row r (label index r / rows_per_label).  loss_sum += loss_scale * sum_r CE_r;
// dlogits = grad_scale * (softmax - onehot), may alias logits.
__global__ void __launch_bounds__(256) softmax_ce_kernel(const float* logits, int64_t ld, const int64_t* __restrict__ labels,
                                                         int64_t R, int64_t rows_per_label, int C, float* __restrict__ loss_rows,
                                                         float* loss_sum, float loss_scale, float* dlogits, int64_t ldd,
                                                         float grad_scale) {
    __shared__ float part[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    float my_loss = 0.f;
    // register-resident rows (C <= 1024, 16-byte aligned): ONE 16-byte read per 4 logits, one exp per logit, one 16-byte write -- the
    // general path below reads the row three times and evaluates exp twice (this kernel sits on the step's critical path between the
    // logit GEMM and the d P_hat GEMM: it is pure latency at B = 128)
#ifdef SOFTMAX_NOVEC
    const bool vec = false;
#else
    const bool vec = C <= 1024 && (C & 3) == 0 && (ld & 3) == 0 && ((reinterpret_cast<uintptr_t>(logits) & 15u) == 0) &&
                     (!dlogits || ((ldd & 3) == 0 && (reinterpret_cast<uintptr_t>(dlogits) & 15u) == 0));
#endif
    if (row < R && vec) {
        const float4* x4 = reinterpret_cast<const float4*>(logits + row * ld);
        const int lab = (int)labels[row / rows_per_label];
        const int C4 = C >> 2;
        float4 v[8];
        float m = -FLT_MAX;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int j4 = lane + 32 * u;
            v[u] = (j4 < C4) ? x4[j4] : make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX);
            m = fmaxf(m, fmaxf(fmaxf(v[u].x, v[u].y), fmaxf(v[u].z, v[u].w)));
        }
        m = warp_max(m);
        float s = 0.f, xl = 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int j4 = lane + 32 * u;
            if (j4 < C4) {
                if ((lab >> 2) == j4) xl = (lab & 3) == 0 ? v[u].x : ((lab & 3) == 1 ? v[u].y : ((lab & 3) == 2 ? v[u].z : v[u].w));
                v[u].x = expf(v[u].x - m); v[u].y = expf(v[u].y - m); v[u].z = expf(v[u].z - m); v[u].w = expf(v[u].w - m);
                s += (v[u].x + v[u].y) + (v[u].z + v[u].w);
            }
        }
        s = warp_sum(s);
        xl = warp_sum(xl);                                   // exactly one lane holds the label's logit
        my_loss = m + logf(s) - xl;
        if (loss_rows && lane == 0) loss_rows[row] = my_loss;
        if (dlogits) {
            const float gs = grad_scale / s;
            float4* g4 = reinterpret_cast<float4*>(dlogits + row * ldd);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int j4 = lane + 32 * u;
                if (j4 < C4) {
                    float4 o = make_float4(v[u].x * gs, v[u].y * gs, v[u].z * gs, v[u].w * gs);
                    if ((lab >> 2) == j4) {
                        if ((lab & 3) == 0) o.x -= grad_scale; else if ((lab & 3) == 1) o.y -= grad_scale;
                        else if ((lab & 3) == 2) o.z -= grad_scale; else o.w -= grad_scale;
                    }
                    g4[j4] = o;
                }
            }
        }
    } else if (row < R) {
        const float* x = logits + row * ld;
        const int lab = (int)labels[row / rows_per_label];
        float m = -FLT_MAX;
        for (int j = lane; j < C; j += 32) m = fmaxf(m, x[j]);
        m = warp_max(m);
        float s = 0.f;
        for (int j = lane; j < C; j += 32) s += expf(x[j] - m);
        s = warp_sum(s);
        const float lse = m + logf(s);
        const float xl = x[lab];
        my_loss = lse - xl;
        if (loss_rows && lane == 0) loss_rows[row] = my_loss;
        if (dlogits) {
            const float invs = 1.f / s;
            float* g = dlogits + row * ldd;
            for (int j = lane; j < C; j += 32) {
                const float p = expf(x[j] - m) * invs;
                g[j] = grad_scale * (p - (j == lab ? 1.f : 0.f));
            }
        }
    }
    if (loss_sum) {
        if (lane == 0) part[warp] = (row < R) ? my_loss : 0.f;
        __syncthreads();
        if (threadIdx.x == 0) {
            float t = 0.f;
            for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += part[k];
            atomicAdd(loss_sum, t * loss_scale);
        }
    }
}

// loss += coef * ||W - I||_F^2 ; dW += 2 coef (W - I)      (adapter.py:468-476)
__global__ void __launch_bounds__(256) l2_identity_kernel(const float* __restrict__ W, int D, float coef,
                                                          float* __restrict__ dW, float* loss_sum) {
    __shared__ float red[32];
    float q = 0.f;
    const int64_t n = (int64_t)D * D;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / D), c = (int)(i - (int64_t)r * D);
        const float v = W[i] - (r == c ? 1.f : 0.f);
        q = fmaf(v, v, q);
        if (dW) dW[i] += 2.f * coef * v;
    }
    const float tot = block_sum(q, red);
    if (threadIdx.x == 0 && loss_sum) atomicAdd(loss_sum, coef * tot);
}

// torch.optim.AdamW semantics (decoupled weight decay, bias correction, eps added after sqrt):
//   p *= 1 - lr*wd ; m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ;
//   p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
// `step` is read from device memory (so a captured CUDA graph advances it); row_mask (optional) restricts the
// update to elements with mask != 0 semantics-free (mask multiplies the gradient first, gp_template_weigher.py:76-79).
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                    float* __restrict__ v, int64_t n, float lr, float b1, float b2, float eps,
                                                    float wd, const int64_t* __restrict__ step_ptr, const float* __restrict__ lr_ptr) {
    if (lr_ptr) lr = *lr_ptr;                       // device-resident learning rate: schedules advance without re-capturing the graph
    const float t = (float)(*step_ptr);
    const float bc1 = 1.f - powf(b1, t), bc2 = 1.f - powf(b2, t);
    const float step_size = lr / bc1;
    const float inv_sqrt_bc2 = rsqrtf(bc2);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float gi = g[i];
        float pi = p[i] * (1.f - lr * wd);
        const float mi = b1 * m[i] + (1.f - b1) * gi;
        const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        m[i] = mi; v[i] = vi;
        pi -= step_size * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
        p[i] = pi;
    }
}

// AdamW on a [R, K] weight (K % 4 == 0) that also emits the updated weight as the bf16 operand of the next step's tensor-core GEMM
// (layouts of clipgp_cast_bf16): the trainable Tip-Adapter-F keys are re-cast every step, this saves that pass over the weight.
__global__ void __launch_bounds__(256) adamw_cast_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                         float* __restrict__ v, int64_t n4, int K4, const float* __restrict__ lr_ptr, float b1,
                                                         float b2, float eps, float wd, const int64_t* __restrict__ step_ptr,
                                                         __nv_bfloat16* __restrict__ out, int64_t out_ld, int64_t seg_stride, int mode) {
    const float lr = *lr_ptr;
    const float t = (float)(*step_ptr);
    const float bc1 = 1.f - powf(b1, t), bc2 = 1.f - powf(b2, t);
    const float step_size = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2), decay = 1.f - lr * wd;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 gi = reinterpret_cast<const float4*>(g)[i];
        float4 pi = reinterpret_cast<float4*>(p)[i], mi = reinterpret_cast<float4*>(m)[i], vi = reinterpret_cast<float4*>(v)[i];
#define CLIPGP_ADAMW1(c)                                                        \
        mi.c = b1 * mi.c + (1.f - b1) * gi.c;                                   \
        vi.c = b2 * vi.c + (1.f - b2) * gi.c * gi.c;                            \
        pi.c = pi.c * decay - step_size * mi.c / (sqrtf(vi.c) * inv_sqrt_bc2 + eps);
        CLIPGP_ADAMW1(x) CLIPGP_ADAMW1(y) CLIPGP_ADAMW1(z) CLIPGP_ADAMW1(w)
#undef CLIPGP_ADAMW1
        reinterpret_cast<float4*>(p)[i] = pi; reinterpret_cast<float4*>(m)[i] = mi; reinterpret_cast<float4*>(v)[i] = vi;
        const int64_t r = i / K4; const int k = (int)(i - r * K4) << 2;
        const __nv_bfloat162 h0 = __floats2bfloat162_rn(pi.x, pi.y), h1 = __floats2bfloat162_rn(pi.z, pi.w);
        uint2 hi; hi.x = *reinterpret_cast<const uint32_t*>(&h0); hi.y = *reinterpret_cast<const uint32_t*>(&h1);
        __nv_bfloat16* o = out + r * out_ld + k;
        *reinterpret_cast<uint2*>(o) = hi;
        if (mode != 0) {
            const __nv_bfloat162 l0 = __floats2bfloat162_rn(pi.x - __low2float(h0), pi.y - __high2float(h0));
            const __nv_bfloat162 l1 = __floats2bfloat162_rn(pi.z - __low2float(h1), pi.w - __high2float(h1));
            uint2 lo; lo.x = *reinterpret_cast<const uint32_t*>(&l0); lo.y = *reinterpret_cast<const uint32_t*>(&l1);
            *reinterpret_cast<uint2*>(o + seg_stride) = (mode == 1) ? hi : lo;
            *reinterpret_cast<uint2*>(o + 2 * seg_stride) = (mode == 1) ? lo : hi;
        }
    }
}

__global__ void __launch_bounds__(256) sum_accumulate_kernel(const float* __restrict__ x, int64_t n, float scale, float* out) {
    __shared__ float red[32];
    float q = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) q += x[i];
    const float tot = block_sum(q, red);
    if (threadIdx.x == 0) atomicAdd(out, tot * scale);
}

// fp32 -> bf16 operand for the tensor-core GEMMs.  mode 0: out[r, K] = bf16(x).  Split modes emulate fp32 products
// with three bf16 MMAs (x = hi + lo, a.b ~= a_hi b_hi + a_hi b_lo + a_lo b_hi, relative error ~2^-16):
// mode 1 (A side): out[r, 3K] = [hi | hi | lo];  mode 2 (B side): out[r, 3K] = [hi | lo | hi].
// out row r, segment g, column k lives at out[r*out_ld + g*seg_stride + k] (seg_stride = K for the layouts above;
// callers that interleave MC samples along K pass their own strides).
// Vector form (K, ldx, out_ld, seg_stride multiples of 4; 16-byte aligned x, 8-byte aligned out): one 16-byte read and up to
// three 8-byte writes per thread and step.
__global__ void __launch_bounds__(256) cast_bf16_vec4_kernel(const float* __restrict__ x, int64_t R, int K, int64_t ldx,
                                                             __nv_bfloat16* __restrict__ out, int64_t out_ld, int64_t seg_stride, int mode) {
    const int K4 = K >> 2;
    const int64_t total = R * (int64_t)K4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / K4; const int k = (int)(i - r * K4) << 2;
        const float4 v = __ldg(reinterpret_cast<const float4*>(x + r * ldx + k));
        const __nv_bfloat162 h0 = __floats2bfloat162_rn(v.x, v.y), h1 = __floats2bfloat162_rn(v.z, v.w);
        uint2 hi; hi.x = *reinterpret_cast<const uint32_t*>(&h0); hi.y = *reinterpret_cast<const uint32_t*>(&h1);
        __nv_bfloat16* o = out + r * out_ld + k;
        *reinterpret_cast<uint2*>(o) = hi;
        if (mode != 0) {
            const __nv_bfloat162 l0 = __floats2bfloat162_rn(v.x - __low2float(h0), v.y - __high2float(h0));
            const __nv_bfloat162 l1 = __floats2bfloat162_rn(v.z - __low2float(h1), v.w - __high2float(h1));
            uint2 lo; lo.x = *reinterpret_cast<const uint32_t*>(&l0); lo.y = *reinterpret_cast<const uint32_t*>(&l1);
            *reinterpret_cast<uint2*>(o + seg_stride) = (mode == 1) ? hi : lo;
            *reinterpret_cast<uint2*>(o + 2 * seg_stride) = (mode == 1) ? lo : hi;
        }
    }
}

// Eight elements per thread (K, ldx, out_ld, seg_stride multiples of 8; 16-byte aligned x and out; R * K / 8 < 2^31): two 16-byte
// reads and up to three 16-byte writes per item, two items in flight per thread, 32-bit index arithmetic.  The streaming form of
// the big operand casts (50 000 x 512 eval features: 102 MB in, 154 MB out).
__device__ __forceinline__ void split8(const float4 a, const float4 b, uint4& hi, uint4& lo) {
    const __nv_bfloat162 h0 = __floats2bfloat162_rn(a.x, a.y), h1 = __floats2bfloat162_rn(a.z, a.w);
    const __nv_bfloat162 h2 = __floats2bfloat162_rn(b.x, b.y), h3 = __floats2bfloat162_rn(b.z, b.w);
    const __nv_bfloat162 l0 = __floats2bfloat162_rn(a.x - __low2float(h0), a.y - __high2float(h0));
    const __nv_bfloat162 l1 = __floats2bfloat162_rn(a.z - __low2float(h1), a.w - __high2float(h1));
    const __nv_bfloat162 l2 = __floats2bfloat162_rn(b.x - __low2float(h2), b.y - __high2float(h2));
    const __nv_bfloat162 l3 = __floats2bfloat162_rn(b.z - __low2float(h3), b.w - __high2float(h3));
    hi = make_uint4(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1),
                    *reinterpret_cast<const uint32_t*>(&h2), *reinterpret_cast<const uint32_t*>(&h3));
    lo = make_uint4(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1),
                    *reinterpret_cast<const uint32_t*>(&l2), *reinterpret_cast<const uint32_t*>(&l3));
}

__global__ void __launch_bounds__(256) cast_bf16_vec8_kernel(const float* __restrict__ x, unsigned total, unsigned K8, int64_t ldx,
                                                             __nv_bfloat16* __restrict__ out, int64_t out_ld, int64_t seg_stride, int mode) {
    const unsigned stride = gridDim.x * blockDim.x;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += 2 * stride) {
        const unsigned i1 = i + stride;
        const bool two = i1 < total;
        const unsigned r0 = i / K8, k0 = (i - r0 * K8) << 3;
        const unsigned r1 = two ? i1 / K8 : r0, k1 = two ? (i1 - r1 * K8) << 3 : k0;
        const float4* p0 = reinterpret_cast<const float4*>(x + (int64_t)r0 * ldx + k0);
        const float4* p1 = reinterpret_cast<const float4*>(x + (int64_t)r1 * ldx + k1);
        const float4 a0 = __ldg(p0), b0 = __ldg(p0 + 1), a1 = __ldg(p1), b1 = __ldg(p1 + 1);
        uint4 hi, lo;
        split8(a0, b0, hi, lo);
        __nv_bfloat16* o = out + (int64_t)r0 * out_ld + k0;
        *reinterpret_cast<uint4*>(o) = hi;
        if (mode != 0) {
            *reinterpret_cast<uint4*>(o + seg_stride) = (mode == 1) ? hi : lo;
            *reinterpret_cast<uint4*>(o + 2 * seg_stride) = (mode == 1) ? lo : hi;
        }
        if (two) {
            split8(a1, b1, hi, lo);
            o = out + (int64_t)r1 * out_ld + k1;
            *reinterpret_cast<uint4*>(o) = hi;
            if (mode != 0) {
                *reinterpret_cast<uint4*>(o + seg_stride) = (mode == 1) ? hi : lo;
                *reinterpret_cast<uint4*>(o + 2 * seg_stride) = (mode == 1) ? lo : hi;
            }
        }
    }
}

__global__ void __launch_bounds__(256) cast_bf16_kernel(const float* __restrict__ x, int64_t R, int K, int64_t ldx,
                                                        __nv_bfloat16* __restrict__ out, int64_t out_ld, int64_t seg_stride, int mode) {
    const int64_t total = R * (int64_t)K;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / K; const int k = (int)(i - r * K);
        const float v = x[r * ldx + k];
        const __nv_bfloat16 hi = __float2bfloat16_rn(v);
        __nv_bfloat16* o = out + r * out_ld + k;
        o[0] = hi;
        if (mode != 0) {
            const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
            o[seg_stride] = (mode == 1) ? hi : lo;
            o[2 * seg_stride] = (mode == 1) ? lo : hi;
        }
    }
}

// Transposing variant: x fp32 [R, K] (row stride ldx) -> out bf16 [K, segs*R] with out[k*out_ld + g*seg_stride + r];
// 32x32 shared-memory tiles keep both the global read and the global write coalesced.  Modes as cast_bf16_kernel.
__global__ void __launch_bounds__(256) cast_bf16_t_kernel(const float* __restrict__ x, int64_t R, int64_t K, int64_t ldx,
                                                          __nv_bfloat16* __restrict__ out, int64_t out_ld, int64_t seg_stride, int mode) {
    __shared__ float tile[32][33];
    const int64_t r0 = (int64_t)blockIdx.y * 32, k0 = (int64_t)blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t r = r0 + ty + 8 * i, k = k0 + tx;
        tile[ty + 8 * i][tx] = (r < R && k < K) ? x[r * ldx + k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t k = k0 + ty + 8 * i, r = r0 + tx;
        if (k < K && r < R) {
            const float v = tile[tx][ty + 8 * i];
            const __nv_bfloat16 hi = __float2bfloat16_rn(v);
            __nv_bfloat16* o = out + k * out_ld + r;
            o[0] = hi;
            if (mode != 0) {
                const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
                o[seg_stride] = (mode == 1) ? hi : lo;
                o[2 * seg_stride] = (mode == 1) ? lo : hi;
            }
        }
    }
}


// ---------------------------------------------------------------------------------------------------------------
// Dual-layout bf16 operand producer: one read of an fp32 [R, K] matrix (optionally transformed on the fly) feeds BOTH
// K-major operands the tensor-core step needs from it: the row-major copy out[r, k] (A/B operand of the forward GEMM) and
// the transposed copy outT[k, r] (operand of the adjoint GEMM that contracts over r).  Modes as cast_bf16_kernel
// (0 plain, 1 A-split [hi|hi|lo], 2 B-split [hi|lo|hi]); either output may be NULL.  32 x 32 tiles through shared memory keep
// both global writes contiguous.
// Element transforms of the dual-layout producer.  `col(k)` is evaluated once per thread and column (the column of a thread is
// fixed while it walks the rows of a tile), `apply(v, r, ctx)` once per element.
struct IdentityOp {
    struct Ctx {};
    __device__ __forceinline__ Ctx col(int64_t) const { return Ctx(); }
    __device__ __forceinline__ float apply(float v, int64_t, const Ctx&) const { return v; }
};
// dlogits = grad_scale * (softmax(x) - onehot) from per-row statistics (max, 1/sum); logits viewed as [B, S*C], row (b, s).
struct SoftmaxGradOp {
    const float2* stats; const int64_t* labels; int S, C; float grad_scale;
    struct Ctx { int sidx, j; };
    __device__ __forceinline__ Ctx col(int64_t k) const {
        Ctx c; c.sidx = (int)(k / C); c.j = (int)(k - (int64_t)c.sidx * C);
        return c;
    }
    __device__ __forceinline__ float apply(float v, int64_t r, const Ctx& c) const {
        const float2 st = __ldg(stats + r * S + c.sidx);
        const float p = __expf(v - st.x) * st.y;
        return grad_scale * (p - (c.j == (int)__ldg(labels + r) ? 1.f : 0.f));
    }
};

__device__ __forceinline__ void store_split(__nv_bfloat16* o, int64_t seg_stride, int mode, float v) {
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    o[0] = hi;
    if (mode != 0) {
        const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
        o[seg_stride] = (mode == 1) ? hi : lo;
        o[2 * seg_stride] = (mode == 1) ? lo : hi;
    }
}

__device__ __forceinline__ void store_split2(__nv_bfloat16* o, int64_t seg_stride, int mode, float v0, float v1) {
    const __nv_bfloat162 hi = __floats2bfloat162_rn(v0, v1);
    *reinterpret_cast<__nv_bfloat162*>(o) = hi;
    if (mode != 0) {
        const __nv_bfloat162 lo = __floats2bfloat162_rn(v0 - __low2float(hi), v1 - __high2float(hi));
        *reinterpret_cast<__nv_bfloat162*>(o + seg_stride) = (mode == 1) ? hi : lo;
        *reinterpret_cast<__nv_bfloat162*>(o + 2 * seg_stride) = (mode == 1) ? lo : hi;
    }
}

// 64 x 64 tiles, 256 threads (32 x 8): every thread moves column pairs, so the fp32 reads are 8-byte and both bf16 writes
// 4-byte (128 contiguous bytes per warp and row); `vec` = all strides even and bases 4-byte aligned (checked by the host).
template <typename Op>
__global__ void __launch_bounds__(256) cast_dual_kernel(const float* __restrict__ x, int64_t R, int64_t K, int64_t ldx, Op op,
                                                        __nv_bfloat16* __restrict__ out, int64_t out_ld, int64_t seg_stride, int mode,
                                                        __nv_bfloat16* __restrict__ outT, int64_t outT_ld, int64_t segT_stride, int modeT,
                                                        int vec) {
    __shared__ float tile[64][65];
    const int64_t r0 = (int64_t)blockIdx.y * 64, k0 = (int64_t)blockIdx.x * 64;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t k = k0 + 2 * tx;
    const typename Op::Ctx c0 = op.col(k < K ? k : 0), c1 = op.col(k + 1 < K ? k + 1 : 0);
    // the row loads of the whole tile column are issued before any is transformed (8 x 8 bytes in flight per thread)
    float2 xin[8];
    const bool pair = vec && k + 1 < K;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t r = r0 + ty + 8 * i;
        xin[i] = make_float2(0.f, 0.f);
        if (r < R) {
            if (pair) xin[i] = __ldg(reinterpret_cast<const float2*>(x + r * ldx + k));
            else { if (k < K) xin[i].x = __ldg(x + r * ldx + k); if (k + 1 < K) xin[i].y = __ldg(x + r * ldx + k + 1); }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int rr = ty + 8 * i;
        const int64_t r = r0 + rr;
        float v0 = 0.f, v1 = 0.f;
        if (r < R) {
            if (k < K) v0 = op.apply(xin[i].x, r, c0);
            if (k + 1 < K) v1 = op.apply(xin[i].y, r, c1);
            if (out) {
                if (pair) store_split2(out + r * out_ld + k, seg_stride, mode, v0, v1);
                else { if (k < K) store_split(out + r * out_ld + k, seg_stride, mode, v0); if (k + 1 < K) store_split(out + r * out_ld + k + 1, seg_stride, mode, v1); }
            }
        }
        tile[rr][2 * tx] = v0; tile[rr][2 * tx + 1] = v1;
    }
    if (outT == nullptr) return;
    __syncthreads();
    const int64_t r = r0 + 2 * tx;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int kk = ty + 8 * i;
        const int64_t kq = k0 + kk;
        if (kq < K) {
            const float v0 = tile[2 * tx][kk], v1 = tile[2 * tx + 1][kk];
            __nv_bfloat16* o = outT + kq * outT_ld + r;
            if (vec && r + 1 < R) store_split2(o, segT_stride, modeT, v0, v1);
            else {
                if (r < R) store_split(o, segT_stride, modeT, v0);
                if (r + 1 < R) store_split(o + 1, segT_stride, modeT, v1);
            }
        }
    }
}

// Row statistics of the softmax cross-entropy (phase 1 of the tensor-core step's loss): stats[r] = (max, 1 / sum exp(x - max)),
// loss_sum += loss_scale * sum_r (lse_r - x_r[label]).  One warp per row; row r uses labels[r / rows_per_label].
__global__ void __launch_bounds__(256) softmax_stats_kernel(const float* __restrict__ logits, int64_t ld, const int64_t* __restrict__ labels,
                                                            int64_t R, int64_t rows_per_label, int C, float2* __restrict__ stats,
                                                            float* loss_sum, float loss_scale) {
    __shared__ float part[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    float my_loss = 0.f;
    if (row < R) {
        const float* x = logits + row * ld;
        float m = -FLT_MAX, sum = 0.f;
        if (C <= 1024 && (C & 3) == 0 && (ld & 3) == 0 && (reinterpret_cast<uintptr_t>(logits) & 15u) == 0) {
            // the row (<= 1024 classes) is read ONCE, 16 bytes per load, and stays in registers for both reductions
            const float4* x4 = reinterpret_cast<const float4*>(x);
            const int C4 = C >> 2;
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int j = lane + 32 * u;
                v[u] = (j < C4) ? __ldg(x4 + j) : make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX);
                m = fmaxf(m, fmaxf(fmaxf(v[u].x, v[u].y), fmaxf(v[u].z, v[u].w)));
            }
            m = warp_max(m);
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (lane + 32 * u < C4) sum += expf(v[u].x - m) + expf(v[u].y - m) + expf(v[u].z - m) + expf(v[u].w - m);
        } else {
            for (int j = lane; j < C; j += 32) m = fmaxf(m, x[j]);
            m = warp_max(m);
            for (int j = lane; j < C; j += 32) sum += expf(x[j] - m);
        }
        sum = warp_sum(sum);
        my_loss = m + logf(sum) - x[(int)labels[row / rows_per_label]];
        if (lane == 0) stats[row] = make_float2(m, 1.f / sum);
    }
    if (loss_sum) {
        if (lane == 0) part[warp] = (row < R) ? my_loss : 0.f;
        __syncthreads();
        if (threadIdx.x == 0) {
            float t = 0.f;
            for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += part[k];
            atomicAdd(loss_sum, t * loss_scale);
        }
    }
}

// End of the optimisation step in one launch: the learnable inducing row (z_last [C,d], part of the flat parameter buffer, just
// updated by AdamW) is scattered into Z[:, n-1, :], and the two device counters (AdamW step, RNG draw index) advance.
__global__ void __launch_bounds__(256) step_epilogue_kernel(const float* __restrict__ z_last, float* __restrict__ Z, int64_t C, int n, int d,
                                                            int64_t* a, int64_t* b, int64_t by) {
    const int64_t total = C * (int64_t)d;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = i / d; const int k = (int)(i - c * d);
        Z[(c * n + (n - 1)) * d + k] = z_last[i];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { *a += by; *b += by; }
}

// Softmax cross-entropy of the tensor-core step in ONE launch: a CTA owns 64 batch rows of one MC sample s, i.e. the block
// logits[b0 .. b0+64, s*C .. (s+1)*C).  Phase 1: one warp per row reduces max / sum exp (row in registers, C <= 1024) and the CTA adds
// its share of the loss.  Phase 2: the block is streamed again (it is L2 resident: 256 KB per CTA) in 64 x 64 tiles that write
// dlogits = grad_scale (softmax - onehot) as both bf16 operands: row-major out[b, s*C + j] and transposed outT[s*C + j, b].
__global__ void __launch_bounds__(256) softmax_ce_fused_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels, int64_t B,
                                                               int S, int C, float* loss_sum, float loss_scale, float grad_scale,
                                                               __nv_bfloat16* __restrict__ out, int64_t out_ld, int64_t seg_stride, int mode,
                                                               __nv_bfloat16* __restrict__ outT, int64_t outT_ld, int64_t segT_stride, int modeT,
                                                               int vec) {
    __shared__ float tile[64][65];
    __shared__ float s_m[64], s_inv[64];
    __shared__ int s_lab[64];
    __shared__ float s_part[8];
    const int sidx = blockIdx.y;
    const int64_t b0 = (int64_t)blockIdx.x * 64;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t K = (int64_t)S * C;
    const float* base = logits + (int64_t)sidx * C;           // column offset of this sample inside a [B, S*C] row
    // ---- phase 1: row statistics + loss (8 rows per warp)
    float wloss = 0.f;
    for (int rr = warp; rr < 64; rr += 8) {
        const int64_t b = b0 + rr;
        if (b >= B) { if (lane == 0) { s_m[rr] = 0.f; s_inv[rr] = 0.f; s_lab[rr] = -1; } continue; }
        const float* x = base + b * K;
        float m = -FLT_MAX, sum = 0.f;
        if (vec && C <= 1024 && (C & 3) == 0 && (K & 3) == 0) {
            const float4* x4 = reinterpret_cast<const float4*>(x);
            const int C4 = C >> 2;
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int j = lane + 32 * u;
                v[u] = (j < C4) ? __ldg(x4 + j) : make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX);
                m = fmaxf(m, fmaxf(fmaxf(v[u].x, v[u].y), fmaxf(v[u].z, v[u].w)));
            }
            m = warp_max(m);
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (lane + 32 * u < C4) sum += expf(v[u].x - m) + expf(v[u].y - m) + expf(v[u].z - m) + expf(v[u].w - m);
        } else {
            for (int j = lane; j < C; j += 32) m = fmaxf(m, x[j]);
            m = warp_max(m);
            for (int j = lane; j < C; j += 32) sum += expf(x[j] - m);
        }
        sum = warp_sum(sum);
        const int lab = (int)labels[b];
        if (lane == 0) { s_m[rr] = m; s_inv[rr] = 1.f / sum; s_lab[rr] = lab; wloss += m + logf(sum) - x[lab]; }
    }
    if (lane == 0) s_part[warp] = wloss;
    __syncthreads();
    if (threadIdx.x == 0 && loss_sum) {
        float t = 0.f;
        for (int k = 0; k < 8; ++k) t += s_part[k];
        atomicAdd(loss_sum, t * loss_scale);
    }
    // ---- phase 2: dlogits into both operand layouts, 64 x 64 tiles (thread = column pair, eight rows)
    const int tx = lane, ty = warp;
    for (int j0 = 0; j0 < C; j0 += 64) {
        const int j = j0 + 2 * tx;
        const int64_t k = (int64_t)sidx * C + j;               // column in the [B, S*C] matrix
        const bool pair = vec && j + 1 < C;
        float2 xin[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int64_t b = b0 + ty + 8 * i;
            xin[i] = make_float2(0.f, 0.f);
            if (b < B) {
                if (pair) xin[i] = __ldg(reinterpret_cast<const float2*>(logits + b * K + k));
                else { if (j < C) xin[i].x = __ldg(logits + b * K + k); if (j + 1 < C) xin[i].y = __ldg(logits + b * K + k + 1); }
            }
        }
        __syncthreads();                                       // the previous tile has been written out
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int rr = ty + 8 * i;
            const int64_t b = b0 + rr;
            float v0 = 0.f, v1 = 0.f;
            if (b < B) {
                const float m = s_m[rr], inv = s_inv[rr];
                const int lab = s_lab[rr];
                if (j < C) v0 = grad_scale * (__expf(xin[i].x - m) * inv - (j == lab ? 1.f : 0.f));
                if (j + 1 < C) v1 = grad_scale * (__expf(xin[i].y - m) * inv - (j + 1 == lab ? 1.f : 0.f));
                if (out) {
                    if (pair) store_split2(out + b * out_ld + k, seg_stride, mode, v0, v1);
                    else { if (j < C) store_split(out + b * out_ld + k, seg_stride, mode, v0); if (j + 1 < C) store_split(out + b * out_ld + k + 1, seg_stride, mode, v1); }
                }
            }
            tile[rr][2 * tx] = v0; tile[rr][2 * tx + 1] = v1;
        }
        __syncthreads();
        if (outT) {
            const int64_t b = b0 + 2 * tx;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int jj = j0 + ty + 8 * i;
                if (jj < C) {
                    const float v0 = tile[2 * tx][ty + 8 * i], v1 = tile[2 * tx + 1][ty + 8 * i];
                    __nv_bfloat16* o = outT + ((int64_t)sidx * C + jj) * outT_ld + b;
                    if (vec && b + 1 < B) store_split2(o, segT_stride, modeT, v0, v1);
                    else { if (b < B) store_split(o, segT_stride, modeT, v0); if (b + 1 < B) store_split(o + 1, segT_stride, modeT, v1); }
                }
            }
        }
    }
}

__global__ void increment2_kernel(int64_t* a, int64_t* b, int64_t by) { if (threadIdx.x == 0 && blockIdx.x == 0) { *a += by; *b += by; } }

__global__ void increment_kernel(int64_t* p, int64_t by) { if (threadIdx.x == 0 && blockIdx.x == 0) *p += by; }


// out[k][r] = x[r][k] for an fp32 [R, K] matrix (row pitches ldx / out_ld): 32 x 32 tiles through shared memory, both sides coalesced.
// Used to hand the SMALL operand of a long-K TF32 GEMM to the tensor cores K-major: an MN-major B operand (the 256-wide tile) costs
// the tcgen05 pipeline ~35 % of its rate (tools/micro/tf32_layouts.py), a transposed copy of a [S*C, D] matrix costs microseconds.
__global__ void __launch_bounds__(256) transpose_f32_kernel(const float* __restrict__ x, int64_t R, int64_t K, int64_t ldx,
                                                            float* __restrict__ out, int64_t out_ld) {
    __shared__ float tile[32][33];
    const int64_t r0 = (int64_t)blockIdx.y * 32, k0 = (int64_t)blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
        const int64_t r = r0 + ty + j, k = k0 + tx;
        tile[ty + j][tx] = (r < R && k < K) ? x[r * ldx + k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
        const int64_t k = k0 + ty + j, r = r0 + tx;
        if (k < K && r < R) out[k * out_ld + r] = tile[tx][ty + j];
    }
}


// Tail of the single-GPU optimisation step in ONE launch (was: KL sum -> AdamW of the gp_weighter group -> scatter of the learnable
// inducing row + counters: three dependent launches behind the adjoint kernel, ~19 us on the critical path).
//   * AdamW (same arithmetic as adamw_kernel) over p / g / m / v [n], 16 bytes per thread and step;
//   * elements of the z_last segment [z_off, z_off + C d) are also written to Z[c, nrows - 1, k] (gp_template_weigher.py:72-79);
//   * the LAST CTA of the grid adds kl_scale * sum(kl) to loss[0] (one writer, fixed summation order);
//   * whoever takes the last ticket advances the AdamW step and the RNG draw index (every CTA has read the step by then).
__global__ void __launch_bounds__(256) adamw_tail_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                         float* __restrict__ v, int64_t n, const float* __restrict__ lr_ptr, float b1, float b2,
                                                         float eps, float wd, int64_t* step_ptr, int64_t z_off, int64_t zn, int nrows, int d,
                                                         float* __restrict__ Z, const float* __restrict__ kl, int64_t kl_n, float kl_scale,
                                                         float* loss, int64_t* counter_b, int64_t by, unsigned int* ticket) {
    __shared__ float red[32];
    const float lr = *lr_ptr;
    const float t = (float)(*step_ptr);
    const float bc1 = 1.f - powf(b1, t), bc2 = 1.f - powf(b2, t);
    const float step_size = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2), decay = 1.f - lr * wd;
    const unsigned workers = gridDim.x - 1;
    if (blockIdx.x == workers) {
        if (kl != nullptr) {
            float q = 0.f;
            for (int64_t i = threadIdx.x; i < kl_n; i += blockDim.x) q += kl[i];
            const float tot = block_sum(q, red);
            if (threadIdx.x == 0) loss[0] += tot * kl_scale;
        }
    } else {
        const int64_t n4 = n >> 2;
        for (int64_t i4 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i4 < n4; i4 += (int64_t)workers * blockDim.x) {
            const float4 gi = reinterpret_cast<const float4*>(g)[i4];
            float4 pi = reinterpret_cast<float4*>(p)[i4], mi = reinterpret_cast<float4*>(m)[i4], vi = reinterpret_cast<float4*>(v)[i4];
#define CLIPGP_ADAMW1(c)                                                        \
            mi.c = b1 * mi.c + (1.f - b1) * gi.c;                               \
            vi.c = b2 * vi.c + (1.f - b2) * gi.c * gi.c;                        \
            pi.c = pi.c * decay - step_size * mi.c / (sqrtf(vi.c) * inv_sqrt_bc2 + eps);
            CLIPGP_ADAMW1(x) CLIPGP_ADAMW1(y) CLIPGP_ADAMW1(z) CLIPGP_ADAMW1(w)
#undef CLIPGP_ADAMW1
            reinterpret_cast<float4*>(p)[i4] = pi; reinterpret_cast<float4*>(m)[i4] = mi; reinterpret_cast<float4*>(v)[i4] = vi;
            const int64_t i = i4 << 2;
            if (Z != nullptr && i + 3 >= z_off && i < z_off + zn) {
                const float pv[4] = {pi.x, pi.y, pi.z, pi.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int64_t r = i + e - z_off;
                    if (r >= 0 && r < zn) { const int64_t c = r / d; Z[(c * nrows + (nrows - 1)) * d + (r - c * d)] = pv[e]; }
                }
            }
        }
        if (blockIdx.x == 0) {
            for (int64_t i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) {     // n % 4 tail
                const float gi = g[i];
                const float mi = b1 * m[i] + (1.f - b1) * gi, vi = b2 * v[i] + (1.f - b2) * gi * gi;
                const float pi = p[i] * decay - step_size * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
                m[i] = mi; v[i] = vi; p[i] = pi;
                const int64_t r = i - z_off;
                if (Z != nullptr && r >= 0 && r < zn) { const int64_t c = r / d; Z[(c * nrows + (nrows - 1)) * d + (r - c * d)] = pi; }
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(ticket, 1u) == gridDim.x - 1u) {
            *step_ptr += by; *counter_b += by;
            *ticket = 0u;
        }
    }
}

}  // namespace clipgp

using namespace clipgp;

extern "C" int clipgp_rownorm_forward(const float* x, int64_t R, int64_t D, float* y, float* inv_norm, void* y_bf16, void* stream) {
    CLIPGP_REQUIRE(R >= 0 && D >= 1 && D < (1ll << 31), "rownorm_forward: bad shape");
    if (R == 0) return CLIPGP_OK;
    CLIPGP_REQUIRE(x && (y || y_bf16), "rownorm_forward: NULL pointer");
    const int64_t blocks = (R + 7) / 8;
    rownorm_fwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, R, (int)D, y, inv_norm, (__nv_bfloat16*)y_bf16);
    return check_launch("rownorm_fwd_kernel");
}

extern "C" int clipgp_rownorm_cast(const float* x, int64_t R, int64_t D, float* y, float* inv_norm, void* out_bf16, int64_t out_ld,
                                   int64_t seg_stride, int mode, void* stream) {
    CLIPGP_REQUIRE(R >= 0 && D >= 4 && (D % 4) == 0 && D < (1ll << 31), "rownorm_cast: D must be a positive multiple of 4");
    CLIPGP_REQUIRE(mode >= 0 && mode <= 2, "rownorm_cast: mode must be 0, 1 or 2");
    if (R == 0) return CLIPGP_OK;
    CLIPGP_REQUIRE(x && out_bf16, "rownorm_cast: NULL pointer");
    CLIPGP_REQUIRE(((out_ld | seg_stride) & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15u) == 0 && (reinterpret_cast<uintptr_t>(out_bf16) & 7u) == 0 &&
                       (y == nullptr || (reinterpret_cast<uintptr_t>(y) & 15u) == 0),
                   "rownorm_cast: operands must be 16-byte (fp32) / 8-byte (bf16) aligned with strides that are multiples of 4");
    const int64_t blocks = (R + 7) / 8;
    rownorm_cast_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, R, (int)D, y, inv_norm, (__nv_bfloat16*)out_bf16, out_ld,
                                                                           seg_stride, mode);
    return check_launch("rownorm_cast_kernel");
}

extern "C" int clipgp_rownorm_backward(const float* dy, const float* y, const float* inv_norm, int64_t R, int64_t D, float* dx,
                                       void* stream) {
    CLIPGP_REQUIRE(R >= 0 && D >= 1 && D < (1ll << 31), "rownorm_backward: bad shape");
    if (R == 0) return CLIPGP_OK;
    CLIPGP_REQUIRE(dy && y && inv_norm && dx, "rownorm_backward: NULL pointer");
    const int64_t blocks = (R + 7) / 8;
    rownorm_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(dy, y, inv_norm, R, (int)D, dx);
    return check_launch("rownorm_bwd_kernel");
}

extern "C" int clipgp_softmax_ce(const float* logits, int64_t ld, const int64_t* labels, int64_t R, int64_t rows_per_label,
                                 int64_t C, float* loss_rows, float* loss_sum, float loss_scale, float* dlogits, int64_t ldd,
                                 float grad_scale, void* stream) {
    CLIPGP_REQUIRE(R >= 0 && C >= 1 && C < (1ll << 31) && rows_per_label >= 1, "softmax_ce: bad shape");
    if (R == 0) return CLIPGP_OK;
    CLIPGP_REQUIRE(logits && labels && ld >= C, "softmax_ce: bad input");
    CLIPGP_REQUIRE(!dlogits || ldd >= C, "softmax_ce: ldd < C");
    const int64_t blocks = (R + 7) / 8;
    softmax_ce_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(logits, ld, labels, R, rows_per_label, (int)C, loss_rows,
                                                                         loss_sum, loss_scale, dlogits, ldd, grad_scale);
    return check_launch("softmax_ce_kernel");
}

extern "C" int clipgp_l2_identity(const float* W, int64_t D, float coef, float* dW, float* loss_sum, void* stream) {
    CLIPGP_REQUIRE(D >= 1 && W, "l2_identity: bad input");
    int64_t blocks = (D * D + 255) / 256;
    if (blocks > 1184) blocks = 1184;
    l2_identity_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(W, (int)D, coef, dW, loss_sum);
    return check_launch("l2_identity_kernel");
}

extern "C" int clipgp_adamw_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                                 float eps, float weight_decay, const int64_t* step, void* stream) {
    CLIPGP_REQUIRE(n >= 0, "adamw_step: n < 0");
    if (n == 0) return CLIPGP_OK;
    CLIPGP_REQUIRE(p && g && m && v && step, "adamw_step: NULL pointer");
    int64_t blocks = (n + 255) / 256;
    const int64_t cap = (int64_t)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    adamw_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, step, nullptr);
    return check_launch("adamw_kernel");
}

extern "C" int clipgp_adamw_step_lrptr(float* p, const float* g, float* m, float* v, int64_t n, const float* lr_dev, float beta1,
                                       float beta2, float eps, float weight_decay, const int64_t* step, void* stream) {
    CLIPGP_REQUIRE(n >= 0, "adamw_step_lrptr: n < 0");
    if (n == 0) return CLIPGP_OK;
    CLIPGP_REQUIRE(p && g && m && v && step && lr_dev, "adamw_step_lrptr: NULL pointer");
    int64_t blocks = (n + 255) / 256;
    const int64_t cap = (int64_t)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    adamw_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, 0.f, beta1, beta2, eps, weight_decay, step, lr_dev);
    return check_launch("adamw_kernel");
}

extern "C" int clipgp_adamw_step_cast(float* p, const float* g, float* m, float* v, int64_t R, int64_t K, const float* lr_dev, float beta1,
                                      float beta2, float eps, float weight_decay, const int64_t* step, void* out_bf16, int64_t out_ld,
                                      int64_t seg_stride, int mode, void* stream) {
    CLIPGP_REQUIRE(R >= 0 && K >= 4 && (K % 4) == 0 && K < (1ll << 31), "adamw_step_cast: K must be a positive multiple of 4");
    CLIPGP_REQUIRE(mode >= 0 && mode <= 2, "adamw_step_cast: mode must be 0, 1 or 2");
    if (R == 0) return CLIPGP_OK;
    CLIPGP_REQUIRE(p && g && m && v && step && lr_dev && out_bf16, "adamw_step_cast: NULL pointer");
    CLIPGP_REQUIRE(((out_ld | seg_stride) & 3) == 0 && (reinterpret_cast<uintptr_t>(out_bf16) & 7u) == 0 &&
                   ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15u) == 0,
                   "adamw_step_cast: misaligned buffers / strides");
    const int64_t n4 = R * (K / 4);
    int64_t blocks = (n4 + 255) / 256;
    const int64_t cap = (int64_t)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    adamw_cast_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n4, (int)(K / 4), lr_dev, beta1, beta2, eps, weight_decay, step,
                                                                         (__nv_bfloat16*)out_bf16, out_ld, seg_stride, mode);
    return check_launch("adamw_cast_kernel");
}

extern "C" int clipgp_increment(int64_t* counter, int64_t by, void* stream) {
    CLIPGP_REQUIRE(counter, "increment: NULL pointer");
    increment_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(counter, by);
    return check_launch("increment_kernel");
}

extern "C" int clipgp_sum_accumulate(const float* x, int64_t n, float scale, float* out, void* stream) {
    CLIPGP_REQUIRE(n >= 0 && out, "sum_accumulate: bad input");
    if (n == 0) return CLIPGP_OK;
    CLIPGP_REQUIRE(x, "sum_accumulate: NULL input");
    int64_t blocks = (n + 255) / 256;
    if (blocks > 148) blocks = 148;
    sum_accumulate_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, n, scale, out);
    return check_launch("sum_accumulate_kernel");
}

extern "C" int clipgp_cast_bf16(const float* x, int64_t R, int64_t K, int64_t ldx, void* out, int64_t out_ld, int64_t seg_stride,
                                int mode, void* stream) {
    CLIPGP_REQUIRE(R >= 0 && K >= 1 && K < (1ll << 31) && ldx >= K, "cast_bf16: bad shape");
    CLIPGP_REQUIRE(mode >= 0 && mode <= 2, "cast_bf16: mode must be 0 (plain), 1 (A split) or 2 (B split)");
    if (R == 0) return CLIPGP_OK;
    CLIPGP_REQUIRE(x && out, "cast_bf16: NULL pointer");
    const int64_t cap = (int64_t)num_sms() * 16;
    if (((K | ldx | out_ld | seg_stride) & 7) == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0 &&
        R * (K / 8) < (1ll << 31) - (1ll << 24)) {
        const int64_t total = R * (K / 8);
        int64_t blocks = (total + 511) / 512;
        if (blocks > cap) blocks = cap;
        cast_bf16_vec8_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, (unsigned)total, (unsigned)(K / 8), ldx, (__nv_bfloat16*)out, out_ld,
                                                                                   seg_stride, mode);
        return check_launch("cast_bf16_vec8_kernel");
    }
    if (((K | ldx | out_ld | seg_stride) & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15u) == 0 && (reinterpret_cast<uintptr_t>(out) & 7u) == 0) {
        int64_t blocks = (R * (K / 4) + 255) / 256;
        if (blocks > cap) blocks = cap;
        cast_bf16_vec4_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, R, (int)K, ldx, (__nv_bfloat16*)out, out_ld, seg_stride, mode);
        return check_launch("cast_bf16_vec4_kernel");
    }
    int64_t blocks = (R * K + 255) / 256;
    if (blocks > cap) blocks = cap;
    cast_bf16_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, R, (int)K, ldx, (__nv_bfloat16*)out, out_ld, seg_stride, mode);
    return check_launch("cast_bf16_kernel");
}

extern "C" int clipgp_cast_bf16_transpose(const float* x, int64_t R, int64_t K, int64_t ldx, void* out, int64_t out_ld,
                                          int64_t seg_stride, int mode, void* stream) {
    CLIPGP_REQUIRE(R >= 0 && K >= 0 && ldx >= K, "cast_bf16_transpose: bad shape");
    CLIPGP_REQUIRE(mode >= 0 && mode <= 2, "cast_bf16_transpose: mode must be 0, 1 or 2");
    if (R == 0 || K == 0) return CLIPGP_OK;
    CLIPGP_REQUIRE(x && out, "cast_bf16_transpose: NULL pointer");
    dim3 grid((unsigned)((K + 31) / 32), (unsigned)((R + 31) / 32));
    CLIPGP_REQUIRE(grid.y <= 65535, "cast_bf16_transpose: R too large");
    cast_bf16_t_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, R, K, ldx, (__nv_bfloat16*)out, out_ld, seg_stride, mode);
    return check_launch("cast_bf16_t_kernel");
}

static int check_dual(const char* who, int64_t R, int64_t K, int64_t ldx, const void* x, const void* out, const void* outT, int mode, int modeT) {
    CLIPGP_REQUIRE(R >= 0 && K >= 0 && ldx >= K, "%s: bad shape", who);
    CLIPGP_REQUIRE(mode >= 0 && mode <= 2 && modeT >= 0 && modeT <= 2, "%s: modes must be 0, 1 or 2", who);
    CLIPGP_REQUIRE(R == 0 || K == 0 || (x && (out || outT)), "%s: NULL pointer", who);
    CLIPGP_REQUIRE((R + 63) / 64 <= 65535, "%s: R too large", who);
    return CLIPGP_OK;
}

static int dual_vec_ok(const void* x, int64_t ldx, const void* out, int64_t out_ld, int64_t seg, const void* outT, int64_t outT_ld,
                       int64_t segT) {
    if ((ldx & 1) || (reinterpret_cast<uintptr_t>(x) & 7u)) return 0;
    if (out && ((out_ld & 1) || (seg & 1) || (reinterpret_cast<uintptr_t>(out) & 3u))) return 0;
    if (outT && ((outT_ld & 1) || (segT & 1) || (reinterpret_cast<uintptr_t>(outT) & 3u))) return 0;
    return 1;
}

extern "C" int clipgp_cast_bf16_dual(const float* x, int64_t R, int64_t K, int64_t ldx, void* out, int64_t out_ld, int64_t seg_stride,
                                     int mode, void* outT, int64_t outT_ld, int64_t segT_stride, int modeT, void* stream) {
    int rc = check_dual("cast_bf16_dual", R, K, ldx, x, out, outT, mode, modeT);
    if (rc != CLIPGP_OK) return rc;
    if (R == 0 || K == 0) return CLIPGP_OK;
    dim3 grid((unsigned)((K + 63) / 64), (unsigned)((R + 63) / 64));
    cast_dual_kernel<IdentityOp><<<grid, 256, 0, (cudaStream_t)stream>>>(x, R, K, ldx, IdentityOp(), (__nv_bfloat16*)out, out_ld, seg_stride,
                                                                         mode, (__nv_bfloat16*)outT, outT_ld, segT_stride, modeT,
                                                                         dual_vec_ok(x, ldx, out, out_ld, seg_stride, outT, outT_ld, segT_stride));
    return check_launch("cast_dual_kernel");
}

extern "C" int clipgp_softmax_ce_stats(const float* logits, int64_t ld, const int64_t* labels, int64_t R, int64_t rows_per_label,
                                       int64_t C, float* stats, float* loss_sum, float loss_scale, void* stream) {
    CLIPGP_REQUIRE(R >= 0 && C >= 1 && C < (1ll << 31) && rows_per_label >= 1, "softmax_ce_stats: bad shape");
    if (R == 0) return CLIPGP_OK;
    CLIPGP_REQUIRE(logits && labels && stats && ld >= C, "softmax_ce_stats: bad input");
    const int64_t blocks = (R + 7) / 8;
    softmax_stats_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(logits, ld, labels, R, rows_per_label, (int)C,
                                                                            reinterpret_cast<float2*>(stats), loss_sum, loss_scale);
    return check_launch("softmax_stats_kernel");
}

extern "C" int clipgp_softmax_grad_bf16_dual(const float* logits, const float* stats, const int64_t* labels, int64_t B, int64_t S,
                                             int64_t C, float grad_scale, void* out, int64_t out_ld, int64_t seg_stride, int mode,
                                             void* outT, int64_t outT_ld, int64_t segT_stride, int modeT, void* stream) {
    const int64_t K = S * C;
    int rc = check_dual("softmax_grad_bf16_dual", B, K, K, logits, out, outT, mode, modeT);
    if (rc != CLIPGP_OK) return rc;
    if (B == 0 || K == 0) return CLIPGP_OK;
    CLIPGP_REQUIRE(stats && labels && S >= 1 && C >= 1 && C < (1ll << 31), "softmax_grad_bf16_dual: bad input");
    SoftmaxGradOp op{reinterpret_cast<const float2*>(stats), labels, (int)S, (int)C, grad_scale};
    dim3 grid((unsigned)((K + 63) / 64), (unsigned)((B + 63) / 64));
    cast_dual_kernel<SoftmaxGradOp><<<grid, 256, 0, (cudaStream_t)stream>>>(logits, B, K, K, op, (__nv_bfloat16*)out, out_ld, seg_stride, mode,
                                                                            (__nv_bfloat16*)outT, outT_ld, segT_stride, modeT,
                                                                            dual_vec_ok(logits, K, out, out_ld, seg_stride, outT, outT_ld, segT_stride));
    return check_launch("cast_dual_kernel(softmax_grad)");
}

extern "C" int clipgp_increment2(int64_t* a, int64_t* b, int64_t by, void* stream) {
    CLIPGP_REQUIRE(a && b, "increment2: NULL pointer");
    increment2_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(a, b, by);
    return check_launch("increment2_kernel");
}

extern "C" int clipgp_step_epilogue(const float* z_last, float* Z, int64_t C, int64_t n, int64_t d, int64_t* counter_a,
                                    int64_t* counter_b, int64_t by, void* stream) {
    CLIPGP_REQUIRE(C >= 0 && n >= 1 && d >= 1 && n < (1ll << 31) && d < (1ll << 31), "step_epilogue: bad shape");
    CLIPGP_REQUIRE(z_last && Z && counter_a && counter_b, "step_epilogue: NULL pointer");
    int64_t blocks = (C * d + 255) / 256;
    if (blocks < 1) blocks = 1;
    if (blocks > 1184) blocks = 1184;
    step_epilogue_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(z_last, Z, C, (int)n, (int)d, counter_a, counter_b, by);
    return check_launch("step_epilogue_kernel");
}

extern "C" int clipgp_softmax_ce_bf16_dual(const float* logits, const int64_t* labels, int64_t B, int64_t S, int64_t C, float* loss_sum,
                                           float loss_scale, float grad_scale, void* out, int64_t out_ld, int64_t seg_stride, int mode,
                                           void* outT, int64_t outT_ld, int64_t segT_stride, int modeT, void* stream) {
    const int64_t K = S * C;
    int rc = check_dual("softmax_ce_bf16_dual", B, K, K, logits, out, outT, mode, modeT);
    if (rc != CLIPGP_OK) return rc;
    if (B == 0 || K == 0) return CLIPGP_OK;
    CLIPGP_REQUIRE(labels && S >= 1 && S <= 65535 && C >= 1 && C < (1ll << 31), "softmax_ce_bf16_dual: bad input");
    dim3 grid((unsigned)((B + 63) / 64), (unsigned)S);
    const int vec = dual_vec_ok(logits, K, out, out_ld, seg_stride, outT, outT_ld, segT_stride) && (C % 2 == 0) &&
                    ((reinterpret_cast<uintptr_t>(logits) & 15u) == 0);
    softmax_ce_fused_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(logits, labels, B, (int)S, (int)C, loss_sum, loss_scale, grad_scale,
                                                                   (__nv_bfloat16*)out, out_ld, seg_stride, mode, (__nv_bfloat16*)outT, outT_ld,
                                                                   segT_stride, modeT, vec);
    return check_launch("softmax_ce_fused_kernel");
}

extern "C" int clipgp_transpose_f32(const float* x, int64_t R, int64_t K, int64_t ldx, float* out, int64_t out_ld, void* stream) {
    CLIPGP_REQUIRE(R >= 0 && K >= 0 && ldx >= K && out_ld >= R, "transpose_f32: bad shape / pitches");
    if (R == 0 || K == 0) return CLIPGP_OK;
    CLIPGP_REQUIRE(x && out, "transpose_f32: NULL pointer");
    const int64_t gy = (R + 31) / 32, gx = (K + 31) / 32;
    CLIPGP_REQUIRE(gy <= 65535, "transpose_f32: too many rows for one launch");
    transpose_f32_kernel<<<dim3((unsigned)gx, (unsigned)gy), 256, 0, (cudaStream_t)stream>>>(x, R, K, ldx, out, out_ld);
    return check_launch("transpose_f32_kernel");
}

extern "C" int clipgp_adamw_tail(float* p, const float* g, float* m, float* v, int64_t n, const float* lr_dev, float beta1, float beta2,
                                 float eps, float weight_decay, int64_t* step, int64_t z_off, int64_t C, int64_t nrows, int64_t d, float* Z,
                                 const float* kl, int64_t kl_n, float kl_scale, float* loss, int64_t* counter_b, int64_t by,
                                 unsigned int* ticket, void* stream) {
    CLIPGP_REQUIRE(n >= 0 && p && g && m && v && lr_dev && step && counter_b && ticket, "adamw_tail: NULL pointer / bad size");
    CLIPGP_REQUIRE(((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15u) == 0,
                   "adamw_tail: p / g / m / v must be 16-byte aligned");
    CLIPGP_REQUIRE(Z == nullptr || (z_off >= 0 && C >= 0 && nrows >= 1 && d >= 1 && z_off + C * d <= n), "adamw_tail: z_last segment outside the group");
    CLIPGP_REQUIRE(kl == nullptr || (loss != nullptr && kl_n >= 0), "adamw_tail: kl without loss");
    int64_t blocks = ((n >> 2) + 255) / 256;
    const int64_t cap = (int64_t)num_sms() * 4;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    adamw_tail_kernel<<<(unsigned)blocks + 1u, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr_dev, beta1, beta2, eps, weight_decay, step, z_off,
                                                                              C * d, (int)nrows, (int)d, Z, kl, kl_n, kl_scale, loss, counter_b, by, ticket);
    return check_launch("adamw_tail_kernel");
}
