// Shared helpers for the clipgp sm_100a kernels (error reporting, warp/block reductions, Philox).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/clipgp.h"

namespace clipgp {

// ---- error channel (int status + last_error(); the Python side raises RuntimeError) ----
void set_error(const char* fmt, ...);
int  check_launch(const char* what);   // cudaGetLastError -> status

#define CLIPGP_REQUIRE(cond, ...)                                   \
    do {                                                            \
        if (!(cond)) {                                              \
            ::clipgp::set_error(__VA_ARGS__);                       \
            return CLIPGP_ERR_INVALID;                              \
        }                                                           \
    } while (0)

#define CLIPGP_CUDA(call)                                                                   \
    do {                                                                                    \
        cudaError_t e__ = (call);                                                           \
        if (e__ != cudaSuccess) {                                                           \
            ::clipgp::set_error("%s failed: %s", #call, cudaGetErrorString(e__));           \
            return CLIPGP_ERR_CUDA;                                                         \
        }                                                                                   \
    } while (0)

int num_sms();

// ---- device helpers ----
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Block-wide sum; `red` must hold >= 32 floats of shared memory. All threads get the result.
__device__ __forceinline__ float block_sum(float v, float* red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    float r = (lane < nw) ? red[lane] : 0.f;
    r = warp_sum(r);
    return r;
}
__device__ __forceinline__ double block_sum(double v, double* red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    double r = (lane < nw) ? red[lane] : 0.0;
    r = warp_sum(r);
    return r;
}

// Ask the memory system to bring [p, p + bytes) into L2 (one `cp.async.bulk.prefetch.L2` per 32 KB, issued by the calling thread;
// no register or shared-memory cost, nothing to wait for).  The per-class GP kernels call it at the top of a class CTA for the
// operands of their LATER streaming phases (text bank rows, inducing points, saved factors): the DRAM latency of those streams
// then overlaps the sequential factorisation phases instead of being paid chunk by chunk when the phase starts.
__device__ __forceinline__ void l2_prefetch(const void* p, size_t bytes) {
    uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uintptr_t end = (a + bytes + 15) & ~(uintptr_t)15;
    a &= ~(uintptr_t)15;
    while (a < end) {
        const unsigned n = (unsigned)((end - a) < 32768 ? (end - a) : 32768);
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a), "r"(n) : "memory");
        a += n;
    }
}

__device__ __forceinline__ float softplusf(float x) {
    // torch.nn.functional.softplus (beta=1, threshold=20)
    return x > 20.f ? x : log1pf(expf(x));
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// ---- Philox4x32-10 counter RNG + Box-Muller (perf-mode base noise; oracle/philox.py restates it) ----
__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                       uint32_t k0, uint32_t k1, uint32_t out[4]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// Standard normal for element index `idx` of draw number `step` under `seed`.
// u = (x + 0.5) / 2^32 in (0,1);  z = sqrt(-2 ln u0) cos(2 pi u1).
__device__ __forceinline__ float philox_normal(uint64_t seed, uint64_t step, uint64_t idx) {
    uint32_t o[4];
    philox4x32_10((uint32_t)idx, (uint32_t)(idx >> 32), (uint32_t)step, (uint32_t)(step >> 32),
                  (uint32_t)seed, (uint32_t)(seed >> 32), o);
    const float u0 = ((float)o[0] + 0.5f) * 2.3283064365386963e-10f;
    const float u1 = ((float)o[1] + 0.5f) * 2.3283064365386963e-10f;
    // (float)o[0] can round up to 2^32 -> u0 == 1.0f exactly -> log = 0, fine; never 0.
    return sqrtf(-2.f * logf(u0)) * cospif(2.f * u1);
}

}  // namespace clipgp
