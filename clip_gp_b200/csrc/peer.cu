// Multi-GPU optimiser step over NVLink peer memory: gradient reduce-scatter + AdamW + parameter all-gather in ONE kernel.
//
// Replaces, for the data-parallel GP-Adapter step (engine.py, shard = "batch"), the sequence
//     ncclAllReduce(flat gradient, 7.6 MB)  ->  adamw(W)  ->  adamw(gp)          (adapter.py:537-549 on every rank)
// Every rank owns 1/world of the flat parameter vector (and ONLY that slice of the Adam moments matters on it).  The kernel
//   (A) tells every peer "my gradient buffer is complete" (one release store per peer) and waits for the peers' flags,
//   (B) reads its slice of all `world` gradient buffers through peer pointers (16-byte loads, all peers in flight at once),
//       sums them in rank order (deterministic, identical on every rank because each element has exactly one owner),
//       applies AdamW to its slice and stores the new parameters into EVERY rank's parameter buffer,
//   (C) after the last CTA's stores are fenced, tells every peer "my slice is written everywhere" and waits for theirs.
// Bytes over NVLink per rank and step: (world-1)/world of the vector in and out (6.6 MB each way at world = 8) instead of the
// ring / tree traffic of an all-reduce, no staging copies, and the optimiser pass over HBM shrinks to 1/world.
// The buffers are plain cudaMalloc allocations exported with CUDA IPC (one process per GPU, torch.distributed only carries the
// 64-byte handles at set-up).  Spin loops carry a wall-clock timeout: a missing peer sets `status` instead of hanging the GPU.
#include "common.cuh"

namespace clipgp {

__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ float4 ld_sys_f4(const float* p) {
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float ld_sys_f1(const float* p) {
    float v;
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_sys_f4(float* p, float4 v) {
    asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void st_sys_f1(float* p, float v) { asm volatile("st.relaxed.sys.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory"); }
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// wait until *flag >= want; false on timeout
__device__ __forceinline__ bool spin_until(const unsigned long long* flag, unsigned long long want, unsigned long long timeout_ns) {
    if (ld_acquire_sys_u64(flag) >= want) return true;
    const unsigned long long t0 = globaltimer_ns();
    for (;;) {
        for (int i = 0; i < 64; ++i)
            if (ld_acquire_sys_u64(flag) >= want) return true;
        if (globaltimer_ns() - t0 > timeout_ns) return false;
    }
}

struct AdamScal { float step0, step1, decay0, decay1, inv_sqrt_bc2, b1, b2, eps; };

__device__ __forceinline__ float adam1(float g, float p, float& m, float& v, const AdamScal& s, bool grp0) {
    m = s.b1 * m + (1.f - s.b1) * g;
    v = s.b2 * v + (1.f - s.b2) * g * g;
    return p * (grp0 ? s.decay0 : s.decay1) - (grp0 ? s.step0 : s.step1) * m / (sqrtf(v) * s.inv_sqrt_bc2 + s.eps);
}

__global__ void __launch_bounds__(256) peer_adamw_kernel(const clipgp_peer_args a) {
    const int W = a.world, r = a.rank, tid = threadIdx.x;
    __shared__ unsigned long long s_epoch;
    __shared__ int s_last, s_fail;
    if (tid == 0) { s_epoch = *reinterpret_cast<volatile unsigned long long*>(a.local) + 1ull; s_fail = 0; }
    __syncthreads();
    const unsigned long long e = s_epoch;
    unsigned long long* mine = a.flags[r];
    // ---- (A) my gradients are complete (written by earlier kernels of this stream); wait for everybody's
    if (blockIdx.x == 0 && a.kl != nullptr) {
        // this rank's KL share joins its loss slot first (adapter.py:462-465; saves the separate reduction launch in front of this kernel)
        __shared__ float red[32];
        float q = 0.f;
        for (int64_t i = tid; i < a.kl_n; i += blockDim.x) q += a.kl[i];
        const float tot = block_sum(q, red);
        if (tid == 0) { float* slot = const_cast<float*>(a.g[r]) + a.n; *slot += tot * a.kl_scale; __threadfence_system(); }
        __syncthreads();
    }
    if (blockIdx.x == 0 && tid < W) st_release_sys_u64(a.flags[tid] + r, e);
    if (tid < W && !spin_until(mine + tid, e, a.timeout_ns)) s_fail = 1;
    __syncthreads();
    if (s_fail && tid == 0) atomicExch(a.status, 1);

    // ---- (B) my slice: [lo4, hi4) in units of four floats; the last rank also takes the n % 4 tail
    AdamScal s;
    {
        const float t = (float)(*a.step);
        const float bc1 = 1.f - powf(a.beta1, t), bc2 = 1.f - powf(a.beta2, t);
        const float lr0 = a.lr_dev[0], lr1 = a.lr_dev[1];
        s.step0 = lr0 / bc1; s.step1 = lr1 / bc1; s.decay0 = 1.f - lr0 * a.weight_decay; s.decay1 = 1.f - lr1 * a.weight_decay;
        s.inv_sqrt_bc2 = rsqrtf(bc2); s.b1 = a.beta1; s.b2 = a.beta2; s.eps = a.eps;
    }
    const int64_t n4 = a.n >> 2, per = (n4 + W - 1) / W;
    const int64_t lo4 = per * r, hi4 = (lo4 + per < n4) ? lo4 + per : n4;
    float* pm = a.p[r];
    for (int64_t i4 = lo4 + (int64_t)blockIdx.x * blockDim.x + tid; i4 < hi4; i4 += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = i4 << 2;
        float4 gq[CLIPGP_PEER_MAX];
#pragma unroll
        for (int q = 0; q < CLIPGP_PEER_MAX; ++q)
            if (q < W) gq[q] = ld_sys_f4(a.g[q] + i);
        float4 gs = gq[0];
#pragma unroll
        for (int q = 1; q < CLIPGP_PEER_MAX; ++q)
            if (q < W) { gs.x += gq[q].x; gs.y += gq[q].y; gs.z += gq[q].z; gs.w += gq[q].w; }
        float4 pv = *reinterpret_cast<const float4*>(pm + i);
        float4 mv = *reinterpret_cast<const float4*>(a.m + i), vv = *reinterpret_cast<const float4*>(a.v + i);
        pv.x = adam1(gs.x, pv.x, mv.x, vv.x, s, i + 0 < a.n_group0);
        pv.y = adam1(gs.y, pv.y, mv.y, vv.y, s, i + 1 < a.n_group0);
        pv.z = adam1(gs.z, pv.z, mv.z, vv.z, s, i + 2 < a.n_group0);
        pv.w = adam1(gs.w, pv.w, mv.w, vv.w, s, i + 3 < a.n_group0);
        *reinterpret_cast<float4*>(a.m + i) = mv;
        *reinterpret_cast<float4*>(a.v + i) = vv;
#pragma unroll
        for (int q = 0; q < CLIPGP_PEER_MAX; ++q)
            if (q < W) st_sys_f4(a.p[q] + i, pv);
    }
    if (r == W - 1 && blockIdx.x == 0) {
        for (int64_t i = (n4 << 2) + tid; i < a.n; i += blockDim.x) {
            float gsum = 0.f;
            for (int q = 0; q < W; ++q) gsum += ld_sys_f1(a.g[q] + i);
            float mv = a.m[i], vv = a.v[i];
            const float pv = adam1(gsum, pm[i], mv, vv, s, i < a.n_group0);
            a.m[i] = mv; a.v[i] = vv;
            for (int q = 0; q < W; ++q) st_sys_f1(a.p[q] + i, pv);
        }
    }
    if (blockIdx.x == 0 && tid == 0 && a.loss_out != nullptr) {         // the loss rides in slot n of every gradient buffer
        float l = 0.f;
        for (int q = 0; q < W; ++q) l += ld_sys_f1(a.g[q] + a.n);
        *a.loss_out = l;
    }
    // ---- (C) everything of mine is written everywhere; nobody may touch the buffers of the next step before all are
    __threadfence_system();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(reinterpret_cast<unsigned int*>(a.local + 1), 1u) == gridDim.x - 1u);
    __syncthreads();
    if (s_last) {
        __threadfence_system();
        if (tid < W) st_release_sys_u64(a.flags[tid] + CLIPGP_PEER_MAX + r, e);
        if (tid < W && !spin_until(mine + CLIPGP_PEER_MAX + tid, e, a.timeout_ns)) atomicExch(a.status, 2);
        __syncthreads();
        if (tid == 0) {
            *reinterpret_cast<volatile unsigned int*>(a.local + 1) = 0u;
            *reinterpret_cast<volatile unsigned long long*>(a.local) = e;
            __threadfence();
        }
    }
}

}  // namespace clipgp

using namespace clipgp;

extern "C" int clipgp_peer_alloc(int64_t bytes, void** out) {
    CLIPGP_REQUIRE(bytes > 0 && out != nullptr, "peer_alloc: bad arguments");
    void* p = nullptr;
    CLIPGP_CUDA(cudaMalloc(&p, (size_t)bytes));
    CLIPGP_CUDA(cudaMemset(p, 0, (size_t)bytes));
    CLIPGP_CUDA(cudaDeviceSynchronize());
    *out = p;
    return CLIPGP_OK;
}

extern "C" int clipgp_peer_free(void* p) {
    if (p != nullptr) CLIPGP_CUDA(cudaFree(p));
    return CLIPGP_OK;
}

extern "C" int clipgp_ipc_export(const void* base, unsigned char* handle64) {
    CLIPGP_REQUIRE(base != nullptr && handle64 != nullptr, "ipc_export: NULL pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    CLIPGP_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(base)));
    memcpy(handle64, &h, 64);
    return CLIPGP_OK;
}

extern "C" int clipgp_ipc_open(const unsigned char* handle64, void** out) {
    CLIPGP_REQUIRE(handle64 != nullptr && out != nullptr, "ipc_open: NULL pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void* p = nullptr;
    CLIPGP_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *out = p;
    return CLIPGP_OK;
}

extern "C" int clipgp_ipc_close(void* p) {
    if (p != nullptr) CLIPGP_CUDA(cudaIpcCloseMemHandle(p));
    return CLIPGP_OK;
}

extern "C" int clipgp_peer_adamw(const clipgp_peer_args* a, void* stream) {
    CLIPGP_REQUIRE(a != nullptr, "peer_adamw: args is NULL");
    CLIPGP_REQUIRE(a->world >= 1 && a->world <= CLIPGP_PEER_MAX && a->rank >= 0 && a->rank < a->world, "peer_adamw: need 1 <= world <= %d and 0 <= rank < world",
                   CLIPGP_PEER_MAX);
    CLIPGP_REQUIRE(a->n >= 0 && a->n_group0 >= 0 && a->n_group0 <= a->n, "peer_adamw: bad sizes");
    CLIPGP_REQUIRE(a->m && a->v && a->lr_dev && a->step && a->local && a->status, "peer_adamw: NULL pointer");
    for (int q = 0; q < a->world; ++q) {
        CLIPGP_REQUIRE(a->g[q] && a->p[q] && a->flags[q], "peer_adamw: peer %d buffers missing", q);
        CLIPGP_REQUIRE(((reinterpret_cast<uintptr_t>(a->g[q]) | reinterpret_cast<uintptr_t>(a->p[q])) & 15u) == 0, "peer_adamw: peer %d buffers must be 16-byte aligned", q);
    }
    CLIPGP_REQUIRE(((reinterpret_cast<uintptr_t>(a->m) | reinterpret_cast<uintptr_t>(a->v)) & 15u) == 0, "peer_adamw: m / v must be 16-byte aligned");
    // one CTA per SM: enough 16-byte requests in flight to cover the NVLink round trip (world loads per thread), and the flag
    // waits cost nothing when the grid is co-resident
    const int64_t n4 = a->n >> 2, per = (n4 + a->world - 1) / a->world;
    int64_t blocks = (per + 255) / 256;
    if (blocks > num_sms()) blocks = num_sms();
    if (blocks < 1) blocks = 1;
    peer_adamw_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(*a);
    return check_launch("peer_adamw_kernel");
}
