// Tensor-core path of the logit / projection / cache-affinity contractions:  D[M,N] = alpha * A[M,K] B[N,K]^T with
// bf16 operands (K-major, i.e. plain row-major [rows, K]) and fp32 accumulation in TMEM.
//
//   * operands are staged by TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B) into a 4-stage shared-memory ring,
//   * one elected thread issues tcgen05.mma.cta_group::1.kind::f16 (UMMA 128 x 256 x 16), accumulators live in TMEM
//     (2 x 256 columns, double buffered so the epilogue of tile i overlaps the MMAs of tile i+1),
//   * four epilogue warps read the accumulator with tcgen05.ld (thread <-> TMEM lane <-> output row) and either
//     store alpha*D (EPI_STORE) or reduce it on the fly (EPI_ROWSTATS): running softmax max / arg-max / sum-exp per
//     row across all class tiles -> confidence, hit flag, top-1 count and the ECE histogram (utils/metrics.py:71-82),
//     so the [N_img, C] logits of the eval path never touch HBM.
//   * `Ka` < K makes A wrap along K (A column = k mod Ka): with B = [p_hat_1 | ... | p_hat_S] along K this
//     accumulates sum_s f_hat . p_hat_s in TMEM, i.e. the MC-averaged logits of adapter.py:247-249 as ONE GEMM of
//     2*N*S*C*D flops without materialising [N,S,C].
//
// Persistent kernel: grid = min(#SMs, #work items); warp 0 = TMA producer, warp 1 = MMA issuer + TMEM allocator,
// warps 2..5 = epilogue.  Roofline: tensor pipe (bf16) for large M; HBM/latency for M = 128.
#include <cuda.h>
#include <float.h>
#include <stdlib.h>
#include <mutex>

#include "common.cuh"

namespace clipgp {
namespace tc {

constexpr int BM = 128, BN = 256, BK = 64, UK = 16;
constexpr int STAGES = 4;
constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int THREADS = 192;        // TMA warp + MMA warp + 4 epilogue warps (long-K configuration)
constexpr int THREADS_WIDE = 320;   // ... + 8 epilogue warps (short-K store configuration: the epilogue is the bottleneck there)
constexpr int ACC_COLS = BN;       // fp32 accumulator columns per tile
constexpr int TMEM_COLS = 512;     // two accumulator stages
constexpr int CST_BYTES = 32 * 32 * 4;             // one 32 x 32 fp32 staging tile of the TMA-store epilogue
// staging: long-K config = 4 stages + 4 warps x 1 tile (212 KB); short-K config = 3 stages + 8 warps x 2 tiles (212 KB); the
// 227 KB limit (static barriers / histogram included) rules out 4 stages with more staging

constexpr int EPI_STORE = 0, EPI_ROWSTATS = 1, EPI_TIP = 2;

struct Params {
    int M, N, K, Ka;
    int num_m, num_n, num_k;
    int n_per_item;             // consecutive n tiles one work item covers (num_n for ROWSTATS, 1 for STORE)
    int k_splits, kb_per_split; // EPI_STORE only: the K blocks are split over k_splits work items that accumulate atomically
    int full_items, tail_split, tail_kb;   // work items >= full_items are the tiles of the last, partial wave, each split over
                                           // tail_split work items (tail_kb K blocks each) so that the wave fills all SMs
    int mode;
    int tma_store;              // C is written by TMA from swizzled shared-memory staging tiles (full-line stores)
    int stages, cbufs;          // operand ring depth (3 or 4) and staging tiles per epilogue warp (1 or 2)
    float alpha;
    float* C; long long ldc;    // STORE target / optional logits copy in ROWSTATS
    const long long* labels; float* conf; int* pred; unsigned char* correct;
    const float* boundaries; int n_bins;
    long long* bin_count; unsigned long long* bin_conf_fx; long long* bin_correct; long long* top1;
    const int* key_class; float beta; float tip_alpha;   // EPI_TIP: class of every key (column), exp(-beta(1-aff)), alpha
    int elt;                    // operand element size: 2 = bf16 (kind::f16), 4 = fp32 read as TF32 (kind::tf32; no operand cast at all)
    int a_mn, b_mn;             // operand given MN-major: A as [K, M] / B as [K, N] row-major (e.g. dlogits [B, S*C] as the A = dlogits^T
                                // operand of d P_hat = dlogits^T f_hat): the transposed copy is never made, the tensor core reads it in place
    int norm_cols;              // EPI_ROWSTATS: the first norm_cols columns (a multiple of BN) are not classes but the projected feature
                                // y = f W^T; their squared sum gives 1 / max(|y|, 1e-12), which scales the class columns that follow
                                // (F.normalize of the projection, adapter.py:239-240, without materialising it)
};

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a pipeline whose producer never arrives (a bad tensor map / descriptor, a byte count that does not match the boxes) would
// otherwise spin for ever and take the GPU with it.  One counter increment per failed poll (a version that read %globaltimer and
// backed off with nanosleep cost the issuing thread 4-5 % of the GEMM's rate); 2^28 failed polls are seconds, then the launch traps
// and fails with an error instead of hanging.  The TMA producer and the epilogue warps wait this way; the single MMA-issuing thread keeps
// the bare loop below (every instruction between a successful poll and the next tcgen05.mma is on the kernel's critical path: the
// counter alone cost 5 % there) -- a stalled pipeline still ends, because one of the bounded waiters traps the whole grid.
__device__ __forceinline__ void mbar_wait_issuer(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t polls = 0;
    while (!mbar_try_wait(bar, parity)) {
#ifndef CLIPGP_MBAR_UNBOUNDED
        if (++polls == (1u << 28)) __trap();
#endif
    }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address >> 4 in bits
// [0,14), LBO = 1 (ignored for swizzled K-major) in [16,30), SBO = 1024 B >> 4 (8 rows x 128 B) in [32,46),
// version = 1 in [46,48), layout type SWIZZLE_128B = 2 in [61,64).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
    const uint32_t lo = ((smem_addr >> 4) & 0x3FFFu) | (1u << 16);
    const uint32_t hi = 64u | (1u << 14) | (2u << 29);
    return ((uint64_t)hi << 32) | lo;
}
// MN-major operands (see cute/atom/mma_traits_sm100.hpp, make_umma_desc<Major::MN>): a 128-byte line runs along M/N, a group of
// K rows forms one swizzle atom, the stage's K groups are stacked SBO apart and consecutive 128-byte M/N blocks LBO apart (one TMA
// box of [k rows of the stage][128 B] per M/N block, box after box).
//   16-bit elements: SWIZZLE_128B (16-byte chunks, 8 K rows per atom: SBO = 1024 B), layout type 2;
//   32-bit elements (TF32): the only MN-major layout the tensor core accepts is SWIZZLE_128B_BASE32B (32-byte chunks XOR-ed with
//   the row index mod 4, i.e. 4 K rows per atom: SBO = 512 B), layout type 1, loaded with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.
__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes, int elt) {
    const uint32_t lo = ((smem_addr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
    const uint32_t hi = (elt == 4) ? (32u | (1u << 14) | (1u << 29)) : (64u | (1u << 14) | (2u << 29));
    return ((uint64_t)hi << 32) | lo;
}
// cute::UMMA::InstrDescriptor: D = F32 (1 << 4), A / B format in [7,10) / [10,13) (1 = BF16, 2 = TF32), A / B major in bit 15 / 16
// (0 = K-major, 1 = MN-major), N >> 3 in [17,23), M >> 4 in [24,29).
__device__ __forceinline__ uint32_t make_idesc(int elt, int a_mn, int b_mn) {
    const uint32_t fmt = (elt == 4) ? 2u : 1u;
    return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(a_mn != 0) << 15) | ((uint32_t)(b_mn != 0) << 16) |
           ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// Work item -> (m block, first n block, K-block range, accumulate-atomically flag).
struct Item { int m_blk, n_first, kb0, kb1, atomic; };
__device__ __forceinline__ Item decode_item(const Params& p, int item) {
    Item it;
    const int groups = p.num_n / p.n_per_item;
    if (item < p.full_items) {
        const int ks = item % p.k_splits, mg = item / p.k_splits;
        it.m_blk = mg / groups; it.n_first = (mg - it.m_blk * groups) * p.n_per_item;
        it.kb0 = ks * p.kb_per_split; it.kb1 = min(p.num_k, it.kb0 + p.kb_per_split);
        it.atomic = p.k_splits > 1;
    } else {
        const int r = item - p.full_items;
        const int tile = p.full_items + r / p.tail_split, ks = r - (r / p.tail_split) * p.tail_split;
        it.m_blk = tile / groups; it.n_first = (tile - it.m_blk * groups) * p.n_per_item;
        it.kb0 = ks * p.tail_kb; it.kb1 = min(p.num_k, it.kb0 + p.tail_kb);
        it.atomic = 1;
    }
    return it;
}

__device__ __forceinline__ unsigned long long conf_to_fx(float c) { return (unsigned long long)((double)c * 1099511627776.0); }

// EPI_TIP: a finished run of equal-class keys leaves the registers: into the row's slot of the warp's class tile when the class
// lies in the tile's 32-class window, else straight to the output row.  Out of line on purpose: it is called from 33 places of
// the unrolled epilogue and runs two or three times per 32 columns.
__device__ __noinline__ void tip_flush_run(float* tip_tile, int lane, int cls, int tip_first, float sum, float* orow, float tip_alpha) {
    if (cls < 0 || orow == nullptr) return;
    const int slot = cls - tip_first;
    if (slot >= 0 && slot < 32) tip_tile[lane * 32 + ((slot + lane) & 31)] += sum;
    else atomicAdd(orow + cls, tip_alpha * sum);
}

// ------------------------------------------------------------------------------------------------ kernel
__global__ void __launch_bounds__(THREADS_WIDE, 1) tc_gemm_kernel(const __grid_constant__ CUtensorMap map_a,
                                                             const __grid_constant__ CUtensorMap map_b,
                                                             const __grid_constant__ CUtensorMap map_c, const Params p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], tmem_full[2], tmem_empty[2];
    __shared__ uint32_t tmem_base_smem;
    __shared__ float s_b[CLIPGP_MAX_BINS + 1];
    __shared__ unsigned int s_cnt[CLIPGP_MAX_BINS], s_cor[CLIPGP_MAX_BINS], s_top1;
    __shared__ unsigned long long s_fx[CLIPGP_MAX_BINS];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], (blockDim.x >> 5) - 2); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_b)) : "memory");
        if (p.tma_store) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_c)) : "memory");
    }
    if (p.mode == EPI_ROWSTATS) {
        for (int i = threadIdx.x; i < CLIPGP_MAX_BINS; i += blockDim.x) { s_cnt[i] = 0; s_cor[i] = 0; s_fx[i] = 0ull; }
        for (int i = threadIdx.x; i <= p.n_bins; i += blockDim.x) s_b[i] = p.boundaries ? p.boundaries[i] : 0.f;
        if (threadIdx.x == 0) s_top1 = 0;
    }
    if (warp == 1) tmem_alloc(&tmem_base_smem, TMEM_COLS);
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    const int num_items = p.full_items + (p.num_m * (p.num_n / p.n_per_item) * p.k_splits - p.full_items) * p.tail_split;

    if (warp == 0) {
        // ===================================================== TMA producer
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
                const Item it = decode_item(p, item);
                const int m_blk = it.m_blk, kb0 = it.kb0, kb1 = it.kb1;
                const int bke = 128 / p.elt;                              // K elements per 128-byte line: 64 bf16 / 32 tf32
                const int mnb = 128 / p.elt;                              // M/N elements per 128-byte line of an MN-major operand
                for (int nn = 0; nn < p.n_per_item; ++nn) {
                    const int n_blk = it.n_first + nn;
                    for (int kb = kb0; kb < kb1; ++kb) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        unsigned char* sa = smem + stage * STAGE_BYTES;
                        mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
                        if (!p.a_mn) tma_load_2d(sa, &map_a, &full_bar[stage], (kb * bke) % p.Ka, m_blk * BM);
                        else
                            for (int blk = 0; blk < BM / mnb; ++blk)      // one [bke k rows][128 B] box per 128-byte block of rows
                                tma_load_2d(sa + blk * (bke * 128), &map_a, &full_bar[stage], m_blk * BM + blk * mnb, kb * bke);
                        if (!p.b_mn) tma_load_2d(sa + A_BYTES, &map_b, &full_bar[stage], kb * bke, n_blk * BN);
                        else
                            for (int blk = 0; blk < BN / mnb; ++blk)
                                tma_load_2d(sa + A_BYTES + blk * (bke * 128), &map_b, &full_bar[stage], n_blk * BN + blk * mnb, kb * bke);
                        if (++stage == p.stages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================================================== MMA issuer
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            const uint32_t idesc = make_idesc(p.elt, p.a_mn, p.b_mn);
            const int bke = 128 / p.elt;
            // one instruction consumes 32 bytes of K: K-major -> the next 32 bytes of every 128-byte row (descriptor + 2);
            // MN-major -> the next 32 / elt K rows = (32 / elt) / 8 swizzle atoms of 1024 bytes (descriptor + 64 per atom)
            const uint64_t a_step = p.a_mn ? (uint64_t)((32 / p.elt) / 8 * 64) : 2ull;
            const uint64_t b_step = p.b_mn ? (uint64_t)((32 / p.elt) / 8 * 64) : 2ull;
            const uint32_t lbo = (uint32_t)(bke * 128);                   // MN-major: distance between 128-byte M/N blocks
            for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
                const Item it = decode_item(p, item);
                const int kb0 = it.kb0, kb1 = it.kb1;
                for (int nn = 0; nn < p.n_per_item; ++nn) {
                    mbar_wait_issuer(&tmem_empty[acc], acc_phase ^ 1);
                    fence_after();
                    const uint32_t tmem_d = tmem_base + (uint32_t)(acc * ACC_COLS);
                    for (int kb = kb0; kb < kb1; ++kb) {
                        mbar_wait_issuer(&full_bar[stage], phase);
                        fence_after();
                        const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
                        const uint64_t adesc = p.a_mn ? make_smem_desc_mn(sa, lbo, p.elt) : make_smem_desc(sa);
                        const uint64_t bdesc = p.b_mn ? make_smem_desc_mn(sa + A_BYTES, lbo, p.elt) : make_smem_desc(sa + A_BYTES);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {                     // 128-byte K line / 32 bytes per instruction
                            const uint32_t accum = (kb > kb0 || k != 0) ? 1u : 0u;
                            if (p.elt == 4) umma_tf32(tmem_d, adesc + (uint64_t)k * a_step, bdesc + (uint64_t)k * b_step, idesc, accum);
                            else umma_bf16(tmem_d, adesc + (uint64_t)k * a_step, bdesc + (uint64_t)k * b_step, idesc, accum);
                        }
                        umma_commit(&empty_bar[stage]);                 // frees the smem slot when these MMAs retire
                        if (kb == kb1 - 1) umma_commit(&tmem_full[acc]);
                        if (++stage == p.stages) { stage = 0; phase ^= 1; }
                    }
                    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else {
        // ===================================================== epilogue (4 warps; thread <-> TMEM lane <-> row)
        const int q = warp & 3;
        // eight epilogue warps: warps q and q+4 share the TMEM lane quadrant and split the 256 accumulator columns in halves
        const int n_epi = (blockDim.x >> 5) - 2;
        const int c_lo = (n_epi == 8) ? ((warp - 2) >> 2) * (BN / 2) : 0;
        const int c_hi = (n_epi == 8) ? c_lo + BN / 2 : BN;
        int cbuf = 0;
        int acc = 0; uint32_t acc_phase = 0;
        for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
            const Item it = decode_item(p, item);
            const int m_blk = it.m_blk, n_first = it.n_first;
            const int row = m_blk * BM + q * 32 + lane;
            float run_m = -FLT_MAX, run_s = 0.f; int run_am = 0x7fffffff;
            float sumsq = 0.f, sc = p.alpha;                     // EPI_ROWSTATS with norm columns: |f W^T|^2 of this row, then alpha / |f W^T|
            for (int nn = 0; nn < p.n_per_item; ++nn) {
                const int n_blk = n_first + nn;
                mbar_wait(&tmem_full[acc], acc_phase);
                fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * ACC_COLS);
                // EPI_TIP: per-warp class tile [32 rows][32 class slots] (slot index rotated by the row: conflict-free column walks)
                float* tip_tile = reinterpret_cast<float*>(smem + p.stages * STAGE_BYTES + (warp - 2) * CST_BYTES);
                int tip_cur = -1, tip_first = 0; float tip_sum = 0.f;
                if (p.mode == EPI_TIP) {
                    tip_first = __ldg(p.key_class + min(n_blk * BN, p.N - 1));
#pragma unroll
                    for (int i = 0; i < 8; ++i) reinterpret_cast<float4*>(tip_tile)[lane + 32 * i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    __syncwarp();
                }
#pragma unroll 1
                for (int c0 = c_lo; c0 < c_hi; c0 += 32) {
                    const int col0 = n_blk * BN + c0;
                    if (col0 >= p.N) break;                              // warp-uniform
                    uint32_t r[32];
                    tmem_ld32(taddr + (uint32_t)c0, r);
                    const int ncols = min(32, p.N - col0);
                    if (p.mode == EPI_TIP) {
                        // cache_logits[b, class(j)] += exp(-beta (1 - aff[b,j])): run-length sums over equal classes, carried across
                        // the chunks of the tile; a finished run goes to this row's slot of the warp's 32 x 32 class tile in shared
                        // memory (classes tip_first .. tip_first + 31; keys are sorted by class on the host, so a 256-key tile
                        // spans few classes) and the tile leaves with 32 row-contiguous reductions after the accumulator is
                        // released.  Classes outside the window (unsorted keys, 1-shot caches) fall back to a direct atomic.
                        // The chunk's 32 key classes come from one coalesced load (lane = column); run boundaries become a ballot
                        // mask, so the element loop carries no memory-dependent branch.
                        const float bl = p.beta * 1.4426950408889634f, bla = bl * p.alpha;
                        const int my_cls = (lane < ncols) ? __ldg(p.key_class + col0 + lane) : -2;
                        const int prev_cls = __shfl_up_sync(0xffffffffu, my_cls, 1);
                        const unsigned bmask = __ballot_sync(0xffffffffu, lane < ncols && my_cls != (lane == 0 ? tip_cur : prev_cls));
                        float* orow = row < p.M ? p.C + (long long)row * p.ldc : nullptr;
#pragma unroll
                        for (int j4 = 0; j4 < 32; j4 += 4) {             // four columns at a time: a group without a boundary is 4 FFMA + 4 EX2 + 4 FADD
                            float e4[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e4[u]) : "f"(fmaf(bla, __uint_as_float(r[j4 + u]), -bl)));
                                if (ncols < 32 && j4 + u >= ncols) e4[u] = 0.f;
                            }
                            const unsigned bits = (bmask >> j4) & 0xFu;  // warp-uniform
                            if (bits == 0u) tip_sum += (e4[0] + e4[1]) + (e4[2] + e4[3]);
                            else {
#pragma unroll
                                for (int u = 0; u < 4; ++u) {
                                    if ((bits >> u) & 1u) {               // a new class starts at column j4 + u
                                        tip_flush_run(tip_tile, lane, tip_cur, tip_first, tip_sum, orow, p.tip_alpha);
                                        tip_cur = __shfl_sync(0xffffffffu, my_cls, j4 + u); tip_sum = 0.f;
                                    }
                                    tip_sum += e4[u];
                                }
                            }
                        }
                    } else if (p.C != nullptr && p.tma_store && !it.atomic) {
                        // stage the 32 x 32 chunk in shared memory in the 128-byte-swizzled layout of the output tensor map
                        // (lane = row; 16-byte chunk index XOR (row & 7): the four 8-lane phases of each vector store hit
                        // disjoint banks), then one TMA store writes full 128-byte lines; rows / columns past M / N are clipped
                        unsigned char* stg = smem + p.stages * STAGE_BYTES + ((warp - 2) * p.cbufs + cbuf) * CST_BYTES;
                        if (lane == 0) {                                  // the previous store from this tile has been read
                            if (p.cbufs == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                            else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                        }
                        __syncwarp();
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 v = make_float4(p.alpha * __uint_as_float(r[j]), p.alpha * __uint_as_float(r[j + 1]),
                                                         p.alpha * __uint_as_float(r[j + 2]), p.alpha * __uint_as_float(r[j + 3]));
                            *reinterpret_cast<float4*>(stg + lane * 128 + ((((j >> 2) ^ (lane & 7))) << 4)) = v;
                        }
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        __syncwarp();
                        if (lane == 0) tma_store_2d(&map_c, stg, col0, m_blk * BM + q * 32);
                        if (p.cbufs == 2) cbuf ^= 1;
                    } else if (p.C != nullptr && row < p.M) {
                        float* dst = p.C + (long long)row * p.ldc + col0;
                        if (it.atomic) {
                            if (ncols == 32 && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0)) {
#pragma unroll
                                for (int j = 0; j < 32; j += 4)        // vector reduction: one L2 atomic per 16 bytes
                                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j),
                                                 "f"(p.alpha * __uint_as_float(r[j])), "f"(p.alpha * __uint_as_float(r[j + 1])),
                                                 "f"(p.alpha * __uint_as_float(r[j + 2])), "f"(p.alpha * __uint_as_float(r[j + 3])) : "memory");
                            } else {
#pragma unroll
                                for (int j = 0; j < 32; ++j)
                                    if (j < ncols) atomicAdd(dst + j, p.alpha * __uint_as_float(r[j]));
                            }
                        } else if (ncols == 32 && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0)) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4)
                                *reinterpret_cast<float4*>(dst + j) = make_float4(p.alpha * __uint_as_float(r[j]), p.alpha * __uint_as_float(r[j + 1]),
                                                                                   p.alpha * __uint_as_float(r[j + 2]), p.alpha * __uint_as_float(r[j + 3]));
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (j < ncols) dst[j] = p.alpha * __uint_as_float(r[j]);
                        }
                    }
                    if (p.mode == EPI_ROWSTATS && col0 < p.norm_cols) {
                        // projection columns: accumulate the squared norm; the last chunk fixes the scale of the class columns
#pragma unroll
                        for (int j = 0; j < 32; ++j) sumsq = fmaf(__uint_as_float(r[j]), __uint_as_float(r[j]), sumsq);
                        if (col0 + 32 >= p.norm_cols) sc = p.alpha / fmaxf(sqrtf(sumsq), 1e-12f);
                    } else if (p.mode == EPI_ROWSTATS) {
                        // chunk max / arg-max (lowest index on ties), then one rescale of the running sum
                        float cm = -FLT_MAX; int cam = 0;
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const float v = sc * __uint_as_float(r[j]);
                            if (j < ncols && v > cm) { cm = v; cam = j; }
                        }
                        if (cm > run_m) {
                            run_s *= exp2f((run_m - cm) * 1.4426950408889634f);
                            run_m = cm; run_am = col0 - p.norm_cols + cam;
                        }
                        float s = 0.f;
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const float v = sc * __uint_as_float(r[j]);
                            if (j < ncols) s += exp2f((v - run_m) * 1.4426950408889634f);
                        }
                        run_s += s;
                    }
                }
                fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                if (p.mode == EPI_TIP) {
                    tip_flush_run(tip_tile, lane, tip_cur, tip_first, tip_sum, row < p.M ? p.C + (long long)row * p.ldc : nullptr, p.tip_alpha);
                    __syncwarp();
                    const int row0 = m_blk * BM + q * 32;
#pragma unroll 4
                    for (int rr = 0; rr < 32; ++rr) {                     // lane = class slot: one contiguous reduction per row
                        const float v = tip_tile[rr * 32 + ((lane + rr) & 31)];
                        if (v != 0.f && row0 + rr < p.M)
                            asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p.C + (long long)(row0 + rr) * p.ldc + tip_first + lane),
                                         "f"(p.tip_alpha * v) : "memory");
                    }
                    __syncwarp();
                }
            }
            if (p.mode == EPI_ROWSTATS && row < p.M) {
                const float cf = 1.0f / run_s;
                const int ok = (p.labels != nullptr) ? ((long long)run_am == p.labels[row]) : 0;
                if (p.conf) p.conf[row] = cf;
                if (p.pred) p.pred[row] = run_am;
                if (p.correct) p.correct[row] = (unsigned char)ok;
                if (ok) atomicAdd(&s_top1, 1u);
                if (p.bin_count != nullptr) {
                    int bi = -1;
                    for (int i = 0; i < p.n_bins; ++i)
                        if (cf > s_b[i] && cf <= s_b[i + 1]) { bi = i; break; }
                    if (bi >= 0) { atomicAdd(&s_cnt[bi], 1u); atomicAdd(&s_cor[bi], (unsigned int)ok); atomicAdd(&s_fx[bi], conf_to_fx(cf)); }
                }
            }
        }
    }
    if (p.tma_store && warp >= 2 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // staging tiles drained
    fence_before();
    __syncthreads();
    if (warp == 1) { fence_after(); tmem_dealloc(tmem_base, TMEM_COLS); }
    if (p.mode == EPI_ROWSTATS) {
        if (p.bin_count != nullptr)
            for (int i = threadIdx.x; i < p.n_bins; i += blockDim.x)
                if (s_cnt[i]) {
                    atomicAdd((unsigned long long*)&p.bin_count[i], (unsigned long long)s_cnt[i]);
                    atomicAdd(&p.bin_conf_fx[i], s_fx[i]);
                    atomicAdd((unsigned long long*)&p.bin_correct[i], (unsigned long long)s_cor[i]);
                }
        if (p.top1 != nullptr && threadIdx.x == 0 && s_top1) atomicAdd((unsigned long long*)p.top1, (unsigned long long)s_top1);
    }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

// Encoded descriptors are pure functions of (base pointer, geometry, layout kind): a small cache keeps eager callers (the autograd
// heads, the trainers' per-batch GEMMs on persistent buffers) from re-encoding three maps per launch on the host.
struct MapKey {
    const void* base; long long a, b, c; int kind;
    bool operator==(const MapKey& o) const { return base == o.base && a == o.a && b == o.b && c == o.c && kind == o.kind; }
};
struct MapSlot { MapKey key; CUtensorMap map; bool used; };
static MapSlot g_map_cache[128];
static std::mutex g_map_mutex;
static unsigned map_slot(const MapKey& k) {
    unsigned long long h = reinterpret_cast<uintptr_t>(k.base) * 0x9E3779B97F4A7C15ull;
    h ^= (unsigned long long)k.a * 0xC2B2AE3D27D4EB4Full; h ^= (unsigned long long)k.b * 0x165667B19E3779F9ull;
    h ^= (unsigned long long)k.c * 0x27D4EB2F165667C5ull; h ^= (unsigned long long)k.kind * 0x85EBCA77C2B2AE63ull;
    return (unsigned)(h >> 40) & 127u;
}
static bool map_lookup(const MapKey& k, CUtensorMap* out) {
    std::lock_guard<std::mutex> lock(g_map_mutex);
    const MapSlot& sl = g_map_cache[map_slot(k)];
    if (sl.used && sl.key == k) { *out = sl.map; return true; }
    return false;
}
static void map_store(const MapKey& k, const CUtensorMap& m) {
    std::lock_guard<std::mutex> lock(g_map_mutex);
    MapSlot& sl = g_map_cache[map_slot(k)];
    sl.key = k; sl.map = m; sl.used = true;
}

// Operand tensor map (128-byte swizzle, zero OOB fill), element size `elt` (2 = bf16, 4 = fp32 consumed as TF32).
//   K-major  operand [rows, K] row-major: box = [128 B of K][box_rows rows];
//   MN-major operand [K, rows] row-major: box = [128 B of rows][128 / elt ... K rows of one stage] (see make_smem_desc_mn).
static int make_map(CUtensorMap* map, const void* base, long long rows, long long K, int box_rows, int elt, int mn_major) {
    const MapKey key{base, rows, K, (long long)box_rows, elt * 2 + (mn_major ? 1 : 0)};
    if (map_lookup(key, map)) return CLIPGP_OK;
    EncodeTiledFn enc = get_encode();
    if (enc == nullptr) { set_error("tc_gemm: cuTensorMapEncodeTiled is not available from the driver"); return CLIPGP_ERR_CUDA; }
    const cuuint32_t line = (cuuint32_t)(128 / elt);
    cuuint64_t dims[2], strides[1];
    cuuint32_t box[2];
    if (!mn_major) { dims[0] = (cuuint64_t)K; dims[1] = (cuuint64_t)rows; strides[0] = (cuuint64_t)K * elt; box[0] = line; box[1] = (cuuint32_t)box_rows; }
    else { dims[0] = (cuuint64_t)rows; dims[1] = (cuuint64_t)K; strides[0] = (cuuint64_t)rows * elt; box[0] = line; box[1] = line; }
    cuuint32_t estr[2] = {1, 1};
    const CUtensorMapSwizzle swz = (mn_major && elt == 4) ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B;
    CUresult r = enc(map, elt == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims,
                     strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("tc_gemm: cuTensorMapEncodeTiled failed (CUresult %d)", (int)r); return CLIPGP_ERR_CUDA; }
    map_store(key, *map);
    return CLIPGP_OK;
}

// fp32 row-major [rows, cols] (row pitch ld elements) -> 2D tensor map with a 32 x 32 box, 128B swizzle (one box row = 128 bytes)
static int make_map_c(CUtensorMap* map, const float* base, long long rows, long long cols, long long ld) {
    const MapKey key{base, rows, cols, ld, 100};
    if (map_lookup(key, map)) return CLIPGP_OK;
    EncodeTiledFn enc = get_encode();
    if (enc == nullptr) { set_error("tc_gemm: cuTensorMapEncodeTiled is not available from the driver"); return CLIPGP_ERR_CUDA; }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {32, 32};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("tc_gemm: cuTensorMapEncodeTiled(C) failed (CUresult %d)", (int)r); return CLIPGP_ERR_CUDA; }
    map_store(key, *map);
    return CLIPGP_OK;
}

static int launch(const void* A, long long M, long long Ka, const void* B, long long N, long long K, Params& p, cudaStream_t st,
                  bool allow_split_k = false) {
    CLIPGP_REQUIRE(M >= 1 && N >= 1 && K >= 1 && Ka >= 1, "tc_gemm: empty problem");
    CLIPGP_REQUIRE(M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), "tc_gemm: size too large");
    CLIPGP_REQUIRE(A && B, "tc_gemm: NULL operand");
    if (p.elt == 0) p.elt = 2;
    const int bke = 128 / p.elt, pitch = 16 / p.elt;           // K elements per 128-byte line; elements per 16 bytes (TMA row pitch)
    CLIPGP_REQUIRE(!p.a_mn || Ka == K, "tc_gemm: an MN-major A operand cannot wrap along K");
    CLIPGP_REQUIRE((p.a_mn ? M : Ka) % pitch == 0 && (p.b_mn ? N : K) % pitch == 0,
                   "tc_gemm: operand row pitch must be a multiple of 16 bytes (M=%lld N=%lld K=%lld Ka=%lld, element size %d)", M, N, K, Ka, p.elt);
    CLIPGP_REQUIRE(K % Ka == 0 && (Ka == K || Ka % bke == 0), "tc_gemm: K must be a multiple of Ka, and Ka a multiple of %d when A wraps", bke);
    CLIPGP_REQUIRE(((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) & 15u) == 0, "tc_gemm: operands must be 16-byte aligned");
    CUtensorMap ma, mb;
    int rc = make_map(&ma, A, M, Ka, BM, p.elt, p.a_mn);
    if (rc != CLIPGP_OK) return rc;
    rc = make_map(&mb, B, N, K, BN, p.elt, p.b_mn);
    if (rc != CLIPGP_OK) return rc;
    p.M = (int)M; p.N = (int)N; p.K = (int)K; p.Ka = (int)Ka;
    p.num_m = (int)((M + BM - 1) / BM); p.num_n = (int)((N + BN - 1) / BN); p.num_k = (int)((K + bke - 1) / bke);
    p.n_per_item = (p.mode == EPI_ROWSTATS) ? p.num_n : 1;
    p.k_splits = 1; p.kb_per_split = p.num_k;
    if (p.mode == EPI_STORE && allow_split_k) {
        // split K when the output grid cannot fill the SMs and K is long (skinny adjoint GEMMs, e.g. d f_hat = dlogits P_hat)
        const int tiles = p.num_m * p.num_n;
        if (tiles * 2 <= num_sms() && p.num_k >= 4) {
            int want = num_sms() / tiles;
            if (want > p.num_k / 2) want = p.num_k / 2;          // at least two K blocks (2 x 48 KB of operands) per work item
            if (want > 1) {
                p.kb_per_split = (p.num_k + want - 1) / want;
                p.k_splits = (p.num_k + p.kb_per_split - 1) / p.kb_per_split;
                CLIPGP_CUDA(cudaMemset2DAsync(p.C, sizeof(float) * p.ldc, 0, sizeof(float) * N, M, st));
            }
        }
    }
    p.full_items = p.num_m * (p.num_n / p.n_per_item) * p.k_splits; p.tail_split = 1; p.tail_kb = p.num_k;
    if (p.mode == EPI_STORE && allow_split_k && p.k_splits == 1) {
        // tail wave: tiles % SMs work items would occupy a whole extra round of the persistent grid (d P_hat of the full-batch
        // step: 158 tiles on 148 SMs = 2 rounds).  Split the K range of just those tiles over the idle SMs (stream-K for the tail).
        const int tiles = p.num_m * p.num_n, sms = num_sms();
        const int tail = tiles % sms;
        if (tiles > sms && tail > 0 && tail * 2 <= sms && p.num_k >= 8) {
            int want = sms / tail;
            if (want > p.num_k / 2) want = p.num_k / 2;
            if (want > 1) {
                p.tail_kb = (p.num_k + want - 1) / want;
                p.tail_split = (p.num_k + p.tail_kb - 1) / p.tail_kb;
                p.full_items = tiles - tail;
                // the tail tiles accumulate atomically: zero the rows they cover (from the first tail tile's row block to M; the few
                // whole-K tiles that share that row block are stored afterwards, so zeroing them too is harmless)
                const long long row0 = (long long)(p.full_items / p.num_n) * BM;
                CLIPGP_CUDA(cudaMemset2DAsync(p.C + row0 * p.ldc, sizeof(float) * p.ldc, 0, sizeof(float) * N, M - row0, st));
            }
        }
    }
    const int items = p.full_items + (p.num_m * (p.num_n / p.n_per_item) * p.k_splits - p.full_items) * p.tail_split;
    // TMA-store epilogue for materialised outputs (store mode, or the optional logits copy of the row-statistics mode)
    CUtensorMap mc = ma;
    p.tma_store = 0;
    static const bool no_tma_store = (getenv("CLIPGP_TC_NO_TMA_STORE") != nullptr);
    if (!no_tma_store && p.C != nullptr && p.mode != EPI_TIP && p.k_splits == 1 && (p.ldc % 4) == 0 &&
        ((reinterpret_cast<uintptr_t>(p.C) & 15u) == 0)) {
        rc = make_map_c(&mc, p.C, M, N, p.ldc);
        if (rc != CLIPGP_OK) return rc;
        p.tma_store = 1;
    }
    // short-K materialised outputs are epilogue bound: 3-stage ring, 8 epilogue warps, double-buffered staging tiles
    const bool wide = p.tma_store && p.mode == EPI_STORE && p.kb_per_split <= 8;      // K <= 512 (measured: K = 1024 prefers 4 stages)
    p.stages = wide ? 3 : STAGES;
    p.cbufs = wide ? 2 : 1;
    const int threads = wide ? THREADS_WIDE : THREADS;
    const size_t smem = (size_t)STAGES * STAGE_BYTES + 4 * CST_BYTES + 1024;      // both configurations need the same 212 KB
    static bool attr_set = false;
    if (!attr_set) {
        CLIPGP_CUDA(cudaFuncSetAttribute(tc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    const int grid = items < num_sms() ? items : num_sms();
    tc_gemm_kernel<<<grid, threads, smem, st>>>(ma, mb, mc, p);
    return check_launch("tc_gemm_kernel");
}

}  // namespace tc
}  // namespace clipgp

using namespace clipgp;

extern "C" int clipgp_tc_gemm_store(const void* A_bf16, int64_t M, int64_t Ka, const void* B_bf16, int64_t N, int64_t K,
                                    float alpha, float* C, int64_t ldc, void* stream) {
    CLIPGP_REQUIRE(C != nullptr && ldc >= N, "tc_gemm_store: bad output");
    tc::Params p = {};
    p.mode = tc::EPI_STORE; p.alpha = alpha; p.C = C; p.ldc = ldc;
    return tc::launch(A_bf16, M, Ka, B_bf16, N, K, p, (cudaStream_t)stream);
}

extern "C" int clipgp_tc_gemm_store_splitk(const void* A_bf16, int64_t M, int64_t Ka, const void* B_bf16, int64_t N, int64_t K,
                                           float alpha, float* C, int64_t ldc, void* stream) {
    CLIPGP_REQUIRE(C != nullptr && ldc >= N, "tc_gemm_store_splitk: bad output");
    tc::Params p = {};
    p.mode = tc::EPI_STORE; p.alpha = alpha; p.C = C; p.ldc = ldc;
    return tc::launch(A_bf16, M, Ka, B_bf16, N, K, p, (cudaStream_t)stream, true);
}

// TF32 on fp32 operands read in place (kind::tf32: 10-bit mantissa products, fp32 accumulation -- the reference's own GPU
// arithmetic, adapter.py:23 `allow_tf32`).  a_layout / b_layout: 0 = K-major ([rows, K] row-major), 1 = MN-major ([K, rows]
// row-major, i.e. the operand is the TRANSPOSE of a row-major tensor and is consumed without a transposed copy).
extern "C" int clipgp_tc_gemm_tf32(const float* A, int a_layout, int64_t M, const float* B, int b_layout, int64_t N, int64_t K,
                                   float alpha, float* C, int64_t ldc, int allow_split_k, void* stream) {
    CLIPGP_REQUIRE(C != nullptr && ldc >= N, "tc_gemm_tf32: bad output");
    tc::Params p = {};
    p.mode = tc::EPI_STORE; p.alpha = alpha; p.C = C; p.ldc = ldc;
    p.elt = 4; p.a_mn = a_layout != 0; p.b_mn = b_layout != 0;
    return tc::launch(A, M, K, B, N, K, p, (cudaStream_t)stream, allow_split_k != 0);
}

extern "C" int clipgp_tc_logits_calibration_tf32(const float* A, int64_t M, int64_t Ka, const float* B, int64_t N, int64_t K, int64_t norm_cols,
                                                 float alpha, const int64_t* labels, float* conf, int32_t* pred, uint8_t* correct,
                                                 const float* boundaries, int n_bins, int64_t* bin_count,
                                                 unsigned long long* bin_conf_fx, int64_t* bin_correct, int64_t* top1,
                                                 float* logits_out, int64_t ld_logits, void* stream) {
    if (M == 0) return CLIPGP_OK;
    CLIPGP_REQUIRE(norm_cols >= 0 && norm_cols % tc::BN == 0 && norm_cols < N,
                   "tc_logits_calibration_tf32: norm_cols (%lld) must be a multiple of %d below N", (long long)norm_cols, tc::BN);
    CLIPGP_REQUIRE(n_bins >= 0 && n_bins <= CLIPGP_MAX_BINS, "tc_logits_calibration_tf32: n_bins must be in [0,%d]", CLIPGP_MAX_BINS);
    CLIPGP_REQUIRE(bin_count == nullptr || (boundaries && bin_conf_fx && bin_correct && n_bins >= 1), "tc_logits_calibration_tf32: histogram outputs incomplete");
    CLIPGP_REQUIRE(labels != nullptr || (correct == nullptr && top1 == nullptr && bin_count == nullptr), "tc_logits_calibration_tf32: labels is NULL");
    CLIPGP_REQUIRE(logits_out == nullptr || (norm_cols == 0 && ld_logits >= N), "tc_logits_calibration_tf32: bad logits_out");
    tc::Params p = {};
    p.mode = tc::EPI_ROWSTATS; p.alpha = alpha; p.C = logits_out; p.ldc = ld_logits; p.norm_cols = (int)norm_cols;
    p.elt = 4;
    p.labels = reinterpret_cast<const long long*>(labels); p.conf = conf; p.pred = pred; p.correct = correct;
    p.boundaries = boundaries; p.n_bins = n_bins;
    p.bin_count = reinterpret_cast<long long*>(bin_count); p.bin_conf_fx = bin_conf_fx;
    p.bin_correct = reinterpret_cast<long long*>(bin_correct); p.top1 = reinterpret_cast<long long*>(top1);
    return tc::launch(A, M, Ka, B, N, K, p, (cudaStream_t)stream);
}

extern "C" int clipgp_tc_logits_calibration(const void* A_bf16, int64_t M, int64_t Ka, const void* B_bf16, int64_t N, int64_t K,
                                            float alpha, const int64_t* labels, float* conf, int32_t* pred, uint8_t* correct,
                                            const float* boundaries, int n_bins, int64_t* bin_count,
                                            unsigned long long* bin_conf_fx, int64_t* bin_correct, int64_t* top1,
                                            float* logits_out, int64_t ld_logits, void* stream) {
    if (M == 0) return CLIPGP_OK;
    CLIPGP_REQUIRE(n_bins >= 0 && n_bins <= CLIPGP_MAX_BINS, "tc_logits_calibration: n_bins must be in [0,%d]", CLIPGP_MAX_BINS);
    CLIPGP_REQUIRE(bin_count == nullptr || (boundaries && bin_conf_fx && bin_correct && n_bins >= 1), "tc_logits_calibration: histogram outputs incomplete");
    CLIPGP_REQUIRE(labels != nullptr || (correct == nullptr && top1 == nullptr && bin_count == nullptr), "tc_logits_calibration: labels is NULL");
    CLIPGP_REQUIRE(logits_out == nullptr || ld_logits >= N, "tc_logits_calibration: ld_logits < N");
    tc::Params p = {};
    p.mode = tc::EPI_ROWSTATS; p.alpha = alpha; p.C = logits_out; p.ldc = ld_logits;
    p.labels = reinterpret_cast<const long long*>(labels); p.conf = conf; p.pred = pred; p.correct = correct;
    p.boundaries = boundaries; p.n_bins = n_bins;
    p.bin_count = reinterpret_cast<long long*>(bin_count); p.bin_conf_fx = bin_conf_fx;
    p.bin_correct = reinterpret_cast<long long*>(bin_correct); p.top1 = reinterpret_cast<long long*>(top1);
    return tc::launch(A_bf16, M, Ka, B_bf16, N, K, p, (cudaStream_t)stream);
}

extern "C" int clipgp_tc_proj_logits_calibration(const void* A_bf16, int64_t M, int64_t Ka, const void* B_bf16, int64_t N_total, int64_t K,
                                                 int64_t norm_cols, float alpha, const int64_t* labels, float* conf, int32_t* pred,
                                                 uint8_t* correct, const float* boundaries, int n_bins, int64_t* bin_count,
                                                 unsigned long long* bin_conf_fx, int64_t* bin_correct, int64_t* top1, void* stream) {
    if (M == 0) return CLIPGP_OK;
    CLIPGP_REQUIRE(norm_cols > 0 && norm_cols % tc::BN == 0 && norm_cols < N_total,
                   "tc_proj_logits_calibration: norm_cols (%lld) must be a positive multiple of %d below N_total", (long long)norm_cols, tc::BN);
    CLIPGP_REQUIRE(n_bins >= 0 && n_bins <= CLIPGP_MAX_BINS, "tc_proj_logits_calibration: n_bins must be in [0,%d]", CLIPGP_MAX_BINS);
    CLIPGP_REQUIRE(bin_count == nullptr || (boundaries && bin_conf_fx && bin_correct && n_bins >= 1), "tc_proj_logits_calibration: histogram outputs incomplete");
    CLIPGP_REQUIRE(labels != nullptr || (correct == nullptr && top1 == nullptr && bin_count == nullptr), "tc_proj_logits_calibration: labels is NULL");
    tc::Params p = {};
    p.mode = tc::EPI_ROWSTATS; p.alpha = alpha; p.C = nullptr; p.ldc = 0; p.norm_cols = (int)norm_cols;
    p.labels = reinterpret_cast<const long long*>(labels); p.conf = conf; p.pred = pred; p.correct = correct;
    p.boundaries = boundaries; p.n_bins = n_bins;
    p.bin_count = reinterpret_cast<long long*>(bin_count); p.bin_conf_fx = bin_conf_fx;
    p.bin_correct = reinterpret_cast<long long*>(bin_correct); p.top1 = reinterpret_cast<long long*>(top1);
    return tc::launch(A_bf16, M, Ka, B_bf16, N_total, K, p, (cudaStream_t)stream);
}

extern "C" int clipgp_tc_tip_logits(const void* F_bf16, int64_t M, const void* keys_bf16, int64_t N_tr, int64_t K,
                                    const int32_t* key_class, float beta, float alpha, float* out, int64_t ldo, void* stream) {
    if (M == 0 || N_tr == 0) return CLIPGP_OK;
    CLIPGP_REQUIRE(key_class && out, "tc_tip_logits: NULL pointer");
    tc::Params p = {};
    p.mode = tc::EPI_TIP; p.alpha = 1.0f; p.C = out; p.ldc = ldo;
    p.key_class = key_class; p.beta = beta; p.tip_alpha = alpha;
    return tc::launch(F_bf16, M, K, keys_bf16, N_tr, K, p, (cudaStream_t)stream);
}
