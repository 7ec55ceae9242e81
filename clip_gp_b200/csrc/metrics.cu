// Calibration metrics kernels: softmax-max confidence + equal-width histogram (ECE), and the
// sort-free exact rank partition for AECE.  Replaces /root/reference/utils/metrics.py:9-229.
//
// Roofline: HBM.  calib_rows_kernel streams the [N,C] logits once (N*C*4 algorithmic bytes; the second
// pass over a row for sum-exp hits L1/L2), everything else is O(N).
#include <float.h>

#include <stdlib.h>
#include "common.cuh"

namespace clipgp {

static constexpr double kFx = 1099511627776.0;  // 2^40: conf in [2^-16,1] is an exact integer at this scale

__device__ __forceinline__ unsigned long long conf_to_fx(float c) {
    return (unsigned long long)((double)c * kFx);
}

__device__ __forceinline__ int find_bin(float cf, const float* b, int n_bins) {
    // metrics.py:77  in_bin = (conf > b[i]) * (conf <= b[i+1]);  conf == 0 or NaN falls in no bin
    for (int i = 0; i < n_bins; ++i)
        if (cf > b[i] && cf <= b[i + 1]) return i;
    return -1;
}

struct HistSmem {
    float b[CLIPGP_MAX_BINS + 1];
    unsigned int cnt[CLIPGP_MAX_BINS];
    unsigned int cor[CLIPGP_MAX_BINS];
    unsigned long long fx[CLIPGP_MAX_BINS];
    unsigned int top1;
};

__device__ __forceinline__ void hist_init(HistSmem& h, const float* boundaries, int n_bins) {
    for (int i = threadIdx.x; i < CLIPGP_MAX_BINS; i += blockDim.x) {
        h.cnt[i] = 0; h.cor[i] = 0; h.fx[i] = 0ull;
    }
    if (boundaries != nullptr)
        for (int i = threadIdx.x; i <= n_bins; i += blockDim.x) h.b[i] = boundaries[i];
    if (threadIdx.x == 0) h.top1 = 0;
    __syncthreads();
}

__device__ __forceinline__ void hist_flush(HistSmem& h, int n_bins, int64_t* bin_count,
                                           unsigned long long* bin_conf_fx, int64_t* bin_correct, int64_t* top1) {
    __syncthreads();
    if (bin_count != nullptr) {
        for (int i = threadIdx.x; i < n_bins; i += blockDim.x) {
            if (h.cnt[i]) {
                atomicAdd((unsigned long long*)&bin_count[i], (unsigned long long)h.cnt[i]);
                atomicAdd(&bin_conf_fx[i], h.fx[i]);
                atomicAdd((unsigned long long*)&bin_correct[i], (unsigned long long)h.cor[i]);
            }
        }
    }
    if (top1 != nullptr && threadIdx.x == 0 && h.top1) atomicAdd((unsigned long long*)top1, (unsigned long long)h.top1);
}

// One warp per logits row.  VEC: rows are 16-byte aligned and C % 4 == 0 -> float4 loads.
template <bool VEC>
__global__ void __launch_bounds__(256) calib_rows_kernel(const float* __restrict__ logits, int64_t ld,
                                                         const int64_t* __restrict__ labels, int64_t N, int64_t C,
                                                         float* __restrict__ conf, int32_t* __restrict__ pred,
                                                         uint8_t* __restrict__ correct,
                                                         const float* __restrict__ boundaries, int n_bins,
                                                         int64_t* bin_count, unsigned long long* bin_conf_fx,
                                                         int64_t* bin_correct, int64_t* top1) {
    __shared__ HistSmem h;
    hist_init(h, boundaries, n_bins);
    const int lane = threadIdx.x & 31;
    const int64_t warps_per_block = blockDim.x >> 5;
    const int64_t gw = (int64_t)blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    const int64_t nw = (int64_t)gridDim.x * warps_per_block;
    for (int64_t row = gw; row < N; row += nw) {
        const float* x = logits + row * ld;
        float m = -FLT_MAX;
        int am = 0x7fffffff;
        float s = 0.f;
        if (VEC && C <= 1024) {
            // the whole row in registers: eight 16-byte loads in flight per lane, one pass over HBM, sum-exp from registers
            const float4* x4 = reinterpret_cast<const float4*>(x);
            const int C4 = (int)(C >> 2);
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int j = lane + 32 * u;
                v[u] = (j < C4) ? __ldg(x4 + j) : make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int base = (lane + 32 * u) << 2;
                if (v[u].x > m) { m = v[u].x; am = base; }
                if (v[u].y > m) { m = v[u].y; am = base + 1; }
                if (v[u].z > m) { m = v[u].z; am = base + 2; }
                if (v[u].w > m) { m = v[u].w; am = base + 3; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                float om = __shfl_xor_sync(0xffffffffu, m, o);
                int oa = __shfl_xor_sync(0xffffffffu, am, o);
                if (om > m || (om == m && oa < am)) { m = om; am = oa; }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (lane + 32 * u < C4) s += expf(v[u].x - m) + expf(v[u].y - m) + expf(v[u].z - m) + expf(v[u].w - m);
        } else {
        if (VEC) {
            const float4* x4 = reinterpret_cast<const float4*>(x);
            const int C4 = (int)(C >> 2);
            for (int j = lane; j < C4; j += 32) {
                float4 v = __ldg(x4 + j);
                const int base = j << 2;
                if (v.x > m) { m = v.x; am = base; }
                if (v.y > m) { m = v.y; am = base + 1; }
                if (v.z > m) { m = v.z; am = base + 2; }
                if (v.w > m) { m = v.w; am = base + 3; }
            }
        } else {
            for (int j = lane; j < C; j += 32) {
                float v = __ldg(x + j);
                if (v > m) { m = v; am = j; }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            float om = __shfl_xor_sync(0xffffffffu, m, o);
            int oa = __shfl_xor_sync(0xffffffffu, am, o);
            if (om > m || (om == m && oa < am)) { m = om; am = oa; }
        }
        if (VEC) {
            const float4* x4 = reinterpret_cast<const float4*>(x);
            const int C4 = (int)(C >> 2);
            for (int j = lane; j < C4; j += 32) {
                float4 v = __ldg(x4 + j);
                s += expf(v.x - m) + expf(v.y - m) + expf(v.z - m) + expf(v.w - m);
            }
        } else {
            for (int j = lane; j < C; j += 32) s += expf(__ldg(x + j) - m);
        }
        }
        s = warp_sum(s);
        if (lane == 0) {
            const float cf = 1.0f / s;   // == max_j softmax_j (exp(0)/sum)
            const int ok = (labels != nullptr) ? ((int64_t)am == labels[row]) : 0;
            if (conf) conf[row] = cf;
            if (pred) pred[row] = am;
            if (correct) correct[row] = (uint8_t)ok;
            if (ok) atomicAdd(&h.top1, 1u);
            if (bin_count != nullptr) {
                const int bi = find_bin(cf, h.b, n_bins);
                if (bi >= 0) {
                    atomicAdd(&h.cnt[bi], 1u);
                    atomicAdd(&h.cor[bi], (unsigned int)ok);
                    atomicAdd(&h.fx[bi], conf_to_fx(cf));
                }
            }
        }
    }
    hist_flush(h, n_bins, bin_count, bin_conf_fx, bin_correct, top1);
}

__global__ void __launch_bounds__(256) ece_hist_kernel(const float* __restrict__ conf, const uint8_t* __restrict__ correct,
                                                       int64_t N, const float* __restrict__ boundaries, int n_bins,
                                                       int64_t* bin_count, unsigned long long* bin_conf_fx,
                                                       int64_t* bin_correct) {
    __shared__ HistSmem h;
    hist_init(h, boundaries, n_bins);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        const float cf = conf[i];
        const int bi = find_bin(cf, h.b, n_bins);
        if (bi >= 0) {
            atomicAdd(&h.cnt[bi], 1u);
            atomicAdd(&h.cor[bi], (unsigned int)correct[i]);
            atomicAdd(&h.fx[bi], conf_to_fx(cf));
        }
    }
    hist_flush(h, n_bins, bin_count, bin_conf_fx, bin_correct, nullptr);
}

// ---------------------------------------------------------------------------------------------------
// AECE: exact prefix sums at rank edges by 5-level (8 bits each) radix select over key40 = conf_bits<<8|correct.
// Elements with equal key have equal (conf, correct), so a rank cut inside a run of equal keys is exact.
//
// Multi-CTA: every CTA histograms its slice of the images into shared memory (warp-aggregated: confidences share their
// leading bytes, so a plain atomicAdd per element serialises on one or two counters), merges the non-empty counters into the
// level's global histogram, and after a grid barrier every CTA derives the same digit choice from it.  The grid is at most
// one CTA per SM and is launched cooperatively, so all CTAs are co-resident and the spin barrier cannot deadlock.
// ---------------------------------------------------------------------------------------------------
static constexpr int kQ = 10;        // edges resolved per sweep (the default n_bins = 10 has 9 interior edges: one sweep)
static constexpr int kLevels = 5;

struct AeceBin { unsigned int cnt, cor; unsigned long long fx; };
struct AeceWs {                      // zero-initialised by the launcher
    unsigned long long tot_fx, tot_cor;
    unsigned int barrier, pad[3];
    // followed by AeceBin hist[sweeps][kLevels][kQ][256]
};

__device__ __forceinline__ void grid_barrier(unsigned int* ctr, unsigned int& target) {
    target += gridDim.x;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(ctr, 1u);
        unsigned int v;
        do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory"); } while (v < target);
        __threadfence();
    }
    __syncthreads();
}

__global__ void __launch_bounds__(1024) aece_select_kernel(const float* __restrict__ conf,
                                                           const uint8_t* __restrict__ correct, int64_t N,
                                                           const int64_t* __restrict__ edges, int n_bins, AeceWs* ws,
                                                           unsigned long long* out_conf_fx, int64_t* out_correct,
                                                           int64_t* out_count) {
    __shared__ AeceBin h[kQ][256];
    __shared__ unsigned long long q_prefix[kQ], q_fx_less[kQ], slot_prefix[kQ];
    __shared__ long long q_rank[kQ], q_cnt_less[kQ], q_cor_less[kQ];
    __shared__ int q_slot[kQ], q_edge[kQ];
    __shared__ int n_slots, nq, scan_pos;
    __shared__ unsigned long long F_fx[CLIPGP_MAX_BINS + 1];
    __shared__ long long F_cor[CLIPGP_MAX_BINS + 1];
    __shared__ long long e_clamped[CLIPGP_MAX_BINS + 1];
    __shared__ unsigned long long tot_fx;
    __shared__ unsigned long long tot_cor;

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int64_t gstride = (int64_t)gridDim.x * blockDim.x;
    const int64_t gfirst = (int64_t)blockIdx.x * blockDim.x;         // warp-uniform trip counts: i - lane is the same for a warp
    AeceBin* ghist = reinterpret_cast<AeceBin*>(ws + 1);
    unsigned int bar_target = 0;
    if (tid == 0) { tot_fx = 0ull; tot_cor = 0ull; }
    __syncthreads();
    {   // totals (prefix sum at rank N)
        unsigned long long fx = 0ull, cor = 0ull;
        for (int64_t i = gfirst + tid; i < N; i += gstride) { fx += conf_to_fx(conf[i]); cor += correct[i]; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            fx += __shfl_xor_sync(0xffffffffu, fx, o);
            cor += __shfl_xor_sync(0xffffffffu, cor, o);
        }
        if (lane == 0) { atomicAdd(&tot_fx, fx); atomicAdd(&tot_cor, cor); }
    }
    __syncthreads();
    if (tid == 0) { atomicAdd(&ws->tot_fx, tot_fx); atomicAdd(&ws->tot_cor, tot_cor); }
    grid_barrier(&ws->barrier, bar_target);
    if (tid <= n_bins) {
        long long e = edges[tid];
        e = e < 0 ? 0 : (e > N ? N : e);
        e_clamped[tid] = e;
        if (e <= 0) { F_fx[tid] = 0ull; F_cor[tid] = 0; }
        else if (e >= N) { F_fx[tid] = __ldcg(&ws->tot_fx); F_cor[tid] = (long long)__ldcg(&ws->tot_cor); }
    }
    __syncthreads();

    int next_edge = 0;  // uniform across the grid (same inputs)
    for (int sweep = 0;; ++sweep) {
        // gather up to kQ interior edges that still need a select
        if (tid == 0) {
            int c = 0, k = next_edge;
            for (; k <= n_bins && c < kQ; ++k) {
                long long e = e_clamped[k];
                if (e > 0 && e < N) {
                    q_edge[c] = k; q_rank[c] = e; q_prefix[c] = 0ull; q_cnt_less[c] = 0; q_cor_less[c] = 0;
                    q_fx_less[c] = 0ull; q_slot[c] = 0; ++c;
                }
            }
            nq = c; n_slots = (c > 0) ? 1 : 0; slot_prefix[0] = 0ull;
            scan_pos = k;
        }
        __syncthreads();
        if (nq == 0) break;
        next_edge = scan_pos;

        for (int level = 0; level < kLevels; ++level) {
            const int shift = 32 - 8 * level;
            const int ns = n_slots;
            AeceBin* gl = ghist + ((size_t)sweep * kLevels + level) * kQ * 256;
            for (int i = tid; i < ns * 256; i += blockDim.x) { (&h[0][0])[i].cnt = 0u; (&h[0][0])[i].cor = 0u; (&h[0][0])[i].fx = 0ull; }
            __syncthreads();
            for (int64_t i0 = gfirst + (tid - lane); i0 < N; i0 += gstride) {
                const int64_t i = i0 + lane;
                const bool valid = i < N;
                const float cf = valid ? conf[i] : 0.f;
                const unsigned int cr = valid ? correct[i] : 0u;
                const unsigned long long key = ((unsigned long long)__float_as_uint(cf) << 8) | cr;
                const unsigned long long hi = (level == 0) ? 0ull : (key >> (shift + 8));
                const int digit = (int)((key >> shift) & 255ull);
                int slot = -1;
                for (int u = 0; u < ns; ++u) if (hi == slot_prefix[u]) slot = u;      // the slot prefixes are distinct
                const bool in = valid && slot >= 0;
                const unsigned int gid = in ? (unsigned int)(slot * 256 + digit) : 0x80000000u;   // ONE shared group for the lanes that sit this level out
                //                    (a distinct id per idle lane made match_any resolve 32 groups in every iteration of the deep levels: 117 -> 106 us)
                const unsigned int m = __match_any_sync(0xffffffffu, gid);
                if (in) {
                    const unsigned long long fx = conf_to_fx(cf);                      // <= 2^40: two 20-bit halves sum in 32 bits
                    const unsigned int scor = __reduce_add_sync(m, cr);
                    const unsigned int slo = __reduce_add_sync(m, (unsigned int)(fx & 0xFFFFFull));
                    const unsigned int shi = __reduce_add_sync(m, (unsigned int)(fx >> 20));
                    if (lane == __ffs(m) - 1) {
                        atomicAdd(&h[slot][digit].cnt, (unsigned int)__popc(m));
                        if (scor) atomicAdd(&h[slot][digit].cor, scor);
                        atomicAdd(&h[slot][digit].fx, ((unsigned long long)shi << 20) + slo);
                    }
                }
            }
            __syncthreads();
            if (gridDim.x > 1) {
                for (int i = tid; i < ns * 256; i += blockDim.x) {
                    const AeceBin v = (&h[0][0])[i];
                    if (v.cnt) {
                        atomicAdd(&gl[i].cnt, v.cnt);
                        if (v.cor) atomicAdd(&gl[i].cor, v.cor);
                        atomicAdd(&gl[i].fx, v.fx);
                    }
                }
                grid_barrier(&ws->barrier, bar_target);
            }
            if (wid < nq) {                        // one warp per edge: 8 digits per lane, warp scan, the owning lane walks its 8
                const int u = q_slot[wid];
                const long long r = q_rank[wid];
                unsigned int c8[8], o8[8];
                unsigned long long f8[8];
                long long csum = 0, osum = 0;
                unsigned long long fsum = 0ull;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int bidx = u * 256 + lane * 8 + j;
                    if (gridDim.x > 1) { c8[j] = __ldcg(&gl[bidx].cnt); o8[j] = __ldcg(&gl[bidx].cor); f8[j] = __ldcg(&gl[bidx].fx); }
                    else { const AeceBin v = (&h[0][0])[bidx]; c8[j] = v.cnt; o8[j] = v.cor; f8[j] = v.fx; }
                    csum += c8[j]; osum += o8[j]; fsum += f8[j];
                }
                long long cinc = csum, oinc = osum;
                unsigned long long finc = fsum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const long long c2 = __shfl_up_sync(0xffffffffu, cinc, o), o2 = __shfl_up_sync(0xffffffffu, oinc, o);
                    const unsigned long long f2 = __shfl_up_sync(0xffffffffu, finc, o);
                    if (lane >= o) { cinc += c2; oinc += o2; finc += f2; }
                }
                const bool mine = (cinc - csum <= r) && (r < cinc);
                const unsigned int found = __ballot_sync(0xffffffffu, mine);
                if (mine || (found == 0u && lane == 31)) {
                    long long cum = mine ? cinc - csum : cinc, cumcor = mine ? oinc - osum : oinc;
                    unsigned long long cumfx = mine ? finc - fsum : finc;
                    int sel = 255;
                    if (mine) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            if (cum + (long long)c8[j] > r) { sel = lane * 8 + j; break; }
                            cum += c8[j]; cumcor += o8[j]; cumfx += f8[j];
                        }
                    }
                    q_cnt_less[wid] += cum; q_cor_less[wid] += cumcor; q_fx_less[wid] += cumfx;
                    q_rank[wid] = r - cum;
                    q_prefix[wid] = (q_prefix[wid] << 8) | (unsigned long long)sel;
                }
            }
            __syncthreads();
            if (tid == 0) {   // dedupe prefixes -> histogram slots for the next level
                int ns2 = 0;
                for (int q = 0; q < nq; ++q) {
                    int found = -1;
                    for (int u = 0; u < ns2; ++u) if (slot_prefix[u] == q_prefix[q]) { found = u; break; }
                    if (found < 0) { slot_prefix[ns2] = q_prefix[q]; found = ns2++; }
                    q_slot[q] = found;
                }
                n_slots = ns2;
            }
            __syncthreads();
        }
        if (tid < nq) {
            const unsigned long long v = q_prefix[tid];
            const float vconf = __uint_as_float((unsigned int)(v >> 8));
            const long long rem = q_rank[tid];   // copies of v inside the prefix
            const int k = q_edge[tid];
            F_fx[k] = q_fx_less[tid] + (unsigned long long)rem * conf_to_fx(vconf);
            F_cor[k] = q_cor_less[tid] + rem * (long long)(v & 255ull);
        }
        __syncthreads();
    }
    if (blockIdx.x == 0 && tid < n_bins) {
        const long long lo = e_clamped[tid], hi = e_clamped[tid + 1];
        if (hi > lo) {
            out_conf_fx[tid] = F_fx[tid + 1] - F_fx[tid];
            out_correct[tid] = F_cor[tid + 1] - F_cor[tid];
            out_count[tid] = hi - lo;
        } else {
            out_conf_fx[tid] = 0ull; out_correct[tid] = 0; out_count[tid] = 0;
        }
    }
}

}  // namespace clipgp

using namespace clipgp;

extern "C" int clipgp_calibration_from_logits(const float* logits, int64_t ld_logits, const int64_t* labels, int64_t N,
                                              int64_t C, float* conf, int32_t* pred, uint8_t* correct,
                                              const float* boundaries, int n_bins, int64_t* bin_count,
                                              unsigned long long* bin_conf_fx, int64_t* bin_correct, int64_t* top1,
                                              void* stream) {
    CLIPGP_REQUIRE(N >= 0 && C >= 1, "calibration_from_logits: bad shape N=%lld C=%lld", (long long)N, (long long)C);
    CLIPGP_REQUIRE(C < (1ll << 31), "calibration_from_logits: C too large");
    if (N == 0) return CLIPGP_OK;
    CLIPGP_REQUIRE(logits != nullptr, "calibration_from_logits: logits is NULL");
    CLIPGP_REQUIRE(ld_logits >= C, "calibration_from_logits: ld_logits < C");
    const bool want_hist = bin_count != nullptr;
    if (want_hist) {
        CLIPGP_REQUIRE(boundaries && bin_conf_fx && bin_correct, "calibration_from_logits: histogram outputs incomplete");
        CLIPGP_REQUIRE(n_bins >= 1 && n_bins <= CLIPGP_MAX_BINS, "calibration_from_logits: n_bins must be in [1,%d]", CLIPGP_MAX_BINS);
    }
    CLIPGP_REQUIRE(labels != nullptr || (correct == nullptr && top1 == nullptr && !want_hist),
                   "calibration_from_logits: labels is NULL");
    const int threads = 256;
    const int64_t rows_per_block = threads / 32;
    int64_t blocks = (N + rows_per_block - 1) / rows_per_block;
    const bool vec = (C % 4 == 0) && (ld_logits % 4 == 0) && ((reinterpret_cast<uintptr_t>(logits) & 15u) == 0);
    // exactly one resident wave (the register-resident row keeps 5 CTAs per SM: a grid of 8 per SM ran 1.6 waves with an idle tail)
    static int occ_vec = 0, occ_scalar = 0;
    int& occ = vec ? occ_vec : occ_scalar;
    if (occ == 0) {
        if (vec) CLIPGP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, calib_rows_kernel<true>, threads, 0));
        else CLIPGP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, calib_rows_kernel<false>, threads, 0));
        if (occ < 1) occ = 1;
    }
    const int64_t cap = (int64_t)num_sms() * occ;
    if (blocks > cap) blocks = cap;
    cudaStream_t st = (cudaStream_t)stream;
    if (vec)
        calib_rows_kernel<true><<<(unsigned)blocks, threads, 0, st>>>(logits, ld_logits, labels, N, C, conf, pred, correct,
                                                                     boundaries, n_bins, bin_count, bin_conf_fx, bin_correct, top1);
    else
        calib_rows_kernel<false><<<(unsigned)blocks, threads, 0, st>>>(logits, ld_logits, labels, N, C, conf, pred, correct,
                                                                      boundaries, n_bins, bin_count, bin_conf_fx, bin_correct, top1);
    return check_launch("calib_rows_kernel");
}

extern "C" int clipgp_ece_hist(const float* conf, const uint8_t* correct, int64_t N, const float* boundaries, int n_bins,
                               int64_t* bin_count, unsigned long long* bin_conf_fx, int64_t* bin_correct, void* stream) {
    CLIPGP_REQUIRE(N >= 0, "ece_hist: N < 0");
    CLIPGP_REQUIRE(n_bins >= 1 && n_bins <= CLIPGP_MAX_BINS, "ece_hist: n_bins must be in [1,%d]", CLIPGP_MAX_BINS);
    if (N == 0) return CLIPGP_OK;
    CLIPGP_REQUIRE(conf && correct && boundaries && bin_count && bin_conf_fx && bin_correct, "ece_hist: NULL pointer");
    int64_t blocks = (N + 255) / 256;
    const int64_t cap = (int64_t)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    ece_hist_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(conf, correct, N, boundaries, n_bins, bin_count,
                                                                       bin_conf_fx, bin_correct);
    return check_launch("ece_hist_kernel");
}

static size_t aece_ws_bytes(int n_bins) {
    const int sweeps = (n_bins - 1 + kQ - 1) / kQ + 1;
    return sizeof(AeceWs) + (size_t)sweeps * kLevels * kQ * 256 * sizeof(AeceBin);
}

extern "C" int64_t clipgp_aece_workspace_bytes(int n_bins) { return (int64_t)aece_ws_bytes(n_bins < 1 ? 1 : n_bins); }

extern "C" int clipgp_aece_bins(const float* conf, const uint8_t* correct, int64_t N, const int64_t* edges, int n_bins,
                                unsigned long long* out_conf_fx, int64_t* out_correct, int64_t* out_count, void* workspace,
                                int64_t workspace_bytes, void* stream) {
    CLIPGP_REQUIRE(N >= 0, "aece_bins: N < 0");
    CLIPGP_REQUIRE(n_bins >= 1 && n_bins < CLIPGP_MAX_BINS, "aece_bins: n_bins must be in [1,%d)", CLIPGP_MAX_BINS);
    CLIPGP_REQUIRE(edges && out_conf_fx && out_correct && out_count, "aece_bins: NULL pointer");
    CLIPGP_REQUIRE(N == 0 || (conf && correct), "aece_bins: NULL input");
    // one CTA per 4096 images, at most one per SM (co-resident: the kernel synchronises the grid); stream-ordered zeroed workspace
    cudaStream_t st = (cudaStream_t)stream;
    int64_t blocks = (N + 4095) / 4096;
    blocks = blocks < 1 ? 1 : (blocks > (int64_t)num_sms() ? (int64_t)num_sms() : blocks);
    {
        static int forced = -2;
        if (forced == -2) { const char* e = getenv("CLIPGP_AECE_BLOCKS"); forced = e ? atoi(e) : -1; }
        if (forced > 0) blocks = forced;
    }
    const size_t ws_bytes = aece_ws_bytes(n_bins);
    void* ws = workspace;
    CLIPGP_REQUIRE(ws == nullptr || (size_t)workspace_bytes >= ws_bytes, "aece_bins: workspace too small (%lld < %lld bytes)",
                   (long long)workspace_bytes, (long long)ws_bytes);
    if (ws == nullptr) CLIPGP_CUDA(cudaMallocAsync(&ws, ws_bytes, st));
    CLIPGP_CUDA(cudaMemsetAsync(ws, 0, blocks > 1 ? ws_bytes : sizeof(AeceWs), st));
    // cooperative launch: the runtime places the whole grid or nothing, so the in-kernel grid barrier cannot wait for CTAs that
    // another stream's kernels keep off the SMs; a grid the device cannot hold at once falls back to one CTA (no barrier needed)
    AeceWs* wsp = reinterpret_cast<AeceWs*>(ws);
    void* kargs[] = {(void*)&conf, (void*)&correct, (void*)&N, (void*)&edges, (void*)&n_bins, (void*)&wsp, (void*)&out_conf_fx,
                     (void*)&out_correct, (void*)&out_count};
    cudaError_t le = cudaErrorUnknown;
    if (blocks > 1) le = cudaLaunchCooperativeKernel((const void*)aece_select_kernel, dim3((unsigned)blocks), dim3(1024), kargs, 0, st);
    if (le != cudaSuccess) {
        (void)cudaGetLastError();
        aece_select_kernel<<<1, 1024, 0, st>>>(conf, correct, N, edges, n_bins, wsp, out_conf_fx, out_correct, out_count);
    }
    const int rc = check_launch("aece_select_kernel");
    if (workspace == nullptr) CLIPGP_CUDA(cudaFreeAsync(ws, st));
    return rc;
}
