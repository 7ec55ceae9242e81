/*
 * clipgp.h — C ABI of the B200-native CLIP-GP few-shot adapter hot path (libclipgp.so).
 *
 * The reference (paulmerceur/CLIP-GP) is pure Python/PyTorch and has no FFI today; each entry point
 * below replaces the torch / gpytorch / entmax call sequence of the cited reference lines
 * (paths relative to the reference root).  Conventions:
 *   - every pointer is a DEVICE pointer unless the name ends in _host; sizes are int64_t;
 *   - tensors are dense row-major fp32 unless stated; labels are int64;
 *   - `stream` is a cudaStream_t passed as void* (the caller's current torch stream);
 *   - the return value is 0 on success, non-zero on failure; clipgp_last_error() then returns a
 *     thread-local message.  There is NO CPU fallback: without a CUDA device every compute entry fails;
 *   - no entry point allocates device memory or synchronises the stream unless documented.
 */
#ifndef CLIPGP_H_
#define CLIPGP_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CLIPGP_OK 0
#define CLIPGP_ERR_INVALID 1 /* bad argument (shape, null pointer, unsupported size) */
#define CLIPGP_ERR_CUDA 2    /* CUDA runtime / launch failure */

#define CLIPGP_KERNEL_RBF 0      /* ScaleKernel(RBFKernel(ARD))      gp_template_weigher.py:102-114 */
#define CLIPGP_KERNEL_MATERN12 1 /* MaternKernel(nu=0.5, ARD)        gp_template_weigher.py:115-117 */
#define CLIPGP_KERNEL_LINEAR 2   /* LinearKernel                     gp_template_weigher.py:118-120 */

#define CLIPGP_MAX_BINS 64
#define CLIPGP_GP_MAX_T 64 /* templates per class; inducing points n = T+1 <= 65 */

/* ------------------------------------------------------------------------------------------------ */
const char* clipgp_last_error(void);
int clipgp_version(void);
/* Number of kernels this library has launched since load (bench.py's gpu_launches counter). */
int64_t clipgp_launch_count(void);

/* ================================================================================================
 * Calibration metrics — utils/metrics.py
 * ================================================================================================ */

/* metrics.py:71-73 (softmax -> max -> eq) fused with the 10-bin equal-width histogram of :75-82 and
 * the top-1 count of compute_accuracy (:9-36), one pass over the logits.
 *   logits [N,C] (row stride ld_logits elements), labels [N].
 *   conf [N] = max softmax prob, pred [N] (int32) = argmax (lowest index on ties), correct [N] (uint8);
 *     any of the three may be NULL.
 *   boundaries [n_bins+1] fp32 (torch.linspace(0,1,n_bins+1)); membership b[i] < conf <= b[i+1].
 *   bin_count [n_bins] int64, bin_conf_fx [n_bins] uint64 (sum of conf in 2^-40 fixed point: exact and
 *     order independent), bin_correct [n_bins] int64, top1 [1] int64 — all ACCUMULATED into (+=), so a
 *     caller can stream chunks / shards; zero them first.  The hist pointers may be NULL (conf only). */
int clipgp_calibration_from_logits(const float* logits, int64_t ld_logits, const int64_t* labels, int64_t N,
                                   int64_t C, float* conf, int32_t* pred, uint8_t* correct,
                                   const float* boundaries, int n_bins, int64_t* bin_count,
                                   unsigned long long* bin_conf_fx, int64_t* bin_correct, int64_t* top1,
                                   void* stream);

/* Equal-width histogram of precomputed (conf, correct) — same accumulators as above. */
int clipgp_ece_hist(const float* conf, const uint8_t* correct, int64_t N, const float* boundaries, int n_bins,
                    int64_t* bin_count, unsigned long long* bin_conf_fx, int64_t* bin_correct, void* stream);

/* metrics.py:107-133 (sort by confidence, contiguous equal-count rank bins) without materialising the
 * sort: an exact 5-level radix select of the n_bins-1 interior rank edges over the 40-bit key
 * (conf_bits<<8 | correct), then prefix sums at those ranks.
 *   edges [n_bins+1] int64 DEVICE (torch.linspace(0,N,n_bins+1).round().long(), edges[0]=0, edges[-1]=N)
 *   out_conf_fx [n_bins] uint64 (2^-40 fixed point), out_correct [n_bins] int64, out_count [n_bins] int64
 *   — overwritten.  One CTA per 4096 images (at most one per SM) with grid-wide barriers between the levels; the per-level
 *   histograms live in `workspace` (device, >= clipgp_aece_workspace_bytes(n_bins) bytes, contents irrelevant on entry).
 *   workspace == NULL: a stream-ordered allocation (cudaMallocAsync / cudaFreeAsync) is made for the call. */
int64_t clipgp_aece_workspace_bytes(int n_bins);
int clipgp_aece_bins(const float* conf, const uint8_t* correct, int64_t N, const int64_t* edges, int n_bins,
                     unsigned long long* out_conf_fx, int64_t* out_correct, int64_t* out_count, void* workspace,
                     int64_t workspace_bytes, void* stream);

/* ================================================================================================
 * GP template weighter — trainers/gp_template_weigher.py:166-222 + gpytorch VariationalStrategy
 * (whitened), MultivariateNormal.rsample, entmax.sparsemax, kl_divergence.
 * One CTA per class.  Per class: K_ZZ(+1e-4 I), K_ZX, K_XX -> L = chol64(K_ZZ) -> A = L^-1 K_ZX ->
 * mu = A^T m + mean_x, Sigma = K_XX + 1e-4 I + A^T (Lq Lq^T - I) A -> R = chol32(Sigma) (psd_safe jitter
 * retries 1e-6,1e-5,1e-4) -> f_s = mu + R eps_s -> w_s = sparsemax(f_s);  KL(q(u)||N(0,I)).
 * ================================================================================================ */
typedef struct clipgp_gp_args {
    int32_t kernel_type;          /* CLIPGP_KERNEL_* */
    int32_t x_is_z_prefix;        /* X[c] == Z[c,:T] bit-for-bit lets the kernel evaluate only K_ZZ: 0 never assume it,
                                     1 test it on the device per class, 2 the caller guarantees it (no test) */
    int64_t C, T, n, d, S;        /* classes, templates, inducing points (T+1), kernel input dim, MC samples */
    const float* Z;               /* [C,n,d] variational_strategy.inducing_points */
    const float* X;               /* [C,T,d] _templates_red (test inputs) */
    const float* raw_lengthscale; /* [C,d]  (rbf, matern) else NULL; lengthscale = softplus(raw) */
    const float* raw_outputscale; /* [C]    (rbf) else NULL */
    const float* raw_variance;    /* [C]    (linear) else NULL */
    const float* var_mean;        /* [C,n]   variational_mean */
    const float* chol_var;        /* [C,n,n] chol_variational_covar (raw; lower triangle is used) */
    const float* mean_x;          /* [C,T] prior mean at the test inputs, or NULL for 0 */
    const float* eps;             /* explicit base noise, element (c,t,s) at eps[c*eps_sc + t*eps_st + s*eps_ss]; */
    int64_t eps_sc, eps_st, eps_ss; /*   NULL -> counter RNG: Philox4x32-10(seed, step)[(c*T+t)*S_total + s_offset+s] */
    const uint64_t* rng_state;    /* device [2] = {seed, step}; used when eps == NULL */
    int64_t s_offset, S_total;    /* this rank's slice of the MC samples (S-sharding keeps the draws identical) */
    float* w;                     /* out [S,C,T] template weights (gp_weighter.scores) */
    float* kl;                    /* out [C] KL(q(u) || N(0,I)), or NULL */
    double* L;                    /* out [C,n,n] saved: lower Cholesky factor of K_ZZ + 1e-4 I (fp64) */
    float* A;                     /* out [C,n,T] saved: interp_term L^-1 K_ZX */
    float* R;                     /* out [C,T,T] saved: lower Cholesky factor of Sigma */
    int32_t* status;              /* out [C]: 0 ok; k>0: Sigma needed k jitter retries; <0: not positive definite */
    float* Ksave;                 /* out [C, 1 + n*n + n*T + T*T] saved kernel blocks: [alias flag | K_ZZ (no jitter) | K_ZX | K_XX];
                                     K_ZX / K_XX are written only when the alias flag is 0.  Required by clipgp_gp_backward
                                     (which uses the K_ZX / K_XX part of aliased classes as scratch) */
    int64_t c_begin, c_count;     /* class shard: only classes [c_begin, c_begin + c_count) are processed (c_count == 0: through C-1);
                                     every pointer still addresses the full C-class tensors, so multi-GPU class sharding needs no
                                     re-layout and the Philox draws do not depend on the shard */
    float* eps_save;              /* optional [S,C,T] (layout of w): with the counter RNG, the warp-path forward kernel stores the base
                                     noise it drew and the warp-path adjoint reads it back instead of regenerating the Philox
                                     stream (the general kernels ignore it and regenerate) */
    /* Optional fused prototype stage of the warp-path forward kernel (all NULL / 0: off).  The CTA that produced w[:, c, :] also
     * builds the unit prototypes of class c (gp_template_weigher.py:221 + F.normalize, adapter.py:425), i.e. what
     * clipgp_proto_forward + clipgp_cast_bf16 would do, without a grid-wide barrier in between: */
    const float* proto_E;         /* [C,T,D] text bank (D % 4 == 0, D <= 1024) */
    int64_t proto_D;
    float* proto_P_hat;           /* out [S,C,D] unit prototypes (fp32) */
    float* proto_norm;            /* out [S,C]   |P_raw| */
    void* proto_bf16;             /* out, optional: bf16 operand rows (s*C + c) with the layouts of clipgp_cast_bf16 */
    int64_t proto_bf16_ld, proto_bf16_seg;
    int32_t proto_bf16_mode;
    float* proto_mean_hat;        /* out, optional [C,D]: (1/S) sum_s P_hat[s,c,:], the collapsed prototype of the logit-mean eval
                                     (adapter.py:243-249); proto_P_hat may then be NULL */
} clipgp_gp_args;

/* Dynamic shared memory the forward / backward kernel needs for (T, n, d); 0 if unsupported. */
int64_t clipgp_gp_smem_bytes(int64_t T, int64_t n, int64_t d, int backward);

int clipgp_gp_forward(const clipgp_gp_args* args, void* stream);
/* 1 if the fused prototype stage (clipgp_gp_args.proto_*) can run for these sizes (warp path, D <= 512, S <= 136); the caller must
 * also pass x_is_z_prefix == 2 and process all classes in one launch.  Otherwise use clipgp_proto_forward. */
int clipgp_gp_fused_proto_ok(int64_t T, int64_t n, int64_t d, int64_t D, int64_t S);
/* 1 if (T, n, d) is served by the warp-per-class fast path (2 <= T <= 32, n == T+1, d % 4 == 0). */
int clipgp_gp_warp_path_ok(int64_t T, int64_t n, int64_t d);

/* Adjoint of clipgp_gp_forward.  `fwd` must be the argument block of the forward call (same inputs, with
 * w/L/A/R holding its outputs).  Upstream: dw [S,C,T]; dkl [C] or NULL (then dkl_scalar multiplies every
 * class, i.e. loss += dkl_scalar * sum_c KL_c).  Outputs are OVERWRITTEN:
 *   dZ_last [C,d]  gradient of the learnable inducing row Z[:, n-1] (rows < T are masked by the reference,
 *                  gp_template_weigher.py:72-79),
 *   draw_lengthscale [C,d], draw_outputscale [C], draw_variance [C] (those the kernel type has; NULL otherwise),
 *   dvar_mean [C,n], dchol_var [C,n,n] (lower triangle, zeros above), dmean_x [C,T] (may be NULL). */
typedef struct clipgp_gp_bwd_args {
    const float* dw;
    const float* dkl;
    float dkl_scalar;
    float* dZ_last;
    float* draw_lengthscale;
    float* draw_outputscale;
    float* draw_variance;
    float* dvar_mean;
    float* dchol_var;
    float* dmean_x;
    /* Optional fused prototype adjoint of the warp path (proto_dP == NULL: off, `dw` is the upstream gradient as above).  The CTA
     * of class c derives dw[:, c, :] itself from the gradient of the UNIT prototypes, i.e. what clipgp_proto_backward does, without a
     * grid-wide barrier in between and without reading P_hat:
     *     a[s,t] = <dP_hat[s,c,:], E[c,t,:]>,  q_s = <w_s, a_s> / |P_s|,  dw[s,t] = (a[s,t] - q_s (w_s G_c)[t] / |P_s|) / |P_s|
     * with G_c = E[c] E[c]^T (frozen, [C,T,T], supplied by the caller).  Needs S * (D + 64) * 4 <= 26208 bytes of shared memory. */
    const float* proto_dP;        /* [S,C,D] gradient of the unit prototypes (sample stride proto_dP_stride_s elements, 0 = broadcast) */
    int64_t proto_dP_stride_s;
    float proto_dP_scale;
    const float* proto_norm;      /* [S,C] |P_raw| from the forward pass */
    const float* proto_E;         /* [C,T,D] */
    const float* proto_EEt;       /* [C,T,T] */
    int64_t proto_D;
    float* dw_out;                /* optional [S,C,T]: the derived dw (for inspection) */
    /* Small-batch source of a[s,t] for the fused prototype adjoint (tl_Z == NULL: off, a comes from proto_dP as above):
     *     a[s,t] = tl_scale * sum_b dlogits[b,s,c] Zt[b,c,t],     Zt [B, C*T] = f_hat E_flat^T  (per-template cosines)
     * which is the same number as <dP_hat[s,c,:], E[c,t,:]> with dP_hat = scale dlogits^T f_hat, contracted in the other order.
     * Zt is ONE tensor-core GEMM on the feature branch (it does not depend on the GP), dlogits^T are the bf16 operand rows the
     * softmax kernel writes anyway, so the d P_hat GEMM leaves the critical path and the class CTA reads B*T*4 + S*B*6 bytes
     * instead of T*D*4 (24 KB instead of 64 KB at B = 128, T = 32, D = 512) for S*B*T instead of S*T*D multiply-adds. */
    const float* tl_Z;            /* [B, C*T] fp32, row stride tl_Z_ld */
    int64_t tl_Z_ld;
    const void* tl_dlT;           /* bf16 [S*C, tl_dlT_ld], row s*C + c; tl_mode 0: [v], 1: segments [hi|hi|lo] of tl_seg elements */
    int64_t tl_dlT_ld;
    int64_t tl_seg;
    int64_t tl_B;
    int tl_mode;
    float tl_scale;
} clipgp_gp_bwd_args;

/* 1 if the fused prototype adjoint can run for these sizes (warp path and the shared-memory bound above). */
int clipgp_gp_fused_proto_bwd_ok(int64_t T, int64_t n, int64_t d, int64_t D, int64_t S);

int clipgp_gp_backward(const clipgp_gp_args* fwd, const clipgp_gp_bwd_args* bwd, void* stream);

/* ================================================================================================
 * Weighted prototypes — gp_template_weigher.py:221 (einsum "skm,kmd->skd") fused with the row
 * normalisation of adapter.py:246/425, taskres.py:109-113, clip_adapter.py:94, tip_adapter.py:136.
 *   w [S,C,T], E [C,T,D] (frozen text bank, D % 4 == 0).
 *   residual [C,D] + alpha: TaskRes branch t_s = normalize(p_hat_s + alpha x) (taskres.py:111-113); NULL otherwise.
 *   Outputs (each may be NULL): P_raw [S,C,D] un-normalised prototypes (the reference's return value),
 *   P_hat [S,C,D] unit rows, norm [S,C] = |P_raw|, P_hat_bf16 [S,C,D] (operand of the tensor-core logit GEMM),
 *   mean_hat [C,D] = sum_s P_hat_s, mean_raw [C,D] = sum_s P_raw_s; with finish_mean != 0 they are finished
 *   to (1/S) sum_s P_hat_s (collapsed logit-mean prototype) and normalize((1/S) sum_s P_raw_s)
 *   (prototype init of taskres.py:281-285, clip_adapter.py:284-288, tip_adapter.py:152-156).
 * ================================================================================================ */
int clipgp_proto_forward(const float* w, const float* E, int64_t S, int64_t C, int64_t T, int64_t D,
                         const float* residual, float alpha, float* P_raw, float* P_hat, float* norm,
                         void* P_hat_bf16, float* mean_hat, float* mean_raw, int finish_mean, void* stream);

/* dw [S,C,T] = <dP[s,c,:], E[c,t,:]>.  Row (s,c) of the upstream gradient is dP_scale * dP[s*dP_stride_s + c*D ...]
 * (dP_stride_s = C*D for a dense [S,C,D] gradient, 0 to broadcast one [C,D] gradient over the samples, which with
 * dP_scale = 1/S is the adjoint of the logit-mean heads).  It is the gradient of P_raw when P_hat == NULL, else the
 * gradient of the unit rows P_hat (then dP = (dP_hat - P_hat <P_hat,dP_hat>) / norm is applied first). */
int clipgp_proto_backward(const float* dP, int64_t dP_stride_s, float dP_scale, const float* P_hat, const float* norm,
                          const float* E, int64_t S, int64_t C, int64_t T, int64_t D, float* dw, void* stream);

/* ================================================================================================
 * Cosine-logit heads — adapter.py:230-252 (forward_features), :401-428 (per-sample MC cross-entropy),
 * taskres.py:96-123, clip_adapter.py:85-100, tip_adapter.py:250-260 — exact-fp32 building blocks.
 * (The tensor-core path with fused epilogues is clipgp_tc_*.)
 * ================================================================================================ */

/* C[M,N] = alpha * op(A) op(B) (+ C if accumulate & 1), fp32 FFMA.  accumulate & 2: never split K (split-K adds partial sums with
 * atomics, so their order -- and the last bits of C -- vary from run to run; set-up code that must be reproducible sets this bit).  A(m,k) = A[m*sam + k*sak], B(k,n) = B[k*sbk + n*sbn];
 * each operand needs one unit stride.  Covers f W^T (adapter.py:239), f_hat P_hat^T (:248,:426), f keys^T
 * (tip_adapter.py:250) and their adjoints dlogits^T f_hat, dlogits P_hat. */
int clipgp_gemm_f32(const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbk, int64_t sbn, float* C,
                    int64_t ldc, int64_t M, int64_t N, int64_t K, float alpha, int accumulate, void* stream);

/* F.normalize(x, dim=-1) (adapter.py:240): y = x / max(|x|, 1e-12); inv_norm [R]; optional bf16 copy of y. */
int clipgp_rownorm_forward(const float* x, int64_t R, int64_t D, float* y, float* inv_norm, void* y_bf16, void* stream);
/* F.normalize fused with the operand cast of the next tensor-core GEMM: out_bf16[r*out_ld + g*seg_stride + k] in the layouts of
 * clipgp_cast_bf16 (mode 0 / 1 / 2); y (fp32 unit rows) and inv_norm may be NULL.  D % 4 == 0. */
int clipgp_rownorm_cast(const float* x, int64_t R, int64_t D, float* y, float* inv_norm, void* out_bf16, int64_t out_ld,
                        int64_t seg_stride, int mode, void* stream);
/* dx = (dy - y <y,dy>) * inv_norm. */
int clipgp_rownorm_backward(const float* dy, const float* y, const float* inv_norm, int64_t R, int64_t D, float* dx,
                            void* stream);

/* F.cross_entropy over R logits rows of C classes (adapter.py:427, taskres.py:270); row r uses labels[r / rows_per_label]
 * (rows_per_label = S for the [B,S,C] per-sample layout).  loss_rows [R] (may be NULL); loss_sum[0] += loss_scale * sum_r CE_r
 * (may be NULL); dlogits (may be NULL, may alias logits) = grad_scale * (softmax - onehot). */
int clipgp_softmax_ce(const float* logits, int64_t ld, const int64_t* labels, int64_t R, int64_t rows_per_label, int64_t C,
                      float* loss_rows, float* loss_sum, float loss_scale, float* dlogits, int64_t ldd, float grad_scale,
                      void* stream);

/* adapter.py:468-476: loss_sum[0] += coef * ||W - I||_F^2 ; dW += 2 coef (W - I)  (dW / loss_sum may be NULL). */
int clipgp_l2_identity(const float* W, int64_t D, float coef, float* dW, float* loss_sum, void* stream);

/* torch.optim.AdamW update (utils/optimization.py:57-216 builds it; adapter.py:549 steps it) on one flat fp32 buffer.
 * `step` is a DEVICE int64 (1-based) so captured CUDA graphs advance it with clipgp_increment. */
int clipgp_adamw_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                      float weight_decay, const int64_t* step, void* stream);
/* Same update with the learning rate read from DEVICE memory (lr_dev[0]): per-epoch / per-step schedules
 * (utils/optimization.py:218-281, stepped at adapter.py:1054-1056) advance inside a captured CUDA graph. */
int clipgp_adamw_step_lrptr(float* p, const float* g, float* m, float* v, int64_t n, const float* lr_dev, float beta1, float beta2,
                            float eps, float weight_decay, const int64_t* step, void* stream);
/* clipgp_adamw_step_lrptr on a [R, K] weight that also writes the UPDATED weight as the bf16 operand of the next tensor-core GEMM
 * (layouts / modes of clipgp_cast_bf16 below): the trainable cache keys of Tip-Adapter-F (tip_adapter.py:229-233) need no separate
 * cast pass per step. */
int clipgp_adamw_step_cast(float* p, const float* g, float* m, float* v, int64_t R, int64_t K, const float* lr_dev, float beta1,
                           float beta2, float eps, float weight_decay, const int64_t* step, void* out_bf16, int64_t out_ld,
                           int64_t seg_stride, int mode, void* stream);
int clipgp_increment(int64_t* counter, int64_t by, void* stream);
/* out[0] += scale * sum(x[0..n))   (e.g. gp_beta * sum_c KL_c, adapter.py:462-465). */
int clipgp_sum_accumulate(const float* x, int64_t n, float scale, float* out, void* stream);

/* ================================================================================================
 * Tensor-core (tcgen05 / TMEM / TMA) path of the same contractions: D[M,N] = alpha * A[M,K] B[N,K]^T, bf16 operands in
 * plain row-major [rows, K] (K % 8 == 0, 16-byte aligned), fp32 accumulation.  Ka <= K with K % Ka == 0 makes A wrap
 * along K (A column = k mod Ka, Ka % 64 == 0): with B = [p_hat_1 | ... | p_hat_S] along K the MC sum over samples
 * (adapter.py:247-249) accumulates inside TMEM.  fp32-grade products: feed the split operands of clipgp_cast_bf16.
 * ================================================================================================ */

/* fp32 [R,K] (row stride ldx) -> bf16.  mode 0: out[r*out_ld + k].  mode 1 / 2: three segments at out[r*out_ld + g*seg_stride + k]
 * holding [hi|hi|lo] (A side) / [hi|lo|hi] (B side), so that one K-tripled GEMM yields a_hi b_hi + a_hi b_lo + a_lo b_hi. */
int clipgp_cast_bf16(const float* x, int64_t R, int64_t K, int64_t ldx, void* out, int64_t out_ld, int64_t seg_stride, int mode,
                     void* stream);

/* Transposing form: fp32 [R,K] -> bf16 [K, R] (or the split [K, 3R] layouts) with out[k*out_ld + g*seg_stride + r]; produces the
 * K-major operands of the adjoint GEMMs (dlogits^T, f_hat^T, P_hat^T) without an fp32 transpose. */
/* out[k][r] = x[r][k] (fp32; row pitches ldx >= K, out_ld >= R).  The small operand of a long-K TF32 GEMM is handed to the tensor
 * cores K-major through this copy: an MN-major B operand costs the MMA pipeline ~35 % of its rate, the copy microseconds. */
int clipgp_transpose_f32(const float* x, int64_t R, int64_t K, int64_t ldx, float* out, int64_t out_ld, void* stream);
int clipgp_cast_bf16_transpose(const float* x, int64_t R, int64_t K, int64_t ldx, void* out, int64_t out_ld, int64_t seg_stride,
                               int mode, void* stream);

/* One read, both layouts: out[r*out_ld + g*seg_stride + k] (mode) and outT[k*outT_ld + g*segT_stride + r] (modeT); either may be
 * NULL.  Feeds the forward GEMM and the adjoint GEMM that contracts over r from the same fp32 source. */
int clipgp_cast_bf16_dual(const float* x, int64_t R, int64_t K, int64_t ldx, void* out, int64_t out_ld, int64_t seg_stride, int mode,
                          void* outT, int64_t outT_ld, int64_t segT_stride, int modeT, void* stream);

/* Tensor-core form of F.cross_entropy (adapter.py:427): phase 1 reduces every logits row to stats[r] = (max, 1/sum exp(x - max))
 * (float2) and accumulates loss_sum[0] += loss_scale * sum_r CE_r; phase 2 streams the logits [B, S*C] once more and writes
 * dlogits = grad_scale * (softmax - onehot) directly as the two bf16 operands of the adjoint GEMMs (row (b,s) uses labels[b]):
 * out [B, S*C] for d f_hat = dlogits P_hat and outT [S*C, B] for d P_hat = dlogits^T f_hat.  No fp32 dlogits reach HBM. */
int clipgp_softmax_ce_stats(const float* logits, int64_t ld, const int64_t* labels, int64_t R, int64_t rows_per_label, int64_t C,
                            float* stats, float* loss_sum, float loss_scale, void* stream);
int clipgp_softmax_grad_bf16_dual(const float* logits, const float* stats, const int64_t* labels, int64_t B, int64_t S, int64_t C,
                                  float grad_scale, void* out, int64_t out_ld, int64_t seg_stride, int mode, void* outT,
                                  int64_t outT_ld, int64_t segT_stride, int modeT, void* stream);
/* Both phases in one launch (a CTA owns 64 batch rows of one MC sample; the second read of its logits block comes from L2):
 * loss_sum[0] += loss_scale * sum CE, and dlogits = grad_scale (softmax - onehot) in both bf16 operand layouts as above. */
int clipgp_softmax_ce_bf16_dual(const float* logits, const int64_t* labels, int64_t B, int64_t S, int64_t C, float* loss_sum,
                                float loss_scale, float grad_scale, void* out, int64_t out_ld, int64_t seg_stride, int mode, void* outT,
                                int64_t outT_ld, int64_t segT_stride, int modeT, void* stream);
int clipgp_increment2(int64_t* a, int64_t* b, int64_t by, void* stream);
/* Tail of the single-GPU optimisation step in ONE launch (adapter.py:462-465 KL term, :537-549 optimizer.step() of the gp_weighter group,
 * gp_template_weigher.py:72-79 learnable inducing row): AdamW over p / g / m / v [n] (16-byte aligned; arithmetic of
 * clipgp_adamw_step_lrptr, lr_dev[0]); elements [z_off, z_off + C d) are also scattered to Z[c, nrows - 1, :]; loss[0] += kl_scale *
 * sum(kl[0 .. kl_n)) (kl may be NULL); then *step += by and *counter_b += by.  ticket: device uint32, zero before the first call. */
int clipgp_adamw_tail(float* p, const float* g, float* m, float* v, int64_t n, const float* lr_dev, float beta1, float beta2, float eps,
                      float weight_decay, int64_t* step, int64_t z_off, int64_t C, int64_t nrows, int64_t d, float* Z, const float* kl,
                      int64_t kl_n, float kl_scale, float* loss, int64_t* counter_b, int64_t by, unsigned int* ticket, void* stream);
/* End of an optimisation step in one launch: Z[:, n-1, :] <- z_last [C,d] (the learnable inducing row lives in the flat parameter
 * buffer; gp_template_weigher.py:72-79 freezes the other rows) and both device counters += by. */
int clipgp_step_epilogue(const float* z_last, float* Z, int64_t C, int64_t n, int64_t d, int64_t* counter_a, int64_t* counter_b,
                         int64_t by, void* stream);

/* C[M,N] (fp32, row stride ldc) = alpha * A B^T.  Deterministic (one accumulator per output tile). */
int clipgp_tc_gemm_store(const void* A_bf16, int64_t M, int64_t Ka, const void* B_bf16, int64_t N, int64_t K, float alpha,
                         float* C, int64_t ldc, void* stream);
/* Same product for the small / skinny GEMMs of the training step (few output tiles, long K): when the output grid cannot fill
 * the SMs the K blocks are split over work items that accumulate into C with vector reductions (red.global.add.v4.f32), so the
 * fp32 summation order, and the last bits of C, vary from run to run (as with any split-K GEMM). */
int clipgp_tc_gemm_store_splitk(const void* A_bf16, int64_t M, int64_t Ka, const void* B_bf16, int64_t N, int64_t K, float alpha,
                                float* C, int64_t ldc, void* stream);

/* TF32 tensor-core GEMM on fp32 operands READ IN PLACE (tcgen05.mma kind::tf32, fp32 accumulation in TMEM): the arithmetic of the
 * reference's own GPU path (trainers/adapter.py:23, `torch.backends.cuda.matmul.allow_tf32 = True`), with no operand cast kernels.
 * C[M,N] = alpha * op(A) op(B)^T.  a_layout / b_layout: 0 = the operand is [rows, K] row-major (K-major), 1 = it is [K, rows]
 * row-major (MN-major): the transpose of a row-major tensor is consumed as is -- e.g. d P_hat = dlogits^T f_hat
 * (trainers/adapter.py:426 adjoint) takes dlogits [B, S*C] with a_layout = 1 and f_hat [B, D] with b_layout = 1, K = B.
 * allow_split_k != 0: skinny outputs may split K over work items (fp32 atomic accumulation, not bit-reproducible). */
int clipgp_tc_gemm_tf32(const float* A, int a_layout, int64_t M, const float* B, int b_layout, int64_t N, int64_t K, float alpha,
                        float* C, int64_t ldc, int allow_split_k, void* stream);

/* clipgp_tc_logits_calibration / clipgp_tc_proj_logits_calibration on fp32 operands (TF32): A [M,Ka] features, B [N,K] =
 * [W ; prototypes] rows (norm_cols > 0: the first norm_cols rows of B are the visual projection, see
 * clipgp_tc_proj_logits_calibration) or the class prototypes only (norm_cols = 0). */
int clipgp_tc_logits_calibration_tf32(const float* A, int64_t M, int64_t Ka, const float* B, int64_t N, int64_t K, int64_t norm_cols,
                                      float alpha, const int64_t* labels, float* conf, int32_t* pred, uint8_t* correct,
                                      const float* boundaries, int n_bins, int64_t* bin_count, unsigned long long* bin_conf_fx,
                                      int64_t* bin_correct, int64_t* top1, float* logits_out, int64_t ld_logits, void* stream);

/* Logits alpha * A B^T reduced on the fly per row: max-softmax confidence, arg-max, hit flag, top-1 count and the equal-width
 * ECE histogram (utils/metrics.py:9-36,71-82) -- same outputs / accumulate semantics as clipgp_calibration_from_logits, but the
 * [M,N] logits never reach HBM unless logits_out != NULL. */
int clipgp_tc_logits_calibration(const void* A_bf16, int64_t M, int64_t Ka, const void* B_bf16, int64_t N, int64_t K, float alpha,
                                 const int64_t* labels, float* conf, int32_t* pred, uint8_t* correct, const float* boundaries,
                                 int n_bins, int64_t* bin_count, unsigned long long* bin_conf_fx, int64_t* bin_correct,
                                 int64_t* top1, float* logits_out, int64_t ld_logits, void* stream);

/* ================================================================================================
 * Tip-Adapter cache model — trainers/tip_adapter.py:43-80 (_build_cache, _search_hyperparams), :250-260 (train),
 * :281-290, :309-318, :331-333, :371-383 (eval):  affinity = f keys^T; cache = exp(-(beta - beta*affinity)) @ one_hot(labels_tr);
 * tip = clip_logits + alpha * cache.  The one-hot GEMM is a label-segmented sum here.
 * ================================================================================================ */

/* Exact-fp32 row pass over a materialised affinity block [B, N_tr] (row stride lda, from clipgp_gemm_f32):
 * out[b,c] = clip_logits[b,c] (0 if NULL) + alpha * sum_{j: labels_tr[j]==c} exp(-(beta - beta*aff[b,j])).
 * store_e != 0 overwrites the block with the exponentials (saved for the adjoint). */
int clipgp_tip_forward(float* affinity, int64_t lda, const int64_t* labels_tr, int64_t B, int64_t N_tr, int64_t C, float beta,
                       float alpha, const float* clip_logits, int64_t ldc, float* out, int64_t ldo, int store_e, void* stream);

/* In place over the saved exponentials: G[b,j] = dout[b, labels_tr[j]] * alpha * beta * e[b,j] = d loss / d affinity[b,j].
 * (The key gradient of Tip-Adapter-F, tip_adapter.py:229-269, is then G^T f via clipgp_gemm_f32.) */
int clipgp_tip_backward(float* e, int64_t lda, const int64_t* labels_tr, int64_t B, int64_t N_tr, const float* dout, int64_t ldd,
                        float beta, float alpha, void* stream);

/* The eval pass of the Adapter head in ONE GEMM (adapter.py:239-249): B = [W ; Q] with W [norm_cols = D, K] the visual projection and
 * Q = (mean_s p_hat_s) W [C, K] the prototypes pulled back through it, A = the RAW features.  The epilogue of a row block first sums
 * the squares of the D projection columns (|f W^T|^2), then treats the remaining N_total - norm_cols columns as class logits
 * alpha * (f . Q_c) / max(|f W^T|, 1e-12) with the calibration reduction of clipgp_tc_logits_calibration.  The projected and the
 * normalised features are never written.  norm_cols must be a multiple of 256 (the tile width); pred / labels index classes. */
int clipgp_tc_proj_logits_calibration(const void* A_bf16, int64_t M, int64_t Ka, const void* B_bf16, int64_t N_total, int64_t K,
                                      int64_t norm_cols, float alpha, const int64_t* labels, float* conf, int32_t* pred, uint8_t* correct,
                                      const float* boundaries, int n_bins, int64_t* bin_count, unsigned long long* bin_conf_fx,
                                      int64_t* bin_correct, int64_t* top1, void* stream);

/* Tensor-core fused form for evaluation: out[b, key_class[j]] += alpha * exp(-beta (1 - f_b . key_j)) for every key j, computed in
 * the epilogue of the tcgen05 affinity GEMM ([M,K] x [N_tr,K]^T, bf16 or split operands): the [M, N_tr] affinity never reaches
 * HBM.  `out` [M, C] must already hold clip_logits; key_class [N_tr] int32 (sorting the cache by class minimises atomics). */
int clipgp_tc_tip_logits(const void* F_bf16, int64_t M, const void* keys_bf16, int64_t N_tr, int64_t K, const int32_t* key_class,
                         float beta, float alpha, float* out, int64_t ldo, void* stream);

/* ================================================================================================
 * One-time GP setup — gp_template_weigher.py:103-107 (median-heuristic RBF length-scale):
 *     pdist = torch.cdist(flat, flat);  ls = pdist[pdist > 0].median()       flat = unit rows [N = C*T, d]
 * without the N x N matrix: exact radix select over the bit patterns of the distances.  One call = one histogram pass that
 * recomputes all pairwise distances sqrt(max(|a|^2 + |b|^2 - 2 a.b, 0)), i != j, tile by tile:
 *   hist[b] += #{ordered pairs with dist > 0, (bits >> prefix_shift) == prefix (if use_prefix), ((bits >> shift) & (nbins-1)) == b}
 *   positives[0] += #{ordered pairs with dist > 0}   (when positives != NULL)
 * hist (uint64 [nbins], nbins a power of two <= 4096) and positives are ACCUMULATED into: zero them first.
 * ================================================================================================ */
int clipgp_row_sqnorm(const float* X, int64_t N, int64_t d, float* sq, void* stream);
int clipgp_pairdist_radix_hist(const float* X, const float* sqnorm, int64_t N, int64_t d, int prefix_shift, uint32_t prefix,
                               int use_prefix, int shift, int nbins, unsigned long long* hist, unsigned long long* positives,
                               void* stream);

/* ================================================================================================
 * Multi-GPU optimiser step over NVLink peer memory (one process per GPU; csrc/peer.cu) — replaces, for the data-parallel
 * GP-Adapter step, ncclAllReduce(flat gradient) + optimizer.step() (adapter.py:537-549; AdamW as utils/optimization.py:57-216):
 * gradient reduce-scatter + AdamW on the owned 1/world slice + parameter all-gather in ONE kernel, through peer pointers.
 *
 * Set-up: every rank allocates one block with clipgp_peer_alloc (cudaMalloc, zeroed), exports it (64-byte CUDA IPC handle, carried
 * to the peers by any host channel, e.g. torch.distributed.all_gather_object) and opens the peers' blocks; g / p / flags below point
 * into those blocks (index = rank; the entry of the calling rank is its own block).
 *   g[q]     : rank q's flat gradient buffer, n + 1 floats (slot n: its loss share), 16-byte aligned
 *   p[q]     : rank q's flat parameter buffer, n floats, 16-byte aligned (every rank ends the call with identical parameters)
 *   flags[q] : rank q's 2 * CLIPGP_PEER_MAX uint64 flags, zero before the first call
 *   m, v     : LOCAL Adam moments, n floats (only the owned slice is read / written)
 *   n_group0 : elements [0, n_group0) use lr_dev[0], the rest lr_dev[1] (visual-projection / gp_weighter groups, adapter.py:298-309)
 *   step     : device int64, the 1-based Adam step (read, not advanced);  local: device uint64[2], zero before the first call
 *   loss_out : optional device float: sum over ranks of slot n
 *   kl       : optional device vector whose scaled sum is added to this rank's loss slot first (the KL term of the step)
 *   status   : device int, set to 1 / 2 if a peer's "gradients ready" / "parameters written" flag did not arrive within timeout_ns
 * Every rank must make the same sequence of calls.  The call is stream-ordered and can be captured in a CUDA graph.
 * ================================================================================================ */
#define CLIPGP_PEER_MAX 8
typedef struct clipgp_peer_args {
    int32_t world, rank;
    const float* g[CLIPGP_PEER_MAX];
    float* p[CLIPGP_PEER_MAX];
    unsigned long long* flags[CLIPGP_PEER_MAX];
    float* m;
    float* v;
    int64_t n, n_group0;
    const float* lr_dev;
    float beta1, beta2, eps, weight_decay;
    const int64_t* step;
    unsigned long long* local;
    float* loss_out;
    int32_t* status;
    unsigned long long timeout_ns;
    const float* kl;                 /* optional: g[rank][n] += kl_scale * sum(kl[0 .. kl_n)) before the exchange (this rank's KL share) */
    int64_t kl_n;
    float kl_scale;
} clipgp_peer_args;
int clipgp_peer_alloc(int64_t bytes, void** out);
int clipgp_peer_free(void* block);
int clipgp_ipc_export(const void* block, unsigned char* handle64);
int clipgp_ipc_open(const unsigned char* handle64, void** out);
int clipgp_ipc_close(void* mapped);
int clipgp_peer_adamw(const clipgp_peer_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CLIPGP_H_ */
