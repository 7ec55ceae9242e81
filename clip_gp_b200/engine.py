"""Fused GP-Adapter engine: the reference's training step (trainers/adapter.py:328-549, compute_loss
:387-476) and MC-averaged evaluation (:230-252 + utils/metrics.py) as a fixed sequence of clipgp kernel
launches on device-resident buffers — no autograd graph, no per-step allocation, CUDA-graph capturable.

One step = GP forward (kernel build, chol64, predictive, chol32, Philox MC sampling, sparsemax, KL) ->
weighted unit prototypes -> visual projection + normalisation -> logits for every MC sample ->
per-sample cross-entropy -> adjoints of all of the above -> ||W-I||^2 regulariser -> AdamW on both
parameter groups (adapter.py:290-311).  The three logging-only GP passes and the per-step full test
evaluation of the reference are not part of the step.

Multi-GPU: MC samples are sharded across ranks for training (every rank draws its slice of
the same Philox stream), gradients are summed with ONE all-reduce of the flat gradient buffer; evaluation
shards the images and all-reduces only the integer counters.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, Optional

import torch

from . import _lib, dist, metrics
from ._lib import GpArgs, GpBwdArgs, KERNEL_IDS
from .gp_template_weigher import GaussianProcessTemplateWeighter


@dataclass
class EngineConfig:
    S_train: int = 10
    S_eval: int = 10
    batch_size: int = 128
    logit_scale: float = 100.0          # logit_scale.exp() of CLIP (adapter.py:241)
    gp_beta: float = 0.01               # configs/trainers/gp.yaml GP_BETA
    l2_lambda: float = 0.5              # default.yaml L2_LAMBDA
    shots: int = 16
    lr: float = 0.01                    # OPTIM.LR (visual_proj group)
    gp_lr: float = 1e-3                 # GP_LR (gp_weighter group)
    weight_decay: float = 0.0
    betas: tuple = (0.9, 0.999)
    adam_eps: float = 1e-8
    loss_mode: str = "per_sample"       # "per_sample" (adapter.py:422-428) | "logit_mean" (taskres.py:268-270)
    train_visual_proj: bool = True      # FREEZE_VISUAL_PROJ False
    precision: str = "fp32"             # GEMMs of the step: "fp32" (FFMA, exact comparator) | "tf32" (tcgen05 kind::tf32 on the fp32
                                        # tensors in place, no operand casts, transposed operands read in place: the reference's own GPU
                                        # arithmetic, adapter.py:23) | "bf16x3" (tcgen05, split bf16 operands, fp32-grade products) | "bf16"
    graph_collectives: bool = False     # world > 1: capture the step (its NCCL all-reduces included) in the CUDA graph as well
                                        # (1886 -> 2088 steps/s at 2 GPUs).  The owner must drop the graph (engine._graph = None)
                                        # before destroying the process group, otherwise NCCL teardown hangs; bench.py does
    shard_classes: bool = False         # world > 1: additionally run the per-class GP kernels on this rank's class shard only (w and dw
                                        # cross NVLink as two 1.3 MB all-reduces).  Off by default: it rules out the per-class fusion of
                                        # the prototype stages into the GP kernels, which is worth more (2 GPUs: 1876 vs ~2200 steps/s)
    fuse_prototypes: bool = True        # build the prototypes inside the GP forward kernel's CTA when the sizes allow it
    template_logit_adjoint: bool = False  # small batches: a[s,t] of the prototype adjoint from Zt = f_hat E^T (one tensor-core GEMM) and
                                        # dlogits^T instead of the d P_hat GEMM + a 64 KB stream over E[c] per class (clipgp.h, tl_*).
                                        # Measured at cfg2: GP adjoint 220 -> 194 us and no d P_hat GEMM, but the Zt GEMM (36 us, 125 CTAs
                                        # of 212 KB shared memory, 98 MB of split-bf16 text bank) cannot hide next to the one-wave GP
                                        # kernels: step 0.428 -> 0.442 ms.  Off by default; wins when the GEMM has a free slot.
    fuse_eval_projection: bool = True   # tensor-core eval: projection + normalisation + logits + calibration in ONE GEMM (B = [W ; P W])
    overlap: bool = True                # run the feature branch of the step on a side stream next to the GP branch
    clear_on_side: bool = True          # clear the gradient buffer at the head of the side stream instead of in front of the GP forward
    fuse_tail: bool = True              # single GPU: KL sum + AdamW of the gp_weighter group + inducing-row scatter + counters in ONE launch
    seed: int = 0
    rank: int = 0
    world: int = 1
    shard: str = "samples"              # world > 1, what the ranks split in a training step:
                                        #   "samples": the S MC samples of ONE batch (north star; strong scaling of a latency-bound step)
                                        #   "batch":   data parallel -- every rank takes its own [B, D] batch and all S samples, the global
                                        #              batch is world * B (weak scaling when B is fixed per rank, strong when the caller
                                        #              splits a fixed batch); loss / gradients are means over the global batch
    peer_update: bool = False           # world > 1, shard = "batch": replace ncclAllReduce(flat gradient) + AdamW by ONE kernel over NVLink peer
                                        # memory (csrc/peer.cu: gradient reduce-scatter + AdamW on the owned 1/world slice + parameter
                                        # all-gather through CUDA-IPC peer pointers).  All ranks must live on one node
    shard_eval_classes: bool = True     # world > 1: the eval GP forward runs on this rank's C / world classes only and the [C, D] mean
                                        # prototypes are completed by one 2 MB all-reduce (instead of replicating the per-class chain)


class GPAdapterEngine:
    def __init__(self, gpw: GaussianProcessTemplateWeighter, cfg: EngineConfig, visual_proj_weight: Optional[torch.Tensor] = None):
        self.lib = _lib.load()
        self.cfg = cfg
        dev = gpw._templates.device
        if dev.type != "cuda":
            raise RuntimeError("GPAdapterEngine needs the weighter on a CUDA device (no CPU fallback)")
        self.dev = dev
        self.kernel_type = gpw.kernel_type
        gpw.variational_strategy._maybe_init()
        self.C, self.T, self.D = gpw.num_classes, gpw.num_templates, gpw.dim
        self.n, self.d = self.T + 1, gpw.red_dim
        Cn, T, D, n, d = self.C, self.T, self.D, self.n, self.d
        # MC samples of this rank: an even split, the first S % world ranks take one more (S=10 over 8 ranks
        # is 2,2,1,1,1,1,1,1); every rank draws its slice [s_offset, s_offset + S_local) of the same Philox stream
        if cfg.shard not in ("samples", "batch"):
            raise ValueError(f"unknown shard mode {cfg.shard!r}")
        self.batch_sharded = bool(cfg.shard == "batch" and cfg.world > 1)
        if self.batch_sharded:
            self.s_offset, self.S_local = 0, cfg.S_train
        else:
            if cfg.S_train < cfg.world:
                raise ValueError(f"S_train={cfg.S_train} < world={cfg.world}: use EngineConfig(shard='batch')")
            self.s_offset, self.S_local = dist.sample_split(cfg.S_train, cfg.rank, cfg.world)
        self.class_sharded = bool(cfg.shard_classes and cfg.world > 1)
        self.c_lo, self.c_hi = dist.shard_range(self.C, cfg.rank, cfg.world) if self.class_sharded else (0, self.C)
        self.E = gpw._templates.detach().contiguous()
        self.X = gpw._templates_red.detach().contiguous()
        self.Z = gpw.variational_strategy.inducing_points.detach().clone().contiguous()
        q = gpw.variational_strategy._variational_distribution
        raw_ls, raw_os, raw_var = gpw._kernel_raw()
        # ---- flat parameter / gradient / Adam buffers: [W | z_last | m | Lq | hyper...]
        segs = [("W", D * D), ("z_last", Cn * d), ("m", Cn * n), ("Lq", Cn * n * n)]
        if raw_ls is not None: segs.append(("ls", Cn * d))
        if raw_os is not None: segs.append(("os", Cn))
        if raw_var is not None: segs.append(("var", Cn))
        self.offsets: Dict[str, tuple] = {}
        off = 0
        for name, sz in segs:
            self.offsets[name] = (off, sz)
            off += sz
        self.n_params = off
        self.peer = None
        if cfg.peer_update and cfg.world > 1:
            if not self.batch_sharded:
                raise ValueError("EngineConfig.peer_update needs shard='batch'")
            a16 = lambda x: (x + 255) & ~255
            lay, o = {}, 0
            for nm, nb in (("flags", 8 * 2 * _lib.PEER_MAX), ("g", 4 * (off + 1)), ("p", 4 * off)):
                lay[nm] = (o, nb); o = a16(o + nb)
            try:
                self.peer = dist.PeerBlock(lay, dev, cfg.rank, cfg.world)
            except RuntimeError as e:          # raised on every rank together: all of them take the NCCL path
                if cfg.rank == 0:
                    import sys
                    print(f"[clipgp] peer_update unavailable, using ncclAllReduce + AdamW: {e}", file=sys.stderr)
                self.peer = None
        if self.peer is not None:
            self.flat_p = self.peer.local("p", torch.float32)
            self.flat_g = self.peer.local("g", torch.float32)
            self.peer_local = torch.zeros(2, dtype=torch.int64, device=dev)         # epoch, CTA arrival counter
            self.peer_status = torch.zeros(1, dtype=torch.int32, device=dev)
            self.loss_global = torch.zeros(1, dtype=torch.float32, device=dev)
        else:
            self.flat_p = torch.zeros(off, dtype=torch.float32, device=dev)
            self.flat_g = torch.zeros(off + 1, dtype=torch.float32, device=dev)     # last slot: the loss (all-reduced with the grads)
        self.flat_m = torch.zeros(off, dtype=torch.float32, device=dev)
        self.flat_v = torch.zeros(off, dtype=torch.float32, device=dev)
        W0 = torch.eye(D, device=dev) if visual_proj_weight is None else visual_proj_weight.detach().to(dev).float()
        self.p("W").copy_(W0.reshape(-1))
        self.p("z_last").copy_(self.Z[:, n - 1, :].reshape(-1))
        self.p("m").copy_(q.variational_mean.detach().reshape(-1))
        self.p("Lq").copy_(q.chol_variational_covar.detach().reshape(-1))
        if raw_ls is not None: self.p("ls").copy_(raw_ls.detach().reshape(-1))
        if raw_os is not None: self.p("os").copy_(raw_os.detach().reshape(-1))
        if raw_var is not None: self.p("var").copy_(raw_var.detach().reshape(-1))
        self.mean_x = gpw.mean_module.test_mean(T).detach().contiguous()      # per-class constant: no effect on w
        # ---- device scalars
        self.rng_state = torch.tensor([int(cfg.seed), 0], dtype=torch.int64, device=dev)
        # evaluation draws from its OWN counter stream (the reference draws fresh, independent noise for every eval call,
        # adapter.py:242-249): a different key, advanced once per eval pass, never touched by the captured train graph
        self.eval_rng_state = torch.tensor([(int(cfg.seed) ^ 0x5DEECE66D) & 0x7FFFFFFFFFFFFFFF, 0], dtype=torch.int64, device=dev)
        self.eval_eps = None            # optional explicit base noise [C, >=T, S] for evaluation (parity tests)
        self.adam_step = torch.ones(1, dtype=torch.int64, device=dev)
        self.lr_dev = torch.tensor([cfg.lr, cfg.gp_lr], dtype=torch.float32, device=dev)   # [visual_proj group, gp_weighter group]
        self._side_stream = torch.cuda.Stream(dev)
        self._alloc_train(cfg.batch_size)
        self._graph = None

    # ------------------------------------------------------------------ views
    def p(self, name):
        o, s = self.offsets[name]
        return self.flat_p[o:o + s]

    def g(self, name):
        o, s = self.offsets[name]
        return self.flat_g[o:o + s]

    @property
    def loss(self) -> torch.Tensor:
        """Loss of the last step (device scalar; the global mean when sharded)."""
        if self.peer is not None:
            return self.loss_global
        return self.flat_g[self.n_params:self.n_params + 1]

    @property
    def _loss_acc(self) -> torch.Tensor:
        """Where this rank's kernels accumulate their loss share (slot n of the gradient buffer)."""
        return self.flat_g[self.n_params:self.n_params + 1]

    def _ptr(self, buf, name):
        o, _ = self.offsets[name]
        return buf.data_ptr() + 4 * o

    # ------------------------------------------------------------------ buffers
    def _alloc_train(self, B):
        dev, Cn, T, D, n, d, S = self.dev, self.C, self.T, self.D, self.n, self.d, self.S_local
        f32 = dict(dtype=torch.float32, device=dev)
        self.B = B
        self.in_feat = torch.zeros(B, D, **f32)
        self.in_lab = torch.zeros(B, dtype=torch.int64, device=dev)
        self.Y = torch.empty(B, D, **f32)
        self.f_hat = torch.empty(B, D, **f32)
        self.f_inv = torch.empty(B, **f32)
        # class-sharded GP: w / dw hold ALL S_train samples (this rank fills its classes / its samples, an all-reduce completes them)
        Sg = self.cfg.S_train if self.class_sharded else S
        self.w_all = torch.zeros(Sg, Cn, T, **f32)
        self.dw_all = torch.zeros(Sg, Cn, T, **f32)
        so = self.s_offset if self.class_sharded else 0
        self.w = self.w_all[so:so + S]
        self.kl = torch.empty(Cn, **f32)
        self.Lsave = torch.empty(Cn, n, n, dtype=torch.float64, device=dev)
        self.Asave = torch.empty(Cn, n, T, **f32)
        self.Rsave = torch.empty(Cn, T, T, **f32)
        self.status = torch.zeros(Cn, dtype=torch.int32, device=dev)
        self.Ksave = torch.empty(Cn, 1 + n * n + n * T + T * T, **f32)
        self.P_hat = torch.empty(S, Cn, D, **f32)
        self.P_norm = torch.empty(S, Cn, **f32)
        self.P_mean = torch.empty(Cn, D, **f32)
        self.logits = torch.empty(B, (S if self.cfg.loss_mode == "per_sample" else 1) * Cn, **f32)
        self.dP = torch.empty((S if self.cfg.loss_mode == "per_sample" else 1), Cn, D, **f32)
        self.df_hat = torch.empty(B, D, **f32)
        self.dY = torch.empty(B, D, **f32)
        self.dw = self.dw_all[so:so + S]
        if self.cfg.precision == "tf32":
            SCt = (S if self.cfg.loss_mode == "per_sample" else 1) * Cn
            if D % 4 or SCt % 4:
                raise ValueError(f"precision='tf32' needs D ({D}) and S*C ({SCt}) to be multiples of 4 (16-byte TMA row pitch); use 'bf16x3'")
            # large batches: the adjoint GEMMs take their SMALL operand (P_hat, f_hat) K-major from a transposed copy -- as the 256-wide
            # B tile an MN-major operand costs the MMA pipeline ~35 % (d f_hat 504 -> 331 us, d P_hat 520 -> 307 us at B = 16 000), the
            # copies ~10 us each; the big operand (dlogits, 640 MB) is still read in place.  Minibatches are latency bound: in place
            self.tf32_kmajor_b = bool(B >= 1024 and B % 4 == 0)
            if self.tf32_kmajor_b:
                self.PT32 = torch.empty(D, SCt, **f32)
                self.fhT32 = torch.empty(D, B, **f32)
        elif self.cfg.precision != "fp32":
            self._alloc_tc()
        a = GpArgs()
        a.kernel_type = KERNEL_IDS[self.kernel_type]
        a.x_is_z_prefix = 2     # the engine never writes the frozen template rows of Z (only z_last is scattered back)
        a.C, a.T, a.n, a.d, a.S = Cn, T, n, d, Sg
        a.c_begin, a.c_count = self.c_lo, self.c_hi - self.c_lo
        self.eps_save = torch.empty(Sg, Cn, T, **f32)       # base noise of the step: drawn once (forward), re-read by the adjoint
        a.eps_save = self.eps_save.data_ptr()
        # fused prototype stage: the GP forward kernel's CTA of class c also builds the unit prototypes of class c and the bf16
        # operand rows of the logit GEMM (no grid-wide barrier, no separate prototype / cast launches)
        self.fused_proto = bool(self.cfg.fuse_prototypes and self.cfg.loss_mode == "per_sample" and not self.class_sharded and
                                self.lib.clipgp_gp_fused_proto_ok(T, n, d, D, S))
        if self.fused_proto:
            a.proto_E, a.proto_D = self.E.data_ptr(), D
            a.proto_P_hat, a.proto_norm = self.P_hat.data_ptr(), self.P_norm.data_ptr()
            if self.cfg.precision in ("bf16x3", "bf16"):
                a.proto_bf16, a.proto_bf16_ld, a.proto_bf16_seg, a.proto_bf16_mode = self.Pb.data_ptr(), self.Pb.stride(0), D, self.tc_mb
        a.Z, a.X = self.Z.data_ptr(), self.X.data_ptr()
        a.raw_lengthscale = self._ptr(self.flat_p, "ls") if "ls" in self.offsets else None
        a.raw_outputscale = self._ptr(self.flat_p, "os") if "os" in self.offsets else None
        a.raw_variance = self._ptr(self.flat_p, "var") if "var" in self.offsets else None
        a.var_mean, a.chol_var = self._ptr(self.flat_p, "m"), self._ptr(self.flat_p, "Lq")
        a.mean_x = self.mean_x.data_ptr()
        a.eps = None
        a.rng_state = self.rng_state.data_ptr()
        a.s_offset, a.S_total = (0 if self.class_sharded else self.s_offset), self.cfg.S_train
        a.w, a.kl, a.L, a.A, a.R, a.status = (self.w_all.data_ptr(), self.kl.data_ptr(), self.Lsave.data_ptr(),
                                              self.Asave.data_ptr(), self.Rsave.data_ptr(), self.status.data_ptr())
        a.Ksave = self.Ksave.data_ptr()
        self.gp_args = a
        b = GpBwdArgs()
        b.dw = self.dw_all.data_ptr()
        b.dkl = None
        # KL term: every rank holds all classes (pre-divided by the world size) or only its class shard (full weight)
        self.kl_weight = float(self.cfg.gp_beta) / (1 if self.class_sharded else self.cfg.world)
        b.dkl_scalar = self.kl_weight
        b.dZ_last = self._ptr(self.flat_g, "z_last")
        b.draw_lengthscale = self._ptr(self.flat_g, "ls") if "ls" in self.offsets else None
        b.draw_outputscale = self._ptr(self.flat_g, "os") if "os" in self.offsets else None
        b.draw_variance = self._ptr(self.flat_g, "var") if "var" in self.offsets else None
        b.dvar_mean, b.dchol_var, b.dmean_x = self._ptr(self.flat_g, "m"), self._ptr(self.flat_g, "Lq"), None
        # fused prototype adjoint: the GP adjoint kernel's CTA of class c derives dw[:, c, :] from dP_hat itself
        self.fused_proto_bwd = bool(self.cfg.fuse_prototypes and self.cfg.loss_mode == "per_sample" and not self.class_sharded and
                                    self.lib.clipgp_gp_fused_proto_bwd_ok(T, n, d, D, S))
        if self.fused_proto_bwd:
            if getattr(self, "EEt", None) is None:
                self.EEt = torch.bmm(self.E, self.E.transpose(1, 2)).contiguous()       # [C,T,T], frozen text bank: one-time setup
            b.proto_dP, b.proto_dP_stride_s, b.proto_dP_scale = self.dP.data_ptr(), Cn * D, 1.0
            b.proto_norm, b.proto_E, b.proto_EEt, b.proto_D = self.P_norm.data_ptr(), self.E.data_ptr(), self.EEt.data_ptr(), D
            b.dw_out = self.dw_all.data_ptr()
        # small-batch form of the same adjoint (needs the bf16 dlogits^T operand of the tensor-core step)
        self.tl_adjoint = bool(self.fused_proto_bwd and self.cfg.template_logit_adjoint and self.cfg.precision in ("bf16x3", "bf16") and B <= 256 and
                               S <= 12 and D % 8 == 0)
        if self.tl_adjoint:
            seg = self.tc_seg
            if getattr(self, "Eb", None) is None:           # frozen text bank as the K-major B operand [C*T, seg*D]: one-time setup
                self.Eb = torch.empty(Cn * T, seg * D, dtype=torch.bfloat16, device=self.dev)
                self._cast(self.E.data_ptr(), Cn * T, D, D, self.Eb, D, self.tc_mb)
            self.Zt = torch.empty(B, Cn * T, dtype=torch.float32, device=self.dev)
            b.tl_Z, b.tl_Z_ld = self.Zt.data_ptr(), Cn * T
            b.tl_dlT, b.tl_dlT_ld, b.tl_seg = self.dlTb.data_ptr(), self.dlTb.stride(0), self.Bp
            b.tl_B, b.tl_mode, b.tl_scale = B, (1 if seg == 3 else 0), float(self._dims()[3])
        self.gp_bwd_args = b

    # ------------------------------------------------------------------ tensor-core operand buffers
    def _alloc_tc(self):
        """bf16 operands of the five GEMMs of the step, all K-major ([rows, K]); in split mode every operand is K-tripled
        ([hi|hi|lo] on the A side, [hi|lo|hi] on the B side).  K is padded to a multiple of 8 (16-byte TMA row pitch) with
        zero columns that no cast ever writes."""
        if self.cfg.precision not in ("bf16x3", "bf16"):
            raise ValueError(f"unknown precision {self.cfg.precision!r}")
        if self.D % 8:
            raise ValueError("tensor-core step needs D % 8 == 0")
        B, Cn, D, S = self.B, self.C, self.D, self.S_local
        SC = (S if self.cfg.loss_mode == "per_sample" else 1) * Cn
        seg = 3 if self.cfg.precision == "bf16x3" else 1
        self.tc_seg, self.tc_ma, self.tc_mb = seg, (1 if seg == 3 else 0), (2 if seg == 3 else 0)
        self.Bp, self.SCp = (B + 7) // 8 * 8, (SC + 7) // 8 * 8
        z = lambda r, k: torch.zeros(r, seg * k, dtype=torch.bfloat16, device=self.dev)
        self.fb, self.Wb, self.fhb, self.Pb = z(B, D), z(D, D), z(B, D), z(SC, D)
        self.dlb, self.dlTb = z(B, self.SCp), z(SC, self.Bp)
        self.fhTb, self.PTb = z(D, self.Bp), z(D, self.SCp)
        self.dYTb, self.fTb = z(D, self.Bp), z(D, self.Bp)
        self.sm_stats = torch.empty(B * (S if self.cfg.loss_mode == "per_sample" else 1), 2, dtype=torch.float32, device=self.dev)

    def _cast(self, src_ptr, R, K, ldx, out, Kp, mode, transpose=False):
        fn = self.lib.clipgp_cast_bf16_transpose if transpose else self.lib.clipgp_cast_bf16
        _lib.check(fn(src_ptr, R, K, ldx, out.data_ptr(), out.stride(0), Kp, mode, _lib.stream_ptr(self.dev)), "cast_bf16")

    def _cast2(self, src_ptr, R, K, ldx, out, Kp, mode, outT, Rp, modeT):
        """One read of an fp32 [R,K] source -> row-major operand `out` [R, seg*Kp] and transposed operand `outT` [K, seg*Rp]."""
        _lib.check(self.lib.clipgp_cast_bf16_dual(src_ptr, R, K, ldx, _lib.ptr(out), out.stride(0) if out is not None else 0, Kp, mode,
                                                  _lib.ptr(outT), outT.stride(0) if outT is not None else 0, Rp, modeT,
                                                  _lib.stream_ptr(self.dev)), "cast_bf16_dual")

    def _tf32(self, A_ptr, a_t, M, B_ptr, b_t, N, K, alpha, out_ptr, ldc, split_k=True):
        """C[M,N] = alpha op(A) op(B)^T in TF32 on fp32 tensors in place (a_t / b_t: the operand is stored [K, rows])."""
        _lib.check(self.lib.clipgp_tc_gemm_tf32(A_ptr, int(a_t), M, B_ptr, int(b_t), N, K, float(alpha), out_ptr, ldc, int(split_k),
                                                _lib.stream_ptr(self.dev)), "tc_gemm_tf32")

    def _tc(self, A, Bm, alpha, out_ptr, ldc):
        _lib.check(self.lib.clipgp_tc_gemm_store_splitk(A.data_ptr(), A.shape[0], A.shape[1], Bm.data_ptr(), Bm.shape[0], Bm.shape[1],
                                                        float(alpha), out_ptr, ldc, _lib.stream_ptr(self.dev)), "tc_gemm_store_splitk")

    # ------------------------------------------------------------------ one training step (launch only)
    # The step is a small dependency graph with two independent branches on either side of the logit GEMM:
    #   features:   cast -> projection GEMM -> row normalise -> casts            | d f_hat GEMM -> normalise adjoint -> dW GEMM -> L2
    #   prototypes: GP forward (Gram + algebra) -> prototypes -> casts           | d P_hat GEMM -> prototype adjoint -> GP adjoint -> KL
    # The feature branch (few, latency-bound launches) runs on a side stream next to the GP branch (one-wave kernels that leave
    # issue slots free); inside the captured CUDA graph this becomes two parallel chains.
    def _launch_step(self):
        cfg = self.cfg
        main = torch.cuda.current_stream(self.dev)
        side = self._side_stream if cfg.overlap else None
        if side is not None:
            if not cfg.clear_on_side:
                self.flat_g.zero_()
            side.wait_stream(main)
            with torch.cuda.stream(side):
                if cfg.clear_on_side:
                    self.flat_g.zero_()       # off the critical path: the GP forward does not touch the gradient buffer, and every writer of
                self._fwd_features()          # it (loss terms, adjoint kernels, split-K GEMMs) runs after main has joined this stream again
        else:
            self.flat_g.zero_()
            self._fwd_features()
        self._fwd_prototypes()
        if side is not None:
            main.wait_stream(side)
        self._logits()
        tl_done = None
        if self.tl_adjoint and cfg.precision != "fp32":
            if side is not None:
                side.wait_stream(main)                       # after the logit GEMM, next to the softmax kernels
                with torch.cuda.stream(side):
                    self._template_logits()
                    tl_done = torch.cuda.Event()
                    tl_done.record(side)
            else:
                self._template_logits()
        self._loss()
        if side is not None:
            side.wait_stream(main)
            with torch.cuda.stream(side):
                self._bwd_features()
        else:
            self._bwd_features()
        if tl_done is not None:
            main.wait_event(tl_done)
        self._bwd_prototypes()
        if side is not None:
            main.wait_stream(side)
        if self.peer is not None and not getattr(self, "skip_update", False):
            self._launch_peer_update()                        # reduce-scatter + AdamW + all-gather over NVLink peer memory, one kernel
            return
        if cfg.world > 1:
            torch.distributed.all_reduce(self.flat_g)         # ONE fused all-reduce: gradients + loss
            if self.peer is not None:
                self.loss_global.copy_(self._loss_acc)
        if not getattr(self, "skip_update", False):
            self._launch_update()

    def _dims(self):
        per_sample = self.cfg.loss_mode == "per_sample"
        S = self.S_local
        SC = (S if per_sample else 1) * self.C
        alpha = self.cfg.logit_scale * (1.0 if per_sample else 1.0 / S)
        return per_sample, S, SC, alpha

    def _fwd_features(self):
        """visual projection + normalisation (adapter.py:419-420)"""
        lib, cfg, ck, st = self.lib, self.cfg, _lib.check, _lib.stream_ptr(self.dev)
        B, D = self.B, self.D
        W = self._ptr(self.flat_p, "W")
        if cfg.precision == "tf32":
            self._tf32(self.in_feat.data_ptr(), False, B, W, False, D, D, 1.0, self.Y.data_ptr(), D)         # Y = f W^T, both K-major
        elif cfg.precision != "fp32":
            ma, mb = self.tc_ma, self.tc_mb
            self._cast2(self.in_feat.data_ptr(), B, D, D, self.fb, D, ma, self.fTb if cfg.train_visual_proj else None, self.Bp, mb)
            self._cast(W, D, D, D, self.Wb, D, mb)
            self._tc(self.fb, self.Wb, 1.0, self.Y.data_ptr(), D)
        else:
            ck(lib.clipgp_gemm_f32(self.in_feat.data_ptr(), D, 1, W, 1, D, self.Y.data_ptr(), D, B, D, D, 1.0, 0, st), "gemm(proj)")
        ck(lib.clipgp_rownorm_forward(self.Y.data_ptr(), B, D, self.f_hat.data_ptr(), self.f_inv.data_ptr(), None, st), "rownorm")
        if cfg.precision in ("bf16x3", "bf16"):
            self._cast2(self.f_hat.data_ptr(), B, D, D, self.fhb, D, self.tc_ma, None if self.tl_adjoint else self.fhTb, self.Bp, self.tc_mb)

    def _template_logits(self):
        """Zt = f_hat E_flat^T [B, C*T]: the per-template cosines the GP adjoint contracts with dlogits^T (clipgp.h, tl_*).  Independent
        of the GP; launched on the feature stream AFTER the logit GEMM so that its 125 tiles (212 KB of shared memory each) neither
        push the GP forward CTAs off the SMs nor delay the logit GEMM, and run next to the softmax kernels instead."""
        _lib.check(self.lib.clipgp_tc_gemm_store(self.fhb.data_ptr(), self.B, self.fhb.shape[1], self.Eb.data_ptr(), self.Eb.shape[0],
                                                 self.Eb.shape[1], 1.0, self.Zt.data_ptr(), self.Zt.stride(0), _lib.stream_ptr(self.dev)),
                   "tc_gemm_store(Zt)")

    def _fwd_prototypes(self):
        """GP weights + unit prototypes (adapter.py:404, 424-425)"""
        lib, cfg, ck, st = self.lib, self.cfg, _lib.check, _lib.stream_ptr(self.dev)
        Cn, T, D = self.C, self.T, self.D
        per_sample, S, SC, _ = self._dims()
        if self.class_sharded:
            self.w_all.zero_()
        ck(lib.clipgp_gp_forward(C.byref(self.gp_args), st), "gp_forward")
        if self.class_sharded:
            torch.distributed.all_reduce(self.w_all)          # every rank wrote its classes (all samples); the sum completes w
        if self.fused_proto:
            return                                            # prototypes and their GEMM operand came out of the GP kernel
        ck(lib.clipgp_proto_forward(self.w.data_ptr(), self.E.data_ptr(), S, Cn, T, D, None, 0.0, None, self.P_hat.data_ptr(),
                                    self.P_norm.data_ptr(), None, None if per_sample else self.P_mean.data_ptr(), None, 0, st), "proto_forward")
        if cfg.precision in ("bf16x3", "bf16"):
            Bmat = self.P_hat if per_sample else self.P_mean
            # only the row-major operand is on the critical path (logit GEMM); the transposed copy feeds d f_hat on the feature branch
            self._cast(Bmat.data_ptr(), SC, D, D, self.Pb, D, self.tc_mb)

    def _logits(self):
        """logits (adapter.py:426)"""
        lib, cfg, ck, st = self.lib, self.cfg, _lib.check, _lib.stream_ptr(self.dev)
        B, Cn, D = self.B, self.C, self.D
        per_sample, S, SC, alpha = self._dims()
        Bmat = self.P_hat if per_sample else self.P_mean
        if cfg.precision == "tf32":
            # no split-K here: at B = 128 (40 tiles) a 3-way split with its memset and atomic epilogue measured exactly as fast as 40 whole-K
            # tiles, and without it the logits are bit-reproducible
            self._tf32(self.f_hat.data_ptr(), False, B, Bmat.data_ptr(), False, SC, D, alpha, self.logits.data_ptr(), SC, split_k=False)
        elif cfg.precision != "fp32":
            self._tc(self.fhb, self.Pb, alpha, self.logits.data_ptr(), SC)
        else:
            ck(lib.clipgp_gemm_f32(self.f_hat.data_ptr(), D, 1, Bmat.data_ptr(), 1, D, self.logits.data_ptr(), SC, B, SC, D,
                                   alpha, 0, st), "gemm(logits)")

    def _loss(self):
        """cross-entropy + gradient of the logits (adapter.py:427-428)"""
        lib, cfg, ck, st = self.lib, self.cfg, _lib.check, _lib.stream_ptr(self.dev)
        B, Cn, D = self.B, self.C, self.D
        per_sample, S, SC, alpha = self._dims()
        tcm = cfg.precision in ("bf16x3", "bf16")          # tf32: the fp32 softmax kernel writes dlogits in place, the GEMMs read it as is
        # mean over the (global) batch and the S_train samples; every rank contributes its share and the all-reduce sums them
        dp = cfg.world if self.batch_sharded else 1
        if per_sample:
            rows, rpl = B * S, S
            loss_scale = 1.0 / (B * cfg.S_train * dp)
        else:
            rows, rpl = B, 1
            loss_scale = 1.0 / (B * cfg.world)
            if cfg.world > 1 and not self.batch_sharded:
                raise NotImplementedError("logit_mean loss is not S-sharded (it is not a sum over samples); use shard='batch'")
        if tcm:
            # two-phase softmax cross-entropy: row statistics + loss, then dlogits straight into the bf16 operands of the two
            # adjoint GEMMs (dlogits [B, SC] and dlogits^T [SC, B]); no fp32 dlogits
            dlb_ptr = _lib.ptr(self.dlb) if cfg.train_visual_proj else None
            if ((B + 63) // 64) * rpl >= 296:
                # large batches: one launch, a CTA owns 64 batch rows of one sample (>= two CTAs per SM; second read from L2)
                ck(lib.clipgp_softmax_ce_bf16_dual(self.logits.data_ptr(), self.in_lab.data_ptr(), B, rpl, Cn, self._loss_acc.data_ptr(), loss_scale,
                                                   loss_scale, dlb_ptr, self.dlb.stride(0), self.SCp, self.tc_ma, self.dlTb.data_ptr(),
                                                   self.dlTb.stride(0), self.Bp, self.tc_ma, st), "softmax_ce_bf16_dual")
            else:
                # minibatches: two launches with full-GPU grids (row statistics, then 64 x 64 tiles over the whole [B, S*C] matrix)
                ck(lib.clipgp_softmax_ce_stats(self.logits.data_ptr(), Cn, self.in_lab.data_ptr(), rows, rpl, Cn, self.sm_stats.data_ptr(),
                                               self._loss_acc.data_ptr(), loss_scale, st), "softmax_ce_stats")
                ck(lib.clipgp_softmax_grad_bf16_dual(self.logits.data_ptr(), self.sm_stats.data_ptr(), self.in_lab.data_ptr(), B, rpl, Cn,
                                                     loss_scale, dlb_ptr, self.dlb.stride(0), self.SCp, self.tc_ma, self.dlTb.data_ptr(),
                                                     self.dlTb.stride(0), self.Bp, self.tc_ma, st), "softmax_grad")
        else:
            ck(lib.clipgp_softmax_ce(self.logits.data_ptr(), Cn, self.in_lab.data_ptr(), rows, rpl, Cn, None, self._loss_acc.data_ptr(),
                                     loss_scale, self.logits.data_ptr(), Cn, loss_scale, st), "softmax_ce")

    def _bwd_features(self):
        """df_hat = scale * dlogits P_hat -> normalisation adjoint -> dW = dY^T f (+ L2 regulariser, adapter.py:468-476)"""
        lib, cfg, ck, st = self.lib, self.cfg, _lib.check, _lib.stream_ptr(self.dev)
        if not cfg.train_visual_proj:
            return
        B, D = self.B, self.D
        per_sample, S, SC, alpha = self._dims()
        Bmat = self.P_hat if per_sample else self.P_mean
        W = self._ptr(self.flat_p, "W")
        tcm = cfg.precision in ("bf16x3", "bf16")
        tf = cfg.precision == "tf32"
        if tf:
            # d f_hat = alpha dlogits P_hat: A = dlogits [B, SC] (K-major), B operand = P_hat^T, i.e. P_hat [SC, D] read MN-major in place
            if self.tf32_kmajor_b:
                ck(lib.clipgp_transpose_f32(Bmat.data_ptr(), SC, D, D, self.PT32.data_ptr(), SC, st), "transpose(P_hat)")
                self._tf32(self.logits.data_ptr(), False, B, self.PT32.data_ptr(), False, D, SC, alpha, self.df_hat.data_ptr(), D)
            else:
                self._tf32(self.logits.data_ptr(), False, B, Bmat.data_ptr(), True, D, SC, alpha, self.df_hat.data_ptr(), D)
        elif tcm:
            # dlogits [B, SC] and P_hat^T [D, SC] (K = samples x classes; split over K inside the GEMM)
            self._cast2(Bmat.data_ptr(), SC, D, D, None, 0, 0, self.PTb, self.SCp, self.tc_mb)
            self._tc(self.dlb, self.PTb, alpha, self.df_hat.data_ptr(), D)
        else:
            ck(lib.clipgp_gemm_f32(self.logits.data_ptr(), SC, 1, Bmat.data_ptr(), D, 1, self.df_hat.data_ptr(), D, B, D, SC,
                                   alpha, 0, st), "gemm(df)")
        ck(lib.clipgp_rownorm_backward(self.df_hat.data_ptr(), self.f_hat.data_ptr(), self.f_inv.data_ptr(), B, D,
                                       self.dY.data_ptr(), st), "rownorm_bwd")
        if tf:
            # dW = dY^T f: both operands are [B, .] tensors contracted over the batch -> both read MN-major in place
            self._tf32(self.dY.data_ptr(), True, D, self.in_feat.data_ptr(), True, D, B, 1.0, self._ptr(self.flat_g, "W"), D)
        elif tcm:
            self._cast2(self.dY.data_ptr(), B, D, D, None, 0, 0, self.dYTb, self.Bp, self.tc_ma)
            self._tc(self.dYTb, self.fTb, 1.0, self._ptr(self.flat_g, "W"), D)
        else:
            ck(lib.clipgp_gemm_f32(self.dY.data_ptr(), 1, D, self.in_feat.data_ptr(), D, 1, self._ptr(self.flat_g, "W"), D, D, D, B,
                                   1.0, 0, st), "gemm(dW)")
        coef = float(cfg.l2_lambda) / float(cfg.shots) / cfg.world
        ck(lib.clipgp_l2_identity(W, D, coef, self._ptr(self.flat_g, "W"), self._loss_acc.data_ptr(), st), "l2_identity")
        if cfg.world == 1 and not getattr(self, "skip_update", False):
            self._adamw_w()                                   # dW is complete here: update W next to the GP adjoint

    def _bwd_prototypes(self):
        """dP_hat = scale * dlogits^T f_hat -> prototype + GP adjoints (dkl_scalar = gp_beta: adapter.py:462-465)"""
        lib, cfg, ck, st = self.lib, self.cfg, _lib.check, _lib.stream_ptr(self.dev)
        B, Cn, T, D = self.B, self.C, self.T, self.D
        per_sample, S, SC, alpha = self._dims()
        if cfg.precision == "tf32":
            # d P_hat = alpha dlogits^T f_hat (contraction over the batch): dlogits [B, SC] and f_hat [B, D] read MN-major in place
            if self.tf32_kmajor_b:
                ck(lib.clipgp_transpose_f32(self.f_hat.data_ptr(), B, D, D, self.fhT32.data_ptr(), B, st), "transpose(f_hat)")
                self._tf32(self.logits.data_ptr(), True, SC, self.fhT32.data_ptr(), False, D, B, alpha, self.dP.data_ptr(), D)
            else:
                self._tf32(self.logits.data_ptr(), True, SC, self.f_hat.data_ptr(), True, D, B, alpha, self.dP.data_ptr(), D)
        elif cfg.precision != "fp32":
            # K-major operands: dlogits^T [SC, B] and f_hat^T [D, B] (K = batch)
            if not self.tl_adjoint:
                self._tc(self.dlTb, self.fhTb, alpha, self.dP.data_ptr(), D)
        else:
            ck(lib.clipgp_gemm_f32(self.logits.data_ptr(), 1, SC, self.f_hat.data_ptr(), D, 1, self.dP.data_ptr(), D, SC, D, B,
                                   alpha, 0, st), "gemm(dP)")
        # per-sample: dense [S,C,D] gradient; logit-mean: the mean over samples happened on unit rows, every sample receives
        # dP_mean (the 1/S is inside alpha)
        if self.class_sharded:
            self.dw_all.zero_()
        if not self.fused_proto_bwd:
            ck(lib.clipgp_proto_backward(self.dP.data_ptr(), Cn * D if per_sample else 0, 1.0, self.P_hat.data_ptr(), self.P_norm.data_ptr(),
                                         self.E.data_ptr(), S, Cn, T, D, self.dw.data_ptr(), st), "proto_backward")
        if self.class_sharded:
            torch.distributed.all_reduce(self.dw_all)         # every rank wrote its samples (all classes); rows of other ranks are zero
        ck(lib.clipgp_gp_backward(C.byref(self.gp_args), C.byref(self.gp_bwd_args), st), "gp_backward")
        if not (self._fused_tail() or (self.peer is not None and not getattr(self, "skip_update", False))):
            # (single GPU: the KL sum rides in the fused tail kernel of _launch_update; peer mode: in clipgp_peer_adamw)
            ck(lib.clipgp_sum_accumulate(self.kl.data_ptr() + 4 * self.c_lo, self.c_hi - self.c_lo, self.kl_weight, self._loss_acc.data_ptr(), st),
               "kl_sum")

    def _adamw_w(self):
        lib, cfg, st = self.lib, self.cfg, _lib.stream_ptr(self.dev)
        _, nW = self.offsets["W"]
        b1, b2 = cfg.betas
        _lib.check(lib.clipgp_adamw_step_lrptr(self.flat_p.data_ptr(), self.flat_g.data_ptr(), self.flat_m.data_ptr(), self.flat_v.data_ptr(),
                                               nW, self.lr_dev.data_ptr(), b1, b2, cfg.adam_eps, cfg.weight_decay, self.adam_step.data_ptr(),
                                               st), "adamw(W)")

    def _peer_args(self):
        a = _lib.PeerArgs()
        cfg, pb = self.cfg, self.peer
        a.world, a.rank = cfg.world, cfg.rank
        for q in range(cfg.world):
            a.g[q], a.p[q], a.flags[q] = pb.ptr(q, "g"), pb.ptr(q, "p"), pb.ptr(q, "flags")
        a.m, a.v = self.flat_m.data_ptr(), self.flat_v.data_ptr()
        a.n = self.n_params
        a.n_group0 = self.offsets["W"][1] if cfg.train_visual_proj else 0
        a.lr_dev = self.lr_dev.data_ptr()
        a.beta1, a.beta2 = cfg.betas
        a.eps, a.weight_decay = cfg.adam_eps, cfg.weight_decay
        a.step, a.local = self.adam_step.data_ptr(), self.peer_local.data_ptr()
        a.loss_out, a.status = self.loss_global.data_ptr(), self.peer_status.data_ptr()
        a.timeout_ns = int(5e9)
        a.kl, a.kl_n, a.kl_scale = self.kl.data_ptr() + 4 * self.c_lo, self.c_hi - self.c_lo, self.kl_weight      # the KL sum rides in the kernel
        return a

    def _launch_peer_update(self):
        lib, st = self.lib, _lib.stream_ptr(self.dev)
        if getattr(self, "_peer_args_c", None) is None:
            self._peer_args_c = self._peer_args()
        _lib.check(lib.clipgp_peer_adamw(C.byref(self._peer_args_c), st), "peer_adamw")
        _lib.check(lib.clipgp_step_epilogue(self._ptr(self.flat_p, "z_last"), self.Z.data_ptr(), self.C, self.n, self.d, self.adam_step.data_ptr(),
                                            self.rng_state.data_ptr() + 8, 1, st), "step_epilogue")

    def check_peer_status(self):
        """Raise if a peer flag timed out inside clipgp_peer_adamw (host read: call outside the timed region)."""
        if self.peer is not None and int(self.peer_status.item()) != 0:
            raise RuntimeError(f"clipgp_peer_adamw: peer flag timeout (status {int(self.peer_status.item())}) on rank {self.cfg.rank}")

    def close_peer(self):
        """Drop the captured graph and unmap / free the NVLink peer block (every rank, before destroy_process_group)."""
        self._graph = None
        if self.peer is not None:
            torch.cuda.synchronize(self.dev)
            self.check_peer_status()
            torch.distributed.barrier()
            fp, fg = self.flat_p.clone(), self.flat_g.clone()
            self.peer.close()
            self.flat_p, self.flat_g, self.peer = fp, fg, None

    def _fused_tail(self) -> bool:
        """Single GPU, updating step: KL sum + AdamW(gp group) + inducing-row scatter + counters are ONE launch (clipgp_adamw_tail)."""
        return bool(self.cfg.world == 1 and self.cfg.fuse_tail and not getattr(self, "skip_update", False) and self.offsets["W"][1] % 4 == 0)

    def _launch_update(self):
        lib, cfg, st = self.lib, self.cfg, _lib.stream_ptr(self.dev)
        ck = _lib.check
        oW, nW = self.offsets["W"]
        b1, b2 = cfg.betas
        if self._fused_tail():
            if getattr(self, "_tail_ticket", None) is None:
                self._tail_ticket = torch.zeros(1, dtype=torch.int32, device=self.dev)
            rest = self.n_params - nW
            zo, _ = self.offsets["z_last"]
            ck(lib.clipgp_adamw_tail(self.flat_p.data_ptr() + 4 * nW, self.flat_g.data_ptr() + 4 * nW, self.flat_m.data_ptr() + 4 * nW,
                                     self.flat_v.data_ptr() + 4 * nW, rest, self.lr_dev.data_ptr() + 4, b1, b2, cfg.adam_eps, cfg.weight_decay,
                                     self.adam_step.data_ptr(), zo - nW, self.C, self.n, self.d, self.Z.data_ptr(),
                                     self.kl.data_ptr() + 4 * self.c_lo, self.c_hi - self.c_lo, self.kl_weight, self._loss_acc.data_ptr(),
                                     self.rng_state.data_ptr() + 8, 1, self._tail_ticket.data_ptr(), st), "adamw_tail")
            return
        if cfg.train_visual_proj and cfg.world > 1:
            self._adamw_w()                                   # multi-GPU: after the gradient all-reduce
        rest = self.n_params - nW
        ck(lib.clipgp_adamw_step_lrptr(self.flat_p.data_ptr() + 4 * nW, self.flat_g.data_ptr() + 4 * nW, self.flat_m.data_ptr() + 4 * nW,
                                       self.flat_v.data_ptr() + 4 * nW, rest, self.lr_dev.data_ptr() + 4, b1, b2, cfg.adam_eps,
                                       cfg.weight_decay, self.adam_step.data_ptr(), st), "adamw(gp)")
        ck(lib.clipgp_step_epilogue(self._ptr(self.flat_p, "z_last"), self.Z.data_ptr(), self.C, self.n, self.d, self.adam_step.data_ptr(),
                                    self.rng_state.data_ptr() + 8, 1, st), "step_epilogue")

    # ------------------------------------------------------------------ public API
    def set_lr(self, lr: Optional[float] = None, gp_lr: Optional[float] = None) -> None:
        """Learning rates of the two parameter groups (adapter.py:298-309) for the next steps.  They live in device memory, so a
        scheduler (CosineAnnealingLR stepped per epoch in the reference, adapter.py:1054-1056) needs no graph re-capture."""
        if lr is not None:
            self.cfg.lr = float(lr)
        if gp_lr is not None:
            self.cfg.gp_lr = float(gp_lr)
        self.lr_dev.copy_(torch.tensor([self.cfg.lr, self.cfg.gp_lr], dtype=torch.float32), non_blocking=False)

    def cosine_lr(self, epoch: int, max_epoch: int, base_lr: float, base_gp_lr: float, eta_min: float = 0.0) -> None:
        """torch CosineAnnealingLR(T_max=max_epoch, eta_min) evaluated in closed form at `epoch` (utils/optimization.py:232-238)."""
        import math
        f = 0.5 * (1.0 + math.cos(math.pi * epoch / max_epoch))
        self.set_lr(eta_min + (base_lr - eta_min) * f, eta_min + (base_gp_lr - eta_min) * f)

    def train_step(self, features: torch.Tensor, labels: torch.Tensor, use_graph: bool = True) -> torch.Tensor:
        """One optimisation step on a [B,D] feature batch (host or device tensors).  Returns the device scalar loss."""
        if features.shape[0] != self.B:
            self._alloc_train(features.shape[0])
            self._graph = None
        self.in_feat.copy_(features, non_blocking=True)
        self.in_lab.copy_(labels, non_blocking=True)
        with torch.cuda.device(self.dev):
            if use_graph and (self.cfg.world == 1 or self.cfg.graph_collectives):
                if self._graph is None:
                    self._capture()
                self._graph.replay()
            else:
                self._launch_step()
        return self.loss

    def train_steps_host(self, f_host: torch.Tensor, y_host: torch.Tensor, batches, use_graph: bool = True) -> torch.Tensor:
        """Run one optimisation step per (lo, hi) row range of PINNED host tensors and return the per-step losses (host tensor).
        The reference moves every batch to the device and reads the loss back synchronously (adapter.py:729-746, 355-361); here the
        next batch's host-to-device copy runs on a copy stream while the current step computes (two staging buffers), and each step's
        loss is copied to pinned host memory asynchronously; the only host synchronisation is the one at the end."""
        batches = list(batches)
        if not batches:
            return torch.empty(0)
        B = batches[0][1] - batches[0][0]
        if any(hi - lo != B for lo, hi in batches):
            raise ValueError("train_steps_host: all batches must have the same size (drop_last, utils/data_manager.py:79)")
        if not (f_host.is_pinned() and y_host.is_pinned()):
            raise ValueError("train_steps_host: host tensors must be pinned (tensor.pin_memory())")
        if B != self.B:
            self._alloc_train(B)
            self._graph = None
        dev = self.dev
        with torch.cuda.device(dev):
            main = torch.cuda.current_stream(dev)
            if getattr(self, "_copy_stream", None) is None:
                self._copy_stream = torch.cuda.Stream(dev)
            cp = self._copy_stream
            if getattr(self, "_stage", None) is None or self._stage[0][0].shape[0] != B:
                self._stage = [(torch.empty_like(self.in_feat), torch.empty_like(self.in_lab)) for _ in range(2)]
            losses = torch.empty(len(batches), dtype=torch.float32).pin_memory()
            if use_graph and self._graph is None and (self.cfg.world == 1 or self.cfg.graph_collectives):
                self._capture()
            loaded = [torch.cuda.Event(), torch.cuda.Event()]
            consumed = [torch.cuda.Event(), torch.cuda.Event()]

            def prefetch(i):
                lo, hi = batches[i]
                sf, sy = self._stage[i % 2]
                with torch.cuda.stream(cp):
                    if i >= 2:
                        cp.wait_event(consumed[i % 2])         # step i - 2 has copied this staging buffer into the step's input
                    sf.copy_(f_host[lo:hi], non_blocking=True)
                    sy.copy_(y_host[lo:hi], non_blocking=True)
                    loaded[i % 2].record(cp)

            cp.wait_stream(main)
            prefetch(0)
            for i in range(len(batches)):
                if i + 1 < len(batches):
                    prefetch(i + 1)
                sf, sy = self._stage[i % 2]
                main.wait_event(loaded[i % 2])
                self.in_feat.copy_(sf, non_blocking=True)
                self.in_lab.copy_(sy, non_blocking=True)
                consumed[i % 2].record(main)
                if use_graph and self._graph is not None:
                    self._graph.replay()
                else:
                    self._launch_step()
                losses[i:i + 1].copy_(self.loss.view(1), non_blocking=True)
            main.synchronize()
        return losses

    def _capture(self):
        # warm-up outside capture (sets kernel attributes), with all state restored afterwards
        snap = (self.flat_p.clone(), self.flat_m.clone(), self.flat_v.clone(), self.Z.clone(), self.adam_step.clone(),
                self.rng_state.clone())
        s = torch.cuda.Stream(self.dev)
        s.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(s):
            self._launch_step()
        torch.cuda.current_stream(self.dev).wait_stream(s)
        torch.cuda.synchronize(self.dev)
        for dst, src in zip((self.flat_p, self.flat_m, self.flat_v, self.Z, self.adam_step, self.rng_state), snap):
            dst.copy_(src)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._launch_step()
        # the capture did not execute anything; state is still the snapshot
        self._graph = g

    def _eval_noise(self, a: GpArgs, S: int):
        """Point a copy of the GP arguments at the evaluation noise: explicit `self.eval_eps` (kept alive by the caller) or the
        evaluation counter stream, which is advanced by one step per call (a device-side increment, so a captured eval graph
        draws fresh noise on every replay)."""
        if self.eval_eps is not None:
            e = self.eval_eps
            if e.shape[0] != self.C or e.shape[1] < self.T or e.shape[2] != S or e.dtype != torch.float32 or e.device != self.dev:
                raise ValueError(f"eval_eps must be a float32 [C, >=T, S={S}] tensor on {self.dev}; got {tuple(e.shape)} {e.dtype} {e.device}")
            a.eps, a.eps_sc, a.eps_st, a.eps_ss, a.rng_state = e.data_ptr(), e.stride(0), e.stride(1), e.stride(2), None
            return False
        a.eps, a.rng_state = None, self.eval_rng_state.data_ptr()
        return True

    def _eval_ksave(self):
        if getattr(self, "_ksave_eval", None) is None:
            self._ksave_eval = torch.empty_like(self.Ksave)
        return self._ksave_eval.data_ptr()

    @torch.no_grad()
    def eval_prototypes(self, S: Optional[int] = None) -> torch.Tensor:
        """(1/S) sum_s p_hat_s [C,D] with the current parameters (collapsed form of adapter.py:243-249)."""
        S = int(S or self.cfg.S_eval)
        lib, st = self.lib, _lib.stream_ptr(self.dev)
        Cn, T, D, n = self.C, self.T, self.D, self.n
        f32 = dict(dtype=torch.float32, device=self.dev)
        w = torch.empty(S, Cn, T, **f32)
        a = GpArgs.from_buffer_copy(self.gp_args)
        a.S, a.s_offset, a.S_total = S, 0, S
        a.c_begin, a.c_count = 0, 0
        a.eps_save = None
        a.proto_E = a.proto_P_hat = a.proto_norm = a.proto_bf16 = a.proto_mean_hat = None
        a.w, a.kl = w.data_ptr(), None
        a.L = a.A = a.R = None
        a.Ksave = self._eval_ksave()                            # hand-over buffer of the two-kernel fast path
        bump = self._eval_noise(a, S)
        fused = bool(self.cfg.fuse_prototypes and lib.clipgp_gp_fused_proto_ok(T, n, self.d, D, S))
        # multi-GPU: this rank's CTAs cover its class shard only; the all-reduce below completes the [C, D] matrix on every rank
        # (the draws are keyed by (seed, step, c, t, s), so the shard does not change them)
        shard = bool(fused and self.cfg.world > 1 and self.cfg.shard_eval_classes)
        Pm = torch.zeros(Cn, D, **f32) if shard else torch.empty(Cn, D, **f32)
        if shard:
            lo, hi = dist.shard_range(Cn, self.cfg.rank, self.cfg.world)
            a.c_begin, a.c_count = lo, max(hi - lo, 0)
        if fused:       # the class's CTA also averages its unit prototypes: no separate prototype pass
            a.proto_E, a.proto_D, a.proto_mean_hat = self.E.data_ptr(), D, Pm.data_ptr()
        with torch.cuda.device(self.dev):
            if not shard or a.c_count > 0:
                _lib.check(lib.clipgp_gp_forward(C.byref(a), st), "gp_forward(eval)")
            if not fused:
                _lib.check(lib.clipgp_proto_forward(w.data_ptr(), self.E.data_ptr(), S, Cn, T, D, None, 0.0, None, None, None, None,
                                                    Pm.data_ptr(), None, 1, st), "proto_forward(eval)")
        if shard:
            torch.distributed.all_reduce(Pm)
        if bump:
            self.eval_rng_state[1:2].add_(1)
        self.last_eval_w = w
        return Pm

    @torch.no_grad()
    def eval_logits(self, features: torch.Tensor, prototypes_mean: Optional[torch.Tensor] = None, S: Optional[int] = None):
        """MC-averaged logits [N,C] = scale * normalize(f W^T) . mean_s p_hat_s  (adapter.py:239-249)."""
        lib, st = self.lib, _lib.stream_ptr(self.dev)
        Pm = prototypes_mean if prototypes_mean is not None else self.eval_prototypes(S)
        f = features.to(self.dev, non_blocking=True).float().contiguous()
        N, D, Cn = f.shape[0], self.D, self.C
        Y = torch.empty(N, D, dtype=torch.float32, device=self.dev)
        logits = torch.empty(N, Cn, dtype=torch.float32, device=self.dev)
        with torch.cuda.device(self.dev):
            for lo in range(0, N, 32768):                      # grid.y limit of the fp32 GEMM
                hi = min(N, lo + 32768)
                _lib.check(lib.clipgp_gemm_f32(f[lo:hi].data_ptr(), D, 1, self._ptr(self.flat_p, "W"), 1, D, Y[lo:hi].data_ptr(), D,
                                               hi - lo, D, D, 1.0, 0, st), "gemm(proj)")
            _lib.check(lib.clipgp_rownorm_forward(Y.data_ptr(), N, D, Y.data_ptr(), None, None, st), "rownorm")
            for lo in range(0, N, 32768):
                hi = min(N, lo + 32768)
                _lib.check(lib.clipgp_gemm_f32(Y[lo:hi].data_ptr(), D, 1, Pm.data_ptr(), 1, D, logits[lo:hi].data_ptr(), Cn,
                                               hi - lo, Cn, D, self.cfg.logit_scale, 0, st), "gemm(logits)")
        return logits

    @torch.no_grad()
    def evaluate(self, features: torch.Tensor, labels: torch.Tensor, S: Optional[int] = None, n_bins: int = 10,
                 precision: str = "fp32", mc: str = "collapsed"):
        """Accuracy / ECE / AECE of the MC-averaged logits for this rank's shard.

        precision: "fp32" (FFMA GEMMs, logits materialised, exact-mode comparator) | "bf16" | "bf16x3" (tcgen05 GEMMs with the
        calibration epilogue fused; bf16x3 = split operands, fp32-grade products).  mc: "collapsed" (one [N,D]x[C,D]^T GEMM
        against mean_s p_hat_s, exact for the reference's logit-mean) | "materialised" (all S samples: [N,D]x[S*C,D]^T flops,
        accumulated over s inside TMEM; tensor modes only)."""
        labels = labels.to(self.dev)
        if precision == "fp32":
            logits = self.eval_logits(features, S=S)
            return metrics.evaluate_calibration(logits, labels, n_bins)
        conf, correct, hist = self.eval_calibration_tc(features, labels, S=S, n_bins=n_bins, precision=precision, mc=mc)
        n = int(labels.numel())
        cnt = metrics.counters_from_hist(hist, n)
        ece, calib = metrics.ece_from_counters(cnt)
        _, out = metrics.aece_pass(conf, correct, n_bins)
        aece, acalib = metrics.aece_from_bins(out, n, n_bins)
        return {"top1_acc": cnt.top1 * (100.0 / max(n, 1)), "ece": ece, "aece": aece, "calibration": calib,
                "adaptive_calibration": acalib, "n": n, "top1_count": cnt.top1, "counters": cnt}

    @torch.no_grad()
    def eval_operands_tc(self, S: Optional[int] = None, precision: str = "bf16", mc: str = "collapsed"):
        """bf16 B operand of the eval GEMM: mean_s p_hat_s [C, D] (collapsed) or [C, S*D] with the samples along K."""
        from . import tc
        S = int(S or self.cfg.S_eval)
        split = precision == "bf16x3"
        lib, st = self.lib, _lib.stream_ptr(self.dev)
        Cn, T, D = self.C, self.T, self.D
        if mc == "collapsed":
            Pm = self.eval_prototypes(S)
            return tc.cast_bf16(Pm, tc.SPLIT_B if split else tc.PLAIN), 1.0
        f32 = dict(dtype=torch.float32, device=self.dev)
        w = torch.empty(S, Cn, T, **f32)
        a = GpArgs.from_buffer_copy(self.gp_args)
        a.S, a.s_offset, a.S_total = S, 0, S
        a.c_begin, a.c_count = 0, 0                             # evaluation: every rank needs the prototypes of all classes
        a.eps_save = None
        a.proto_E = a.proto_P_hat = a.proto_norm = a.proto_bf16 = a.proto_mean_hat = None
        a.w, a.kl = w.data_ptr(), None
        a.L = a.A = a.R = None
        a.Ksave = self._eval_ksave()                            # hand-over buffer of the two-kernel fast path
        bump = self._eval_noise(a, S)
        P_hat = torch.empty(S, Cn, D, **f32)
        seg = 3 if split else 1
        Bop = torch.empty(Cn, S * seg * D, dtype=torch.bfloat16, device=self.dev)
        with torch.cuda.device(self.dev):
            _lib.check(lib.clipgp_gp_forward(C.byref(a), st), "gp_forward(eval)")
            _lib.check(lib.clipgp_proto_forward(w.data_ptr(), self.E.data_ptr(), S, Cn, T, D, None, 0.0, None, P_hat.data_ptr(), None,
                                                None, None, None, 0, st), "proto_forward(eval)")
            for s_ in range(S):     # row c of the operand = [p_hat_1c | ... | p_hat_Sc] (each optionally [hi|lo|hi])
                _lib.check(lib.clipgp_cast_bf16(P_hat[s_].data_ptr(), Cn, D, D, Bop.data_ptr() + 2 * s_ * seg * D, S * seg * D, D,
                                                2 if split else 0, st), "cast_bf16(P)")
        if bump:
            self.eval_rng_state[1:2].add_(1)
        return Bop, 1.0 / S

    @torch.no_grad()
    def eval_calibration_tc(self, features, labels, S=None, n_bins=10, precision="bf16", mc="collapsed", want_logits=False):
        """Tensor-core eval: cast -> projection GEMM -> row normalise -> fused logits + softmax-max + histogram GEMM."""
        from . import tc
        if precision not in ("bf16", "bf16x3", "tf32"):
            raise ValueError(f"unknown precision {precision}")
        split = precision == "bf16x3"
        lib, st = self.lib, _lib.stream_ptr(self.dev)
        f = features.to(self.dev, non_blocking=True).float().contiguous()
        N, D = f.shape
        if precision == "tf32":
            return self._eval_calibration_tf32(f, labels, S, n_bins, mc, want_logits)
        # the prototype chain (GP forward -> unit prototypes -> operand cast; latency bound) runs on a side stream next to the feature
        # chain (cast -> projection GEMM -> normalise; bandwidth / tensor bound); they meet at the logit GEMM
        cur = torch.cuda.current_stream(self.dev)
        if getattr(self, "_eval_stream", None) is None:
            self._eval_stream = torch.cuda.Stream(self.dev)
        side = self._eval_stream
        side.wait_stream(cur)
        fuse_proj = (mc == "collapsed" and not want_logits and D % 256 == 0 and self.cfg.fuse_eval_projection)
        if fuse_proj:
            # ONE GEMM for projection + normalisation + logits (clipgp_tc_proj_logits_calibration): B = [W ; Q], Q = mean_s p_hat_s W, so
            # that f . Q_c = (f W^T) . p_c; the epilogue divides by |f W^T|.  The projected / normalised features are never written.
            mA, mB = (tc.SPLIT_A, tc.SPLIT_B) if split else (tc.PLAIN, tc.PLAIN)
            seg = 3 if split else 1
            W = self.p("W").view(D, D)
            Bop = torch.empty(D + self.C, seg * D, dtype=torch.bfloat16, device=self.dev)
            with torch.cuda.device(self.dev):
                _lib.check(lib.clipgp_cast_bf16(W.data_ptr(), D, D, D, Bop.data_ptr(), seg * D, D, mB, st), "cast_bf16(W)")
            WT = tc.cast_bf16_transpose(W, mB)                                          # K-major operand of Q = P W
            with torch.cuda.stream(side):
                WT.record_stream(side); Bop.record_stream(side)
                side.wait_stream(cur)
                Pm = self.eval_prototypes(S)
                Q = tc.gemm_store(tc.cast_bf16(Pm, mA), WT, 1.0)
                with torch.cuda.device(self.dev):
                    _lib.check(lib.clipgp_cast_bf16(Q.data_ptr(), self.C, D, D, Bop.data_ptr() + 2 * D * seg * D, seg * D, D, mB,
                                                    _lib.stream_ptr(self.dev)), "cast_bf16(Q)")
            fb = tc.cast_bf16(f, mA)
            cur.wait_stream(side)
            conf, correct, hist = tc.proj_logits_calibration(fb, Bop, D, self.cfg.logit_scale, labels, n_bins)
            self.last_eval_logits = None
            return conf, correct, hist
        with torch.cuda.stream(side):
            Bop, mc_scale = self.eval_operands_tc(S, precision, mc)
        Bop.record_stream(cur)
        Wb = tc.cast_bf16(self.p("W").view(D, D), tc.SPLIT_B if split else tc.PLAIN)
        fb = tc.cast_bf16(f, tc.SPLIT_A if split else tc.PLAIN)
        Y = tc.gemm_store(fb, Wb, 1.0)                                              # adapter.py:239
        seg = 3 if split else 1
        fhat_b = torch.empty(N, seg * D, dtype=torch.bfloat16, device=self.dev)
        with torch.cuda.device(self.dev):
            # adapter.py:240 fused with the operand cast: the fp32 unit rows are never written
            _lib.check(lib.clipgp_rownorm_cast(Y.data_ptr(), N, D, None, None, fhat_b.data_ptr(), seg * D, D, 1 if split else 0, st),
                       "rownorm_cast")
        cur.wait_stream(side)
        conf, correct, hist, logits = tc.logits_calibration(fhat_b, Bop, self.cfg.logit_scale * mc_scale, labels, n_bins,
                                                            want_conf=True, want_logits=want_logits)
        self.last_eval_logits = logits
        return conf, correct, hist

    def _eval_calibration_tf32(self, f, labels, S, n_bins, mc, want_logits):
        """The eval pass in TF32 on the fp32 tensors in place: no operand casts at all.  Collapsed + D % 256 == 0: ONE GEMM over the
        raw features against B = [W ; Q], Q = mean_s p_hat_s W (written by a small TF32 GEMM straight into the operand, W read
        transposed in place); otherwise projection GEMM -> row normalise -> logits GEMM."""
        from . import tc
        lib, st = self.lib, _lib.stream_ptr(self.dev)
        N, D = f.shape
        Cn = self.C
        if D % 4:
            raise ValueError("precision='tf32' needs D % 4 == 0")
        W = self.p("W").view(D, D)
        cur = torch.cuda.current_stream(self.dev)
        if getattr(self, "_eval_stream", None) is None:
            self._eval_stream = torch.cuda.Stream(self.dev)
        side = self._eval_stream
        scale = self.cfg.logit_scale
        fuse_proj = (mc == "collapsed" and not want_logits and D % 256 == 0 and self.cfg.fuse_eval_projection)
        if mc == "collapsed":
            if fuse_proj:
                Bop = torch.empty(D + Cn, D, dtype=torch.float32, device=self.dev)
                Bop[:D].copy_(W)
                Bop.record_stream(side)
                side.wait_stream(cur)
                with torch.cuda.stream(side):                   # prototype chain next to nothing: the features need no preparation
                    Pm = self.eval_prototypes(S)
                    tc.gemm_tf32(Pm, W, 1.0, b_t=True, out=Bop[D:])                    # Q = P_mean W  ([C, D_out] x [D_out, D_in])
                cur.wait_stream(side)
                conf, correct, hist, _ = tc.logits_calibration_tf32(f, Bop, scale, labels, n_bins, norm_cols=D)
                self.last_eval_logits = None
                return conf, correct, hist
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                Pm = self.eval_prototypes(S)
            Pm.record_stream(cur)
            Bmat, alpha = Pm, scale
        else:
            S_ = int(S or self.cfg.S_eval)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                P_hat = self._eval_unit_prototypes(S_)                                   # [S, C, D]
                Bmat = P_hat.permute(1, 0, 2).reshape(Cn, S_ * D).contiguous()           # row c = [p_hat_1c | ... | p_hat_Sc]: A wraps along K
            Bmat.record_stream(cur)
            alpha = scale / S_
        Y = tc.gemm_tf32(f, W, 1.0)                                                       # adapter.py:239
        with torch.cuda.device(self.dev):
            _lib.check(lib.clipgp_rownorm_forward(Y.data_ptr(), N, D, Y.data_ptr(), None, None, st), "rownorm")
        cur.wait_stream(side)
        conf, correct, hist, logits = tc.logits_calibration_tf32(Y, Bmat, alpha, labels, n_bins, want_logits=want_logits)
        self.last_eval_logits = logits
        return conf, correct, hist

    @torch.no_grad()
    def _eval_unit_prototypes(self, S: int) -> torch.Tensor:
        """p_hat_s [S, C, D] of one evaluation draw (materialised MC form)."""
        lib, st = self.lib, _lib.stream_ptr(self.dev)
        Cn, T, D = self.C, self.T, self.D
        f32 = dict(dtype=torch.float32, device=self.dev)
        w = torch.empty(S, Cn, T, **f32)
        a = GpArgs.from_buffer_copy(self.gp_args)
        a.S, a.s_offset, a.S_total = S, 0, S
        a.c_begin, a.c_count = 0, 0
        a.eps_save = None
        a.proto_E = a.proto_P_hat = a.proto_norm = a.proto_bf16 = a.proto_mean_hat = None
        a.w, a.kl = w.data_ptr(), None
        a.L = a.A = a.R = None
        a.Ksave = self._eval_ksave()
        bump = self._eval_noise(a, S)
        P_hat = torch.empty(S, Cn, D, **f32)
        with torch.cuda.device(self.dev):
            _lib.check(lib.clipgp_gp_forward(C.byref(a), st), "gp_forward(eval)")
            _lib.check(lib.clipgp_proto_forward(w.data_ptr(), self.E.data_ptr(), S, Cn, T, D, None, 0.0, None, P_hat.data_ptr(), None,
                                                None, None, None, 0, st), "proto_forward(eval)")
        if bump:
            self.eval_rng_state[1:2].add_(1)
        return P_hat

    @torch.no_grad()
    def eval_graph(self, features: torch.Tensor, labels: torch.Tensor, S=None, n_bins=10, precision="bf16x3", mc="collapsed"):
        """Capture the tensor-core eval pass over a fixed (features, labels) device shard in one CUDA graph and return a callable
        that replays it: () -> (conf, correct, hist), static output tensors.  The graph reads the engine's parameter buffers in
        place, so it stays valid across training steps (the reference evaluates the same cached test features after every step,
        adapter.py:363-380); ~15 launches and their allocations collapse into one replay."""
        f = features.to(self.dev).float().contiguous()
        y = labels.to(self.dev).to(torch.int64).contiguous()
        from . import metrics as _m
        _m._boundaries(n_bins, self.dev)                        # host-to-device copies must happen before the capture
        for _ in range(2):                                      # warm-up: kernel attributes, tensor-map encoder, side stream
            self.eval_calibration_tc(f, y, S=S, n_bins=n_bins, precision=precision, mc=mc)
        torch.cuda.synchronize(self.dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = self.eval_calibration_tc(f, y, S=S, n_bins=n_bins, precision=precision, mc=mc)

        holder = [g]
        del g

        def replay():
            holder[0].replay()
            return out
        replay.inputs = (f, y)
        replay.release = holder.clear          # a captured graph that holds NCCL work must be dropped before the process group goes away
        return replay

    @torch.no_grad()
    def eval_metrics_graph(self, features: torch.Tensor, labels: torch.Tensor, n_total: int, S=None, n_bins=10,
                           precision="bf16x3", mc="collapsed"):
        """The WHOLE evaluation metric of this rank's image shard as one captured CUDA graph: the tensor-core eval pass
        (eval_calibration_tc) -> all-reduce of the integer calibration counters -> all-gather of the (confidence, hit) pairs ->
        AECE rank-select over the global set.  Returns a callable () -> (conf, correct, hist_global [4, n_bins] int64,
        aece_bins [3, nb] int64) with static outputs; `.inputs` are the static (features, labels) to copy new data into,
        `.release()` drops the graph (before the process group is destroyed).  `n_total` = images over all ranks."""
        from . import metrics as _m
        f = features.to(self.dev).float().contiguous()
        y = labels.to(self.dev).to(torch.int64).contiguous()
        world = self.cfg.world
        _m._boundaries(n_bins, self.dev)

        def whole():
            conf, correct, hist = self.eval_calibration_tc(f, y, S=S, n_bins=n_bins, precision=precision, mc=mc)
            hg, cg, og = dist.global_calibration(hist, conf, correct, n_total, world)
            _, aout = _m.aece_pass(cg, og, n_bins)
            return conf, correct, hg, aout
        for _ in range(2):                                      # warm-up: kernel attributes, edge cache, NCCL channels
            whole()
        torch.cuda.synchronize(self.dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = whole()
        holder = [g]
        del g

        def replay():
            holder[0].replay()
            return out
        replay.inputs = (f, y)
        replay.release = holder.clear
        return replay

    # ------------------------------------------------------------------ sync back to the nn.Module
    @torch.no_grad()
    def export_to_module(self, gpw: GaussianProcessTemplateWeighter, visual_proj: Optional[torch.nn.Linear] = None):
        gpw.variational_strategy.inducing_points.copy_(self.Z)
        q = gpw.variational_strategy._variational_distribution
        q.variational_mean.copy_(self.p("m").view_as(q.variational_mean))
        q.chol_variational_covar.copy_(self.p("Lq").view_as(q.chol_variational_covar))
        raw_ls, raw_os, raw_var = gpw._kernel_raw()
        if raw_ls is not None: raw_ls.copy_(self.p("ls").view_as(raw_ls))
        if raw_os is not None: raw_os.copy_(self.p("os").view_as(raw_os))
        if raw_var is not None: raw_var.copy_(self.p("var").view_as(raw_var))
        if visual_proj is not None:
            visual_proj.weight.copy_(self.p("W").view(self.D, self.D))
