"""Tip-Adapter-F training step (trainers/tip_adapter.py:227-269) as a fixed launch sequence on preallocated buffers.

The reference trains an ``nn.Linear(D, N_tr, bias=False)`` initialised to the cache keys with AdamW(lr, eps) and a per-step
cosine schedule; per batch:  affinity = f W^T,  cache = exp(-(beta - beta affinity)) @ one_hot(labels_tr),
tip = clip_logits + alpha cache,  CE.  Here the step is ten launches through the C ABI (no autograd, no allocation), captured in
one CUDA graph:

    cast f -> bf16 operand + transposed operand | cast W -> bf16 operand | tcgen05 affinity GEMM | exp / class-segmented sum (+ clip
    logits) | softmax CE + dlogits | d affinity in place | transposed cast | tcgen05 key-gradient GEMM (contraction over the batch) |
    AdamW | step counter

``precision``: "bf16x3" (split-bf16 operands, fp32-grade products; default), "bf16" (stated tolerance) or "fp32" (FFMA GEMMs,
the exact comparator).  The learning rate lives in device memory, so the per-step cosine schedule needs no re-capture.
"""
from __future__ import annotations

import math
from typing import Optional

import torch

from . import _lib

_MODE = {"bf16": (0, 0), "bf16x3": (1, 2)}


class TipAdapterEngine:
    def __init__(self, keys: torch.Tensor, key_labels: torch.Tensor, num_classes: int, batch_size: int, beta: float, alpha: float,
                 lr: float = 1e-3, eps: float = 1e-4, betas=(0.9, 0.999), weight_decay: float = 1e-2, total_steps: int = 1,
                 precision: str = "bf16x3"):
        dev = _lib.require_cuda(keys, key_labels)
        if precision not in ("fp32", "bf16x3", "bf16"):
            raise ValueError(f"TipAdapterEngine: unknown precision {precision!r}")
        self.lib = _lib.load()
        self.dev = dev
        self.N_tr, self.D = keys.shape
        if precision != "fp32" and self.D % 8:
            precision = "fp32"
        self.precision = precision
        self.C, self.B = int(num_classes), int(batch_size)
        self.beta, self.alpha = float(beta), float(alpha)
        self.betas, self.eps, self.weight_decay = betas, float(eps), float(weight_decay)
        self.base_lr, self.total_steps = float(lr), max(1, int(total_steps))
        f32 = dict(dtype=torch.float32, device=dev)
        self.keys = keys.detach().to(**f32).clone().contiguous()               # the trainable nn.Linear weight
        self.key_labels = key_labels.to(torch.int64).contiguous()
        self.m, self.v = torch.zeros_like(self.keys), torch.zeros_like(self.keys)
        self.dkeys = torch.empty_like(self.keys)
        self.adam_step = torch.ones(1, dtype=torch.int64, device=dev)
        self.lr_dev = torch.full((1,), self.base_lr, **f32)
        t = torch.arange(self.total_steps + 1, dtype=torch.float64)
        self.lr_table = (0.5 * self.base_lr * (1.0 + torch.cos(math.pi * t / self.total_steps))).to(**f32)   # CosineAnnealingLR, eta_min 0
        self.steps_done = 0
        self.loss = torch.zeros(1, **f32)
        self._graph: Optional[torch.cuda.CUDAGraph] = None
        self._kb_valid = False                                  # bf16 operand of the keys: cast once, then maintained by the AdamW kernel
        self._per_batch = {}                                    # batch size -> (buffers, captured graph): the last partial batch of an epoch
        self._use(self.B)

    _BUFS = ("B", "Bp", "in_feat", "in_clip", "in_lab", "aff", "logits", "fa", "fT", "kb", "GT", "_graph")

    def _use(self, B: int):
        """Switch to the buffer set (and graph) of batch size B, allocating it on first use."""
        if getattr(self, "in_feat", None) is not None:
            self._per_batch[self.B] = {k: getattr(self, k, None) for k in self._BUFS}
        if B in self._per_batch:
            for k, v in self._per_batch[B].items():
                setattr(self, k, v)
            return
        self._graph = None
        self._alloc(B)

    def _alloc(self, B: int):
        dev, D, N_tr, C = self.dev, self.D, self.N_tr, self.C
        f32 = dict(dtype=torch.float32, device=dev)
        self.B = B
        self.Bp = (B + 7) // 8 * 8
        self.in_feat = torch.zeros(B, D, **f32)
        self.in_clip = torch.zeros(B, C, **f32)
        self.in_lab = torch.zeros(B, dtype=torch.int64, device=dev)
        self.aff = torch.empty(B, N_tr, **f32)                                  # affinity -> e -> d loss / d affinity, in place
        self.logits = torch.empty(B, C, **f32)                                  # tip logits -> dlogits, in place
        if self.precision != "fp32":
            seg = 3 if self.precision == "bf16x3" else 1
            bf = dict(dtype=torch.bfloat16, device=dev)
            self.fa = torch.zeros(B, seg * D, **bf)
            self.fT = torch.zeros(D, seg * self.Bp, **bf)
            self.kb = self._per_batch[next(iter(self._per_batch))]["kb"] if self._per_batch else torch.empty(N_tr, seg * D, **bf)
            self.GT = torch.zeros(N_tr, seg * self.Bp, **bf)
        else:
            self.fa = self.fT = self.kb = self.GT = None

    # ------------------------------------------------------------------ the launch sequence
    def _launch_step(self):
        lib, ck, st = self.lib, _lib.check, _lib.stream_ptr(self.dev)
        B, D, N_tr, C, Bp = self.B, self.D, self.N_tr, self.C, self.Bp
        self.loss.zero_()
        if self.precision == "fp32":
            ck(lib.clipgp_gemm_f32(self.in_feat.data_ptr(), D, 1, self.keys.data_ptr(), 1, D, self.aff.data_ptr(), N_tr, B, N_tr, D, 1.0, 0, st),
               "gemm_f32(affinity)")
        else:
            mA, mB = _MODE[self.precision]
            ck(lib.clipgp_cast_bf16_dual(self.in_feat.data_ptr(), B, D, D, self.fa.data_ptr(), self.fa.stride(0), D, mA,
                                         self.fT.data_ptr(), self.fT.stride(0), Bp, mB, st), "cast_bf16_dual(f)")
            if not self._kb_valid:      # afterwards the AdamW kernel keeps the bf16 operand of the keys up to date
                ck(lib.clipgp_cast_bf16(self.keys.data_ptr(), N_tr, D, D, self.kb.data_ptr(), self.kb.stride(0), D, mB, st), "cast_bf16(keys)")
            ck(lib.clipgp_tc_gemm_store_splitk(self.fa.data_ptr(), B, self.fa.shape[1], self.kb.data_ptr(), N_tr, self.kb.shape[1], 1.0,
                                               self.aff.data_ptr(), N_tr, st), "tc_gemm(affinity)")
        ck(lib.clipgp_tip_forward(self.aff.data_ptr(), N_tr, self.key_labels.data_ptr(), B, N_tr, C, self.beta, self.alpha,
                                  self.in_clip.data_ptr(), C, self.logits.data_ptr(), C, 1, st), "tip_forward")
        ck(lib.clipgp_softmax_ce(self.logits.data_ptr(), C, self.in_lab.data_ptr(), B, 1, C, None, self.loss.data_ptr(), 1.0 / B,
                                 self.logits.data_ptr(), C, 1.0 / B, st), "softmax_ce")
        ck(lib.clipgp_tip_backward(self.aff.data_ptr(), N_tr, self.key_labels.data_ptr(), B, N_tr, self.logits.data_ptr(), C, self.beta,
                                   self.alpha, st), "tip_backward")
        if self.precision == "fp32":
            ck(lib.clipgp_gemm_f32(self.aff.data_ptr(), 1, N_tr, self.in_feat.data_ptr(), D, 1, self.dkeys.data_ptr(), D, N_tr, D, B, 1.0, 0, st),
               "gemm_f32(dkeys)")
        else:
            mA, _ = _MODE[self.precision]
            ck(lib.clipgp_cast_bf16_transpose(self.aff.data_ptr(), B, N_tr, N_tr, self.GT.data_ptr(), self.GT.stride(0), Bp, mA, st),
               "cast_bf16_transpose(G)")
            ck(lib.clipgp_tc_gemm_store_splitk(self.GT.data_ptr(), N_tr, self.GT.shape[1], self.fT.data_ptr(), D, self.fT.shape[1], 1.0,
                                               self.dkeys.data_ptr(), D, st), "tc_gemm(dkeys)")
        b1, b2 = self.betas
        if self.precision == "fp32":
            ck(lib.clipgp_adamw_step_lrptr(self.keys.data_ptr(), self.dkeys.data_ptr(), self.m.data_ptr(), self.v.data_ptr(), self.keys.numel(),
                                           self.lr_dev.data_ptr(), b1, b2, self.eps, self.weight_decay, self.adam_step.data_ptr(), st),
               "adamw(keys)")
        else:       # update + bf16 operand of the updated keys for the next step's affinity GEMM, one pass
            ck(lib.clipgp_adamw_step_cast(self.keys.data_ptr(), self.dkeys.data_ptr(), self.m.data_ptr(), self.v.data_ptr(), N_tr, D,
                                          self.lr_dev.data_ptr(), b1, b2, self.eps, self.weight_decay, self.adam_step.data_ptr(),
                                          self.kb.data_ptr(), self.kb.stride(0), D, _MODE[self.precision][1], st), "adamw_cast(keys)")
        ck(lib.clipgp_increment(self.adam_step.data_ptr(), 1, st), "increment")
        self._kb_valid = self.precision != "fp32"

    def _recast_keys(self):
        if self.precision != "fp32":
            _lib.check(self.lib.clipgp_cast_bf16(self.keys.data_ptr(), self.N_tr, self.D, self.D, self.kb.data_ptr(), self.kb.stride(0), self.D,
                                                 _MODE[self.precision][1], _lib.stream_ptr(self.dev)), "cast_bf16(keys)")
            self._kb_valid = True

    def _capture(self):
        snap = (self.keys.clone(), self.m.clone(), self.v.clone(), self.adam_step.clone())
        s = torch.cuda.Stream(self.dev)
        s.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(s):
            self._launch_step()                                 # warm-up outside capture; state restored below
        torch.cuda.current_stream(self.dev).wait_stream(s)
        torch.cuda.synchronize(self.dev)
        for dst, src in zip((self.keys, self.m, self.v, self.adam_step), snap):
            dst.copy_(src)
        self._recast_keys()                                     # the warm-up step left the operand of the UPDATED keys behind
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._launch_step()
        self._graph = g

    # ------------------------------------------------------------------ public API
    def train_step(self, feats_hat: torch.Tensor, clip_logits: torch.Tensor, labels: torch.Tensor, use_graph: bool = True) -> torch.Tensor:
        """One optimisation step on unit-norm features [B,D] with their frozen CLIP logits [B,C].  Returns the device scalar loss."""
        if feats_hat.shape[0] != self.B:
            self._use(feats_hat.shape[0])
        self.in_feat.copy_(feats_hat, non_blocking=True)
        self.in_clip.copy_(clip_logits, non_blocking=True)
        self.in_lab.copy_(labels, non_blocking=True)
        self.lr_dev.copy_(self.lr_table[min(self.steps_done, self.total_steps):][:1], non_blocking=True)
        with torch.cuda.device(self.dev):
            if use_graph:
                if self._graph is None:
                    self._capture()
                self._graph.replay()
            else:
                self._launch_step()
        self.steps_done += 1
        return self.loss
