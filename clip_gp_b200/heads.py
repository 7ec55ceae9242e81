"""Cosine-logit heads of the reference trainers, on CACHED frozen-CLIP features (the hot path starts after feature
extraction).  Same class / method names and argument meaning as the reference's ``CustomCLIP``
variants; every tensor op on the path is a clipgp kernel (ops.py), torch only owns parameters and autograd plumbing.

* ``AdapterHead``      — trainers/adapter.py:145-259 (visual_proj + GP / uniform template weighting)
* ``TaskResHead``      — trainers/taskres.py:35-47, 96-123
* ``ClipAdapterHead``  — trainers/clip_adapter.py:16-32, 77-100
* ``tip_*``            — trainers/tip_adapter.py:43-80, 250-260
* ``gp_pretrain``      — the full-batch ELBO loop of taskres.py:254-289 == clip_adapter.py:257-290 == tip_adapter.py:122-157
"""
from __future__ import annotations

import math
from typing import Any, Optional, Tuple

import torch
import torch.nn as nn

from . import metrics, ops
from .gp_template_weigher import GaussianProcessTemplateWeighter


def _precision(config) -> str:
    """GEMM path of the head modules: config.adapter.clipgp_precision, default "tf32" (tensor cores on the fp32 tensors in place, the
    reference's own GPU arithmetic); "fp32" selects the FFMA comparator.  The split-bf16 / bf16 modes exist in the fused engines only."""
    p = getattr(getattr(config, "adapter", config), "clipgp_precision", None) or "tf32"
    return "tf32" if p in ("tf32", "bf16x3", "bf16") else "fp32"


def _cosine_logits(feats_hat: torch.Tensor, prototypes: torch.Tensor, scale, precision: str = "fp32") -> torch.Tensor:
    """scale * f_hat . normalize(p): [B,C] for 2-D prototypes, mean over samples for [S,C,D] (adapter.py:246-251).
    The MC mean is taken over the unit prototypes BEFORE the contraction (the logits are linear in them): one [B,D] x [C,D]^T
    product instead of the reference's [B,S,C] einsum + mean -- identical values, S times fewer flops."""
    if prototypes.dim() == 3:
        S, C, D = prototypes.shape
        p_bar = ops.row_normalize(prototypes).mean(dim=0)
        return ops.matmul_nt(feats_hat, p_bar, 1.0, precision) * scale
    return ops.matmul_nt(feats_hat, ops.row_normalize(prototypes), 1.0, precision) * scale


class AdapterHead(nn.Module):
    """``CustomCLIP`` of trainers/adapter.py without the encoders: text_embeddings [K,M,D] are given."""

    def __init__(self, config: Any, text_embeddings: torch.Tensor, logit_scale: float = math.log(100.0)):
        super().__init__()
        self.config = config
        a = getattr(config, "adapter", config)
        self.register_buffer("text_embeddings", text_embeddings.detach().float())
        self.logit_scale = nn.Parameter(torch.tensor(float(logit_scale)), requires_grad=False)
        self.gp_num_mc_samples_train = int(getattr(a, "gp_num_mc_samples_train", 1) or 1)
        self.gp_num_mc_samples_eval = int(getattr(a, "gp_num_mc_samples_eval", 1) or 1)
        self.gp_weighter = GaussianProcessTemplateWeighter(text_embeddings, config) if bool(getattr(a, "use_gp", False)) else None
        dim = int(text_embeddings.shape[-1])
        self.visual_proj = nn.Linear(dim, dim, bias=False)
        with torch.no_grad():
            self.visual_proj.weight.copy_(torch.eye(dim))                      # adapter.py:187-192

    def get_prototypes(self, num_samples: int = 1, visual_embeddings: Optional[torch.Tensor] = None):
        if self.gp_weighter is not None:                                        # adapter.py:206-208
            return self.gp_weighter.sample_prototypes(max(1, num_samples), visual_embeddings)
        return self.text_embeddings.mean(dim=1)                                 # adapter.py:225 (uniform template weights)

    def forward_features(self, features: torch.Tensor, num_samples: Optional[int] = None) -> torch.Tensor:
        prec = _precision(self.config)
        projected = ops.matmul_nt(features, self.visual_proj.weight, 1.0, prec)  # adapter.py:239
        f_hat = ops.row_normalize(projected)
        scale = self.logit_scale.exp()
        S = self.gp_num_mc_samples_train if self.training else self.gp_num_mc_samples_eval     # adapter.py:242
        return _cosine_logits(f_hat, self.get_prototypes(S, projected), scale, prec)

    forward = forward_features

    def compute_loss(self, features, labels, num_samples: int, gp_beta: float, l2_lambda: float, shots: int):
        """Trainer.compute_loss, adapter.py:387-476."""
        prec = _precision(self.config)
        f_hat = ops.row_normalize(ops.matmul_nt(features, self.visual_proj.weight, 1.0, prec))
        scale = float(self.logit_scale.exp())
        if self.gp_weighter is not None and num_samples > 1:
            protos = self.gp_weighter.sample_prototypes(num_samples)            # adapter.py:404
            S, C, D = protos.shape
            p_hat = ops.row_normalize(protos).reshape(S * C, D)
            logits = ops.matmul_nt(f_hat, p_hat, scale, prec).view(-1, C)       # rows (b, s)
            ce = ops.cross_entropy(logits, labels, rows_per_label=S)            # mean_s mean_b CE (adapter.py:422-428)
        else:
            ce = ops.cross_entropy(_cosine_logits(f_hat, self.get_prototypes(num_samples), scale, prec), labels)
        total = ce
        if self.gp_weighter is not None:
            total = total + self.gp_weighter.variational_strategy.kl_divergence().sum() * float(gp_beta)
        W = self.visual_proj.weight
        if W.requires_grad:
            eye = torch.eye(W.shape[0], device=W.device, dtype=W.dtype)
            total = total + (W - eye).pow(2).sum() * (float(l2_lambda) / shots)
        return total


class TaskResHead(nn.Module):
    """``TaskResLearner`` + ``CustomCLIP.forward`` of trainers/taskres.py on cached features."""

    def __init__(self, config: Any, base_text_features: torch.Tensor, logit_scale: float = math.log(100.0)):
        super().__init__()
        self.config = config
        a = getattr(config, "adapter", config)
        self.alpha = float(getattr(a, "taskres_residual_scale", 0.5))
        self.register_buffer("base_text_features", base_text_features.detach().float().clone())
        self.text_feature_residuals = nn.Parameter(torch.zeros_like(self.base_text_features))
        self.logit_scale = nn.Parameter(torch.tensor(float(logit_scale)), requires_grad=False)
        self.gp_weighter: Optional[GaussianProcessTemplateWeighter] = None
        self.gp_num_mc_samples_train = int(getattr(a, "gp_num_mc_samples_train", 1) or 1)
        self.gp_num_mc_samples_eval = int(getattr(a, "gp_num_mc_samples_eval", 1) or 1)

    def forward(self, image_features: torch.Tensor) -> torch.Tensor:
        f_hat = ops.row_normalize(image_features)                               # taskres.py:99
        scale = self.logit_scale.exp()
        if self.gp_weighter is not None and bool(getattr(getattr(self.config, "adapter", self.config), "use_gp", False)):
            S = max(1, self.gp_num_mc_samples_train if self.training else self.gp_num_mc_samples_eval)
            protos = self.gp_weighter.sample_prototypes(S)                      # taskres.py:107-109
            p_hat = ops.row_normalize(protos)
            text_s = p_hat + (self.alpha * self.text_feature_residuals).unsqueeze(0)       # :111-112
            return _cosine_logits(f_hat, text_s, scale, _precision(self.config))            # :113-116
        return _cosine_logits(f_hat, self.base_text_features + self.alpha * self.text_feature_residuals, scale,
                              _precision(self.config))                          # :119-121


class AdapterMLP(nn.Module):
    """trainers/clip_adapter.py:16-32 (bias-free 2-layer MLP with ReLU); the linears run on the clipgp fp32 GEMM."""

    def __init__(self, in_dim: int, reduction: int = 4, precision: str = "fp32"):
        super().__init__()
        hidden = max(1, in_dim // max(1, int(reduction)))
        self.fc1 = nn.Linear(in_dim, hidden, bias=False)
        self.fc2 = nn.Linear(hidden, in_dim, bias=False)
        self.precision = precision

    def forward(self, x):
        x = torch.relu(ops.matmul_nt(x, self.fc1.weight, 1.0, self.precision))
        return torch.relu(ops.matmul_nt(x, self.fc2.weight, 1.0, self.precision))


class ClipAdapterHead(nn.Module):
    """``CustomCLIP`` of trainers/clip_adapter.py on cached features; clip_weights is [D,K] as in the reference."""

    def __init__(self, config: Any, clip_weights: torch.Tensor, logit_scale: float = math.log(100.0)):
        super().__init__()
        self.config = config
        a = getattr(config, "adapter", config)
        in_dim = int(clip_weights.shape[0])
        self.adapter = AdapterMLP(in_dim, int(getattr(a, "clip_adapter_reduction", 4)), _precision(config))
        self.register_buffer("_blend_ratio", torch.tensor(float(getattr(a, "clip_adapter_ratio", 0.2))))
        self.register_buffer("clip_weights", clip_weights.detach().float().clone())
        self.logit_scale = nn.Parameter(torch.tensor(float(logit_scale)), requires_grad=False)
        self.gp_weighter: Optional[GaussianProcessTemplateWeighter] = None
        self.gp_num_mc_samples_train = int(getattr(a, "gp_num_mc_samples_train", 1) or 1)
        self.gp_num_mc_samples_eval = int(getattr(a, "gp_num_mc_samples_eval", 1) or 1)

    def _apply_adapter(self, feats):
        ratio = float(self._blend_ratio.item())
        return ratio * self.adapter(feats) + (1.0 - ratio) * feats            # clip_adapter.py:77-80

    def logits_from_features(self, features: torch.Tensor, training: bool = False) -> torch.Tensor:
        feats = self._apply_adapter(features.float())
        f_hat = ops.row_normalize(feats)
        scale = self.logit_scale.exp()
        if self.gp_weighter is not None and bool(getattr(getattr(self.config, "adapter", self.config), "use_gp", False)):
            S = max(1, self.gp_num_mc_samples_train if training else self.gp_num_mc_samples_eval)
            return _cosine_logits(f_hat, self.gp_weighter.sample_prototypes(S), scale, _precision(self.config))   # clip_adapter.py:90-96
        return _cosine_logits(f_hat, self.clip_weights.t().contiguous(), scale, _precision(self.config))           # :97-100 (normalize(dim=0) of [D,K])

    def forward(self, features):
        return self.logits_from_features(features, training=self.training)


# ---------------------------------------------------------------------------------------------------- Tip-Adapter
def tip_build_cache(features_hat: torch.Tensor, labels: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """_build_cache (tip_adapter.py:43-50) with the keys sorted by class so that label-segmented sums are contiguous;
    the one-hot value matrix is never formed (the labels are the values)."""
    order = torch.argsort(labels, stable=True)
    return features_hat[order].contiguous(), labels[order].contiguous()


def tip_search(feats_hat, labels, keys, key_labels, clip_logits, num_classes: int, init_beta: float, init_alpha: float,
               betas=(1.0, 2.0, 5.0), alphas=(1.0, 5.0, 10.0, 20.0, 50.0), precision: str = "fp32", chunk: int = 8192):
    """_search_hyperparams (tip_adapter.py:52-80): first (beta, alpha) with the strictly best top-1.  One affinity GEMM is shared
    by all 15 pairs."""
    best_acc, best_beta, best_alpha = -1.0, float(init_beta), float(init_alpha)
    N = feats_hat.shape[0]
    pairs = [(float(b), float(a)) for b in betas for a in alphas]
    hits = torch.zeros(len(pairs), dtype=torch.int64, device=feats_hat.device)
    with torch.no_grad():
        for lo in range(0, N, chunk):                       # image chunks bound the [chunk, N_tr] affinity (0.5 GB at 16 000 keys)
            aff = ops.tip_affinity(feats_hat[lo:lo + chunk], keys, precision)
            y = labels[lo:lo + chunk]
            cl = clip_logits[lo:lo + chunk]
            for i, (beta, alpha) in enumerate(pairs):
                tl = ops.tip_logits_from_affinity(aff, key_labels, cl, beta, alpha, num_classes)
                hits[i] += metrics.calibration_pass(tl, y, 1, want_conf=False)[2][3, 0]
    for (beta, alpha), h in zip(pairs, hits.tolist()):      # one host read; first strictly best pair, as the reference's loop order
        acc = 100.0 * h / max(N, 1)
        if acc > best_acc:
            best_acc, best_beta, best_alpha = acc, beta, alpha
    return best_beta, best_alpha, best_acc


# ---------------------------------------------------------------------------------------------------- GP pre-training
def gp_pretrain(gp_weighter: GaussianProcessTemplateWeighter, feats_hat: torch.Tensor, labels: torch.Tensor, epochs: int,
                gp_lr: float, beta_kl: float, num_samples: int, weight_decay: float = 0.0, scale: float = 100.0, log_every: int = 10,
                precision: str = "auto", seed: int = 0):
    """Full-batch ELBO optimisation of the weighter (taskres.py:254-280; identical in clip_adapter.py:257-279 and
    tip_adapter.py:122-146): CE(mean_s 100 f . normalize(protos_s), y) + beta * sum KL, AdamW(gp_lr) + CosineAnnealingLR(epochs),
    one step per epoch on ALL few-shot features.

    Runs on the fused engine (engine.py) in its collapsed ``logit_mean`` form: the MC mean is taken over unit prototypes, so the
    step is ONE [N_tr, D] x [C, D]^T tcgen05 GEMM forward and one backward instead of the reference's [N_tr, S, C] einsum, with
    the engine's own AdamW and the device-resident cosine rate; no autograd graph, no torch.optim.  The features are unit rows
    already and the visual projection is the frozen identity, exactly the reference's setting.  Returns the loss history."""
    from .engine import EngineConfig, GPAdapterEngine
    if float(weight_decay) != 0.0:
        # the reference's AdamW decays every gp_weighter tensor that has a gradient, including the hook-masked template rows of
        # the inducing points; the engine keeps those rows frozen (K_ZX = K_ZZ[:, :T] aliasing), so decay is not drop-in
        raise NotImplementedError("gp_pretrain: weight_decay > 0 is not supported by the fused engine (it would move the frozen "
                                  "template rows of the inducing points, gp_template_weigher.py:72-79)")
    N = int(feats_hat.shape[0])
    S = max(1, int(num_samples))
    if precision in ("auto", None):
        # TF32 on the fp32 tensors in place when the row pitches allow it (the collapsed logits are [N, C]), else split-bf16
        precision = "tf32" if (gp_weighter.dim % 4 == 0 and gp_weighter.num_classes % 4 == 0) else "bf16x3"
    cfg = EngineConfig(S_train=S, S_eval=S, batch_size=N, logit_scale=float(scale), gp_beta=float(beta_kl), l2_lambda=0.0, shots=1,
                       lr=0.0, gp_lr=float(gp_lr), weight_decay=0.0, loss_mode="logit_mean", train_visual_proj=False,
                       precision=precision, seed=int(seed))
    eng = GPAdapterEngine(gp_weighter, cfg)
    f = feats_hat.detach().float().contiguous()
    y = labels.to(torch.int64).contiguous()
    losses = torch.empty(int(epochs), dtype=torch.float32, device=f.device)
    for ep in range(int(epochs)):
        eng.cosine_lr(ep, int(epochs), 0.0, float(gp_lr))                    # scheduler.step() after every epoch (taskres.py:276)
        losses[ep:ep + 1].copy_(eng.train_step(f, y))
    eng.export_to_module(gp_weighter)
    if getattr(gp_weighter, "rng", None) == "philox" and int(gp_weighter._rng_state[0]) == int(seed):
        gp_weighter._rng_state[1] += int(epochs)                             # the engine consumed steps 0 .. epochs-1 of this stream
    hist = losses.tolist()                                                   # one host read for the whole loop
    if log_every:
        for ep, l in enumerate(hist):
            if ep == 0 or (ep + 1) % log_every == 0:
                print(f"[GP] epoch {ep + 1}/{epochs} loss={l:.4f}")
    eng._graph = None
    return hist


def gp_pretrain_autograd(gp_weighter: GaussianProcessTemplateWeighter, feats_hat: torch.Tensor, labels: torch.Tensor, epochs: int,
                         gp_lr: float, beta_kl: float, num_samples: int, weight_decay: float = 0.0, scale: float = 100.0):
    """The same loop on the autograd surface (ops.py) with torch.optim — the comparator of tools/bench_paths.py and the route for
    weight_decay > 0.  Materialises the [N_tr, S*C] logits exactly like the reference."""
    opt = torch.optim.AdamW(gp_weighter.parameters(), lr=gp_lr, weight_decay=weight_decay)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, epochs)
    hist = []
    for ep in range(epochs):
        gp_weighter.train()
        prot = gp_weighter.sample_prototypes(num_samples=max(1, num_samples))
        loss = ops.cross_entropy(_cosine_logits(feats_hat, prot, scale), labels) + \
            beta_kl * gp_weighter.variational_strategy.kl_divergence().sum()
        opt.zero_grad()
        loss.backward()
        opt.step()
        sched.step()
        hist.append(float(loss.detach()))
    return hist


# ---------------------------------------------------------------------------------------------------------------------
@torch.no_grad()
def template_accuracy_scores(text_embeddings: torch.Tensor, features: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
    """S[k,m] = zero-shot accuracy of template m on the cached features of class k (adapter.py:103-115).

    One tcgen05 GEMM per template ([N,D] x [K,D]^T on split-bf16 operands, fp32-grade products) whose epilogue reduces every logits
    row to its arg-max and the hit flag `pred == label` (the row-statistics epilogue of the eval path): the [N,K] logits of the
    M zero-shot passes never reach HBM.  The logit scale is a positive factor and does not move the arg-max."""
    from . import _lib, tc
    dev = _lib.require_cuda(text_embeddings)
    K, M, D = text_embeddings.shape
    f_hat = ops.row_normalize(features.to(dev).float().contiguous())
    fb = tc.cast_bf16(f_hat, tc.SPLIT_A)
    labels_i64 = labels.to(dev, torch.int64).contiguous()
    counts_k = torch.bincount(labels_i64, minlength=K).to(torch.float32).clamp_min(1)
    scores = torch.zeros(K, M, dtype=torch.float32, device=dev)
    for m in range(M):
        prot = ops.row_normalize(text_embeddings[:, m, :].to(dev).float().contiguous())
        _, correct, _, _ = tc.logits_calibration(fb, tc.cast_bf16(prot, tc.SPLIT_B), 1.0, labels_i64, want_conf=True)
        sums_k = torch.zeros(K, dtype=torch.float32, device=dev)
        sums_k.index_add_(0, labels_i64, correct.to(torch.float32))
        scores[:, m] = sums_k / counts_k
    return scores


@torch.no_grad()
def get_template_weights(config: Any, text_embeddings: torch.Tensor, features: Optional[torch.Tensor], labels: Optional[torch.Tensor],
                         logit_scale=100.0) -> torch.Tensor:
    """Per-class template weights [K,M], rows sum to 1 — `_get_template_weights`, adapter.py:48-142 (the `prefit_on_full_set`
    image pipeline is outside the cached-feature path).  Methods (config.adapter.template_init_method): "uniform",
    "val_weighted", "top3", "minmax"."""
    adapter_cfg = getattr(config, "adapter", config)
    method = str(getattr(adapter_cfg, "template_init_method", "uniform")).lower()
    E = text_embeddings
    K, M = int(E.shape[0]), int(E.shape[1])
    if M == 0:
        return torch.empty(K, 0, device=E.device, dtype=E.dtype)
    if method == "uniform" or features is None or labels is None:
        return torch.full((K, M), 1.0 / float(M), device=E.device, dtype=E.dtype)
    scores = template_accuracy_scores(E, features, labels)
    if method == "top3":
        top_k = min(3, M)
        _, top_idx = torch.topk(scores.mean(dim=0), k=top_k, largest=True)
        keep = torch.zeros(M, dtype=scores.dtype, device=scores.device)
        keep[top_idx] = 1.0
        scores = scores * keep.view(1, -1)
        zero_rows = scores.sum(dim=1) <= 1e-12
        if bool(zero_rows.any()):
            scores[zero_rows] = (keep / float(top_k)).view(1, -1).expand(int(zero_rows.sum().item()), -1)
    elif method == "minmax":
        s_min = scores.min(dim=1, keepdim=True).values
        s_max = scores.max(dim=1, keepdim=True).values
        rng = s_max - s_min
        scores = torch.where(rng.le(1e-12), torch.full_like(scores, 1.0 / float(M)), (scores - s_min) / rng.clamp_min(1e-12))
    return torch.softmax(torch.log(scores.clamp_min(1e-12)), dim=1).to(device=E.device, dtype=E.dtype)
