"""CPU oracle: cosine-logit heads, losses and the Tip-Adapter cache affinity
(TEST INFRASTRUCTURE, see oracle/__init__.py).  Paths relative to /root/reference.

All functions are differentiable torch; autograd through them is the gradient oracle.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F


# ---------------------------------------------------------------- Adapter (trainers/adapter.py)
def adapter_logits(features, visual_proj_w, prototypes, scale):
    """CustomCLIP.forward_features, adapter.py:230-252.

    features [B,D]; visual_proj_w [D,D] (nn.Linear weight, y = x W^T); prototypes [S,C,D]
    (GP samples, un-normalised) or [C,D]; scale = logit_scale.exp().
    GP branch: mean over s of scale * f_hat . p_hat_s   (adapter.py:247-249).
    """
    projected = features @ visual_proj_w.t()
    f_hat = F.normalize(projected, p=2, dim=-1)
    p_hat = F.normalize(prototypes, p=2, dim=-1)
    if prototypes.dim() == 3:
        return (scale * torch.einsum("bd,skd->bsk", f_hat, p_hat)).mean(dim=1)
    return scale * (f_hat @ p_hat.t())


def adapter_mc_ce(features, labels, visual_proj_w, prototypes, scale):
    """Trainer.compute_loss GP branch, adapter.py:401-428: mean over s of CE(scale f_hat p_hat_s^T, y)."""
    f_hat = F.normalize(features @ visual_proj_w.t(), p=2, dim=-1)
    ce = []
    for s in range(prototypes.shape[0]):
        p_hat = F.normalize(prototypes[s], p=2, dim=-1)
        ce.append(F.cross_entropy(scale * (f_hat @ p_hat.t()), labels))
    return torch.stack(ce, dim=0).mean()


def adapter_total_loss(ce, kl_per_class, gp_beta, visual_proj_w, l2_lambda, shots, w_trainable=True):
    """adapter.py:453-476: CE + gp_beta * sum_c KL_c + (l2_lambda/shots) * ||W - I||_F^2."""
    total = ce
    if kl_per_class is not None:
        total = total + kl_per_class.sum() * float(gp_beta)
    if w_trainable:
        eye = torch.eye(visual_proj_w.shape[0], dtype=visual_proj_w.dtype, device=visual_proj_w.device)
        total = total + (visual_proj_w - eye).pow(2).sum() * (float(l2_lambda) / shots)
    return total


# ---------------------------------------------------------------- TaskRes (trainers/taskres.py)
def taskres_logits(features, base_text_features, residuals, alpha, scale,
                   prototypes: Optional[torch.Tensor] = None):
    """CustomCLIP.forward on already-encoded features, taskres.py:96-123 (+ :45-47).

    non-GP: normalize(base + alpha x); GP: p_hat_s = normalize(protos_s); t_s = normalize(p_hat_s + alpha x);
    logits = mean_s scale f_hat . t_s.
    """
    f_hat = F.normalize(features, p=2, dim=-1)
    if prototypes is not None:
        p = prototypes / prototypes.norm(dim=-1, keepdim=True)
        t = p + (alpha * residuals).unsqueeze(0)
        t = t / t.norm(dim=-1, keepdim=True)
        return (scale * torch.einsum("bd,skd->bsk", f_hat, t)).mean(dim=1)
    t = F.normalize(base_text_features + alpha * residuals, p=2, dim=-1)
    return scale * (f_hat @ t.t())


def gp_pretrain_loss(features_hat, labels, prototypes, kl_per_class, beta_kl, scale=100.0):
    """Full-batch ELBO step of taskres.py:261-272 == clip_adapter.py:264-275 == tip_adapter.py:130-142:
    CE(mean_s 100 f . normalize(protos_s), y) + beta * sum KL.  features are already unit-norm."""
    prot = prototypes / prototypes.norm(dim=-1, keepdim=True)
    logits = (scale * torch.einsum("bd,skd->bsk", features_hat, prot)).mean(dim=1)
    return F.cross_entropy(logits, labels) + beta_kl * kl_per_class.sum(), logits


def gp_mean_prototypes(prototypes):
    """taskres.py:281-285 / clip_adapter.py:284-288 / tip_adapter.py:152-156: normalize(mean_s protos_s)."""
    p = prototypes.mean(dim=0)
    return p / p.norm(dim=-1, keepdim=True)


# ---------------------------------------------------------------- CLIP-Adapter (trainers/clip_adapter.py)
def clip_adapter_features(feats, fc1_w, fc2_w, ratio):
    """AdapterMLP + blend, clip_adapter.py:16-32, 77-80."""
    g = F.relu(F.relu(feats @ fc1_w.t()) @ fc2_w.t())
    return ratio * g + (1.0 - ratio) * feats


def clip_adapter_logits(feats_adapted, scale, clip_weights=None, prototypes=None):
    """_compute_logits_from_embeddings, clip_adapter.py:85-100. clip_weights is [D,C]."""
    f_hat = F.normalize(feats_adapted, p=2, dim=-1)
    if prototypes is not None:
        p = prototypes / prototypes.norm(dim=-1, keepdim=True)
        return (scale * torch.einsum("bd,skd->bsk", f_hat, p)).mean(dim=1)
    return scale * (f_hat @ F.normalize(clip_weights, p=2, dim=0))


# ---------------------------------------------------------------- Tip-Adapter (trainers/tip_adapter.py)
def tip_cache_vals(labels, num_classes, dtype=torch.float32):
    """_build_cache, tip_adapter.py:43-50: one-hot values."""
    vals = torch.zeros(labels.shape[0], num_classes, dtype=dtype, device=labels.device)
    vals.scatter_(1, labels.view(-1, 1).long(), 1.0)
    return vals


def tip_logits(feats_hat, keys, cache_vals, clip_logits, beta, alpha):
    """tip_adapter.py:250-260 (also :69-74, :281-290, :309-318, :331-333, :371-383):
    affinity = f keys^T (nn.Linear(D,N_tr) initialised with keys);
    cache = exp(-(beta - beta*affinity)) @ vals;  tip = clip_logits + alpha * cache."""
    affinity = feats_hat @ keys.t()
    cache_logits = ((-1.0) * (beta - beta * affinity)).exp() @ cache_vals
    return clip_logits + cache_logits * alpha


def tip_search(feats_hat, labels, keys, cache_vals, clip_logits, init_beta, init_alpha,
               betas=(1.0, 2.0, 5.0), alphas=(1.0, 5.0, 10.0, 20.0, 50.0)):
    """_search_hyperparams, tip_adapter.py:52-80: first (beta, alpha) with strictly best top-1."""
    best_acc, best_beta, best_alpha = -1.0, float(init_beta), float(init_alpha)
    affinity = feats_hat @ keys.t()
    for beta in betas:
        cache_logits = ((-1.0) * (beta - beta * affinity)).exp() @ cache_vals
        for alpha in alphas:
            tl = clip_logits + cache_logits * alpha
            acc = float(tl.argmax(1).eq(labels).float().sum().mul(100.0 / labels.numel()).item())
            if acc > best_acc:
                best_acc, best_beta, best_alpha = acc, float(beta), float(alpha)
    return best_beta, best_alpha, best_acc


# ---------------------------------------------------------------- template-weight initialisation (trainers/adapter.py)
@torch.no_grad()
def template_weights(method, text_embeddings, features, labels, logit_scale, temperature: float = 1.0):
    """_get_template_weights, adapter.py:48-142 (without the prefit_on_full_set image pipeline): per-class template weights [K,M].

    zero-shot accuracy of every template per class (:107-115) -> optional top3 (:116-129) / minmax (:130-137) transform ->
    softmax(log(clamp(S, 1e-12)) / temperature) (:138-139).  "uniform" (or missing features) returns 1/M (:99-100)."""
    E = text_embeddings
    K, M = int(E.shape[0]), int(E.shape[1])
    method = str(method).lower()
    if method == "uniform" or features is None or labels is None:
        return torch.full((K, M), 1.0 / float(M), dtype=E.dtype)
    feats = F.normalize(features, p=2, dim=-1)
    labels = labels.to(torch.int64)
    counts_k = torch.bincount(labels, minlength=K).to(feats.dtype).clamp_min(1)
    scores = torch.zeros(K, M, dtype=feats.dtype)
    for m in range(M):
        prot_m = F.normalize(E[:, m, :], p=2, dim=-1)
        preds = (logit_scale * (feats @ prot_m.t())).argmax(dim=1)
        corr = (preds == labels).to(feats.dtype)
        sums_k = torch.zeros(K, dtype=feats.dtype)
        sums_k.index_add_(0, labels, corr)
        scores[:, m] = sums_k / counts_k
    if method == "top3":
        top_k = min(3, M)
        _, top_idx = torch.topk(scores.mean(dim=0), k=top_k, largest=True)
        keep = torch.zeros(M, dtype=scores.dtype)
        keep[top_idx] = 1.0
        scores = scores * keep.view(1, -1)
        zero_rows = scores.sum(dim=1) <= 1e-12
        if bool(zero_rows.any()):
            scores[zero_rows] = (keep / float(top_k)).view(1, -1).expand(int(zero_rows.sum()), -1)
    elif method == "minmax":
        s_min = scores.min(dim=1, keepdim=True).values
        s_max = scores.max(dim=1, keepdim=True).values
        rng = s_max - s_min
        scores = torch.where(rng.le(1e-12), torch.full_like(scores, 1.0 / float(M)), (scores - s_min) / rng.clamp_min(1e-12))
    return torch.softmax(torch.log(scores.clamp_min(1e-12)) / max(temperature, 1e-6), dim=1), scores
