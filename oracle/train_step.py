"""CPU oracle: one GP-Adapter optimisation step exactly as the engine defines it (TEST INFRASTRUCTURE;
also the `cpu_baseline` / `--impl reference` leg of bench.py).

Restates trainers/adapter.py:387-476 (compute_loss) + :537-549 (backward, optimizer step) with the oracle GP,
torch autograd and torch.optim.AdamW (utils/optimization.py builds AdamW for OPTIM.NAME "adamw", two parameter
groups as adapter.py:298-309).
"""
from __future__ import annotations

import torch

from . import gp as ogp
from . import heads


def state_to_device(st: ogp.GPState, device) -> ogp.GPState:
    """Copy of the oracle state on `device` (bench.py's torch-eager GPU comparator: the same restated reference step, run by
    torch on the GPU as the reference itself would)."""
    import copy
    s2 = copy.deepcopy(st)
    for k in ("templates", "templates_red", "inducing_points", "var_mean", "chol_var", "f0", "cls_bias", "tmp_bias", "pca_mean", "pca_W"):
        setattr(s2, k, getattr(s2, k).detach().to(device))
    for k in ("raw_lengthscale", "raw_outputscale", "raw_variance"):
        v = getattr(s2.kernel, k)
        if v is not None:
            setattr(s2.kernel, k, v.detach().to(device))
    return s2


class OracleAdapter:
    def __init__(self, st: ogp.GPState, D: int, scale=100.0, gp_beta=0.01, l2_lambda=0.5, shots=16, lr=0.01, gp_lr=1e-3,
                 weight_decay=0.0, loss_mode="per_sample"):
        self.st = st
        self.W = torch.eye(D, dtype=st.templates.dtype, device=st.templates.device, requires_grad=True)
        self.scale, self.gp_beta, self.l2_lambda, self.shots = scale, gp_beta, l2_lambda, shots
        self.loss_mode = loss_mode
        self.gp_params = [st.inducing_points, st.var_mean, st.chol_var]
        for p in (st.kernel.raw_lengthscale, st.kernel.raw_outputscale, st.kernel.raw_variance):
            if p is not None:
                self.gp_params.append(p)
        for p in self.gp_params:
            p.requires_grad_(True)
        self.mask = torch.zeros_like(st.inducing_points)
        self.mask[:, -1, :] = 1.0                                           # gp_template_weigher.py:72-79
        self.opt = torch.optim.AdamW([{"params": [self.W], "lr": lr, "weight_decay": weight_decay},
                                      {"params": self.gp_params, "lr": gp_lr, "weight_decay": weight_decay}])

    def loss(self, feats, labels, eps):
        feats, eps = feats.to(self.W.dtype), eps.to(self.W.dtype)          # float64 twin: state_to(st, torch.float64)
        protos, aux = ogp.sample_prototypes(self.st, eps)
        kl = ogp.kl_divergence(self.st.var_mean, self.st.chol_var)
        if self.loss_mode == "per_sample":
            ce = heads.adapter_mc_ce(feats, labels, self.W, protos, self.scale)
        else:
            ce = torch.nn.functional.cross_entropy(heads.adapter_logits(feats, self.W, protos, self.scale), labels)
        return heads.adapter_total_loss(ce, kl, self.gp_beta, self.W, self.l2_lambda, self.shots)

    def step(self, feats, labels, eps):
        self.opt.zero_grad(set_to_none=True)
        loss = self.loss(feats, labels, eps)
        loss.backward()
        self.st.inducing_points.grad.mul_(self.mask)
        self.opt.step()
        return float(loss.detach())

    @torch.no_grad()
    def eval_logits(self, feats, eps):
        feats, eps = feats.to(self.W.dtype), eps.to(self.W.dtype)
        protos, _ = ogp.sample_prototypes(self.st, eps)
        return heads.adapter_logits(feats, self.W, protos, self.scale)
