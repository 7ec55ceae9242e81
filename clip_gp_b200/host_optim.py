"""Host-side optimizer / scheduler selection for the small non-GP parameter sets (TaskRes residuals [C,D], the CLIP-Adapter MLP,
the baseline visual projection): same names, defaults and error behaviour as the reference's builders
(utils/optimization.py:57-238).  These stay torch optimizers (host code); the GP path has its own fused
AdamW kernel in the engine."""
from __future__ import annotations

from typing import Any, Iterable

import torch


def _get(cfg: Any, name: str, default):
    return getattr(cfg, name, default) if cfg is not None else default


def build_optimizer(params_or_groups: Iterable, optim_cfg: Any, name: str | None = None, lr: float | None = None) -> torch.optim.Optimizer:
    """utils/optimization.py:57-98 / :147-170 (sgd | adam | adamw; `muon` needs a package that is not in this image)."""
    name = str(name if name is not None else _get(optim_cfg, "name", "sgd")).lower()
    lr = float(lr if lr is not None else _get(optim_cfg, "lr", 1e-3))
    wd = float(_get(optim_cfg, "weight_decay", 0.0))
    betas = tuple(_get(optim_cfg, "betas", (0.9, 0.999)))
    eps = float(_get(optim_cfg, "eps", 1e-8))
    if name == "sgd":
        return torch.optim.SGD(params_or_groups, lr=lr, weight_decay=wd, momentum=float(_get(optim_cfg, "momentum", 0.9)),
                               nesterov=bool(_get(optim_cfg, "nesterov", False)))
    if name == "adam":
        return torch.optim.Adam(params_or_groups, lr=lr, weight_decay=wd, betas=betas, eps=eps)
    if name == "adamw":
        return torch.optim.AdamW(params_or_groups, lr=lr, weight_decay=wd, betas=betas, eps=eps)
    raise ValueError(f"Unsupported optimizer: {name}")


def build_lr_scheduler(optimizer: torch.optim.Optimizer, optim_cfg: Any, max_epoch: int | None = None):
    """utils/optimization.py:218-270; stepped once per epoch by the trainers (utils/trainer.py:466-470)."""
    S = torch.optim.lr_scheduler
    name = str(_get(optim_cfg, "lr_scheduler", "constant")).lower()
    max_epoch = int(max_epoch if max_epoch is not None else _get(optim_cfg, "max_epoch", 1))
    if name == "cosine":
        return S.CosineAnnealingLR(optimizer, T_max=max_epoch, eta_min=float(_get(optim_cfg, "eta_min", 0.0)))
    if name == "step":
        return S.StepLR(optimizer, step_size=int(_get(optim_cfg, "step_size", max_epoch // 3)), gamma=float(_get(optim_cfg, "gamma", 0.1)))
    if name == "multistep":
        return S.MultiStepLR(optimizer, milestones=list(_get(optim_cfg, "milestones", [max_epoch // 2, max_epoch * 3 // 4])),
                             gamma=float(_get(optim_cfg, "gamma", 0.1)))
    if name == "exponential":
        return S.ExponentialLR(optimizer, gamma=float(_get(optim_cfg, "gamma", 0.95)))
    if name == "constant":
        return S.ConstantLR(optimizer, factor=1.0)
    if name == "linear":
        return S.LinearLR(optimizer, start_factor=float(_get(optim_cfg, "start_factor", 1.0)), end_factor=float(_get(optim_cfg, "end_factor", 0.0)),
                          total_iters=int(_get(optim_cfg, "total_iters", max_epoch)))
    raise ValueError(f"Unsupported scheduler: {name}")
