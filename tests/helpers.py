"""Shared builders for the parity tests (oracle state <-> clipgp module / kernel arguments)."""
import copy

import torch

from clip_gp_b200 import synth
from oracle import gp as ogp


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| / max |b|  (the 1e-3 gate of BASELINE.json is relative to the tensor's scale)."""
    a = a.detach().double().cpu(); b = b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def make_state(name: str, kernel: str, trained: bool = True, seed: int = 5, perturb: bool = True):
    """Oracle GPState for a named synthetic workload with a non-trivial q(u) and hyper-parameters."""
    wl = synth.make_workload(name)
    shp = wl["shape"]
    st = ogp.build_state(wl["E"], kernel, shp.d)
    if trained:
        m, Lq = synth.trained_like_q(shp.C, shp.T + 1, seed)
        st.var_mean, st.chol_var = m.clone(), Lq.clone()
    if perturb:
        g = torch.Generator().manual_seed(seed + 3)
        kp = st.kernel
        if kp.raw_lengthscale is not None:
            kp.raw_lengthscale = kp.raw_lengthscale + 0.1 * torch.randn(kp.raw_lengthscale.shape, generator=g)
        if kp.raw_outputscale is not None:
            kp.raw_outputscale = kp.raw_outputscale + 0.1 * torch.randn(kp.raw_outputscale.shape, generator=g)
        if kp.raw_variance is not None:
            kp.raw_variance = kp.raw_variance + 0.1 * torch.randn(kp.raw_variance.shape, generator=g)
        st.inducing_points = st.inducing_points.clone()
        st.inducing_points[:, -1] += 0.05 * torch.randn(shp.C, st.inducing_points.shape[-1], generator=g)
    return wl, st


def state_to(st, dtype=None, device=None):
    s2 = copy.deepcopy(st)
    for k in ["templates", "templates_red", "inducing_points", "var_mean", "chol_var", "f0", "cls_bias", "tmp_bias",
              "pca_mean", "pca_W"]:
        setattr(s2, k, getattr(s2, k).to(dtype=dtype, device=device))
    for k in ["raw_lengthscale", "raw_outputscale", "raw_variance"]:
        v = getattr(s2.kernel, k)
        if v is not None:
            setattr(s2.kernel, k, v.to(dtype=dtype, device=device))
    return s2


def oracle_grads(st, eps, dw, dkl, dtype=torch.float64):
    """Autograd through oracle.gp in `dtype` for loss = <w, dw> + <kl, dkl>.  Returns (w, kl, grads dict)."""
    s2 = state_to(st, dtype=dtype)
    names = {"Z": s2.inducing_points, "m": s2.var_mean, "chol": s2.chol_var}
    if s2.kernel.raw_lengthscale is not None: names["ls"] = s2.kernel.raw_lengthscale
    if s2.kernel.raw_outputscale is not None: names["os"] = s2.kernel.raw_outputscale
    if s2.kernel.raw_variance is not None: names["var"] = s2.kernel.raw_variance
    for p in names.values():
        p.requires_grad_(True)
    w, aux = ogp.gp_weights(s2, eps.to(dtype))
    kl = ogp.kl_divergence(s2.var_mean, s2.chol_var)
    loss = (w * dw.to(dtype)).sum() + (kl * dkl.to(dtype)).sum()
    g = torch.autograd.grad(loss, list(names.values()))
    return w.detach(), kl.detach(), dict(zip(names.keys(), [x.detach() for x in g])), aux
