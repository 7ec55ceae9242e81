"""Golden vectors produced by EXECUTING THE REFERENCE'S OWN FILES (build container only).

    python tests/golden/make_ref_golden.py            # writes tests/golden/ref_gp.npz, ref_heads.npz, ref_train.npz

``/root/reference/trainers/{gp_template_weigher,adapter,taskres,clip_adapter,tip_adapter}.py`` and
``utils/{trainer,optimization,config,metrics}.py`` are imported UNMODIFIED (tests/golden/_ref_env.py) on top of the
minimal gpytorch / linear_operator / entmax stand-ins of ``oracle/_shim`` (those libraries cannot be installed offline;
the shim restates the ~20 library routines the reference reaches, see oracle/_shim/README.md).  Everything the
reference tree itself holds on the hot path is therefore executed, not restated: PCA, f0 prior, mean-module tail, the
``[:, :, :N_templates]`` slice, the ``batch == K`` branch, the prototype einsum, ``compute_loss``, the TaskRes /
CLIP-Adapter / Tip-Adapter heads, ``_get_template_weights``, ``_build_cache``, ``_search_hyperparams``, the optimizer /
scheduler builders and (with a stand-in CLIP that returns cached features) the whole ``Trainer.train()`` of Tip-Adapter,
TaskRes and CLIP-Adapter including their GP pre-training loops.

The fixtures pin (tests/test_ref_golden.py, CPU) ``oracle/*.py`` and (tests/test_gpu_ref_golden.py, GPU) the CUDA path.
The GPU box has no /root/reference: only the .npz files travel.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
import _ref_env  # noqa: E402

_ref_env.setup()
sys.path.append(ROOT)                      # after the reference: its `utils` / `datasets` packages must win

import gpytorch  # noqa: E402  (the shim)
from gpytorch import distributions as shim_dist  # noqa: E402

from clip_gp_b200 import synth  # noqa: E402  (synthetic text bank / features only)
from oracle import philox  # noqa: E402

ref_gpw = _ref_env.ref_module("trainers.gp_template_weigher")
ref_cfg = _ref_env.ref_module("utils.config")


def np_(t):
    return t.detach().cpu().numpy().copy()


def make_config(**adapter):
    cfg = ref_cfg.Config()
    cfg.use_cuda = False
    for k, v in adapter.items():
        setattr(cfg.adapter, k, v)
    return cfg


class eps_hook:
    """Feed explicit base noise [C, Nx, S] to every rsample inside the block (records what was requested)."""

    def __init__(self, provider):
        self.provider, self.calls = provider, []

    def __enter__(self):
        def hook(shape, dtype, device):
            e = self.provider(tuple(shape), len(self.calls))
            self.calls.append(tuple(shape))
            assert tuple(e.shape) == tuple(shape), (e.shape, shape)
            return e.to(dtype=dtype, device=device)
        shim_dist.BASE_SAMPLES_HOOK = hook
        return self

    def __exit__(self, *a):
        shim_dist.BASE_SAMPLES_HOOK = None


def perturb_gp(gp, seed, bias=True):
    """A trained-like state for every learnable tensor of the reference module (after its first-call initialisation)."""
    g = torch.Generator().manual_seed(seed)
    vs = gp.variational_strategy
    q = vs._variational_distribution
    C, n = q.variational_mean.shape
    with torch.no_grad():
        q.variational_mean.copy_(0.5 * torch.randn(C, n, generator=g))
        q.chol_variational_covar.copy_(torch.eye(n).repeat(C, 1, 1) + 0.1 * torch.randn(C, n, n, generator=g))  # NOT tril: masked inside
        vs.inducing_points[:, -1] += 0.05 * torch.randn(C, vs.inducing_points.shape[-1], generator=g)
        for name, p in gp.covar_module.named_parameters():
            p.add_(0.1 * torch.randn(p.shape, generator=g))
        if bias:
            gp.mean_module.cls_bias.copy_(0.3 * torch.randn(gp.mean_module.cls_bias.shape, generator=g))
            gp.mean_module.tmp_bias.copy_(0.3 * torch.randn(gp.mean_module.tmp_bias.shape, generator=g))


def gp_param_dict(gp):
    out = {"Z": gp.variational_strategy.inducing_points,
           "m": gp.variational_strategy._variational_distribution.variational_mean,
           "chol": gp.variational_strategy._variational_distribution.chol_variational_covar,
           "cls_bias": gp.mean_module.cls_bias, "tmp_bias": gp.mean_module.tmp_bias}
    for name, p in gp.covar_module.named_parameters():
        out[name.split(".")[-1]] = p                 # raw_lengthscale / raw_outputscale / raw_variance
    return out


# ------------------------------------------------------------------------------------------------------------------ GP
GP_CASES = {
    # name: (C, T, D, pca_dim, S)
    "tiny": (12, 5, 64, 16, 3),
    "t32": (6, 32, 96, 48, 4),        # the headline per-class shape n = 33
    "t1": (5, 1, 32, 8, 2),           # single template (n = 2)
    "lowrank": (3, 4, 40, 256, 3),    # gp_pca_dim > rank: red_dim = min(256, C*T) = 12  (gp_template_weigher.py:33-34)
}


def gp_goldens():
    out = {}
    for cname, (C, T, D, pd, S) in GP_CASES.items():
        E, mu_c = synth.make_text_bank(C, T, D, seed=4242 + C)
        for kern in ("rbf", "matern", "linear"):
            key = f"{cname}/{kern}"
            cfg = make_config(gp_pca_dim=pd, gp_kernel_type=kern, gp_prior_temp=1.0)
            torch.manual_seed(99)
            gp = ref_gpw.GaussianProcessTemplateWeighter(text_embeddings=E, cfg=cfg)
            gp.train()
            out[f"{cname}/E"] = np_(E)
            out[f"{key}/pca_mean"] = np_(gp._pca_mean); out[f"{key}/pca_W"] = np_(gp._pca_W)
            out[f"{key}/templates_red"] = np_(gp._templates_red)
            out[f"{key}/Z0"] = np_(gp.variational_strategy.inducing_points)
            out[f"{key}/f0"] = np_(gp.mean_module.f0)
            out[f"{key}/cls_mean_init"] = np_(gp._cls_mean_init)
            if kern == "rbf":
                out[f"{key}/raw_lengthscale0"] = np_(gp.covar_module.base_kernel.raw_lengthscale)
                out[f"{key}/lengthscale0"] = np_(gp.covar_module.base_kernel.lengthscale)
            # first call: gpytorch initialises q(u) (consumes torch RNG: randn_like [C,n]) BEFORE the base noise is drawn
            torch.manual_seed(7)
            p_first = gp.sample_prototypes(S)
            out[f"{key}/first_call/m"] = np_(gp.variational_strategy._variational_distribution.variational_mean)
            out[f"{key}/first_call/chol"] = np_(gp.variational_strategy._variational_distribution.chol_variational_covar)
            out[f"{key}/first_call/w"] = np_(gp.scores); out[f"{key}/first_call/protos"] = np_(p_first)
            torch.manual_seed(7)
            m_chk = 1e-3 * torch.randn(C, T + 1); eps_chk = torch.randn(C, T, S)
            assert torch.equal(m_chk, gp.variational_strategy._variational_distribution.variational_mean.detach())
            out[f"{key}/first_call/eps"] = np_(eps_chk)

            perturb_gp(gp, seed=31 + T)
            params = gp_param_dict(gp)
            for k, p in params.items():
                out[f"{key}/param/{k}"] = np_(p)
            # (1) standard call: Nx = T
            eps = torch.randn(C, T, S, generator=torch.Generator().manual_seed(555))
            for p in gp.parameters():
                p.grad = None
            with eps_hook(lambda shape, i: eps):
                protos = gp.sample_prototypes(S)
            w = gp.scores
            kl = gp.variational_strategy.kl_divergence()
            qf = gp(gp._templates_red)                              # the predictive the samples came from (no RNG)
            g = torch.Generator().manual_seed(777)
            dP = torch.randn(protos.shape, generator=g); dkl = torch.randn(kl.shape, generator=g)
            ((protos * dP).sum() + (kl * dkl).sum()).backward()
            out[f"{key}/eps"] = np_(eps); out[f"{key}/dP"] = np_(dP); out[f"{key}/dkl"] = np_(dkl)
            out[f"{key}/w"] = np_(w); out[f"{key}/protos"] = np_(protos); out[f"{key}/kl"] = np_(kl)
            out[f"{key}/mu"] = np_(qf.mean); out[f"{key}/Sigma"] = np_(qf.covariance_matrix)
            for k, p in params.items():
                out[f"{key}/grad/{k}"] = np_(p.grad if p.grad is not None else torch.zeros_like(p))
            # (2) the `visual_embeddings.shape[0] == num_classes` branch (gp_template_weigher.py:198-203): Nx = T + 1
            vis = synth.make_features(mu_c, torch.arange(C), seed=5, noise=1.0)
            eps1 = torch.randn(C, T + 1, S, generator=torch.Generator().manual_seed(556))
            with eps_hook(lambda shape, i: eps1) as h, torch.no_grad():
                protos_v = gp.sample_prototypes(S, visual_embeddings=vis)
                assert h.calls == [(C, T + 1, S)]
            out[f"{key}/vis/features"] = np_(vis); out[f"{key}/vis/eps"] = np_(eps1)
            out[f"{key}/vis/w"] = np_(gp.scores); out[f"{key}/vis/protos"] = np_(protos_v)
            # a visual batch whose size differs from K is ignored (:204-210)
            with eps_hook(lambda shape, i: eps) as h, torch.no_grad():
                protos_o = gp.sample_prototypes(S, visual_embeddings=torch.cat([vis, vis])[: C + 1])
                assert h.calls == [(C, T, S)]
            # under no_grad gpytorch's sq_dist takes its `x1_eq_x2 and not requires_grad` branch (diagonal forced to 0): the
            # result differs from the grad-mode call by fp32 rounding only
            out[f"{key}/nograd/protos"] = np_(protos_o); out[f"{key}/nograd/w"] = np_(gp.scores)
            print(f"  {key}: no_grad vs grad-mode prototypes max |diff| = {float((protos_o - protos.detach()).abs().max()):.2e}")
            # (3) eval mode re-uses the memoised chol(K_ZZ) of the last call; from a fresh cache it equals the train-mode result
            gp.eval(); gp.variational_strategy._clear_cache()
            with eps_hook(lambda shape, i: eps), torch.no_grad():
                protos_e = gp.sample_prototypes(S)
            out[f"{key}/eval_mode/protos"] = np_(protos_e)
            # (4) initialize_from_weights is a no-op for T > 1 (shape mismatch swallowed) and a broadcast copy for T == 1
            m_before = gp.variational_strategy._variational_distribution.variational_mean.detach().clone()
            gp.initialize_from_weights(torch.full((C, T), 1.0 / T))
            out[f"{key}/init_from_weights/m_after"] = np_(gp.variational_strategy._variational_distribution.variational_mean)
            out[f"{key}/init_from_weights/changed"] = np.array(
                not torch.equal(m_before, gp.variational_strategy._variational_distribution.variational_mean.detach()))
            out[f"{key}/state_dict_keys"] = np.array(sorted(gp.state_dict().keys()))
    return out


def main():
    torch.set_num_threads(4)
    out = gp_goldens()
    np.savez_compressed(os.path.join(HERE, "ref_gp.npz"), **out)
    print("wrote ref_gp.npz:", len(out), "arrays,", os.path.getsize(os.path.join(HERE, "ref_gp.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
