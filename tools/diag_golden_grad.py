"""Diagnostic: error of the CUDA gradients and of the reference-executed golden gradients against float64 (per golden case)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.test_gpu_ref_golden import module_from_golden, T_
from tests.test_ref_golden import state_from_golden, CASES, KERNELS
from tests.helpers import state_to, max_err
from oracle import gp as ogp
G = np.load("tests/golden/ref_gp.npz")
for case in CASES:
    for kernel in KERNELS:
        key = f"{case}/{kernel}"
        gp = module_from_golden(G, case, kernel)
        eps = T_(G, f"{key}/eps")
        protos = gp.sample_prototypes(eps.shape[2], eps=eps.cuda())
        kl = gp.variational_strategy.kl_divergence()
        ((protos * T_(G, f"{key}/dP").cuda()).sum() + (kl * T_(G, f"{key}/dkl").cuda()).sum()).backward()
        s64 = state_to(state_from_golden(G, case, kernel), dtype=torch.float64)
        ps = {"Z": s64.inducing_points, "m": s64.var_mean, "chol": s64.chol_var}
        for p in ps.values(): p.requires_grad_(True)
        pr, _ = ogp.sample_prototypes(s64, eps.double())
        l64 = (pr * T_(G, f"{key}/dP").double()).sum() + (ogp.kl_divergence(s64.var_mean, s64.chol_var) * T_(G, f"{key}/dkl").double()).sum()
        g64 = dict(zip(ps, torch.autograd.grad(l64, list(ps.values()))))
        q = gp.variational_strategy._variational_distribution
        got = {"Z": gp.variational_strategy.inducing_points.grad[:, -1], "m": q.variational_mean.grad, "chol": q.chol_variational_covar.grad}
        line = key
        for n in ("Z", "m", "chol"):
            t = g64[n][:, -1] if n == "Z" else g64[n]
            r = T_(G, f"{key}/grad/{n}"); r = r[:, -1] if n == "Z" else r
            line += f" | d{n}: cuda {max_err(got[n], t):.1e} ref {max_err(r, t):.1e}"
        print(line)
