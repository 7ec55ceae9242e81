"""Debug aid: aliased vs un-aliased general-path adjoint of the learnable inducing row for n > 33 against the float64 oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tests.helpers import make_state, oracle_grad_pair
from tests.test_gpu_gp import run_kernel

for name in sys.argv[1:] or ["t40"]:
    for kernel in ("rbf", "matern", "linear"):
        wl, st = make_state(name, kernel)
        shp = wl["shape"]
        g = torch.Generator().manual_seed(11)
        eps = torch.randn(shp.C, shp.T, shp.S, generator=g)
        dw = torch.randn(shp.S, shp.C, shp.T, generator=g)
        dkl = torch.rand(shp.C, generator=g)
        G32, G = oracle_grad_pair(st, eps, dw, dkl)
        out = {}
        for alias in (True, False):
            w, kl, _, P = run_kernel(st, eps, kernel, alias_check=alias, need_grad=True)
            ((w * dw.cuda()).sum() + (kl * dkl.cuda()).sum()).backward()
            out[alias] = {k: v.grad.detach().cpu().double() for k, v in P.items() if v is not None and v.grad is not None}
        t = G["Z"][:, -1]
        for alias in (True, False):
            x = out[alias]["Z"][:, -1]
            per_class = ((x - t).abs().amax(1) / t.abs().amax()).tolist()
            print(name, kernel, "alias" if alias else "general", "dZ_last max|x-t|/max|t| per class:", " ".join(f"{v:.1e}" for v in per_class),
                  "| ref32:", f"{float((G32['Z'][:, -1].double() - t).abs().max() / t.abs().max()):.1e}")
        for k in ("ls", "os", "var", "m", "chol"):
            if k in out[True] and k in G:
                print("      ", k, " ".join(f"{float((out[a][k] - G[k]).abs().max() / G[k].abs().max()):.1e}" for a in (True, False)))
