"""How sparse are the sparsemax template weights of the bench state?  (rows of E[c] a class actually needs per step)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch, bench
from clip_gp_b200 import synth
from clip_gp_b200.engine import EngineConfig, GPAdapterEngine
from clip_gp_b200.gp_template_weigher import GaussianProcessTemplateWeighter
wl = synth.make_workload("cfg2", n_test=256); shp = wl["shape"]
dev = torch.device("cuda", 0)
torch.manual_seed(1)
gpw = GaussianProcessTemplateWeighter(wl["E"].to(dev), bench._Cfg(shp.kernel, shp.d), lengthscale=1.0).to(dev)
ls = bench.product_lengthscale(gpw); gpw.covar_module.base_kernel.initialize(lengthscale=ls)
eng = GPAdapterEngine(gpw, EngineConfig(S_train=shp.S, S_eval=shp.S, batch_size=shp.B, shots=shp.shots, seed=1234, precision="tf32"))
f, y = wl["f_train"].to(dev), wl["y_train"].to(dev)
for steps in (0, 50, 300):
    for it in range(steps):
        b = it % (f.shape[0] // shp.B)
        eng.train_step(f[b * shp.B:(b + 1) * shp.B], y[b * shp.B:(b + 1) * shp.B])
    eng.eval_prototypes()
    w = eng.last_eval_w                      # [S, C, T]
    per_sample = (w > 0).sum(-1).float()
    union = ((w > 0).any(0)).sum(-1).float()
    print(f"after {steps:4d} more steps: support per sample mean {float(per_sample.mean()):.1f} (min {int(per_sample.min())}, max {int(per_sample.max())}); "
          f"union over the {w.shape[0]} samples per class mean {float(union.mean()):.1f} of {w.shape[2]} (max {int(union.max())})")
