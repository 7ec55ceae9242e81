"""Run a few EAGER GP-Adapter steps of the bench workload (for ncu: graph replays hide the kernels' names).
    python tools/prof_step.py [workload] [steps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from clip_gp_b200 import synth
from clip_gp_b200.engine import EngineConfig, GPAdapterEngine
from clip_gp_b200.gp_template_weigher import GaussianProcessTemplateWeighter

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
precision = sys.argv[3] if len(sys.argv) > 3 else "bf16x3"
wl = synth.make_workload(name, n_test=4096); shp = wl["shape"]
dev = torch.device("cuda", 0)
ls = 1.4146 if shp.kernel == "rbf" else None
torch.manual_seed(1)
gpw = GaussianProcessTemplateWeighter(wl["E"].to(dev), bench._Cfg(shp.kernel, shp.d), lengthscale=ls).to(dev)
eng = GPAdapterEngine(gpw, EngineConfig(S_train=shp.S, S_eval=shp.S, batch_size=shp.B, shots=shp.shots, seed=1234,
                                       precision=precision))
f, y = wl["f_train"].to(dev), wl["y_train"].to(dev)
for it in range(steps):
    loss = eng.train_step(f[it * shp.B:(it + 1) * shp.B], y[it * shp.B:(it + 1) * shp.B], use_graph=False)
torch.cuda.synchronize()
print("loss", float(loss), "status", int(eng.status.abs().max()))
res = eng.evaluate(wl["f_test"].to(dev), wl["y_test"].to(dev))
print("eval", res["top1_acc"], res["ece"], res["aece"])
