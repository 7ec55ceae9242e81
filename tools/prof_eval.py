"""Run the 50k-image eval pass of the bench (split-bf16, collapsed MC form) a few times, for the ncu launch list.
    python tools/prof_eval.py [precision] [mc] [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from clip_gp_b200 import synth
from clip_gp_b200.engine import EngineConfig, GPAdapterEngine
from clip_gp_b200.gp_template_weigher import GaussianProcessTemplateWeighter

precision = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
mc = sys.argv[2] if len(sys.argv) > 2 else "collapsed"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
wl = synth.make_workload("cfg2"); shp = wl["shape"]
dev = torch.device("cuda", 0)
ls = bench.bench_lengthscale(wl["E"], shp.d)
torch.manual_seed(1)
gpw = GaussianProcessTemplateWeighter(wl["E"].to(dev), bench._Cfg(shp.kernel, shp.d), lengthscale=ls).to(dev)
eng = GPAdapterEngine(gpw, EngineConfig(S_train=shp.S, S_eval=shp.S, batch_size=shp.B, shots=shp.shots, seed=1234, precision="bf16x3"))
f, y = wl["f_test"].to(dev), wl["y_test"].to(dev)
for _ in range(reps):
    conf, correct, hist = eng.eval_calibration_tc(f, y, precision=precision, mc=mc)
torch.cuda.synchronize()
print("top1", int(hist[3, 0]), "of", f.shape[0])
