"""Two eager full-batch steps (B = N_train = 16 000, bf16 GEMMs) for ncu.   python tools/prof_fullbatch.py [precision]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from clip_gp_b200 import synth
from clip_gp_b200.engine import EngineConfig, GPAdapterEngine
from clip_gp_b200.gp_template_weigher import GaussianProcessTemplateWeighter

precision = sys.argv[1] if len(sys.argv) > 1 else "bf16"
wl = synth.make_workload("cfg2", n_test=1024); shp = wl["shape"]
dev = torch.device("cuda", 0)
torch.manual_seed(1)
gpw = GaussianProcessTemplateWeighter(wl["E"].to(dev), bench._Cfg(shp.kernel, shp.d), lengthscale=1.4146).to(dev)
f, y = wl["f_train"].to(dev), wl["y_train"].to(dev)
eng = GPAdapterEngine(gpw, EngineConfig(S_train=shp.S, S_eval=shp.S, batch_size=f.shape[0], shots=shp.shots, seed=1234, precision=precision,
                                       overlap=False))
for it in range(2):
    loss = eng.train_step(f, y, use_graph=False)
torch.cuda.synchronize()
print("loss", float(loss))
