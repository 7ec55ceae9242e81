"""Graph-replayed cfg2 train step in ms (L2 flushed between steps), for quick A/B runs of kernel experiments (env toggles)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from clip_gp_b200 import _lib
if os.environ.get("CLIPGP_LIB"):        # A/B runs: another build of the library (tools/micro/ab_step.py)
    _lib.LIB_PATH = os.path.join(ROOT, "clip_gp_b200", "lib", os.environ["CLIPGP_LIB"])
import bench
from clip_gp_b200 import synth
from clip_gp_b200.engine import EngineConfig, GPAdapterEngine
from clip_gp_b200.gp_template_weigher import GaussianProcessTemplateWeighter

precision = sys.argv[1] if len(sys.argv) > 1 else "tf32"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
wl = synth.make_workload("cfg2", n_test=256); shp = wl["shape"]
dev = torch.device("cuda", 0)
torch.manual_seed(1)
gpw = GaussianProcessTemplateWeighter(wl["E"].to(dev), bench._Cfg(shp.kernel, shp.d), lengthscale=1.4146).to(dev)
eng = GPAdapterEngine(gpw, EngineConfig(S_train=shp.S, S_eval=shp.S, batch_size=shp.B, shots=shp.shots, seed=1234, precision=precision,
                                       fuse_tail=not os.environ.get("CLIPGP_NO_FUSE_TAIL"),
                                       clear_on_side=not os.environ.get("CLIPGP_CLEAR_ON_MAIN")))
f, y = wl["f_train"].to(dev), wl["y_train"].to(dev)
flush = torch.empty(64 * 1024 * 1024, device=dev)
nb = f.shape[0] // shp.B
for it in range(20):
    eng.train_step(f[(it % nb) * shp.B:(it % nb + 1) * shp.B], y[(it % nb) * shp.B:(it % nb + 1) * shp.B])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
tot = 0.0
for it in range(steps):
    flush.fill_(0.0)
    b = it % nb
    e0.record(); eng.train_step(f[b * shp.B:(b + 1) * shp.B], y[b * shp.B:(b + 1) * shp.B]); e1.record()
    torch.cuda.synchronize()
    tot += e0.elapsed_time(e1)
print(f"{precision}: {tot / steps:.4f} ms/step = {steps / tot * 1e3:.0f} steps/s  status {int(eng.status.abs().max())}")
