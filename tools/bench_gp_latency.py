"""GP forward / adjoint kernel time against the number of classes in flight (CUDA events, L2 flushed): separates the per-class
dependent-chain latency (C = 148: one class per SM) from the throughput regime (C = 1000: seven classes per SM).
    python tools/bench_gp_latency.py [C ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from clip_gp_b200 import synth
from clip_gp_b200.engine import EngineConfig, GPAdapterEngine
from clip_gp_b200.gp_template_weigher import GaussianProcessTemplateWeighter

dev = torch.device("cuda", 0)
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
Cs = [int(a) for a in sys.argv[1:]] or [148, 296, 1000]
for Cn in Cs:
    name = f"lat{Cn}"
    synth.CONFIGS[name] = synth.WorkloadShape(name, noise=6.0, C=Cn, T=32, D=512, d=256, S=10, shots=2, B=128, N_test=256, kernel="rbf")
    wl = synth.make_workload(name); shp = wl["shape"]
    torch.manual_seed(1)
    gpw = GaussianProcessTemplateWeighter(wl["E"].to(dev), bench._Cfg(shp.kernel, shp.d), lengthscale=1.4146).to(dev)
    eng = GPAdapterEngine(gpw, EngineConfig(S_train=shp.S, S_eval=shp.S, batch_size=shp.B, shots=shp.shots, seed=1234, precision="bf16x3"))
    f, y = wl["f_train"][:shp.B].to(dev), wl["y_train"][:shp.B].to(dev)
    kt = bench.profile_step_kernels(eng, f.repeat(2, 1), y.repeat(2), shp, flush, reps=7)
    ev = []
    for _ in range(7):
        flush.fill_(0.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.eval_prototypes(shp.S); e1.record(); torch.cuda.synchronize()
        ev.append(e0.elapsed_time(e1))
    ev.sort()
    print(f"C={Cn:5d}: gp_forward {kt.get('gp_forward', {'ms': 0})['ms'] * 1e3:7.1f} us   gp_backward {kt.get('gp_backward', {'ms': 0})['ms'] * 1e3:7.1f} us   "
          f"eval gp_forward(+MC-mean prototypes) {ev[3] * 1e3:7.1f} us", flush=True)
    del eng, gpw
