"""Per-phase clock64() timestamps of ONE class CTA of the GP forward / adjoint kernels (debug library built with -DCLIPGP_PHASE_TS):
    nvcc <flags of __graft_entry__> -DCLIPGP_PHASE_TS clip_gp_b200/csrc/*.cu -o clip_gp_b200/lib/libclipgp_ts.so ; python tools/gp_phase_ts.py [C]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from clip_gp_b200 import _lib, synth
from clip_gp_b200.engine import EngineConfig, GPAdapterEngine
from clip_gp_b200.gp_template_weigher import GaussianProcessTemplateWeighter

_lib.LIB_PATH = os.path.join(ROOT, "clip_gp_b200", "lib", "libclipgp_ts.so")      # the -DCLIPGP_PHASE_TS build
dev = torch.device("cuda", 0)
for Cn in [int(a) for a in sys.argv[1:]] or [148, 1000]:
    name = f"lat{Cn}"
    synth.CONFIGS[name] = synth.WorkloadShape(name, noise=6.0, C=Cn, T=32, D=512, d=256, S=10, shots=2, B=128, N_test=256, kernel="rbf")
    wl = synth.make_workload(name); shp = wl["shape"]
    torch.manual_seed(1)
    gpw = GaussianProcessTemplateWeighter(wl["E"].to(dev), bench._Cfg(shp.kernel, shp.d), lengthscale=1.4146).to(dev)
    eng = GPAdapterEngine(gpw, EngineConfig(S_train=shp.S, S_eval=shp.S, batch_size=shp.B, shots=shp.shots, seed=1234, precision="bf16x3"))
    f, y = wl["f_train"][:shp.B].to(dev), wl["y_train"][:shp.B].to(dev)
    for _ in range(3):
        eng.train_step(f, y, use_graph=False)
    torch.cuda.synchronize()
    lib = ctypes.CDLL(_lib.load()._name)
    buf = (ctypes.c_longlong * 64)()
    names_f = ["invls+gram", "handover+stage Lq", "chol64+trsm64 (warp 0) | KL", "A->f32, mu, Bm", "Sigma", "chol32", "saves", "sample+sparsemax",
               "prototypes (E stream, norms, bf16 rows)"]
    lib.clipgp_debug_phase_ts(buf, 0)
    ts = list(buf)[:10]
    print(f"C={Cn} forward (cycles of class 0; total {ts[9] - ts[0]}):")
    for i, nm in enumerate(names_f):
        print(f"   {nm:45s} {ts[i + 1] - ts[i]:8d}")
    full = list(buf)
    print(f"   detail: chol64 {full[20] - ts[2]}  trsm64 {full[21] - full[20]}  wait-for-barrier {ts[3] - full[21]} | chol32 attempts (copy, factor): "
          + ", ".join(f"({full[23 + 2 * a] - full[22 + 2 * a]} of {full[22 + 2 * a] - ts[5]}+)" for a in range(4) if full[22 + 2 * a] > ts[5]))
    names_b = ["stage dP rows + EEt", "a = <g, E>: E stream", "dw", "P1 stage R, sparsemax adj, dR", "chol32 adjoint", "P2 products (H, dBm, dA, dLq, dm)",
               "P3 stage L", "trsmT64", "dL", "chol64 adjoint", "dK assembly", "kernel adjoint (Z stream)"]
    lib.clipgp_debug_phase_ts_bwd(buf)
    ts = list(buf)[:13]
    print(f"C={Cn} adjoint (cycles of class 0; total {ts[12] - ts[0]}):")
    for i, nm in enumerate(names_b):
        print(f"   {nm:45s} {ts[i + 1] - ts[i]:8d}")
    del eng, gpw
