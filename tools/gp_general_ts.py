"""Per-phase clock64() timestamps of the FIRST class CTA of the general GP block kernels (n up to 65: cfg5 = SUN397, T = 64), from the
debug library built by tools/build_ts.sh (-DCLIPGP_PHASE_TS):   bash tools/build_ts.sh ; python tools/gp_general_ts.py [C ...]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from clip_gp_b200 import _lib, ops, synth
from oracle import gp as ogp          # state construction only (PCA, prior mean); the timed calls are the CUDA kernels

_lib.LIB_PATH = os.path.join(ROOT, "clip_gp_b200", "lib", "libclipgp_ts.so")
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
wl = synth.make_workload("cfg5"); shp = wl["shape"]
st = ogp.build_state(wl["E"], "matern", shp.d)
st.var_mean, st.chol_var = synth.trained_like_q(shp.C, shp.T + 1, 3)
S, n = 100, shp.T + 1
for Cn in [int(a) for a in sys.argv[1:]] or [148, shp.C]:
    eps = torch.randn(Cn, shp.T, S, generator=g).to(dev)
    mean_x = ogp.residual_mean(st.f0, st.cls_bias, st.tmp_bias, n + shp.T)[:Cn, n:].contiguous().to(dev)
    Z, X = st.inducing_points[:Cn].contiguous().to(dev), st.templates_red[:Cn].contiguous().to(dev)
    ls, vm, cv = st.kernel.raw_lengthscale[:Cn].contiguous().to(dev), st.var_mean[:Cn].contiguous().to(dev), st.chol_var[:Cn].contiguous().to(dev)
    Zr, lsr, vmr, cvr = (t.clone().requires_grad_(True) for t in (Z, ls, vm, cv))
    for _ in range(3):
        for t in (Zr, lsr, vmr, cvr):
            t.grad = None
        w, kl, _ = ops.gp_weights(Zr, X, lsr, None, None, vmr, cvr, mean_x, eps, "matern", S)
        (w.square().sum() + kl.sum()).backward()
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    for t in (Zr, lsr, vmr, cvr):
        t.grad = None
    e0.record()
    w, kl, _ = ops.gp_weights(Zr, X, lsr, None, None, vmr, cvr, mean_x, eps, "matern", S)
    e1.record()
    (w.square().sum() + kl.sum()).backward()
    e2.record()
    torch.cuda.synchronize()
    lib = ctypes.CDLL(_lib.load()._name)
    buf = (ctypes.c_longlong * 32)()
    lib.clipgp_debug_general_ts(buf)
    ts = list(buf)
    nf = ["stage hyper, Lq, m, alias check", "Gram K_ZZ", "save K, K->fp64 L / A / Sigma", "chol64 + fused forward solve", "A->f32, Bm, mu", "Sigma",
          "chol32 (+ retries)", "saves, KL", "sampling + sparsemax"]
    print(f"C={Cn}  T={shp.T} S={S}: forward {e0.elapsed_time(e1) * 1e3:.0f} us, adjoint (incl. autograd glue) {e1.elapsed_time(e2) * 1e3:.0f} us")
    print(f"  forward, cycles of the first class CTA (total {ts[9] - ts[0]}):")
    for i, nm in enumerate(nf):
        print(f"     {nm:40s} {ts[i + 1] - ts[i]:8d}")
    lib.clipgp_debug_general_ts_bwd(buf)
    ts = list(buf)
    nb = ["stage L, A, m", "B1 sparsemax adj, dmu, dR (S chunks)", "chol32 adjoint", "B3 products (Bm, dBm, dA, dLq, dm)", "trsmT64 (dK_ZX)", "dL = -tril(dK_ZX A^T)",
          "chol64 adjoint", "dKzz -> f32", "B5 kernel adjoint (Z stream)"]
    print(f"  adjoint, cycles of the first class CTA (total {ts[9] - ts[0]}):")
    for i, nm in enumerate(nb):
        print(f"     {nm:40s} {ts[i + 1] - ts[i]:8d}")
