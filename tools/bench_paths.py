"""Timings of the hot-path kernels that bench.py's headline step does not isolate (CUDA events, L2 flushed between reps, median):
the calibration / binning passes, the stand-alone prototype kernels, the Tip-Adapter cache affinity at cfg4 and the T=64 Matern GP
of cfg5.  HBM-bound passes are reported as achieved GB/s over their ALGORITHMIC bytes against MEASURED_PEAKS.json.
    python tools/bench_paths.py [section ...]      sections: metrics proto tip cfg5"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F
from clip_gp_b200 import _lib, metrics as gm, ops, synth, tc

dev = torch.device("cuda", 0)
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
pk = {}
if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
    pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
HBM = float(pk.get("hbm_gbs", 6544.0))
TF = float(pk.get("bf16_tflops", 1626.7))
want = set(sys.argv[1:]) or {"metrics", "proto", "tip", "cfg5"}


def timeit(fn, reps=9):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms = []
    for _ in range(reps):
        flush.fill_(0.0)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    ms.sort()
    return ms[len(ms) // 2]


def line(name, ms, nbytes=None, flops=None, note=""):
    s = f"{name:<58s} {ms * 1e3:9.1f} us"
    if nbytes is not None:
        gbs = nbytes / ms / 1e6
        s += f"  {nbytes / 1e6:8.1f} MB  {gbs:7.0f} GB/s ({gbs / HBM * 100:4.1f}% of {HBM:.0f})"
    if flops is not None:
        tf = flops / ms / 1e9
        s += f"  {tf:7.1f} TFLOP/s ({tf / TF * 100:4.1f}% of {TF:.0f})"
    print(s + ("  " + note if note else ""), flush=True)


g = torch.Generator().manual_seed(0)

if "metrics" in want:
    N, C = 50000, 1000
    logits = (3.0 * torch.randn(N, C, generator=g)).to(dev)
    y = torch.randint(0, C, (N,), generator=g).to(dev)
    out = {}
    def calib():
        out["r"] = gm.calibration_pass(logits, y, 10, want_conf=True)
    ms = timeit(calib)
    line("calibration pass (softmax max/argmax + ECE histogram) [50000,1000]", ms, N * C * 4 + N * 8 + N * 5)
    conf, correct = out["r"][0], out["r"][1]
    ms = timeit(lambda: gm.aece_pass(conf, correct, 10))
    line("AECE radix select over (conf, hit) N=50000", ms, N * 5, note="(5 passes over 250 KB: launch-latency bound)")
    ms = timeit(lambda: gm.evaluate_calibration(logits, y, 10))
    line("evaluate_calibration: accuracy + ECE + AECE (incl. host reads)", ms, N * C * 4)
    del logits

if "proto" in want:
    wl = synth.make_workload("cfg2"); shp = wl["shape"]
    E = wl["E"].to(dev)
    w = torch.softmax(torch.randn(shp.S, shp.C, shp.T, generator=g), -1).to(dev).requires_grad_(True)
    ms = timeit(lambda: ops.prototypes(w.detach(), E))
    line("prototypes forward  w[10,1000,32] x E[1000,32,512]", ms, E.numel() * 4 + shp.S * shp.C * shp.D * 4 + w.numel() * 4)
    P = ops.prototypes(w, E)
    dP = torch.randn_like(P)
    def bwd():
        w.grad = None
        P.backward(dP, retain_graph=True)
    ms = timeit(bwd)
    line("prototypes adjoint  dP[10,1000,512] -> dw", ms, E.numel() * 4 + shp.S * shp.C * shp.D * 4 + w.numel() * 4)
    ms = timeit(lambda: ops.prototypes_reduced(w.detach(), E, want_mean_hat=True))
    line("prototypes reduced (unit rows, MC mean)", ms, E.numel() * 4 + shp.C * shp.D * 4 + w.numel() * 4)
    del E, P, dP

if "tip" in want:
    B, N_tr, C, D, N = 128, 16000, 1000, 1024, 50000
    mu = torch.randn(C, D, generator=g)
    lab = torch.arange(C).repeat_interleave(16)
    keys = F.normalize(mu[lab] + 2.0 * torch.randn(N_tr, D, generator=g), dim=-1).to(dev)
    labd = lab.to(dev)
    yb = torch.randint(0, C, (B,), generator=g)
    f = F.normalize(mu[yb] + 2.0 * torch.randn(B, D, generator=g), dim=-1).to(dev)
    clipw = F.normalize(mu, dim=-1).to(dev)
    clip = ops.matmul_nt(f, clipw, 100.0)
    kd = keys.clone().requires_grad_(True)
    opt = torch.optim.AdamW([kd], lr=1e-3, eps=1e-4)
    ybd = yb.to(dev)
    def tip_step():
        out = ops.tip_logits(f, kd, labd, clip, 2.0, 20.0, C)
        loss = ops.cross_entropy(out, ybd)
        opt.zero_grad(); loss.backward(); opt.step()
    ms = timeit(tip_step)
    line("Tip-Adapter-F train step cfg4 (fp32 FFMA affinity, autograd, torch AdamW)", ms, 5 * N_tr * D * 4,
         flops=6.0 * B * N_tr * D, note="bytes: keys r + grad w/r + Adam m,v r/w lower bound")
    for prec in ("bf16x3", "bf16"):
        def tip_step_tc():
            out = ops.tip_logits(f, kd, labd, clip, 2.0, 20.0, C, prec)
            loss = ops.cross_entropy(out, ybd)
            opt.zero_grad(); loss.backward(); opt.step()
        ms = timeit(tip_step_tc)
        line(f"Tip-Adapter-F train step cfg4 (tcgen05 {prec} affinity + key-gradient GEMMs, torch AdamW)", ms, 5 * N_tr * D * 4,
             flops=6.0 * B * N_tr * D)
    from clip_gp_b200.tip_engine import TipAdapterEngine
    for prec in ("bf16x3", "bf16", "fp32"):
        eng = TipAdapterEngine(keys, labd, C, B, 2.0, 20.0, total_steps=1000, precision=prec)
        eng.train_step(f, clip, ybd)
        ms = timeit(lambda: eng.train_step(f, clip, ybd))
        line(f"Tip-Adapter-F train step cfg4, fused engine ({prec}, CUDA graph, own AdamW)", ms, 5 * N_tr * D * 4, flops=6.0 * B * N_tr * D)
        del eng
    ms = timeit(lambda: ops.tip_logits(f, keys, labd, clip, 2.0, 20.0, C))
    line("Tip logits forward B=128, fp32 FFMA affinity + cache kernel", ms, N_tr * D * 4, flops=2.0 * B * N_tr * D)
    lib = _lib.load()
    li = labd.to(torch.int32).contiguous()
    fa, kb = tc.cast_bf16(f, tc.SPLIT_A), tc.cast_bf16(keys, tc.SPLIT_B)
    o3 = clip.clone()
    def tip_tc(fa_, rows, out_):
        _lib.check(lib.clipgp_tc_tip_logits(fa_.data_ptr(), rows, kb.data_ptr(), N_tr, 3 * D, li.data_ptr(), 2.0, 20.0, out_.data_ptr(), C,
                                            _lib.stream_ptr(dev)), "clipgp_tc_tip_logits")
    ms = timeit(lambda: tip_tc(fa, B, o3))
    line("Tip logits forward B=128, tcgen05 split-bf16, fused exp / class-sum epilogue", ms, N_tr * 3 * D * 2, flops=2.0 * B * N_tr * D)
    yt = torch.randint(0, C, (N,), generator=g)
    ft = F.normalize(mu[yt] + 2.0 * torch.randn(N, D, generator=g), dim=-1).to(dev)
    fta = tc.cast_bf16(ft, tc.SPLIT_A)
    oN = ops.matmul_nt(ft, clipw, 100.0)
    ms = timeit(lambda: tip_tc(fta, N, oN), reps=5)
    line("Tip logits eval N=50000 x 16000 keys, tcgen05 split-bf16 fused epilogue", ms, None, flops=2.0 * N * N_tr * D,
         note=f"(issued flops x3: {3 * 2.0 * N * N_tr * D / ms / 1e9:.0f} TFLOP/s)")
    fb, kbb = tc.cast_bf16(ft, tc.PLAIN), tc.cast_bf16(keys, tc.PLAIN)
    def tip_tc_plain():
        _lib.check(lib.clipgp_tc_tip_logits(fb.data_ptr(), N, kbb.data_ptr(), N_tr, D, li.data_ptr(), 2.0, 20.0, oN.data_ptr(), C,
                                            _lib.stream_ptr(dev)), "clipgp_tc_tip_logits")
    ms = timeit(tip_tc_plain, reps=5)
    line("Tip logits eval N=50000 x 16000 keys, tcgen05 bf16 fused epilogue", ms, None, flops=2.0 * N * N_tr * D)
    del keys, kd, ft, fta, fb, kbb, oN

if "cfg5" in want:
    sys.path.insert(0, ROOT)
    from oracle import gp as ogp          # state construction only (PCA, prior mean); the timed calls are the CUDA kernels
    wl = synth.make_workload("cfg5"); shp = wl["shape"]
    st = ogp.build_state(wl["E"], "matern", shp.d)
    st.var_mean, st.chol_var = synth.trained_like_q(shp.C, shp.T + 1, 3)
    S, n = 100, shp.T + 1
    eps = torch.randn(shp.C, shp.T, S, generator=g).to(dev)
    mean_x = ogp.residual_mean(st.f0, st.cls_bias, st.tmp_bias, n + shp.T)[:, n:].contiguous().to(dev)
    Z, X = st.inducing_points.to(dev), st.templates_red.to(dev)
    ls, vm, cv = st.kernel.raw_lengthscale.to(dev), st.var_mean.to(dev), st.chol_var.to(dev)
    Et = st.templates.to(dev)
    ms = timeit(lambda: ops.gp_weights(Z, X, ls, None, None, vm, cv, mean_x, eps, "matern", S), reps=5)
    by = shp.C * (n * shp.d * 4 + shp.T * shp.d * 4 + n * n * 4 + S * shp.T * 4 * 2)
    line("cfg5 GP forward C=397 T=64 S=100 Matern (general block kernel)", ms, by)
    Zr, lsr, vmr, cvr = Z.clone().requires_grad_(True), ls.clone().requires_grad_(True), vm.clone().requires_grad_(True), cv.clone().requires_grad_(True)
    def fb5():
        for t in (Zr, lsr, vmr, cvr):
            t.grad = None
        w, kl, _ = ops.gp_weights(Zr, X, lsr, None, None, vmr, cvr, mean_x, eps, "matern", S)
        P = ops.prototypes(w, Et)
        (P.square().sum() * 1e-3 + kl.sum()).backward()
    ms = timeit(fb5, reps=5)
    line("cfg5 GP + prototypes forward + adjoint (autograd surface)", ms)
    w, _, _ = ops.gp_weights(Z, X, ls, None, None, vm, cv, mean_x, eps, "matern", S)
    ms = timeit(lambda: ops.prototypes_reduced(w, Et, want_mean_raw=True), reps=5)
    line("cfg5 TaskRes init: normalised MC-mean prototypes (taskres.py:281-285)", ms, Et.numel() * 4 + w.numel() * 4)
