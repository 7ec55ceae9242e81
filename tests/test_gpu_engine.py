"""GPU: the fused GP-Adapter engine (train step + MC-averaged eval) against the CPU oracle step."""
import copy

import pytest
import torch

from clip_gp_b200 import synth
from clip_gp_b200.engine import EngineConfig, GPAdapterEngine
from clip_gp_b200.gp_template_weigher import GaussianProcessTemplateWeighter
from oracle import gp as ogp
from oracle import metrics as om
from oracle import philox
from oracle.train_step import OracleAdapter
from tests.helpers import assert_parity, fix_eval_noise, max_err, rel_err, state_to, within

pytestmark = pytest.mark.gpu


class _Cfg:
    def __init__(self, kernel, pca):
        self.adapter = type("A", (), {"gp_pca_dim": pca, "gp_kernel_type": kernel})()


def build(kernel, name="small", loss_mode="per_sample", S=4, seed=3, precision="fp32"):
    wl = synth.make_workload(name); shp = wl["shape"]
    torch.manual_seed(0)
    gpw = GaussianProcessTemplateWeighter(wl["E"], _Cfg(kernel, shp.d)).to("cuda")
    q = gpw.variational_strategy._variational_distribution
    m, Lq = synth.trained_like_q(shp.C, shp.T + 1, 5)
    gpw.variational_strategy._maybe_init()
    with torch.no_grad():
        q.variational_mean.copy_(m); q.chol_variational_covar.copy_(Lq)
    cfg = EngineConfig(S_train=S, S_eval=S, batch_size=shp.B, shots=shp.shots, loss_mode=loss_mode, seed=seed,
                       precision=precision)
    eng = GPAdapterEngine(gpw, cfg)
    # the oracle twin
    st = ogp.build_state(wl["E"], kernel, shp.d)
    st.inducing_points = gpw.variational_strategy.inducing_points.detach().cpu().clone()
    st.var_mean, st.chol_var = m.clone(), Lq.clone()
    raw_ls, raw_os, raw_var = gpw._kernel_raw()
    if raw_ls is not None: st.kernel.raw_lengthscale = raw_ls.detach().cpu().clone()
    if raw_os is not None: st.kernel.raw_outputscale = raw_os.detach().cpu().clone()
    if raw_var is not None: st.kernel.raw_variance = raw_var.detach().cpu().clone()
    mk = lambda s_: OracleAdapter(s_, shp.D, scale=cfg.logit_scale, gp_beta=cfg.gp_beta, l2_lambda=cfg.l2_lambda, shots=cfg.shots,
                                  lr=cfg.lr, gp_lr=cfg.gp_lr, loss_mode=loss_mode)
    st64 = state_to(st, dtype=torch.float64)
    orc = mk(st)
    orc.twin64 = mk(st64)          # the same step in float64: `truth64` of tests.helpers.assert_parity
    return wl, shp, eng, orc, cfg


def oracle_grads_pair(orc, f, y, eps):
    """name -> (fp32 oracle gradient, float64 oracle gradient) for one loss evaluation of both twins; returns (loss32, dict)."""
    out = []
    for o in (orc, orc.twin64):
        loss = o.loss(f, y, eps)
        loss.backward()
        k = o.st.kernel
        g = {"W": o.W.grad, "m": o.st.var_mean.grad, "Lq": o.st.chol_var.grad, "z_last": o.st.inducing_points.grad[:, -1]}
        if k.raw_lengthscale is not None: g["ls"] = k.raw_lengthscale.grad
        if k.raw_outputscale is not None: g["os"] = k.raw_outputscale.grad
        if k.raw_variance is not None: g["var"] = k.raw_variance.grad
        out.append((loss, g))
    return out[0][0], {n: (out[0][1][n], out[1][1][n]) for n in out[0][1]}


@pytest.mark.parametrize("kernel", ["rbf", "matern", "linear"])
@pytest.mark.parametrize("loss_mode", ["per_sample", "logit_mean"])
def test_step_loss_and_gradients(kernel, loss_mode):
    wl, shp, eng, orc, cfg = build(kernel, loss_mode=loss_mode)
    f, y = wl["f_train"][: shp.B], wl["y_train"][: shp.B]
    eps = philox.eps_tensor(cfg.seed, 0, shp.C, shp.T, cfg.S_train)
    loss_ref, G = oracle_grads_pair(orc, f, y, eps)
    eng.skip_update = True
    loss = eng.train_step(f.cuda(), y.cuda(), use_graph=False)
    assert float(loss) == pytest.approx(float(loss_ref), rel=1e-3)
    for name, (g32, g64) in G.items():
        got = eng.g(name).view(g32.shape)
        # elementwise gate against the reference-arithmetic (fp32 oracle) gradient, widened by its own distance to float64
        assert_parity(got, g32, g64, rtol=2e-3, name=f"d{name}")
        assert max_err(got, g64) < (2e-2 if name == "z_last" else 2e-3), name       # norm-wise against exact arithmetic
    assert int(eng.status.abs().max()) == 0


@pytest.mark.parametrize("precision", ["tf32", "bf16x3", "bf16"])
@pytest.mark.parametrize("loss_mode", ["per_sample", "logit_mean"])
@pytest.mark.parametrize("name", ["small", "tiny"])
def test_tensor_core_step_loss_and_gradients(precision, loss_mode, name):
    """The same step with all five GEMMs on tcgen05: split operands (bf16x3) meet the fp32 gate; plain bf16 meets its stated
    tolerance (2 % of the largest gradient entry; the reference's own GPU path is TF32, adapter.py:23)."""
    if name == "small" and loss_mode == "logit_mean" and precision == "tf32":
        with pytest.raises(ValueError, match="multiples of 4"):     # C = 37: the collapsed [B, C] logits have a 148-byte row pitch
            build("rbf", name=name, loss_mode=loss_mode, precision=precision)
        return
    wl, shp, eng, orc, cfg = build("rbf", name=name, loss_mode=loss_mode, precision=precision)
    f, y = wl["f_train"][: shp.B], wl["y_train"][: shp.B]
    eps = philox.eps_tensor(cfg.seed, 0, shp.C, shp.T, cfg.S_train)
    loss_ref, G = oracle_grads_pair(orc, f, y, eps)
    eng.skip_update = True
    loss = eng.train_step(f.cuda(), y.cuda(), use_graph=False)
    # tf32: the reference's own GPU arithmetic (adapter.py:23); 10-bit mantissa products -> 5e-3 norm-wise on the gradients
    tol, ltol = {"bf16x3": (2e-3, 1e-3), "tf32": (5e-3, 1e-3), "bf16": (3e-2, 1e-2)}[precision]
    assert float(loss) == pytest.approx(float(loss_ref), rel=ltol)
    for pn in ("W", "m", "Lq", "ls", "os"):
        g32, g64 = G[pn]
        got = eng.g(pn).view(g32.shape)
        if precision == "bf16x3":
            assert_parity(got, g32, g64, rtol=tol, name=f"d{pn}")                 # split operands: the fp32 gate, elementwise
        assert max_err(got, g64) < tol, pn                                         # bf16: stated tolerance, 3 % of the largest entry
    # graph replay of the tensor-core step reproduces the eager launch sequence
    _, _, eng2, _, _ = build("rbf", name=name, loss_mode=loss_mode, precision=precision)
    eng2.skip_update = True
    loss2 = eng2.train_step(f.cuda(), y.cuda(), use_graph=True)
    assert float(loss2) == pytest.approx(float(loss), rel=1e-5)
    assert within(eng2.flat_g, eng.flat_g, 1e-5)            # split-K reductions: fp32 summation order varies run to run


@pytest.mark.parametrize("S", [10, 12, 23])
def test_step_with_more_samples_than_one_prototype_pass(S):
    """S = 10 is the single-pass form of the fused prototype stages (16-byte weight loads, zero-weight skipping); S = 12 / 23 take the
    chunked passes over E[c] (10 samples per pass) and the fused adjoint's S <= 12 limit / the un-fused fallback."""
    wl, shp, eng, orc, cfg = build("rbf", name="tiny", S=S)
    f, y = wl["f_train"][: shp.B], wl["y_train"][: shp.B]
    eps = philox.eps_tensor(cfg.seed, 0, shp.C, shp.T, cfg.S_train)
    loss_ref, G = oracle_grads_pair(orc, f, y, eps)
    eng.skip_update = True
    loss = eng.train_step(f.cuda(), y.cuda(), use_graph=False)
    assert float(loss) == pytest.approx(float(loss_ref), rel=1e-4)
    for pn in ("W", "m", "Lq", "ls", "os"):
        g32, g64 = G[pn]
        assert_parity(eng.g(pn).view(g32.shape), g32, g64, rtol=2e-3, name=f"d{pn}")
        assert max_err(eng.g(pn).view(g32.shape), g64) < 2e-3, pn


def test_tf32_large_batch_uses_kmajor_copies_and_matches_in_place_operands():
    """B >= 1024: the adjoint GEMMs read P_hat / f_hat from transposed copies (K-major B operand) instead of MN-major in place.
    Same products, same TF32 rounding of the operands: the gradients agree to accumulation order, and with the float64 oracle."""
    wl, shp, eng, orc, cfg = build("rbf", name="tiny", precision="tf32")
    g = torch.Generator().manual_seed(3)
    B = 1024
    idx = torch.randint(0, wl["f_train"].shape[0], (B,), generator=g)
    f = (wl["f_train"][idx] + 0.05 * torch.randn(B, shp.D, generator=g)).contiguous(); y = wl["y_train"][idx].contiguous()
    eng.skip_update = True
    loss = eng.train_step(f.cuda(), y.cuda(), use_graph=False)
    assert eng.tf32_kmajor_b
    _, _, eng2, _, _ = build("rbf", name="tiny", precision="tf32")
    eng2.skip_update = True
    eng2._alloc_train(B); eng2.tf32_kmajor_b = False
    loss2 = eng2.train_step(f.cuda(), y.cuda(), use_graph=False)
    assert float(loss) == pytest.approx(float(loss2), rel=1e-6)
    assert within(eng.flat_g, eng2.flat_g, 1e-5)
    eps = philox.eps_tensor(cfg.seed, 0, shp.C, shp.T, cfg.S_train)
    loss_ref, G = oracle_grads_pair(orc, f, y, eps)
    assert float(loss) == pytest.approx(float(loss_ref), rel=1e-3)
    for pn in ("W", "m", "Lq", "ls", "os"):
        assert max_err(eng.g(pn).view(G[pn][1].shape), G[pn][1]) < 5e-3, pn


def test_adamw_update_and_graph_replay_match_eager():
    wl, shp, eng_a, orc, cfg = build("rbf")
    _, _, eng_b, _, _ = build("rbf")
    f, y = wl["f_train"], wl["y_train"]
    losses_ref = []
    for it in range(3):
        fb, yb = f[it * shp.B:(it + 1) * shp.B], y[it * shp.B:(it + 1) * shp.B]
        eps = philox.eps_tensor(cfg.seed, it, shp.C, shp.T, cfg.S_train)
        losses_ref.append(orc.step(fb, yb, eps))
        la = float(eng_a.train_step(fb.cuda(), yb.cuda(), use_graph=False))
        lb = float(eng_b.train_step(fb.cuda(), yb.cuda(), use_graph=True))
        assert la == pytest.approx(losses_ref[-1], rel=2e-3)
        assert lb == pytest.approx(la, rel=1e-5)
    assert torch.allclose(eng_a.flat_p, eng_b.flat_p, rtol=1e-5, atol=1e-6)
    # parameters after three AdamW steps (sign-like first steps: compare with an absolute budget of one lr step)
    W = eng_a.p("W").view(shp.D, shp.D).cpu()
    assert float((W - orc.W.detach()).abs().max()) < 0.5 * cfg.lr
    m = eng_a.p("m").view(shp.C, -1).cpu()
    bad = ((m - orc.st.var_mean.detach()).abs() > 0.5 * cfg.gp_lr).float().mean()
    assert float(bad) < 0.01
    assert int(eng_a.adam_step) == 4 and int(eng_a.rng_state[1]) == 3
    assert torch.equal(eng_a.Z[:, :-1].cpu(), orc.st.inducing_points.detach()[:, :-1])     # frozen rows untouched


def test_eval_matches_oracle_logit_mean_and_metrics():
    wl, shp, eng, orc, cfg = build("rbf", S=6)
    f, y = wl["f_test"], wl["y_test"]
    eps = fix_eval_noise(eng, 6)
    logits_ref = orc.eval_logits(f, eps)                       # materialised [B,S,C] mean (adapter.py:247-249)
    logits = eng.eval_logits(f.cuda(), S=6)                    # collapsed: one GEMM against mean_s p_hat_s
    assert float((logits.cpu() - logits_ref).abs().max()) < 1e-3 * float(logits_ref.abs().max())
    res = eng.evaluate(f.cuda(), y.cuda(), S=6)
    assert res["top1_count"] == om.top1_count(logits_ref, y)
    e, b = om.compute_ece_with_bins(logits_ref, y)
    conf, _, _ = om.confidence(logits_ref, y)
    gap = float((conf[:, None] - torch.linspace(0, 1, 11)[None]).abs().min())
    assert res["calibration"]["bin_count"] == b["bin_count"], f"min |conf-boundary| = {gap:.2e}"
    assert res["ece"] == pytest.approx(e, rel=1e-3, abs=1e-3)
    assert res["aece"] == pytest.approx(om.compute_aece(logits_ref, y), rel=1e-3, abs=1e-3)


@pytest.mark.parametrize("precision,mc", [("bf16x3", "collapsed"), ("bf16x3", "materialised"), ("bf16", "collapsed"), ("bf16", "materialised"),
                                          ("tf32", "collapsed"), ("tf32", "materialised")])
def test_tensor_core_eval_modes(precision, mc):
    """tcgen05 eval (fused calibration epilogue) against the oracle's materialised logit-mean: bf16x3 meets the fp32 gate
    (1e-3 relative on logits, identical top-1 / bin counts up to boundary ties); bf16 meets its stated tolerance."""
    wl, shp, eng, orc, cfg = build("rbf", S=6)
    f, y = wl["f_test"], wl["y_test"]
    eps = fix_eval_noise(eng, 6)
    logits_ref = orc.eval_logits(f, eps)
    conf, correct, hist = eng.eval_calibration_tc(f.cuda(), y.cuda(), S=6, precision=precision, mc=mc, want_logits=True)
    logits = eng.last_eval_logits.cpu()
    err = float((logits - logits_ref).abs().max())
    scale = float(logits_ref.abs().max())
    res = eng.evaluate(f.cuda(), y.cuda(), S=6, precision=precision, mc=mc)
    e_ref, b_ref = om.compute_ece_with_bins(logits_ref, y)
    if precision == "tf32":
        # TF32 (the reference's GPU arithmetic): logits within 1e-3 of their scale, accuracy / ECE within the fragile-sample band
        assert err < 1e-3 * scale
        top2 = logits_ref.topk(2, dim=1).values
        fragile = int(((top2[:, 0] - top2[:, 1]) < 2 * err).sum())
        assert abs(res["top1_count"] - om.top1_count(logits_ref, y)) <= fragile
        assert res["ece"] == pytest.approx(e_ref, abs=0.15)          # percentage points, 1000 images: a handful of boundary crossings
    elif precision == "bf16x3":
        assert err < 1e-3 * scale
        conf_ref, _, _ = om.confidence(logits_ref, y)
        gap = float((conf_ref[:, None] - torch.linspace(0, 1, 11)[None]).abs().min())
        assert res["top1_count"] == om.top1_count(logits_ref, y)
        if gap > 1e-3:
            assert res["calibration"]["bin_count"] == b_ref["bin_count"]
        assert res["ece"] == pytest.approx(e_ref, rel=1e-3, abs=2e-3)
        assert res["aece"] == pytest.approx(om.compute_aece(logits_ref, y), rel=1e-3, abs=2e-3)
    else:
        # stated bf16 tolerance: |dlogit| <= 0.25 at logit scale 100 (operands rounded to 8 mantissa bits three times:
        # features, projected features, prototypes); top-1 may flip only for samples whose top-2 margin is inside that band
        assert err < 0.25
        top2 = logits_ref.topk(2, dim=1).values
        fragile = int(((top2[:, 0] - top2[:, 1]) < 0.5).sum())
        assert abs(res["top1_count"] - om.top1_count(logits_ref, y)) <= fragile
        assert res["ece"] == pytest.approx(e_ref, abs=0.5)          # percentage points
    assert sum(res["calibration"]["bin_count"]) == len(y)


def test_device_resident_learning_rate_follows_the_schedule():
    """set_lr / cosine_lr change the step size of a CAPTURED graph (the rate lives in device memory): a zero rate freezes the
    parameters, the cosine rate at epoch e equals torch's CosineAnnealingLR (utils/optimization.py:232-238)."""
    wl, shp, eng, orc, cfg = build("rbf")
    f, y = wl["f_train"].cuda(), wl["y_train"].cuda()
    eng.train_step(f[: shp.B], y[: shp.B], use_graph=True)              # captures the graph with lr = cfg.lr
    p0 = eng.flat_p.clone()
    eng.set_lr(0.0, 0.0)
    eng.train_step(f[: shp.B], y[: shp.B], use_graph=True)
    assert torch.equal(eng.flat_p, p0)
    eng.cosine_lr(epoch=25, max_epoch=100, base_lr=0.01, base_gp_lr=1e-3)
    opt = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=0.01)
    sch = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=100)
    for _ in range(25):
        opt.step(); sch.step()
    assert float(eng.lr_dev[0]) == pytest.approx(opt.param_groups[0]["lr"], rel=1e-6)
    assert float(eng.lr_dev[1]) == pytest.approx(0.1 * opt.param_groups[0]["lr"], rel=1e-6)
    eng.train_step(f[: shp.B], y[: shp.B], use_graph=True)
    assert not torch.equal(eng.flat_p, p0)


def test_eval_graph_replays_track_the_parameters():
    """engine.eval_graph: the captured eval pass (two-stream GP / feature chains + fused calibration GEMM) equals the eager pass, and a
    replay after a training step sees the updated parameters (the graph reads the engine's buffers in place)."""
    wl, shp, eng, _, _ = build("rbf", precision="bf16x3")
    f, y = wl["f_test"].cuda(), wl["y_test"].cuda()
    fix_eval_noise(eng)
    replay = eng.eval_graph(f, y, precision="bf16x3", mc="collapsed")
    conf_e, cor_e, hist_e = eng.eval_calibration_tc(f, y, precision="bf16x3", mc="collapsed")
    conf_g, cor_g, hist_g = replay()
    assert torch.equal(hist_g, hist_e) and torch.equal(conf_g, conf_e) and torch.equal(cor_g, cor_e)
    for it in range(3):
        eng.train_step(wl["f_train"][:shp.B].cuda(), wl["y_train"][:shp.B].cuda())
    conf_e, cor_e, hist_e = eng.eval_calibration_tc(f, y, precision="bf16x3", mc="collapsed")
    before = hist_g.clone()
    conf_g, cor_g, hist_g = replay()
    assert torch.equal(hist_g, hist_e) and torch.equal(conf_g, conf_e)
    assert not torch.equal(before[1], hist_g[1])              # the confidences moved with the parameters


@pytest.mark.parametrize("precision", ["bf16x3", "bf16", "tf32"])
def test_fused_projection_eval_matches_the_three_kernel_form(precision):
    """clipgp_tc_proj_logits_calibration (B = [W ; P W], norm from the projection columns) against the cast -> projection GEMM ->
    normalise -> logits GEMM form of the same engine, with a non-trivial visual projection; D = 256 so that the fused path is taken."""
    from clip_gp_b200 import metrics as gm
    synth.CONFIGS["d256"] = synth.WorkloadShape("d256", C=37, T=8, D=256, d=32, S=4, shots=4, B=48, N_test=3000, kernel="rbf", noise=5.0)
    wl, shp, eng, _, _ = build("rbf", name="d256", precision=precision)
    g = torch.Generator().manual_seed(3)
    with torch.no_grad():
        eng.p("W").add_(0.05 * torch.randn(shp.D * shp.D, generator=g).cuda())
    f, y = wl["f_test"].cuda(), wl["y_test"].cuda()
    assert eng.cfg.fuse_eval_projection
    fix_eval_noise(eng)
    conf_f, cor_f, hist_f = eng.eval_calibration_tc(f, y, precision=precision, mc="collapsed")
    eng.cfg.fuse_eval_projection = False
    conf_u, cor_u, hist_u = eng.eval_calibration_tc(f, y, precision=precision, mc="collapsed")
    N = f.shape[0]
    tol = {"bf16x3": 1e-3, "tf32": 1e-2, "bf16": 3e-2}[precision]
    assert float((conf_f - conf_u).abs().max()) < tol
    assert int((cor_f != cor_u).sum()) <= (0 if precision == "bf16x3" else N // 100) + int(((conf_f - conf_u).abs() > 0).sum() > 0) * 3
    assert int(hist_f[0].sum()) == N
    ece_f, _ = gm.ece_from_counters(gm.counters_from_hist(hist_f, N))
    ece_u, _ = gm.ece_from_counters(gm.counters_from_hist(hist_u, N))
    assert ece_f == pytest.approx(ece_u, abs=0.05 if precision == "bf16x3" else 0.5)
    # and against the exact fp32 path
    ref = eng.evaluate(f, y, precision="fp32")
    assert abs(ref["top1_count"] - int(hist_f[3, 0])) <= (3 if precision == "bf16x3" else N // 100)
    if precision == "tf32":
        assert ece_f == pytest.approx(ref["ece"], abs=0.1)


def test_template_logit_adjoint_matches_the_default_step():
    """EngineConfig.template_logit_adjoint: a[s,t] = scale * sum_b dlogits[b,s,c] (f_hat . E[c,t]) from the per-template cosine GEMM and
    the bf16 dlogits^T operand (clipgp_gp_bwd_args.tl_*) gives the same gradients as the d P_hat GEMM + stream over E[c]."""
    wl, shp, eng0, _, _ = build("rbf", name="t32", precision="bf16x3")
    _, _, eng1, _, _ = build("rbf", name="t32", precision="bf16x3")
    eng1.cfg.template_logit_adjoint = True
    eng1._alloc_train(shp.B)
    assert eng1.tl_adjoint and not eng0.tl_adjoint
    f, y = wl["f_train"][:shp.B].cuda(), wl["y_train"][:shp.B].cuda()
    for e in (eng0, eng1):
        e.skip_update = True
        e.train_step(f, y, use_graph=False)
    torch.cuda.synchronize()
    assert float(eng0.loss) == pytest.approx(float(eng1.loss), rel=1e-6)
    for name in ("m", "Lq", "ls", "z_last"):
        assert max_err(eng1.g(name), eng0.g(name)) < 2e-4, name          # two fp32 routes to the same contraction: norm-wise


def test_host_batch_pipeline_matches_step_by_step():
    """train_steps_host (copy-stream prefetch + asynchronous loss read-back) == the same batches through train_step one by one."""
    wl, shp, eng_a, _, _ = build("rbf", precision="bf16x3")
    _, _, eng_b, _, _ = build("rbf", precision="bf16x3")
    f, y = wl["f_train"].pin_memory(), wl["y_train"].pin_memory()
    nb = min(5, f.shape[0] // shp.B)
    batches = [(i * shp.B, (i + 1) * shp.B) for i in range(nb)]
    ref = [float(eng_a.train_step(f[lo:hi].cuda(), y[lo:hi].cuda())) for lo, hi in batches]
    losses = eng_b.train_steps_host(f, y, batches)
    assert losses.tolist() == pytest.approx(ref, rel=1e-5)
    assert within(eng_b.flat_p, eng_a.flat_p, 1e-5)
    with pytest.raises(ValueError):
        eng_b.train_steps_host(wl["f_train"], wl["y_train"], batches)          # pageable host memory
