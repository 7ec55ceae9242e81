"""GPU: BASELINE.json's configurations at FULL size.  The CPU oracle cannot run them whole in seconds, so each test combines
(1) size-independent properties of the outputs and (2) exact-path parity with the oracle on a subset that the oracle can afford:
the GP is independent per class (subset of classes), logits / cache affinities are independent per image (subset of rows)."""
import copy
import dataclasses

import pytest
import torch
import torch.nn.functional as F

from clip_gp_b200 import dist as cd
from clip_gp_b200 import metrics as gm
from clip_gp_b200 import ops, synth, tc
from clip_gp_b200.engine import EngineConfig, GPAdapterEngine
from clip_gp_b200.gp_template_weigher import GaussianProcessTemplateWeighter
from oracle import gp as ogp
from oracle import heads as oh
from oracle import metrics as om
from oracle import philox
from tests.helpers import assert_parity, fix_eval_noise, max_err, oracle_grad_pair, oracle_grads, oracle_pair, rel_err, state_to, within

pytestmark = pytest.mark.gpu


class _Cfg:
    def __init__(self, kernel, pca):
        self.adapter = type("A", (), {"gp_pca_dim": pca, "gp_kernel_type": kernel})()


def _class_subset(st, idx):
    """Oracle GPState restricted to the classes `idx` (the GP never couples classes)."""
    C = st.templates.shape[0]
    kw = {}
    for fld in dataclasses.fields(st):
        v = getattr(st, fld.name)
        if fld.name == "kernel":
            kp = copy.copy(v)
            for nm in ("raw_lengthscale", "raw_outputscale", "raw_variance"):
                t = getattr(kp, nm)
                if t is not None:
                    setattr(kp, nm, t[idx].clone())
            kw[fld.name] = kp
        elif torch.is_tensor(v) and v.dim() >= 1 and v.shape[0] == C and fld.name != "pca_W":
            kw[fld.name] = v[idx].clone()
        else:
            kw[fld.name] = v
    return type(st)(**kw)


def _engine(name, kernel, S, precision="bf16x3", lengthscale=None, **kw):
    wl = synth.make_workload(name, n_test=kw.pop("n_test", None)); shp = wl["shape"]
    torch.manual_seed(0)
    # built on the CPU like the oracle (same SVD -> same PCA basis), then moved
    gpw = GaussianProcessTemplateWeighter(wl["E"], _Cfg(kernel, shp.d), lengthscale=lengthscale).cuda()
    m, Lq = synth.trained_like_q(shp.C, shp.T + 1, 5)
    gpw.variational_strategy._maybe_init()
    q = gpw.variational_strategy._variational_distribution
    with torch.no_grad():
        q.variational_mean.copy_(m); q.chol_variational_covar.copy_(Lq)
    eng = GPAdapterEngine(gpw, EngineConfig(S_train=S, S_eval=S, batch_size=shp.B, shots=shp.shots, seed=11, precision=precision, **kw))
    # the oracle twin (same PCA, same parameters)
    st = ogp.build_state(wl["E"], kernel, shp.d, lengthscale=lengthscale)
    st.inducing_points = gpw.variational_strategy.inducing_points.detach().cpu().clone()
    st.templates_red = gpw._templates_red.detach().cpu().clone()
    st.var_mean, st.chol_var = m.clone(), Lq.clone()
    raw_ls, raw_os, raw_var = gpw._kernel_raw()
    if raw_ls is not None: st.kernel.raw_lengthscale = raw_ls.detach().cpu().clone()
    if raw_os is not None: st.kernel.raw_outputscale = raw_os.detach().cpu().clone()
    if raw_var is not None: st.kernel.raw_variance = raw_var.detach().cpu().clone()
    return wl, shp, eng, st


def test_cfg2_train_step_full_size_subset_parity():
    """ImageNet shape (C=1000, T=32, D=512, d=256, S=10, B=128): one engine step; the template weights and the GP parameter
    gradients of 41 classes spread over the class range against float64 autograd through the oracle, driven by the engine's own dw."""
    wl, shp, eng, st = _engine("cfg2", "rbf", 10, lengthscale=1.4146)
    f, y = wl["f_train"][: shp.B].cuda(), wl["y_train"][: shp.B].cuda()
    eng.skip_update = True
    loss = eng.train_step(f, y, use_graph=False)
    torch.cuda.synchronize()
    assert torch.isfinite(loss).all() and int(eng.status.abs().max()) == 0
    w = eng.w.cpu()                                                    # [S, C, T]
    assert float((w.sum(-1) - 1).abs().max()) < 1e-5 and float(w.min()) >= 0.0
    assert torch.isfinite(eng.flat_g).all()
    idx = torch.unique(torch.cat([torch.tensor([0, 1, 17, 500, 998, 999]), torch.arange(5, shp.C, 29)]))   # 41 classes across every SM's slots
    sub = _class_subset(st, idx)
    eps = philox.eps_tensor(11, 0, shp.C, shp.T, 10)[idx]
    dw = eng.dw.cpu()[:, idx]
    dkl = torch.full((len(idx),), eng.cfg.gp_beta)
    G32, G = oracle_grad_pair(sub, eps, dw, dkl)                      # reference arithmetic (fp32) / exact (float64)
    w32, _, w64, _ = oracle_pair(sub, eps)
    n = shp.T + 1
    assert_parity(w[:, idx], w32, w64, name="w")
    assert rel_err(w[:, idx], w64) < max(1e-3, 2.0 * rel_err(w32, w64))   # vs EXACT arithmetic: no further off than the reference's own fp32 path
    got = {"m": eng.g("m").view(shp.C, n).cpu()[idx], "chol": eng.g("Lq").view(shp.C, n, n).cpu()[idx],
           "ls": eng.g("ls").view(shp.C, 1, -1).cpu()[idx], "os": eng.g("os").cpu()[idx]}
    for k_, g_ in got.items():
        assert_parity(g_, G32[k_], G[k_], rtol=3e-3, name="d" + k_)     # bf16x3 GEMMs feed dw: 3e-3 (the reference's GPU path is TF32)
        assert max_err(g_, G[k_]) < 2e-3, k_
    assert_parity(eng.g("z_last").view(shp.C, -1).cpu()[idx], G32["Z"][:, -1], G["Z"][:, -1], rtol=3e-3, name="dz_last")
    assert max_err(eng.g("z_last").view(shp.C, -1).cpu()[idx], G["Z"][:, -1]) < 2e-2


def test_cfg3_eval_full_size_properties():
    """50 000 ImageNet-shaped test features: tensor-core eval (split operands, fused calibration) against the exact fp32 path of
    the same engine; counters are shard-invariant bit for bit; a row subset of the logits against the oracle."""
    wl, shp, eng, st = _engine("cfg3", "rbf", 10, lengthscale=1.4146)
    f, y = wl["f_test"].cuda(), wl["y_test"].cuda()
    N = f.shape[0]
    assert N == 50000
    fix_eval_noise(eng, 10)                                 # one draw for every eval call of this test (seed 11, step 0)
    conf, correct, hist = eng.eval_calibration_tc(f, y, precision="bf16x3", mc="collapsed")
    cnt = gm.counters_from_hist(hist, N)
    ece, bins = gm.ece_from_counters(cnt)
    assert sum(bins["bin_count"]) == N and int(hist[3, 0]) == int(correct.sum())
    # exact-mode comparator (fp32 FFMA GEMMs, logits materialised): same top-1 up to near-ties, same ECE to 1e-3
    ref = eng.evaluate(f, y, precision="fp32")
    logits = eng.eval_logits(f)
    top2 = logits.topk(2, dim=1).values
    fragile = int(((top2[:, 0] - top2[:, 1]) < 1e-3).sum())
    assert abs(ref["top1_count"] - cnt.top1) <= fragile
    assert ece == pytest.approx(ref["ece"], rel=1e-3, abs=1e-3)
    # shards: 8 contiguous image shards, integer counters summed == single pass, bit for bit
    tot = torch.zeros_like(hist)
    for r in range(8):
        lo, hi = cd.shard_range(N, r, 8)
        _, _, h = eng.eval_calibration_tc(f[lo:hi], y[lo:hi], precision="bf16x3", mc="collapsed")
        tot += h
    assert torch.equal(tot, hist)
    # oracle on a row subset: MC-averaged logits (adapter.py:243-249) with the same Philox noise
    rows = torch.arange(0, N, 997)
    eps = philox.eps_tensor(11, 0, shp.C, shp.T, 10)
    protos, _ = ogp.sample_prototypes(st, eps)
    lg_ref = oh.adapter_logits(wl["f_test"][rows], torch.eye(shp.D), protos, 100.0)
    assert float((logits[rows.cuda()].cpu() - lg_ref).abs().max()) < 1e-3 * float(lg_ref.abs().max())


def test_cfg4_tip_adapter_full_size():
    """Tip-Adapter-F shape: 16 000 cache keys, D=1024, C=1000, B=128: exact fp32 kernels and the fused tcgen05 form against the
    oracle's one-hot formulation (tip_adapter.py:250-251) on the whole batch; key gradient on a sample of keys."""
    g = torch.Generator().manual_seed(4)
    B, N_tr, C, D = 128, 16000, 1000, 1024
    mu = torch.randn(C, D, generator=g)
    lab = torch.arange(C).repeat_interleave(16)
    keys = F.normalize(mu[lab] + 2.0 * torch.randn(N_tr, D, generator=g), dim=-1)
    yb = torch.randint(0, C, (B,), generator=g)
    f = F.normalize(mu[yb] + 2.0 * torch.randn(B, D, generator=g), dim=-1)
    clip = 100.0 * f @ F.normalize(mu, dim=-1).t()
    dout = torch.randn(B, C, generator=g)
    kr = keys.clone().requires_grad_(True)
    ref = oh.tip_logits(f, kr, oh.tip_cache_vals(lab, C), clip, 2.0, 20.0)
    ref.backward(dout)
    kd = keys.cuda().requires_grad_(True)
    out = ops.tip_logits(f.cuda(), kd, lab.cuda(), clip.cuda(), 2.0, 20.0, C)
    out.backward(dout.cuda())
    assert within(out, ref, 1e-5)
    pick = torch.arange(0, N_tr, 37)
    assert within(kd.grad.cpu()[pick], kr.grad[pick], 1e-4)
    from clip_gp_b200 import _lib
    o3 = clip.cuda().clone()
    fa, kb, li = tc.cast_bf16(f.cuda(), tc.SPLIT_A), tc.cast_bf16(keys.cuda(), tc.SPLIT_B), lab.to(torch.int32).cuda().contiguous()
    _lib.check(_lib.load().clipgp_tc_tip_logits(fa.data_ptr(), B, kb.data_ptr(), N_tr, 3 * D, li.data_ptr(), 2.0, 20.0, o3.data_ptr(), C,
                                                _lib.stream_ptr(o3.device)), "clipgp_tc_tip_logits")
    assert rel_err(o3, ref) < 1e-3
    assert int((o3.argmax(1).cpu() == ref.argmax(1)).sum()) >= B - 1


def test_cfg5_matern_T64_S100_full_size():
    """SUN397 shape (C=397, T=64, n=65, S=100, Matern-1/2): forward of all classes on the general block kernel; properties for all,
    oracle parity for a few classes; the MC-mean prototypes that initialise TaskRes (taskres.py:281-285) are unit rows."""
    wl = synth.make_workload("cfg5"); shp = wl["shape"]
    st = ogp.build_state(wl["E"], "matern", shp.d)
    st.var_mean, st.chol_var = synth.trained_like_q(shp.C, shp.T + 1, 3)
    S = 100
    eps = torch.randn(shp.C, shp.T, S, generator=torch.Generator().manual_seed(5))
    dev = "cuda"
    n = shp.T + 1
    mean_x = ogp.residual_mean(st.f0, st.cls_bias, st.tmp_bias, n + shp.T)[:, n:].contiguous()
    w, kl, status = ops.gp_weights(st.inducing_points.to(dev), st.templates_red.to(dev), st.kernel.raw_lengthscale.to(dev), None, None,
                                   st.var_mean.to(dev), st.chol_var.to(dev), mean_x.to(dev), eps.to(dev), "matern", S)
    assert w.shape == (S, shp.C, shp.T) and int(status.abs().max()) == 0
    assert float((w.sum(-1) - 1).abs().max()) < 1e-5 and float(w.min()) >= 0.0
    idx = torch.tensor([0, 7, 200, 396])
    w_ref, _, w64, _ = oracle_pair(_class_subset(st, idx), eps[idx])
    assert_parity(w[:, idx], w_ref, w64, name="w")
    assert rel_err(w[:, idx], w64) < 2e-3
    assert within(kl.cpu()[idx], ogp.kl_divergence(st.var_mean[idx], st.chol_var[idx]), 1e-5)
    _, _, base = ops.prototypes_reduced(w, st.templates.to(dev), want_mean_raw=True)
    assert float((base.norm(dim=-1) - 1).abs().max()) < 1e-5
    P_ref = torch.einsum("skm,kmd->skd", w_ref, st.templates[idx]).mean(0)
    P_64 = torch.einsum("skm,kmd->skd", w64, st.templates[idx].double()).mean(0)
    assert_parity(base[idx.to(dev)], F.normalize(P_ref, dim=-1), F.normalize(P_64, dim=-1), name="mean prototypes")


def test_cfg5_adjoint_full_size_against_float64_autograd():
    """The same shape through the adjoint (wide-CTA general kernels, blocked factorisations): ALL 397 classes run, the gradients of a
    class subset are compared with float64 autograd through the oracle."""
    wl = synth.make_workload("cfg5"); shp = wl["shape"]
    st = ogp.build_state(wl["E"], "matern", shp.d)
    st.var_mean, st.chol_var = synth.trained_like_q(shp.C, shp.T + 1, 3)
    S = 20
    g = torch.Generator().manual_seed(6)
    eps = torch.randn(shp.C, shp.T, S, generator=g)
    dw = torch.randn(S, shp.C, shp.T, generator=g); dkl = torch.rand(shp.C, generator=g)
    dev, n = "cuda", shp.T + 1
    mean_x = ogp.residual_mean(st.f0, st.cls_bias, st.tmp_bias, n + shp.T)[:, n:].contiguous()
    t = lambda x: x.detach().clone().to(dev).requires_grad_(True)
    Z, ls, m, chol = t(st.inducing_points), t(st.kernel.raw_lengthscale), t(st.var_mean), t(st.chol_var)
    w, kl, status = ops.gp_weights(Z, st.templates_red.to(dev), ls, None, None, m, chol, mean_x.to(dev), eps.to(dev), "matern", S)
    assert int(status.abs().max()) == 0
    ((w * dw.to(dev)).sum() + (kl * dkl.to(dev)).sum()).backward()
    assert all(bool(torch.isfinite(x.grad).all()) for x in (Z, ls, m, chol))
    assert float(Z.grad[:, :-1].abs().max()) == 0.0 and float(chol.grad.triu(1).abs().max()) == 0.0
    idx = torch.tensor([0, 7, 200, 396])
    from tests.helpers import oracle_grad_pair
    G32, G64 = oracle_grad_pair(_class_subset(st, idx), eps[idx], dw[:, idx], dkl[idx])
    for name, got in (("m", m.grad), ("chol", chol.grad), ("ls", ls.grad)):
        assert_parity(got[idx.to(dev)], G32[name], G64[name], rtol=2e-3, name="d" + name)
        assert max_err(got[idx.to(dev)], G64[name]) < 1e-3, name
    assert max_err(Z.grad[idx.to(dev), -1], G64["Z"][:, -1]) < 5e-3


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
def test_cfg1_whole_step_and_eval_against_the_oracle(precision):
    """BASELINE configs[0] at full size (Caltech101 shape: C=100, T=8, D=1024, 4-shot, S=4, RBF) — small enough for the CPU oracle to
    run WHOLE: loss, every gradient of one optimisation step, and the MC-averaged eval logits / accuracy / ECE / AECE."""
    from oracle.train_step import OracleAdapter
    wl, shp, eng, st = _engine("cfg1", "rbf", 4, precision=precision, lengthscale=1.3)
    assert (shp.C, shp.T, shp.D, shp.S, shp.shots) == (100, 8, 1024, 4, 4)
    cfg = eng.cfg
    orc = OracleAdapter(st, shp.D, scale=cfg.logit_scale, gp_beta=cfg.gp_beta, l2_lambda=cfg.l2_lambda, shots=cfg.shots, lr=cfg.lr,
                        gp_lr=cfg.gp_lr, loss_mode="per_sample")
    f, y = wl["f_train"][: shp.B], wl["y_train"][: shp.B]
    eps = philox.eps_tensor(cfg.seed, 0, shp.C, shp.T, 4)
    orc64 = OracleAdapter(state_to(st, dtype=torch.float64), shp.D, scale=cfg.logit_scale, gp_beta=cfg.gp_beta, l2_lambda=cfg.l2_lambda,
                          shots=cfg.shots, lr=cfg.lr, gp_lr=cfg.gp_lr, loss_mode="per_sample")
    loss_ref = orc.loss(f, y, eps)
    loss_ref.backward()
    orc64.loss(f, y, eps).backward()
    eng.skip_update = True
    loss = eng.train_step(f.cuda(), y.cuda(), use_graph=False)
    n = shp.T + 1
    tol = 1e-3 if precision == "fp32" else 2e-3
    assert float(loss) == pytest.approx(float(loss_ref), rel=1e-3)
    assert int(eng.status.abs().max()) == 0
    for name, g32, g64 in (("W", orc.W.grad, orc64.W.grad), ("m", orc.st.var_mean.grad, orc64.st.var_mean.grad),
                           ("Lq", orc.st.chol_var.grad, orc64.st.chol_var.grad),
                           ("ls", orc.st.kernel.raw_lengthscale.grad, orc64.st.kernel.raw_lengthscale.grad),
                           ("os", orc.st.kernel.raw_outputscale.grad, orc64.st.kernel.raw_outputscale.grad)):
        got = eng.g(name).view(g32.shape)
        assert_parity(got, g32, g64, rtol=2e-3, name="d" + name)
        assert max_err(got, g64) < tol, name
    # eval over the whole test split: exact path vs oracle logits and the reference-pinned metrics oracle
    ft, yt = wl["f_test"], wl["y_test"]
    eps_e = fix_eval_noise(eng, 4)
    protos, _ = ogp.sample_prototypes(st, eps_e)
    lg_ref = oh.adapter_logits(ft, torch.eye(shp.D), protos, cfg.logit_scale)
    res = eng.evaluate(ft.cuda(), yt.cuda(), precision="fp32" if precision == "fp32" else "bf16x3")
    top1_ref = int((lg_ref.argmax(1) == yt).sum())
    top2 = lg_ref.topk(2, dim=1).values
    fragile = int(((top2[:, 0] - top2[:, 1]) < 1e-3 * float(lg_ref.abs().max())).sum())
    assert abs(res["top1_count"] - top1_ref) <= fragile
    assert res["ece"] == pytest.approx(om.compute_ece(lg_ref, yt), rel=1e-3, abs=2e-2)
    # AECE re-ranks the confidences: with 200 images per bin, one near-tie swapped across a rank edge moves it by 0.05 pp; the split
    # operands perturb confidences by ~1e-5, the exact path reproduces the order
    assert res["aece"] == pytest.approx(om.compute_aece(lg_ref, yt), rel=1e-3, abs=2e-2 if precision == "fp32" else 0.25)
