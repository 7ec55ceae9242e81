"""GPU: GP template weighter kernels (forward + hand-derived adjoint) against the CPU oracle."""
import os

import numpy as np
import pytest
import torch

from clip_gp_b200 import ops, synth
from clip_gp_b200.gp_template_weigher import GaussianProcessTemplateWeighter
from oracle import gp as ogp
from tests.helpers import assert_parity, make_state, max_err, oracle_grad_pair, oracle_grads, oracle_pair, rel_err, within

pytestmark = pytest.mark.gpu
KERNELS = ["rbf", "matern", "linear"]
TOL = 1e-3          # BASELINE.json: GP weights within 1e-3 relative in fp32


def run_kernel(st, eps, kernel, alias_check=True, need_grad=False, X=None):
    dev = "cuda"
    kp = st.kernel
    t = lambda x: None if x is None else x.detach().clone().to(dev).requires_grad_(need_grad)
    Z, m, chol = t(st.inducing_points), t(st.var_mean), t(st.chol_var)
    ls, os_, var = t(kp.raw_lengthscale), t(kp.raw_outputscale), t(kp.raw_variance)
    Xd = (st.templates_red if X is None else X).to(dev)
    C, T = st.templates_red.shape[:2]
    n = st.inducing_points.shape[1]
    mean_x = ogp.residual_mean(st.f0, st.cls_bias, st.tmp_bias, n + T)[:, n:].contiguous().to(dev)
    w, kl, status = ops.gp_weights(Z, Xd, ls, os_, var, m, chol, mean_x, eps.to(dev), kernel, eps.shape[2],
                                   alias_check=alias_check)
    return w, kl, status, dict(Z=Z, m=m, chol=chol, ls=ls, os=os_, var=var)


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("name", ["tiny", "small", "cfg1", "t32", "t31", "t40", "t64"])
def test_forward_matches_oracle(kernel, name):
    wl, st = make_state(name, kernel)
    shp = wl["shape"]
    eps = torch.randn(shp.C, shp.T, shp.S, generator=torch.Generator().manual_seed(21))
    w32, _, w64, _ = oracle_pair(st, eps)
    kl_ref = ogp.kl_divergence(st.var_mean, st.chol_var)
    w, kl, status, _ = run_kernel(st, eps, kernel)
    assert int(status.abs().max()) == 0
    assert_parity(w, w32, w64, rtol=TOL, name="w")                     # vs the reference's fp32 arithmetic (pinned oracle)
    assert rel_err(w, w64) < 2 * TOL          # and close to exact arithmetic (Sigma is factorised in fp32, as the reference does)
    assert within(kl, kl_ref, 1e-5)
    assert float((w.sum(-1) - 1).abs().max()) < 1e-5 and float(w.min()) >= 0
    # the general three-block path (no aliasing of X with Z[:T]) gives the same answer
    w2, _, _, _ = run_kernel(st, eps, kernel, alias_check=False)
    assert rel_err(w2, w) < TOL and max_err(w2, w) < 1e-4


@pytest.mark.parametrize("kernel", KERNELS)
def test_forward_general_inputs(kernel):
    """X different from the inducing rows (e.g. weight decay moved Z): oracle parity on the general path."""
    wl, st = make_state("small", kernel)
    shp = wl["shape"]
    g = torch.Generator().manual_seed(4)
    st.inducing_points = st.inducing_points + 0.02 * torch.randn(st.inducing_points.shape, generator=g)
    eps = torch.randn(shp.C, shp.T, shp.S, generator=g)
    w32, _, w64, _ = oracle_pair(st, eps)
    w, _, status, _ = run_kernel(st, eps, kernel)         # alias check on, but rows differ -> general path
    assert int(status.abs().max()) == 0
    assert_parity(w, w32, w64, rtol=TOL, name="w")
    assert rel_err(w, w64) < TOL


@pytest.mark.parametrize("kernel", KERNELS)
def test_selfgolden_fixture(golden_dir, kernel):
    g = np.load(os.path.join(golden_dir, "gp_selfgolden.npz"))
    wl = synth.make_workload("tiny"); shp = wl["shape"]
    st = ogp.build_state(wl["E"], kernel, shp.d)
    st.var_mean, st.chol_var = synth.trained_like_q(shp.C, shp.T + 1, 5)
    eps = torch.randn(shp.C, shp.T, shp.S, generator=torch.Generator().manual_seed(77))
    w, kl, _, _ = run_kernel(st, eps, kernel)
    _, _, w64, P64 = oracle_pair(st, eps)
    assert_parity(w, torch.from_numpy(g[f"{kernel}/w"]), w64, rtol=TOL, name="w")
    assert within(kl, torch.from_numpy(g[f"{kernel}/kl"]), 1e-5)
    P = ops.prototypes(w, st.templates.cuda())
    assert_parity(P, torch.from_numpy(g[f"{kernel}/protos"]), P64, rtol=TOL, name="prototypes")


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("name", ["tiny", "small", "t32", "t31", "t40", "t64"])
@pytest.mark.parametrize("alias", [True, False])
def test_adjoint_matches_oracle_autograd(kernel, name, alias):
    wl, st = make_state(name, kernel)
    shp = wl["shape"]
    g = torch.Generator().manual_seed(11)
    eps = torch.randn(shp.C, shp.T, shp.S, generator=g)
    dw = torch.randn(shp.S, shp.C, shp.T, generator=g)
    dkl = torch.rand(shp.C, generator=g)
    G32, G = oracle_grad_pair(st, eps, dw, dkl)                     # fp32 (reference arithmetic) and float64 autograd through the oracle
    w, kl, _, P = run_kernel(st, eps, kernel, alias_check=alias, need_grad=True)
    loss = (w * dw.cuda()).sum() + (kl * dkl.cuda()).sum()
    loss.backward()
    assert float(P["Z"].grad[:, :-1].abs().max()) == 0.0            # frozen template rows (:72-79)
    assert float(P["chol"].grad.triu(1).abs().max()) == 0.0
    # elementwise parity gate against the reference-arithmetic gradient (budget widened by ITS OWN distance to float64) ...
    assert_parity(P["Z"].grad[:, -1], G32["Z"][:, -1], G["Z"][:, -1], rtol=2e-3, name="dZ_last")
    for k_ in ("m", "chol", "ls", "os", "var"):
        if k_ in G:
            assert_parity(P[k_].grad, G32[k_], G[k_], rtol=2e-3, name="d" + k_)
    # ... and norm-wise against exact arithmetic
    assert max_err(P["Z"].grad[:, -1], G["Z"][:, -1]) < 5e-3
    for k_ in ("m", "chol", "ls", "os", "var"):
        if k_ in G:
            assert max_err(P[k_].grad, G[k_]) < TOL


def test_max_size_T64_n65_matern():
    """cfg5 shape per class (T=64, n=65, Matern) on a handful of classes, forward + adjoint."""
    g = torch.Generator().manual_seed(8)
    C, T, D, d, S = 6, 64, 512, 256, 7
    E, _ = synth.make_text_bank(C, T, D, 99)
    st = ogp.build_state(E, "matern", d)
    st.var_mean, st.chol_var = synth.trained_like_q(C, T + 1, 3)
    eps = torch.randn(C, T, S, generator=g)
    dw = torch.randn(S, C, T, generator=g); dkl = torch.rand(C, generator=g)
    G32, G = oracle_grad_pair(st, eps, dw, dkl)
    w32, _, w64, _ = oracle_pair(st, eps)
    w, kl, status, P = run_kernel(st, eps, "matern", need_grad=True)
    assert int(status.abs().max()) == 0
    assert_parity(w, w32, w64, rtol=TOL, name="w")
    assert rel_err(w, w64) < 2 * TOL
    ((w * dw.cuda()).sum() + (kl * dkl.cuda()).sum()).backward()
    for k_ in ("m", "chol", "ls"):
        assert_parity(P[k_].grad, G32[k_], G[k_], rtol=2e-3, name="d" + k_)
        assert max_err(P[k_].grad, G[k_]) < TOL


def test_identity_q_and_jitter_retry_status():
    wl, st = make_state("small", "rbf", trained=False, perturb=False)
    shp = wl["shape"]
    eps = torch.randn(shp.C, shp.T, shp.S, generator=torch.Generator().manual_seed(6))
    w32, _, w64, _ = oracle_pair(st, eps)
    w, kl, status, _ = run_kernel(st, eps, "rbf")
    assert_parity(w, w32, w64, rtol=TOL, name="w")
    assert float(kl.abs().max()) < 1e-6 and int(status.abs().max()) == 0
    # make Sigma indefinite for one class: L_q = 0 => Sigma = K_XX + jI - A^T A ~ jitter-level, may need retries;
    # the status channel must report it instead of silently returning garbage
    st.chol_var = st.chol_var.clone(); st.chol_var[0] = 0.0; st.chol_var[0].diagonal().fill_(1e-3)
    _, _, status, _ = run_kernel(st, eps, "rbf")
    assert int(status[1:].abs().max()) == 0


def test_philox_stream_is_reproducible_and_shardable():
    wl, st = make_state("small", "rbf")
    shp = wl["shape"]
    dev = "cuda"
    kp = st.kernel
    args = [x.to(dev) for x in (st.inducing_points, st.templates_red, kp.raw_lengthscale, kp.raw_outputscale)]
    m, chol = st.var_mean.to(dev), st.chol_var.to(dev)
    rng = torch.tensor([1234, 7], dtype=torch.int64, device=dev)
    S = 8
    full, _, _ = ops.gp_weights(*args, None, m, chol, None, None, "rbf", S, rng_state=rng, s_offset=0, S_total=S)
    again, _, _ = ops.gp_weights(*args, None, m, chol, None, None, "rbf", S, rng_state=rng, s_offset=0, S_total=S)
    assert torch.equal(full, again)
    lo, _, _ = ops.gp_weights(*args, None, m, chol, None, None, "rbf", 3, rng_state=rng, s_offset=0, S_total=S)
    hi, _, _ = ops.gp_weights(*args, None, m, chol, None, None, "rbf", 5, rng_state=rng, s_offset=3, S_total=S)
    assert torch.equal(torch.cat([lo, hi], 0), full)                # sharding S never changes the draws
    rng2 = torch.tensor([1234, 8], dtype=torch.int64, device=dev)
    other, _, _ = ops.gp_weights(*args, None, m, chol, None, None, "rbf", S, rng_state=rng2, s_offset=0, S_total=S)
    assert not torch.equal(other, full)


class _Cfg:
    class adapter:
        gp_pca_dim = 32
        gp_kernel_type = "rbf"


@pytest.mark.parametrize("kernel", KERNELS)
def test_module_drop_in_surface(kernel):
    """Same constructor / methods / state_dict names as the reference class; differentiable end to end."""
    wl = synth.make_workload("small"); shp = wl["shape"]
    cfg = _Cfg(); cfg.adapter.gp_kernel_type = kernel
    torch.manual_seed(0)
    gpw = GaussianProcessTemplateWeighter(text_embeddings=wl["E"], cfg=cfg).to("cuda")
    keys = set(gpw.state_dict().keys())
    must = {"variational_strategy.inducing_points", "variational_strategy._variational_distribution.variational_mean",
            "variational_strategy._variational_distribution.chol_variational_covar", "mean_module.f0",
            "mean_module.cls_bias", "mean_module.tmp_bias", "mean_module.neg_tail", "A.weight",
            "likelihood.noise_covar.raw_noise", "_ind_mask", "_templates", "_templates_red", "_cls_mean_init"}
    must |= {"rbf": {"covar_module.raw_outputscale", "covar_module.base_kernel.raw_lengthscale"},
             "matern": {"covar_module.raw_lengthscale"}, "linear": {"covar_module.raw_variance"}}[kernel]
    assert must <= keys
    gpw.train()
    protos = gpw.sample_prototypes(num_samples=shp.S)
    assert protos.shape == (shp.S, shp.C, shp.D) and gpw.scores.shape == (shp.S, shp.C, shp.T)
    kl = gpw.variational_strategy.kl_divergence()
    assert kl.shape == (shp.C,)
    (protos.pow(2).sum() + 0.01 * kl.sum()).backward()
    ip = gpw.variational_strategy.inducing_points
    assert ip.grad is not None and float(ip.grad[:, :-1].abs().max()) == 0.0
    q = gpw.variational_strategy._variational_distribution
    assert q.variational_mean.grad is not None and torch.isfinite(q.chol_variational_covar.grad).all()
    # oracle parity of the whole module given the same parameters and noise
    st = ogp.build_state(wl["E"], kernel, 32)
    st.inducing_points = ip.detach().cpu(); st.var_mean = q.variational_mean.detach().cpu()
    st.chol_var = q.chol_variational_covar.detach().cpu()
    if kernel == "rbf":
        st.kernel.raw_lengthscale = gpw.covar_module.base_kernel.raw_lengthscale.detach().cpu()
    eps = torch.randn(shp.C, shp.T, 3, generator=torch.Generator().manual_seed(1))
    _, P_ref, _, P64 = oracle_pair(st, eps)
    with torch.no_grad():
        P = gpw.sample_prototypes(3, eps=eps.cuda())
        Pm = gpw.mean_prototypes(3, eps=eps.cuda())
        Pc = gpw.collapsed_prototypes(3, eps=eps.cuda())
    assert_parity(P, P_ref, P64, rtol=TOL, name="prototypes")
    nrm = lambda x: x / x.norm(dim=-1, keepdim=True)
    assert_parity(Pm, nrm(P_ref.mean(0)), nrm(P64.mean(0)), rtol=TOL, name="mean prototypes")
    assert_parity(Pc, nrm(P_ref).mean(0), nrm(P64).mean(0), rtol=TOL, name="collapsed prototypes")
    # initialize_from_weights is the reference's no-op for T > 1
    before = q.variational_mean.detach().clone()
    gpw.initialize_from_weights(torch.full((shp.C, shp.T), 1.0 / shp.T, device="cuda"))
    assert torch.equal(before, q.variational_mean.detach())
    with pytest.raises(ValueError):
        cfg.adapter.gp_kernel_type = "periodic"
        GaussianProcessTemplateWeighter(text_embeddings=wl["E"], cfg=cfg)


@pytest.mark.parametrize("N,d", [(40, 16), (333, 32), (2048, 64), (3001, 256)])
def test_median_lengthscale_radix_select(N, d):
    """gp_template_weigher.py:103-107 without the N x N matrix: the three-pass radix select returns the element of the positive
    OFF-DIAGONAL pairwise distances whose rank is the lower median.  The reference's `pdist > 0` also keeps those diagonal entries
    that torch's matmul-form cdist leaves at a rounding-noise value (~3e-4) instead of 0, which moves its rank by at most N/2 of
    ~N^2 (7e-7 relative at the ImageNet shape, profiles/r1_setup_kernels.txt); the oracle must therefore sit within N ranks."""
    g = torch.Generator().manual_seed(N + d)
    X = torch.nn.functional.normalize(torch.randn(N, d, generator=g), dim=-1)
    X[5] = X[3]                                            # an exact duplicate: its zero distance must not be counted
    got = ops.median_pairwise_distance(X.cuda())
    D = torch.cdist(X.double(), X.double())
    off = ~torch.eye(N, dtype=torch.bool)
    pos = D[off & (D > 1e-7)]
    k = (pos.numel() - 1) // 2
    below = int((pos < got - 1e-7).sum()); not_above = int((pos <= got + 1e-7).sum())
    # `got` is the k-th smallest up to the fp32 evaluation of the distances; the duplicated row may contribute two rounding-noise
    # "positive" distances at the bottom of the list (the matmul form |a|^2+|b|^2-2ab does not return an exact 0 for it)
    assert below - 2 <= k < not_above + 2
    ref = ogp.median_lengthscale(X.view(1, N, d))
    assert abs(int((pos < ref).sum()) - k) <= N + 2


def test_module_uses_the_setup_kernel_for_the_rbf_lengthscale():
    wl = synth.make_workload("small"); shp = wl["shape"]
    cfg = _Cfg(); cfg.adapter.gp_kernel_type = "rbf"
    gpw = GaussianProcessTemplateWeighter(text_embeddings=wl["E"].cuda(), cfg=cfg)
    st = ogp.build_state(wl["E"], "rbf", 32)
    ls = float(gpw.covar_module.base_kernel.lengthscale.flatten()[0])
    ls_ref = float(torch.nn.functional.softplus(st.kernel.raw_lengthscale).flatten()[0])
    assert ls == pytest.approx(ls_ref, rel=1e-3)           # N = 296 points: the diagonal-noise rank ambiguity is ~1/N


@pytest.mark.parametrize("alias", [True, False])
@pytest.mark.parametrize("name", ["small", "t32"])
def test_philox_adjoint_uses_the_forward_draw(name, alias):
    """Counter-RNG mode: the caller advances (seed, step) right after the forward launch (sample_weights does); the adjoint must
    still differentiate the draw the forward used.  Gradients equal those of the explicit-eps call with oracle/philox's stream."""
    from oracle import philox
    wl, st = make_state(name, "rbf")
    shp = wl["shape"]
    g = torch.Generator().manual_seed(12)
    dw = torch.randn(shp.S, shp.C, shp.T, generator=g).cuda()
    seed, step = 11, 4
    eps = philox.eps_tensor(seed, step, shp.C, shp.T, shp.S)
    w_ref, kl_ref, _, P_ref = run_kernel(st, eps, "rbf", alias_check=alias, need_grad=True)
    (w_ref * dw).sum().backward()
    dev = "cuda"
    kp = st.kernel
    t = lambda x: x.detach().clone().to(dev).requires_grad_(True)
    Z, m, chol, ls, os_ = t(st.inducing_points), t(st.var_mean), t(st.chol_var), t(kp.raw_lengthscale), t(kp.raw_outputscale)
    n = st.inducing_points.shape[1]
    mean_x = ogp.residual_mean(st.f0, st.cls_bias, st.tmp_bias, n + shp.T)[:, n:].contiguous().to(dev)
    rng = torch.tensor([seed, step], dtype=torch.int64, device=dev)
    w, kl, status = ops.gp_weights(Z, st.templates_red.to(dev), ls, os_, None, m, chol, mean_x, None, "rbf", shp.S, rng_state=rng,
                                   alias_check=alias)
    rng[1] += 1                                            # what GaussianProcessTemplateWeighter.sample_weights does
    torch.cuda.synchronize()
    assert max_err(w, w_ref) < 1e-5                        # device Philox normals vs the numpy restatement: last-bit differences
    (w * dw).sum().backward()
    for a, b in ((Z, P_ref["Z"]), (m, P_ref["m"]), (chol, P_ref["chol"]), (ls, P_ref["ls"]), (os_, P_ref["os"])):
        assert max_err(a.grad, b.grad) < 1e-3              # (a step + 1 draw would be O(1) away)


def test_module_philox_mode_backward_and_kl_cache():
    """rng='philox' through the module: gradients equal the explicit-eps module call; kl_divergence() after a consumed graph
    re-launches instead of handing out a tensor whose graph was freed."""
    from oracle import philox
    wl = synth.make_workload("small"); shp = wl["shape"]
    cfg = type("Cfg", (), {"adapter": type("A", (), {"gp_pca_dim": shp.d, "gp_kernel_type": "rbf"})()})()
    mods = []
    for rng in ("philox", "torch"):
        torch.manual_seed(0)
        gp = GaussianProcessTemplateWeighter(wl["E"], cfg, rng=rng, seed=9, lengthscale=1.3).cuda()
        gp.variational_strategy._maybe_init()
        mods.append(gp)
    a, b = mods
    eps = philox.eps_tensor(9, 0, shp.C, shp.T, shp.S).cuda()
    Pa = a.sample_prototypes(shp.S)
    Pb = b.sample_prototypes(shp.S, eps=eps)
    assert max_err(Pa, Pb) < 1e-5 and int(a._rng_state[1]) == 1
    (Pa.pow(2).sum() + a.variational_strategy.kl_divergence().sum()).backward()
    (Pb.pow(2).sum() + b.variational_strategy.kl_divergence().sum()).backward()
    for (na, pa), (_, pb) in zip(a.named_parameters(), b.named_parameters()):
        if pa.grad is not None and "mean_module" not in na:         # the mean module's gradients are rounding noise around 0
            assert max_err(pa.grad, pb.grad) < 1e-3, na
    # the graph of the last launch is gone now: a new kl call must be differentiable again
    kl = a.variational_strategy.kl_divergence().sum()
    kl.backward()


@pytest.mark.parametrize("C,T,D,d", [(40, 8, 64, 16), (100, 32, 512, 256), (5, 3, 64, 256)])
def test_device_pca_spans_the_reference_subspace(C, T, D, d):
    """gp_template_weigher.py:26-37 on the device (Gram matrix by the clipgp GEMM + symmetric eigen-solve) against the reference's
    thin SVD on the CPU: same reduced dimension, same principal subspace (projector), same projected geometry."""
    E, _ = synth.make_text_bank(C, T, D, seed=C + T)
    X = E.reshape(-1, D)
    Xc = X - X.mean(0, keepdim=True)
    _, Sv, Vt = torch.linalg.svd(Xc, full_matrices=False)
    dd = min(d, Vt.shape[0])
    W_ref = Vt[:dd].T
    W = GaussianProcessTemplateWeighter._pca_axes_device(Xc.cuda(), dd).cpu()
    assert W.shape == W_ref.shape
    assert float((W.t() @ W - torch.eye(dd)).abs().max()) < 1e-4                       # orthonormal axes
    # compare on the numerically non-degenerate part of the spectrum (the centred bank has rank <= C*T - 1)
    keep = int((Sv[:dd] > 1e-4 * Sv[0]).sum())
    Pr = W_ref[:, :keep] @ W_ref[:, :keep].t()
    Pg = W[:, :keep] @ W[:, :keep].t()
    if keep == dd or float(Sv[keep - 1] / Sv[min(keep, len(Sv) - 1)]) > 1.01:          # a spectral gap at the cut makes the subspace unique
        assert float((Pr - Pg).abs().max()) < 2e-3
    Zr, Zg = Xc @ W_ref, Xc @ W
    assert max_err(Zg @ Zg.t(), Zr @ Zr.t()) < 1e-3                                    # what the kernels consume: inner products of the reduced points
