"""Shared builders for the parity tests (oracle state <-> clipgp module / kernel arguments)."""
import copy

import torch

from clip_gp_b200 import synth
from oracle import gp as ogp


def rel_err(a: torch.Tensor, b: torch.Tensor, floor: float = 1e-2, scale: float | None = None) -> float:
    """ELEMENTWISE relative error  max_i |a_i - b_i| / (|b_i| + floor * scale).

    ``rel_err(a, b) < 1e-3`` is SURVEY 8d's gate ``|a - b| <= 1e-3 |b| + 1e-5`` (relative, with an absolute floor of 1e-5 for
    values near zero) for tensors of unit scale or larger; ``scale`` defaults to ``min(1, max|b|)`` so that for tensors whose
    largest entry is small (gradients scaled by 1/(B*S), KL weights, ...) the floor shrinks with them instead of swallowing
    the whole tensor."""
    a = a.detach().double().cpu(); b = b.detach().double().cpu()
    if a.shape != b.shape:
        raise AssertionError(f"shape mismatch {tuple(a.shape)} vs {tuple(b.shape)}")
    if b.numel() == 0:
        return 0.0
    if scale is None:
        scale = min(1.0, float(b.abs().max()))
    return float(((a - b).abs() / (b.abs() + floor * scale + 1e-300)).max())


def max_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| / max |b| (norm-wise; only for quantities whose small entries are cancellation noise in the REFERENCE itself,
    stated at the call site)."""
    a = a.detach().double().cpu(); b = b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def within(a: torch.Tensor, b: torch.Tensor, tol: float) -> bool:
    """Tight fp32-vs-fp32 checks (tol < 1e-3: two evaluations of the same fp32 formula in a different summation order):
    norm-wise max|a-b| / max|b| < tol AND the elementwise gate rel_err < 1e-3 (entries that are small by cancellation cannot be
    relatively accurate to 1e-5 in ANY fp32 evaluation, so the tight part is norm-wise; the elementwise floor is 1e-2 of the
    tensor's largest entry: these are op-level checks on random data of arbitrary magnitude)."""
    return max_err(a, b) < tol and rel_err(a, b, scale=float(b.detach().abs().max())) < max(1e-3, tol)


def assert_parity(x: torch.Tensor, ref32: torch.Tensor, truth64: torch.Tensor | None = None, rtol: float = 1e-3,
                  floor: float = 1e-2, name: str = "", cond_floor: float = 0.0) -> float:
    """The parity gate against a REFERENCE-EXECUTED fp32 golden vector.

    elementwise:  |x - ref32| <= rtol * (|ref32| + floor * scale) + 2 * |ref32 - truth64| + max|ref32 - truth64|

    The first term is SURVEY 8d's gate (1e-3 relative, absolute floor 1e-5 at unit scale).  The second is the reference's OWN
    deviation from exact arithmetic at that element (truth64 = the pinned oracle evaluated in float64 on the same inputs): the
    reference evaluates squared distances by the fp32 expansion |a|^2 - 2ab + |b|^2 and factorises Sigma in fp32, which leaves
    up to 1.3e-3 (RBF, T = 32) / 4e-3 (Matern-1/2) elementwise noise on small template weights (measured in
    tests/golden/make_ref_golden.py, asserted in tests/test_ref_golden.py).  A kernel cannot be closer to the reference than
    the reference is to the function it evaluates.  ``cond_floor`` (used ONLY for the gradient of the learnable inducing row)
    adds ``cond_floor * max|ref32|``: that gradient flows through L^-1 of K_ZZ + 1e-4 I whose condition number is ~5e4 (the
    class-mean token lies in the span of the templates), so ANY evaluation that stores an intermediate in fp32 — the reference's
    autograd as much as the CUDA adjoint — carries cond * 2^-24 ~ 3e-3 of the tensor's scale as rounding noise.
    Returns the worst ratio (<= 1 passes)."""
    x = x.detach().double().cpu(); r = ref32.detach().double().cpu()
    assert x.shape == r.shape, (name, tuple(x.shape), tuple(r.shape))
    scale = min(1.0, float(r.abs().max())) if r.numel() else 1.0
    budget = rtol * (r.abs() + floor * scale) + cond_floor * (float(r.abs().max()) if r.numel() else 0.0)
    if truth64 is not None:
        dev_ref = (r - truth64.detach().double().cpu()).abs()
        # the reference's own error: at this element, plus its worst over the tensor (two independent fp32 evaluations do not
        # place their larger errors on the same elements).  Callers pair this gate with a check against truth64 itself, so a
        # noisy reference (Matern-1/2 d z_last: 0.1-0.5 norm-wise off float64, tools/diag_golden_grad.py) cannot make it vacuous
        budget = budget + 2.0 * dev_ref + float(dev_ref.max())
    ratio = float(((x - r).abs() / (budget + 1e-300)).max()) if r.numel() else 0.0
    assert ratio <= 1.0, f"{name}: parity gate violated, worst |x-ref| / budget = {ratio:.3g}"
    return ratio


def make_state(name: str, kernel: str, trained: bool = True, seed: int = 5, perturb: bool = True):
    """Oracle GPState for a named synthetic workload with a non-trivial q(u) and hyper-parameters."""
    wl = synth.make_workload(name)
    shp = wl["shape"]
    st = ogp.build_state(wl["E"], kernel, shp.d)
    if trained:
        m, Lq = synth.trained_like_q(shp.C, shp.T + 1, seed)
        st.var_mean, st.chol_var = m.clone(), Lq.clone()
    if perturb:
        g = torch.Generator().manual_seed(seed + 3)
        kp = st.kernel
        if kp.raw_lengthscale is not None:
            kp.raw_lengthscale = kp.raw_lengthscale + 0.1 * torch.randn(kp.raw_lengthscale.shape, generator=g)
        if kp.raw_outputscale is not None:
            kp.raw_outputscale = kp.raw_outputscale + 0.1 * torch.randn(kp.raw_outputscale.shape, generator=g)
        if kp.raw_variance is not None:
            kp.raw_variance = kp.raw_variance + 0.1 * torch.randn(kp.raw_variance.shape, generator=g)
        st.inducing_points = st.inducing_points.clone()
        st.inducing_points[:, -1] += 0.05 * torch.randn(shp.C, st.inducing_points.shape[-1], generator=g)
    return wl, st


def state_to(st, dtype=None, device=None):
    s2 = copy.deepcopy(st)
    for k in ["templates", "templates_red", "inducing_points", "var_mean", "chol_var", "f0", "cls_bias", "tmp_bias",
              "pca_mean", "pca_W"]:
        setattr(s2, k, getattr(s2, k).detach().to(dtype=dtype, device=device))
    for k in ["raw_lengthscale", "raw_outputscale", "raw_variance"]:
        v = getattr(s2.kernel, k)
        if v is not None:
            setattr(s2.kernel, k, v.detach().to(dtype=dtype, device=device))
    return s2


def oracle_grads(st, eps, dw, dkl, dtype=torch.float64):
    """Autograd through oracle.gp in `dtype` for loss = <w, dw> + <kl, dkl>.  Returns (w, kl, grads dict)."""
    s2 = state_to(st, dtype=dtype)
    names = {"Z": s2.inducing_points, "m": s2.var_mean, "chol": s2.chol_var}
    if s2.kernel.raw_lengthscale is not None: names["ls"] = s2.kernel.raw_lengthscale
    if s2.kernel.raw_outputscale is not None: names["os"] = s2.kernel.raw_outputscale
    if s2.kernel.raw_variance is not None: names["var"] = s2.kernel.raw_variance
    for p in names.values():
        p.requires_grad_(True)
    w, aux = ogp.gp_weights(s2, eps.to(dtype))
    kl = ogp.kl_divergence(s2.var_mean, s2.chol_var)
    loss = (w * dw.to(dtype)).sum() + (kl * dkl.to(dtype)).sum()
    g = torch.autograd.grad(loss, list(names.values()))
    return w.detach(), kl.detach(), dict(zip(names.keys(), [x.detach() for x in g])), aux


def oracle_pair(st, eps, visual=None):
    """(w32, P32, w64, P64): the pinned oracle in the reference's fp32 arithmetic and in float64 on the same inputs — the
    `ref32` / `truth64` pair of assert_parity for shapes that have no reference-executed golden vector."""
    P32, a32 = ogp.sample_prototypes(st, eps, visual)
    s64 = state_to(st, dtype=torch.float64)
    P64, a64 = ogp.sample_prototypes(s64, eps.double(), None if visual is None else visual.double())
    return a32["w"].detach(), P32.detach(), a64["w"].detach(), P64.detach()


def oracle_grad_pair(st, eps, dw, dkl):
    """(G32, G64): autograd through the oracle in fp32 (the reference's arithmetic) and float64."""
    _, _, G32, _ = oracle_grads(st, eps, dw, dkl, torch.float32)
    _, _, G64, _ = oracle_grads(st, eps, dw, dkl, torch.float64)
    return G32, G64


def fix_eval_noise(eng, S=None, seed=None, step=0):
    """Give a GPAdapterEngine explicit evaluation noise (the counter stream of oracle/philox.py) so that repeated eval calls of a
    test see the same draw; returns the CPU tensor [C, T, S] for the oracle side."""
    from oracle import philox
    S = int(S or eng.cfg.S_eval)
    eps = philox.eps_tensor(int(eng.cfg.seed if seed is None else seed), step, eng.C, eng.T, S)
    eng.eval_eps = eps.to(eng.dev)
    return eps
