"""CPU oracle: the GP template weighter (TEST INFRASTRUCTURE, see oracle/__init__.py).

PINNED (tests/test_ref_golden.py) on tests/golden/ref_gp.npz: vectors produced by running the reference's own
trainers/gp_template_weigher.py, unmodified, on the gpytorch / linear_operator / entmax stand-ins of oracle/_shim
(tests/golden/make_ref_golden.py).  Every function cites the reference call site it restates (paths relative to
/root/reference) and, where the arithmetic lives in gpytorch / linear_operator / entmax, the library routine whose
published algorithm is restated.

Everything is plain differentiable torch, so ``torch.autograd`` through these functions
is the gradient oracle for the hand-written CUDA backward.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional, Tuple

import torch
import torch.nn.functional as F

VARIATIONAL_JITTER = 1e-4   # gpytorch settings.variational_cholesky_jitter (float32 default)
CHOL_JITTER = 1e-6          # gpytorch settings.cholesky_jitter (float32 default)
CHOL_MAX_TRIES = 3          # gpytorch settings.cholesky_max_tries


# ----------------------------------------------------------------------------------------
# constraints (gpytorch.constraints.Positive == softplus transform)
# ----------------------------------------------------------------------------------------
def softplus(x: torch.Tensor) -> torch.Tensor:
    return F.softplus(x)


def inv_softplus(y: float | torch.Tensor) -> torch.Tensor:
    """gpytorch.utils.transforms.inv_softplus: raw = y + log(-expm1(-y))."""
    y = torch.as_tensor(y, dtype=torch.float32)
    return y + torch.log(-torch.expm1(-y))


# ----------------------------------------------------------------------------------------
# distances (gpytorch.kernels.kernel.sq_dist / dist)
# ----------------------------------------------------------------------------------------
def sq_dist(x1: torch.Tensor, x2: torch.Tensor) -> torch.Tensor:
    """gpytorch ``sq_dist``: subtract x1's row mean from both inputs, expand |a|^2 - 2ab + |b|^2 as one matmul, clamp at 0.

    gpytorch has a second branch, ``x1_eq_x2 and not x1.requires_grad and not x2.requires_grad``, that re-uses x1's norms and
    forces the diagonal to exactly 0.  The reference's parameters always require grad (inducing points, length-scales), so the
    branch is taken exactly when autograd is disabled (``torch.no_grad()`` evaluation, e.g. adapter.py:362, taskres.py:280):
    here it is keyed on ``torch.is_grad_enabled()``.  It matters for Matern-1/2 only: the expansion leaves ~1e-6 noise on the
    diagonal, and sqrt turns that into a 1e-3 deficit of K_ii in grad mode (pinned by tests/golden/ref_gp.npz)."""
    x1_eq_x2 = (x1.shape == x2.shape) and torch.equal(x1, x2) and not torch.is_grad_enabled()
    adjustment = x1.mean(-2, keepdim=True)
    x1 = x1 - adjustment
    x1_norm = x1.pow(2).sum(dim=-1, keepdim=True)
    x1_pad = torch.ones_like(x1_norm)
    if x1_eq_x2:
        x2, x2_norm, x2_pad = x1, x1_norm, x1_pad
    else:
        x2 = x2 - adjustment
        x2_norm = x2.pow(2).sum(dim=-1, keepdim=True)
        x2_pad = torch.ones_like(x2_norm)
    x1_ = torch.cat([-2.0 * x1, x1_norm, x1_pad], dim=-1)
    x2_ = torch.cat([x2, x2_pad, x2_norm], dim=-1)
    res = x1_.matmul(x2_.transpose(-2, -1))
    if x1_eq_x2:
        res.diagonal(dim1=-2, dim2=-1).fill_(0)
    return res.clamp_min(0)


def dist(x1: torch.Tensor, x2: torch.Tensor) -> torch.Tensor:
    return sq_dist(x1, x2).clamp_min(1e-30).sqrt()


# ----------------------------------------------------------------------------------------
# kernels (trainers/gp_template_weigher.py:101-122)
# ----------------------------------------------------------------------------------------
@dataclass
class KernelParams:
    """Raw (unconstrained) kernel parameters with gpytorch's names/shapes.

    rbf    : raw_lengthscale [C,1,d], raw_outputscale [C]   (ScaleKernel(RBFKernel ARD))
    matern : raw_lengthscale [C,1,d]                         (MaternKernel nu=0.5 ARD, no scale)
    linear : raw_variance    [C,1,1]                         (LinearKernel, no scale)
    """
    kind: str
    raw_lengthscale: Optional[torch.Tensor] = None
    raw_outputscale: Optional[torch.Tensor] = None
    raw_variance: Optional[torch.Tensor] = None


def kernel_matrix(kp: KernelParams, x1: torch.Tensor, x2: torch.Tensor) -> torch.Tensor:
    """K(x1, x2) for batched inputs [C,n1,d], [C,n2,d] -> [C,n1,n2] (fp32)."""
    if kp.kind == "rbf":
        ls = softplus(kp.raw_lengthscale)                      # [C,1,d]
        os_ = softplus(kp.raw_outputscale)                     # [C]
        k = sq_dist(x1 / ls, x2 / ls).div(-2).exp()            # RBFKernel.forward + postprocess_rbf
        return k * os_.view(-1, 1, 1)                          # ScaleKernel.forward
    if kp.kind == "matern":
        ls = softplus(kp.raw_lengthscale)
        # MaternKernel.forward: centre by the mean over ALL rows of x1 (all batches), nu=0.5
        mean = x1.reshape(-1, x1.size(-1)).mean(0)[(None,) * (x1.dim() - 1)]
        x1_ = (x1 - mean) / ls
        x2_ = (x2 - mean) / ls
        return torch.exp(-math.sqrt(0.5 * 2) * dist(x1_, x2_))
    if kp.kind == "linear":
        v = softplus(kp.raw_variance)                          # [C,1,1]
        return (x1 * v.sqrt()).matmul((x2 * v.sqrt()).transpose(-2, -1))
    raise ValueError(f"Unsupported kernel: {kp.kind}")         # gp_template_weigher.py:122


# ----------------------------------------------------------------------------------------
# psd_safe_cholesky (linear_operator.utils.cholesky)
# ----------------------------------------------------------------------------------------
def psd_safe_cholesky(A: torch.Tensor, jitter: Optional[float] = None,
                      max_tries: int = CHOL_MAX_TRIES) -> torch.Tensor:
    L, info = torch.linalg.cholesky_ex(A)
    if not torch.any(info):
        return L
    if torch.isnan(A).any():
        raise RuntimeError("cholesky: input has NaNs")
    if jitter is None:
        jitter = CHOL_JITTER if A.dtype == torch.float32 else 1e-8
    Aprime = A.clone()
    jitter_prev = 0.0
    for i in range(max_tries):
        jitter_new = jitter * (10 ** i)
        diag_add = ((info > 0) * (jitter_new - jitter_prev)).unsqueeze(-1).expand(*Aprime.shape[:-1])
        Aprime = Aprime + torch.diag_embed(diag_add.to(Aprime.dtype))
        jitter_prev = jitter_new
        L, info = torch.linalg.cholesky_ex(Aprime)
        if not torch.any(info):
            return L
    raise RuntimeError(f"Matrix not positive definite after repeatedly adding jitter up to {jitter_new:.1e}.")


# ----------------------------------------------------------------------------------------
# mean module (trainers/gp_template_weigher.py:225-244)
# ----------------------------------------------------------------------------------------
def residual_mean(f0: torch.Tensor, cls_bias: torch.Tensor, tmp_bias: torch.Tensor, N: int) -> torch.Tensor:
    K, M = f0.shape
    base = f0 + cls_bias + tmp_bias
    if N == M:
        return base
    extra = N - M
    tail = (cls_bias + tmp_bias.mean(dim=1, keepdim=True)).expand(K, extra)
    return torch.cat([base, tail], dim=1)


# ----------------------------------------------------------------------------------------
# whitened VariationalStrategy.forward (gpytorch.variational.VariationalStrategy)
# called from gp_template_weigher.py:213 via ApproximateGP.__call__
# ----------------------------------------------------------------------------------------
def variational_predictive(kp: KernelParams, Z: torch.Tensor, X: torch.Tensor,
                           var_mean: torch.Tensor, chol_var: torch.Tensor,
                           test_mean: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, dict]:
    """q(f(X)) = N(mu, Sigma) for the whitened variational strategy.

    Z [C,n,d] inducing points, X [C,Nx,d] test inputs, var_mean [C,n], chol_var [C,n,n]
    (raw; masked to lower-triangular here as CholeskyVariationalDistribution.forward does),
    test_mean [C,Nx] = mean_module(cat[Z;X])[:, n:].
    """
    C, n, _ = Z.shape
    eye_n = torch.eye(n, dtype=Z.dtype, device=Z.device)
    K_ZZ = kernel_matrix(kp, Z, Z) + VARIATIONAL_JITTER * eye_n          # add_jitter, fp32
    K_ZX = kernel_matrix(kp, Z, X)
    K_XX = kernel_matrix(kp, X, X)
    L = psd_safe_cholesky(K_ZZ.double())                                   # _cholesky_factor, float64
    A = torch.linalg.solve_triangular(L, K_ZX.double(), upper=False).to(Z.dtype)   # interp_term
    Lq = chol_var * torch.ones(n, n, dtype=Z.dtype, device=Z.device).tril(0)
    mu = (A.transpose(-1, -2) @ var_mean.unsqueeze(-1)).squeeze(-1) + test_mean
    middle_A = Lq @ (Lq.transpose(-1, -2) @ A) - A                         # (S_u - I) @ interp
    eye_x = torch.eye(X.shape[-2], dtype=Z.dtype, device=Z.device)
    Sigma = K_XX + VARIATIONAL_JITTER * eye_x + A.transpose(-1, -2) @ middle_A
    return mu, Sigma, {"K_ZZ": K_ZZ, "K_ZX": K_ZX, "K_XX": K_XX, "L": L, "A": A, "Lq": Lq}


def rsample(mu: torch.Tensor, Sigma: torch.Tensor, eps: torch.Tensor) -> torch.Tensor:
    """MultivariateNormal.rsample([S]) with explicit base noise.

    linear_operator ``zero_mean_mvn_samples``: R = chol_fp32(Sigma) (psd_safe), eps ~ randn(C,Nx,S),
    samples = (R @ eps).permute(-1, 0, 1) + mu  ->  [S,C,Nx].
    """
    R = psd_safe_cholesky(Sigma)
    return (R @ eps).permute(2, 0, 1) + mu.unsqueeze(0)


# ----------------------------------------------------------------------------------------
# entmax.sparsemax (entmax.activations.SparsemaxFunction), gp_template_weigher.py:217
# ----------------------------------------------------------------------------------------
class _Sparsemax(torch.autograd.Function):
    @staticmethod
    def forward(ctx, X):
        max_val, _ = X.max(dim=-1, keepdim=True)
        X = X - max_val
        topk, _ = torch.sort(X, dim=-1, descending=True)
        topk_cumsum = topk.cumsum(-1) - 1
        d = X.shape[-1]
        rhos = torch.arange(1, d + 1, device=X.device, dtype=X.dtype).view(*([1] * (X.dim() - 1)), d)
        support = rhos * topk > topk_cumsum
        support_size = support.sum(dim=-1).unsqueeze(-1)
        tau = topk_cumsum.gather(-1, support_size - 1)
        tau = tau / support_size.to(X.dtype)
        output = torch.clamp(X - tau, min=0)
        ctx.save_for_backward(support_size, output)
        return output

    @staticmethod
    def backward(ctx, grad_output):
        supp_size, output = ctx.saved_tensors
        grad_input = grad_output.clone()
        grad_input[output == 0] = 0
        v_hat = grad_input.sum(dim=-1) / supp_size.to(output.dtype).squeeze(-1)
        v_hat = v_hat.unsqueeze(-1)
        grad_input = torch.where(output != 0, grad_input - v_hat, grad_input)
        return grad_input


def sparsemax(X: torch.Tensor) -> torch.Tensor:
    return _Sparsemax.apply(X)


# ----------------------------------------------------------------------------------------
# KL(q(u) || N(0,I)) per class (gpytorch kl_mvn_mvn; call sites adapter.py:463 etc.)
# ----------------------------------------------------------------------------------------
def kl_divergence(var_mean: torch.Tensor, chol_var: torch.Tensor) -> torch.Tensor:
    n = var_mean.shape[-1]
    Lq = chol_var * torch.ones(n, n, dtype=chol_var.dtype, device=chol_var.device).tril(0)
    logdet_p = Lq.diagonal(dim1=-2, dim2=-1).pow(2).log().sum(-1)
    trace_plus_inv_quad = var_mean.pow(2).sum(-1) + Lq.pow(2).sum((-2, -1))
    return 0.5 * (-logdet_p + trace_plus_inv_quad - float(n))


# ----------------------------------------------------------------------------------------
# The weighter state + sample_prototypes (gp_template_weigher.py:13-132, 183-222)
# ----------------------------------------------------------------------------------------
@dataclass
class GPState:
    templates: torch.Tensor        # [C,T,D]  (_templates)
    templates_red: torch.Tensor    # [C,T,d]  (_templates_red)
    inducing_points: torch.Tensor  # [C,T+1,d]
    var_mean: torch.Tensor         # [C,T+1]
    chol_var: torch.Tensor         # [C,T+1,T+1]
    kernel: KernelParams
    f0: torch.Tensor               # [C,T]
    cls_bias: torch.Tensor         # [C,1]
    tmp_bias: torch.Tensor         # [1,T]
    pca_mean: torch.Tensor         # [D]
    pca_W: torch.Tensor            # [D,d]

    def trainable(self):
        ps = [self.inducing_points, self.var_mean, self.chol_var, self.cls_bias, self.tmp_bias]
        for p in (self.kernel.raw_lengthscale, self.kernel.raw_outputscale, self.kernel.raw_variance):
            if p is not None:
                ps.append(p)
        return ps


def pca_setup(text_embeddings: torch.Tensor, pca_dim: int):
    """gp_template_weigher.py:22-52."""
    K, M, D = text_embeddings.shape
    X = text_embeddings.reshape(-1, D)
    mu = X.mean(dim=0, keepdim=True)
    Xc = X - mu
    _, _, Vt = torch.linalg.svd(Xc, full_matrices=False)
    red = min(int(pca_dim), Vt.shape[0])
    W = Vt[:red].T.contiguous()
    pca_mean = mu.squeeze(0)
    templates_red = ((X - pca_mean) @ W).view(K, M, red)
    cls_mean = text_embeddings.mean(dim=1, keepdim=True)
    cls_mean_red = ((cls_mean.view(-1, D) - pca_mean) @ W).view(K, 1, red)
    return pca_mean, W, templates_red, cls_mean_red, cls_mean


def median_lengthscale(templates_red: torch.Tensor, chunk: int = 4096) -> float:
    """gp_template_weigher.py:103-107: median of the non-zero pairwise distances of the
    unit-normalised reduced templates over all C*T points (torch.median = lower median)."""
    flat = F.normalize(templates_red.reshape(-1, templates_red.shape[-1]), p=2, dim=-1)
    vals = []
    for i in range(0, flat.shape[0], chunk):
        pd = torch.cdist(flat[i:i + chunk], flat)
        vals.append(pd[pd > 0])
    return torch.cat(vals).median().item()


def build_state(text_embeddings: torch.Tensor, kernel_type: str = "rbf", pca_dim: int = 256,
                prior_temp: float = 1.0, lengthscale: Optional[float] = None) -> GPState:
    """GaussianProcessTemplateWeighter.__init__ (gp_template_weigher.py:13-132) followed by
    gpytorch's first-call initialisation of q(u) WITHOUT the 1e-3 noise (var_mean = 0,
    chol = I); tests overwrite var_mean / chol_var explicitly."""
    C, T, D = text_embeddings.shape
    pca_mean, W, templates_red, cls_mean_red, _ = pca_setup(text_embeddings, pca_dim)
    d = templates_red.shape[-1]
    Z = torch.cat([templates_red, cls_mean_red], dim=1).clone()
    class_mean = text_embeddings.mean(dim=1, keepdim=True)
    mean_init = (F.normalize(text_embeddings, dim=-1) * F.normalize(class_mean, dim=-1)).sum(-1)
    tau = float(prior_temp or 1.0)
    f0 = torch.log(torch.softmax(mean_init / max(tau, 1e-6), dim=-1).clamp_min(1e-12)).float()
    if kernel_type == "rbf":
        ls = float(lengthscale) if lengthscale is not None else median_lengthscale(templates_red)
        kp = KernelParams("rbf",
                          raw_lengthscale=inv_softplus(ls).expand(C, 1, d).clone().contiguous(),
                          raw_outputscale=torch.zeros(C))
    elif kernel_type == "matern":
        kp = KernelParams("matern", raw_lengthscale=torch.zeros(C, 1, d))
    elif kernel_type == "linear":
        kp = KernelParams("linear", raw_variance=torch.zeros(C, 1, 1))
    else:
        raise ValueError(f"Unsupported kernel: {kernel_type}")
    n = T + 1
    return GPState(templates=text_embeddings, templates_red=templates_red, inducing_points=Z,
                   var_mean=torch.zeros(C, n), chol_var=torch.eye(n).repeat(C, 1, 1), kernel=kp,
                   f0=f0, cls_bias=torch.zeros(C, 1), tmp_bias=torch.zeros(1, T),
                   pca_mean=pca_mean, pca_W=W)


def gp_weights(st: GPState, eps: torch.Tensor, visual_embeddings: Optional[torch.Tensor] = None):
    """sample_prototypes up to the template weights (gp_template_weigher.py:194-217).

    eps: [C, Nx, S] base noise with Nx = T (or T+1 in the ``batch == C`` branch, :198-203).
    Returns (w [S,C,T], aux dict).
    """
    C, T, _ = st.templates_red.shape
    if (visual_embeddings is not None) and (visual_embeddings.shape[0] == C):
        ve = ((visual_embeddings - st.pca_mean) @ st.pca_W).unsqueeze(1)
        gp_input = torch.cat([st.templates_red, ve], dim=1)
    else:
        gp_input = st.templates_red
    n = st.inducing_points.shape[1]
    Nx = gp_input.shape[1]
    full_mean = residual_mean(st.f0, st.cls_bias, st.tmp_bias, n + Nx)
    test_mean = full_mean[:, n:]
    mu, Sigma, aux = variational_predictive(st.kernel, st.inducing_points, gp_input,
                                            st.var_mean, st.chol_var, test_mean)
    f = rsample(mu, Sigma, eps)[:, :, :T]
    w = sparsemax(f)
    aux.update(mu=mu, Sigma=Sigma, f=f)
    return w, aux


def sample_prototypes(st: GPState, eps: torch.Tensor, visual_embeddings: Optional[torch.Tensor] = None):
    """gp_template_weigher.py:183-222 -> un-normalised prototypes [S,C,D]."""
    w, aux = gp_weights(st, eps, visual_embeddings)
    protos = torch.einsum("skm,kmd->skd", w, st.templates)
    aux["w"] = w
    return protos, aux
