"""``GaussianProcessTemplateWeighter`` — drop-in for trainers/gp_template_weigher.py:8-222, without
gpytorch / linear_operator / entmax: the per-class variational GP, the re-parameterised sampling,
sparsemax and the prototype contraction run in the clipgp sm_100a kernels (ops.py).

The module tree reproduces gpytorch's attribute and ``state_dict`` names so reference checkpoints
round-trip: ``variational_strategy.inducing_points``,
``variational_strategy._variational_distribution.{variational_mean,chol_variational_covar}``,
``covar_module.raw_outputscale`` / ``covar_module.base_kernel.raw_lengthscale`` (rbf),
``covar_module.raw_lengthscale`` (matern), ``covar_module.raw_variance`` (linear),
``mean_module.{f0,cls_bias,tmp_bias,neg_tail}``, ``A.weight``, ``likelihood.noise_covar.raw_noise``,
buffers ``_ind_mask, _templates, _templates_red, _cls_mean_init``.
"""
from __future__ import annotations

import math
from typing import Any, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops


def _inv_softplus(y: float) -> float:
    # gpytorch.utils.transforms.inv_softplus
    return y + math.log(-math.expm1(-y))


class _Positive(nn.Module):
    """gpytorch.constraints.Positive (softplus transform); carries the bound buffers gpytorch registers."""

    def __init__(self):
        super().__init__()
        self.register_buffer("lower_bound", torch.tensor(0.0))
        self.register_buffer("upper_bound", torch.tensor(float("inf")))

    def transform(self, raw):
        return F.softplus(raw)


class ResidualMeanWithBias(nn.Module):
    """trainers/gp_template_weigher.py:225-244 (same parameters / buffers; pure host-side torch: [C,T] floats)."""

    def __init__(self, f0_logits: torch.Tensor):
        super().__init__()
        K, M = f0_logits.shape
        self.register_buffer("f0", f0_logits.clone())
        self.cls_bias = nn.Parameter(torch.zeros(K, 1))
        self.tmp_bias = nn.Parameter(torch.zeros(1, M))
        self.register_buffer("neg_tail", torch.tensor(50.0))

    def forward(self, x):
        K, M = self.f0.shape
        N = x.size(-2)
        base = self.f0 + self.cls_bias + self.tmp_bias
        if N == M:
            return base
        tail = (self.cls_bias + self.tmp_bias.mean(dim=1, keepdim=True)).expand(K, N - M)
        return torch.cat([base, tail], dim=1)

    def test_mean(self, n_test: int) -> torch.Tensor:
        """mean_module(cat[Z; X])[:, n:] — every test row gets the per-class tail constant (:241-244)."""
        K, _ = self.f0.shape
        return (self.cls_bias + self.tmp_bias.mean(dim=1, keepdim=True)).expand(K, n_test)


class _RBFKernel(nn.Module):
    def __init__(self, C: int, d: int):
        super().__init__()
        self.raw_lengthscale = nn.Parameter(torch.zeros(C, 1, d))
        self.raw_lengthscale_constraint = _Positive()

    @property
    def lengthscale(self):
        return F.softplus(self.raw_lengthscale)

    def initialize(self, lengthscale: float):
        with torch.no_grad():
            self.raw_lengthscale.fill_(_inv_softplus(float(lengthscale)))
        return self


class _ScaleKernel(nn.Module):
    kind = "rbf"

    def __init__(self, base_kernel: _RBFKernel, C: int):
        super().__init__()
        self.base_kernel = base_kernel
        self.raw_outputscale = nn.Parameter(torch.zeros(C))
        self.raw_outputscale_constraint = _Positive()

    @property
    def outputscale(self):
        return F.softplus(self.raw_outputscale)


class _MaternKernel(nn.Module):
    kind = "matern"

    def __init__(self, C: int, d: int, nu: float = 0.5):
        super().__init__()
        if nu != 0.5:
            raise NotImplementedError("only nu=0.5 is used by the reference (gp_template_weigher.py:116)")
        self.nu = nu
        self.raw_lengthscale = nn.Parameter(torch.zeros(C, 1, d))
        self.raw_lengthscale_constraint = _Positive()

    @property
    def lengthscale(self):
        return F.softplus(self.raw_lengthscale)


class _LinearKernel(nn.Module):
    kind = "linear"

    def __init__(self, C: int):
        super().__init__()
        self.raw_variance = nn.Parameter(torch.zeros(C, 1, 1))
        self.raw_variance_constraint = _Positive()

    @property
    def variance(self):
        return F.softplus(self.raw_variance)


class _CholeskyVariationalDistribution(nn.Module):
    def __init__(self, n: int, C: int, mean_init_std: float = 1e-3):
        super().__init__()
        self.mean_init_std = mean_init_std
        self.variational_mean = nn.Parameter(torch.zeros(C, n))
        self.chol_variational_covar = nn.Parameter(torch.eye(n).repeat(C, 1, 1))

    @torch.no_grad()
    def initialize_variational_distribution(self):
        # gpytorch: mean <- prior mean (0) + 1e-3 * randn_like ; chol <- chol(I) = I.  Consumes torch RNG.
        self.variational_mean.zero_()
        self.variational_mean.add_(torch.randn_like(self.variational_mean), alpha=self.mean_init_std)
        n = self.chol_variational_covar.shape[-1]
        self.chol_variational_covar.copy_(torch.eye(n, device=self.chol_variational_covar.device).expand_as(self.chol_variational_covar))


class _VariationalStrategy(nn.Module):
    """Whitened gpytorch.variational.VariationalStrategy(learn_inducing_locations=True)."""

    def __init__(self, owner: "GaussianProcessTemplateWeighter", inducing_points: torch.Tensor, dist: _CholeskyVariationalDistribution):
        super().__init__()
        object.__setattr__(self, "_owner", owner)     # not a sub-module (avoids a cycle in state_dict)
        self.inducing_points = nn.Parameter(inducing_points.clone())
        self._variational_distribution = dist
        self.register_buffer("variational_params_initialized", torch.tensor(0))
        self.register_buffer("updated_strategy", torch.tensor(True))

    def _maybe_init(self):
        if getattr(self, "_inited", False):
            return
        if not bool(self.variational_params_initialized.item()):
            self._variational_distribution.initialize_variational_distribution()
            self.variational_params_initialized.fill_(1)
        self._inited = True

    def kl_divergence(self) -> torch.Tensor:
        """KL(q(u) || N(0, I)) per class, shape [C] (call sites: adapter.py:463, taskres.py:271,
        clip_adapter.py:274, tip_adapter.py:141).  Re-uses the value the last ``sample_prototypes`` kernel
        produced when q(u) has not changed since (same autograd node, so one backward kernel serves both)."""
        return self._owner._kl()


class _NoiseCovar(nn.Module):
    def __init__(self, C: int):
        super().__init__()
        self.raw_noise = nn.Parameter(torch.zeros(C, 1))
        self.raw_noise_constraint = _Positive()


class _GaussianLikelihood(nn.Module):
    """Unused by the reference's hot path (gp_template_weigher.py:126); kept for state_dict parity."""

    def __init__(self, C: int):
        super().__init__()
        self.noise_covar = _NoiseCovar(C)


class _PriorMVN:
    """What ``GaussianProcessTemplateWeighter.forward`` returns in place of gpytorch.distributions.MultivariateNormal:
    the prior's mean [K,N] and dense covariance [K,N,N]."""

    def __init__(self, mean: torch.Tensor, covariance_matrix: torch.Tensor):
        self.mean = self.loc = mean
        self.covariance_matrix = self.lazy_covariance_matrix = covariance_matrix

    @property
    def variance(self) -> torch.Tensor:
        return self.covariance_matrix.diagonal(dim1=-2, dim2=-1)


class GaussianProcessTemplateWeighter(nn.Module):
    """Per-class variational GP over templates -> sparsemax template weights -> prototypes.

    Same constructor and public surface as the reference class: ``sample_prototypes(num_samples,
    visual_embeddings=None) -> [S,K,D]`` (differentiable), ``.scores``, ``.variational_strategy
    .kl_divergence()``, ``.initialize_from_weights``, ``.covar_module``, ``.mean_module``.

    ``rng`` selects the base-noise source: ``"torch"`` draws ``torch.randn(C, Nx, S)`` exactly where the
    reference does (MultivariateNormal.rsample), ``"philox"`` uses the on-device counter RNG (no eps tensor).
    """

    def __init__(self, text_embeddings: torch.Tensor, cfg: Any, rng: str = "torch", seed: int = 0,
                 lengthscale: Optional[float] = None, **kwargs) -> None:
        super().__init__()
        self.orig_device = text_embeddings.device
        self.num_classes, self.num_templates, self.dim = text_embeddings.shape
        K, M, D = self.num_classes, self.num_templates, self.dim
        adapter_cfg = getattr(cfg, "adapter", cfg)
        self.input_dim = D
        self.red_dim = int(getattr(adapter_cfg, "gp_pca_dim", 128))
        text_embeddings = text_embeddings.detach().float()

        # ---- PCA (gp_template_weigher.py:26-37)
        with torch.no_grad():
            X = text_embeddings.reshape(-1, D)
            mu = X.mean(dim=0, keepdim=True)
            max_rank = min(X.shape[0], D)                                  # Vt.shape[0] of the reference's thin SVD
            self.red_dim = min(self.red_dim, max_rank)
            if X.is_cuda:
                W = self._pca_axes_device(X - mu, self.red_dim)
            else:
                _, _, Vt = torch.linalg.svd(X - mu, full_matrices=False)
                W = Vt[: self.red_dim].T.contiguous()
        self._pca_mean = mu.squeeze(0)
        self._pca_W = W
        with torch.no_grad():
            cls_mean = text_embeddings.mean(dim=1, keepdim=True)
            templates_red = self._project(text_embeddings.view(-1, D)).view(K, M, self.red_dim).contiguous()
            cls_mean_red = self._project(cls_mean.view(-1, D)).view(K, 1, self.red_dim)
        self.scores = None

        # ---- inducing points = [templates_red ; class-mean token] (:60-63)
        n_ind = M + 1
        inducing = torch.cat([templates_red, cls_mean_red], dim=1)
        dist = _CholeskyVariationalDistribution(n_ind, K)
        self.variational_strategy = _VariationalStrategy(self, inducing, dist)

        # unused learnable map (:68-70; _to_kernel_space is commented out at :167)
        self.A = nn.Linear(self.red_dim, self.red_dim, bias=False)
        with torch.no_grad():
            self.A.weight.copy_(torch.eye(self.red_dim))

        mask = torch.zeros(K, n_ind, self.red_dim, device=inducing.device)
        mask[:, M:, :] = 1.0
        self.register_buffer("_ind_mask", mask)
        self.variational_strategy.inducing_points.register_hook(lambda g: g * self._ind_mask)   # :76-79

        # ---- prior mean f0 = log softmax(cos(template, class mean) / tau) (:83-98)
        with torch.no_grad():
            mean_init = (F.normalize(text_embeddings, p=2, dim=-1) * F.normalize(cls_mean, p=2, dim=-1)).sum(-1)
            tau = float(getattr(adapter_cfg, "gp_prior_temp", 1.0) or 1.0)
            f0 = torch.log(torch.softmax(mean_init / max(tau, 1e-6), dim=-1).clamp_min(1e-12))
        self.mean_module = ResidualMeanWithBias(f0_logits=f0.to(torch.float32))

        # ---- kernel (:101-122)
        kernel_type = getattr(adapter_cfg, "gp_kernel_type", "rbf")
        self.kernel_type = kernel_type
        if kernel_type == "rbf":
            # `lengthscale` (extension) skips the O((C*T)^2) median heuristic when the caller already knows the value
            ls_cfg = float(lengthscale) if lengthscale is not None else self._median_lengthscale(templates_red)
            print(f"[GP] Auto length-scale (normalised median): {ls_cfg:.4f}")
            base = _RBFKernel(K, self.red_dim).initialize(lengthscale=ls_cfg)
            self.covar_module = _ScaleKernel(base, K)
        elif kernel_type == "matern":
            self.covar_module = _MaternKernel(K, self.red_dim, nu=0.5)
        elif kernel_type == "linear":
            self.covar_module = _LinearKernel(K)
        else:
            raise ValueError(f"Unsupported kernel: {kernel_type}")
        self.likelihood = _GaussianLikelihood(K)

        self.register_buffer("_templates", text_embeddings.contiguous())
        self.register_buffer("_templates_red", templates_red.detach())
        self.register_buffer("_cls_mean_init", cls_mean)
        self.register_buffer("_pca_mean_buf", self._pca_mean, persistent=False)
        self.register_buffer("_pca_W_buf", self._pca_W, persistent=False)

        # ---- clipgp runtime state
        self.rng = rng
        self.register_buffer("_rng_state", torch.tensor([int(seed), 0], dtype=torch.int64), persistent=False)
        self.eval_eps = None        # optional explicit base noise [K, >=M, S] used by eval-mode calls whose S matches (parity tests)
        self._last = None           # (kl tensor, versions) of the most recent kernel launch
        self.last_status = None     # int32 [C]: 0 ok, k>0 jitter retries, <0 not positive definite

    # ------------------------------------------------------------------ helpers
    def _project(self, x):
        return (x - self._pca_mean.to(x.device)) @ self._pca_W.to(x.device)

    def _lift(self, z):
        return z @ self._pca_W.to(z.device).T + self._pca_mean.to(z.device)

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self._pca_mean = self._pca_mean_buf
        self._pca_W = self._pca_W_buf
        return out

    @staticmethod
    @torch.no_grad()
    def _pca_axes_device(Xc: torch.Tensor, d: int) -> torch.Tensor:
        """Leading d right-singular vectors of the centred [C*T, D] text bank WITHOUT the SVD of the tall matrix
        (gp_template_weigher.py:26-37 runs torch.linalg.svd on [32 000, 512] at the ImageNet shape): the D x D
        Gram matrix Xc^T Xc comes from the clipgp fp32 GEMM (one pass over the bank, contraction over the rows read in place),
        its symmetric eigen-decomposition (512 x 512, a library call in one-time setup) gives the axes in descending order.
        Axes are defined up to sign (and rotation inside degenerate eigenspaces), exactly like the SVD's; everything downstream
        consumes inner products / distances of the projected points, which depend on the subspace only."""
        N, D = Xc.shape
        Xc = Xc.float().contiguous()
        G = torch.empty(D, D, dtype=torch.float32, device=Xc.device)
        ops._gemm(Xc, 1, D, Xc, D, 1, G, D, D, N, 1.0, accumulate=2)       # G[i,j] = sum_r Xc[r,i] Xc[r,j]; flag 2: no split-K, so
        #                                                                    two constructions give bit-identical axes (set-up: 16 CTAs suffice)
        G = 0.5 * (G + G.t())
        evals, evecs = torch.linalg.eigh(G.double())                       # ascending
        return evecs[:, -d:].flip(-1).float().contiguous()                 # [D, d], largest variance first

    @staticmethod
    @torch.no_grad()
    def _median_lengthscale(templates_red: torch.Tensor, chunk: int = 2048) -> float:
        """:103-107 median of the non-zero pairwise distances of the unit-normalised reduced templates.
        Chunked so C*T = 32 000 points never materialise the 4 GB distance matrix.
        torch.median over an even count returns the LOWER middle element -> k = (m-1)//2 smallest."""
        flat = F.normalize(templates_red.reshape(-1, templates_red.shape[-1]), p=2, dim=-1)
        if flat.is_cuda:
            from . import ops
            return ops.median_pairwise_distance(flat)      # exact radix select, no N x N matrix (csrc/setup.cu)
        parts = []
        for i in range(0, flat.shape[0], chunk):
            pd = torch.cdist(flat[i:i + chunk], flat)
            parts.append(pd[pd > 0])
        vals = torch.cat(parts)
        return vals.median().item()

    def _kernel_raw(self):
        cm = self.covar_module
        if self.kernel_type == "rbf":
            return cm.base_kernel.raw_lengthscale, cm.raw_outputscale, None
        if self.kernel_type == "matern":
            return cm.raw_lengthscale, None, None
        return None, None, cm.raw_variance

    def _versions(self):
        q = self.variational_strategy._variational_distribution
        return (q.variational_mean._version, q.chol_variational_covar._version)

    def _kl(self) -> torch.Tensor:
        # re-use the KL by-product of the last launch only while q(u) is unchanged AND (under autograd) that launch's graph is
        # still alive: a backward through w or kl frees it (the hooks below flag that), and then a fresh launch is needed
        if self._last is not None and self._last[1] == self._versions() and (
                not torch.is_grad_enabled() or (self._last[0].requires_grad and not self._last[2]["consumed"])):
            return self._last[0]
        # q(u) changed since the last launch: evaluate the kernel once for its KL by-product
        self._launch(num_samples=1, eps=torch.zeros(self.num_classes, self.num_templates, 1, device=self._templates.device))
        return self._last[0]

    def _launch(self, num_samples: int, eps: Optional[torch.Tensor]):
        vs = self.variational_strategy
        q = vs._variational_distribution
        raw_ls, raw_os, raw_var = self._kernel_raw()
        mean_x = self.mean_module.test_mean(self.num_templates)
        w, kl, status = ops.gp_weights(vs.inducing_points, self._templates_red, raw_ls, raw_os, raw_var,
                                       q.variational_mean, q.chol_variational_covar, mean_x, eps, self.kernel_type,
                                       num_samples, rng_state=self._rng_state if eps is None else None)
        flag = {"consumed": False}
        if kl.requires_grad:
            def _mark(_g, flag=flag):
                flag["consumed"] = True
            w.register_hook(_mark); kl.register_hook(_mark)
        self._last = (kl, self._versions(), flag)
        self.last_status = status
        return w

    # ------------------------------------------------------------------ reference API
    @torch.no_grad()
    def initialize_from_weights(self, weights_km: torch.Tensor, temperature: float = 1.0) -> None:
        """:139-164.  In the reference this is effectively a no-op: the mean module has no ``mean_param`` and
        ``variational_mean`` is [K, M+1] so ``copy_([K, M])`` raises and is swallowed (it broadcasts only for M == 1).
        Reproduced as such."""
        w = torch.clamp(weights_km.to(device=self._templates.device), min=1e-12)
        f_init = torch.log(w) / max(float(temperature), 1e-6)
        try:
            self.variational_strategy._variational_distribution.variational_mean.data.copy_(f_init)
        except Exception:
            pass

    def forward(self, x):
        """:167-175 -> the PRIOR over f at ``x`` [K,N,d]: mean_module(x), covar_module(x).  In the reference this is what the
        variational strategy calls on cat[Z; X]; here that product is fused into the clipgp GP kernel (``sample_prototypes``),
        so this method is host-side torch for callers that inspect the prior (not a hot path, differentiable by autograd).
        Returns an object with ``mean`` [K,N], ``covariance_matrix`` [K,N,N] (also ``loc`` / ``lazy_covariance_matrix``)."""
        mean_x = self.mean_module(x)
        if not isinstance(mean_x, torch.Tensor):
            raise TypeError("Mean module must return a tensor")
        return _PriorMVN(mean_x, self._prior_covariance(x))

    def _prior_covariance(self, x: torch.Tensor) -> torch.Tensor:
        """covar_module(x) of :101-122 with gpytorch's arithmetic: squared distances as ONE product of the augmented rows
        [-2a, |a|^2, 1] . [b, 1, |b|^2] after shifting by the row mean (with autograd off, equal inputs get an exactly-zero
        diagonal), RBF = outputscale * exp(-r^2 / 2) on x / lengthscale, Matern-1/2 = exp(-r) on (x - mean of all rows) /
        lengthscale with r clamped at 1e-15, linear = variance * x x^T."""
        cm = self.covar_module

        def sqd(a):
            exact_diag = not torch.is_grad_enabled()
            a = a - a.mean(-2, keepdim=True)
            n2 = a.pow(2).sum(-1, keepdim=True)
            one = torch.ones_like(n2)
            r = torch.cat([-2.0 * a, n2, one], -1) @ torch.cat([a, one, n2], -1).transpose(-2, -1)
            if exact_diag:
                r.diagonal(dim1=-2, dim2=-1).fill_(0)
            return r.clamp_min(0)

        if self.kernel_type == "rbf":
            return sqd(x / cm.base_kernel.lengthscale).div(-2).exp() * cm.outputscale.view(-1, 1, 1)
        if self.kernel_type == "matern":
            centre = x.reshape(-1, x.size(-1)).mean(0)
            return torch.exp(-sqd((x - centre) / cm.lengthscale).clamp_min(1e-30).sqrt())
        sx = x * cm.variance.sqrt()
        return sx @ sx.transpose(-2, -1)

    def sample_weights(self, num_samples: int, visual_embeddings: Optional[torch.Tensor] = None,
                       eps: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Template weights w [S,K,M] = sparsemax(f_s) (:213-217).  ``eps`` overrides the base noise (tests)."""
        self.variational_strategy._maybe_init()
        S = max(1, int(num_samples))
        K, M = self.num_classes, self.num_templates
        if eps is None and not self.training and self.eval_eps is not None and self.eval_eps.shape[-1] == S:
            eps = self.eval_eps
        if eps is None and self.rng == "torch":
            # :198-203 — a visual batch whose size happens to equal K adds one test row; it never reaches the
            # first M outputs (lower-triangular chol), it only changes how much RNG is consumed.
            nx = M + 1 if (visual_embeddings is not None and visual_embeddings.shape[0] == K) else M
            eps = torch.randn(K, nx, S, dtype=torch.float32, device=self._templates.device)
        w = self._launch(S, eps)
        if eps is None:
            self._rng_state[1] += 1
        return w

    def sample_prototypes(self, num_samples: int, visual_embeddings: Optional[torch.Tensor] = None,
                          eps: Optional[torch.Tensor] = None) -> torch.Tensor:
        """:183-222 -> un-normalised prototypes [S,K,D]."""
        w = self.sample_weights(num_samples, visual_embeddings, eps)
        self.scores = w
        return ops.prototypes(w, self._templates)

    @torch.no_grad()
    def mean_prototypes(self, num_samples: int, eps: Optional[torch.Tensor] = None) -> torch.Tensor:
        """normalize(mean_s prototypes_s) [K,D] — the prototype init of taskres.py:281-285,
        clip_adapter.py:284-288, tip_adapter.py:152-156, in one fused pass."""
        w = self.sample_weights(num_samples, None, eps)
        self.scores = w
        return ops.prototypes_reduced(w, self._templates, want_mean_raw=True)[2]

    @torch.no_grad()
    def collapsed_prototypes(self, num_samples: int, eps: Optional[torch.Tensor] = None, residual=None, alpha: float = 0.0):
        """(1/S) sum_s p_hat_s [K,D]: mean_s scale*f.p_hat_s == scale*f.(mean_s p_hat_s), exact for the
        reference's logit-mean eval (adapter.py:247-249 etc.)."""
        w = self.sample_weights(num_samples, None, eps)
        self.scores = w
        return ops.prototypes_reduced(w, self._templates, residual=residual, alpha=alpha, want_mean_hat=True)[1]
