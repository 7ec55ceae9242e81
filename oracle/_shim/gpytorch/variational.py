"""gpytorch.variational: CholeskyVariationalDistribution + whitened VariationalStrategy (1.11-1.14)."""
import torch

from linear_operator.utils.cholesky import psd_safe_cholesky

from . import settings
from .distributions import MultivariateNormal, _Chol, _Diag, kl_mvn_mvn
from .lazy import LazyEvaluatedKernelTensor
from .module import Module


class _VariationalDistribution(Module):
    def __init__(self, num_inducing_points, batch_shape=torch.Size([]), mean_init_std=1e-3):
        super().__init__()
        self.num_inducing_points = num_inducing_points
        self.batch_shape = batch_shape
        self.mean_init_std = mean_init_std

    def shape(self):
        return torch.Size([*self.batch_shape, self.num_inducing_points])

    def __call__(self):
        return self.forward()


class CholeskyVariationalDistribution(_VariationalDistribution):
    def __init__(self, num_inducing_points, batch_shape=torch.Size([]), mean_init_std=1e-3, **kwargs):
        super().__init__(num_inducing_points=num_inducing_points, batch_shape=batch_shape, mean_init_std=mean_init_std)
        mean_init = torch.zeros(num_inducing_points)
        covar_init = torch.eye(num_inducing_points, num_inducing_points)
        mean_init = mean_init.repeat(*batch_shape, 1)
        covar_init = covar_init.repeat(*batch_shape, 1, 1)
        self.register_parameter(name="variational_mean", parameter=torch.nn.Parameter(mean_init))
        self.register_parameter(name="chol_variational_covar", parameter=torch.nn.Parameter(covar_init))

    def forward(self):
        chol_variational_covar = self.chol_variational_covar
        dtype, device = chol_variational_covar.dtype, chol_variational_covar.device
        # First make the cholesky factor is upper triangular
        lower_mask = torch.ones(self.chol_variational_covar.shape[-2:], dtype=dtype, device=device).tril(0)
        chol_variational_covar = chol_variational_covar.mul(lower_mask)
        # Now construct the actual matrix (CholLinearOperator(TriangularLinearOperator(...)))
        return MultivariateNormal(self.variational_mean, _Chol(chol_variational_covar))

    def initialize_variational_distribution(self, prior_dist):
        self.variational_mean.data.copy_(prior_dist.mean)
        self.variational_mean.data.add_(torch.randn_like(prior_dist.mean), alpha=self.mean_init_std)
        # prior covariance is DiagLinearOperator(ones): its Cholesky factor is the identity
        self.chol_variational_covar.data.copy_(torch.diag_embed(prior_dist.lazy_covariance_matrix.diag.sqrt()))


class VariationalStrategy(Module):
    """Whitened variational strategy (Matthews 2017): u = L^-1 (f(Z) - mu_Z), p(u) = N(0, I)."""

    def __init__(self, model, inducing_points, variational_distribution, learn_inducing_locations=True, jitter_val=None):
        super().__init__()
        self._jitter_val = jitter_val
        # Model
        object.__setattr__(self, "model", model)
        # Inducing points
        inducing_points = inducing_points.clone()
        if inducing_points.dim() == 1:
            inducing_points = inducing_points.unsqueeze(-1)
        if learn_inducing_locations:
            self.register_parameter(name="inducing_points", parameter=torch.nn.Parameter(inducing_points))
        else:
            self.register_buffer("inducing_points", inducing_points)
        # Variational distribution
        self._variational_distribution = variational_distribution
        self.register_buffer("variational_params_initialized", torch.tensor(0))
        self.register_buffer("updated_strategy", torch.tensor(True))
        self._memoize_cache = {}

    @property
    def jitter_val(self):
        if self._jitter_val is None:
            return settings.variational_cholesky_jitter.value(dtype=self.inducing_points.dtype)
        return self._jitter_val

    def _clear_cache(self):
        self._memoize_cache = {}

    def _cholesky_factor(self, induc_induc_covar):
        """@cached(name="cholesky_factor", ignore_args=True): float64 psd_safe_cholesky of K_ZZ + jitter I."""
        if "cholesky_factor" not in self._memoize_cache:
            L = psd_safe_cholesky(induc_induc_covar.type(settings._linalg_dtype_cholesky.value()))
            self._memoize_cache["cholesky_factor"] = L
        return self._memoize_cache["cholesky_factor"]

    @property
    def prior_distribution(self):
        zeros = torch.zeros(self._variational_distribution.shape(), dtype=self._variational_distribution.variational_mean.dtype,
                            device=self._variational_distribution.variational_mean.device)
        ones = torch.ones_like(zeros)
        return MultivariateNormal(zeros, _Diag(ones))

    @property
    def variational_distribution(self):
        return self._variational_distribution()

    def kl_divergence(self):
        return kl_mvn_mvn(self.variational_distribution, self.prior_distribution)

    def forward(self, x, inducing_points, inducing_values, variational_inducing_covar=None, **kwargs):
        # Compute full prior distribution
        full_inputs = torch.cat([inducing_points, x], dim=-2)
        full_output = self.model.forward(full_inputs, **kwargs)
        full_covar = full_output.lazy_covariance_matrix
        assert isinstance(full_covar, LazyEvaluatedKernelTensor)

        # Covariance terms
        num_induc = inducing_points.size(-2)
        test_mean = full_output.mean[..., num_induc:]
        induc_induc_covar = full_covar[..., :num_induc, :num_induc].add_jitter(self.jitter_val)
        induc_data_covar = full_covar[..., :num_induc, num_induc:].to_dense()
        data_data_covar = full_covar[..., num_induc:, num_induc:]

        # Compute interpolation terms
        # K_ZZ^{-1/2} K_ZX
        L = self._cholesky_factor(induc_induc_covar)
        interp_term = torch.linalg.solve_triangular(
            L, induc_data_covar.type(settings._linalg_dtype_cholesky.value()), upper=False).to(full_inputs.dtype)

        # Compute the mean of q(f):  k_XZ K_ZZ^{-1/2} m + \mu_X
        predictive_mean = (interp_term.mT @ inducing_values.unsqueeze(-1)).squeeze(-1) + test_mean

        # Compute the covariance of q(f):  K_XX + k_XZ K_ZZ^{-1/2} (S - I) K_ZZ^{-1/2} k_ZX
        # middle_term = SumLinearOperator(variational_inducing_covar, -I); (A + B) @ rhs = A @ rhs + B @ rhs,
        # with CholLinearOperator._matmul(rhs) = L @ (L^T @ rhs)
        Lq = variational_inducing_covar.L
        middle_interp = Lq @ (Lq.mT @ interp_term) + interp_term.mul(-1)
        predictive_covar = data_data_covar.add_jitter(self.jitter_val) + interp_term.mT @ middle_interp
        return MultivariateNormal(predictive_mean, predictive_covar)

    def __call__(self, x, prior=False, **kwargs):
        # If we're in prior mode, then we're done!
        if prior:
            return self.model.forward(x, **kwargs)
        # Delete previously cached items from the training distribution
        if self.training:
            self._clear_cache()
        # (Maybe) initialize variational distribution
        if not self.variational_params_initialized.item():
            prior_dist = self.prior_distribution
            self._variational_distribution.initialize_variational_distribution(prior_dist)
            self.variational_params_initialized.fill_(1)
        # Ensure inducing_points and x are the same size
        inducing_points = self.inducing_points
        if inducing_points.shape[:-2] != x.shape[:-2]:
            raise NotImplementedError("shim: _expand_inputs is not on the hot path")
        # Get p(u)/q(u)
        variational_dist_u = self.variational_distribution
        return self.forward(x, inducing_points, inducing_values=variational_dist_u.mean,
                            variational_inducing_covar=variational_dist_u.lazy_covariance_matrix, **kwargs)
