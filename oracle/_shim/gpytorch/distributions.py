"""gpytorch.distributions.MultivariateNormal (dense covariance) + its registered KL (kl_mvn_mvn)."""
import torch
from torch.distributions import MultivariateNormal as TMultivariateNormal  # noqa: F401  (name parity)

from linear_operator import settings as lo_settings
from linear_operator.utils.cholesky import psd_safe_cholesky

from .lazy import LazyEvaluatedKernelTensor


# Test hook (NOT part of gpytorch): when set to a callable ``f(shape, dtype, device) -> Tensor`` the base samples of
# ``rsample`` come from it instead of ``torch.randn`` -- the golden generator feeds the counter-based noise stream of the
# CUDA path (oracle/philox.py) through the unmodified reference code this way.
BASE_SAMPLES_HOOK = None


class _Chol:
    """CholLinearOperator(TriangularLinearOperator(L)): covariance given by its lower factor."""

    def __init__(self, L):
        self.L = L

    def to_dense(self):
        return self.L @ self.L.mT


class _Diag:
    """DiagLinearOperator(diag)."""

    def __init__(self, diag):
        self.diag = diag

    def to_dense(self):
        return torch.diag_embed(self.diag)


class MultivariateNormal:
    def __init__(self, mean, covariance_matrix, validate_args=False):
        self.loc = mean
        self._covar = covariance_matrix

    @property
    def mean(self):
        return self.loc

    @property
    def lazy_covariance_matrix(self):
        return self._covar

    @property
    def covariance_matrix(self):
        c = self._covar
        return c.to_dense() if isinstance(c, (LazyEvaluatedKernelTensor, _Chol, _Diag)) else c

    @property
    def batch_shape(self):
        return self.loc.shape[:-1]

    @property
    def event_shape(self):
        return self.loc.shape[-1:]

    def _root(self):
        """LinearOperator.root_decomposition() for size <= max_cholesky_size: psd_safe_cholesky of the
        evaluated matrix in ITS OWN dtype (fp32 on the hot path), jitter retries 1e-6, 1e-5, 1e-4."""
        c = self._covar
        if isinstance(c, _Chol):
            return c.L
        if isinstance(c, _Diag):
            return torch.diag_embed(c.diag.sqrt())
        dense = self.covariance_matrix
        if dense.size(-1) > lo_settings.max_cholesky_size.value():
            raise NotImplementedError("shim: Lanczos root decomposition is not on the hot path")
        if dense.shape[-2:] == torch.Size([1, 1]):
            return dense.clamp_min(0.0).sqrt()
        return psd_safe_cholesky(dense).contiguous()

    def rsample(self, sample_shape=torch.Size(), base_samples=None):
        """MultivariateNormal.rsample -> LinearOperator.zero_mean_mvn_samples:
        base_samples = randn(*batch, N, num_samples); samples = (root @ base).permute(-1, batch..., N) + loc."""
        if base_samples is not None:
            raise NotImplementedError("shim: explicit base_samples are not used by the reference")
        num_samples = sample_shape.numel() or 1
        covar_root = self._root()
        dense_dim = covar_root.dim()
        shape = (*self.batch_shape, covar_root.size(-1), num_samples)
        if BASE_SAMPLES_HOOK is not None:
            base = BASE_SAMPLES_HOOK(shape, self.loc.dtype, self.loc.device)
        else:
            base = torch.randn(*shape, dtype=self.loc.dtype, device=self.loc.device)
        samples = covar_root.matmul(base).permute(-1, *range(dense_dim - 1)).contiguous()
        res = samples + self.loc.unsqueeze(0)
        return res.view(sample_shape + self.loc.shape)


def kl_mvn_mvn(p_dist, q_dist):
    """gpytorch.distributions.multivariate_normal.kl_mvn_mvn for p = N(m, L L^T) (CholLinearOperator),
    q = N(0, I) (DiagLinearOperator of ones)."""
    q_mean, q_covar = q_dist.loc, q_dist.lazy_covariance_matrix
    p_mean, p_covar = p_dist.loc, p_dist.lazy_covariance_matrix
    assert isinstance(p_covar, _Chol) and isinstance(q_covar, _Diag)
    root_p_covar = p_covar.L                                            # root_decomposition().root.to_dense()
    mean_diffs = p_mean - q_mean
    inv_quad_rhs = torch.cat([mean_diffs.unsqueeze(-1), root_p_covar], -1)
    logdet_p_covar = p_covar.L.diagonal(dim1=-2, dim2=-1).pow(2).log().sum(-1)   # CholLinearOperator.logdet
    # DiagLinearOperator.inv_quad_logdet
    trace_plus_inv_quad_form = inv_quad_rhs.div(q_covar.diag.unsqueeze(-1)).mul(inv_quad_rhs).sum((-2, -1))
    logdet_q_covar = q_covar.diag.log().sum(-1)
    return 0.5 * sum([logdet_q_covar, logdet_p_covar.mul(-1), trace_plus_inv_quad_form, -float(mean_diffs.size(-1))])
