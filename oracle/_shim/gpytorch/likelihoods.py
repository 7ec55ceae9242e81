"""gpytorch.likelihoods.GaussianLikelihood: parameters only (unused by the hot path, gp_template_weigher.py:126)."""
import torch

from .constraints import GreaterThan
from .module import Module


class HomoskedasticNoise(Module):
    def __init__(self, noise_prior=None, noise_constraint=None, batch_shape=torch.Size(), num_tasks=1):
        super().__init__()
        if noise_constraint is None:
            noise_constraint = GreaterThan(1e-4)
        self.register_parameter(name="raw_noise", parameter=torch.nn.Parameter(torch.zeros(*batch_shape, num_tasks)))
        self.register_constraint("raw_noise", noise_constraint)

    @property
    def noise(self):
        return self.raw_noise_constraint.transform(self.raw_noise)


class GaussianLikelihood(Module):
    def __init__(self, noise_prior=None, noise_constraint=None, batch_shape=torch.Size(), **kwargs):
        super().__init__()
        self.noise_covar = HomoskedasticNoise(noise_prior=noise_prior, noise_constraint=noise_constraint, batch_shape=batch_shape)

    @property
    def noise(self):
        return self.noise_covar.noise
