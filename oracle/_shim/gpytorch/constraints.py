"""gpytorch.constraints: Interval / GreaterThan / Positive with the softplus transform."""
import math

import torch
from torch import nn
from torch.nn.functional import softplus


def inv_softplus(x):
    """gpytorch.utils.transforms.inv_softplus."""
    return x + torch.log(-torch.expm1(-x))


class Interval(nn.Module):
    def __init__(self, lower_bound, upper_bound, transform=torch.sigmoid, inv_transform=None, initial_value=None):
        super().__init__()
        self.register_buffer("lower_bound", torch.as_tensor(lower_bound, dtype=torch.get_default_dtype()))
        self.register_buffer("upper_bound", torch.as_tensor(upper_bound, dtype=torch.get_default_dtype()))
        self._transform = transform
        self._inv_transform = inv_transform
        self._initial_value = initial_value

    @property
    def initial_value(self):
        return self._initial_value

    @property
    def enforced(self):
        return self._transform is not None


class GreaterThan(Interval):
    def __init__(self, lower_bound, transform=softplus, inv_transform=inv_softplus, initial_value=None):
        super().__init__(lower_bound=lower_bound, upper_bound=math.inf, transform=transform,
                         inv_transform=inv_transform, initial_value=initial_value)

    def transform(self, tensor):
        return self._transform(tensor) + self.lower_bound if self.enforced else tensor

    def inverse_transform(self, transformed_tensor):
        return self._inv_transform(transformed_tensor - self.lower_bound) if self.enforced else transformed_tensor


class Positive(GreaterThan):
    def __init__(self, transform=softplus, inv_transform=inv_softplus, initial_value=None):
        super().__init__(lower_bound=0.0, transform=transform, inv_transform=inv_transform, initial_value=initial_value)

    def transform(self, tensor):
        return self._transform(tensor) if self.enforced else tensor

    def inverse_transform(self, transformed_tensor):
        return self._inv_transform(transformed_tensor) if self.enforced else transformed_tensor
