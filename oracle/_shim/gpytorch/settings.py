"""gpytorch.settings: the values the hot path reads."""
import torch

from linear_operator.settings import (  # noqa: F401
    _linalg_dtype_cholesky, cholesky_jitter, cholesky_max_tries, max_cholesky_size)


class variational_cholesky_jitter:
    """Jitter added to K_ZZ (and K_XX) by the variational strategies: 1e-4 float, 1e-6 double."""
    _global_float_value = 1e-4
    _global_double_value = 1e-6

    @classmethod
    def value(cls, dtype):
        if torch.is_tensor(dtype):
            dtype = dtype.dtype
        if dtype == torch.float:
            return cls._global_float_value
        if dtype == torch.double:
            return cls._global_double_value
        raise RuntimeError(f"Unsupported dtype for {cls.__name__}.")
