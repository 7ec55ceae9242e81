"""linear_operator.settings: only the values the hot path reads (float32 / float64 defaults)."""
import torch


class _dtype_value_context:
    _global_float_value = None
    _global_double_value = None
    _global_half_value = None

    @classmethod
    def value(cls, dtype):
        if torch.is_tensor(dtype):
            dtype = dtype.dtype
        if dtype == torch.float:
            return cls._global_float_value
        elif dtype == torch.double:
            return cls._global_double_value
        elif dtype == torch.half:
            return cls._global_half_value
        raise RuntimeError(f"Unsupported dtype for {cls.__name__}.")


class cholesky_jitter(_dtype_value_context):
    """The jitter value used by psd_safe_cholesky when using cholesky solves (1e-6 float, 1e-8 double)."""
    _global_float_value = 1e-6
    _global_double_value = 1e-8


class cholesky_max_tries:
    _global_value = 3

    @classmethod
    def value(cls):
        return cls._global_value


class max_cholesky_size:
    _global_value = 800

    @classmethod
    def value(cls):
        return cls._global_value


class _linalg_dtype_cholesky:
    """dtype of the variational strategy's Cholesky factor and triangular solve (default float64)."""
    _global_value = torch.double

    @classmethod
    def value(cls):
        return cls._global_value
