from . import cholesky  # noqa: F401
