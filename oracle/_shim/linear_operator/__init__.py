"""Minimal linear_operator stand-in (dense tensors) — TEST INFRASTRUCTURE, see oracle/_shim/README.md."""
from . import settings, utils  # noqa: F401
