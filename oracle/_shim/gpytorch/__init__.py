"""Minimal gpytorch stand-in (dense tensors) — TEST INFRASTRUCTURE, see oracle/_shim/README.md.

Restates, routine by routine, the gpytorch 1.11-1.14 code paths reached from
/root/reference/trainers/gp_template_weigher.py (:8, :62-63, :110-120, :126, :173, :215, :225)
and the KL call sites (adapter.py:463, taskres.py:271, clip_adapter.py:274, tip_adapter.py:141).
"""
from . import constraints, distributions, kernels, lazy, likelihoods, means, models, settings, variational  # noqa: F401
from .module import Module  # noqa: F401

__version__ = "1.13+shim"
