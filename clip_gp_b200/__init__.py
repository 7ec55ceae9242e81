"""clip_gp_b200 — B200-native (sm_100a) implementation of the CLIP-GP few-shot adapter hot path.

Host code is Python/PyTorch; the arithmetic runs in hand-written CUDA kernels behind the C ABI of
include/clipgp.h (libclipgp.so, loaded with ctypes).  No CPU fallback.  See DESIGN.md.
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
