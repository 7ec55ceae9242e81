"""CPU oracle: the counter-based base-noise stream of the perf path (TEST INFRASTRUCTURE).

Restates clip_gp_b200/csrc/common.cuh::philox_normal: Philox4x32-10 keyed by (seed), counter
(idx_lo, idx_hi, step_lo, step_hi), two uniforms u = (x + 0.5) * 2^-32 in fp32, Box-Muller
z = sqrt(-2 ln u0) * cos(2 pi u1).  The reference draws torch.randn at this point
(MultivariateNormal.rsample); any i.i.d. N(0,1) stream is distribution-equivalent, and parity tests
feed identical noise to both sides.
"""
from __future__ import annotations

import numpy as np
import torch

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = (np.asarray(x, dtype=np.uint32) for x in (c0, c1, c2, c3))
    k0 = np.uint32(k0); k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            n0 = ((p1 >> np.uint64(32)) & MASK).astype(np.uint32) ^ c1 ^ k0
            n1 = (p1 & MASK).astype(np.uint32)
            n2 = ((p0 >> np.uint64(32)) & MASK).astype(np.uint32) ^ c3 ^ k1
            n3 = (p0 & MASK).astype(np.uint32)
            c0, c1, c2, c3 = n0, n1, n2, n3
            k0 = np.uint32((int(k0) + int(W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def normal(seed: int, step: int, idx: np.ndarray) -> np.ndarray:
    idx = np.asarray(idx, dtype=np.uint64)
    lo = (idx & MASK).astype(np.uint32)
    hi = (idx >> np.uint64(32)).astype(np.uint32)
    z = np.zeros_like(lo)
    o0, o1, _, _ = philox4x32_10(lo, hi, z + np.uint32(step & 0xFFFFFFFF), z + np.uint32((step >> 32) & 0xFFFFFFFF),
                                 seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    scale = np.float32(2.3283064365386963e-10)
    u0 = (o0.astype(np.float32) + np.float32(0.5)) * scale
    u1 = (o1.astype(np.float32) + np.float32(0.5)) * scale
    r = np.sqrt(np.float32(-2.0) * np.log(u0))
    return (r * np.cos(np.float32(2.0 * np.pi) * u1.astype(np.float64)).astype(np.float32)).astype(np.float32)


def eps_tensor(seed: int, step: int, C: int, T: int, S: int, s_offset: int = 0, S_total: int | None = None) -> torch.Tensor:
    """[C, T, S] base noise the GP kernel draws for (seed, step): idx = (c*T + t) * S_total + s_offset + s."""
    S_total = S if S_total is None else S_total
    c = np.arange(C, dtype=np.uint64)[:, None, None]
    t = np.arange(T, dtype=np.uint64)[None, :, None]
    s = np.arange(S, dtype=np.uint64)[None, None, :]
    idx = (c * np.uint64(T) + t) * np.uint64(S_total) + np.uint64(s_offset) + s
    return torch.from_numpy(normal(seed, step, idx))
