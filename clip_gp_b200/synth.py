"""Synthetic cached-feature workloads of the shapes BASELINE.json names.

No CLIP weights or datasets are available offline, so the text bank and the image features are
drawn from a class-centre model that mimics frozen-CLIP geometry: templates of one class are
close to each other, all rows are unit-norm (reference: trainers/tip_adapter.py:101,
trainers/adapter.py:240).  The feature noise is larger than SURVEY.md 8d's 1.0: at 1.0 every shape is
100 % accurate with zero loss and zero ECE, which would make the loss / gradient / calibration parity
checks vacuous; throughput does not depend on it.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict

import torch
import torch.nn.functional as F


@dataclass(frozen=True)
class WorkloadShape:
    name: str
    C: int          # classes
    T: int          # templates per class
    D: int          # CLIP embedding dim
    d: int          # PCA dim fed to the GP kernel
    S: int          # MC samples
    shots: int
    B: int          # train minibatch
    N_test: int
    kernel: str
    noise: float = 4.0   # feature noise: chosen per shape so that zero-shot accuracy is ~60-90% and ECE is non-trivial


CONFIGS: Dict[str, WorkloadShape] = {
    # BASELINE.json configs[0..4]
    "cfg1": WorkloadShape("cfg1", C=100, T=8, D=1024, d=256, S=4, shots=4, B=128, N_test=2000, kernel="rbf", noise=10.0),
    "cfg2": WorkloadShape("cfg2", noise=6.0, C=1000, T=32, D=512, d=256, S=10, shots=16, B=128, N_test=50000, kernel="rbf"),
    "cfg3": WorkloadShape("cfg3", noise=6.0, C=1000, T=32, D=512, d=256, S=10, shots=16, B=128, N_test=50000, kernel="rbf"),
    "cfg4": WorkloadShape("cfg4", noise=6.0, C=1000, T=32, D=1024, d=256, S=10, shots=16, B=128, N_test=50000, kernel="linear"),
    "cfg5": WorkloadShape("cfg5", noise=6.0, C=397, T=64, D=512, d=256, S=100, shots=16, B=128, N_test=19850, kernel="matern"),
    # small shapes for parity tests
    "tiny": WorkloadShape("tiny", C=12, T=5, D=64, d=16, S=3, shots=4, B=16, N_test=257, kernel="rbf"),
    "small": WorkloadShape("small", C=37, T=8, D=128, d=32, S=4, shots=4, B=48, N_test=1000, kernel="rbf"),
    # the headline per-class shape (T=32, n=33) and its neighbour (n=32) on a few classes
    "t32": WorkloadShape("t32", C=19, T=32, D=128, d=64, S=5, shots=4, B=32, N_test=500, kernel="rbf"),
    "t31": WorkloadShape("t31", C=7, T=31, D=128, d=48, S=3, shots=4, B=32, N_test=500, kernel="rbf"),
    # n > 33: the general block kernels with wide CTAs (SUN397-like template counts; t40: partial 4 x 4 tiles); d ~ 2.5 n as in cfg5, so
    # that K_ZZ is not rank deficient beyond the class-mean row (with d < n the fp32 noise of the REFERENCE arithmetic exceeds the gate)
    "t64": WorkloadShape("t64", C=5, T=64, D=256, d=160, S=4, shots=4, B=32, N_test=500, kernel="rbf"),
    "t40": WorkloadShape("t40", C=6, T=40, D=192, d=100, S=3, shots=4, B=32, N_test=500, kernel="rbf"),
}


def make_text_bank(C: int, T: int, D: int, seed: int, spread: float = 0.3):
    """Class centres mu_c ~ N(0,I); E[c,t] = normalize(mu_c + spread*N(0,I)).  Returns (E [C,T,D], mu [C,D])."""
    g = torch.Generator().manual_seed(seed)
    mu = torch.randn(C, D, generator=g)
    E = F.normalize(mu.unsqueeze(1) + spread * torch.randn(C, T, D, generator=g), dim=-1)
    return E, mu


def make_features(mu: torch.Tensor, labels: torch.Tensor, seed: int, noise: float = 1.0):
    """f_i = normalize(mu_{y_i} + noise*N(0,I))."""
    g = torch.Generator().manual_seed(seed)
    return F.normalize(mu[labels] + noise * torch.randn(labels.shape[0], mu.shape[1], generator=g), dim=-1)


def make_train_set(mu: torch.Tensor, shots: int, seed: int, noise: float = 1.0):
    C = mu.shape[0]
    labels = torch.arange(C).repeat_interleave(shots)
    g = torch.Generator().manual_seed(seed + 7)
    perm = torch.randperm(labels.numel(), generator=g)
    labels = labels[perm].contiguous()
    return make_features(mu, labels, seed, noise), labels


def make_test_set(mu: torch.Tensor, N: int, seed: int, noise: float = 1.0):
    g = torch.Generator().manual_seed(seed + 13)
    labels = torch.randint(0, mu.shape[0], (N,), generator=g)
    return make_features(mu, labels, seed + 1, noise), labels


def make_workload(name: str, seed_base: int = 1234, n_test: int | None = None):
    """All host tensors for one named config (fp32 / int64, CPU)."""
    shp = CONFIGS[name]
    cfg_id = {"cfg1": 1, "cfg2": 2, "cfg3": 3, "cfg4": 4, "cfg5": 5}.get(name, 9)
    seed = seed_base + cfg_id
    E, mu = make_text_bank(shp.C, shp.T, shp.D, seed)
    f_tr, y_tr = make_train_set(mu, shp.shots, seed + 100, shp.noise)
    f_te, y_te = make_test_set(mu, n_test if n_test is not None else shp.N_test, seed + 200, shp.noise)
    return {"shape": shp, "E": E, "mu": mu, "f_train": f_tr, "y_train": y_tr, "f_test": f_te, "y_test": y_te,
            "seed": seed}


def trained_like_q(C: int, n: int, seed: int):
    """A non-trivial variational state: m ~ 0.5 N(0,1), L_q = I + 0.1 tril(N(0,1))."""
    g = torch.Generator().manual_seed(seed + 31)
    m = 0.5 * torch.randn(C, n, generator=g)
    Lq = torch.eye(n).repeat(C, 1, 1) + 0.1 * torch.randn(C, n, n, generator=g).tril()
    return m, Lq


def make_mixed_calibration_set(mu: torch.Tensor, prototypes: torch.Tensor, N: int, seed: int, noise: float = 6.0,
                               delta: float = 0.05):
    """A test set whose reliability diagram errs in BOTH directions, so that ECE (equal-width bins) and AECE (equal-mass bins)
    differ.  Half of the images are the usual noisy features (over-confident at logit scale 100).  The other half are placed
    exactly between FOUR class prototypes: f is the combination of unit rows P_y, P_j1..3 with f . P_y = 1 and f . P_jk = 1 - delta
    (a 4 x 4 Gram solve per image), so the classifier is right but only ~60-75 % confident (under-confident).  On make_test_set
    alone every bin errs the same way and both metrics collapse to |mean conf - acc|.  `prototypes` [C, D]: the unit class
    prototypes the evaluated model uses (or close to them)."""
    g = torch.Generator().manual_seed(seed + 29)
    C, D = mu.shape
    n_a = N // 2
    n_b = N - n_a
    f_a, y_a = make_test_set(mu, n_a, seed + 31, noise)
    P = F.normalize(prototypes.detach().float().cpu(), dim=-1)
    y_b = torch.randint(0, C, (n_b,), generator=g)
    offs = torch.stack([torch.randperm(C - 1, generator=g)[:3] + 1 for _ in range(n_b)])           # three distinct other classes
    idx = torch.cat([y_b[:, None], (y_b[:, None] + offs) % C], dim=1)                               # [n_b, 4]
    Ps = P[idx]                                                                                     # [n_b, 4, D]
    G = Ps @ Ps.transpose(1, 2)
    d = delta * (0.6 + 0.8 * torch.rand(n_b, 1, generator=g))
    t = torch.cat([torch.ones(n_b, 1), (1.0 - d).expand(n_b, 3)], dim=1)
    coef = torch.linalg.solve(G, t.unsqueeze(-1))                                                   # f . P_i = t_i
    f_b = F.normalize((coef * Ps).sum(1), dim=-1)
    perm = torch.randperm(N, generator=g)
    return torch.cat([f_a, f_b])[perm].contiguous(), torch.cat([y_a, y_b])[perm].contiguous()
