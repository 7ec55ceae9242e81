"""Hand-derived forward/backward of the GP template weighter, written as batched torch on the CPU.

TEST INFRASTRUCTURE (see oracle/__init__.py).  This file is the *algorithm sheet* of the CUDA
kernels in clip_gp_b200/csrc/gp_*.cu: same intermediate quantities, same adjoint formulas, no
autograd.  tests/test_oracle_gp.py checks it against ``torch.autograd`` through ``oracle.gp`` so the
derivation is validated on the CPU before it is trusted on the GPU.

Notation per class: Z [n,d] inducing points, X [T,d] templates (test inputs), l lengthscale,
K_ZZ (+1e-4 I), K_ZX, K_XX, L = chol64(K_ZZ), A = L^-1 K_ZX, Lq = tril(chol_var),
Bm = Lq^T A, mu = A^T m + mean_x, Sigma = K_XX + 1e-4 I + Bm^T Bm - A^T A, R = chol32(Sigma),
f_s = mu + R eps_s, w_s = sparsemax(f_s).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

JIT = 1e-4


def gram(kind, a, b, ls=None, os_=None, var=None):
    """Direct-difference Gram matrix (what the CUDA kernel computes). a [C,na,d], b [C,nb,d]."""
    if kind == "linear":
        return var.view(-1, 1, 1) * (a @ b.transpose(-1, -2))
    diff = (a.unsqueeze(2) - b.unsqueeze(1)) / ls.unsqueeze(1).unsqueeze(1)   # [C,na,nb,d]
    r2 = diff.pow(2).sum(-1)
    if kind == "rbf":
        return os_.view(-1, 1, 1) * torch.exp(-0.5 * r2)
    if kind == "matern":
        return torch.exp(-torch.sqrt(r2.clamp_min(1e-30)))
    raise ValueError(kind)


def sparsemax_fwd(f):
    """Sort-free sparsemax over the last dim: rank by (value desc, index asc)."""
    z = f - f.max(-1, keepdim=True).values
    zi, zj = z.unsqueeze(-1), z.unsqueeze(-2)                           # i: element, j: other
    idx = torch.arange(z.shape[-1])
    before = (zj > zi) | ((zj == zi) & (idx.view(1, -1) <= idx.view(-1, 1)))   # j sorted at or before i
    k = before.sum(-1).to(z.dtype)
    cs = (before.to(z.dtype) * zj).sum(-1)
    supp = k * z > cs - 1
    ksz = supp.sum(-1, keepdim=True)
    tau = ((supp.to(z.dtype) * z).sum(-1, keepdim=True) - 1) / ksz.to(z.dtype)
    return torch.clamp(z - tau, min=0), ksz


def chol_bwd(L, dL):
    """Adjoint of L = chol(A): dA = 1/2 (X + X^T), X = L^-T Phi(L^T dL) L^-1, Phi = tril with halved diagonal."""
    P = L.transpose(-1, -2) @ dL.tril()
    P = P.tril()
    P = P - 0.5 * torch.diag_embed(P.diagonal(dim1=-2, dim2=-1))
    X = torch.linalg.solve_triangular(L.transpose(-1, -2), P, upper=True, left=True)
    X = torch.linalg.solve_triangular(L, X, upper=False, left=False)
    return 0.5 * (X + X.transpose(-1, -2))


def forward(kind, Z, X, raw_ls, raw_os, raw_var, m, chol_var, mean_x, eps):
    """Returns (w [S,C,T], kl [C], saved)."""
    C, n, d = Z.shape
    T = X.shape[1]
    ls = F.softplus(raw_ls).view(C, d) if raw_ls is not None else None
    os_ = F.softplus(raw_os) if raw_os is not None else None
    var = F.softplus(raw_var).view(C) if raw_var is not None else None
    K_ZZ = gram(kind, Z, Z, ls, os_, var) + JIT * torch.eye(n)
    K_ZX = gram(kind, Z, X, ls, os_, var)
    K_XX = gram(kind, X, X, ls, os_, var)
    L = torch.linalg.cholesky(K_ZZ.double())
    A64 = torch.linalg.solve_triangular(L, K_ZX.double(), upper=False)
    A = A64.float()
    Lq = chol_var.tril()
    Bm = Lq.transpose(-1, -2) @ A
    mu = (A.transpose(-1, -2) @ m.unsqueeze(-1)).squeeze(-1) + mean_x
    Sigma = K_XX + JIT * torch.eye(T) + Bm.transpose(-1, -2) @ Bm - A.transpose(-1, -2) @ A
    R = torch.linalg.cholesky(Sigma)
    f = (R @ eps).permute(2, 0, 1) + mu.unsqueeze(0)                     # [S,C,T]
    w, ksz = sparsemax_fwd(f)
    dg = Lq.diagonal(dim1=-2, dim2=-1)
    kl = 0.5 * (Lq.pow(2).sum((-2, -1)) + m.pow(2).sum(-1) - n - dg.pow(2).log().sum(-1))
    saved = dict(kind=kind, Z=Z, X=X, ls=ls, os=os_, var=var, raw_ls=raw_ls, raw_os=raw_os, raw_var=raw_var,
                 K_ZZ=K_ZZ, K_ZX=K_ZX, K_XX=K_XX, L=L, A=A, Lq=Lq, Bm=Bm, R=R, w=w, ksz=ksz, m=m, eps=eps)
    return w, kl, saved


def gram_bwd(kind, a, b, K, dK, ls, os_, var, need_a=True, need_b=True):
    """Adjoints of one Gram block.  Returns dict(dls [C,d], dos [C], dvar [C], da, db)."""
    out = {}
    if kind == "linear":
        ab = a @ b.transpose(-1, -2)
        out["dvar"] = (dK * ab).sum((-2, -1))
        out["da"] = var.view(-1, 1, 1) * (dK @ b) if need_a else None
        out["db"] = var.view(-1, 1, 1) * (dK.transpose(-1, -2) @ a) if need_b else None
        return out
    diff = a.unsqueeze(2) - b.unsqueeze(1)                                # [C,na,nb,d]
    G = dK * K
    if kind == "rbf":
        out["dos"] = G.sum((-2, -1)) / os_
        Wt = G                                                           # dK/dr2 = -1/2 K  ->  -1/2 G
        coef = -0.5
    else:
        r2 = (diff / ls.unsqueeze(1).unsqueeze(1)).pow(2).sum(-1)
        r = torch.sqrt(r2.clamp_min(1e-30))
        Wt = torch.where(r2 > 1e-30, G / r, torch.zeros_like(G))         # dK/dr2 = -K/(2r); clamp kills the grad
        coef = -0.5
    # dr2 = coef * Wt ; r2 = sum_k diff_k^2 / ls_k^2
    out["dls"] = (coef * Wt.unsqueeze(-1) * (-2.0) * diff.pow(2)).sum((1, 2)) / ls.pow(3)
    g_diff = coef * Wt.unsqueeze(-1) * 2.0 * diff / ls.pow(2).unsqueeze(1).unsqueeze(1)
    out["da"] = g_diff.sum(2) if need_a else None
    out["db"] = -g_diff.sum(1) if need_b else None
    return out


def backward(saved, dw, dkl):
    """Adjoints given dw [S,C,T] and dkl [C] (upstream gradient of the per-class KL).

    Returns dict with dZ (all rows; the module masks rows < T), draw_ls, draw_os, draw_var, dm, dchol, dmean_x.
    """
    sv = saved
    kind, Z, X, L, A, Lq, Bm, R, w, m, eps = (sv[k] for k in ("kind", "Z", "X", "L", "A", "Lq", "Bm", "R", "w", "m", "eps"))
    C, n, d = Z.shape
    T = X.shape[1]
    # sparsemax adjoint (entmax): zero outside the support, subtract the support mean
    g = torch.where(w > 0, dw, torch.zeros_like(dw))
    vhat = g.sum(-1, keepdim=True) / sv["ksz"].to(dw.dtype)
    df = torch.where(w > 0, g - vhat, g)                                  # [S,C,T]
    dmu = df.sum(0)                                                       # [C,T]
    dR = torch.einsum("sct,cks->ctk", df, eps).tril()                     # [C,T,T]
    dSigma = chol_bwd(R, dR)
    # Sigma = K_XX + jI + Bm^T Bm - A^T A ;  Bm = Lq^T A ; mu = A^T m + mean_x
    dK_XX = dSigma
    dBm = 2.0 * Bm @ dSigma
    dA = -2.0 * A @ dSigma + Lq @ dBm + m.unsqueeze(-1) * dmu.unsqueeze(-2)
    dLq = (A @ dBm.transpose(-1, -2)).tril()
    dm = (A @ dmu.unsqueeze(-1)).squeeze(-1)
    # KL
    dg = Lq.diagonal(dim1=-2, dim2=-1)
    dm = dm + dkl.unsqueeze(-1) * m
    dLq = dLq + dkl.view(-1, 1, 1) * (Lq - torch.diag_embed(1.0 / dg))
    # A = L^-1 K_ZX  (float64)
    dA64 = dA.double()
    dK_ZX64 = torch.linalg.solve_triangular(L.transpose(-1, -2), dA64, upper=True)
    dL = -(dK_ZX64 @ A.double().transpose(-1, -2)).tril()
    dK_ZZ = chol_bwd(L, dL).float()
    dK_ZX = dK_ZX64.float()
    # kernel adjoints
    ls, os_, var = sv["ls"], sv["os"], sv["var"]
    K_ZZ0 = sv["K_ZZ"] - JIT * torch.eye(n)
    gzz = gram_bwd(kind, Z, Z, K_ZZ0, dK_ZZ, ls, os_, var)
    gzx = gram_bwd(kind, Z, X, sv["K_ZX"], dK_ZX, ls, os_, var, need_b=False)
    gxx = gram_bwd(kind, X, X, sv["K_XX"], dK_XX, ls, os_, var, need_a=False, need_b=False)
    out = {"dm": dm, "dchol": dLq, "dmean_x": dmu}
    out["dZ"] = gzz["da"] + gzz["db"] + gzx["da"]
    if kind != "linear":
        dls = gzz["dls"] + gzx["dls"] + gxx["dls"]
        out["draw_ls"] = (dls * torch.sigmoid(sv["raw_ls"].view(C, d))).view_as(sv["raw_ls"])
    if kind == "rbf":
        dos = gzz["dos"] + gzx["dos"] + gxx["dos"]
        out["draw_os"] = dos * torch.sigmoid(sv["raw_os"])
    if kind == "linear":
        dvar = gzz["dvar"] + gzx["dvar"] + gxx["dvar"]
        out["draw_var"] = (dvar * torch.sigmoid(sv["raw_var"].view(C))).view_as(sv["raw_var"])
    return out
