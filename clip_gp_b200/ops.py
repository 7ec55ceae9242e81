"""``torch.autograd.Function`` wrappers over the clipgp C ABI (the differentiable drop-in surface).

Everything here launches hand-written sm_100a kernels from libclipgp.so on the caller's current CUDA
stream; torch only owns the memory.  Gradients are the hand-derived adjoints (gp_backward.cu,
proto.cu), not autograd traces.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import GpArgs, GpBwdArgs, KERNEL_IDS


def _c(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    if t is None:
        return None
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


class GPWeightsFunction(torch.autograd.Function):
    """(Z, X, raw_ls, raw_os, raw_var, m, chol, mean_x, eps) -> (w [S,C,T], kl [C]).

    Replaces gp_template_weigher.py:213-217 (``self(gp_input)``, ``rsample``, ``sparsemax``) and
    ``variational_strategy.kl_divergence()`` (adapter.py:463).  eps is the explicit base noise
    ``[C,Nx,S]`` with Nx >= T (only rows < T are read; a strided view is fine), or None for the
    on-device Philox stream keyed by ``rng_state = [seed, step]`` (int64 tensor on the device).
    """

    @staticmethod
    def forward(ctx, Z, X, raw_ls, raw_os, raw_var, var_mean, chol_var, mean_x, eps, kernel_type: str, S: int,
                rng_state, s_offset: int, S_total: int, alias_check: bool):
        dev = _lib.require_cuda(Z, X, var_mean, chol_var)
        lib = _lib.load()
        Z, X, var_mean, chol_var = _c(Z), _c(X), _c(var_mean), _c(chol_var)
        raw_ls, raw_os, raw_var, mean_x = _c(raw_ls), _c(raw_os), _c(raw_var), _c(mean_x)
        Cn, n, d = Z.shape
        T = X.shape[1]
        if X.shape[0] != Cn or X.shape[2] != d:
            raise ValueError(f"X {tuple(X.shape)} does not match Z {tuple(Z.shape)}")
        if eps is not None:
            if eps.dtype != torch.float32:
                eps = eps.float()
            if eps.shape[0] != Cn or eps.shape[1] < T or eps.shape[2] != S:
                raise ValueError(f"eps must be [C, >=T, S]; got {tuple(eps.shape)}")
        elif rng_state is None:
            raise ValueError("need eps or rng_state")
        w = torch.empty(S, Cn, T, dtype=torch.float32, device=dev)
        kl = torch.empty(Cn, dtype=torch.float32, device=dev)
        Lf = torch.empty(Cn, n, n, dtype=torch.float64, device=dev)
        Af = torch.empty(Cn, n, T, dtype=torch.float32, device=dev)
        Rf = torch.empty(Cn, T, T, dtype=torch.float32, device=dev)
        status = torch.empty(Cn, dtype=torch.int32, device=dev)
        Ksave = torch.empty(Cn, 1 + n * n + n * T + T * T, dtype=torch.float32, device=dev)
        a = GpArgs()
        a.kernel_type = KERNEL_IDS[kernel_type]
        a.x_is_z_prefix = 1 if (alias_check and n >= T) else 0
        a.C, a.T, a.n, a.d, a.S = Cn, T, n, d, S
        a.Z, a.X = Z.data_ptr(), X.data_ptr()
        a.raw_lengthscale, a.raw_outputscale, a.raw_variance = _lib.ptr(raw_ls), _lib.ptr(raw_os), _lib.ptr(raw_var)
        a.var_mean, a.chol_var, a.mean_x = var_mean.data_ptr(), chol_var.data_ptr(), _lib.ptr(mean_x)
        if eps is not None:
            a.eps = eps.data_ptr()
            a.eps_sc, a.eps_st, a.eps_ss = eps.stride(0), eps.stride(1), eps.stride(2)
            a.rng_state = None
        rng_snap = eps_save = None
        else_branch = eps is None
        if else_branch:
            # counter RNG: the adjoint must see the forward's draw even if the caller advances its (seed, step) counter before
            # backward runs -> private snapshot of the state + the drawn noise itself (warp path re-reads it, general path
            # regenerates from the snapshot)
            a.eps = None
            rng_snap = rng_state.detach().clone()
            a.rng_state = rng_snap.data_ptr()
            eps_save = torch.empty(S, Cn, T, dtype=torch.float32, device=dev)
            a.eps_save = eps_save.data_ptr()
        a.s_offset, a.S_total = int(s_offset), int(S_total if S_total else S)
        a.w, a.kl, a.L, a.A, a.R, a.status = (w.data_ptr(), kl.data_ptr(), Lf.data_ptr(), Af.data_ptr(), Rf.data_ptr(),
                                              status.data_ptr())
        a.Ksave = Ksave.data_ptr()
        with torch.cuda.device(dev):
            _lib.check(lib.clipgp_gp_forward(C.byref(a), _lib.stream_ptr(dev)), "clipgp_gp_forward")
        ctx.kernel_type = kernel_type
        ctx.args = a
        # inputs and outputs go through save_for_backward so that autograd's version counters catch an in-place update
        # (optimizer.step(), a hand-written copy_) between forward and backward; kernel-private buffers ride on ctx
        ctx.save_for_backward(Z, X, raw_ls, raw_os, raw_var, var_mean, chol_var, mean_x, eps, w)
        ctx.keep = (rng_snap, eps_save, Lf, Af, Rf)
        ctx.ksave = Ksave
        ctx.status = status
        ctx.mark_non_differentiable(status)
        return w, kl, status

    @staticmethod
    def backward(ctx, dw, dkl, _dstatus):
        a = ctx.args
        (Z, X, raw_ls, raw_os, raw_var, var_mean, chol_var, mean_x, eps, w) = ctx.saved_tensors
        dev = Z.device
        lib = _lib.load()
        Cn, n, d = Z.shape
        T = X.shape[1]
        dw = torch.zeros_like(w) if dw is None else _c(dw)
        dkl = None if dkl is None else _c(dkl)
        dZ = torch.zeros_like(Z)                      # rows < T stay zero (reference masks them, :72-79)
        dZ_last = torch.empty(Cn, d, dtype=torch.float32, device=dev)
        dls = torch.empty_like(raw_ls) if raw_ls is not None else None
        dos = torch.empty_like(raw_os) if raw_os is not None else None
        dvar = torch.empty_like(raw_var) if raw_var is not None else None
        dm = torch.empty_like(var_mean)
        dchol = torch.empty_like(chol_var)
        dmean = torch.empty_like(mean_x) if mean_x is not None else None
        b = GpBwdArgs()
        b.dw = dw.data_ptr()
        b.dkl = _lib.ptr(dkl)
        b.dkl_scalar = 0.0
        b.dZ_last = dZ_last.data_ptr()
        b.draw_lengthscale, b.draw_outputscale, b.draw_variance = _lib.ptr(dls), _lib.ptr(dos), _lib.ptr(dvar)
        b.dvar_mean, b.dchol_var, b.dmean_x = dm.data_ptr(), dchol.data_ptr(), _lib.ptr(dmean)
        with torch.cuda.device(dev):
            _lib.check(lib.clipgp_gp_backward(C.byref(a), C.byref(b), _lib.stream_ptr(dev)), "clipgp_gp_backward")
        dZ[:, n - 1, :] = dZ_last
        return (dZ, None, dls, dos, dvar, dm, dchol, dmean, None, None, None, None, None, None, None)


def gp_weights(Z, X, raw_ls, raw_os, raw_var, var_mean, chol_var, mean_x, eps, kernel_type: str, S: int,
               rng_state=None, s_offset: int = 0, S_total: int = 0, alias_check: bool = True):
    return GPWeightsFunction.apply(Z, X, raw_ls, raw_os, raw_var, var_mean, chol_var, mean_x, eps, kernel_type, int(S),
                                   rng_state, s_offset, S_total, alias_check)


class PrototypesFunction(torch.autograd.Function):
    """w [S,C,T], E [C,T,D] -> un-normalised prototypes [S,C,D] (gp_template_weigher.py:221)."""

    @staticmethod
    def forward(ctx, w, E):
        dev = _lib.require_cuda(w, E)
        lib = _lib.load()
        w, E = _c(w), _c(E)
        S, Cn, T = w.shape
        D = E.shape[2]
        P = torch.empty(S, Cn, D, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.clipgp_proto_forward(w.data_ptr(), E.data_ptr(), S, Cn, T, D, None, 0.0, P.data_ptr(), None, None,
                                                None, None, None, 0, _lib.stream_ptr(dev)), "clipgp_proto_forward")
        ctx.save_for_backward(E)
        ctx.shape = (S, Cn, T, D)
        return P

    @staticmethod
    def backward(ctx, dP):
        (E,) = ctx.saved_tensors
        S, Cn, T, D = ctx.shape
        lib = _lib.load()
        dev = E.device
        dP = _c(dP)
        dw = torch.empty(S, Cn, T, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.clipgp_proto_backward(dP.data_ptr(), Cn * D, 1.0, None, None, E.data_ptr(), S, Cn, T, D,
                                                 dw.data_ptr(), _lib.stream_ptr(dev)), "clipgp_proto_backward")
        return dw, None


def prototypes(w, E):
    return PrototypesFunction.apply(w, E)


def prototypes_reduced(w, E, residual=None, alpha: float = 0.0, want_hat=False, want_mean_hat=False, want_mean_raw=False):
    """Inference-side prototype products (no autograd): unit rows, the collapsed logit-mean prototype
    ``(1/S) sum_s p_hat_s`` and the prototype-init ``normalize(mean_s P_s)``."""
    dev = _lib.require_cuda(w, E)
    lib = _lib.load()
    w, E = _c(w.detach()), _c(E.detach())
    residual = _c(residual.detach()) if residual is not None else None
    S, Cn, T = w.shape
    D = E.shape[2]
    P_hat = torch.empty(S, Cn, D, dtype=torch.float32, device=dev) if want_hat else None
    mh = torch.empty(Cn, D, dtype=torch.float32, device=dev) if want_mean_hat else None
    mr = torch.empty(Cn, D, dtype=torch.float32, device=dev) if want_mean_raw else None
    with torch.cuda.device(dev):
        _lib.check(lib.clipgp_proto_forward(w.data_ptr(), E.data_ptr(), S, Cn, T, D, _lib.ptr(residual), float(alpha), None,
                                            _lib.ptr(P_hat), None, None, _lib.ptr(mh), _lib.ptr(mr), 1,
                                            _lib.stream_ptr(dev)), "clipgp_proto_forward")
    return P_hat, mh, mr


# =====================================================================================================
# Generic differentiable building blocks of the cosine-logit heads (exact-fp32 kernels)
# =====================================================================================================
def _gemm(A, sam, sak, B, sbk, sbn, out, M, N, K, alpha, accumulate=0):
    dev = out.device
    with torch.cuda.device(dev):
        _lib.check(_lib.load().clipgp_gemm_f32(A.data_ptr(), sam, sak, B.data_ptr(), sbk, sbn, out.data_ptr(), out.stride(0), M, N, K,
                                               float(alpha), accumulate, _lib.stream_ptr(dev)), "clipgp_gemm_f32")


class MatmulNT(torch.autograd.Function):
    """C = alpha * A @ B^T for A [M,K], B [N,K] (adapter.py:239 f W^T; :251 / :426 f_hat p_hat^T; tip_adapter.py:250)."""

    @staticmethod
    def forward(ctx, A, B, alpha: float):
        dev = _lib.require_cuda(A, B)
        A, B = _c(A), _c(B)
        M, K = A.shape
        N = B.shape[0]
        if B.shape[1] != K:
            raise ValueError(f"matmul_nt: {tuple(A.shape)} x {tuple(B.shape)}^T")
        out = torch.empty(M, N, dtype=torch.float32, device=dev)
        if M and N:
            _gemm(A, K, 1, B, 1, K, out, M, N, K, alpha)
        ctx.save_for_backward(A, B)
        ctx.alpha = float(alpha)
        return out

    @staticmethod
    def backward(ctx, dC):
        A, B = ctx.saved_tensors
        dC = _c(dC)
        M, K = A.shape
        N = B.shape[0]
        dA = dB = None
        if ctx.needs_input_grad[0]:
            dA = torch.empty_like(A)
            _gemm(dC, N, 1, B, K, 1, dA, M, K, N, ctx.alpha)          # dA = alpha dC B
        if ctx.needs_input_grad[1]:
            dB = torch.empty_like(B)
            _gemm(dC, 1, N, A, K, 1, dB, N, K, M, ctx.alpha)          # dB = alpha dC^T A
        return dA, dB, None


class MatmulNT_TF32(torch.autograd.Function):
    """MatmulNT on the tensor cores: tcgen05 kind::tf32 over the fp32 tensors in place (csrc/gemm_tc.cu); the two adjoint
    products read their operands transposed in place (MN-major), so neither direction makes a cast or a transposed copy.
    TF32 is the reference's own GPU arithmetic (trainers/adapter.py:23, allow_tf32)."""

    @staticmethod
    def forward(ctx, A, B, alpha: float):
        from . import tc
        _lib.require_cuda(A, B)
        A, B = _c(A), _c(B)
        if B.shape[1] != A.shape[1]:
            raise ValueError(f"matmul_nt: {tuple(A.shape)} x {tuple(B.shape)}^T")
        out = tc.gemm_tf32(A, B, float(alpha))
        ctx.save_for_backward(A, B)
        ctx.alpha = float(alpha)
        return out

    @staticmethod
    def backward(ctx, dC):
        from . import tc
        A, B = ctx.saved_tensors
        dC = _c(dC)
        dA = dB = None
        if ctx.needs_input_grad[0]:
            dA = tc.gemm_tf32(dC, B, ctx.alpha, b_t=True, split_k=True)              # dA = alpha dC B        (B [N,K] read as [K(=N), .])
        if ctx.needs_input_grad[1]:
            dB = tc.gemm_tf32(dC, A, ctx.alpha, a_t=True, b_t=True, split_k=True)    # dB = alpha dC^T A      (contraction over rows)
        return dA, dB, None


def tf32_ok(A, B) -> bool:
    """Row pitches of all three products (forward and both adjoints) are multiples of 16 bytes and there is something to do."""
    M, K = A.shape
    N = B.shape[0]
    return M > 0 and N > 0 and K % 4 == 0 and N % 4 == 0 and A.is_cuda


def matmul_nt(A, B, alpha: float = 1.0, precision: str = "fp32"):
    """alpha * A @ B^T.  precision: "fp32" (FFMA kernel, exact comparator) | "tf32" (tensor cores; falls back to fp32 when a row
    pitch is not a multiple of 16 bytes)."""
    if precision == "tf32" and tf32_ok(A, B):
        return MatmulNT_TF32.apply(A, B, alpha)
    return MatmulNT.apply(A, B, alpha)


class RowNormalize(torch.autograd.Function):
    """F.normalize(x, p=2, dim=-1) for a 2-D tensor (adapter.py:240,246)."""

    @staticmethod
    def forward(ctx, x):
        dev = _lib.require_cuda(x)
        x = _c(x)
        R, D = x.shape
        y = torch.empty_like(x)
        inv = torch.empty(R, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().clipgp_rownorm_forward(x.data_ptr(), R, D, y.data_ptr(), inv.data_ptr(), None, _lib.stream_ptr(dev)),
                       "clipgp_rownorm_forward")
        ctx.save_for_backward(y, inv)
        return y

    @staticmethod
    def backward(ctx, dy):
        y, inv = ctx.saved_tensors
        dy = _c(dy)
        dx = torch.empty_like(y)
        with torch.cuda.device(y.device):
            _lib.check(_lib.load().clipgp_rownorm_backward(dy.data_ptr(), y.data_ptr(), inv.data_ptr(), y.shape[0], y.shape[1], dx.data_ptr(),
                                                           _lib.stream_ptr(y.device)), "clipgp_rownorm_backward")
        return dx


def row_normalize(x):
    shp = x.shape
    return RowNormalize.apply(x.reshape(-1, shp[-1])).reshape(shp)


class SoftmaxCrossEntropy(torch.autograd.Function):
    """mean over rows of F.cross_entropy(logits [R,C], labels[r // rows_per_label])  (adapter.py:427, taskres.py:270)."""

    @staticmethod
    def forward(ctx, logits, labels, rows_per_label: int):
        dev = _lib.require_cuda(logits, labels)
        logits = _c(logits)
        labels = labels.to(torch.int64).contiguous()
        R, Cn = logits.shape
        loss = torch.zeros(1, dtype=torch.float32, device=dev)
        dlogits = torch.empty_like(logits)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().clipgp_softmax_ce(logits.data_ptr(), Cn, labels.data_ptr(), R, int(rows_per_label), Cn, None, loss.data_ptr(),
                                                     1.0 / max(R, 1), dlogits.data_ptr(), Cn, 1.0 / max(R, 1), _lib.stream_ptr(dev)),
                       "clipgp_softmax_ce")
        ctx.save_for_backward(dlogits)
        return loss[0]

    @staticmethod
    def backward(ctx, dloss):
        (dlogits,) = ctx.saved_tensors
        return dlogits * dloss, None, None


def cross_entropy(logits, labels, rows_per_label: int = 1):
    return SoftmaxCrossEntropy.apply(logits, labels, rows_per_label)


class TipCacheLogits(torch.autograd.Function):
    """tip_logits = clip_logits + alpha * exp(-(beta - beta * f keys^T)) @ one_hot(labels_tr)   (tip_adapter.py:250-260).
    Differentiable w.r.t. the cache keys (Tip-Adapter-F's nn.Linear weight) and clip_logits."""

    @staticmethod
    def forward(ctx, feats, keys, labels_tr, clip_logits, beta: float, alpha: float, num_classes: int, precision: str = "fp32"):
        dev = _lib.require_cuda(feats, keys, labels_tr)
        if precision not in ("fp32", "bf16x3", "bf16"):
            raise ValueError(f"tip_logits: unknown precision {precision!r}")
        lib = _lib.load()
        feats, keys = _c(feats), _c(keys)
        labels_tr = labels_tr.to(torch.int64).contiguous()
        clip_logits = _c(clip_logits) if clip_logits is not None else None
        B, D = feats.shape
        N_tr = keys.shape[0]
        aff = torch.empty(B, N_tr, dtype=torch.float32, device=dev)
        out = torch.empty(B, num_classes, dtype=torch.float32, device=dev)
        if B and N_tr:
            if precision == "fp32" or D % 8:
                precision = "fp32"
                _gemm(feats, D, 1, keys, 1, D, aff, B, N_tr, D, 1.0)
            else:       # tcgen05 affinity GEMM on bf16 / split-bf16 operands (the key matrix is re-cast every step: it is the trainable weight)
                from . import tc
                split = precision == "bf16x3"
                tc.gemm_store(tc.cast_bf16(feats, tc.SPLIT_A if split else tc.PLAIN), tc.cast_bf16(keys, tc.SPLIT_B if split else tc.PLAIN),
                              1.0, out=aff, split_k=True)
        with torch.cuda.device(dev):
            _lib.check(lib.clipgp_tip_forward(aff.data_ptr(), N_tr, labels_tr.data_ptr(), B, N_tr, num_classes, float(beta), float(alpha),
                                              _lib.ptr(clip_logits), num_classes, out.data_ptr(), num_classes, 1, _lib.stream_ptr(dev)),
                       "clipgp_tip_forward")
        ctx.save_for_backward(aff, feats, labels_tr)
        ctx.consts = (float(beta), float(alpha), precision)
        return out

    @staticmethod
    def backward(ctx, dout):
        e, feats, labels_tr = ctx.saved_tensors
        beta, alpha, precision = ctx.consts
        dev = e.device
        dout = _c(dout)
        B, N_tr = e.shape
        D = feats.shape[1]
        dkeys = None
        if ctx.needs_input_grad[1]:
            G = e.clone()
            with torch.cuda.device(dev):
                _lib.check(_lib.load().clipgp_tip_backward(G.data_ptr(), N_tr, labels_tr.data_ptr(), B, N_tr, dout.data_ptr(), dout.shape[1],
                                                           beta, alpha, _lib.stream_ptr(dev)), "clipgp_tip_backward")
            dkeys = torch.empty(N_tr, D, dtype=torch.float32, device=dev)
            if precision == "fp32":
                _gemm(G, 1, N_tr, feats, D, 1, dkeys, N_tr, D, B, 1.0)      # dkeys = G^T f
            else:       # K-major operands of the contraction over the batch: G^T [N_tr, B] and f^T [D, B]
                from . import tc
                split = precision == "bf16x3"
                tc.gemm_store(tc.cast_bf16_transpose(G, tc.SPLIT_A if split else tc.PLAIN),
                              tc.cast_bf16_transpose(feats, tc.SPLIT_B if split else tc.PLAIN), 1.0, out=dkeys, split_k=True)
        dclip = dout if ctx.needs_input_grad[3] else None
        return None, dkeys, None, dclip, None, None, None, None


@torch.no_grad()
def tip_affinity(feats, keys, precision: str = "fp32") -> torch.Tensor:
    """aff = feats @ keys^T [B, N_tr] (tip_adapter.py:69), materialised once so that a (beta, alpha) grid can re-use it."""
    dev = _lib.require_cuda(feats, keys)
    feats, keys = _c(feats), _c(keys)
    B, D = feats.shape
    N_tr = keys.shape[0]
    aff = torch.empty(B, N_tr, dtype=torch.float32, device=dev)
    if B and N_tr:
        if precision == "fp32" or D % 8:
            _gemm(feats, D, 1, keys, 1, D, aff, B, N_tr, D, 1.0)
        else:
            from . import tc
            split = precision == "bf16x3"
            tc.gemm_store(tc.cast_bf16(feats, tc.SPLIT_A if split else tc.PLAIN), tc.cast_bf16(keys, tc.SPLIT_B if split else tc.PLAIN), 1.0, out=aff)
    return aff


@torch.no_grad()
def tip_logits_from_affinity(aff, labels_tr, clip_logits, beta: float, alpha: float, num_classes: int) -> torch.Tensor:
    """clip_logits + alpha * exp(-(beta - beta * aff)) @ one_hot(labels_tr) from a saved affinity (left untouched)."""
    dev = _lib.require_cuda(aff, labels_tr)
    B, N_tr = aff.shape
    labels_tr = labels_tr.to(torch.int64).contiguous()
    clip_logits = _c(clip_logits) if clip_logits is not None else None
    out = torch.empty(B, num_classes, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().clipgp_tip_forward(aff.data_ptr(), aff.stride(0), labels_tr.data_ptr(), B, N_tr, num_classes, float(beta),
                                                  float(alpha), _lib.ptr(clip_logits), num_classes, out.data_ptr(), num_classes, 0,
                                                  _lib.stream_ptr(dev)), "clipgp_tip_forward")
    return out


def tip_logits(feats, keys, labels_tr, clip_logits, beta: float, alpha: float, num_classes: int, precision: str = "fp32"):
    """precision: "fp32" (FFMA GEMMs, exact comparator), "bf16x3" (tcgen05 on split-bf16 operands, fp32-grade products) or "bf16"."""
    return TipCacheLogits.apply(feats, keys, labels_tr, clip_logits, beta, alpha, num_classes, precision)


# ----------------------------------------------------------------------------------------------- GP setup
@torch.no_grad()
def median_pairwise_distance(X: torch.Tensor) -> float:
    """Lower median (torch.median) of the positive pairwise Euclidean distances between the rows of X [N,d]
    (gp_template_weigher.py:103-107 with X = unit-normalised reduced templates), by an exact three-pass radix select over the
    distance bit patterns: the N x N matrix (4 GB at the ImageNet shape) is never materialised."""
    dev = _lib.require_cuda(X)
    X = X.detach().float().contiguous()
    N, d = X.shape
    lib, st = _lib.load(), _lib.stream_ptr(dev)
    sq = torch.empty(N, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.clipgp_row_sqnorm(X.data_ptr(), N, d, sq.data_ptr(), st), "clipgp_row_sqnorm")
        prefix, prefix_bits, k = 0, 0, None
        for shift, bits in ((21, 11), (10, 11), (0, 10)):
            hist = torch.zeros(1 << bits, dtype=torch.int64, device=dev)
            pos = torch.zeros(1, dtype=torch.int64, device=dev) if k is None else None
            _lib.check(lib.clipgp_pairdist_radix_hist(X.data_ptr(), sq.data_ptr(), N, d, shift + bits, prefix, 1 if prefix_bits else 0,
                                                      shift, 1 << bits, hist.data_ptr(), _lib.ptr(pos), st), "clipgp_pairdist_radix_hist")
            h = hist.cpu()
            if k is None:
                m = int(pos.item())
                if m == 0:
                    raise RuntimeError("median_pairwise_distance: no positive pairwise distance")
                k = (m - 1) // 2                           # torch.median returns the lower middle element
            cum = torch.cumsum(h, 0)
            b = int(torch.searchsorted(cum, torch.tensor(k, dtype=torch.int64), right=True))
            k -= int(cum[b - 1]) if b > 0 else 0
            prefix = (prefix << bits) | b
            prefix_bits += bits
    return float(torch.tensor([prefix], dtype=torch.int64).to(torch.int32).view(torch.float32).item())
