"""Thin Python surface of the tcgen05 tensor-core GEMM entry points (gemm_tc.cu) and the bf16 operand casts."""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib

PLAIN, SPLIT_A, SPLIT_B = 0, 1, 2


def cast_bf16(x: torch.Tensor, mode: int = PLAIN) -> torch.Tensor:
    """fp32 [R,K] -> bf16 [R,K] (mode 0) or the K-tripled split operand [R,3K] ([hi|hi|lo] for A, [hi|lo|hi] for B)."""
    dev = _lib.require_cuda(x)
    if x.dtype != torch.float32 or x.dim() != 2 or x.stride(1) != 1:
        x = x.float().contiguous()
    R, K = x.shape
    out = torch.empty(R, K * (3 if mode else 1), dtype=torch.bfloat16, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().clipgp_cast_bf16(x.data_ptr(), R, K, x.stride(0), out.data_ptr(), out.stride(0), K, mode,
                                                _lib.stream_ptr(dev)), "clipgp_cast_bf16")
    return out


def cast_bf16_transpose(x: torch.Tensor, mode: int = PLAIN) -> torch.Tensor:
    """fp32 [R,K] -> bf16 [K, Rp] (or the split [K, 3*Rp] layouts), Rp = R padded to 8: the K-major operand of a contraction over R."""
    dev = _lib.require_cuda(x)
    if x.dtype != torch.float32 or x.dim() != 2 or x.stride(1) != 1:
        x = x.float().contiguous()
    R, K = x.shape
    Rp = (R + 7) // 8 * 8
    out = torch.zeros(K, Rp * (3 if mode else 1), dtype=torch.bfloat16, device=dev) if Rp != R else \
        torch.empty(K, Rp * (3 if mode else 1), dtype=torch.bfloat16, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().clipgp_cast_bf16_transpose(x.data_ptr(), R, K, x.stride(0), out.data_ptr(), out.stride(0), Rp, mode,
                                                          _lib.stream_ptr(dev)), "clipgp_cast_bf16_transpose")
    return out


def _pad8(v: int) -> int:
    return (v + 7) // 8 * 8


def cast_bf16_dual(x: torch.Tensor, mode: int = PLAIN, modeT: int = PLAIN):
    """One read of fp32 [R,K] -> (row-major operand [R, seg*Kp], transposed operand [K, segT*Rp]); Kp / Rp = K / R padded to 8."""
    dev = _lib.require_cuda(x)
    x = x.float().contiguous()
    R, K = x.shape
    Kp, Rp = _pad8(K), _pad8(R)
    out = torch.zeros(R, Kp * (3 if mode else 1), dtype=torch.bfloat16, device=dev)
    outT = torch.zeros(K, Rp * (3 if modeT else 1), dtype=torch.bfloat16, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().clipgp_cast_bf16_dual(x.data_ptr(), R, K, K, out.data_ptr(), out.stride(0), Kp, mode, outT.data_ptr(),
                                                     outT.stride(0), Rp, modeT, _lib.stream_ptr(dev)), "clipgp_cast_bf16_dual")
    return out, outT


def softmax_ce_operands(logits: torch.Tensor, labels: torch.Tensor, S: int, grad_scale: float, mode: int = PLAIN, fused: bool = True):
    """logits [B, S*C] (row (b,s) uses labels[b]) -> (mean CE over the B*S rows, dlogits bf16 [B, seg*SCp], dlogits^T bf16 [S*C, seg*Bp])
    with dlogits = grad_scale * (softmax - onehot): the two-phase tensor-core form of F.cross_entropy + its gradient."""
    dev = _lib.require_cuda(logits, labels)
    logits = logits.float().contiguous()
    B, SC = logits.shape
    Cn = SC // S
    stats = torch.empty(B * S, 2, dtype=torch.float32, device=dev)
    loss = torch.zeros(1, dtype=torch.float32, device=dev)
    SCp, Bp = _pad8(SC), _pad8(B)
    seg = 3 if mode else 1
    out = torch.zeros(B, seg * SCp, dtype=torch.bfloat16, device=dev)
    outT = torch.zeros(SC, seg * Bp, dtype=torch.bfloat16, device=dev)
    lib = _lib.load()
    if fused:       # one launch (the form the engine uses)
        with torch.cuda.device(dev):
            _lib.check(lib.clipgp_softmax_ce_bf16_dual(logits.data_ptr(), labels.data_ptr(), B, S, Cn, loss.data_ptr(), 1.0 / (B * S),
                                                       float(grad_scale), out.data_ptr(), out.stride(0), SCp, mode, outT.data_ptr(),
                                                       outT.stride(0), Bp, mode, _lib.stream_ptr(dev)), "clipgp_softmax_ce_bf16_dual")
        return loss, out, outT
    with torch.cuda.device(dev):
        st = _lib.stream_ptr(dev)
        _lib.check(lib.clipgp_softmax_ce_stats(logits.data_ptr(), Cn, labels.data_ptr(), B * S, S, Cn, stats.data_ptr(), loss.data_ptr(),
                                               1.0 / (B * S), st), "clipgp_softmax_ce_stats")
        _lib.check(lib.clipgp_softmax_grad_bf16_dual(logits.data_ptr(), stats.data_ptr(), labels.data_ptr(), B, S, Cn, float(grad_scale),
                                                     out.data_ptr(), out.stride(0), SCp, mode, outT.data_ptr(), outT.stride(0), Bp, mode,
                                                     st), "clipgp_softmax_grad_bf16_dual")
    return loss, out, outT


def gemm_store(A: torch.Tensor, B: torch.Tensor, alpha: float = 1.0, out: Optional[torch.Tensor] = None,
               split_k: bool = False) -> torch.Tensor:
    """C = alpha * A @ B^T for bf16 A [M,Ka], B [N,K] (K % Ka == 0: A wraps along K).  split_k: the training-step entry that may
    split K over work items (skinny outputs, tail wave) and accumulate with vector reductions (non-deterministic last bits)."""
    dev = _lib.require_cuda(A, B)
    assert A.dtype == torch.bfloat16 and B.dtype == torch.bfloat16 and A.is_contiguous() and B.is_contiguous()
    M, Ka = A.shape
    N, K = B.shape
    C = out if out is not None else torch.empty(M, N, dtype=torch.float32, device=dev)
    lib = _lib.load()
    fn = lib.clipgp_tc_gemm_store_splitk if split_k else lib.clipgp_tc_gemm_store
    with torch.cuda.device(dev):
        _lib.check(fn(A.data_ptr(), M, Ka, B.data_ptr(), N, K, float(alpha), C.data_ptr(), C.stride(0), _lib.stream_ptr(dev)),
                   "clipgp_tc_gemm_store")
    return C


def logits_calibration(A: torch.Tensor, B: torch.Tensor, alpha: float, labels: torch.Tensor, n_bins: int = 10,
                       want_conf: bool = True, want_logits: bool = False):
    """Fused logits + softmax confidence + hit flag + ECE histogram.  Returns (conf, correct, hist[4,n_bins], logits|None)
    with the same hist layout as metrics.calibration_pass (rows: count, conf_fx, correct; hist[3,0] = top-1 count)."""
    from .metrics import _boundaries
    dev = _lib.require_cuda(A, B, labels)
    assert A.dtype == torch.bfloat16 and B.dtype == torch.bfloat16 and A.is_contiguous() and B.is_contiguous()
    M, Ka = A.shape
    N, K = B.shape
    labels = labels.to(torch.int64).contiguous()
    conf = torch.empty(M, dtype=torch.float32, device=dev) if want_conf else None
    correct = torch.empty(M, dtype=torch.uint8, device=dev) if want_conf else None
    hist = torch.zeros(4, max(n_bins, 1), dtype=torch.int64, device=dev)
    logits = torch.empty(M, N, dtype=torch.float32, device=dev) if want_logits else None
    b = _boundaries(n_bins, dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().clipgp_tc_logits_calibration(
            A.data_ptr(), M, Ka, B.data_ptr(), N, K, float(alpha), labels.data_ptr(), _lib.ptr(conf), None, _lib.ptr(correct),
            b.data_ptr(), n_bins, hist[0].data_ptr(), hist[1].data_ptr(), hist[2].data_ptr(), hist[3].data_ptr(),
            _lib.ptr(logits), N, _lib.stream_ptr(dev)), "clipgp_tc_logits_calibration")
    return conf, correct, hist, logits


def proj_logits_calibration(A: torch.Tensor, B: torch.Tensor, norm_cols: int, alpha: float, labels: torch.Tensor, n_bins: int = 10):
    """One GEMM for projection + normalisation + logits + calibration: A = raw features [M,Ka] bf16, B = [W ; Q] [norm_cols + C, K]
    bf16 (clipgp_tc_proj_logits_calibration).  Returns (conf, correct, hist[4,n_bins]) like logits_calibration."""
    from .metrics import _boundaries
    dev = _lib.require_cuda(A, B, labels)
    assert A.dtype == torch.bfloat16 and B.dtype == torch.bfloat16 and A.is_contiguous() and B.is_contiguous()
    M, Ka = A.shape
    N, K = B.shape
    labels = labels.to(torch.int64).contiguous()
    conf = torch.empty(M, dtype=torch.float32, device=dev)
    correct = torch.empty(M, dtype=torch.uint8, device=dev)
    hist = torch.zeros(4, max(n_bins, 1), dtype=torch.int64, device=dev)
    b = _boundaries(n_bins, dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().clipgp_tc_proj_logits_calibration(
            A.data_ptr(), M, Ka, B.data_ptr(), N, K, int(norm_cols), float(alpha), labels.data_ptr(), conf.data_ptr(), None,
            correct.data_ptr(), b.data_ptr(), n_bins, hist[0].data_ptr(), hist[1].data_ptr(), hist[2].data_ptr(), hist[3].data_ptr(),
            _lib.stream_ptr(dev)), "clipgp_tc_proj_logits_calibration")
    return conf, correct, hist


# ---------------------------------------------------------------------------------------------------- TF32 (fp32 operands in place)
def gemm_tf32(A: torch.Tensor, B: torch.Tensor, alpha: float = 1.0, a_t: bool = False, b_t: bool = False, out: Optional[torch.Tensor] = None,
              split_k: bool = False) -> torch.Tensor:
    """C = alpha * op(A) @ op(B)^T on the tensor cores in TF32, fp32 operands read in place (no cast kernels).

    a_t False: A is [M, K];  a_t True: A is [K, M] (its transpose is the operand: an "MN-major" read, no transposed copy).
    b_t False: B is [N, K];  b_t True: B is [K, N].  Examples (trainers/adapter.py:419-428 and adjoints):
        logits = gemm_tf32(f_hat, P_hat, scale);  d f_hat = gemm_tf32(dlogits, P_hat, scale, b_t=True)
        d P_hat = gemm_tf32(dlogits, f_hat, scale, a_t=True, b_t=True)."""
    dev = _lib.require_cuda(A, B)
    A, B = A.float().contiguous(), B.float().contiguous()
    M, K = (A.shape[1], A.shape[0]) if a_t else A.shape
    N, Kb = (B.shape[1], B.shape[0]) if b_t else B.shape
    if K != Kb:
        raise ValueError(f"gemm_tf32: contraction sizes differ ({K} vs {Kb})")
    C_ = out if out is not None else torch.empty(M, N, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().clipgp_tc_gemm_tf32(A.data_ptr(), int(a_t), M, B.data_ptr(), int(b_t), N, K, float(alpha), C_.data_ptr(),
                                                   C_.stride(0), int(split_k), _lib.stream_ptr(dev)), "clipgp_tc_gemm_tf32")
    return C_


def logits_calibration_tf32(A: torch.Tensor, B: torch.Tensor, alpha: float, labels: torch.Tensor, n_bins: int = 10, norm_cols: int = 0,
                            want_logits: bool = False):
    """logits_calibration / proj_logits_calibration on fp32 operands (TF32, no casts): A [M,K] features, B [N,K] = class prototypes
    (norm_cols = 0) or [W ; Q] (norm_cols = D: projection columns first).  Returns (conf, correct, hist[4,n_bins], logits|None)."""
    from .metrics import _boundaries
    dev = _lib.require_cuda(A, B, labels)
    assert A.dtype == torch.float32 and B.dtype == torch.float32 and A.is_contiguous() and B.is_contiguous()
    M, Ka = A.shape
    N, K = B.shape
    labels = labels.to(torch.int64).contiguous()
    conf = torch.empty(M, dtype=torch.float32, device=dev)
    correct = torch.empty(M, dtype=torch.uint8, device=dev)
    hist = torch.zeros(4, max(n_bins, 1), dtype=torch.int64, device=dev)
    logits = torch.empty(M, N, dtype=torch.float32, device=dev) if want_logits else None
    b = _boundaries(n_bins, dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().clipgp_tc_logits_calibration_tf32(
            A.data_ptr(), M, Ka, B.data_ptr(), N, K, int(norm_cols), float(alpha), labels.data_ptr(), conf.data_ptr(), None,
            correct.data_ptr(), b.data_ptr(), n_bins, hist[0].data_ptr(), hist[1].data_ptr(), hist[2].data_ptr(), hist[3].data_ptr(),
            _lib.ptr(logits), N, _lib.stream_ptr(dev)), "clipgp_tc_logits_calibration_tf32")
    return conf, correct, hist, logits
