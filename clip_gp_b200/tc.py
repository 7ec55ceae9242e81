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


def gemm_store(A: torch.Tensor, B: torch.Tensor, alpha: float = 1.0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """C = alpha * A @ B^T for bf16 A [M,Ka], B [N,K] (K % Ka == 0: A wraps along K)."""
    dev = _lib.require_cuda(A, B)
    assert A.dtype == torch.bfloat16 and B.dtype == torch.bfloat16 and A.is_contiguous() and B.is_contiguous()
    M, Ka = A.shape
    N, K = B.shape
    C = out if out is not None else torch.empty(M, N, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().clipgp_tc_gemm_store(A.data_ptr(), M, Ka, B.data_ptr(), N, K, float(alpha), C.data_ptr(),
                                                    C.stride(0), _lib.stream_ptr(dev)), "clipgp_tc_gemm_store")
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
