"""Accuracy / ECE / AECE on the GPU — same names, arguments and return conventions as the reference's
``utils/metrics.py`` (:9-36, :59-83, :86-135, :138-176, :179-229), computed by the clipgp CUDA kernels.

One pass over the logits yields (conf, correct) and the equal-width histogram; AECE runs an exact
sort-free rank partition on the confidences.  Percent floats, per-bin dicts ``{bin_acc, bin_conf,
bin_count}`` exactly as the reference writes them into ``metrics.json`` (utils/trainer.py:625-637).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch

from . import _lib

FX_SCALE = float(1 << 40)   # bin_conf_fx is a sum of confidences in 2^-40 fixed point (exact, order independent)


@dataclass
class CalibrationCounters:
    """Integer sufficient statistics of one evaluation shard; additive across shards / ranks."""
    n: int
    top1: int
    bin_count: torch.Tensor      # [n_bins] int64 (CPU)
    bin_correct: torch.Tensor    # [n_bins] int64
    bin_conf_fx: torch.Tensor    # [n_bins] int64 (uint64 payload; sums stay < 2^63 for N < 2^23)

    def merged(self, other: "CalibrationCounters") -> "CalibrationCounters":
        return CalibrationCounters(self.n + other.n, self.top1 + other.top1, self.bin_count + other.bin_count,
                                   self.bin_correct + other.bin_correct, self.bin_conf_fx + other.bin_conf_fx)


_BOUNDARY_CACHE: Dict[tuple, torch.Tensor] = {}


def _boundaries(n_bins: int, device) -> torch.Tensor:
    """torch.linspace(0, 1, n_bins + 1) in fp32 (metrics.py:70) on the device; cached, so that a captured CUDA graph (engine.eval_graph)
    sees no host-to-device copy."""
    key = (int(n_bins), str(device))
    if key not in _BOUNDARY_CACHE:
        _BOUNDARY_CACHE[key] = torch.linspace(0, 1, n_bins + 1, dtype=torch.float32).to(device)
    return _BOUNDARY_CACHE[key]


def _prep(logits: torch.Tensor, labels: torch.Tensor):
    dev = _lib.require_cuda(logits, labels)
    if logits.dim() != 2:
        raise ValueError(f"logits must be [N, C], got {tuple(logits.shape)}")
    if logits.dtype != torch.float32:
        logits = logits.float()
    if logits.stride(1) != 1:
        logits = logits.contiguous()
    labels = labels.to(dtype=torch.int64).contiguous()
    if labels.numel() != logits.shape[0]:
        raise ValueError("labels must have one entry per logits row")
    return dev, logits, labels


def calibration_pass(logits: torch.Tensor, labels: torch.Tensor, n_bins: int = 10, want_conf: bool = True):
    """Run the fused kernel.  Returns device tensors (conf, correct, hist[4, n_bins] int64) where
    hist rows are count, conf_fx, correct, and hist[3,0] is the top-1 count.  No host sync."""
    dev, logits, labels = _prep(logits, labels)
    lib = _lib.load()
    N, Cc = logits.shape
    conf = torch.empty(N, dtype=torch.float32, device=dev) if want_conf else None
    correct = torch.empty(N, dtype=torch.uint8, device=dev) if want_conf else None
    hist = torch.zeros(4, max(n_bins, 1), dtype=torch.int64, device=dev)
    b = _boundaries(n_bins, dev)
    with torch.cuda.device(dev):
        _lib.check(lib.clipgp_calibration_from_logits(
            _lib.ptr(logits), logits.stride(0) if N > 0 else Cc, _lib.ptr(labels), N, Cc, _lib.ptr(conf), None,
            _lib.ptr(correct), _lib.ptr(b), n_bins, hist[0].data_ptr(), hist[1].data_ptr(), hist[2].data_ptr(),
            hist[3].data_ptr(), _lib.stream_ptr(dev)), "clipgp_calibration_from_logits")
    return conf, correct, hist


def counters_from_hist(hist: torch.Tensor, n: int) -> CalibrationCounters:
    h = hist.cpu()
    return CalibrationCounters(n=n, top1=int(h[3, 0]), bin_count=h[0].clone(), bin_correct=h[2].clone(), bin_conf_fx=h[1].clone())


def ece_from_counters(cnt: CalibrationCounters) -> Tuple[float, Dict[str, list]]:
    """metrics.py:159-176 on integer counters."""
    n_bins = cnt.bin_count.numel()
    ece = 0.0
    bin_acc: List[float] = []
    bin_conf: List[float] = []
    bin_cnt: List[int] = []
    for i in range(n_bins):
        c = int(cnt.bin_count[i])
        if c > 0:
            a = int(cnt.bin_correct[i]) / c
            cf = (int(cnt.bin_conf_fx[i]) / FX_SCALE) / c
            ece += abs(cf - a) * (c / float(cnt.n))
            bin_acc.append(a); bin_conf.append(cf); bin_cnt.append(c)
        else:
            bin_acc.append(0.0); bin_conf.append((i + 0.5) / n_bins); bin_cnt.append(0)
    return ece * 100.0, {"bin_acc": bin_acc, "bin_conf": bin_conf, "bin_count": bin_cnt}


def aece_edges(n: int, n_bins: int) -> torch.Tensor:
    """metrics.py:110-120: rank edges linspace(0, N, n_bins+1).round().long() (fp32, half-to-even)."""
    nb = max(1, min(int(n_bins), int(n)))
    edges = torch.linspace(0, n, nb + 1).round().long()
    edges[0] = 0
    edges[-1] = n
    return edges


_EDGE_CACHE: Dict[tuple, tuple] = {}


def aece_pass(conf: torch.Tensor, correct: torch.Tensor, n_bins: int = 10):
    """Exact equal-count binning of (conf, correct) on the device.  Returns (edges CPU, out[3, nb] int64 device)
    with rows conf_fx, correct, count.  No host sync."""
    dev = _lib.require_cuda(conf, correct)
    lib = _lib.load()
    n = conf.numel()
    key = (n, n_bins, str(dev))
    if key not in _EDGE_CACHE:                              # the rank edges depend on (N, n_bins) only: one H2D copy per shape
        if len(_EDGE_CACHE) > 64:
            _EDGE_CACHE.clear()
        e = aece_edges(n, n_bins)
        _EDGE_CACHE[key] = (e, e.to(dev))
    edges, edges_d = _EDGE_CACHE[key]
    nb = edges.numel() - 1
    out = torch.zeros(3, nb, dtype=torch.int64, device=dev)
    ws = torch.empty(int(lib.clipgp_aece_workspace_bytes(nb)), dtype=torch.uint8, device=dev)     # torch's caching allocator: no driver call
    with torch.cuda.device(dev):
        _lib.check(lib.clipgp_aece_bins(_lib.ptr(conf), _lib.ptr(correct), n, _lib.ptr(edges_d), nb, out[0].data_ptr(),
                                        out[1].data_ptr(), out[2].data_ptr(), ws.data_ptr(), ws.numel(), _lib.stream_ptr(dev)),
                   "clipgp_aece_bins")
    return edges, out


def aece_from_bins(out: torch.Tensor, n: int, n_bins_requested: int) -> Tuple[float, Dict[str, list]]:
    """metrics.py:207-229 on the per-bin integer sums."""
    o = out.cpu()
    nb = o.shape[1]
    aece = 0.0
    bin_acc: List[float] = []
    bin_conf: List[float] = []
    bin_cnt: List[int] = []
    for i in range(nb):
        c = int(o[2, i])
        if c <= 0:
            bin_acc.append(0.0); bin_conf.append((i + 0.5) / float(nb)); bin_cnt.append(0)
            continue
        cf = (int(o[0, i]) / FX_SCALE) / c
        a = int(o[1, i]) / c
        aece += abs(cf - a) * (c / float(n))
        bin_acc.append(a); bin_conf.append(cf); bin_cnt.append(c)
    return aece * 100.0, {"bin_acc": bin_acc, "bin_conf": bin_conf, "bin_count": bin_cnt}


# ------------------------------------------------------------------------------------------------
# reference-named entry points
# ------------------------------------------------------------------------------------------------
def compute_accuracy(logits: torch.Tensor, labels: torch.Tensor, topk: Tuple[int, ...] = (1,)) -> List[float]:
    """utils/metrics.py:9-36: top-k accuracies in percent.  Top-1 (the only one the reference trainers use) comes from the fused
    calibration pass (arg-max with first-index tie break, as `topk`'s largest-first order); k > 1 counts the rows whose label logit
    is beaten by fewer than k others (ties resolved in favour of the label, the documented-undefined case of torch.topk)."""
    n = labels.size(0)
    if n == 0:
        return [0.0] * len(topk)
    out = []
    for k in topk:
        if k == 1:
            _, _, hist = calibration_pass(logits, labels, n_bins=1, want_conf=False)
            out.append(int(hist[3, 0].item()) * (100.0 / n))
        else:
            lab = labels.to(logits.device, torch.int64).view(-1, 1)
            rank = (logits > logits.gather(1, lab)).sum(dim=1)               # streaming compare + row sum: no sort
            out.append(float((rank < int(k)).sum().item()) * (100.0 / n))
    return out


def compute_ece_with_bins(logits, labels, n_bins: int = 10) -> Tuple[float, Dict[str, list]]:
    """utils/metrics.py:138-176."""
    _, _, hist = calibration_pass(logits, labels, n_bins=n_bins, want_conf=False)
    return ece_from_counters(counters_from_hist(hist, int(labels.numel())))


def compute_ece(logits, labels, n_bins: int = 10) -> float:
    """utils/metrics.py:59-83."""
    if labels.numel() == 0:
        return 0.0
    return compute_ece_with_bins(logits, labels, n_bins)[0]


def compute_aece_with_bins(logits, labels, n_bins: int = 10) -> Tuple[float, Dict[str, list]]:
    """utils/metrics.py:179-229."""
    if logits.numel() == 0:
        return 0.0, {"bin_acc": [], "bin_conf": [], "bin_count": []}
    conf, correct, _ = calibration_pass(logits, labels, n_bins=1, want_conf=True)
    _, out = aece_pass(conf, correct, n_bins)
    return aece_from_bins(out, conf.numel(), n_bins)


def compute_aece(logits, labels, n_bins: int = 10) -> float:
    """utils/metrics.py:86-135."""
    return compute_aece_with_bins(logits, labels, n_bins)[0]


def evaluate_calibration(logits, labels, n_bins: int = 10) -> Dict[str, object]:
    """Everything ``BaseTrainer.test`` reports (utils/trainer.py:474-557) from ONE pass over the logits
    (the reference makes four softmax passes): top-1 %, ECE, AECE and both per-bin tables."""
    n = int(labels.numel())
    if n == 0:
        return {"top1_acc": 0.0, "ece": 0.0, "aece": 0.0, "calibration": {"bin_acc": [], "bin_conf": [], "bin_count": []},
                "adaptive_calibration": {"bin_acc": [], "bin_conf": [], "bin_count": []}, "n": 0, "top1_count": 0}
    conf, correct, hist = calibration_pass(logits, labels, n_bins=n_bins, want_conf=True)
    _, out = aece_pass(conf, correct, n_bins)
    cnt = counters_from_hist(hist, n)
    ece, calib = ece_from_counters(cnt)
    aece, acalib = aece_from_bins(out, n, n_bins)
    return {"top1_acc": cnt.top1 * (100.0 / n), "ece": ece, "aece": aece, "calibration": calib,
            "adaptive_calibration": acalib, "n": n, "top1_count": cnt.top1, "counters": cnt}
