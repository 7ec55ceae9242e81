"""CPU oracle: accuracy / ECE / AECE (TEST INFRASTRUCTURE, see oracle/__init__.py).

Restates /root/reference/utils/metrics.py:9-229.  PINNED: checked against outputs of the
reference module itself (tests/golden/metrics_golden.npz, made by tests/golden/make_golden.py).

One pass computes (conf, pred, correct); the four reference entry points are thin views on it.
Float summaries reproduce the reference's reduction order (masked select -> fp32 mean) so that
they agree to the last bit with the golden values on the same torch build.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch


def confidence(logits: torch.Tensor, labels: torch.Tensor):
    """metrics.py:71-73 — softmax, max, hit flag."""
    probs = torch.softmax(logits, dim=-1)
    conf, pred = probs.max(dim=-1)
    return conf, pred, pred.eq(labels)


def compute_accuracy(logits: torch.Tensor, labels: torch.Tensor, topk: Tuple[int, ...] = (1,)) -> List[float]:
    """metrics.py:9-36 — percent of rows whose label is among the k largest logits."""
    n = labels.size(0)
    if n == 0:
        return [0.0] * len(topk)
    idx = logits.topk(max(topk), dim=1, largest=True, sorted=True).indices       # [N,maxk]
    hit = idx.eq(labels.view(-1, 1))
    out = []
    for k in topk:
        cnt = hit[:, :k].reshape(-1).float().sum(0, keepdim=True)
        out.append(cnt.mul_(100.0 / n).item())
    return out


def top1_count(logits: torch.Tensor, labels: torch.Tensor) -> int:
    return int(logits.argmax(dim=1).eq(labels).sum().item())


def ece_bins(conf: torch.Tensor, correct: torch.Tensor, n_bins: int = 10):
    """Equal-width binning of metrics.py:75-82 / :154-175 on precomputed (conf, correct).

    Returns (ece_percent, bin_acc, bin_conf, bin_count[int]).  Boundaries are
    torch.linspace(0,1,n_bins+1) in fp32; membership is  b_i < conf <= b_{i+1}.
    """
    acc = correct.float()
    b = torch.linspace(0, 1, n_bins + 1, device=conf.device)
    ece = torch.zeros(1, device=conf.device)
    bin_acc, bin_conf, bin_cnt = [], [], []
    for i in range(n_bins):
        in_bin = (conf > b[i]) * (conf <= b[i + 1])
        count = int(in_bin.sum().item())
        if count > 0:
            a = acc[in_bin].mean().item()
            c = conf[in_bin].mean().item()
            ece += abs(c - a) * (float(count) / float(conf.numel()))
            bin_acc.append(float(a)); bin_conf.append(float(c)); bin_cnt.append(count)
        else:
            bin_acc.append(0.0); bin_conf.append((i + 0.5) / n_bins); bin_cnt.append(0)
    return float(ece.item() * 100.0), bin_acc, bin_conf, bin_cnt


def compute_ece(logits: torch.Tensor, labels: torch.Tensor, n_bins: int = 10) -> float:
    """metrics.py:59-83.  (Tensor-valued |conf-acc|*prop accumulation, as the reference.)"""
    conf, _, correct = confidence(logits, labels)
    acc = correct.float()
    b = torch.linspace(0, 1, n_bins + 1, device=logits.device)
    ece = torch.zeros(1, device=logits.device)
    for i in range(n_bins):
        in_bin = (conf > b[i]) * (conf <= b[i + 1])
        prop = in_bin.float().mean()
        if prop.item() > 0:
            ece += torch.abs(conf[in_bin].mean() - acc[in_bin].mean()) * prop
    return float(ece.item() * 100)


def compute_ece_with_bins(logits, labels, n_bins: int = 10) -> Tuple[float, Dict[str, list]]:
    """metrics.py:138-176."""
    conf, _, correct = confidence(logits, labels)
    e, a, c, n = ece_bins(conf, correct, n_bins)
    return e, {"bin_acc": a, "bin_conf": c, "bin_count": n}


def aece_bins(conf: torch.Tensor, correct: torch.Tensor, n_bins: int = 10):
    """Equal-count binning of metrics.py:107-133 / :197-229 on precomputed (conf, correct)."""
    N = conf.numel()
    if N == 0:
        return 0.0, [], [], []
    n_bins = max(1, min(int(n_bins), int(N)))
    sconf, order = torch.sort(conf)
    sacc = correct.float()[order]
    edges = torch.linspace(0, N, n_bins + 1, device=conf.device).round().long()
    edges[0] = 0
    edges[-1] = N
    aece = torch.zeros(1, device=conf.device)
    bin_acc, bin_conf, bin_cnt = [], [], []
    for i in range(n_bins):
        lo, hi = int(edges[i].item()), int(edges[i + 1].item())
        if hi <= lo:
            bin_acc.append(0.0); bin_conf.append((i + 0.5) / n_bins); bin_cnt.append(0)
            continue
        c = sconf[lo:hi].mean().item()
        a = sacc[lo:hi].mean().item()
        aece += abs(c - a) * ((hi - lo) / float(N))
        bin_acc.append(float(a)); bin_conf.append(float(c)); bin_cnt.append(hi - lo)
    return float(aece.item() * 100.0), bin_acc, bin_conf, bin_cnt


def compute_aece(logits, labels, n_bins: int = 10) -> float:
    """metrics.py:86-135 (tensor-valued accumulation)."""
    if logits.numel() == 0:
        return 0.0
    conf, _, correct = confidence(logits, labels)
    N = conf.numel()
    n_bins = max(1, min(int(n_bins), int(N)))
    sconf, order = torch.sort(conf)
    sacc = correct.float()[order]
    edges = torch.linspace(0, N, n_bins + 1, device=logits.device).round().long()
    edges[0] = 0
    edges[-1] = N
    aece = torch.zeros(1, device=logits.device)
    for i in range(n_bins):
        lo, hi = int(edges[i].item()), int(edges[i + 1].item())
        if hi <= lo:
            continue
        aece += torch.abs(sconf[lo:hi].mean() - sacc[lo:hi].mean()) * ((hi - lo) / float(N))
    return float(aece.item() * 100)


def compute_aece_with_bins(logits, labels, n_bins: int = 10) -> Tuple[float, Dict[str, list]]:
    """metrics.py:179-229."""
    if logits.numel() == 0:
        return 0.0, {"bin_acc": [], "bin_conf": [], "bin_count": []}
    conf, _, correct = confidence(logits, labels)
    e, a, c, n = aece_bins(conf, correct, n_bins)
    return e, {"bin_acc": a, "bin_conf": c, "bin_count": n}
