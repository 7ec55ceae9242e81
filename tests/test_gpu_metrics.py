"""GPU: the clipgp calibration kernels against the reference's own outputs (golden) and the oracle."""
import os

import numpy as np
import pytest
import torch

from clip_gp_b200 import metrics as gm
from oracle import metrics as om

pytestmark = pytest.mark.gpu
CASES = ["rand_600x50", "peaked_257x12", "flat_123x7", "tiny_5x4", "binary_40x2", "ties_64x10", "saturated_30x3"]


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "metrics_golden.npz"))


@pytest.mark.parametrize("name", CASES)
def test_kernels_match_reference_golden(gold, name):
    lg = torch.from_numpy(gold[f"{name}/logits"]).cuda()
    lb = torch.from_numpy(gold[f"{name}/labels"]).cuda()
    conf, correct, hist = gm.calibration_pass(lg, lb, 10)
    # confidences: same formula (1/sum exp(x - max)); summation order differs -> a few ulp
    np.testing.assert_allclose(conf.cpu().numpy(), gold[f"{name}/conf"], rtol=4e-7, atol=0)
    res = gm.evaluate_calibration(lg, lb, 10)
    gap = np.abs(gold[f"{name}/conf"][:, None] - gold["boundaries"][None, :]).min()
    # bit-exact integer outputs (the ulp-level conf difference can only matter within ~1e-7 of a boundary)
    assert res["calibration"]["bin_count"] == gold[f"{name}/ece_bin_count"].tolist(), f"min |conf-boundary| = {gap:.3e}"
    assert res["adaptive_calibration"]["bin_count"] == gold[f"{name}/aece_bin_count"].tolist()
    if name != "ties_64x10":   # exact logit ties: reference topk/argmax tie order is implementation-defined
        assert res["top1_acc"] == pytest.approx(float(gold[f"{name}/acc"]), rel=1e-6)
    assert res["ece"] == pytest.approx(float(gold[f"{name}/ece"]), rel=1e-3, abs=1e-4)
    np.testing.assert_allclose(res["calibration"]["bin_conf"], gold[f"{name}/ece_bin_conf"], rtol=1e-5)
    if name != "ties_64x10":
        np.testing.assert_allclose(res["calibration"]["bin_acc"], gold[f"{name}/ece_bin_acc"], rtol=1e-5, atol=1e-7)
        assert res["aece"] == pytest.approx(float(gold[f"{name}/aece"]), rel=1e-3, abs=1e-4)
        np.testing.assert_allclose(res["adaptive_calibration"]["bin_acc"], gold[f"{name}/aece_bin_acc"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(res["adaptive_calibration"]["bin_conf"], gold[f"{name}/aece_bin_conf"], rtol=1e-5)
    # reference-named entry points
    assert gm.compute_ece(lg, lb) == pytest.approx(res["ece"])
    assert gm.compute_aece(lg, lb) == pytest.approx(res["aece"])
    assert gm.compute_accuracy(lg, lb)[0] == pytest.approx(res["top1_acc"])


@pytest.mark.parametrize("N,C", [(1, 3), (33, 1000), (4097, 37), (1000, 1001), (50000, 1000)])
def test_against_oracle_on_shared_logits(N, C):
    g = torch.Generator().manual_seed(N * 7 + C)
    lg = 4.0 * torch.randn(N, C, generator=g)
    lb = torch.randint(0, C, (N,), generator=g)
    lb[: N // 2] = lg[: N // 2].argmax(1)
    res = gm.evaluate_calibration(lg.cuda(), lb.cuda(), 10)
    conf, pred, correct = om.confidence(lg, lb)
    assert res["top1_count"] == int(correct.sum())                                   # bit-exact top-1 count
    e, b = om.compute_ece_with_bins(lg, lb)
    gap = float((conf[:, None] - torch.linspace(0, 1, 11)[None]).abs().min())
    assert res["calibration"]["bin_count"] == b["bin_count"], f"min |conf-boundary| = {gap:.3e}"
    assert res["ece"] == pytest.approx(e, rel=1e-3, abs=1e-4)
    ea, ba = om.compute_aece_with_bins(lg, lb)
    assert res["adaptive_calibration"]["bin_count"] == ba["bin_count"]
    assert res["aece"] == pytest.approx(ea, rel=1e-3, abs=1e-4)
    np.testing.assert_allclose(res["adaptive_calibration"]["bin_conf"], ba["bin_conf"], rtol=1e-5)


def test_aece_is_the_sorted_partition():
    """Property at full size: per-bin sums from the radix select == sums over the torch-sorted ranges."""
    g = torch.Generator().manual_seed(5)
    N = 50000
    conf = torch.rand(N, generator=g).cuda()
    conf[::7] = conf[0]                                   # long runs of equal keys across bin edges
    correct = (torch.rand(N, generator=g) < 0.6).to(torch.uint8).cuda()
    edges, out = gm.aece_pass(conf, correct, 10)
    key = conf.double() * 2 + correct.double() * 1e-12    # sort by (conf, correct), the kernel's tie order
    order = torch.argsort(key)
    sc, sa = conf[order].double(), correct[order].double()
    o = out.cpu()
    for i in range(10):
        lo, hi = int(edges[i]), int(edges[i + 1])
        assert int(o[2, i]) == hi - lo
        assert int(o[1, i]) == int(sa[lo:hi].sum().item())
        assert float(o[0, i]) / gm.FX_SCALE == pytest.approx(float(sc[lo:hi].sum().item()), rel=1e-9)


@pytest.mark.parametrize("N,n_bins", [(1000003, 25), (4097, 10), (37, 10)])
def test_aece_multi_cta_multi_sweep(N, n_bins):
    """Grid-wide radix select: many CTAs, more interior edges than one sweep resolves, saturated confidences (conf == 1 runs)."""
    g = torch.Generator().manual_seed(N)
    conf = torch.rand(N, generator=g).pow(0.05)           # piled up against 1.0 like real softmax confidences
    conf[torch.rand(N, generator=g) < 0.3] = 1.0
    conf = conf.cuda()
    correct = (torch.rand(N, generator=g) < 0.7).to(torch.uint8).cuda()
    edges, out = gm.aece_pass(conf, correct, n_bins)
    order = torch.argsort(conf.double() * 4 + correct.double() * 1e-9)
    sc, sa = conf[order].double(), correct[order].double()
    o = out.cpu()
    for i in range(edges.numel() - 1):
        lo, hi = int(edges[i]), int(edges[i + 1])
        assert int(o[2, i]) == hi - lo
        assert int(o[1, i]) == int(sa[lo:hi].sum().item())
        assert float(o[0, i]) / gm.FX_SCALE == pytest.approx(float(sc[lo:hi].sum().item()), rel=1e-9, abs=1e-9)


def test_counters_are_shard_invariant():
    g = torch.Generator().manual_seed(9)
    lg = 3.0 * torch.randn(3001, 100, generator=g).cuda()
    lb = torch.randint(0, 100, (3001,), generator=g).cuda()
    _, _, full = gm.calibration_pass(lg, lb, 10, want_conf=False)
    acc = None
    for lo, hi in [(0, 700), (700, 701), (701, 3001)]:
        _, _, h = gm.calibration_pass(lg[lo:hi], lb[lo:hi], 10, want_conf=False)
        acc = h if acc is None else acc + h
    assert torch.equal(acc, full)                         # counts AND fixed-point conf sums are exactly additive


def test_strided_logits_and_empty():
    g = torch.Generator().manual_seed(2)
    big = torch.randn(64, 50, generator=g).cuda()
    view = big[:, :30]
    lb = torch.randint(0, 30, (64,), generator=g).cuda()
    a = gm.evaluate_calibration(view, lb)
    b = gm.evaluate_calibration(view.contiguous(), lb)
    assert a["calibration"] == b["calibration"] and a["top1_count"] == b["top1_count"]
    e = gm.evaluate_calibration(torch.zeros(0, 5).cuda(), torch.zeros(0, dtype=torch.long).cuda())
    assert e["n"] == 0 and e["ece"] == 0.0
    assert gm.compute_accuracy(torch.zeros(0, 5).cuda(), torch.zeros(0, dtype=torch.long).cuda()) == [0.0]
