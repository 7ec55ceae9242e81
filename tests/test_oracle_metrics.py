"""CPU: oracle/metrics.py against the golden outputs of the REFERENCE's utils/metrics.py."""
import os

import numpy as np
import pytest
import torch

from oracle import metrics as om

CASES = ["rand_600x50", "peaked_257x12", "flat_123x7", "tiny_5x4", "binary_40x2", "ties_64x10", "saturated_30x3"]


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "metrics_golden.npz"))


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_metrics(gold, name):
    lg = torch.from_numpy(gold[f"{name}/logits"])
    lb = torch.from_numpy(gold[f"{name}/labels"])
    assert om.compute_accuracy(lg, lb)[0] == pytest.approx(float(gold[f"{name}/acc"]), abs=0)
    assert om.compute_ece(lg, lb) == pytest.approx(float(gold[f"{name}/ece"]), rel=1e-6, abs=1e-9)
    assert om.compute_aece(lg, lb) == pytest.approx(float(gold[f"{name}/aece"]), rel=1e-6, abs=1e-9)
    e, b = om.compute_ece_with_bins(lg, lb)
    assert e == pytest.approx(float(gold[f"{name}/ece_b"]), rel=1e-6, abs=1e-9)
    assert b["bin_count"] == gold[f"{name}/ece_bin_count"].tolist()          # bit-exact integer counts
    np.testing.assert_allclose(b["bin_acc"], gold[f"{name}/ece_bin_acc"], rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(b["bin_conf"], gold[f"{name}/ece_bin_conf"], rtol=1e-6, atol=1e-9)
    e, b = om.compute_aece_with_bins(lg, lb)
    assert e == pytest.approx(float(gold[f"{name}/aece_b"]), rel=1e-6, abs=1e-9)
    assert b["bin_count"] == gold[f"{name}/aece_bin_count"].tolist()
    np.testing.assert_allclose(b["bin_acc"], gold[f"{name}/aece_bin_acc"], rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(b["bin_conf"], gold[f"{name}/aece_bin_conf"], rtol=1e-6, atol=1e-9)
    conf, pred, _ = om.confidence(lg, lb)
    np.testing.assert_array_equal(conf.numpy(), gold[f"{name}/conf"])
    np.testing.assert_array_equal(pred.numpy(), gold[f"{name}/pred"])


def test_boundaries_are_the_fp32_linspace(gold):
    b = torch.linspace(0, 1, 11)
    np.testing.assert_array_equal(b.numpy(), gold["boundaries"])
    # 0.7 and 0.9 round DOWN in fp32
    assert float(b[7]) < 0.7 and float(b[9]) < 0.9


def test_empty_inputs():
    lg = torch.zeros(0, 5); lb = torch.zeros(0, dtype=torch.long)
    assert om.compute_accuracy(lg, lb) == [0.0]
    assert om.compute_aece(lg, lb) == 0.0
    assert om.compute_aece_with_bins(lg, lb)[1] == {"bin_acc": [], "bin_conf": [], "bin_count": []}
