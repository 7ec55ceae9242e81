"""Generate the golden fixtures under tests/golden/.

Run in the BUILD container only (needs /root/reference):  python tests/golden/make_golden.py

* metrics_golden.npz  – inputs + outputs of the REFERENCE's own utils/metrics.py (imported by file
  path; importing it as ``utils.metrics`` fails because utils/__init__.py pulls in clip -> ftfy).
  This pins oracle/metrics.py and the CUDA binning kernels.
* gp_selfgolden.npz   – outputs of oracle/gp.py (self-golden, PARITY UNPINNED: gpytorch/entmax are
  not installable here).  Guards the oracle against silent drift and travels to the GPU box.
"""
from __future__ import annotations

import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def load_reference_metrics():
    spec = importlib.util.spec_from_file_location("ref_metrics", "/root/reference/utils/metrics.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def metric_cases():
    """name -> (logits, labels).  Covers ragged N, N < n_bins, peaked and flat confidences, exact ties."""
    g = torch.Generator().manual_seed(20261018)
    cases = {}
    lg = 3.0 * torch.randn(600, 50, generator=g); lb = torch.randint(0, 50, (600,), generator=g)
    cases["rand_600x50"] = (lg, lb)
    lg = 8.0 * torch.randn(257, 12, generator=g); lb = lg.argmax(1).clone(); lb[::3] = torch.randint(0, 12, (86,), generator=g)
    cases["peaked_257x12"] = (lg, lb)
    lg = 0.05 * torch.randn(123, 7, generator=g); lb = torch.randint(0, 7, (123,), generator=g)
    cases["flat_123x7"] = (lg, lb)
    lg = torch.randn(5, 4, generator=g); lb = torch.randint(0, 4, (5,), generator=g)
    cases["tiny_5x4"] = (lg, lb)
    lg = torch.zeros(40, 2); lg[:, 0] = torch.linspace(-4, 4, 40); lb = (torch.arange(40) % 2)
    cases["binary_40x2"] = (lg, lb)
    lg = torch.randn(64, 10, generator=g).round(); lb = torch.randint(0, 10, (64,), generator=g)  # many exact ties
    cases["ties_64x10"] = (lg, lb)
    lg = torch.zeros(30, 3); lg[:, 1] = 40.0; lb = torch.ones(30, dtype=torch.long)               # conf == 1.0 exactly
    cases["saturated_30x3"] = (lg, lb)
    return cases


def main():
    ref = load_reference_metrics()
    out = {}
    for name, (lg, lb) in metric_cases().items():
        out[f"{name}/logits"] = lg.numpy()
        out[f"{name}/labels"] = lb.numpy()
        out[f"{name}/acc"] = np.float64(ref.compute_accuracy(lg, lb)[0])
        out[f"{name}/ece"] = np.float64(ref.compute_ece(lg, lb))
        out[f"{name}/aece"] = np.float64(ref.compute_aece(lg, lb))
        e, b = ref.compute_ece_with_bins(lg, lb)
        out[f"{name}/ece_b"] = np.float64(e)
        out[f"{name}/ece_bin_acc"] = np.array(b["bin_acc"], np.float64)
        out[f"{name}/ece_bin_conf"] = np.array(b["bin_conf"], np.float64)
        out[f"{name}/ece_bin_count"] = np.array(b["bin_count"], np.int64)
        e, b = ref.compute_aece_with_bins(lg, lb)
        out[f"{name}/aece_b"] = np.float64(e)
        out[f"{name}/aece_bin_acc"] = np.array(b["bin_acc"], np.float64)
        out[f"{name}/aece_bin_conf"] = np.array(b["bin_conf"], np.float64)
        out[f"{name}/aece_bin_count"] = np.array(b["bin_count"], np.int64)
        # the intermediate the CUDA kernel must reproduce: confidences, predictions
        conf, pred = torch.softmax(lg, -1).max(-1)
        out[f"{name}/conf"] = conf.numpy()
        out[f"{name}/pred"] = pred.numpy()
    out["boundaries"] = torch.linspace(0, 1, 11).numpy()
    np.savez_compressed(os.path.join(HERE, "metrics_golden.npz"), **out)
    print("wrote metrics_golden.npz with", len(out), "arrays")

    # ---- GP self-golden (oracle output; unpinned) ----
    from oracle import gp
    from clip_gp_b200 import synth
    gout = {}
    wl = synth.make_workload("tiny")
    shp = wl["shape"]
    for kern in ("rbf", "matern", "linear"):
        st = gp.build_state(wl["E"], kern, shp.d)
        m, Lq = synth.trained_like_q(shp.C, shp.T + 1, 5)
        st.var_mean, st.chol_var = m, Lq
        eps = torch.randn(shp.C, shp.T, shp.S, generator=torch.Generator().manual_seed(77))
        P, aux = gp.sample_prototypes(st, eps)
        gout[f"{kern}/w"] = aux["w"].numpy()
        gout[f"{kern}/mu"] = aux["mu"].numpy()
        gout[f"{kern}/Sigma"] = aux["Sigma"].numpy()
        gout[f"{kern}/protos"] = P.numpy()
        gout[f"{kern}/kl"] = gp.kl_divergence(m, Lq).numpy()
    np.savez_compressed(os.path.join(HERE, "gp_selfgolden.npz"), **gout)
    print("wrote gp_selfgolden.npz with", len(gout), "arrays")


if __name__ == "__main__":
    main()
