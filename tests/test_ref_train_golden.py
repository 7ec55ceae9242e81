"""CPU: the oracle's training step / heads / metrics against tests/golden/ref_train.npz — the reference's OWN trainers run end to
end (Trainer.train() of trainers/adapter.py etc., unmodified) on a stand-in CLIP that returns cached features
(tests/golden/make_ref_golden.py, tests/golden/_fake_clip.py).  Pins oracle/train_step.py, oracle/heads.py and the trainer-level
semantics (optimizer groups, per-epoch cosine schedule, loss assembly, MC-averaged evaluation)."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import gp as ogp
from oracle import heads as oh
from oracle import metrics as om
from oracle import philox
from oracle.train_step import OracleAdapter
from tests.helpers import rel_err

KERNELS = ["rbf", "matern", "linear"]


@pytest.fixture(scope="module")
def G(golden_dir):
    return np.load(os.path.join(golden_dir, "ref_train.npz"))


def T_(G, key):
    return torch.from_numpy(G[key])


def adapter_state(G, kernel, which="init"):
    key = f"adapter/{kernel}"
    kp = ogp.KernelParams(kernel)
    for name in ("raw_lengthscale", "raw_outputscale", "raw_variance"):
        if f"{key}/{which}/{name}" in G:
            setattr(kp, name, T_(G, f"{key}/{which}/{name}").clone())
    C, T = G[f"{key}/f0"].shape
    return ogp.GPState(templates=T_(G, f"{key}/text_embeddings"), templates_red=T_(G, f"{key}/templates_red"),
                       inducing_points=T_(G, f"{key}/{which}/Z").clone(), var_mean=T_(G, f"{key}/{which}/m").clone(),
                       chol_var=T_(G, f"{key}/{which}/chol").clone(), kernel=kp, f0=T_(G, f"{key}/f0"),
                       cls_bias=torch.zeros(C, 1), tmp_bias=torch.zeros(1, T), pca_mean=T_(G, f"{key}/pca_mean"), pca_W=T_(G, f"{key}/pca_W"))


def cosine(base, epoch, t_max):
    return base * 0.5 * (1.0 + math.cos(math.pi * epoch / t_max))


@pytest.mark.parametrize("kernel", KERNELS)
def test_adapter_trainer_trajectory(G, kernel):
    """Six optimisation steps (3 epochs x 2 batches) of the reference's Adapter trainer: per-step loss, learning rates and the
    final parameters (adapter.py:328-385, 387-476, 290-311; utils/optimization.py:147-238; utils/trainer.py:466-470)."""
    key = f"adapter/{kernel}"
    st = adapter_state(G, kernel)
    D = st.templates.shape[-1]
    orc = OracleAdapter(st, D, scale=100.0, gp_beta=0.01, l2_lambda=0.5, shots=4, lr=0.01, gp_lr=1e-3, loss_mode="per_sample")
    with torch.no_grad():
        orc.W.copy_(T_(G, f"{key}/init/W"))
    bf, by, losses_ref, lrs_ref = T_(G, f"{key}/batches_f"), T_(G, f"{key}/batches_y"), G[f"{key}/losses"], G[f"{key}/lrs"]
    seed = int(G[f"{key}/philox_seed"])
    C, T = G[f"{key}/f0"].shape
    steps_per_epoch = 2
    for it in range(bf.shape[0]):
        ep = it // steps_per_epoch
        lr, gp_lr = cosine(0.01, ep, 3), cosine(1e-3, ep, 3)              # CosineAnnealingLR(T_max=optim.max_epoch) stepped per epoch
        assert lrs_ref[it] == pytest.approx([lr, gp_lr], rel=1e-6)
        orc.opt.param_groups[0]["lr"], orc.opt.param_groups[1]["lr"] = lr, gp_lr
        loss = orc.step(bf[it], by[it], philox.eps_tensor(seed, it, C, T, 3))
        assert loss == pytest.approx(float(losses_ref[it]), rel=2e-3), it
    # final parameters after six AdamW steps: AdamW's first steps are sign-like (|update| ~ lr), compare with a budget of 5 % of
    # the distance travelled
    for name, got in (("W", orc.W), ("m", orc.st.var_mean), ("chol", orc.st.chol_var), ("Z", orc.st.inducing_points)):
        ref0, ref1 = T_(G, f"{key}/init/{name}"), T_(G, f"{key}/final/{name}")
        moved = float((ref1 - ref0).abs().max())
        assert moved > 0
        # Matern-1/2: the reference's own fp32 gradient of the learnable inducing row is 10-50 % off float64 (sq_dist expansion +
        # sqrt, tools/diag_golden_grad.py) and Adam turns gradient noise into sign flips of whole steps
        budget = 0.5 if (kernel == "matern" and name == "Z") else 0.05
        assert float((got.detach() - ref1).abs().max()) < budget * moved + 1e-6, name
    assert torch.equal(orc.st.inducing_points.detach()[:, :-1], T_(G, f"{key}/init/Z")[:, :-1])      # frozen template rows


@pytest.mark.parametrize("kernel", KERNELS)
def test_adapter_trainer_final_and_zero_shot_metrics(G, kernel):
    """BaseTrainer.test() of the reference (utils/trainer.py:474-557) on the final model, and the zero-shot pass before training
    (adapter.py:589-611), with the eval noise the golden run used: accuracy, ECE, AECE and the integer bin counts."""
    key = f"adapter/{kernel}"
    f_te, y_te = T_(G, "world/f_te"), T_(G, "world/y_te")
    eps_eval = T_(G, f"{key}/eps_eval")
    for which, mkey in (("init", "zero_shot"), ("final", "final_metrics")):
        st = adapter_state(G, kernel, which)
        W = T_(G, f"{key}/{which}/W")
        with torch.no_grad():
            protos, _ = ogp.sample_prototypes(st, eps_eval)
            logits = oh.adapter_logits(f_te, W, protos, 100.0)
        acc_key = "top1_acc" if f"{key}/{mkey}/top1_acc" in G else "accuracy"
        assert om.compute_accuracy(logits, y_te)[0] == pytest.approx(float(G[f"{key}/{mkey}/{acc_key}"]), abs=1e-9)
        assert om.compute_ece(logits, y_te) == pytest.approx(float(G[f"{key}/{mkey}/ece"]), rel=1e-3, abs=1e-3)
        assert om.compute_aece(logits, y_te) == pytest.approx(float(G[f"{key}/{mkey}/aece"]), rel=1e-3, abs=1e-3)
        _, bins = om.compute_ece_with_bins(logits, y_te)
        assert list(bins["bin_count"]) == list(G[f"{key}/{mkey}/calibration/bin_count"])


# ======================================================================================================================
# TaskRes / CLIP-Adapter / Tip-Adapter(-F): the reference's whole train() per variant (tests/golden/make_ref_golden.py:
# head_trainer_goldens) against oracle/heads.py
# ======================================================================================================================
import torch.nn.functional as F

SEED, S_TRAIN, S_EVAL, BETA_KL, GP_LR, PRE_EPOCHS = 21, 3, 5, 0.01, 1e-3, 4


def gp_state(G, key, kernel):
    kp = ogp.KernelParams(kernel)
    for name in ("raw_lengthscale", "raw_outputscale", "raw_variance"):
        if f"{key}/{name}" in G:
            setattr(kp, name, T_(G, f"{key}/{name}").clone())
    C, T = G[f"{key}/f0"].shape
    return ogp.GPState(templates=T_(G, f"{key}/templates"), templates_red=T_(G, f"{key}/templates_red"),
                       inducing_points=T_(G, f"{key}/Z").clone(), var_mean=T_(G, f"{key}/m").clone(), chol_var=T_(G, f"{key}/chol").clone(),
                       kernel=kp, f0=T_(G, f"{key}/f0"), cls_bias=torch.zeros(C, 1), tmp_bias=torch.zeros(1, T),
                       pca_mean=T_(G, f"{key}/pca_mean"), pca_W=T_(G, f"{key}/pca_W"))


def gp_trainables(st):
    ps = [st.inducing_points, st.var_mean, st.chol_var]
    for p in (st.kernel.raw_lengthscale, st.kernel.raw_outputscale, st.kernel.raw_variance):
        if p is not None:
            ps.append(p)
    return ps


def replay_pretrain(G, key, kernel):
    """taskres.py:254-280 == clip_adapter.py:257-279 == tip_adapter.py:122-146 with the oracle; returns the trained state."""
    st = gp_state(G, f"{key}/gp_before_pretrain", kernel)
    f = F.normalize(T_(G, f"{key}/pretrain_f"), dim=-1); y = T_(G, f"{key}/pretrain_y")
    ps = gp_trainables(st)
    for p in ps:
        p.requires_grad_(True)
    opt = torch.optim.AdamW(ps, lr=GP_LR, weight_decay=0.0)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, PRE_EPOCHS)
    C, T = st.f0.shape
    mask = torch.zeros_like(st.inducing_points); mask[:, -1] = 1.0
    for ep in range(PRE_EPOCHS):
        protos, _ = ogp.sample_prototypes(st, philox.eps_tensor(SEED, ep, C, T, S_TRAIN))
        loss, _ = oh.gp_pretrain_loss(f, y, protos, ogp.kl_divergence(st.var_mean, st.chol_var), BETA_KL)
        opt.zero_grad(); loss.backward()
        st.inducing_points.grad.mul_(mask)
        opt.step(); sched.step()
    return st


HEAD_KERNEL = {"taskres": "rbf", "clip_adapter": "linear", "tip": "matern"}


@pytest.mark.parametrize("which", ["taskres/gp", "clip_adapter/gp", "tip/gp/cache"])
def test_gp_pretrain_loop_and_prototype_init(G, which):
    kernel = HEAD_KERNEL[which.split("/")[0]]
    st = replay_pretrain(G, which, kernel)
    ref = gp_state(G, f"{which}/gp_after_pretrain", kernel)
    st0 = gp_state(G, f"{which}/gp_before_pretrain", kernel)
    for a, b, b0, name in zip(gp_trainables(st), gp_trainables(ref), gp_trainables(st0), ("Z", "m", "chol", "h1", "h2")):
        moved = float((b - b0).abs().max())
        assert moved > 0, name
        budget = 0.5 if (kernel == "matern" and name == "Z") else 0.05     # Matern d z_last: the reference's fp32 gradient is noise
        assert float((a.detach() - b).abs().max()) < budget * moved + 1e-7, name
    # prototype init from the trained weighter: normalize(mean_s protos) with the eval draw (taskres.py:281-289 etc.)
    with torch.no_grad():
        protos, _ = ogp.sample_prototypes(ref, T_(G, f"{which}/eps_eval"))
    init = oh.gp_mean_prototypes(protos)
    got_key = {"taskres/gp": "base_text_features", "clip_adapter/gp": "clip_weights", "tip/gp/cache": "clip_weights"}[which]
    target = T_(G, f"{which}/{got_key}")
    if got_key == "clip_weights":
        target = target.t()                                                 # reference stores [D,K]
    assert rel_err(init, target) < 1e-3


@pytest.mark.parametrize("variant", ["plain", "gp"])
def test_taskres_trainer_trajectory_and_metrics(G, variant):
    """TaskResLearner + CustomCLIP.forward (taskres.py:45-47, 96-123), Adam(taskres_lr) + CosineAnnealingLR(T_max=taskres_epochs)
    stepped per epoch while the loop runs clip_adapter_epochs epochs (:158-173; utils/trainer.py:256)."""
    key = f"taskres/{variant}"
    base = T_(G, f"{key}/base_text_features")
    x = torch.zeros_like(base, requires_grad=True)
    opt = torch.optim.Adam([x], lr=2e-3)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=5)
    bf, by, losses = T_(G, f"{key}/batches_f"), T_(G, f"{key}/batches_y"), G[f"{key}/losses"]
    st = gp_state(G, f"{key}/gp_after_pretrain", "rbf") if variant == "gp" else None
    C, T = (st.f0.shape if st is not None else (0, 0))
    nb = bf.shape[0] // 3
    for it in range(bf.shape[0]):
        assert G[f"{key}/lrs"][it][0] == pytest.approx(opt.param_groups[0]["lr"], rel=1e-6)
        protos = None
        if st is not None:
            with torch.no_grad():
                protos, _ = ogp.sample_prototypes(st, philox.eps_tensor(SEED, PRE_EPOCHS + it, C, T, S_TRAIN))
        loss = F.cross_entropy(oh.taskres_logits(bf[it], base, x, 0.5, 100.0, protos), by[it])
        assert float(loss) == pytest.approx(float(losses[it]), rel=2e-3, abs=1e-4), it
        opt.zero_grad(); loss.backward(); opt.step()
        if (it + 1) % nb == 0:
            sched.step()
    ref_x = T_(G, f"{key}/final/residuals")
    assert float((x.detach() - ref_x).abs().max()) < 0.05 * float(ref_x.abs().max())
    with torch.no_grad():
        protos = None
        if st is not None:
            protos, _ = ogp.sample_prototypes(st, T_(G, f"{key}/eps_eval"))
        logits = oh.taskres_logits(T_(G, "world/f_te"), base, ref_x, 0.5, 100.0, protos)
    y = T_(G, "world/y_te")
    assert om.compute_accuracy(logits, y)[0] == pytest.approx(float(G[f"{key}/final_metrics/top1_acc"]), abs=1e-9)
    assert om.compute_ece(logits, y) == pytest.approx(float(G[f"{key}/final_metrics/ece"]), rel=1e-3, abs=1e-3)
    assert om.compute_aece(logits, y) == pytest.approx(float(G[f"{key}/final_metrics/aece"]), rel=1e-3, abs=1e-3)


@pytest.mark.parametrize("variant", ["plain", "gp"])
def test_clip_adapter_trainer_trajectory_and_metrics(G, variant):
    """AdapterMLP + blend + logits (clip_adapter.py:16-32, 77-100), Adam(clip_adapter_lr) + CosineAnnealingLR(T_max=clip_adapter_epochs)."""
    key = f"clip_adapter/{variant}"
    fc1 = T_(G, f"{key}/init/fc1").clone().requires_grad_(True); fc2 = T_(G, f"{key}/init/fc2").clone().requires_grad_(True)
    clip_w = T_(G, f"{key}/clip_weights")                                   # [D,K]; replaced by the GP prototypes in the gp variant
    opt = torch.optim.Adam([fc1, fc2], lr=1e-3)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=3)
    bf, by, losses = T_(G, f"{key}/batches_f"), T_(G, f"{key}/batches_y"), G[f"{key}/losses"]
    st = gp_state(G, f"{key}/gp_after_pretrain", "linear") if variant == "gp" else None
    C, T = (st.f0.shape if st is not None else (0, 0))
    nb = bf.shape[0] // 3
    for it in range(bf.shape[0]):
        protos = None
        if st is not None:
            with torch.no_grad():
                protos, _ = ogp.sample_prototypes(st, philox.eps_tensor(SEED, PRE_EPOCHS + it, C, T, S_TRAIN))
        fa = oh.clip_adapter_features(bf[it], fc1, fc2, 0.2)
        loss = F.cross_entropy(oh.clip_adapter_logits(fa, 100.0, clip_weights=clip_w, prototypes=protos), by[it])
        assert float(loss) == pytest.approx(float(losses[it]), rel=2e-3, abs=1e-4), it
        opt.zero_grad(); loss.backward(); opt.step()
        if (it + 1) % nb == 0:
            sched.step()
    for got, name in ((fc1, "fc1"), (fc2, "fc2")):
        ref0, ref1 = T_(G, f"{key}/init/{name}"), T_(G, f"{key}/final/{name}")
        assert float((got.detach() - ref1).abs().max()) < 0.05 * float((ref1 - ref0).abs().max()) + 1e-7
    with torch.no_grad():
        protos = None
        if st is not None:
            protos, _ = ogp.sample_prototypes(st, T_(G, f"{key}/eps_eval"))
        fa = oh.clip_adapter_features(T_(G, "world/f_te"), T_(G, f"{key}/final/fc1"), T_(G, f"{key}/final/fc2"), 0.2)
        logits = oh.clip_adapter_logits(fa, 100.0, clip_weights=clip_w, prototypes=protos)
    y = T_(G, "world/y_te")
    assert om.compute_accuracy(logits, y)[0] == pytest.approx(float(G[f"{key}/final_metrics/top1_acc"]), abs=1e-9)
    assert om.compute_ece(logits, y) == pytest.approx(float(G[f"{key}/final_metrics/ece"]), rel=1e-3, abs=1e-3)


@pytest.mark.parametrize("variant", ["plain", "gp"])
@pytest.mark.parametrize("mode", ["cache", "F"])
def test_tip_adapter_trainer(G, variant, mode):
    """_build_cache, affinity / cache logits / blend, the Tip-Adapter-F loop (AdamW(lr, eps) + per-step cosine) and
    _search_hyperparams (tip_adapter.py:43-80, 227-296, 298-334) + final metrics (:364-398)."""
    key = f"tip/{variant}/{mode}"
    K = 10
    f_tr, y_tr = F.normalize(T_(G, "world/f_tr"), dim=-1), T_(G, "world/y_tr")
    keys_ref, vals_ref = T_(G, f"{key}/cache_keys0"), T_(G, f"{key}/cache_vals")
    if mode == "cache":
        assert torch.equal(keys_ref, T_(G, f"{key}/cache_keys"))
    else:   # the trainable nn.Linear shares storage with cache_keys (tip_adapter.py:230): the attribute ends up trained
        assert torch.equal(T_(G, f"{key}/cache_keys"), T_(G, f"{key}/final/adapter_w"))
    # the cache is built from one shuffled pass over the few-shot loader: same multiset of (key, one-hot) rows
    assert torch.equal(vals_ref.argmax(1).sort().values, y_tr.sort().values)
    assert rel_err(keys_ref.norm(dim=-1), torch.ones(keys_ref.shape[0])) < 1e-5
    assert torch.equal(oh.tip_cache_vals(vals_ref.argmax(1), K), vals_ref)
    clip_w = T_(G, f"{key}/clip_weights")                                   # [D,K]
    st = gp_state(G, f"{key}/gp_after_pretrain", "matern") if variant == "gp" else None

    def clip_logits(fh):
        if st is None:
            return 100.0 * fh @ clip_w
        with torch.no_grad():
            protos, _ = ogp.sample_prototypes(st, T_(G, f"{key}/eps_eval"))
        p = protos / protos.norm(dim=-1, keepdim=True)
        return (100.0 * torch.einsum("bd,skd->bsk", fh, p)).mean(1)
    w = keys_ref.clone()
    if mode == "F":
        w.requires_grad_(True)
        bf, by = T_(G, f"{key}/batches_f"), T_(G, f"{key}/batches_y")
        opt = torch.optim.AdamW([w], lr=1e-3, eps=1e-4)
        sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, bf.shape[0])
        for it in range(bf.shape[0]):
            fh = F.normalize(bf[it], dim=-1)
            loss = F.cross_entropy(oh.tip_logits(fh, w, vals_ref, clip_logits(fh), 2.0, 20.0), by[it])
            opt.zero_grad(); loss.backward(); opt.step(); sched.step()
        ref_w = T_(G, f"{key}/final/adapter_w")
        assert float((w.detach() - ref_w).abs().max()) < 0.05 * float((ref_w - keys_ref).abs().max()) + 1e-7
        w = ref_w
    fv, yv = F.normalize(T_(G, "world/f_va"), dim=-1), T_(G, "world/y_va")
    with torch.no_grad():
        bb, ba, _ = oh.tip_search(fv, yv, w, vals_ref, clip_logits(fv), 2.0, 20.0)
        assert (bb, ba) == (float(G[f"{key}/best_beta"]), float(G[f"{key}/best_alpha"]))
        ft, yt = F.normalize(T_(G, "world/f_te"), dim=-1), T_(G, "world/y_te")
        logits = oh.tip_logits(ft, w, vals_ref, clip_logits(ft), bb, ba)
    assert om.compute_accuracy(logits, yt)[0] == pytest.approx(float(G[f"{key}/final_metrics/top1_acc"]), abs=1e-9)
    assert om.compute_ece(logits, yt) == pytest.approx(float(G[f"{key}/final_metrics/ece"]), rel=1e-3, abs=1e-3)
    assert om.compute_aece(logits, yt) == pytest.approx(float(G[f"{key}/final_metrics/aece"]), rel=1e-3, abs=1e-3)
