"""GPU: the drop-in trainers (clip_gp_b200/trainers.py: "Adapter", "TaskRes", "CLIP-Adapter", "Tip-Adapter" through the
reference-named registry) replay the reference's OWN Trainer.train() runs recorded in tests/golden/ref_train.npz
(tests/golden/make_ref_golden.py: reference trainers executed unmodified on a stand-in CLIP): same cached features, same batch
order, same base-noise stream -> per-step losses, learning rates, final parameters, zero-shot and final accuracy / ECE / AECE."""
import os
import types

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from clip_gp_b200 import trainers
from tests.helpers import max_err, rel_err

pytestmark = pytest.mark.gpu
SEED, S_TRAIN, S_EVAL, PRE_EPOCHS = 21, 3, 5, 4


@pytest.fixture(scope="module")
def G(golden_dir):
    return np.load(os.path.join(golden_dir, "ref_train.npz"))


def T_(G, key):
    return torch.from_numpy(G[key])


def make_config(name, kernel, use_gp, precision="bf16x3", **adapter):
    ns = types.SimpleNamespace
    a = ns(use_gp=use_gp, gp_kernel_type=kernel, gp_pca_dim=8, num_templates=4, gp_num_mc_samples_train=S_TRAIN,
           gp_num_mc_samples_eval=S_EVAL, gp_lr=1e-3, gp_beta=0.01, l2_lambda=0.5, clip_adapter_epochs=3, gp_prior_temp=1.0,
           freeze_visual_proj=False, clipgp_precision=precision, template_init_method="uniform")
    for k, v in adapter.items():
        setattr(a, k, v)
    return ns(trainer_name=name, adapter=a, seed=SEED, output_dir=None,
              optim=ns(name="adamw", lr=0.01, max_epoch=3, lr_scheduler="cosine", weight_decay=0.0, betas=(0.9, 0.999), momentum=0.9),
              dataset=ns(num_shots=4, name="synthetic"), dataloader=ns(batch_size_train=16, batch_size_test=32),
              model=ns(backbone_name="fake"))


def data_manager(G, f_tr=None, y_tr=None):
    return trainers.FeatureDataManager(text_embeddings=T_(G, "world/E"), features_train=T_(G, "world/f_tr") if f_tr is None else f_tr,
                                       labels_train=T_(G, "world/y_tr") if y_tr is None else y_tr, features_test=T_(G, "world/f_te"),
                                       labels_test=T_(G, "world/y_te"), features_val=T_(G, "world/f_va"), labels_val=T_(G, "world/y_va"))


def load_gp(gp, G, key, buf_key=None):
    """Give the drop-in module the reference module's buffers (PCA axes are unique up to sign only) and parameters."""
    gp.variational_strategy._maybe_init()
    q = gp.variational_strategy._variational_distribution
    bk = buf_key or key
    with torch.no_grad():
        gp._templates_red.copy_(T_(G, f"{bk}/templates_red")); gp._pca_W_buf.copy_(T_(G, f"{bk}/pca_W")); gp._pca_mean_buf.copy_(T_(G, f"{bk}/pca_mean"))
        gp.variational_strategy.inducing_points.copy_(T_(G, f"{key}/Z"))
        q.variational_mean.copy_(T_(G, f"{key}/m")); q.chol_variational_covar.copy_(T_(G, f"{key}/chol"))
        raw_ls, raw_os, raw_var = gp._kernel_raw()
        if raw_ls is not None: raw_ls.copy_(T_(G, f"{key}/raw_lengthscale"))
        if raw_os is not None: raw_os.copy_(T_(G, f"{key}/raw_outputscale"))
        if raw_var is not None: raw_var.copy_(T_(G, f"{key}/raw_variance"))


def gp_tensors(gp):
    q = gp.variational_strategy._variational_distribution
    out = {"Z": gp.variational_strategy.inducing_points, "m": q.variational_mean, "chol": q.chol_variational_covar}
    raw_ls, raw_os, raw_var = gp._kernel_raw()
    if raw_ls is not None: out["raw_lengthscale"] = raw_ls
    if raw_os is not None: out["raw_outputscale"] = raw_os
    if raw_var is not None: out["raw_variance"] = raw_var
    return out


def record_losses(tr):
    losses = []
    orig = tr.forward_backward

    def fb(batch):
        out = orig(batch)
        losses.append(out["loss"].detach().clone().reshape(()))
        return out
    tr.forward_backward = fb
    return losses


def check_metrics(G, key, m, ece_abs=2e-3):
    acc_key = "top1_acc" if f"{key}/top1_acc" in G else "accuracy"
    assert m["top1_acc"] == pytest.approx(float(G[f"{key}/{acc_key}"]), abs=1e-9)
    assert m["ece"] == pytest.approx(float(G[f"{key}/ece"]), rel=1e-3, abs=ece_abs)
    assert m["aece"] == pytest.approx(float(G[f"{key}/aece"]), rel=1e-3, abs=ece_abs)
    if f"{key}/calibration/bin_count" in G:
        assert list(m["calibration"]["bin_count"]) == list(G[f"{key}/calibration/bin_count"])


def moved_budget(got, ref0, ref1, frac):
    moved = float((ref1 - ref0).abs().max())
    return float((got.detach().cpu() - ref1).abs().max()) <= frac * moved + 1e-6


# ---------------------------------------------------------------------------------------------------------------- Adapter
@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
@pytest.mark.parametrize("kernel", ["rbf", "matern", "linear"])
def test_adapter_trainer_replays_the_reference_run(G, kernel, precision):
    key = f"adapter/{kernel}"
    cfg = make_config("Adapter", kernel, True, precision, template_init_method="val_weighted")
    tr = trainers.build_trainer(cfg, data_manager(G))
    assert isinstance(tr, trainers.AdapterTrainer)
    tr.build_model()
    load_gp(tr.model.gp_weighter, G, f"{key}/init", buf_key=key)
    with torch.no_grad():
        tr.model.gp_weighter.mean_module.f0.copy_(T_(G, f"{key}/f0"))
        tr.model.visual_proj.weight.copy_(T_(G, f"{key}/init/W"))
    eng = tr.build_engine()
    eng.eval_eps = T_(G, f"{key}/eps_eval").cuda()
    bf, by = T_(G, f"{key}/batches_f"), T_(G, f"{key}/batches_y")
    tr.batch_plan = iter(list(zip(bf, by)))
    losses = record_losses(tr)
    lrs = []
    orig_step = eng.train_step
    eng.train_step = lambda f, y, **k: (lrs.append(eng.lr_dev.clone()), orig_step(f, y, **k))[1]
    tr.train()
    check_metrics(G, f"{key}/zero_shot", tr.zero_shot_metrics)
    got = torch.stack(losses).cpu().numpy()
    assert got == pytest.approx(G[f"{key}/losses"], rel=3e-3)
    assert torch.stack(lrs).cpu().numpy() == pytest.approx(G[f"{key}/lrs"], rel=1e-5)
    gp = tr.model.gp_weighter
    for name, p in gp_tensors(gp).items():
        frac = 0.5 if (kernel == "matern" and name == "Z") else 0.05        # see tests/test_ref_train_golden.py
        assert moved_budget(p, T_(G, f"{key}/init/{name}"), T_(G, f"{key}/final/{name}"), frac), name
    assert moved_budget(tr.model.visual_proj.weight, T_(G, f"{key}/init/W"), T_(G, f"{key}/final/W"), 0.05)
    assert torch.equal(gp.variational_strategy.inducing_points[:, :-1].cpu(), T_(G, f"{key}/init/Z")[:, :-1])
    # final evaluation of OUR trained state (fused projection + logits + calibration GEMM; never materialises logits)
    m = tr._compute_final_metrics()
    check_metrics(G, f"{key}/final_metrics", m, ece_abs=5e-2)               # parameters differ by the budget above
    # ... and of the REFERENCE's final state through the same path: exact accuracy / bin counts, ECE to 1e-3
    load_gp(gp, G, f"{key}/final", buf_key=key)
    with torch.no_grad():
        tr.model.visual_proj.weight.copy_(T_(G, f"{key}/final/W"))
    eng = tr.build_engine()
    eng.eval_eps = T_(G, f"{key}/eps_eval").cuda()
    check_metrics(G, f"{key}/final_metrics", tr._compute_final_metrics())


# ---------------------------------------------------------------------------------------------------------------- GP pre-training
HEAD_KERNEL = {"taskres": "rbf", "clip_adapter": "linear", "tip": "matern"}


def pretrained_trainer(G, which, name, precision, train_set=None, **adapter):
    """Trainer whose weighter starts from the reference's pre-training start state and sees the reference's few-shot feature order."""
    kernel = HEAD_KERNEL[which.split("/")[0]]
    cfg = make_config(name, kernel, True, precision, **adapter)
    cfg.optim.max_epoch = PRE_EPOCHS
    cfg.dataloader.batch_size_train = 8
    f_tr, y_tr = train_set if train_set is not None else (T_(G, f"{which}/pretrain_f"), T_(G, f"{which}/pretrain_y"))
    tr = trainers.build_trainer(cfg, data_manager(G, f_tr, y_tr))
    from clip_gp_b200.gp_template_weigher import GaussianProcessTemplateWeighter
    E = T_(G, f"{which}/gp_before_pretrain/templates").cuda()
    gp = GaussianProcessTemplateWeighter(E, cfg, rng="philox", seed=SEED).cuda()
    load_gp(gp, G, f"{which}/gp_before_pretrain")
    with torch.no_grad():
        gp.mean_module.f0.copy_(T_(G, f"{which}/gp_before_pretrain/f0"))
    gp.eval_eps = T_(G, f"{which}/eps_eval").cuda()
    tr.gp_weighter = gp
    tr.eval_eps = gp.eval_eps
    return tr, gp, kernel


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
@pytest.mark.parametrize("which,name", [("taskres/gp", "TaskRes"), ("clip_adapter/gp", "CLIP-Adapter"), ("tip/gp/cache", "Tip-Adapter")])
def test_gp_pretrain_on_the_engine_matches_the_reference_loop(G, which, name, precision):
    """heads.gp_pretrain (fused engine, collapsed logit-mean, own AdamW + cosine) against the reference's full-batch ELBO loop
    (taskres.py:254-289 == clip_adapter.py:257-290 == tip_adapter.py:122-157) and its prototype initialisation."""
    tr, gp, kernel = pretrained_trainer(G, which, name, precision)
    protos = tr._maybe_gp_pretrain(name)
    assert protos is not None and int(gp._rng_state[1]) == PRE_EPOCHS
    for nm, p in gp_tensors(gp).items():
        frac = 0.5 if (kernel == "matern" and nm == "Z") else 0.05
        assert moved_budget(p, T_(G, f"{which}/gp_before_pretrain/{nm}"), T_(G, f"{which}/gp_after_pretrain/{nm}"), frac), nm
    target = T_(G, f"{which}/base_text_features") if name == "TaskRes" else T_(G, f"{which}/clip_weights").t()
    assert rel_err(protos, target) < 2e-3
    assert max_err(protos, target) < 1e-3


# ---------------------------------------------------------------------------------------------------------------- TaskRes
@pytest.mark.parametrize("variant", ["plain", "gp"])
def test_taskres_trainer_replays_the_reference_run(G, variant):
    key = f"taskres/{variant}"
    extra = dict(taskres_optimizer="adam", taskres_lr=2e-3, taskres_epochs=5, taskres_residual_scale=0.5)
    if variant == "gp":
        tr, gp, _ = pretrained_trainer(G, key, "TaskRes", "fp32", **extra)
        load_gp(gp, G, f"{key}/gp_after_pretrain")           # the main loop is checked from the reference's pre-trained state
        tr._maybe_gp_pretrain = lambda tag: gp.mean_prototypes(S_EVAL, eps=gp.eval_eps)
        gp._rng_state[1] = PRE_EPOCHS
    else:
        cfg = make_config("TaskRes", "rbf", False, **extra)
        cfg.dataloader.batch_size_train = 8
        tr = trainers.build_trainer(cfg, data_manager(G))
    tr.build_model()
    bf, by = T_(G, f"{key}/batches_f"), T_(G, f"{key}/batches_y")
    tr.batch_plan = iter(list(zip(bf, by)))
    losses = record_losses(tr)
    lrs = []
    orig = tr.forward_backward
    tr.forward_backward = lambda b: (lrs.append(tr.optim.param_groups[0]["lr"]), orig(b))[1]
    tr.train()
    assert tr.zero_shot_metrics["top1_acc"] == pytest.approx(float(G[f"{key}/zero_shot_acc"]), abs=1e-9)
    assert rel_err(tr.model.base_text_features, T_(G, f"{key}/base_text_features")) < 2e-3
    assert torch.stack(losses).cpu().numpy() == pytest.approx(G[f"{key}/losses"], rel=3e-3, abs=1e-4)
    assert np.array(lrs) == pytest.approx(G[f"{key}/lrs"][:, 0], rel=1e-6)
    ref_x = T_(G, f"{key}/final/residuals")
    assert float((tr.model.text_feature_residuals.detach().cpu() - ref_x).abs().max()) < 0.05 * float(ref_x.abs().max())
    check_metrics(G, f"{key}/final_metrics", tr._compute_final_metrics(), ece_abs=5e-2)


# ---------------------------------------------------------------------------------------------------------------- CLIP-Adapter
@pytest.mark.parametrize("variant", ["plain", "gp"])
def test_clip_adapter_trainer_replays_the_reference_run(G, variant):
    key = f"clip_adapter/{variant}"
    extra = dict(clip_adapter_optimizer="adam", clip_adapter_lr=1e-3, clip_adapter_epochs=3, clip_adapter_ratio=0.2, clip_adapter_reduction=4)
    if variant == "gp":
        tr, gp, _ = pretrained_trainer(G, key, "CLIP-Adapter", "fp32", **extra)
        load_gp(gp, G, f"{key}/gp_after_pretrain")
        tr._maybe_gp_pretrain = lambda tag: gp.mean_prototypes(S_EVAL, eps=gp.eval_eps)
        gp._rng_state[1] = PRE_EPOCHS
    else:
        cfg = make_config("CLIP-Adapter", "linear", False, **extra)
        cfg.dataloader.batch_size_train = 8
        tr = trainers.build_trainer(cfg, data_manager(G))
    tr.build_model()
    assert rel_err(tr.model.clip_weights, T_(G, f"{key}/init/clip_weights")) < 1e-4          # _get_clip_weights
    with torch.no_grad():
        tr.model.adapter.fc1.weight.copy_(T_(G, f"{key}/init/fc1")); tr.model.adapter.fc2.weight.copy_(T_(G, f"{key}/init/fc2"))
    bf, by = T_(G, f"{key}/batches_f"), T_(G, f"{key}/batches_y")
    tr.batch_plan = iter(list(zip(bf, by)))
    losses = record_losses(tr)
    tr.train()
    assert tr.zero_shot_metrics["top1_acc"] == pytest.approx(float(G[f"{key}/zero_shot_acc"]), abs=1e-9)
    assert rel_err(tr.model.clip_weights, T_(G, f"{key}/clip_weights")) < 2e-3
    assert torch.stack(losses).cpu().numpy() == pytest.approx(G[f"{key}/losses"], rel=3e-3, abs=1e-4)
    for nm, p in (("fc1", tr.model.adapter.fc1.weight), ("fc2", tr.model.adapter.fc2.weight)):
        assert moved_budget(p, T_(G, f"{key}/init/{nm}"), T_(G, f"{key}/final/{nm}"), 0.05), nm
    check_metrics(G, f"{key}/final_metrics", tr._compute_final_metrics(), ece_abs=5e-2)


# ---------------------------------------------------------------------------------------------------------------- Tip-Adapter
@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
@pytest.mark.parametrize("variant", ["plain", "gp"])
@pytest.mark.parametrize("mode", ["cache", "F"])
def test_tip_adapter_trainer_replays_the_reference_run(G, variant, mode, precision):
    key = f"tip/{variant}/{mode}"
    extra = dict(tip_adapter_trainable=(mode == "F"), tip_adapter_lr=1e-3, tip_adapter_eps=1e-4, tip_adapter_epochs=3,
                 tip_adapter_init_alpha=20.0, tip_adapter_init_beta=2.0)
    f0, y0 = T_(G, f"{key}/cache_keys0"), T_(G, f"{key}/cache_labels0")     # the reference's (shuffled) cache order
    if variant == "gp":
        tr, gp, _ = pretrained_trainer(G, key, "Tip-Adapter", precision, train_set=(f0, y0), **extra)
        load_gp(gp, G, f"{key}/gp_after_pretrain")
        tr._maybe_gp_pretrain = lambda tag: gp.mean_prototypes(S_EVAL, eps=gp.eval_eps)
        gp.eval()
    else:
        cfg = make_config("Tip-Adapter", "matern", False, precision, **extra)
        cfg.dataloader.batch_size_train = 8
        tr = trainers.build_trainer(cfg, data_manager(G, f0, y0))
    if mode == "F":
        tr.batch_plan = iter(list(zip(T_(G, f"{key}/batches_f"), T_(G, f"{key}/batches_y"))))
    tr.train()
    assert tr.zero_shot_metrics["top1_acc"] == pytest.approx(float(G[f"{key}/zero_shot_acc"]), abs=1e-9)
    assert (tr._tip_adapter_best_beta, tr._tip_adapter_best_alpha) == (float(G[f"{key}/best_beta"]), float(G[f"{key}/best_alpha"]))
    if mode == "F":
        # keys are class-sorted here; compare the trained keys as a set of (label, row) pairs via their class sums
        ref_w, lab_ref = T_(G, f"{key}/final/adapter_w"), T_(G, f"{key}/cache_labels0")
        got_w, lab_got = tr.cache_keys.cpu(), tr.cache_labels.cpu()
        order = torch.argsort(lab_ref, stable=True)
        assert torch.equal(lab_ref[order], lab_got)
        moved = float((ref_w - T_(G, f"{key}/cache_keys0")).abs().max())
        assert float((got_w - ref_w[order]).abs().max()) < 0.05 * moved + 1e-6
    check_metrics(G, f"{key}/final_metrics", tr._compute_final_metrics(), ece_abs=5e-3)
