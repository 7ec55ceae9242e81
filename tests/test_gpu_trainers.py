"""GPU: the drop-in trainer registry on cached features: every registered trainer trains, evaluates and writes the
reference's metrics.json schema; losses go down; Tip-Adapter without training equals the oracle's closed form."""
import json
import os
from types import SimpleNamespace as NS

import pytest
import torch
import torch.nn.functional as F

from clip_gp_b200 import synth
from clip_gp_b200.trainers import TRAINER_REGISTRY, FeatureDataManager, build_trainer
from oracle import heads as oh
from oracle import metrics as om

pytestmark = pytest.mark.gpu


def make_dm(name="small"):
    wl = synth.make_workload(name)
    nv = 200
    return wl, FeatureDataManager(text_embeddings=wl["E"], features_train=wl["f_train"], labels_train=wl["y_train"],
                                  features_test=wl["f_test"][nv:], labels_test=wl["y_test"][nv:],
                                  features_val=wl["f_test"][:nv], labels_val=wl["y_test"][:nv])


def make_cfg(trainer, tmp_path, **adapter):
    a = dict(use_gp=False, gp_kernel_type="rbf", gp_pca_dim=32, gp_num_mc_samples_train=4, gp_num_mc_samples_eval=8, gp_lr=1e-2,
             gp_beta=0.01, l2_lambda=0.5, clip_adapter_epochs=3, taskres_epochs=3, taskres_lr=1e-3, taskres_residual_scale=0.5,
             clip_adapter_ratio=0.2, clip_adapter_reduction=4, tip_adapter_trainable=False, tip_adapter_lr=1e-3, tip_adapter_eps=1e-4,
             tip_adapter_epochs=2, tip_adapter_init_alpha=20.0, tip_adapter_init_beta=2.0, freeze_visual_proj=False)
    a.update(adapter)
    return NS(trainer_name=trainer, adapter=NS(**a), optim=NS(lr=1e-3, max_epoch=5, weight_decay=0.0), dataset=NS(name="synthetic", num_shots=4),
              dataloader=NS(batch_size_train=48), model=NS(backbone_name="synthetic"), seed=1, output_dir=str(tmp_path))


def test_registry_names_and_errors():
    assert {"Adapter", "TaskRes", "CLIP-Adapter", "Tip-Adapter"} <= set(TRAINER_REGISTRY.list_trainers())
    with pytest.raises(ValueError, match="Unknown trainer"):
        TRAINER_REGISTRY.get("Adapter-CoOp-nope")


@pytest.mark.parametrize("trainer,adapter", [("Adapter", dict(use_gp=True)), ("Adapter", dict(use_gp=False)),
                                             ("TaskRes", dict(use_gp=True)), ("TaskRes", dict()),
                                             ("CLIP-Adapter", dict(use_gp=True)), ("CLIP-Adapter", dict()),
                                             ("Tip-Adapter", dict(tip_adapter_trainable=True)), ("Tip-Adapter", dict(use_gp=True))])
def test_trainers_run_and_write_metrics_json(tmp_path, trainer, adapter):
    torch.manual_seed(0)
    wl, dm = make_dm()
    cfg = make_cfg(trainer, tmp_path, **adapter)
    tr = build_trainer(cfg, dm)
    acc = tr.train()
    assert 0.0 <= acc <= 100.0
    payload = json.load(open(os.path.join(str(tmp_path), "metrics.json")))
    assert set(payload) >= {"timestamp", "dataset", "shots", "seed", "method", "backbone", "zero_shot", "metrics", "config", "output_dir", "train_time_s"}
    m = payload["metrics"]
    assert set(m) >= {"top1_acc", "ece", "aece", "calibration", "adaptive_calibration"}
    assert len(m["calibration"]["bin_count"]) == 10 and sum(m["calibration"]["bin_count"]) == dm.labels_test.numel()
    assert payload["method"] == ("gp" if adapter.get("use_gp") else "baseline")
    assert payload["zero_shot"]["top1_acc"] > 100.0 / dm.num_classes       # synthetic features are informative


def test_tip_adapter_closed_form_matches_oracle(tmp_path):
    wl, dm = make_dm()
    cfg = make_cfg("Tip-Adapter", tmp_path)
    dm.features_val = None; dm.labels_val = None                              # keep the initial (beta, alpha)
    tr = build_trainer(cfg, dm)
    tr.train()
    f = F.normalize(dm.features_test, dim=-1); keys = F.normalize(dm.features_train, dim=-1)
    clip_w = F.normalize(F.normalize(dm.text_embeddings, dim=-1).mean(1), dim=-1)
    ref = oh.tip_logits(f, keys, oh.tip_cache_vals(dm.labels_train, dm.num_classes), 100.0 * f @ clip_w.t(), 2.0, 20.0)
    got = tr.model_inference(dm.features_test.cuda())
    assert float((got.cpu() - ref).abs().max()) < 1e-3 * float(ref.abs().max())
    assert tr._compute_final_metrics()["top1_acc"] == pytest.approx(om.compute_accuracy(ref, dm.labels_test)[0], abs=0.2)


def test_gp_adapter_training_reduces_loss(tmp_path):
    torch.manual_seed(0)
    wl, dm = make_dm()
    cfg = make_cfg("Adapter", tmp_path, use_gp=True, clip_adapter_epochs=1)
    tr = build_trainer(cfg, dm)
    tr.build_model()
    losses = []
    for ep in range(12):
        for batch in tr._epoch_batches():
            losses.append(float(tr.forward_backward(batch)["loss"]))
    assert sum(losses[-3:]) / 3 < sum(losses[:3]) / 3
    assert int(tr.engine.status.abs().max()) == 0


def test_engine_export_state_dict_round_trip():
    """engine.export_to_module -> module.state_dict() -> a fresh module (reference checkpoint flow, utils/trainer.py:347-414):
    the reloaded module reproduces the trained engine's prototypes bit for bit given the same base noise."""
    from clip_gp_b200.engine import EngineConfig, GPAdapterEngine
    from clip_gp_b200.gp_template_weigher import GaussianProcessTemplateWeighter
    wl = synth.make_workload("small"); shp = wl["shape"]
    cfg = NS(adapter=NS(gp_pca_dim=shp.d, gp_kernel_type="rbf"))
    torch.manual_seed(0)
    gpw = GaussianProcessTemplateWeighter(wl["E"], cfg, lengthscale=1.3).cuda()
    eng = GPAdapterEngine(gpw, EngineConfig(S_train=4, S_eval=4, batch_size=shp.B, shots=shp.shots, seed=5))
    for it in range(3):
        eng.train_step(wl["f_train"][it * shp.B:(it + 1) * shp.B].cuda(), wl["y_train"][it * shp.B:(it + 1) * shp.B].cuda())
    proj = torch.nn.Linear(shp.D, shp.D, bias=False).cuda()
    eng.export_to_module(gpw, proj)
    sd = {k: v.detach().cpu().clone() for k, v in gpw.state_dict().items()}
    fresh = GaussianProcessTemplateWeighter(wl["E"], cfg, lengthscale=9.9).cuda()       # different init on purpose
    missing = fresh.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    eps = torch.randn(shp.C, shp.T, 4, generator=torch.Generator().manual_seed(3)).cuda()
    with torch.no_grad():
        assert torch.equal(fresh.sample_prototypes(4, eps=eps), gpw.sample_prototypes(4, eps=eps))
    assert torch.equal(proj.weight, eng.p("W").view(shp.D, shp.D))
    assert int(fresh.variational_strategy.variational_params_initialized) == 1           # no re-initialisation of q(u) after a reload
    # the reloaded module drives a new engine to the same next step
    eng2 = GPAdapterEngine(fresh, EngineConfig(S_train=4, S_eval=4, batch_size=shp.B, shots=shp.shots, seed=5), proj.weight)
    eng2.rng_state.copy_(eng.rng_state)
    eng.skip_update = eng2.skip_update = True
    f, y = wl["f_train"][:shp.B].cuda(), wl["y_train"][:shp.B].cuda()
    assert float(eng2.train_step(f, y, use_graph=False)) == pytest.approx(float(eng.train_step(f, y, use_graph=False)), rel=1e-6)


def test_topk_accuracy_matches_torch_topk():
    from clip_gp_b200 import metrics as gm
    g = torch.Generator().manual_seed(4)
    logits = torch.randn(500, 37, generator=g); labels = torch.randint(0, 37, (500,), generator=g)
    got = gm.compute_accuracy(logits.cuda(), labels.cuda(), topk=(1, 3, 5))
    ref = []
    for k in (1, 3, 5):                                                     # utils/metrics.py:20-34
        pred = logits.topk(k, 1, True, True)[1].t()
        ref.append(float(pred.eq(labels.view(1, -1).expand_as(pred))[:k].reshape(-1).float().sum() * 100.0 / 500))
    assert got == pytest.approx(ref, abs=1e-5)                              # the reference accumulates the percentage in fp32
