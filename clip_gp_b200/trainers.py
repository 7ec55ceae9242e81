"""Trainer registry + trainers on cached frozen-CLIP features — the drop-in boundary of SURVEY.md 8b.

``TRAINER_REGISTRY`` / ``build_trainer(config, dataset_manager)`` have the reference's API
(utils/trainer_registry.py:9-42) and the trainers are registered under the reference's names ("Adapter", "TaskRes",
"CLIP-Adapter", "Tip-Adapter").  The hot path starts AFTER feature extraction, so the ``dataset_manager`` these trainers
take is a ``FeatureDataManager``: cached (features, labels) per split plus the per-class, per-template text embeddings
(what adapter.py:886-926 / tip_adapter.py:15-25 produce in memory).  CLIP encoders, image datasets and
prompt learners (CoOp / CoCoOp) are out of scope.

Trainer contract kept from utils/trainer.py:240-663: ``train()``, ``test() -> accuracy %``, ``forward_backward(batch)``,
``_compute_final_metrics()``, ``metrics.json`` schema (:599-639), attributes ``features_train/test``,
``labels_train/test``, ``zero_shot_metrics``, ``time_start``.
"""
from __future__ import annotations

import datetime
import json
import math
import os
import time
from dataclasses import dataclass
from typing import Any, Dict, Optional, Type

import torch
import torch.nn.functional as F

from . import heads, metrics, ops
from .engine import EngineConfig, GPAdapterEngine
from .gp_template_weigher import GaussianProcessTemplateWeighter


# ---------------------------------------------------------------------------------------------------- registry
class TrainerRegistry:
    """utils/trainer_registry.py:9-31."""

    def __init__(self):
        self._trainers: Dict[str, Type] = {}

    def register(self, name: str):
        def wrapper(trainer_cls):
            self._trainers[name] = trainer_cls
            return trainer_cls
        return wrapper

    def get(self, name: str):
        if name not in self._trainers:
            raise ValueError(f"Unknown trainer: {name}. Available: {list(self._trainers.keys())}")
        return self._trainers[name]

    def list_trainers(self):
        return list(self._trainers.keys())


TRAINER_REGISTRY = TrainerRegistry()


def build_trainer(config, dataset_manager):
    """utils/trainer_registry.py:37-42."""
    return TRAINER_REGISTRY.get(config.trainer_name)(config, dataset_manager)


# ---------------------------------------------------------------------------------------------------- cached features
@dataclass
class FeatureDataManager:
    """Cached frozen-CLIP features.  Image features may be un-normalised (Adapter, CLIP-Adapter) — the heads normalise."""
    text_embeddings: torch.Tensor                 # [C, T, D]
    features_train: torch.Tensor                  # [N_tr, D]
    labels_train: torch.Tensor                    # [N_tr] int64
    features_test: torch.Tensor
    labels_test: torch.Tensor
    features_val: Optional[torch.Tensor] = None
    labels_val: Optional[torch.Tensor] = None
    classnames: Optional[list] = None

    @property
    def num_classes(self) -> int:
        return int(self.text_embeddings.shape[0])

    def save(self, path: str) -> None:
        torch.save({k: getattr(self, k) for k in self.__dataclass_fields__}, path)

    @staticmethod
    def load(path: str) -> "FeatureDataManager":
        return FeatureDataManager(**torch.load(path, map_location="cpu"))


def _get(cfg, path: str, default=None):
    cur = cfg
    for part in path.split("."):
        if cur is None or not hasattr(cur, part):
            return default
        cur = getattr(cur, part)
    return cur


class BaseTrainer:
    """The part of utils/trainer.py:240-663 the cached-feature trainers need."""

    def __init__(self, config, dataset_manager: FeatureDataManager):
        self.config = config
        self.dm = dataset_manager
        if not torch.cuda.is_available():
            raise RuntimeError("clip_gp_b200 trainers need a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.max_epoch = int(_get(config, "adapter.clip_adapter_epochs", _get(config, "optim.max_epoch", 10)))   # trainer.py:256
        self.output_dir = _get(config, "output_dir", None)
        self.num_classes = dataset_manager.num_classes
        self.batch_size = int(_get(config, "dataloader.batch_size_train", 128))
        self.shots = int(_get(config, "dataset.num_shots", 1) or 1)
        d = self.device
        self.features_train, self.labels_train = dataset_manager.features_train.to(d).float(), dataset_manager.labels_train.to(d)
        self.features_test, self.labels_test = dataset_manager.features_test.to(d).float(), dataset_manager.labels_test.to(d)
        self.text_embeddings = dataset_manager.text_embeddings.to(d).float()
        self.zero_shot_metrics = None
        self.time_start = time.time()
        self.epoch = 0
        self.model = None

    # -- evaluation ------------------------------------------------------------------------------------------------
    def model_inference(self, features: torch.Tensor) -> torch.Tensor:
        return self.model(features)

    @torch.no_grad()
    def _compute_final_metrics(self) -> Dict[str, Any]:
        if self.model is not None:
            self.model.eval()
        logits = self.model_inference(self.features_test)
        res = metrics.evaluate_calibration(logits, self.labels_test)
        return {"top1_acc": float(res["top1_acc"]), "ece": float(res["ece"]), "aece": float(res["aece"]),
                "calibration": res["calibration"], "adaptive_calibration": res["adaptive_calibration"]}

    def test(self, split=None) -> float:
        m = self._compute_final_metrics()
        print("=> result")
        print(f"* total: {int(self.labels_test.numel()):,}")
        print(f"* accuracy: {m['top1_acc']:.1f}%")
        print(f"* ECE: {m['ece']:.2f}%")
        print(f"* AECE: {m['aece']:.2f}%")
        self._write_run_summary_json(m, start_time=self.time_start)
        return m["top1_acc"]

    def zero_shot(self) -> Dict[str, Any]:
        """Zero-shot baseline on the test features: uniform template mean (adapter.py:589-611)."""
        with torch.no_grad():
            protos = ops.row_normalize(self.text_embeddings.mean(dim=1))
            logits = ops.matmul_nt(ops.row_normalize(self.features_test), protos, 100.0)
            res = metrics.evaluate_calibration(logits, self.labels_test)
        self.zero_shot_metrics = {"top1_acc": res["top1_acc"], "ece": res["ece"], "aece": res["aece"],
                                  "calibration": res["calibration"], "adaptive_calibration": res["adaptive_calibration"]}
        print("Zero-Shot accuracy on test: " + str(round(res["top1_acc"], 2)))
        return self.zero_shot_metrics

    def _write_run_summary_json(self, metrics_: Dict[str, Any], start_time: float) -> None:
        """utils/trainer.py:599-639 (same keys; consumed by scripts/aggregate_results.py)."""
        if not self.output_dir:
            return
        os.makedirs(self.output_dir, exist_ok=True)
        method = "gp" if bool(_get(self.config, "adapter.use_gp", False)) else "baseline"
        payload = {
            "timestamp": datetime.datetime.now().isoformat(),
            "dataset": _get(self.config, "dataset.name", "cached"),
            "shots": int(self.shots),
            "seed": int(_get(self.config, "seed", 0) or 0),
            "method": method,
            "backbone": _get(self.config, "model.backbone_name", "cached-features"),
            "zero_shot": self.zero_shot_metrics,
            "metrics": metrics_,
            "config": {},
            "output_dir": str(self.output_dir),
            "train_time_s": float(max(0.0, time.time() - start_time)),
        }
        with open(os.path.join(self.output_dir, "metrics.json"), "w") as f:
            json.dump(payload, f, indent=2)

    def _epoch_batches(self, generator: Optional[torch.Generator] = None):
        """Device-side shuffle + drop_last batches (adapter.py:729-746, utils/data_manager.py:79)."""
        n = self.features_train.shape[0]
        perm = torch.randperm(n, device=self.device, generator=generator)
        bs = min(self.batch_size, n)
        nb = n // bs if n >= bs else 1
        for i in range(nb):
            idx = perm[i * bs:(i + 1) * bs]
            yield self.features_train[idx], self.labels_train[idx]


# ---------------------------------------------------------------------------------------------------- Adapter
@TRAINER_REGISTRY.register("Adapter")
class AdapterTrainer(BaseTrainer):
    """trainers/adapter.py:262-884 on cached features.  With USE_GP the step runs in the fused engine (engine.py)."""

    def build_model(self):
        cfg = self.config
        self.model = heads.AdapterHead(cfg, self.text_embeddings).to(self.device)
        self.use_gp = self.model.gp_weighter is not None
        a = cfg.adapter
        if self.use_gp:
            ecfg = EngineConfig(S_train=int(getattr(a, "gp_num_mc_samples_train", 1) or 1), S_eval=int(getattr(a, "gp_num_mc_samples_eval", 1) or 1),
                                batch_size=min(self.batch_size, self.features_train.shape[0]), logit_scale=float(self.model.logit_scale.exp()),
                                gp_beta=float(getattr(a, "gp_beta", 1.0)), l2_lambda=float(getattr(a, "l2_lambda", 0.5)), shots=self.shots,
                                lr=float(_get(cfg, "optim.lr", 0.01)), gp_lr=float(getattr(a, "gp_lr", 1e-3)),
                                weight_decay=float(_get(cfg, "optim.weight_decay", 0.0)), loss_mode="per_sample" if int(getattr(a, "gp_num_mc_samples_train", 1) or 1) > 1 else "logit_mean",
                                train_visual_proj=not bool(getattr(a, "freeze_visual_proj", False)), seed=int(_get(cfg, "seed", 0) or 0),
                                # GEMMs of the step: split-bf16 tensor-core path by default (fp32-grade products; the reference's own
                                # GPU path is TF32, adapter.py:23); "fp32" = FFMA comparator, "bf16" = stated tolerance
                                precision=str(getattr(a, "clipgp_precision", "bf16x3")))
            self.engine = GPAdapterEngine(self.model.gp_weighter, ecfg, self.model.visual_proj.weight)
        else:
            params = [p for p in self.model.visual_proj.parameters()]
            self.optim = torch.optim.AdamW(params, lr=float(_get(cfg, "optim.lr", 0.01)), weight_decay=float(_get(cfg, "optim.weight_decay", 0.0)))

    def forward_backward(self, batch):
        feats, labels = batch
        if self.use_gp:
            loss = self.engine.train_step(feats, labels)
            return {"loss": loss}
        a = self.config.adapter
        self.model.train()
        loss = self.model.compute_loss(feats, labels, 1, 0.0, float(getattr(a, "l2_lambda", 0.5)), self.shots)
        self.optim.zero_grad()
        loss.backward()
        self.optim.step()
        return {"loss": loss.detach()}

    def model_inference(self, features):
        if self.use_gp:
            return self.engine.eval_logits(features)
        return self.model(features)

    def train(self):
        self.time_start = time.time()
        self.build_model()
        self.zero_shot()
        last = None
        cfg = self.config
        sched = str(_get(cfg, "optim.lr_scheduler", "cosine") or "constant").lower()
        base_lr, base_gp_lr = float(_get(cfg, "optim.lr", 0.01)), float(getattr(cfg.adapter, "gp_lr", 1e-3))
        for self.epoch in range(self.max_epoch):
            if self.use_gp and sched == "cosine":
                # CosineAnnealingLR(T_max=max_epoch) stepped once per epoch (utils/optimization.py:232-238, adapter.py:1054-1056);
                # the rates live in device memory, so the captured step graph is not re-captured
                self.engine.cosine_lr(self.epoch, self.max_epoch, base_lr, base_gp_lr)
            for batch in self._epoch_batches():
                last = self.forward_backward(batch)
            if last is not None and ((self.epoch + 1) % 10 == 0 or self.epoch == 0):
                print(f"epoch [{self.epoch + 1}/{self.max_epoch}] loss {float(last['loss']):.4f}")
        if self.use_gp:
            self.engine.export_to_module(self.model.gp_weighter, self.model.visual_proj)
        return self.test()


# ---------------------------------------------------------------------------------------------------- GP-initialised heads
class _GPInitMixin:
    def _maybe_gp_pretrain(self, tag: str) -> Optional[torch.Tensor]:
        """taskres.py:209-293 / clip_adapter.py:235-294 / tip_adapter.py:90-160: optional GP pre-training on the few-shot
        features, then normalize(mean_s prototypes) as the new class weights.  Any failure disables GP, as in the reference."""
        cfg = self.config
        if not bool(_get(cfg, "adapter.use_gp", False)):
            return None
        try:
            gpw = GaussianProcessTemplateWeighter(text_embeddings=F.normalize(self.text_embeddings, dim=-1), cfg=cfg).to(self.device)
            self.gp_weighter = gpw
            a = cfg.adapter
            heads.gp_pretrain(gpw, ops.row_normalize(self.features_train), self.labels_train, epochs=int(_get(cfg, "optim.max_epoch", 50)),
                              gp_lr=float(getattr(a, "gp_lr", 1e-3)), beta_kl=float(getattr(a, "gp_beta", 1e-3)),
                              num_samples=int(getattr(a, "gp_num_mc_samples_train", 30) or 1),
                              weight_decay=float(_get(cfg, "optim.weight_decay", 0.0)))
            protos = gpw.mean_prototypes(int(getattr(a, "gp_num_mc_samples_eval", 100) or 1))
            print(f"[{tag}] Using trained GP-based template weighter for prototypes.")
            return protos
        except Exception as e:                                              # reference behaviour: warn and continue without GP
            print(f"[{tag}][WARN] GP weighting failed ({e}); continuing without GP.")
            self.gp_weighter = None
            return None


@TRAINER_REGISTRY.register("TaskRes")
class TaskResTrainer(BaseTrainer, _GPInitMixin):
    """trainers/taskres.py:126-439 on cached features."""

    def build_model(self):
        base = self.text_embeddings.mean(dim=1)                              # taskres.py:88-92
        self.model = heads.TaskResHead(self.config, base).to(self.device)
        a = self.config.adapter
        self.optim = torch.optim.AdamW([self.model.text_feature_residuals], lr=float(getattr(a, "taskres_lr", _get(self.config, "optim.lr", 1e-3))),
                                       weight_decay=float(_get(self.config, "optim.weight_decay", 0.0)))
        self.max_epoch = int(getattr(a, "taskres_epochs", _get(self.config, "optim.max_epoch", 100)))

    def forward_backward(self, batch):
        feats, labels = batch
        self.model.train()
        loss = ops.cross_entropy(self.model(feats), labels)                 # taskres.py:189-195
        self.optim.zero_grad()
        loss.backward()
        self.optim.step()
        return {"loss": loss.detach()}

    def train(self):
        self.time_start = time.time()
        self.build_model()
        self.zero_shot()
        protos = self._maybe_gp_pretrain("TaskRes")
        if protos is not None:
            with torch.no_grad():
                self.model.base_text_features.copy_(protos)                  # taskres.py:286-289
            self.model.gp_weighter = self.gp_weighter
        for self.epoch in range(self.max_epoch):
            for batch in self._epoch_batches():
                self.forward_backward(batch)
        return self.test()


@TRAINER_REGISTRY.register("CLIP-Adapter")
class ClipAdapterTrainer(BaseTrainer, _GPInitMixin):
    """trainers/clip_adapter.py:114-386 on cached features."""

    def build_model(self):
        clip_w = F.normalize(F.normalize(self.text_embeddings, dim=-1).mean(dim=1), dim=-1).t().contiguous()   # _get_clip_weights: [D,K]
        self.model = heads.ClipAdapterHead(self.config, clip_w).to(self.device)
        self.optim = torch.optim.AdamW(self.model.adapter.parameters(), lr=float(_get(self.config, "optim.lr", 1e-3)),
                                       weight_decay=float(_get(self.config, "optim.weight_decay", 0.0)))

    def forward_backward(self, batch):
        feats, labels = batch
        self.model.train()
        loss = ops.cross_entropy(self.model.logits_from_features(feats, training=True), labels)     # clip_adapter.py:218-228
        self.optim.zero_grad()
        loss.backward()
        self.optim.step()
        return {"loss": loss.detach()}

    def train(self):
        self.time_start = time.time()
        self.build_model()
        self.zero_shot()
        protos = self._maybe_gp_pretrain("CLIP-Adapter")
        if protos is not None:
            with torch.no_grad():
                self.model.clip_weights.copy_(protos.t())                    # clip_adapter.py:284-290
            self.model.gp_weighter = self.gp_weighter
        for self.epoch in range(self.max_epoch):
            for batch in self._epoch_batches():
                self.forward_backward(batch)
        return self.test()


@TRAINER_REGISTRY.register("Tip-Adapter")
class TipAdapterTrainer(BaseTrainer, _GPInitMixin):
    """trainers/tip_adapter.py:28-398 on cached features (``tip_adapter_trainable`` selects Tip-Adapter-F)."""

    def build_model(self):
        self.clip_weights = F.normalize(F.normalize(self.text_embeddings, dim=-1).mean(dim=1), dim=-1)    # [K,D] (transpose of the reference's [D,K])
        self.gp_weighter = None

    def _clip_logits(self, feats_hat):
        if self.gp_weighter is not None:
            S = int(_get(self.config, "adapter.gp_num_mc_samples_eval", 100) or 1)
            with torch.no_grad():
                pm = self.gp_weighter.collapsed_prototypes(max(1, S))       # mean_s 100 f.p_hat_s == 100 f.(mean_s p_hat_s) (tip_adapter.py:211-217)
            return ops.matmul_nt(feats_hat, pm, 100.0)
        return ops.matmul_nt(feats_hat, self.clip_weights, 100.0)           # tip_adapter.py:219

    def model_inference(self, features):
        f_hat = ops.row_normalize(features)
        return ops.tip_logits(f_hat, self.cache_keys, self.cache_labels, self._clip_logits(f_hat), self.best_beta, self.best_alpha, self.num_classes)

    def train(self):
        self.time_start = time.time()
        self.build_model()
        a = self.config.adapter
        protos = self._maybe_gp_pretrain("Tip-Adapter")
        if protos is not None:
            self.clip_weights = protos                                       # tip_adapter.py:152-157
        self.zero_shot()
        f_tr = ops.row_normalize(self.features_train)
        self.cache_keys, self.cache_labels = heads.tip_build_cache(f_tr.detach(), self.labels_train)
        self.best_beta, self.best_alpha = float(getattr(a, "tip_adapter_init_beta", 2.0)), float(getattr(a, "tip_adapter_init_alpha", 20.0))
        prec = str(getattr(a, "clipgp_precision", "bf16x3"))                 # GEMM path of the affinity and key-gradient contractions
        if bool(getattr(a, "tip_adapter_trainable", False)):                 # Tip-Adapter-F, tip_adapter.py:227-296
            epochs = int(getattr(a, "tip_adapter_epochs", 20))
            nb = max(1, self.features_train.shape[0] // min(self.batch_size, self.features_train.shape[0]))
            # the step runs in the fused engine (tip_engine.py): AdamW(lr, eps) with torch's default weight decay and the per-step
            # CosineAnnealingLR over epochs * len(loader) steps (tip_adapter.py:231-235).  The reference's "best epoch" bookkeeping
            # keeps a reference to the live state_dict (:289-292), i.e. the final weights are always the last epoch's: same here.
            from .tip_engine import TipAdapterEngine
            eng = TipAdapterEngine(self.cache_keys, self.cache_labels, self.num_classes, min(self.batch_size, self.features_train.shape[0]),
                                   self.best_beta, self.best_alpha, lr=float(getattr(a, "tip_adapter_lr", 1e-3)),
                                   eps=float(getattr(a, "tip_adapter_eps", 1e-4)), total_steps=epochs * nb, precision=prec)
            with torch.no_grad():
                for self.epoch in range(epochs):
                    for feats, labels in self._epoch_batches():
                        f_hat = ops.row_normalize(feats)
                        eng.train_step(f_hat, self._clip_logits(f_hat), labels)
            self.tip_engine = eng
            self.cache_keys = eng.keys
        if self.dm.features_val is not None:                                  # tip_adapter.py:298-304
            fv = ops.row_normalize(self.dm.features_val.to(self.device).float())
            yv = self.dm.labels_val.to(self.device)
            self.best_beta, self.best_alpha, _ = heads.tip_search(fv, yv, self.cache_keys, self.cache_labels, self._clip_logits(fv), self.num_classes,
                                                                  self.best_beta, self.best_alpha, precision=prec)
        self._tip_adapter_best_beta, self._tip_adapter_best_alpha = self.best_beta, self.best_alpha
        return self.test()
