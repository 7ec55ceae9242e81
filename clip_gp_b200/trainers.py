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

from . import heads, host_optim, metrics, ops
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
CACHE_MAGIC = "clipgp-feature-cache"
CACHE_VERSION = 1


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

    # -- on-disk cache: what adapter.py:886-926 / tip_adapter.py:15-25 recompute from images on every run ------
    def save(self, path: str, meta: Optional[Dict[str, Any]] = None) -> None:
        """Versioned single-file cache: {"magic", "version", "meta", "tensors": {name: fp32 / int64 CPU tensor}, "classnames"}.
        `meta` should carry what identifies the extraction (backbone, dataset, shots, seed, template list) so that a stale cache
        is detected by the caller; it is stored verbatim."""
        tensors = {}
        for k in ("text_embeddings", "features_train", "labels_train", "features_test", "labels_test", "features_val", "labels_val"):
            v = getattr(self, k)
            if v is not None:
                v = v.detach().cpu().contiguous()
                tensors[k] = v.to(torch.int64) if k.startswith("labels") else v.to(torch.float32)
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        tmp = path + ".tmp"
        torch.save({"magic": CACHE_MAGIC, "version": CACHE_VERSION, "meta": dict(meta or {}), "tensors": tensors,
                    "classnames": list(self.classnames) if self.classnames is not None else None}, tmp)
        os.replace(tmp, path)                                   # atomic: a killed run never leaves a half-written cache behind

    @staticmethod
    def load(path: str, expect_meta: Optional[Dict[str, Any]] = None) -> "FeatureDataManager":
        blob = torch.load(path, map_location="cpu", weights_only=True)
        if not isinstance(blob, dict) or blob.get("magic") != CACHE_MAGIC:
            raise ValueError(f"{path}: not a clipgp feature cache")
        if int(blob.get("version", -1)) != CACHE_VERSION:
            raise ValueError(f"{path}: feature-cache version {blob.get('version')} != {CACHE_VERSION}")
        for k, v in (expect_meta or {}).items():
            if blob["meta"].get(k) != v:
                raise ValueError(f"{path}: cache was extracted with {k}={blob['meta'].get(k)!r}, this run needs {v!r}")
        t = blob["tensors"]
        N = {k: t[k].shape[0] for k in t}
        for a, b in (("features_train", "labels_train"), ("features_test", "labels_test"), ("features_val", "labels_val")):
            if (a in t) != (b in t) or (a in t and N[a] != N[b]):
                raise ValueError(f"{path}: {a} / {b} are inconsistent")
        dm = FeatureDataManager(**{k: t.get(k) for k in ("text_embeddings", "features_train", "labels_train", "features_test",
                                                         "labels_test", "features_val", "labels_val")},
                                classnames=blob.get("classnames"))
        dm.meta = blob["meta"]
        return dm

    # -- adaptor at the registry boundary: train.py:89 hands `build_trainer` the reference's image DataManager -----------------
    @staticmethod
    def from_reference(dataset_manager: Any, text_embeddings: torch.Tensor, encode_image=None, device="cpu") -> "FeatureDataManager":
        """Build the cached-feature manager from the reference's ``DataManager`` (utils/data_manager.py: ``train_loader_x`` /
        ``val_loader`` / ``test_loader`` yielding {"img", "label"}, ``dataset.classnames``) — the loop of adapter.py:886-926 /
        tip_adapter.py:15-25.  ``encode_image`` is the frozen CLIP visual encoder (``clip_model.visual`` /
        ``clip_model.encode_image``); None means the loaders already yield feature vectors."""
        def extract(loader):
            if loader is None:
                return None, None
            if isinstance(loader, torch.utils.data.DataLoader) and loader.drop_last:
                # the few-shot train loader drops its last partial batch (utils/data_manager.py:79); feature extraction must not
                # lose samples (adapter.py:895-903)
                loader = torch.utils.data.DataLoader(loader.dataset, batch_size=loader.batch_size, shuffle=False,
                                                     num_workers=loader.num_workers, drop_last=False)
            fs, ys = [], []
            with torch.no_grad():
                for batch in loader:
                    x = batch["img"].to(device)
                    fs.append((encode_image(x) if encode_image is not None else x).float().cpu())
                    ys.append(torch.as_tensor(batch["label"]).to(torch.int64).cpu())
            return torch.cat(fs, 0), torch.cat(ys, 0)
        f_tr, y_tr = extract(dataset_manager.train_loader_x)
        f_te, y_te = extract(dataset_manager.test_loader)
        f_va, y_va = extract(getattr(dataset_manager, "val_loader", None))
        names = getattr(getattr(dataset_manager, "dataset", None), "classnames", None)
        return FeatureDataManager(text_embeddings.detach().float().cpu(), f_tr, y_tr, f_te, y_te, f_va, y_va,
                                  list(names) if names is not None else None)


def as_feature_manager(dataset_manager: Any) -> FeatureDataManager:
    """What every trainer constructor does with its ``dataset_manager`` argument: a FeatureDataManager passes through, a path is
    loaded from the cache file, and a reference-style DataManager that carries ``text_embeddings`` (and feature-yielding loaders,
    or an ``encode_image`` attribute) is converted."""
    if isinstance(dataset_manager, FeatureDataManager):
        return dataset_manager
    if isinstance(dataset_manager, (str, os.PathLike)):
        return FeatureDataManager.load(os.fspath(dataset_manager))
    if hasattr(dataset_manager, "train_loader_x") and hasattr(dataset_manager, "text_embeddings"):
        return FeatureDataManager.from_reference(dataset_manager, dataset_manager.text_embeddings,
                                                 getattr(dataset_manager, "encode_image", None))
    raise TypeError("clip_gp_b200 trainers take a FeatureDataManager, a feature-cache path, or a reference DataManager with "
                    "`text_embeddings` attached (FeatureDataManager.from_reference)")


def _get(cfg, path: str, default=None):
    cur = cfg
    for part in path.split("."):
        if cur is None or not hasattr(cur, part):
            return default
        cur = getattr(cur, part)
    return cur


class BaseTrainer:
    """The part of utils/trainer.py:240-663 the cached-feature trainers need."""

    def __init__(self, config, dataset_manager):
        self.config = config
        self.dm = dataset_manager = as_feature_manager(dataset_manager)
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
        """Zero-shot baseline on the test features with `_get_clip_weights` (utils/trainer.py:222-237): per-class mean of the
        unit-normalised template embeddings, re-normalised; logits 100 f_hat . w (taskres.py:203-206, tip_adapter.py:219)."""
        with torch.no_grad():
            protos = F.normalize(F.normalize(self.text_embeddings, dim=-1).mean(dim=1), dim=-1)
            logits = ops.matmul_nt(ops.row_normalize(self.features_test), protos.contiguous(), 100.0, heads._precision(self.config))
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

    def _num_batches(self) -> int:
        n = self.features_train.shape[0]
        bs = min(self.batch_size, n)
        return n // bs if n >= bs else 1

    def _epoch_batches(self, generator: Optional[torch.Generator] = None):
        """Device-side shuffle + drop_last batches (adapter.py:729-746, utils/data_manager.py:79).  `self.batch_plan`, when set
        (an iterable of (features, labels) per step), replaces the shuffle: parity tests replay the reference's batch order."""
        plan = getattr(self, "batch_plan", None)
        if plan is not None:
            nb = self._num_batches()
            for _ in range(nb):
                f, y = next(plan)
                yield f.to(self.device).float(), y.to(self.device)
            return
        n = self.features_train.shape[0]
        perm = torch.randperm(n, device=self.device, generator=generator)
        bs = min(self.batch_size, n)
        for i in range(self._num_batches()):
            idx = perm[i * bs:(i + 1) * bs]
            yield self.features_train[idx], self.labels_train[idx]

    def _fit_epochs(self):
        """BaseTrainer.train() loop of the reference (utils/trainer.py:648-659): run_epoch, then scheduler.step() once per epoch
        (after_epoch, :466-470)."""
        for self.epoch in range(self.max_epoch):
            last = None
            for batch in self._epoch_batches():
                last = self.forward_backward(batch)
            if getattr(self, "sched", None) is not None:
                self.sched.step()
            if last is not None and ((self.epoch + 1) % 10 == 0 or self.epoch == 0):
                print(f"epoch [{self.epoch + 1}/{self.max_epoch}] loss {float(last['loss']):.4f}")


# ---------------------------------------------------------------------------------------------------- Adapter
@TRAINER_REGISTRY.register("Adapter")
class AdapterTrainer(BaseTrainer):
    """trainers/adapter.py:262-884 on cached features.  With USE_GP the step runs in the fused engine (engine.py)."""

    def build_model(self):
        cfg = self.config
        self.model = heads.AdapterHead(cfg, self.text_embeddings).to(self.device)
        self.use_gp = self.model.gp_weighter is not None
        self.optim = self.sched = None
        if self.use_gp:
            self.build_engine()
        else:
            # baseline: visual projection only (uniform template weights); optimizer / scheduler as adapter.py:312-326
            params = [p for p in self.model.visual_proj.parameters()]
            if bool(getattr(cfg.adapter, "freeze_visual_proj", False)):
                for p in params:
                    p.requires_grad_(False)
                params = []
            if params:
                self.optim = host_optim.build_optimizer(params, _get(cfg, "optim"))
                self.sched = host_optim.build_lr_scheduler(self.optim, _get(cfg, "optim"))

    def build_engine(self):
        """(Re)create the fused engine from the current module state (adapter.py:290-311: two AdamW groups, lr / gp_lr)."""
        cfg, a = self.config, self.config.adapter
        opt_name = str(_get(cfg, "optim.name", "adamw")).lower()
        wd = float(_get(cfg, "optim.weight_decay", 0.0))
        if opt_name not in ("adamw", "adam"):
            raise NotImplementedError(f"the fused GP-Adapter engine implements Adam / AdamW (optim.name={opt_name!r}); "
                                      "the reference's GP config uses adamw (configs/trainers/default.yaml)")
        if wd != 0.0:
            raise NotImplementedError("optim.weight_decay > 0 with use_gp: the reference decays the hook-frozen template rows of the "
                                      "inducing points too (gp_template_weigher.py:72-79); the engine keeps them fixed")
        S_tr = int(getattr(a, "gp_num_mc_samples_train", 1) or 1)
        ecfg = EngineConfig(S_train=S_tr, S_eval=int(getattr(a, "gp_num_mc_samples_eval", 1) or 1),
                            batch_size=min(self.batch_size, self.features_train.shape[0]), logit_scale=float(self.model.logit_scale.exp()),
                            gp_beta=float(getattr(a, "gp_beta", 1.0)), l2_lambda=float(getattr(a, "l2_lambda", 0.5)), shots=self.shots,
                            lr=float(_get(cfg, "optim.lr", 0.01)), gp_lr=float(getattr(a, "gp_lr", 1e-3)), weight_decay=0.0,
                            betas=tuple(_get(cfg, "optim.betas", (0.9, 0.999))), adam_eps=float(_get(cfg, "optim.eps", 1e-8)),
                            loss_mode="per_sample" if S_tr > 1 else "logit_mean",                      # adapter.py:401 vs :444-452
                            train_visual_proj=not bool(getattr(a, "freeze_visual_proj", False)), seed=int(_get(cfg, "seed", 0) or 0),
                            # GEMMs of the step: TF32 on the fp32 tensors in place by default (the reference's own GPU arithmetic,
                            # adapter.py:23); "bf16x3" = split-bf16 (fp32-grade products), "fp32" = FFMA comparator, "bf16" = stated tolerance
                            precision=self._precision())
        try:
            self.engine = GPAdapterEngine(self.model.gp_weighter, ecfg, self.model.visual_proj.weight)
        except ValueError as e:
            if ecfg.precision != "tf32" or "multiples of 4" not in str(e) or getattr(a, "clipgp_precision", None) is not None:
                raise
            ecfg.precision = "bf16x3"                        # row pitch not a multiple of 16 bytes (e.g. odd S*C): split-bf16 has no such rule
            self.engine = GPAdapterEngine(self.model.gp_weighter, ecfg, self.model.visual_proj.weight)
        return self.engine

    def _precision(self) -> str:
        return str(getattr(self.config.adapter, "clipgp_precision", None) or "tf32")

    def forward_backward(self, batch):
        feats, labels = batch
        if self.use_gp:
            loss = self.engine.train_step(feats, labels)
            return {"loss": loss}
        if self.optim is None:
            return {"loss": torch.zeros((), device=self.device)}
        a = self.config.adapter
        self.model.train()
        loss = self.model.compute_loss(feats, labels, 1, 0.0, float(getattr(a, "l2_lambda", 0.5)), self.shots)
        self.optim.zero_grad()
        loss.backward()
        self.optim.step()
        return {"loss": loss.detach()}

    def model_inference(self, features):
        """Logits [N,C] (the reference API).  `test()` does not come through here for the GP model: it uses the fused
        projection + logits + calibration GEMM and never materialises them."""
        if self.use_gp:
            return self.engine.eval_logits(features)
        return self.model(features)

    @torch.no_grad()
    def _compute_final_metrics(self) -> Dict[str, Any]:
        if not self.use_gp:
            return super()._compute_final_metrics()
        res = self.engine.evaluate(self.features_test, self.labels_test, precision=self.engine.cfg.precision)
        return {"top1_acc": float(res["top1_acc"]), "ece": float(res["ece"]), "aece": float(res["aece"]),
                "calibration": res["calibration"], "adaptive_calibration": res["adaptive_calibration"]}

    def zero_shot(self) -> Dict[str, Any]:
        """adapter.py:589-611: the model's own forward on the test features BEFORE training (for the GP model that is the
        MC-averaged GP prediction at its initial state, not the uniform template mean)."""
        m = self._compute_final_metrics()
        self.zero_shot_metrics = m
        print("Zero-Shot accuracy on test: " + str(round(m["top1_acc"], 2)))
        return m

    def train(self):
        self.time_start = time.time()
        if self.model is None:
            self.build_model()
        self.zero_shot()
        cfg = self.config
        if not self.use_gp:
            self._fit_epochs()
            return self.test()
        sched = str(_get(cfg, "optim.lr_scheduler", "cosine") or "constant").lower()
        if sched not in ("cosine", "constant"):
            raise NotImplementedError(f"fused GP-Adapter engine: lr_scheduler={sched!r} (cosine | constant)")
        base_lr, base_gp_lr = float(_get(cfg, "optim.lr", 0.01)), float(getattr(cfg.adapter, "gp_lr", 1e-3))
        t_max = int(_get(cfg, "optim.max_epoch", self.max_epoch))           # build_lr_scheduler: T_max = OPTIM.MAX_EPOCH, while the
        last = None                                                          # loop runs clip_adapter_epochs epochs (utils/trainer.py:256)
        for self.epoch in range(self.max_epoch):
            if sched == "cosine":
                # CosineAnnealingLR(T_max) stepped once per epoch (utils/optimization.py:232-238, utils/trainer.py:466-470); the
                # rates live in device memory, so the captured step graph is not re-captured
                self.engine.cosine_lr(self.epoch, t_max, base_lr, base_gp_lr, eta_min=float(_get(cfg, "optim.eta_min", 0.0)))
            for batch in self._epoch_batches():
                last = self.forward_backward(batch)
            if last is not None and ((self.epoch + 1) % 10 == 0 or self.epoch == 0):
                print(f"epoch [{self.epoch + 1}/{self.max_epoch}] loss {float(last['loss']):.4f}")
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
            gpw = getattr(self, "gp_weighter", None)
            if gpw is None:
                # taskres.py:240-245 and tip_adapter.py:96-102 unit-normalise the per-template text features; clip_adapter.py:240
                # hands the weighter the raw `encode_text` outputs
                E = self.text_embeddings if tag == "CLIP-Adapter" else F.normalize(self.text_embeddings, dim=-1)
                gpw = GaussianProcessTemplateWeighter(text_embeddings=E, cfg=cfg, rng="philox",
                                                      seed=int(_get(cfg, "seed", 0) or 0)).to(self.device)
                self.gp_weighter = gpw
            a = cfg.adapter
            heads.gp_pretrain(gpw, ops.row_normalize(self.features_train), self.labels_train, epochs=int(_get(cfg, "optim.max_epoch", 50)),
                              gp_lr=float(getattr(a, "gp_lr", 1e-3)), beta_kl=float(getattr(a, "gp_beta", 1e-3)),
                              num_samples=int(getattr(a, "gp_num_mc_samples_train", 30) or 1),
                              weight_decay=float(_get(cfg, "optim.weight_decay", 0.0)),
                              precision=str(getattr(a, "clipgp_precision", None) or "auto"), seed=int(_get(cfg, "seed", 0) or 0))
            protos = gpw.mean_prototypes(int(getattr(a, "gp_num_mc_samples_eval", 100) or 1), eps=getattr(self, "eval_eps", None))
            print(f"[{tag}] Using trained GP-based template weighter for prototypes.")
            return protos
        except NotImplementedError:
            raise
        except Exception as e:                                              # reference behaviour: warn and continue without GP
            print(f"[{tag}][WARN] GP weighting failed ({e}); continuing without GP.")
            self.gp_weighter = None
            return None


def _tmp_optim(cfg, name, lr, max_epoch):
    """The `_TmpOptim` record the reference fills for the head-specific optimizer (taskres.py:158-171, clip_adapter.py:152-165):
    name / lr / max_epoch from the adapter section, everything else from OPTIM."""
    o = _get(cfg, "optim")
    t = type("TmpOptim", (), {})()
    t.name, t.lr, t.max_epoch = name, float(lr), int(max_epoch)
    t.lr_scheduler = _get(o, "lr_scheduler", "cosine")
    t.weight_decay = float(_get(o, "weight_decay", 0.0))
    t.momentum = _get(o, "momentum", 0.9)
    t.betas = _get(o, "betas", (0.9, 0.999))
    return t


@TRAINER_REGISTRY.register("TaskRes")
class TaskResTrainer(BaseTrainer, _GPInitMixin):
    """trainers/taskres.py:126-439 on cached features."""

    def build_model(self):
        cfg, a = self.config, self.config.adapter
        base = self.text_embeddings.mean(dim=1)                              # taskres.py:88-92
        self.model = heads.TaskResHead(cfg, base).to(self.device)
        # taskres.py:158-173: optimizer from taskres_optimizer / taskres_lr, CosineAnnealingLR(T_max = taskres_epochs) per epoch.
        # The epoch LOOP length is BaseTrainer's max_epoch = clip_adapter_epochs (utils/trainer.py:256), as in the reference.
        tmp = _tmp_optim(cfg, getattr(a, "taskres_optimizer", _get(cfg, "optim.name", "adam")), getattr(a, "taskres_lr", _get(cfg, "optim.lr", 1e-3)),
                         getattr(a, "taskres_epochs", _get(cfg, "optim.max_epoch", 100)))
        self.optim = host_optim.build_optimizer([self.model.text_feature_residuals], tmp)
        self.sched = host_optim.build_lr_scheduler(self.optim, tmp)

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
        if self.model is None:
            self.build_model()
        self.zero_shot()
        protos = self._maybe_gp_pretrain("TaskRes")
        if protos is not None:
            with torch.no_grad():
                self.model.base_text_features.copy_(protos)                  # taskres.py:281-289
            self.model.gp_weighter = self.gp_weighter
        self._fit_epochs()
        return self.test()


@TRAINER_REGISTRY.register("CLIP-Adapter")
class ClipAdapterTrainer(BaseTrainer, _GPInitMixin):
    """trainers/clip_adapter.py:114-386 on cached features."""

    def build_model(self):
        cfg, a = self.config, self.config.adapter
        clip_w = F.normalize(F.normalize(self.text_embeddings, dim=-1).mean(dim=1), dim=-1).t().contiguous()   # _get_clip_weights: [D,K]
        self.model = heads.ClipAdapterHead(cfg, clip_w).to(self.device)
        # clip_adapter.py:141-167: lr = clip_adapter_lr, optimizer = clip_adapter_optimizer, T_max = clip_adapter_epochs
        lr = float(getattr(a, "clip_adapter_lr", _get(cfg, "optim.lr", 1e-3)))
        tmp = _tmp_optim(cfg, getattr(a, "clip_adapter_optimizer", _get(cfg, "optim.name", "adam")), lr,
                         getattr(a, "clip_adapter_epochs", _get(cfg, "optim.max_epoch", 100)))
        self.optim = host_optim.build_optimizer([{"params": list(self.model.adapter.parameters()), "lr": lr, "weight_decay": tmp.weight_decay}], tmp)
        self.sched = host_optim.build_lr_scheduler(self.optim, tmp)

    def forward_backward(self, batch):
        feats, labels = batch
        self.model.train()
        loss = ops.cross_entropy(self.model.logits_from_features(feats, training=True), labels)     # clip_adapter.py:218-228
        self.optim.zero_grad()
        loss.backward()
        self.optim.step()
        return {"loss": loss.detach()}

    def zero_shot(self) -> Dict[str, Any]:
        """clip_adapter.py:180-183: the model's own logits on the test features before training."""
        self.zero_shot_metrics = self._compute_final_metrics()
        print("Zero-Shot accuracy on test: " + str(round(self.zero_shot_metrics["top1_acc"], 2)))
        return self.zero_shot_metrics

    def train(self):
        self.time_start = time.time()
        if self.model is None:
            self.build_model()
        self.zero_shot()
        protos = self._maybe_gp_pretrain("CLIP-Adapter")
        if protos is not None:
            with torch.no_grad():
                self.model.clip_weights.copy_(protos.t())                    # clip_adapter.py:284-290
            self.model.gp_weighter = self.gp_weighter
        self._fit_epochs()
        return self.test()


@TRAINER_REGISTRY.register("Tip-Adapter")
class TipAdapterTrainer(BaseTrainer, _GPInitMixin):
    """trainers/tip_adapter.py:28-398 on cached features (``tip_adapter_trainable`` selects Tip-Adapter-F)."""

    def build_model(self):
        self.clip_weights = F.normalize(F.normalize(self.text_embeddings, dim=-1).mean(dim=1), dim=-1)    # [K,D] (transpose of the reference's [D,K])
        self.gp_weighter = getattr(self, "gp_weighter", None)
        self.model = None

    def zero_shot(self) -> Dict[str, Any]:
        """tip_adapter.py:206-221: CLIP logits on the test features (MC-averaged GP logits once the weighter is trained)."""
        with torch.no_grad():
            res = metrics.evaluate_calibration(self._clip_logits(ops.row_normalize(self.features_test)), self.labels_test)
        self.zero_shot_metrics = {"top1_acc": res["top1_acc"], "ece": res["ece"], "aece": res["aece"],
                                  "calibration": res["calibration"], "adaptive_calibration": res["adaptive_calibration"]}
        print("Zero-Shot accuracy on test: " + str(round(res["top1_acc"], 2)))
        return self.zero_shot_metrics

    def _clip_logits(self, feats_hat):
        if self.gp_weighter is not None:
            S = int(_get(self.config, "adapter.gp_num_mc_samples_eval", 100) or 1)
            with torch.no_grad():       # mean_s 100 f.p_hat_s == 100 f.(mean_s p_hat_s) (tip_adapter.py:211-217)
                pm = self.gp_weighter.collapsed_prototypes(max(1, S), eps=getattr(self, "eval_eps", None))
            return ops.matmul_nt(feats_hat, pm, 100.0, heads._precision(self.config))
        return ops.matmul_nt(feats_hat, self.clip_weights, 100.0, heads._precision(self.config))           # tip_adapter.py:219

    def model_inference(self, features):
        f_hat = ops.row_normalize(features)
        return ops.tip_logits(f_hat, self.cache_keys, self.cache_labels, self._clip_logits(f_hat), self.best_beta, self.best_alpha, self.num_classes)

    @torch.no_grad()
    def _compute_final_metrics(self) -> Dict[str, Any]:
        logits = self.model_inference(self.features_test)
        res = metrics.evaluate_calibration(logits, self.labels_test)
        return {"top1_acc": float(res["top1_acc"]), "ece": float(res["ece"]), "aece": float(res["aece"]),
                "calibration": res["calibration"], "adaptive_calibration": res["adaptive_calibration"]}

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
            nb = self._num_batches()
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
