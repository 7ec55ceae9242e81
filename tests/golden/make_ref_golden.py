"""Golden vectors produced by EXECUTING THE REFERENCE'S OWN FILES (build container only).

    python tests/golden/make_ref_golden.py            # writes tests/golden/ref_gp.npz, ref_heads.npz, ref_train.npz

``/root/reference/trainers/{gp_template_weigher,adapter,taskres,clip_adapter,tip_adapter}.py`` and
``utils/{trainer,optimization,config,metrics}.py`` are imported UNMODIFIED (tests/golden/_ref_env.py) on top of the
minimal gpytorch / linear_operator / entmax stand-ins of ``oracle/_shim`` (those libraries cannot be installed offline;
the shim restates the ~20 library routines the reference reaches, see oracle/_shim/README.md).  Everything the
reference tree itself holds on the hot path is therefore executed, not restated: PCA, f0 prior, mean-module tail, the
``[:, :, :N_templates]`` slice, the ``batch == K`` branch, the prototype einsum, ``compute_loss``, the TaskRes /
CLIP-Adapter / Tip-Adapter heads, ``_get_template_weights``, ``_build_cache``, ``_search_hyperparams``, the optimizer /
scheduler builders and (with a stand-in CLIP that returns cached features) the whole ``Trainer.train()`` of Tip-Adapter,
TaskRes and CLIP-Adapter including their GP pre-training loops.

The fixtures pin (tests/test_ref_golden.py, CPU) ``oracle/*.py`` and (tests/test_gpu_ref_golden.py, GPU) the CUDA path.
The GPU box has no /root/reference: only the .npz files travel.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
import _ref_env  # noqa: E402

_ref_env.setup()
sys.path.append(ROOT)                      # after the reference: its `utils` / `datasets` packages must win

import gpytorch  # noqa: E402  (the shim)
from gpytorch import distributions as shim_dist  # noqa: E402

from clip_gp_b200 import synth  # noqa: E402  (synthetic text bank / features only)
from oracle import philox  # noqa: E402

ref_gpw = _ref_env.ref_module("trainers.gp_template_weigher")
ref_cfg = _ref_env.ref_module("utils.config")


def np_(t):
    return t.detach().cpu().numpy().copy()


def make_config(**adapter):
    cfg = ref_cfg.Config()
    cfg.use_cuda = False
    for k, v in adapter.items():
        setattr(cfg.adapter, k, v)
    return cfg


class eps_hook:
    """Feed explicit base noise [C, Nx, S] to every rsample inside the block (records what was requested)."""

    def __init__(self, provider):
        self.provider, self.calls = provider, []

    def __enter__(self):
        def hook(shape, dtype, device):
            e = self.provider(tuple(shape), len(self.calls))
            self.calls.append(tuple(shape))
            assert tuple(e.shape) == tuple(shape), (e.shape, shape)
            return e.to(dtype=dtype, device=device)
        shim_dist.BASE_SAMPLES_HOOK = hook
        return self

    def __exit__(self, *a):
        shim_dist.BASE_SAMPLES_HOOK = None


def perturb_gp(gp, seed, bias=True):
    """A trained-like state for every learnable tensor of the reference module (after its first-call initialisation)."""
    g = torch.Generator().manual_seed(seed)
    vs = gp.variational_strategy
    q = vs._variational_distribution
    C, n = q.variational_mean.shape
    with torch.no_grad():
        q.variational_mean.copy_(0.5 * torch.randn(C, n, generator=g))
        q.chol_variational_covar.copy_(torch.eye(n).repeat(C, 1, 1) + 0.1 * torch.randn(C, n, n, generator=g))  # NOT tril: masked inside
        vs.inducing_points[:, -1] += 0.05 * torch.randn(C, vs.inducing_points.shape[-1], generator=g)
        for name, p in gp.covar_module.named_parameters():
            p.add_(0.1 * torch.randn(p.shape, generator=g))
        if bias:
            gp.mean_module.cls_bias.copy_(0.3 * torch.randn(gp.mean_module.cls_bias.shape, generator=g))
            gp.mean_module.tmp_bias.copy_(0.3 * torch.randn(gp.mean_module.tmp_bias.shape, generator=g))


def gp_param_dict(gp):
    out = {"Z": gp.variational_strategy.inducing_points,
           "m": gp.variational_strategy._variational_distribution.variational_mean,
           "chol": gp.variational_strategy._variational_distribution.chol_variational_covar,
           "cls_bias": gp.mean_module.cls_bias, "tmp_bias": gp.mean_module.tmp_bias}
    for name, p in gp.covar_module.named_parameters():
        out[name.split(".")[-1]] = p                 # raw_lengthscale / raw_outputscale / raw_variance
    return out


# ------------------------------------------------------------------------------------------------------------------ GP
GP_CASES = {
    # name: (C, T, D, pca_dim, S)
    "tiny": (12, 5, 64, 16, 3),
    "t32": (6, 32, 96, 48, 4),        # the headline per-class shape n = 33
    "t1": (5, 1, 32, 8, 2),           # single template (n = 2)
    "lowrank": (3, 4, 40, 256, 3),    # gp_pca_dim > rank: red_dim = min(256, C*T) = 12  (gp_template_weigher.py:33-34)
    "t64": (4, 64, 128, 48, 3),       # the cfg5 per-class shape n = 65 (general block kernels); pins the ORACLE on CPU (tests/test_ref_golden.py)
}


def gp_goldens():
    out = {}
    for cname, (C, T, D, pd, S) in GP_CASES.items():
        E, mu_c = synth.make_text_bank(C, T, D, seed=4242 + C)
        for kern in ("rbf", "matern", "linear"):
            key = f"{cname}/{kern}"
            cfg = make_config(gp_pca_dim=pd, gp_kernel_type=kern, gp_prior_temp=1.0)
            torch.manual_seed(99)
            gp = ref_gpw.GaussianProcessTemplateWeighter(text_embeddings=E, cfg=cfg)
            gp.train()
            out[f"{cname}/E"] = np_(E)
            out[f"{key}/pca_mean"] = np_(gp._pca_mean); out[f"{key}/pca_W"] = np_(gp._pca_W)
            out[f"{key}/templates_red"] = np_(gp._templates_red)
            out[f"{key}/Z0"] = np_(gp.variational_strategy.inducing_points)
            out[f"{key}/f0"] = np_(gp.mean_module.f0)
            out[f"{key}/cls_mean_init"] = np_(gp._cls_mean_init)
            if kern == "rbf":
                out[f"{key}/raw_lengthscale0"] = np_(gp.covar_module.base_kernel.raw_lengthscale)
                out[f"{key}/lengthscale0"] = np_(gp.covar_module.base_kernel.lengthscale)
            # first call: gpytorch initialises q(u) (consumes torch RNG: randn_like [C,n]) BEFORE the base noise is drawn
            torch.manual_seed(7)
            p_first = gp.sample_prototypes(S)
            out[f"{key}/first_call/m"] = np_(gp.variational_strategy._variational_distribution.variational_mean)
            out[f"{key}/first_call/chol"] = np_(gp.variational_strategy._variational_distribution.chol_variational_covar)
            out[f"{key}/first_call/w"] = np_(gp.scores); out[f"{key}/first_call/protos"] = np_(p_first)
            torch.manual_seed(7)
            m_chk = 1e-3 * torch.randn(C, T + 1); eps_chk = torch.randn(C, T, S)
            assert torch.equal(m_chk, gp.variational_strategy._variational_distribution.variational_mean.detach())
            out[f"{key}/first_call/eps"] = np_(eps_chk)

            perturb_gp(gp, seed=31 + T)
            params = gp_param_dict(gp)
            for k, p in params.items():
                out[f"{key}/param/{k}"] = np_(p)
            # (1) standard call: Nx = T
            eps = torch.randn(C, T, S, generator=torch.Generator().manual_seed(555))
            for p in gp.parameters():
                p.grad = None
            with eps_hook(lambda shape, i: eps):
                protos = gp.sample_prototypes(S)
            w = gp.scores
            kl = gp.variational_strategy.kl_divergence()
            qf = gp(gp._templates_red)                              # the predictive the samples came from (no RNG)
            g = torch.Generator().manual_seed(777)
            dP = torch.randn(protos.shape, generator=g); dkl = torch.randn(kl.shape, generator=g)
            ((protos * dP).sum() + (kl * dkl).sum()).backward()
            out[f"{key}/eps"] = np_(eps); out[f"{key}/dP"] = np_(dP); out[f"{key}/dkl"] = np_(dkl)
            out[f"{key}/w"] = np_(w); out[f"{key}/protos"] = np_(protos); out[f"{key}/kl"] = np_(kl)
            out[f"{key}/mu"] = np_(qf.mean); out[f"{key}/Sigma"] = np_(qf.covariance_matrix)
            for k, p in params.items():
                out[f"{key}/grad/{k}"] = np_(p.grad if p.grad is not None else torch.zeros_like(p))
            # (2) the `visual_embeddings.shape[0] == num_classes` branch (gp_template_weigher.py:198-203): Nx = T + 1
            vis = synth.make_features(mu_c, torch.arange(C), seed=5, noise=1.0)
            eps1 = torch.randn(C, T + 1, S, generator=torch.Generator().manual_seed(556))
            with eps_hook(lambda shape, i: eps1) as h, torch.no_grad():
                protos_v = gp.sample_prototypes(S, visual_embeddings=vis)
                assert h.calls == [(C, T + 1, S)]
            out[f"{key}/vis/features"] = np_(vis); out[f"{key}/vis/eps"] = np_(eps1)
            out[f"{key}/vis/w"] = np_(gp.scores); out[f"{key}/vis/protos"] = np_(protos_v)
            # a visual batch whose size differs from K is ignored (:204-210)
            with eps_hook(lambda shape, i: eps) as h, torch.no_grad():
                protos_o = gp.sample_prototypes(S, visual_embeddings=torch.cat([vis, vis])[: C + 1])
                assert h.calls == [(C, T, S)]
            # under no_grad gpytorch's sq_dist takes its `x1_eq_x2 and not requires_grad` branch (diagonal forced to 0): the
            # result differs from the grad-mode call by fp32 rounding only
            out[f"{key}/nograd/protos"] = np_(protos_o); out[f"{key}/nograd/w"] = np_(gp.scores)
            print(f"  {key}: no_grad vs grad-mode prototypes max |diff| = {float((protos_o - protos.detach()).abs().max()):.2e}")
            # (3) eval mode re-uses the memoised chol(K_ZZ) of the last call; from a fresh cache it equals the train-mode result
            gp.eval(); gp.variational_strategy._clear_cache()
            with eps_hook(lambda shape, i: eps), torch.no_grad():
                protos_e = gp.sample_prototypes(S)
            out[f"{key}/eval_mode/protos"] = np_(protos_e)
            # (4) initialize_from_weights is a no-op for T > 1 (shape mismatch swallowed) and a broadcast copy for T == 1
            m_before = gp.variational_strategy._variational_distribution.variational_mean.detach().clone()
            gp.initialize_from_weights(torch.full((C, T), 1.0 / T))
            out[f"{key}/init_from_weights/m_after"] = np_(gp.variational_strategy._variational_distribution.variational_mean)
            out[f"{key}/init_from_weights/changed"] = np.array(
                not torch.equal(m_before, gp.variational_strategy._variational_distribution.variational_mean.detach()))
            out[f"{key}/state_dict_keys"] = np.array(sorted(gp.state_dict().keys()))
            # (5) the PRIOR gp.forward(x) (gp_template_weigher.py:167-175) at the template rows (N = T) and at the inducing rows
            #     (N = T + 1: mean-module tail), in grad mode and under no_grad (gpytorch's exact-zero-diagonal distance branch)
            for tag, xin in (("templates", gp._templates_red), ("inducing", gp.variational_strategy.inducing_points.detach())):
                prior = gp.forward(xin)
                out[f"{key}/prior/{tag}/mean"] = np_(prior.mean); out[f"{key}/prior/{tag}/covar"] = np_(prior.covariance_matrix)
                with torch.no_grad():                              # the kernel is evaluated lazily: densify inside the block
                    out[f"{key}/prior/{tag}/covar_nograd"] = np_(gp.forward(xin).covariance_matrix)
    return out


# ------------------------------------------------------------------------------------------- whole trainers (fake CLIP)
import contextlib
import io
import json
import tempfile

import _fake_clip  # noqa: E402

ref_adapter = _ref_env.ref_module("trainers.adapter")
ref_taskres = _ref_env.ref_module("trainers.taskres")
ref_clipad = _ref_env.ref_module("trainers.clip_adapter")
ref_tip = _ref_env.ref_module("trainers.tip_adapter")
ref_utrainer = _ref_env.ref_module("utils.trainer")

TR = dict(K=10, M=4, D=32, pca=8, shots=4, N_test=50, N_val=30, bs_train=16, bs_test=32, S_train=3, S_eval=5, seed=21)


def trainer_world(kernel="rbf"):
    """Synthetic cached features + text bank shared by all whole-trainer goldens (raw, un-normalised image features)."""
    K, M, D = TR["K"], TR["M"], TR["D"]
    E, mu = synth.make_text_bank(K, M, D, seed=9001)
    g = torch.Generator().manual_seed(9002)
    y_tr = torch.arange(K).repeat_interleave(TR["shots"])[torch.randperm(K * TR["shots"], generator=g)]
    mk = lambda y, sd: (mu[y] + 2.5 * torch.randn(y.shape[0], D, generator=torch.Generator().manual_seed(sd))) * 1.7
    y_te = torch.randint(0, K, (TR["N_test"],), generator=g); y_va = torch.randint(0, K, (TR["N_val"],), generator=g)
    return dict(E=E, mu=mu, f_tr=mk(y_tr, 1), y_tr=y_tr, f_te=mk(y_te, 2), y_te=y_te, f_va=mk(y_va, 3), y_va=y_va,
                classnames=[f"class{i:03d}" for i in range(K)])


def trainer_config(name, kernel="rbf", use_gp=True, **adapter):
    cfg = ref_cfg.Config()
    cfg.trainer_name = name
    cfg.use_cuda = False
    cfg.seed = 1
    cfg.dataset.num_shots = TR["shots"]
    cfg.dataloader.batch_size_train, cfg.dataloader.batch_size_test = TR["bs_train"], TR["bs_test"]
    cfg.train.enable_tensorboard = False
    cfg.train.enable_adapter_checkpoints = False
    cfg.train.print_freq = 100
    cfg.optim.name, cfg.optim.lr, cfg.optim.max_epoch, cfg.optim.lr_scheduler = "adamw", 0.01, 3, "cosine"
    a = cfg.adapter
    a.use_gp, a.gp_kernel_type, a.gp_pca_dim, a.num_templates = use_gp, kernel, TR["pca"], TR["M"]
    a.gp_num_mc_samples_train, a.gp_num_mc_samples_eval = TR["S_train"], TR["S_eval"]
    a.gp_lr, a.gp_beta, a.l2_lambda = 1e-3, 0.01, 0.5
    a.clip_adapter_epochs = 3
    for k, v in adapter.items():
        assert hasattr(a, k), k
        setattr(a, k, v)
    return cfg


class NoiseBook:
    """Base noise for every rsample of a whole-trainer run.

    * calls made while the step loss is being computed (flag set by the compute_loss / GP pre-train wrappers) get the counter
      stream of the CUDA path: philox.eps_tensor(seed, step) -- the draws the engine will make for the same (seed, step);
    * every eval-mode call (S == S_eval) gets ONE fixed tensor (the reference re-samples per test batch; feeding the same noise to
      every batch makes the MC-averaged eval a deterministic function the CUDA path can reproduce with one pass);
    * the remaining train-mode calls (logging-only passes of adapter.py:339,357) get an unrelated stream."""

    def __init__(self, C, T, sequential=False):
        self.C, self.T = C, T
        self.seed, self.step, self.in_loss, self.log = TR["seed"], 0, False, []
        self.eps_eval = {}
        self.sequential = sequential    # trainers without compute_loss: EVERY train-mode draw (S == S_train) is a step of the stream

    def __call__(self, shape, idx):
        C, Nx, S = shape
        if self.sequential and S == TR["S_train"]:
            assert Nx == self.T, shape
            self.log.append(("loss", self.step, shape))
            self.step += 1
            return philox.eps_tensor(self.seed, self.step - 1, C, Nx, S)
        if self.in_loss:
            assert Nx == self.T, shape
            self.log.append(("loss", self.step, shape))
            return philox.eps_tensor(self.seed, self.step, C, Nx, S)
        if S == TR["S_eval"]:
            if Nx not in self.eps_eval:
                self.eps_eval[Nx] = philox.eps_tensor(self.seed + 1000, 0, C, Nx, S)
            self.log.append(("eval", -1, shape))
            return self.eps_eval[Nx]
        self.log.append(("other", idx, shape))
        return philox.eps_tensor(self.seed + 5000, idx, C, Nx, S)


class RecordingLoader:
    """Wraps the few-shot train DataLoader: same attributes, but every pass over it is recorded (the shuffle order is the one
    piece of host-side randomness a replay needs)."""

    def __init__(self, loader, passes):
        self._loader, self._passes = loader, passes

    def __iter__(self):
        cur = []
        self._passes.append(cur)
        for batch in self._loader:
            cur.append((batch["img"].clone(), batch["label"].clone()))
            yield batch

    def __len__(self):
        return len(self._loader)

    def __getattr__(self, name):
        return getattr(self._loader, name)


def run_reference_trainer(ref_mod, cfg, world, pretrain_epochs_attr=None, init_hook=None, sequential=False, bs_train=None):
    """Instantiate the reference trainer on the fake data manager, run its own train(), return (trainer, record)."""
    register = _fake_clip.install([ref_adapter, ref_taskres, ref_clipad, ref_tip, ref_utrainer], world["E"], world["classnames"],
                                  ref_utrainer._get_templates)
    with contextlib.redirect_stdout(io.StringIO()):
        register(cfg)
    dm = _fake_clip.FakeDataManager(world["classnames"], world["f_tr"], world["y_tr"], world["f_te"], world["y_te"], world["f_va"],
                                    world["y_va"], bs_train or TR["bs_train"], TR["bs_test"])
    out_dir = tempfile.mkdtemp(prefix="refgolden_")
    cfg.output_dir = out_dir
    trainer = ref_mod.Trainer(cfg, dm)
    book = NoiseBook(TR["K"], TR["M"], sequential=sequential)
    rec = {"batches_f": [], "batches_y": [], "losses": [], "lrs": [], "passes": []}
    if sequential:
        trainer.train_loader_x = RecordingLoader(trainer.train_loader_x, rec["passes"])
    # instance-level wrappers (the reference classes stay untouched): record each step's batch / loss, flag the loss pass
    orig_fb = trainer.forward_backward

    def fb(batch):
        if isinstance(batch, dict):
            rec["batches_f"].append(batch["img"].clone()); rec["batches_y"].append(batch["label"].clone())
        else:
            rec["batches_f"].append(torch.as_tensor(batch[0]).clone()); rec["batches_y"].append(torch.as_tensor(batch[1]).clone())
        if getattr(trainer, "optim", None) is not None:
            rec["lrs"].append([g_["lr"] for g_ in trainer.optim.param_groups])
        res = orig_fb(batch)
        rec["losses"].append(float(res["loss"]))
        return res
    trainer.forward_backward = fb
    if hasattr(trainer, "compute_loss"):
        orig_cl = trainer.compute_loss

        def cl(*a, **k):
            book.in_loss = True
            try:
                return orig_cl(*a, **k)
            finally:
                book.in_loss = False
                book.step += 1
        trainer.compute_loss = cl
    # gpytorch memoises chol(K_ZZ) in eval mode (`@cached("cholesky_factor", ignore_args=True)`): an eval pass right after an
    # optimizer step would combine the PREVIOUS parameters' factor with the new kernel blocks.  SURVEY 8c(4) classifies that as a
    # reference artefact not to reproduce ("run parity from a fresh state"): the final evaluation starts from a cleared cache.
    def fresh(fn):
        def wrapped(*a, **k):
            for holder in (getattr(trainer, "model", None), trainer):
                gpw = getattr(holder, "gp_weighter", None)
                if gpw is not None:
                    gpw.variational_strategy._clear_cache()
            return fn(*a, **k)
        return wrapped
    trainer.test = fresh(trainer.test)
    trainer._compute_final_metrics = fresh(trainer._compute_final_metrics)
    if init_hook is not None:
        orig_bm = trainer.build_model

        def bm():
            orig_bm()
            init_hook(trainer, rec)
        trainer.build_model = bm
    # the head trainers create their weighter inside train(): a recording factory (the reference CLASS is untouched) triggers
    # gpytorch's first-call initialisation of q(u) right away and keeps the state the pre-training loop starts from
    real_cls = ref_gpw.GaussianProcessTemplateWeighter

    def factory(*a, **k):
        gp = real_cls(*a, **k)
        gp.train()
        with torch.no_grad():
            gp.sample_prototypes(1)
        gp.variational_strategy._clear_cache()
        rec["gp_before"] = {}
        _gp_state(rec["gp_before"], "s", gp)
        return gp
    if sequential:
        for m_ in (ref_taskres, ref_clipad, ref_tip):
            m_.GaussianProcessTemplateWeighter = factory
    torch.manual_seed(1234); np.random.seed(1234)
    log = io.StringIO()
    try:
        with eps_hook(book) as h, contextlib.redirect_stdout(log):
            trainer.train()
    finally:
        for m_ in (ref_taskres, ref_clipad, ref_tip):
            m_.GaussianProcessTemplateWeighter = real_cls
    rec["stdout"] = log.getvalue()
    rec["noise_log"] = book.log
    rec["book"] = book
    mj = os.path.join(out_dir, "metrics.json")
    rec["metrics_json"] = json.load(open(mj)) if os.path.exists(mj) else None
    return trainer, rec


def pack_metrics(out, key, m):
    for k in ("top1_acc", "accuracy", "ece", "aece"):
        if k in m:
            out[f"{key}/{k}"] = np.float64(m[k])
    for cal in ("calibration", "adaptive_calibration"):
        if cal in m and m[cal]:
            out[f"{key}/{cal}/bin_count"] = np.array(m[cal]["bin_count"], np.int64)
            out[f"{key}/{cal}/bin_acc"] = np.array(m[cal]["bin_acc"], np.float64)
            out[f"{key}/{cal}/bin_conf"] = np.array(m[cal]["bin_conf"], np.float64)


def adapter_trainer_goldens(out):
    """The reference's Adapter trainer end to end (trainers/adapter.py:582-700 train(), :702-884 run_epoch, :328-385
    forward_backward, :387-476 compute_loss, optimizer / scheduler from utils/optimization.py) for each kernel."""
    world = trainer_world()
    for k_, v_ in world.items():
        if torch.is_tensor(v_):
            out[f"world/{k_}"] = np_(v_)
    for kern in ("rbf", "matern", "linear"):
        key = f"adapter/{kern}"
        cfg = trainer_config("Adapter", kern, template_init_method="val_weighted")

        def init_hook(trainer, rec):
            gp = trainer.model.gp_weighter
            gp.train()
            gp.sample_prototypes(1)                          # gpytorch's first-call initialisation of q(u) (consumes torch RNG)
            perturb_gp(gp, seed=77, bias=False)              # start from a non-trivial state so that every gradient path is live
            gp.variational_strategy._clear_cache()
            rec["init"] = {k: p.detach().clone() for k, p in gp_param_dict(gp).items()}
            rec["init"]["W"] = trainer.model.visual_proj.weight.detach().clone()
        trainer, rec = run_reference_trainer(ref_adapter, cfg, world, init_hook=init_hook)
        gp = trainer.model.gp_weighter
        for n_, t_ in rec["init"].items():
            out[f"{key}/init/{n_}"] = np_(t_)
        out[f"{key}/templates_red"] = np_(gp._templates_red); out[f"{key}/pca_W"] = np_(gp._pca_W); out[f"{key}/pca_mean"] = np_(gp._pca_mean)
        out[f"{key}/f0"] = np_(gp.mean_module.f0)
        out[f"{key}/text_embeddings"] = np_(trainer.model.text_embeddings)
        out[f"{key}/batches_f"] = np.stack([np_(b) for b in rec["batches_f"]]); out[f"{key}/batches_y"] = np.stack([np_(b) for b in rec["batches_y"]])
        out[f"{key}/losses"] = np.array(rec["losses"]); out[f"{key}/lrs"] = np.array(rec["lrs"])
        for n_, p_ in gp_param_dict(gp).items():
            out[f"{key}/final/{n_}"] = np_(p_)
        out[f"{key}/final/W"] = np_(trainer.model.visual_proj.weight)
        out[f"{key}/eps_eval"] = np_(rec["book"].eps_eval[TR["M"]])
        out[f"{key}/philox_seed"] = np.int64(TR["seed"])
        pack_metrics(out, f"{key}/zero_shot", trainer.zero_shot_metrics)
        pack_metrics(out, f"{key}/final_metrics", rec["metrics_json"]["metrics"])
        n_loss = sum(1 for t in rec["noise_log"] if t[0] == "loss")
        assert n_loss == len(rec["losses"]) == 6, (n_loss, len(rec["losses"]))
        print(f"  {key}: losses {np.round(rec['losses'], 4).tolist()} acc {rec['metrics_json']['metrics']['accuracy']:.2f} "
              f"(zero-shot {trainer.zero_shot_metrics['top1_acc']:.2f}) ece {rec['metrics_json']['metrics']['ece']:.3f}")
    return out


def _gp_state(out, key, gp):
    for n_, p_ in gp_param_dict(gp).items():
        out[f"{key}/{n_}"] = np_(p_)
    out[f"{key}/templates_red"] = np_(gp._templates_red); out[f"{key}/pca_W"] = np_(gp._pca_W); out[f"{key}/pca_mean"] = np_(gp._pca_mean)
    out[f"{key}/f0"] = np_(gp.mean_module.f0); out[f"{key}/templates"] = np_(gp._templates)


def _zero_shot_from_stdout(text):
    import re
    m = re.search(r"Zero-Shot accuracy on test: ([0-9.]+)", text)
    return float(m.group(1)) if m else float("nan")


def head_trainer_goldens(out):
    """TaskRes / CLIP-Adapter / Tip-Adapter(-F): the reference's whole train() (GP pre-training loop included) per variant."""
    world = trainer_world()
    BS = 8                                                   # 40 few-shot features -> 5 full batches (no drop_last loss)
    # ---------------------------------------------------------------- TaskRes (taskres.py:197-398)
    for use_gp in (False, True):
        key = f"taskres/{'gp' if use_gp else 'plain'}"
        cfg = trainer_config("TaskRes", "rbf", use_gp=use_gp, taskres_optimizer="adam", taskres_lr=2e-3, taskres_epochs=5,
                             taskres_residual_scale=0.5, template_init_method="uniform")
        cfg.optim.max_epoch = 4                              # GP pre-training epochs (taskres.py:257); main loop: clip_adapter_epochs = 3
        trainer, rec = run_reference_trainer(ref_taskres, cfg, world, sequential=True, bs_train=BS)
        out[f"{key}/zero_shot_acc"] = np.float64(_zero_shot_from_stdout(rec["stdout"]))
        passes = rec["passes"]
        first_epoch_pass = 0
        if use_gp:
            out[f"{key}/pretrain_f"] = np.concatenate([np_(f) for f, _ in passes[0]]); out[f"{key}/pretrain_y"] = np.concatenate([np_(y) for _, y in passes[0]])
            first_epoch_pass = 1
            _gp_state(out, f"{key}/gp_after_pretrain", trainer.model.gp_weighter)
            out.update({f"{key}/gp_before_pretrain/{k_[2:]}": v_ for k_, v_ in rec["gp_before"].items()})
            assert "GP weighting failed" not in rec["stdout"], rec["stdout"][-2000:]
        out[f"{key}/base_text_features"] = np_(trainer.model.taskres_learner.base_text_features)
        out[f"{key}/batches_f"] = np.stack([np_(b) for b in rec["batches_f"]]); out[f"{key}/batches_y"] = np.stack([np_(b) for b in rec["batches_y"]])
        out[f"{key}/losses"] = np.array(rec["losses"]); out[f"{key}/lrs"] = np.array(rec["lrs"])
        out[f"{key}/final/residuals"] = np_(trainer.model.taskres_learner.text_feature_residuals)
        if use_gp:
            out[f"{key}/eps_eval"] = np_(rec["book"].eps_eval[TR["M"]])
        pack_metrics(out, f"{key}/final_metrics", rec["metrics_json"]["metrics"])
        print(f"  {key}: zs {out[f'{key}/zero_shot_acc']:.2f} losses {np.round(rec['losses'][:3], 4).tolist()}..{rec['losses'][-1]:.4f} "
              f"final {rec['metrics_json']['metrics']}"[:260])
    # ---------------------------------------------------------------- CLIP-Adapter (clip_adapter.py:230-355)
    for use_gp in (False, True):
        key = f"clip_adapter/{'gp' if use_gp else 'plain'}"
        cfg = trainer_config("CLIP-Adapter", "linear", use_gp=use_gp, clip_adapter_optimizer="adam", clip_adapter_lr=1e-3,
                             clip_adapter_epochs=3, clip_adapter_ratio=0.2, clip_adapter_reduction=4)
        cfg.optim.max_epoch = 4

        def init_hook(trainer, rec):
            rec["fc1"] = trainer.model.adapter.fc1.weight.detach().clone(); rec["fc2"] = trainer.model.adapter.fc2.weight.detach().clone()
            rec["clip_weights0"] = trainer.model.clip_weights.detach().clone()
        trainer, rec = run_reference_trainer(ref_clipad, cfg, world, sequential=True, bs_train=BS, init_hook=init_hook)
        out[f"{key}/zero_shot_acc"] = np.float64(_zero_shot_from_stdout(rec["stdout"]))
        out[f"{key}/init/fc1"] = np_(rec["fc1"]); out[f"{key}/init/fc2"] = np_(rec["fc2"]); out[f"{key}/init/clip_weights"] = np_(rec["clip_weights0"])
        passes = rec["passes"]
        if use_gp:
            out[f"{key}/pretrain_f"] = np.concatenate([np_(f) for f, _ in passes[0]]); out[f"{key}/pretrain_y"] = np.concatenate([np_(y) for _, y in passes[0]])
            _gp_state(out, f"{key}/gp_after_pretrain", trainer.model.gp_weighter)
            out.update({f"{key}/gp_before_pretrain/{k_[2:]}": v_ for k_, v_ in rec["gp_before"].items()})
            out[f"{key}/eps_eval"] = np_(rec["book"].eps_eval[TR["M"]])
            assert "GP weighting failed" not in rec["stdout"], rec["stdout"][-2000:]
        out[f"{key}/clip_weights"] = np_(trainer.model.clip_weights)
        out[f"{key}/batches_f"] = np.stack([np_(b) for b in rec["batches_f"]]); out[f"{key}/batches_y"] = np.stack([np_(b) for b in rec["batches_y"]])
        out[f"{key}/losses"] = np.array(rec["losses"]); out[f"{key}/lrs"] = np.array(rec["lrs"])
        out[f"{key}/final/fc1"] = np_(trainer.model.adapter.fc1.weight); out[f"{key}/final/fc2"] = np_(trainer.model.adapter.fc2.weight)
        pack_metrics(out, f"{key}/final_metrics", rec["metrics_json"]["metrics"])
        print(f"  {key}: zs {out[f'{key}/zero_shot_acc']:.2f} losses {np.round(rec['losses'][:3], 4).tolist()}..{rec['losses'][-1]:.4f}")
    # ---------------------------------------------------------------- Tip-Adapter / Tip-Adapter-F (tip_adapter.py:82-362)
    for use_gp in (False, True):
        for trainable in (False, True):
            key = f"tip/{'gp' if use_gp else 'plain'}/{'F' if trainable else 'cache'}"
            cfg = trainer_config("Tip-Adapter", "matern", use_gp=use_gp, tip_adapter_trainable=trainable, tip_adapter_lr=1e-3,
                                 tip_adapter_eps=1e-4, tip_adapter_epochs=3, tip_adapter_init_alpha=20.0, tip_adapter_init_beta=2.0)
            cfg.optim.max_epoch = 4
            captured = {}
            trainer_holder = {}

            def init_hook(trainer, rec, captured=captured):
                orig = trainer._compute_final_metrics_tip_adapter

                def wrapped(adapter=None, beta=None, alpha=None):
                    if adapter is not None:
                        captured["adapter_w"] = adapter.weight.detach().clone()
                    gpw = getattr(trainer, "gp_weighter", None)
                    if gpw is not None:
                        gpw.variational_strategy._clear_cache()
                    return orig(adapter=adapter, beta=beta, alpha=alpha)
                trainer._compute_final_metrics_tip_adapter = wrapped
            trainer, rec = run_reference_trainer(ref_tip, cfg, world, sequential=True, bs_train=BS, init_hook=init_hook)
            out[f"{key}/zero_shot_acc"] = np.float64(_zero_shot_from_stdout(rec["stdout"]))
            passes = rec["passes"]
            pi = 0
            if use_gp:
                out[f"{key}/pretrain_f"] = np.concatenate([np_(f) for f, _ in passes[0]]); out[f"{key}/pretrain_y"] = np.concatenate([np_(y) for _, y in passes[0]])
                _gp_state(out, f"{key}/gp_after_pretrain", trainer.gp_weighter)
                out.update({f"{key}/gp_before_pretrain/{k_[2:]}": v_ for k_, v_ in rec["gp_before"].items()})
                out[f"{key}/eps_eval"] = np_(rec["book"].eps_eval[TR["M"]])
                assert "GP weighting failed" not in rec["stdout"], rec["stdout"][-2000:]
                pi = 1
            out[f"{key}/clip_weights"] = np_(trainer.clip_weights)                     # [D,K]
            out[f"{key}/cache_keys"] = np_(trainer.cache_keys); out[f"{key}/cache_vals"] = np_(trainer.cache_vals)
            # pass `pi` built the cache (tip_adapter.py:43-50); the following passes are the Tip-Adapter-F epochs.  The trainable
            # nn.Linear SHARES STORAGE with cache_keys (`adapter.weight = nn.Parameter(self.cache_keys)`, :230), so after
            # training `trainer.cache_keys` holds the trained weights: the initial keys are re-derived from the recorded pass
            cf = torch.cat([f for f, _ in passes[pi]]); cy = torch.cat([y for _, y in passes[pi]])
            out[f"{key}/cache_keys0"] = np_(cf / cf.norm(dim=-1, keepdim=True)); out[f"{key}/cache_labels0"] = np_(cy)
            if trainable:
                ep_passes = passes[pi + 1:pi + 1 + 3]
                out[f"{key}/batches_f"] = np.stack([np_(f) for ps in ep_passes for f, _ in ps]); out[f"{key}/batches_y"] = np.stack([np_(y) for ps in ep_passes for _, y in ps])
                out[f"{key}/final/adapter_w"] = np_(captured["adapter_w"])
            out[f"{key}/best_beta"] = np.float64(trainer._tip_adapter_best_beta); out[f"{key}/best_alpha"] = np.float64(trainer._tip_adapter_best_alpha)
            pack_metrics(out, f"{key}/final_metrics", rec["metrics_json"]["metrics"])
            print(f"  {key}: zs {out[f'{key}/zero_shot_acc']:.2f} beta {trainer._tip_adapter_best_beta} alpha {trainer._tip_adapter_best_alpha} "
                  f"final acc {rec['metrics_json']['metrics']['top1_acc']:.2f} ece {rec['metrics_json']['metrics']['ece']:.3f}")
    return out


def main():
    torch.set_num_threads(4)
    which = sys.argv[1:] or ["gp", "train"]
    if "gp" in which:
        out = gp_goldens()
        np.savez_compressed(os.path.join(HERE, "ref_gp.npz"), **out)
        print("wrote ref_gp.npz:", len(out), "arrays,", os.path.getsize(os.path.join(HERE, "ref_gp.npz")) // 1024, "KiB")
    if "train" in which:
        out = {}
        adapter_trainer_goldens(out)
        head_trainer_goldens(out)
        np.savez_compressed(os.path.join(HERE, "ref_train.npz"), **out)
        print("wrote ref_train.npz:", len(out), "arrays,", os.path.getsize(os.path.join(HERE, "ref_train.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
