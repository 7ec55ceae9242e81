#!/usr/bin/env python
"""bench.py — GP-adapter train steps/s (and eval img/s) on synthetic cached features of BASELINE.json's cfg2
shape (ViT-B/16 D=512, ImageNet-1k C=1000, T=32 templates, 16-shot, S=10 MC samples, B=128).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

One "step" = one GP-Adapter optimisation step over one batch of cached features (engine.py): GP forward,
prototypes, projection, logits for every MC sample, per-sample CE, all adjoints, L2 regulariser, AdamW.
`value` = steps/s with inputs resident in HBM (CUDA events around each step, L2 flushed between steps, max over
ranks); `e2e` = the same step driven from pinned HOST batches (H2D copy of the batch and D2H read of the loss inside
the timed region).  N > 1 (weak scaling): data parallel over the batch -- every rank steps its OWN 128-image batch with all S
MC samples, the global batch is N * 128, gradients + loss go through ONE NCCL all-reduce inside the captured step; `value`
counts 128-image batch-steps per second over all ranks (= N x optimisation steps/s; identical to steps/s at N = 1).  The north
star's S-sharded form of ONE batch (strong scaling of a latency-bound step) is timed next to it (`train_sample_sharded`), and
so is the full-batch step with the batch split over the ranks.  Eval images are sharded over ranks, the eval GP forward over
classes.
`--impl reference` times the CPU oracle (the restated reference path: torch CPU + autograd + AdamW) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOAD = "cfg2"
METRIC = "gp_adapter_train_steps_per_s"
UNIT = "steps/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=WORKLOAD)
    ap.add_argument("--eval-n", type=int, default=50000)
    ap.add_argument("--precision", default="tf32", choices=["fp32", "tf32", "bf16x3", "bf16"],
                    help="GEMMs of the train step and the eval pass: fp32 FFMA | tcgen05 TF32 on the fp32 tensors in place (default: the "
                         "reference's own GPU arithmetic) | tcgen05 split-bf16 (fp32-grade) | tcgen05 bf16")
    ap.add_argument("--no-fullbatch", action="store_true", help="skip the B = N_train full-batch (tensor-bound) leg")
    ap.add_argument("--no-variants", action="store_true", help="skip the other-precision timings of the step")
    ap.add_argument("--no-peer-update", action="store_true",
                    help="N > 1: gradient ncclAllReduce + AdamW instead of the fused NVLink peer-memory optimiser step (csrc/peer.cu)")
    ap.add_argument("--no-graph-collectives", action="store_true",
                    help="N > 1: launch the step eagerly instead of capturing it (NCCL all-reduces included) in a CUDA graph")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-kernels", action="store_true", help="also print the per-kernel time table")
    return ap.parse_args()


TRAFFIC_JSON = "r2_ncu_traffic.json"


def ncu_traffic(kernel_base):
    """DRAM bytes per launch (read + write) of the kernels behind one C-ABI call, from the committed `ncu --set full` capture
    (profiles/r2_ncu_traffic.json; bench.py itself never runs under a profiler)."""
    path = os.path.join(ROOT, "profiles", TRAFFIC_JSON)
    try:
        return json.load(open(path)).get(kernel_base, {}).get("dram_bytes")
    except Exception:
        return None


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"], "bf16_tflops_sustained": p.get("bf16_tflops_sustained"),
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md).  NVML is polled from a thread every
    ~2 ms (the timed region of the default run lasts tens of milliseconds, shorter than nvidia-smi's start-up); nvidia-smi in
    loop mode is the fallback when the NVML binding is missing."""

    _REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index: int):
        self.index, self.rows, self.proc, self.h, self._stop = index, [], None, None, threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.h = None

    def _poll(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                mhz = int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.rows.append((mhz, mask))
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.h is not None:
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index),
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.h is not None:
            self._stop.set()
            self.thread.join(timeout=1.0)
            sm = sorted(r[0] for r in self.rows)
            reasons = [n for n, bit in self._REASONS if any(r[1] & bit for r in self.rows)]
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(sm),
                    "source": "NVML polled every 2 ms during the timed region"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm), "source": "nvidia-smi -lms 20"}


def workload_string(name, shp, S):
    """`config.workload` of BOTH arms (the driver compares them verbatim)."""
    return (f"{name}: GP-Adapter optimisation step on cached features, C={shp.C} T={shp.T} D={shp.D} d={shp.d} S={S} B={shp.B} "
            f"shots={shp.shots} kernel={shp.kernel}; per-sample MC cross-entropy + KL + L2, AdamW")


def bench_lengthscale(E, d):
    """REFERENCE ARM ONLY (CPU oracle): median-heuristic length-scale (gp_template_weigher.py:103-107) on the templates of the
    first 64 classes; the full C*T = 32 000-point cdist is one-time setup outside the step."""
    from oracle import gp as ogp
    _, _, tr, _, _ = ogp.pca_setup(E, d)
    return ogp.median_lengthscale(tr[:64])


def product_lengthscale(gpw):
    """PRODUCT ARM: the same quantity from the repo's own setup kernel (csrc/setup.cu, exact radix select of the pairwise
    distances; no oracle import on this arm), on the same 64-class subset so that both arms start from the same value."""
    import torch.nn.functional as F
    from clip_gp_b200 import ops
    tr = gpw._templates_red[:64]
    return float(ops.median_pairwise_distance(F.normalize(tr.reshape(-1, tr.shape[-1]), p=2, dim=-1).contiguous()))


class _Cfg:
    def __init__(self, kernel, pca, ls=None):
        self.adapter = type("A", (), {"gp_pca_dim": pca, "gp_kernel_type": kernel})()


# ======================================================================================================= reference arm
def run_reference(args):
    """The reference's CPU path for the same step (oracle restatement: the reference itself cannot run offline without
    gpytorch/entmax).  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from clip_gp_b200 import synth
    from oracle import gp as ogp
    from oracle.train_step import OracleAdapter
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    wl = synth.make_workload(args.workload)
    shp = wl["shape"]
    S = shp.S
    ls = bench_lengthscale(wl["E"], shp.d) if shp.kernel == "rbf" else None
    st = ogp.build_state(wl["E"], shp.kernel, shp.d, lengthscale=ls)
    g = torch.Generator().manual_seed(1)
    st.var_mean = 1e-3 * torch.randn(shp.C, shp.T + 1, generator=g)
    orc = OracleAdapter(st, shp.D, shots=shp.shots)
    f, y = wl["f_train"], wl["y_train"]
    nb = f.shape[0] // shp.B

    def one(it):
        lo = (it % nb) * shp.B
        eps = torch.randn(shp.C, shp.T, S, generator=g)
        return orc.step(f[lo:lo + shp.B], y[lo:lo + shp.B], eps)

    for it in range(args.warmup):
        one(it)
    t0 = time.perf_counter()
    for it in range(args.steps):
        one(args.warmup + it)
    dt = time.perf_counter() - t0
    v = args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_string(args.workload, shp, S),
                       "arm": "CPU oracle port of the reference path (torch CPU + autograd + torch AdamW) on all host cores"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{args.steps} full optimisation steps (fwd + autograd bwd + AdamW) of the same workload"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ======================================================================================================= our arm
def algorithmic_bytes(shp, S, kernel_name, eng=None):
    """Algorithmic HBM bytes per launch of the step's kernels (DESIGN.md section 'kernels')."""
    C, T, D, d, n, B = shp.C, shp.T, shp.D, shp.d, shp.T + 1, shp.B
    f = 4
    seg = getattr(eng, "tc_seg", 0) if eng is not None and eng.cfg.precision != "fp32" else 0
    # warp path (test inputs alias the frozen inducing rows, so X is never read).  Gram: reads Z [C,n,d] + lengthscales, writes the K_ZZ
    # record; algebra: reads Lq, m, writes w [S,C,T], the base noise, the saved L (fp64), A, R, KL
    gp_fwd = f * (C * n * d + C * d + C * n * n) + f * (C * n * n + C * n + 2 * S * C * T + C * n * T + C * T * T + C) + 8 * C * n * n
    # algebra adjoint: reads R, w, eps, dw, A, Lq, m, L (fp64), writes dLq, dm and the dK scratch block (write, add, read);
    # kernel adjoint: reads dK, K_ZZ, Z, lengthscales, writes the length-scale / learnable-row / output-scale gradients
    gp_bwd = f * (C * T * T + 3 * S * C * T + C * n * T + C * n * n + C * n) + 8 * C * n * n + f * (4 * C * n * n + C * n) \
        + f * (C * n * n + C * n * d + C * d + 2 * C * d + C)
    # prototypes: reads E [C,T,D] (+ w); writes P_hat [S,C,D], norms and (tensor-core step) the bf16 operand rows
    proto_fwd = f * (C * T * D + S * C * T + S * C * D + S * C)
    # prototype adjoint: reads dP_hat [S,C,D], E (once per chunk of <= 8 samples), P_hat; writes dw
    proto_bwd = f * (2 * S * C * D + ((S + 7) // 8) * C * T * D + S * C * T + S * C)
    if eng is not None and getattr(eng, "fused_proto", False):
        gp_fwd += f * (C * T * D + S * C * D + S * C) + 2 * seg * S * C * D      # fused prototype stage (w stays on chip)
    if eng is not None and getattr(eng, "fused_proto_bwd", False):
        gp_bwd += f * (S * C * D + C * T * D + C * T * T + S * C + S * C * T)       # dP_hat, E, E E^T, norms in; dw out (inspection copy)
    table = {"gp_forward": gp_fwd, "gp_backward": gp_bwd, "proto_forward": proto_fwd, "proto_backward": proto_bwd}
    return table.get(kernel_name)


def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the clipgp kernels have no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    import __graft_entry__ as ge
    if rank == 0:
        ge.build()
    if world > 1:
        torch.distributed.barrier()
    from clip_gp_b200 import _lib, metrics, synth
    from clip_gp_b200.dist import sample_split as cdist_sample_split
    from clip_gp_b200.engine import EngineConfig, GPAdapterEngine
    from clip_gp_b200.gp_template_weigher import GaussianProcessTemplateWeighter

    wl = synth.make_workload(args.workload)
    shp = wl["shape"]
    S = shp.S
    torch.manual_seed(1)
    gpw = GaussianProcessTemplateWeighter(wl["E"].to(dev), _Cfg(shp.kernel, shp.d), lengthscale=1.0).to(dev)
    ls = None
    if shp.kernel == "rbf":
        ls = product_lengthscale(gpw)
        gpw.covar_module.base_kernel.initialize(lengthscale=ls)

    def engine_config(**kw):
        base = dict(S_train=S, S_eval=S, batch_size=shp.B, shots=shp.shots, seed=1234, rank=rank, world=world, precision=args.precision,
                    graph_collectives=not args.no_graph_collectives, shard="batch", peer_update=world > 1 and not args.no_peer_update)
        base.update(kw)
        if base["shard"] != "batch" or base["world"] == 1:
            base["peer_update"] = False
        return EngineConfig(**base)

    # N > 1: data parallel over the batch (every rank its own B-image batch, all S samples; ONE gradient all-reduce per step)
    cfg = engine_config()
    eng = GPAdapterEngine(gpw, cfg)
    f_all = wl["f_train"].to(dev)
    y_all = wl["y_train"].to(dev)
    nb = f_all.shape[0] // shp.B
    f_host = wl["f_train"].pin_memory()
    y_host = wl["y_train"].pin_memory()
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)     # > 126 MB L2

    def batch(it):
        lo = ((it * world + rank) % nb) * shp.B            # rank r takes batch it * world + r of the (shared) few-shot set
        return lo, lo + shp.B

    def sync_all():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize(dev)

    def time_steps(e, steps, warmup):
        """Device time of `steps` optimisation steps (CUDA events around each step on the launching stream, L2 flushed
        between steps outside the events, max over ranks).  Returns total ms."""
        Bsz = e.B
        nbb = max(1, f_all.shape[0] // Bsz)
        dp = world if e.batch_sharded else 1
        for it in range(max(warmup, 3)):
            lo = ((it * dp + (rank if dp > 1 else 0)) % nbb) * Bsz
            e.train_step(f_all[lo:lo + Bsz], y_all[lo:lo + Bsz])
        sync_all()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        sync_all()
        for it in range(steps):
            lo = (((warmup + it) * dp + (rank if dp > 1 else 0)) % nbb) * Bsz
            flush.fill_(0.0)                               # L2 flush between timed iterations (outside the events)
            e.in_feat.copy_(f_all[lo:lo + Bsz]); e.in_lab.copy_(y_all[lo:lo + Bsz])
            evs[it][0].record()
            if e._graph is not None:
                e._graph.replay()
            else:
                e._launch_step()
            evs[it][1].record()
        sync_all()
        tt = torch.tensor([sum(a.elapsed_time(b) for a, b in evs)], dtype=torch.float64, device=dev)
        if world > 1:
            torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
        return float(tt.item())

    # ---------------- device-resident timing
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms_total = time_steps(eng, args.steps, args.warmup)
    loss_last = float(eng.loss.item())
    # kernels per step: counted from one eager (un-graphed) step: graph replays do not pass through the C ABI
    eng_launch0 = _lib.launch_count()
    eng.skip_update = False
    lo, hi = batch(0)
    eng.in_feat.copy_(f_all[lo:hi]); eng.in_lab.copy_(y_all[lo:hi])
    eng._launch_step()
    torch.cuda.synchronize(dev)
    launches_per_step = _lib.launch_count() - eng_launch0

    # ---------------- e2e: pinned host batches in, loss out, every step
    sync_all()
    t0 = time.perf_counter()
    for it in range(args.steps):
        lo, hi = batch(args.warmup + it)
        loss = eng.train_step(f_host[lo:hi], y_host[lo:hi])
        _ = float(loss.item())
    sync_all()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    e2e_s = float(t.item())
    # the same through the engine's host-batch API (next batch's H2D on a copy stream, asynchronous loss read-back into pinned memory,
    # one synchronisation at the end): every step still copies its own inputs from pinned host memory and its loss back to the host
    eng.train_steps_host(f_host, y_host, [batch(args.warmup + it) for it in range(min(3, args.steps))])
    sync_all()
    t0 = time.perf_counter()
    losses_host = eng.train_steps_host(f_host, y_host, [batch(args.warmup + it) for it in range(args.steps)])
    sync_all()
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    e2e_pipe_s = float(t.item())
    assert losses_host.numel() == args.steps and bool(torch.isfinite(losses_host).all())
    clk = clocks.stop() if rank == 0 else None

    # ---------------- per-kernel times (CUDA events on the launching stream, eager launches, L2 flushed)
    ktimes = profile_step_kernels(eng, f_all, y_all, shp, flush, reps=5)

    # ---------------- the same step at the other GEMM precisions (short runs)
    train_variants = {}
    if not args.no_variants:
        for prec in ("fp32", "tf32", "bf16x3", "bf16"):
            if prec == args.precision:
                train_variants[prec] = {"ms_per_step": ms_total / args.steps, "steps_per_s": args.steps / (ms_total * 1e-3)}
                continue
            ev = GPAdapterEngine(gpw, engine_config(precision=prec))
            msv = time_steps(ev, 10, 3)
            train_variants[prec] = {"ms_per_step": msv / 10, "steps_per_s": world * 10 / (msv * 1e-3)}
            ev.close_peer()
            del ev

    # ---------------- N > 1: the north star's S-sharded form of ONE 128-image batch (strong scaling of a latency-bound step)
    train_sample_sharded = None
    if world > 1 and S >= world:
        es = GPAdapterEngine(gpw, engine_config(shard="samples"))
        mss = time_steps(es, 10, 3) / 10
        train_sample_sharded = {"ms_per_step": mss, "steps_per_s": 1e3 / mss, "scaling": "strong",
                                "split": [cdist_sample_split(S, r, world)[1] for r in range(world)],
                                "note": "MC samples of one batch sharded over the ranks (same Philox stream), ONE all-reduce of the flat gradient "
                                        "buffer; the per-class GP chain is replicated on every rank, so this form cannot beat one GPU"}
        es.close_peer()
        del es

    # ---------------- full-batch leg (B = N_train = C * shots): the tensor-bound form of the same step
    fullbatch = None
    if not args.no_fullbatch and args.precision != "fp32":
        Btot = f_all.shape[0]
        Bf = Btot // world                                   # N > 1: the 16 000 cached training features are split over the ranks
        fullbatch = {"B": Btot, "B_per_rank": Bf, "scaling": "strong",
                     "note": "per-sample MC cross-entropy over the whole cached training set (batch rows split over the ranks, gradient "
                             "all-reduce); logits [B, S*C] materialised in fp32"}
        for prec in ("bf16", "bf16x3", "tf32"):
            ef = GPAdapterEngine(gpw, engine_config(batch_size=Bf, precision=prec))
            msf = time_steps(ef, 5, 3) / 5
            kt = profile_step_kernels(ef, f_all[rank * Bf:(rank + 1) * Bf], y_all[rank * Bf:(rank + 1) * Bf], shp, flush, reps=2, B=Bf)
            entry = {"ms_per_step": msf, "steps_per_s": 1e3 / msf, "img_per_s": Btot * 1e3 / msf}
            gem = {k: v for k, v in kt.items() if k.startswith("tc_gemm_") and v.get("flops")}
            if gem:
                # the three logit contractions: logits, d P_hat, d f_hat  (2*B*S*C*D algorithmic flops each)
                big = sorted(gem.items(), key=lambda kv: -kv[1]["flops"])[:3]
                fl = sum(v["flops"] for _, v in big); tm = sum(v["ms"] for _, v in big) * 1e-3
                pk_ = peaks()
                entry["logit_gemms"] = {"bound": "tensor", "flops": fl, "ms": tm * 1e3, "achieved": fl / tm / 1e12, "unit": "TFLOP/s",
                                        "peak": pk_["bf16_tflops"], "frac": fl / tm / 1e12 / pk_["bf16_tflops"],
                                        "hw_flops_factor": 3 if prec == "bf16x3" else 1,
                                        "tf32_peak_note": ("kind::tf32 runs at half the bf16 rate: frac of the TF32 peak = 2 x frac; measured library "
                                                           "TF32 peak (torch.matmul allow_tf32, 8192^3) 687 TFLOP/s, profiles/r2_tc_gemm_bench.txt") if prec == "tf32" else None,
                                        "per_gemm": {k: {"ms": round(v["ms"], 4), "tflops": round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 1)} for k, v in big}}
                entry["kernel_ms_per_step"] = {k: round(v["ms"], 4) for k, v in kt.items()}
            fullbatch[prec] = entry
            ef.close_peer()
            del ef
        torch.cuda.empty_cache()

    # ---------------- eval leg (on the INITIAL parameters: a fresh engine, so that accuracy / ECE / AECE do not depend on how many
    # optimisation steps the timing loops above happened to run, and are identical for every GPU count): MC-averaged logits + acc/ECE/AECE over this rank's shard of the test features
    eng_train = eng
    eng = GPAdapterEngine(gpw, engine_config())
    n_eval = args.eval_n
    f_te, y_te = wl["f_test"][:n_eval], wl["y_test"][:n_eval]
    from clip_gp_b200 import dist as cdist
    lo_e, hi_e = cdist.shard_range(n_eval, rank, world)
    sl = slice(lo_e, hi_e)
    f_sh, y_sh = f_te[sl].to(dev), y_te[sl].to(dev)
    def time_eval(fn, reps=5):
        for _ in range(2):
            out = fn()
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tot = 0.0
        for _ in range(reps):
            flush.fill_(0.0)
            e0.record()
            out = fn()
            e1.record()
            torch.cuda.synchronize(dev)
            tot += e0.elapsed_time(e1)
        tt = torch.tensor([tot / reps], dtype=torch.float64, device=dev)
        if world > 1:
            torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
        return float(tt.item()), out

    def eval_fp32():
        logits = eng.eval_logits(f_sh)
        return metrics.calibration_pass(logits, y_sh, 10)

    # headline eval: tcgen05 GEMMs with split-bf16 operands (fp32-grade products, >= the reference's TF32), collapsed MC form
    # the whole pass (GP forward -> prototypes | cast -> projection -> normalise | logits + calibration) is one captured CUDA graph over
    # the rank's resident test shard, as a trainer that evaluates after every step (adapter.py:363-380) would hold it
    eprec = args.precision if args.precision != "fp32" else "tf32"
    eval_replay = eng.eval_graph(f_sh, y_sh, precision=eprec, mc="collapsed")
    # the whole metric -- logits + accuracy + ECE histogram, counter all-reduce, (conf, hit) all-gather, AECE rank-select -- as ONE graph
    eval_pass = eng.eval_metrics_graph(f_sh, y_sh, n_eval, precision=eprec, mc="collapsed")
    eval_ms, _ = time_eval(eval_pass)
    eval_graph_only_ms, _ = time_eval(eval_replay)
    # weak-scaling eval (N > 1): every rank holds a FULL n_eval-image set (the test set rotated by the rank's shard offset, so the
    # ranks' images differ), the metric is global over world * n_eval images: same graph, counters all-reduced, (conf, hit) gathered,
    # AECE rank-select over the global set.  50 000 images are 0.3 ms of work for ONE B200; this is the regime 8 of them are for.
    eval_weak = None
    if world > 1:
        f_wk, y_wk = torch.roll(f_te, -lo_e, 0).to(dev), torch.roll(y_te, -lo_e, 0).to(dev)
        weak_pass = eng.eval_metrics_graph(f_wk, y_wk, n_eval * world, precision=eprec, mc="collapsed")
        weak_ms, wout = time_eval(weak_pass)
        cnt_w = metrics.counters_from_hist(wout[2], n_eval * world)
        eval_weak = {"images_total": n_eval * world, "images_per_rank": n_eval, "ms": weak_ms, "img_per_s": n_eval * world / (weak_ms * 1e-3),
                     "top1_acc": cnt_w.top1 * 100.0 / max(1, n_eval * world), "ece": metrics.ece_from_counters(cnt_w)[0],
                     "aece": metrics.aece_from_bins(wout[3], n_eval * world, 10)[0], "scaling": "weak"}
        weak_pass.release()
        del weak_pass, f_wk, y_wk

    # e2e eval: pinned HOST features of this rank's shard in, metrics out (H2D of the shard + D2H of counters inside the timed region)
    f_host_sh, y_host_sh = f_te[sl].contiguous().pin_memory(), y_te[sl].contiguous().pin_memory()
    f_static, y_static = eval_pass.inputs

    def eval_e2e_once():
        f_static.copy_(f_host_sh, non_blocking=True)
        y_static.copy_(y_host_sh, non_blocking=True)
        _, _, hg, aout = eval_pass()
        cnt_ = metrics.counters_from_hist(hg, n_eval)                      # D2H: 32 integer counters
        return cnt_, metrics.ece_from_counters(cnt_)[0], metrics.aece_from_bins(aout, n_eval, 10)[0]

    for _ in range(2):
        eval_e2e_once()
    sync_all()
    t0 = time.perf_counter()
    E2E_REPS = 5
    for _ in range(E2E_REPS):
        eval_e2e_once()
    sync_all()
    tt = torch.tensor([(time.perf_counter() - t0) / E2E_REPS], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
    eval_e2e_s = float(tt.item())
    cnt, ece, aece = eval_e2e_once()                           # the reported metrics: ONE pass (counters and AECE bins of the same draw)
    eval_variants = {}
    tf32_replay = eng.eval_graph(f_sh, y_sh, precision="bf16x3", mc="collapsed")
    for nm, fn in (("bf16x3_collapsed_graph", tf32_replay),
                   ("tf32_collapsed_eager_launches", lambda: eng.eval_calibration_tc(f_sh, y_sh, precision="tf32", mc="collapsed")),
                   ("bf16x3_collapsed_eager_launches", lambda: eng.eval_calibration_tc(f_sh, y_sh, precision="bf16x3", mc="collapsed")),
                   ("fp32_ffma_collapsed", eval_fp32),
                   ("bf16_collapsed", lambda: eng.eval_calibration_tc(f_sh, y_sh, precision="bf16", mc="collapsed")),
                   ("bf16_materialised", lambda: eng.eval_calibration_tc(f_sh, y_sh, precision="bf16", mc="materialised")),
                   ("bf16x3_materialised", lambda: eng.eval_calibration_tc(f_sh, y_sh, precision="bf16x3", mc="materialised"))):
        ms_v, _ = time_eval(fn, reps=3)
        eval_variants[nm] = {"ms": ms_v, "img_per_s": n_eval / (ms_v * 1e-3)}
    # the fused logit GEMM alone, materialised MC form (north star: [N,D] x [S*C,D]^T), against the tensor peak
    from clip_gp_b200 import tc as _tc
    Bop, mcs = eng.eval_operands_tc(S, "bf16", "materialised")
    fb = _tc.cast_bf16(f_sh)
    gemm_ms, _ = time_eval(lambda: _tc.logits_calibration(fb, Bop, 100.0 * mcs, y_sh, 10), reps=5)
    gemm_flops = 2.0 * f_sh.shape[0] * shp.C * shp.D * S
    # global metrics: all-reduce only the integer counters; AECE needs the gathered confidences
    # calibration check on a set whose bins err in both directions (on the plain synthetic set every bin is over-confident and
    # ECE == AECE to rounding, so an AECE rank-edge bug could hide): rank 0, single pass, small
    calib_check = None
    if rank == 0:
        with torch.no_grad():
            Pm0 = torch.nn.functional.normalize(wl["E"].mean(1), dim=-1)
        f_mx, y_mx = synth.make_mixed_calibration_set(wl["mu"], Pm0, 20000, 7, shp.noise)
        eng1 = GPAdapterEngine(gpw, engine_config(world=1, rank=0))
        r_mx = eng1.evaluate(f_mx.to(dev), y_mx.to(dev), precision=eprec)
        calib_check = {"n_images": 20000, "top1_acc": r_mx["top1_acc"], "ece": r_mx["ece"], "aece": r_mx["aece"],
                       "bin_gap_signs": [int((a_ > c_) - (a_ < c_)) for a_, c_ in zip(r_mx["calibration"]["bin_acc"], r_mx["calibration"]["bin_conf"])],
                       "set": "synth.make_mixed_calibration_set: half over-confident noisy features, half under-confident 4-way ambiguous features"}
        del eng1

    def shutdown():
        # captured graphs hold NCCL work: release them before the process group goes away
        eval_replay.release()
        eval_pass.release()
        tf32_replay.release()
        eng.close_peer()
        eng_train.check_peer_status()
        eng_train.close_peer()
        import gc
        gc.collect()
        torch.cuda.synchronize(dev)
        if world > 1:
            torch.distributed.barrier()
            torch.distributed.destroy_process_group()

    if rank != 0:
        shutdown()
        return
    pk = peaks()
    steps_per_s = world * args.steps / (ms_total * 1e-3)      # 128-image batch-steps per second over all ranks
    # dominant kernel of the step -> roofline
    own = {k: v for k, v in ktimes.items() if not k.startswith("nccl_")}       # collectives are not this repo's kernels
    dom = max(own, key=lambda k: own[k]["ms"]) if own else None
    roof = None
    if dom is not None:
        base = dom.split("(")[0]
        ab = algorithmic_bytes(shp, eng_train.S_local, base, eng_train)
        dur = ktimes[dom]["ms"] * 1e-3 / max(1, ktimes[dom]["calls"])
        if ab is not None:
            ach = ab / dur / 1e9
            roof = {"kernel": dom, "bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
                    "traffic": ncu_traffic(base), "traffic_source": "profiles/" + TRAFFIC_JSON + " (ncu --set full, DRAM read + write per launch)",
                    "algorithmic_bytes": ab, "avg_launch_us": dur * 1e6, "share_of_step": ktimes[dom]["ms"] / sum(v["ms"] for v in ktimes.values()),
                    "peak_source": pk["source"]}
            try:        # what actually limits the kernel (committed ncu capture): issue slots and one class's dependent chain, not bytes
                prof = json.load(open(os.path.join(ROOT, "profiles", TRAFFIC_JSON))).get(base, {})
                if "issue_active_pct" in prof:
                    roof["limiter"] = {"issue_slots_active_pct": prof["issue_active_pct"], "warps_active_pct": prof["warps_active_pct"],
                                       "warp_instructions_M": prof["inst_executed_M"],
                                       "note": "per-class dependent chain (Cholesky factorisations / adjoints, triangular sweeps) at 7 classes per SM; "
                                               "per-phase cycles in profiles/r2_gp_hotspots.txt / r1_gp_phase_cycles.txt"}
            except Exception:
                pass
        else:
            C_, D_, B_ = shp.C, shp.D, shp.B
            fl = ktimes[dom].get("flops") or 2.0 * B_ * eng.S_local * C_ * D_
            roof = {"kernel": dom, "bound": "tensor", "achieved": fl / dur / 1e12, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                    "frac": fl / dur / 1e12 / pk["bf16_tflops"], "traffic": None, "avg_launch_us": dur * 1e6,
                    "share_of_step": ktimes[dom]["ms"] / sum(v["ms"] for v in ktimes.values()), "peak_source": pk["source"],
                    "note": ("tcgen05 GEMM, algorithmic flops (split operands issue 3x the MMAs)" if "tc_gemm" in dom else
                             "fp32 FFMA GEMM (exact mode) reported against the bf16 tensor peak")}
    line = {
        "metric": METRIC, "value": steps_per_s, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": {"fp32": "f32", "tf32": "tf32 (fp32 operands read in place, 10-bit mantissa products, fp32 accumulate)",
                                      "bf16x3": "bf16x3 (split bf16 operands, fp32 accumulate)", "bf16": "bf16"}[args.precision],
        "data": "synthetic",
        "config": {"workload": workload_string(args.workload, shp, S),
                   "global_batch": world * shp.B, "parallelism": f"dp{world}" if world > 1 else "single GPU",
                   "value_counts": "128-image batch-steps per second summed over the ranks (one optimisation step consumes one batch per rank)",
                   "l2_flush": "256 MB device buffer written between timed steps", "precision": {"fp32": "fp32 (FFMA GEMMs, fp64 K_ZZ Cholesky)",
                                 "tf32": "tcgen05 kind::tf32 GEMMs on the fp32 tensors in place (no operand casts; the reference's GPU arithmetic, adapter.py:23), fp32 everywhere else, fp64 K_ZZ Cholesky",
                                 "bf16x3": "tcgen05 GEMMs on split-bf16 operands (fp32-grade products; the reference's GPU path is TF32), fp32 everywhere else, fp64 K_ZZ Cholesky",
                                 "bf16": "tcgen05 bf16 GEMMs, fp32 everywhere else, fp64 K_ZZ Cholesky"}[args.precision],
                   "cuda_graph": eng_train._graph is not None, "two_stream_overlap": bool(eng_train.cfg.overlap),
                   "multi_gpu": (f"data parallel over the batch: {world} ranks x {shp.B} images, all {S} MC samples on every rank (same Philox draw); "
                                 + ("gradient reduce-scatter + AdamW on the owned 1/world slice + parameter all-gather in ONE kernel over NVLink peer "
                                    "memory (CUDA IPC pointers, csrc/peer.cu), inside the step's CUDA graph" if eng_train.peer is not None else
                                    "ONE ncclAllReduce of the flat gradient buffer + loss (7.6 MB) + replicated AdamW; step captured in a CUDA graph incl. NCCL"))
                   if world > 1 else None,
                   "loss_last": loss_last},
        "e2e": {"value": world * args.steps / e2e_pipe_s, "unit": UNIT, "h2d_bytes_per_step": shp.B * shp.D * 4 + shp.B * 8, "d2h_bytes_per_step": 4,
                "api": "GPAdapterEngine.train_steps_host: per step H2D of the batch from pinned host memory (copy stream, double buffered) + "
                       "asynchronous D2H of the loss into pinned memory, one host synchronisation at the end of the timed region",
                "synchronous_value": world * args.steps / e2e_s,
                "synchronous_api": "GPAdapterEngine.train_step(host batch) + loss.item() every step (blocking read-back, as the reference logs)"},
        "gpu_launches": int(launches_per_step) * args.steps,
        "gpu_launches_per_step": int(launches_per_step),
        "clocks": clk,
        "roofline": roof,
        "kernel_ms_per_step": {k: round(v["ms"], 4) for k, v in ktimes.items()},
        "train_variants": train_variants,
        "train_sample_sharded": train_sample_sharded,
        "train_fullbatch": fullbatch,
        "eval": {"metric": "eval_img_per_s (projection + MC-averaged logits + accuracy + ECE histogram + AECE rank-select, device resident)",
                 "value": n_eval / (eval_ms * 1e-3), "unit": "img/s", "n_images": n_eval, "ms": eval_ms, "S_eval": S, "scaling": "strong",
                 "graph_only_ms": eval_graph_only_ms,
                 "e2e": {"value": n_eval / eval_e2e_s, "unit": "img/s", "ms": eval_e2e_s * 1e3,
                         "h2d_bytes_per_pass": int(f_host_sh.numel() * 4 + y_host_sh.numel() * 8), "d2h_bytes_per_pass": 32 * 8 + 10 * 3 * 8,
                         "api": "pinned host features of the rank's shard -> H2D -> eval graph replay -> counter all-reduce -> AECE select -> "
                                "D2H of the counters (accuracy / ECE / AECE on the host)"},
                 "top1_acc": cnt.top1 * 100.0 / n_eval, "ece": ece, "aece": aece, "calibration_check": calib_check,
                 "multi_gpu": (f"images sharded over {world} ranks; the eval GP forward runs on C/{world} classes per rank and one 2 MB "
                               "all-reduce completes the mean prototypes; counters: one all-reduce, AECE: all-gather of (conf, hit)") if world > 1 else None,
                 "form": "one CUDA graph (engine.eval_metrics_graph; graph_only_ms = engine.eval_graph without the AECE / collective tail): GP forward + prototypes on a side stream next to cast / projection / normalise, then ONE tcgen05 GEMM over the raw features against [W ; mean_s p_hat_s W] (projection, normalisation, collapsed logit-mean and the calibration epilogue fused), precision = " + eprec,
                 "weak": eval_weak,
                 "variants": eval_variants},
        "roofline_eval_gemm": {"kernel": "tc_gemm_kernel (EPI_ROWSTATS), materialised MC logits [N,D]x[S*C,D]^T accumulated over s in TMEM",
                               "bound": "tensor", "achieved": gemm_flops / (gemm_ms * 1e-3) / 1e12, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                               "frac": gemm_flops / (gemm_ms * 1e-3) / 1e12 / pk["bf16_tflops"], "traffic": None, "avg_launch_us": gemm_ms * 1e3,
                               "flops_per_launch": gemm_flops, "peak_source": pk["source"]},
    }
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline(wl, shp, S, ls)
        line["torch_eager_gpu"] = torch_eager_gpu(wl, shp, S, ls, dev)
    print(json.dumps(line), flush=True)
    if args.profile_kernels:
        for k, v in sorted(ktimes.items(), key=lambda kv: -kv[1]["ms"]):
            print(f"# {k:28s} {v['ms']*1e3:9.1f} us/step  ({v['calls']} launches)", file=sys.stderr)
    shutdown()


def profile_step_kernels(eng, f_all, y_all, shp, flush, reps=5, B=None):
    """Average device time of every kernel of one step, from CUDA events around each C-ABI launch (eager mode)."""
    from clip_gp_b200 import _lib
    lib = eng.lib
    names = ["clipgp_tc_gemm_tf32", "clipgp_gemm_f32", "clipgp_rownorm_forward", "clipgp_gp_forward", "clipgp_proto_forward", "clipgp_softmax_ce",
             "clipgp_rownorm_backward", "clipgp_l2_identity", "clipgp_proto_backward", "clipgp_gp_backward", "clipgp_sum_accumulate",
             "clipgp_adamw_step", "clipgp_adamw_step_lrptr", "clipgp_increment", "clipgp_tc_gemm_store", "clipgp_tc_gemm_store_splitk", "clipgp_cast_bf16", "clipgp_cast_bf16_transpose", "clipgp_cast_bf16_dual",
             "clipgp_softmax_ce_stats", "clipgp_softmax_grad_bf16_dual", "clipgp_softmax_ce_bf16_dual", "clipgp_increment2", "clipgp_step_epilogue",
             "clipgp_peer_adamw", "clipgp_transpose_f32", "clipgp_adamw_tail"]
    multi = ("tc_gemm_tf32", "gemm_f32", "transpose_f32", "adamw_step", "adamw_step_lrptr", "increment", "tc_gemm_store", "tc_gemm_store_splitk", "cast_bf16", "cast_bf16_transpose", "cast_bf16_dual")
    seg = getattr(eng, "tc_seg", 1)
    B = B or shp.B
    records = []

    class Wrap:
        def __init__(self, fn, name):
            self.fn, self.name = fn, name

        def __call__(self, *a):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = self.fn(*a)
            e1.record()
            fl = 2.0 * a[1] * a[4] * a[5] / seg if self.name.startswith("clipgp_tc_gemm_store") else None
            if self.name == "clipgp_tc_gemm_tf32":
                fl = 2.0 * a[2] * a[5] * a[6]
            records.append((self.name, e0, e1, fl))
            return rc

    class Proxy:
        def __getattr__(self, n):
            fn = getattr(lib, n)
            return Wrap(fn, n) if n in names else fn

    eng.lib = Proxy()
    real_allreduce = torch.distributed.all_reduce
    if eng.cfg.world > 1:
        def timed_allreduce(t, *a, **kw):                  # NCCL collectives of the step, timed like the kernels
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = real_allreduce(t, *a, **kw)
            e1.record()
            records.append((f"nccl_all_reduce[{t.numel() * t.element_size() / 1e6:.1f}MB]", e0, e1, None))
            return r
        torch.distributed.all_reduce = timed_allreduce
    out = {}
    overlap_was = eng.cfg.overlap
    eng.cfg.overlap = False          # serialise the two branches of the step so that the per-kernel events do not overlap
    try:
        for r in range(reps):
            records.clear()
            flush.fill_(0.0)
            eng.in_feat.copy_(f_all[:B]); eng.in_lab.copy_(y_all[:B])
            eng._launch_step()
            torch.cuda.synchronize()
            seen = {}
            for name, e0, e1, fl in records:
                short = name.replace("clipgp_", "")
                k = seen.get(short, 0); seen[short] = k + 1
                key = f"{short}({k})" if (short in multi or short.startswith("nccl_")) else short
                d = out.setdefault(key, {"ms": 0.0, "calls": 0})
                d["ms"] += e0.elapsed_time(e1) / reps
                d["calls"] = 1
                if fl: d["flops"] = fl
    finally:
        eng.lib = lib
        eng.cfg.overlap = overlap_was
        torch.distributed.all_reduce = real_allreduce
    return out


def cpu_baseline(wl, shp, S, ls, budget_s=20.0):
    from oracle import gp as ogp
    from oracle.train_step import OracleAdapter
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    st = ogp.build_state(wl["E"], shp.kernel, shp.d, lengthscale=ls)
    g = torch.Generator().manual_seed(1)
    st.var_mean = 1e-3 * torch.randn(shp.C, shp.T + 1, generator=g)
    orc = OracleAdapter(st, shp.D, shots=shp.shots)
    f, y = wl["f_train"], wl["y_train"]
    orc.step(f[: shp.B], y[: shp.B], torch.randn(shp.C, shp.T, S, generator=g))   # warm-up
    n, t0 = 0, time.perf_counter()
    while n < 3 or (time.perf_counter() - t0 < budget_s and n < 50):
        lo = (n % (f.shape[0] // shp.B)) * shp.B
        orc.step(f[lo:lo + shp.B], y[lo:lo + shp.B], torch.randn(shp.C, shp.T, S, generator=g))
        n += 1
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n} full steps (fwd + autograd bwd + AdamW) of the same {shp.name} workload on the host CPU, torch {torch.__version__}"}


def torch_eager_gpu(wl, shp, S, ls, dev, steps=20):
    """Second comparator of BASELINE.md section 4: the restated reference step (oracle: torch ops + autograd + torch AdamW) run
    eagerly by torch ON THE SAME B200, TF32 matmuls allowed as in the reference (adapter.py:23) -- i.e. the reference's GPU path
    without gpytorch's Python overhead.  Reported next to the CPU baseline; not the product path."""
    try:
        from oracle import gp as ogp
        from oracle.train_step import OracleAdapter, state_to_device
        torch.backends.cuda.matmul.allow_tf32 = True
        st = state_to_device(ogp.build_state(wl["E"], shp.kernel, shp.d, lengthscale=ls), dev)
        g = torch.Generator(device=dev).manual_seed(1)
        st.var_mean = 1e-3 * torch.randn(shp.C, shp.T + 1, generator=g, device=dev)
        orc = OracleAdapter(st, shp.D, shots=shp.shots)
        f, y = wl["f_train"].to(dev), wl["y_train"].to(dev)
        nb = f.shape[0] // shp.B

        def one(it):
            lo = (it % nb) * shp.B
            return orc.step(f[lo:lo + shp.B], y[lo:lo + shp.B], torch.randn(shp.C, shp.T, S, generator=g, device=dev))

        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # eval leg first (initial parameters, like our own eval leg): sample prototypes -> [N,S,C] logits -> mean over s
        # (adapter.py:243-249) in image chunks of 128 as the reference's test loader does (default.yaml:5), then accuracy / ECE / AECE
        # with the restated utils/metrics.py
        from oracle import metrics as om
        ft, yt = wl["f_test"].to(dev), wl["y_test"].to(dev)
        n_img = min(ft.shape[0], 12800)

        def eval_pass():
            with torch.no_grad():
                chunks = []
                for lo in range(0, n_img, 128):
                    eps = torch.randn(shp.C, shp.T, S, generator=g, device=dev)
                    chunks.append(orc.eval_logits(ft[lo:lo + 128], eps))          # the reference re-samples per forward_features call
                lg = torch.cat(chunks)
                return om.compute_accuracy(lg, yt[:n_img])[0], om.compute_ece(lg, yt[:n_img]), om.compute_aece(lg, yt[:n_img])

        eval_pass()
        torch.cuda.synchronize(dev)
        e0.record()
        res = eval_pass()
        e1.record()
        torch.cuda.synchronize(dev)
        ev = {"value": n_img / (e0.elapsed_time(e1) * 1e-3), "unit": "img/s", "n_images": n_img,
              "note": "GP forward + prototypes re-sampled for every 128-image batch as CustomCLIP.forward_features does; "
                      "utils/metrics.py restatement for accuracy / ECE / AECE", "top1_acc": res[0], "ece": res[1]}
        for it in range(3):
            one(it)
        torch.cuda.synchronize(dev)
        e0.record()
        for it in range(steps):
            one(3 + it)
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / steps
        out = {"value": 1e3 / ms, "unit": UNIT, "ms_per_step": ms, "kind": "port",
               "sample": f"{steps} steps of the oracle restatement of the reference step, torch {torch.__version__} eager on the same GPU, "
                         "TF32 matmuls, fp64 Cholesky of K_ZZ via cuSOLVER (loss read back every step, as the reference logs it)",
               "eval": ev}
        return out
    except Exception as e:  # pragma: no cover
        return {"unavailable": f"{type(e).__name__}: {e}"[:200]}


if __name__ == "__main__":
    a = parse()
    # stdout carries exactly ONE JSON line; everything else (build log, module prints) goes to stderr
    _real_stdout = sys.stdout
    sys.stdout = sys.stderr
    _print = print

    def print(*args, **kw):  # noqa: A001
        if "file" not in kw:
            kw["file"] = _real_stdout
        _print(*args, **kw)

    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
