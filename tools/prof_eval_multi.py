"""Per-stage device times of the sharded eval pass (run under torchrun): GP forward (class shard) + prototype all-reduce, operand
casts, fused GEMM, counter all-reduce, (conf, hit) all-gather, AECE select.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/prof_eval_multi.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as td
import bench
from clip_gp_b200 import dist as cdist, metrics, synth, tc, _lib
from clip_gp_b200.engine import EngineConfig, GPAdapterEngine
from clip_gp_b200.gp_template_weigher import GaussianProcessTemplateWeighter

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    td.init_process_group("nccl", device_id=dev)
wl = synth.make_workload("cfg2"); shp = wl["shape"]
torch.manual_seed(1)
gpw = GaussianProcessTemplateWeighter(wl["E"].to(dev), bench._Cfg(shp.kernel, shp.d), lengthscale=1.4).to(dev)
eng = GPAdapterEngine(gpw, EngineConfig(S_train=10, S_eval=10, batch_size=128, shots=16, seed=1, rank=rank, world=world, precision="bf16x3", shard="batch"))
n = 50000
lo, hi = cdist.shard_range(n, rank, world)
f, y = wl["f_test"][lo:hi].to(dev), wl["y_test"][lo:hi].to(dev)
flush = torch.empty(64 * 1024 * 1024, device=dev)


def timeit(name, fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        flush.fill_(0.0)
        if world > 1:
            td.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    t = torch.tensor([tot / reps], device=dev, dtype=torch.float64)
    if world > 1:
        td.all_reduce(t, op=td.ReduceOp.MAX)
    if rank == 0:
        print(f"{name:55s} {float(t) * 1e3:8.1f} us", flush=True)


timeit("eval_prototypes (GP fwd shard + proto all-reduce)", lambda: eng.eval_prototypes(10))
eng.cfg.shard_eval_classes = False
timeit("eval_prototypes (GP fwd all classes, no collective)", lambda: eng.eval_prototypes(10))
eng.cfg.shard_eval_classes = True
Pm = torch.zeros(1000, 512, device=dev)
if world > 1:
    timeit("all_reduce [C,D] fp32 2 MB (eager)", lambda: td.all_reduce(Pm))
timeit("cast features -> split bf16", lambda: tc.cast_bf16(f, tc.SPLIT_A))
timeit("eval_calibration_tc eager (whole pass, no AECE)", lambda: eng.eval_calibration_tc(f, y, precision="bf16x3", mc="collapsed"))
rp = eng.eval_graph(f, y, precision="bf16x3", mc="collapsed")
timeit("eval_graph replay", rp)
conf, correct, hist = rp()
if world > 1:
    timeit("global_calibration (counter all-reduce + 2 all-gathers)", lambda: cdist.global_calibration(hist, conf, correct, n, world))
hg, cg, og = cdist.global_calibration(hist, conf, correct, n, world)
timeit("aece_pass over the global 50k (eager)", lambda: metrics.aece_pass(cg, og, 10))
mp = eng.eval_metrics_graph(f, y, n, precision="bf16x3", mc="collapsed")
timeit("eval_metrics_graph replay (whole metric)", mp)
rp.release(); mp.release()
if world > 1:
    td.barrier(); td.destroy_process_group()
