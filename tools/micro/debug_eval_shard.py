import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch, torch.distributed as td
import bench
from clip_gp_b200 import dist as cdist, synth
from clip_gp_b200.engine import EngineConfig, GPAdapterEngine
from clip_gp_b200.gp_template_weigher import GaussianProcessTemplateWeighter
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
td.init_process_group("nccl", device_id=dev)
wl = synth.make_workload("cfg2", n_test=8192); shp = wl["shape"]
ls = bench.bench_lengthscale(wl["E"], shp.d)
def make(world_, rank_, **kw):
    torch.manual_seed(1)
    gpw = GaussianProcessTemplateWeighter(wl["E"].to(dev), bench._Cfg(shp.kernel, shp.d), lengthscale=ls).to(dev)
    return GPAdapterEngine(gpw, EngineConfig(S_train=10, S_eval=10, batch_size=shp.B, shots=shp.shots, seed=77, rank=rank_, world=world_, precision="bf16x3", **kw))
eng, ref = make(world, rank), make(1, 0)
print(rank, "state equal:", torch.equal(eng.flat_p, ref.flat_p), torch.equal(eng.Z, ref.Z), torch.equal(eng.eval_rng_state, ref.eval_rng_state))
Pa, Pb = eng.eval_prototypes(), ref.eval_prototypes()
d = (Pa - Pb).abs()
print(rank, "Pm equal:", torch.equal(Pa, Pb), "max diff", float(d.max()), "rows differing", int((d.amax(1) > 0).sum()), "w equal on my classes:",
      torch.equal(eng.last_eval_w[:, rank * 500:(rank + 1) * 500], ref.last_eval_w[:, rank * 500:(rank + 1) * 500]))
ft, yt = wl["f_test"].to(dev), wl["y_test"].to(dev)
lo, hi = cdist.shard_range(8192, rank, world)
eng2, ref2 = make(world, rank, shard_eval_classes=False), make(1, 0)
c2, o2, h2 = eng2.eval_calibration_tc(ft[lo:hi], yt[lo:hi], precision="bf16x3", mc="collapsed")
c1, o1, h1 = ref2.eval_calibration_tc(ft, yt, precision="bf16x3", mc="collapsed")
print(rank, "unsharded-GP eval: conf equal on my rows:", torch.equal(c2, c1[lo:hi]), float((c2 - c1[lo:hi]).abs().max()))
td.destroy_process_group()
