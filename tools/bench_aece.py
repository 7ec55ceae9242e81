import os, sys, torch
sys.path.insert(0, "/root/repo")
from clip_gp_b200 import metrics
torch.manual_seed(0)
for N in (6250, 50000, 200000):
    conf = torch.rand(N, device="cuda") * 0.5 + 0.5
    cor = (torch.rand(N, device="cuda") < conf).to(torch.uint8)
    for _ in range(3): metrics.aece_pass(conf, cor, 10)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): _, out = metrics.aece_pass(conf, cor, 10)
    e1.record(); torch.cuda.synchronize()
    print(os.environ.get("CLIPGP_AECE_BLOCKS", "auto"), N, f"{e0.elapsed_time(e1)/20*1e3:.1f} us", out[2].tolist()[:3])
