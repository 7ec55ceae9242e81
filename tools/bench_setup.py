"""Time the GP setup kernels at the ImageNet shape (C*T = 32 000 points, d = 256): median-heuristic length-scale by radix select
(csrc/setup.cu) against the reference formulation (torch.cdist chunks + masked median) on the same GPU."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F
from clip_gp_b200 import ops

N, d = 32000, 256
g = torch.Generator().manual_seed(0)
X = F.normalize(torch.randn(N, d, generator=g) + 2.0 * torch.randn(1, d, generator=g), dim=-1).cuda()
for _ in range(2):
    v = ops.median_pairwise_distance(X)
torch.cuda.synchronize(); t0 = time.perf_counter()
v = ops.median_pairwise_distance(X)
torch.cuda.synchronize(); t1 = time.perf_counter()
print(f"radix select   : {v:.7f}  {1e3 * (t1 - t0):8.1f} ms  (3 passes, {3 * N * N * d / 2 / (t1 - t0) / 1e12:.1f} TMAC/s, no N x N matrix)")
torch.cuda.synchronize(); t0 = time.perf_counter()
parts = []
for i in range(0, N, 4096):
    pdist = torch.cdist(X[i:i + 4096], X)
    parts.append(pdist[pdist > 0])
ref = torch.cat(parts).median().item()
torch.cuda.synchronize(); t1 = time.perf_counter()
print(f"torch reference: {ref:.7f}  {1e3 * (t1 - t0):8.1f} ms  (cdist + mask + median over {sum(p.numel() for p in parts):,} values, "
      f"{torch.cuda.max_memory_allocated() / 2**30:.1f} GiB peak)")
print(f"relative difference {abs(v - ref) / ref:.2e}")
