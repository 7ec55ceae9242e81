import os, sys
sys.path.insert(0, "/root/repo")
import torch
from clip_gp_b200 import metrics as gm
g = torch.Generator().manual_seed(0)
N, C = 50000, 1000
logits = (3.0 * torch.randn(N, C, generator=g)).cuda()
y = torch.randint(0, C, (N,), generator=g).cuda()
for _ in range(3):
    gm.calibration_pass(logits, y, 10, want_conf=True)
torch.cuda.synchronize()
