"""The calibration pass (softmax max / arg-max + ECE histogram) over fp32 logits [50 000, 1000] a few times, for ncu:
    ncu --set full -k regex:calib_rows python tools/prof_calib.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clip_gp_b200 import metrics as gm
g = torch.Generator().manual_seed(0)
N, C = 50000, 1000
logits = (3.0 * torch.randn(N, C, generator=g)).cuda()
y = torch.randint(0, C, (N,), generator=g).cuda()
for _ in range(3):
    gm.calibration_pass(logits, y, 10, want_conf=True)
torch.cuda.synchronize()
