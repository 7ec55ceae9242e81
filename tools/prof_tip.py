"""The fused Tip-Adapter evaluation GEMM (50 000 images x 16 000 keys, bf16) a few times, for ncu:  python tools/prof_tip.py [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F
from clip_gp_b200 import _lib, tc

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
N, N_tr, C, D = 50000, 16000, 1000, 1024
mu = torch.randn(C, D, generator=g)
lab = torch.arange(C).repeat_interleave(16)
keys = F.normalize(mu[lab] + 2.0 * torch.randn(N_tr, D, generator=g), dim=-1).to(dev)
yt = torch.randint(0, C, (N,), generator=g)
ft = F.normalize(mu[yt] + 2.0 * torch.randn(N, D, generator=g), dim=-1).to(dev)
fb, kb, li = tc.cast_bf16(ft), tc.cast_bf16(keys), lab.to(torch.int32).to(dev).contiguous()
out = torch.zeros(N, C, device=dev)
lib = _lib.load()
for _ in range(reps):
    _lib.check(lib.clipgp_tc_tip_logits(fb.data_ptr(), N, kb.data_ptr(), N_tr, D, li.data_ptr(), 2.0, 20.0, out.data_ptr(), C,
                                        _lib.stream_ptr(dev)), "clipgp_tc_tip_logits")
torch.cuda.synchronize()
print("checksum", float(out.sum()))
