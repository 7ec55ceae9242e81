"""Which operand layout costs what: the two full-batch adjoint GEMM shapes in TF32 with every K-major / MN-major combination."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from clip_gp_b200 import tc
dev = torch.device("cuda")
g = torch.Generator().manual_seed(0)
def timeit(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for name, M, N, K in (("d f_hat", 16000, 512, 10000), ("d P_hat", 10000, 512, 16000), ("square", 8192, 8192, 8192), ("wide-N long-K", 16000, 2048, 10000)):
    for a_t in (False, True):
        for b_t in (False, True):
            A = torch.randn((K, M) if a_t else (M, K), generator=g).to(dev)
            B = torch.randn((K, N) if b_t else (N, K), generator=g).to(dev)
            C = torch.empty(M, N, device=dev)
            ms = timeit(lambda: tc.gemm_tf32(A, B, 1.0, a_t=a_t, b_t=b_t, out=C, split_k=True))
            print(f"{name:14s} M={M} N={N} K={K} A {'MN' if a_t else 'K '}-major, B {'MN' if b_t else 'K '}-major: {ms*1e3:8.1f} us {2.0*M*N*K/ms/1e9:7.1f} TFLOP/s")
            del A, B, C
