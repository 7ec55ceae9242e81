"""A/B two builds of the library on the same box: python tools/micro/ab_lib.py  (libclipgp.so vs libclipgp_ts.so)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
code = r'''
import sys, os
sys.path.insert(0, %r)
from clip_gp_b200 import _lib
if sys.argv[1] == "alt":
    _lib.LIB_PATH = os.path.join(%r, "clip_gp_b200", "lib", "libclipgp_ts.so")
import torch
from clip_gp_b200 import tc
dev = torch.device("cuda"); g = torch.Generator().manual_seed(0)
def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for M, N, K in ((16000, 10000, 512), (8192, 8192, 8192), (128, 10000, 512)):
    A = torch.randn(M, K, generator=g).to(dev); B = torch.randn(N, K, generator=g).to(dev); C = torch.empty(M, N, device=dev)
    ms = timeit(lambda: tc.gemm_tf32(A, B, 1.0, out=C))
    Ab, Bb = tc.cast_bf16(A), tc.cast_bf16(B)
    ms2 = timeit(lambda: tc.gemm(Ab, Bb, 1.0, out=C)) if hasattr(tc, "gemm") else float("nan")
    print(f"{sys.argv[1]:4s} M={M} N={N} K={K}: tf32 {ms*1e3:8.1f} us {2.0*M*N*K/ms/1e9:7.1f} TF | bf16 {ms2*1e3:8.1f} us")
''' % (ROOT, ROOT)
for rnd in range(2):
    for which in ("main", "alt"):
        print(subprocess.run([sys.executable, "-c", code, which], capture_output=True, text=True).stdout.strip())
