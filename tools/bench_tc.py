"""Micro-benchmark of the tcgen05 GEMM (CUDA events, L2 flushed between reps):  python tools/bench_tc.py"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from clip_gp_b200 import tc

dev = torch.device("cuda", 0)
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 1590.0

def timeit(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms = []
    for _ in range(reps):
        flush.fill_(0.0)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    ms.sort()
    return ms[len(ms) // 2]

g = torch.Generator().manual_seed(0)
cases = [("eval collapsed       ", 50000, 1000, 512, 512), ("eval materialised S10", 50000, 1000, 5120, 512),
         ("eval materialised S100", 50000, 1000, 51200, 512), ("full-batch logits    ", 16000, 10000, 512, 512),
         ("minibatch logits     ", 128, 10000, 512, 512), ("projection           ", 50000, 512, 512, 512),
         ("tip affinity eval    ", 50000, 16000, 1024, 1024), ("square 8192          ", 8192, 8192, 8192, 8192)]
for name, M, N, K, Ka in cases:
    A = torch.randn(M, Ka, generator=g).to(dev).to(torch.bfloat16)
    B = (torch.randn(N, min(K, 4096), generator=g).to(dev).to(torch.bfloat16)).repeat(1, (K + 4095) // 4096)[:, :K].contiguous()
    y = torch.randint(0, N, (M,), generator=g).to(dev)
    C = torch.empty(M, N, device=dev) if M * N * 4 < 8e9 else None
    fl = 2.0 * M * N * K
    if C is not None:
        ms = timeit(lambda: tc.gemm_store(A, B, 1.0, out=C))
        print(f"{name} store    M={M} N={N} K={K}: {ms*1e3:9.1f} us  {fl/ms/1e9:8.1f} TFLOP/s  ({fl/ms/1e9/peak*100:5.1f}% of {peak})")
    ms = timeit(lambda: tc.logits_calibration(A, B, 1.0, y, 10))
    print(f"{name} rowstats M={M} N={N} K={K}: {ms*1e3:9.1f} us  {fl/ms/1e9:8.1f} TFLOP/s  ({fl/ms/1e9/peak*100:5.1f}% of {peak})")
    del A, B, C
