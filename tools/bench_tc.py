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

# ---- TF32 on fp32 operands read in place (no cast kernels); the measured TF32 library peak next to it (torch.matmul, allow_tf32)
torch.backends.cuda.matmul.allow_tf32 = True
Xa, Xb = torch.randn(8192, 8192, device=dev), torch.randn(8192, 8192, device=dev)
ms_lib = timeit(lambda: torch.matmul(Xa, Xb.t()))
tf32_peak = 2.0 * 8192 ** 3 / ms_lib / 1e9
print(f"torch.matmul TF32 8192^3: {ms_lib*1e3:.1f} us = {tf32_peak:.1f} TFLOP/s (library TF32 peak used below)")
del Xa, Xb
tcases = [("full-batch logits   f_hat P_hat^T  ", 16000, 10000, 512, False, False),
          ("full-batch d f_hat  dlogits P_hat  ", 16000, 512, 10000, False, True),
          ("full-batch d P_hat  dlogits^T f_hat", 10000, 512, 16000, True, True),
          ("minibatch logits                   ", 128, 10000, 512, False, False),
          ("minibatch d f_hat                  ", 128, 512, 10000, False, True),
          ("minibatch d P_hat                  ", 10000, 512, 128, True, True),
          ("projection                         ", 50000, 512, 512, False, False),
          ("eval collapsed store               ", 50000, 1000, 512, False, False),
          ("square 8192                        ", 8192, 8192, 8192, False, False)]
for name, M, N, K, a_t, b_t in tcases:
    A = torch.randn((K, M) if a_t else (M, K), generator=g).to(dev)
    B = torch.randn((K, N) if b_t else (N, K), generator=g).to(dev)
    C = torch.empty(M, N, device=dev)
    fl = 2.0 * M * N * K
    for sk in (False, True):
        ms = timeit(lambda: tc.gemm_tf32(A, B, 1.0, a_t=a_t, b_t=b_t, out=C, split_k=sk))
        print(f"tf32 {name} M={M} N={N} K={K} a_t={int(a_t)} b_t={int(b_t)} split_k={int(sk)}: {ms*1e3:9.1f} us  {fl/ms/1e9:8.1f} TFLOP/s  ({fl/ms/1e9/tf32_peak*100:5.1f}% of TF32 {tf32_peak:.0f})")
    del A, B, C
