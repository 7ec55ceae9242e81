#!/bin/bash
# Debug build with per-phase clock64() timestamps in the GP kernels (read by tools/gp_phase_ts.py)
set -e
cd "$(dirname "$0")/.."
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared -cudart shared \
    -DCLIPGP_PHASE_TS -o clip_gp_b200/lib/libclipgp_ts.so clip_gp_b200/csrc/*.cu 2>&1 | grep -i "error" || true
ls -la clip_gp_b200/lib/libclipgp_ts.so
