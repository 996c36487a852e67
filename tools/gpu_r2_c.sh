#!/bin/bash
# ncu --set full of the attention v2 kernels (one launch each), source-level sampling
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_tc -s 4 -c 2 -f -o gpurun_out/r2c_attn python tools/bench_attn.py 256 12 > gpurun_out/r2c_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r2c_ncu.log
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "resize or preprocess" -s 2>&1 | tail -5
