#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "layernorm" 2>&1 | tail -2
timeout 120 python tools/bench_ln.py
timeout 600 python tools/bench_teacher.py dinov2 128 | tail -1
