#!/bin/bash
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "ln3d" 2>&1 | tail -2
timeout 200 python tools/bench_ln.py | grep ln3d
THEIA_B200_LIB=/root/repo/gpurun_in/lib_old.so timeout 200 python tools/bench_ln.py | grep ln3d
