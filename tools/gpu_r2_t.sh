#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -k "resize or preprocess" > gpurun_out/r2t_k.log 2>&1; echo "resize kernel tests rc=$?"; tail -12 gpurun_out/r2t_k.log
