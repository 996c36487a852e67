#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "attention" > gpurun_out/r2o_attn.log 2>&1; echo "attn tests rc=$?"; tail -3 gpurun_out/r2o_attn.log
for i in 1 2 3; do timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "attention_fwd_long" 2>&1 | tail -1; done
timeout 120 python tools/bench_attn.py 128 16 257
timeout 600 python tools/bench_teacher.py dinov2 128 | tail -1
