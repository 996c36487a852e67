#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_teachers_gpu.py -q -s > gpurun_out/r2aa_t.log 2>&1; echo "teacher tests rc=$?"; grep "rel L2\|replay\|residual\|passed\|failed\|Error" gpurun_out/r2aa_t.log | head -50
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -k "gemm" 2>&1 | tail -2
for k in dinov2 clip vith; do timeout 600 python tools/bench_teacher.py $k 128 | tail -1; timeout 600 python tools/bench_teacher.py $k 128 --bf16-residual | tail -1; done
