#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -k "attention" > gpurun_out/r2y_k.log 2>&1; echo "attention kernel tests rc=$?"; tail -4 gpurun_out/r2y_k.log
timeout 900 python -m pytest tests/test_teachers_gpu.py -q > gpurun_out/r2y_t.log 2>&1; echo "teacher tests rc=$?"; tail -3 gpurun_out/r2y_t.log
timeout 600 python tools/bench_teacher.py vith 128 --hf | tail -1
