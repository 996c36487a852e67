#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -k "attention or layernorm" > gpurun_out/r2x_k.log 2>&1; echo "attention/ln kernel tests rc=$?"; tail -6 gpurun_out/r2x_k.log
timeout 900 python -m pytest tests/test_teachers_gpu.py -q -s -k "arch6 or unsupported or arch0" > gpurun_out/r2x_t.log 2>&1; echo "teacher tests rc=$?"; grep "rel L2\|replay\|passed\|failed\|Error" gpurun_out/r2x_t.log | head -20
