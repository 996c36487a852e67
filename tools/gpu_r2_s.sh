#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "resize or preprocess" > gpurun_out/r2s_k.log 2>&1; echo "resize kernel tests rc=$?"; tail -5 gpurun_out/r2s_k.log
timeout 600 python -m pytest tests/test_model_gpu.py -x -q -k "any_image_extent or golden or readme" > gpurun_out/r2s_m.log 2>&1; echo "model tests rc=$?"; tail -5 gpurun_out/r2s_m.log
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2s_all.log 2>&1; echo "all gpu tests rc=$?"; tail -4 gpurun_out/r2s_all.log
