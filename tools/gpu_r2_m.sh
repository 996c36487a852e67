#!/bin/bash
# teacher path bring-up: long-sequence attention, teacher parity, throughput
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "attention" > gpurun_out/r2m_attn.log 2>&1; echo "attn tests rc=$?"; tail -4 gpurun_out/r2m_attn.log
timeout 900 python -m pytest tests/test_teachers_gpu.py -q -s > gpurun_out/r2m_teach.log 2>&1; echo "teacher tests rc=$?"; tail -15 gpurun_out/r2m_teach.log
timeout 600 python tools/bench_teacher.py dinov2 128 --hf > gpurun_out/r2m_bench_dinov2.log 2>&1; tail -2 gpurun_out/r2m_bench_dinov2.log
timeout 600 python tools/bench_teacher.py clip 128 > gpurun_out/r2m_bench_clip.log 2>&1; tail -1 gpurun_out/r2m_bench_clip.log
