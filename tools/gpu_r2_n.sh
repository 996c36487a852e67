#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_teachers_gpu.py -q -s > gpurun_out/r2n_teach.log 2>&1; echo "teacher tests rc=$?"; grep "rel L2\|replay\|passed\|failed" gpurun_out/r2n_teach.log | head -40
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2n_teacher_launches.csv python tools/bench_teacher.py dinov2 128 > gpurun_out/r2n_ncu.log 2>&1; echo "ncu rc=$?"
