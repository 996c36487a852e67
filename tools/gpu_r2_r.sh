#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_teachers_gpu.py -q -s > gpurun_out/r2r_teach.log 2>&1; echo "teacher tests rc=$?"; grep "passed\|failed" gpurun_out/r2r_teach.log
timeout 600 python tools/bench_teacher.py dinov2 128 --hf > gpurun_out/r2r_bench_dinov2.json 2>gpurun_out/r2r_err.log; tail -1 gpurun_out/r2r_bench_dinov2.json
timeout 600 python tools/bench_teacher.py clip 128 --hf > gpurun_out/r2r_bench_clip.json 2>>gpurun_out/r2r_err.log; tail -1 gpurun_out/r2r_bench_clip.json
timeout 600 python tools/bench_teacher.py dinov2 256 > gpurun_out/r2r_bench_dinov2_b256.json 2>>gpurun_out/r2r_err.log; tail -1 gpurun_out/r2r_bench_dinov2_b256.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2r_teacher_launches.csv python tools/bench_teacher.py dinov2 128 > gpurun_out/r2r_ncu.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_tc_fwd_long -s 30 -c 1 -o gpurun_out/r2r_attn_long python tools/bench_teacher.py dinov2 128 > gpurun_out/r2r_ncu2.log 2>&1; echo "ncu full rc=$?"
