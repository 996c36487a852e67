#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_teachers_gpu.py -x -q -k "deterministic or unsupported" 2>&1 | tail -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/bench_teacher.py dinov2 128 2>gpurun_out/r2v_err.log | tee gpurun_out/r2v_teacher_n2.json | tail -1
timeout 300 python tools/bench_teacher.py dinov2 128 | tee gpurun_out/r2v_teacher_n1.json | tail -1
