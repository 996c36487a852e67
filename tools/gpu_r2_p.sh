#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2p_tiny_launches.csv python bench.py --backbone tiny --steps 1 --warmup 1 --no-eager --no-parity --no-cpu-baseline --no-e2e > gpurun_out/r2p_ncu.log 2>&1; echo "ncu rc=$?"
timeout 600 python bench.py --backbone tiny --no-eager --no-cpu-baseline --gemm-csv gpurun_out/r2p_gemm_tiny.csv > gpurun_out/r2p_bench_tiny.log 2>&1; tail -1 gpurun_out/r2p_bench_tiny.log | cut -c1-400
