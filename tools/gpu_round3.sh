#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" | tee -a gpurun_out/summary.txt; timeout -k 10 "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/summary.txt; tail -n 4 gpurun_out/$name.log | cut -c1-900 | tee -a gpurun_out/summary.txt; }
: > gpurun_out/summary.txt
PTA="python -m pytest -m gpu -q -p no:cacheprovider --durations=12"
TMO=400 run tests_kernels $PTA tests/test_kernels_gpu.py
TMO=600 run tests_model $PTA tests/test_model_gpu.py
TMO=300 run smoke python __graft_entry__.py --smoke
TMO=600 run bench_base python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --gemm-csv gpurun_out/gemm_base.csv
TMO=600 run bench_tiny python bench.py --backbone tiny --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --gemm-csv gpurun_out/gemm_tiny.csv
