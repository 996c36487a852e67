#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" | tee -a gpurun_out/summary.txt; timeout -k 10 "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/summary.txt; tail -n 4 gpurun_out/$name.log | cut -c1-600 | tee -a gpurun_out/summary.txt; }
: > gpurun_out/summary.txt
PTA="python -m pytest -m gpu -q -p no:cacheprovider"
TMO=600 run tests_all $PTA tests
TMO=300 run smoke python __graft_entry__.py --smoke
TMO=600 run bench_base python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --gemm-csv gpurun_out/gemm_base.csv
TMO=600 run bench_tiny python bench.py --backbone tiny --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --gemm-csv gpurun_out/gemm_tiny.csv
TMO=900 run ncu_base ncu --metrics gpu__time_duration.sum --clock-control none -s 1800 -c 700 --csv --log-file gpurun_out/launches_base.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e
cat gpurun_out/summary.txt
