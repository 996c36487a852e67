#!/bin/bash
# round 2, run G: full GPU suite + bench (all configs) + launch list on the current build
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2g_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2g_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 400 python bench.py --steps 20 --warmup 5 --gemm-csv gpurun_out/r2g_gemm.csv > gpurun_out/r2g_bench.log 2>&1; echo "bench rc=$?"; tail -c 1500 gpurun_out/r2g_bench.log | cut -c1-1500
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2g_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-eager --no-parity > gpurun_out/r2g_ncu_launch.log 2>&1; echo "ncu launch list rc=$?"
for cfg in "--teachers cddsv" "--backbone tiny" "--backbone tiny --teachers cddsv"; do
  timeout 300 python bench.py $cfg --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-eager 2>&1 | tail -1 >> gpurun_out/r2g_bench_other.log
done
cut -c1-420 gpurun_out/r2g_bench_other.log
