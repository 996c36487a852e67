#!/bin/bash
# round 2, run A: full GPU suite on the refactored step (MN-major dgrad, pack table, segmented packs), smoke, bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2a_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2a_smoke.log
python bench.py --steps 10 --warmup 3 --gemm-csv gpurun_out/r2a_gemm.csv > gpurun_out/r2a_bench.log 2>&1; echo "bench rc=$?"; tail -c 2500 gpurun_out/r2a_bench.log
python tools/bench_attn.py 256 12 > gpurun_out/r2a_attn.log 2>&1; cat gpurun_out/r2a_attn.log
