#!/bin/bash
# round 2, run B: attention v2 kernels (forward: register-resident softmax, double-buffered scores; backward: no recompute)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k attention > gpurun_out/r2b_attn_tests.log 2>&1; echo "attn tests rc=$?"; tail -15 gpurun_out/r2b_attn_tests.log
timeout 120 python tools/bench_attn.py 256 12 > gpurun_out/r2b_attn.log 2>&1; echo "bench_attn rc=$?"; cat gpurun_out/r2b_attn.log
THEIA_ATTN_FWD_V1=1 THEIA_ATTN_BWD_V1=1 timeout 120 python tools/bench_attn.py 256 12 2>&1 | sed 's/^/v1 /'
timeout 120 python tools/bench_attn.py 256 3 2>&1 | sed 's/^/H3 /'
timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -x -q > gpurun_out/r2b_model_tests.log 2>&1; echo "model tests rc=$?"; tail -5 gpurun_out/r2b_model_tests.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager --no-e2e > gpurun_out/r2b_bench.log 2>&1; echo "bench rc=$?"; tail -c 900 gpurun_out/r2b_bench.log
