#!/bin/bash
# round 2, run E: elected-lane MMA issue with uniform descriptors (GEMM + attention), explicit shared ops
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q > gpurun_out/r2e_ktests.log 2>&1; echo "kernel tests rc=$?"; tail -4 gpurun_out/r2e_ktests.log
timeout 120 python tools/bench_attn.py 256 12 2>&1 | tee gpurun_out/r2e_attn.log
timeout 120 python tools/bench_attn.py 256 3 2>&1 | sed 's/^/H3 /'
timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -x -q > gpurun_out/r2e_model_tests.log 2>&1; echo "model tests rc=$?"; tail -3 gpurun_out/r2e_model_tests.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager --no-e2e --gemm-csv gpurun_out/r2e_gemm.csv > gpurun_out/r2e_bench.log 2>&1; echo "bench rc=$?"; python - <<'PY'
import json
l=[x for x in open('gpurun_out/r2e_bench.log') if x.startswith('{')][-1]
d=json.loads(l); print(d['value'], d['ms_per_step'], d['clocks'], d['roofline']['achieved'], d['roofline']['kernel_ms_per_step'])
PY
timeout 300 python bench.py --backbone tiny --steps 10 --warmup 3 --no-cpu-baseline --no-eager --no-e2e 2>&1 | tail -1 | cut -c1-400
