#!/bin/bash
# round 2, run F: bwd2 with double-buffered score sets + lean math; GEMM producer under elect; conv pair A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q > gpurun_out/r2f_ktests.log 2>&1; echo "kernel tests rc=$?"; tail -4 gpurun_out/r2f_ktests.log
timeout 120 python tools/bench_attn.py 256 12 2>&1 | tee gpurun_out/r2f_attn.log
timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -x -q > gpurun_out/r2f_model_tests.log 2>&1; echo "model tests rc=$?"; tail -3 gpurun_out/r2f_model_tests.log
for env in "X=1" "THEIA_GEMM_PAIR_CONV=1"; do
env $env timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager --no-e2e --no-parity --gemm-csv gpurun_out/r2f_gemm_$env.csv > gpurun_out/r2f_bench.log 2>&1; echo "bench $env rc=$?"; python - <<'PY'
import json
l=[x for x in open('gpurun_out/r2f_bench.log') if x.startswith('{')][-1]
d=json.loads(l); print(d['value'], d['ms_per_step'], d['clocks'], d['roofline']['achieved'], d['roofline']['kernel_ms_per_step'])
PY
done
