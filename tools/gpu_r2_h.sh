#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2h_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r2h_tests.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager > gpurun_out/r2h_bench.log 2>&1; echo "bench rc=$?"; python - <<'PY'
import json
l=[x for x in open('gpurun_out/r2h_bench.log') if x.startswith('{')][-1]
d=json.loads(l); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'], d['gpu_launches'], d['roofline']['achieved'], d['roofline']['kernel_ms_per_step'])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2h_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-eager --no-parity > gpurun_out/r2h_ncu_launch.log 2>&1; echo "ncu launch list rc=$?"
python tools/summarize_launches.py gpurun_out/r2h_launches.csv | head -30
