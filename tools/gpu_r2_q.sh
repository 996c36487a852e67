#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "layernorm" 2>&1 | tail -2
timeout 120 python tools/bench_ln.py
THEIA_B200_LIB=/root/repo/gpurun_in/lib_old.so timeout 120 python tools/bench_ln.py
timeout 600 python bench.py --backbone tiny --no-eager --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('tiny', d['value'], d['ms_per_step'], d['parity_check'])"
