#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_model_gpu.py tests/test_kernels_gpu.py -x -q > gpurun_out/ab_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/ab_tests.log
timeout 600 python __graft_entry__.py --smoke 2>&1 | grep "smoke ok"
timeout 900 python bench.py --no-eager --no-cpu-baseline --gemm-csv gpurun_out/ab_gemm.csv 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('base', d['value'], d['ms_per_step'], d['roofline']['achieved'], d['clocks'], d['parity_check']['loss_rel_max'], d['parity_check']['pred_rel_l2_max'])"
THEIA_B200_LIB=/root/repo/gpurun_in/lib_old.so timeout 900 python bench.py --no-eager --no-cpu-baseline --no-parity 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('old ', d['value'], d['ms_per_step'], d['roofline']['achieved'], d['clocks'])"
