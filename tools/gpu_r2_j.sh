#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_tc -s 6 -c 2 -f -o gpurun_out/r2j_attn python tools/bench_attn.py 256 12 > gpurun_out/r2j_ncu.log 2>&1; echo "ncu rc=$?"
# in-step capture of GEMM launches (traffic per launch for the roofline line) -- 8 consecutive launches inside the step
timeout 900 ncu --set full --clock-control none -k regex:gemm_tc_kernel -s 40 -c 8 -f -o gpurun_out/r2j_gemm_step python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-eager --no-parity > gpurun_out/r2j_ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
ls -la gpurun_out/*.ncu-rep
