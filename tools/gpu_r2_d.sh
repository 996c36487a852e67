#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_tc_bwd2 -s 1 -c 1 -f -o gpurun_out/r2d_attnbwd python tools/bench_attn.py 256 12 > gpurun_out/r2d_ncu.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/r2d_ncu.log
