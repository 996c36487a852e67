#!/bin/bash
# Round-end evidence (one B200): the bench line, the ncu launch list of a short run (tools/summarize_launches.py
# cuts the last full step out of it), and ncu --set full of six consecutive GEMM launches inside the step.
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_base_full.log 2>&1
tail -c 300 gpurun_out/bench_base_full.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_all.csv python bench.py --steps 1 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 30 -c 6 -f -o gpurun_out/prof_gemm_step python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep
# the other BASELINE configurations on the same build (device-resident value only)
for cfg in "--teachers cddsv" "--backbone tiny" "--backbone tiny --teachers cddsv"; do
  python bench.py $cfg --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 >> gpurun_out/bench_other.log
done
cut -c1-330 gpurun_out/bench_other.log
