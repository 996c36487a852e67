#!/bin/bash
# Other BASELINE configurations and the ncu launch list of one base step, final build (one B200 box)
mkdir -p gpurun_out
rm -f gpurun_out/ev_other.jsonl
for cfg in "--teachers cddsv" "--backbone tiny" "--backbone tiny --teachers cddsv"; do
  timeout 900 python bench.py $cfg --no-eager --no-cpu-baseline 2>/dev/null | tail -1 >> gpurun_out/ev_other.jsonl
done
python -c "
import json
for l in open('gpurun_out/ev_other.jsonl'):
    d=json.loads(l); print(d['config']['backbone'], d['config']['teachers'], round(d['value'],1), round(d['ms_per_step'],2), d['gpu_launches'], d['parity_check']['ok'])"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ev_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-eager --no-parity > gpurun_out/ev_ncu.log 2>&1; echo "ncu rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:gemm_tc_kernel -s 40 -c 8 -f -o gpurun_out/ev_gemm_step python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-eager --no-parity > gpurun_out/ev_ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
