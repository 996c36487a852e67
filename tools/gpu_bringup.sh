#!/bin/bash
# First-contact GPU run: each group under its own timeout so a hang in one does not hide the rest.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
run() { name=$1; shift; echo "=== $name" | tee -a gpurun_out/summary.txt; timeout -k 10 "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/summary.txt; tail -n 3 gpurun_out/$name.log | tee -a gpurun_out/summary.txt; }
: > gpurun_out/summary.txt
PT="python -m pytest -m gpu -q -p no:cacheprovider -x"
PTA="python -m pytest -m gpu -q -p no:cacheprovider"
TMO=300 run k_misc $PTA tests/test_kernels_gpu.py -k "layernorm or ln3d or loss or preprocess or pack"
TMO=300 run k_attn $PTA tests/test_kernels_gpu.py -k "attention"
TMO=200 run k_gemm_k2d $PTA tests/test_kernels_gpu.py -k "k2d"
TMO=200 run k_gemm_epi $PTA tests/test_kernels_gpu.py -k "epilogues or poscls"
TMO=200 run k_gemm_mn $PTA tests/test_kernels_gpu.py -k "wgrad_mn or mixed"
TMO=200 run k_conv $PTA tests/test_kernels_gpu.py -k "conv"
TMO=600 run model $PTA tests/test_model_gpu.py
TMO=300 run smoke python __graft_entry__.py --smoke
TMO=300 run bench_tiny python bench.py --backbone tiny --steps 5 --warmup 3 --no-cpu-baseline
TMO=600 run bench_base python bench.py --steps 5 --warmup 3 --no-cpu-baseline
cat gpurun_out/summary.txt
