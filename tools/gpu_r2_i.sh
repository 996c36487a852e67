#!/bin/bash
# GELU epilogue variants: default (5-term erf, select-based cdf) vs alt (3-term erf 7.1.25), A/B/A on one box
for rep in 1 2; do
 for lib in "" theia_b200/libtheia_b200_alt.so; do
  for epi in 2 1088 0; do
   THEIA_B200_LIB=$lib timeout 60 python tools/bench_gemm.py 50432 3072 768 $epi 20 2>&1 | sed "s|^|[${lib:-default}] |"
  done
 done
done
