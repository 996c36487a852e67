#!/bin/bash
# Full verification on one B200 box (run through gpurun): GPU test suite, smoke(), the bench line (both arms).
# Everything is written under gpurun_out/verify_*; nothing here reads /root/reference.
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/verify_tests.log 2>&1; echo "pytest -m gpu rc=$?"; tail -3 gpurun_out/verify_tests.log
timeout 900 python __graft_entry__.py --smoke > gpurun_out/verify_smoke.log 2>&1; echo "smoke rc=$?"; grep "smoke ok" gpurun_out/verify_smoke.log
timeout 1500 python bench.py --gemm-csv gpurun_out/verify_gemm.csv > gpurun_out/verify_bench.json 2> gpurun_out/verify_bench.err; echo "bench rc=$?"; tail -1 gpurun_out/verify_bench.json | cut -c1-600
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/verify_bench_ref.json 2> gpurun_out/verify_bench_ref.err; echo "bench reference rc=$?"; tail -1 gpurun_out/verify_bench_ref.json | cut -c1-400
