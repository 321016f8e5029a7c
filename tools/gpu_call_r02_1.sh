#!/bin/bash
# Round 2, first GPU call: opt-in variants vs default engine, the A/B table, the per-panel trace.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== opt-in variants vs default engine" | tee gpurun_out/next_call.log
SPRAL_B200_EXPERIMENTAL_TESTS=1 timeout 1500 python -m pytest tests/test_gpu_experimental.py -q -m gpu >> gpurun_out/next_call.log 2>&1
tail -8 gpurun_out/next_call.log
echo "== A/B table (27-pt 100^3)" | tee -a gpurun_out/next_call.log
timeout 2400 python tools/ab_variants.py 100 2 > gpurun_out/ab_variants.log 2>&1
tail -16 gpurun_out/ab_variants.log
echo "== per-panel trace of the default engine"
SPRAL_B200_TRACE=1 SPRAL_B200_TRACE_PANELS=1 timeout 600 python tools/profile_factor.py 100 > gpurun_out/panels_trace.out 2> gpurun_out/panels_trace.log
grep -c "\[panel\]" gpurun_out/panels_trace.log; grep "\[level" gpurun_out/panels_trace.log | tail -4
