#!/bin/bash
# Round 2, second GPU call: the one-warp diagonal-block kernel (default) and the rewritten speculative panel
# kernels (SPRAL_B200_PANEL_V2=1): parity first, then the A/B table and a per-panel trace.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== default engine (k_diag_w): GPU suite" | tee gpurun_out/call2.log
timeout 900 python -m pytest tests -q -m gpu -x >> gpurun_out/call2.log 2>&1
tail -4 gpurun_out/call2.log
echo "== PANEL_V2=1: parity tests" | tee -a gpurun_out/call2.log
SPRAL_B200_PANEL_V2=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_widened.py -q -m gpu >> gpurun_out/call2.log 2>&1
tail -6 gpurun_out/call2.log
echo "== A/B table (27-pt 100^3)" | tee -a gpurun_out/call2.log
timeout 1500 python tools/ab_variants.py 100 2 > gpurun_out/ab_variants2.log 2>&1
tail -12 gpurun_out/ab_variants2.log
echo "== per-panel trace, PANEL_V2 + BULK_PRIO"
SPRAL_B200_PANEL_V2=1 SPRAL_B200_BULK_PRIO=1 SPRAL_B200_TRACE=1 SPRAL_B200_TRACE_PANELS=1 timeout 600 python tools/profile_factor.py 100 > gpurun_out/panels_trace2.out 2> gpurun_out/panels_trace2.log
grep "\[level" gpurun_out/panels_trace2.log | tail -5; tail -3 gpurun_out/panels_trace2.out
