#!/bin/bash
# Round 2, call 4: after the latency fixes (loads before stores, branch-free maxima, reciprocals off the critical path)
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== parity, default and PANEL_V2" | tee gpurun_out/call4.log
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x >> gpurun_out/call4.log 2>&1; tail -2 gpurun_out/call4.log
SPRAL_B200_PANEL_V2=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_widened.py -q -m gpu >> gpurun_out/call4.log 2>&1; tail -2 gpurun_out/call4.log
echo "== A/B"
timeout 1500 python tools/ab_variants.py 100 2 base panel_v2 panel_v2+bulk_prio > gpurun_out/ab_variants4.log 2>&1
tail -6 gpurun_out/ab_variants4.log
echo "== ncu launch list, warm caches, PANEL_V2"
SPRAL_B200_NOPROFILE=1 SPRAL_B200_PANEL_V2=1 timeout 1200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv \
   --log-file gpurun_out/launches_r2_v2b.csv python tools/profile_factor.py 100 > gpurun_out/prof_run_r2_v2b.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_r2_v2b.csv "PANEL_V2, warm caches" | head -24
echo "== ncu launch list, warm caches, default"
SPRAL_B200_NOPROFILE=1 timeout 1200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv \
   --log-file gpurun_out/launches_r2_baseb.csv python tools/profile_factor.py 100 > gpurun_out/prof_run_r2_baseb.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_r2_baseb.csv "default, warm caches" | head -20
gzip -f gpurun_out/launches_r2_v2b.csv gpurun_out/launches_r2_baseb.csv
