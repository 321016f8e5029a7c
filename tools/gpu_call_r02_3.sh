#!/bin/bash
# Round 2, third GPU call: where does a panel's time go?  (a) host enqueue vs wait per panel (no profiling events),
# (b) ncu launch list with warm caches of the speculative-panel variant and of the default.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== panel trace without profiling events (PANEL_V2 + BULK_PRIO)"
SPRAL_B200_NOPROFILE=1 SPRAL_B200_PANEL_V2=1 SPRAL_B200_BULK_PRIO=1 SPRAL_B200_TRACE_PANELS=1 timeout 600 python tools/profile_factor.py 100 > gpurun_out/trace3_v2.out 2> gpurun_out/trace3_v2.log
grep "m 16349" gpurun_out/trace3_v2.log | tail -64 | awk 'NR%8==1' | cut -c60-400
echo "== same, default engine"
SPRAL_B200_NOPROFILE=1 SPRAL_B200_TRACE_PANELS=1 timeout 600 python tools/profile_factor.py 100 > gpurun_out/trace3_base.out 2> gpurun_out/trace3_base.log
grep "m 16349" gpurun_out/trace3_base.log | tail -64 | awk 'NR%8==1' | cut -c60-400
echo "== ncu launch list, warm caches, PANEL_V2"
SPRAL_B200_NOPROFILE=1 SPRAL_B200_PANEL_V2=1 timeout 1200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv \
   --log-file gpurun_out/launches_r2_v2.csv python tools/profile_factor.py 100 > gpurun_out/prof_run_r2_v2.log 2>&1
tail -2 gpurun_out/prof_run_r2_v2.log
python tools/summarize_launches.py gpurun_out/launches_r2_v2.csv "PANEL_V2, warm caches" | head -40
echo "== ncu launch list, warm caches, default"
SPRAL_B200_NOPROFILE=1 timeout 1200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv \
   --log-file gpurun_out/launches_r2_base.csv python tools/profile_factor.py 100 > gpurun_out/prof_run_r2_base.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_r2_base.csv "default, warm caches" | head -40
gzip -f gpurun_out/launches_r2_v2.csv gpurun_out/launches_r2_base.csv
