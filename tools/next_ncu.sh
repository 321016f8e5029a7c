#!/bin/bash
# Profiles of the next session, AFTER tools/next_gpu_call.sh picked the variants to keep:
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'VARIANT_ENV="SPRAL_B200_PANEL_V2=1 SPRAL_B200_BULK_PRIO=1" bash tools/next_ncu.sh'
# 1. launch list of one factorisation (serialised, cold cache: compare SHARES) -> gpurun_out/launches_r2.csv
# 2. ncu --set full of the largest Schur-complement launch and of the speculative panel kernels
# tools/summarize_launches.py turns the list into the table kept under profiles/.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export $VARIANT_ENV
echo "variant: $VARIANT_ENV"
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file gpurun_out/launches_r2.csv python tools/profile_factor.py 100 > gpurun_out/prof_run_r2.log 2>&1
tail -3 gpurun_out/prof_run_r2.log
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on --nvtx \
   --nvtx-include "upd_contrib/" -k regex:k_update_ws --launch-skip 12 --launch-count 1 \
   -o gpurun_out/prof_contrib_r2 python tools/profile_factor.py 100 > gpurun_out/prof_full_r2.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on \
   -k regex:"k_panel_chain|k_panel_tiles|k_seg_commit" --launch-skip 300 --launch-count 6 \
   -o gpurun_out/prof_panel_r2 python tools/profile_factor.py 100 >> gpurun_out/prof_full_r2.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_r2.csv 2>/dev/null
