#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
SPRAL_B200_NOPROFILE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv \
   --log-file gpurun_out/launches_solve45_64.csv python tools/profile_factor.py 100 indef solve 64 > gpurun_out/prof_solve45_64.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_solve45_64.csv "solve nrhs=64" 2>/dev/null | head -16
