#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for nr in 1 64; do
SPRAL_B200_NOPROFILE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv \
   --log-file gpurun_out/launches_solve22_$nr.csv python tools/profile_factor.py 100 indef solve $nr > gpurun_out/prof_solve22_$nr.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_solve22_$nr.csv "solve nrhs=$nr" 2>/dev/null | head -16
done
gzip -f gpurun_out/launches_solve22_1.csv gpurun_out/launches_solve22_64.csv
