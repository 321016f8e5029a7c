#!/bin/bash
# Round 2, call 5: where do the solves spend their time?  launch lists (warm caches) for 1 and 64 right-hand sides
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for nr in 1 64; do
SPRAL_B200_NOPROFILE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv \
   --log-file gpurun_out/launches_solve_$nr.csv python tools/profile_factor.py 100 indef solve $nr > gpurun_out/prof_solve_$nr.log 2>&1
tail -3 gpurun_out/prof_solve_$nr.log
python tools/summarize_launches.py gpurun_out/launches_solve_$nr.csv "solve nrhs=$nr" | head -24
done
python tools/profile_factor.py 100 indef solve 1 | tail -2
python tools/profile_factor.py 100 indef solve 64 | tail -2
gzip -f gpurun_out/launches_solve_1.csv gpurun_out/launches_solve_64.csv
