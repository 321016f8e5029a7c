#!/bin/bash
# Round 2, call 14 (1 GPU): everything so far -- parity suite, A/B, solve timings, launch list
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/call14.log 2>&1; tail -3 gpurun_out/call14.log
timeout 900 python tools/ab_variants.py 100 2 base > gpurun_out/ab_variants14.log 2>&1; tail -3 gpurun_out/ab_variants14.log
SPRAL_B200_NOPROFILE=1 python tools/profile_factor.py 100 indef solve 1 | tail -1
SPRAL_B200_NOPROFILE=1 python tools/profile_factor.py 100 indef solve 64 | tail -1
SPRAL_B200_NOPROFILE=1 timeout 1200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv \
   --log-file gpurun_out/launches_r2_14.csv python tools/profile_factor.py 100 > gpurun_out/prof_run_r2_14.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_r2_14.csv "default, warm caches" | head -16
SPRAL_B200_NOPROFILE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv \
   --log-file gpurun_out/launches_solve14_64.csv python tools/profile_factor.py 100 indef solve 64 > gpurun_out/prof_solve14_64.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_solve14_64.csv "solve nrhs=64" | head -14
gzip -f gpurun_out/launches_r2_14.csv gpurun_out/launches_solve14_64.csv
