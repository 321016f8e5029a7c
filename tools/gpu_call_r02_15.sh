#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "parity or solve or dense" > gpurun_out/call15.log 2>&1; tail -2 gpurun_out/call15.log
for nr in 1 4 64; do SPRAL_B200_NOPROFILE=1 python tools/profile_factor.py 100 indef solve $nr | tail -1; done
SPRAL_B200_NOPROFILE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv \
   --log-file gpurun_out/launches_solve15_64.csv python tools/profile_factor.py 100 indef solve 64 > gpurun_out/prof_solve15_64.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_solve15_64.csv "solve nrhs=64" 2>/dev/null | head -14
SPRAL_B200_NOPROFILE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv \
   --log-file gpurun_out/launches_solve15_1.csv python tools/profile_factor.py 100 indef solve 1 > gpurun_out/prof_solve15_1.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_solve15_1.csv "solve nrhs=1" 2>/dev/null | head -14
gzip -f gpurun_out/launches_solve15_64.csv gpurun_out/launches_solve15_1.csv
