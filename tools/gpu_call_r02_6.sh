#!/bin/bash
# Round 2, call 6: rewritten wide-solve G kernels (bandwidth path + DMMA path), accumulator-based backward step
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_widened.py -q -m gpu -x > gpurun_out/call6.log 2>&1; tail -3 gpurun_out/call6.log
for wm in 8 1; do
echo "== SOLVE_WIDE_MIN=$wm"
SPRAL_B200_NOPROFILE=1 SPRAL_B200_SOLVE_WIDE_MIN=$wm python tools/profile_factor.py 100 indef solve 1 | tail -1
SPRAL_B200_NOPROFILE=1 SPRAL_B200_SOLVE_WIDE_MIN=$wm python tools/profile_factor.py 100 indef solve 64 | tail -1
done
for nr in 1 64; do
SPRAL_B200_NOPROFILE=1 SPRAL_B200_SOLVE_WIDE_MIN=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv \
   --log-file gpurun_out/launches_solve6_$nr.csv python tools/profile_factor.py 100 indef solve $nr > gpurun_out/prof_solve6_$nr.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_solve6_$nr.csv "solve nrhs=$nr, WIDE_MIN=1" | head -20
done
gzip -f gpurun_out/launches_solve6_1.csv gpurun_out/launches_solve6_64.csv
