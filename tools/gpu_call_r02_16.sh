#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
SPRAL_B200_NOPROFILE=1 timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on \
   -k regex:"k_bwd_wide_T|k_fwd_wide_T" --launch-skip 230 --launch-count 2 -o gpurun_out/r02_solveT64 -f python tools/profile_factor.py 100 indef solve 64 > gpurun_out/prof_solveT.log 2>&1
SPRAL_B200_NOPROFILE=1 timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on \
   -k regex:"k_fwd_wide_G" --launch-skip 100 --launch-count 1 -o gpurun_out/r02_solveG1 -f python tools/profile_factor.py 100 indef solve 1 > gpurun_out/prof_solveG.log 2>&1
ls -la gpurun_out/r02_solve*.ncu-rep
