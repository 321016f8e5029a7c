#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
SPRAL_B200_NOPROFILE=1 timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on \
   -k regex:"k_bwd_wide_T" --launch-skip 3 --launch-count 1 -o gpurun_out/r02_solveTb64 -f python tools/profile_factor.py 100 indef solve 64 > gpurun_out/prof_solveTb.log 2>&1
SPRAL_B200_NOPROFILE=1 timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on \
   -k regex:"k_fwd_wide_T" --launch-skip 110 --launch-count 1 -o gpurun_out/r02_solveTf64 -f python tools/profile_factor.py 100 indef solve 64 > gpurun_out/prof_solveTf.log 2>&1
ls -la gpurun_out/r02_solveT?64.ncu-rep
