#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for nr in 1 64; do
SPRAL_B200_TRACE_SOLVE=1 SPRAL_B200_NOPROFILE=1 timeout 600 python tools/profile_factor.py 100 indef solve $nr > gpurun_out/solve_tl_$nr.out 2> gpurun_out/solve_tl_$nr.log
grep 'solve nrhs' gpurun_out/solve_tl_$nr.out | tail -1
done
