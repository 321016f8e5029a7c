#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for nr in 1 16 64 128; do
echo "nrhs=$nr: $(SPRAL_B200_NOPROFILE=1 timeout 600 python tools/profile_factor.py 100 indef solve $nr 2>&1 | grep 'solve nrhs' | tail -1)"
done
echo "nrhs=128 no lookahead: $(SPRAL_B200_SOLVE_LOOKAHEAD=0 SPRAL_B200_NOPROFILE=1 timeout 600 python tools/profile_factor.py 100 indef solve 128 2>&1 | grep 'solve nrhs' | tail -1)"
echo "nrhs=128 one lane: $(SPRAL_B200_SOLVE_LANES=1 SPRAL_B200_NOPROFILE=1 timeout 600 python tools/profile_factor.py 100 indef solve 128 2>&1 | grep 'solve nrhs' | tail -1)"
SPRAL_B200_TRACE_SOLVE=1 SPRAL_B200_NOPROFILE=1 timeout 600 python tools/profile_factor.py 100 indef solve 64 > gpurun_out/solve_tl_64.out 2> gpurun_out/solve_tl_64.log
