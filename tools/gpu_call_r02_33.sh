#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for la in 0 1; do for nr in 1 64; do
echo "lookahead=$la nrhs=$nr"; SPRAL_B200_TRACE=1 SPRAL_B200_SOLVE_LOOKAHEAD=$la SPRAL_B200_NOPROFILE=1 timeout 600 python tools/profile_factor.py 100 indef solve $nr 2>&1 | grep -E '^\[solve\]|solve nrhs' | tail -4
done; done
