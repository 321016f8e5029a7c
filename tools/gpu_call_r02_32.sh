#!/bin/bash
# look-ahead sweeps replayed as CUDA graphs (the host enqueues ~8 calls per 256-column step otherwise)
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for g in 0 1; do for la in 0 1; do for nr in 1 16 64; do
echo "graphs=$g lookahead=$la nrhs=$nr: $(SPRAL_B200_SOLVE_GRAPHS=$g SPRAL_B200_SOLVE_LOOKAHEAD=$la SPRAL_B200_NOPROFILE=1 timeout 600 python tools/profile_factor.py 100 indef solve $nr 2>&1 | grep -E 'solve nrhs|rror' | tail -2 | tr '\n' ' ')"
done; done; done
