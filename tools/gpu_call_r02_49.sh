#!/bin/bash
# two cheap sweeps over existing switches: levels that use the speculative panels, levels that use the wide sweeps with 1 RHS
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for v in 32 64 128 256; do
echo "PANEL_V2_FRONTS=$v: $(SPRAL_B200_PANEL_V2_FRONTS=$v SPRAL_B200_NOPROFILE=1 timeout 600 python tools/profile_factor.py 100 2>&1 | grep '^factor' | cut -c1-60)"
done
for v in 8 4 2 1; do
echo "SOLVE_WIDE_MIN=$v: $(SPRAL_B200_SOLVE_WIDE_MIN=$v SPRAL_B200_NOPROFILE=1 timeout 600 python tools/profile_factor.py 100 indef solve 1 2>&1 | grep 'solve nrhs' | tail -1)"
done
