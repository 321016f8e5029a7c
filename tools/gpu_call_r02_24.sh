#!/bin/bash
# GPU timeline of the large levels (both streams), and the solve lanes at 128 right-hand sides
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
SPRAL_B200_NOPROFILE=1 SPRAL_B200_TRACE=1 SPRAL_B200_TRACE_PANELS=2 timeout 600 python tools/profile_factor.py 100 > gpurun_out/timeline24.out 2> gpurun_out/timeline24.log
tail -5 gpurun_out/timeline24.out
for lanes in 1 2; do for nr in 128; do
echo "lanes=$lanes nrhs=$nr: $(SPRAL_B200_SOLVE_LANES=$lanes SPRAL_B200_NOPROFILE=1 timeout 600 python tools/profile_factor.py 100 indef solve $nr 2>&1 | grep 'solve nrhs' | tail -2 | tr '\n' ' ')"
done; done
