#!/bin/bash
# look-ahead decided front by front (a failed pivot in one front of a level no longer serialises the others)
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for i in 1 2 3; do
echo "factor: $(SPRAL_B200_NOPROFILE=1 timeout 600 python tools/profile_factor.py 100 2>&1 | grep '^factor' | cut -c1-60)"
done
