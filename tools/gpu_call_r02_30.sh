#!/bin/bash
# tiles per CTA of the bulk launches
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
run() { echo "$*: $(env "$@" SPRAL_B200_NOPROFILE=1 timeout 600 python tools/profile_factor.py 100 2>&1 | grep '^factor' | cut -c1-60 | tr '\n' ' ')"; }
for t in 1 2 3 4 6 1 2 3 4 6; do run SPRAL_B200_BULK_TPC=$t; done
