#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for tpc in 1 0 2 4; do echo "== BULK_TPC=$tpc (0 = adaptive)"; SPRAL_B200_BULK_TPC=$tpc timeout 600 python tools/ab_variants.py 100 2 base 2>/dev/null | grep "^| base"; done
