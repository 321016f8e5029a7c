#!/bin/bash
# epilogue with pipelined loads; full / half / mixed tile kernels
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
run() { echo "$*: $(env "$@" SPRAL_B200_NOPROFILE=1 timeout 600 python tools/profile_factor.py 100 2>&1 | grep '^factor' | cut -c1-60 | tr '\n' ' ')"; }
run SPRAL_B200_UPD_HALF=0
run SPRAL_B200_UPD_HALF=1
run SPRAL_B200_UPD_HALF=3
run SPRAL_B200_UPD_HALF=0
run SPRAL_B200_UPD_HALF=1
run SPRAL_B200_UPD_HALF=3
SPRAL_B200_UPD_HALF=3 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "dense_fronts or doctored or sparse or cfg3" 2>&1 | tail -3
SPRAL_B200_UPD_HALF=3 SPRAL_B200_NOPROFILE=1 SPRAL_B200_TRACE=1 SPRAL_B200_TRACE_PANELS=2 timeout 600 python tools/profile_factor.py 100 > gpurun_out/timeline28.out 2> gpurun_out/timeline28.log
