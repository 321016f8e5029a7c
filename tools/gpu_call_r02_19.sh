#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "parity or solve or dense or refine" > gpurun_out/call19.log 2>&1; tail -2 gpurun_out/call19.log
for nr in 1 16 64; do SPRAL_B200_NOPROFILE=1 python tools/profile_factor.py 100 indef solve $nr | tail -1; done
