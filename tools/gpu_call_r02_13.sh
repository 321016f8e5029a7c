#!/bin/bash
# Round 2, call 13 (1 GPU): compute-sanitizer passes; ncu --set full of the largest Schur-complement launch for the
# default tile order and two blocked orders (DRAM traffic), and of the panel kernels
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
SAN="timeout 900 compute-sanitizer --error-exitcode 9 --print-limit 20"
echo "== memcheck" | tee gpurun_out/sanitizer.log
$SAN --tool memcheck python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "dense_fronts and (300 or 600) or smalllead-300 or st27_32 or lap3d_14" >> gpurun_out/sanitizer.log 2>&1; echo "memcheck rc=$?" | tee -a gpurun_out/sanitizer.log
echo "== racecheck" | tee -a gpurun_out/sanitizer.log
$SAN --tool racecheck python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "dense_fronts and 300 or smalllead-300 or st27_32" >> gpurun_out/sanitizer.log 2>&1; echo "racecheck rc=$?" | tee -a gpurun_out/sanitizer.log
echo "== synccheck" | tee -a gpurun_out/sanitizer.log
$SAN --tool synccheck python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "dense_fronts and 600 or st27_32" >> gpurun_out/sanitizer.log 2>&1; echo "synccheck rc=$?" | tee -a gpurun_out/sanitizer.log
grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY|rc=" gpurun_out/sanitizer.log | tail -12
for cb in 0 8 12; do
SPRAL_B200_CTILE_BLOCK=$cb timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on --nvtx \
   --nvtx-include "upd_contrib/" -k regex:k_update_ws --launch-skip 12 --launch-count 1 \
   -o gpurun_out/r02_contrib_cb$cb -f python tools/profile_factor.py 100 > gpurun_out/prof_full_cb$cb.log 2>&1
ncu -i gpurun_out/r02_contrib_cb$cb.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
h=rows[0]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active','lts__t_sector_hit_rate.pct','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','sm__throughput.avg.pct_of_peak_sustained_elapsed','smsp__inst_executed_pipe_fp64.sum']
for r in rows[2:]:
    print('CTILE_BLOCK=$cb', {w:r[h.index(w)] for w in want if w in h}, 'units', {w:rows[1][h.index(w)] for w in want if w in h})
"
done
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on \
   -k regex:"k_panel_chain|k_panel_tiles|k_seg_commit" --launch-skip 300 --launch-count 6 \
   -o gpurun_out/r02_panel -f python tools/profile_factor.py 100 > gpurun_out/prof_full_panel.log 2>&1
ls -la gpurun_out/*.ncu-rep
