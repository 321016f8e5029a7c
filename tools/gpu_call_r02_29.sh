#!/bin/bash
# ncu --set full: one K = 256 panel update (persistent launch, look-ahead off so that every panel update is one launch)
# and the largest Schur-complement launch, with the setmaxnreg kernel
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
pick() { ncu -i $1 --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
want=['Grid Size','Block Size','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active','lts__t_sector_hit_rate.pct','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','sm__throughput.avg.pct_of_peak_sustained_elapsed','launch__registers_per_thread','smsp__cycles_active.avg','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','lts__t_bytes.sum','smsp__average_warp_latency_issue_stalled_long_scoreboard.pct','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio']
for r in rows[2:]:
    for w in want:
        if w in h: print(w, r[h.index(w)], rows[1][h.index(w)])
    for i,name in enumerate(h):
        if 'issue_stalled' in name and 'per_warp_active' in name: print(name, r[i])
"; }
SPRAL_B200_UPD_HALF=0 SPRAL_B200_LOOKAHEAD=0 timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on \
   -k regex:k_update_ws --launch-skip 110 --launch-count 1 -o gpurun_out/r02_panel_update -f python tools/profile_factor.py 100 > gpurun_out/prof_full_pu.log 2>&1
pick gpurun_out/r02_panel_update.ncu-rep
SPRAL_B200_UPD_HALF=0 timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on --nvtx \
   --nvtx-include "upd_contrib/" -k regex:k_update_ws --launch-skip 12 --launch-count 1 \
   -o gpurun_out/r02_contrib_v3 -f python tools/profile_factor.py 100 > gpurun_out/prof_full_c3.log 2>&1
pick gpurun_out/r02_contrib_v3.ncu-rep
