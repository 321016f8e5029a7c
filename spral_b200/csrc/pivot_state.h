/* The pivoting state machine of a front (engine.h: Front), advanced on the device by ONE
 * thread per front.  Written once and compiled twice: by nvcc as device code for the kernels of
 * factor_kernels.cu / gemm_dmma.cu, and by g++ for tests/c/pivot_state_emu.cpp, which drives
 * it together with a copy of the host mirror of subtree.cu (factor_fronts) through random
 * sequences of accepted / failed block columns, accepted / given-up / rolled-back speculative
 * segments, passes and delays. */
#pragma once
#include "engine.h"
#include "panel_v2.h"

#ifdef __CUDACC__
#define PS_FN __host__ __device__ __forceinline__
#else
#define PS_FN inline
#endif

namespace b200 {

PS_FN bool ps_isinf(double x) { return x > 1.7976931348623157e308 || x < -1.7976931348623157e308; }

/* Number of columns of the current block column that are accepted: the first
 * failing column, moved back by one if that would split a 2x2 pivot (the
 * second column of a 2x2 carries +Inf in dinv[2*j], block_ldlt.hxx:403-406). */
PS_FN int calc_ne(const Front* f) {
   int ne = f->first_fail;
   if (ne > 0 && ne < f->bs && ps_isinf(f->ws->dinv[2 * ne])) ne--;
   return ne;
}

/* Accounts a speculative segment (panel_v2.h): accepted -> CW more columns are done;
 * given up or rolled back -> the rest of the panel is done step by step.  spec_fails is a
 * penalty account: +SPEC_PENALTY per segment that was given up, -1 per accepted one; a front whose
 * account reaches SPEC_MAX_FAILS (more than one segment in SPEC_PENALTY + 1 fails, sustained) stops
 * speculating for good, sporadic failures (the benchmark's top fronts: 5 in 128 segments) cost one
 * panel each. */
PS_FN void account_segment(Front* f) {
   if (!f->seg_valid) return;
   if (f->seg_ok && !f->seg_fail) { f->done += CW; if (f->spec_fails > 0) f->spec_fails--; }
   else { f->spec_off = 1; f->spec_fails += SPEC_PENALTY; }
   f->seg_valid = 0;
}

/* Accounts the last inner step and opens / closes panels and passes.
 * Executed by ONE thread per front (first thing in k_diag, k_panel_chain and k_finalize). */
PS_FN void advance_state(Front* f, bool new_panel) {
   if (f->finished) return;
   account_segment(f);
   if (f->step_valid) {
      int ne = calc_ne(f);
      f->done += ne;
      f->pend -= f->bs - ne;
      f->step_valid = 0;
   }
   if (new_panel && f->panel_open) {
      f->end -= f->pend0 - f->pend;   // failed columns were swapped to the end
      f->panel_open = 0;
   }
   if (!f->panel_open) {
      if (f->done == f->end) {
         if (f->first_pass_done < 0) f->first_pass_done = f->done;
         if (f->end == f->n) f->finished = 1;
         else if (f->done > f->pass_start) { f->pass_start = f->done; f->end = f->n; }
         else f->finished = 1;
      }
      if (!f->finished) {
         f->p0 = f->done;
         f->pend0 = (f->done + PW < f->end) ? f->done + PW : f->end;
         f->pend = f->pend0;
         f->panel_open = 1;
         f->spec_off = 0;
      }
   }
   if (f->finished) f->nelim = f->done;
}

/* Panel snapshot for the host (k_snapshot): accounts what is pending and reports
 * p0, done, pend, pend0, end, finished, flag, spec_fails. */
PS_FN void snapshot_state(Front* f, int* o) {
   if (!f->finished) account_segment(f);
   if (!f->finished && f->step_valid) {
      int ne = calc_ne(f);
      f->done += ne;
      f->pend -= f->bs - ne;
      f->step_valid = 0;
   }
   o[0] = f->p0; o[1] = f->done; o[2] = f->pend; o[3] = f->pend0; o[4] = f->end;
   o[5] = f->finished; o[6] = f->flag; o[7] = f->spec_fails;
}

/* Whether k_panel_chain may speculate on the segment of CW columns at `done` (after advance_state). */
PS_FN bool segment_may_start(const Front* f) {
   return !f->finished && f->panel_open && !f->spec_off && f->sws != nullptr && !f->step_valid
          && f->spec_fails < SPEC_MAX_FAILS && (f->pend - f->done >= CW);
}

} // namespace b200
