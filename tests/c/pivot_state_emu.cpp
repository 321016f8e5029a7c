/* TEST INFRASTRUCTURE (CPU only): model check of the pivoting protocol between the device state machine
 * (spral_b200/csrc/pivot_state.h -- the SAME code the kernels run: advance_state, account_segment,
 * calc_ne, snapshot_state, segment_may_start) and the host mirror of factor_fronts (subtree.cu), which is
 * restated here launch for launch without the numerics: every "kernel" only does what the real one does
 * to the state, and the outcomes that depend on the matrix are drawn at random --
 *   the first failing column of a block column (k_apply), 2x2 pivots that must not be split,
 *   speculative segments that are accepted / given up by the chain / rolled back by the tiles.
 * Checked on thousands of random fronts, several per level: the host mirror never diverges from the
 * device (the exception factor_fronts would throw), every panel is complete after the launches the host
 * issues, the loop terminates, and at the end eliminated + delayed columns = n, with the statistics the
 * level-end code reads.
 *
 * Build: g++ -O2 -std=c++17 -I/usr/local/cuda/include -Iinclude tests/c/pivot_state_emu.cpp */
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <random>
#include <stdexcept>
#include <vector>

#include "../../spral_b200/csrc/pivot_state.h"

using namespace b200;
static const double INF = std::numeric_limits<double>::infinity();

struct HostState {            // subtree.cu
   int fi = 0, m = 0, n = 0;
   int done = 0, end = 0, pass_start = 0, p0 = 0, pend0 = 0, pend = 0;
   bool finished = false;
   bool spec_dead = false;
};

struct Sim {
   std::mt19937_64 rng;
   double p_fail, p_chain_giveup, p_tile_fail;
   bool v2;
   long launches = 0, panels = 0, segments_ok = 0, segments_lost = 0, steps = 0;
   explicit Sim(uint64_t seed) : rng(seed) {}
   double U() { return std::uniform_real_distribution<double>(0, 1)(rng); }
   int I(int lo, int hi) { return std::uniform_int_distribution<int>(lo, hi)(rng); }

   /* ---- the device side: what each kernel does to the state ---- */
   void k_diag(Front* f, bool new_panel) {
      advance_state(f, new_panel);
      if (!f->finished && f->done < f->pend) {
         f->bs = std::min(BS, f->pend - f->done);
         f->first_fail = f->bs;
         f->step_valid = 1;
         for (int j = 0; j < 2 * BS; ++j) f->ws->dinv[j] = 1.0;
         for (int j = 0; j + 1 < f->bs;) {                      // random 1x1 / 2x2 structure of the block
            if (U() < 0.3) { f->ws->dinv[2 * (j + 1)] = INF; j += 2; } else j += 1;
         }
      } else f->bs = 0;
      ++launches;
   }
   void k_apply(Front* f) {
      if (f->step_valid && U() < p_fail) f->first_fail = std::min(f->first_fail, I(0, f->bs - 1));
      ++launches;
   }
   void k_panel_chain(Front* f, bool new_panel) {
      advance_state(f, new_panel);
      if (segment_may_start(f)) {
         f->seg_valid = 1; f->seg_fail = 0;
         f->seg_ok = U() < p_chain_giveup ? 0 : 1;
      }
      ++launches;
   }
   void k_panel_tiles(Front* f) {
      if (f->seg_valid && f->seg_ok && U() < p_tile_fail) f->seg_fail = 1;
      ++launches;
   }

   /* ---- the host side: factor_fronts of subtree.cu for one level ---- */
   void factor_fronts(std::vector<Front>& F) {
      std::vector<HostState> H(F.size());
      for (size_t i = 0; i < F.size(); ++i) {
         HostState& h = H[i];
         h.fi = (int)i; h.n = F[i].n; h.m = F[i].m;
         h.done = 0; h.end = F[i].n; h.pass_start = 0;
         h.finished = (F[i].n == 0);
         h.p0 = 0; h.pend0 = std::min(PW, F[i].n); h.pend = h.pend0;
      }
      std::vector<int> snap_host;
      for (int guard = 0;; ++guard) {
         if (guard > 100000) throw std::runtime_error("the panel loop does not terminate");
         std::vector<int> act;
         for (size_t i = 0; i < H.size(); ++i) if (!H[i].finished) act.push_back((int)i);
         if (act.empty()) break;
         std::stable_sort(act.begin(), act.end(), [&](int a, int b) { return H[a].pend0 - H[a].p0 > H[b].pend0 - H[b].p0; });
         const int na_all = (int)act.size();
         std::vector<int> cand(na_all);
         for (int k = 0; k < na_all; ++k) cand[k] = H[act[k]].pend0 - H[act[k]].p0;
         const int nsteps = (cand[0] + BS - 1) / BS;
         auto count_gt = [&](int thr) { int c = 0; while (c < na_all && cand[c] > thr) ++c; return c; };
         auto take_snapshot = [&]() {
            snap_host.assign((size_t)na_all * 8, 0);
            for (int k = 0; k < na_all; ++k) snapshot_state(&F[H[act[k]].fi], &snap_host[(size_t)k * 8]);
            ++launches;
         };
         ++panels;
         bool steps_new_panel = true;
         int steps_todo = nsteps;
         bool use_v2 = v2;
         if (use_v2) {
            bool any_alive = false;
            for (int k = 0; k < na_all; ++k) any_alive = any_alive || !H[act[k]].spec_dead;
            use_v2 = any_alive;
         }
         if (use_v2) {
            const int nseg = PW / CW;
            for (int seg = 0; seg < nseg; ++seg) {
               for (int k = 0; k < na_all; ++k) k_panel_chain(&F[H[act[k]].fi], seg == 0);
               for (int k = 0; k < na_all; ++k) k_panel_tiles(&F[H[act[k]].fi]);
               for (int k = 0; k < na_all; ++k) {
                  const Front& f = F[H[act[k]].fi];
                  if (f.seg_valid) { if (f.seg_ok && !f.seg_fail) ++segments_ok; else ++segments_lost; }
               }
               launches += 2;                                // commit, UPD_SEG: no state change
            }
            take_snapshot();
            int maxrem = 0;
            for (int k = 0; k < na_all; ++k) {
               const int* sn = &snap_host[(size_t)k * 8];
               if (sn[6] < 0 || sn[5]) continue;
               maxrem = std::max(maxrem, sn[2] - sn[1]);
            }
            steps_new_panel = false;
            steps_todo = (maxrem + BS - 1) / BS;
         }
         for (int st = 0; st < steps_todo; ++st) {
            const int na = use_v2 ? na_all : count_gt(st * BS);
            if (na == 0) break;
            for (int k = 0; k < na; ++k) k_diag(&F[H[act[k]].fi], st == 0 && steps_new_panel);
            for (int k = 0; k < na; ++k) k_apply(&F[H[act[k]].fi]);
            launches += 3;                                   // commit, inner update, swap: no state change
            ++steps;
         }
         if (!use_v2 || steps_todo > 0) take_snapshot();
         /* what happened in the panel */
         for (int k = 0; k < na_all; ++k) {
            HostState& h = H[act[k]];
            const int* sn = &snap_host[(size_t)k * 8];
            if (sn[6] < 0) { h.finished = true; continue; }
            if (sn[0] != h.p0 || sn[3] != h.pend0 || sn[4] != h.end)
               throw std::runtime_error("host mirror of the pivoting state diverged from the device");
            h.done = sn[1]; h.pend = sn[2];
            h.spec_dead = sn[7] >= SPEC_MAX_FAILS;
            if (h.done != h.pend) throw std::runtime_error("a panel was left incomplete by the launches of the host");
            if (h.done < h.p0 || h.pend > h.pend0) throw std::runtime_error("state out of range");
         }
         /* mirror of advance_state(new_panel = true) */
         for (int k = 0; k < na_all; ++k) {
            HostState& h = H[act[k]];
            if (h.finished) continue;
            h.end -= h.pend0 - h.pend;
            if (h.done == h.end) {
               if (h.end == h.n) h.finished = true;
               else if (h.done > h.pass_start) { h.pass_start = h.done; h.end = h.n; }
               else h.finished = true;
            }
            if (!h.finished) { h.p0 = h.done; h.pend0 = std::min(h.done + PW, h.end); h.pend = h.pend0; }
         }
      }
      /* level end: k_finalize */
      for (size_t i = 0; i < F.size(); ++i) {
         advance_state(&F[i], true);
         if (!F[i].finished) throw std::runtime_error("front not finished after the panel loop");
         if (F[i].nelim != F[i].done || F[i].nelim > F[i].n || F[i].nelim != H[i].done)
            throw std::runtime_error("eliminated columns disagree");
         if (F[i].seg_valid || F[i].step_valid) throw std::runtime_error("something was left unaccounted");
      }
   }
};

int main(int argc, char** argv) {
   const int ntrial = argc > 1 ? atoi(argv[1]) : 4000;
   long total_cols = 0, total_elim = 0, total_launch = 0, seg_ok = 0, seg_lost = 0;
   int failures = 0;
   for (int trial = 0; trial < ntrial; ++trial) {
      Sim sim(1000 + trial);
      sim.v2 = trial % 4 != 0;
      const double pf[] = {0.0, 0.02, 0.2, 0.9};
      sim.p_fail = pf[trial % 4 == 0 ? (trial / 4) % 4 : sim.I(0, 3)];
      sim.p_chain_giveup = sim.U() < 0.5 ? 0.0 : sim.U() * 0.6;
      sim.p_tile_fail = sim.U() < 0.5 ? 0.0 : sim.U() * 0.6;
      const int nfront = sim.I(1, 6);
      std::vector<Front> F(nfront);
      std::vector<BlockWS> ws(nfront);
      static SegWS* dummy = reinterpret_cast<SegWS*>(&ws);       // only tested against nullptr
      for (int i = 0; i < nfront; ++i) {
         Front& f = F[i];
         f = Front();
         const int n = sim.U() < 0.1 ? sim.I(0, 3) : sim.I(1, 1700);
         f.n = n; f.m = n + sim.I(0, 500);
         f.end = f.n; f.first_pass_done = -1;
         f.ws = &ws[i];
         f.sws = sim.v2 ? dummy : nullptr;
      }
      try {
         sim.factor_fronts(F);
         for (const Front& f : F) { total_cols += f.n; total_elim += f.nelim; }
         if (sim.p_fail == 0.0)
            for (const Front& f : F) if (f.nelim != f.n) throw std::runtime_error("columns delayed although nothing failed");
      } catch (const std::exception& e) {
         printf("trial %d (v2=%d p_fail=%.2f giveup=%.2f tile_fail=%.2f): %s\n", trial, (int)sim.v2, sim.p_fail,
                sim.p_chain_giveup, sim.p_tile_fail, e.what());
         ++failures;
      }
      total_launch += sim.launches; seg_ok += sim.segments_ok; seg_lost += sim.segments_lost;
   }
   printf("pivot_state_emu: %d levels, %ld columns, %ld eliminated, %ld delayed, %ld segments accepted, %ld given up / rolled "
          "back, %ld launches, %d failures\n", ntrial, total_cols, total_elim, total_cols - total_elim, seg_ok, seg_lost,
          total_launch, failures);
   return failures ? 1 : 0;
}
