/* TEST INFRASTRUCTURE (CPU only): model check of the TWO-STREAM schedule of factor_fronts (subtree.cu).
 *
 * The panel loop runs the latency-bound panel kernels on the main stream while the bulk of the previous
 * panels' trailing updates (UPD_EXPLICIT) runs on a second stream; the only ordering between the two is
 *   - the host's stream synchronisation at every panel snapshot (main stream only),
 *   - cudaStreamWaitEvent(main, ev_bulk | ev_bulk_all) where factor_fronts issues them.
 * This program restates that host code launch for launch (as tests/c/pivot_state_emu.cpp does for the state
 * machine, which it shares: spral_b200/csrc/pivot_state.h), gives every launch the footprint of the real
 * kernel -- the rectangles of L, L*D and the backup it reads and writes, at the tile granularity of the TMA
 * operand loads -- and checks, over random fronts with failed block columns, passes, delays and
 * (SPRAL_B200_PANEL_V2) accepted / given-up / rolled-back speculative segments:
 *   1. no two launches that may run concurrently (one per stream, not ordered by a sync or an event wait)
 *      touch the same entries unless both only read them;
 *   2. every column has received the update of EVERY previously eliminated column when its diagonal block
 *      is factorised (counted per column, carried through the symmetric swaps), and all of them at the end.
 * 1 + 2 together say the look-ahead computes what the in-order schedule computes.
 *
 * Build: g++ -O2 -std=c++17 -I/usr/local/cuda/include -Iinclude tests/c/lookahead_race_emu.cpp */
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../spral_b200/csrc/pivot_state.h"

using namespace b200;
static const double INF = std::numeric_limits<double>::infinity();
static const int T = 128, Ti = 64;   // update_tile_size(big), inner_tile_size(big)  (gemm_dmma.cu)

enum Arr { A_L = 0, A_LD = 1, A_BK = 2 };
struct Rect { int arr, front, r0, r1, c0, c1; bool w; };       // half-open; L / LD in front coordinates
struct Launch { const char* name; std::vector<Rect> rects; };

static bool conflict(const Rect& a, const Rect& b) {
   if (a.arr != b.arr || a.front != b.front || !(a.w || b.w)) return false;
   return a.r0 < b.r1 && b.r0 < a.r1 && a.c0 < b.c1 && b.c0 < a.c1;
}

struct HostState {            // subtree.cu
   int fi = 0, m = 0, n = 0;
   int done = 0, end = 0, pass_start = 0, p0 = 0, pend0 = 0, pend = 0;
   bool finished = false;
   bool spec_dead = false;
};
struct MatTile_ { int front, ti, tj; };

struct Sim {
   std::mt19937_64 rng;
   double p_fail = 0, p_chain_giveup = 0, p_tile_fail = 0;
   bool v2 = false, lookahead_on = true;
   int sm_count = 148;
   int inject = 0;                   // fault injection (the checker must notice): 1 drop the ev_bulk wait, 2 drop the
                                     // ev_bulk_all wait, 3 the bulk starts one tile column early, 4 the second bulk part is lost
   long n_main = 0, n_bulk = 0, n_pairs = 0, n_waits = 0;
   explicit Sim(uint64_t seed) : rng(seed) {}
   double U() { return std::uniform_real_distribution<double>(0, 1)(rng); }
   int I(int lo, int hi) { return std::uniform_int_distribution<int>(lo, hi)(rng); }

   std::vector<Front>* Fp = nullptr;
   std::vector<std::vector<int>> upd;            // per front, per column: updates received (count of pivot columns)

   /* ---- stream model ---- */
   std::vector<Launch> main_unsynced;            // main-stream launches since the last host sync of that stream
   std::vector<Launch> s2_pending;               // second-stream launches not known to be complete
   long s2_issued = 0, s2_base = 0;              // s2_pending[i] has sequence number s2_base + i
   long ev_bulk = -1, ev_bulk_all = -1;          // sequence numbers covered by the last record of each event

   void check(const Launch& x, const Launch& y) {
      for (const Rect& a : x.rects)
         for (const Rect& b : y.rects) {
            ++n_pairs;
            if (conflict(a, b)) {
               char buf[512];
               snprintf(buf, sizeof buf, "race: %s {arr %d front %d rows [%d,%d) cols [%d,%d) %s} on the main stream vs %s "
                        "{rows [%d,%d) cols [%d,%d) %s} on the bulk stream", x.name, a.arr, a.front, a.r0, a.r1, a.c0, a.c1,
                        a.w ? "W" : "R", y.name, b.r0, b.r1, b.c0, b.c1, b.w ? "W" : "R");
               throw std::runtime_error(buf);
            }
         }
   }
   void issue_main(Launch&& x) {
      ++n_main;
      for (const Launch& y : s2_pending) check(x, y);
      main_unsynced.push_back(std::move(x));
   }
   void issue_s2(Launch&& y) {
      ++n_bulk;
      for (const Launch& x : main_unsynced) check(x, y);
      s2_pending.push_back(std::move(y));
      ++s2_issued;
   }
   void sync_main() { main_unsynced.clear(); }
   void record(long& ev) { ev = s2_issued; }
   void wait_main(long ev) {                     // cudaStreamWaitEvent(main, ev): launches before the record are complete
      ++n_waits;
      /* launches issued on the main stream BEFORE the wait are not ordered by it, but every one of them is
       * ordered before the later main launches, which is all that issue_main() compares against */
      while (s2_base < ev && !s2_pending.empty()) { s2_pending.erase(s2_pending.begin()); ++s2_base; }
   }

   /* ---- footprints of the kernels (factor_kernels.cu, gemm_dmma.cu, panel_v2.h) ---- */
   static void sym_rects(std::vector<Rect>& out, int fi, int s, int m) {     // k_swap: position s of the symmetric front
      out.push_back({A_L, fi, s, s + 1, 0, s + 1, true});
      out.push_back({A_L, fi, s, m, s, s + 1, true});
   }
   /* A(r, c) -= L(r, K) LD(c, K)^T on columns [c_lo, c_hi) of tile columns [tj_lo, tj_hi], tile size tile */
   static void upd_rects(std::vector<Rect>& out, const Front& f, int fi, int k0, int k1, int c_lo, int c_hi, int tj_lo,
         int tj_hi, int tile) {
      const int cl = std::max(c_lo, tj_lo * tile), ch = std::min(c_hi, (tj_hi + 1) * tile);
      if (cl >= ch || k1 <= k0) return;
      const int rlo = (cl / tile) * tile;                                   // operand tiles are tile-aligned
      out.push_back({A_L, fi, cl, f.m, cl, ch, true});
      out.push_back({A_L, fi, rlo, f.ldl, k0, k1, false});
      out.push_back({A_LD, fi, rlo, std::min(f.ldl, ((ch + tile - 1) / tile) * tile), k0, k1, false});
   }

   void k_diag(int fi, bool new_panel) {
      Front* f = &(*Fp)[fi];
      advance_state(f, new_panel);
      Launch x{"k_diag", {}};
      if (!f->finished && f->done < f->pend) {
         f->bs = std::min(BS, f->pend - f->done);
         f->first_fail = f->bs;
         f->step_valid = 1;
         for (int j = 0; j < 2 * BS; ++j) f->ws->dinv[j] = 1.0;
         for (int j = 0; j + 1 < f->bs;) { if (U() < 0.3) { f->ws->dinv[2 * (j + 1)] = INF; j += 2; } else j += 1; }
         x.rects.push_back({A_L, fi, f->done, f->done + f->bs, f->done, f->done + f->bs, false});
         for (int c = f->done; c < f->done + f->bs; ++c)
            if (upd[fi][c] != f->done) throw std::runtime_error("k_diag: a column of the block has not received every update");
      } else f->bs = 0;
      issue_main(std::move(x));
   }
   void k_apply(int fi) {
      Front* f = &(*Fp)[fi];
      Launch x{"k_apply", {}};
      if (f->step_valid) {
         if (U() < p_fail) f->first_fail = std::min(f->first_fail, I(0, f->bs - 1));
         const int d = f->done, b = f->bs;
         x.rects.push_back({A_L, fi, d + b, f->m, d, d + b, true});
         x.rects.push_back({A_LD, fi, d + b, f->m, d, d + b, true});
         x.rects.push_back({A_BK, fi, d + b, f->m, 0, b, true});
      }
      issue_main(std::move(x));
   }
   void k_commit(int fi) {
      Front* f = &(*Fp)[fi];
      Launch x{"k_commit", {}};
      if (f->step_valid) {
         const int d = f->done, b = f->bs, ne = calc_ne(f);
         x.rects.push_back({A_L, fi, d, d + b, 0, d, true});
         x.rects.push_back({A_L, fi, d, d + b, d, d + b, true});
         x.rects.push_back({A_LD, fi, d + ne, d + b, d, d + ne, true});
         if (ne < b) { x.rects.push_back({A_L, fi, d + b, f->m, d + ne, d + b, true}); x.rects.push_back({A_BK, fi, d + b, f->m, ne, b, false}); }
      }
      issue_main(std::move(x));
   }
   void upd_inner(int fi, const HostState& h) {
      Front* f = &(*Fp)[fi];
      Launch x{"UPD_INNER", {}};
      if (f->step_valid) {
         const int ne = calc_ne(f);
         if (ne > 0) {
            /* tile list of the host: tile columns p0 / Ti .. (pend0 - 1) / Ti */
            upd_rects(x.rects, *f, fi, f->done, f->done + ne, f->done + ne, f->pend0, h.p0 / Ti, (h.pend0 - 1) / Ti, Ti);
            for (int c = f->done + ne; c < f->pend0; ++c) upd[fi][c] += ne;
            if (f->pend0 > ((h.pend0 - 1) / Ti + 1) * Ti) throw std::runtime_error("inner tile list does not cover the panel");
         }
      }
      issue_main(std::move(x));
   }
   void swap_cols(int fi, int a0, int b0, int nswap, std::vector<Rect>& out) {
      Front* f = &(*Fp)[fi];
      for (int t = 0; t < nswap; ++t) {
         sym_rects(out, fi, a0 + t, f->m); sym_rects(out, fi, b0 + t, f->m);
         std::swap(upd[fi][a0 + t], upd[fi][b0 + t]);
      }
   }
   void k_swap_inner(int fi) {
      Front* f = &(*Fp)[fi];
      Launch x{"k_swap(inner)", {}};
      if (f->step_valid) {
         const int ne = calc_ne(f), nfail = f->bs - ne;
         if (nfail > 0) {
            const int a0 = f->done + ne, rem = f->pend - (f->done + f->bs), nswap = std::min(nfail, rem);
            if (nswap > 0) swap_cols(fi, a0, f->pend - nswap, nswap, x.rects);
         }
      }
      issue_main(std::move(x));
   }
   void k_swap_outer(int fi) {
      Front* f = &(*Fp)[fi];
      Launch x{"k_swap(outer)", {}};
      if (f->panel_open && !f->finished) {
         int pend = f->pend;
         if (f->step_valid) pend -= f->bs - calc_ne(f);
         const int nf = f->pend0 - pend;
         if (nf > 0) {
            const int rem = f->end - f->pend0, nswap = std::min(nf, rem);
            if (nswap > 0) swap_cols(fi, pend, f->end - nswap, nswap, x.rects);
         }
      }
      issue_main(std::move(x));
   }
   void k_panel_chain(int fi, bool new_panel) {
      Front* f = &(*Fp)[fi];
      advance_state(f, new_panel);
      Launch x{"k_panel_chain", {}};
      if (segment_may_start(f)) {
         f->seg_valid = 1; f->seg_fail = 0;
         f->seg_ok = U() < p_chain_giveup ? 0 : 1;
         x.rects.push_back({A_L, fi, f->done, f->done + CW, f->done, f->done + CW, false});
         for (int c = f->done; c < f->done + CW; ++c)
            if (upd[fi][c] != f->done) throw std::runtime_error("k_panel_chain: a column of the segment has not received every update");
      }
      issue_main(std::move(x));
   }
   void k_panel_tiles(int fi) {
      Front* f = &(*Fp)[fi];
      Launch x{"k_panel_tiles", {}};
      if (f->seg_valid && f->seg_ok) {
         if (U() < p_tile_fail) f->seg_fail = 1;
         const int p = f->done;
         x.rects.push_back({A_L, fi, p + CW, f->m, p, p + CW, true});
         x.rects.push_back({A_LD, fi, p + CW, f->m, p, p + CW, true});
         x.rects.push_back({A_BK, fi, p + CW, f->m, 0, CW, true});
      }
      issue_main(std::move(x));
   }
   void k_seg_commit(int fi) {
      Front* f = &(*Fp)[fi];
      Launch x{"k_seg_commit", {}};
      if (f->seg_valid && f->seg_ok) {
         const int p = f->done;
         if (f->seg_fail) { x.rects.push_back({A_L, fi, p + CW, f->m, p, p + CW, true}); x.rects.push_back({A_BK, fi, p + CW, f->m, 0, CW, false}); }
         else { x.rects.push_back({A_L, fi, p, p + CW, 0, p, true}); x.rects.push_back({A_L, fi, p, p + CW, p, p + CW, true}); }
      }
      issue_main(std::move(x));
   }
   void upd_seg(int fi, const HostState& h) {
      Front* f = &(*Fp)[fi];
      Launch x{"UPD_SEG", {}};
      if (f->seg_valid && f->seg_ok && !f->seg_fail) {
         upd_rects(x.rects, *f, fi, f->done, f->done + CW, f->done + CW, f->pend0, h.p0 / Ti, (h.pend0 - 1) / Ti, Ti);
         /* the chain / tiles update the segment's own columns; UPD_SEG the rest of the panel */
         for (int c = f->done; c < f->done + CW; ++c) upd[fi][c] = -1;           // eliminated
         for (int c = f->done + CW; c < f->pend0; ++c) upd[fi][c] += CW;
      }
      issue_main(std::move(x));
   }

   /* ---- the host side: factor_fronts of subtree.cu for one level of large fronts ---- */
   void factor_fronts(std::vector<Front>& F) {
      Fp = &F;
      upd.assign(F.size(), {});
      for (size_t i = 0; i < F.size(); ++i) upd[i].assign(F[i].n, 0);
      std::vector<HostState> H(F.size());
      for (size_t i = 0; i < F.size(); ++i) {
         HostState& h = H[i];
         h.fi = (int)i; h.n = F[i].n; h.m = F[i].m;
         h.done = 0; h.end = F[i].n; h.pass_start = 0;
         h.finished = (F[i].n == 0);
         h.p0 = 0; h.pend0 = std::min(PW, F[i].n); h.pend = h.pend0;
      }
      std::vector<int> snap_host;
      bool bulk_pending = false;
      for (int guard = 0;; ++guard) {
         if (guard > 100000) throw std::runtime_error("the panel loop does not terminate");
         std::vector<int> act;
         for (size_t i = 0; i < H.size(); ++i) if (!H[i].finished) act.push_back((int)i);
         if (act.empty()) {
            if (bulk_pending) wait_main(ev_bulk_all);
            break;
         }
         std::stable_sort(act.begin(), act.end(), [&](int a, int b) { return H[a].pend0 - H[a].p0 > H[b].pend0 - H[b].p0; });
         const int na_all = (int)act.size();
         std::vector<int> cand(na_all);
         for (int k = 0; k < na_all; ++k) cand[k] = H[act[k]].pend0 - H[act[k]].p0;
         const int nsteps = (cand[0] + BS - 1) / BS;
         auto count_gt = [&](int thr) { int c = 0; while (c < na_all && cand[c] > thr) ++c; return c; };
         auto take_snapshot = [&]() {
            snap_host.assign((size_t)na_all * 8, 0);
            for (int k = 0; k < na_all; ++k) snapshot_state(&F[H[act[k]].fi], &snap_host[(size_t)k * 8]);
            sync_main();                                    // cudaStreamSynchronize(s)
         };
         /* the columns a block step eliminates count as updated-by-everything (they leave the game) */
         auto account_elims = [&](int fi, int from, int to) { for (int c = from; c < to; ++c) upd[fi][c] = -1; };
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
               for (int k = 0; k < na_all; ++k) k_panel_chain(H[act[k]].fi, seg == 0);
               for (int k = 0; k < na_all; ++k) k_panel_tiles(H[act[k]].fi);
               for (int k = 0; k < na_all; ++k) k_seg_commit(H[act[k]].fi);
               if (seg + 1 < nseg) for (int k = 0; k < na_all; ++k) upd_seg(H[act[k]].fi, H[act[k]]);
               else for (int k = 0; k < na_all; ++k) {        // last segment: nothing of the panel is left to update
                  Front* f = &F[H[act[k]].fi];
                  if (f->seg_valid && f->seg_ok && !f->seg_fail) {
                     for (int c = f->done; c < f->done + CW; ++c) upd[H[act[k]].fi][c] = -1;
                     for (int c = f->done + CW; c < f->pend0; ++c) upd[H[act[k]].fi][c] += 0;   // (PW == 2 CW: none)
                     if (f->done + CW < f->pend0)
                        throw std::runtime_error("columns of the panel right of the last segment get no segment update");
                  }
               }
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
            for (int k = 0; k < na; ++k) k_diag(H[act[k]].fi, st == 0 && steps_new_panel);
            for (int k = 0; k < na; ++k) k_apply(H[act[k]].fi);
            for (int k = 0; k < na; ++k) k_commit(H[act[k]].fi);
            for (int k = 0; k < na; ++k) {
               Front* f = &F[H[act[k]].fi];
               upd_inner(H[act[k]].fi, H[act[k]]);
               if (f->step_valid) account_elims(H[act[k]].fi, f->done, f->done + calc_ne(f));
            }
            for (int k = 0; k < na; ++k) k_swap_inner(H[act[k]].fi);
         }
         if (!use_v2 || steps_todo > 0) take_snapshot();

         /* ---- what happened in the panel; outer update, look-ahead bulk, swaps (subtree.cu) ---- */
         std::vector<MatTile_> outer, bulk, bulk_b;
         struct Reg { int front, k0, k1, c_lo; };
         std::vector<Reg> bulk_regs;
         std::vector<int> swap_fronts;
         bool any_fail = false;
         for (int k = 0; k < na_all; ++k) {
            const int* sn = &snap_host[(size_t)k * 8];
            if (sn[6] >= 0 && H[act[k]].pend0 - sn[2] > 0) any_fail = true;
         }
         const bool lookahead = !any_fail && lookahead_on;          // big == true
         for (int k = 0; k < na_all; ++k) {
            HostState& h = H[act[k]];
            const int* sn = &snap_host[(size_t)k * 8];
            if (sn[0] != h.p0 || sn[3] != h.pend0 || sn[4] != h.end)
               throw std::runtime_error("host mirror of the pivoting state diverged from the device");
            h.done = sn[1]; h.pend = sn[2];
            h.spec_dead = sn[7] >= SPEC_MAX_FAILS;
            if (h.done != h.pend) throw std::runtime_error("a panel was left incomplete by the launches of the host");
            if (h.done > h.p0 && h.pend0 < h.n) {
               int mt = (h.m + T - 1) / T, nt = (h.n + T - 1) / T;
               int tj_urgent = (std::min(h.pend0 + PW, h.n) - 1) / T;
               if (inject == 3 && tj_urgent > h.pend0 / T) --tj_urgent;      // the last tile column of the next panel goes to the bulk
               int tj_next = (std::min(h.pend0 + 2 * PW, h.n) - 1) / T;
               bool has_bulk = lookahead && tj_urgent + 1 < nt;
               if (has_bulk) bulk_regs.push_back({h.fi, h.p0, h.done, (tj_urgent + 1) * T});
               for (int tj = h.pend0 / T; tj < nt; ++tj)
                  for (int ti = tj; ti < mt; ++ti) {
                     if (has_bulk && tj > tj_next) bulk_b.push_back({(int)bulk_regs.size() - 1, ti, tj});
                     else if (has_bulk && tj > tj_urgent) bulk.push_back({(int)bulk_regs.size() - 1, ti, tj});
                     else outer.push_back({h.fi, ti, tj});
                  }
            }
            if (h.pend0 - h.pend > 0 && h.end - h.pend0 > 0) swap_fronts.push_back(h.fi);
         }
         if (lookahead && (int)(bulk.size() + bulk_b.size()) < sm_count) {
            for (const MatTile_& t : bulk) outer.push_back({bulk_regs[t.front].front, t.ti, t.tj});
            for (const MatTile_& t : bulk_b) outer.push_back({bulk_regs[t.front].front, t.ti, t.tj});
            bulk.clear(); bulk_b.clear();
         }
         const bool have_bulk = !bulk.empty() || !bulk_b.empty();
         if (bulk_pending && (!outer.empty() || !swap_fronts.empty())) {
            const bool part = lookahead && have_bulk;
            if (!((inject == 1 && part) || (inject == 2 && !part))) wait_main(part ? ev_bulk : ev_bulk_all);
            if (!(lookahead && have_bulk)) bulk_pending = false;
         }
         /* tile lists -> per-front tile-column ranges (the lists are whole tile columns, rows tj .. mt - 1) */
         auto col_ranges = [&](const std::vector<MatTile_>& lst, bool explicit_regs) {
            std::vector<std::pair<int, std::pair<int, int>>> out;   // (front or region, [tj_lo, tj_hi])
            for (const MatTile_& t : lst) {
               bool found = false;
               for (auto& o : out) if (o.first == t.front) { o.second.first = std::min(o.second.first, t.tj); o.second.second = std::max(o.second.second, t.tj); found = true; }
               if (!found) out.push_back({t.front, {t.tj, t.tj}});
            }
            (void)explicit_regs;
            return out;
         };
         if (!outer.empty()) {
            Launch x{"UPD_OUTER", {}};
            for (auto& o : col_ranges(outer, false)) {
               const Front& f = F[o.first];
               /* device region (make_region, UPD_OUTER): K = [p0, done), columns [pend0, n) */
               if (!f.panel_open || f.finished) continue;
               const int tl = o.second.first, th = o.second.second;
               upd_rects(x.rects, f, o.first, f.p0, f.done, f.pend0, f.n, tl, th, T);
               for (int c = std::max(f.pend0, tl * T); c < std::min(f.n, (th + 1) * T); ++c) upd[o.first][c] += f.done - f.p0;
            }
            issue_main(std::move(x));
         }
         if (have_bulk) {
            for (int part = 0; part < 2; ++part) {
               const std::vector<MatTile_>& lst = part == 0 ? bulk : bulk_b;
               if (!lst.empty() && !(inject == 4 && part == 1)) {
                  Launch y{part == 0 ? "UPD_EXPLICIT(a)" : "UPD_EXPLICIT(b)", {}};
                  for (auto& o : col_ranges(lst, true)) {
                     const Reg& rg = bulk_regs[o.first];
                     const Front& f = F[rg.front];
                     const int tl = o.second.first, th = o.second.second;
                     upd_rects(y.rects, f, rg.front, rg.k0, rg.k1, rg.c_lo, f.n, tl, th, T);
                     for (int c = std::max(rg.c_lo, tl * T); c < std::min(f.n, (th + 1) * T); ++c) upd[rg.front][c] += rg.k1 - rg.k0;
                  }
                  issue_s2(std::move(y));
               }
               record(part == 0 ? ev_bulk : ev_bulk_all);
            }
            bulk_pending = true;
         }
         if (!swap_fronts.empty()) for (int fi : swap_fronts) k_swap_outer(fi);
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
      /* level end: k_finalize, then the Schur complement reads every eliminated column on the main stream */
      if (!s2_pending.empty()) throw std::runtime_error("bulk updates still in flight when the panel loop ends");
      for (size_t i = 0; i < F.size(); ++i) {
         advance_state(&F[i], true);
         if (!F[i].finished) throw std::runtime_error("front not finished after the panel loop");
         for (int c = 0; c < F[i].n; ++c) {
            if (c < F[i].nelim) { if (upd[i][c] != -1) throw std::runtime_error("an eliminated column is not marked"); }
            else if (upd[i][c] != F[i].nelim) throw std::runtime_error("a delayed column has not received every update");
         }
      }
   }
};

static int run(int ntrial, int inject, long* stats) {
   int failures = 0;
   for (int trial = 0; trial < ntrial; ++trial) {
      Sim sim(5000 + trial);
      sim.inject = inject;
      sim.v2 = trial % 3 == 2;
      const double pf[] = {0.0, 0.0, 0.02, 0.2};
      sim.p_fail = pf[sim.I(0, 3)];
      sim.p_chain_giveup = sim.U() < 0.6 ? 0.0 : sim.U() * 0.5;
      sim.p_tile_fail = sim.U() < 0.6 ? 0.0 : sim.U() * 0.5;
      sim.sm_count = sim.U() < 0.5 ? 148 : 8;            // small value: the "not worth a second stream" fold rarely triggers
      const int nfront = sim.I(1, 3);
      std::vector<Front> F(nfront);
      std::vector<BlockWS> ws(nfront);
      static SegWS* dummy = reinterpret_cast<SegWS*>(&ws);
      for (int i = 0; i < nfront; ++i) {
         Front& f = F[i];
         f = Front();
         const int n = sim.U() < 0.1 ? sim.I(1, 300) : sim.I(300, 2600);
         f.n = n; f.m = n + sim.I(0, 700); f.ldl = (f.m + 1) / 2 * 2;
         f.end = f.n; f.first_pass_done = -1;
         f.ws = &ws[i];
         f.sws = sim.v2 ? dummy : nullptr;
      }
      try {
         sim.factor_fronts(F);
         if (inject) continue;
      } catch (const std::exception& e) {
         if (inject) { ++failures; continue; }            // counted as "detected"
         printf("trial %d (v2=%d p_fail=%.2f): %s\n", trial, (int)sim.v2, sim.p_fail, e.what());
         ++failures;
      }
      stats[0] += sim.n_main; stats[1] += sim.n_bulk; stats[2] += sim.n_pairs; stats[3] += sim.n_waits;
   }
   return failures;
}

int main(int argc, char** argv) {
   const int ntrial = argc > 1 ? atoi(argv[1]) : 600;
   long st[4] = {0, 0, 0, 0}, st2[4] = {0, 0, 0, 0};
   const int failures = run(ntrial, false, st);
   /* the checker must see schedules that are wrong */
   int detected[5] = {0, 0, 0, 0, 0};
   bool blind = false;
   for (int inj = 1; inj <= 4; ++inj) { detected[inj] = run(std::min(ntrial, 300), inj, st2); blind = blind || detected[inj] == 0; }
   printf("lookahead_race_emu: %d levels, %ld main-stream launches, %ld bulk launches, %ld event waits, %ld footprint pairs "
          "compared, %d failures; fault injection detected in %d (ev_bulk wait dropped) / %d (ev_bulk_all wait dropped) / %d (bulk "
          "one tile column early) / %d (second bulk part lost) levels\n", ntrial, st[0], st[1], st[3], st[2], failures,
          detected[1], detected[2], detected[3], detected[4]);
   return (failures || blind) ? 1 : 0;
}
