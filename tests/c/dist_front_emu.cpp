/* TEST INFRASTRUCTURE (CPU only): model check of the PROTOCOL PLANNED for the distributed top fronts
 * ("bulk offload").  Nothing of it runs on a GPU yet: this program fixes the
 * protocol -- who copies / updates / signals what, on which stream, behind which flag -- before the CUDA
 * side is written, and will stay as its regression check.
 *
 * One top front, an owner rank R0 and a helper rank R1.  R0 runs factor_fronts (restated launch for launch,
 * on the shared state machine of spral_b200/csrc/pivot_state.h, as tests/c/lookahead_race_emu.cpp does) but
 * does not run the look-ahead bulk update itself while the split is active:
 *   set-up   R0 main:  copy the far columns (blocks J >= 2 of 256 columns) into R1's mirror (full shape of the
 *                      front, for L and for L*D), set `init`
 *   panel k  R0 main:  wait updated[k+1], copy block k+1 back out of the mirror, urgent update of block k+1
 *                      with panel k                                                           (as today)
 *            R0 copy:  copy L, L*D of panel k, rows >= 256 (k+2), into the same columns of the mirror, set ready[k]
 *            R1      :  wait ready[k]; update block k+2 with panel k (the UPD_EXPLICIT kernel against the mirror),
 *                      set updated[k+2]; update the blocks > k+2 with panel k
 *   failure  the first panel with a failed pivot ends the split by DRAINING: R0 sets ready[k] = DRAIN; R1, in
 *            stream order behind everything it was given, sets `drained`; R0 waits for it, copies every far
 *            column back and carries on alone (full outer update, swaps, local look-ahead).
 * This is the host-driven first version that spral_b200/csrc/split_front.h implements (flags raised from host
 * callbacks behind the stream work they announce, polled by the other host).
 * Streams are in-order; the only cross-stream orderings are R0's host synchronisation of its main stream
 * at each panel snapshot (which orders what R0's host issues afterwards, on either of its streams) and
 * the flags.  Checked, over random fronts / failure patterns / speculative segments:
 *   1. the flag graph has no deadlock (every wait is eventually satisfied);
 *   2. no two operations that are not ordered (vector clocks over the three streams) touch the same
 *      entries of R0's front, R1's mirror or R1's panel buffers unless both only read them;
 *   3. every column has received the update of every eliminated column when its block is factorised,
 *      and R1 applied exactly the panels R0 accounts for when a block comes back.
 * Injected faults (a wait dropped, a block announced before it is updated, the drain not awaited, a panel copy
 * that reaches into rows the owner is permuting) must be detected.
 *
 * Build: g++ -O2 -std=c++17 -I/usr/local/cuda/include -Iinclude tests/c/dist_front_emu.cpp */
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

enum Arr { A_L = 0, A_LD = 1, A_BK = 2, A_HM = 3, A_HMD = 4 };   // HM / HMD: R1's mirror of L / of L*D (full shape of the front)
enum Stream { S_MAIN = 0, S_COPY = 1, S_HELP = 2, NSTREAM = 3 };

struct Rect { int arr, front, r0, r1, c0, c1; bool w; };       // half-open; L / LD in front coordinates
struct Launch { const char* name; std::vector<Rect> rects; };
struct Op { int kind; Launch l; int flag; };          // kind 0 launch, 1 set flag, 2 wait flag

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
   int split_inject = 0;             // faults of the split protocol: 1 no wait for the updated block, 2 block announced before its update, 3 drain not awaited,
                                     // 4 the panel copy also reads the rows of the next panel (which R0 is permuting)
   int inject = 0;                   // fault injection (the checker must notice): 1 drop the ev_bulk wait, 2 drop the
                                     // ev_bulk_all wait, 3 the bulk starts one tile column early, 4 the second bulk part is lost
   long n_main = 0, n_bulk = 0, n_pairs = 0, n_waits = 0;
   explicit Sim(uint64_t seed) : rng(seed) {}
   double U() { return std::uniform_real_distribution<double>(0, 1)(rng); }
   int I(int lo, int hi) { return std::uniform_int_distribution<int>(lo, hi)(rng); }

   std::vector<Front>* Fp = nullptr;
   std::vector<std::vector<int>> upd;            // per front, per column: updates received (count of pivot columns)

   /* ---- stream model: operations are recorded per stream, ordered and checked afterwards ---- */
   std::vector<Op> ops[NSTREAM];
   int nflags = 0;
   int pending_sync = -1;                         // flag set at R0's last host sync of its main stream
   int synced_copy = -1;                          // ... the last one the copy stream already waits for
   int new_flag() { return nflags++; }
   void push_set(int st, int flag) { ops[st].push_back({1, {"set", {}}, flag}); }
   void push_wait(int st, int flag) { ops[st].push_back({2, {"wait", {}}, flag}); ++n_waits; }
   void issue_main(Launch&& x) { ++n_main; ops[S_MAIN].push_back({0, std::move(x), -1}); }
   void host_order_copy() {                       // what R0's host issues after a sync is behind everything synced
      if (pending_sync >= 0 && pending_sync != synced_copy) { push_wait(S_COPY, pending_sync); synced_copy = pending_sync; }
   }
   void issue_s2(Launch&& y) { ++n_bulk; host_order_copy(); ops[S_COPY].push_back({0, std::move(y), -1}); }
   void sync_main() { pending_sync = new_flag(); push_set(S_MAIN, pending_sync); }
   void record(long& ev) { host_order_copy(); ev = new_flag(); push_set(S_COPY, (int)ev); }
   void wait_main(long ev) { if (ev >= 0) push_wait(S_MAIN, (int)ev); }
   long ev_bulk = -1, ev_bulk_all = -1;

   /* ---- the split ---- */
   static int B(int j) { return j * PW; }         // first column of block j (no failed pivot so far: panel j == block j)
   bool offload_active = false;
   int drain_at = -1;                             // panel whose failure ended the split
   int flag_init = -1, flag_drained = -1;
   std::vector<int> flag_ready, flag_updated;     // by panel / by block (-1: never created)
   std::vector<int> accounted;                    // by block: panels R0 counted when the block came back
   long n_helper = 0, n_drains = 0, n_offloaded = 0;
   int& slot(std::vector<int>& v, int i) { if ((int)v.size() <= i) v.resize(i + 1, -1); if (v[i] < 0) v[i] = new_flag(); return v[i]; }
   void account_block(int J, int npanels, int n) {
      if ((int)accounted.size() <= J) accounted.resize(J + 1, -1);
      accounted[J] = npanels;
      for (int c = B(J); c < std::min(B(J + 1), n); ++c) upd[0][c] += npanels * PW;
   }
   /* R1's schedule (split_front.h: split_helper_serve): a function of the geometry and of the flags R0 raised */
   void helper_schedule(const Front& f) {
      if (flag_init < 0) return;
      push_wait(S_HELP, flag_init);
      std::vector<int> cnt(64 + f.n / PW, 0);
      for (int k = 0; k < (int)flag_ready.size(); ++k) {
         if (flag_ready[k] < 0) break;
         push_wait(S_HELP, flag_ready[k]);
         if (k == drain_at) {                               /* in stream order behind everything it was given */
            for (int J = std::max(2, k + 1); B(J) < f.n; ++J) check_count(J, cnt[J]);
            push_set(S_HELP, flag_drained);
            return;
         }
         for (int J = k + 2; B(J) < f.n; ++J) {
            if (J == k + 2 && split_inject == 2) push_set(S_HELP, slot(flag_updated, J));      // fault: announced too early
            Launch y{"helper update", {}};
            y.rects.push_back({A_HM, 0, B(J), f.ldl, panel_k0[k], panel_k1[k], false});
            y.rects.push_back({A_HMD, 0, B(J), std::min(f.ldl, B(J + 1)), panel_k0[k], panel_k1[k], false});
            y.rects.push_back({A_HM, 0, B(J), f.m, B(J), std::min(B(J + 1), f.n), true});
            ops[S_HELP].push_back({0, std::move(y), -1}); ++n_helper;
            ++cnt[J];
            if (J == k + 2) {
               if (split_inject != 2) push_set(S_HELP, slot(flag_updated, J));
               if ((int)accounted.size() > J && accounted[J] >= 0 && !(drain_at >= 0 && J > drain_at)) check_count(J, cnt[J]);
            }
         }
      }
   }
   void check_count(int J, int applied) {
      if ((int)accounted.size() <= J || accounted[J] != applied) {
         char buf[160];
         snprintf(buf, sizeof buf, "block %d comes back with %d panels applied, R0 accounts for %d", J, applied,
                  (int)accounted.size() > J ? accounted[J] : -1);
         throw std::runtime_error(buf);
      }
   }
   std::vector<int> panel_k0, panel_k1;           // columns of the panels R0 pushed

   /* Vector clocks + pairwise check.  Returns the number of unordered pairs compared. */
   void order_and_check() {
      struct Clk { long v[NSTREAM]; };
      std::vector<Clk> setclk(nflags, Clk{{-1, -1, -1}});
      std::vector<char> isset(nflags, 0);
      std::vector<Clk> clk[NSTREAM];
      size_t pos[NSTREAM] = {0, 0, 0};
      Clk cur[NSTREAM];
      for (int st = 0; st < NSTREAM; ++st) { clk[st].resize(ops[st].size()); for (int x = 0; x < NSTREAM; ++x) cur[st].v[x] = -1; }
      for (bool progress = true; progress;) {
         progress = false;
         for (int st = 0; st < NSTREAM; ++st)
            while (pos[st] < ops[st].size()) {
               Op& o = ops[st][pos[st]];
               if (o.kind == 2) {
                  if (!isset[o.flag]) break;                       // blocked
                  for (int x = 0; x < NSTREAM; ++x) cur[st].v[x] = std::max(cur[st].v[x], setclk[o.flag].v[x]);
               }
               cur[st].v[st] = (long)pos[st];
               clk[st][pos[st]] = cur[st];
               if (o.kind == 1) { setclk[o.flag] = cur[st]; isset[o.flag] = 1; }
               ++pos[st];
               progress = true;
            }
      }
      for (int st = 0; st < NSTREAM; ++st)
         if (pos[st] < ops[st].size()) {
            char buf[200];
            snprintf(buf, sizeof buf, "deadlock: stream %d waits for flag %d that is never set (operation %zu of %zu)", st,
                     ops[st][pos[st]].flag, pos[st], ops[st].size());
            throw std::runtime_error(buf);
         }
      for (int s1 = 0; s1 < NSTREAM; ++s1)
         for (int s2 = s1 + 1; s2 < NSTREAM; ++s2)
            for (size_t i = 0; i < ops[s1].size(); ++i) {
               if (ops[s1][i].kind != 0 || ops[s1][i].l.rects.empty()) continue;
               for (size_t j = 0; j < ops[s2].size(); ++j) {
                  if (ops[s2][j].kind != 0 || ops[s2][j].l.rects.empty()) continue;
                  if (clk[s2][j].v[s1] >= (long)i || clk[s1][i].v[s2] >= (long)j) continue;      // ordered
                  check(ops[s1][i].l, ops[s2][j].l, s1, s2);
               }
            }
   }
   void check(const Launch& x, const Launch& y, int s1, int s2) {
      for (const Rect& a : x.rects)
         for (const Rect& b : y.rects) {
            ++n_pairs;
            if (conflict(a, b)) {
               char buf[512];
               snprintf(buf, sizeof buf, "race: %s {arr %d rows [%d,%d) cols [%d,%d) %s} on stream %d vs %s "
                        "{rows [%d,%d) cols [%d,%d) %s} on stream %d", x.name, a.arr, a.r0, a.r1, a.c0, a.c1,
                        a.w ? "W" : "R", s1, y.name, b.r0, b.r1, b.c0, b.c1, b.w ? "W" : "R", s2);
               throw std::runtime_error(buf);
            }
         }
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
      if (F.size() == 1 && F[0].n > 2 * PW) {          /* set-up of the split: the far columns go to R1's mirror */
         Launch x{"init copy", {}};
         x.rects.push_back({A_L, 0, B(2), F[0].m, B(2), F[0].n, false});
         x.rects.push_back({A_HM, 0, B(2), F[0].m, B(2), F[0].n, true});
         issue_main(std::move(x));
         flag_init = new_flag(); flag_drained = new_flag();
         push_set(S_MAIN, flag_init);
         offload_active = true;
      }
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
         const bool split_now = offload_active && !any_fail;       // this panel's far update goes to R1
         const int kpanel = H[act[0]].p0 / PW;
         if (offload_active) {
            const HostState& h = H[act[0]];
            if (h.p0 % PW != 0) throw std::runtime_error("split active although the panels are no longer block aligned");
            if (any_fail) {                                         /* drain: everything comes back, R0 carries on alone */
               host_order_copy();
               push_set(S_COPY, slot(flag_ready, kpanel));
               drain_at = kpanel; ++n_drains;
               if (split_inject != 3) push_wait(S_MAIN, flag_drained);
               const int J0 = std::max(2, kpanel + 1);
               if (B(J0) < h.n) {
                  Launch x{"pull the far columns", {}};
                  x.rects.push_back({A_HM, 0, B(J0), F[0].m, B(J0), h.n, false});
                  x.rects.push_back({A_L, 0, B(J0), F[0].m, B(J0), h.n, true});
                  issue_main(std::move(x));
               }
               for (int J = J0; B(J) < h.n; ++J) account_block(J, std::min(J - 1, kpanel), h.n);
               offload_active = false;
            }
         }
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
         if (lookahead && !split_now && (int)(bulk.size() + bulk_b.size()) < sm_count) {
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
         if (split_now) {
            const HostState& h = H[act[0]];
            const int J = kpanel + 1;                               // the block the urgent update is about to touch
            if (J >= 2 && B(J) < h.n) {
               if (split_inject != 1) push_wait(S_MAIN, slot(flag_updated, J));
               else slot(flag_updated, J);
               Launch x{"pull block", {}};
               x.rects.push_back({A_HM, 0, B(J), F[0].m, B(J), std::min(B(J + 1), h.n), false});
               x.rects.push_back({A_L, 0, B(J), F[0].m, B(J), std::min(B(J + 1), h.n), true});
               issue_main(std::move(x));
               account_block(J, J - 1, h.n);
            }
         }
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
         if (split_now) {
            const HostState& h = H[act[0]];
            const Front& f = F[0];
            if (h.done != h.pend0) throw std::runtime_error("split: a panel without failure is not complete");
            if (B(kpanel + 2) < h.n) {                              /* panel k to R1: rows of the far blocks only */
               host_order_copy();
               Launch y{"copy panel", {}};
               const int rtop = B(kpanel + (split_inject == 4 ? 1 : 2));      // fault 4: the rows of the NEXT panel are copied too
               y.rects.push_back({A_L, 0, rtop, f.ldl, h.p0, h.done, false});
               y.rects.push_back({A_LD, 0, rtop, f.ldl, h.p0, h.done, false});
               y.rects.push_back({A_HM, 0, rtop, f.ldl, h.p0, h.done, true});
               y.rects.push_back({A_HMD, 0, rtop, f.ldl, h.p0, h.done, true});
               if ((int)panel_k0.size() <= kpanel) { panel_k0.resize(kpanel + 1, 0); panel_k1.resize(kpanel + 1, 0); }
               panel_k0[kpanel] = h.p0; panel_k1[kpanel] = h.done;
               ops[S_COPY].push_back({0, std::move(y), -1}); ++n_offloaded;
               push_set(S_COPY, slot(flag_ready, kpanel));
               /* the far columns get this panel from R1: nothing to add here, account_block() does it when they return */
            } else offload_active = false;                          // nothing is left on R1
         } else if (have_bulk) {
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
      if (offload_active) throw std::runtime_error("the split is still active when the front is finished");
      helper_schedule(F[0]);
      order_and_check();
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
      Sim sim(9000 + trial);
      sim.inject = 0;                                // (the look-ahead faults of lookahead_race_emu.cpp are not used here)
      sim.v2 = trial % 3 == 2;
      const double pf[] = {0.0, 0.0, 0.004, 0.03};
      sim.p_fail = pf[sim.I(0, 3)];
      sim.p_chain_giveup = sim.U() < 0.6 ? 0.0 : sim.U() * 0.5;
      sim.p_tile_fail = sim.U() < 0.6 ? 0.0 : sim.U() * 0.5;
      sim.sm_count = sim.U() < 0.5 ? 148 : 8;
      std::vector<Front> F(1);
      std::vector<BlockWS> ws(1);
      static SegWS* dummy = reinterpret_cast<SegWS*>(&ws);
      Front& f = F[0];
      f = Front();
      const int n = sim.U() < 0.15 ? sim.I(1, 600) : sim.I(600, 5000);
      f.n = n; f.m = n + (sim.U() < 0.3 ? 0 : sim.I(0, 900)); f.ldl = (f.m + 1) / 2 * 2;
      f.end = f.n; f.first_pass_done = -1;
      f.ws = &ws[0];
      f.sws = sim.v2 ? dummy : nullptr;
      sim.split_inject = inject;
      try {
         sim.factor_fronts(F);
         if (inject) continue;
      } catch (const std::exception& e) {
         if (inject) { ++failures; continue; }            // counted as "detected"
         printf("trial %d (n=%d m=%d v2=%d p_fail=%.3f): %s\n", trial, f.n, f.m, (int)sim.v2, sim.p_fail, e.what());
         ++failures;
      }
      stats[0] += sim.n_main; stats[1] += sim.n_offloaded; stats[2] += sim.n_pairs; stats[3] += sim.n_waits;
      stats[4] += sim.n_helper; stats[5] += sim.n_drains;
   }
   return failures;
}

int main(int argc, char** argv) {
   const int ntrial = argc > 1 ? atoi(argv[1]) : 400;
   long st[6] = {0, 0, 0, 0, 0, 0}, st2[6] = {0, 0, 0, 0, 0, 0};
   const int failures = run(ntrial, 0, st);
   /* the checker must see protocols that are wrong */
   int detected[5] = {0, 0, 0, 0, 0};
   bool blind = false;
   for (int inj = 1; inj <= 4; ++inj) { detected[inj] = run(std::min(ntrial, 300), inj, st2); blind = blind || detected[inj] == 0; }
   printf("dist_front_emu: %d fronts, %ld owner launches, %ld panels sent, %ld helper operations, %ld drains, %ld flag waits, "
          "%ld unordered footprint pairs compared, %d failures; fault injection detected in %d (wait for the updated block "
          "dropped) / %d (block announced before its update) / %d (drain not awaited) / %d (panel copy reaches into the rows R0 is permuting) fronts\n", ntrial, st[0], st[1],
          st[4], st[5], st[3], st[2], failures, detected[1], detected[2], detected[3], detected[4]);
   return (failures || blind) ? 1 : 0;
}
