/* Host pre-processing for ssids_factor's options%scaling:
 * matching-based (MC64 / "Hungarian") and infinity-norm equilibration scalings of a
 * symmetric matrix held as its lower triangle in CSC.  They produce the vector that
 * ssids_factor applies as  S A S  (src/ssids/ssids.f90:861-1028) and that the numeric
 * engine takes through its `scaling` argument.
 *
 * Restated from the reference (ralna/spral):
 *   hungarian_scale_sym  src/scaling.f90:134-170  -> hungarian_wrapper :597-801
 *       log|a|, column maxima, cost c_ij = cmax_j - log|a_ij|, duals u, v of the
 *       minimum-sum assignment, s_i = exp((u_i + v_i - cmax_i) / 2); structurally
 *       singular matrices: matching on the matched sub-matrix and the Duff-Pralet
 *       correction s_i = 1 / max_k |a_ik s_k| for the unmatched variables (:708-801)
 *   equilib_scale_sym    src/scaling.f90:480-521 (Knight, Ruiz, Ucar, Algorithm 1)
 * The assignment itself (hungarian_match, :938-1171, adapted there from HSL_MC64) is
 * NOT a line-by-line restatement: this file solves the same linear assignment problem
 * with a sparse shortest-augmenting-path method (row-minima / greedy start, Dijkstra
 * with a binary heap, MC64 dual update).  The optimal value and the defining property of
 * the scaling (|s_i a_ij s_j| <= 1) are the same; where the optimal duals are not unique
 * the scaling factors can differ from the reference's (parity unpinned: no Fortran
 * compiler here; tests check optimality against scipy and the scaling property).
 */
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <queue>
#include <utility>
#include <vector>

#include "spral_ssids_b200.h"

namespace {

constexpr double RINF = std::numeric_limits<double>::max();

struct Csc {
   int n = 0;
   std::vector<int64_t> ptr;      // 0-based
   std::vector<int> row;          // 0-based
   std::vector<double> val;
};

/* Minimum-sum assignment on the sparse cost matrix C (costs >= 0).  On return
 * rowmatch[i] = column matched to row i or -1, u / v are dual variables with
 * u_i + v_j <= c_ij for every entry and equality on the matched ones.  Returns
 * the number of matched columns. */
int assignment(const Csc& C, std::vector<int>& rowmatch, std::vector<double>& u, std::vector<double>& v) {
   const int n = C.n;
   std::vector<int> colmatch(n, -1);
   std::vector<int64_t> colent(n, -1);
   rowmatch.assign(n, -1);
   u.assign(n, RINF);
   v.assign(n, 0.0);
   for (int j = 0; j < n; ++j)
      for (int64_t e = C.ptr[j]; e < C.ptr[j + 1]; ++e) u[C.row[e]] = std::min(u[C.row[e]], C.val[e]);
   for (int i = 0; i < n; ++i) if (u[i] == RINF) u[i] = 0.0;
   int matched = 0;
   for (int j = 0; j < n; ++j) {
      double vmin = RINF;
      for (int64_t e = C.ptr[j]; e < C.ptr[j + 1]; ++e) vmin = std::min(vmin, C.val[e] - u[C.row[e]]);
      v[j] = (vmin == RINF) ? 0.0 : vmin;
      for (int64_t e = C.ptr[j]; e < C.ptr[j + 1]; ++e) {
         const int i = C.row[e];
         if (rowmatch[i] < 0 && C.val[e] - u[i] - v[j] <= 0.0) {
            rowmatch[i] = j; colmatch[j] = i; colent[j] = e; ++matched;
            break;
         }
      }
   }
   std::vector<double> d(n, RINF);
   std::vector<int> pred_col(n, -1);
   std::vector<int64_t> pred_ent(n, -1);
   std::vector<char> done(n, 0);
   std::vector<int> touched, tree_rows, tree_cols;
   using Item = std::pair<double, int>;
   for (int j0 = 0; j0 < n; ++j0) {
      if (colmatch[j0] >= 0) continue;
      std::priority_queue<Item, std::vector<Item>, std::greater<Item>> heap;
      touched.clear(); tree_rows.clear(); tree_cols.clear();
      int j = j0, sink = -1;
      double lsp = 0.0;
      for (;;) {
         tree_cols.push_back(j);
         for (int64_t e = C.ptr[j]; e < C.ptr[j + 1]; ++e) {
            const int i = C.row[e];
            if (done[i]) continue;
            const double dn = lsp + (C.val[e] - u[i] - v[j]);
            if (dn < d[i]) {
               if (d[i] == RINF) touched.push_back(i);
               d[i] = dn; pred_col[i] = j; pred_ent[i] = e;
               heap.push({dn, i});
            }
         }
         int i = -1;
         while (!heap.empty()) {
            Item it = heap.top(); heap.pop();
            if (!done[it.second] && it.first == d[it.second]) { i = it.second; break; }
         }
         if (i < 0) break;                        // no augmenting path: column j0 stays unmatched
         done[i] = 1; tree_rows.push_back(i);
         lsp = d[i];
         if (rowmatch[i] < 0) { sink = i; break; }
         j = rowmatch[i];
      }
      if (sink >= 0) {
         const double lsap = d[sink];
         for (int i : tree_rows) u[i] += d[i] - lsap;
         for (int i = sink;;) {                   // augment along the predecessor columns
            const int jc = pred_col[i];
            const int prev = colmatch[jc];
            rowmatch[i] = jc; colmatch[jc] = i; colent[jc] = pred_ent[i];
            if (jc == j0) break;
            i = prev;
         }
         for (int jc : tree_cols)
            if (colmatch[jc] >= 0) v[jc] = C.val[colent[jc]] - u[colmatch[jc]];
         ++matched;
      }
      for (int i : touched) { d[i] = RINF; done[i] = 0; }
   }
   return matched;
}

/* lower triangle (1-based CSC, explicit zeros dropped) -> full symmetric pattern with log|a| */
Csc expand_log_abs(int n, const int64_t* ptr, const int* row, const double* val) {
   Csc F;
   F.n = n;
   std::vector<int64_t> cnt(n + 1, 0);
   for (int j = 0; j < n; ++j)
      for (int64_t e = ptr[j] - 1; e < ptr[j + 1] - 1; ++e) {
         if (val[e] == 0.0) continue;
         const int i = row[e] - 1;
         cnt[j + 1]++;
         if (i != j) cnt[i + 1]++;
      }
   F.ptr.assign(n + 1, 0);
   for (int j = 0; j < n; ++j) F.ptr[j + 1] = F.ptr[j] + cnt[j + 1];
   F.row.resize(F.ptr[n]); F.val.resize(F.ptr[n]);
   std::vector<int64_t> pos(F.ptr.begin(), F.ptr.end() - 1);
   for (int j = 0; j < n; ++j)
      for (int64_t e = ptr[j] - 1; e < ptr[j + 1] - 1; ++e) {
         if (val[e] == 0.0) continue;
         const int i = row[e] - 1;
         const double lv = std::log(std::fabs(val[e]));
         F.row[pos[j]] = i; F.val[pos[j]++] = lv;
         if (i != j) { F.row[pos[i]] = j; F.val[pos[i]++] = lv; }
      }
   return F;
}

} // namespace

extern "C" {

/* hungarian_scale_sym (src/scaling.f90:134-170).  match (may be NULL): match[i] = 1-based
 * column matched to row i (src/scaling.f90:597-801 conventions: after the singular-case
 * post-processing unmatched variables carry a negative number).  Returns inform%flag:
 * 0, 1 = WARNING_SINGULAR (scale_if_singular), -2 = ERROR_SINGULAR (identity scaling). */
int spral_ssids_b200_hungarian_scale_sym(int n, const int64_t* ptr, const int* row, const double* val,
      double* scaling, int* match, int scale_if_singular, int* matched_out) {
   Csc C = expand_log_abs(n, ptr, row, val);
   std::vector<double> cmax(n, 0.0);
   for (int j = 0; j < n; ++j) {
      double mx = -RINF;
      for (int64_t e = C.ptr[j]; e < C.ptr[j + 1]; ++e) mx = std::max(mx, C.val[e]);
      cmax[j] = (C.ptr[j] == C.ptr[j + 1]) ? 0.0 : mx;
      for (int64_t e = C.ptr[j]; e < C.ptr[j + 1]; ++e) C.val[e] = cmax[j] - C.val[e];
   }
   std::vector<int> rowmatch;
   std::vector<double> u, v;
   const int matched = assignment(C, rowmatch, u, v);
   if (matched_out) *matched_out = matched;
   int flag = 0;
   if (matched != n) {
      if (!scale_if_singular) {
         for (int i = 0; i < n; ++i) scaling[i] = 1.0;       // exp((0 + 0) / 2)
         if (match) for (int i = 0; i < n; ++i) match[i] = rowmatch[i] >= 0 ? rowmatch[i] + 1 : -1;
         return -2;
      }
      flag = 1;
   }
   if (matched == n) {
      for (int i = 0; i < n; ++i) scaling[i] = std::exp((u[i] + v[i] - cmax[i]) / 2);
      if (match) for (int i = 0; i < n; ++i) match[i] = rowmatch[i] + 1;
      return flag;
   }
   /* structurally rank deficient: matching on the sub-matrix of the matched variables,
    * then the Duff-Pralet correction for the others (src/scaling.f90:690-801) */
   std::vector<int> old_to_new(n), new_to_old;
   {
      int jn = matched + 1;
      for (int i = 0; i < n; ++i) {
         if (rowmatch[i] < 0) old_to_new[i] = -(jn++);
         else { old_to_new[i] = (int)new_to_old.size(); new_to_old.push_back(i); }
      }
   }
   Csc Sub;
   Sub.n = (int)new_to_old.size();
   Sub.ptr.assign(1, 0);
   for (int j = 0; j < n; ++j) {
      if (rowmatch[j] < 0) continue;
      for (int64_t e = C.ptr[j]; e < C.ptr[j + 1]; ++e) {
         const int i = C.row[e];
         if (rowmatch[i] < 0) continue;
         Sub.row.push_back(old_to_new[i]); Sub.val.push_back(C.val[e]);
      }
      Sub.ptr.push_back((int64_t)Sub.row.size());
   }
   std::vector<int> cperm;
   std::vector<double> du, dv;
   assignment(Sub, cperm, du, dv);
   std::vector<double> rs(n);
   for (int i = 0; i < n; ++i) {
      const int j = old_to_new[i];
      rs[i] = (j < 0) ? -RINF : (du[j] + dv[j] - cmax[i]) / 2;
   }
   if (match) {
      for (int i = 0; i < n; ++i) match[i] = -1;
      for (int k = 0; k < Sub.n; ++k) match[new_to_old[k]] = cperm[k] >= 0 ? new_to_old[cperm[k]] + 1 : -1;
      for (int i = 0; i < n; ++i) if (old_to_new[i] < 0) match[i] = old_to_new[i];
   }
   std::vector<double> cs(rs);
   for (int j = 0; j < n; ++j)
      for (int64_t e = ptr[j] - 1; e < ptr[j + 1] - 1; ++e) {
         if (val[e] == 0.0) continue;
         const int k = row[e] - 1;
         const double la = std::log(std::fabs(val[e]));
         if (cs[j] == -RINF && cs[k] != -RINF) rs[j] = std::max(rs[j], la + rs[k]);
         if (cs[k] == -RINF && cs[j] != -RINF) rs[k] = std::max(rs[k], la + rs[j]);
      }
   for (int i = 0; i < n; ++i) {
      if (cs[i] != -RINF) continue;
      rs[i] = (rs[i] == -RINF) ? 0.0 : -rs[i];
   }
   for (int i = 0; i < n; ++i) scaling[i] = std::exp(rs[i]);
   return flag;
}

/* auction_scale_sym (src/scaling.f90:269-309) -> auction_match (:1481-1586) -> auction_match_core
 * (:1328-1466), restated line by line: an approximate maximum-product matching by single-column auction
 * rounds with epsilon scaling; the row / column prices become the scaling after match_postproc
 * (:1588-1696, square case: the two dual vectors are shifted to the same mean).  opts: max_iterations,
 * max_unchanged[3], min_proportion[3], eps_initial as auction_options (:33-38; NULL = defaults 30000,
 * {10,100,100}, {0.90,0,0}, 0.01).  match[i] (may be NULL) = 1-based column matched to row i, 0 if none.
 * Returns 0. */
struct spral_ssids_b200_auction_options_ { int max_iterations; int max_unchanged[3]; float min_proportion[3]; float eps_initial; };
int spral_ssids_b200_auction_scale_sym(int n, const int64_t* ptr, const int* row, const double* val,
      double* scaling, int* match, const void* opts_in, int* matched_out, int* iterations_out) {
   spral_ssids_b200_auction_options_ o{30000, {10, 100, 100}, {0.90f, 0.0f, 0.0f}, 0.01f};
   if (opts_in) o = *static_cast<const spral_ssids_b200_auction_options_*>(opts_in);
   /* expand, drop explicit zeros, log|a|; cost = maxentry - (cmax - log|a|) */
   Csc C = expand_log_abs(n, ptr, row, val);
   std::vector<double> cmax(n, 0.0), rsc(n, 0.0), csc(n, 0.0);
   double maxentry = -RINF;
   for (int j = 0; j < n; ++j) {
      if (C.ptr[j + 1] <= C.ptr[j]) { cmax[j] = 0.0; continue; }
      double mx = -RINF;
      for (int64_t e = C.ptr[j]; e < C.ptr[j + 1]; ++e) mx = std::max(mx, C.val[e]);
      cmax[j] = mx;
      for (int64_t e = C.ptr[j]; e < C.ptr[j + 1]; ++e) { C.val[e] = mx - C.val[e]; maxentry = std::max(maxentry, C.val[e]); }
   }
   if (maxentry == -RINF) maxentry = 0.0;
   maxentry = 2 * maxentry + 1;               // prefers high-cardinality matchings (+1 avoids zero columns)
   for (double& x : C.val) x = maxentry - x;
   for (int j = 0; j < n; ++j) csc[j] = -cmax[j];
   /* ---- auction_match_core ---- */
   std::vector<int> cmatch(n, 0), owner(n, -1), next(n);
   std::vector<double>& dualu = rsc;
   std::vector<double>& dualv = csc;
   int unmatched = n, prev = -1, nunchanged = 0, tail = n, itr = 1;
   for (int i = 0; i < n; ++i) next[i] = i;
   double eps = o.eps_initial;
   for (; itr <= o.max_iterations; ++itr) {
      if (unmatched == 0) break;
      if (unmatched != prev) nunchanged = 0;
      prev = unmatched;
      ++nunchanged;
      const float prop = (float)(n - unmatched) / (float)n;
      if (nunchanged >= o.max_unchanged[0] && prop >= o.min_proportion[0]) break;
      if (nunchanged >= o.max_unchanged[1] && prop >= o.min_proportion[1]) break;
      if (nunchanged >= o.max_unchanged[2] && prop >= o.min_proportion[2]) break;
      eps = std::min(1.0, eps + 1.0 / (n + 1));
      int insert = 0;
      for (int cp = 0; cp < tail; ++cp) {
         const int col = next[cp];
         if (cmatch[col] != 0) continue;                          // matched (> 0) or ineligible (-1)
         if (C.ptr[col] == C.ptr[col + 1]) continue;              // empty column
         int64_t e = C.ptr[col];
         int bestr = C.row[e];
         double bestu = C.val[e] - dualu[bestr], bestv = -RINF;
         for (e = C.ptr[col] + 1; e < C.ptr[col + 1]; ++e) {
            const double u = C.val[e] - dualu[C.row[e]];
            if (u > bestu) { bestv = bestu; bestr = C.row[e]; bestu = u; }
            else if (u > bestv) bestv = u;
         }
         if (bestv == -RINF) bestv = 0.0;                         // no second best
         if (bestu > 0) {
            dualu[bestr] += bestu - bestv + eps;
            dualv[col] = bestv - eps;
            cmatch[col] = bestr + 1;
            --unmatched;
            const int k = owner[bestr];
            owner[bestr] = col;
            if (k >= 0) { cmatch[k] = 0; ++unmatched; next[insert++] = k; }
         } else {
            cmatch[col] = -1;                                     // no net benefit: never considered again
            --unmatched;
         }
      }
      tail = insert;
   }
   if (iterations_out) *iterations_out = itr - 1;
   int matched = 0;
   for (int j = 0; j < n; ++j) { if (cmatch[j] == -1) cmatch[j] = 0; if (cmatch[j] != 0) ++matched; }
   if (matched_out) *matched_out = matched;
   /* undo the pre-processing; match_postproc for a square matrix */
   for (int i = 0; i < n; ++i) { rsc[i] = -rsc[i] + maxentry; csc[i] = -csc[i] - cmax[i]; }
   if (match) {
      for (int i = 0; i < n; ++i) match[i] = 0;
      for (int j = 0; j < n; ++j) if (cmatch[j] != 0) match[cmatch[j] - 1] = j + 1;
   }
   double ravg = 0, cavg = 0;
   for (int i = 0; i < n; ++i) { ravg += rsc[i]; cavg += csc[i]; }
   if (n > 0) { ravg /= n; cavg /= n; }
   const double adjust = (ravg - cavg) / 2;
   for (int i = 0; i < n; ++i) scaling[i] = std::exp(((rsc[i] - adjust) + (csc[i] + adjust)) / 2);
   return 0;
}

/* match_order_metis (src/match_order.f90:51-208): matching-based ordering for options%ordering = 2.
 * The MC64-type matching and scaling (mo_scale / mo_match there; hungarian_scale_sym here, with
 * scale_if_singular), then mo_split (:220-396): the matching is split into 1- and 2-cycles, the matrix
 * is condensed so that a matched pair is one vertex, METIS orders the condensed graph, and the order is
 * expanded so that the two variables of a pair are consecutive -- the 2x2 pivot the matching suggests is
 * available to the factorisation without delays.  order[i] = position of variable i+1 in the pivot
 * sequence (1-based).  Returns 0, 1 (structurally singular: warning) or a negative flag. */
int spral_ssids_b200_match_order_metis(int n, const int64_t* ptr, const int* row, const double* val,
      int* order, double* scaling) {
   if (n < 0) return -1;
   if (n == 0) return 0;
   std::vector<int> cperm(n);
   int matched = 0;
   int flag = spral_ssids_b200_hungarian_scale_sym(n, ptr, row, val, scaling, cperm.data(), 1, &matched);
   if (flag < 0) return flag;
   for (int i = 0; i < n; ++i) if (cperm[i] < 0) cperm[i] = -1;
   /* full pattern (both triangles), explicit zeros dropped, 1-based like the reference */
   std::vector<int64_t> ptr2(n + 2, 0);
   for (int j = 0; j < n; ++j)
      for (int64_t e = ptr[j] - 1; e < ptr[j + 1] - 1; ++e) {
         if (val[e] == 0.0) continue;
         const int i = row[e] - 1;
         ptr2[j + 2]++;
         if (i != j) ptr2[i + 2]++;
      }
   ptr2[1] = 1;
   for (int j = 1; j <= n; ++j) ptr2[j + 1] += ptr2[j];          // ptr2[j] = start of column j (1-based)
   std::vector<int> row2((size_t)(ptr2[n + 1] - 1) + 1);
   {
      std::vector<int64_t> pos(ptr2.begin(), ptr2.end());
      for (int j = 0; j < n; ++j)
         for (int64_t e = ptr[j] - 1; e < ptr[j + 1] - 1; ++e) {
            if (val[e] == 0.0) continue;
            const int i = row[e] - 1;
            row2[pos[j + 1]++] = i + 1;
            if (i != j) row2[pos[i + 1]++] = j + 1;
         }
   }
   /* ---- mo_split ---- (1-based arrays as in the reference) */
   std::vector<int> iwork(n + 1, 0), cp(n + 1);
   for (int i = 1; i <= n; ++i) cp[i] = cperm[i - 1];
   for (int i = 1; i <= n; ++i) {
      if (iwork[i] != 0) continue;
      int j = i;
      for (;;) {
         if (cp[j] == -1) { iwork[j] = -2; break; }               // unmatched
         else if (cp[j] == i) { iwork[j] = -1; break; }           // singleton (or the end of an odd cycle)
         const int jj = cp[j];
         iwork[j] = jj; iwork[jj] = j;                            // pair j with cperm(j)
         j = cp[jj];
         if (j == i) break;
      }
   }
   for (int i = 1; i <= n; ++i) cp[i] = iwork[i];
   std::vector<int> old_to_new(n + 1, 0), new_to_old(n + 1, 0);
   int k = 1;
   for (int i = 1; i <= n; ++i) {
      const int j = cp[i];
      if (j < i && j > 0) continue;
      old_to_new[i] = k; new_to_old[k] = i;
      if (j > 0) old_to_new[j] = k;
      ++k;
   }
   const int ncomp_matched = k - 1;
   std::vector<int64_t> ptr3(n + 2, 0);
   std::vector<int> row3(row2.size() + 1);
   std::fill(iwork.begin(), iwork.end(), 0);
   ptr3[1] = 1;
   int ncomp = 1;
   int64_t jj = 1;
   for (int i = 1; i <= n; ++i) {
      const int j = cp[i];
      if (j < i && j > 0) continue;
      for (int pass = 0; pass < 2; ++pass) {
         const int col = pass == 0 ? i : j;
         if (pass == 1 && j <= 0) break;
         for (int64_t kl = ptr2[col]; kl < ptr2[col + 1]; ++kl) {
            const int krow = old_to_new[row2[kl]];
            if (iwork[krow] == i) continue;
            if (krow > ncomp_matched) continue;
            row3[jj++] = krow;
            iwork[krow] = i;
         }
      }
      ptr3[ncomp + 1] = jj;
      ++ncomp;
   }
   --ncomp;
   /* lower triangle of the condensed pattern for metis_order */
   {
      int64_t out = 1, j1 = 1;
      for (int i = 1; i <= ncomp; ++i) {
         const int64_t j2 = ptr3[i + 1];
         for (int64_t q = j1; q < j2; ++q) {
            const int krow = row3[q];
            if (krow < i) continue;
            row3[out++] = krow;
         }
         ptr3[i + 1] = out;
         j1 = j2;
      }
   }
   std::vector<int> corder(ncomp);
   {
      std::vector<int64_t> p3((size_t)ncomp + 1);
      for (int i = 0; i <= ncomp; ++i) p3[i] = ptr3[i + 1];
      std::vector<int> r3((size_t)std::max<int64_t>(p3[ncomp] - 1, 1), 0);
      for (int64_t q = 1; q < p3[ncomp]; ++q) r3[q - 1] = row3[q];
      const int rc = spral_ssids_b200_metis_order(ncomp, p3.data(), r3.data(), corder.data());
      if (rc != 0) return rc;
   }
   std::vector<int> inv(ncomp + 1);
   for (int i = 1; i <= ncomp; ++i) inv[corder[i - 1]] = i;        // iwork(order(i)) = i
   k = 1;
   for (int i = 1; i <= ncomp; ++i) {
      int j = new_to_old[inv[i]];
      order[j - 1] = k++;
      if (cp[j] > 0) { j = cp[j]; order[j - 1] = k++; }
   }
   return flag;
}

/* equilib_scale_sym -> inf_norm_equilib_sym (src/scaling.f90:480-521); defaults of
 * equilib_options: max_iterations = 10, tol = 1e-8 (:48-51). */
int spral_ssids_b200_equilib_scale_sym(int n, const int64_t* ptr, const int* row, const double* val,
      double* scaling, int max_iterations, double tol, int* iterations) {
   std::vector<double> maxentry(n);
   for (int i = 0; i < n; ++i) scaling[i] = 1.0;
   int itr = 1;
   for (; itr <= max_iterations; ++itr) {
      std::fill(maxentry.begin(), maxentry.end(), 0.0);
      for (int c = 0; c < n; ++c)
         for (int64_t e = ptr[c] - 1; e < ptr[c + 1] - 1; ++e) {
            const int r = row[e] - 1;
            const double v = std::fabs(scaling[r] * val[e] * scaling[c]);
            maxentry[r] = std::max(maxentry[r], v);
            maxentry[c] = std::max(maxentry[c], v);
         }
      double dev = 0.0;
      for (int i = 0; i < n; ++i) {
         if (maxentry[i] > 0) scaling[i] /= std::sqrt(maxentry[i]);
         dev = std::max(dev, std::fabs(1 - maxentry[i]));
      }
      if (dev < tol) break;
   }
   if (iterations) *iterations = itr - 1;
   return 0;
}

} /* extern "C" */
