/* The factor/solve subset of SPRAL's C interface on top of the B200 engine.
 *
 * Semantics follow interfaces/C/ssids.f90 (:135-229 analyse, :505-590 factor,
 * :727-772 solve, :600-720 enquire/alter/free) and src/ssids/ssids.f90
 * (analyse :148-389, factor :767-1109, solve :1140-1250): array_base handling,
 * data checking (clean_cscl_oop of matrix_util.f90, restated minimally), flag
 * codes (src/ssids/datatypes.f90:25-59), inform fill-in (anal.F90:1100-1116),
 * x permutation and scaling (fkeep.F90:252-266,300-315).
 *
 * One process drives all parts of the subtree partition one after the other on
 * device 0 (ngpu = 1 gives a single part); the one-process-per-GPU scheduler
 * lives in spral_b200/dist.py.
 */
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>
#include "spral_ssids_b200.h"
#include "spral_ssids_compat.h"

namespace {

enum { OK = 0, E_CALL_SEQUENCE = -1, E_A_N_OOR = -2, E_A_PTR = -3, E_A_ALL_OOR = -4, E_SINGULAR = -5, E_ORDER = -8, E_VAL = -9,
       E_X_SIZE = -10, E_JOB_OOR = -11, E_NOT_LLT = -13, E_NOT_LDLT = -14, E_ALLOCATION = -50,
       E_NO_SAVED_SCALING = -15, E_UNIMPLEMENTED = -98,
       W_IDX_OOR = 1, W_DUP_IDX = 2, W_DUP_AND_OOR = 3, W_MISSING_DIAGONAL = 4, W_MISS_DIAG_OORDUP = 5,
       W_ANAL_SINGULAR = 6, W_FACT_SINGULAR = 7, W_MATCH_ORD_NO_SCALE = 8 };

struct Akeep {
   int n = 0;
   bool check = false;
   std::vector<int64_t> ptr;            // cleaned lower-triangular CSC, 1-based
   std::vector<int> row;
   std::vector<int64_t> map_ptr, map;   // cleaned entry k sums user entries map[map_ptr[k]..map_ptr[k+1])
   std::vector<double> mo_scaling;      // scaling found by the matching-based ordering (options%ordering = 2)
   bool analyse_only = false;           // built with SPRAL_B200_ANALYSE_ONLY: cannot be factorised
   spral_ssids_b200_analysis* an = nullptr;
   spral_ssids_b200_analysis_view v{};
   std::vector<void*> symbolic;         // one per part
   spral_ssids_inform inform{};         // analyse-time values (the flag of a failed analyse included)
   ~Akeep() {
      for (void* s : symbolic) if (s) spral_ssids_gpu_destroy_symbolic_subtree(s);
      if (an) spral_ssids_b200_analysis_free(an);
   }
};

struct Fkeep {
   bool posdef = false;
   std::vector<void*> numeric;
   std::vector<double> scaling;         // in pivot order (fkeep%scaling(i) = scale(invp(i)))
   ~Fkeep() { for (void* p : numeric) if (p) spral_ssids_gpu_destroy_num_subtree_dbl(posdef, p); }
};

spral_ssids_b200_options engine_options(const spral_ssids_options* o) {
   spral_ssids_b200_options e;
   std::memset(&e, 0, sizeof(e));
   e.print_level = o->print_level; e.action = o->action; e.small = o->small; e.u = o->u;
   e.multiplier = 1.1; e.small_subtree_threshold = o->small_subtree_threshold;
   e.cpu_block_size = o->cpu_block_size;
   e.pivot_method = std::min(3, std::max(1, o->pivot_method));
   e.failed_pivot_method = 1;
   return e;
}

/* Data checking of ssids_analyse(check=true): out-of-range entries are dropped,
 * upper-triangular entries moved to the lower triangle, duplicates summed
 * (clean_cscl_oop, src/matrix_util.f90; counts reported as the reference does). */
int clean_columns(int n, int64_t ne, int64_t oor, std::vector<std::vector<std::pair<int, int64_t>>>& cols, Akeep& A,
      spral_ssids_inform* inf);

int clean_matrix(int n, int base, const int64_t* ptr, const int* row, Akeep& A, spral_ssids_inform* inf) {
   if (ptr[0] < base) return E_A_PTR;          /* ptr(1) < 1 (matrix_util.f90, clean_cscl_oop) */
   int64_t ne = ptr[n] - base;
   std::vector<std::vector<std::pair<int, int64_t>>> cols(n);   // per cleaned column: (row, source index)
   int64_t oor = 0;
   for (int j = 0; j < n; ++j) {
      if (ptr[j + 1] < ptr[j]) return E_A_PTR;
      for (int64_t k = ptr[j] - base; k < ptr[j + 1] - base; ++k) {
         int i = row[k] - base;
         if (i < 0 || i >= n) { oor++; continue; }
         if (i >= j) cols[j].push_back({i, k}); else cols[i].push_back({j, k});
      }
   }
   return clean_columns(n, ne, oor, cols, A, inf);
}

/* Coordinate input of ssids_analyse_coord (src/ssids/ssids.f90:392-700; clean_coord, matrix_util.f90):
 * entries with an index out of range are dropped, upper-triangular ones mirrored, duplicates summed. */
int clean_coord(int n, int base, int64_t ne, const int* row, const int* col, Akeep& A, spral_ssids_inform* inf) {
   std::vector<std::vector<std::pair<int, int64_t>>> cols(n);
   int64_t oor = 0;
   for (int64_t k = 0; k < ne; ++k) {
      int i = row[k] - base, j = col[k] - base;
      if (i < 0 || i >= n || j < 0 || j >= n) { oor++; continue; }
      if (i >= j) cols[j].push_back({i, k}); else cols[i].push_back({j, k});
   }
   return clean_columns(n, ne, oor, cols, A, inf);
}

int clean_columns(int n, int64_t ne, int64_t oor, std::vector<std::vector<std::pair<int, int64_t>>>& cols, Akeep& A,
      spral_ssids_inform* inf) {
   if (ne > 0 && oor == ne) return E_A_ALL_OOR;
   int64_t dup = 0;
   int missing = 0;
   A.ptr.assign(n + 1, 1);
   A.row.clear(); A.map_ptr.assign(1, 0); A.map.clear();
   for (int j = 0; j < n; ++j) {
      auto& c = cols[j];
      std::stable_sort(c.begin(), c.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
      bool diag = false;
      for (size_t q = 0; q < c.size(); ++q) {
         if (q > 0 && c[q].first == c[q - 1].first) { dup++; A.map.push_back(c[q].second); A.map_ptr.back() = (int64_t)A.map.size(); continue; }
         A.row.push_back(c[q].first + 1);
         A.map.push_back(c[q].second);
         A.map_ptr.push_back((int64_t)A.map.size());
         if (c[q].first == j) diag = true;
      }
      if (!diag) missing++;
      A.ptr[j + 1] = (int64_t)A.row.size() + 1;
   }
   inf->matrix_dup = (int)dup; inf->matrix_outrange = (int)oor; inf->matrix_missing_diag = missing;
   int flag = OK;
   if (oor && dup) flag = W_DUP_AND_OOR; else if (oor) flag = W_IDX_OOR; else if (dup) flag = W_DUP_IDX;
   if (missing) flag = (flag == OK) ? W_MISSING_DIAGONAL : W_MISS_DIAG_OORDUP;
   return flag;
}

void analyse_common(bool check, int n, int* order, const int64_t* ptr, const int* row, const double* val,
      void** akeep, const spral_ssids_options* opt, spral_ssids_inform* inf,
      int64_t coord_ne = -1, const int* coord_col = nullptr) {
   std::memset(inf, 0, sizeof(*inf));
   if (*akeep) { delete static_cast<Akeep*>(*akeep); *akeep = nullptr; }
   if (n < 0) { inf->flag = E_A_N_OOR; return; }
   const int base = opt->array_base ? 1 : 0;
   Akeep* A = new (std::nothrow) Akeep;
   if (!A) { inf->flag = E_ALLOCATION; return; }
   *akeep = A;
   A->n = n; A->check = check;
   /* every error exit leaves the flag in the akeep as well: a later ssids_factor then answers
    * SSIDS_ERROR_CALL_SEQUENCE (ssids.f90:806-811) instead of walking a half-built analysis */
   auto fail = [&](int flag) { inf->flag = flag; A->inform = *inf; };
   int wflag = OK;
   if (coord_ne >= 0) {                 /* coordinate input: always cleaned; `row` / coord_col are the triplets */
      A->check = true; check = true;
      wflag = clean_coord(n, base, coord_ne, row, coord_col, *A, inf);
      if (wflag < 0) { fail(wflag); return; }
   } else if (check) {
      wflag = clean_matrix(n, base, ptr, row, *A, inf);
      if (wflag < 0) { fail(wflag); return; }
   } else {
      A->ptr.resize(n + 1);
      for (int j = 0; j <= n; ++j) A->ptr[j] = ptr[j] - base + 1;
      int64_t ne = A->ptr[n] - 1;
      A->row.resize(ne);
      for (int64_t k = 0; k < ne; ++k) A->row[k] = row[k] - base + 1;
   }
   if (n == 0) { inf->flag = wflag; A->inform = *inf; return; }
   /* ordering (ssids.f90:287-353) */
   std::vector<int> ord(n);
   if (opt->ordering == 0) {
      if (!order) { fail(E_ORDER); return; }
      std::vector<char> seen(n + 1, 0);
      for (int i = 0; i < n; ++i) {
         int p = order[i] - base + 1;
         if (p < 1 || p > n || seen[p]) { fail(E_ORDER); return; }
         seen[p] = 1; ord[i] = p;
      }
   } else if (opt->ordering == 1) {
      int rc = spral_ssids_b200_metis_order(n, A->ptr.data(), A->row.data(), ord.data());
      if (rc != 0) { fail(rc); return; }
   } else if (opt->ordering == 2) {
      /* matching-based ordering (ssids.f90:312-353): needs the values; the scaling is kept for
       * options%scaling = 3 at factor time */
      if (!val) { fail(E_VAL); return; }
      std::vector<double> cleaned;
      const double* aval = val;
      if (check) {
         cleaned.assign(A->row.size(), 0.0);
         for (size_t k = 0; k < cleaned.size(); ++k)
            for (int64_t q = A->map_ptr[k]; q < A->map_ptr[k + 1]; ++q) cleaned[k] += val[A->map[q]];
         aval = cleaned.data();
      }
      A->mo_scaling.resize(n);
      int rc = spral_ssids_b200_match_order_metis(n, A->ptr.data(), A->row.data(), aval, ord.data(), A->mo_scaling.data());
      if (rc < 0) { fail((rc == -1 || rc == -50) ? E_ALLOCATION : rc); return; }
      if (rc == 1) wflag = W_ANAL_SINGULAR;
   } else { fail(E_ORDER); return; }
   int aflag = 0;
   A->an = spral_ssids_b200_analyse(n, A->ptr.data(), A->row.data(), ord.data(), opt->nemin, -1,
                                    0, opt->max_load_inbalance, opt->gpu_perf_coeff, &aflag);
   if (!A->an) { fail(aflag < 0 ? aflag : E_ALLOCATION); return; }
   spral_ssids_b200_analysis_get(A->an, &A->v);
   if (order) for (int i = 0; i < n; ++i) order[i] = ord[i] - 1 + base;
   spral_ssids_b200_options eo = engine_options(opt);
   /* SPRAL_B200_ANALYSE_ONLY (test hook): no symbolic subtrees, i.e. no device is touched; such an akeep
    * can be inspected and freed but not factorised (the CPU test-suite checks data cleaning, orderings
    * and the analyse-time inform this way) */
   const bool analyse_only = getenv("SPRAL_B200_ANALYSE_ONLY") != nullptr;
   A->analyse_only = analyse_only;
   for (int p = 0; p < A->v.nparts && !analyse_only; ++p) {
      int lo = A->v.contrib_ptr[p] - 1, hi = A->v.contrib_ptr[p + 1] - 1;
      void* s = spral_ssids_gpu_create_symbolic_subtree(0, n, A->v.part[p], A->v.part[p + 1], A->v.sptr,
            A->v.sparent, A->v.rptr, A->v.rlist, A->v.nptr, A->v.nlist, hi - lo, A->v.contrib_dest + lo, &eo);
      if (!s) { fail(-51); return; }
      A->symbolic.push_back(s);
   }
   /* inform (anal.F90:1100-1116) */
   inf->matrix_rank = A->v.sptr[A->v.nnodes] - 1;
   inf->num_sup = A->v.nnodes;
   inf->maxdepth = A->v.maxdepth; inf->maxfront = A->v.maxfront; inf->maxsupernode = A->v.maxsupernode;
   inf->num_factor = A->v.num_factor; inf->num_flops = A->v.num_flops;
   inf->flag = (aflag == W_ANAL_SINGULAR) ? W_ANAL_SINGULAR : wflag;
   A->inform = *inf;
}

/* x2(i) = x(invp(i)) [* scaling(i)]; the sweeps; back (fkeep.F90:252-315) */
void solve_common(int job, int nrhs, double* x, int ldx, Akeep* A, Fkeep* F, spral_ssids_inform* inf) {
   const int n = A->n;
   if (n == 0 || nrhs == 0) return;
   std::vector<double> x2((size_t)n * nrhs);
   const int* invp = A->v.invp;
   const bool sc = !F->scaling.empty();
   for (int r = 0; r < nrhs; ++r)
      for (int i = 0; i < n; ++i) {
         double v = x[(size_t)r * ldx + invp[i] - 1];
         if (sc && (job == 0 || job == 1)) v *= F->scaling[i];
         x2[(size_t)r * n + i] = v;
      }
   int rc = 0;
   const int np = (int)F->numeric.size();
   if (job == 0 || job == 1)
      for (int p = 0; p < np && rc == 0; ++p) rc = spral_ssids_gpu_subtree_solve_fwd_dbl(F->posdef, F->numeric[p], nrhs, x2.data(), n);
   if (job == 2)
      for (int p = 0; p < np && rc == 0; ++p) rc = spral_ssids_gpu_subtree_solve_diag_dbl(F->posdef, F->numeric[p], nrhs, x2.data(), n);
   if (job == 3)
      for (int p = np - 1; p >= 0 && rc == 0; --p) rc = spral_ssids_gpu_subtree_solve_bwd_dbl(F->posdef, F->numeric[p], nrhs, x2.data(), n);
   if (job == 0 || job == 4)
      for (int p = np - 1; p >= 0 && rc == 0; --p) rc = spral_ssids_gpu_subtree_solve_diag_bwd_dbl(F->posdef, F->numeric[p], nrhs, x2.data(), n);
   if (rc < 0) { inf->flag = rc; return; }
   for (int r = 0; r < nrhs; ++r)
      for (int i = 0; i < n; ++i) {
         double v = x2[(size_t)r * n + i];
         if (sc && (job == 0 || job == 3 || job == 4)) v *= F->scaling[i];
         x[(size_t)r * ldx + invp[i] - 1] = v;
      }
}

} // namespace

extern "C" {

/* defaults of ssids_options (src/ssids/datatypes.f90:188-284) */
void spral_ssids_default_options(struct spral_ssids_options* o) {
   std::memset(o, 0, sizeof(*o));
   o->array_base = 0; o->print_level = 0; o->unit_diagnostics = 6; o->unit_error = 6; o->unit_warning = 6;
   o->ordering = 1; o->nemin = 32; o->ignore_numa = true; o->use_gpu = true;
   o->min_gpu_work = 5000000000LL; o->max_load_inbalance = 1.2f; o->gpu_perf_coeff = 1.0f;
   o->scaling = 0; o->small_subtree_threshold = 4000000; o->cpu_block_size = 256;
   o->action = true; o->pivot_method = 2; o->small = 1e-20; o->u = 0.01;
}

void spral_ssids_analyse(bool check, int n, int* order, const int64_t* ptr, const int* row,
      const double* val, void** akeep, const struct spral_ssids_options* options,
      struct spral_ssids_inform* inform) {
   analyse_common(check, n, order, ptr, row, val, akeep, options, inform);
}

void spral_ssids_analyse_topology(bool check, int n, int* order, const int64_t* ptr, const int* row,
      const double* val, void** akeep, const struct spral_ssids_options* options,
      struct spral_ssids_inform* inform, int, const struct spral_numa_region*) {
   analyse_common(check, n, order, ptr, row, val, akeep, options, inform);
}

void spral_ssids_analyse_ptr32(bool check, int n, int* order, const int* ptr, const int* row,
      const double* val, void** akeep, const struct spral_ssids_options* options,
      struct spral_ssids_inform* inform) {
   std::vector<int64_t> p64(ptr, ptr + (n >= 0 ? n + 1 : 0));
   analyse_common(check, n, order, p64.data(), row, val, akeep, options, inform);
}

void spral_ssids_analyse_coord(int n, int* order, int64_t ne, const int* row, const int* col, const double* val,
      void** akeep, const struct spral_ssids_options* options, struct spral_ssids_inform* inform) {
   if (ne < 0) { std::memset(inform, 0, sizeof(*inform)); inform->flag = E_A_ALL_OOR; return; }
   analyse_common(true, n, order, nullptr, row, val, akeep, options, inform, ne, col);
}

void spral_ssids_factor(bool posdef, const int64_t*, const int*, const double* val, double* scale,
      void* akeep, void** fkeep, const struct spral_ssids_options* options,
      struct spral_ssids_inform* inform) {
   Akeep* A = static_cast<Akeep*>(akeep);
   if (!A) { inform->flag = E_CALL_SEQUENCE; return; }
   *inform = A->inform;
   if (A->inform.flag < 0 || A->analyse_only) { inform->flag = E_CALL_SEQUENCE; return; }
   if (A->n > 0 && (!A->an || (int)A->symbolic.size() != A->v.nparts)) { inform->flag = E_CALL_SEQUENCE; return; }
   /* options%scaling (ssids.f90:899-1028): <= 0 user vector, 1 Hungarian (MC64), 4.. equilibration;
    * 2 auction, 3 the scaling of the matching-based ordering (options%ordering = 2) */
   if (options->scaling == 3 && A->mo_scaling.empty()) { inform->flag = E_NO_SAVED_SCALING; return; }
   if (*fkeep) { delete static_cast<Fkeep*>(*fkeep); *fkeep = nullptr; }
   Fkeep* F = new (std::nothrow) Fkeep;
   if (!F) { inform->flag = E_ALLOCATION; return; }
   *fkeep = F;
   F->posdef = posdef;
   const int n = A->n;
   if (n == 0) return;
   /* apply_conversion_map (ssids.f90:862-867) */
   std::vector<double> cleaned;
   const double* aval = val;
   if (A->check) {
      cleaned.assign(A->row.size(), 0.0);
      for (size_t k = 0; k < cleaned.size(); ++k)
         for (int64_t q = A->map_ptr[k]; q < A->map_ptr[k + 1]; ++q) cleaned[k] += val[A->map[q]];
      aval = cleaned.data();
   }
   if (options->scaling <= 0) {
      if (scale) {                     /* user scaling: fkeep%scaling(i) = scale(invp(i)) (ssids.f90:921-926) */
         F->scaling.resize(n);
         for (int i = 0; i < n; ++i) F->scaling[i] = scale[A->v.invp[i] - 1];
      }
   } else {
      std::vector<double> sc(n);
      if (options->scaling == 3) {     /* the scaling saved by the matching-based ordering (:989-997) */
         sc = A->mo_scaling;
      } else if (options->scaling == 2) {   /* auction_scale_sym with default auction_options (:961-987) */
         spral_ssids_b200_auction_scale_sym(n, A->ptr.data(), A->row.data(), aval, sc.data(), nullptr, nullptr,
                                            nullptr, nullptr);
      } else if (options->scaling == 1) {     /* hungarian_scale_sym, scale_if_singular = options%action (:927-959) */
         int matched = 0;
         int hf = spral_ssids_b200_hungarian_scale_sym(n, A->ptr.data(), A->row.data(), aval, sc.data(), nullptr,
                                                       options->action ? 1 : 0, &matched);
         if (hf == -2) { inform->flag = E_SINGULAR; return; }
      } else {                         /* equilib_scale_sym with default equilib_options (:998-1028) */
         spral_ssids_b200_equilib_scale_sym(n, A->ptr.data(), A->row.data(), aval, sc.data(), 10, 1e-8, nullptr);
      }
      F->scaling.resize(n);
      for (int i = 0; i < n; ++i) F->scaling[i] = sc[A->v.invp[i] - 1];
      if (scale) for (int i = 0; i < n; ++i) scale[i] = sc[i];
   }
   spral_ssids_b200_options eo = engine_options(options);
   const int np = A->v.nparts;
   std::vector<spral_ssids_b200_contrib> slots(np + 1);
   std::memset((void*)slots.data(), 0, slots.size() * sizeof(slots[0]));
   inform->num_delay = 0; inform->num_neg = 0; inform->num_two = 0;
   inform->num_factor = 0; inform->num_flops = 0; inform->maxfront = 0; inform->maxsupernode = 0;
   for (int p = 0; p < np; ++p) {
      int lo = A->v.contrib_ptr[p] - 1, hi = A->v.contrib_ptr[p + 1] - 1;
      std::vector<void*> cc(std::max(1, hi - lo));
      for (int i = lo; i < hi; ++i) cc[i - lo] = &slots[i];
      spral_ssids_b200_stats st;
      void* ns = spral_ssids_gpu_create_num_subtree_dbl(posdef, A->symbolic[p], aval,
            F->scaling.empty() ? nullptr : F->scaling.data(), cc.data(), &eo, &st);
      F->numeric.push_back(ns);
      for (int i = lo; i < hi; ++i)      /* the consumer releases what it was handed (contrib_free.f90) */
         if (slots[i].owner_ptr) spral_ssids_gpu_subtree_free_contrib_dbl(posdef, slots[i].owner_ptr);
      /* cpu_copy_stats_out (src/ssids/cpu/cpu_iface.f90:74-94) */
      if (st.flag < 0) { inform->flag = st.flag; inform->cuda_error = st.cuda_error; return; }
      inform->flag = std::max(inform->flag, st.flag);
      inform->num_delay += st.num_delay; inform->num_neg += st.num_neg; inform->num_two += st.num_two;
      inform->num_factor += st.num_factor; inform->num_flops += st.num_flops;
      inform->maxfront = std::max(inform->maxfront, st.maxfront);
      inform->maxsupernode = std::max(inform->maxsupernode, st.maxsupernode);
      inform->matrix_rank -= st.num_zero;
      int idx = A->v.contrib_idx[p] - 1;
      if (idx < np) spral_ssids_b200_contrib_fill(&slots[idx], posdef, ns, true);
   }
   /* rank deficient: always WARNING_FACT_SINGULAR, whatever the analyse phase warned (ssids.f90:1050-1058;
    * action = false never gets here: the engine returned -5) */
   if (inform->matrix_rank < n && inform->flag >= 0) inform->flag = W_FACT_SINGULAR;
   /* matching-based ordering without its scaling (ssids.f90:1060-1063) */
   else if (inform->flag >= 0 && !A->mo_scaling.empty() && options->scaling != 3) inform->flag = W_MATCH_ORD_NO_SCALE;
}

void spral_ssids_solve(int job, int nrhs, double* x, int ldx, void* akeep, void* fkeep,
      const struct spral_ssids_options*, struct spral_ssids_inform* inform) {
   Akeep* A = static_cast<Akeep*>(akeep);
   Fkeep* F = static_cast<Fkeep*>(fkeep);
   inform->flag = OK;
   if (!A || !F) { inform->flag = E_CALL_SEQUENCE; return; }
   if (job < 0 || job > 4) { inform->flag = E_JOB_OOR; return; }
   if (F->posdef && (job == 2 || job == 4)) { inform->flag = E_JOB_OOR; return; }   /* ssids.f90:1187-1191 */
   if (ldx < A->n || nrhs < 1) { inform->flag = E_X_SIZE; return; }
   solve_common(job, nrhs, x, ldx, A, F, inform);
}

void spral_ssids_solve1(int job, double* x1, void* akeep, void* fkeep,
      const struct spral_ssids_options* options, struct spral_ssids_inform* inform) {
   Akeep* A = static_cast<Akeep*>(akeep);
   spral_ssids_solve(job, 1, x1, A ? std::max(1, A->n) : 1, akeep, fkeep, options, inform);
}

int spral_ssids_free_akeep(void** akeep) {
   if (akeep && *akeep) { delete static_cast<Akeep*>(*akeep); *akeep = nullptr; }
   return 0;
}
int spral_ssids_free_fkeep(void** fkeep) {
   if (fkeep && *fkeep) { delete static_cast<Fkeep*>(*fkeep); *fkeep = nullptr; }
   return 0;
}
int spral_ssids_free(void** akeep, void** fkeep) {      /* fkeep first: it refers to akeep (ssids.f90:1429) */
   spral_ssids_free_fkeep(fkeep);
   return spral_ssids_free_akeep(akeep);
}

void spral_ssids_enquire_posdef(const void* akeep, const void* fkeep, const struct spral_ssids_options*,
      struct spral_ssids_inform* inform, double* d) {
   const Akeep* A = static_cast<const Akeep*>(akeep);
   const Fkeep* F = static_cast<const Fkeep*>(fkeep);
   inform->flag = OK;
   if (!A || !F) { inform->flag = E_CALL_SEQUENCE; return; }
   if (!F->posdef) { inform->flag = E_NOT_LLT; return; }
   for (size_t p = 0; p < F->numeric.size(); ++p) {
      spral_ssids_gpu_subtree_enquire_dbl(true, F->numeric[p], nullptr, d);
      d += A->v.sptr[A->v.part[p + 1] - 1] - A->v.sptr[A->v.part[p] - 1];
   }
}

void spral_ssids_enquire_indef(const void* akeep, const void* fkeep, const struct spral_ssids_options*,
      struct spral_ssids_inform* inform, int* piv_order, double* d) {
   const Akeep* A = static_cast<const Akeep*>(akeep);
   const Fkeep* F = static_cast<const Fkeep*>(fkeep);
   inform->flag = OK;
   if (!A || !F) { inform->flag = E_CALL_SEQUENCE; return; }
   if (F->posdef) { inform->flag = E_NOT_LDLT; return; }
   const int n = A->n;
   /* The subtrees report in pivot order -- position in the part's pivot sequence, 0-based, negative for a variable of
    * a 2x2 pivot (NumericSubtree::enquire, src/ssids/cpu/NumericSubtree.hxx:424-470) -- and d in elimination order.
    * The user gets piv_order per ORIGINAL variable, 1-based (ssids.f90:1288-1350).  Two things the reference's CPU
    * path gets wrong are done properly here: (i) position 0 cannot carry a sign, so the first variable of a LEADING
    * 2x2 pivot would come back positive -- the 2x2 marker is taken from d instead (off-diagonal entry non-zero);
    * (ii) with several parts every part's positions start at 0 and the reference reads part 1 only
    * (fkeep.F90:386-405, "FIXME") -- here the positions and d are concatenated part by part. */
   std::vector<int> po(n, 0), tmp(n);
   std::vector<double> dd(2 * (size_t)n, 0.0), dpart(2 * (size_t)n);
   const int UNSET = -2147483647 - 1;
   int off = 0;                                              // pivots reported by the parts before this one
   for (size_t p = 0; p < F->numeric.size(); ++p) {
      std::fill(tmp.begin(), tmp.end(), UNSET);
      std::fill(dpart.begin(), dpart.end(), 0.0);
      spral_ssids_gpu_subtree_enquire_dbl(false, F->numeric[p], tmp.data(), dpart.data());
      int cnt = 0;
      for (int i = 0; i < n; ++i) {
         if (tmp[i] == UNSET) continue;
         const int pos = tmp[i] < 0 ? -tmp[i] : tmp[i];       // 0-based position inside the part
         const bool two = tmp[i] < 0 || dpart[2 * (size_t)pos + 1] != 0.0;      // first variable of a 2x2: d(2, pos) != 0
         po[i] = two ? -(off + pos + 1) : (off + pos + 1);
         ++cnt;
      }
      for (int k = 0; k < cnt && off + k < n; ++k) { dd[2 * (size_t)(off + k)] = dpart[2 * (size_t)k]; dd[2 * (size_t)(off + k) + 1] = dpart[2 * (size_t)k + 1]; }
      off += cnt;
   }
   if (piv_order) for (int i = 0; i < n; ++i) piv_order[A->v.invp[i] - 1] = po[i];
   if (d) std::copy(dd.begin(), dd.end(), d);
}

void spral_ssids_alter(const double* d, const void* akeep, void* fkeep, const struct spral_ssids_options*,
      struct spral_ssids_inform* inform) {
   Fkeep* F = static_cast<Fkeep*>(fkeep);
   inform->flag = OK;
   if (!akeep || !F) { inform->flag = E_CALL_SEQUENCE; return; }
   if (F->posdef) { inform->flag = E_NOT_LDLT; return; }
   /* part by part, each taking the d entries of the pivots it eliminated (alter_cpu, fkeep.F90:420-437; the
    * offsets follow the parts' eliminated counts, as enquire_indef reports them) */
   const Akeep* A = static_cast<const Akeep*>(akeep);
   const int n = A->n;
   std::vector<int> tmp(n);
   const int UNSET = -2147483647 - 1;
   size_t off = 0;
   for (size_t p = 0; p < F->numeric.size(); ++p) {
      std::fill(tmp.begin(), tmp.end(), UNSET);
      spral_ssids_gpu_subtree_enquire_dbl(false, F->numeric[p], tmp.data(), nullptr);
      size_t cnt = 0;
      for (int i = 0; i < n; ++i) if (tmp[i] != UNSET) ++cnt;
      spral_ssids_gpu_subtree_alter_dbl(false, F->numeric[p], d + 2 * off);
      off += cnt;
   }
}

/* 32-bit column pointers: the factor phase takes the pattern from akeep, ptr / row are not read
 * (interfaces/C/ssids.f90: they may be NULL when check = true). */
void spral_ssids_factor_ptr32(bool posdef, const int*, const int* row, const double* val, double* scale,
      void* akeep, void** fkeep, const struct spral_ssids_options* options, struct spral_ssids_inform* inform) {
   spral_ssids_factor(posdef, nullptr, row, val, scale, akeep, fkeep, options, inform);
}

} /* extern "C" */
