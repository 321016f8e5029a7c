/* Host-side symbolic analysis: C++ restatement of the Fortran-only pieces of
 * the reference that produce the numeric hot path's inputs.
 *
 * The north star reuses the reference's CPU analyse unchanged; this image has
 * no Fortran compiler, so the routines are restated here one-for-one (1-based
 * index VALUES and tie-breaking preserved) so that sptr/sparent/rptr/rlist/
 * nptr/nlist/part/contrib_* are what the reference would hand to a subtree.
 * PARITY UNPINNED at this boundary: the Fortran cannot be executed here; the
 * outputs are instead checked by invariants (tests/test_analyse.py) and by
 * feeding the SAME arrays to the compiled reference CPU engine.
 *
 * Each function cites the reference routine it follows.
 */
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <vector>

#include "spral_ssids_b200.h"

/* METIS 5 (64-bit idx_t, static lib shipped in the CUDA toolkit; no metis.h) */
extern "C" {
int METIS_SetDefaultOptions(int64_t* options);
int METIS_NodeND(int64_t* nvtxs, int64_t* xadj, int64_t* adjncy, int64_t* vwgt,
                 int64_t* options, int64_t* perm, int64_t* iperm);
}

namespace {

typedef int64_t i64;
/* 1-based vector helper: v[1..n] valid. */
template <typename T> using vec = std::vector<T>;

/* ---- src/metis5_wrapper.F90:276-326 half_to_full_drop_diag ---- */
void half_to_full_drop_diag(int n, const i64* ptr, const int* row,
                            vec<i64>& ptr2, vec<i64>& row2) {
   /* ptr,row: 1-based values, C arrays (ptr[j-1]). */
   ptr2.assign(n + 2, 0);
   for (int j = 1; j <= n; ++j)
      for (i64 k = ptr[j - 1]; k <= ptr[j] - 1; ++k) {
         int i = row[k - 1];
         if (j != i) { ptr2[i]++; ptr2[j]++; }
      }
   for (int j = 2; j <= n; ++j) ptr2[j] = ptr2[j - 1] + ptr2[j];
   ptr2[n + 1] = ptr2[n] + 1;
   row2.assign(ptr2[n + 1] + 1, 0);
   for (int j = 1; j <= n; ++j)
      for (i64 k = ptr[j - 1]; k <= ptr[j] - 1; ++k) {
         int i = row[k - 1];
         if (j != i) {
            row2[ptr2[i]] = j;
            row2[ptr2[j]] = i;
            ptr2[i]--; ptr2[j]--;
         }
      }
   for (int j = 1; j <= n; ++j) ptr2[j]++;
}

/* ---- src/ssids/anal.F90:42-85 expand_pattern ---- */
void expand_pattern(int n, i64 nz, const i64* ptr, const int* row,
                    vec<i64>& aptr, vec<int>& arow) {
   aptr.assign(n + 2, 0);
   arow.assign(2 * nz + 1, 0);
   for (int j = 1; j <= n; ++j)
      for (i64 kk = ptr[j - 1]; kk <= ptr[j] - 1; ++kk) {
         int i = row[kk - 1];
         aptr[i]++;
         if (j == i) continue;
         aptr[j]++;
      }
   for (int j = 2; j <= n; ++j) aptr[j] = aptr[j - 1] + aptr[j];
   aptr[n + 1] = aptr[n] + 1;
   for (int j = 1; j <= n; ++j)
      for (i64 kk = ptr[j - 1]; kk <= ptr[j] - 1; ++kk) {
         int i = row[kk - 1];
         arow[aptr[i]] = j;
         aptr[i]--;
         if (j == i) continue;
         arow[aptr[j]] = i;
         aptr[j]--;
      }
   for (int j = 1; j <= n; ++j) aptr[j]++;
}

/* ---- src/core_analyse.f90:173-224 find_etree (Liu) ---- */
void find_etree(int n, const vec<i64>& ptr, const vec<int>& row,
                const vec<int>& perm, const vec<int>& invp, vec<int>& parent) {
   vec<int> vforest(n + 2, n + 1);
   parent.assign(n + 2, 0);
   for (int piv = 1; piv <= n; ++piv) {
      int rowidx = invp[piv];
      for (i64 i = ptr[rowidx]; i <= ptr[rowidx + 1] - 1; ++i) {
         int j = perm[row[i]];
         if (j >= piv) continue;
         int k = j;
         while (vforest[k] < piv) {
            int l = vforest[k];
            vforest[k] = piv;
            k = l;
         }
         if (vforest[k] == piv) continue;
         parent[k] = piv;
         vforest[k] = piv;
      }
      parent[piv] = n + 1;
   }
}

/* ---- src/core_analyse.f90:233-357 find_postorder ---- */
void find_postorder(int n, int& realn, const vec<i64>& ptr, vec<int>& perm,
                    vec<int>& invp, vec<int>& parent) {
   realn = n;
   vec<int> chead(n + 2, -1), cnext(n + 2, -1);
   for (int i = n; i >= 1; --i) {
      int j = parent[i];
      cnext[i] = chead[j];
      chead[j] = i;
   }
   vec<int> map(n + 2, 0), stack(n + 2, 0);
   int shead = 1;
   stack[shead] = n + 1;
   int id = n + 1;
   while (shead != 0) {
      int node = stack[shead];
      shead--;
      map[node] = id;
      id--;
      if (node == n + 1) {
         int i = chead[node];
         while (i != -1) {
            if (ptr[invp[i] + 1] - ptr[invp[i]] == 0) { i = cnext[i]; continue; }
            stack[++shead] = i;
            i = cnext[i];
         }
         i = chead[node];
         while (i != -1) {
            if (ptr[invp[i] + 1] - ptr[invp[i]] != 0) { i = cnext[i]; continue; }
            realn--;
            stack[++shead] = i;
            i = cnext[i];
         }
      } else {
         int i = chead[node];
         while (i != -1) {
            stack[++shead] = i;
            i = cnext[i];
         }
      }
   }
   for (int i = 1; i <= n; ++i) stack[i] = invp[i];
   for (int i = 1; i <= n; ++i) invp[map[i]] = stack[i];
   for (int i = 1; i <= n; ++i) perm[invp[i]] = i;
   for (int i = 1; i <= n; ++i) stack[i] = map[parent[i]];
   for (int i = 1; i <= n; ++i) parent[map[i]] = stack[i];
}

/* ---- src/core_analyse.f90:506-523 FIND ---- */
int vf_find(vec<int>& vforest, int u) {
   int prev = -1, current = u;
   while (vforest[current] != 0) {
      prev = current;
      current = vforest[current];
      if (vforest[current] != 0) vforest[prev] = vforest[current];
   }
   return current;
}

/* ---- src/core_analyse.f90:387-501 find_col_counts (Gilbert/Ng/Peyton) ---- */
void find_col_counts(int n, const vec<i64>& ptr, const vec<int>& row,
                     const vec<int>& perm, const vec<int>& invp,
                     const vec<int>& parent, vec<int>& cc) {
   vec<int> first(n + 2);
   cc.assign(n + 2, 0);
   for (int i = 1; i <= n + 1; ++i) first[i] = i;
   for (int i = 1; i <= n; ++i) {
      int par = parent[i];
      first[par] = std::min(first[i], first[par]);
      cc[i] = (first[i] == i) ? 1 : 0;
   }
   cc[n + 1] = n + 1;
   vec<int> vforest(n + 2, 0), last_p(n + 2, 0), last_nbr(n + 2, 0);
   for (int piv = 1; piv <= n; ++piv) {
      int col = invp[piv];
      for (i64 ii = ptr[col]; ii <= ptr[col + 1] - 1; ++ii) {
         int u = perm[row[ii]];
         if (u <= piv) continue;
         if (first[piv] > last_nbr[u]) {
            cc[piv] += 1;
            int pp = last_p[u];
            if (pp != 0) {
               int lca = vf_find(vforest, pp);
               cc[lca] -= 1;
            }
            last_p[u] = piv;
         }
         last_nbr[u] = piv;
      }
      int par = parent[piv];
      cc[par] = cc[par] + cc[piv] - 1;
      vforest[piv] = par;
   }
}

/* ---- src/core_analyse.f90:712-804 sort_by_val / sort_by_val_ms ----
 * Sorts idx[0..n) into DECREASING val(idx), stable in the reference's sense. */
const int minsz_ms = 16;
void sort_by_val(int n, int* idx, const vec<int>& val);
void sort_by_val_ms(int n, int* idx, const vec<int>& val) {
   if (n <= 1) return;
   if (n < minsz_ms) { sort_by_val(n, idx, val); return; }
   int mid = (n - 1) / 2 + 1;
   sort_by_val_ms(mid, idx, val);
   sort_by_val_ms(n - mid, idx + mid, val);
   std::vector<int> work(idx, idx + mid);
   /* 1-based transliteration: j over work, k over idx(mid+1:n) */
   int j = 1, k = mid + 1;
   int jj = work[j - 1], jj2 = val[jj];
   int kk = idx[k - 1], kk2 = val[kk];
   int i;
   for (i = 1; i <= n; ++i) {
      if (jj2 >= kk2) {
         idx[i - 1] = jj;
         j++;
         if (j > mid) break;
         jj = work[j - 1]; jj2 = val[jj];
      } else {
         idx[i - 1] = kk;
         k++;
         if (k > n) break;
         kk = idx[k - 1]; kk2 = val[kk];
      }
   }
   if (j <= mid)
      for (int t = 0; t <= mid - j; ++t) idx[i + t] = work[j - 1 + t];  /* idx(i+1:n) = work(j:mid) */
}
void sort_by_val(int n, int* idx, const vec<int>& val) {
   if (n >= minsz_ms) { sort_by_val_ms(n, idx, val); return; }
   int kor = n;
   for (int kdummy = 2; kdummy <= n; ++kdummy) {
      int ice_idx = idx[kor - 2];
      int ice_val = val[ice_idx];
      int k;
      for (k = kor; k <= n; ++k) {
         int ik_idx = idx[k - 1];
         int ik_val = val[ik_idx];
         if (ice_val >= ik_val) break;
         idx[k - 2] = ik_idx;
      }
      idx[k - 2] = ice_idx;
      kor--;
   }
}

/* ---- src/core_analyse.f90:536-710 find_supernodes (+do_merge :806-820,
 *      merge_nodes :827-857) ---- */
void find_supernodes(int n, int realn, const vec<int>& parent, const vec<int>& cc,
                     vec<int>& sperm, int& nnodes, vec<int>& sptr,
                     vec<int>& sparent, vec<int>& scc, int nemin) {
   const i64 HUGE = std::numeric_limits<i64>::max();
   vec<int> nelim(n + 2, 1), nvert(n + 2, 1), vhead(n + 2, -1), vnext(n + 2, -1),
       stack(n + 2, 0), height(n + 2, 1);
   std::vector<char> mark(n + 2, 0);
   vec<int> map(n + 2, 0), npar(n + 2, 0);
   vec<i64> ezero(n + 2, 0);
   int totalwt = n;
   ezero[n + 1] = HUGE;
   nelim[n + 1] = totalwt + 1 + nemin;

   vec<int> chead(n + 2, -1), cnext(n + 2, -1), child(n + 2, 0);
   for (int i = realn; i >= 1; --i) {
      int j = parent[i];
      cnext[i] = chead[j];
      chead[j] = i;
   }

   sperm.assign(n + 2, 0);
   sptr.assign(n + 2, 0);
   sparent.assign(n + 2, 0);
   scc.assign(n + 2, 0);

   int v = 1;
   nnodes = 0;
   for (int par = 1; par <= n + 1; ++par) {
      int nchild = 0;
      int node = chead[par];
      while (node != -1) {
         child[nchild++] = node;
         node = cnext[node];
      }
      sort_by_val(nchild, child.data(), cc);
      for (int j = 0; j < nchild; ++j) {
         node = child[j];
         bool merge;
         if (ezero[par] == HUGE) merge = false;
         else
            merge = ((cc[par] == cc[node] - 1) && (nelim[par] == 1)) ||
                    ((nelim[par] < nemin) && (nelim[node] < nemin));
         if (merge) {
            vnext[node] = vhead[par];
            vhead[par] = node;
            ezero[par] = ezero[par] + ezero[node] +
                         ((i64)cc[par] - 1 + nelim[par] - cc[node] + 1) * nelim[par];
            nelim[par] += nelim[node];
            nvert[par] += nvert[node];
            height[par] = std::max(height[par], height[node]);
            mark[node] = 0;
         } else {
            mark[node] = 1;
         }
      }
   }

   for (int node = 1; node <= realn; ++node) {
      if (!mark[node]) continue;
      nnodes++;
      sptr[nnodes] = v;
      npar[nnodes] = parent[node];
      scc[nnodes] = cc[node] + nelim[node] - 1;
      height[parent[node]] = std::max(height[parent[node]], height[node] + 1);
      v += nvert[node];
      int k = v;
      int shead = 1;
      stack[shead] = node;
      while (shead > 0) {
         int i = stack[shead];
         shead--;
         k--;
         sperm[i] = k;
         map[i] = nnodes;
         if (vnext[i] != -1) stack[++shead] = vnext[i];
         if (vhead[i] != -1) stack[++shead] = vhead[i];
      }
   }
   sptr[nnodes + 1] = v;
   map[n + 1] = nnodes + 1;
   npar[nnodes + 1] = n + 1;
   for (int i = realn + 1; i <= n; ++i) sperm[i] = i;
   for (int node = 1; node <= nnodes; ++node) sparent[node] = map[npar[node]];
}

/* ---- src/core_analyse.f90:1069-1098 apply_perm ---- */
void apply_perm(int n, const vec<int>& perm, vec<int>& order, vec<int>& invp,
                vec<int>& cc) {
   for (int i = 1; i <= n; ++i) order[i] = cc[i];
   for (int i = 1; i <= n; ++i) cc[perm[i]] = order[i];
   for (int i = 1; i <= n; ++i) order[i] = invp[i];
   for (int i = 1; i <= n; ++i) invp[perm[i]] = order[i];
   for (int i = 1; i <= n; ++i) order[invp[i]] = i;
}

/* ---- src/core_analyse.f90:911-998 find_row_lists ---- */
void find_row_lists(int n, const vec<i64>& ptr, const vec<int>& row,
                    const vec<int>& perm, const vec<int>& invp, int nnodes,
                    const vec<int>& sptr, const vec<int>& sparent,
                    const vec<int>& scc, vec<i64>& rptr, vec<int>& rlist) {
   vec<int> seen(n + 2, 0), chead(nnodes + 2, -1), cnext(nnodes + 2, -1);
   for (int node = nnodes; node >= 1; --node) {
      int i = sparent[node];
      cnext[node] = chead[i];
      chead[i] = node;
   }
   i64 total = 0;
   for (int node = 1; node <= nnodes; ++node) total += scc[node];
   rptr.assign(nnodes + 2, 0);
   rlist.assign(total + 1, 0);
   rptr[1] = 1;
   for (int node = 1; node <= nnodes; ++node) {
      rptr[node + 1] = rptr[node] + scc[node];
      i64 idx = rptr[node];
      for (int piv = sptr[node]; piv <= sptr[node + 1] - 1; ++piv) {
         seen[piv] = node;
         rlist[idx++] = piv;
      }
      int child = chead[node];
      while (child != -1) {
         for (i64 i = rptr[child]; i <= rptr[child + 1] - 1; ++i) {
            int j = rlist[i];
            if (j < sptr[node]) continue;
            if (seen[j] == node) continue;
            seen[j] = node;
            rlist[idx++] = j;
         }
         child = cnext[child];
      }
      for (int piv = sptr[node]; piv <= sptr[node + 1] - 1; ++piv) {
         int col = invp[piv];
         for (i64 i = ptr[col]; i <= ptr[col + 1] - 1; ++i) {
            int j = perm[row[i]];
            if (j < piv) continue;
            if (seen[j] == node) continue;
            seen[j] = node;
            rlist[idx++] = j;
         }
      }
   }
}

/* ---- src/core_analyse.f90:1007-1064 dbl_tr_sort ---- */
void dbl_tr_sort(int n, int nnodes, const vec<i64>& rptr, vec<int>& rlist) {
   vec<i64> ptr(n + 3, 0);
   for (int node = 1; node <= nnodes; ++node)
      for (i64 ii = rptr[node]; ii <= rptr[node + 1] - 1; ++ii) ptr[rlist[ii] + 2]++;
   ptr[1] = 1; ptr[2] = 1;
   for (int i = 1; i <= n; ++i) ptr[i + 2] = ptr[i + 1] + ptr[i + 2];
   i64 jj = ptr[n + 2] - 1;
   vec<int> col(jj + 1, 0);
   for (int node = 1; node <= nnodes; ++node)
      for (i64 ii = rptr[node]; ii <= rptr[node + 1] - 1; ++ii) {
         int j = rlist[ii];
         col[ptr[j + 1]] = node;
         ptr[j + 1]++;
      }
   vec<i64> nptr(nnodes + 2, 0);
   for (int node = 1; node <= nnodes; ++node) nptr[node] = rptr[node];
   for (int i = 1; i <= n; ++i)
      for (i64 k = ptr[i]; k <= ptr[i + 1] - 1; ++k) {
         int node = col[k];
         rlist[nptr[node]] = i;
         nptr[node]++;
      }
}

/* ---- src/ssids/anal.F90:1137-1239 build_map ---- */
void build_map(int n, const i64* ptr, const int* row, const vec<int>& perm,
               const vec<int>& invp, int nnodes, const vec<int>& sptr,
               const vec<i64>& rptr, const vec<int>& rlist, vec<i64>& nptr,
               vec<i64>& nlist /* 2*nz, pairs, 0-based storage */) {
   i64 nz = ptr[n] - 1;
   vec<int> map(n + 2, 0);
   vec<i64> ptr2(n + 4, 0);
   vec<int> row2(nz + 1, 0);
   vec<i64> origin(nz + 1, 0);
   for (int i = 1; i <= n; ++i)
      for (i64 jj = ptr[i - 1]; jj <= ptr[i] - 1; ++jj) {
         int k = row[jj - 1];
         if (k == i) continue;
         ptr2[k + 2]++;
      }
   ptr2[1] = 1; ptr2[2] = 1;
   for (int i = 1; i <= n; ++i) ptr2[i + 2] = ptr2[i + 2] + ptr2[i + 1];
   for (int i = 1; i <= n; ++i)
      for (i64 jj = ptr[i - 1]; jj <= ptr[i] - 1; ++jj) {
         int k = row[jj - 1];
         if (k == i) continue;
         row2[ptr2[k + 1]] = i;
         origin[ptr2[k + 1]] = jj;
         ptr2[k + 1]++;
      }
   nptr.assign(nnodes + 2, 0);
   nlist.assign(2 * nz, 0);
   i64 pp = 1;
   for (int node = 1; node <= nnodes; ++node) {
      int blkm = (int)(rptr[node + 1] - rptr[node]);
      nptr[node] = pp;
      for (i64 jj = rptr[node]; jj <= rptr[node + 1] - 1; ++jj)
         map[rlist[jj]] = (int)(jj - rptr[node] + 1);
      for (int j = sptr[node]; j <= sptr[node + 1] - 1; ++j) {
         int col = invp[j];
         for (i64 i = ptr2[col]; i <= ptr2[col + 1] - 1; ++i) {
            int k = std::abs(perm[row2[i]]);
            if (k < j) continue;
            nlist[2 * (pp - 1) + 1] = (i64)(j - sptr[node]) * blkm + map[k];
            nlist[2 * (pp - 1) + 0] = origin[i];
            pp++;
         }
      }
      for (int j = sptr[node]; j <= sptr[node + 1] - 1; ++j) {
         int col = invp[j];
         for (i64 ii = ptr[col - 1]; ii <= ptr[col] - 1; ++ii) {
            int k = std::abs(perm[row[ii - 1]]);
            if (k < j) continue;
            nlist[2 * (pp - 1) + 1] = (i64)(j - sptr[node]) * blkm + map[k];
            nlist[2 * (pp - 1) + 0] = ii;
            pp++;
         }
      }
   }
   nptr[nnodes + 1] = pp;
}

/* ---- src/ssids/anal.F90:210-230 compute_flops ---- */
i64 compute_flops(const vec<int>& sptr, const vec<i64>& rptr, int node) {
   i64 m = rptr[node + 1] - rptr[node];
   i64 n = sptr[node + 1] - sptr[node];
   i64 f = 0;
   for (i64 jj = m - n + 1; jj <= m; ++jj) f += jj * jj;
   return f;
}

/* ---- src/ssids/anal.F90:734-753 create_size_order ---- */
void create_size_order(int nparts, const vec<int>& part, const vec<i64>& flops,
                       vec<int>& size_order) {
   for (int i = 1; i <= nparts; ++i) {
      i64 iflops = flops[part[i + 1] - 1];
      /* NB: the reference compares against flops(part(j+1)-1), i.e. it indexes
       * parts by POSITION j rather than by size_order(j); kept as written. */
      int j;
      for (j = 1; j <= i - 1; ++j)
         if (iflops > flops[part[j + 1] - 1]) break;
      for (int t = i; t >= j + 1; --t) size_order[t] = size_order[t - 1];
      size_order[j] = i;
   }
}

struct Topology { int nregion; std::vector<int> ngpus; /* per region */ };

/* ---- src/ssids/anal.F90:508-621 calc_exec_alloc ----
 * gpu_only (extension, the reference declares but never reads options%gpu_only):
 * CPU regions are left out of the resource map so every child part lands on
 * a GPU regardless of min_gpu_work. */
float calc_exec_alloc(int nparts, const vec<int>& part, const vec<int>& size_order,
                      const std::vector<char>& is_child, const vec<i64>& flops,
                      const Topology& topo, i64 min_gpu_work, float gpu_perf_coeff,
                      bool gpu_only, vec<int>& exec_loc) {
   int nregion = topo.nregion, ngpu = 0, max_gpu = 0;
   for (int i = 0; i < nregion; ++i) {
      ngpu += topo.ngpus[i];
      max_gpu = std::max(max_gpu, topo.ngpus[i]);
   }
   std::vector<int> map;
   if (gpu_only) {
      for (int i = 1; i <= nregion; ++i)
         for (int p = 1; p <= topo.ngpus[i - 1]; ++p) map.push_back(p * nregion + i);
   } else if (gpu_perf_coeff > 1.0f) {
      for (int i = 1; i <= nregion; ++i)
         for (int p = 1; p <= topo.ngpus[i - 1]; ++p) map.push_back(p * nregion + i);
      for (int i = 1; i <= nregion; ++i) map.push_back(i);
   } else {
      for (int i = 1; i <= nregion; ++i) map.push_back(i);
      for (int i = 1; i <= nregion; ++i)
         for (int p = 1; p <= topo.ngpus[i - 1]; ++p) map.push_back(p * nregion + i);
   }
   int nmap = (int)map.size();
   int next = 1;
   for (int i = 1; i <= nparts; ++i) {
      int p = size_order[i];
      if (!is_child[p]) { exec_loc[p] = -1; continue; }
      i64 pflops = flops[part[p + 1] - 1];
      if (pflops < min_gpu_work && !gpu_only) {
         while (map[next - 1] > nregion) {
            next++;
            if (next > nmap) next = 1;
         }
      }
      exec_loc[p] = map[next - 1];
      next++;
      if (next > nmap) next = 1;
   }
   std::vector<float> load_balance(nregion * (1 + max_gpu) + 1, 0.0f);
   float total_balance = 0.0f;
   for (int p = 1; p <= nparts; ++p) {
      if (exec_loc[p] == -1) continue;
      i64 pflops = flops[part[p + 1] - 1];
      if (exec_loc[p] > nregion) {
         load_balance[exec_loc[p]] += (float)pflops / gpu_perf_coeff;
         total_balance += (float)pflops / gpu_perf_coeff;
      } else {
         load_balance[exec_loc[p]] += (float)pflops;
         total_balance += (float)pflops;
      }
   }
   float mx = 0.0f;
   for (size_t i = 1; i < load_balance.size(); ++i) mx = std::max(mx, load_balance[i]);
   int nres = gpu_only ? ngpu : (nregion + ngpu);
   return nres * mx / total_balance;
}

/* ---- src/ssids/anal.F90:645-722 split_tree ---- */
void split_tree(int& nparts, vec<int>& part, vec<int>& size_order,
                std::vector<char>& is_child, const vec<int>& sparent,
                const vec<i64>& flops, int ngpu, i64 min_gpu_work) {
   int to_split = 1;
   while (!is_child[size_order[to_split]]) to_split++;
   int to_split_pos = to_split;
   to_split = size_order[to_split];
   std::vector<int> children;
   int root = part[to_split + 1] - 1;
   for (int i = part[to_split]; i <= root - 1; ++i)
      if (sparent[i] == root) children.push_back(i);
   int nchild = (int)children.size();
   if (nchild == 0) return;
   int nbig = 0;
   /* NB: the reference loops i = to_split+1..nparts over size_order using the
    * PART index, not the position in size_order; kept as written. */
   (void)to_split_pos;
   for (int i = to_split + 1; i <= nparts; ++i) {
      int p = size_order[i];
      if (!is_child[p]) continue;
      int r = part[p + 1] - 1;
      if (flops[r] < min_gpu_work) break;
      nbig++;
   }
   if (nbig + 1 >= ngpu) {
      for (int i = 0; i < nchild; ++i)
         if (flops[children[i]] >= min_gpu_work) nbig++;
      if (nbig < ngpu) return;
   }
   /* shift later parts back (overlapping: copy from the end) */
   for (int t = nparts + 1; t >= to_split + 1; --t) part[t + nchild] = part[t];
   for (int t = nparts; t >= to_split + 1; --t) is_child[t + nchild] = is_child[t];
   for (int i = 1; i <= nchild; ++i) part[to_split + i] = children[i - 1] + 1;
   for (int t = to_split; t <= to_split + nchild - 1; ++t) is_child[t] = 1;
   is_child[to_split + nchild] = 0;
   nparts += nchild;
   create_size_order(nparts, part, flops, size_order);
}

} /* anon */

struct spral_ssids_b200_analysis {
   int n = 0, nnodes = 0, nparts = 0;
   std::vector<int> sptr, sparent, rlist, invp, part, exec_loc, contrib_ptr,
       contrib_idx, contrib_dest;
   std::vector<i64> rptr, nptr, nlist;
   i64 num_factor = 0, num_flops = 0;
   int maxfront = 0, maxsupernode = 0, maxdepth = 0;
   bool gpu_only = false;
};

/* contrib_ptr / contrib_idx / contrib_dest of a partition (the tail of find_subtree_partition, anal.F90:418-454):
 * part p sends its root contribution to slot contrib_idx(p) of the part that holds the parent of its last node.
 * part / exec_loc are 1-based work arrays (part[1..nparts+1], exec_loc[1..nparts]). */
static void store_partition(spral_ssids_b200_analysis& A, int nparts, const vec<int>& part, const vec<int>& exec_loc,
                            const vec<int>& sparent) {
   const int nnodes = A.nnodes;
   vec<int> contrib_ptr(nparts + 4, 0), contrib_idx(nparts + 1, 0), contrib_dest(nparts + 1, 0);
   for (int i = 1; i <= nparts - 1; ++i) {
      int jn = sparent[part[i + 1] - 1];
      if (jn > nnodes) continue;
      int kp = i + 1;
      while (jn >= part[kp + 1]) kp++;
      contrib_ptr[kp + 2]++;
   }
   contrib_ptr[1] = 1; contrib_ptr[2] = 1;
   for (int i = 1; i <= nparts; ++i) contrib_ptr[i + 2] = contrib_ptr[i + 1] + contrib_ptr[i + 2];
   for (int i = 1; i <= nparts - 1; ++i) {
      int jn = sparent[part[i + 1] - 1];
      if (jn > nnodes) { contrib_idx[i] = nparts + 1; continue; }
      int kp = i + 1;
      while (jn >= part[kp + 1]) kp++;
      contrib_idx[i] = contrib_ptr[kp + 1];
      contrib_dest[contrib_idx[i]] = jn;
      contrib_ptr[kp + 1]++;
   }
   contrib_idx[nparts] = nparts + 1;

   A.nparts = nparts;
   A.part.assign(part.begin() + 1, part.begin() + nparts + 2);
   A.exec_loc.assign(exec_loc.begin() + 1, exec_loc.begin() + nparts + 1);
   A.contrib_ptr.assign(contrib_ptr.begin() + 1, contrib_ptr.begin() + nparts + 4);
   A.contrib_idx.assign(contrib_idx.begin() + 1, contrib_idx.begin() + nparts + 1);
   A.contrib_dest.assign(contrib_dest.begin() + 1, contrib_dest.begin() + nparts + 1);
}

/* ---- src/ssids/anal.F90:289-464 find_subtree_partition ---- */
static void find_subtree_partition(spral_ssids_b200_analysis& A, const vec<int>& sptr,
                                   const vec<int>& sparent, const vec<i64>& rptr,
                                   const Topology& topo, i64 min_gpu_work,
                                   float max_load_inbalance, float gpu_perf_coeff,
                                   bool gpu_only) {
   int nnodes = A.nnodes;
   vec<i64> flops(nnodes + 2, 0);
   for (int node = 1; node <= nnodes; ++node) {
      flops[node] += compute_flops(sptr, rptr, node);
      int j = std::min(sparent[node], nnodes + 1);
      flops[j] += flops[node];
   }
   int cap = 2 * nnodes + 8;
   vec<int> part(cap, 0), size_order(cap, 0), exec_loc(cap, 0);
   std::vector<char> is_child(cap, 0);
   int nparts = 0;
   part[1] = 1;
   for (int i = 1; i <= nnodes; ++i)
      if (sparent[i] > nnodes) {
         nparts++;
         part[nparts + 1] = i + 1;
         is_child[nparts] = 1;
      }
   create_size_order(nparts, part, flops, size_order);
   int nregion = topo.nregion, ngpu = 0;
   for (int g : topo.ngpus) ngpu += g;
   int nres = gpu_only ? ngpu : nregion + ngpu;
   for (int i = 1; i <= 2 * nres; ++i) {
      float lb = calc_exec_alloc(nparts, part, size_order, is_child, flops, topo,
                                 min_gpu_work, gpu_perf_coeff, gpu_only, exec_loc);
      if (lb < max_load_inbalance) break;
      split_tree(nparts, part, size_order, is_child, sparent, flops, ngpu, min_gpu_work);
   }
   /* consolidate adjacent non-children */
   int j = 1;
   for (int i = 2; i <= nparts; ++i) {
      part[j + 1] = part[i];
      if (is_child[i] || is_child[j]) {
         j++;
         is_child[j] = is_child[i];
      }
   }
   part[j + 1] = part[nparts + 1];
   nparts = j;
   create_size_order(nparts, part, flops, size_order);
   calc_exec_alloc(nparts, part, size_order, is_child, flops, topo, min_gpu_work,
                   gpu_perf_coeff, gpu_only, exec_loc);
   /* merge adjacent subtrees on the same location with <=1 contribution to parent */
   j = 1;
   int k = sparent[part[j + 1] - 1];
   bool has_parent = (k <= nnodes);
   for (int i = 2; i <= nparts; ++i) {
      part[j + 1] = part[i];
      exec_loc[j + 1] = exec_loc[i];
      k = sparent[part[i + 1] - 1];
      if (exec_loc[i] != exec_loc[j] || (has_parent && k <= nnodes)) {
         j++;
         has_parent = false;
      }
      has_parent = has_parent || (k <= nnodes);
   }
   part[j + 1] = part[nparts + 1];
   nparts = j;

   store_partition(A, nparts, part, exec_loc, sparent);
}


extern "C" {

/* follows metis_order (src/metis5_wrapper.F90:102-182) */
int spral_ssids_b200_metis_order(int n, const int64_t* ptr, const int* row, int* order) {
   if (n < 1) return -1;
   if (n == 1) { order[0] = 1; return 0; }
   vec<i64> ptr2, row2;
   half_to_full_drop_diag(n, ptr, row, ptr2, row2);
   i64 nz2 = ptr2[n + 1] - 1;
   std::vector<i64> xadj(n + 1), adj(std::max<i64>(nz2, 1));
   for (int j = 1; j <= n + 1; ++j) xadj[j - 1] = ptr2[j] - 1;
   for (i64 k = 1; k <= nz2; ++k) adj[k - 1] = row2[k] - 1;
   i64 opts[40];
   METIS_SetDefaultOptions(opts);
   std::vector<i64> mperm(n), miperm(n);
   i64 nn = n;
   int rc = METIS_NodeND(&nn, xadj.data(), adj.data(), nullptr, opts, mperm.data(), miperm.data());
   if (rc != 1) return (rc == -3) ? -50 : -99;
   /* METIS perm -> SPRAL invp, METIS iperm -> SPRAL perm/order (:163,179-180) */
   for (int i = 0; i < n; ++i) order[i] = (int)miperm[i] + 1;
   return 0;
}

/* follows analyse_phase (src/ssids/anal.F90:945-1129) */
struct spral_ssids_b200_analysis* spral_ssids_b200_analyse(
      int n, const int64_t* ptr, const int* row, int* order_io, int nemin, int ngpu,
      int64_t min_gpu_work, float max_load_inbalance, float gpu_perf_coeff, int* flag) {
   auto* A = new spral_ssids_b200_analysis;
   *flag = 0;
   A->n = n;
   if (nemin < 1) nemin = 32; /* nemin_default, datatypes.f90 */
   if (n <= 0) {
      /* trivial matrix: ssids_analyse returns with akeep%nnodes = 0 (src/ssids/ssids.f90:212-218)
       * and factor / solve return immediately (:848-852, :1193) */
      A->n = 0;
      A->sptr.assign(1, 1); A->rptr.assign(1, 1); A->nptr.assign(1, 1); A->part.assign(1, 1);
      A->contrib_ptr.assign(3, 1);
      return A;
   }
   i64 nz = ptr[n] - 1;
   vec<i64> ptr2;
   vec<int> row2;
   expand_pattern(n, nz, ptr, row, ptr2, row2);

   /* basic_analyse (core_analyse.f90:38-156) */
   vec<int> perm(n + 2, 0), invp(n + 2, 0);
   for (int i = 1; i <= n; ++i) perm[i] = order_io[i - 1];
   for (int i = 1; i <= n; ++i) invp[perm[i]] = i;
   int realn = n;
   vec<int> parent, cc, tperm, sptr, sparent, scc;
   find_etree(n, ptr2, row2, perm, invp, parent);
   find_postorder(n, realn, ptr2, perm, invp, parent);
   if (n != realn) *flag = 6; /* SSIDS_WARNING_ANAL_SINGULAR */
   if (realn == 0) {
      /* no entries at all: every variable is structurally unused, there is nothing to factorise
       * (the numeric phases see nnodes = 0 like for n = 0; matrix_rank = 0) */
      A->sptr.assign(1, 1); A->rptr.assign(1, 1); A->nptr.assign(1, 1); A->part.assign(1, 1);
      A->contrib_ptr.assign(3, 1);
      A->invp.resize(n);
      for (int i = 0; i < n; ++i) A->invp[i] = i + 1;
      return A;
   }
   find_col_counts(n, ptr2, row2, perm, invp, parent, cc);
   int nnodes = 0;
   find_supernodes(n, realn, parent, cc, tperm, nnodes, sptr, sparent, scc, nemin);
   apply_perm(n, tperm, perm, invp, cc);
   vec<i64> rptr;
   vec<int> rlist;
   find_row_lists(n, ptr2, row2, perm, invp, nnodes, sptr, sparent, scc, rptr, rlist);
   /* calc_stats (core_analyse.f90:862-903) */
   i64 nfact = 0, nflops = 0;
   for (int node = 1; node <= nnodes; ++node) {
      i64 nel = sptr[node + 1] - sptr[node];
      i64 m = scc[node] - nel;
      nfact += nel * (nel + 1) / 2 + nel * m;
      for (i64 j = 1; j <= nel; ++j) nflops += (m + j) * (m + j);
   }
   dbl_tr_sort(n, nnodes, rptr, rlist);

   /* analyse_phase :1001-1009: invp = inverse of order; unused variables get 0 */
   for (int i = 1; i <= n; ++i) invp[perm[i]] = i;
   for (int j = sptr[nnodes + 1]; j <= n; ++j) perm[invp[j]] = 0;

   vec<i64> nptr, nlist;
   build_map(n, ptr, row, perm, invp, nnodes, sptr, rptr, rlist, nptr, nlist);

   A->nnodes = nnodes;
   A->num_factor = nfact;
   A->num_flops = nflops;

   bool gpu_only = ngpu < 0;           /* negative ngpu: |ngpu| GPUs, no CPU resource */
   if (ngpu == 0) gpu_only = false;
   Topology topo;
   topo.nregion = 1;
   topo.ngpus.assign(1, std::abs(ngpu));
   A->gpu_only = gpu_only;
   find_subtree_partition(*A, sptr, sparent, rptr, topo, min_gpu_work,
                          max_load_inbalance, gpu_perf_coeff, gpu_only);

   /* inform (anal.F90:1100-1116) */
   vec<int> level(nnodes + 2, 0);
   for (int i = nnodes; i >= 1; --i) {
      int blkn = sptr[i + 1] - sptr[i];
      int blkm = (int)(rptr[i + 1] - rptr[i]);
      level[i] = level[std::min(sparent[i], nnodes + 1)] + 1;
      A->maxfront = std::max(A->maxfront, blkm);
      A->maxsupernode = std::max(A->maxsupernode, blkn);
      A->maxdepth = std::max(A->maxdepth, level[i]);
   }

   /* export as 0-based C arrays holding 1-based values */
   A->sptr.assign(sptr.begin() + 1, sptr.begin() + nnodes + 2);
   A->sparent.assign(sparent.begin() + 1, sparent.begin() + nnodes + 1);
   A->rptr.assign(rptr.begin() + 1, rptr.begin() + nnodes + 2);
   A->rlist.assign(rlist.begin() + 1, rlist.begin() + rptr[nnodes + 1]);
   A->nptr.assign(nptr.begin() + 1, nptr.begin() + nnodes + 2);
   A->nlist.swap(nlist);
   A->invp.assign(invp.begin() + 1, invp.begin() + n + 1);
   for (int i = 1; i <= n; ++i) order_io[i - 1] = perm[i];
   return A;
}

void spral_ssids_b200_analysis_free(struct spral_ssids_b200_analysis* A) { delete A; }

/* Replaces the subtree partition of an analysis (one-process-per-GPU driver: proportional mapping of the top of the
 * tree, spral_b200/dist.py).  part[0..nparts] are 1-based first nodes (part[0] = 1, part[nparts] = nnodes + 1),
 * exec_loc[0..nparts-1] the reference's location codes.  A part must be connected with a single exit: only its last
 * node may have its parent outside.  Returns 0, or -1 when the partition is not valid (the analysis is unchanged). */
int spral_ssids_b200_analysis_set_partition(struct spral_ssids_b200_analysis* A, int nparts, const int* part_in,
      const int* exec_loc_in) {
   if (!A || nparts < 1 || part_in[0] != 1 || part_in[nparts] != A->nnodes + 1) return -1;
   const int nnodes = A->nnodes;
   vec<int> sparent(nnodes + 2, 0), part(nparts + 3, 0), exec_loc(nparts + 2, 0);
   for (int i = 1; i <= nnodes; ++i) sparent[i] = A->sparent[i - 1];
   for (int p = 1; p <= nparts; ++p) {
      part[p] = part_in[p - 1]; exec_loc[p] = exec_loc_in[p - 1];
      if (part_in[p] <= part_in[p - 1]) return -1;
      for (int i = part_in[p - 1]; i < part_in[p] - 1; ++i)              /* all but the last node stay inside */
         if (sparent[i] >= part_in[p] || sparent[i] < part_in[p - 1]) return -1;
   }
   part[nparts + 1] = part_in[nparts];
   store_partition(*A, nparts, part, exec_loc, sparent);
   return 0;
}

void spral_ssids_b200_analysis_get(const struct spral_ssids_b200_analysis* A,
                                   struct spral_ssids_b200_analysis_view* v) {
   v->n = A->n; v->nnodes = A->nnodes; v->nparts = A->nparts;
   v->sptr = A->sptr.data(); v->sparent = A->sparent.data();
   v->rptr = A->rptr.data(); v->rlist = A->rlist.data();
   v->nptr = A->nptr.data(); v->nlist = A->nlist.data();
   v->invp = A->invp.data(); v->part = A->part.data();
   v->exec_loc = A->exec_loc.data(); v->contrib_ptr = A->contrib_ptr.data();
   v->contrib_idx = A->contrib_idx.data(); v->contrib_dest = A->contrib_dest.data();
   v->num_factor = A->num_factor; v->num_flops = A->num_flops;
   v->maxfront = A->maxfront; v->maxsupernode = A->maxsupernode; v->maxdepth = A->maxdepth;
}

} /* extern "C" */
