/* Panel kernels of the B200 SSIDS numeric engine (sm_100a).
 *
 * What they replace in the reference (ralna/spral):
 *   k_scatter_a        cu_load_nodes[_sc]       src/ssids/gpu/kernels/assemble.cu:41-98
 *   k_assemble         assemble<>               src/ssids/gpu/kernels/assemble.cu:172-232
 *   k_delays           add_delays               src/ssids/gpu/kernels/assemble.cu:248-274
 *   k_diag/k_apply/... cu_multiblock_ldlt/chol + setup + reorder family
 *                      src/ssids/gpu/kernels/dense_factor.cu:228-334,1092-1122,1350-1373
 *                      src/ssids/gpu/kernels/reorder.cu
 *   k_finalize         cu_collect_stats         src/ssids/gpu/kernels/dense_factor.cu:1380-1430
 * and whose numerical semantics follow the reference CPU kernels (the oracle):
 *   in-block pivoting  block_ldlt               src/ssids/cpu/kernels/block_ldlt.hxx:289-413
 *   a-posteriori test  check_threshold          src/ssids/cpu/kernels/ldlt_app.cxx:303-321
 *   statistics         NumericSubtree ctor      src/ssids/cpu/NumericSubtree.hxx:248-280
 *
 * Pivoting state machine (all decisions on the device, see engine.h:Front).
 * A front's n fully-summed columns are processed in outer panels of PW columns;
 * each panel in inner steps of BS columns:
 *   k_diag    advance state; factor the BS x BS diagonal block with full
 *             pivoting inside the block (1x1 / 2x2), keep L11, L11*D, D^-1, lperm
 *   k_apply   rows below the block: solve against L11^T, scale by D^-1, write
 *             the candidate L in place (originals are backed up), record the
 *             first column that violates |l_ij| <= 1/u
 *   k_commit  accept the columns before the first failure; restore the failed
 *             ones from the backup; permute the already-factored rows
 *   update    (gemm_dmma.cu) trailing update restricted to the panel
 *   k_swap    move the failed columns to the end of the panel
 * and at the end of a panel the outer DMMA update of everything right of the
 * panel, then k_swap(outer) moves the panel's failed columns to the end of the
 * candidate range.  Failed columns are retried in later passes; what is left
 * is delayed to the parent.
 */
#include <algorithm>
#include <cstdlib>
#include "engine.h"
#include "device_utils.cuh"
#include "pivot_state.h"
#include "diag_warp.cuh"
#include "panel_v2.h"

namespace b200 {
#ifndef COUNT_LAUNCH
#define COUNT_LAUNCH() (void)g_launches.fetch_add(1, std::memory_order_relaxed)
#endif

/* ------------------------------------------------------------------------ */
/* A scatter (init of a front)                                               */
/* ------------------------------------------------------------------------ */

constexpr int SCATTER_CHUNK = 1024;

/* lcol(dest) = aval(src) [* scaling]; follows add_a_block
 * (src/ssids/cpu/kernels/assemble.hxx:50-86): dest is relative to the
 * undelayed m0 x n0 node, rows >= n0 shift down by ndin. */
__global__ void __launch_bounds__(256)
k_scatter_a(Front* fronts, const int2* work, const int64_t* __restrict__ nlist,
      const int64_t* __restrict__ nptr, const int* __restrict__ node_of_front,
      const double* __restrict__ aval, const double* __restrict__ scaling) {
   int2 w = work[blockIdx.x];
   const Front* f = &fronts[w.x];
   int node = node_of_front[w.x];
   int64_t beg = nptr[node] + (int64_t)w.y * SCATTER_CHUNK;
   int64_t end = min(beg + SCATTER_CHUNK, nptr[node + 1]);
   int m0 = f->m0, n0 = f->n0, ndin = f->ndin, ldl = f->ldl;
   double* L = f->L;
   const int* rows = f->rows;
   if (w.y == 0)   /* perm of the node's own columns (assemble.hxx:198-200) */
      for (int i = threadIdx.x; i < n0; i += blockDim.x) f->perm[i] = rows[i];
   for (int64_t i = beg + threadIdx.x; i < end; i += blockDim.x) {
      int64_t src = nlist[2 * i] - 1;
      int64_t dest = nlist[2 * i + 1] - 1;
      int c = (int)(dest / m0);
      int r = (int)(dest % m0);
      double v = aval[src];
      if (scaling) v = scaling[rows[r] - 1] * v * scaling[rows[c] - 1];
      if (r >= n0) r += ndin;
      L[r + (size_t)c * ldl] = v;
   }
}

void launch_scatter_a(Front* fronts, const int2* work, int nwork, const int64_t* nlist,
      const int64_t* nptr, const int* node_of_front, const double* aval,
      const double* scaling, cudaStream_t s) {
   if (nwork == 0) return;
   k_scatter_a<<<nwork, 256, 0, s>>>(fronts, work, nlist, nptr, node_of_front, aval, scaling); COUNT_LAUNCH();
}

/* ------------------------------------------------------------------------ */
/* Extend-add                                                                */
/* ------------------------------------------------------------------------ */

constexpr int ASM_COLS = 8;   // child columns per CTA

/* dest(map(i), map(j)) += C(i,j), i >= j.  Child columns whose image is one
 * of the parent's own columns go to the parent's L (to_contrib = false, run
 * before the parent is factorised), the others to the parent's contribution
 * block (to_contrib = true, run after the parent's Schur complement was
 * formed).  Follows assemble_expected / assemble_expected_contrib
 * (src/ssids/cpu/kernels/assemble.hxx:88-139).  Children of one parent are
 * processed in separate launches (one launch per child rank) so that the
 * summation order is fixed; ATOMIC is used only for high-degree parents. */
template <bool ATOMIC>
__global__ void __launch_bounds__(256)
k_assemble(Front* fronts, const AsmSrc* srcs, const int2* work, int to_contrib) {
   int2 w = work[blockIdx.x];
   const AsmSrc* s = &srcs[w.x];
   if (!s->C) return;
   const Front* p = &fronts[s->parent];
   int cm = s->cm, ldc = s->ldc, npassl = s->npassl;
   int j0 = w.y * ASM_COLS, j1 = min(j0 + ASM_COLS, cm);
   if (to_contrib) { j0 = max(j0, npassl); } else { j1 = min(j1, npassl); }
   if (j0 >= j1) return;
   const int* __restrict__ map = s->map;
   const double* __restrict__ C = s->C;
   int n0 = p->n0, ndin = p->ndin;
   double* dest;
   size_t ldd;
   int roff;
   if (to_contrib) { dest = p->C; ldd = p->ldc; roff = -n0; }
   else            { dest = p->L; ldd = p->ldl; roff = 0; }
   /* column images of this CTA's child columns */
   __shared__ int s_col[ASM_COLS];
   if (threadIdx.x < ASM_COLS) s_col[threadIdx.x] = (j0 + (int)threadIdx.x < j1) ? map[j0 + threadIdx.x] - 1 + roff : 0;
   __syncthreads();
   const int nj = j1 - j0;
   for (int i = j0 + threadIdx.x; i < cm; i += blockDim.x) {
      int pi = map[i] - 1;
      int r = to_contrib ? pi - n0 : (pi < n0 ? pi : pi + ndin);
      /* all loads of the row first, then the stores: a load issued after a
       * possibly-aliasing store would wait for it */
      double v[ASM_COLS], old[ASM_COLS];
      #pragma unroll
      for (int q = 0; q < ASM_COLS; ++q) {
         bool on = (q < nj) && (j0 + q <= i);
         v[q] = on ? C[i + (size_t)(j0 + q) * ldc] : 0.0;
         if (!ATOMIC) old[q] = on ? dest[(size_t)r + (size_t)s_col[q] * ldd] : 0.0;
      }
      #pragma unroll
      for (int q = 0; q < ASM_COLS; ++q) {
         bool on = (q < nj) && (j0 + q <= i);
         if (on) {
            size_t d = (size_t)r + (size_t)s_col[q] * ldd;
            if (ATOMIC) atomicAdd(&dest[d], v[q]);
            else dest[d] = old[q] + v[q];
         }
      }
   }
}

void launch_assemble(Front* fronts, const AsmSrc* srcs, const int2* work, int nwork,
      bool to_contrib, bool use_atomics, cudaStream_t s) {
   if (nwork == 0) return;
   if (use_atomics) k_assemble<true><<<nwork, 256, 0, s>>>(fronts, srcs, work, to_contrib);
   else k_assemble<false><<<nwork, 256, 0, s>>>(fronts, srcs, work, to_contrib); COUNT_LAUNCH();
}
int assemble_cols_per_cta() { return ASM_COLS; }
int scatter_chunk() { return SCATTER_CHUNK; }

/* Delayed columns of a child become extra fully-summed columns of the parent,
 * placed after its own n0 columns.  Follows assemble_pre
 * (src/ssids/cpu/kernels/assemble.hxx:244-263): the ndelay x ndelay lower
 * square goes to the diagonal, the child's contribution rows either below the
 * new column or, when they are own columns of the parent, transposed into the
 * new ROW.  work[i] = (src, delayed column). */
__global__ void __launch_bounds__(128)
k_delays(Front* fronts, const AsmSrc* srcs, const int2* work) {
   int2 w = work[blockIdx.x];
   const AsmSrc* s = &srcs[w.x];
   const Front* p = &fronts[s->parent];
   int j = w.y;
   int nd = s->ndelay, cm = s->cm, ldd = s->lddelay;
   int pc = s->delay_col + j;
   size_t ldl = p->ldl;
   double* L = p->L;
   const double* __restrict__ src = s->dval + (size_t)j * ldd;
   if (threadIdx.x == 0) p->perm[pc] = s->dperm[j];
   for (int i = j + threadIdx.x; i < nd; i += blockDim.x)
      L[(size_t)(s->delay_col + i) + pc * ldl] = src[i];
   int n0 = p->n0, ndin = p->ndin;
   const int* __restrict__ map = s->map;
   for (int k = threadIdx.x; k < cm; k += blockDim.x) {
      int pi = map[k] - 1;
      double v = src[nd + k];
      if (pi < n0) L[pc + (size_t)pi * ldl] = v;
      else         L[(size_t)(pi + ndin) + pc * ldl] = v;
   }
}

void launch_delays(Front* fronts, const AsmSrc* srcs, const int2* work, int nwork, cudaStream_t s) {
   if (nwork == 0) return;
   k_delays<<<nwork, 128, 0, s>>>(fronts, srcs, work); COUNT_LAUNCH();
}

/* ------------------------------------------------------------------------ */
/* State machine                                                             */
/* ------------------------------------------------------------------------ */

/* advance_state(), account_segment(), snapshot_state(): pivot_state.h (shared with the host model
 * tests/c/pivot_state_emu.cpp) */

/* ------------------------------------------------------------------------ */
/* Diagonal block                                                            */
/* ------------------------------------------------------------------------ */

/* ONE warp per front (diag_warp.cuh): no block-wide barrier inside the chain of pivots, the decision taken
 * redundantly by every lane, every load of a rank-1 / rank-2 update issued before its stores.  (Round 1 used a
 * 1024-thread CTA, thread per entry, two block barriers per pivot: 51 us per block; a 4-warp variant measured the
 * same.  This kernel: 25 us per block on a B200, tools/micro/bench_diag.cu.) */
template <bool POSDEF>
__global__ void __launch_bounds__(32)
k_diag_w(Front* fronts, const int* __restrict__ flist, int new_panel, FactorParams prm) {
   Front* f = &fronts[flist[blockIdx.x]];
   __shared__ double S[BS * DW_LD];
   __shared__ double dinv[2 * BS];
   __shared__ int lperm[BS];
   const int lane = threadIdx.x;
   int go = 0;
   if (lane == 0) {
      advance_state(f, new_panel != 0);
      if (!f->finished && f->done < f->pend) {
         f->bs = min(BS, f->pend - f->done);
         f->first_fail = f->bs;
         f->step_valid = 1;
         go = 1;
      } else f->bs = 0;
   }
   go = __shfl_sync(0xffffffffu, go, 0);
   if (!go) return;
   __syncwarp();
   const int bs = f->bs, done = f->done, ldl = f->ldl;
   double* Ld = f->L + (size_t)done * ldl + done;   // the diagonal block
   BlockWS* ws = f->ws;
   /* lane = row: coalesced column segments */
   #pragma unroll 8
   for (int c = 0; c < BS; ++c)
      S[lane * DW_LD + c] = (lane < bs && c < bs && lane >= c) ? Ld[lane + (size_t)c * ldl] : 0.0;
   __syncwarp();
   if (POSDEF) {
      const int rc = diag_warp_chol(S, dinv, bs);
      if (rc != DW_OK) {
         if (lane == 0) { f->flag = rc; f->finished = 1; f->step_valid = 0; f->nelim = f->done; }
         return;
      }
      #pragma unroll 8
      for (int c = 0; c < BS; ++c) {
         const bool in = lane < bs && c < bs && lane >= c;
         const double l = in ? S[lane * DW_LD + c] : 0.0;
         if (in) Ld[lane + (size_t)c * ldl] = l;
         ws->l11[lane + c * BS] = l;
      }
      ws->dinv[lane] = (lane < bs) ? dinv[lane] : 0.0;
      return;
   }
   /* keep the unfactorised block (full symmetric) for k_commit */
   #pragma unroll 8
   for (int c = 0; c < BS; ++c) ws->a0[lane + c * BS] = (lane >= c) ? S[lane * DW_LD + c] : S[c * DW_LD + lane];
   __syncwarp();
   int zfrom = BS;
   const int rc = diag_warp_ldlt(S, dinv, lperm, bs, prm.small, prm.action, CUDART_INF, zfrom);
   if (rc != DW_OK) {
      if (lane == 0) { f->flag = rc; f->finished = 1; f->step_valid = 0; f->nelim = f->done; }
      return;
   }
   /* publish L11 (unit lower), L11*D, D^-1 and the local permutation */
   #pragma unroll 8
   for (int c = 0; c < BS; ++c) {
      double l = 0.0, y = 0.0;
      if (lane < bs && c < bs) {
         if (lane > c) { l = S[lane * DW_LD + c]; y = S[c * DW_LD + lane]; }
         else if (lane == c) l = 1.0;
      }
      ws->l11[lane + c * BS] = l;
      ws->ld11[lane + c * BS] = y;
   }
   ws->dinv[2 * lane] = (lane < bs) ? dinv[2 * lane] : 0.0;
   ws->dinv[2 * lane + 1] = (lane < bs) ? dinv[2 * lane + 1] : 0.0;
   ws->lperm[lane] = lperm[lane];
   if (lane == 0) ws->zfrom = zfrom;
}

void launch_diag(Front* fronts, const int* flist, int count, bool posdef, bool new_panel,
      const FactorParams& prm, cudaStream_t s) {
   if (count == 0) return;
   if (posdef) k_diag_w<true><<<count, 32, 0, s>>>(fronts, flist, new_panel, prm);
   else k_diag_w<false><<<count, 32, 0, s>>>(fronts, flist, new_panel, prm);
   COUNT_LAUNCH();
}

/* ------------------------------------------------------------------------ */
/* Speculative panel segments (panel_v2.h)                                   */
/* ------------------------------------------------------------------------ */

constexpr int CH_NT = 256;           // threads of the chain kernel
constexpr int CLD = CW + 4;          // column stride of the segment in shared memory: = 4 mod 16 doubles, so the 8 x 4
                                     // DMMA fragment loads are bank-conflict free along rows AND along the mirrored L*D
constexpr int TNT = 256;             // threads of the tiles kernel
constexpr int PBLD = PV_BS + 4;      // stride of the small k-major operand tiles (same rule)

struct ChainSm {
   double S[CW * CLD];               // S[c * CLD + r]: lower triangle A / L; (L D)(r, c) mirrored to S[r * CLD + c]
   double B[PV_BS * DW_LD];          // the 32 x 32 block being factorised (diag_warp.cuh storage)
   double X[PV_BS * DW_LD];          // inverse of its lower factor, X(r, c) = X[r * DW_LD + c]
   double dinv[2 * PV_BS];
   double c0[PV_BS], c1[PV_BS], c2[PV_BS];
   int lperm[PV_BS];
   int status;                       // 1: the block cannot be taken (failed / zero pivot, not positive definite)
   int fail;                         // a row of the diagonal block failed the a-posteriori test
};

/* One CTA per front: opens the panel / accounts the previous segment, then factorises the CW x CW diagonal
 * block at `done` in shared memory.  Nothing of the front is modified; the factors go to f->sws. */
template <bool POSDEF>
__global__ void __launch_bounds__(CH_NT)
k_panel_chain(Front* fronts, const int* __restrict__ flist, int new_panel, FactorParams prm) {
   extern __shared__ __align__(16) unsigned char pv_smem[];
   ChainSm& sh = *reinterpret_cast<ChainSm*>(pv_smem);
   Front* f = &fronts[flist[blockIdx.x]];
   __shared__ int s_go;
   if (threadIdx.x == 0) {
      advance_state(f, new_panel != 0);
      int go = segment_may_start(f) ? 1 : 0;
      if (go) { f->seg_valid = 1; f->seg_ok = 0; f->seg_fail = 0; }
      s_go = go;
   }
   __syncthreads();
   if (!s_go) return;
   const int p = f->done;
   const size_t ldl = (size_t)f->ldl;
   SegWS* out = f->sws;
   const double* Lseg = f->L + p + (size_t)p * ldl;
   const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
   double* S = sh.S;

   for (int c0 = warp; c0 < CW; c0 += 4 * (CH_NT / 32)) {      // four columns (16 independent loads) per thread and round
      double v[4][CW / 32];
      #pragma unroll
      for (int q = 0; q < 4; ++q)
         #pragma unroll
         for (int i = 0; i < CW / 32; ++i) {
            const int c = c0 + q * (CH_NT / 32), r = lane + 32 * i;
            v[q][i] = (r >= c) ? Lseg[r + (size_t)c * ldl] : 0.0;
         }
      #pragma unroll
      for (int q = 0; q < 4; ++q)
         #pragma unroll
         for (int i = 0; i < CW / 32; ++i) S[(c0 + q * (CH_NT / 32)) * CLD + lane + 32 * i] = v[q][i];
   }
   if (tid == 0) { sh.fail = 0; sh.status = 0; }
   __syncthreads();
   const double lim = 1.0 / prm.u;
   int ok = 1;
   for (int jb = 0; jb < CW; jb += PV_BS) {
      for (int e = tid; e < PV_BS * PV_BS; e += CH_NT) {
         const int r = e & 31, c = e >> 5;
         sh.B[r * DW_LD + c] = (r >= c) ? S[(jb + c) * CLD + jb + r] : 0.0;
      }
      __syncthreads();
      if (warp == 0) {
         int zfrom = PV_BS, rc;
         if (POSDEF) rc = diag_warp_chol(sh.B, sh.dinv, PV_BS);
         else rc = diag_warp_ldlt(sh.B, sh.dinv, sh.lperm, PV_BS, prm.small, 1, CUDART_INF, zfrom);
         /* not positive definite / zero pivots: the step-by-step path reports it */
         if (rc != DW_OK || zfrom < PV_BS) { if (lane == 0) sh.status = 1; }
         else {
            double v0, v1 = 0.0, v2 = 0.0;
            if (POSDEF) { v0 = sh.dinv[lane]; sh.lperm[lane] = lane; }
            else pv_dinv_coeffs(sh.dinv, lane, CUDART_INF, v0, v1, v2);
            sh.c0[lane] = v0; sh.c1[lane] = v1; sh.c2[lane] = v2;
         }
      }
      __syncthreads();
      if (sh.status) { ok = 0; break; }

      if (warp == 1) {
         /* X = L_jj^-1 (lower): lane c owns column c, kept in registers (no store / load round trip through shared
          * memory between rows): x_r = -(sum_{c <= k < r} L(r, k) x_k) / l_rr */
         const int c = lane;
         double x[PV_BS];
         #pragma unroll
         for (int r = 0; r < PV_BS; ++r) {
            double s0 = 0.0, s1 = 0.0;
            #pragma unroll
            for (int k = 0; k < r; ++k) {
               const double t = sh.B[r * DW_LD + k] * x[k];      // x[k] == 0 for k < c
               if (k & 1) s1 += t; else s0 += t;
            }
            double v = -(s0 + s1);
            if (POSDEF) v *= sh.c0[r];
            x[r] = (r < c) ? 0.0 : (r == c ? (POSDEF ? sh.c0[c] : 1.0) : v);
         }
         #pragma unroll
         for (int r = 0; r < PV_BS; ++r) sh.X[r * DW_LD + c] = x[r];
      } else if (warp >= 2 && warp <= 4) {
         /* rows of the diagonal block below the 32 x 32 block: Y = A21(:, lperm) L11^-T, W = Y D^-1,
          * a-posteriori test |w| <= 1/u (ldlt_app.cxx:303-321) */
         const int t = jb + PV_BS + (tid - 64);
         if (t < CW) {
            double y[PV_BS], wv[PV_BS];
            #pragma unroll
            for (int j = 0; j < PV_BS; ++j) y[j] = S[(jb + sh.lperm[j]) * CLD + t];
            int bad = 0;
            if (POSDEF) {                                    /* l_tj = (a_tj - sum_k l_tk l_jk) / l_jj */
               #pragma unroll
               for (int j = 0; j < PV_BS; ++j) {
                  double s = y[j];
                  #pragma unroll
                  for (int k = 0; k < j; ++k) s -= y[k] * sh.B[j * DW_LD + k];
                  y[j] = s * sh.c0[j];
                  wv[j] = y[j];
               }
            } else {
               #pragma unroll
               for (int j = 0; j < PV_BS; ++j) {
                  double s = y[j];
                  #pragma unroll
                  for (int k = 0; k < j; ++k) s -= y[k] * sh.B[j * DW_LD + k];
                  y[j] = s;
               }
               #pragma unroll
               for (int j = 0; j < PV_BS; ++j) {
                  double w = sh.c0[j] * y[j];
                  if (j + 1 < PV_BS) w += sh.c1[j] * y[(j + 1) % PV_BS];
                  if (j > 0) w += sh.c2[j] * y[(j + PV_BS - 1) % PV_BS];
                  wv[j] = w;
                  if (!(fabs(w) <= lim)) bad = 1;
               }
            }
            #pragma unroll
            for (int j = 0; j < PV_BS; ++j) {
               S[(jb + j) * CLD + t] = wv[j];          // L(t, jb + j)
               S[t * CLD + jb + j] = y[j];             // (L D)(t, jb + j), mirrored (== L for Cholesky)
            }
            if (bad) sh.fail = 1;
         }
      } else if (warp == 5) {
         const int i = lane;                            // a row of the 32 x 32 block itself
         double lv[PV_BS], dv[PV_BS];
         #pragma unroll
         for (int c = 0; c < PV_BS; ++c) { lv[c] = sh.B[i * DW_LD + c]; dv[c] = POSDEF ? lv[c] : sh.B[c * DW_LD + i]; }
         #pragma unroll
         for (int c = 0; c < PV_BS; ++c) {
            if (c < i) {
               S[(jb + c) * CLD + jb + i] = lv[c];
               S[(jb + i) * CLD + jb + c] = dv[c];
            } else if (c == i) S[(jb + c) * CLD + jb + i] = POSDEF ? lv[c] : 1.0;
         }
      } else if (!POSDEF) {
         /* earlier columns of the segment (warps 0, 6, 7): the block's permutation is a row permutation of L */
         const int t0 = (warp == 0) ? lane : (warp - 5) * 32 + lane;
         if (t0 < jb) {
            double v[PV_BS];
            #pragma unroll
            for (int i = 0; i < PV_BS; ++i) v[i] = S[t0 * CLD + jb + sh.lperm[i]];
            #pragma unroll
            for (int i = 0; i < PV_BS; ++i) S[t0 * CLD + jb + i] = v[i];
         }
      }
      if (warp == 0) {
         if (POSDEF) out->dinv[jb + lane] = sh.dinv[lane];      // 1 / l_jj, one per column
         else { out->dinv[2 * jb + lane] = sh.dinv[lane]; out->dinv[2 * jb + 32 + lane] = sh.dinv[32 + lane]; }
         out->lperm[jb + lane] = sh.lperm[lane];
      }
      __syncthreads();
      if (sh.fail) { ok = 0; break; }
      /* the inverse block for the tiles kernel (column-major, ld = 32) */
      for (int e = tid; e < PV_BS * PV_BS; e += CH_NT) {
         const int r = e & 31, c = e >> 5;
         out->invl[jb / PV_BS][e] = sh.X[r * DW_LD + c];
      }
      /* trailing update of the rest of the diagonal block on the tensor cores, 32 x 32 tiles, one per warp:
       * A(t, c) -= sum_k L(t, jb + k) (L D)(c, jb + k), t >= c >= jb + 32 */
      const int nb = (CW - jb - PV_BS) / PV_BS;
      for (int tile = warp; tile < nb * (nb + 1) / 2; tile += CH_NT / 32) {
         int bi = 0, rem = tile;
         while (rem > bi) { rem -= bi + 1; ++bi; }       // tile -> (bi >= bj)
         const int bj = rem;
         const int rb = jb + PV_BS + bi * PV_BS, cb = jb + PV_BS + bj * PV_BS;
         double acc[4][4][2];
         #pragma unroll
         for (int j = 0; j < 4; ++j)
            #pragma unroll
            for (int i = 0; i < 4; ++i) { acc[j][i][0] = 0.0; acc[j][i][1] = 0.0; }
         #pragma unroll
         for (int kk = 0; kk < PV_BS; kk += 4) {
            double af[4], bf[4];
            #pragma unroll
            for (int i = 0; i < 4; ++i) af[i] = S[(jb + kk + (lane & 3)) * CLD + rb + i * 8 + (lane >> 2)];
            #pragma unroll
            for (int j = 0; j < 4; ++j) bf[j] = S[(cb + j * 8 + (lane >> 2)) * CLD + jb + kk + (lane & 3)];
            #pragma unroll
            for (int j = 0; j < 4; ++j)
               #pragma unroll
               for (int i = 0; i < 4; ++i) pv_dmma(acc[j][i][0], acc[j][i][1], bf[j], af[i]);
         }
         #pragma unroll
         for (int j = 0; j < 4; ++j) {
            const int c = cb + j * 8 + (lane >> 2);
            #pragma unroll
            for (int i = 0; i < 4; ++i) {
               const int r = rb + i * 8 + 2 * (lane & 3);
               if (r >= c) S[c * CLD + r] -= acc[j][i][0];
               if (r + 1 >= c) S[c * CLD + r + 1] -= acc[j][i][1];
            }
         }
      }
      __syncthreads();
   }
   if (ok) {
      for (int c = warp; c < CW; c += CH_NT / 32) {
         #pragma unroll
         for (int i = 0; i < CW / 32; ++i) {
            const int t = lane + 32 * i;
            out->l11[t + (size_t)c * CW] = (t > c) ? S[c * CLD + t] : (t == c ? (POSDEF ? S[c * CLD + c] : 1.0) : 0.0);
            out->ld11[t + (size_t)c * CW] = (t > c) ? S[t * CLD + c] : 0.0;
         }
      }
   }
   if (tid == 0) { out->ok = ok; f->seg_ok = ok; }
}

struct TileSm {
   double T[CW * CLD];               // the tile, column-major: T[c * CLD + r], r < RT
   double Ys[PV_BS * CLD];           // Y = A21 P L11^-T of the current block column, same layout
   double Bs[(CW - PV_BS) * PBLD];   // Bs[k * PBLD + n] = (L11 D)(jb + n, k), k < jb
   double Xs[PV_BS * PBLD];          // Xs[k * PBLD + c] = X(c, k), X the inverse of the diagonal block of L11
   double c0[PV_BS], c1[PV_BS], c2[PV_BS];
   int lperm[PV_BS];
};

/* One CTA per (front, 128-row tile): the rows below the segment.  Per 32-column block: (U) left-looking DMMA
 * update with the earlier blocks, T_j -= W_{<j} (L11 D)(j, <j)^T; (S) Y = T_j(:, lperm) X^T, the triangular
 * solve as a multiplication by the inverse diagonal block; (D) W = Y D^-1, threshold test |w| <= 1/u
 * (ldlt_app.cxx:303-321), W -> L, Y -> L*D.  The originals go to the backup when the tile is loaded. */
template <bool POSDEF>
__global__ void __launch_bounds__(TNT)
k_panel_tiles(Front* fronts, const RowTile* work, FactorParams prm) {
   extern __shared__ __align__(16) unsigned char pv_smem[];
   TileSm& sh = *reinterpret_cast<TileSm*>(pv_smem);
   const RowTile w = work[blockIdx.x];
   Front* f = &fronts[w.front];
   if (!f->seg_valid || !f->seg_ok) return;
   const int p = f->done, m = f->m;
   const int r0 = w.tile * RT;
   if (r0 + RT <= p + CW || r0 >= m) return;
   const size_t ldl = (size_t)f->ldl;
   double* Lp = f->L + (size_t)p * ldl;
   double* LDp = f->LD + (size_t)p * ldl;
   double* BK = f->BK;
   const SegWS* ws = f->sws;
   const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
   double* T = sh.T;

   {  /* load: two rows per thread, 64 threads per column (1 KB contiguous) */
      const int rr = (tid & 63) * 2, cq = tid >> 6;
      const int r = r0 + rr;
      const bool a0 = (r >= p + CW) && (r < m), a1 = (r + 1 >= p + CW) && (r + 1 < m);
      constexpr int NBL = 8;                                  // columns in flight per thread: all loads, then all stores
      for (int c0 = cq; c0 < CW; c0 += NBL * (TNT / 64)) {
         double2 v[NBL];
         #pragma unroll
         for (int q = 0; q < NBL; ++q) {
            const int c = c0 + q * (TNT / 64);
            v[q] = make_double2(0.0, 0.0);
            const double* src = Lp + r + (size_t)c * ldl;
            if (a0 && a1) v[q] = *reinterpret_cast<const double2*>(src);
            else { if (a0) v[q].x = src[0]; if (a1) v[q].y = src[1]; }
         }
         #pragma unroll
         for (int q = 0; q < NBL; ++q) {
            const int c = c0 + q * (TNT / 64);
            *reinterpret_cast<double2*>(&T[c * CLD + rr]) = v[q];
            if (!POSDEF) {                                      // Cholesky never rolls back
               double* dst = BK + r + (size_t)c * ldl;
               if (a0 && a1) *reinterpret_cast<double2*>(dst) = v[q];
               else { if (a0) dst[0] = v[q].x; if (a1) dst[1] = v[q].y; }
            }
         }
      }
   }
   const double lim = 1.0 / prm.u;
   int bad = 0;
   const int rbase = warp * 16;                              // the warp's rows in phases U and S
   for (int jb = 0; jb < CW; jb += PV_BS) {
      {  /* operands of this step: all loads of a thread first (up to 12 + 4), then the stores */
         double bv[(CW - PV_BS) * PV_BS / TNT], xv[PV_BS * PV_BS / TNT];
         #pragma unroll
         for (int q = 0; q < (CW - PV_BS) * PV_BS / TNT; ++q) {
            const int e = tid + q * TNT, n = e & 31, k = e >> 5;
            bv[q] = (e < jb * PV_BS) ? ws->ld11[(jb + n) + (size_t)k * CW] : 0.0;
         }
         #pragma unroll
         for (int q = 0; q < PV_BS * PV_BS / TNT; ++q) xv[q] = ws->invl[jb / PV_BS][tid + q * TNT];
         #pragma unroll
         for (int q = 0; q < (CW - PV_BS) * PV_BS / TNT; ++q) {
            const int e = tid + q * TNT, n = e & 31, k = e >> 5;
            if (e < jb * PV_BS) sh.Bs[k * PBLD + n] = bv[q];
         }
         #pragma unroll
         for (int q = 0; q < PV_BS * PV_BS / TNT; ++q) {
            const int e = tid + q * TNT, c = e & 31, k = e >> 5;
            sh.Xs[k * PBLD + c] = xv[q];
         }
      }
      if (tid < PV_BS) {
         double v0, v1 = 0.0, v2 = 0.0;
         if (POSDEF) v0 = 1.0;
         else pv_dinv_coeffs(ws->dinv + 2 * jb, tid, CUDART_INF, v0, v1, v2);
         sh.c0[tid] = v0; sh.c1[tid] = v1; sh.c2[tid] = v2;
         sh.lperm[tid] = ws->lperm[jb + tid];
      }
      __syncthreads();
      /* (U) the warp's 16 rows x the block's 32 columns */
      if (jb > 0) {
         double acc[4][2][2];
         #pragma unroll
         for (int j = 0; j < 4; ++j)
            #pragma unroll
            for (int i = 0; i < 2; ++i) { acc[j][i][0] = 0.0; acc[j][i][1] = 0.0; }
         for (int kk = 0; kk < jb; kk += 4) {
            double af[2], bf[4];
            #pragma unroll
            for (int i = 0; i < 2; ++i) af[i] = T[(kk + (lane & 3)) * CLD + rbase + i * 8 + (lane >> 2)];
            #pragma unroll
            for (int j = 0; j < 4; ++j) bf[j] = sh.Bs[(kk + (lane & 3)) * PBLD + j * 8 + (lane >> 2)];
            #pragma unroll
            for (int j = 0; j < 4; ++j)
               #pragma unroll
               for (int i = 0; i < 2; ++i) pv_dmma(acc[j][i][0], acc[j][i][1], bf[j], af[i]);
         }
         #pragma unroll
         for (int j = 0; j < 4; ++j)
            #pragma unroll
            for (int i = 0; i < 2; ++i) {
               double2* q = reinterpret_cast<double2*>(&T[(jb + j * 8 + (lane >> 2)) * CLD + rbase + i * 8 + 2 * (lane & 3)]);
               double2 v = *q;
               v.x -= acc[j][i][0]; v.y -= acc[j][i][1];
               *q = v;
            }
         __syncwarp();                                       // (S) reads what the other lanes of this warp wrote
      }
      /* (S) */
      {
         double acc[4][2][2];
         #pragma unroll
         for (int j = 0; j < 4; ++j)
            #pragma unroll
            for (int i = 0; i < 2; ++i) { acc[j][i][0] = 0.0; acc[j][i][1] = 0.0; }
         #pragma unroll
         for (int kk = 0; kk < PV_BS; kk += 4) {
            double af[2], bf[4];
            const int col = jb + sh.lperm[kk + (lane & 3)];
            #pragma unroll
            for (int i = 0; i < 2; ++i) af[i] = T[col * CLD + rbase + i * 8 + (lane >> 2)];
            #pragma unroll
            for (int j = 0; j < 4; ++j) bf[j] = sh.Xs[(kk + (lane & 3)) * PBLD + j * 8 + (lane >> 2)];
            #pragma unroll
            for (int j = 0; j < 4; ++j)
               #pragma unroll
               for (int i = 0; i < 2; ++i) pv_dmma(acc[j][i][0], acc[j][i][1], bf[j], af[i]);
         }
         #pragma unroll
         for (int j = 0; j < 4; ++j)
            #pragma unroll
            for (int i = 0; i < 2; ++i)
               *reinterpret_cast<double2*>(&sh.Ys[(j * 8 + (lane >> 2)) * CLD + rbase + i * 8 + 2 * (lane & 3)]) =
                  make_double2(acc[j][i][0], acc[j][i][1]);
      }
      __syncthreads();
      /* (D) one row and 16 columns per thread */
      {
         const int r = tid & (RT - 1), h = tid / RT;
         const int grow = r0 + r;
         const bool active = (grow >= p + CW) && (grow < m);
         double yv[18];                                       // columns h*16 - 1 .. h*16 + 16: loads first, then the stores
         #pragma unroll
         for (int q = 0; q < 18; ++q) {
            const int j = h * 16 - 1 + q;
            yv[q] = (j >= 0 && j < PV_BS) ? sh.Ys[j * CLD + r] : 0.0;
         }
         #pragma unroll
         for (int q = 1; q < 17; ++q) {
            const int j = h * 16 - 1 + q;
            const double y = yv[q];
            double wv = y;
            if (!POSDEF) {
               wv = sh.c0[j] * y + sh.c1[j] * yv[q + 1] + sh.c2[j] * yv[q - 1];
               if (active && !(fabs(wv) <= lim)) bad = 1;
            }
            T[(jb + j) * CLD + r] = wv;
            if (active) {
               Lp[grow + (size_t)(jb + j) * ldl] = wv;
               if (!POSDEF) LDp[grow + (size_t)(jb + j) * ldl] = y;      // Cholesky: L*D is L itself (f->LD == f->L)
            }
         }
      }
      __syncthreads();
   }
   if (bad) f->seg_fail = 1;
}

/* Accept (diagonal block, D, perm, row permutation of the earlier columns) or roll back.  One CTA per
 * (front, 128-row tile): on failure the rows below the segment are restored from the backup; on success
 * the CTAs left of the segment permute the segment's rows in their 128 already-factored columns (one
 * thread per row of the segment: coalesced), and the CTA whose tile holds row p writes the diagonal block,
 * D^-1 and the pivot order. */
template <bool POSDEF>
__global__ void __launch_bounds__(RT)
k_seg_commit(Front* fronts, const RowTile* work) {
   const RowTile w = work[blockIdx.x];
   Front* f = &fronts[w.front];
   if (!f->seg_valid || !f->seg_ok) return;
   const int p = f->done, m = f->m;
   const size_t ldl = (size_t)f->ldl;
   const int r0 = w.tile * RT;
   const int t = threadIdx.x;
   const SegWS* ws = f->sws;
   double* L = f->L;
   const bool diag_cta = (w.tile == p / RT);
   /* the diagonal block from the workspace to the front: thread t owns row t, 16 columns in flight (a load behind a
    * store through another pointer would wait for it) */
   auto write_diag_block = [&]() {
      for (int c0 = 0; c0 < CW; c0 += 16) {
         double v[16];
         #pragma unroll
         for (int q = 0; q < 16; ++q) v[q] = ws->l11[t + (size_t)(c0 + q) * CW];
         #pragma unroll
         for (int q = 0; q < 16; ++q) if (t >= c0 + q) L[(size_t)(p + t) + (size_t)(p + c0 + q) * ldl] = v[q];
      }
   };
   if (POSDEF) {                      /* no permutation, no D: only the diagonal block is left to write */
      if (diag_cta) write_diag_block();
      return;
   }
   if (f->seg_fail) {
      const int r = r0 + t;
      if (r >= p + CW && r < m) {
         const double* BKr = f->BK + r;
         double* Lr = L + r + (size_t)p * ldl;
         #pragma unroll 16
         for (int c = 0; c < CW; ++c) Lr[(size_t)c * ldl] = BKr[(size_t)c * ldl];
      }
      return;
   }
   __shared__ int s_lperm[CW];
   __shared__ int s_perm[CW];
   __shared__ int s_moved;
   if (t == 0) s_moved = 0;
   __syncthreads();
   const int src = (t & ~(PV_BS - 1)) + ws->lperm[t];
   s_lperm[t] = src;
   if (src != t) s_moved = 1;
   __syncthreads();
   if (r0 < p && s_moved) {
      const int cend = min(r0 + RT, p);
      constexpr int NB = 16;
      for (int c0 = r0; c0 < cend; c0 += NB) {
         double v[NB];
         #pragma unroll
         for (int q = 0; q < NB; ++q) v[q] = (c0 + q < cend) ? L[(size_t)(c0 + q) * ldl + p + src] : 0.0;
         __syncthreads();
         #pragma unroll
         for (int q = 0; q < NB; ++q) if (c0 + q < cend) L[(size_t)(c0 + q) * ldl + p + t] = v[q];
      }
   }
   if (diag_cta) {
      write_diag_block();
      for (int e = t; e < 2 * CW; e += RT) f->D[2 * p + e] = ws->dinv[e];
      s_perm[t] = f->perm[p + src];
      __syncthreads();
      f->perm[p + t] = s_perm[t];
   }
}

int panel_segment_width() { return CW; }
size_t panel_segment_ws_bytes() { return sizeof(SegWS); }

void configure_panel_kernels() {
   cudaFuncSetAttribute(k_panel_chain<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ChainSm));
   cudaFuncSetAttribute(k_panel_chain<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ChainSm));
   cudaFuncSetAttribute(k_panel_tiles<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TileSm));
   cudaFuncSetAttribute(k_panel_tiles<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TileSm));
}

void launch_panel_chain(Front* fronts, const int* flist, int count, bool posdef, bool new_panel, const FactorParams& prm,
      cudaStream_t s) {
   if (count == 0) return;
   if (posdef) k_panel_chain<true><<<count, CH_NT, sizeof(ChainSm), s>>>(fronts, flist, new_panel ? 1 : 0, prm);
   else k_panel_chain<false><<<count, CH_NT, sizeof(ChainSm), s>>>(fronts, flist, new_panel ? 1 : 0, prm);
   COUNT_LAUNCH();
}
void launch_panel_tiles(Front* fronts, const RowTile* work, int nwork, bool posdef, const FactorParams& prm, cudaStream_t s) {
   if (nwork == 0) return;
   if (posdef) k_panel_tiles<true><<<nwork, TNT, sizeof(TileSm), s>>>(fronts, work, prm);
   else k_panel_tiles<false><<<nwork, TNT, sizeof(TileSm), s>>>(fronts, work, prm);
   COUNT_LAUNCH();
}
void launch_seg_commit(Front* fronts, const RowTile* work, int nwork, bool posdef, cudaStream_t s) {
   if (nwork == 0) return;
   if (posdef) k_seg_commit<true><<<nwork, RT, 0, s>>>(fronts, work);
   else k_seg_commit<false><<<nwork, RT, 0, s>>>(fronts, work);
   COUNT_LAUNCH();
}

/* ------------------------------------------------------------------------ */
/* Apply the block pivots to the rows below                                  */
/* ------------------------------------------------------------------------ */

/* One thread per row.  Indefinite: Y = A21(:,lperm) L11^-T, W = Y D^-1, the
 * a-posteriori threshold test |w| <= 1/u (check_threshold,
 * src/ssids/cpu/kernels/ldlt_app.cxx:303-321; apply_pivot :332-377).
 * W overwrites the block column, Y goes to LD, the originals to BK. */
template <bool POSDEF>
__global__ void __launch_bounds__(RT)
k_apply(Front* fronts, const RowTile* work, FactorParams prm) {
   RowTile w = work[blockIdx.x];
   Front* f = &fronts[w.front];
   if (!f->step_valid) return;
   const int done = f->done, bs = f->bs, m = f->m;
   const int r0 = w.tile * RT;
   const int rbeg = done + bs;
   if (r0 + RT <= rbeg || r0 >= m) return;

   __shared__ double l11[BS][BS + 1];   // l11[j][k]
   __shared__ double c0[BS], c1[BS], c2[BS];
   __shared__ int lperm[BS];
   __shared__ int s_fail;
   const BlockWS* ws = f->ws;
   for (int i = threadIdx.x; i < BS * BS; i += RT) l11[i % BS][i / BS] = ws->l11[i];
   if (threadIdx.x < BS) {
      int j = threadIdx.x;
      if (POSDEF) {
         c0[j] = ws->dinv[j]; c1[j] = 0.0; c2[j] = 0.0; lperm[j] = j;
      } else {
         const double* d = ws->dinv;
         double v0, v1 = 0.0, v2 = 0.0;
         if (j < bs && isinf(d[2 * j])) { v0 = d[2 * j + 1]; v2 = d[2 * j - 1]; }          // second of a 2x2
         else if (j + 1 < bs && isinf(d[2 * j + 2])) { v0 = d[2 * j]; v1 = d[2 * j + 1]; } // first of a 2x2
         else v0 = (j < bs) ? d[2 * j] : 0.0;
         c0[j] = v0; c1[j] = v1; c2[j] = v2;
         lperm[j] = ws->lperm[j];
      }
   }
   if (threadIdx.x == 0) s_fail = BS;
   __syncthreads();
   /* columns >= zfrom were declared zero pivots from the diagonal block alone;
    * they are only zero columns if the rows below are (numerically) zero too
    * (ldlt_tpp.cxx:179-188 tests the whole column) -- otherwise they fail */
   const int zfrom = POSDEF ? BS : ws->zfrom;

   const int r = r0 + threadIdx.x;
   const bool active = (r >= rbeg) && (r < m);
   const size_t ldl = f->ldl;
   double* Lr = f->L + r + (size_t)done * ldl;
   double y[BS];
   #pragma unroll
   for (int j = 0; j < BS; ++j)
      y[j] = (active && j < bs) ? Lr[(size_t)lperm[j] * ldl] : 0.0;

   if (POSDEF) {
      #pragma unroll
      for (int j = 0; j < BS; ++j) {
         double s = y[j];
         #pragma unroll
         for (int k = 0; k < j; ++k) s -= y[k] * l11[j][k];
         y[j] = s * c0[j];
      }
      if (active) {
         #pragma unroll
         for (int j = 0; j < BS; ++j)
            if (j < bs) Lr[(size_t)j * ldl] = y[j];
      }
      return;
   }

   double* BKr = f->BK + r;
   double* LDr = f->LD + r + (size_t)done * ldl;
   if (active) {
      #pragma unroll
      for (int j = 0; j < BS; ++j)
         if (j < bs) BKr[(size_t)j * ldl] = y[j];
   }
   #pragma unroll
   for (int j = 0; j < BS; ++j) {
      double s = y[j];
      #pragma unroll
      for (int k = 0; k < j; ++k) s -= y[k] * l11[j][k];
      y[j] = s;
   }
   int myfail = BS;
   const double lim = 1.0 / prm.u;
   if (active) {
      #pragma unroll
      for (int j = 0; j < BS; ++j) {
         if (j < bs) {
            double wv = c0[j] * y[j];
            if (j + 1 < BS) wv += c1[j] * y[(j + 1) % BS];
            if (j > 0) wv += c2[j] * y[(j + BS - 1) % BS];
            Lr[(size_t)j * ldl] = wv;
            LDr[(size_t)j * ldl] = y[j];
            bool bad = !(fabs(wv) <= lim) || (j >= zfrom && !(fabs(y[j]) < prm.small));
            if (bad && myfail == BS) myfail = j;
         }
      }
   }
   #pragma unroll
   for (int off = 16; off > 0; off >>= 1)
      myfail = min(myfail, __shfl_xor_sync(0xffffffffu, myfail, off));
   if ((threadIdx.x & 31) == 0 && myfail < BS) atomicMin(&s_fail, myfail);
   __syncthreads();
   if (threadIdx.x == 0 && s_fail < BS) atomicMin(&f->first_fail, s_fail);
}

void launch_apply(Front* fronts, const RowTile* work, int nwork, bool posdef,
      const FactorParams& prm, cudaStream_t s) {
   if (nwork == 0) return;
   if (posdef) k_apply<true><<<nwork, RT, 0, s>>>(fronts, work, prm);
   else k_apply<false><<<nwork, RT, 0, s>>>(fronts, work, prm); COUNT_LAUNCH();
}

/* ------------------------------------------------------------------------ */
/* Commit an inner step (indefinite only)                                    */
/* ------------------------------------------------------------------------ */

__global__ void __launch_bounds__(RT)
k_commit(Front* fronts, const RowTile* work) {
   RowTile w = work[blockIdx.x];
   Front* f = &fronts[w.front];
   if (!f->step_valid) return;
   const int done = f->done, bs = f->bs, m = f->m;
   const int ne = calc_ne(f);
   const size_t ldl = f->ldl;
   const int r0 = w.tile * RT;
   const BlockWS* ws = f->ws;
   double* L = f->L;

   __shared__ int lperm[BS];
   __shared__ int s_ident;
   __shared__ int s_perm[BS];
   if (threadIdx.x == 0) s_ident = 1;
   __syncthreads();
   if (threadIdx.x < BS) {
      int q = (threadIdx.x < bs) ? ws->lperm[threadIdx.x] : threadIdx.x;
      lperm[threadIdx.x] = q;
      if (q != threadIdx.x) s_ident = 0;
   }
   __syncthreads();

   /* (1) rows of the block in the already-factored columns c < done: the
    * block's local permutation is a row permutation of L */
   if (r0 < done && !s_ident) {
      int c = r0 + threadIdx.x;
      if (c < done) {
         double* col = L + (size_t)c * ldl + done;
         double v[BS];
         #pragma unroll
         for (int i = 0; i < BS; ++i) v[i] = (i < bs) ? col[lperm[i]] : 0.0;
         #pragma unroll
         for (int i = 0; i < BS; ++i) if (i < bs) col[i] = v[i];
      }
   }

   /* (2) the diagonal block itself, D and perm: one CTA per front */
   if (w.tile == done / RT) {
      double* LD = f->LD;
      for (int e = threadIdx.x; e < BS * BS; e += RT) {
         int i = e % BS, j = e / BS;
         if (i < bs && j < bs && i >= j) {
            double v = (j < ne) ? ws->l11[i + j * BS] : ws->a0[lperm[i] + lperm[j] * BS];
            L[(size_t)(done + i) + (size_t)(done + j) * ldl] = v;
            if (j < ne && i >= ne) LD[(size_t)(done + i) + (size_t)(done + j) * ldl] = ws->ld11[i + j * BS];
         }
      }
      if (threadIdx.x < 2 * ne) f->D[2 * done + threadIdx.x] = ws->dinv[threadIdx.x];
      if (threadIdx.x < bs) s_perm[threadIdx.x] = f->perm[done + lperm[threadIdx.x]];
      __syncthreads();
      if (threadIdx.x < bs) f->perm[done + threadIdx.x] = s_perm[threadIdx.x];
   }

   /* (3) rows below the block: restore the failed columns (permuted originals) */
   if (ne < bs) {
      int r = r0 + threadIdx.x;
      if (r >= done + bs && r < m) {
         const double* BKr = f->BK + r;
         double* Lr = L + r + (size_t)done * ldl;
         for (int j = ne; j < bs; ++j) Lr[(size_t)j * ldl] = BKr[(size_t)j * ldl];
      }
   }
}

void launch_commit(Front* fronts, const RowTile* work, int nwork, cudaStream_t s) {
   if (nwork == 0) return;
   k_commit<<<nwork, RT, 0, s>>>(fronts, work); COUNT_LAUNCH();
}

/* ------------------------------------------------------------------------ */
/* Move failed columns out of the way (symmetric swaps, lower storage)       */
/* ------------------------------------------------------------------------ */

__device__ __forceinline__ double* sym_addr(double* L, size_t ldl, int i, int x) {
   return (i > x) ? L + i + (size_t)x * ldl : L + x + (size_t)i * ldl;
}

/* Swaps positions a0+t <-> b0+t, t < nswap (a0+nswap <= b0), of the symmetric
 * front.  inner: failed columns of the current block go to the end of the
 * panel; outer: failed columns of the panel go to the end of the candidates. */
__global__ void __launch_bounds__(RT)
k_swap(Front* fronts, const RowTile* work, int outer, int slice) {
   RowTile w = work[blockIdx.x];
   Front* f = &fronts[w.front];
   int a0, b0, nswap;
   if (!outer) {
      if (!f->step_valid) return;
      int ne = calc_ne(f);
      int nfail = f->bs - ne;
      if (nfail == 0) return;
      a0 = f->done + ne;
      int rem = f->pend - (f->done + f->bs);
      nswap = min(nfail, rem);
      b0 = f->pend - nswap;
   } else {
      if (!f->panel_open || f->finished) return;
      int pend = f->pend;
      if (f->step_valid) pend -= f->bs - calc_ne(f);
      int nf = f->pend0 - pend;
      if (nf == 0) return;
      a0 = pend;
      int rem = f->end - f->pend0;
      nswap = min(nf, rem);
      b0 = f->end - nswap;
   }
   /* disjoint transpositions commute: the pairs are applied BS at a time */
   a0 += slice * BS; b0 += slice * BS; nswap = min(BS, nswap - slice * BS);
   if (nswap <= 0) return;
   const size_t ldl = f->ldl;
   const int m = f->m;
   double* L = f->L;
   int x = w.tile * RT + threadIdx.x;
   bool inS = (x >= a0 && x < a0 + nswap) || (x >= b0 && x < b0 + nswap);
   if (x < m && !inS) {
      for (int t = 0; t < nswap; ++t) {
         double* pa = sym_addr(L, ldl, a0 + t, x);
         double* pb = sym_addr(L, ldl, b0 + t, x);
         double v = *pa; *pa = *pb; *pb = v;
      }
   }
   if (w.tile == 0) {
      /* the S x S block: entry (s_k1, s_k2) <- (partner(k1), partner(k2)) */
      __shared__ double sbuf[2 * BS * 2 * BS];
      int ns2 = 2 * nswap;
      for (int e = threadIdx.x; e < ns2 * ns2; e += RT) {
         int k1 = e % ns2, k2 = e / ns2;
         if (k1 >= k2) {
            int s1 = (k1 < nswap) ? a0 + k1 : b0 + k1 - nswap;
            int s2 = (k2 < nswap) ? a0 + k2 : b0 + k2 - nswap;
            sbuf[k1 + k2 * ns2] = L[(size_t)s1 + (size_t)s2 * ldl];
         }
      }
      __syncthreads();
      for (int e = threadIdx.x; e < ns2 * ns2; e += RT) {
         int k1 = e % ns2, k2 = e / ns2;
         if (k1 >= k2) {
            int s1 = (k1 < nswap) ? a0 + k1 : b0 + k1 - nswap;
            int s2 = (k2 < nswap) ? a0 + k2 : b0 + k2 - nswap;
            int q1 = (k1 < nswap) ? k1 + nswap : k1 - nswap;
            int q2 = (k2 < nswap) ? k2 + nswap : k2 - nswap;
            double v = (q1 >= q2) ? sbuf[q1 + q2 * ns2] : sbuf[q2 + q1 * ns2];
            L[(size_t)s1 + (size_t)s2 * ldl] = v;
         }
      }
      int* perm = f->perm;
      for (int t = threadIdx.x; t < nswap; t += RT) {
         int q = perm[a0 + t]; perm[a0 + t] = perm[b0 + t]; perm[b0 + t] = q;
      }
   }
}

void launch_swap(Front* fronts, const RowTile* work, int nwork, bool outer, cudaStream_t s) {
   if (nwork == 0) return;
   int nslice = outer ? PW / BS : 1;
   for (int sl = 0; sl < nslice; ++sl)
      k_swap<<<nwork, RT, 0, s>>>(fronts, work, outer ? 1 : 0, sl); COUNT_LAUNCH();
}

/* ------------------------------------------------------------------------ */
/* Panel snapshot for the host                                               */
/* ------------------------------------------------------------------------ */

/* Accounts the last inner step of the panel and reports the state the host
 * needs to size the outer update exactly: p0, done, pend, pend0, end,
 * finished, flag (8 ints per front). */
__global__ void k_snapshot(Front* fronts, const int* __restrict__ flist, int count, int* __restrict__ snap) {
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= count) return;
   snapshot_state(&fronts[flist[i]], snap + (size_t)i * 8);
}

void launch_snapshot(Front* fronts, const int* flist, int count, int* snap, cudaStream_t s) {
   if (count == 0) return;
   k_snapshot<<<(count + 127) / 128, 128, 0, s>>>(fronts, flist, count, snap); COUNT_LAUNCH();
}

/* ------------------------------------------------------------------------ */
/* Finalise a level: close the state machine, collect statistics             */
/* ------------------------------------------------------------------------ */

template <bool POSDEF>
__global__ void __launch_bounds__(128)
k_finalize(Front* fronts, const int* __restrict__ flist) {
   Front* f = &fronts[flist[blockIdx.x]];
   if (threadIdx.x == 0) advance_state(f, true);
   __syncthreads();
   if (!f->finished || POSDEF) return;
   __shared__ int s_neg, s_two, s_zero;
   if (threadIdx.x == 0) { s_neg = 0; s_two = 0; s_zero = 0; }
   __syncthreads();
   const int nelim = f->nelim;
   const double* d = f->D;
   int neg = 0, two = 0, zero = 0;
   for (int i = threadIdx.x; i < nelim; i += blockDim.x) {
      double a11 = d[2 * i];
      if (isinf(a11)) continue;                       // second column of a 2x2
      double a21 = d[2 * i + 1];
      if (i + 1 == nelim || !isinf(d[2 * i + 2])) {   // 1x1 (or zero)
         if (a11 == 0.0) zero++;
         if (a11 < 0.0) neg++;
      } else {
         double a22 = d[2 * i + 3];
         two++;
         double det = a11 * a22 - a21 * a21;
         double trace = a11 + a22;
         if (det < 0) neg++;
         else if (trace < 0) neg += 2;
      }
   }
   atomicAdd(&s_neg, neg); atomicAdd(&s_two, two); atomicAdd(&s_zero, zero);
   __syncthreads();
   if (threadIdx.x == 0) { f->num_neg = s_neg; f->num_two = s_two; f->num_zero = s_zero; }
}

void launch_finalize(Front* fronts, const int* flist, int count, bool posdef, cudaStream_t s) {
   if (count == 0) return;
   if (posdef) k_finalize<true><<<count, 128, 0, s>>>(fronts, flist);
   else k_finalize<false><<<count, 128, 0, s>>>(fronts, flist); COUNT_LAUNCH();
}

} // namespace b200
