/* Speculative panel factorisation, version 2 (opt-in, SPRAL_B200_PANEL_V2=1).
 *
 * The step-by-step path (factor_kernels.cu) spends five dependent launches on every
 * 32 columns: diagonal block, apply + threshold test, commit, in-panel update, swap.
 * On the large fronts at the top of the tree that chain -- not the FP64 tensor work --
 * is the critical path.  APP pivoting (reference CPU engine, ldlt_app.cxx) chooses
 * pivots inside the diagonal block only and tests the rows below a posteriori, so a
 * segment of CW = 128 columns can be factorised in two passes instead of twenty launches:
 *
 *   chain  ONE CTA per front factorises the CW x CW diagonal block of the segment in
 *          shared memory: four 32 x 32 blocks (diag_block.h, the same pivoting rules),
 *          each followed by the triangular solve, D^-1 scaling, threshold test and
 *          trailing update of the rows of the diagonal block below it.  Nothing is
 *          written to the front; L11, L11*D, D^-1 and the block permutations go to a
 *          workspace (SegWS).  Any failed pivot / zero pivot gives up (ok = 0).
 *   tiles  every 128-row tile of the rows below reads its CW columns ONCE, runs the
 *          four steps locally (solve against L11^T, scale, threshold test, update of the
 *          tile's later columns with L11*D from the workspace) and writes L, L*D and a
 *          backup of the originals.  No communication between tiles.
 *   commit if every tile passed: diagonal block, D, perm and the row permutation of the
 *          earlier columns are written and the state machine advances by CW columns;
 *          otherwise the tiles restore their rows and the step-by-step path redoes the
 *          segment (failed pivots are rare: 23 delays in 10^6 columns on the benchmark).
 *
 * The bodies are written against a context (tid / sync / shuffles) and compiled twice:
 * by nvcc for the kernels in factor_kernels.cu and by g++ for tests/c/panel_v2_emu.cpp,
 * which runs them on host threads and checks P A P^T = L D L^T on the diagonal block,
 * A21 P^T = (W D) L11^T on the rows below, and the give-up paths.
 */
#pragma once
#include <cmath>
#include <cstddef>
#include "diag_block.h"
#include "solve_types.h"

#ifdef __CUDACC__
#define PV_FN __device__ __forceinline__
#else
#define PV_FN inline
#endif

namespace b200 {

constexpr int CW = 128;            // segment width (columns per chain / tiles / commit round)
constexpr int CLD = CW + 1;        // shared-memory column stride (bank = row + column)
constexpr int CNT = CW;            // threads of the chain kernel: one per row of the diagonal block
constexpr int PV_YLD = DB_BS + 2;   // even: rows stay 16-byte aligned, the broadcast reads of a row vectorise (LDS.128)

/* Per-front workspace of a segment (device global memory). */
struct SegWS {
   double l11[CW * CW];     // unit lower factor of the diagonal block, final row order, column-major ld = CW
   double ld11[CW * CW];    // (L11 D)(i, k), i > k, row i in the order it had when column k was eliminated
   double dinv[2 * CW];     // D^-1, reference CPU layout (block_ldlt.hxx:375-406)
   int lperm[CW];           // per 32-block: position jb + i holds old position jb + lperm[jb + i]
   int ok;                  // the chain kernel factorised the whole diagonal block without a failed pivot
   int pad_[3];
};

struct ChainShared {
   DiagShared<4> dg;
   double S[CW * CLD];      // column-major: A / L below the diagonal, L*D mirrored above it
   double c0[DB_BS], c1[DB_BS], c2[DB_BS];
   int lperm[DB_BS];
   int fail;
};

/* D^-1 of a 32-column block as three coefficient vectors: w_j = c0 y_j + c1 y_{j+1} + c2 y_{j-1} */
PV_FN void pv_dinv_coeffs(const double* d, int j, double inf, double& v0, double& v1, double& v2) {
   v1 = 0.0; v2 = 0.0;
   if (d[2 * j] == inf) { v0 = d[2 * j + 1]; v2 = d[2 * j - 1]; }                          // second column of a 2x2
   else if (j + 1 < DB_BS && d[2 * j + 2] == inf) { v0 = d[2 * j]; v1 = d[2 * j + 1]; }   // first column of a 2x2
   else v0 = d[2 * j];
}

/* Chain: factorises the CW x CW diagonal block whose lower triangle is at Lseg (ld = ldl).
 * CNT threads.  Returns 1 (every thread) when `out` holds the factors, 0 when it gave up. */
template <bool POSDEF, class Ctx>
PV_FN int chain_segment(Ctx& cx, ChainShared& sh, const double* Lseg, size_t ldl, double u, double small,
      double inf, SegWS* out) {
   constexpr int BS = DB_BS;
   const int t = cx.tid();
   double* S = sh.S;
   #pragma unroll 8
   for (int c = 0; c < CW; ++c) S[(size_t)c * CLD + t] = (t >= c) ? Lseg[t + (size_t)c * ldl] : 0.0;
   if (t == 0) sh.fail = 0;
   cx.sync();
   const double lim = 1.0 / u;
   for (int jb = 0; jb < CW; jb += BS) {
      int cur = 0, zfrom = BS;
      const int rc = diag_block_factor<4, POSDEF>(cx, sh.dg, S + (size_t)jb * CLD + jb, (size_t)CLD, BS, small, 1, inf,
                                                  (double*)nullptr, cur, zfrom);
      if (rc != DB_OK || zfrom < BS) return 0;     // not positive definite / zero pivots: the step-by-step path reports it
      if (t < BS) {
         double v0, v1 = 0.0, v2 = 0.0;
         if (POSDEF) v0 = sh.dg.dinv[t];            // 1 / l_tt (cholesky_factor, cholesky.cxx:33-187)
         else pv_dinv_coeffs(sh.dg.dinv, t, inf, v0, v1, v2);
         sh.c0[t] = v0; sh.c1[t] = v1; sh.c2[t] = v2;
         sh.lperm[t] = POSDEF ? t : sh.dg.lperm[t];
      }
      cx.sync();
      double wv[BS];
      #pragma unroll
      for (int j = 0; j < BS; ++j) wv[j] = 0.0;
      if (t >= jb + BS) {
         /* rows of the diagonal block below the 32 x 32 block: Y = A21(:, lperm) L11^-T, W = Y D^-1,
          * a-posteriori test |w| <= 1/u (ldlt_app.cxx:303-321) */
         double y[BS];
         #pragma unroll
         for (int j = 0; j < BS; ++j) y[j] = S[(size_t)(jb + sh.lperm[j]) * CLD + t];
         int bad = 0;
         if (POSDEF) {                                    /* l_tj = (a_tj - sum_k l_tk l_jk) / l_jj */
            #pragma unroll
            for (int j = 0; j < BS; ++j) {
               double s = y[j];
               #pragma unroll
               for (int k = 0; k < j; ++k) s -= y[k] * sh.dg.A[cur][j][k];
               y[j] = s * sh.c0[j];
               wv[j] = y[j];
            }
         } else {
            #pragma unroll
            for (int j = 0; j < BS; ++j) {
               double s = y[j];
               #pragma unroll
               for (int k = 0; k < j; ++k) s -= y[k] * sh.dg.A[cur][j][k];
               y[j] = s;
            }
            #pragma unroll
            for (int j = 0; j < BS; ++j) {
               double w = sh.c0[j] * y[j];
               if (j + 1 < BS) w += sh.c1[j] * y[(j + 1) % BS];
               if (j > 0) w += sh.c2[j] * y[(j + BS - 1) % BS];
               wv[j] = w;
               if (!(fabs(w) <= lim)) bad = 1;
            }
         }
         #pragma unroll
         for (int j = 0; j < BS; ++j) {
            S[(size_t)(jb + j) * CLD + t] = wv[j];       // L(t, jb + j)
            S[(size_t)t * CLD + jb + j] = y[j];          // (L D)(t, jb + j), mirrored (== L for Cholesky)
         }
         if (bad) sh.fail = 1;
      } else if (t >= jb) {
         const int i = t - jb;                            // a row of the 32 x 32 block itself
         #pragma unroll
         for (int c = 0; c < BS; ++c) {
            if (c < i) {
               S[(size_t)(jb + c) * CLD + t] = sh.dg.A[cur][i][c];
               S[(size_t)t * CLD + jb + c] = POSDEF ? sh.dg.A[cur][i][c] : sh.dg.LDm[cur][i][c];
            } else if (c == i) S[(size_t)(jb + c) * CLD + t] = POSDEF ? sh.dg.A[cur][i][i] : 1.0;
         }
      } else if (!POSDEF) {
         /* an earlier column of the segment: the block's permutation is a row permutation of L */
         double v[BS];
         #pragma unroll
         for (int i = 0; i < BS; ++i) v[i] = S[(size_t)t * CLD + jb + sh.lperm[i]];
         #pragma unroll
         for (int i = 0; i < BS; ++i) S[(size_t)t * CLD + jb + i] = v[i];
      }
      cx.sync();
      if (sh.fail) return 0;
      if (t >= jb + BS) {
         for (int c = jb + BS; c <= t; ++c) {            // A(t, c) -= sum_k L(t, jb+k) (L D)(c, jb+k)
            const double* Yc = S + (size_t)c * CLD + jb;
            double s = S[(size_t)c * CLD + t];
            #pragma unroll
            for (int k = 0; k < BS; ++k) s -= wv[k] * Yc[k];
            S[(size_t)c * CLD + t] = s;
         }
      }
      if (POSDEF) { if (t < BS) out->dinv[jb + t] = sh.dg.dinv[t]; }      // 1 / l_jj, one per column
      else if (t < 2 * BS) out->dinv[2 * jb + t] = sh.dg.dinv[t];
      if (t < BS) out->lperm[jb + t] = sh.lperm[t];
      cx.sync();
   }
   for (int c = 0; c < CW; ++c) {
      out->l11[t + (size_t)c * CW] = (t > c) ? S[(size_t)c * CLD + t] : (t == c ? (POSDEF ? S[(size_t)c * CLD + c] : 1.0) : 0.0);
      out->ld11[t + (size_t)c * CW] = (t > c) ? S[(size_t)t * CLD + c] : 0.0;
   }
   return 1;
}

struct TileShared {
   double T[CW * CLD];                 // the tile, column-major (RT rows, stride CLD)
   double l11[DB_BS * PV_YLD];         // diagonal 32 x 32 block of L11 for the current step
   double ys[(CW - DB_BS) * PV_YLD];   // (L11 D)(c, jb + k) for the later columns c of the segment
   double c0[DB_BS], c1[DB_BS], c2[DB_BS];
   int lperm[DB_BS];
};

/* Tiles: rows [r0, r0 + RT) of the front below the segment (r >= p + CW), RT threads.
 * Lp / LDp: column p of L / of the L*D scratch; BK: backup, same layout, CW columns.
 * Sets *fail when an entry violates |l_ij| <= 1/u (the segment is then rolled back). */
template <bool POSDEF, class Ctx>
PV_FN void panel_tile(Ctx& cx, TileShared& sh, double* Lp, double* LDp, double* BK, size_t ldl, int m, int r0,
      int p, double u, double inf, const SegWS* ws, int* fail) {
   constexpr int BS = DB_BS;
   const int t = cx.tid();
   const int r = r0 + t;
   const bool active = (r >= p + CW) && (r < m);
   #pragma unroll 16
   for (int c = 0; c < CW; ++c) {       /* independent loads: many in flight per thread */
      const double v = active ? Lp[r + (size_t)c * ldl] : 0.0;
      sh.T[(size_t)c * CLD + t] = v;
      if (!POSDEF && active) BK[r + (size_t)c * ldl] = v;       // Cholesky never rolls back
   }
   const double lim = 1.0 / u;
   int bad = 0;
   for (int jb = 0; jb < CW; jb += BS) {
      for (int e = t; e < BS * BS; e += RT) {
         const int i = e % BS, j = e / BS;
         sh.l11[i * PV_YLD + j] = ws->l11[(jb + i) + (size_t)(jb + j) * CW];
      }
      const int ncc = CW - jb - BS;                       // later columns of the segment
      for (int e = t; e < ncc * BS; e += RT) {
         const int cc = e % ncc, k = e / ncc;
         sh.ys[cc * PV_YLD + k] = ws->ld11[(jb + BS + cc) + (size_t)(jb + k) * CW];
      }
      if (t < BS) {
         double v0, v1 = 0.0, v2 = 0.0;
         if (POSDEF) v0 = ws->dinv[jb + t];
         else pv_dinv_coeffs(ws->dinv + 2 * jb, t, inf, v0, v1, v2);
         sh.c0[t] = v0; sh.c1[t] = v1; sh.c2[t] = v2;
         sh.lperm[t] = ws->lperm[jb + t];
      }
      cx.sync();
      if (active) {
         double y[BS], wv[BS];
         #pragma unroll
         for (int j = 0; j < BS; ++j) y[j] = sh.T[(size_t)(jb + sh.lperm[j]) * CLD + t];
         if (POSDEF) {
            #pragma unroll
            for (int j = 0; j < BS; ++j) {
               double s = y[j];
               #pragma unroll
               for (int k = 0; k < j; ++k) s -= y[k] * sh.l11[j * PV_YLD + k];
               y[j] = s * sh.c0[j];
               wv[j] = y[j];
            }
         } else {
            #pragma unroll
            for (int j = 0; j < BS; ++j) {
               double s = y[j];
               #pragma unroll
               for (int k = 0; k < j; ++k) s -= y[k] * sh.l11[j * PV_YLD + k];
               y[j] = s;
            }
            #pragma unroll
            for (int j = 0; j < BS; ++j) {
               double w = sh.c0[j] * y[j];
               if (j + 1 < BS) w += sh.c1[j] * y[(j + 1) % BS];
               if (j > 0) w += sh.c2[j] * y[(j + BS - 1) % BS];
               wv[j] = w;
               if (!(fabs(w) <= lim)) bad = 1;
            }
         }
         #pragma unroll
         for (int j = 0; j < BS; ++j) {
            sh.T[(size_t)(jb + j) * CLD + t] = wv[j];
            if (!POSDEF) LDp[r + (size_t)(jb + j) * ldl] = y[j];      // Cholesky: L*D is L itself (f->LD == f->L)
         }
         for (int cc = 0; cc < ncc; ++cc) {               // A(r, c) -= sum_k L(r, jb+k) (L D)(c, jb+k)
            const double* Yc = sh.ys + cc * PV_YLD;
            double s = sh.T[(size_t)(jb + BS + cc) * CLD + t];
            #pragma unroll
            for (int k = 0; k < BS; ++k) s -= wv[k] * Yc[k];
            sh.T[(size_t)(jb + BS + cc) * CLD + t] = s;
         }
      }
      cx.sync();
   }
   if (active) {
      #pragma unroll 16
      for (int c = 0; c < CW; ++c) Lp[r + (size_t)c * ldl] = sh.T[(size_t)c * CLD + t];
      if (bad) *fail = 1;
   }
}

struct CommitShared { int lperm[CW]; int perm[CW]; };

/* Commit of a segment by the CTA of row tile r0 (RT threads): on failure the rows below
 * the segment are restored from the backup; on success (1) the rows of the segment in the
 * already-factored columns c < p are permuted block by block, (2) the CTA whose tile holds
 * row p (diag_cta) writes the diagonal block, D^-1 and the pivot order. */
template <bool POSDEF, class Ctx>
PV_FN void seg_commit(Ctx& cx, CommitShared& sh, double* L, double* D, int* perm, const double* BK, size_t ldl,
      int m, int p, int r0, bool diag_cta, int seg_fail, const SegWS* ws) {
   constexpr int BS = DB_BS;
   const int t = cx.tid();
   if (POSDEF) {                      /* no permutation, no D: only the diagonal block is left to write */
      if (diag_cta)
         for (int e = t; e < CW * CW; e += RT) {
            const int i = e % CW, c = e / CW;
            if (i >= c) L[(size_t)(p + i) + (size_t)(p + c) * ldl] = ws->l11[e];
         }
      return;
   }
   if (seg_fail) {
      const int r = r0 + t;
      if (r >= p + CW && r < m) {
         const double* BKr = BK + r;
         double* Lr = L + r + (size_t)p * ldl;
         #pragma unroll 16
         for (int c = 0; c < CW; ++c) Lr[(size_t)c * ldl] = BKr[(size_t)c * ldl];
      }
      return;
   }
   for (int i = t; i < CW; i += RT) sh.lperm[i] = ws->lperm[i];
   cx.sync();
   if (r0 < p) {
      const int c = r0 + t;
      if (c < p) {
         double* col = L + (size_t)c * ldl + p;
         for (int jb = 0; jb < CW; jb += BS) {
            double v[BS];
            #pragma unroll
            for (int i = 0; i < BS; ++i) v[i] = col[jb + sh.lperm[jb + i]];
            #pragma unroll
            for (int i = 0; i < BS; ++i) col[jb + i] = v[i];
         }
      }
   }
   if (diag_cta) {
      for (int e = t; e < CW * CW; e += RT) {
         const int i = e % CW, c = e / CW;
         if (i >= c) L[(size_t)(p + i) + (size_t)(p + c) * ldl] = ws->l11[e];
      }
      for (int e = t; e < 2 * CW; e += RT) D[2 * p + e] = ws->dinv[e];
      for (int i = t; i < CW; i += RT) sh.perm[i] = perm[p + (i / BS) * BS + sh.lperm[i]];
      cx.sync();
      for (int i = t; i < CW; i += RT) perm[p + i] = sh.perm[i];
   }
}

} // namespace b200
