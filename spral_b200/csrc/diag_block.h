/* Factorisation of one BS x BS diagonal block with pivoting inside the block:
 * the body of k_diag_v2 (factor_kernels.cu), written once and compiled twice --
 * by nvcc as device code, and by g++ for tests/c/diag_block_emu.cpp, which runs
 * it on host threads (one per CUDA thread, barriers and shuffles emulated) and
 * compares it bit for bit with a sequential model of the thread-per-entry k_diag.
 *
 * Semantics: block_ldlt of the reference CPU engine
 * (src/ssids/cpu/kernels/block_ldlt.hxx:289-413: largest remaining entry; on the
 * diagonal -> 1x1; else 2x2 if |a11 a22 / |a21| - |a21|| >= |a21| / 2, else 1x1 on
 * the larger diagonal entry; all remaining entries < small -> zero pivots or
 * error) and cholesky_factor (cholesky.cxx:33-187) for the positive-definite case.
 *
 * Work split: NW warps; thread (warp q, lane c) owns rows q*RPT .. q*RPT+RPT-1 of
 * column c of the full symmetric block, which lives in shared memory, double
 * buffered (a pivot reads the old buffer through the permutation that brings
 * the pivot to the front and writes the new one).  k_diag gave every entry its
 * own thread (32 warps): its pivots cost two 1024-thread barriers and a
 * 32-warp-wide update each; here a pivot costs two NW-warp barriers and RPT
 * entries per thread, with the same expression per entry, so the results are
 * identical.  Entry (r,c) and its mirror (c,r) are computed with the same
 * expression, which keeps the block exactly symmetric.
 */
#pragma once
#include <cmath>
#include <cstddef>

#ifdef __CUDACC__
#define DB_FN __device__ __forceinline__
#else
#define DB_FN inline
#endif

namespace b200 {

constexpr int DB_BS = 32;

constexpr int DB_OK = 0;
constexpr int DB_SINGULAR = -5;      // SPRAL_SSIDS_ERROR_SINGULAR
constexpr int DB_NOT_POS_DEF = -6;   // SPRAL_SSIDS_ERROR_NOT_POS_DEF

template <int NW>
struct DiagShared {
   double A[2][DB_BS][DB_BS + 1];
   double LDm[2][DB_BS][DB_BS + 1];
   double dinv[2 * DB_BS];
   double pmax[NW][DB_BS];    // per warp: largest |entry| of its rows of column c (remaining lower triangle)
   int prow[NW][DB_BS];       // ... and the smallest row that attains it
   int lperm[DB_BS];          // position j of the permuted block holds old position lperm[j]
   int piv_i[4];
   double piv_d[4];
};

/* Ctx: tid(), sync() (all NW*32 threads), shfl_xor(double|int, offset) within a warp.
 * Ld: the block in the front (column-major, lower triangle valid), leading dimension ldl.
 * a0_out (may be null): receives the unfactorised block, full symmetric, ld = BS.
 * On return (DB_OK) buffer `cur` of sh.A holds L11 (strict lower part; the unit
 * diagonal is implied), sh.LDm[cur] holds L11*D, sh.dinv / sh.lperm are final and
 * columns >= zfrom are tentative zero pivots.  Every thread returns the same code. */
template <int NW, bool POSDEF, class Ctx>
DB_FN int diag_block_factor(Ctx& cx, DiagShared<NW>& sh, const double* Ld, size_t ldl, int bs,
      double small, int action, double inf, double* a0_out, int& cur_out, int& zfrom_out) {
   constexpr int BS = DB_BS, RPT = BS / NW;
   const int c = cx.tid() & 31, q = cx.tid() >> 5;
   const int rlo = q * RPT;

   /* load the lower triangle, mirror it */
   #pragma unroll
   for (int i = 0; i < RPT; ++i) {
      const int r = rlo + i;
      double v = 0.0;
      if (r < bs && c < bs && r >= c) v = Ld[r + (size_t)c * ldl];
      sh.A[0][r][c] = v;
      sh.LDm[0][r][c] = 0.0; sh.LDm[1][r][c] = 0.0;
   }
   if (q == 0) { sh.lperm[c] = c; sh.dinv[2 * c] = 0.0; sh.dinv[2 * c + 1] = 0.0; }
   cx.sync();
   #pragma unroll
   for (int i = 0; i < RPT; ++i) {
      const int r = rlo + i;
      if (r < c) sh.A[0][r][c] = sh.A[0][c][r];
   }
   cx.sync();
   int cur = 0;
   zfrom_out = BS;

   if (POSDEF) {
      for (int p = 0; p < bs; ++p) {
         const double d = sh.A[cur][p][p];
         if (!(d > 0.0)) { cur_out = cur; return DB_NOT_POS_DEF; }
         const double lpp = sqrt(d);
         #pragma unroll
         for (int i = 0; i < RPT; ++i) {
            const int r = rlo + i;
            const int R = r > c ? r : c, C = r > c ? c : r;
            double v = sh.A[cur][R][C];
            if (C == p) v = (R == p) ? lpp : v / lpp;
            else if (C > p) v -= (sh.A[cur][R][p] / lpp) * (sh.A[cur][C][p] / lpp);
            sh.A[cur ^ 1][r][c] = v;
         }
         if (cx.tid() == 0) sh.dinv[p] = 1.0 / lpp;
         cx.sync();
         cur ^= 1;
      }
      cur_out = cur;
      return DB_OK;
   }

   if (a0_out) {
      #pragma unroll
      for (int i = 0; i < RPT; ++i) a0_out[(rlo + i) + c * BS] = sh.A[0][rlo + i][c];
   }

   /* largest remaining entry of this thread's rows of column c (ties: smallest row);
    * the search for the next pivot is folded into the update of the current one */
   double vn[RPT];
   #pragma unroll
   for (int i = 0; i < RPT; ++i) vn[i] = sh.A[0][rlo + i][c];
   int p = 0;
   auto partial_max = [&](int pp) {
      double v = -1.0;
      int rr = rlo;
      #pragma unroll
      for (int i = 0; i < RPT; ++i) {
         const int r = rlo + i;
         const double x = (r >= c && c >= pp && r < bs) ? fabs(vn[i]) : -1.0;
         if (x > v) { v = x; rr = r; }
      }
      sh.pmax[q][c] = v; sh.prow[q][c] = rr;
   };
   partial_max(0);

   while (p < bs) {
      cx.sync();
      /* warp 0 takes the decision (lane r looks at column r); everybody else waits */
      if (q == 0) {
         double best = sh.pmax[0][c];
         int row = sh.prow[0][c];
         #pragma unroll
         for (int w = 1; w < NW; ++w) {
            const double x = sh.pmax[w][c];
            if (x > best) { best = x; row = sh.prow[w][c]; }
         }
         int bidx = c * BS + row;
         #pragma unroll
         for (int off = 16; off > 0; off >>= 1) {
            const double ob = cx.shfl_xor(best, off);
            const int oi = cx.shfl_xor(bidx, off);
            if (ob > best || (ob == best && oi < bidx)) { best = ob; bidx = oi; }
         }
         if (c == 0) {
            int m_ = bidx / BS, t_ = bidx % BS;      // column m <= row t
            int ps = 1;
            double e11 = 0, e21 = 0, e22 = 0;
            if (!(best >= small)) ps = 0;
            else if (t_ == m_) e11 = 1.0 / sh.A[cur][t_][t_];
            else {
               const double a11 = sh.A[cur][m_][m_], a22 = sh.A[cur][t_][t_], a21 = sh.A[cur][t_][m_];
               const double detscale = 1.0 / fabs(a21);
               const double detpiv = (a11 * detscale) * a22 - fabs(a21);
               if (fabs(detpiv) >= fabs(a21) / 2) {
                  ps = 2;
                  e11 = (a22 * detscale) / detpiv;
                  e22 = (a11 * detscale) / detpiv;
                  e21 = (-a21 * detscale) / detpiv;
               } else {
                  if (fabs(a11) > fabs(a22)) t_ = m_;    // a11 as 1x1, else a22 (row/col t)
                  e11 = 1.0 / sh.A[cur][t_][t_];
               }
            }
            sh.piv_i[0] = ps; sh.piv_i[1] = t_; sh.piv_i[2] = m_;
            sh.piv_d[0] = e11; sh.piv_d[1] = e21; sh.piv_d[2] = e22;
         }
      }
      cx.sync();
      const int pivsiz = sh.piv_i[0], t = sh.piv_i[1], m = sh.piv_i[2];
      const double d11 = sh.piv_d[0], d21 = sh.piv_d[1], d22 = sh.piv_d[2];

      if (pivsiz == 0) {
         /* everything left is (numerically) zero: block_ldlt.hxx:303-317 */
         if (!action) { cur_out = cur; return DB_SINGULAR; }
         zfrom_out = p;
         #pragma unroll
         for (int i = 0; i < RPT; ++i) {
            const int r = rlo + i;
            const int R = r > c ? r : c, C = r > c ? c : r;
            if (C >= p) { sh.A[cur][r][c] = (R == C) ? 1.0 : 0.0; sh.LDm[cur][r][c] = 0.0; }
         }
         cx.sync();
         break;
      }

      const double (*Ao)[BS + 1] = sh.A[cur];
      const double (*Lo)[BS + 1] = sh.LDm[cur];
      if (pivsiz == 1) {
         #pragma unroll
         for (int i = 0; i < RPT; ++i) {
            const int r = rlo + i;
            const int R = r > c ? r : c, C = r > c ? c : r;
            /* new position x holds old position pi(x): swap p <-> t */
            const int oR = (R == p) ? t : (R == t ? p : R);
            const int oC = (C == p) ? t : (C == t ? p : C);
            double vnew, ldnew;
            if (C < p) { vnew = Ao[oR][C]; ldnew = Lo[oR][C]; }
            else if (C == p) {
               const double wr = Ao[oR][t];
               vnew = (R == p) ? 1.0 : wr * d11;
               ldnew = (R == p) ? 0.0 : wr;
            } else {
               vnew = Ao[oR][oC] - (Ao[oR][t] * d11) * Ao[oC][t];
               ldnew = 0.0;
            }
            vn[i] = vnew;
            sh.A[cur ^ 1][r][c] = vnew;
            /* LD is only meaningful strictly below the diagonal of eliminated columns */
            sh.LDm[cur ^ 1][r][c] = (r > c) ? ldnew : 0.0;
         }
         if (cx.tid() == 0) {
            sh.dinv[2 * p] = d11; sh.dinv[2 * p + 1] = 0.0;
            const int x = sh.lperm[p]; sh.lperm[p] = sh.lperm[t]; sh.lperm[t] = x;
         }
      } else {
         #pragma unroll
         for (int i = 0; i < RPT; ++i) {
            const int r = rlo + i;
            const int R = r > c ? r : c, C = r > c ? c : r;
            /* swap p <-> m, then p+1 <-> t */
            auto pi1 = [&](int y) { return y == p ? m : (y == m ? p : y); };
            auto pi = [&](int x) { return x == p + 1 ? pi1(t) : (x == t ? pi1(p + 1) : pi1(x)); };
            const int oR = pi(R), oC = pi(C);
            double vnew, ldnew;
            if (C < p) { vnew = Ao[oR][C]; ldnew = Lo[oR][C]; }
            else if (C <= p + 1) {
               if (R <= p + 1) {              // the 2x2 diagonal block of L is the identity
                  vnew = (R == C) ? 1.0 : 0.0; ldnew = 0.0;
               } else {
                  const double w1 = Ao[oR][m], w2 = Ao[oR][t];
                  if (C == p) { vnew = d11 * w1 + d21 * w2; ldnew = w1; }
                  else        { vnew = d21 * w1 + d22 * w2; ldnew = w2; }
               }
            } else {
               const double w1 = Ao[oR][m], w2 = Ao[oR][t];
               const double l1 = d11 * w1 + d21 * w2, l2 = d21 * w1 + d22 * w2;
               vnew = Ao[oR][oC] - (Ao[oC][m] * l1 + Ao[oC][t] * l2);
               ldnew = 0.0;
            }
            vn[i] = vnew;
            sh.A[cur ^ 1][r][c] = vnew;
            sh.LDm[cur ^ 1][r][c] = (r > c) ? ldnew : 0.0;
         }
         if (cx.tid() == 0) {
            sh.dinv[2 * p] = d11; sh.dinv[2 * p + 1] = d21;
            sh.dinv[2 * p + 2] = inf; sh.dinv[2 * p + 3] = d22;
            int x = sh.lperm[p]; sh.lperm[p] = sh.lperm[m]; sh.lperm[m] = x;
            x = sh.lperm[p + 1]; sh.lperm[p + 1] = sh.lperm[t]; sh.lperm[t] = x;
         }
      }
      cur ^= 1;
      p += pivsiz;
      if (p < bs) partial_max(p);      // pmax was consumed before the previous barrier
   }
   cx.sync();
   cur_out = cur;
   return DB_OK;
}

} // namespace b200
