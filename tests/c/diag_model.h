/* TEST INFRASTRUCTURE (CPU only): sequential model of the thread-per-entry diagonal-block
 * kernel k_diag (spral_b200/csrc/factor_kernels.cu): one "thread" per entry, double
 * buffered, the same expression per entry.  Shared by the emulation tests. */
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include "../../spral_b200/csrc/diag_block.h"

namespace diag_model {
using namespace b200;
constexpr int BS = DB_BS;
static const double INF = std::numeric_limits<double>::infinity();

/* what k_diag / k_diag_v2 leave in the front's BlockWS (+ the block itself for posdef) */
struct Published {
   double l11[BS * BS], ld11[BS * BS], a0[BS * BS], dinv[2 * BS];
   int lperm[BS], zfrom, rc;
   double lblock[BS * BS];       // posdef: the block written back to the front
};

/* ------------------------------------------------------------------ */
/* sequential model of k_diag: one "thread" per entry, double buffered */
/* ------------------------------------------------------------------ */
static void model_v1(bool posdef, const double* Ld, int ldl, int bs, double small, int action, Published& out) {
   static double A[2][BS][BS + 1], LDm[2][BS][BS + 1];
   double dinv[2 * BS], cmax[BS];
   int crow[BS], lperm[BS];
   std::memset(&out, 0, sizeof(out));
   out.zfrom = BS; out.rc = 0;
   for (int c = 0; c < BS; ++c)
      for (int r = 0; r < BS; ++r) {
         double v = 0.0;
         if (r < bs && c < bs && r >= c) v = Ld[r + (size_t)c * ldl];
         A[0][r][c] = v; LDm[0][r][c] = 0.0; LDm[1][r][c] = 0.0;
      }
   for (int r = 0; r < BS; ++r) { lperm[r] = r; dinv[2 * r] = 0.0; dinv[2 * r + 1] = 0.0; }
   for (int c = 0; c < BS; ++c) for (int r = 0; r < c; ++r) A[0][r][c] = A[0][c][r];
   int cur = 0;
   if (posdef) {
      for (int p = 0; p < bs; ++p) {
         double d = A[cur][p][p];
         if (!(d > 0.0)) { out.rc = DB_NOT_POS_DEF; return; }
         double lpp = std::sqrt(d);
         for (int c = 0; c < BS; ++c)
            for (int r = 0; r < BS; ++r) {
               const int R = std::max(r, c), C = std::min(r, c);
               double v = A[cur][R][C];
               if (C == p) v = (R == p) ? lpp : v / lpp;
               else if (C > p) v -= (A[cur][R][p] / lpp) * (A[cur][C][p] / lpp);
               A[cur ^ 1][r][c] = v;
            }
         dinv[p] = 1.0 / lpp;
         cur ^= 1;
      }
      for (int c = 0; c < BS; ++c)
         for (int r = 0; r < BS; ++r) {
            double l = (r < bs && c < bs && r >= c) ? A[cur][r][c] : 0.0;
            if (r < bs && c < bs && r >= c) out.lblock[r + c * BS] = l;
            out.l11[r + c * BS] = l;
         }
      for (int r = 0; r < BS; ++r) out.dinv[r] = (r < bs) ? dinv[r] : 0.0;
      return;
   }
   for (int c = 0; c < BS; ++c) for (int r = 0; r < BS; ++r) out.a0[r + c * BS] = A[0][r][c];
   int zfrom = BS, p = 0;
   auto column_max = [&](int pp) {       // warp c reduces over lanes r: max value, smallest row on ties
      for (int c = 0; c < BS; ++c) {
         double v = -1.0; int rr = 0;
         bool first = true;
         for (int r = 0; r < BS; ++r) {
            double x = (r >= c && c >= pp && r < bs) ? std::fabs(A[cur][r][c]) : -1.0;
            if (first) { v = x; rr = r; first = false; }
            else if (x > v || (x == v && r < rr)) { v = x; rr = r; }
         }
         cmax[c] = v; crow[c] = rr;
      }
   };
   column_max(0);
   while (p < bs) {
      double best = cmax[0]; int bidx = 0 * BS + crow[0];
      for (int c = 1; c < BS; ++c) {
         int oi = c * BS + crow[c];
         if (cmax[c] > best || (cmax[c] == best && oi < bidx)) { best = cmax[c]; bidx = oi; }
      }
      int m = bidx / BS, t = bidx % BS, ps = 1;
      double d11 = 0, d21 = 0, d22 = 0;
      if (!(best >= small)) ps = 0;
      else if (t == m) d11 = 1.0 / A[cur][t][t];
      else {
         double a11 = A[cur][m][m], a22 = A[cur][t][t], a21 = A[cur][t][m];
         double detscale = 1.0 / std::fabs(a21);
         double detpiv = (a11 * detscale) * a22 - std::fabs(a21);
         if (std::fabs(detpiv) >= std::fabs(a21) / 2) {
            ps = 2;
            d11 = (a22 * detscale) / detpiv;
            d22 = (a11 * detscale) / detpiv;
            d21 = (-a21 * detscale) / detpiv;
         } else {
            if (std::fabs(a11) > std::fabs(a22)) t = m;
            d11 = 1.0 / A[cur][t][t];
         }
      }
      const int pivsiz = ps;
      if (pivsiz == 0) {
         if (!action) { out.rc = DB_SINGULAR; return; }
         zfrom = p;
         for (int c = 0; c < BS; ++c)
            for (int r = 0; r < BS; ++r) {
               const int R = std::max(r, c), C = std::min(r, c);
               if (C >= p) { A[cur][r][c] = (R == C) ? 1.0 : 0.0; LDm[cur][r][c] = 0.0; }
            }
         break;
      }
      const double (*Ao)[BS + 1] = A[cur];
      const double (*Lo)[BS + 1] = LDm[cur];
      for (int c = 0; c < BS; ++c)
         for (int r = 0; r < BS; ++r) {
            const int R = std::max(r, c), C = std::min(r, c);
            double vnew, ldnew;
            if (pivsiz == 1) {
               auto pi = [&](int x) { return x == p ? t : (x == t ? p : x); };
               const int oR = pi(R), oC = pi(C);
               if (C < p) { vnew = Ao[oR][C]; ldnew = Lo[oR][C]; }
               else if (C == p) {
                  double wr = Ao[oR][t];
                  vnew = (R == p) ? 1.0 : wr * d11;
                  ldnew = (R == p) ? 0.0 : wr;
               } else {
                  vnew = Ao[oR][oC] - (Ao[oR][t] * d11) * Ao[oC][t];
                  ldnew = 0.0;
               }
            } else {
               auto pi1 = [&](int y) { return y == p ? m : (y == m ? p : y); };
               auto pi = [&](int x) { return x == p + 1 ? pi1(t) : (x == t ? pi1(p + 1) : pi1(x)); };
               const int oR = pi(R), oC = pi(C);
               if (C < p) { vnew = Ao[oR][C]; ldnew = Lo[oR][C]; }
               else if (C <= p + 1) {
                  if (R <= p + 1) { vnew = (R == C) ? 1.0 : 0.0; ldnew = 0.0; }
                  else {
                     double w1 = Ao[oR][m], w2 = Ao[oR][t];
                     if (C == p) { vnew = d11 * w1 + d21 * w2; ldnew = w1; }
                     else        { vnew = d21 * w1 + d22 * w2; ldnew = w2; }
                  }
               } else {
                  double w1 = Ao[oR][m], w2 = Ao[oR][t];
                  double l1 = d11 * w1 + d21 * w2, l2 = d21 * w1 + d22 * w2;
                  vnew = Ao[oR][oC] - (Ao[oC][m] * l1 + Ao[oC][t] * l2);
                  ldnew = 0.0;
               }
            }
            A[cur ^ 1][r][c] = vnew;
            LDm[cur ^ 1][r][c] = (r > c) ? ldnew : 0.0;
         }
      if (pivsiz == 1) {
         dinv[2 * p] = d11; dinv[2 * p + 1] = 0.0;
         std::swap(lperm[p], lperm[t]);
      } else {
         dinv[2 * p] = d11; dinv[2 * p + 1] = d21; dinv[2 * p + 2] = INF; dinv[2 * p + 3] = d22;
         std::swap(lperm[p], lperm[m]);
         std::swap(lperm[p + 1], lperm[t]);
      }
      cur ^= 1;
      p += pivsiz;
      if (p < bs) column_max(p);
   }
   for (int c = 0; c < BS; ++c)
      for (int r = 0; r < BS; ++r) {
         double l = 0.0, y = 0.0;
         if (r < bs && c < bs) {
            if (r > c) { l = A[cur][r][c]; y = LDm[cur][r][c]; }
            else if (r == c) l = 1.0;
         }
         out.l11[r + c * BS] = l; out.ld11[r + c * BS] = y;
      }
   for (int r = 0; r < BS; ++r) {
      out.dinv[2 * r] = (r < bs) ? dinv[2 * r] : 0.0;
      out.dinv[2 * r + 1] = (r < bs) ? dinv[2 * r + 1] : 0.0;
      out.lperm[r] = lperm[r];
   }
   out.zfrom = zfrom;
}


} // namespace diag_model
