/* Factorisation of one 32 x 32 diagonal block by ONE warp, without a block-wide barrier.
 *
 * Semantics: block_ldlt of the reference CPU engine (src/ssids/cpu/kernels/block_ldlt.hxx:289-413:
 * largest remaining entry; on the diagonal -> 1x1; else 2x2 if |a11 a22 / |a21| - |a21|| >= |a21| / 2,
 * else 1x1 on the larger diagonal entry; all remaining entries < small -> zero pivots or error) and
 * cholesky_factor (cholesky.cxx:33-187) for the positive-definite case -- the same rules, the same
 * expressions per entry and the same tie-breaking (first entry in column-major order) as k_diag.
 *
 * Why one warp: the pivots of a block form a serial chain (search, decision, swap, rank-1 / rank-2
 * update).  With 32 or 4 warps every link of the chain pays two CTA barriers and a round trip through
 * one deciding thread (k_diag: ~1.6 us per pivot, measured); a single warp needs only __syncwarp, takes
 * the decision redundantly in every lane and finds the largest entry with three warp reductions.
 *
 * Storage: S(r, c) = S[r * 33 + c].  Lower triangle (r >= c): the not yet eliminated part of the block,
 * and L in the eliminated columns.  Upper triangle: (L D)(r, c) of an eliminated column c is mirrored
 * to S(c, r).  Lane c owns column c during the updates (conflict-free: consecutive lanes, consecutive
 * words) and row `lane` when a pivot column is scaled (stride 33: conflict-free as well).
 */
#pragma once
#include <cmath>
#include <cuda_runtime.h>

namespace b200 {

constexpr int DW_LD = 33;
constexpr int DW_OK = 0;
constexpr int DW_SINGULAR = -5;      // SPRAL_SSIDS_ERROR_SINGULAR
constexpr int DW_NOT_POS_DEF = -6;   // SPRAL_SSIDS_ERROR_NOT_POS_DEF
constexpr unsigned DW_FULL = 0xffffffffu;

/* Symmetric swap of positions a < b (a >= p) of the block in lower storage: rows a, b of the eliminated
 * columns k < p (L and the mirrored L*D), and the trailing symmetric matrix.  One lane per index k; the
 * lanes touch disjoint entries.  The caller synchronises the warp before and after. */
__device__ __forceinline__ void dw_swap(double* S, int a, int b, int p, int lane) {
   const int k = lane;
   int i1, j1, i2, j2;
   if (k < a)       { i1 = a; j1 = k; i2 = b; j2 = k; }      // rows a and b left of column a (L for k < p)
   else if (k == a) { i1 = a; j1 = a; i2 = b; j2 = b; }      // the two diagonal entries
   else if (k < b)  { i1 = k; j1 = a; i2 = b; j2 = k; }      // column a below the diagonal <-> row b
   else if (k == b) { i1 = b; j1 = a; i2 = b; j2 = a; }      // S(b, a) stays
   else             { i1 = k; j1 = a; i2 = k; j2 = b; }      // columns a and b below row b
   const double x = S[i1 * DW_LD + j1], y = S[i2 * DW_LD + j2];
   S[i1 * DW_LD + j1] = y; S[i2 * DW_LD + j2] = x;
   if (k < p) {                                              // mirrored L*D of the eliminated column k
      const double u = S[k * DW_LD + a], v = S[k * DW_LD + b];
      S[k * DW_LD + a] = v; S[k * DW_LD + b] = u;
   }
}

/* Largest |entry| over the lanes with valid = true (ties: smallest lane): returns the winning lane and
 * the value.  v >= 0 (or -1 / NaN, which then win and are rejected by the caller's `best >= small`). */
__device__ __forceinline__ int dw_argmax(double v, bool valid, double& best) {
   const unsigned hi = (unsigned)__double2hiint(v), lo = (unsigned)__double2loint(v);
   const unsigned mhi = __reduce_max_sync(DW_FULL, valid ? hi : 0u);
   const bool c1 = valid && hi == mhi;
   const unsigned mlo = __reduce_max_sync(DW_FULL, c1 ? lo : 0u);
   const unsigned win = __ballot_sync(DW_FULL, c1 && lo == mlo);
   best = __hiloint2double((int)mhi, (int)mlo);
   return __ffs((int)win) - 1;
}

/* Indefinite: P A P^T = L D L^T with 1x1 / 2x2 pivots chosen inside the block.  bs <= 32 rows / columns
 * are valid (the rest of S must be zero).  On return (DW_OK): strict lower triangle of S = L (unit
 * diagonal implied), upper triangle = mirrored L*D, dinv = D^-1 in the reference CPU layout
 * (block_ldlt.hxx:375-406), lperm[j] = old position of the column now at j, columns >= zfrom are
 * tentative zero pivots.  Every lane returns the same code.  dinv (64) and lperm (32) live in shared memory. */
__device__ __forceinline__ int diag_warp_ldlt(double* S, double* dinv, int* lperm, int bs, double small, int action,
      double inf, int& zfrom_out) {
   const int lane = threadIdx.x & 31;
   zfrom_out = 32;
   lperm[lane] = lane; dinv[2 * lane] = 0.0; dinv[2 * lane + 1] = 0.0;
   /* largest entry of column `lane` of the remaining lower triangle (ties: smallest row) */
   double cmaxv = -1.0;
   int cmaxr = lane;
   if (lane < bs)
      for (int r = lane; r < bs; ++r) {
         const double av = fabs(S[r * DW_LD + lane]);
         if (av > cmaxv) { cmaxv = av; cmaxr = r; }
      }
   int p = 0;
   while (p < bs) {
      double best;
      const int m = dw_argmax(cmaxv, lane >= p && lane < bs, best);     // column m <= row t
      int t = __shfl_sync(DW_FULL, cmaxr, m);
      __syncwarp();
      int ps = 1;
      double e11 = 0.0, e21 = 0.0, e22 = 0.0;
      if (!(best >= small)) ps = 0;
      else if (t == m) e11 = 1.0 / S[t * DW_LD + t];
      else {
         const double a11 = S[m * DW_LD + m], a22 = S[t * DW_LD + t], a21 = S[t * DW_LD + m];
         const double detscale = 1.0 / fabs(a21);
         const double detpiv = (a11 * detscale) * a22 - fabs(a21);
         if (fabs(detpiv) >= fabs(a21) / 2) {
            ps = 2;
            e11 = (a22 * detscale) / detpiv;
            e22 = (a11 * detscale) / detpiv;
            e21 = (-a21 * detscale) / detpiv;
         } else {
            if (fabs(a11) > fabs(a22)) t = m;      // a11 as 1x1, else a22 (row / column t)
            e11 = 1.0 / S[t * DW_LD + t];
         }
      }
      __syncwarp();                       // every lane has read the pivot entries before anybody swaps them
      if (ps == 0) {
         /* everything left is (numerically) zero: block_ldlt.hxx:303-317 */
         if (!action) return DW_SINGULAR;
         zfrom_out = p;
         if (lane >= p && lane < bs)
            for (int r = lane + 1; r < bs; ++r) { S[r * DW_LD + lane] = 0.0; S[lane * DW_LD + r] = 0.0; }
         break;
      }
      cmaxv = -1.0; cmaxr = lane;
      if (ps == 1) {
         if (t != p) {
            dw_swap(S, p, t, p, lane);
            if (lane == 0) { const int x = lperm[p]; lperm[p] = lperm[t]; lperm[t] = x; }
            __syncwarp();
         }
         double w = 0.0;
         if (lane > p && lane < bs) {
            w = S[lane * DW_LD + p];
            S[lane * DW_LD + p] = w * e11;          // L
            S[p * DW_LD + lane] = w;                // L*D, mirrored
         }
         if (lane == 0) { dinv[2 * p] = e11; dinv[2 * p + 1] = 0.0; }
         __syncwarp();
         if (lane > p && lane < bs) {
            #pragma unroll 4
            for (int r = p + 1; r < bs; ++r) {
               if (r >= lane) {
                  const double vnew = S[r * DW_LD + lane] - S[r * DW_LD + p] * w;
                  S[r * DW_LD + lane] = vnew;
                  const double av = fabs(vnew);
                  if (av > cmaxv) { cmaxv = av; cmaxr = r; }
               }
            }
         }
      } else {
         /* swap p <-> m, then p+1 <-> t */
         if (m != p) {
            dw_swap(S, p, m, p, lane);
            if (lane == 0) { const int x = lperm[p]; lperm[p] = lperm[m]; lperm[m] = x; }
            __syncwarp();
         }
         if (t != p + 1) {
            dw_swap(S, p + 1, t, p, lane);
            if (lane == 0) { const int x = lperm[p + 1]; lperm[p + 1] = lperm[t]; lperm[t] = x; }
            __syncwarp();
         }
         double w1 = 0.0, w2 = 0.0;
         if (lane > p + 1 && lane < bs) {
            w1 = S[lane * DW_LD + p]; w2 = S[lane * DW_LD + p + 1];
            S[lane * DW_LD + p] = e11 * w1 + e21 * w2;
            S[lane * DW_LD + p + 1] = e21 * w1 + e22 * w2;
            S[p * DW_LD + lane] = w1;
            S[(p + 1) * DW_LD + lane] = w2;
         }
         if (lane == 0) {
            S[(p + 1) * DW_LD + p] = 0.0;           // the 2x2 diagonal block of L is the identity
            S[p * DW_LD + p + 1] = 0.0;
            dinv[2 * p] = e11; dinv[2 * p + 1] = e21;
            dinv[2 * p + 2] = inf; dinv[2 * p + 3] = e22;
         }
         __syncwarp();
         if (lane > p + 1 && lane < bs) {
            #pragma unroll 4
            for (int r = p + 2; r < bs; ++r) {
               if (r >= lane) {
                  const double vnew = S[r * DW_LD + lane] - (w1 * S[r * DW_LD + p] + w2 * S[r * DW_LD + p + 1]);
                  S[r * DW_LD + lane] = vnew;
                  const double av = fabs(vnew);
                  if (av > cmaxv) { cmaxv = av; cmaxr = r; }
               }
            }
         }
      }
      p += ps;
   }
   __syncwarp();
   return DW_OK;
}

/* Positive definite: A = L L^T (cholesky_factor, cholesky.cxx:33-187).  On return the lower triangle of S
 * holds L (diagonal included), dinv[j] = 1 / l_jj. */
__device__ __forceinline__ int diag_warp_chol(double* S, double* dinv, int bs) {
   const int lane = threadIdx.x & 31;
   for (int p = 0; p < bs; ++p) {
      const double d = S[p * DW_LD + p];
      if (!(d > 0.0)) return DW_NOT_POS_DEF;
      const double lpp = sqrt(d);
      __syncwarp();                        // every lane has read the diagonal entry
      double w = 0.0;
      if (lane > p && lane < bs) { w = S[lane * DW_LD + p] / lpp; S[lane * DW_LD + p] = w; }
      if (lane == p) { S[p * DW_LD + p] = lpp; dinv[p] = 1.0 / lpp; }
      __syncwarp();
      if (lane > p && lane < bs) {
         #pragma unroll 4
         for (int r = p + 1; r < bs; ++r)
            if (r >= lane) S[r * DW_LD + lane] -= S[r * DW_LD + p] * w;
      }
      __syncwarp();
   }
   return DW_OK;
}

} // namespace b200
