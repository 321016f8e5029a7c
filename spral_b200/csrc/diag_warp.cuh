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
   /* all loads before the stores: the compiler cannot tell that the addresses differ and would otherwise wait
    * for every store before the next load */
   const bool el = k < p;                                    // mirrored L*D of the eliminated column k
   const double x = S[i1 * DW_LD + j1], y = S[i2 * DW_LD + j2];
   const double u = el ? S[k * DW_LD + a] : 0.0, v = el ? S[k * DW_LD + b] : 0.0;
   S[i1 * DW_LD + j1] = y; S[i2 * DW_LD + j2] = x;
   if (el) { S[k * DW_LD + a] = v; S[k * DW_LD + b] = u; }
}

/* Running maximum of |v| over the rows a lane visits.  Branch-free (a divergent branch per row costs more than
 * the arithmetic): the key of a candidate is the bit pattern of |v| -- monotonic for non-negative doubles, so the
 * comparison runs on the integer pipe -- or -1 for a row that does not take part.  Four independent accumulators
 * (row mod 4) keep the compare / select chain short; merged with ties going to the smallest row. */
struct DwMax {
   long long k[4]; int r[4];
   __device__ __forceinline__ void reset(int lane) {
      #pragma unroll
      for (int i = 0; i < 4; ++i) { k[i] = -1; r[i] = lane; }
   }
   __device__ __forceinline__ void add(int slot, bool on, double v, int row) {
      const long long key = on ? __double_as_longlong(fabs(v)) : -1LL;
      const bool better = key > k[slot];
      k[slot] = better ? key : k[slot];
      r[slot] = better ? row : r[slot];
   }
   __device__ __forceinline__ void result(double& best, int& row) const {
      long long bk = k[0]; row = r[0];
      #pragma unroll
      for (int i = 1; i < 4; ++i) {
         const bool better = k[i] > bk || (k[i] == bk && r[i] < row);
         bk = better ? k[i] : bk; row = better ? r[i] : row;
      }
      best = bk < 0 ? -1.0 : __longlong_as_double(bk);
   }
};

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

/* Trailing update of column `lane` with one (NP = 1) or two (NP = 2) pivot columns at p, rows R0 .. 31, and the
 * column's new maximum.  Straight-line code: every load is independent of the stores before it in program order only
 * by its position, so the loads of a row batch are issued ahead of the stores (a load behind a store to the same
 * array would wait for it); R0 (0, 8, 16, 24) skips the rows above the pivot with ONE uniform branch in the caller. */
template <int R0, int NP>
__device__ __forceinline__ void dw_update(double* S, int p, int lane, int bs, bool on, double w1, double w2, DwMax& mx) {
   constexpr int NR = 32 - R0;
   double l1[NR], l2[NR], cc[NR];
   #pragma unroll
   for (int i = 0; i < NR; ++i) {
      const int r = R0 + i;
      l1[i] = S[r * DW_LD + p];
      if (NP == 2) l2[i] = S[r * DW_LD + p + 1];
      cc[i] = S[r * DW_LD + lane];
   }
   #pragma unroll
   for (int i = 0; i < NR; ++i) {
      const int r = R0 + i;
      const bool act = on && r > p + NP - 1 && r >= lane && r < bs;
      const double vnew = (NP == 2) ? cc[i] - (w1 * l1[i] + w2 * l2[i]) : cc[i] - l1[i] * w1;
      if (act) S[r * DW_LD + lane] = vnew;
      mx.add(r & 3, act, vnew, r);
   }
}
template <int NP>
__device__ __forceinline__ void dw_update_from(double* S, int p, int lane, int bs, bool on, double w1, double w2, DwMax& mx) {
   const int first = p + NP;            // first row that changes
   if (first >= 24) dw_update<24, NP>(S, p, lane, bs, on, w1, w2, mx);
   else if (first >= 16) dw_update<16, NP>(S, p, lane, bs, on, w1, w2, mx);
   else if (first >= 8) dw_update<8, NP>(S, p, lane, bs, on, w1, w2, mx);
   else dw_update<0, NP>(S, p, lane, bs, on, w1, w2, mx);
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
   int lp = lane;                       // lperm[lane], kept in a register: a swap is two shuffles off the critical path
   dinv[2 * lane] = 0.0; dinv[2 * lane + 1] = 0.0;
   double rdiag = 0.0;                  // 1 / S(lane, lane), computed while the trailing update runs
   if (lane < bs) rdiag = 1.0 / S[lane * DW_LD + lane];
   /* largest entry of column `lane` of the remaining lower triangle (ties: smallest row) */
   double cmaxv = -1.0;
   int cmaxr = lane;
   {
      DwMax mx; mx.reset(lane);
      #pragma unroll
      for (int r = 0; r < 32; ++r) {
         const double av = fabs(S[r * DW_LD + lane]);
         mx.add(r & 3, r >= lane && r < bs && lane < bs, av, r);
      }
      mx.result(cmaxv, cmaxr);
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
      else if (t == m) e11 = __shfl_sync(DW_FULL, rdiag, t);      // 1 / S(t, t)
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
            e11 = __shfl_sync(DW_FULL, rdiag, t);
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
      if (ps == 1) {
         if (t != p) {
            dw_swap(S, p, t, p, lane);
            { const int la = __shfl_sync(DW_FULL, lp, p), lb = __shfl_sync(DW_FULL, lp, t); lp = (lane == p) ? lb : (lane == t ? la : lp);
              const double ra = __shfl_sync(DW_FULL, rdiag, p), rb = __shfl_sync(DW_FULL, rdiag, t); rdiag = (lane == p) ? rb : (lane == t ? ra : rdiag); }
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
         {
            /* rank-1 update of column `lane`.  The reciprocal of the lane's new diagonal entry -- the next 1x1
             * pivot's D^-1 -- is computed in the same straight-line code, so its latency hides behind the update
             * instead of sitting between the search and the swap of the next pivot. */
            const bool on = lane > p && lane < bs;
            const double dnew = S[lane * DW_LD + lane] - (w * e11) * w;       // == the value the update stores at (lane, lane)
            if (on) rdiag = 1.0 / dnew;
            DwMax mx; mx.reset(lane);
            dw_update_from<1>(S, p, lane, bs, on, w, 0.0, mx);
            mx.result(cmaxv, cmaxr);
         }
      } else {
         /* swap p <-> m, then p+1 <-> t */
         if (m != p) {
            dw_swap(S, p, m, p, lane);
            { const int la = __shfl_sync(DW_FULL, lp, p), lb = __shfl_sync(DW_FULL, lp, m); lp = (lane == p) ? lb : (lane == m ? la : lp);
              const double ra = __shfl_sync(DW_FULL, rdiag, p), rb = __shfl_sync(DW_FULL, rdiag, m); rdiag = (lane == p) ? rb : (lane == m ? ra : rdiag); }
            __syncwarp();
         }
         if (t != p + 1) {
            dw_swap(S, p + 1, t, p, lane);
            { const int la = __shfl_sync(DW_FULL, lp, p + 1), lb = __shfl_sync(DW_FULL, lp, t); lp = (lane == p + 1) ? lb : (lane == t ? la : lp);
              const double ra = __shfl_sync(DW_FULL, rdiag, p + 1), rb = __shfl_sync(DW_FULL, rdiag, t); rdiag = (lane == p + 1) ? rb : (lane == t ? ra : rdiag); }
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
         {
            const bool on = lane > p + 1 && lane < bs;
            const double dnew = S[lane * DW_LD + lane] - (w1 * (e11 * w1 + e21 * w2) + w2 * (e21 * w1 + e22 * w2));
            if (on) rdiag = 1.0 / dnew;
            DwMax mx; mx.reset(lane);
            dw_update_from<2>(S, p, lane, bs, on, w1, w2, mx);
            mx.result(cmaxv, cmaxr);
         }
      }
      p += ps;
   }
   lperm[lane] = lp;
   __syncwarp();
   return DW_OK;
}

/* Positive definite: A = L L^T (cholesky_factor, cholesky.cxx:33-187).  On return the lower triangle of S
 * holds L (diagonal included), dinv[j] = 1 / l_jj. */
__device__ __forceinline__ int diag_warp_chol(double* S, double* dinv, int bs) {
   const int lane = threadIdx.x & 31;
   /* every lane keeps its own diagonal entry, its square root and the reciprocal of the root up to date inside the
    * straight-line code of the trailing update, so that no square root / division sits between two columns */
   double dg = (lane < bs) ? S[lane * DW_LD + lane] : 1.0;
   double sq = sqrt(dg), rq = 1.0 / sq;
   for (int p = 0; p < bs; ++p) {
      const double d = __shfl_sync(DW_FULL, dg, p);
      if (!(d > 0.0)) return DW_NOT_POS_DEF;
      const double lpp = __shfl_sync(DW_FULL, sq, p), rl = __shfl_sync(DW_FULL, rq, p);
      double w = 0.0;
      if (lane > p && lane < bs) { w = S[lane * DW_LD + p] * rl; S[lane * DW_LD + p] = w; }     // v / l_pp as v * (1 / l_pp)
      if (lane == p) { S[p * DW_LD + p] = lpp; dinv[p] = rl; }
      __syncwarp();
      {
         const bool on = lane > p && lane < bs;
         if (on) { dg = dg - w * w; sq = sqrt(dg); rq = 1.0 / sq; }
         double lc[32], cc[32];
         #pragma unroll
         for (int r = 0; r < 32; ++r) { lc[r] = S[r * DW_LD + p]; cc[r] = S[r * DW_LD + lane]; }
         #pragma unroll
         for (int r = 0; r < 32; ++r)
            if (on && r > p && r >= lane && r < bs) S[r * DW_LD + lane] = cc[r] - lc[r] * w;
      }
      __syncwarp();
   }
   return DW_OK;
}

} // namespace b200
