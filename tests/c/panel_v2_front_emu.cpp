/* TEST INFRASTRUCTURE (CPU only): a whole dense front factorised with the speculative
 * panel kernels of spral_b200/csrc/panel_v2.h on host threads (tests/c/emu.h), in the
 * order the host scheduler issues them (subtree.cu: factor_fronts, v2 branch):
 *   panel [0,256):   chain/tiles/commit at p = 0, UPD_SEG, chain/tiles/commit at p = 128,
 *                    outer update of the columns right of the panel
 *   panel [256,384): chain/tiles/commit at p = 256 (row permutation of 256 earlier columns)
 * The DMMA updates (gemm_dmma.cu: UPD_SEG, UPD_OUTER) are replaced by plain loops over the
 * same regions.  Checks P A P^T = L D L^T over the eliminated columns with P taken from
 * the front's perm array, and the roll-back of a segment whose rows below fail. */
#include <cmath>
#include <cstdio>
#include <limits>
#include <memory>
#include <random>
#include <vector>

#include "../../spral_b200/csrc/panel_v2.h"
#include "emu.h"

using namespace b200;
static const double INF = std::numeric_limits<double>::infinity();

struct FrontEmu {
   int m, n, ldl;
   std::vector<double> A, L, LD, BK, D;
   std::vector<int> perm;
   std::unique_ptr<SegWS> ws{new SegWS};
   int done = 0;
};

/* one speculative segment at p = done: returns 1 accepted, 0 chain gave up, -1 rolled back */
template <bool POSDEF>
static int segment(FrontEmu& f) {
   const int p = f.done, m = f.m, ldl = f.ldl;
   auto csh = std::make_unique<ChainShared>();
   int ok = 0;
   emu::run_cta(CNT, [&](emu::Ctx& cx) {
      int r = chain_segment<POSDEF>(cx, *csh, f.L.data() + p + (size_t)p * ldl, (size_t)ldl, 0.01, 1e-20, INF, f.ws.get());
      if (cx.tid() == 0) ok = r;
   });
   if (!ok) return 0;
   int seg_fail = 0;
   auto tsh = std::make_unique<TileShared>();
   for (int r0 = 0; r0 < m; r0 += RT) {
      if (r0 + RT <= p + CW) continue;
      emu::run_cta(RT, [&](emu::Ctx& cx) {
         panel_tile<POSDEF>(cx, *tsh, f.L.data() + (size_t)p * ldl, (POSDEF ? f.L.data() : f.LD.data()) + (size_t)p * ldl, POSDEF ? nullptr : f.BK.data(), (size_t)ldl, m, r0,
                    p, 0.01, INF, f.ws.get(), &seg_fail);
      });
   }
   CommitShared cs;
   for (int r0 = 0; r0 < m; r0 += RT)
      emu::run_cta(RT, [&](emu::Ctx& cx) {
         seg_commit<POSDEF>(cx, cs, f.L.data(), f.D.data(), f.perm.data(), f.BK.data(), (size_t)ldl, m, p, r0, r0 / RT == p / RT,
                    seg_fail, f.ws.get());
      });
   if (seg_fail) return -1;
   f.done += CW;                       // account_segment
   return 1;
}

/* C(r, c) -= sum_{k in [k0, k1)} L(r, k) LD(c, k) for c in [c_lo, c_hi), r >= c */
static void update(FrontEmu& f, int k0, int k1, int c_lo, int c_hi, bool posdef = false) {
   const std::vector<double>& B = posdef ? f.L : f.LD;        // f->LD == f->L for Cholesky
   for (int c = c_lo; c < c_hi; ++c)
      for (int r = c; r < f.m; ++r) {
         double s = 0;
         for (int k = k0; k < k1; ++k) s += f.L[r + (size_t)k * f.ldl] * B[c + (size_t)k * f.ldl];
         f.L[r + (size_t)c * f.ldl] -= s;
      }
}

static double check(const FrontEmu& f, int nelim) {
   const int m = f.m, ldl = f.ldl;
   std::vector<double> Dm((size_t)nelim * nelim, 0.0);
   for (int j = 0; j < nelim;) {
      if (j + 1 < nelim && f.D[2 * j + 2] == INF) {
         double e11 = f.D[2 * j], e21 = f.D[2 * j + 1], e22 = f.D[2 * j + 3], det = e11 * e22 - e21 * e21;
         Dm[j + (size_t)j * nelim] = e22 / det; Dm[j + 1 + (size_t)(j + 1) * nelim] = e11 / det;
         Dm[j + 1 + (size_t)j * nelim] = Dm[j + (size_t)(j + 1) * nelim] = -e21 / det;
         j += 2;
      } else { Dm[j + (size_t)j * nelim] = 1.0 / f.D[2 * j]; ++j; }
   }
   auto a = [&](int r, int c) { return r >= c ? f.A[r + (size_t)c * ldl] : f.A[c + (size_t)r * ldl]; };
   auto orig = [&](int i) { return i < f.n ? f.perm[i] - 1 : i; };
   std::vector<double> W((size_t)m * nelim, 0.0);            // L D
   for (int i = 0; i < m; ++i)
      for (int k = 0; k < nelim; ++k) {
         double s = 0;
         for (int q = (k > 0 ? k - 1 : 0); q <= k + 1 && q < nelim; ++q)
            if (q <= i && Dm[q + (size_t)k * nelim] != 0.0) s += (q == i ? 1.0 : f.L[i + (size_t)q * ldl]) * Dm[q + (size_t)k * nelim];
         W[i + (size_t)k * m] = s;
      }
   double err = 0;
   for (int c = 0; c < nelim; ++c)
      for (int i = c; i < m; ++i) {
         double s = 0;
         for (int k = 0; k <= c; ++k) s += W[i + (size_t)k * m] * (k == c ? 1.0 : f.L[c + (size_t)k * ldl]);
         err = std::max(err, std::fabs(s - a(orig(i), orig(c))));
      }
   return err;
}

static FrontEmu make_front(int m, int n, std::mt19937_64& rng) {
   std::uniform_real_distribution<double> U(-1.0, 1.0);
   FrontEmu f;
   f.m = m; f.n = n; f.ldl = (m + 1) / 2 * 2;
   f.A.assign((size_t)f.ldl * n, std::nan(""));
   for (int c = 0; c < n; ++c)
      for (int r = c; r < m; ++r) f.A[r + (size_t)c * f.ldl] = (r == c) ? (c % 2 ? -1.0 : 1.0) * (4.0 + U(rng)) : U(rng);
   f.L = f.A;
   f.LD.assign((size_t)f.ldl * n, std::nan(""));
   f.BK.assign((size_t)f.ldl * CW, std::nan(""));
   f.D.assign(2 * n, 0.0);
   f.perm.resize(n);
   for (int i = 0; i < n; ++i) f.perm[i] = i + 1;
   return f;
}

int main() {
   std::mt19937_64 rng(4242);
   int failures = 0;
   for (int m : {384, 500}) {
      FrontEmu f = make_front(m, 384, rng);
      int a = segment<false>(f);                       // p = 0
      update(f, 0, 128, 128, 256);              // UPD_SEG
      int b = segment<false>(f);                       // p = 128
      update(f, 0, 256, 256, f.n);              // UPD_OUTER
      int c = segment<false>(f);                       // p = 256
      double err = check(f, 384);
      bool ok = a == 1 && b == 1 && c == 1 && f.done == 384 && err < 1e-10;
      printf("m=%d n=384: segments %d %d %d, |P A P' - L D L'| = %.2e %s\n", m, a, b, c, err, ok ? "ok" : "FAIL");
      failures += !ok;
   }
   {  /* the rows below the second segment are huge: it must roll back and leave everything as it was */
      FrontEmu f = make_front(500, 384, rng);
      for (int c = 128; c < 256; ++c)
         for (int r = 256; r < f.m; ++r) { f.A[r + (size_t)c * f.ldl] *= 1e9; f.L[r + (size_t)c * f.ldl] *= 1e9; }
      int a = segment<false>(f);
      update(f, 0, 128, 128, 256);
      std::vector<double> Lb = f.L, Db = f.D;
      std::vector<int> pb = f.perm;
      int b = segment<false>(f);
      bool same = (f.L.size() == Lb.size());
      for (size_t e = 0; e < Lb.size() && same; ++e) same = (f.L[e] == Lb[e]) || (std::isnan(f.L[e]) && std::isnan(Lb[e]));
      same = same && f.D == Db && f.perm == pb && f.done == 128;
      printf("roll-back: segments %d %d, front unchanged = %d %s\n", a, b, (int)same, (a == 1 && b == -1 && same) ? "ok" : "FAIL");
      failures += !(a == 1 && b == -1 && same);
   }
   {  /* a segment that does not start on a tile boundary: 37 columns eliminated by plain 1x1 pivots first
         (what the step-by-step path leaves behind after a panel with delays), then segments at p = 37, 165 */
      const int lead = 37, m = 430, n = lead + 256;
      FrontEmu f = make_front(m, n, rng);
      for (int j = 0; j < lead; ++j) {
         const double d = f.L[j + (size_t)j * f.ldl];
         f.D[2 * j] = 1.0 / d; f.D[2 * j + 1] = 0.0;
         for (int i = j + 1; i < m; ++i) {
            f.LD[i + (size_t)j * f.ldl] = f.L[i + (size_t)j * f.ldl];
            f.L[i + (size_t)j * f.ldl] /= d;
         }
         f.L[j + (size_t)j * f.ldl] = 1.0;
         update(f, j, j + 1, j + 1, n);
      }
      f.done = lead;
      int a = segment<false>(f);                // p = 37
      update(f, lead, lead + 128, lead + 128, n);
      int b = segment<false>(f);                // p = 165
      double err = check(f, n);
      bool ok = a == 1 && b == 1 && f.done == n && err < 1e-10;
      printf("unaligned segments (p = 37, 165), m=%d: segments %d %d, |P A P' - L D L'| = %.2e %s\n", m, a, b, err,
             ok ? "ok" : "FAIL");
      failures += !ok;
   }
   for (int m : {384, 470}) {
      /* positive definite: Cholesky segments, no permutation, no D, no backup */
      FrontEmu f = make_front(m, 384, rng);
      for (int c = 0; c < f.n; ++c) {                     // make it strictly diagonally dominant
         f.A[c + (size_t)c * f.ldl] = 40.0 + c % 7;
         f.L[c + (size_t)c * f.ldl] = f.A[c + (size_t)c * f.ldl];
      }
      int a = segment<true>(f);
      update(f, 0, 128, 128, 256, true);
      int b = segment<true>(f);
      update(f, 0, 256, 256, f.n, true);
      int c = segment<true>(f);
      double err = 0;
      for (int cc = 0; cc < f.n; ++cc)
         for (int i = cc; i < m; ++i) {
            double sum = 0;
            for (int k = 0; k <= cc; ++k) sum += f.L[i + (size_t)k * f.ldl] * f.L[cc + (size_t)k * f.ldl];
            err = std::max(err, std::fabs(sum - f.A[i + (size_t)cc * f.ldl]));
         }
      bool ok = a == 1 && b == 1 && c == 1 && f.done == 384 && err < 1e-10;
      printf("posdef m=%d n=384: segments %d %d %d, |A - L L'| = %.2e %s\n", m, a, b, c, err, ok ? "ok" : "FAIL");
      failures += !ok;
   }
   {  /* not positive definite: the chain gives up and changes nothing */
      FrontEmu f = make_front(300, 256, rng);
      std::vector<double> Lb = f.L;
      int a = segment<true>(f);                           // make_front's diagonal alternates in sign
      bool same = true;
      for (size_t e = 0; e < Lb.size() && same; ++e) same = (f.L[e] == Lb[e]) || (std::isnan(f.L[e]) && std::isnan(Lb[e]));
      printf("not positive definite: segment %d, front unchanged = %d %s\n", a, (int)same, (a == 0 && same) ? "ok" : "FAIL");
      failures += !(a == 0 && same);
   }
   printf("panel_v2_front_emu: %d failures\n", failures);
   return failures ? 1 : 0;
}
