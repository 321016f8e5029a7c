/* TEST INFRASTRUCTURE (CPU only): the bodies of the wide solve kernels
 * (spral_b200/csrc/solve_wide.h) run on host threads (tests/c/emu.h) and are compared
 * with a plain gather / substitute / scatter sweep over one front, as
 * NumericSubtree::solve_fwd / solve_bwd do it node by node in the reference CPU engine
 * (src/ssids/cpu/NumericSubtree.hxx:286-418).  Cases: nelim a multiple of 256 or not,
 * delayed columns (n > nelim, n > n0), a root front (m == n), a front with one short
 * block, 1 and 4 right-hand sides, indefinite (unit diagonal) and positive definite. */
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <numeric>
#include <random>
#include <vector>

#include "../../spral_b200/csrc/solve_wide.h"
#include "emu.h"

using namespace b200;

struct Case { int m, n, n0, nelim; bool posdef; };

template <int NR, bool POSDEF>
static double run_case(const Case& c, std::mt19937_64& rng) {
   const int m = c.m, n = c.n, n0 = c.n0, nelim = c.nelim, ndin = n - n0, m0 = m - ndin;
   const int ldl = (m + 1) / 2 * 2;
   const int NG = 3 * m + 17;                       // global variables
   std::uniform_real_distribution<double> U(-1.0, 1.0);
   std::vector<double> L((size_t)ldl * n, std::nan(""));
   for (int j = 0; j < n; ++j)
      for (int i = j; i < m; ++i)
         L[i + (size_t)j * ldl] = (i == j) ? (POSDEF ? 1.5 + 0.5 * U(rng) : 1.0) : U(rng) / std::sqrt((double)m);
   std::vector<int> glob(NG);
   std::iota(glob.begin(), glob.end(), 1);
   std::shuffle(glob.begin(), glob.end(), rng);
   std::vector<int> perm(glob.begin(), glob.begin() + n);              // 1-based global index of local column i < n
   std::vector<int> rows(m0, -12345);                                  // entries [n0, m0) are the contribution rows
   for (int i = n; i < m; ++i) rows[n0 + i - n] = glob[i];
   SolveFront f{L.data(), nullptr, perm.data(), rows.data(), ldl, m, n, n0, m0, nelim};
   auto idx = [&](int i) { return (i < n ? perm[i] : rows[n0 + i - n]) - 1; };

   std::vector<double> x0((size_t)NG * NR);
   for (auto& v : x0) v = U(rng);
   const int ntile = (m + RT - 1) / RT;
   const int nblk = (nelim + SWB - 1) / SWB;
   double worst = 0;

   /* ---------------- forward ---------------- */
   {
      std::vector<double> xr = x0;                                      // reference
      for (int k = 0; k < NR; ++k) {
         std::vector<double> xf(m);
         for (int i = 0; i < m; ++i) xf[i] = xr[(size_t)idx(i) * NR + k];
         for (int j = 0; j < nelim; ++j) {
            if (POSDEF) xf[j] /= L[j + (size_t)j * ldl];
            for (int i = j + 1; i < m; ++i) xf[i] -= L[i + (size_t)j * ldl] * xf[j];
         }
         for (int i = 0; i < m; ++i) xr[(size_t)idx(i) * NR + k] = xf[i];
      }
      std::vector<double> x = x0, ywork((size_t)NG * NR, std::nan(""));
      std::vector<double> smT(sw_T_smem_doubles<NR>()), smG(sw_fG_smem_doubles<NR>());
      for (int b = 0; b < nblk + 1; ++b) {                              // one block too many: must be a no-op
         std::fill(smT.begin(), smT.end(), std::nan(""));
         emu::run_cta(SW_TT, [&](emu::Ctx& cx) { fwd_wide_T<NR, NR, POSDEF>(cx, f, b, x.data(), ywork.data(), smT.data()); });
         for (int t = 0; t < ntile + 1; ++t) {
            std::fill(smG.begin(), smG.end(), std::nan(""));
            for (int ch = 0; ch < SW_FSPLIT; ++ch)
               emu::run_cta(SW_GT, [&](emu::Ctx& cx) { fwd_wide_G<NR>(cx, f, t, b, ch, x.data(), ywork.data(), smG.data()); });
         }
      }
      for (int j = 0; j < nelim; ++j)                                   // k_fwd_flush
         for (int k = 0; k < NR; ++k) x[(size_t)(perm[j] - 1) * NR + k] = ywork[(size_t)(perm[j] - 1) * NR + k];
      for (size_t e = 0; e < x.size(); ++e) {
         double d = std::fabs(x[e] - xr[e]);
         if (!(d <= 1e300)) d = 1e300;
         worst = std::max(worst, d);
      }
   }
   /* ---------------- backward ---------------- */
   {
      std::vector<double> xr = x0;
      for (int k = 0; k < NR; ++k)
         for (int j = nelim - 1; j >= 0; --j) {
            double s = xr[(size_t)idx(j) * NR + k];
            for (int i = j + 1; i < m; ++i) s -= L[i + (size_t)j * ldl] * xr[(size_t)idx(i) * NR + k];
            if (POSDEF) s /= L[j + (size_t)j * ldl];
            xr[(size_t)idx(j) * NR + k] = s;
         }
      std::vector<double> x = x0;
      std::vector<double> pbuf((size_t)SWB * NR, 0.0);                  // the front's accumulator: zero when the sweep starts
      std::vector<double> smT(sw_T_smem_doubles<NR>()), smG(sw_bG_smem_doubles<NR>());
      for (int st = 0; st < nblk + 1; ++st) {
         for (int t = 0; t < ntile + 1; ++t) {
            std::fill(smG.begin(), smG.end(), std::nan(""));
            emu::run_cta(SW_GT, [&](emu::Ctx& cx) {
               bwd_wide_G<NR>(cx, f, t, st, x.data(), pbuf.data(), smG.data()); });
         }
         std::fill(smT.begin(), smT.end(), std::nan(""));
         emu::run_cta(SW_TT, [&](emu::Ctx& cx) { bwd_wide_T<NR, NR, POSDEF>(cx, f, st, x.data(), pbuf.data(), smT.data()); });
      }
      for (size_t e = 0; e < x.size(); ++e) {
         double d = std::fabs(x[e] - xr[e]);
         if (!(d <= 1e300)) d = 1e300;
         worst = std::max(worst, d);
      }
   }
   return worst;
}

int main() {
   std::mt19937_64 rng(7);
   const Case cases[] = {
      {700, 600, 600, 600, false},   // two full blocks + one of 88, rows below
      {520, 300, 257, 257, false},   // delayed columns received (n > n0), block of 1 column
      {300, 300, 300, 300, false},   // root: no rows below
      {1000, 256, 256, 256, false},  // exactly one block
      {900, 40, 40, 33, false},      // not everything eliminated (n > nelim), one short block
      {640, 512, 512, 512, true},    // positive definite
      {333, 300, 280, 270, true},
   };
   int failures = 0;
   for (const Case& c : cases) {
      double e1 = c.posdef ? run_case<1, true>(c, rng) : run_case<1, false>(c, rng);
      /* (4 right-hand sides on the smaller cases only: the emulation spends its time in barriers) */
      double e4 = c.m > 600 ? 0.0 : c.posdef ? run_case<4, true>(c, rng) : run_case<4, false>(c, rng);
      bool ok = e1 < 1e-11 && e4 < 1e-11;
      printf("m=%d n=%d n0=%d nelim=%d posdef=%d: max |wide - reference| = %.2e (1 rhs) %.2e (4 rhs) %s\n",
             c.m, c.n, c.n0, c.nelim, (int)c.posdef, e1, e4, ok ? "ok" : "FAIL");
      failures += !ok;
   }
   {  /* several right-hand sides per warp (NRW = 2 and 4) */
      const Case c{520, 300, 257, 257, false};
      double e16 = run_case<16, false>(c, rng), e32 = run_case<32, false>(c, rng);
      const Case cp{333, 300, 280, 270, true};
      double p32 = run_case<32, true>(cp, rng);
      bool ok = e16 < 1e-11 && e32 < 1e-11 && p32 < 1e-11;
      printf("16 / 32 right-hand sides: %.2e %.2e, posdef 32: %.2e %s\n", e16, e32, p32, ok ? "ok" : "FAIL");
      failures += !ok;
   }
   printf("solve_wide_emu: %d failures\n", failures);
   return failures ? 1 : 0;
}
