/* TEST INFRASTRUCTURE (CPU only): the bodies of the speculative panel kernels
 * (spral_b200/csrc/panel_v2.h: chain_segment, panel_tile) run on host threads
 * (tests/c/emu.h) and are checked against the mathematics they implement:
 *   chain   P A11 P^T = L11 D L11^T on the 128 x 128 diagonal block, |L11| <= 1/u,
 *           D^-1 in the reference CPU layout (2x2 marked by +Inf, block_ldlt.hxx:403-406)
 *   tiles   A21 P^T = (W D) L11^T and L*D = W D for the rows below, backup == originals
 * and on the give-up paths: a failed a-posteriori test inside the diagonal block, in the
 * rows below, and an all-zero block (zero pivots are left to the step-by-step path). */
#include <algorithm>
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

struct Result { int ok; int tile_fail; double err11, err21, errld, maxl; int n2x2; bool bk_ok; };

static Result run(int m, std::mt19937_64& rng, int kind) {
   std::uniform_real_distribution<double> U(-1.0, 1.0);
   const int ldl = (m + 1) / 2 * 2;
   const int p = 0;
   std::vector<double> A((size_t)ldl * CW, std::nan(""));
   for (int c = 0; c < CW; ++c)
      for (int r = c; r < m; ++r) {
         double v = U(rng);
         if (r == c) v = (c % 3 == 0 ? -1.0 : 1.0) * (2.0 + U(rng));
         if (kind == 1 && c < 32 && r < 32) v *= 1e-9;            // tiny leading block, large rows inside the diagonal block
         if (kind == 2 && r >= CW) v *= 1e7;                       // huge rows below: the tiles must fail
         if (kind == 3 && c >= 64 && c < 96 && r >= 64 && r < 96) v = 0.0;   // an exactly zero 32 x 32 block ...
         if (kind == 3 && c < 64 && r >= 64 && r < 96) v = 0.0;              // ... that stays zero after the updates
         A[r + (size_t)c * ldl] = v;
      }
   std::vector<double> L = A, LD((size_t)ldl * CW, std::nan("")), BK((size_t)ldl * CW, std::nan(""));
   auto ws = std::make_unique<SegWS>();
   auto csh = std::make_unique<ChainShared>();
   std::fill((char*)csh.get(), (char*)csh.get() + sizeof(ChainShared), (char)0xAB);
   int ok = -1;
   emu::run_cta(CNT, [&](emu::Ctx& cx) {
      int r = chain_segment<false>(cx, *csh, L.data() + p + (size_t)p * ldl, (size_t)ldl, 0.01, 1e-20, INF, ws.get());
      if (cx.tid() == 0) ok = r;
   });
   Result res{ok, 0, 0, 0, 0, 0, 0, true};
   if (!ok) return res;
   /* ---- diagonal block ---- */
   int P[CW];
   for (int i = 0; i < CW; ++i) P[i] = (i / 32) * 32 + ws->lperm[i];
   std::vector<double> D(CW * CW, 0.0);
   for (int j = 0; j < CW;) {
      if (j + 1 < CW && ws->dinv[2 * j + 2] == INF) {
         double e11 = ws->dinv[2 * j], e21 = ws->dinv[2 * j + 1], e22 = ws->dinv[2 * j + 3];
         double det = e11 * e22 - e21 * e21;
         D[j + j * CW] = e22 / det; D[j + 1 + (j + 1) * CW] = e11 / det; D[j + 1 + j * CW] = D[j + (j + 1) * CW] = -e21 / det;
         res.n2x2++; j += 2;
      } else { D[j + j * CW] = 1.0 / ws->dinv[2 * j]; j += 1; }
   }
   auto a11 = [&](int r, int c) { return r >= c ? A[r + (size_t)c * ldl] : A[c + (size_t)r * ldl]; };
   std::vector<double> LDm(CW * CW, 0.0);              // L11 * D
   for (int i = 0; i < CW; ++i)
      for (int k = 0; k < CW; ++k) {
         double s = 0;
         for (int q = 0; q < CW; ++q) if (D[q + k * CW] != 0.0) s += ws->l11[i + (size_t)q * CW] * D[q + k * CW];
         LDm[i + k * CW] = s;
      }
   for (int i = 0; i < CW; ++i)
      for (int c = 0; c <= i; ++c) {
         double s = 0;
         for (int k = 0; k < CW; ++k) s += LDm[i + k * CW] * ws->l11[c + (size_t)k * CW];
         res.err11 = std::max(res.err11, std::fabs(s - a11(P[i], P[c])));
         if (i > c) res.maxl = std::max(res.maxl, std::fabs(ws->l11[i + (size_t)c * CW]));
      }
   /* ---- rows below ---- */
   auto tsh = std::make_unique<TileShared>();
   for (int r0 = 0; r0 < m; r0 += RT) {
      if (r0 + RT <= p + CW) continue;
      std::fill((char*)tsh.get(), (char*)tsh.get() + sizeof(TileShared), (char)0xAB);
      emu::run_cta(RT, [&](emu::Ctx& cx) {
         panel_tile<false>(cx, *tsh, L.data() + (size_t)p * ldl, LD.data() + (size_t)p * ldl, BK.data(), (size_t)ldl, m, r0, p,
                    0.01, INF, ws.get(), &res.tile_fail);
      });
   }
   for (int r = CW; r < m; ++r)
      for (int c = 0; c < CW; ++c) {
         double s = 0, sd = 0;
         for (int k = 0; k < CW; ++k) {
            double wd = 0;                                   // (W D)(r, k)
            for (int q = 0; q < CW; ++q) if (D[q + k * CW] != 0.0) wd += L[r + (size_t)q * ldl] * D[q + k * CW];
            s += wd * ws->l11[c + (size_t)k * CW];
            if (k == c) sd = wd;
         }
         res.err21 = std::max(res.err21, std::fabs(s - A[r + (size_t)P[c] * ldl]));
         res.errld = std::max(res.errld, std::fabs(sd - LD[r + (size_t)c * ldl]));
         res.maxl = std::max(res.maxl, std::fabs(L[r + (size_t)c * ldl]));
         if (BK[r + (size_t)c * ldl] != A[r + (size_t)c * ldl]) res.bk_ok = false;
      }
   return res;
}

int main() {
   std::mt19937_64 rng(99);
   int failures = 0;
   for (int m : {128, 200, 256, 391, 700}) {
      Result r = run(m, rng, 0);
      bool good = r.ok == 1 && !r.tile_fail && r.err11 < 1e-11 && r.err21 < 1e-10 && r.errld < 1e-11 && r.maxl <= 100.0 && r.bk_ok;
      printf("m=%d: ok=%d tile_fail=%d |PAP'-LDL'|=%.1e |A21P'-WDL'|=%.1e |LD-WD|=%.1e max|l|=%.2f 2x2=%d backup=%d %s\n",
             m, r.ok, r.tile_fail, r.err11, r.err21, r.errld, r.maxl, r.n2x2, (int)r.bk_ok, good ? "ok" : "FAIL");
      failures += !good;
   }
   {  Result r = run(300, rng, 1);
      printf("tiny leading block: chain ok=%d (expected 0)\n", r.ok); failures += (r.ok != 0); }
   {  Result r = run(300, rng, 2);
      printf("huge rows below: chain ok=%d tile_fail=%d (expected 1, 1)\n", r.ok, r.tile_fail);
      failures += !(r.ok == 1 && r.tile_fail == 1 && r.bk_ok); }
   {  Result r = run(300, rng, 3);
      printf("zero block: chain ok=%d (expected 0)\n", r.ok); failures += (r.ok != 0); }
   printf("panel_v2_emu: %d failures\n", failures);
   return failures ? 1 : 0;
}
