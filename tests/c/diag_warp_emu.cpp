/* TEST INFRASTRUCTURE (CPU only): the one-warp diagonal-block factorisation (spral_b200/csrc/diag_warp.cuh,
 * the body of k_diag_w and of the chain kernel) on the SIMT emulator of tests/emu (32 fibers, warp collectives
 * as rendezvous points -- an unsynchronised shared-memory hazard between lanes shows up as a wrong result,
 * because a fiber runs alone until its next rendezvous).  Checks P A P^T = L D L^T, the mirrored L*D, the
 * bound |l_ij| <= 1 inside a block for 1x1 pivots, zero pivots, short blocks and A = L L^T.
 * Build: g++ -O1 -std=c++17 -Itests/emu -Ispral_b200/csrc -Iinclude -I$CUDA/include -include tests/emu/cuda_emu.h
 *        tests/c/diag_warp_emu.cpp tests/emu/cuda_emu_rt.cpp -lrt */
#include "diag_warp.cuh"
#include <cstdio>
#include <random>
using namespace b200;
static double S[32 * DW_LD], dinv[64];
static int lperm[32];

static int check_ldlt(int bs, int seed, int kind) {
   std::mt19937 g(seed);
   std::uniform_real_distribution<double> u(-1, 1);
   double A[32][32] = {};
   for (int i = 0; i < bs; i++)
      for (int j = 0; j <= i; j++) {
         double v = u(g);
         if (kind == 1) v = std::round(4 * v) / 4;                 // many ties
         if (kind == 2 && i >= bs / 2) v = 0.0;                     // zero rows: zero pivots
         A[i][j] = A[j][i] = v;
      }
   for (int i = 0; i < 32 * DW_LD; i++) S[i] = 0;
   for (int i = 0; i < bs; i++) for (int j = 0; j <= i; j++) S[i * DW_LD + j] = A[i][j];
   int zf = 0, rc = 0;
   emu::launch(1, 32, 0, [&]() { int z; int r = diag_warp_ldlt(S, dinv, lperm, bs, 1e-20, 1, INFINITY, z); if (threadIdx.x == 0) { zf = z; rc = r; } });
   if (rc != DW_OK) { printf("bs %d seed %d kind %d: rc %d\n", bs, seed, kind, rc); return 1; }
   if (kind == 0 && zf != 32) { printf("bs %d seed %d: unexpected zero pivots from %d\n", bs, seed, zf); return 1; }
   if (kind == 2 && zf > bs / 2 + 1) { printf("bs %d seed %d: zfrom %d, expected <= %d\n", bs, seed, zf, bs / 2 + 1); return 1; }
   double L[32][32] = {}, D[32][32] = {};
   const int ne = zf < bs ? zf : bs;
   for (int i = 0; i < bs; i++) { L[i][i] = 1; for (int j = 0; j < i; j++) L[i][j] = S[i * DW_LD + j]; }
   for (int i = 0; i < ne;) {
      if (i + 1 < ne && std::isinf(dinv[2 * i + 2])) {
         double a = dinv[2 * i], b = dinv[2 * i + 1], c = dinv[2 * i + 3], det = a * c - b * b;
         D[i][i] = c / det; D[i + 1][i + 1] = a / det; D[i][i + 1] = D[i + 1][i] = -b / det; i += 2;
      } else { D[i][i] = 1.0 / dinv[2 * i]; i++; }
   }
   double err = 0, e2 = 0;
   for (int i = 0; i < bs; i++)
      for (int j = 0; j < bs; j++) {
         double s = 0;
         for (int k = 0; k < bs; k++) for (int l = 0; l < bs; l++) s += L[i][k] * D[k][l] * L[j][l];
         err = std::max(err, fabs(s - A[lperm[i]][lperm[j]]));
      }
   for (int i = 0; i < bs; i++)
      for (int j = 0; j < i && j < ne; j++) {
         if (i == j + 1 && std::isinf(dinv[2 * i])) continue;      // inside a 2x2 pivot L*D is not kept
         double s = 0;
         for (int k = 0; k < bs; k++) s += L[i][k] * D[k][j];
         e2 = std::max(e2, fabs(s - S[j * DW_LD + i]));
      }
   bool perm_ok = true; int seen[32] = {};
   for (int i = 0; i < 32; i++) { if (lperm[i] < 0 || lperm[i] > 31 || seen[lperm[i]]++) perm_ok = false; }
   if (err > 1e-13 || e2 > 1e-13 || !perm_ok) { printf("bs %d seed %d kind %d: err %g LD err %g perm %d\n", bs, seed, kind, err, e2, (int)perm_ok); return 1; }
   return 0;
}

static int check_chol(int bs, int seed) {
   std::mt19937 g(seed);
   std::uniform_real_distribution<double> u(-1, 1);
   double A[32][32] = {};
   for (int i = 0; i < bs; i++) for (int j = 0; j <= i; j++) A[i][j] = A[j][i] = u(g) + (i == j ? bs : 0);
   for (int i = 0; i < 32 * DW_LD; i++) S[i] = 0;
   for (int i = 0; i < bs; i++) for (int j = 0; j <= i; j++) S[i * DW_LD + j] = A[i][j];
   int rc = 0;
   emu::launch(1, 32, 0, [&]() { int r = diag_warp_chol(S, dinv, bs); if (threadIdx.x == 0) rc = r; });
   if (rc != DW_OK) return 1;
   double err = 0;
   for (int i = 0; i < bs; i++)
      for (int j = 0; j <= i; j++) {
         double s = 0;
         for (int k = 0; k <= j; k++) s += S[i * DW_LD + k] * S[j * DW_LD + k];
         err = std::max(err, fabs(s - A[i][j]));
      }
   if (err > 1e-12) { printf("chol bs %d seed %d: err %g\n", bs, seed, err); return 1; }
   /* not positive definite */
   S[0] = -1.0;
   emu::launch(1, 32, 0, [&]() { int r = diag_warp_chol(S, dinv, bs); if (threadIdx.x == 0) rc = r; });
   return rc == DW_NOT_POS_DEF ? 0 : 1;
}

int main() {
   int failures = 0, cases = 0;
   const int sizes[] = {1, 2, 3, 5, 8, 17, 31, 32};
   for (int bs : sizes)
      for (int seed = 1; seed <= 6; ++seed)
         for (int kind = 0; kind < 3; ++kind) { if (kind == 2 && bs < 4) continue; failures += check_ldlt(bs, seed, kind); ++cases; }
   for (int bs : sizes) { failures += check_chol(bs, bs); ++cases; }
   printf("diag_warp_emu: %d cases, %d failures\n", cases, failures);
   return failures != 0;
}
