/* TEST INFRASTRUCTURE (CPU only): runs the body of k_diag_v2
 * (spral_b200/csrc/diag_block.h) on host threads -- one pthread per CUDA thread,
 * __syncthreads and the warp shuffles emulated with barriers -- and compares the
 * published results BIT FOR BIT with a sequential model of the thread-per-entry
 * kernel k_diag (factor_kernels.cu), on random / tie-ridden / singular / short
 * blocks, indefinite and positive definite.  Also checks P A P^T = L D L^T and,
 * when the reference tree is present (-DHAVE_REF), that the pivot sequence and D
 * agree with the reference's own block_ldlt<double,32>
 * (src/ssids/cpu/kernels/block_ldlt.hxx:289-413).
 *
 * Build: g++ -O2 -std=c++17 -ffp-contract=off -pthread [-DHAVE_REF -mavx2 -I$REF/src] */
#include <pthread.h>
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <random>
#include <vector>

#include "../../spral_b200/csrc/diag_block.h"

#ifdef HAVE_REF
#include "ssids/cpu/kernels/block_ldlt.hxx"
#endif

#include "diag_model.h"
using namespace b200;
using diag_model::BS; using diag_model::INF; using diag_model::Published; using diag_model::model_v1;
constexpr int NW = 4;
constexpr int NT = NW * 32;

/* ------------------------------------------------------------------ */
/* emulation of a CTA of NT threads                                    */
/* ------------------------------------------------------------------ */
struct EmuBlock {
   pthread_barrier_t bar;
   struct Warp { pthread_barrier_t bar; double xd[32]; int xi[32]; } warp[NW];
   DiagShared<NW> sh;
};
struct EmuCtx {
   EmuBlock* b; int t;
   int tid() const { return t; }
   void sync() { pthread_barrier_wait(&b->bar); }
   double shfl_xor(double v, int off) {
      auto& w = b->warp[t >> 5];
      w.xd[t & 31] = v; pthread_barrier_wait(&w.bar);
      double r = w.xd[(t & 31) ^ off]; pthread_barrier_wait(&w.bar);
      return r;
   }
   int shfl_xor(int v, int off) {
      auto& w = b->warp[t >> 5];
      w.xi[t & 31] = v; pthread_barrier_wait(&w.bar);
      int r = w.xi[(t & 31) ^ off]; pthread_barrier_wait(&w.bar);
      return r;
   }
};
struct Job {
   EmuBlock* blk; int t; bool posdef; const double* Ld; int ldl, bs; double small; int action; Published* out;
};

/* mirrors the publishing code of k_diag_v2 */
template <bool POSDEF>
static void thread_body(Job* j) {
   EmuCtx cx{j->blk, j->t};
   DiagShared<NW>& sh = j->blk->sh;
   Published& out = *j->out;
   constexpr int RPT = BS / NW;
   const int c = j->t & 31, q = j->t >> 5, bs = j->bs;
   int cur = 0, zfrom = BS;
   int rc = diag_block_factor<NW, POSDEF>(cx, sh, j->Ld, (size_t)j->ldl, bs, j->small, j->action, INF,
                                          POSDEF ? nullptr : out.a0, cur, zfrom);
   if (rc != DB_OK) { if (j->t == 0) out.rc = rc; return; }
   if (POSDEF) {
      for (int i = 0; i < RPT; ++i) {
         const int r = q * RPT + i;
         const double l = (r < bs && c < bs && r >= c) ? sh.A[cur][r][c] : 0.0;
         if (r < bs && c < bs && r >= c) out.lblock[r + c * BS] = l;
         out.l11[r + c * BS] = l;
      }
      if (q == 0) out.dinv[c] = (c < bs) ? sh.dinv[c] : 0.0;
   } else {
      for (int i = 0; i < RPT; ++i) {
         const int r = q * RPT + i;
         double l = 0.0, y = 0.0;
         if (r < bs && c < bs) {
            if (r > c) { l = sh.A[cur][r][c]; y = sh.LDm[cur][r][c]; }
            else if (r == c) l = 1.0;
         }
         out.l11[r + c * BS] = l; out.ld11[r + c * BS] = y;
      }
      if (q == 0) {
         out.dinv[2 * c] = (c < bs) ? sh.dinv[2 * c] : 0.0;
         out.dinv[2 * c + 1] = (c < bs) ? sh.dinv[2 * c + 1] : 0.0;
         out.lperm[c] = sh.lperm[c];
         if (c == 0) out.zfrom = zfrom;
      }
   }
}
static void* thread_entry(void* p) {
   Job* j = (Job*)p;
   if (j->posdef) thread_body<true>(j); else thread_body<false>(j);
   return nullptr;
}

static void emu_v2(bool posdef, const double* Ld, int ldl, int bs, double small, int action, Published& out) {
   static EmuBlock blk;
   std::memset(&out, 0, sizeof(out));
   out.zfrom = BS;
   std::memset(&blk.sh, 0xAB, sizeof(blk.sh));          // shared memory is not zeroed on the device either
   pthread_barrier_init(&blk.bar, nullptr, NT);
   for (int w = 0; w < NW; ++w) pthread_barrier_init(&blk.warp[w].bar, nullptr, 32);
   std::vector<pthread_t> th(NT);
   std::vector<Job> jobs(NT);
   pthread_attr_t at; pthread_attr_init(&at); pthread_attr_setstacksize(&at, 256 << 10);
   for (int t = 0; t < NT; ++t) {
      jobs[t] = Job{&blk, t, posdef, Ld, ldl, bs, small, action, &out};
      if (pthread_create(&th[t], &at, thread_entry, &jobs[t]) != 0) { perror("pthread_create"); exit(2); }
   }
   for (int t = 0; t < NT; ++t) pthread_join(th[t], nullptr);
   pthread_attr_destroy(&at);
   pthread_barrier_destroy(&blk.bar);
   for (int w = 0; w < NW; ++w) pthread_barrier_destroy(&blk.warp[w].bar);
}

/* ------------------------------------------------------------------ */
static bool same_bits(const void* a, const void* b, size_t n) { return std::memcmp(a, b, n) == 0; }

static int compare(const char* what, int idx, bool posdef, const Published& a, const Published& b) {
   int bad = 0;
   if (a.rc != b.rc) { printf("%s[%d]: rc %d vs %d\n", what, idx, a.rc, b.rc); return 1; }
   if (a.rc != 0) return 0;
   bad += !same_bits(a.l11, b.l11, sizeof(a.l11));
   bad += !same_bits(a.dinv, b.dinv, sizeof(a.dinv));
   if (posdef) bad += !same_bits(a.lblock, b.lblock, sizeof(a.lblock));
   else {
      bad += !same_bits(a.ld11, b.ld11, sizeof(a.ld11));
      bad += !same_bits(a.a0, b.a0, sizeof(a.a0));
      bad += !same_bits(a.lperm, b.lperm, sizeof(a.lperm));
      bad += (a.zfrom != b.zfrom);
   }
   if (bad) printf("%s[%d]: model and emulated kernel differ (%d fields)\n", what, idx, bad);
   return bad ? 1 : 0;
}

/* || P A P^T - L D L^T ||_max / ||A||_max over the eliminated (non-zero-pivot) part */
static double reconstruction_error(const double* Afull, int bs, const Published& o) {
   std::vector<double> D(BS * BS, 0.0);
   for (int j = 0; j < o.zfrom && j < bs;) {
      if (j + 1 < bs && std::isinf(o.dinv[2 * j + 2])) {       // invert the stored 2x2 of D^-1
         double e11 = o.dinv[2 * j], e21 = o.dinv[2 * j + 1], e22 = o.dinv[2 * j + 3];
         double det = e11 * e22 - e21 * e21;
         D[j + j * BS] = e22 / det; D[j + 1 + (j + 1) * BS] = e11 / det;
         D[j + 1 + j * BS] = D[j + (j + 1) * BS] = -e21 / det;
         j += 2;
      } else { D[j + j * BS] = 1.0 / o.dinv[2 * j]; j += 1; }
   }
   double err = 0, nrm = 0;
   const int ne = std::min(o.zfrom, bs);
   for (int c = 0; c < bs; ++c)
      for (int r = c; r < bs; ++r) {
         double s = 0;
         for (int k = 0; k < ne; ++k)
            for (int l = 0; l < ne; ++l)
               if (D[k + l * BS] != 0.0) s += o.l11[r + k * BS] * D[k + l * BS] * o.l11[c + l * BS];
         double a = Afull[o.lperm[r] + o.lperm[c] * BS];
         nrm = std::max(nrm, std::fabs(a));
         if (r < ne || c < ne || true) err = std::max(err, std::fabs(a - s));
      }
   return nrm > 0 ? err / nrm : err;
}

int main(int argc, char** argv) {
   int ncase = argc > 1 ? atoi(argv[1]) : 40;
   std::mt19937_64 rng(20261017);
   std::uniform_real_distribution<double> U(-1.0, 1.0);
   int failures = 0, checked = 0, ref_checked = 0, n2x2 = 0, nzero = 0;
   double worst_rec = 0;
   const int ldl = BS + 6;
   for (int it = 0; it < ncase; ++it) {
      const int kind = it % 8;
      int bs = BS;
      bool posdef = false;
      int action = 1;
      std::vector<double> Afull(BS * BS, 0.0);
      auto sym = [&](int r, int c, double v) { Afull[r + c * BS] = v; Afull[c + r * BS] = v; };
      if (kind == 0 || kind == 1) {                      // random dense indefinite
         for (int c = 0; c < BS; ++c) for (int r = c; r < BS; ++r) sym(r, c, U(rng));
      } else if (kind == 2) {                            // exact ties: stencil-like integers, zero diagonal part
         for (int c = 0; c < BS; ++c) for (int r = c; r < BS; ++r)
            sym(r, c, r == c ? (c % 3 == 0 ? 0.0 : 13.0 - (c % 5)) : ((r + c) % 4 == 0 ? -1.0 : 0.0));
      } else if (kind == 3) {                            // short block
         bs = 1 + (int)(rng() % 31);
         for (int c = 0; c < bs; ++c) for (int r = c; r < bs; ++r) sym(r, c, U(rng));
      } else if (kind == 4) {                            // rank deficient: rank-k outer product, exact zeros
         int k = 1 + (int)(rng() % 20);
         for (int c = 0; c < BS; ++c) for (int r = c; r < BS; ++r) sym(r, c, (r < k && c < k) ? U(rng) : 0.0);
         action = (it % 16 == 4) ? 0 : 1;
      } else if (kind == 5) {                            // positive definite
         posdef = true;
         std::vector<double> G(BS * BS);
         for (auto& g : G) g = U(rng);
         for (int c = 0; c < BS; ++c) for (int r = c; r < BS; ++r) {
            double s = (r == c) ? 1.0 : 0.0;
            for (int k = 0; k < BS; ++k) s += G[r + k * BS] * G[c + k * BS];
            sym(r, c, s);
         }
         if (it % 16 == 13) { bs = 17; }
      } else if (kind == 6) {                            // not positive definite under posdef=true
         posdef = true;
         for (int c = 0; c < BS; ++c) for (int r = c; r < BS; ++r) sym(r, c, r == c ? (c == 9 ? -1.0 : 4.0) : 0.25);
      } else {                                           // KKT-like: zero (2,2) block
         for (int c = 0; c < BS; ++c) for (int r = c; r < BS; ++r)
            sym(r, c, (c >= 20) ? 0.0 : (r == c ? 2.0 + U(rng) : 0.3 * U(rng)));
      }
      /* the front stores the lower triangle only; poison the rest */
      std::vector<double> Ld((size_t)ldl * BS, std::nan(""));
      for (int c = 0; c < BS; ++c) for (int r = c; r < BS; ++r) Ld[r + (size_t)c * ldl] = Afull[r + c * BS];

      Published m1, e2;
      model_v1(posdef, Ld.data(), ldl, bs, 1e-20, action, m1);
      emu_v2(posdef, Ld.data(), ldl, bs, 1e-20, action, e2);
      failures += compare("case", it, posdef, m1, e2);
      ++checked;
      if (!posdef && m1.rc == 0) {
         double rec = reconstruction_error(Afull.data(), bs, m1);
         /* zero-pivot blocks: the remainder is exactly zero in these cases, so the bound holds */
         worst_rec = std::max(worst_rec, rec);
         if (!(rec < 1e-10)) { printf("case[%d]: reconstruction error %.3e\n", it, rec); ++failures; }
         for (int j = 0; j < bs; ++j) if (std::isinf(m1.dinv[2 * j])) ++n2x2;
         if (m1.zfrom < bs) ++nzero;
      }
#ifdef HAVE_REF
      if (!posdef && bs == BS && m1.rc == 0 && m1.zfrom == BS && (kind == 0 || kind == 1 || kind == 7)) {
         using namespace spral::ssids::cpu;
         alignas(64) static double a[BS * BS], d[2 * BS], ldwork[2 * BS * BS + BS];   // SimdVec uses aligned loads
         std::fill(a, a + BS * BS, 0.0); std::fill(d, d + 2 * BS, 0.0); std::fill(ldwork, ldwork + 2 * BS * BS + BS, 0.0);
         for (int c = 0; c < BS; ++c) for (int r = c; r < BS; ++r) a[r + c * BS] = Afull[r + c * BS];
         int perm[BS], lperm[BS];
         for (int i = 0; i < BS; ++i) { perm[i] = i; lperm[i] = i; }
         block_ldlt<double, BS>(0, perm, a, BS, d, ldwork, true, 0.01, 1e-20, lperm);
         bool ok = true;
         for (int i = 0; i < BS; ++i) ok = ok && (lperm[i] == m1.lperm[i]);
         for (int i = 0; i < 2 * BS && ok; ++i) {
            double x = d[i], y = m1.dinv[i];
            if (std::isinf(x) || std::isinf(y)) ok = (x == y);
            else ok = std::fabs(x - y) <= 1e-9 * std::max(1.0, std::fabs(x));
         }
         if (!ok) { printf("case[%d]: pivot sequence / D differ from the reference block_ldlt\n", it); ++failures; }
         ++ref_checked;
      }
#endif
   }
   printf("diag_block_emu: %d blocks, %d failures, %d second columns of 2x2 pivots, %d zero-pivot blocks, "
          "worst reconstruction error %.2e, %d compared with the reference block_ldlt\n",
          checked, failures, n2x2, nzero, worst_rec, ref_checked);
   return failures ? 1 : 0;
}
