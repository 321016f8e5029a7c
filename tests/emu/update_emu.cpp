/* TEST INFRASTRUCTURE (CPU only): stands in for spral_b200/csrc/gemm_dmma.cu in the emulated library -- that file is
 * inline PTX (mma.sync, cp.async.bulk, mbarrier) and cannot be emulated.  Same regions (make_region, explicit_region),
 * same tile clipping (load_job, store_tile), the contraction as a plain loop over k. */
#include "cuda_emu.h"
#include "../../spral_b200/csrc/engine.h"
#include "../../spral_b200/csrc/pivot_state.h"

namespace b200 {
namespace {
struct Region {
   const double* A; const double* B; double* C;
   size_t lda, ldb, ldc;
   int m, c_lo, c_hi, k0, k1;
   bool accumulate, valid;
};
Region make_region(const Front* f, int mode) {
   Region g;
   g.valid = false;
   g.A = f->L; g.B = f->LD; g.lda = g.ldb = (size_t)f->ldl; g.m = f->m;
   if (mode == UPD_INNER) {
      if (!f->step_valid) return g;
      int ne = calc_ne(f);
      if (ne == 0) return g;
      g.k0 = f->done; g.k1 = f->done + ne;
      g.c_lo = f->done + ne; g.c_hi = f->pend0;
      g.C = f->L; g.ldc = g.lda; g.accumulate = true;
   } else if (mode == UPD_OUTER) {
      if (!f->panel_open || f->finished) return g;
      int done = f->done;
      if (f->step_valid) done += calc_ne(f);
      g.k0 = f->p0; g.k1 = done;
      if (g.k1 <= g.k0) return g;
      g.c_lo = f->pend0; g.c_hi = f->n;
      g.C = f->L; g.ldc = g.lda; g.accumulate = true;
   } else if (mode == UPD_SEG) {
      if (!f->seg_valid || !f->seg_ok || f->seg_fail) return g;
      g.k0 = f->done; g.k1 = f->done + CW;
      g.c_lo = f->done + CW; g.c_hi = f->pend0;
      g.C = f->L; g.ldc = g.lda; g.accumulate = true;
   } else {
      if (f->m == f->n || !f->C) return g;
      g.k0 = 0; g.k1 = f->nelim;
      g.c_lo = f->n; g.c_hi = f->m;
      g.ldc = (size_t)f->ldc;
      g.C = f->C - (ptrdiff_t)f->n - (ptrdiff_t)f->n * (ptrdiff_t)g.ldc;
      g.accumulate = false;
   }
   if (g.c_lo >= g.c_hi) return g;
   g.valid = true;
   return g;
}
Region explicit_region(const Front* fronts, const int4 xr) {
   const Front* f = &fronts[xr.x];
   Region g;
   g.A = f->L; g.B = f->LD; g.C = f->L;
   g.lda = g.ldb = g.ldc = (size_t)f->ldl; g.m = f->m;
   g.k0 = xr.y; g.k1 = xr.z; g.c_lo = xr.w; g.c_hi = f->n;
   g.accumulate = true;
   g.valid = (g.k1 > g.k0) && (g.c_lo < g.c_hi);
   return g;
}
}

int update_tile_size(bool big_tiles) { return big_tiles ? 128 : 64; }
int inner_tile_size(bool big_tiles) { return (big_tiles && getenv("SPRAL_B200_INNER128")) ? 128 : 64; }
void configure_update_kernels() {}
int device_sm_count() { return 148; }

void launch_update(Front* fronts, const MatTile* work, int nwork, UpdateMode mode, bool big_tiles, cudaStream_t,
      int, const int4* xregs) {
   if (nwork == 0) return;
   g_launches.fetch_add(1, std::memory_order_relaxed);
   const int T = (mode == UPD_INNER) ? inner_tile_size(big_tiles) : (big_tiles ? 128 : 64);
   for (int it = 0; it < nwork; ++it) {
      const MatTile w = work[it];
      const Region g = (mode == UPD_EXPLICIT) ? explicit_region(fronts, xregs[w.front]) : make_region(&fronts[w.front], (int)mode);
      if (!g.valid) continue;
      const int r0 = w.ti * T, c0 = w.tj * T;
      if (c0 + T <= g.c_lo || c0 >= g.c_hi || r0 >= g.m) continue;
      for (int c = std::max(c0, g.c_lo); c < std::min(c0 + T, g.c_hi); ++c) {
         double* Cc = g.C + (ptrdiff_t)c * (ptrdiff_t)g.ldc;
         for (int r = std::max(r0, c); r < std::min(r0 + T, g.m); ++r) {
            double acc = 0.0;
            for (int k = g.k0; k < g.k1; ++k) acc += g.A[r + (size_t)k * g.lda] * g.B[c + (size_t)k * g.ldb];
            Cc[r] = (g.accumulate ? Cc[r] : 0.0) - acc;
         }
      }
   }
}
} // namespace b200

extern "C" double spral_ssids_b200_fp64_peak_tflops(int) { return 1.0; }
