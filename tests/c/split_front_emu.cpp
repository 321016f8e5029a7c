/* TEST INFRASTRUCTURE (CPU only): runs the REAL host code of the distributed top front
 * (spral_b200/csrc/split_front.h: SplitOwner, split_helper_serve, the shared-memory protocol) against a mock of the
 * few CUDA runtime calls it makes -- streams are in-order worker threads, "device" memory is host memory, an IPC
 * handle carries a plain pointer, the UPD_EXPLICIT kernel is a triple loop with the same region / tile semantics
 * (gemm_dmma.cu: explicit_region, load_job, store_tile) -- with the owner and the helper as two threads of this
 * process.  What it checks is the arithmetic the model check (dist_front_emu.cpp) cannot see: offsets, pitches,
 * tile lists, region tables, which block comes back when.
 *
 * The "front" is a dense m x n lower trapezoid.  A panel is frozen (L*D := 0.5 L on its columns -- any fixed
 * function does) and every column to its right gets A(r,c) -= sum_j L(r,j) LD(c,j), panel after panel, exactly
 * the data flow of factor_fronts without pivoting.  The split run (urgent update of the next block by the owner,
 * everything further right by the helper, optional "failed pivot" at some panel -> drain) must reproduce the
 * single-process run BIT FOR BIT.
 *
 * Build: g++ -O2 -std=c++17 -pthread -DSPRAL_B200_SPLIT -I/usr/local/cuda/include -Iinclude tests/c/split_front_emu.cpp -lrt */
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <mutex>
#include <random>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "../../spral_b200/csrc/engine.h"

/* ---- mock CUDA runtime (only what split_front.h calls) ---------------------------------------------------- */
struct MockStream {
   std::mutex mtx;
   std::condition_variable cv, idle;
   std::deque<std::function<void()>> q;
   bool busy = false, stop = false;
   std::thread th;
   MockStream() : th([this] { run(); }) {}
   ~MockStream() { { std::lock_guard<std::mutex> l(mtx); stop = true; } cv.notify_all(); th.join(); }
   void run() {
      for (;;) {
         std::function<void()> fn;
         {
            std::unique_lock<std::mutex> l(mtx);
            cv.wait(l, [&] { return stop || !q.empty(); });
            if (q.empty()) return;
            fn = std::move(q.front()); q.pop_front(); busy = true;
         }
         fn();
         { std::lock_guard<std::mutex> l(mtx); busy = false; }
         idle.notify_all();
      }
   }
   void push(std::function<void()> fn) { { std::lock_guard<std::mutex> l(mtx); q.push_back(std::move(fn)); } cv.notify_all(); }
   void sync() { std::unique_lock<std::mutex> l(mtx); idle.wait(l, [&] { return q.empty() && !busy; }); }
};
static MockStream* MS(cudaStream_t s) { return reinterpret_cast<MockStream*>(s); }

extern "C" {
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned int) { *s = reinterpret_cast<cudaStream_t>(new MockStream); return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) { delete MS(s); return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t s) { MS(s)->sync(); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t n, cudaMemcpyKind, cudaStream_t s) {
   std::vector<char> stage((const char*)src, (const char*)src + n);          // pageable source: consumed at call time
   MS(s)->push([dst, stage] { std::memcpy(dst, stage.data(), stage.size()); });
   return cudaSuccess;
}
cudaError_t cudaMemcpy2DAsync(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height,
      cudaMemcpyKind, cudaStream_t s) {
   MS(s)->push([=] { for (size_t j = 0; j < height; ++j) std::memcpy((char*)dst + j * dpitch, (const char*)src + j * spitch, width); });
   return cudaSuccess;
}
cudaError_t cudaLaunchHostFunc(cudaStream_t s, cudaHostFn_t fn, void* arg) { MS(s)->push([=] { fn(arg); }); return cudaSuccess; }
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p) { std::memset(h, 0, sizeof(*h)); std::memcpy(h, &p, sizeof(p)); return cudaSuccess; }
cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned int) { std::memcpy(p, &h, sizeof(*p)); return cudaSuccess; }
cudaError_t cudaIpcCloseMemHandle(void*) { return cudaSuccess; }
}

namespace b200 {
struct CudaError { cudaError_t code; };
#define CUDA_TRY(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) throw CudaError{e_}; } while (0)
struct MockPool {
   void* alloc(size_t bytes) { void* p = std::malloc(bytes); std::memset(p, 0xEE, bytes); return p; }    // garbage, not zeros
   void release(void* p) { std::free(p); }
} g_pool;
std::atomic<long> g_launches{0};
int update_tile_size(bool big) { return big ? 128 : 64; }
void configure_update_kernels() {}

/* UPD_EXPLICIT as gemm_dmma.cu defines it: region {front, k0, k1, c_lo} from xregs[tile.front]; tile (ti, tj) in
 * absolute T x T coordinates of the front; output columns [c_lo, n), rows r >= c, r < m:
 * L(r, c) -= sum_{j in [k0, k1)} L(r, j) * LD(c, j). */
void launch_update(Front* fronts, const MatTile* work, int nwork, UpdateMode mode, bool big, cudaStream_t s, int,
      const int4* xregs) {
   if (mode != UPD_EXPLICIT || !big) throw std::runtime_error("mock: unexpected update mode");
   ++g_launches;
   MS(s)->push([=] {
      const int T = 128;
      for (int w = 0; w < nwork; ++w) {
         const MatTile t = work[w];
         const int4 xr = xregs[t.front];
         const Front& f = fronts[xr.x];
         const int k0 = xr.y, k1 = xr.z, c_lo = xr.w, c_hi = f.n;
         if (k1 <= k0 || c_lo >= c_hi) continue;
         const int r0 = t.ti * T, c0 = t.tj * T;
         if (c0 + T <= c_lo || c0 >= c_hi || r0 >= f.m) continue;
         for (int c = std::max(c0, c_lo); c < std::min(c0 + T, c_hi); ++c)
            for (int r = std::max(r0, c); r < std::min(r0 + T, f.m); ++r) {
               double sum = 0.0;
               for (int j = k0; j < k1; ++j) sum += f.L[r + (size_t)j * f.ldl] * f.LD[c + (size_t)j * f.ldl];
               f.L[r + (size_t)c * f.ldl] -= sum;
            }
      }
   });
}

#include "../../spral_b200/csrc/split_front.h"
} // namespace b200

using namespace b200;

/* the same update on a column range, for the owner's urgent update and the single-process reference */
static void update_cols(double* L, const double* LD, int ldl, int m, int n, int k0, int k1, int c_lo, int c_hi) {
   for (int c = c_lo; c < std::min(c_hi, n); ++c)
      for (int r = c; r < m; ++r) {
         double sum = 0.0;
         for (int j = k0; j < k1; ++j) sum += L[r + (size_t)j * ldl] * LD[c + (size_t)j * ldl];
         L[r + (size_t)c * ldl] -= sum;
      }
}
static void freeze_panel(const double* L, double* LD, int ldl, int m, int k0, int k1) {
   for (int j = k0; j < k1; ++j)
      for (int r = k0; r < m; ++r) LD[r + (size_t)j * ldl] = 0.5 * L[r + (size_t)j * ldl];
}

static int run_case(int m, int n, int fail_at, bool posdef_like, unsigned seed, const std::string& shm) {
   const int ldl = (m + 1) / 2 * 2;
   std::mt19937_64 rng(seed);
   std::uniform_real_distribution<double> U(-1.0, 1.0);
   std::vector<double> A((size_t)ldl * n, 0.0);
   for (int c = 0; c < n; ++c) for (int r = c; r < m; ++r) A[r + (size_t)c * ldl] = U(rng) / 16.0;

   /* ---- single-process reference ---- */
   std::vector<double> Lr(A), LDr((size_t)ldl * n, 0.0);
   for (int k0 = 0; k0 < n; k0 += PW) {
      const int k1 = std::min(k0 + PW, n);
      if (posdef_like) { /* L*D == L */ }
      else freeze_panel(Lr.data(), LDr.data(), ldl, m, k0, k1);
      update_cols(Lr.data(), posdef_like ? Lr.data() : LDr.data(), ldl, m, n, k0, k1, k1, n);
   }

   /* ---- split run: owner thread (this one) + helper thread ---- */
   std::vector<double> L(A), LD((size_t)ldl * n, 0.0);
   int helper_rc = -99;
   std::thread helper([&] {
      try { helper_rc = split_helper_serve(shm.c_str(), 0, 10.0, 0); }
      catch (const std::exception& e) { printf("helper: %s\n", e.what()); helper_rc = -1; }
   });
   cudaStream_t s = nullptr, s2 = nullptr;
   cudaStreamCreateWithFlags(&s, 0); cudaStreamCreateWithFlags(&s2, 0);
   SplitOwner* sp = SplitOwner::create(shm.c_str(), 1);
   if (!sp) throw std::runtime_error("cannot create the shared-memory segment");
   sp->timeout_s = 10.0;
   Front f;
   std::memset(&f, 0, sizeof(f));
   f.L = L.data(); f.LD = posdef_like ? L.data() : LD.data(); f.ldl = ldl; f.m = m; f.n = n;
   const bool started = sp->begin_front(f, posdef_like, s);
   int pushed = 0;
   for (int k = 0, k0 = 0; k0 < n; k0 += PW, ++k) {
      const int k1 = std::min(k0 + PW, n);
      /* "the panel kernels" on the main stream, then the snapshot synchronisation */
      double* Lp = L.data(); double* LDp = LD.data();
      if (!posdef_like) MS(s)->push([=] { freeze_panel(Lp, LDp, ldl, m, k0, k1); });
      cudaStreamSynchronize(s);
      const double* B = posdef_like ? L.data() : LD.data();
      if (sp->active && k == fail_at) sp->drain(k, k + 1, s, s2);            // "a pivot failed in this panel"
      if (sp->active) {
         sp->need_block(k + 1, s);
         MS(s)->push([=] { update_cols(Lp, B, ldl, m, n, k0, k1, k1, k1 + PW); });     // urgent: the next block only
         if (sp->has_far(k)) { sp->push_panel(k, k0, k1, s2); ++pushed; }
         else sp->end_front(s);
      } else {
         MS(s)->push([=] { update_cols(Lp, B, ldl, m, n, k0, k1, k1, n); });            // alone: everything to the right
      }
   }
   cudaStreamSynchronize(s); cudaStreamSynchronize(s2);
   if (sp->active) throw std::runtime_error("split still active at the end of the front");
   delete sp;                                                                            // phase 4
   helper.join();
   cudaStreamDestroy(s); cudaStreamDestroy(s2);

   long bad = 0;
   for (int c = 0; c < n; ++c) for (int r = c; r < m; ++r) if (L[r + (size_t)c * ldl] != Lr[r + (size_t)c * ldl]) ++bad;
   const bool expect_split = n >= 4 * PW;
   int rc = 0;
   if (bad) rc = 1;
   if (started != expect_split) rc = 1;
   if (expect_split && helper_rc != 0) rc = 1;
   printf("  m %5d n %5d fail_at %3d %s: split %s, %d panels pushed, helper rc %d, %ld entries differ%s\n", m, n, fail_at,
          posdef_like ? "L*D==L" : "L*D   ", started ? "yes" : "no ", pushed, helper_rc, bad, rc ? "   <-- FAIL" : "");
   return rc;
}

int main() {
   const std::string shm = "/spral_b200_split_emu_" + std::to_string((long)getpid());
   int failures = 0, cases = 0;
   struct C { int m, n, fail_at; bool pd; };
   const C cs[] = {
      {1500, 1500, -1, false},   /* square root front, no failure: natural end of the split */
      {1700, 1300, -1, false},   /* with contribution rows, n not a multiple of the block */
      {1290, 1290, -1, true},    /* Cholesky flavour (L*D == L): one mirror */
      {1500, 1500, 0, false},    /* failure in the first panel: nothing was updated remotely */
      {1500, 1500, 1, false},
      {1801, 1793, 3, false},    /* odd sizes, failure in the middle */
      {1600, 1536, 4, false},    /* failure at the last panel that still has far columns */
      {1600, 1536, 5, false},    /* ... after the split ended by itself */
      {1100, 1025, -1, false},   /* four blocks and one column */
      {900, 800, -1, false},     /* too small: never split (the helper sees the part end) */
   };
   for (const C& c : cs) {
      try { failures += run_case(c.m, c.n, c.fail_at, c.pd, 77 + cases, shm); }
      catch (const std::exception& e) { printf("  case %d: %s   <-- FAIL\n", cases, e.what()); ++failures; }
      ++cases;
   }
   printf("split_front_emu: %d cases, %d failures\n", cases, failures);
   return failures ? 1 : 0;
}
