/* TEST INFRASTRUCTURE (CPU only): runs a CUDA kernel BODY that is written against a small
 * context (tid / sync / shfl / atomic_add) on host threads: one pthread per CUDA thread of
 * a CTA, __syncthreads and the warp shuffles emulated with pthread barriers.  CTAs of a
 * launch are run one after the other. */
#pragma once
#include <pthread.h>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <mutex>
#include <vector>

namespace emu {

struct Block {
   int nthreads;
   pthread_barrier_t bar;
   struct Warp { pthread_barrier_t bar; double xd[32]; int xi[32]; };
   std::vector<Warp> warps;
   std::mutex atomics;
   explicit Block(int nt) : nthreads(nt), warps((nt + 31) / 32) {
      pthread_barrier_init(&bar, nullptr, nt);
      for (size_t w = 0; w < warps.size(); ++w) {
         int cnt = (int)std::min<size_t>(32, nt - 32 * w);
         pthread_barrier_init(&warps[w].bar, nullptr, cnt);
      }
   }
   ~Block() {
      pthread_barrier_destroy(&bar);
      for (auto& w : warps) pthread_barrier_destroy(&w.bar);
   }
};

struct Ctx {
   Block* b; int t;
   int tid() const { return t; }
   void sync() { pthread_barrier_wait(&b->bar); }
   void sync_warp() { pthread_barrier_wait(&b->warps[t >> 5].bar); }
   double shfl(double v, int src) {
      auto& w = b->warps[t >> 5];
      w.xd[t & 31] = v; pthread_barrier_wait(&w.bar);
      double r = w.xd[src & 31]; pthread_barrier_wait(&w.bar);
      return r;
   }
   double shfl_xor(double v, int off) {
      auto& w = b->warps[t >> 5];
      w.xd[t & 31] = v; pthread_barrier_wait(&w.bar);
      double r = w.xd[(t & 31) ^ off]; pthread_barrier_wait(&w.bar);
      return r;
   }
   int shfl_xor(int v, int off) {
      auto& w = b->warps[t >> 5];
      w.xi[t & 31] = v; pthread_barrier_wait(&w.bar);
      int r = w.xi[(t & 31) ^ off]; pthread_barrier_wait(&w.bar);
      return r;
   }
   /* mma.sync.m8n8k4 (FP64): lane l holds A[l/4][l%4], B[l%4][l/4], D[l/4][2(l%4) .. +1] */
   void mma(double& d0, double& d1, double a, double bb) {
      const int lane = t & 31, i = lane >> 2, j0 = 2 * (lane & 3);
      for (int k = 0; k < 4; ++k) {
         const double av = shfl(a, i * 4 + k), b0 = shfl(bb, j0 * 4 + k), b1 = shfl(bb, (j0 + 1) * 4 + k);
         d0 += av * b0; d1 += av * b1;
      }
   }
   void atomic_add(double* p, double v) { std::lock_guard<std::mutex> lock(b->atomics); *p += v; }
};

/* one CTA: body(ctx) on nthreads host threads */
inline void run_cta(int nthreads, const std::function<void(Ctx&)>& body) {
   Block blk(nthreads);
   struct Arg { Block* b; int t; const std::function<void(Ctx&)>* body; };
   std::vector<Arg> args(nthreads);
   std::vector<pthread_t> th(nthreads);
   pthread_attr_t at; pthread_attr_init(&at); pthread_attr_setstacksize(&at, 512 << 10);
   auto entry = [](void* p) -> void* { Arg* a = (Arg*)p; Ctx cx{a->b, a->t}; (*a->body)(cx); return nullptr; };
   for (int t = 0; t < nthreads; ++t) {
      args[t] = Arg{&blk, t, &body};
      if (pthread_create(&th[t], &at, entry, &args[t]) != 0) { perror("pthread_create"); exit(2); }
   }
   for (int t = 0; t < nthreads; ++t) pthread_join(th[t], nullptr);
   pthread_attr_destroy(&at);
}

} // namespace emu
