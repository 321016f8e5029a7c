/* FP64 tensor-core (DMMA) symmetric update kernels of the B200 SSIDS engine.
 *
 * One kernel template serves the three dense contractions of a front
 *   UPD_INNER    A(r,c) -= L(r,K) * LD(c,K)^T   K = columns of the current block
 *                column, c restricted to the open outer panel
 *   UPD_OUTER    same with K = all columns eliminated in the panel, c right of it
 *   UPD_CONTRIB  C(r,c)  = -L(r,K) * LD(c,K)^T  K = all eliminated columns,
 *                r,c >= n: the Schur complement / contribution block
 * (lower triangle only, r >= c).  It replaces cu_multisyrk_lc_r4x4 /
 * cu_multisyrk_r4x4 / cu_syrk_r4x4 (src/ssids/gpu/kernels/syrk.cu:180-600) and
 * the cublasDgemm calls of src/ssids/gpu/dense_factor.f90:146,672,785,1032;
 * the mathematics is form_contrib / update of the reference CPU engine
 * (src/ssids/cpu/kernels/ldlt_app.cxx:1082-1186).
 *
 * sm_100a design: tcgen05 has no FP64 kind, so the FP64 tensor path is
 * mma.sync.m8n8k4 (SASS DMMA).  Operand tiles are staged global -> shared by
 * the TMA engine with 1-D bulk async copies (cp.async.bulk, SASS UBLKCP): the
 * fronts are column-major, so a T x BK operand tile is BK contiguous column
 * segments, each one bulk copy completing on an mbarrier.  Shared tiles are
 * k-major with a row stride of T+4 doubles, which makes the 8x4 DMMA fragment
 * loads bank-conflict free.  A 4-stage mbarrier pipeline keeps the copies
 * ahead of the math.  The accumulator tile is computed transposed (D = B A^T)
 * so that each thread owns two consecutive ROWS of a column and the epilogue
 * uses 16-byte accesses on the column-major output.
 *
 * Tiles are aligned to absolute front coordinates (multiples of T from row 0),
 * which keeps every bulk copy 16-byte aligned whatever the pivoting state is.
 */
#include <algorithm>
#include <cstdlib>
#include "engine.h"
#include "device_utils.cuh"

namespace b200 {
#ifndef COUNT_LAUNCH
#define COUNT_LAUNCH() (void)g_launches.fetch_add(1, std::memory_order_relaxed)
#endif

namespace {

constexpr int BK = 16;       // K columns per pipeline stage
constexpr int NSTAGE = 4;
#ifndef WS_NS
#define WS_NS 3
#endif
#ifndef WS_BK
#define WS_BK 32
#endif


__device__ __forceinline__ uint32_t smem_u32(const void* p) {
   return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
   asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
   asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
   asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra.uni WAIT_DONE;\n"
      "bra.uni WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
   asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
/* TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier */
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
   asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
   asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

struct Region {
   const double* A; const double* B; double* C;
   size_t lda, ldb, ldc;
   int rows_alloc;        // rows that may be read per column of A/B (ldl, even)
   int m;                 // row limit of the output
   int c_lo, c_hi;        // output columns
   int k0, k1;            // contraction range (columns of A and B)
   int tj_base;           // absolute tile column of relative tile 0
   bool accumulate;       // C -= ... (true) or C = -... (false)
   bool valid;
};

__device__ __forceinline__ Region make_region(const Front* f, int mode, int T) {
   Region g;
   g.valid = false;
   g.A = f->L; g.B = f->LD; g.lda = g.ldb = (size_t)f->ldl; g.rows_alloc = f->ldl; g.m = f->m;
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
      const int cw = 128;                       // panel_v2.h: CW
      g.k0 = f->done; g.k1 = f->done + cw;
      g.c_lo = f->done + cw; g.c_hi = f->pend0;
      g.C = f->L; g.ldc = g.lda; g.accumulate = true;
   } else {
      if (f->m == f->n || !f->C) return g;
      g.k0 = 0; g.k1 = f->nelim;
      g.c_lo = f->n; g.c_hi = f->m;
      g.ldc = (size_t)f->ldc;
      g.C = f->C - (ptrdiff_t)f->n - (ptrdiff_t)f->n * (ptrdiff_t)g.ldc;   // absolute front coordinates
      g.accumulate = false;
   }
   if (g.c_lo >= g.c_hi) return g;
   g.tj_base = g.c_lo / T;
   g.valid = true;
   return g;
}

/* Epilogue of one warp: thread holds rows r, r+1 of column c for NC x NR
 * 8x8 sub-tiles.  The loads of column batch j + 1 are issued before the stores of batch j (the compiler keeps a
 * load behind a possibly-aliasing store, so batch after batch would pay one memory latency each). */
template <int NR>
__device__ __forceinline__ void load_old(const Region& g, double2 (&old)[NR], int c, int rfirst, int lane) {
   #pragma unroll
   for (int i = 0; i < NR; ++i) { old[i].x = 0.0; old[i].y = 0.0; }
   if (!g.accumulate || c < g.c_lo || c >= g.c_hi) return;
   const double* Cc = g.C + (ptrdiff_t)c * (ptrdiff_t)g.ldc;
   #pragma unroll
   for (int i = 0; i < NR; ++i) {
      const int r = rfirst + i * 8 + 2 * (lane & 3);
      const bool v0 = (r >= c) && (r < g.m);
      const bool v1 = (r + 1 >= c) && (r + 1 < g.m);
      const double* p = Cc + r;
      if (v0 && v1 && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) old[i] = *reinterpret_cast<const double2*>(p);
      else { if (v0) old[i].x = p[0]; if (v1) old[i].y = p[1]; }
   }
}
template <int NC, int NR>
__device__ __forceinline__ void store_tile(const Region& g, const double (&acc)[NC][NR][2],
      int r0, int c0, int rbase, int cbase, int lane) {
   double2 old[2][NR];
   load_old<NR>(g, old[0], c0 + cbase + (lane >> 2), r0 + rbase, lane);
   #pragma unroll
   for (int j = 0; j < NC; ++j) {
      const int c = c0 + cbase + j * 8 + (lane >> 2);
      if (j + 1 < NC) load_old<NR>(g, old[(j + 1) & 1], c + 8, r0 + rbase, lane);
      if (c < g.c_lo || c >= g.c_hi) continue;
      double* Cc = g.C + (ptrdiff_t)c * (ptrdiff_t)g.ldc;
      #pragma unroll
      for (int i = 0; i < NR; ++i) {
         const int r = r0 + rbase + i * 8 + 2 * (lane & 3);
         const bool v0 = (r >= c) && (r < g.m);
         const bool v1 = (r + 1 >= c) && (r + 1 < g.m);
         double* p = Cc + r;
         double2 o;
         o.x = old[j & 1][i].x - acc[j][i][0]; o.y = old[j & 1][i].y - acc[j][i][1];
         if (v0 && v1 && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) *reinterpret_cast<double2*>(p) = o;
         else { if (v0) p[0] = o.x; if (v1) p[1] = o.y; }
      }
   }
}

/* L2 prefetch of the warp's part of the output tile (accumulate mode): issued
 * when the tile starts so that the epilogue's loads hit the L2. */
template <int NC, int NR>
__device__ __forceinline__ void prefetch_tile(const Region& g, int r0, int c0, int rbase, int cbase, int lane) {
   if (!g.accumulate) return;
   #pragma unroll
   for (int j = 0; j < NC; ++j) {
      const int c = c0 + cbase + j * 8 + (lane >> 2);
      if (c < g.c_lo || c >= g.c_hi) continue;
      const double* Cc = g.C + (ptrdiff_t)c * (ptrdiff_t)g.ldc;
      #pragma unroll
      for (int i = 0; i < NR; i += 2) {      // 8 rows of a column = 64 B: two sub-tiles share a 128-B line
         const int r = r0 + rbase + i * 8;
         if ((lane & 3) == 0 && r < g.m) asm volatile("prefetch.global.L2 [%0];" :: "l"(Cc + r));
      }
   }
}

/* What one CTA needs to know about one output tile. */
struct TileJob {
   Region g;
   int r0, c0;
   int nchunk;
};

/* UPD_EXPLICIT: the region does not come from the front's (moving) pivoting
 * state but from a table written by the host: {front, k0, k1, c_lo}.  Used by
 * the look-ahead bulk update, which runs while the next panel advances the state. */
__device__ __forceinline__ Region explicit_region(const Front* fronts, const int4 xr) {
   const Front* f = &fronts[xr.x];
   Region g;
   g.A = f->L; g.B = f->LD; g.C = f->L;
   g.lda = g.ldb = g.ldc = (size_t)f->ldl; g.rows_alloc = f->ldl; g.m = f->m;
   g.k0 = xr.y; g.k1 = xr.z; g.c_lo = xr.w; g.c_hi = f->n;
   g.tj_base = 0; g.accumulate = true;
   g.valid = (g.k1 > g.k0) && (g.c_lo < g.c_hi);
   return g;
}

template <int T, int BKT = BK>
__device__ __forceinline__ bool load_job(const Front* fronts, const MatTile* work, int item, int mode, TileJob& j,
      const int4* xregs = nullptr) {
   MatTile w = work[item];
   if (mode == UPD_EXPLICIT) j.g = explicit_region(fronts, xregs[w.front]);
   else j.g = make_region(&fronts[w.front], mode, T);
   if (!j.g.valid) return false;
   j.r0 = w.ti * T; j.c0 = w.tj * T;          // absolute tile coordinates of the front
   if (j.c0 + T <= j.g.c_lo) return false;
   if (j.c0 >= j.g.c_hi || j.r0 >= j.g.m) return false;
   j.nchunk = (j.g.k1 - j.g.k0 + BKT - 1) / BKT;
   return true;
}

/* Persistent kernel: CTA b processes work items b, b+grid, ... ; T x T output
 * tile per item, NWR x NWC warps, NS pipeline stages.  The TMA producer (warp
 * 0) runs NS-1 K-chunks ahead of the math ACROSS tile boundaries, so the
 * operand loads of the next tile are in flight while the current tile's
 * epilogue (read-modify-write of C) drains. */
template <int T, int NWR, int NWC, int NS>
__global__ void __launch_bounds__(NWR * NWC * 32, 1)
k_update(Front* fronts, const MatTile* work, int nwork, int mode) {
   constexpr int LDS = T + 4;
   constexpr int WTR = T / NWR, WTC = T / NWC;
   constexpr int NR = WTR / 8, NC = WTC / 8;
   constexpr int STAGE_DOUBLES = 2 * BK * LDS;

   extern __shared__ __align__(128) unsigned char smem_raw[];
   double* tiles = reinterpret_cast<double*>(smem_raw);
   uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)NS * STAGE_DOUBLES * sizeof(double));

   const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
   const int wr = warp % NWR, wc = warp / NWR;
   const int rbase = wr * WTR, cbase = wc * WTC;

   if (tid == 0) {
      for (int s = 0; s < NS; ++s) mbar_init(&full[s], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
   }
   __syncthreads();

   /* producer cursor (used by warp 0 only): next (item, chunk) to load */
   int p_item = blockIdx.x - gridDim.x, p_chunk = 0, p_g = 0;   // p_g: global chunk counter
   TileJob pj; pj.nchunk = 0;
   auto produce_one = [&]() -> bool {      // issues the loads of one K chunk; false when out of work
      while (p_chunk >= pj.nchunk) {
         p_item += gridDim.x; p_chunk = 0; pj.nchunk = 0;
         if (p_item >= nwork) return false;
         if (!load_job<T>(fronts, work, p_item, mode, pj)) pj.nchunk = 0;
      }
      const Region& g = pj.g;
      const int klen = g.k1 - g.k0;
      const int s = p_g % NS;
      const int kc = min(BK, klen - p_chunk * BK);
      const int rowsA = min(T, g.rows_alloc - pj.r0);   // even
      const int rowsB = min(T, g.rows_alloc - pj.c0);
      double* st = tiles + (size_t)s * STAGE_DOUBLES;
      if (lane == 0) mbar_expect_tx(&full[s], (uint32_t)(kc * (rowsA + rowsB) * sizeof(double)));
      __syncwarp();
      const double* Ag = g.A + pj.r0 + (size_t)(g.k0 + p_chunk * BK) * g.lda;
      const double* Bg = g.B + pj.c0 + (size_t)(g.k0 + p_chunk * BK) * g.ldb;
      for (int idx = lane; idx < 2 * kc; idx += 32) {
         int op = idx >= kc;
         int col = idx - op * kc;
         if (op) bulk_g2s(st + BK * LDS + col * LDS, Bg + (size_t)col * g.ldb, rowsB * sizeof(double), &full[s]);
         else    bulk_g2s(st + col * LDS, Ag + (size_t)col * g.lda, rowsA * sizeof(double), &full[s]);
      }
      const int kc4 = (kc + 3) & ~3;
      if (kc4 != kc) {                 // zero the K tail of both operands
         for (int col = kc; col < kc4; ++col)
            for (int i = lane; i < T; i += 32) { st[col * LDS + i] = 0.0; st[BK * LDS + col * LDS + i] = 0.0; }
      }
      ++p_chunk; ++p_g;
      return true;
   };

   if (warp == 0) {
      for (int q = 0; q < NS - 1; ++q) if (!produce_one()) break;
   }
   __syncthreads();

   int c_g = 0;                          // consumer's global chunk counter
   for (int item = blockIdx.x; item < nwork; item += gridDim.x) {
      TileJob cj;
      if (!load_job<T>(fronts, work, item, mode, cj)) continue;
      const Region& g = cj.g;
      const int r0 = cj.r0, c0 = cj.c0;
      /* a warp whose sub-tile lies strictly above the diagonal has nothing to do */
      const bool warp_active = (r0 + rbase + WTR > c0 + cbase) && (r0 + rbase < g.m)
                               && (c0 + cbase < g.c_hi) && (c0 + cbase + WTC > g.c_lo);
      const int klen = g.k1 - g.k0;

      double acc[NC][NR][2];
      #pragma unroll
      for (int j = 0; j < NC; ++j)
         #pragma unroll
         for (int i = 0; i < NR; ++i) { acc[j][i][0] = 0.0; acc[j][i][1] = 0.0; }

      for (int chunk = 0; chunk < cj.nchunk; ++chunk, ++c_g) {
         if (warp == 0) produce_one();            // keeps NS-1 chunks in flight (may belong to the next tile)
         const int s = c_g % NS;
         mbar_wait(&full[s], (uint32_t)((c_g / NS) & 1));
         if (warp_active) {
            const int kc4 = (min(BK, klen - chunk * BK) + 3) & ~3;
            const double* As = tiles + (size_t)s * STAGE_DOUBLES + (lane & 3) * LDS + rbase + (lane >> 2);
            const double* Bs = As + BK * LDS - rbase + cbase;
            #pragma unroll
            for (int kk = 0; kk < BK; kk += 4) {
               if (kk < kc4) {
                  double af[NR], bf[NC];
                  #pragma unroll
                  for (int i = 0; i < NR; ++i) af[i] = As[kk * LDS + i * 8];
                  #pragma unroll
                  for (int j = 0; j < NC; ++j) bf[j] = Bs[kk * LDS + j * 8];
                  #pragma unroll
                  for (int j = 0; j < NC; ++j)
                     #pragma unroll
                     for (int i = 0; i < NR; ++i)
                        dmma(acc[j][i][0], acc[j][i][1], bf[j], af[i]);
               }
            }
         }
         __syncthreads();   // everyone is done with stage s before it is refilled
      }

      if (!warp_active) continue;
      store_tile<NC, NR>(g, acc, r0, c0, rbase, cbase, lane);
   }
}

/* The DMMAs of one K chunk of a warp tile.  A full chunk runs without guards and with the fragments of the next
 * 4-column step loaded (second register set) while the DMMAs of the current step issue: with one load batch per
 * step in front of its DMMAs every step starts with a shared-memory latency bubble on the tensor pipe. */
template <int NC, int NR, int BKT, int LDA_S, int LDB_S>
__device__ __forceinline__ void mma_chunk(double (&acc)[NC][NR][2], const double* __restrict__ As,
      const double* __restrict__ Bs, int kc4) {
   if (kc4 == BKT) {
      double af[2][NR], bf[2][NC];
      #pragma unroll
      for (int i = 0; i < NR; ++i) af[0][i] = As[i * 8];
      #pragma unroll
      for (int j = 0; j < NC; ++j) bf[0][j] = Bs[j * 8];
      #pragma unroll
      for (int kk = 0; kk < BKT; kk += 4) {
         const int cur = (kk >> 2) & 1, nxt = cur ^ 1;
         if (kk + 4 < BKT) {
            #pragma unroll
            for (int i = 0; i < NR; ++i) af[nxt][i] = As[(kk + 4) * LDA_S + i * 8];
            #pragma unroll
            for (int j = 0; j < NC; ++j) bf[nxt][j] = Bs[(kk + 4) * LDB_S + j * 8];
         }
         #pragma unroll
         for (int j = 0; j < NC; ++j)
            #pragma unroll
            for (int i = 0; i < NR; ++i)
               dmma(acc[j][i][0], acc[j][i][1], bf[cur][j], af[cur][i]);
      }
   } else {
      #pragma unroll 1
      for (int kk = 0; kk < kc4; kk += 4) {
         double af[NR], bf[NC];
         #pragma unroll
         for (int i = 0; i < NR; ++i) af[i] = As[kk * LDA_S + i * 8];
         #pragma unroll
         for (int j = 0; j < NC; ++j) bf[j] = Bs[kk * LDB_S + j * 8];
         #pragma unroll
         for (int j = 0; j < NC; ++j)
            #pragma unroll
            for (int i = 0; i < NR; ++i)
               dmma(acc[j][i][0], acc[j][i][1], bf[j], af[i]);
      }
   }
}

/* Warp-specialised variant for the large tiles: NWR x NWC consumer warps issue
 * the DMMAs, one extra producer warp drives the TMA bulk copies.  Stages are
 * handed back and forth with full[] / empty[] mbarriers, so there is no
 * block-wide barrier in the main loop: a consumer warp only waits for data, the
 * producer only for the slowest consumer of the stage it wants to refill, and
 * it keeps loading across tile boundaries while the consumers run the epilogue. */
template <int T, int NWR, int NWC, int NS, int BKT>
__global__ void __launch_bounds__((NWR * NWC + 4) * 32, 1)
k_update_ws(Front* fronts, const MatTile* work, int nwork, int mode, const int4* xregs) {
   constexpr int LDS = T + 4;
   constexpr int WTR = T / NWR, WTC = T / NWC;
   constexpr int NR = WTR / 8, NC = WTC / 8;
   constexpr int NCONS = NWR * NWC;
   constexpr int STAGE_DOUBLES = 2 * BKT * LDS;

   extern __shared__ __align__(128) unsigned char smem_raw[];
   double* tiles = reinterpret_cast<double*>(smem_raw);
   uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)NS * STAGE_DOUBLES * sizeof(double));
   uint64_t* empty = full + NS;

   const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
   if (tid == 0) {
      for (int s = 0; s < NS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], NCONS); }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
   }
   __syncthreads();

   /* Registers: the block is three warpgroups (two of consumers, one that holds the producer warp), compiled for
    * 168 registers a thread; the producer group hands most of its share to the consumers (64 accumulators = 128
    * registers, two sets of operand fragments = 48), 128 x 40 + 256 x 232 <= 65536. */
   if (warp >= NCONS) {
      asm volatile("setmaxnreg.dec.sync.aligned.u32 40;\n");
      if (warp > NCONS) return;
      /* ---------------- producer warp ---------------- */
      int g = 0;
      for (int item = blockIdx.x; item < nwork; item += gridDim.x) {
         TileJob pj;
         if (!load_job<T, BKT>(fronts, work, item, mode, pj, xregs)) continue;
         const Region& rg = pj.g;
         const int klen = rg.k1 - rg.k0;
         const int rowsA = min(T, rg.rows_alloc - pj.r0);   // even
         const int rowsB = min(T, rg.rows_alloc - pj.c0);
         for (int chunk = 0; chunk < pj.nchunk; ++chunk, ++g) {
            const int s = g % NS;
            if (g >= NS) mbar_wait(&empty[s], (uint32_t)(((g / NS) & 1) ^ 1));
            const int kc = min(BKT, klen - chunk * BKT);
            double* st = tiles + (size_t)s * STAGE_DOUBLES;
            const int kc4 = (kc + 3) & ~3;
            if (kc4 != kc) {              // zero the K tail of both operands
               for (int col = kc; col < kc4; ++col)
                  for (int i = lane; i < T; i += 32) { st[col * LDS + i] = 0.0; st[BKT * LDS + col * LDS + i] = 0.0; }
            }
            __syncwarp();
            if (lane == 0) mbar_expect_tx(&full[s], (uint32_t)(kc * (rowsA + rowsB) * sizeof(double)));
            __syncwarp();
            const double* Ag = rg.A + pj.r0 + (size_t)(rg.k0 + chunk * BKT) * rg.lda;
            const double* Bg = rg.B + pj.c0 + (size_t)(rg.k0 + chunk * BKT) * rg.ldb;
            for (int idx = lane; idx < 2 * kc; idx += 32) {
               int op = idx >= kc;
               int col = idx - op * kc;
               if (op) bulk_g2s(st + BKT * LDS + col * LDS, Bg + (size_t)col * rg.ldb, rowsB * sizeof(double), &full[s]);
               else    bulk_g2s(st + col * LDS, Ag + (size_t)col * rg.lda, rowsA * sizeof(double), &full[s]);
            }
         }
      }
      return;
   }

   /* ---------------- consumer warps ---------------- */
   asm volatile("setmaxnreg.inc.sync.aligned.u32 232;\n");
   const int wr = warp % NWR, wc = warp / NWR;
   const int rbase = wr * WTR, cbase = wc * WTC;
   int g = 0;
   for (int item = blockIdx.x; item < nwork; item += gridDim.x) {
      TileJob cj;
      if (!load_job<T, BKT>(fronts, work, item, mode, cj, xregs)) continue;
      const Region& rg = cj.g;
      const int r0 = cj.r0, c0 = cj.c0;
      const bool warp_active = (r0 + rbase + WTR > c0 + cbase) && (r0 + rbase < rg.m)
                               && (c0 + cbase < rg.c_hi) && (c0 + cbase + WTC > rg.c_lo);
      const int klen = rg.k1 - rg.k0;
      if (warp_active) prefetch_tile<NC, NR>(rg, r0, c0, rbase, cbase, lane);
      double acc[NC][NR][2];
      #pragma unroll
      for (int j = 0; j < NC; ++j)
         #pragma unroll
         for (int i = 0; i < NR; ++i) { acc[j][i][0] = 0.0; acc[j][i][1] = 0.0; }

      for (int chunk = 0; chunk < cj.nchunk; ++chunk, ++g) {
         const int s = g % NS;
         mbar_wait(&full[s], (uint32_t)((g / NS) & 1));
         if (warp_active) {
            const int kc4 = (min(BKT, klen - chunk * BKT) + 3) & ~3;
            const double* As = tiles + (size_t)s * STAGE_DOUBLES + (lane & 3) * LDS + rbase + (lane >> 2);
            const double* Bs = As + BKT * LDS - rbase + cbase;
            mma_chunk<NC, NR, BKT, LDS, LDS>(acc, As, Bs, kc4);
         }
         __syncwarp();
         if (lane == 0) mbar_arrive(&empty[s]);      // this warp is done with stage s
      }

      if (!warp_active) continue;
      store_tile<NC, NR>(rg, acc, r0, c0, rbase, cbase, lane);
   }
}

template <int T, int NS, int BKT = BK>
constexpr size_t update_smem_bytes() {
   return (size_t)NS * 2 * BKT * (T + 4) * sizeof(double) + 2 * NS * sizeof(uint64_t);
}

} // namespace

/* Tile sizes: the inner (K <= 32) updates always use 64 x 64 tiles with a
 * 2-stage pipeline (several CTAs per SM hide the latency of the short K loop);
 * outer / contribution updates use 128 x 128 tiles on large fronts. */
int update_tile_size(bool big_tiles) { return big_tiles ? 128 : 64; }
int inner_tile_size(bool big_tiles) { return (big_tiles && getenv("SPRAL_B200_INNER128")) ? 128 : 64; }

static int g_num_sms = 0;

void configure_update_kernels() {
   cudaFuncSetAttribute(k_update_ws<128, 2, 4, WS_NS, WS_BK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                        (int)update_smem_bytes<128, WS_NS, WS_BK>());
   cudaFuncSetAttribute(k_update<64, 2, 2, NSTAGE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                        (int)update_smem_bytes<64, NSTAGE>());
   cudaFuncSetAttribute(k_update<64, 2, 2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                        (int)update_smem_bytes<64, 2>());
   int dev = 0;
   cudaGetDevice(&dev);
   cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
}

int device_sm_count() { return g_num_sms > 0 ? g_num_sms : 148; }

void launch_update(Front* fronts, const MatTile* work, int nwork, UpdateMode mode,
      bool big_tiles, cudaStream_t s, int max_ctas, const int4* xregs) {
   if (nwork == 0) return;
   int sms = g_num_sms > 0 ? g_num_sms : 148;
   if (max_ctas > 0) sms = std::min(sms, max_ctas);
   if (mode == UPD_INNER && inner_tile_size(big_tiles) == 64) {
      int grid = std::min(nwork, sms * 5);
      k_update<64, 2, 2, 2><<<grid, 128, update_smem_bytes<64, 2>(), s>>>(fronts, work, nwork, (int)mode);
   } else if (big_tiles) {
      /* max_ctas < 0: one tile per CTA, not persistent (the bulk stream: kernels of the urgent stream get an SM whenever
       * a CTA retires; 2 .. 6 tiles per CTA measured no faster, tools/gpu_call_r02_31.sh) */
      int grid = max_ctas < 0 ? nwork : std::min(nwork, sms);
      k_update_ws<128, 2, 4, WS_NS, WS_BK><<<grid, 384, update_smem_bytes<128, WS_NS, WS_BK>(), s>>>(fronts, work, nwork, (int)mode, xregs);
   } else {
      int grid = std::min(nwork, sms * 3);
      k_update<64, 2, 2, NSTAGE><<<grid, 128, update_smem_bytes<64, NSTAGE>(), s>>>(fronts, work, nwork, (int)mode);
   }
   COUNT_LAUNCH();
}

} // namespace b200

/* FP64 tensor-pipe peak of the device this runs on: a register-resident DMMA
 * issue loop (no memory traffic).  Used by bench.py as the roofline
 * denominator because MEASURED_PEAKS.json holds no FP64 figure. */
namespace b200 {
__global__ void __launch_bounds__(256) k_dmma_peak(double* out, int iters) {
   double acc[16][2];
   for (int i = 0; i < 16; ++i) { acc[i][0] = threadIdx.x * 1e-9; acc[i][1] = i * 1e-9; }
   double a = 1.0 + threadIdx.x * 1e-12, b = 1.0 - threadIdx.x * 1e-12;
   for (int it = 0; it < iters; ++it) {
      #pragma unroll
      for (int i = 0; i < 16; ++i)
         asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                      : "+d"(acc[i][0]), "+d"(acc[i][1]) : "d"(a), "d"(b));
   }
   double s = 0;
   for (int i = 0; i < 16; ++i) s += acc[i][0] + acc[i][1];
   out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
} // namespace b200

extern "C" double spral_ssids_b200_fp64_peak_tflops(int device) {
   if (cudaSetDevice(device) != cudaSuccess) return -1.0;
   cudaDeviceProp p;
   if (cudaGetDeviceProperties(&p, device) != cudaSuccess) return -1.0;
   const int blocks = p.multiProcessorCount * 2, iters = 20000;
   double* out = nullptr;
   if (cudaMalloc(&out, (size_t)blocks * 256 * sizeof(double)) != cudaSuccess) return -1.0;
   cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
   double best = 0;
   b200::k_dmma_peak<<<blocks, 256>>>(out, 200);
   for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      b200::k_dmma_peak<<<blocks, 256>>>(out, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
      double tf = (double)blocks * 8 * iters * 16 * 512.0 / ms / 1e9;
      if (tf > best) best = tf;
   }
   cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
   return best;
}
