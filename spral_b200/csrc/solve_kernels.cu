/* Triangular solve kernels of the B200 SSIDS engine (sm_100a), batched over
 * the fronts of one tree level and over right-hand sides.
 *
 * They replace the per-RHS host loop and the kernels of the reference GPU path
 *   src/ssids/gpu/solve.f90:177-291 (fwd_solve_gpu), :17-56 (bwd), :116-155 (d_solve)
 *   src/ssids/gpu/kernels/solve.cu:593-735, src/ssids/gpu/kernels/dtrsv.h:462-813
 * and compute what the reference CPU engine computes node by node
 *   src/ssids/cpu/NumericSubtree.hxx:286-418 (gather, trsv/gemv or trsm/gemm, scatter)
 *   src/ssids/cpu/kernels/ldlt_app.cxx:2538-2589, cholesky.cxx:190-213.
 *
 * A front of m rows with nelim eliminated columns is processed in column
 * blocks of SB = 32.  Row i of the front is entry idx(i) of x: perm[i] for
 * i < n (eliminated and delayed columns), rows[n0 + i - n] below.
 *
 * Forward, step s (one launch per step, all fronts of the level):
 *   every CTA (front, 128-row tile) solves the 32x32 diagonal block of the
 *   step redundantly in shared memory, then subtracts L(tile, block) * y from
 *   its rows of x.  Rows >= nelim are shared with sibling fronts -> atomics.
 *   The final y is parked in `ywork` and flushed to x when the part is done
 *   (nobody reads an eliminated variable again during the forward sweep).
 * Backward, step s (two launches): partial L(tile, block)^T x(tile) per tile
 *   into a scratch, then one warp per front sums the partials in a fixed
 *   order and solves the transposed diagonal block.
 */
#include <vector>
#include <cstdio>
#include <cstdlib>
#include "engine.h"
#include "device_utils.cuh"
#include "solve_wide.h"
#include <math_constants.h>

namespace b200 {
#ifndef COUNT_LAUNCH
#define COUNT_LAUNCH() (void)g_launches.fetch_add(1, std::memory_order_relaxed)
#endif

namespace {

constexpr int SB = 32;    // solve block

/* Inside the sweeps the right-hand sides of a chunk are stored RHS-contiguous
 * (x^T: entry (g, k) at g*NR + k), so a row of the front touches one 8*NR-byte
 * segment instead of NR segments ldx apart (the ABI's column-major layout). */
#define XI(g, k) ((size_t)(g) * NR + (k))

__device__ __forceinline__ int row_index(const SolveFront& f, int i) {
   return (i < f.n ? f.perm[i] : f.rows[f.n0 + i - f.n]) - 1;
}

/* ---- forward ---------------------------------------------------------- */
/* Loads of data that other CTAs write during the same launch (cooperative
 * kernels only) must bypass the non-coherent L1. */
template <bool COH>
__device__ __forceinline__ double ld_shared_data(const double* p) { return COH ? __ldcg(p) : *p; }

/* Grid-wide barrier of a cooperative launch (all CTAs co-resident): `target` is
 * the value the arrival counter reaches when every CTA has arrived. */
__device__ __forceinline__ void grid_barrier(unsigned int* bar, unsigned int target) {
   __syncthreads();
   if (threadIdx.x == 0) {
      __threadfence();
      atomicAdd(bar, 1u);
      while (*(volatile unsigned int*)bar < target) { }
      __threadfence();
   }
   __syncthreads();
}

template <int NR, bool POSDEF, bool COH>
__device__ __forceinline__ void fwd_step_body(const SolveFront* fronts, const RowTile* work, int item, int step,
      double* __restrict__ x, double* __restrict__ ywork) {
   RowTile w = work[item];
   const SolveFront f = fronts[w.front];
   const int j0 = step * SB;
   if (j0 >= f.nelim) return;
   const int wd = min(SB, f.nelim - j0);
   const int r0 = w.tile * RT;
   if (r0 + RT <= j0 || r0 >= f.m) return;

   __shared__ double lkk[SB][SB + 1];
   __shared__ double xs[SB][NR];
   __shared__ int gidx[SB];
   const size_t ldl = f.ldl;
   const double* Lb = f.L + (size_t)j0 * ldl;     // block column
   for (int e = threadIdx.x; e < SB * SB; e += RT) {
      int i = e % SB, j = e / SB;
      lkk[i][j] = (i < wd && j < wd && i >= j) ? Lb[(size_t)(j0 + i) + j * ldl] : 0.0;
   }
   if (threadIdx.x < SB) {
      int g = (threadIdx.x < wd) ? f.perm[j0 + threadIdx.x] - 1 : -1;
      gidx[threadIdx.x] = g;
      #pragma unroll
      for (int k = 0; k < NR; ++k) xs[threadIdx.x][k] = (g >= 0) ? ld_shared_data<COH>(&x[XI(g, k)]) : 0.0;   // L2: other CTAs update x within the launch
   }
   __syncthreads();
   {  /* forward substitution: lanes are rows; every warp takes NR/4 right-hand sides and
       * advances them together (independent dependency chains hide the shuffle latency) */
      constexpr int NW = RT / 32;
      constexpr int NRW = (NR + NW - 1) / NW;
      const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
      double v[NRW];
      #pragma unroll
      for (int q = 0; q < NRW; ++q) { int k = warp * NRW + q; v[q] = (k < NR) ? xs[lane][k] : 0.0; }
      for (int j = 0; j < wd; ++j) {
         const double l = (lane > j && lane < wd) ? lkk[lane][j] : 0.0;
         const double dj = POSDEF ? lkk[j][j] : 1.0;
         #pragma unroll
         for (int q = 0; q < NRW; ++q) {
            double yj = __shfl_sync(0xffffffffu, v[q], j);
            if (POSDEF) { yj /= dj; if (lane == j) v[q] = yj; }
            v[q] -= l * yj;
         }
      }
      #pragma unroll
      for (int q = 0; q < NRW; ++q) { int k = warp * NRW + q; if (k < NR) xs[lane][k] = v[q]; }
   }
   __syncthreads();
   /* the CTA owning the block's first row publishes y */
   if (w.tile == j0 / RT && threadIdx.x < wd) {
      #pragma unroll
      for (int k = 0; k < NR; ++k) ywork[XI(gidx[threadIdx.x], k)] = xs[threadIdx.x][k];
   }
   const int r = r0 + threadIdx.x;
   if (r >= j0 + wd && r < f.m) {
      double acc[NR];
      #pragma unroll
      for (int k = 0; k < NR; ++k) acc[k] = 0.0;
      const double* Lr = Lb + r;
      for (int j = 0; j < wd; ++j) {
         double l = Lr[j * ldl];
         #pragma unroll
         for (int k = 0; k < NR; ++k) acc[k] += l * xs[j][k];
      }
      const int g = row_index(f, r);
      /* fire-and-forget reductions (RED.ADD.F64): no load latency on the critical
       * path; rows < nelim are touched by this thread only, rows >= nelim are
       * shared with sibling fronts */
      #pragma unroll
      for (int k = 0; k < NR; ++k) atomicAdd(&x[XI(g, k)], -acc[k]);
   }
}

template <int NR, bool POSDEF>
__global__ void __launch_bounds__(RT)
k_fwd_step(const SolveFront* fronts, const RowTile* work, int step,
      double* __restrict__ x, int ldx, double* __restrict__ ywork) {
   fwd_step_body<NR, POSDEF, false>(fronts, work, blockIdx.x, step, x, ywork);
}

/* All block steps of a level in ONE cooperative launch: a grid barrier replaces
 * the kernel boundary between dependent steps (the sweeps are bound by that
 * latency, not by bandwidth, on the large fronts at the top of the tree). */
template <int NR, bool POSDEF>
__global__ void __launch_bounds__(RT)
k_fwd_level_coop(const SolveFront* fronts, const RowTile* work, int nsteps,
      double* __restrict__ x, double* __restrict__ ywork, unsigned int* bar) {
   for (int step = 0; step < nsteps; ++step) {
      fwd_step_body<NR, POSDEF, true>(fronts, work, blockIdx.x, step, x, ywork);
      if (step + 1 < nsteps) grid_barrier(bar, (unsigned int)(step + 1) * gridDim.x);
   }
}

/* Right-hand sides handled per thread by the flush / diagonal kernels: a row of x is NR contiguous values; consecutive
 * threads take consecutive chunks of it (coalesced), every chunk is loaded completely before it is stored (a load behind
 * a store through the same pointer would wait for it). */
constexpr int RHS_CHUNK = 8;

/* x(eliminated variables) <- ywork.  Grid = fronts x column chunks (block b: front b % count, chunk b / count): the
 * large fronts are not left to one CTA. */
__global__ void __launch_bounds__(256)
k_fwd_flush(const SolveFront* fronts, int first, int count, int nrhs, double* __restrict__ x, int ldx,
      const double* __restrict__ ywork) {
   const int NR = nrhs;
   const SolveFront f = fronts[first + blockIdx.x % count];
   const int chunk = blockIdx.x / count, nchunk = gridDim.x / count;
   const int nkc = (nrhs + RHS_CHUNK - 1) / RHS_CHUNK;
   const int total = f.nelim * nkc;
   for (int idx = chunk * 256 + threadIdx.x; idx < total; idx += nchunk * 256) {
      const int j = idx / nkc, k0 = (idx % nkc) * RHS_CHUNK;
      const int g = f.perm[j] - 1;
      double v[RHS_CHUNK];
      #pragma unroll
      for (int q = 0; q < RHS_CHUNK; ++q) v[q] = (k0 + q < nrhs) ? ywork[XI(g, k0 + q)] : 0.0;
      #pragma unroll
      for (int q = 0; q < RHS_CHUNK; ++q) if (k0 + q < nrhs) x[XI(g, k0 + q)] = v[q];
   }
}

/* ---- diagonal --------------------------------------------------------- */
/* x <- D^-1 x on the eliminated variables; D^-1 is stored (2 per column,
 * second column of a 2x2 marked by +Inf), src/ssids/cpu/kernels/ldlt_app.cxx
 * ldlt_app_solve_diag :2553-2573.  Grid = fronts x column chunks. */
__global__ void __launch_bounds__(256)
k_diag_solve(const SolveFront* fronts, int first, int count, int nrhs, double* __restrict__ x, int ldx) {
   const int NR = nrhs;
   const SolveFront f = fronts[first + blockIdx.x % count];
   const double* d = f.D;
   const int chunk = blockIdx.x / count, nchunk = gridDim.x / count;
   const int nkc = (nrhs + RHS_CHUNK - 1) / RHS_CHUNK;
   const int total = f.nelim * nkc;
   for (int idx = chunk * 256 + threadIdx.x; idx < total; idx += nchunk * 256) {
      const int j = idx / nkc, k0 = (idx % nkc) * RHS_CHUNK;
      const double d11 = d[2 * j];
      if (isinf(d11)) continue;                 // handled by the first column of the pair
      const int g1 = f.perm[j] - 1;
      if (j + 1 < f.nelim && isinf(d[2 * j + 2])) {
         const double d21 = d[2 * j + 1], d22 = d[2 * j + 3];
         const int g2 = f.perm[j + 1] - 1;
         double x1[RHS_CHUNK], x2[RHS_CHUNK];
         #pragma unroll
         for (int q = 0; q < RHS_CHUNK; ++q) {
            x1[q] = (k0 + q < nrhs) ? x[XI(g1, k0 + q)] : 0.0;
            x2[q] = (k0 + q < nrhs) ? x[XI(g2, k0 + q)] : 0.0;
         }
         #pragma unroll
         for (int q = 0; q < RHS_CHUNK; ++q)
            if (k0 + q < nrhs) {
               x[XI(g1, k0 + q)] = d11 * x1[q] + d21 * x2[q];
               x[XI(g2, k0 + q)] = d21 * x1[q] + d22 * x2[q];
            }
      } else {
         double x1[RHS_CHUNK];
         #pragma unroll
         for (int q = 0; q < RHS_CHUNK; ++q) x1[q] = (k0 + q < nrhs) ? x[XI(g1, k0 + q)] : 0.0;
         #pragma unroll
         for (int q = 0; q < RHS_CHUNK; ++q) if (k0 + q < nrhs) x[XI(g1, k0 + q)] = x1[q] * d11;
      }
   }
}

/* ---- backward --------------------------------------------------------- */
/* block index handled by front f at backward step s (last block first) */
__device__ __forceinline__ int bwd_block(const SolveFront& f, int step) {
   int nblk = (f.nelim + SB - 1) / SB;
   return nblk - 1 - step;
}

/* partial(t, j, k) = sum over rows r of tile t below the block of L(r, j0+j) * x(r, k) */
template <int NR, bool COH>
__device__ __forceinline__ void bwd_reduce_body(const SolveFront* fronts, const RowTile* work, int item, int step,
      const double* __restrict__ x, double* __restrict__ pbuf) {
   RowTile w = work[item];
   const SolveFront f = fronts[w.front];
   const int b = bwd_block(f, step);
   if (b < 0) return;
   const int j0 = b * SB;
   const int wd = min(SB, f.nelim - j0);
   const int r0 = w.tile * RT;
   if (r0 + RT <= j0 + wd || r0 >= f.m) return;   // no row of this tile below the block

   extern __shared__ double smem_dyn[];
   double (*tile)[32][SB + 1] = reinterpret_cast<double (*)[32][SB + 1]>(smem_dyn);
   double (*xr)[NR] = reinterpret_cast<double (*)[NR]>(smem_dyn + (RT / 32) * 32 * (SB + 1));
   /* red aliases tile once the tile has been consumed (NR <= 32 < SB + 1) */
   double (*red)[32][SB + 1] = tile;
   const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
   const int r = r0 + threadIdx.x;
   const bool active = (r >= j0 + wd) && (r < f.m);
   const size_t ldl = f.ldl;
   {
      int g = active ? row_index(f, r) : 0;
      #pragma unroll
      for (int k = 0; k < NR; ++k) xr[threadIdx.x][k] = active ? ld_shared_data<COH>(&x[XI(g, k)]) : 0.0;
      const double* Lr = f.L + r + (size_t)j0 * ldl;
      for (int j = 0; j < SB; ++j) tile[warp][lane][j] = (active && j < wd) ? Lr[j * ldl] : 0.0;
   }
   __syncthreads();
   {
      double acc[NR];
      #pragma unroll
      for (int k = 0; k < NR; ++k) acc[k] = 0.0;
      for (int i = 0; i < 32; ++i) {
         double l = tile[warp][i][lane];
         #pragma unroll
         for (int k = 0; k < NR; ++k) acc[k] += l * xr[warp * 32 + i][k];
      }
      __syncthreads();
      #pragma unroll
      for (int k = 0; k < NR; ++k) red[warp][lane][k] = acc[k];
   }
   __syncthreads();
   if (threadIdx.x < SB) {
      double* out = pbuf + (size_t)item * SB * NR;
      #pragma unroll
      for (int k = 0; k < NR; ++k) {
         double s = 0.0;
         for (int q = 0; q < RT / 32; ++q) s += red[q][threadIdx.x][k];
         out[threadIdx.x * NR + k] = s;
      }
   }
}

/* one CTA per front: y = x_blk - sum_t partial(t) (fixed order), solve L_kk^T z = y;
 * lanes are the block's columns, the right-hand sides are dealt to the warps */
template <int NR>
__global__ void __launch_bounds__(RT)
k_bwd_reduce(const SolveFront* fronts, const RowTile* work, int step,
      const double* __restrict__ x, int ldx, double* __restrict__ pbuf) {
   bwd_reduce_body<NR, false>(fronts, work, blockIdx.x, step, x, pbuf);
}

template <int NR, bool POSDEF, bool COH>
__device__ __forceinline__ void bwd_diag_body(const SolveFront* fronts, int fi, const int* __restrict__ wbeg, int step,
      double* __restrict__ x, const double* __restrict__ pbuf) {
   const SolveFront f = fronts[fi];
   const int b = bwd_block(f, step);
   if (b < 0) return;
   const int j0 = b * SB;
   const int wd = min(SB, f.nelim - j0);
   const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
   __shared__ double lkk[SB][SB + 1];
   const size_t ldl = f.ldl;
   for (int e = threadIdx.x; e < SB * SB; e += RT) {
      int i = e % SB, j = e / SB;
      lkk[i][j] = (i < wd && j < wd && i >= j) ? f.L[(size_t)(j0 + i) + (size_t)(j0 + j) * ldl] : 0.0;
   }
   __syncthreads();
   const int g = (lane < wd) ? f.perm[j0 + lane] - 1 : -1;
   /* tiles holding rows below the block, in increasing order */
   const int ntile = (f.m + RT - 1) / RT;
   const int t0 = (j0 + wd) / RT;
   const double* pb = pbuf + (size_t)wbeg[fi] * SB * NR;
   constexpr int NW = RT / 32;
   constexpr int NRW = (NR + NW - 1) / NW;
   double v[NRW];
   #pragma unroll
   for (int q = 0; q < NRW; ++q) {
      int k = warp * NRW + q;
      v[q] = (g >= 0 && k < NR) ? ld_shared_data<COH>(&x[XI(g, k)]) : 0.0;
   }
   for (int t = t0; t < ntile; ++t) {
      #pragma unroll
      for (int q = 0; q < NRW; ++q) {
         int k = warp * NRW + q;
         if (k < NR) v[q] -= ld_shared_data<COH>(&pb[(size_t)t * SB * NR + lane * NR + k]);
      }
   }
   for (int j = wd - 1; j >= 0; --j) {
      const double l = (lane < j) ? lkk[j][lane] : 0.0;
      const double dj = POSDEF ? lkk[j][j] : 1.0;
      #pragma unroll
      for (int q = 0; q < NRW; ++q) {
         double zj = __shfl_sync(0xffffffffu, v[q], j);
         if (POSDEF) { zj /= dj; if (lane == j) v[q] = zj; }
         v[q] -= l * zj;
      }
   }
   #pragma unroll
   for (int q = 0; q < NRW; ++q) {
      int k = warp * NRW + q;
      if (g >= 0 && k < NR) x[XI(g, k)] = v[q];
   }
}

template <int NR, bool POSDEF>
__global__ void __launch_bounds__(RT)
k_bwd_diag(const SolveFront* fronts, int first, const int* __restrict__ wbeg, int step,
      double* __restrict__ x, int ldx, const double* __restrict__ pbuf) {
   bwd_diag_body<NR, POSDEF, false>(fronts, first + blockIdx.x, wbeg, step, x, pbuf);
}

/* Backward counterpart: per step the partial sums of every tile, a grid barrier,
 * the diagonal-block solve by the CTA that owns tile 0 of each front, a barrier. */
template <int NR, bool POSDEF>
__global__ void __launch_bounds__(RT)
k_bwd_level_coop(const SolveFront* fronts, const RowTile* work, const int* __restrict__ wbeg, int nsteps,
      double* __restrict__ x, double* __restrict__ pbuf, unsigned int* bar) {
   const RowTile w = work[blockIdx.x];
   unsigned int target = 0;
   for (int step = 0; step < nsteps; ++step) {
      bwd_reduce_body<NR, true>(fronts, work, blockIdx.x, step, x, pbuf);
      target += gridDim.x; grid_barrier(bar, target);
      if (w.tile == 0) bwd_diag_body<NR, POSDEF, true>(fronts, w.front, wbeg, step, x, pbuf);
      if (step + 1 < nsteps) { target += gridDim.x; grid_barrier(bar, target); }
   }
}

template <int NR>
constexpr size_t reduce_smem() { return ((size_t)(RT / 32) * 32 * (SB + 1) + (size_t)RT * NR) * sizeof(double); }

/* CTAs of a kernel that can be resident at once on the current device */
template <class K>
int coop_capacity(K kernel, size_t smem) {
   int nb = 0, dev = 0, sms = 0;
   cudaGetDevice(&dev);
   cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
   if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, RT, smem) != cudaSuccess) { cudaGetLastError(); return 0; }
   return nb * sms;
}

template <int NR, bool POSDEF>
void fwd_level_t(const SolveFront* fronts, const RowTile* work, int nwork, int nsteps,
      double* x, int ldx, double* ywork, unsigned int* bar, cudaStream_t s) {
   static int cap = -1;
   if (cap < 0) cap = coop_capacity(k_fwd_level_coop<NR, POSDEF>, 0);
   if (bar && nsteps > 1 && nwork <= cap) {
      cudaMemsetAsync(bar, 0, sizeof(unsigned int), s);
      void* args[] = {(void*)&fronts, (void*)&work, (void*)&nsteps, (void*)&x, (void*)&ywork, (void*)&bar};
      if (cudaLaunchCooperativeKernel((const void*)k_fwd_level_coop<NR, POSDEF>, dim3(nwork), dim3(RT), args, 0, s) == cudaSuccess) {
         COUNT_LAUNCH();
         return;
      }
      cudaGetLastError();      // fall back to one launch per step
   }
   for (int st = 0; st < nsteps; ++st) {
      k_fwd_step<NR, POSDEF><<<nwork, RT, 0, s>>>(fronts, work, st, x, ldx, ywork); COUNT_LAUNCH();
   }
}

template <int NR>
void fwd_level(const SolveFront* fronts, const RowTile* work, int nwork, int nsteps, bool posdef,
      double* x, int ldx, double* ywork, unsigned int* bar, cudaStream_t s) {
   if (posdef) fwd_level_t<NR, true>(fronts, work, nwork, nsteps, x, ldx, ywork, bar, s);
   else fwd_level_t<NR, false>(fronts, work, nwork, nsteps, x, ldx, ywork, bar, s);
}

template <int NR, bool POSDEF>
void bwd_level_t(const SolveFront* fronts, int first, int count, const RowTile* work, int nwork,
      const int* wbeg, int nsteps, double* x, int ldx, double* pbuf, unsigned int* bar, cudaStream_t s) {
   static int cap = -1;
   if (cap < 0) {
      cudaFuncSetAttribute(k_bwd_level_coop<NR, POSDEF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)reduce_smem<NR>());
      cap = coop_capacity(k_bwd_level_coop<NR, POSDEF>, reduce_smem<NR>());
   }
   if (bar && nsteps > 1 && nwork <= cap) {
      cudaMemsetAsync(bar, 0, sizeof(unsigned int), s);
      void* args[] = {(void*)&fronts, (void*)&work, (void*)&wbeg, (void*)&nsteps, (void*)&x, (void*)&pbuf, (void*)&bar};
      if (cudaLaunchCooperativeKernel((const void*)k_bwd_level_coop<NR, POSDEF>, dim3(nwork), dim3(RT), args,
                                      reduce_smem<NR>(), s) == cudaSuccess) {
         COUNT_LAUNCH();
         return;
      }
      cudaGetLastError();
   }
   for (int st = 0; st < nsteps; ++st) {
      k_bwd_reduce<NR><<<nwork, RT, reduce_smem<NR>(), s>>>(fronts, work, st, x, ldx, pbuf); COUNT_LAUNCH();
      k_bwd_diag<NR, POSDEF><<<count, RT, 0, s>>>(fronts, first, wbeg, st, x, ldx, pbuf); COUNT_LAUNCH();
   }
}

template <int NR>
void bwd_level(const SolveFront* fronts, int first, int count, const RowTile* work, int nwork,
      const int* wbeg, int nsteps, bool posdef, double* x, int ldx, double* pbuf, unsigned int* bar, cudaStream_t s) {
   if (posdef) bwd_level_t<NR, true>(fronts, first, count, work, nwork, wbeg, nsteps, x, ldx, pbuf, bar, s);
   else bwd_level_t<NR, false>(fronts, first, count, work, nwork, wbeg, nsteps, x, ldx, pbuf, bar, s);
}


/* ---- wide sweeps (solve_wide.h) ---------------------------------------- */
struct SolveDevCtx {
   __device__ __forceinline__ int tid() const { return threadIdx.x; }
   __device__ __forceinline__ void sync() { __syncthreads(); }
   __device__ __forceinline__ void sync_warp() { __syncwarp(); }
   __device__ __forceinline__ double shfl(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
   __device__ __forceinline__ double shfl_xor(double v, int off) { return __shfl_xor_sync(0xffffffffu, v, off); }
   __device__ __forceinline__ void atomic_add(double* p, double v) { atomicAdd(p, v); }
   __device__ __forceinline__ void mma(double& d0, double& d1, double a, double b) { pv_dmma(d0, d1, a, b); }
};

/* grid = fronts x (NRT / NR) slices of right-hand sides: block b handles front b % count, slice b / count */
template <int NR, int NRT, bool POSDEF, int HB = SWB>
__global__ void __launch_bounds__(HB < SW_TT ? HB : SW_TT)
k_fwd_wide_T(const SolveFront* fronts, int first, int count, int blk, const double* __restrict__ x, double* __restrict__ ywork) {
   extern __shared__ double smem_dyn[];
   const SolveFront f = fronts[first + blockIdx.x % count];
   const int k0 = (blockIdx.x / count) * NR;
   SolveDevCtx cx;
   fwd_wide_T_h<NR, NRT, POSDEF, HB>(cx, f, blk, x + k0, ywork + k0, smem_dyn);
}

template <int NR>
__global__ void __launch_bounds__(SW_GT)
k_fwd_wide_G(const SolveFront* fronts, const RowTile* work, int blk, double* __restrict__ x, const double* __restrict__ ywork,
      int part, int first) {
   extern __shared__ double smem_dyn[];
   const RowTile w = work[blockIdx.x / SW_FSPLIT];
   const SolveFront f = fronts[w.front];
   SolveDevCtx cx;
   fwd_wide_G<NR>(cx, f, w.tile, blk, (int)(blockIdx.x % SW_FSPLIT), x, ywork, smem_dyn, part);
}

template <int NR>
__global__ void __launch_bounds__(SW_GT)
k_fwd_wide_G_near(const SolveFront* fronts, int first, int blk, double* __restrict__ x, const double* __restrict__ ywork) {
   extern __shared__ double smem_dyn[];
   const SolveFront f = fronts[first + (int)blockIdx.x / (2 * SW_NSPLIT)];
   SolveDevCtx cx;
   fwd_wide_G<NR, SolveDevCtx, SW_NSPLIT>(cx, f, ((int)blockIdx.x / SW_NSPLIT) & 1, blk, (int)(blockIdx.x % SW_NSPLIT), x, ywork, smem_dyn, SW_NEAR);
}

constexpr int SW_BNSPLIT = 4;      // CTAs that share a tile in the near launches of the backward sweep and of the tensor-core G kernels

template <int NR>
__global__ void __launch_bounds__(SW_GT)
k_bwd_wide_G(const SolveFront* fronts, const RowTile* work, int first, int step, const double* __restrict__ x, double* __restrict__ pbuf,
      int part) {
   extern __shared__ double smem_dyn[];
   const RowTile w = part == SW_NEAR ? RowTile{first + (int)blockIdx.x / (2 * SW_BNSPLIT), ((int)blockIdx.x / SW_BNSPLIT) & 1} : work[blockIdx.x];
   const SolveFront f = fronts[w.front];
   SolveDevCtx cx;
   bwd_wide_G<NR>(cx, f, w.tile, step, x, pbuf + (size_t)(w.front - first) * SWB * NR, smem_dyn, part,
                  part == SW_NEAR ? (int)blockIdx.x % SW_BNSPLIT : 0, part == SW_NEAR ? SW_BNSPLIT : 1);
}

template <int NR, int NRT, bool POSDEF, int HB = SWB>
__global__ void __launch_bounds__(HB < SW_TT ? HB : SW_TT)
k_bwd_wide_T(const SolveFront* fronts, int first, int count, int step, double* __restrict__ x, double* __restrict__ pbuf) {
   extern __shared__ double smem_dyn[];
   const int fl = blockIdx.x % count;
   const SolveFront f = fronts[first + fl];
   const int k0 = (blockIdx.x / count) * NR;
   SolveDevCtx cx;
   bwd_wide_T_h<NR, NRT, POSDEF, HB>(cx, f, step, x + k0, pbuf + (size_t)fl * SWB * NRT + k0, smem_dyn);
}

/* ---- G kernels on the FP64 tensor cores (16 or 32 right-hand sides) ---------------------------------------- */
/* A 128-row tile times the block's <= 256 columns: L is streamed through shared memory in chunks of 32 columns
 * (two buffers; the next chunk travels global -> registers while the current one is multiplied), the right-hand
 * sides of the block (forward) or of the tile's rows (backward) sit in shared memory for the whole tile.  Strides
 * = 4 mod 16 doubles keep the 8 x 4 fragment loads bank-conflict free.  mma.sync.m8n8k4 with M = right-hand side. */
constexpr int SG_LLD = RT + 4;
template <int NR> constexpr int sg_xld() { return NR + 4; }
template <int NR> constexpr size_t sg_f_smem_bytes() { return ((size_t)2 * 32 * SG_LLD + (size_t)SWB * sg_xld<NR>()) * sizeof(double); }
template <int NR> constexpr size_t sg_b_smem_bytes() { return ((size_t)2 * 32 * SG_LLD + (size_t)RT * sg_xld<NR>()) * sizeof(double); }

struct SgChunk { double2 v[8]; };
/* chunk c of the tile: columns kb + 32 c .. + 31, rows r0 .. r0 + 127; rows outside [rlo, rhi) and columns >= kb + w read 0 */
__device__ __forceinline__ void sg_load(SgChunk& ck, const SolveFront& f, int kb, int w, int r0, int rlo, int rhi, int c, int tid) {
   const int rr = (tid & 63) * 2, cq = tid >> 6;
   const int r = r0 + rr;
   const bool a0 = r >= rlo && r < rhi, a1 = r + 1 >= rlo && r + 1 < rhi;
   const size_t ldl = (size_t)f.ldl;
   #pragma unroll
   for (int q = 0; q < 8; ++q) {
      const int col = 32 * c + cq + 4 * q;
      double2 v = make_double2(0.0, 0.0);
      if (col < w && (a0 || a1)) {
         const double* src = f.L + r + (size_t)(kb + col) * ldl;
         if (a0 && a1) v = *reinterpret_cast<const double2*>(src);
         else { if (a0) v.x = src[0]; if (a1) v.y = src[1]; }
      }
      ck.v[q] = v;
   }
}
__device__ __forceinline__ void sg_store(const SgChunk& ck, double* Ls, int tid) {
   const int rr = (tid & 63) * 2, cq = tid >> 6;
   #pragma unroll
   for (int q = 0; q < 8; ++q) *reinterpret_cast<double2*>(&Ls[(cq + 4 * q) * SG_LLD + rr]) = ck.v[q];
}

template <int NR>
__global__ void __launch_bounds__(SW_GT)
k_fwd_wide_G_mma(const SolveFront* fronts, const RowTile* work, int blk, double* __restrict__ x, const double* __restrict__ ywork,
      int part, int first) {
   extern __shared__ __align__(16) double smem_dyn[];
   /* near launches: two tiles per front of the level (no work list), SW_BNSPLIT CTAs share the 8 column chunks of a tile */
   const RowTile wk = part == SW_NEAR ? RowTile{first + (int)blockIdx.x / (2 * SW_BNSPLIT), ((int)blockIdx.x / SW_BNSPLIT) & 1} : work[blockIdx.x];
   const SolveFront f = fronts[wk.front];
   const int kb = blk * SWB;
   if (kb >= f.nelim) return;
   const int w = min(SWB, f.nelim - kb);
   int rlo, rhi;
   sw_part_rows(f, kb, w, part, rlo, rhi);
   const int r0 = (wk.tile + (part == SW_NEAR ? rlo / RT : 0)) * RT;
   if (r0 + RT <= rlo || r0 >= rhi) return;
   constexpr int YLD = sg_xld<NR>();
   double* Ls = smem_dyn;
   double* ys = smem_dyn + 2 * 32 * SG_LLD;
   const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
   const int nchunk = (w + 31) / 32;
   const int cper = part == SW_NEAR ? (SWB / 32) / SW_BNSPLIT : SWB / 32;
   const int c_begin = part == SW_NEAR ? ((int)blockIdx.x % SW_BNSPLIT) * cper : 0;
   const int c_end = min(nchunk, c_begin + cper);
   if (c_begin >= c_end) return;
   /* the right-hand sides of this CTA's columns: consecutive threads read consecutive values of a row (coalesced) */
   {
      constexpr int UN = 8;                                        // loads in flight per thread
      const int total = (c_end - c_begin) * 32 * NR;
      for (int base = tid; base < total; base += SW_GT * UN) {
         double v[UN];
         #pragma unroll
         for (int u = 0; u < UN; ++u) {
            const int idx = base + u * SW_GT;
            const int e = c_begin * 32 + idx / NR;
            v[u] = (idx < total && e < w) ? ywork[XI(f.perm[kb + e] - 1, idx % NR)] : 0.0;
         }
         #pragma unroll
         for (int u = 0; u < UN; ++u) {
            const int idx = base + u * SW_GT;
            if (idx < total) ys[(c_begin * 32 + idx / NR) * YLD + idx % NR] = v[u];
         }
      }
   }
   SgChunk ck;
   sg_load(ck, f, kb, w, r0, rlo, rhi, c_begin, tid);
   sg_store(ck, Ls + (c_begin & 1) * 32 * SG_LLD, tid);
   __syncthreads();
   const int rbase = warp * 16;
   double acc[NR / 8][2][2];
   #pragma unroll
   for (int j = 0; j < NR / 8; ++j)
      #pragma unroll
      for (int i = 0; i < 2; ++i) { acc[j][i][0] = 0.0; acc[j][i][1] = 0.0; }
   for (int c = c_begin; c < c_end; ++c) {
      if (c + 1 < c_end) sg_load(ck, f, kb, w, r0, rlo, rhi, c + 1, tid);
      const double* Lc = Ls + (c & 1) * 32 * SG_LLD;
      #pragma unroll
      for (int kk = 0; kk < 32; kk += 4) {
         double bfr[2], afr[NR / 8];
         #pragma unroll
         for (int i = 0; i < 2; ++i) bfr[i] = Lc[(kk + (lane & 3)) * SG_LLD + rbase + i * 8 + (lane >> 2)];
         #pragma unroll
         for (int j = 0; j < NR / 8; ++j) afr[j] = ys[(32 * c + kk + (lane & 3)) * YLD + j * 8 + (lane >> 2)];
         #pragma unroll
         for (int j = 0; j < NR / 8; ++j)
            #pragma unroll
            for (int i = 0; i < 2; ++i) pv_dmma(acc[j][i][0], acc[j][i][1], afr[j], bfr[i]);
      }
      if (c + 1 < c_end) sg_store(ck, Ls + ((c + 1) & 1) * 32 * SG_LLD, tid);
      __syncthreads();
   }
   /* acc[j][i][e] = sum for right-hand side 8 j + lane / 4 and row rbase + 8 i + 2 (lane % 4) + e */
   #pragma unroll
   for (int i = 0; i < 2; ++i)
      #pragma unroll
      for (int e = 0; e < 2; ++e) {
         const int r = r0 + rbase + i * 8 + 2 * (lane & 3) + e;
         if (r >= rlo && r < rhi) {
            const int g = row_index(f, r);
            #pragma unroll
            for (int j = 0; j < NR / 8; ++j) atomicAdd(&x[XI(g, j * 8 + (lane >> 2))], -acc[j][i][e]);
         }
      }
}

template <int NR>
__global__ void __launch_bounds__(SW_GT)
k_bwd_wide_G_mma(const SolveFront* fronts, const RowTile* work, int first, int step, const double* __restrict__ x,
      double* __restrict__ pbuf, int part) {
   extern __shared__ __align__(16) double smem_dyn[];
   const RowTile wk = part == SW_NEAR ? RowTile{first + (int)blockIdx.x / (2 * SW_BNSPLIT), ((int)blockIdx.x / SW_BNSPLIT) & 1} : work[blockIdx.x];
   const SolveFront f = fronts[wk.front];
   const int b = sw_bwd_block(f, step);
   if (b < 0) return;
   const int kb = b * SWB;
   const int w = min(SWB, f.nelim - kb);
   int rlo, rhi;
   sw_part_rows(f, kb, w, part, rlo, rhi);
   const int r0 = (wk.tile + (part == SW_NEAR ? rlo / RT : 0)) * RT;
   if (r0 + RT <= rlo || r0 >= rhi) return;
   constexpr int XLD = sg_xld<NR>();
   double* Ls = smem_dyn;
   double* xs = smem_dyn + 2 * 32 * SG_LLD;
   double* accb = pbuf + (size_t)(wk.front - first) * SWB * NR;
   const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
   const int nchunk = (w + 31) / 32;
   const int cper = part == SW_NEAR ? (SWB / 32) / SW_BNSPLIT : SWB / 32;
   const int c_begin = part == SW_NEAR ? ((int)blockIdx.x % SW_BNSPLIT) * cper : 0;
   const int c_end = min(nchunk, c_begin + cper);
   if (c_begin >= c_end) return;
   {                                                            // coalesced: consecutive threads, consecutive values of a row
      constexpr int UN = 8;
      for (int base = tid; base < RT * NR; base += SW_GT * UN) {
         double v[UN];
         #pragma unroll
         for (int u = 0; u < UN; ++u) {
            const int idx = base + u * SW_GT;
            const int r = r0 + idx / NR;
            v[u] = (idx < RT * NR && r >= rlo && r < rhi) ? x[XI(row_index(f, r), idx % NR)] : 0.0;
         }
         #pragma unroll
         for (int u = 0; u < UN; ++u) {
            const int idx = base + u * SW_GT;
            if (idx < RT * NR) xs[(idx / NR) * XLD + idx % NR] = v[u];
         }
      }
   }
   SgChunk ck;
   sg_load(ck, f, kb, w, r0, rlo, rhi, c_begin, tid);
   sg_store(ck, Ls + (c_begin & 1) * 32 * SG_LLD, tid);
   __syncthreads();
   constexpr int NB = NR / 16;               // 8 x 8 output blocks per warp and chunk: (NR / 8 right-hand-side groups) x 4 column groups / 8 warps
   for (int c = c_begin; c < c_end; ++c) {
      if (c + 1 < c_end) sg_load(ck, f, kb, w, r0, rlo, rhi, c + 1, tid);
      const double* Lc = Ls + (c & 1) * 32 * SG_LLD;
      double acc[NB][2];
      int jg[NB], cg[NB];
      #pragma unroll
      for (int u = 0; u < NB; ++u) { acc[u][0] = 0.0; acc[u][1] = 0.0; const int bid = warp + 8 * u; jg[u] = bid % (NR / 8); cg[u] = bid / (NR / 8); }
      #pragma unroll 8
      for (int kk = 0; kk < RT; kk += 4) {
         #pragma unroll
         for (int u = 0; u < NB; ++u) {
            const double afr = xs[(kk + (lane & 3)) * XLD + 8 * jg[u] + (lane >> 2)];
            const double bfr = Lc[(8 * cg[u] + (lane >> 2)) * SG_LLD + kk + (lane & 3)];
            pv_dmma(acc[u][0], acc[u][1], afr, bfr);
         }
      }
      /* acc[u][e]: right-hand side 8 jg + lane / 4, column 32 c + 8 cg + 2 (lane % 4) + e of the block */
      #pragma unroll
      for (int u = 0; u < NB; ++u)
         #pragma unroll
         for (int e = 0; e < 2; ++e) {
            const int col = 32 * c + 8 * cg[u] + 2 * (lane & 3) + e;
            if (col < w) atomicAdd(&accb[(size_t)col * NR + 8 * jg[u] + (lane >> 2)], acc[u][e]);
         }
      if (c + 1 < c_end) sg_store(ck, Ls + ((c + 1) & 1) * 32 * SG_LLD, tid);
      __syncthreads();
   }
}

template <int NR> struct UseMma { static constexpr bool value = (NR >= 16); };
/* right-hand sides per T CTA: a block of 32 or 64 is solved by 2 or 4 CTAs side by side (the T kernels are one CTA
 * per front and latency-bound; the SMs are idle while they run) */
template <int NR> constexpr int t_slice() { return NR > 16 ? 16 : NR; }

template <int NR, bool POSDEF>
void fwd_level_wide_t(const SolveFront* fronts, int first, int count, const RowTile* work, int nwork, int nblk,
      double* x, double* ywork, cudaStream_t s, SolveAux* aux, bool half) {
   static bool configured_dev[64] = {false};       // function attributes are per device
   int dev_ = 0;
   cudaGetDevice(&dev_);
   bool& configured = configured_dev[dev_ & 63];
   constexpr int TS = t_slice<NR>();
   constexpr bool MMA = UseMma<NR>::value;
   const size_t smT = sw_T_smem_doubles<TS>() * sizeof(double);
   size_t smG;
   if constexpr (MMA) smG = sg_f_smem_bytes<NR>(); else smG = sw_fG_smem_doubles<NR>() * sizeof(double);
   if (!configured) {
      cudaFuncSetAttribute(k_fwd_wide_T<TS, NR, POSDEF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smT);
      if constexpr (MMA) cudaFuncSetAttribute(k_fwd_wide_T<TS, NR, POSDEF, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)(sw_T_smem_doubles<TS, 128>() * sizeof(double)));
      if constexpr (MMA) cudaFuncSetAttribute(k_fwd_wide_G_mma<NR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smG);
      else {
         cudaFuncSetAttribute(k_fwd_wide_G<NR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smG);
         cudaFuncSetAttribute(k_fwd_wide_G_near<NR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smG);
      }
      configured = true;
   }
   auto G = [&](int b, int part, cudaStream_t st) {
      if constexpr (MMA) {
         const int grid = part == SW_NEAR ? 2 * count * SW_BNSPLIT : nwork;
         k_fwd_wide_G_mma<NR><<<grid, SW_GT, smG, st>>>(fronts, work, b, x, ywork, part, first);
      } else if (part == SW_NEAR) {
         k_fwd_wide_G_near<NR><<<2 * count * SW_NSPLIT, SW_GT, smG, st>>>(fronts, first, b, x, ywork);
      } else {
         k_fwd_wide_G<NR><<<nwork * SW_FSPLIT, SW_GT, smG, st>>>(fronts, work, b, x, ywork, part, first);
      }
      COUNT_LAUNCH();
   };
   auto T = [&](int b) {
      if constexpr (MMA) {
         if (half) {          // every front of the level eliminates at most 128 columns: three T CTAs per SM
            const size_t smTh = sw_T_smem_doubles<TS, 128>() * sizeof(double);
            k_fwd_wide_T<TS, NR, POSDEF, 128><<<count * (NR / TS), 128, smTh, s>>>(fronts, first, count, b, x, ywork);
            COUNT_LAUNCH();
            return;
         }
      }
      k_fwd_wide_T<TS, NR, POSDEF><<<count * (NR / TS), SW_TT, smT, s>>>(fronts, first, count, b, x, ywork); COUNT_LAUNCH();
   };
   if (!aux || nblk < 2) {
      for (int b = 0; b < nblk; ++b) { T(b); G(b, SW_ALL, s); }
      return;
   }
   /* T(b) needs the near part of block b - 1 (same stream) and the far parts of the blocks up to b - 2 */
   static const bool tl = getenv("SPRAL_B200_TRACE_SOLVE") != nullptr;         // timeline of the sweep (events per launch)
   std::vector<cudaEvent_t> tev;
   auto mark = [&](cudaStream_t st) { if (tl && nblk >= 16) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); tev.push_back(e); } };
   mark(s);
   for (int b = 0; b < nblk; ++b) {
      if (b >= 2) cudaStreamWaitEvent(s, aux->evF[(b - 2) & 3], 0);
      mark(s);
      T(b);
      mark(s);
      cudaEventRecord(aux->evT[b & 3], s);
      cudaStream_t fs = aux->far[b & 1];
      cudaStreamWaitEvent(fs, aux->evT[b & 3], 0);
      mark(fs);
      G(b, SW_FAR, fs);
      mark(fs);
      cudaEventRecord(aux->evF[b & 3], fs);
      G(b, SW_NEAR, s);
      mark(s);
   }
   cudaStreamWaitEvent(s, aux->evF[(nblk - 2) & 3], 0);
   cudaStreamWaitEvent(s, aux->evF[(nblk - 1) & 3], 0);
   if (!tev.empty()) {
      cudaStreamSynchronize(s);
      for (int b = 0; b < nblk; ++b) {
         float t[5];
         for (int i = 0; i < 5; ++i) cudaEventElapsedTime(&t[i], tev[0], tev[1 + 5 * b + i]);
         fprintf(stderr, "[fwd nr %d] block %3d: T %8.1f -> %8.1f us, near -> %8.1f, far %8.1f -> %8.1f\n", NR, b,
                 1e3 * t[0], 1e3 * t[1], 1e3 * t[4], 1e3 * t[2], 1e3 * t[3]);
      }
      for (auto e : tev) cudaEventDestroy(e);
   }
}

/* pbuf: one SWB x NR accumulator per front of the level (two with look-ahead: blocks alternate between them), all
 * zero when the sweep of the level starts (the T kernel clears what it consumed) */
template <int NR, bool POSDEF>
void bwd_level_wide_t(const SolveFront* fronts, int first, int count, const RowTile* work, int nwork,
      const int* wbeg, int nblk, double* x, double* pbuf, cudaStream_t s, SolveAux* aux, bool half) {
   static bool configured_dev[64] = {false};       // function attributes are per device
   int dev_ = 0;
   cudaGetDevice(&dev_);
   bool& configured = configured_dev[dev_ & 63];
   constexpr int TS = t_slice<NR>();
   constexpr bool MMA = UseMma<NR>::value;
   const size_t smT = sw_T_smem_doubles<TS>() * sizeof(double);
   if (!configured) {
      cudaFuncSetAttribute(k_bwd_wide_T<TS, NR, POSDEF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smT);
      if constexpr (MMA) cudaFuncSetAttribute(k_bwd_wide_T<TS, NR, POSDEF, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)(sw_T_smem_doubles<TS, 128>() * sizeof(double)));
      if constexpr (MMA)
         cudaFuncSetAttribute(k_bwd_wide_G_mma<NR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sg_b_smem_bytes<NR>());
      configured = true;
   }
   const size_t accsz = (size_t)count * SWB * NR;
   auto G = [&](int st, int part, double* acc, cudaStream_t strm) {
      const int grid = part == SW_NEAR ? 2 * count * SW_BNSPLIT : nwork;
      if constexpr (MMA) k_bwd_wide_G_mma<NR><<<grid, SW_GT, sg_b_smem_bytes<NR>(), strm>>>(fronts, work, first, st, x, acc, part);
      else k_bwd_wide_G<NR><<<grid, SW_GT, 0, strm>>>(fronts, work, first, st, x, acc, part);
      COUNT_LAUNCH();
   };
   auto T = [&](int st, double* acc) {
      if constexpr (MMA) {
         if (half) {
            const size_t smTh = sw_T_smem_doubles<TS, 128>() * sizeof(double);
            k_bwd_wide_T<TS, NR, POSDEF, 128><<<count * (NR / TS), 128, smTh, s>>>(fronts, first, count, st, x, acc);
            COUNT_LAUNCH();
            return;
         }
      }
      k_bwd_wide_T<TS, NR, POSDEF><<<count * (NR / TS), SW_TT, smT, s>>>(fronts, first, count, st, x, acc); COUNT_LAUNCH();
   };
   if (!aux || nblk < 2) {
      cudaMemsetAsync(pbuf, 0, accsz * sizeof(double), s);
      for (int st = 0; st < nblk; ++st) { G(st, SW_ALL, pbuf, s); T(st, pbuf); }
      return;
   }
   /* T(st) needs the near part of its block (the rows T(st - 1) has just solved: same stream) and the far part, which
    * only reads what T(st - 2) and earlier left and therefore runs two steps ahead on the far streams */
   cudaMemsetAsync(pbuf, 0, 2 * accsz * sizeof(double), s);
   cudaEventRecord(aux->evT[3], s);
   for (int st = 0; st < 2; ++st) {
      cudaStreamWaitEvent(aux->far[st], aux->evT[3], 0);
      G(st, SW_FAR, pbuf + st * accsz, aux->far[st]);
      cudaEventRecord(aux->evF[st], aux->far[st]);
   }
   static const bool tl = getenv("SPRAL_B200_TRACE_SOLVE") != nullptr;
   std::vector<cudaEvent_t> tev;
   auto mark = [&](cudaStream_t st) { if (tl && nblk >= 16) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); tev.push_back(e); } };
   mark(s);
   for (int st = 0; st < nblk; ++st) {
      double* acc = pbuf + (st & 1) * accsz;
      mark(s);
      if (st >= 1) G(st, SW_NEAR, acc, s);
      mark(s);
      cudaStreamWaitEvent(s, aux->evF[st & 3], 0);
      mark(s);
      T(st, acc);
      mark(s);
      if (st + 2 < nblk) {
         cudaEventRecord(aux->evT[st & 3], s);
         cudaStream_t fs = aux->far[st & 1];
         cudaStreamWaitEvent(fs, aux->evT[st & 3], 0);
         G(st + 2, SW_FAR, acc, fs);
         cudaEventRecord(aux->evF[(st + 2) & 3], fs);
      }
   }
   if (!tev.empty()) {
      cudaStreamSynchronize(s);
      for (int st = 0; st < nblk; ++st) {
         float t[4];
         for (int i = 0; i < 4; ++i) cudaEventElapsedTime(&t[i], tev[0], tev[1 + 4 * st + i]);
         fprintf(stderr, "[bwd nr %d] step %3d: near %8.1f -> %8.1f us, T %8.1f -> %8.1f\n", NR, st, 1e3 * t[0], 1e3 * t[1], 1e3 * t[2], 1e3 * t[3]);
      }
      for (auto e : tev) cudaEventDestroy(e);
   }
}

} // namespace

/* x (column-major, ld = ldx, nr columns) <-> xt (n rows of nr contiguous values): 32 x 32 tiles through shared memory,
 * coalesced on both sides.  Grid = ceil(n / 32) x ceil(nr / 32) tiles, flattened. */
__global__ void __launch_bounds__(256)
k_transpose_rhs(double* __restrict__ x, int ldx, double* __restrict__ xt, int n, int nr, int to_xt) {
   __shared__ double tile[32][33];
   const int nkt = (nr + 31) / 32;
   const int g0 = (blockIdx.x / nkt) * 32, k0 = (blockIdx.x % nkt) * 32;
   const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 8 warps
   if (to_xt) {
      #pragma unroll
      for (int i = 0; i < 4; ++i) {                                  // read x: lanes along g (contiguous)
         const int k = k0 + ty + 8 * i, g = g0 + tx;
         tile[ty + 8 * i][tx] = (k < nr && g < n) ? x[g + (size_t)k * ldx] : 0.0;
      }
      __syncthreads();
      #pragma unroll
      for (int i = 0; i < 4; ++i) {                                  // write xt: lanes along k (contiguous)
         const int g = g0 + ty + 8 * i, k = k0 + tx;
         if (g < n && k < nr) xt[(size_t)g * nr + k] = tile[tx][ty + 8 * i];
      }
   } else {
      #pragma unroll
      for (int i = 0; i < 4; ++i) {
         const int g = g0 + ty + 8 * i, k = k0 + tx;
         tile[tx][ty + 8 * i] = (g < n && k < nr) ? xt[(size_t)g * nr + k] : 0.0;
      }
      __syncthreads();
      #pragma unroll
      for (int i = 0; i < 4; ++i) {
         const int k = k0 + ty + 8 * i, g = g0 + tx;
         if (k < nr && g < n) x[g + (size_t)k * ldx] = tile[ty + 8 * i][tx];
      }
   }
}

void launch_transpose_rhs(double* x, int ldx, double* xt, int n, int nr, bool to_xt, cudaStream_t s) {
   if (n == 0) return;
   k_transpose_rhs<<<((n + 31) / 32) * ((nr + 31) / 32), 256, 0, s>>>(x, ldx, xt, n, nr, to_xt ? 1 : 0); COUNT_LAUNCH();
}

int solve_block() { return SB; }

/* Wide sweeps of one level (solve_wide.h): nblk = number of 256-column blocks of its largest front. */
int solve_wide_block() { return SWB; }

#define SW_DISPATCH(NRV, CALL_T, CALL_F) case NRV: if (posdef) { CALL_T; } else { CALL_F; } break
void launch_fwd_level_wide(const SolveFront* fronts, int first, int count, const RowTile* work, int nwork,
      int nblk, bool posdef, int nr, double* x, double* ywork, cudaStream_t s, SolveAux* aux, bool half) {
   if (nwork == 0 || nblk == 0 || count == 0) return;
#define SW_F(NRV) SW_DISPATCH(NRV, (fwd_level_wide_t<NRV, true>(fronts, first, count, work, nwork, nblk, x, ywork, s, aux, half)), \
                                   (fwd_level_wide_t<NRV, false>(fronts, first, count, work, nwork, nblk, x, ywork, s, aux, half)))
   switch (nr) { SW_F(64); SW_F(32); SW_F(16); SW_F(8); SW_F(4); SW_F(2); default: SW_F(1); }
#undef SW_F
}

void launch_bwd_level_wide(const SolveFront* fronts, int first, int count, const RowTile* work, int nwork,
      const int* wbeg, int nblk, bool posdef, int nr, double* x, double* pbuf, cudaStream_t s, SolveAux* aux, bool half) {
   if (nwork == 0 || nblk == 0 || count == 0) return;
#define SW_B(NRV) SW_DISPATCH(NRV, (bwd_level_wide_t<NRV, true>(fronts, first, count, work, nwork, wbeg, nblk, x, pbuf, s, aux, half)), \
                                   (bwd_level_wide_t<NRV, false>(fronts, first, count, work, nwork, wbeg, nblk, x, pbuf, s, aux, half)))
   switch (nr) { SW_B(64); SW_B(32); SW_B(16); SW_B(8); SW_B(4); SW_B(2); default: SW_B(1); }
#undef SW_B
}
#undef SW_DISPATCH

/* Inverses of the 32 x 32 diagonal blocks of L for the fronts that carry SolveFront::Linv: one warp per block, lane j
 * solves L x = e_j by substitution (column j of the inverse), the block is written row-major.  Unit diagonal for
 * L D L^T, the stored diagonal for Cholesky; rows and columns past nelim are identity.  An entry beyond LINV_LIMIT (or
 * not finite) marks the whole front: the explicit inverse is only as accurate as the block is well conditioned, and
 * the threshold test bounds the entries of L (by 1 / u), not of its inverse. */
constexpr double LINV_LIMIT = 1e4;
template <bool POSDEF>
__global__ void __launch_bounds__(32)
k_build_linv(const SolveFront* fronts, const int2* work) {
   __shared__ double Lb[32][33];
   __shared__ double X[32][33];
   const int2 w = work[blockIdx.x];
   const SolveFront f = fronts[w.x];
   const int lane = threadIdx.x;
   const int c0 = w.y * 32;
   const int wd = min(32, f.nelim - c0);
   for (int j = 0; j < 32; ++j) {                        // lane = row
      double v = 0.0;
      if (lane < wd && j < wd && lane >= j) v = f.L[(size_t)(c0 + lane) + (size_t)(c0 + j) * (size_t)f.ldl];
      if (lane == j && (!POSDEF || lane >= wd)) v = 1.0;
      Lb[lane][j] = v;
   }
   __syncwarp();
   const int j = lane;                                   // lane = column of the inverse
   bool bad = false;
   for (int i = 0; i < 32; ++i) {
      double x = 0.0;
      if (i >= j) {
         double sum = (i == j) ? 1.0 : 0.0;
         for (int k = j; k < i; ++k) sum -= Lb[i][k] * X[k][j];
         x = sum / Lb[i][i];
      }
      X[i][j] = x;
      if (!(fabs(x) <= LINV_LIMIT)) bad = true;
   }
   __syncwarp();
   double* out = const_cast<double*>(f.Linv) + (size_t)w.y * 1024;
   for (int i = 0; i < 32; ++i) out[i * 32 + lane] = X[i][lane];
   if (__ballot_sync(0xffffffffu, bad) != 0u && lane == 0) *const_cast<int*>(f.linv_bad) = 1;
}

void launch_build_linv(const SolveFront* fronts, const int2* work, int nwork, bool posdef, cudaStream_t s) {
   if (nwork == 0) return;
   if (posdef) k_build_linv<true><<<nwork, 32, 0, s>>>(fronts, work);
   else k_build_linv<false><<<nwork, 32, 0, s>>>(fronts, work);
   COUNT_LAUNCH();
}

void SolveAux::create() {
   if (far[0]) return;
   int least = 0, greatest = 0;
   cudaDeviceGetStreamPriorityRange(&least, &greatest);
   for (int i = 0; i < 2; ++i) cudaStreamCreateWithPriority(&far[i], cudaStreamNonBlocking, least);
   for (int i = 0; i < 4; ++i) {
      cudaEventCreateWithFlags(&evT[i], cudaEventDisableTiming);
      cudaEventCreateWithFlags(&evF[i], cudaEventDisableTiming);
   }
}
void SolveAux::destroy() {
   if (!far[0]) return;
   for (int i = 0; i < 2; ++i) { cudaStreamSynchronize(far[i]); cudaStreamDestroy(far[i]); far[i] = nullptr; }
   for (int i = 0; i < 4; ++i) { cudaEventDestroy(evT[i]); cudaEventDestroy(evF[i]); }
}

void configure_solve_kernels() {
   cudaFuncSetAttribute(k_bwd_reduce<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)reduce_smem<32>());
   cudaFuncSetAttribute(k_bwd_reduce<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)reduce_smem<16>());
}

/* Largest number of right-hand sides one kernel pass handles. */
int solve_rhs_chunk(int nrhs) { return nrhs >= 32 ? 32 : nrhs >= 16 ? 16 : nrhs >= 8 ? 8 : nrhs >= 4 ? 4 : nrhs >= 2 ? 2 : 1; }
/* 64 right-hand sides in one pass exist in the wide kernels only (tensor cores; L is read once instead of twice) */
int solve_max_chunk() { return 64; }

void launch_fwd_level(const SolveFront* fronts, const RowTile* work, int nwork, int nsteps,
      bool posdef, int nr, double* x, int ldx, double* ywork, cudaStream_t s, unsigned int* bar) {
   if (nwork == 0 || nsteps == 0) return;
   switch (nr) {
   case 32: fwd_level<32>(fronts, work, nwork, nsteps, posdef, x, ldx, ywork, bar, s); break;
   case 16: fwd_level<16>(fronts, work, nwork, nsteps, posdef, x, ldx, ywork, bar, s); break;
   case 8: fwd_level<8>(fronts, work, nwork, nsteps, posdef, x, ldx, ywork, bar, s); break;
   case 4: fwd_level<4>(fronts, work, nwork, nsteps, posdef, x, ldx, ywork, bar, s); break;
   case 2: fwd_level<2>(fronts, work, nwork, nsteps, posdef, x, ldx, ywork, bar, s); break;
   default: fwd_level<1>(fronts, work, nwork, nsteps, posdef, x, ldx, ywork, bar, s); break;
   }
}

void launch_fwd_flush(const SolveFront* fronts, int first, int count, int nrhs, double* x, int ldx,
      const double* ywork, cudaStream_t s) {
   if (count == 0) return;
   k_fwd_flush<<<count * 8, 256, 0, s>>>(fronts, first, count, nrhs, x, ldx, ywork); COUNT_LAUNCH();
}

void launch_diag_solve(const SolveFront* fronts, int first, int count, int nrhs, double* x, int ldx,
      cudaStream_t s) {
   if (count == 0) return;
   const int nchunk = count <= 64 ? 16 : 2;
   k_diag_solve<<<count * nchunk, 256, 0, s>>>(fronts, first, count, nrhs, x, ldx); COUNT_LAUNCH();
}

void launch_bwd_level(const SolveFront* fronts, int first, int count, const RowTile* work, int nwork,
      const int* wbeg, int nsteps, bool posdef, int nr, double* x, int ldx, double* pbuf,
      cudaStream_t s, unsigned int* bar) {
   if (nwork == 0 || nsteps == 0) return;
   switch (nr) {
   case 32: bwd_level<32>(fronts, first, count, work, nwork, wbeg, nsteps, posdef, x, ldx, pbuf, bar, s); break;
   case 16: bwd_level<16>(fronts, first, count, work, nwork, wbeg, nsteps, posdef, x, ldx, pbuf, bar, s); break;
   case 8: bwd_level<8>(fronts, first, count, work, nwork, wbeg, nsteps, posdef, x, ldx, pbuf, bar, s); break;
   case 4: bwd_level<4>(fronts, first, count, work, nwork, wbeg, nsteps, posdef, x, ldx, pbuf, bar, s); break;
   case 2: bwd_level<2>(fronts, first, count, work, nwork, wbeg, nsteps, posdef, x, ldx, pbuf, bar, s); break;
   default: bwd_level<1>(fronts, first, count, work, nwork, wbeg, nsteps, posdef, x, ldx, pbuf, bar, s); break;
   }
}

} // namespace b200
