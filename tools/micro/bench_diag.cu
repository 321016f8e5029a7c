/* Micro-benchmark (developer tool): cycles of the one-warp 32 x 32 LDL^T (diag_warp.cuh) by phase.
 *   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I spral_b200/csrc -I include -o build/bench_diag tools/micro/bench_diag.cu */
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <math_constants.h>
#include "diag_warp.cuh"
using namespace b200;
namespace b200 {
__device__ __forceinline__ int ldlt_instr(double* S, double* dinv, int* lperm, int bs, double small, int action,
      double inf, int& zfrom_out, long long* ph) {
   const int lane = threadIdx.x & 31;
   zfrom_out = 32;
   int lp = lane;                       // lperm[lane], kept in a register: a swap is two shuffles off the critical path
   dinv[2 * lane] = 0.0; dinv[2 * lane + 1] = 0.0;
   double rdiag = 0.0;                  // 1 / S(lane, lane), computed while the trailing update runs
   if (lane < bs) rdiag = 1.0 / S[lane * DW_LD + lane];
   /* largest entry of column `lane` of the remaining lower triangle (ties: smallest row) */
   double cmaxv = -1.0;
   int cmaxr = lane;
   {
      DwMax mx; mx.reset(lane);
      #pragma unroll
      for (int r = 0; r < 32; ++r) {
         const double av = fabs(S[r * DW_LD + lane]);
         mx.add(r & 3, r >= lane && r < bs && lane < bs, av, r);
      }
      mx.result(cmaxv, cmaxr);
   }
   int p = 0;
   while (p < bs) {
      long long c0 = clock64();
      double best;
      const int m = dw_argmax(cmaxv, lane >= p && lane < bs, best);     // column m <= row t
      int t = __shfl_sync(DW_FULL, cmaxr, m);
      __syncwarp();
      long long c1 = clock64(); ph[0] += c1 - c0;
      int ps = 1;
      double e11 = 0.0, e21 = 0.0, e22 = 0.0;
      if (!(best >= small)) ps = 0;
      else if (t == m) e11 = __shfl_sync(DW_FULL, rdiag, t);      // 1 / S(t, t)
      else {
         const double a11 = S[m * DW_LD + m], a22 = S[t * DW_LD + t], a21 = S[t * DW_LD + m];
         const double detscale = 1.0 / fabs(a21);
         const double detpiv = (a11 * detscale) * a22 - fabs(a21);
         if (fabs(detpiv) >= fabs(a21) / 2) {
            ps = 2;
            e11 = (a22 * detscale) / detpiv;
            e22 = (a11 * detscale) / detpiv;
            e21 = (-a21 * detscale) / detpiv;
         } else {
            if (fabs(a11) > fabs(a22)) t = m;      // a11 as 1x1, else a22 (row / column t)
            e11 = __shfl_sync(DW_FULL, rdiag, t);
         }
      }
      __syncwarp();                       // every lane has read the pivot entries before anybody swaps them
      long long c2 = clock64(); ph[1] += c2 - c1; if (ps == 2) ph[5] += c2 - c1;
      if (ps == 0) {
         /* everything left is (numerically) zero: block_ldlt.hxx:303-317 */
         if (!action) return DW_SINGULAR;
         zfrom_out = p;
         if (lane >= p && lane < bs)
            for (int r = lane + 1; r < bs; ++r) { S[r * DW_LD + lane] = 0.0; S[lane * DW_LD + r] = 0.0; }
         break;
      }
      if (ps == 1) {
         if (t != p) {
            dw_swap(S, p, t, p, lane);
            { const int la = __shfl_sync(DW_FULL, lp, p), lb = __shfl_sync(DW_FULL, lp, t); lp = (lane == p) ? lb : (lane == t ? la : lp);
              const double ra = __shfl_sync(DW_FULL, rdiag, p), rb = __shfl_sync(DW_FULL, rdiag, t); rdiag = (lane == p) ? rb : (lane == t ? ra : rdiag); }
            __syncwarp();
         }
         long long c3 = clock64(); ph[2] += c3 - c2;
         double w = 0.0;
         if (lane > p && lane < bs) {
            w = S[lane * DW_LD + p];
            S[lane * DW_LD + p] = w * e11;          // L
            S[p * DW_LD + lane] = w;                // L*D, mirrored
         }
         if (lane == 0) { dinv[2 * p] = e11; dinv[2 * p + 1] = 0.0; }
         __syncwarp();
         long long c4 = clock64(); ph[3] += c4 - c3;
         {
            /* rank-1 update of column `lane`.  The reciprocal of the lane's new diagonal entry -- the next 1x1
             * pivot's D^-1 -- is computed in the same straight-line code, so its latency hides behind the update
             * instead of sitting between the search and the swap of the next pivot. */
            const bool on = lane > p && lane < bs;
            const double dnew = S[lane * DW_LD + lane] - (w * e11) * w;       // == the value the update stores at (lane, lane)
            if (on) rdiag = 1.0 / dnew;
            DwMax mx; mx.reset(lane);
            dw_update_from<1>(S, p, lane, bs, on, w, 0.0, mx);
            mx.result(cmaxv, cmaxr);
         }
         ph[4] += clock64() - c4; ph[6] += 1;
      } else {
         /* swap p <-> m, then p+1 <-> t */
         if (m != p) {
            dw_swap(S, p, m, p, lane);
            { const int la = __shfl_sync(DW_FULL, lp, p), lb = __shfl_sync(DW_FULL, lp, m); lp = (lane == p) ? lb : (lane == m ? la : lp);
              const double ra = __shfl_sync(DW_FULL, rdiag, p), rb = __shfl_sync(DW_FULL, rdiag, m); rdiag = (lane == p) ? rb : (lane == m ? ra : rdiag); }
            __syncwarp();
         }
         if (t != p + 1) {
            dw_swap(S, p + 1, t, p, lane);
            { const int la = __shfl_sync(DW_FULL, lp, p + 1), lb = __shfl_sync(DW_FULL, lp, t); lp = (lane == p + 1) ? lb : (lane == t ? la : lp);
              const double ra = __shfl_sync(DW_FULL, rdiag, p + 1), rb = __shfl_sync(DW_FULL, rdiag, t); rdiag = (lane == p + 1) ? rb : (lane == t ? ra : rdiag); }
            __syncwarp();
         }
         double w1 = 0.0, w2 = 0.0;
         if (lane > p + 1 && lane < bs) {
            w1 = S[lane * DW_LD + p]; w2 = S[lane * DW_LD + p + 1];
            S[lane * DW_LD + p] = e11 * w1 + e21 * w2;
            S[lane * DW_LD + p + 1] = e21 * w1 + e22 * w2;
            S[p * DW_LD + lane] = w1;
            S[(p + 1) * DW_LD + lane] = w2;
         }
         if (lane == 0) {
            S[(p + 1) * DW_LD + p] = 0.0;           // the 2x2 diagonal block of L is the identity
            S[p * DW_LD + p + 1] = 0.0;
            dinv[2 * p] = e11; dinv[2 * p + 1] = e21;
            dinv[2 * p + 2] = inf; dinv[2 * p + 3] = e22;
         }
         __syncwarp();
         {
            const bool on = lane > p + 1 && lane < bs;
            const double dnew = S[lane * DW_LD + lane] - (w1 * (e11 * w1 + e21 * w2) + w2 * (e21 * w1 + e22 * w2));
            if (on) rdiag = 1.0 / dnew;
            DwMax mx; mx.reset(lane);
            dw_update_from<2>(S, p, lane, bs, on, w1, w2, mx);
            mx.result(cmaxv, cmaxr);
         }
      }
      if (ps == 2) { ph[7] += clock64() - c2; ph[8] += 1; }
      p += ps;
   }
   lperm[lane] = lp;
   __syncwarp();
   return DW_OK;
}

}

__global__ void k_phases(const double* A, int nblk, long long* ph_out) {
   __shared__ double S[32 * DW_LD];
   __shared__ double dinv[64];
   __shared__ int lperm[32];
   const int lane = threadIdx.x;
   long long ph[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
   for (int b = 0; b < nblk; ++b) {
      for (int c = 0; c < 32; ++c) S[lane * DW_LD + c] = (lane >= c) ? A[(size_t)b * 1024 + lane + c * 32] : 0.0;
      __syncwarp();
      int z;
      ldlt_instr(S, dinv, lperm, 32, 1e-20, 1, CUDART_INF, z, ph);
      __syncwarp();
   }
   if (lane == 0) for (int i = 0; i < 9; ++i) ph_out[i] = ph[i];
}

__global__ void k_time(const double* A, int nblk, long long* cyc, double* sink, int variant) {
   __shared__ double S[32 * DW_LD];
   __shared__ double dinv[64];
   __shared__ int lperm[32];
   const int lane = threadIdx.x;
   long long tot = 0;
   for (int b = 0; b < nblk; ++b) {
      for (int c = 0; c < 32; ++c) S[lane * DW_LD + c] = (lane >= c) ? A[(size_t)b * 1024 + lane + c * 32] : 0.0;
      __syncwarp();
      long long t0 = clock64();
      int z;
      int rc = (variant == 0) ? diag_warp_ldlt(S, dinv, lperm, 32, 1e-20, 1, CUDART_INF, z)
                              : diag_warp_chol(S, dinv, 32);
      long long t1 = clock64();
      tot += t1 - t0;
      if (lane == 0) sink[b] = S[5 * DW_LD + 3] + rc + dinv[7];
      __syncwarp();
   }
   if (lane == 0) cyc[blockIdx.x] = tot;
}

int main() {
   const int nblk = 64;
   std::vector<double> h((size_t)nblk * 1024);
   srand(1);
   for (int b = 0; b < nblk; ++b)
      for (int c = 0; c < 32; ++c)
         for (int r = c; r < 32; ++r) h[(size_t)b * 1024 + r + c * 32] = (double)rand() / RAND_MAX * 2 - 1;
   std::vector<double> hp = h;                       // diagonally dominant copy for Cholesky / 1x1-only pivots
   for (int b = 0; b < nblk; ++b) for (int c = 0; c < 32; ++c) hp[(size_t)b * 1024 + c + c * 32] += 40.0;
   double *dA, *dP, *sink; long long* cyc;
   cudaMalloc(&dA, h.size() * 8); cudaMalloc(&dP, h.size() * 8); cudaMalloc(&sink, nblk * 8); cudaMalloc(&cyc, 8 * 8);
   cudaMemcpy(dA, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
   cudaMemcpy(dP, hp.data(), hp.size() * 8, cudaMemcpyHostToDevice);
   const char* names[3] = {"ldlt random indefinite (1x1 + 2x2)", "ldlt diagonally dominant (1x1 on the diagonal)", "cholesky"};
   for (int cfg = 0; cfg < 3; ++cfg) {
      for (int rep = 0; rep < 2; ++rep) k_time<<<1, 32>>>(cfg == 0 ? dA : dP, nblk, cyc, sink, cfg == 2);
      cudaDeviceSynchronize();
      long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0); k_time<<<1, 32>>>(cfg == 0 ? dA : dP, nblk, cyc, sink, cfg == 2); cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      printf("%-48s %8.0f cycles / block = %6.1f cycles / column; kernel %.1f us / block\n", names[cfg], (double)c / nblk,
             (double)c / nblk / 32, 1e3 * ms / nblk);
   }
   {
      long long* ph; cudaMalloc(&ph, 9 * 8);
      for (int cfg = 0; cfg < 2; ++cfg) {
         k_phases<<<1, 32>>>(cfg == 0 ? dA : dP, nblk, ph); cudaDeviceSynchronize();
         long long h9[9]; cudaMemcpy(h9, ph, 72, cudaMemcpyDeviceToHost);
         double npiv = (double)(h9[6] + h9[8]);
         printf("%s: pivots/block %.1f (1x1 %.1f, 2x2 %.1f); cycles per pivot: argmax %.0f decision %.0f (2x2 decisions: %.0f each) "
                "| 1x1: swap %.0f scale %.0f update %.0f | 2x2 swap+scale+update %.0f\n", cfg == 0 ? "indefinite" : "dominant",
                npiv / nblk, (double)h9[6] / nblk, (double)h9[8] / nblk, h9[0] / npiv, h9[1] / npiv, h9[8] ? (double)h9[5] / h9[8] : 0.0,
                h9[6] ? (double)h9[2] / h9[6] : 0.0, h9[6] ? (double)h9[3] / h9[6] : 0.0, h9[6] ? (double)h9[4] / h9[6] : 0.0,
                h9[8] ? (double)h9[7] / h9[8] : 0.0);
      }
   }
   printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
   return 0;
}
