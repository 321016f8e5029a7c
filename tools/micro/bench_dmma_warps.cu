// Micro-benchmark: FP64 DMMA rate of a realistic warp-tile main loop (fragments from shared memory, as in
// gemm_dmma.cu) against the number of consumer warps per SM and the warp tile shape.  One CTA per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bench_dmma_warps bench_dmma_warps.cu && ./bench_dmma_warps
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
   asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int NC, int NR, int BKT, bool PIPE, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) k(double* out, int iters) {
   extern __shared__ double sm[];
   constexpr int LDS_ = 132;
   for (int i = threadIdx.x; i < 2 * BKT * LDS_; i += blockDim.x) sm[i] = 1e-3 * (i % 7);
   __syncthreads();
   const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
   const double* As = sm + (lane & 3) * LDS_ + (warp & 1) * 8 * NR % 64 + (lane >> 2);
   const double* Bs = sm + BKT * LDS_ + (lane & 3) * LDS_ + (lane >> 2);
   double acc[NC][NR][2];
   for (int j = 0; j < NC; ++j) for (int i = 0; i < NR; ++i) { acc[j][i][0] = 0; acc[j][i][1] = 0; }
   for (int it = 0; it < iters; ++it) {
      if (PIPE) {
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
               for (int i = 0; i < NR; ++i) af[nxt][i] = As[(kk + 4) * LDS_ + i * 8];
               #pragma unroll
               for (int j = 0; j < NC; ++j) bf[nxt][j] = Bs[(kk + 4) * LDS_ + j * 8];
            }
            #pragma unroll
            for (int j = 0; j < NC; ++j)
               #pragma unroll
               for (int i = 0; i < NR; ++i) dmma(acc[j][i][0], acc[j][i][1], bf[cur][j], af[cur][i]);
         }
      } else {
         #pragma unroll
         for (int kk = 0; kk < BKT; kk += 4) {
            double af[NR], bf[NC];
            #pragma unroll
            for (int i = 0; i < NR; ++i) af[i] = As[kk * LDS_ + i * 8];
            #pragma unroll
            for (int j = 0; j < NC; ++j) bf[j] = Bs[kk * LDS_ + j * 8];
            #pragma unroll
            for (int j = 0; j < NC; ++j)
               #pragma unroll
               for (int i = 0; i < NR; ++i) dmma(acc[j][i][0], acc[j][i][1], bf[j], af[i]);
         }
      }
      __syncwarp();
   }
   double s = 0;
   for (int j = 0; j < NC; ++j) for (int i = 0; i < NR; ++i) s += acc[j][i][0] + acc[j][i][1];
   out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NC, int NR, bool PIPE, int MAXT>
void run(const char* name, int warps, double* out) {
   constexpr int BKT = 32;
   auto kern = k<NC, NR, BKT, PIPE, MAXT>;
   if (warps * 32 > MAXT) return;
   int smem = 200 * 1024;
   cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
   int iters = 4000;
   cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
   kern<<<148, warps * 32, smem>>>(out, 100);
   cudaEventRecord(a);
   kern<<<148, warps * 32, smem>>>(out, iters);
   cudaEventRecord(b); cudaEventSynchronize(b);
   float ms; cudaEventElapsedTime(&ms, a, b);
   double flops = 148.0 * warps * iters * (BKT / 4) * NC * NR * 512.0;
   printf("%-28s warps/SM %2d  %7.2f TF/s  (%s)\n", name, warps, flops / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
}
int main() {
   double* out; cudaMalloc(&out, 148 * 512 * 8);
   for (int w : {4, 8}) run<4, 8, false, 256>("64x32 warp tile, unpipelined", w, out);
   for (int w : {4, 8}) run<4, 8, true, 256>("64x32 warp tile, pipelined", w, out);
   for (int w : {4, 8, 12, 16}) run<4, 4, true, 512>("32x32 warp tile, pipelined", w, out);
   for (int w : {4, 8, 12, 16}) run<2, 8, true, 512>("64x16 warp tile, pipelined", w, out);
   for (int w : {4, 8, 12, 16}) run<2, 4, true, 512>("32x16 warp tile, pipelined", w, out);
   for (int w : {4, 8, 12}) run<4, 4, true, 384>("32x32 warp tile, 384 thr", w, out);
   return 0;
}
