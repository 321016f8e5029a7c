// Measures the FP64 roofline denominators on the box: (1) a register-resident
// DMMA (mma.sync.m8n8k4.f64) issue loop, (2) a DFMA loop, (3) cuBLAS DGEMM 8192^3.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu -lcublas
#include <cstdio>
#include <cuda_runtime.h>
#include <cublas_v2.h>

__global__ void __launch_bounds__(256) k_dmma(double* out, int iters) {
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

__global__ void __launch_bounds__(256) k_dfma(double* out, int iters) {
   double acc[16];
   for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 1e-9 + i;
   double a = 1.0 + threadIdx.x * 1e-12, b = 1e-12;
   for (int it = 0; it < iters; ++it) {
      #pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
   }
   double s = 0;
   for (int i = 0; i < 16; ++i) s += acc[i];
   out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
   cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
   int nsm = p.multiProcessorCount;
   double* out; cudaMalloc(&out, (size_t)nsm * 8 * 256 * sizeof(double));
   cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
   float ms;
   for (int wpb = 1; wpb <= 4; wpb *= 2) {
      int iters = 20000, blocks = nsm * wpb * 2;
      k_dmma<<<blocks, 256>>>(out, 100);
      cudaEventRecord(e0); k_dmma<<<blocks, 256>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms, e0, e1);
      double flops = (double)blocks * 8 * iters * 16 * 512.0;   // 8 warps, 16 mma of 8*8*4*2 flops
      printf("DMMA  blocks/SM=%d: %.2f TFLOP/s (%.2f ms)\n", wpb * 2, flops / ms / 1e9, ms);
      k_dfma<<<blocks, 256>>>(out, 100);
      cudaEventRecord(e0); k_dfma<<<blocks, 256>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms, e0, e1);
      flops = (double)blocks * 256 * iters * 16 * 2.0;
      printf("DFMA  blocks/SM=%d: %.2f TFLOP/s (%.2f ms)\n", wpb * 2, flops / ms / 1e9, ms);
   }
   cublasHandle_t h; cublasCreate(&h);
   for (int n : {4096, 8192}) {
      double *A, *B, *C; size_t sz = (size_t)n * n * sizeof(double);
      cudaMalloc(&A, sz); cudaMalloc(&B, sz); cudaMalloc(&C, sz);
      cudaMemset(A, 0, sz); cudaMemset(B, 0, sz); cudaMemset(C, 0, sz);
      double al = 1.0, be = 0.0;
      for (int i = 0; i < 2; ++i) cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_T, n, n, n, &al, A, n, B, n, &be, C, n);
      cudaEventRecord(e0);
      int reps = 5;
      for (int i = 0; i < reps; ++i) cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_T, n, n, n, &al, A, n, B, n, &be, C, n);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms, e0, e1);
      printf("cuBLAS DGEMM NT n=%d: %.2f TFLOP/s\n", n, 2.0 * n * n * n * reps / ms / 1e9);
      // SYRK-like shape of a front: C(10000x10000) = A(10000 x 2048) B^T
      cudaFree(A); cudaFree(B); cudaFree(C);
   }
   printf("device: %s, %d SMs, clock %d MHz\n", p.name, nsm, p.clockRate / 1000);
   return 0;
}
