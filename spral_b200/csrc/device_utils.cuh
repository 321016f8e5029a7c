/* Small device helpers shared by the kernels of the B200 SSIDS engine. */
#pragma once
#include <math_constants.h>
#include "engine.h"
#include "pivot_state.h"

namespace b200 {

/* calc_ne(): pivot_state.h */

/* FP64 tensor-core MMA, m8n8k4: D(8x8) += A(8x4) B(4x8).  Lane l holds A[l/4][l%4], B[l%4][l/4] and
 * D[l/4][2(l%4) .. 2(l%4)+1].  (g++ build of the test emulator: the same contraction through a warp gather.) */
#ifdef __CUDACC__
__device__ __forceinline__ void pv_dmma(double& d0, double& d1, double a, double b) {
   asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
#else
inline void pv_dmma(double& d0, double& d1, double a, double b) {
   double av[32], bv[32];
   emu::warp_gather(&a, av, sizeof(double));
   emu::warp_gather(&b, bv, sizeof(double));
   const int lane = threadIdx.x & 31, i = lane >> 2, j0 = 2 * (lane & 3);
   for (int k = 0; k < 4; ++k) { d0 += av[i * 4 + k] * bv[j0 * 4 + k]; d1 += av[i * 4 + k] * bv[(j0 + 1) * 4 + k]; }
}
#endif


} // namespace b200
