/* TEST INFRASTRUCTURE (CPU only) -- a small SIMT emulator, so that the engine's real CUDA sources can be compiled
 * by g++ and run in this GPU-less container: tests/emu/build_emu.py turns `kernel<<<grid, block, smem, stream>>>(args)`
 * into emu::launch(...) and force-includes this header; __global__ kernels become plain functions that are run one
 * CTA after the other, every CUDA thread of a CTA on a fiber of its own (ucontext), so that __syncthreads() and the
 * warp shuffles are real rendezvous points.  __shared__ becomes `static` (one CTA runs at a time).  The CUDA runtime
 * calls of the host code are answered by tests/emu/cuda_emu_rt.cpp (synchronous streams, "device" memory = host
 * memory).  Nothing of this is part of the product: the library built from it (build/emu/libspral_ssids_b200_emu.so)
 * is loaded by tests only, to check host logic and kernel logic together against the oracle without a GPU. */
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <limits>

#undef __shared__
#define __shared__ static
#undef __global__
#define __global__
#undef __device__
#define __device__
#undef __host__
#define __host__
#undef __launch_bounds__
#define __launch_bounds__(...)
#undef __forceinline__
#define __forceinline__ inline
#undef __align__
#define __align__(n)
#undef CUDART_INF
#define CUDART_INF (std::numeric_limits<double>::infinity())

namespace emu {
struct Idx { unsigned x = 0, y = 0, z = 0; };
void launch(unsigned grid, unsigned block, size_t smem_bytes, const std::function<void()>& body);
void sync_block();
void* dyn_smem();
void shfl_exchange(const void* in, void* out, size_t bytes, int src_lane);
void warp_gather(const void* in, void* out32, size_t bytes);
long switches();
}
extern emu::Idx threadIdx, blockIdx, blockDim, gridDim;

inline void __syncthreads() { emu::sync_block(); }
inline void __threadfence() {}
template <class T> inline T __ldcg(const T* p) { return *p; }
template <class T> inline T __shfl_sync(unsigned, T v, int src) { T r; emu::shfl_exchange(&v, &r, sizeof(T), src); return r; }
template <class T> inline T __shfl_xor_sync(unsigned, T v, int off) {
   T r; emu::shfl_exchange(&v, &r, sizeof(T), (int)(threadIdx.x & 31) ^ off); return r; }
/* warp collectives used by diag_warp.cuh / the panel kernels: every lane of the warp takes part */
inline void __syncwarp(unsigned = 0xffffffffu) { int z = 0, o[32]; emu::warp_gather(&z, o, sizeof(int)); }
inline unsigned __ballot_sync(unsigned, int pred) {
   int v = pred ? 1 : 0, o[32];
   emu::warp_gather(&v, o, sizeof(int));
   unsigned r = 0;
   for (int l = 0; l < 32; ++l) if (o[l]) r |= 1u << l;
   return r;
}
inline unsigned __reduce_max_sync(unsigned, unsigned v) {
   unsigned o[32];
   emu::warp_gather(&v, o, sizeof(unsigned));
   unsigned r = v;
   for (int l = 0; l < 32; ++l) if (o[l] > r) r = o[l];
   return r;
}
inline long long __double_as_longlong(double d) { long long b; std::memcpy(&b, &d, 8); return b; }
inline double __longlong_as_double(long long b) { double d; std::memcpy(&d, &b, 8); return d; }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline int __double2hiint(double d) { long long b; std::memcpy(&b, &d, 8); return (int)(b >> 32); }
inline int __double2loint(double d) { long long b; std::memcpy(&b, &d, 8); return (int)(b & 0xffffffffLL); }
inline double __hiloint2double(int hi, int lo) { long long b = ((long long)(unsigned)hi << 32) | (unsigned)lo; double d; std::memcpy(&d, &b, 8); return d; }
inline double atomicAdd(double* p, double v) { double o = *p; *p = o + v; return o; }
inline int atomicAdd(int* p, int v) { int o = *p; *p = o + v; return o; }
inline unsigned atomicAdd(unsigned* p, unsigned v) { unsigned o = *p; *p = o + v; return o; }
inline int atomicMin(int* p, int v) { int o = *p; if (v < o) *p = v; return o; }
using std::min;
using std::max;
using std::isinf;
inline int min(int a, long b) { return (int)std::min<long>(a, b); }
/* the C++ convenience overloads of cuda_runtime.h exist only for nvcc */
template <class R, class... A> inline cudaError_t cudaFuncSetAttribute(R (*)(A...), cudaFuncAttribute, int) { return cudaSuccess; }
template <class R, class... A> inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, R (*)(A...), int, size_t) { *n = 1; return cudaSuccess; }
template <class R, class... A> inline cudaError_t cudaLaunchCooperativeKernel(R (*)(A...), dim3, dim3, void**, size_t, cudaStream_t) { return cudaErrorNotSupported; }
