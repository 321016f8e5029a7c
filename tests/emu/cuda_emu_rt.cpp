/* TEST INFRASTRUCTURE (CPU only): fiber scheduler of the SIMT emulator (cuda_emu.h) and a synchronous mock of the
 * CUDA runtime calls the engine's host code makes ("device" memory is host memory, a stream executes at once). */
#include <ucontext.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <fcntl.h>
#include <unistd.h>
#include <atomic>
#include <string>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>
#include <stdexcept>
#include <vector>
#include "cuda_emu.h"

emu::Idx threadIdx, blockIdx, blockDim, gridDim;

namespace emu {
namespace {
constexpr size_t STACK = 256 << 10;
struct Fiber { ucontext_t ctx; char* stack = nullptr; bool done = true; int wait = 0; };   // wait: 0 runnable, 1 block barrier, 2 warp barrier
std::vector<Fiber> fibers;
ucontext_t sched;
int cur = -1, nthreads = 0;
const std::function<void()>* body_fn = nullptr;
std::vector<unsigned char> dyn;
struct WarpSlot { unsigned char v[32][16]; };
std::vector<WarpSlot> slots;
long n_switch = 0;
std::mutex big_lock;            // one launch at a time (host threads of different parts)

void entry() {
   (*body_fn)();
   fibers[cur].done = true;
   swapcontext(&fibers[cur].ctx, &sched);
}
void yield_to_scheduler() { ++n_switch; swapcontext(&fibers[cur].ctx, &sched); }
}

long switches() { return n_switch; }
void* dyn_smem() { return dyn.data(); }

void sync_block() { fibers[cur].wait = 1; yield_to_scheduler(); }
static void sync_warp() { fibers[cur].wait = 2; yield_to_scheduler(); }

void shfl_exchange(const void* in, void* out, size_t bytes, int src_lane) {
   const int lane = cur & 31, warp = cur >> 5;
   std::memcpy(slots[warp].v[lane], in, bytes);
   sync_warp();
   const int src = warp * 32 + (src_lane & 31);
   if (src < nthreads && (src_lane & 31) == src_lane) std::memcpy(out, slots[warp].v[src_lane & 31], bytes);
   else std::memcpy(out, in, bytes);
   sync_warp();
}

/* all-gather inside the warp: out[l] = the value lane l passed in (lanes that do not exist: zero bytes) */
void warp_gather(const void* in, void* out32, size_t bytes) {
   const int lane = cur & 31, warp = cur >> 5;
   std::memcpy(slots[warp].v[lane], in, bytes);
   sync_warp();
   for (int l = 0; l < 32; ++l) {
      if (warp * 32 + l < nthreads) std::memcpy((char*)out32 + l * bytes, slots[warp].v[l], bytes);
      else std::memset((char*)out32 + l * bytes, 0, bytes);
   }
   sync_warp();
}

void launch(unsigned grid, unsigned block, size_t smem_bytes, const std::function<void()>& body) {
   std::lock_guard<std::mutex> lock(big_lock);
   if (block == 0 || grid == 0) return;
   if (fibers.size() < block) {
      size_t old = fibers.size();
      fibers.resize(block);
      for (size_t i = old; i < block; ++i) {
         fibers[i].stack = (char*)mmap(nullptr, STACK, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
         if (fibers[i].stack == MAP_FAILED) throw std::runtime_error("emu: cannot map a fiber stack");
      }
   }
   slots.resize((block + 31) / 32);
   dyn.assign(smem_bytes + 64, 0);
   body_fn = &body;
   nthreads = (int)block;
   blockDim.x = block; blockDim.y = blockDim.z = 1;
   gridDim.x = grid; gridDim.y = gridDim.z = 1;
   for (unsigned b = 0; b < grid; ++b) {
      blockIdx.x = b;
      for (int t = 0; t < nthreads; ++t) {
         Fiber& f = fibers[t];
         getcontext(&f.ctx);
         f.ctx.uc_stack.ss_sp = f.stack; f.ctx.uc_stack.ss_size = STACK; f.ctx.uc_link = nullptr;
         makecontext(&f.ctx, entry, 0);
         f.done = false; f.wait = 0;
      }
      for (;;) {
         bool ran = false;
         int live = 0;
         for (int t = 0; t < nthreads; ++t) {
            Fiber& f = fibers[t];
            if (f.done) continue;
            ++live;
            if (f.wait) continue;
            cur = t; threadIdx.x = (unsigned)t;
            swapcontext(&sched, &f.ctx);
            ran = true;
         }
         if (!live) break;
         /* release barriers that every live participant has reached */
         bool released = false;
         int at_block = 0;
         live = 0;
         for (int t = 0; t < nthreads; ++t) if (!fibers[t].done) { ++live; if (fibers[t].wait == 1) ++at_block; }
         if (live && at_block == live) { for (int t = 0; t < nthreads; ++t) fibers[t].wait = 0; released = true; }
         for (int w = 0; w * 32 < nthreads; ++w) {
            int lw = 0, aw = 0;
            for (int t = w * 32; t < std::min(nthreads, w * 32 + 32); ++t) if (!fibers[t].done) { ++lw; if (fibers[t].wait == 2) ++aw; }
            if (lw && aw == lw) { for (int t = w * 32; t < std::min(nthreads, w * 32 + 32); ++t) fibers[t].wait = 0; released = true; }
         }
         if (!ran && !released && live) throw std::runtime_error("emu: divergent barrier (threads of a CTA wait at different rendezvous points)");
      }
   }
   cur = -1;
}
} // namespace emu

/* ---- the CUDA runtime, synchronously --------------------------------------------------------------------- */
namespace {
std::mutex reg_mtx;
std::map<const char*, size_t> allocs;          // base -> bytes
std::map<const char*, std::string> shm_names;  // base -> shared-memory name (SPRAL_B200_EMU_SHM)
struct ShmCleanup { ~ShmCleanup() { for (auto& kv : shm_names) shm_unlink(kv.second.c_str()); } } shm_cleanup;   // the pool keeps blocks until exit
bool is_dev(const void* p) {
   std::lock_guard<std::mutex> l(reg_mtx);
   auto it = allocs.upper_bound((const char*)p);
   if (it == allocs.begin()) return false;
   --it;
   return (const char*)p < it->first + it->second;
}
struct Ev { double t; };
double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
}

extern "C" {
cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
const char* cudaGetErrorString(cudaError_t) { return "emulated CUDA error"; }
/* "Device" memory: malloc, or -- SPRAL_B200_EMU_SHM=1 -- POSIX shared memory, so that an IPC handle (which then carries
 * the segment's name) can be opened by ANOTHER process: the one-process-per-GPU paths run as real processes on the CPU. */
static bool shm_mode() { static int v = -1; if (v < 0) v = getenv("SPRAL_B200_EMU_SHM") ? 1 : 0; return v == 1; }
cudaError_t cudaMallocHost(void** p, size_t bytes) { *p = std::malloc(bytes ? bytes : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaFreeHost(void* p) { std::free(p); return cudaSuccess; }
cudaError_t cudaMalloc(void** p, size_t bytes) {
   if (!bytes) bytes = 1;
   void* q = nullptr;
   std::string name;
   if (shm_mode()) {
      static std::atomic<long> counter{0};
      name = "/spral_emu_" + std::to_string((long)getpid()) + "_" + std::to_string(counter.fetch_add(1));
      int fd = shm_open(name.c_str(), O_CREAT | O_EXCL | O_RDWR, 0600);
      if (fd < 0) return cudaErrorMemoryAllocation;
      if (ftruncate(fd, (off_t)bytes) != 0) { close(fd); shm_unlink(name.c_str()); return cudaErrorMemoryAllocation; }
      q = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
      close(fd);
      if (q == MAP_FAILED) { shm_unlink(name.c_str()); return cudaErrorMemoryAllocation; }
      if (bytes <= ((size_t)64 << 20)) std::memset(q, 0xEE, bytes);
   } else {
      q = std::malloc(bytes);
      if (!q) return cudaErrorMemoryAllocation;
      std::memset(q, 0xEE, bytes);                       // device memory is not zero
   }
   { std::lock_guard<std::mutex> l(reg_mtx); allocs[(const char*)q] = bytes; if (!name.empty()) shm_names[(const char*)q] = name; }
   *p = q; return cudaSuccess;
}
cudaError_t cudaFree(void* p) {
   if (!p) return cudaSuccess;
   size_t bytes = 0; std::string name;
   {
      std::lock_guard<std::mutex> l(reg_mtx);
      auto it = allocs.find((const char*)p);
      if (it != allocs.end()) { bytes = it->second; allocs.erase(it); }
      auto jt = shm_names.find((const char*)p);
      if (jt != shm_names.end()) { name = jt->second; shm_names.erase(jt); }
   }
   if (!name.empty()) { munmap(p, bytes); shm_unlink(name.c_str()); }
   else std::free(p);
   return cudaSuccess;
}
cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { if (n) std::memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { if (n) std::memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpy2D(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, cudaMemcpyKind) {
   for (size_t j = 0; j < h; ++j) std::memmove((char*)d + j * dp, (const char*)s + j * sp, w);
   return cudaSuccess;
}
cudaError_t cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, cudaMemcpyKind k, cudaStream_t) {
   return cudaMemcpy2D(d, dp, s, sp, w, h, k); }
cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { std::memset(p, v, n); return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (cudaStream_t)std::malloc(8); return cudaSuccess; }
cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned, int) { *s = (cudaStream_t)std::malloc(8); return cudaSuccess; }
cudaError_t cudaDeviceGetStreamPriorityRange(int* lo, int* hi) { *lo = 0; *hi = -1; return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) { std::free((void*)s); return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = (cudaEvent_t) new Ev{0}; return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = (cudaEvent_t) new Ev{0}; return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete (Ev*)e; return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { ((Ev*)e)->t = now_ms(); return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventQuery(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)std::max(1e-3, ((Ev*)b)->t - ((Ev*)a)->t); return cudaSuccess; }
cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* at, const void* p) {
   std::memset(at, 0, sizeof(*at));
   at->type = is_dev(p) ? cudaMemoryTypeDevice : cudaMemoryTypeUnregistered;
   at->device = 0;
   return cudaSuccess;
}
cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr, int) { *v = 148; return cudaSuccess; }
cudaError_t cudaFuncSetAttribute(const void*, cudaFuncAttribute, int) { return cudaSuccess; }
cudaError_t cudaDeviceCanAccessPeer(int* can, int, int) { *can = 0; return cudaSuccess; }
cudaError_t cudaDeviceEnablePeerAccess(int, unsigned) { return cudaSuccess; }
cudaError_t cudaLaunchHostFunc(cudaStream_t, cudaHostFn_t fn, void* arg) { fn(arg); return cudaSuccess; }
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p) {
   std::memset(h, 0, sizeof(*h));
   if (shm_mode()) {
      std::lock_guard<std::mutex> l(reg_mtx);
      auto it = shm_names.find((const char*)p);
      if (it == shm_names.end() || it->second.size() > 55) return cudaErrorInvalidValue;      // must be the base of an allocation
      std::memcpy((char*)h, "SHM:", 4);
      std::memcpy((char*)h + 4, it->second.c_str(), it->second.size() + 1);
      return cudaSuccess;
   }
   std::memcpy(h, &p, sizeof(p)); return cudaSuccess;
}
cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned) {
   if (std::memcmp((const char*)&h, "SHM:", 4) == 0) {
      const char* name = (const char*)&h + 4;
      int fd = shm_open(name, O_RDWR, 0600);
      if (fd < 0) return cudaErrorInvalidValue;
      struct stat st;
      if (fstat(fd, &st) != 0) { close(fd); return cudaErrorInvalidValue; }
      void* q = mmap(nullptr, (size_t)st.st_size, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
      close(fd);
      if (q == MAP_FAILED) return cudaErrorInvalidValue;
      { std::lock_guard<std::mutex> l(reg_mtx); allocs[(const char*)q] = (size_t)st.st_size; }     // counts as device memory here too
      *p = q; return cudaSuccess;
   }
   std::memcpy(p, &h, sizeof(*p)); return cudaSuccess;
}
cudaError_t cudaIpcCloseMemHandle(void* p) {
   if (shm_mode()) {
      size_t bytes = 0;
      { std::lock_guard<std::mutex> l(reg_mtx); auto it = allocs.find((const char*)p); if (it != allocs.end()) { bytes = it->second; allocs.erase(it); } }
      if (bytes) munmap(p, bytes);
   }
   return cudaSuccess;
}
/* opt-in paths that the emulator does not serve */
cudaError_t cudaStreamBeginCapture(cudaStream_t, cudaStreamCaptureMode) { return cudaErrorNotSupported; }
cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t*) { return cudaErrorNotSupported; }
cudaError_t cudaGraphInstantiate(cudaGraphExec_t*, cudaGraph_t, unsigned long long) { return cudaErrorNotSupported; }
cudaError_t cudaGraphLaunch(cudaGraphExec_t, cudaStream_t) { return cudaErrorNotSupported; }
cudaError_t cudaGraphDestroy(cudaGraph_t) { return cudaSuccess; }
cudaError_t cudaGraphExecDestroy(cudaGraphExec_t) { return cudaSuccess; }
cudaError_t cudaLaunchCooperativeKernel(const void*, dim3, dim3, void**, size_t, cudaStream_t) { return cudaErrorNotSupported; }
cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, const void*, int, size_t) { *n = 1; return cudaSuccess; }
cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessorWithFlags(int* n, const void*, int, size_t, unsigned) { *n = 1; return cudaSuccess; }
}
