/* Host side of the B200 SSIDS numeric engine: symbolic subtree, level-set
 * factorisation scheduler, solves, and the C ABI of include/spral_ssids_b200.h.
 *
 * Replaces (reference tree, ralna/spral):
 *   src/ssids/gpu/subtree.f90   construct_gpu_symbolic_subtree :76-160, factor :304-460,
 *                               build_child_pointers :162-196, build_rlist_direct :204-234,
 *                               solve_* :555-700, enquire/alter :702-981, get_contrib :522-538
 *   src/ssids/gpu/factor.f90    parfactor :42-153, subtree_factor_gpu :240-656,
 *                               assign_nodes_to_levels :824-879, transfer_contrib :155-221
 *   src/ssids/gpu/solve.f90     setup_gpu_solve :840-1007, fwd/bwd/d_solve_gpu
 *   driver/cuda_helper_gpu.f90  cuda_init :9-43
 * Results are defined by the reference CPU engine (src/ssids/cpu/NumericSubtree.hxx).
 *
 * Differences by design: the whole part is factorised level by level with ONE
 * host synchronisation per level (to learn how many columns each front
 * delayed, which sizes the parents); all pivoting decisions are taken on the
 * device; contribution blocks never leave HBM; every scratch buffer comes from
 * a grow-only pool that survives across factorisations.
 */
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <utility>
#include <new>
#include <stdexcept>
#include <vector>
#include <nvtx3/nvToolsExt.h>
#ifdef SPRAL_B200_SPLIT              /* distributed top front (split_front.h): opt-in build, not run on GPUs yet */
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <string>
#include <thread>
#endif
#include "engine.h"

namespace b200 {

struct CudaError { cudaError_t code; };
static bool g_profile = false;
static bool g_solve_graphs = false;  // SPRAL_B200_SOLVE_GRAPHS=1: replay the sweeps as CUDA graphs (no measured gain)
static bool g_solve_coop = false;    // SPRAL_B200_SOLVE_COOP=1: one cooperative launch per level (experimental: no measured gain yet)
static int g_solve_wide = 1;         // 256-column sweeps (solve_wide.h; SPRAL_B200_SOLVE_WIDE=0: 32-column steps everywhere) on
                                     // levels whose largest front has at least SPRAL_B200_SOLVE_WIDE_MIN (8) 32-column steps
static int g_solve_wide_min = 8;
/* look-ahead inside the wide sweeps only on levels of few, large fronts (two accumulators per front; a level of
 * thousands of fronts keeps the T kernels busy anyway) */
constexpr int SOLVE_LOOKAHEAD_MAX_FRONTS = 64;
/* fronts with at least this many eliminated columns get the inverses of their 32 x 32 diagonal blocks (8 KB per 32
 * columns; cfg5: 46 fronts, 23 MB); SPRAL_B200_SOLVE_LINV=0: none */
constexpr int SOLVE_LINV_MIN_COLS = 256;
static bool g_solve_linv = true;
static bool g_solve_lookahead = true; // SPRAL_B200_SOLVE_LOOKAHEAD=0: the wide sweeps on one stream
static bool g_trace_panels = false;  // SPRAL_B200_TRACE_PANELS=1: per-panel trace lines on stderr (host time between panels)
static bool g_lookahead = true;      // SPRAL_B200_LOOKAHEAD=0 disables the two-stream panel look-ahead
static int g_bulk_ctas = 0;          // SMs given to the overlapped bulk update (SPRAL_B200_BULK_CTAS); < 0: one CTA per tile
static bool g_panel_v2 = true;       // speculative 128-column panel segments (panel_v2.h) on levels of at most g_panel_v2_fronts
                                     // large fronts (SPRAL_B200_PANEL_V2=0: every panel step by step).  Measured on cfg5:
                                     // 376 -> 357 ms together with g_bulk_prio (profiles/r02_ab_variants.md)
static int g_panel_v2_fronts = 32;
static int g_ctile_block = 8;        // Schur-complement tiles in SB x SB blocked order for L2 reuse of the operand panels
                                     // (SPRAL_B200_CTILE_BLOCK=0: column by column).  ncu, largest launch of cfg5: DRAM reads
                                     // 20.8 -> 11.4 GB, L2 hit rate 59 -> 74 %, same 34.0 ms (profiles/r02_ncu_full_upd_contrib.md)
static bool g_bulk_prio = true;      // no static SM split: the panel stream has the highest stream priority and the bulk
                                     // update runs one tile per CTA on the lowest, so the panel kernels take SMs as bulk
                                     // tiles retire (SPRAL_B200_BULK_PRIO=0: persistent bulk kernel on SMs - 28)
/* Clears (and, with SPRAL_B200_DEBUG set, reports) a pending non-sticky CUDA
 * error so that it cannot leak into the host application's own CUDA calls. */
static void clear_cuda_error(const char* where) {
   cudaError_t e = cudaGetLastError();
   if (e != cudaSuccess && getenv("SPRAL_B200_DEBUG"))
      fprintf(stderr, "spral_ssids_b200: pending CUDA error %d (%s) cleared at exit of %s\n",
              (int)e, cudaGetErrorString(e), where);
}
struct AbiExit { const char* w; ~AbiExit() { clear_cuda_error(w); } };
#define ABI_GUARD() AbiExit abi_guard_{__func__}
std::atomic<long> g_launches{0};      // incremented by every launch wrapper
#define CUDA_TRY(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) throw CudaError{e_}; } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

/* Per-device cache of device allocations.  cudaMalloc / cudaFree are slow
 * (milliseconds per GB, much more once peer access between the GPUs of the box
 * is enabled, as NCCL and CUDA IPC do), so blocks released by one factorisation
 * are kept and handed to the next.  Every block is a whole cudaMalloc
 * allocation (IPC handles stay valid). */
class DevicePool {
   struct Blk { void* p; size_t bytes; };
   std::mutex mtx_;
   std::vector<Blk> free_[16];
   std::vector<std::pair<void*, size_t>> live_[16];      // size of every block handed out
   size_t cached_[16] = {0};
   static constexpr size_t kMaxCached = (size_t)96 << 30;
public:
   void* alloc(size_t bytes) {
      int dev = 0;
      CUDA_TRY(cudaGetDevice(&dev));
      bytes = align_up(std::max<size_t>(bytes, 256), 256);
      {
         std::lock_guard<std::mutex> lock(mtx_);
         auto& fl = free_[dev & 15];
         int best = -1;
         for (int i = 0; i < (int)fl.size(); ++i)
            if (fl[i].bytes >= bytes && fl[i].bytes <= bytes + bytes / 4 + (1 << 20) &&
                (best < 0 || fl[i].bytes < fl[best].bytes)) best = i;
         if (best >= 0) {
            Blk b = fl[best];
            fl.erase(fl.begin() + best);
            cached_[dev & 15] -= b.bytes;
            live_[dev & 15].push_back({b.p, b.bytes});
            return b.p;
         }
      }
      void* p = nullptr;
      cudaError_t e = cudaMalloc(&p, bytes);
      if (e != cudaSuccess) {            // out of memory: drop the cache and retry once
         cudaGetLastError();
         trim(dev, 0);
         CUDA_TRY(cudaMalloc(&p, bytes));
      }
      std::lock_guard<std::mutex> lock(mtx_);
      live_[dev & 15].push_back({p, bytes});
      return p;
   }
   void release(void* p) {
      if (!p) return;
      int dev = 0;
      cudaGetDevice(&dev);
      size_t bytes = 0;
      {
         std::lock_guard<std::mutex> lock(mtx_);
         auto& lv = live_[dev & 15];
         for (size_t i = 0; i < lv.size(); ++i)
            if (lv[i].first == p) { bytes = lv[i].second; lv[i] = lv.back(); lv.pop_back(); break; }
         if (bytes) { free_[dev & 15].push_back({p, bytes}); cached_[dev & 15] += bytes; }
      }
      if (!bytes) { cudaFree(p); return; }          // not ours (or another device): plain free
      if (cached_[dev & 15] > kMaxCached) trim(dev, kMaxCached / 2);
   }
   void trim(int dev, size_t keep) {
      std::vector<Blk> drop;
      {
         std::lock_guard<std::mutex> lock(mtx_);
         auto& fl = free_[dev & 15];
         while (!fl.empty() && cached_[dev & 15] > keep) {
            drop.push_back(fl.front());
            cached_[dev & 15] -= fl.front().bytes;
            fl.erase(fl.begin());
         }
      }
      for (auto& b : drop) cudaFree(b.p);
   }
};
static DevicePool g_pool;

/* Grow-only device buffer. */
struct Buf {
   void* p = nullptr;
   size_t cap = 0;
   /* Makes room for `bytes`; when the buffer has to be re-allocated the stream
    * is drained first because kernels in flight may still use the old one. */
   void ensure(size_t bytes, cudaStream_t s) {
      if (bytes <= cap) return;
      if (p) { CUDA_TRY(cudaStreamSynchronize(s)); g_pool.release(p); p = nullptr; cap = 0; }
      size_t want = align_up(bytes + bytes / 8, 256);
      p = g_pool.alloc(want);
      cap = want;
   }
   void release() { if (p) g_pool.release(p); p = nullptr; cap = 0; }
};

/* Bump allocator over a Buf. */
struct Bump {
   char* base = nullptr; size_t off = 0, cap = 0;
   void reset(Buf& b) { base = (char*)b.p; off = 0; cap = b.cap; }
   template <class T> T* take(size_t count) {
      size_t bytes = align_up(count * sizeof(T), 256);
      T* r = (T*)(base + off);
      off += bytes;
      if (off > cap) throw std::runtime_error("bump overflow");
      return r;
   }
};

#include "split_front.h"      /* empty unless built with -DSPRAL_B200_SPLIT */

/* ------------------------------------------------------------------------ */
/* Symbolic subtree                                                          */
/* ------------------------------------------------------------------------ */

/* grow-only pinned host buffer */
struct PinnedInts {
   int* p = nullptr; size_t cap = 0;
   int* ensure(size_t n) {
      if (n > cap) {
         if (p) cudaFreeHost(p);
         cap = std::max<size_t>(2 * n, 1 << 16);
         if (cudaMallocHost((void**)&p, cap * sizeof(int)) != cudaSuccess) { p = nullptr; cap = 0; cudaGetLastError(); throw std::bad_alloc(); }
      }
      return p;
   }
   ~PinnedInts() { if (p) cudaFreeHost(p); }
};

struct Symbolic {
   int device = 0, n = 0, sa = 0, en = 0, nloc = 0;
   spral_ssids_b200_options options;
   std::vector<int> n0, m0, parent;   // local parent (or -1)
   std::vector<char> exported;        // contribution block leaves the part
   std::vector<int64_t> rptr;         // nloc+1, 0-based offsets into rlist
   std::vector<int> rlist;            // global 1-based pivot-order indices
   std::vector<int> rlist_direct;     // 1-based position in the parent's row list
   std::vector<int> npassl;
   std::vector<int64_t> nptr;         // nloc+1, 0-based entry offsets into nlist
   std::vector<int> child_ptr, child_list;     // 0-based local, children in increasing order
   int nlevels = 0;
   std::vector<int> level_ptr, level_list;     // reference format: 1-based
   std::vector<int> front_of_node, node_of_front;
   std::vector<int> contrib_dest;     // local node per incoming contribution
   std::vector<std::vector<int>> contribs_of_node;   // incoming contribution indices per node
   /* device copies */
   int* d_rlist = nullptr; int* d_rlist_direct = nullptr;
   int64_t* d_nlist = nullptr; int64_t* d_nptr = nullptr; int* d_node_of_front = nullptr;
   /* sizes */
   size_t cbuf_bytes[2] = {0, 0};     // contribution ping-pong
   size_t ld_estimate = 0;            // doubles of LD scratch for the largest level (no delays)
   int64_t aval_len = 0;              // entries of aval this part reads: max source index of its nlist slice
   /* scratch pool shared by the factorisations / solves of this subtree */
   std::mutex mtx;
   Buf b_aval, b_scal, b_cbuf[2], b_ld, b_bk, b_ws, b_work, b_retry, b_x, b_y, b_pbuf, b_xt;
   Buf b_y2, b_pbuf2, b_xt2;          // second lane of the solves (two chunks of right-hand sides swept concurrently)
   PinnedInts snap_pinned;            // per-panel snapshot of the pivoting state (D2H)
   Buf b_bar;                         // arrival counter of the cooperative solve kernels
   Buf b_export[2];                   // packed contribution block handed to another process (IPC), double-buffered
   int export_slot = 0;
   Buf b_bulk[2];                     // tile lists of the look-ahead bulk updates (alternating panels)
   Buf b_segws;                       // chain workspaces of the speculative panel segments (panel_v2.h)
#ifdef SPRAL_B200_SPLIT
   std::string split_shm;             // shared-memory name of the split protocol (set for the root part only)
   int split_helpers = 1;             // helper ranks that serve it
#endif

   ~Symbolic() {
      cudaSetDevice(device);
      cudaFree(d_rlist); cudaFree(d_rlist_direct); cudaFree(d_nlist); cudaFree(d_nptr);
      cudaFree(d_node_of_front);
      b_aval.release(); b_scal.release(); b_cbuf[0].release(); b_cbuf[1].release();
      b_ld.release(); b_bk.release(); b_ws.release(); b_work.release(); b_x.release();
      b_y.release(); b_pbuf.release(); b_retry.release(); b_xt.release(); b_y2.release(); b_pbuf2.release(); b_xt2.release(); b_export[0].release(); b_export[1].release(); b_bar.release(); b_bulk[0].release(); b_bulk[1].release();
      b_segws.release();
   }
};

static Symbolic* build_symbolic(int device, int n, int sa, int en, const int* sptr,
      const int* sparent, const int64_t* rptr, const int* rlist, const int64_t* nptr,
      const int64_t* nlist, int ncontrib, const int* contrib_idx,
      const spral_ssids_b200_options* options) {
   auto* S = new Symbolic;
   S->device = device; S->n = n; S->sa = sa; S->en = en;
   const int nloc = S->nloc = en - sa;
   S->options = *options;
   S->n0.resize(nloc); S->m0.resize(nloc); S->parent.resize(nloc); S->exported.assign(nloc, 0);
   S->rptr.resize(nloc + 1); S->nptr.resize(nloc + 1);
   const int64_t rbase = nloc ? rptr[sa - 1] : 1, nbase = nloc ? nptr[sa - 1] : 1;
   for (int i = 0; i < nloc; ++i) {
      int node = sa + i;                      // global 1-based
      S->n0[i] = sptr[node] - sptr[node - 1];
      S->m0[i] = (int)(rptr[node] - rptr[node - 1]);
      int p = sparent[node - 1];
      S->parent[i] = (p < en) ? p - sa : -1;
      S->exported[i] = (p >= en && S->m0[i] > S->n0[i]);
      S->rptr[i] = rptr[node - 1] - rbase;
      S->nptr[i] = nptr[node - 1] - nbase;
   }
   S->rptr[nloc] = nloc ? rptr[en - 1] - rbase : 0;
   S->nptr[nloc] = nloc ? nptr[en - 1] - nbase : 0;
   S->rlist.assign(rlist + (rbase - 1), rlist + (rbase - 1) + S->rptr[nloc]);

   /* children, in increasing node order (build_child_pointers, gpu/subtree.f90:162-196) */
   S->child_ptr.assign(nloc + 2, 0);
   for (int i = 0; i < nloc; ++i) if (S->parent[i] >= 0) S->child_ptr[S->parent[i] + 1]++;
   for (int i = 0; i < nloc; ++i) S->child_ptr[i + 1] += S->child_ptr[i];
   S->child_list.resize(nloc);
   {
      std::vector<int> pos(S->child_ptr.begin(), S->child_ptr.end() - 1);
      for (int i = 0; i < nloc; ++i) if (S->parent[i] >= 0) S->child_list[pos[S->parent[i]]++] = i;
   }

   /* rlist_direct (build_rlist_direct, gpu/subtree.f90:204-234) */
   S->rlist_direct.assign(S->rlist.size(), 0);
   S->npassl.assign(nloc, 0);
   {
      std::vector<int> map(n + 1, 0);
      for (int i = 0; i < nloc; ++i) {
         int p = S->parent[i];
         if (p < 0) continue;
         for (int64_t ii = S->rptr[p]; ii < S->rptr[p + 1]; ++ii) map[S->rlist[ii]] = (int)(ii - S->rptr[p] + 1);
         for (int64_t ii = S->rptr[i]; ii < S->rptr[i + 1]; ++ii) S->rlist_direct[ii] = map[S->rlist[ii]];
         int cnt = 0;
         for (int64_t ii = S->rptr[i] + S->n0[i]; ii < S->rptr[i + 1]; ++ii)
            if (S->rlist_direct[ii] <= S->n0[p]) cnt++;
         S->npassl[i] = cnt;
      }
   }

   /* levels (assign_nodes_to_levels, gpu/factor.f90:824-879): level 1 = deepest */
   {
      std::vector<int> level(nloc + 1, 0), lvlcount(nloc + 2, 0);
      int num_levels = 1;
      for (int i = nloc - 1; i >= 0; --i) {
         int j = S->parent[i] >= 0 ? S->parent[i] : nloc;
         int lvl = level[j] + 1;
         level[i] = lvl;
         lvlcount[lvl]++;
         num_levels = std::max(num_levels, lvl);
      }
      S->nlevels = num_levels;
      S->level_ptr.assign(num_levels + 2, 1);
      for (int lvl = 2; lvl <= num_levels; ++lvl)
         S->level_ptr[lvl] = S->level_ptr[lvl - 1] + lvlcount[num_levels - (lvl - 1) + 1];
      /* level_ptr[k] (0-based k) holds reference lvlptr(k+1) once filled below */
      std::vector<int> ins(num_levels + 2, 1);
      ins[1] = 1;
      for (int lvl = 2; lvl <= num_levels; ++lvl) ins[lvl] = ins[lvl - 1] + lvlcount[num_levels - (lvl - 1) + 1];
      S->level_list.assign(nloc, 0);
      std::vector<int> start(ins);
      for (int i = 0; i < nloc; ++i) {
         int lvl = num_levels - level[i] + 1;
         S->level_list[ins[lvl] - 1] = i + 1;
         ins[lvl]++;
      }
      /* reference lvlptr(1..num_levels+1) */
      for (int lvl = 1; lvl <= num_levels; ++lvl) S->level_ptr[lvl - 1] = start[lvl];
      S->level_ptr[num_levels] = nloc + 1;
      S->level_ptr.resize(num_levels + 1);
   }
   S->front_of_node.resize(nloc); S->node_of_front.resize(nloc);
   for (int fi = 0; fi < nloc; ++fi) {
      int node = S->level_list[fi] - 1;
      S->node_of_front[fi] = node;
      S->front_of_node[node] = fi;
   }

   /* incoming contributions from other parts */
   S->contribs_of_node.assign(nloc, {});
   S->contrib_dest.resize(ncontrib);
   for (int k = 0; k < ncontrib; ++k) {
      int node = contrib_idx[k] - sa;
      S->contrib_dest[k] = node;
      if (node >= 0 && node < nloc) S->contribs_of_node[node].push_back(k);
   }

   /* static scratch sizes */
   for (int lev = 0; lev < S->nlevels; ++lev) {
      size_t cb = 0, ld = 0;
      for (int fi = S->level_ptr[lev] - 1; fi < S->level_ptr[lev + 1] - 1; ++fi) {
         int node = S->node_of_front[fi];
         size_t cm = S->m0[node] - S->n0[node];
         if (!S->exported[node]) cb += align_up(align_up(cm, 2) * cm * sizeof(double), 256);
         ld += align_up((size_t)S->m0[node], 2) * S->n0[node] + 32;
      }
      S->cbuf_bytes[lev & 1] = std::max(S->cbuf_bytes[lev & 1], cb);
      S->ld_estimate = std::max(S->ld_estimate, ld);
   }

   /* device copies */
   CUDA_TRY(cudaSetDevice(device));
   auto up = [](auto*& d, const auto& v) {
      using T = typename std::remove_reference<decltype(v)>::type::value_type;
      CUDA_TRY(cudaMalloc((void**)&d, std::max<size_t>(v.size(), 1) * sizeof(T)));
      if (!v.empty()) CUDA_TRY(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
   };
   up(S->d_rlist, S->rlist);
   up(S->d_rlist_direct, S->rlist_direct);
   up(S->d_nptr, S->nptr);
   up(S->d_node_of_front, S->node_of_front);
   size_t nent = (size_t)S->nptr[nloc];
   CUDA_TRY(cudaMalloc((void**)&S->d_nlist, std::max<size_t>(2 * nent, 1) * sizeof(int64_t)));
   if (nent) CUDA_TRY(cudaMemcpy(S->d_nlist, nlist + 2 * (nbase - 1), 2 * nent * sizeof(int64_t), cudaMemcpyHostToDevice));
   for (size_t i = 0; i < nent; ++i) S->aval_len = std::max(S->aval_len, nlist[2 * (nbase - 1) + 2 * i]);
   return S;
}

/* ------------------------------------------------------------------------ */
/* Numeric subtree                                                           */
/* ------------------------------------------------------------------------ */

struct Numeric {
   Symbolic* S = nullptr;
   bool posdef = false;
   cudaStream_t stream = nullptr;
   cudaStream_t stream2 = nullptr;     // bulk trailing updates overlapped with the next panel (look-ahead)
   cudaEvent_t ev_bulk = nullptr;       // recorded after the part of a bulk update the NEXT urgent update depends on
   cudaEvent_t ev_bulk_all = nullptr;   // recorded after the whole bulk update
   std::vector<void*> chunks;          // factor storage (L, D, perm), stream-ordered allocations
   char* chunk_base = nullptr; size_t chunk_off = 0, chunk_cap = 0;
   Front* d_fronts = nullptr;          // level order
   std::vector<Front> h_fronts;
   SolveFront* d_sfronts = nullptr;
   double* d_linv = nullptr; int* d_linv_bad = nullptr;       // inverse diagonal blocks of the large fronts (solves)
   RowTile* d_swork = nullptr; int* d_wbeg = nullptr;
   std::vector<int> swork_ptr, lvl_steps;
   size_t max_level_work = 0;
   /* exported contribution */
   int export_front = -1;
   double* d_export = nullptr;
   std::vector<double> h_cval, h_dval; std::vector<int> h_dperm;
   /* external contribution staging (device copies owned by this object) */
   std::vector<void*> ext_allocs;
   double timings[8] = {0};
   double class_ms[16] = {0};          // profiling mode: device ms per kernel class (ProfClass order)
   /* CUDA graphs of the solve sweeps, one per (job, right-hand sides per pass) */
   struct SolveGraph { int job, nr; cudaGraphExec_t exec; double* xs; double* ywork; double* pbuf; };
   std::vector<SolveGraph> graphs;
   cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
   /* profiling of the Schur-complement launches (enabled by spral_ssids_b200_set_profile) */
   std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;
   double prof_flops = 0;
   int n_launch = 0;

   int device = 0;                     // copy of S->device: the symbolic object may die first
#ifdef SPRAL_B200_SPLIT
   SplitOwner* split = nullptr;
   cudaStream_t stream3 = nullptr;     // pushes of the distributed front's panels to the helpers
#endif
   cudaStream_t lane2 = nullptr;       // second lane of the solves
   SolveAux solve_aux[2];              // look-ahead streams of the wide sweeps, one set per lane
   cudaEvent_t ev_lane_in = nullptr, ev_lane_out = nullptr;
   ~Numeric() {
      if (!S) return;
      cudaSetDevice(device);
#ifdef SPRAL_B200_SPLIT
      delete split;
      if (stream3) { cudaStreamSynchronize(stream3); cudaStreamDestroy(stream3); }
#endif
      if (stream) cudaStreamSynchronize(stream);
      if (stream2) { cudaStreamSynchronize(stream2); cudaStreamDestroy(stream2); }
      if (lane2) { cudaStreamSynchronize(lane2); cudaStreamDestroy(lane2); }
      solve_aux[0].destroy(); solve_aux[1].destroy();
      if (ev_lane_in) cudaEventDestroy(ev_lane_in);
      if (ev_lane_out) cudaEventDestroy(ev_lane_out);
      for (auto& g : graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
      if (ev_bulk) cudaEventDestroy(ev_bulk);
      if (ev_bulk_all) cudaEventDestroy(ev_bulk_all);
      if (ev_begin) cudaEventDestroy(ev_begin);
      if (ev_end) cudaEventDestroy(ev_end);
      for (auto& e : prof_events) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
      for (void* p : chunks) g_pool.release(p);
      for (void* p : ext_allocs) g_pool.release(p);
      if (d_linv) g_pool.release(d_linv);
      if (d_linv_bad) g_pool.release(d_linv_bad);
      g_pool.release(d_fronts); g_pool.release(d_sfronts); g_pool.release(d_swork); g_pool.release(d_wbeg);
      g_pool.release(d_export);
      if (stream) cudaStreamDestroy(stream);
   }

   /* factor storage: chunked bump allocation */
   void* falloc(size_t bytes) {
      bytes = align_up(bytes, 256);
      if (chunk_off + bytes > chunk_cap) {
         size_t want = std::max(bytes, (size_t)256 << 20);
         void* p = g_pool.alloc(want);
         chunks.push_back(p);
         chunk_base = (char*)p; chunk_off = 0; chunk_cap = want;
      }
      void* r = chunk_base + chunk_off;
      chunk_off += bytes;
      return r;
   }
   void reserve(size_t bytes) {
      if (chunk_cap - chunk_off >= bytes) return;
      void* p = g_pool.alloc(align_up(bytes, 256));
      chunks.push_back(p);
      chunk_base = (char*)p; chunk_off = 0; chunk_cap = align_up(bytes, 256);
   }
};

/* Per-class kernel timing (profiling mode only): CUDA events around every launch. */
enum ProfClass { PC_DIAG = 0, PC_APPLY, PC_COMMIT, PC_INNER, PC_SWAP, PC_OUTER, PC_CONTRIB, PC_ASSEMBLE, PC_INIT, PC_COUNT };
struct Prof {
   struct Rec { int cls; cudaEvent_t a, b; double flops; int ntiles; };
   double next_flops = 0; int next_tiles = 0;   // annotation for the next record (trace output)
   std::vector<Rec> recs;
   double ms[PC_COUNT] = {0};
   cudaEvent_t begin(cudaStream_t s) {
      if (!g_profile) return nullptr;
      cudaEvent_t a; cudaEventCreate(&a); cudaEventRecord(a, s); return a;
   }
   void end(int cls, cudaEvent_t a, cudaStream_t s) {
      if (!a) return;
      cudaEvent_t b; cudaEventCreate(&b); cudaEventRecord(b, s);
      recs.push_back({cls, a, b, next_flops, next_tiles});
      next_flops = 0; next_tiles = 0;
   }
   void collect() {       // call after a stream synchronisation
      std::vector<Rec> keep;
      for (auto& r : recs) {
         if (cudaEventQuery(r.b) != cudaSuccess) { cudaGetLastError(); keep.push_back(r); continue; }
         float t = 0; cudaEventElapsedTime(&t, r.a, r.b); ms[r.cls] += t;
         if (r.flops > 0 && getenv("SPRAL_B200_TRACE"))
            fprintf(stderr, "[launch] class %d tiles %d flops %.3e ms %.3f -> %.2f TF/s\n", r.cls, r.ntiles, r.flops, t,
                    r.flops / t / 1e9);
         cudaEventDestroy(r.a); cudaEventDestroy(r.b);
      }
      recs.swap(keep);
   }
};
static thread_local Prof* g_prof = nullptr;
#define PROF(cls, stmt) do { cudaEvent_t pe_ = g_prof ? g_prof->begin(s) : nullptr; stmt; if (pe_) g_prof->end(cls, pe_, s); } while (0)
#define PROF_ON(cls, strm, stmt) do { cudaEvent_t pe_ = g_prof ? g_prof->begin(strm) : nullptr; stmt; if (pe_) g_prof->end(cls, pe_, strm); } while (0)

template <class T>
static T* upload(Bump& bump, const std::vector<T>& v, cudaStream_t s) {
   T* d = bump.take<T>(std::max<size_t>(v.size(), 1));
   if (!v.empty()) CUDA_TRY(cudaMemcpyAsync(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s));
   return d;
}

static bool is_device_pointer(const void* p, int* device) {
   cudaPointerAttributes at;
   if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
   if (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged) { if (device) *device = at.device; return true; }
   return false;
}

/* Host mirror of a front's pivoting state at panel boundaries.  The device
 * takes every pivoting decision; the host only learns, once per outer panel
 * (k_snapshot + one small D2H), how many columns were eliminated / failed, so
 * that it can size the next launches exactly (no idle tiles). */
struct HostState {
   int fi = 0, m = 0, n = 0;
   int done = 0, end = 0, pass_start = 0, p0 = 0, pend0 = 0, pend = 0;
   bool finished = false;
   bool spec_dead = false;      // panel_v2: the front gave up SPEC_MAX_FAILS segments, no more speculative launches
};

/* Factorises the n fully-summed columns of every front in `fronts` (indices
 * into the level-ordered Front array).  Panels of PW columns, inner steps of
 * BS; failed columns are retried in later passes until a pass eliminates
 * nothing.  Work buffers come from `wb` (grow-only, re-used panel by panel:
 * every panel ends with a stream synchronisation). */
static int factor_fronts(Numeric& N, Front* d_fronts, const std::vector<Front>& F,
      const std::vector<int>& fronts, bool big, const FactorParams& prm, Buf& wb, double& t_sync) {
   cudaStream_t s = N.stream;
   const bool posdef = N.posdef;
   const int T = update_tile_size(big), Ti = inner_tile_size(big);
   std::vector<HostState> H(fronts.size());
   for (size_t i = 0; i < fronts.size(); ++i) {
      HostState& h = H[i];
      const Front& f = F[fronts[i]];
      h.fi = fronts[i]; h.m = f.m; h.n = f.n;
      h.done = 0; h.end = f.n; h.pass_start = 0;
      h.finished = (f.n == 0);
      h.p0 = 0; h.pend0 = std::min(PW, f.n); h.pend = h.pend0;     // advance_state() opens the same panel
   }
   /* the snapshot lands in pinned memory (a pageable destination makes the copy synchronous and slower); the buffer
    * belongs to the symbolic subtree, so it is allocated once, not per factorisation or thread */
   PinnedInts& snap_pinned = N.S->snap_pinned;
   int* snap_host = nullptr;
   int err = 0;
   /* SPRAL_B200_TRACE_PANELS=2: a timeline of the level on both streams (events behind every update launch) */
   struct TlRec { int p0; char what; cudaEvent_t a, b; int tiles; };
   std::vector<TlRec> tl;
   cudaEvent_t tl_base = nullptr;
   const bool timeline = g_trace_panels && big && getenv("SPRAL_B200_TRACE_PANELS")[0] == '2';
   if (timeline) { cudaEventCreate(&tl_base); cudaStreamSynchronize(N.stream2); cudaStreamSynchronize(s); cudaEventRecord(tl_base, s); }
   auto tl_open = [&](cudaStream_t st) { cudaEvent_t e = nullptr; if (timeline) { cudaEventCreate(&e); cudaEventRecord(e, st); } return e; };
   auto tl_close = [&](int p0, char what, cudaEvent_t a, cudaStream_t st, int tiles) {
      if (!timeline) return;
      cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); tl.push_back({p0, what, a, e, tiles});
   };
   bool bulk_pending = false;       // a bulk update is (possibly) still running on stream2
   int bulk_parity = 0;
   for (;;) {
      /* active fronts, most candidates first (so that per-step launches use a prefix) */
      std::vector<int> act;
      for (size_t i = 0; i < H.size(); ++i) if (!H[i].finished) act.push_back((int)i);
      if (act.empty()) {
         if (bulk_pending) CUDA_TRY(cudaStreamWaitEvent(s, N.ev_bulk_all, 0));     // join
         if (timeline) {
            cudaStreamSynchronize(N.stream2); cudaStreamSynchronize(s);
            for (auto& r : tl) {
               float t0 = 0, t1 = 0; cudaEventElapsedTime(&t0, tl_base, r.a); cudaEventElapsedTime(&t1, tl_base, r.b);
               fprintf(stderr, "[timeline] p0 %5d %c tiles %5d  %9.3f -> %9.3f ms (%.3f)\n", r.p0, r.what, r.tiles, t0, t1, t1 - t0);
               cudaEventDestroy(r.a); cudaEventDestroy(r.b);
            }
            cudaEventDestroy(tl_base);
         }
         break;
      }
      std::stable_sort(act.begin(), act.end(), [&](int a, int b) {
         return H[a].pend0 - H[a].p0 > H[b].pend0 - H[b].p0; });
      const int na_all = (int)act.size();
      std::vector<int> flist(na_all), cand(na_all), rows_prefix(1, 0), inner_prefix(1, 0);
      std::vector<RowTile> rows;
      std::vector<MatTile> inner;
      for (int k = 0; k < na_all; ++k) {
         const HostState& h = H[act[k]];
         flist[k] = h.fi; cand[k] = h.pend0 - h.p0;
         int ntr = (h.m + RT - 1) / RT;
         for (int t = 0; t < ntr; ++t) rows.push_back({h.fi, t});
         rows_prefix.push_back((int)rows.size());
         /* inner updates touch columns [p0, pend0) of rows >= p0 */
         int mti = (h.m + Ti - 1) / Ti;
         for (int tj = h.p0 / Ti; tj <= (h.pend0 - 1) / Ti; ++tj)
            for (int ti = tj; ti < mti; ++ti) inner.push_back({h.fi, ti, tj});
         inner_prefix.push_back((int)inner.size());
      }
      size_t need = 8192 + flist.size() * 2 * sizeof(int) * 8 + rows.size() * sizeof(RowTile) + inner.size() * sizeof(MatTile);
      /* room for the outer list of this panel as well (bounded by all lower tiles) */
      size_t outer_max = 0;
      for (int k = 0; k < na_all; ++k) {
         const HostState& h = H[act[k]];
         size_t mt = (h.m + T - 1) / T, nt = (h.n + T - 1) / T;
         outer_max += mt * nt;
      }
      need += outer_max * sizeof(MatTile) + rows.size() * sizeof(RowTile);
      wb.ensure(need, s);
      Bump bump; bump.reset(wb);
      int* d_flist = upload(bump, flist, s);
      RowTile* d_rows = upload(bump, rows, s);
      MatTile* d_inner = upload(bump, inner, s);
      int* d_snap = bump.take<int>((size_t)na_all * 8);

      const int maxcand = cand[0];
      const int nsteps = (maxcand + BS - 1) / BS;
      auto count_gt = [&](int thr) {
         int lo = 0, hi = na_all;
         while (lo < hi) { int mid = (lo + hi) / 2; if (cand[mid] > thr) lo = mid + 1; else hi = mid; }
         return lo;
      };
      double t_wait_panel = 0.0;
      const auto t_panel0 = std::chrono::steady_clock::now();
      auto take_snapshot = [&]() {
         launch_snapshot(d_fronts, d_flist, na_all, d_snap, s);
         snap_host = snap_pinned.ensure((size_t)na_all * 8);
         CUDA_TRY(cudaMemcpyAsync(snap_host, d_snap, (size_t)na_all * 8 * sizeof(int), cudaMemcpyDeviceToHost, s));
         auto ts0 = std::chrono::steady_clock::now();
         CUDA_TRY(cudaStreamSynchronize(s));
         const double dtw = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - ts0).count();
         t_sync += dtw; t_wait_panel += dtw;
         if (g_prof) g_prof->collect();
      };
      /* Speculative path (panel_v2.h): the panel in PW / CW segments of three launches each (chain, tiles,
       * commit) plus one DMMA update of the panel's remaining columns between segments.  A segment that meets
       * a failed or zero pivot changes nothing; the step-by-step loop below then finishes the panel. */
      bool steps_new_panel = true;
      int steps_todo = nsteps;
      bool v2 = F[fronts[0]].sws != nullptr;
      if (v2) {                              // nothing to gain once every active front has stopped speculating
         bool any_alive = false;
         for (int k = 0; k < na_all; ++k) any_alive = any_alive || !H[act[k]].spec_dead;
         v2 = any_alive;
      }
      if (v2) {
         const int nseg = PW / panel_segment_width();
         cudaEvent_t tle = tl_open(s);
         for (int seg = 0; seg < nseg; ++seg) {
            PROF(PC_DIAG, launch_panel_chain(d_fronts, d_flist, na_all, posdef, seg == 0, prm, s));
            PROF(PC_APPLY, launch_panel_tiles(d_fronts, d_rows, rows_prefix[na_all], posdef, prm, s));
            PROF(PC_COMMIT, launch_seg_commit(d_fronts, d_rows, rows_prefix[na_all], posdef, s));
            if (seg + 1 < nseg) PROF(PC_INNER, launch_update(d_fronts, d_inner, inner_prefix[na_all], UPD_SEG, Ti == 128, s));   // the tile size the list was built with
         }
         tl_close(H[act[0]].p0, 'C', tle, s, rows_prefix[na_all]);
         take_snapshot();
         int maxrem = 0;
         for (int k = 0; k < na_all; ++k) {
            const int* sn = &snap_host[(size_t)k * 8];          // p0, done, pend, pend0, end, finished, flag
            if (sn[6] < 0 || sn[5]) continue;
            maxrem = std::max(maxrem, sn[2] - sn[1]);
         }
         steps_new_panel = false;                                // the chain kernel opened the panel
         steps_todo = (maxrem + BS - 1) / BS;
      }
      for (int st = 0; st < steps_todo; ++st) {
         int na = v2 ? na_all : count_gt(st * BS);
         if (na == 0) break;
         PROF(PC_DIAG, launch_diag(d_fronts, d_flist, na, posdef, st == 0 && steps_new_panel, prm, s));
         PROF(PC_APPLY, launch_apply(d_fronts, d_rows, rows_prefix[na], posdef, prm, s));
         if (!posdef) PROF(PC_COMMIT, launch_commit(d_fronts, d_rows, rows_prefix[na], s));
         PROF(PC_INNER, launch_update(d_fronts, d_inner, inner_prefix[na], UPD_INNER, big, s));
         if (!posdef) PROF(PC_SWAP, launch_swap(d_fronts, d_rows, rows_prefix[na], false, s));
      }
      if (!v2 || steps_todo > 0) take_snapshot();

      /* what happened in the panel; exact outer-update and swap work.  Look-ahead:
       * when no front failed a pivot in this panel, only the tile columns that hold
       * the NEXT panel are updated on the main stream; the rest of the trailing
       * update (the bulk of the flops) runs on the second stream, on a capped
       * number of SMs, concurrently with the next panel's latency-bound steps.
       * The two touch disjoint rows/columns. */
      std::vector<MatTile> outer, bulk, bulk_b;   // bulk: columns of the panel after next; bulk_b: the rest
      std::vector<int4> bulk_regs;          // {front, k0, k1, c_lo} of every front in the bulk lists
      std::vector<RowTile> swap_rows;
      bool any_fail = false;
      std::vector<char> failed(na_all, 0);
      for (int k = 0; k < na_all; ++k) {
         const int* sn = &snap_host[(size_t)k * 8];
         if (sn[6] >= 0 && H[act[k]].pend0 - sn[2] > 0) { failed[k] = 1; any_fail = true; }
      }
      /* look-ahead is decided front by front: on a level of several fronts one failed pivot must not put the whole
       * trailing update of every other front on the main stream (the fronts are independent; a front with a failed
       * pivot gets its full update and its swaps in order, behind the whole bulk backlog) */
      const bool lookahead = big && !any_fail && g_lookahead;
      const bool la_level = big && g_lookahead;
#ifdef SPRAL_B200_SPLIT
      /* distributed top front (split_front.h; protocol: tests/c/dist_front_emu.cpp): while the split is active the far
       * columns live on the helper; the first panel with a failed pivot brings them back and ends it */
      bool split_now = false, split_restart = false;
      int split_k = 0;
      if (N.split && N.split->active) {
         const HostState& h0 = H[act[0]];
         split_k = N.split->panel_of(h0.p0);
         /* the split lives on full panels that are its blocks; a failed pivot (columns swapped across the front), a short
          * panel (candidates parked at the end) or a new pass ends it */
         if (any_fail || !lookahead || na_all != 1 || split_k < 0 || h0.pend0 != h0.p0 + PW) {
            const int kd = std::max(0, (h0.p0 - N.split->p_first + PW - 1) / PW);      // panels the helper was given: 0 .. kd-1
            N.split->drain(kd, kd + 1, s, N.stream3);
         } else split_now = true;
      } else if (N.split && N.split->level_ok && N.split->restart && !N.split->dead && !any_fail && lookahead && na_all == 1) {
         /* a drained split starts again behind a clean full panel of the first pass order (sporadic failures must not
          * cost the rest of a 64-panel front): this panel is still updated here, the far columns go out behind its bulk */
         const HostState& h0 = H[act[0]];
         const int* sn0 = &snap_host[0];
         split_restart = sn0[6] >= 0 && h0.pend0 == h0.p0 + PW && sn0[1] == h0.pend0;
      }
#endif
      for (int k = 0; k < na_all; ++k) {
         HostState& h = H[act[k]];
         const int* sn = &snap_host[(size_t)k * 8];   // p0, done, pend, pend0, end, finished, flag
         if (sn[6] < 0) { err = err ? std::max(err, sn[6]) : sn[6]; h.finished = true; continue; }
         if (sn[0] != h.p0 || sn[3] != h.pend0 || sn[4] != h.end)
            throw std::runtime_error("host mirror of the pivoting state diverged from the device");
         h.done = sn[1]; h.pend = sn[2];
         h.spec_dead = sn[7] >= SPEC_MAX_FAILS;
         if (h.done > h.p0 && h.pend0 < h.n) {
            if (g_prof) {
               double K = h.done - h.p0, nn = h.n, c0 = h.pend0, mm = h.m;
               g_prof->next_flops += 2.0 * K * ((nn - c0) * mm - (nn * (nn - 1) - c0 * (c0 - 1)) / 2.0);
            }
            int mt = (h.m + T - 1) / T, nt = (h.n + T - 1) / T;
            /* last tile column that holds a column of the next panel */
            int tj_urgent = (std::min(h.pend0 + PW, h.n) - 1) / T;
            int tj_next = (std::min(h.pend0 + 2 * PW, h.n) - 1) / T;     // last tile column of the panel after next
            bool has_bulk = la_level && !failed[k] && tj_urgent + 1 < nt;
            if (has_bulk) bulk_regs.push_back(make_int4(h.fi, h.p0, h.done, (tj_urgent + 1) * T));
            for (int tj = h.pend0 / T; tj < nt; ++tj) {
#ifdef SPRAL_B200_SPLIT
               /* distributed front: the tile columns of the blocks a helper holds are updated there */
               if (split_now && tj > tj_urgent) {
                  const int J = N.split->block_of_tile(tj, T);
                  if (J >= 2 && !N.split->is_local(J)) continue;
               }
#endif
               for (int ti = tj; ti < mt; ++ti) {
                  if (has_bulk && tj > tj_next) bulk_b.push_back({(int)bulk_regs.size() - 1, ti, tj});
                  else if (has_bulk && tj > tj_urgent) bulk.push_back({(int)bulk_regs.size() - 1, ti, tj});
                  else outer.push_back({h.fi, ti, tj});
               }
            }
         }
         if (!posdef && h.pend0 - h.pend > 0 && h.end - h.pend0 > 0) {
            int ntr = (h.m + RT - 1) / RT;
            for (int t = 0; t < ntr; ++t) swap_rows.push_back({h.fi, t});
         }
      }
      if (err) { if (bulk_pending) cudaStreamSynchronize(N.stream2); return err; }
      if (la_level && (int)(bulk.size() + bulk_b.size()) < device_sm_count()) {   // not worth a second stream
         for (const MatTile& t : bulk) outer.push_back({bulk_regs[t.front].x, t.ti, t.tj});
         for (const MatTile& t : bulk_b) outer.push_back({bulk_regs[t.front].x, t.ti, t.tj});
         bulk.clear(); bulk_b.clear();
      }
      const bool have_bulk = !bulk.empty() || !bulk_b.empty();
      if (g_trace_panels) {              /* SPRAL_B200_TRACE_PANELS=1: one line per panel of a level of large fronts */
         static thread_local std::chrono::steady_clock::time_point last = std::chrono::steady_clock::now();
         auto now = std::chrono::steady_clock::now();
         const HostState& h0 = H[act[0]];
         int failed_cols = 0;
         for (int k = 0; k < na_all; ++k) failed_cols += H[act[k]].pend0 - H[act[k]].pend;
         if (big) fprintf(stderr, "[panel] fronts %d first(m %d n %d p0 %d done %d pend0 %d) failed_cols %d v2 %d steps %d "
                 "lookahead %d tiles urgent %zu bulk %zu+%zu swap %zu  %.1f us since the previous panel "
                 "(this panel: %.1f us from its first launch to the snapshot, of which %.1f us waiting in the sync)\n",
                 na_all, h0.m, h0.n, h0.p0, h0.done, h0.pend0, failed_cols, (int)v2, steps_todo, (int)lookahead,
                 outer.size(), bulk.size(), bulk_b.size(), swap_rows.size(),
                 std::chrono::duration<double, std::micro>(now - last).count(),
                 std::chrono::duration<double, std::micro>(now - t_panel0).count(), 1e3 * t_wait_panel);
         last = now;
      }
      if (bulk_pending && (!outer.empty() || !swap_rows.empty())) {
         /* With look-ahead the urgent update only touches the next panel's columns,
          * which the previous bulk update finished first (ev_bulk); a full update or
          * a swap touches everything, so it waits for the whole backlog. */
         CUDA_TRY(cudaStreamWaitEvent(s, (lookahead && have_bulk) ? N.ev_bulk : N.ev_bulk_all, 0));
         if (!(lookahead && have_bulk)) bulk_pending = false;
      }
#ifdef SPRAL_B200_SPLIT
      if (split_now) N.split->need_block(split_k + 1, s);
#endif
      if (!outer.empty()) {
         MatTile* d_outer = upload(bump, outer, s);
         if (g_prof) g_prof->next_tiles = (int)outer.size();
         cudaEvent_t tle = tl_open(s);
         PROF(PC_OUTER, launch_update(d_fronts, d_outer, (int)outer.size(), UPD_OUTER, big, s));
         tl_close(H[act[0]].p0, 'U', tle, s, (int)outer.size());
      }
#ifdef SPRAL_B200_SPLIT
      if (split_now) {
         const HostState& h0 = H[act[0]];
         /* the panel travels on a stream of its own (the host has synchronised the panel; the bulk stream may be
          * busy with the owner's share of the previous panel for a long time) */
         if (N.split->has_far(split_k)) N.split->push_panel(split_k, h0.p0, h0.done, N.stream3);
         else N.split->end_front(s);          // nothing is left on a helper
      }
#endif
      if (have_bulk) {
         cudaStream_t s2 = N.stream2;
         Buf& bb = N.S->b_bulk[bulk_parity];
         bulk_parity ^= 1;
         size_t regs_bytes = align_up(bulk_regs.size() * sizeof(int4), 256);
         size_t a_bytes = align_up(bulk.size() * sizeof(MatTile), 256);
         size_t need_b = regs_bytes + a_bytes + bulk_b.size() * sizeof(MatTile) + 256;
         if (need_b > bb.cap) bb.ensure(need_b * 2 + 4096, s2);       // drains stream2 before re-allocating
         CUDA_TRY(cudaMemcpyAsync(bb.p, bulk_regs.data(), bulk_regs.size() * sizeof(int4), cudaMemcpyHostToDevice, s2));
         MatTile* d_a = (MatTile*)((char*)bb.p + regs_bytes);
         MatTile* d_b = (MatTile*)((char*)bb.p + regs_bytes + a_bytes);
         if (!bulk.empty()) {
            CUDA_TRY(cudaMemcpyAsync(d_a, bulk.data(), bulk.size() * sizeof(MatTile), cudaMemcpyHostToDevice, s2));
            cudaEvent_t tle = tl_open(s2);
            PROF_ON(PC_OUTER, s2, launch_update(d_fronts, d_a, (int)bulk.size(), UPD_EXPLICIT, big, s2,
                                                 g_bulk_ctas, (const int4*)bb.p));
            tl_close(H[act[0]].p0, 'A', tle, s2, (int)bulk.size());
         }
         CUDA_TRY(cudaEventRecord(N.ev_bulk, s2));       // the panel after next has all its updates from this panel
         if (!bulk_b.empty()) {
            CUDA_TRY(cudaMemcpyAsync(d_b, bulk_b.data(), bulk_b.size() * sizeof(MatTile), cudaMemcpyHostToDevice, s2));
            cudaEvent_t tle = tl_open(s2);
            PROF_ON(PC_OUTER, s2, launch_update(d_fronts, d_b, (int)bulk_b.size(), UPD_EXPLICIT, big, s2,
                                                 g_bulk_ctas, (const int4*)bb.p));
            tl_close(H[act[0]].p0, 'B', tle, s2, (int)bulk_b.size());
         }
         CUDA_TRY(cudaEventRecord(N.ev_bulk_all, s2));
         bulk_pending = true;
      }
#ifdef SPRAL_B200_SPLIT
      if (split_restart) {
         /* the far columns got this panel's update on either stream: the main stream joins the bulk, then copies them */
         if (bulk_pending) { CUDA_TRY(cudaStreamWaitEvent(s, N.ev_bulk_all, 0)); bulk_pending = false; }
         N.split->begin_front(F[H[act[0]].fi], posdef, s, H[act[0]].pend0);
      }
#endif
      if (!swap_rows.empty()) {
         RowTile* d_sw = upload(bump, swap_rows, s);
         PROF(PC_SWAP, launch_swap(d_fronts, d_sw, (int)swap_rows.size(), true, s));
      }
      /* mirror of advance_state(new_panel = true): close the panel, open the next */
      for (int k = 0; k < na_all; ++k) {
         HostState& h = H[act[k]];
         if (h.finished) continue;
         h.end -= h.pend0 - h.pend;
         if (h.done == h.end) {
            if (h.end == h.n) h.finished = true;
            else if (h.done > h.pass_start) { h.pass_start = h.done; h.end = h.n; }
            else h.finished = true;
         }
         if (!h.finished) { h.p0 = h.done; h.pend0 = std::min(h.done + PW, h.end); h.pend = h.pend0; }
      }
   }
   return 0;
}

static void factor_subtree(Numeric& N, const double* aval_in, const double* scaling_in,
      void** child_contrib, const spral_ssids_b200_options* opt, spral_ssids_b200_stats* stats) {
   Symbolic& S = *N.S;
   const int nloc = S.nloc;
   const bool posdef = N.posdef;
   std::lock_guard<std::mutex> lock(S.mtx);
   CUDA_TRY(cudaSetDevice(S.device));
   if (const char* e = getenv("SPRAL_B200_BULK_PRIO")) g_bulk_prio = atoi(e) != 0;
   if (const char* e = getenv("SPRAL_B200_CTILE_BLOCK")) g_ctile_block = atoi(e);
   g_trace_panels = getenv("SPRAL_B200_TRACE_PANELS") != nullptr;
   if (const char* e = getenv("SPRAL_B200_PANEL_V2")) g_panel_v2 = atoi(e) != 0;
   if (const char* e = getenv("SPRAL_B200_PANEL_V2_FRONTS")) g_panel_v2_fronts = std::max(1, atoi(e));
   if (g_panel_v2) configure_panel_kernels();
   if (g_bulk_prio) {
      int least = 0, greatest = 0;
      CUDA_TRY(cudaDeviceGetStreamPriorityRange(&least, &greatest));
      CUDA_TRY(cudaStreamCreateWithPriority(&N.stream, cudaStreamNonBlocking, greatest));
      CUDA_TRY(cudaStreamCreateWithPriority(&N.stream2, cudaStreamNonBlocking, least));
   } else {
      CUDA_TRY(cudaStreamCreateWithFlags(&N.stream, cudaStreamNonBlocking));
      CUDA_TRY(cudaStreamCreateWithFlags(&N.stream2, cudaStreamNonBlocking));
   }
   CUDA_TRY(cudaEventCreateWithFlags(&N.ev_bulk, cudaEventDisableTiming));
   CUDA_TRY(cudaEventCreateWithFlags(&N.ev_bulk_all, cudaEventDisableTiming));
   cudaStream_t s = N.stream;
   configure_update_kernels();
   configure_solve_kernels();
   if (const char* e = getenv("SPRAL_B200_LOOKAHEAD")) g_lookahead = atoi(e) != 0;
   if (const char* e = getenv("SPRAL_B200_SOLVE_GRAPHS")) g_solve_graphs = atoi(e) != 0;
   if (const char* e = getenv("SPRAL_B200_SOLVE_COOP")) g_solve_coop = atoi(e) != 0;
   if (const char* e = getenv("SPRAL_B200_SOLVE_WIDE")) g_solve_wide = atoi(e);
   if (const char* e = getenv("SPRAL_B200_SOLVE_LOOKAHEAD")) g_solve_lookahead = atoi(e) != 0;
   if (const char* e = getenv("SPRAL_B200_SOLVE_LINV")) g_solve_linv = atoi(e) != 0;
   if (const char* e = getenv("SPRAL_B200_SOLVE_WIDE_MIN")) g_solve_wide_min = std::max(1, atoi(e));
   g_bulk_ctas = g_bulk_prio ? -1 : device_sm_count() - 28;
   if (const char* e = getenv("SPRAL_B200_BULK_CTAS")) g_bulk_ctas = atoi(e);
   auto t_begin = std::chrono::steady_clock::now();
   CUDA_TRY(cudaEventCreate(&N.ev_begin));
   CUDA_TRY(cudaEventCreate(&N.ev_end));
   CUDA_TRY(cudaEventRecord(N.ev_begin, s));

   FactorParams prm{opt->u, opt->small, opt->action ? 1 : 0};
   if (nloc == 0) return;
#ifdef SPRAL_B200_SPLIT
   /* the segment exists for the whole part, so that the helper -- which arrives when its own parts are done, before
    * this part can end -- always finds it and always sees its end (phase 4 in ~SplitOwner) */
   if (!S.split_shm.empty()) {
      N.split = SplitOwner::create(S.split_shm.c_str(), S.split_helpers);
      if (!N.stream3) CUDA_TRY(cudaStreamCreateWithFlags(&N.stream3, cudaStreamNonBlocking));
   }
#endif
   Prof prof;
   struct ProfScope { ProfScope(Prof* p) { g_prof = g_profile ? p : nullptr; } ~ProfScope() { g_prof = nullptr; } } prof_scope(&prof);

   /* aval / scaling on the device (the H2D copy of A is part of the factor time,
    * as in gpu/subtree.f90:375-379) */
   const double* d_aval = aval_in;
   const double* d_scal = scaling_in;
   if (!is_device_pointer(aval_in, nullptr)) {
      S.b_aval.ensure((size_t)std::max<int64_t>(S.aval_len, 1) * sizeof(double), s);
      CUDA_TRY(cudaMemcpyAsync(S.b_aval.p, aval_in, (size_t)S.aval_len * sizeof(double), cudaMemcpyHostToDevice, s));
      d_aval = (const double*)S.b_aval.p;
   }
   if (scaling_in && !is_device_pointer(scaling_in, nullptr)) {
      S.b_scal.ensure((size_t)S.n * sizeof(double), s);
      CUDA_TRY(cudaMemcpyAsync(S.b_scal.p, scaling_in, (size_t)S.n * sizeof(double), cudaMemcpyHostToDevice, s));
      d_scal = (const double*)S.b_scal.p;
   }

   /* fixed-size scratch */
   S.b_cbuf[0].ensure(std::max<size_t>(S.cbuf_bytes[0], 256), s);
   S.b_cbuf[1].ensure(std::max<size_t>(S.cbuf_bytes[1], 256), s);
   N.d_fronts = (Front*)g_pool.alloc(nloc * sizeof(Front));
   N.h_fronts.assign(nloc, Front());
   std::vector<Front>& F = N.h_fronts;

   /* factor storage estimate: sum over nodes of L + D + perm, times multiplier */
   {
      double tot = 0;
      for (int i = 0; i < nloc; ++i)
         tot += (double)align_up((size_t)S.m0[i], 2) * S.n0[i] * 8 + 16.0 * S.n0[i] + 4.0 * S.n0[i] + 768;
      double mult = opt->multiplier > 1.0 ? opt->multiplier : 1.0;
      N.reserve((size_t)(tot * mult) + (1 << 20));
   }

   /* external contributions: bring them to this device, compute their maps */
   struct Ext { AsmSrc src; int node; };
   std::vector<Ext> ext(S.contrib_dest.size());
   for (size_t k = 0; k < ext.size(); ++k) {
      auto* c = static_cast<spral_ssids_b200_contrib*>(child_contrib[k]);
      while (!c->ready) { /* the producer part sets ready last (fkeep.F90:163-169) */ }
      int node = S.contrib_dest[k];
      Ext& e = ext[k];
      e.node = node;
      std::memset(&e.src, 0, sizeof(AsmSrc));
      int cn = c->n, nd = c->ndelay;
      /* row list -> positions in the destination node's row list */
      std::vector<int> crl(cn);
      if (cn) CUDA_TRY(cudaMemcpy(crl.data(), c->rlist, cn * sizeof(int), cudaMemcpyDefault));
      std::vector<int> map(cn);
      {
         std::vector<int> pos(S.n + 1, 0);
         for (int64_t ii = S.rptr[node]; ii < S.rptr[node + 1]; ++ii) pos[S.rlist[ii]] = (int)(ii - S.rptr[node] + 1);
         int npl = 0;
         for (int i = 0; i < cn; ++i) { map[i] = pos[crl[i]]; if (map[i] <= S.n0[node]) npl++; }
         e.src.npassl = npl;
      }
      int* d_map = (int*)g_pool.alloc(std::max(cn, 1) * sizeof(int));
      N.ext_allocs.push_back(d_map);
      if (cn) CUDA_TRY(cudaMemcpyAsync(d_map, map.data(), cn * sizeof(int), cudaMemcpyHostToDevice, s));
      CUDA_TRY(cudaStreamSynchronize(s));   // map is a local vector
      e.src.map = d_map; e.src.cm = cn; e.src.ndelay = nd;
      bool local = (c->device == S.device);
      if (!local && c->device >= 0) {      // pull over NVLink: direct P2P when the topology allows it
         int can = 0;
         if (cudaDeviceCanAccessPeer(&can, S.device, c->device) == cudaSuccess && can) {
            cudaError_t pe = cudaDeviceEnablePeerAccess(c->device, 0);
            if (pe != cudaSuccess) cudaGetLastError();   // already enabled
         }
      }
      if (c->val && cn) {
         if (local) { e.src.C = c->val; e.src.ldc = c->ldval; }
         else {
            double* d = (double*)g_pool.alloc((size_t)cn * cn * sizeof(double));
            N.ext_allocs.push_back(d);
            CUDA_TRY(cudaMemcpy2DAsync(d, (size_t)cn * sizeof(double), c->val, (size_t)c->ldval * sizeof(double),
                                       (size_t)cn * sizeof(double), cn, cudaMemcpyDefault, s));
            e.src.C = d; e.src.ldc = cn;
         }
      }
      if (nd > 0) {
         if (local) { e.src.dval = c->delay_val; e.src.lddelay = c->lddelay; e.src.dperm = c->delay_perm; }
         else {
            size_t rows = (size_t)nd + cn;
            double* d = (double*)g_pool.alloc(rows * nd * sizeof(double));
            int* dp = (int*)g_pool.alloc(nd * sizeof(int));
            N.ext_allocs.push_back(d); N.ext_allocs.push_back(dp);
            CUDA_TRY(cudaMemcpy2DAsync(d, rows * sizeof(double), c->delay_val, (size_t)c->lddelay * sizeof(double),
                                       rows * sizeof(double), nd, cudaMemcpyDefault, s));
            CUDA_TRY(cudaMemcpyAsync(dp, c->delay_perm, nd * sizeof(int), cudaMemcpyDefault, s));
            e.src.dval = d; e.src.lddelay = (int)rows; e.src.dperm = dp;
         }
      }
   }

   std::vector<int> nelim(nloc, 0), ncol(nloc, 0);   // by node
   spral_ssids_b200_stats st;
   std::memset(&st, 0, sizeof(st));
   const int ASMC = assemble_cols_per_cta();
   const int SCH = scatter_chunk();
   const int MAXRANK = 8;
   double t_sync = 0;

   const bool trace_levels = getenv("SPRAL_B200_TRACE") != nullptr;
   for (int lev = 0; lev < S.nlevels; ++lev) {
      const int f0 = S.level_ptr[lev] - 1, f1 = S.level_ptr[lev + 1] - 1;
      const int nfl = f1 - f0;
      if (nfl == 0) continue;
      auto tl0 = std::chrono::steady_clock::now();

      /* ---- geometry ---- */
      size_t lbytes = 0, ldd = 0, bkd = 0;
      int maxm = 0;
      for (int fi = f0; fi < f1; ++fi) {
         int node = S.node_of_front[fi];
         int ndin = 0;
         for (int k = S.child_ptr[node]; k < S.child_ptr[node + 1]; ++k) {
            int c = S.child_list[k];
            ndin += ncol[c] - nelim[c];
         }
         for (int k : S.contribs_of_node[node]) ndin += ext[k].src.ndelay;
         Front& f = F[fi];
         std::memset(&f, 0, sizeof(Front));
         f.m0 = S.m0[node]; f.n0 = S.n0[node]; f.ndin = ndin;
         f.m = f.m0 + ndin; f.n = f.n0 + ndin;
         f.ldl = (int)align_up((size_t)f.m, 2);
         f.end = f.n; f.first_pass_done = -1;
         ncol[node] = f.n;
         maxm = std::max(maxm, f.m);
         lbytes += align_up((size_t)f.ldl * f.n * sizeof(double), 256);
         if (!posdef) lbytes += align_up((size_t)2 * f.n * sizeof(double), 256);
         lbytes += align_up((size_t)f.n * sizeof(int), 256);
         ldd += (size_t)f.ldl * f.n + 32;
         bkd += (size_t)f.ldl * BS + 32;
      }
      const bool big = maxm >= 192;
      /* speculative panel segments: a few large fronts per level only (the chain workspace is 260 KB per front) */
      const bool v2 = g_panel_v2 && big && nfl <= g_panel_v2_fronts;
      const int bkw = v2 ? panel_segment_width() : BS;        // columns of the backup scratch
      if (v2) {
         bkd = 0;
         for (int fi = f0; fi < f1; ++fi) bkd += (size_t)F[fi].ldl * bkw + 32;      // (unused when posdef)
         S.b_segws.ensure((size_t)nfl * panel_segment_ws_bytes(), s);
      }
      char* lblock = (char*)N.falloc(lbytes);
      if (!posdef) {
         S.b_ld.ensure(std::max(ldd, S.ld_estimate) * sizeof(double), s);
         S.b_bk.ensure(bkd * sizeof(double), s);
         S.b_ws.ensure((size_t)nfl * sizeof(BlockWS), s);
      } else {
         S.b_ws.ensure((size_t)nfl * sizeof(BlockWS), s);
      }
      {
         size_t off = 0, ldo = 0, bko = 0, co = 0;
         for (int fi = f0; fi < f1; ++fi) {
            int node = S.node_of_front[fi];
            Front& f = F[fi];
            f.L = (double*)(lblock + off); off += align_up((size_t)f.ldl * f.n * sizeof(double), 256);
            if (!posdef) { f.D = (double*)(lblock + off); off += align_up((size_t)2 * f.n * sizeof(double), 256); }
            f.perm = (int*)(lblock + off); off += align_up((size_t)f.n * sizeof(int), 256);
            if (!posdef) {
               f.LD = (double*)S.b_ld.p + ldo; ldo += align_up((size_t)f.ldl * f.n, 32);
               f.BK = (double*)S.b_bk.p + bko; bko += align_up((size_t)f.ldl * bkw, 32);
            } else { f.LD = f.L; f.BK = nullptr; }
            f.ws = (BlockWS*)S.b_ws.p + (fi - f0);
            f.sws = v2 ? (SegWS*)((char*)S.b_segws.p + (size_t)(fi - f0) * panel_segment_ws_bytes()) : nullptr;
            f.rows = S.d_rlist + S.rptr[node];
            int cm = f.m0 - f.n0;
            f.ldc = (int)align_up((size_t)cm, 2);
            if (cm > 0) {
               if (S.exported[node]) {
                  N.d_export = (double*)g_pool.alloc((size_t)f.ldc * cm * sizeof(double));
                  f.C = N.d_export; N.export_front = fi;
               } else {
                  f.C = (double*)((char*)S.b_cbuf[lev & 1].p + co);
                  co += align_up((size_t)f.ldc * cm * sizeof(double), 256);
               }
            }
         }
      }
      CUDA_TRY(cudaMemcpyAsync(N.d_fronts + f0, &F[f0], nfl * sizeof(Front), cudaMemcpyHostToDevice, s));
      PROF(PC_INIT, CUDA_TRY(cudaMemsetAsync(lblock, 0, lbytes, s)));

      /* ---- per-level work lists ---- */
      std::vector<int2> scat;
      std::vector<AsmSrc> srcs;
      std::vector<std::vector<int2>> pre(MAXRANK + 1), post(MAXRANK + 1);
      std::vector<int2> dly;
      std::vector<int> lfronts(nfl), lrem(nfl);
      for (int fi = f0; fi < f1; ++fi) {
         int node = S.node_of_front[fi];
         const Front& f = F[fi];
         lfronts[fi - f0] = fi; lrem[fi - f0] = f.n;
         int64_t nent = S.nptr[node + 1] - S.nptr[node];
         int nch = (int)std::max<int64_t>(1, (nent + SCH - 1) / SCH);
         for (int c = 0; c < nch; ++c) scat.push_back(make_int2(fi, c));
         int rank = 0, delay_col = f.n0;
         auto add_src = [&](const AsmSrc& a) {
            int si = (int)srcs.size();
            srcs.push_back(a);
            AsmSrc& q = srcs.back();
            q.parent = fi; q.delay_col = delay_col;
            delay_col += q.ndelay;
            int r = std::min(rank, MAXRANK);
            if (q.C) {
               int t_split = q.npassl / ASMC;   // tile containing the first contribution column
               int ntile = (q.cm + ASMC - 1) / ASMC;
               for (int t = 0; t < (q.npassl + ASMC - 1) / ASMC; ++t) pre[r].push_back(make_int2(si, t));
               for (int t = t_split; t < ntile; ++t) post[r].push_back(make_int2(si, t));
            }
            for (int j = 0; j < q.ndelay; ++j) dly.push_back(make_int2(si, j));
            rank++;
         };
         for (int k = S.child_ptr[node]; k < S.child_ptr[node + 1]; ++k) {
            int c = S.child_list[k];
            const Front& cf = F[S.front_of_node[c]];
            AsmSrc a;
            std::memset(&a, 0, sizeof(a));
            a.cm = cf.m0 - cf.n0; a.C = cf.C; a.ldc = cf.ldc;
            a.map = S.d_rlist_direct + S.rptr[c] + cf.n0;
            a.ndelay = cf.n - cf.nelim;
            a.dval = cf.L + (size_t)cf.nelim * cf.ldl + cf.nelim; a.lddelay = cf.ldl;
            a.dperm = cf.perm + cf.nelim;
            a.npassl = S.npassl[c];
            add_src(a);
         }
         for (int k : S.contribs_of_node[node]) add_src(ext[k].src);
      }
      std::vector<MatTile> ctiles;
      {
         const int T = update_tile_size(big);
         for (int fi = f0; fi < f1; ++fi) {
            const Front& f = F[fi];
            if (f.m == f.n) continue;
            int mt = (f.m + T - 1) / T;
            const int tj0 = f.n / T;
            const int SB = g_ctile_block;
            if (big && SB > 1 && (mt - tj0) >= 2 * SB) {
               /* blocked order: consecutive groups of ~SB*SB tiles form SB x SB squares, so the CTAs of
                * one wave of the persistent kernel share SB row panels and SB column panels of L / LD in
                * the L2 instead of one column panel and 148 different row panels */
               for (int bj = tj0; bj < mt; bj += SB)
                  for (int bi = bj; bi < mt; bi += SB)
                     for (int tj = bj; tj < std::min(bj + SB, mt); ++tj)
                        for (int ti = std::max(bi, tj); ti < std::min(bi + SB, mt); ++ti) ctiles.push_back({fi, ti, tj});
            } else {
               for (int tj = tj0; tj < mt; ++tj)
                  for (int ti = tj; ti < mt; ++ti) ctiles.push_back({fi, ti, tj});
            }
         }
      }
      size_t wbytes = 4096 + (scat.size() + dly.size()) * sizeof(int2) + srcs.size() * sizeof(AsmSrc)
                    + (lfronts.size() + 64) * sizeof(int) + ctiles.size() * sizeof(MatTile);
      for (auto& v : pre) wbytes += v.size() * sizeof(int2) + 256;
      for (auto& v : post) wbytes += v.size() * sizeof(int2) + 256;
      wbytes += 256 * 32;
      S.b_work.ensure(wbytes + (1 << 16), s);
      Bump bump; bump.reset(S.b_work);

      auto tl1 = std::chrono::steady_clock::now();
      /* ---- init + assemble (fully-summed part) ---- */
      int2* d_scat = upload(bump, scat, s);
      PROF(PC_INIT, launch_scatter_a(N.d_fronts, d_scat, (int)scat.size(), S.d_nlist, S.d_nptr, S.d_node_of_front,
                       d_aval, d_scal, s));
      AsmSrc* d_srcs = upload(bump, srcs, s);
      std::vector<int2*> d_post(MAXRANK + 1, nullptr);
      for (int r = 0; r <= MAXRANK; ++r) {
         if (!pre[r].empty()) {
            int2* d = upload(bump, pre[r], s);
            PROF(PC_ASSEMBLE, launch_assemble(N.d_fronts, d_srcs, d, (int)pre[r].size(), false, r == MAXRANK, s));
         }
         if (!post[r].empty()) d_post[r] = upload(bump, post[r], s);
      }
      if (!dly.empty()) {
         int2* d = upload(bump, dly, s);
         launch_delays(N.d_fronts, d_srcs, d, (int)dly.size(), s);
      }
      MatTile* d_ctiles = upload(bump, ctiles, s);

      auto tl2 = std::chrono::steady_clock::now();
      /* ---- factorise the fully-summed columns (panel by panel, one sync per panel) ---- */
      {
#ifdef SPRAL_B200_SPLIT
         /* a single large front at the top of the root part: its far columns go to the helper GPU (split_front.h) */
         if (N.split) N.split->level_ok = (nfl == 1 && big && g_lookahead);
         if (N.split && N.split->level_ok) N.split->begin_front(F[f0], posdef, s);
#endif
         int err = factor_fronts(N, N.d_fronts, F, lfronts, big, prm, S.b_retry, t_sync);
#ifdef SPRAL_B200_SPLIT
         if (!err && N.split && N.split->active) throw std::runtime_error("split front: still active when the front is finished");
         if (N.split) N.split->level_ok = false;
#endif
         if (err) { st.flag = err; *stats = st; return; }
         int* d_lf = upload(bump, lfronts, s);
         launch_finalize(N.d_fronts, d_lf, nfl, posdef, s);
         CUDA_TRY(cudaMemcpyAsync(&F[f0], N.d_fronts + f0, nfl * sizeof(Front), cudaMemcpyDeviceToHost, s));
         auto ts0 = std::chrono::steady_clock::now();
         CUDA_TRY(cudaStreamSynchronize(s));
         t_sync += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - ts0).count();
         prof.collect();
         for (int fi = f0; fi < f1; ++fi) {
            if (F[fi].flag < 0) { st.flag = F[fi].flag; *stats = st; return; }
            if (!F[fi].finished) throw std::runtime_error("front not finished after the panel loop");
         }
      }

      /* ---- Schur complement, then the children's contributions to it ---- */
      if (g_profile && !ctiles.empty()) {
         cudaEvent_t a, b;
         CUDA_TRY(cudaEventCreate(&a)); CUDA_TRY(cudaEventCreate(&b));
         CUDA_TRY(cudaEventRecord(a, s));
         nvtxRangePushA("upd_contrib");
         PROF(PC_CONTRIB, launch_update(N.d_fronts, d_ctiles, (int)ctiles.size(), UPD_CONTRIB, big, s));
         nvtxRangePop();
         CUDA_TRY(cudaEventRecord(b, s));
         N.prof_events.push_back({a, b});
         for (int fi = f0; fi < f1; ++fi) {
            double cm = F[fi].m - F[fi].n;
            N.prof_flops += cm * (cm + 1) * F[fi].nelim;   // lower triangle, 2 flops per fma
         }
      } else
      launch_update(N.d_fronts, d_ctiles, (int)ctiles.size(), UPD_CONTRIB, big, s);
      for (int r = 0; r <= MAXRANK; ++r)
         if (d_post[r]) PROF(PC_ASSEMBLE, launch_assemble(N.d_fronts, d_srcs, d_post[r], (int)post[r].size(), true, r == MAXRANK, s));
      CUDA_TRY(cudaGetLastError());

      if (trace_levels) {
         auto tl3 = std::chrono::steady_clock::now();
         auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
         fprintf(stderr, "[level %d] fronts %d maxm %d: host prep %.2f ms, init+assemble launches %.2f ms, factor %.2f ms\n",
                 lev, nfl, maxm, ms(tl0, tl1), ms(tl1, tl2), ms(tl2, tl3));
      }
      /* ---- statistics (cpu/factor.hxx:117-124, NumericSubtree.hxx:248-280) ---- */
      for (int fi = f0; fi < f1; ++fi) {
         const Front& f = F[fi];
         int node = S.node_of_front[fi];
         nelim[node] = f.nelim;
         st.num_delay += f.n - f.nelim;
         for (int64_t j = f.m; j >= (int64_t)f.m - f.nelim + 1; --j) { st.num_factor += j; st.num_flops += j * j; }
         st.maxfront = std::max(st.maxfront, f.m);
         st.maxsupernode = std::max(st.maxsupernode, f.n);
         if (!posdef) {
            st.num_neg += f.num_neg; st.num_two += f.num_two; st.num_zero += f.num_zero;
            int fp = f.first_pass_done < 0 ? f.nelim : f.first_pass_done;
            st.not_first_pass += f.n - fp;
            st.not_second_pass += f.n - f.nelim;
         }
      }
      /* the next level re-uses b_work / b_ld / ...: its uploads and kernels are
       * ordered behind this level's kernels on the same stream */
   }
   if (st.num_zero > 0) st.flag = SPRAL_SSIDS_WARNING_FACT_SINGULAR;

   /* ---- solve data ---- */
   {
      std::vector<SolveFront> sf(nloc);
      std::vector<RowTile> sw;
      std::vector<int> wbeg(nloc, 0);
      N.swork_ptr.assign(S.nlevels + 1, 0);
      N.lvl_steps.assign(S.nlevels, 0);
      const int SBk = solve_block();
      for (int lev = 0; lev < S.nlevels; ++lev) {
         int f0 = S.level_ptr[lev] - 1, f1 = S.level_ptr[lev + 1] - 1;
         N.swork_ptr[lev] = (int)sw.size();
         int mx = 0;
         for (int fi = f0; fi < f1; ++fi) {
            const Front& f = F[fi];
            sf[fi] = SolveFront{f.L, f.D, f.perm, f.rows, f.ldl, f.m, f.n, f.n0, f.m0, f.nelim};
            wbeg[fi] = (int)sw.size() - N.swork_ptr[lev];
            int nt = (f.m + RT - 1) / RT;
            for (int t = 0; t < nt; ++t) sw.push_back({fi, t});
            mx = std::max(mx, f.nelim);
         }
         N.lvl_steps[lev] = (mx + SBk - 1) / SBk;
         N.max_level_work = std::max(N.max_level_work, sw.size() - (size_t)N.swork_ptr[lev]);
      }
      N.swork_ptr[S.nlevels] = (int)sw.size();
      /* inverse diagonal blocks for the fronts whose T kernels are the critical path of a sweep (solve_types.h) */
      std::vector<int2> lwork;
      if (g_solve_linv) {
         size_t nb_total = 0;
         for (int fi = 0; fi < nloc; ++fi) if (sf[fi].nelim >= SOLVE_LINV_MIN_COLS) nb_total += (size_t)(sf[fi].nelim + 31) / 32;
         if (nb_total > 0) {
            N.d_linv = (double*)g_pool.alloc(nb_total * 1024 * sizeof(double));
            N.d_linv_bad = (int*)g_pool.alloc(nloc * sizeof(int));
            CUDA_TRY(cudaMemsetAsync(N.d_linv_bad, 0, nloc * sizeof(int), s));
            size_t off = 0;
            for (int fi = 0; fi < nloc; ++fi) {
               if (sf[fi].nelim < SOLVE_LINV_MIN_COLS) continue;
               const int nb = (sf[fi].nelim + 31) / 32;
               sf[fi].Linv = N.d_linv + off * 1024;
               sf[fi].linv_bad = N.d_linv_bad + fi;
               for (int b = 0; b < nb; ++b) lwork.push_back(make_int2(fi, b));
               off += (size_t)nb;
            }
         }
      }
      N.d_sfronts = (SolveFront*)g_pool.alloc(nloc * sizeof(SolveFront));
      N.d_swork = (RowTile*)g_pool.alloc(std::max<size_t>(sw.size(), 1) * sizeof(RowTile));
      N.d_wbeg = (int*)g_pool.alloc(nloc * sizeof(int));
      CUDA_TRY(cudaMemcpyAsync(N.d_sfronts, sf.data(), nloc * sizeof(SolveFront), cudaMemcpyHostToDevice, s));
      if (!sw.empty()) CUDA_TRY(cudaMemcpyAsync(N.d_swork, sw.data(), sw.size() * sizeof(RowTile), cudaMemcpyHostToDevice, s));
      CUDA_TRY(cudaMemcpyAsync(N.d_wbeg, wbeg.data(), nloc * sizeof(int), cudaMemcpyHostToDevice, s));
      int2* d_lwork = nullptr;
      if (!lwork.empty()) {
         d_lwork = (int2*)g_pool.alloc(lwork.size() * sizeof(int2));
         CUDA_TRY(cudaMemcpyAsync(d_lwork, lwork.data(), lwork.size() * sizeof(int2), cudaMemcpyHostToDevice, s));
         launch_build_linv(N.d_sfronts, d_lwork, (int)lwork.size(), posdef, s);
      }
      CUDA_TRY(cudaStreamSynchronize(s));
      if (d_lwork) g_pool.release(d_lwork);
   }
#ifdef SPRAL_B200_SPLIT
   delete N.split; N.split = nullptr;          // phase 4: the helper leaves its service loop
#endif
   CUDA_TRY(cudaEventRecord(N.ev_end, s));
   CUDA_TRY(cudaEventSynchronize(N.ev_end));
   float ems = 0;
   CUDA_TRY(cudaEventElapsedTime(&ems, N.ev_begin, N.ev_end));
   N.timings[0] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
   N.timings[1] = ems;                 // device time of the whole factorisation (CUDA events on its stream)
   N.timings[5] = t_sync;              // host time spent waiting in the per-level syncs
   N.timings[6] = (double)g_launches.exchange(0);   // kernels launched by this factorisation
   if (g_profile) {
      double tot = 0;
      for (auto& e : N.prof_events) { float t = 0; cudaEventElapsedTime(&t, e.first, e.second); tot += t; }
      N.timings[2] = tot;              // ms inside the Schur-complement (UPD_CONTRIB) launches
      N.timings[3] = N.prof_flops;     // their algorithmic flops
      N.timings[4] = (double)N.prof_events.size();
      prof.collect();
      for (int i = 0; i < PC_COUNT; ++i) N.class_ms[i] = prof.ms[i];
   }
   *stats = st;
}

/* ------------------------------------------------------------------------ */
/* Solves                                                                    */
/* ------------------------------------------------------------------------ */

enum SolveJob { JOB_FWD, JOB_DIAG, JOB_DIAG_BWD, JOB_BWD };

static int solve_subtree(const Numeric& Nc, SolveJob job, int nrhs, double* x, int ldx) {
   Numeric& N = const_cast<Numeric&>(Nc);
   Symbolic& S = *N.S;
   if (S.nloc == 0 || nrhs == 0) return 0;
   try {
      std::lock_guard<std::mutex> lock(S.mtx);
      const auto t_solve0 = std::chrono::steady_clock::now();
      CUDA_TRY(cudaSetDevice(S.device));
      cudaStream_t s = N.stream;
      const bool posdef = N.posdef;
      double* dx = x;
      bool host_x = !is_device_pointer(x, nullptr);
      size_t xbytes = (size_t)ldx * nrhs * sizeof(double);
      if (host_x) {
         S.b_x.ensure(xbytes, s);
         dx = (double*)S.b_x.p;
         CUDA_TRY(cudaMemcpyAsync(dx, x, xbytes, cudaMemcpyHostToDevice, s));
      }
      const int maxnr = std::min(std::max(nrhs, 1), solve_max_chunk());
      const size_t chunk_bytes = (size_t)S.n * maxnr * sizeof(double);
      if (job == JOB_FWD) S.b_y.ensure(chunk_bytes, s);
      /* levels swept 256 columns at a time (solve_wide.h) */
      auto wide_level = [&](int lev, int nr) {
         /* with 16+ right-hand sides the wide kernels (FP64 tensor cores) win on every level; with few, the levels of
          * small fronts are better off with the 32-column kernels (measured on cfg5: 15.3 vs 16.6 ms, 84 vs 93 ms) */
         if (!g_solve_wide || N.lvl_steps[lev] < (nr >= 16 ? 1 : g_solve_wide_min)) return false;
         /* the backward sweep keeps one 256 x nr accumulator per front of the level */
         size_t nfr = (size_t)(S.level_ptr[lev + 1] - S.level_ptr[lev]);
         return (nfr <= SOLVE_LOOKAHEAD_MAX_FRONTS ? 2 : 1) * nfr * solve_wide_block() * nr * sizeof(double) <= ((size_t)1 << 30);
      };
      if (job == JOB_DIAG_BWD || job == JOB_BWD) {
         size_t need = std::max<size_t>(N.max_level_work, 1) * solve_block() * 32 * sizeof(double);      // 32-column kernels: per tile
         for (int lev = 0; lev < S.nlevels; ++lev)
            if (wide_level(lev, maxnr)) {
               const size_t nfr = (size_t)(S.level_ptr[lev + 1] - S.level_ptr[lev]);
               need = std::max(need, (nfr <= SOLVE_LOOKAHEAD_MAX_FRONTS ? 2 : 1) * nfr * solve_wide_block() * maxnr * sizeof(double));
            }
         S.b_pbuf.ensure(need, s);
      }
      S.b_bar.ensure(256, s);
      unsigned int* bar = g_solve_coop ? (unsigned int*)S.b_bar.p : nullptr;
      S.b_xt.ensure(chunk_bytes, s);
      /* 64 at a time when every level can be swept by the wide kernels (the 32-column kernels stop at 32) */
      bool all_wide64 = g_solve_wide != 0;
      for (int lev = 0; lev < S.nlevels && all_wide64; ++lev)
         if (N.swork_ptr[lev + 1] > N.swork_ptr[lev] && !wide_level(lev, 64)) all_wide64 = false;
      /* Two lanes: the sweeps of two chunks of right-hand sides are independent (separate copies of x, y and the
       * accumulators; the factors are read-only), so with two or more full chunks two of them run concurrently on two
       * streams: the one-CTA-per-front T kernels of one lane overlap the G kernels of the other, which a single
       * sweep cannot do (T(b+1) needs G(b)).  Measured on cfg5: 64 right-hand sides as 2 x 32 on two lanes 54.6 ms
       * against 52.3 ms as one chunk of 64, so a lane is never narrower than a full chunk.
       * SPRAL_B200_SOLVE_LANES=1: one lane. */
      static int lanes_env = -1;
      if (lanes_env < 0) { const char* e = getenv("SPRAL_B200_SOLVE_LANES"); lanes_env = e ? std::max(1, std::min(2, atoi(e))) : 2; }
      const int full_chunk = all_wide64 ? 64 : solve_rhs_chunk(nrhs);
      const int nlanes = (lanes_env == 2 && nrhs >= 2 * full_chunk && nrhs >= 64 && !g_solve_graphs && !bar) ? 2 : 1;
      cudaStream_t lane_s[2] = {s, s};
      double* lane_xs[2] = {(double*)S.b_xt.p, nullptr};
      double* lane_y[2] = {(double*)S.b_y.p, nullptr};
      double* lane_p[2] = {(double*)S.b_pbuf.p, nullptr};
      if (nlanes == 2) {
         if (!N.lane2) {
            CUDA_TRY(cudaStreamCreateWithFlags(&N.lane2, cudaStreamNonBlocking));
            CUDA_TRY(cudaEventCreateWithFlags(&N.ev_lane_in, cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&N.ev_lane_out, cudaEventDisableTiming));
         }
         lane_s[1] = N.lane2;
         S.b_xt2.ensure(chunk_bytes, s);
         if (job == JOB_FWD) S.b_y2.ensure(chunk_bytes, s);
         if (job == JOB_DIAG_BWD || job == JOB_BWD) S.b_pbuf2.ensure(S.b_pbuf.cap, s);
         lane_xs[1] = (double*)S.b_xt2.p; lane_y[1] = (double*)S.b_y2.p; lane_p[1] = (double*)S.b_pbuf2.p;
         CUDA_TRY(cudaEventRecord(N.ev_lane_in, s));              // x is on the device, the buffers exist
         CUDA_TRY(cudaStreamWaitEvent(N.lane2, N.ev_lane_in, 0));
      }
      int lane = 0;
      for (int r0 = 0; r0 < nrhs;) {
         int nr = (nrhs - r0 >= 64 && all_wide64) ? 64 : solve_rhs_chunk(nrhs - r0);
         double* xcol = dx + (size_t)r0 * ldx;
         cudaStream_t s = lane_s[lane];                           // (shadows the main stream inside the chunk)
         double* xs = lane_xs[lane];
         double* ywork = lane_y[lane];
         double* pbuf = lane_p[lane];
         SolveAux* aux = nullptr;
         if (g_solve_lookahead) { aux = &N.solve_aux[lane]; aux->create(); }
         launch_transpose_rhs(xcol, ldx, xs, S.n, nr, true, s);
         auto sweep = [&]() {
            if (job == JOB_FWD) {
               for (int lev = 0; lev < S.nlevels; ++lev) {
                  const int nwork = N.swork_ptr[lev + 1] - N.swork_ptr[lev];
                  if (wide_level(lev, nr)) {
                     int f0 = S.level_ptr[lev] - 1, f1 = S.level_ptr[lev + 1] - 1;
                     launch_fwd_level_wide(N.d_sfronts, f0, f1 - f0, N.d_swork + N.swork_ptr[lev], nwork,
                           (N.lvl_steps[lev] + 7) / 8, posdef, nr, xs, ywork, s, f1 - f0 <= SOLVE_LOOKAHEAD_MAX_FRONTS ? aux : nullptr,
                           N.lvl_steps[lev] <= 4);
                  } else
                     launch_fwd_level(N.d_sfronts, N.d_swork + N.swork_ptr[lev], nwork, N.lvl_steps[lev], posdef, nr,
                           xs, ldx, ywork, s, bar);
               }
               launch_fwd_flush(N.d_sfronts, 0, S.nloc, nr, xs, ldx, ywork, s);
            } else if (job == JOB_DIAG) {
               if (!posdef) launch_diag_solve(N.d_sfronts, 0, S.nloc, nr, xs, ldx, s);
            } else {
               for (int lev = S.nlevels - 1; lev >= 0; --lev) {
                  int f0 = S.level_ptr[lev] - 1, f1 = S.level_ptr[lev + 1] - 1;
                  if (job == JOB_DIAG_BWD && !posdef) launch_diag_solve(N.d_sfronts, f0, f1 - f0, nr, xs, ldx, s);
                  if (wide_level(lev, nr))
                     launch_bwd_level_wide(N.d_sfronts, f0, f1 - f0, N.d_swork + N.swork_ptr[lev],
                           N.swork_ptr[lev + 1] - N.swork_ptr[lev], N.d_wbeg, (N.lvl_steps[lev] + 7) / 8, posdef, nr,
                           xs, pbuf, s, f1 - f0 <= SOLVE_LOOKAHEAD_MAX_FRONTS ? aux : nullptr, N.lvl_steps[lev] <= 4);
                  else
                  launch_bwd_level(N.d_sfronts, f0, f1 - f0, N.d_swork + N.swork_ptr[lev],
                        N.swork_ptr[lev + 1] - N.swork_ptr[lev], N.d_wbeg, N.lvl_steps[lev], posdef, nr,
                        xs, ldx, pbuf, s, bar);
               }
            }
         };
         Numeric::SolveGraph* sg = nullptr;
         if (g_solve_graphs) {
            for (auto& c : N.graphs)
               if (c.job == (int)job && c.nr == nr) { sg = &c; break; }
            if (sg && (sg->xs != xs || sg->ywork != ywork || sg->pbuf != pbuf)) {   // a pool buffer moved
               cudaGraphExecDestroy(sg->exec); sg->exec = nullptr;
            }
            if (!sg) { N.graphs.push_back(Numeric::SolveGraph{(int)job, nr, nullptr, nullptr, nullptr, nullptr}); sg = &N.graphs.back(); }
            if (!sg->exec) {
               cudaGraph_t graph = nullptr;
               CUDA_TRY(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
               sweep();
               CUDA_TRY(cudaStreamEndCapture(s, &graph));
               CUDA_TRY(cudaGraphInstantiate(&sg->exec, graph, 0));
               cudaGraphDestroy(graph);
               sg->xs = xs; sg->ywork = ywork; sg->pbuf = pbuf;
            }
            CUDA_TRY(cudaGraphLaunch(sg->exec, s));
         } else {
            sweep();
         }
         launch_transpose_rhs(xcol, ldx, xs, S.n, nr, false, s);
         r0 += nr;
         lane = (lane + 1) % nlanes;
      }
      if (nlanes == 2) {                                          // join
         CUDA_TRY(cudaEventRecord(N.ev_lane_out, N.lane2));
         CUDA_TRY(cudaStreamWaitEvent(s, N.ev_lane_out, 0));
      }
      CUDA_TRY(cudaGetLastError());
      if (host_x) CUDA_TRY(cudaMemcpyAsync(x, dx, xbytes, cudaMemcpyDeviceToHost, s));
      const auto t_enq = std::chrono::steady_clock::now();
      CUDA_TRY(cudaStreamSynchronize(s));
      if (getenv("SPRAL_B200_TRACE")) {
         auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
         fprintf(stderr, "[solve] job %d nrhs %d: %.2f ms to enqueue, %.2f ms more until the stream drained\n", (int)job, nrhs,
                 ms(t_solve0, t_enq), ms(t_enq, std::chrono::steady_clock::now()));
      }
   } catch (const CudaError& e) {
      fprintf(stderr, "spral_ssids_b200: CUDA error %d (%s) in solve\n", (int)e.code, cudaGetErrorString(e.code));
      return SPRAL_SSIDS_ERROR_CUDA_UNKNOWN;
   } catch (const std::bad_alloc&) {
      return SPRAL_SSIDS_ERROR_ALLOCATION;
   }
   return 0;
}

} // namespace b200

/* ------------------------------------------------------------------------ */
/* C ABI                                                                     */
/* ------------------------------------------------------------------------ */

using namespace b200;

extern "C" {

int spral_ssids_b200_cuda_init(int* cnt) {
   ABI_GUARD();
   int n = 0;
   cudaError_t e = cudaGetDeviceCount(&n);
   if (e != cudaSuccess) { *cnt = 0; cudaGetLastError(); return (int)e; }
   for (int d = 0; d < n; ++d) {
      if ((e = cudaSetDevice(d)) != cudaSuccess) { *cnt = 0; return (int)e; }
      void* p = nullptr;
      if ((e = cudaMalloc(&p, 1 << 20)) != cudaSuccess) { *cnt = 0; return (int)e; }
      cudaFree(p);
   }
   *cnt = n;
   return 0;
}

void* spral_ssids_gpu_create_symbolic_subtree(
      int device, int n, int sa, int en, const int* sptr, const int* sparent,
      const int64_t* rptr, const int* rlist, const int64_t* nptr,
      const int64_t* nlist, int ncontrib, const int* contrib_idx,
      const struct spral_ssids_b200_options* options) {
   ABI_GUARD();
   try {
      return build_symbolic(device, n, sa, en, sptr, sparent, rptr, rlist, nptr, nlist,
                            ncontrib, contrib_idx, options);
   } catch (const CudaError& e) {
      fprintf(stderr, "spral_ssids_b200: CUDA error %d (%s) in symbolic constructor\n",
              (int)e.code, cudaGetErrorString(e.code));
   } catch (const std::exception& e) {
      fprintf(stderr, "spral_ssids_b200: %s in symbolic constructor\n", e.what());
   }
   return nullptr;
}

void spral_ssids_gpu_destroy_symbolic_subtree(void* p) {
   ABI_GUARD(); delete static_cast<Symbolic*>(p); }

void* spral_ssids_gpu_create_num_subtree_dbl(
      bool posdef, const void* symbolic_subtree, const double* aval,
      const double* scaling, void** child_contrib,
      const struct spral_ssids_b200_options* options,
      struct spral_ssids_b200_stats* stats) {
   ABI_GUARD();
   auto* N = new Numeric;
   N->S = const_cast<Symbolic*>(static_cast<const Symbolic*>(symbolic_subtree));
   N->device = N->S->device;
   N->posdef = posdef;
   std::memset(stats, 0, sizeof(*stats));
   try {
      factor_subtree(*N, aval, scaling, child_contrib, options, stats);
   } catch (const CudaError& e) {
      stats->flag = SPRAL_SSIDS_ERROR_CUDA_UNKNOWN;
      stats->cuda_error = (int)e.code;
      fprintf(stderr, "spral_ssids_b200: CUDA error %d (%s) in factor\n", (int)e.code, cudaGetErrorString(e.code));
   } catch (const std::bad_alloc&) {
      stats->flag = SPRAL_SSIDS_ERROR_ALLOCATION;
   } catch (const std::exception& e) {
      stats->flag = SPRAL_SSIDS_ERROR_ALLOCATION;
      fprintf(stderr, "spral_ssids_b200: %s in factor\n", e.what());
   }
   return N;
}

void spral_ssids_gpu_destroy_num_subtree_dbl(bool, void* p) {
   ABI_GUARD(); delete static_cast<Numeric*>(p); }

int spral_ssids_gpu_subtree_solve_fwd_dbl(bool, const void* p, int nrhs, double* x, int ldx) {
   ABI_GUARD();
   return solve_subtree(*static_cast<const Numeric*>(p), JOB_FWD, nrhs, x, ldx);
}
int spral_ssids_gpu_subtree_solve_diag_dbl(bool, const void* p, int nrhs, double* x, int ldx) {
   ABI_GUARD();
   return solve_subtree(*static_cast<const Numeric*>(p), JOB_DIAG, nrhs, x, ldx);
}
int spral_ssids_gpu_subtree_solve_diag_bwd_dbl(bool, const void* p, int nrhs, double* x, int ldx) {
   ABI_GUARD();
   return solve_subtree(*static_cast<const Numeric*>(p), JOB_DIAG_BWD, nrhs, x, ldx);
}
int spral_ssids_gpu_subtree_solve_bwd_dbl(bool, const void* p, int nrhs, double* x, int ldx) {
   ABI_GUARD();
   return solve_subtree(*static_cast<const Numeric*>(p), JOB_BWD, nrhs, x, ldx);
}

/* Format of NumericSubtree::enquire (src/ssids/cpu/NumericSubtree.hxx:424-470):
 * nodes in node order; posdef: diagonal of L; indefinite: piv_order indexed by
 * pivot-order variable, 2x2 pivots negative; d two entries per column. */
void spral_ssids_gpu_subtree_enquire_dbl(bool posdef, const void* p, int* piv_order, double* d) {
   ABI_GUARD();
   const Numeric& N = *static_cast<const Numeric*>(p);
   const Symbolic& S = *N.S;
   cudaSetDevice(S.device);
   int piv = 0;
   std::vector<double> buf; std::vector<int> perm;
   for (int node = 0; node < S.nloc; ++node) {
      const Front& f = N.h_fronts[S.front_of_node[node]];
      if (posdef) {
         buf.resize(f.nelim);
         if (f.nelim) cudaMemcpy2D(buf.data(), sizeof(double), f.L, (size_t)(f.ldl + 1) * sizeof(double),
                                   sizeof(double), f.nelim, cudaMemcpyDeviceToHost);
         for (int i = 0; i < f.nelim; ++i) *(d++) = buf[i];
         continue;
      }
      buf.resize(2 * (size_t)f.nelim + 2); perm.resize(f.nelim);
      if (f.nelim) {
         cudaMemcpy(buf.data(), f.D, 2 * (size_t)f.nelim * sizeof(double), cudaMemcpyDeviceToHost);
         cudaMemcpy(perm.data(), f.perm, f.nelim * sizeof(int), cudaMemcpyDeviceToHost);
      }
      for (int i = 0; i < f.nelim;) {
         if (i + 1 == f.nelim || std::isfinite(buf[2 * i + 2])) {
            if (piv_order) piv_order[perm[i] - 1] = (piv++);
            if (d) { *(d++) = buf[2 * i]; *(d++) = 0.0; }
            i += 1;
         } else {
            if (piv_order) { piv_order[perm[i] - 1] = -(piv++); piv_order[perm[i + 1] - 1] = -(piv++); }
            if (d) { *(d++) = buf[2 * i]; *(d++) = buf[2 * i + 1]; *(d++) = buf[2 * i + 3]; *(d++) = 0.0; }
            i += 2;
         }
      }
   }
}

/* NumericSubtree::alter (src/ssids/cpu/NumericSubtree.hxx:473-497) */
void spral_ssids_gpu_subtree_alter_dbl(bool posdef, void* p, const double* d) {
   ABI_GUARD();
   if (posdef) return;
   Numeric& N = *static_cast<Numeric*>(p);
   const Symbolic& S = *N.S;
   cudaSetDevice(S.device);
   std::vector<double> buf;
   for (int node = 0; node < S.nloc; ++node) {
      const Front& f = N.h_fronts[S.front_of_node[node]];
      if (!f.nelim) continue;
      buf.resize(2 * (size_t)f.nelim + 2);
      cudaMemcpy(buf.data(), f.D, 2 * (size_t)f.nelim * sizeof(double), cudaMemcpyDeviceToHost);
      for (int i = 0; i < f.nelim;) {
         if (i + 1 == f.nelim || std::isfinite(buf[2 * i + 2])) { buf[2 * i] = *(d++); d++; i += 1; }
         else { buf[2 * i] = *(d++); buf[2 * i + 1] = *(d++); buf[2 * i + 3] = *(d++); d++; i += 2; }
      }
      cudaMemcpy(f.D, buf.data(), 2 * (size_t)f.nelim * sizeof(double), cudaMemcpyHostToDevice);
   }
}

void spral_ssids_gpu_subtree_get_contrib_device_dbl(bool, void* p,
      int* n, const double** val, int* ldval, const int** rlist, int* ndelay,
      const int** delay_perm, const double** delay_val, int* lddelay, int* device) {
   ABI_GUARD();
   Numeric& N = *static_cast<Numeric*>(p);
   const Symbolic& S = *N.S;
   *device = S.device;
   if (N.export_front < 0) {
      *n = 0; *val = nullptr; *ldval = 0; *rlist = nullptr; *ndelay = 0;
      *delay_perm = nullptr; *delay_val = nullptr; *lddelay = 0;
      return;
   }
   const Front& f = N.h_fronts[N.export_front];
   int node = S.node_of_front[N.export_front];
   *n = f.m0 - f.n0;
   *val = f.C; *ldval = f.ldc;
   *rlist = S.d_rlist + S.rptr[node] + f.n0;
   *ndelay = f.n - f.nelim;
   *lddelay = f.ldl;
   *delay_perm = (*ndelay > 0) ? f.perm + f.nelim : nullptr;
   *delay_val = (*ndelay > 0) ? f.L + (size_t)f.nelim * (f.ldl + 1) : nullptr;
}

void spral_ssids_gpu_subtree_get_contrib_dbl(bool posdef, void* p,
      int* n, const double** val, int* ldval, const int** rlist, int* ndelay,
      const int** delay_perm, const double** delay_val, int* lddelay) {
   ABI_GUARD();
   Numeric& N = *static_cast<Numeric*>(p);
   const Symbolic& S = *N.S;
   int dev;
   const double *dval, *ddel; const int *drl, *dperm;
   spral_ssids_gpu_subtree_get_contrib_device_dbl(posdef, p, n, &dval, ldval, &drl, ndelay, &dperm, &ddel, lddelay, &dev);
   if (N.export_front < 0) { *val = nullptr; *rlist = nullptr; *delay_perm = nullptr; *delay_val = nullptr; return; }
   cudaSetDevice(S.device);
   const Front& f = N.h_fronts[N.export_front];
   int node = S.node_of_front[N.export_front];
   int cn = *n;
   N.h_cval.resize((size_t)cn * cn);
   cudaMemcpy2D(N.h_cval.data(), (size_t)cn * sizeof(double), dval, (size_t)f.ldc * sizeof(double),
                (size_t)cn * sizeof(double), cn, cudaMemcpyDeviceToHost);
   *val = N.h_cval.data(); *ldval = cn;
   *rlist = S.rlist.data() + S.rptr[node] + f.n0;
   int nd = *ndelay;
   if (nd > 0) {
      size_t rows = (size_t)nd + cn;
      N.h_dval.resize(rows * nd); N.h_dperm.resize(nd);
      cudaMemcpy2D(N.h_dval.data(), rows * sizeof(double), ddel, (size_t)f.ldl * sizeof(double),
                   rows * sizeof(double), nd, cudaMemcpyDeviceToHost);
      cudaMemcpy(N.h_dperm.data(), dperm, nd * sizeof(int), cudaMemcpyDeviceToHost);
      *delay_val = N.h_dval.data(); *lddelay = (int)rows; *delay_perm = N.h_dperm.data();
   } else { *delay_val = nullptr; *delay_perm = nullptr; }
}

void spral_ssids_gpu_subtree_free_contrib_dbl(bool, void* p) {
   ABI_GUARD();
   Numeric& N = *static_cast<Numeric*>(p);
   cudaSetDevice(N.S->device);
   if (N.d_export) { g_pool.release(N.d_export); N.d_export = nullptr; }
   std::vector<double>().swap(N.h_cval);
   std::vector<double>().swap(N.h_dval);
}

void spral_ssids_b200_contrib_fill(struct spral_ssids_b200_contrib* c, bool posdef,
      void* numeric_subtree, bool device_resident) {
   ABI_GUARD();
   std::memset((void*)c, 0, sizeof(*c));
   int dev = -1;
   if (device_resident)
      spral_ssids_gpu_subtree_get_contrib_device_dbl(posdef, numeric_subtree, &c->n, &c->val, &c->ldval,
            &c->rlist, &c->ndelay, &c->delay_perm, &c->delay_val, &c->lddelay, &dev);
   else
      spral_ssids_gpu_subtree_get_contrib_dbl(posdef, numeric_subtree, &c->n, &c->val, &c->ldval,
            &c->rlist, &c->ndelay, &c->delay_perm, &c->delay_val, &c->lddelay);
   c->owner = 1; c->posdef = posdef; c->owner_ptr = numeric_subtree; c->device = dev;
   c->ready = 1;
}

void spral_ssids_gpu_symbolic_get_maps(const void* p, int* rlist_direct, int* num_levels,
      int* level_ptr, int* level_list) {
   ABI_GUARD();
   const Symbolic& S = *static_cast<const Symbolic*>(p);
   cudaSetDevice(S.device);
   if (rlist_direct && !S.rlist_direct.empty())
      cudaMemcpy(rlist_direct, S.d_rlist_direct, S.rlist_direct.size() * sizeof(int), cudaMemcpyDeviceToHost);
   if (num_levels) *num_levels = S.nlevels;
   if (level_ptr) std::copy(S.level_ptr.begin(), S.level_ptr.end(), level_ptr);
   if (level_list) std::copy(S.level_list.begin(), S.level_list.end(), level_list);
}

void spral_ssids_b200_set_profile(int on) {
   ABI_GUARD(); g_profile = (on != 0); }

#ifdef SPRAL_B200_SPLIT
/* Distributed top front (split_front.h; opt-in build).  enable: the owner rank names the shared-memory segment
 * for the part that holds the top of the tree, before it factorises it.  serve: called by the helper rank once
 * its own parts are done; returns when the owner's part is finished (0), when no owner showed up (1) or no
 * front was split (2), or the raw cudaError_t. */
void spral_ssids_b200_split_enable(void* symbolic_subtree, const char* shm_name, int nhelpers) {
   ABI_GUARD();
   static_cast<Symbolic*>(symbolic_subtree)->split_shm = shm_name ? shm_name : "";
   static_cast<Symbolic*>(symbolic_subtree)->split_helpers = std::max(1, nhelpers);
}
int spral_ssids_b200_split_helper_serve(const char* shm_name, int device, double timeout_s, int helper_index) {
   ABI_GUARD();
   try { return split_helper_serve(shm_name, device, timeout_s, helper_index); }
   catch (const CudaError& e) {
      fprintf(stderr, "spral_ssids_b200: CUDA error %d (%s) in the split helper\n", (int)e.code, cudaGetErrorString(e.code));
      return (int)e.code;
   } catch (const std::exception& e) {
      fprintf(stderr, "spral_ssids_b200: %s in the split helper\n", e.what());
      return -1;
   }
}
#endif

void spral_ssids_gpu_subtree_get_timings(const void* p, double* ms, int n) {
   ABI_GUARD();
   const Numeric& N = *static_cast<const Numeric*>(p);
   for (int i = 0; i < n && i < 8; ++i) ms[i] = N.timings[i];
   /* entries 8.. : profiling-mode device ms per kernel class: diag, apply, commit,
    * inner update, swap, outer update, contrib, assemble, init */
   for (int i = 8; i < n && i < 24; ++i) ms[i] = N.class_ms[i - 8];
}

} /* extern "C" */

/* Debug/introspection (not declared in the public header): copies one front's
 * factor data to the host.  sizes = {m, n, ldl, nelim, ndin}. */
extern "C" void spral_ssids_b200_debug_front(const void* p, int node, int* sizes,
      double* L, double* D, int* perm) {
   const Numeric& N = *static_cast<const Numeric*>(p);
   const Symbolic& S = *N.S;
   cudaSetDevice(S.device);
   const Front& f = N.h_fronts[S.front_of_node[node]];
   sizes[0] = f.m; sizes[1] = f.n; sizes[2] = f.ldl; sizes[3] = f.nelim; sizes[4] = f.ndin;
   if (L) cudaMemcpy(L, f.L, (size_t)f.ldl * f.n * sizeof(double), cudaMemcpyDeviceToHost);
   if (D && f.D) cudaMemcpy(D, f.D, (size_t)2 * f.n * sizeof(double), cudaMemcpyDeviceToHost);
   if (perm) cudaMemcpy(perm, f.perm, (size_t)f.n * sizeof(int), cudaMemcpyDeviceToHost);
}

/* ------------------------------------------------------------------------ */
/* Cross-process hand-over of a contribution block (one process per GPU):    */
/* the producer packs [val | delay_val | delay_perm] into one cudaMalloc'd   */
/* block and publishes its CUDA IPC handle; the consumer maps it and pulls   */
/* it over NVLink with a peer copy.  Replaces transfer_contrib's D2H copy    */
/* (src/ssids/gpu/factor.f90:155-221).                                       */
/* ------------------------------------------------------------------------ */
extern "C" {

/* Packs the root contribution of `numeric_subtree` into a fresh device block
 * owned by the numeric object.  Layout (bytes): val n*n doubles (ld n), then
 * delay_val (ndelay+n)*ndelay doubles (ld ndelay+n), then delay_perm ndelay
 * ints.  handle receives the 64-byte cudaIpcMemHandle_t.  Returns 0 or the
 * raw cudaError_t. */
int spral_ssids_gpu_subtree_export_contrib_ipc(void* numeric_subtree, unsigned char* handle,
      int* n, int* ndelay, int64_t* bytes, void** device_block) {
   ABI_GUARD();
   Numeric& N = *static_cast<Numeric*>(numeric_subtree);
   const Symbolic& S = *N.S;
   cudaSetDevice(S.device);
   *n = 0; *ndelay = 0; *bytes = 0; *device_block = nullptr;
   if (N.export_front < 0) return 0;
   const Front& f = N.h_fronts[N.export_front];
   int cn = f.m0 - f.n0, nd = f.n - f.nelim;
   size_t rows = (size_t)nd + cn;
   size_t b_val = (size_t)cn * cn * sizeof(double), b_del = rows * nd * sizeof(double);
   size_t total = b_val + b_del + (size_t)nd * sizeof(int);
   /* the block lives in the symbolic subtree's pool -- two buffers used alternately, so that a slow consumer can
    * still pull the block of the previous factorisation while this one is packed (the host protocol of dist.py
    * waits for the consumer's acknowledgement before a buffer comes round again); the addresses (and IPC handles)
    * repeat, so the consumer maps each buffer once */
   Symbolic& Sm = *N.S;
   Buf& eb = Sm.b_export[Sm.export_slot];
   Sm.export_slot ^= 1;
   cudaError_t e = cudaSuccess;
   try { eb.ensure(std::max<size_t>(total, 256), N.stream); }
   catch (const CudaError& ce) { return (int)ce.code; }
   char* blk = (char*)eb.p;
   /* packed on the factorisation stream (ordered behind the kernels that produced the block) and complete before
    * the handle leaves this function: the consumer's copy runs in another process, ordered against nothing here */
   if (cn) e = cudaMemcpy2DAsync(blk, (size_t)cn * sizeof(double), f.C, (size_t)f.ldc * sizeof(double),
                                 (size_t)cn * sizeof(double), cn, cudaMemcpyDeviceToDevice, N.stream);
   if (e == cudaSuccess && nd)
      e = cudaMemcpy2DAsync(blk + b_val, rows * sizeof(double), f.L + (size_t)f.nelim * (f.ldl + 1),
                            (size_t)f.ldl * sizeof(double), rows * sizeof(double), nd, cudaMemcpyDeviceToDevice, N.stream);
   if (e == cudaSuccess && nd)
      e = cudaMemcpyAsync(blk + b_val + b_del, f.perm + f.nelim, (size_t)nd * sizeof(int), cudaMemcpyDeviceToDevice, N.stream);
   if (e == cudaSuccess) e = cudaStreamSynchronize(N.stream);
   if (e != cudaSuccess) return (int)e;
   cudaIpcMemHandle_t h;
   e = cudaIpcGetMemHandle(&h, blk);
   if (e != cudaSuccess) return (int)e;
   std::memcpy(handle, &h, sizeof(h));
   *n = cn; *ndelay = nd; *bytes = (int64_t)total; *device_block = blk;
   return 0;
}

/* Consumer side: maps the producer's block and copies `bytes` into dst (a
 * device pointer of the calling process' current device) over NVLink. */
int spral_ssids_b200_ipc_pull(int device, const unsigned char* handle, int64_t bytes, void* dst) {
   /* opened handles are cached for the life of the process: mapping a peer
    * allocation costs milliseconds, the copy itself ~1 ms per GB over NVLink */
   static std::mutex mtx;
   static std::vector<std::pair<std::vector<unsigned char>, void*>> cache;
   void* src = nullptr;
   cudaError_t es = cudaSetDevice(device);      // the caller may be a fresh host thread (current device 0)
   if (es != cudaSuccess) return (int)es;
   {
      std::lock_guard<std::mutex> lock(mtx);
      for (auto& c : cache) if (std::memcmp(c.first.data(), handle, 64) == 0) { src = c.second; break; }
      if (!src) {
         cudaIpcMemHandle_t h;
         std::memcpy(&h, handle, sizeof(h));
         cudaError_t e = cudaIpcOpenMemHandle(&src, h, cudaIpcMemLazyEnablePeerAccess);
         if (e != cudaSuccess) return (int)e;
         cache.push_back({std::vector<unsigned char>(handle, handle + 64), src});
      }
   }
   return (int)cudaMemcpy(dst, src, (size_t)bytes, cudaMemcpyDefault);
}

int spral_ssids_b200_copy_to_host(void* dst, const void* src, int64_t bytes) {
   ABI_GUARD();
   return (int)cudaMemcpy(dst, src, (size_t)bytes, cudaMemcpyDeviceToHost);
}

void* spral_ssids_b200_device_alloc(int device, int64_t bytes) {
   ABI_GUARD();
   void* p = nullptr;
   if (cudaSetDevice(device) != cudaSuccess) return nullptr;
   if (cudaMalloc(&p, (size_t)std::max<int64_t>(bytes, 256)) != cudaSuccess) return nullptr;
   return p;
}
void spral_ssids_b200_device_free(void* p) {
   ABI_GUARD(); if (p) cudaFree(p); }

} /* extern "C" */
