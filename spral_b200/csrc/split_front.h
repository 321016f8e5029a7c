/* Distributed top front, first version ("bulk offload"): the owner of a large front keeps
 * the open panel and the next one, a helper GPU of the same box holds a mirror of the columns further right
 * and applies every trailing update to them.
 *
 * STATUS: part of the default build, switched on at run time (SPRAL_B200_SPLIT=1, spral_b200/dist.py); runs on
 * B200s (round 2: cfg5 on 2 GPUs, inertia / delays / backward error equal to the unsplit run).  The far blocks are
 * dealt round robin over the WORKERS = the owner itself (its blocks stay in its own front and are updated by its
 * look-ahead bulk stream, as without a split) and the helpers.  The protocol is the one tests/c/dist_front_emu.cpp
 * model-checks (no deadlock, no unordered conflicting accesses, complete updates), in its host-driven form: no
 * kernel ever spins.  The two processes talk through a POSIX shared-memory segment; a flag is raised from a
 * cudaLaunchHostFunc callback behind the copy / update it announces and polled by the other host with a
 * time-out.
 *
 *   owner                                              helper (spral_ssids_b200_split_helper_serve)
 *   begin_front : geometry -> shm, phase 1             allocates the mirror (L and L*D, full shape of the
 *                 waits phase 2, maps the mirror,       front), publishes its IPC handles, phase 2
 *                 copies the far columns (>= 2 blocks)
 *   per panel k without a failed pivot:
 *     need_block(k+1): waits updated[k+1], copies       waits ready[k]; UPD_EXPLICIT of block k+2 with panel k
 *                 block k+1 back, then the urgent       (the kernel of the look-ahead bulk update, against a
 *                 update of block k+1 (as today)        Front descriptor that points at the mirror), raises
 *     push_panel(k): L, L*D of panel k, rows of the     updated[k+2]; then the blocks > k+2
 *                 far blocks, into the mirror; ready[k]
 *   first failed pivot -> drain(k): ready[k] = DRAIN,   drains its stream, raises drained
 *                 waits drained, copies every far column back and carries on alone
 *   end_front   : phase 3                               releases the mirror, waits for the next front
 *
 * Included by subtree.cu inside namespace b200 (needs CUDA_TRY, g_pool, Front, launch_update). */
#pragma once
#ifdef SPRAL_B200_SPLIT
/* (system headers: subtree.cu includes <fcntl.h> <sys/mman.h> <sys/stat.h> <unistd.h> <thread> <string> at file scope) */

constexpr int SPLIT_MAXP = 512;          // panels of PW columns per front (n <= 131 072)
constexpr int SPLIT_DRAIN = 2;
constexpr int SPLIT_MAGIC = 0x5b200;

struct SplitShm {
   std::atomic<int> magic;
   std::atomic<int> phase;               // 0 idle (helper), 1 front published (owner), 2 helper ready, 3 front done (owner;
                                         // the helper answers 3 -> 0 when it has released the mirror), 4 exit (owner)
   std::atomic<int> error;               // raised by either side: the other gives up
   int m, n, ldl, ld_is_l;               // geometry of the front; ld_is_l: positive definite (L*D == L)
   int nhelpers, helper_index;           // helpers of this front and this helper's number
   int nworkers, worker_index;           // the far blocks are dealt round robin over the workers: block J >= 2 belongs to worker
                                         // (J - 2) % nworkers; worker 0 is the owner itself when it takes a share
   int base;                             // first column of block 0 (a multiple of the update tile): 0, or where the split was re-started
   unsigned char h_L[64], h_LD[64];      // IPC handles of the helper's mirror
   std::atomic<int> ready[SPLIT_MAXP + 4];   // panel k is in the mirror (1) / the split ends here (SPLIT_DRAIN)
   int k0[SPLIT_MAXP + 4], k1[SPLIT_MAXP + 4];   // its columns
   std::atomic<int> updated[SPLIT_MAXP + 4]; // block J of the mirror has every update the helper owes it
   std::atomic<int> drained;
};

static inline int split_block(int base, int j) { return base + j * PW; }
static inline int split_round_up(int c, int t) { return (c + t - 1) / t * t; }

struct SplitFlagSet { std::atomic<int>* p; int v; };
static void CUDART_CB split_set_flag(void* arg) {
   auto* f = static_cast<SplitFlagSet*>(arg);
   f->p->store(f->v, std::memory_order_release);
   delete f;
}
static inline void split_raise_behind(cudaStream_t s, std::atomic<int>* p, int v) {
   CUDA_TRY(cudaLaunchHostFunc(s, split_set_flag, new SplitFlagSet{p, v}));
}
/* Polls until pred() or the time-out; false on time-out or when the other side raised `error`. */
template <class Pred>
static bool split_wait(SplitShm* sh, double timeout_s, Pred pred) {
   auto t0 = std::chrono::steady_clock::now();
   for (long it = 0;; ++it) {
      if (pred()) return true;
      if (sh && sh->error.load(std::memory_order_acquire)) return false;
      if ((it & 1023) == 1023) {
         if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > timeout_s) return false;
         std::this_thread::yield();
      }
   }
}
static SplitShm* split_map(const char* name, bool create) {
   int fd = shm_open(name, create ? (O_CREAT | O_RDWR) : O_RDWR, 0600);
   if (fd < 0) return nullptr;
   if (create && ftruncate(fd, sizeof(SplitShm)) != 0) { close(fd); return nullptr; }
   struct stat st;
   if (fstat(fd, &st) != 0 || (size_t)st.st_size < sizeof(SplitShm)) { close(fd); return nullptr; }
   void* p = mmap(nullptr, sizeof(SplitShm), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
   close(fd);
   return p == MAP_FAILED ? nullptr : static_cast<SplitShm*>(p);
}

/* ---- owner side ---------------------------------------------------------------------------------------- */
/* One shared-memory segment and one mirror per helper; the far blocks are dealt round robin: block J >= 2 lives on
 * helper (J - 2) % H.  Every helper that still holds a block gets every panel. */
struct SplitOwner {
   struct Link { SplitShm* sh = nullptr; std::string name; double* mL = nullptr; double* mLD = nullptr;
                 int next_k = 0; };      // next_k: panels this helper was given (it waits for ready[next_k] next)
   std::vector<Link> links;
   double timeout_s = 20.0;
   bool active = false;                  // a front is split right now
   bool dead = false;                    // a helper did not answer: do not try again
   bool level_ok = false;                // the level being factorised is one large front (set by factor_subtree)
   bool restart = true;                  // SPRAL_B200_SPLIT_RESTART=0: a drained split stays off for the rest of the front
   bool trace = getenv("SPRAL_B200_TRACE") != nullptr;
   int n_pushed = 0, n_pulled = 0;
   double t_wait_ms = 0, t_push_ms = 0, t_begin_ms = 0, t_drain_ms = 0;   // host time: waiting for blocks, enqueuing pushes, set-up, drains
   const Front* f = nullptr;             // host copy of its descriptor (owner's pointers)
   int base = 0;                         // block j = columns [base + j PW, base + (j + 1) PW): tile aligned, so that what the
                                         // owner's urgent update touches (whole tile columns) ends where the helpers' columns begin
   int p_first = 0;                      // first column of panel 0 of this split (== base unless re-started off a tile boundary)
   bool owner_share = true;              // SPRAL_B200_SPLIT_OWNER_SHARE=0: every far block goes to a helper
   int blk(int j) const { return split_block(base, j); }
   int H() const { return (int)links.size(); }
   int nworkers() const { return H() + (owner_share ? 1 : 0); }
   int worker_of(int J) const { return (J - 2) % nworkers(); }
   bool is_local(int J) const { return owner_share && worker_of(J) == 0; }     // stays in the owner's front
   Link& link_of(int J) { return links[worker_of(J) - (owner_share ? 1 : 0)]; }
   /* block of the tile column tj (T = update tile): < 2 for the columns the owner always keeps */
   int block_of_tile(int tj, int T) const { return tj * T < base ? -1 : (tj * T - base) / PW; }

   static std::string segment_name(const char* shm_name, int h) { return std::string(shm_name) + "_" + std::to_string(h); }
   static SplitOwner* create(const char* shm_name, int nhelpers) {
      auto* o = new SplitOwner;
      for (int h = 0; h < std::max(1, nhelpers); ++h) {
         Link l;
         l.name = segment_name(shm_name, h);
         shm_unlink(l.name.c_str());
         l.sh = split_map(l.name.c_str(), true);
         if (!l.sh) { delete o; return nullptr; }
         std::memset((void*)l.sh, 0, sizeof(SplitShm));
         l.sh->magic.store(SPLIT_MAGIC, std::memory_order_release);
         o->links.push_back(l);
      }
      if (const char* e = getenv("SPRAL_B200_SPLIT_TIMEOUT")) o->timeout_s = atof(e);
      if (const char* e = getenv("SPRAL_B200_SPLIT_RESTART")) o->restart = atoi(e) != 0;
      if (const char* e = getenv("SPRAL_B200_SPLIT_OWNER_SHARE")) o->owner_share = atoi(e) != 0;
      return o;
   }
   ~SplitOwner() {
      for (Link& l : links)
         if (l.sh) { l.sh->phase.store(4, std::memory_order_release); munmap((void*)l.sh, sizeof(SplitShm)); shm_unlink(l.name.c_str()); }
   }
   /* mappings of the helpers' mirrors are kept for the life of the process (opening one costs milliseconds; the helper's
    * allocation cache hands out the same block -- same handle -- for the same front of the next factorisation) */
   static void* open_handle(const unsigned char* h) {
      static std::mutex mtx;
      static std::vector<std::pair<std::vector<unsigned char>, void*>> opened;
      std::lock_guard<std::mutex> lock(mtx);
      for (auto& o : opened) if (std::memcmp(o.first.data(), h, 64) == 0) return o.second;
      cudaIpcMemHandle_t hh; std::memcpy(&hh, h, sizeof(hh));
      void* p = nullptr;
      CUDA_TRY(cudaIpcOpenMemHandle(&p, hh, cudaIpcMemLazyEnablePeerAccess));
      opened.push_back({std::vector<unsigned char>(h, h + 64), p});
      return p;
   }
   void give_up() { for (Link& l : links) l.sh->phase.store(3, std::memory_order_release); dead = true; }
   /* Worth splitting: at least four blocks right of `first_col`, the first column of the next panel (0 at the start of a
    * front; where a drained split is re-started otherwise).  The far columns are copied on stream `s`, in order behind
    * whatever updated them last.  Returns false (and leaves the front alone) when a helper does not answer. */
   bool begin_front(const Front& fr, bool posdef, cudaStream_t s, int first_col = 0) {
      const auto tb0 = std::chrono::steady_clock::now();
      const int T = update_tile_size(true);
      const int b = split_round_up(first_col, T);
      if (active || dead || fr.n - b < 4 * PW || (fr.n - b + PW - 1) / PW > SPLIT_MAXP) return false;
      for (int h = 0; h < H(); ++h) {
         SplitShm* sh = links[h].sh;
         /* the helper has released the previous front (3 -> 0) */
         if (!split_wait(sh, timeout_s, [&] { return sh->phase.load(std::memory_order_acquire) == 0; })) { dead = true; return false; }
         for (int k = 0; k < SPLIT_MAXP + 4; ++k) { sh->ready[k].store(0); sh->updated[k].store(0); }
         sh->drained.store(0);
         sh->m = fr.m; sh->n = fr.n; sh->ldl = fr.ldl; sh->ld_is_l = posdef ? 1 : 0; sh->base = b;
         sh->nhelpers = H(); sh->helper_index = h; links[h].next_k = 0;
         sh->nworkers = nworkers(); sh->worker_index = h + (owner_share ? 1 : 0);
         sh->phase.store(1, std::memory_order_release);
      }
      for (int h = 0; h < H(); ++h) {
         SplitShm* sh = links[h].sh;
         if (!split_wait(sh, timeout_s, [&] { return sh->phase.load(std::memory_order_acquire) == 2; })) { give_up(); return false; }
         links[h].mL = static_cast<double*>(open_handle(sh->h_L));
         links[h].mLD = posdef ? links[h].mL : static_cast<double*>(open_handle(sh->h_LD));
      }
      f = &fr; base = b; p_first = first_col;
      /* the far blocks, each to the helper that owns it, rows from its first column down */
      for (int J = 2; blk(J) < fr.n; ++J) {
         if (is_local(J)) continue;
         const int c0 = blk(J), c1 = std::min(blk(J + 1), fr.n);
         const size_t off = (size_t)c0 + (size_t)c0 * fr.ldl;
         CUDA_TRY(cudaMemcpy2DAsync(link_of(J).mL + off, (size_t)fr.ldl * sizeof(double), fr.L + off, (size_t)fr.ldl * sizeof(double),
                                    (size_t)(fr.m - c0) * sizeof(double), c1 - c0, cudaMemcpyDefault, s));
      }
      active = true; n_pushed = n_pulled = 0;
      t_begin_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tb0).count();
      if (trace) fprintf(stderr, "[split] front m %d n %d: far columns %d.. on %d helper(s) (panels from column %d)\n", fr.m, fr.n,
                         blk(2), H(), first_col);
      return true;
   }
   /* panel index of the panel that starts at column p0, or -1 when the panels are no longer the split's blocks */
   int panel_of(int p0) const { return (p0 >= p_first && (p0 - p_first) % PW == 0) ? (p0 - p_first) / PW : -1; }
   /* a helper still holds a block >= k + 2 */
   bool has_far(int k) const {
      for (int h = 0; h < (int)links.size(); ++h) if (holds_from(h, k + 2)) return true;
      return false;
   }
   /* helper h still holds a block >= J0 */
   bool holds_from(int h, int J0) const {
      for (int J = std::max(2, J0); blk(J) < f->n && J < std::max(2, J0) + nworkers(); ++J)
         if (worker_of(J) == h + (owner_share ? 1 : 0)) return true;
      return false;
   }
   /* Panel k = columns [k0, k1): rows of the far blocks into the mirror of every helper that still holds one, then
    * ready[k] (copy stream). */
   void push_panel(int k, int k0, int k1, cudaStream_t s2) {
      const auto tp0 = std::chrono::steady_clock::now();
      const int r0 = blk(k + 2);
      const size_t off = (size_t)r0 + (size_t)k0 * f->ldl;
      const size_t pitch = (size_t)f->ldl * sizeof(double), width = (size_t)(f->m - r0) * sizeof(double);
      for (int h = 0; h < H(); ++h) {
         if (!holds_from(h, k + 2)) continue;
         Link& l = links[h];
         CUDA_TRY(cudaMemcpy2DAsync(l.mL + off, pitch, f->L + off, pitch, width, k1 - k0, cudaMemcpyDefault, s2));
         if (l.mLD != l.mL) CUDA_TRY(cudaMemcpy2DAsync(l.mLD + off, pitch, f->LD + off, pitch, width, k1 - k0, cudaMemcpyDefault, s2));
         l.sh->k0[k] = k0; l.sh->k1[k] = k1;
         split_raise_behind(s2, &l.sh->ready[k], 1);
         l.next_k = k + 1;
      }
      ++n_pushed;
      t_push_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tp0).count();
   }
   void pull_block(int J, cudaStream_t s) {
      const int c0 = blk(J), c1 = std::min(blk(J + 1), f->n);
      const size_t off = (size_t)c0 + (size_t)c0 * f->ldl;
      CUDA_TRY(cudaMemcpy2DAsync(f->L + off, (size_t)f->ldl * sizeof(double), link_of(J).mL + off, (size_t)f->ldl * sizeof(double),
                                 (size_t)(f->m - c0) * sizeof(double), c1 - c0, cudaMemcpyDefault, s));
      ++n_pulled;
   }
   /* Block J comes back (main stream), in order before the urgent update that touches it. */
   void need_block(int J, cudaStream_t s) {
      if (J < 2 || blk(J) >= f->n || is_local(J)) return;        // a local block is ordered by the bulk stream's events
      SplitShm* sh = link_of(J).sh;
      const auto tw0 = std::chrono::steady_clock::now();
      const bool okw = split_wait(sh, timeout_s, [&] { return sh->updated[J].load(std::memory_order_acquire) != 0; });
      t_wait_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tw0).count();
      if (!okw) {
         sh->error.store(1, std::memory_order_release);
         throw std::runtime_error("split front: a helper did not return a block in time");
      }
      pull_block(J, s);
   }
   /* The split ends at panel k (failed pivot, short panel): every block the helpers still hold -- J >= first_block --
    * comes back; they have applied panels 0 .. k-1 to them. */
   void drain(int k, int first_block, cudaStream_t s, cudaStream_t s2) {
      const auto td0 = std::chrono::steady_clock::now();
      CUDA_TRY(cudaStreamSynchronize(s2));                     // every pushed panel has been announced
      if (trace) fprintf(stderr, "[split] drain at panel %d (%d panels pushed, %d blocks pulled)\n", k, n_pushed, n_pulled);
      for (Link& l : links) l.sh->ready[l.next_k].store(SPLIT_DRAIN, std::memory_order_release);   // where that helper waits
      for (Link& l : links)
         if (!split_wait(l.sh, timeout_s, [&] { return l.sh->drained.load(std::memory_order_acquire) != 0; })) {
            l.sh->error.store(1, std::memory_order_release);
            throw std::runtime_error("split front: a helper did not drain in time");
         }
      for (int J = std::max(2, first_block); blk(J) < f->n; ++J) if (!is_local(J)) pull_block(J, s);
      end_front(s);
      t_drain_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - td0).count();
   }
   void end_front(cudaStream_t s) {
      CUDA_TRY(cudaStreamSynchronize(s));                      // the copies out of the mirrors are done
      if (trace) fprintf(stderr, "[split] front closed (%d panels pushed, %d blocks pulled); host ms so far: set-up %.1f, pushes %.1f, "
                         "waiting for blocks %.1f, drains %.1f\n", n_pushed, n_pulled, t_begin_ms, t_push_ms, t_wait_ms, t_drain_ms);
      for (Link& l : links) l.sh->phase.store(3, std::memory_order_release);
      active = false; f = nullptr;
   }
};

/* ---- helper side --------------------------------------------------------------------------------------- */
static int split_helper_serve(const char* shm_name_base, int device, double timeout_s, int helper_index) {
   const std::string seg = SplitOwner::segment_name(shm_name_base, helper_index);
   const char* shm_name = seg.c_str();
   CUDA_TRY(cudaSetDevice(device));
   SplitShm* sh = nullptr;
   if (!split_wait(nullptr, timeout_s, [&] { sh = split_map(shm_name, false); return sh != nullptr; })) return 1;   // no owner showed up
   if (!split_wait(sh, timeout_s, [&] { return sh->magic.load(std::memory_order_acquire) == SPLIT_MAGIC; })) return 1;
   cudaStream_t s = nullptr;
   CUDA_TRY(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
   configure_update_kernels();
   int fronts_served = 0;
   for (;;) {
      int ph = 0;
      if (!split_wait(sh, timeout_s, [&] { ph = sh->phase.load(std::memory_order_acquire); return ph == 1 || ph == 4; })) break;
      if (ph == 4) break;
      const int m = sh->m, n = sh->n, ldl = sh->ldl, base = sh->base;
      const bool ld_is_l = sh->ld_is_l != 0;
      const int Hn = std::max(1, sh->nworkers), hidx = sh->worker_index;      // workers the blocks are dealt to / this one
      const size_t bytes = (size_t)ldl * n * sizeof(double);
      double* mL = (double*)g_pool.alloc(bytes);
      double* mLD = ld_is_l ? mL : (double*)g_pool.alloc(bytes);
      cudaIpcMemHandle_t h;
      CUDA_TRY(cudaIpcGetMemHandle(&h, mL)); std::memcpy(sh->h_L, &h, 64);
      if (!ld_is_l) { CUDA_TRY(cudaIpcGetMemHandle(&h, mLD)); std::memcpy(sh->h_LD, &h, 64); }
      /* descriptor of the mirror for the update kernel (explicit regions: only L, LD, ldl, m, n are read) */
      Front hf;
      std::memset(&hf, 0, sizeof(hf));
      hf.L = mL; hf.LD = mLD; hf.ldl = ldl; hf.m = m; hf.n = n;
      Front* d_front = (Front*)g_pool.alloc(sizeof(Front));
      CUDA_TRY(cudaMemcpyAsync(d_front, &hf, sizeof(Front), cudaMemcpyHostToDevice, s));
      /* tile lists and regions of every panel, built up front (no failed pivot: panel k == block k) */
      const int T = update_tile_size(true), mt = (m + T - 1) / T, nblk = (n - base + PW - 1) / PW;
      std::vector<MatTile> tiles;
      std::vector<int4> regs(nblk);
      std::vector<size_t> first(nblk + 1, 0), split_at(nblk, 0);
      for (int k = 0; k < nblk; ++k) {
         first[k] = tiles.size();
         for (int J = k + 2; split_block(base, J) < n; ++J) {
            if (J == k + 3) split_at[k] = tiles.size() - first[k];
            if ((J - 2) % Hn != hidx) continue;                     // another helper's block
            const int tj0 = split_block(base, J) / T, tj1 = (std::min(split_block(base, J + 1), n) - 1) / T;
            for (int tj = tj0; tj <= tj1; ++tj)
               for (int ti = tj; ti < mt; ++ti) tiles.push_back({k, ti, tj});
         }
         if (split_block(base, k + 3) >= n) split_at[k] = tiles.size() - first[k];      // block k+2 is the only (or no) far block
      }
      first[nblk] = tiles.size();
      MatTile* d_tiles = (MatTile*)g_pool.alloc(std::max<size_t>(tiles.size(), 1) * sizeof(MatTile));
      int4* d_regs = (int4*)g_pool.alloc((size_t)nblk * sizeof(int4));
      if (!tiles.empty()) CUDA_TRY(cudaMemcpyAsync(d_tiles, tiles.data(), tiles.size() * sizeof(MatTile), cudaMemcpyHostToDevice, s));
      CUDA_TRY(cudaStreamSynchronize(s));
      sh->phase.store(2, std::memory_order_release);

      bool ok = true;
      for (int k = 0; k < nblk && ok; ++k) {
         int r = 0;
         ok = split_wait(sh, timeout_s, [&] {
            r = sh->ready[k].load(std::memory_order_acquire);
            return r != 0 || sh->phase.load(std::memory_order_acquire) >= 3; });
         if (!ok || r == 0) break;                                 // time-out, or the owner closed the front
         if (r == SPLIT_DRAIN) {
            CUDA_TRY(cudaStreamSynchronize(s));
            sh->drained.store(1, std::memory_order_release);
            break;
         }
         /* the panel columns the owner announced; the region table entry is uploaded now (pageable, 16 bytes) */
         regs[k] = make_int4(0, sh->k0[k], sh->k1[k], split_block(base, k + 2));
         CUDA_TRY(cudaMemcpyAsync(d_regs + k, &regs[k], sizeof(int4), cudaMemcpyHostToDevice, s));
         const size_t nt_all = first[k + 1] - first[k], nt_a = split_at[k];
         if (nt_a) launch_update(d_front, d_tiles + first[k], (int)nt_a, UPD_EXPLICIT, true, s, 0, d_regs);
         if (k % Hn == hidx) split_raise_behind(s, &sh->updated[k + 2], 1);   // block k+2 (this helper's) may go back
         if (nt_all > nt_a) launch_update(d_front, d_tiles + first[k] + nt_a, (int)(nt_all - nt_a), UPD_EXPLICIT, true, s, 0, d_regs);
      }
      if (!ok) sh->error.store(2, std::memory_order_release);
      /* the owner reads the mirror until it closes the front */
      split_wait(sh, timeout_s, [&] { return sh->phase.load(std::memory_order_acquire) >= 3; });
      CUDA_TRY(cudaStreamSynchronize(s));
      g_pool.release(d_regs); g_pool.release(d_tiles); g_pool.release(d_front);
      if (!ld_is_l) g_pool.release(mLD);
      g_pool.release(mL);
      ++fronts_served;
      if (!ok) break;
      { int expect = 3; sh->phase.compare_exchange_strong(expect, 0, std::memory_order_acq_rel); }   // (4 stays 4)
      /* phase 3 -> wait for the next front (phase 1 again) or the end (phase 4) */
      if (!split_wait(sh, timeout_s, [&] { int p = sh->phase.load(std::memory_order_acquire); return p == 1 || p == 4; })) break;
   }
   cudaStreamDestroy(s);
   munmap((void*)sh, sizeof(SplitShm));
   return fronts_served > 0 ? 0 : 2;
}
#endif /* SPRAL_B200_SPLIT */
