/* Internal types of the B200 SSIDS numeric engine (not part of the C ABI).
 *
 * Replaces the host scheduling of src/ssids/gpu/{subtree,factor,dense_factor,
 * solve}.f90 and the kernels of the src/ssids/gpu/kernels sources (reference tree)
 * with a level-set scheduler whose pivoting state machine lives on the device:
 * no host round trip per block column, one host sync per tree level.
 *
 * Front layout in HBM (column-major, lower triangle only):
 *   rows/cols 0..n0-1        the node's own fully-summed variables
 *   rows/cols n0..n-1        delayed columns received from children (n = n0+ndin)
 *   rows      n..m-1         contribution rows (m = m0+ndin)
 *   L is m x n with leading dimension ldl (even -> every column 16 B aligned),
 *   D^-1 is 2 x n in the reference CPU layout (block_ldlt.hxx:375-406),
 *   the contribution block is (m-n)^2, lower triangle valid.
 */
#pragma once
#include <atomic>
#include <cstdint>
#include <cuda_runtime.h>
#include <vector>
#include "spral_ssids_b200.h"
#include "solve_types.h"

namespace b200 {

struct SegWS;
extern std::atomic<long> g_launches;   // kernels launched (the gpu_launches figure of bench.py)

constexpr int BS = 32;    // block column width (inner step)
constexpr int PW = 256;   // outer panel width (a multiple of BS)

/* Per-front workspace written by the diagonal-block kernel each inner step. */
struct BlockWS {
   double l11[BS * BS];   // unit lower factor of the permuted diagonal block, col-major ld=BS
   double ld11[BS * BS];  // L11*D (the columns before scaling by D^-1), col-major ld=BS
   double a0[BS * BS];    // the diagonal block before factorisation, full symmetric, unpermuted
   double dinv[2 * BS];   // D^-1, reference CPU layout
   int lperm[BS];         // position j of the permuted block holds old position lperm[j]
   int zfrom;             // columns >= zfrom of the block are tentative zero pivots (BS if none)
   int pad_[3];
};

/* One frontal matrix (device resident). */
struct Front {
   /* geometry (host-filled when the front's level is reached) */
   double* L;        // m x n, ld = ldl
   double* D;        // 2*n (indefinite only)
   int* perm;        // n: global pivot-order index (1-based) of each local column
   double* C;        // contribution block, (m-n)^2, ld = ldc (null if m == n)
   double* LD;       // scratch: L*D of the eliminated columns, same shape as L (== L if posdef)
   double* BK;       // scratch: m x BS backup of the current block column (permuted originals)
   BlockWS* ws;
   const int* rows;  // device rlist of the node (m0 global 1-based indices)
   int ldl, m, n;    // sizes including incoming delays
   int m0, n0, ndin;
   int ldc;
   /* pivoting state machine (device-advanced; see factor_kernels.cu) */
   int done;         // columns eliminated so far
   int end;          // end of the candidate range of this pass
   int pass_start;   // value of done when this pass began
   int p0, pend0;    // current outer panel: first column and initial end
   int pend;         // end of the still-untried candidates of the panel
   int panel_open;
   int bs;           // width of the current block column
   int first_fail;   // first failing column of the current block column (bs if none)
   int step_valid;   // the current block column still has to be accounted
   int finished;
   int first_pass_done;  // columns eliminated at the end of the first pass (-1 until then)
   /* results */
   int nelim;
   int num_neg, num_two, num_zero;
   int flag;         // 0, -5 singular (action=false), -6 not positive definite
   /* speculative segments of the panel (panel_v2.h; null / 0 unless SPRAL_B200_PANEL_V2) */
   SegWS* sws;       // chain workspace of this front
   int seg_valid;    // a segment of CW columns starting at `done` is in flight
   int seg_ok;       // ... its diagonal block was factorised by the chain kernel
   int seg_fail;     // ... a row below failed the a-posteriori test: the tiles roll back
   int spec_off;     // the rest of this panel is done step by step
   int spec_fails;   // penalty account of given-up segments: at SPEC_MAX_FAILS the front stops speculating
   int pad2_;
};
constexpr int SPEC_PENALTY = 3;      // pivot_state.h: account_segment
constexpr int SPEC_MAX_FAILS = 9;

/* A child contribution to be extend-added into a parent front (either a child
 * front of the same part or a contribution block received from another part). */
struct AsmSrc {
   const double* C;   // cm x cm lower, ld = ldc (null: nothing but delays)
   int ldc, cm;
   const int* map;    // cm entries: 1-based position in the parent's symbolic row list (rlist_direct)
   const double* dval; // delayed columns: (ndelay + cm) x ndelay, ld = lddelay
   int lddelay, ndelay;
   const int* dperm;  // ndelay global indices
   int parent;        // index of the parent Front
   int delay_col;     // first local column of these delays in the parent
   int npassl;        // number of leading child columns that map into the parent's own columns
};

struct FactorParams {
   double u, small;
   int action;
};

/* RowTile (front index, row-tile index): solve_types.h */
/* One tile of work for the DMMA update kernels: absolute tile row / column of the front. */
struct MatTile { int front; int ti; int tj; };

/* ---- launch wrappers (factor_kernels.cu) ---- */
void launch_scatter_a(Front* fronts, const int2* work, int nwork, const int64_t* nlist,
      const int64_t* nptr, const int* node_of_front, const double* aval,
      const double* scaling, cudaStream_t s);
void launch_assemble(Front* fronts, const AsmSrc* srcs, const int2* work, int nwork,
      bool to_contrib, bool use_atomics, cudaStream_t s);
void launch_delays(Front* fronts, const AsmSrc* srcs, const int2* work, int nwork, cudaStream_t s);
void launch_diag(Front* fronts, const int* flist, int count, bool posdef, bool new_panel,
      const FactorParams& prm, cudaStream_t s);
void launch_apply(Front* fronts, const RowTile* work, int nwork, bool posdef,
      const FactorParams& prm, cudaStream_t s);
void launch_commit(Front* fronts, const RowTile* work, int nwork, cudaStream_t s);
void launch_swap(Front* fronts, const RowTile* work, int nwork, bool outer, cudaStream_t s);
void launch_finalize(Front* fronts, const int* flist, int count, bool posdef, cudaStream_t s);
void launch_snapshot(Front* fronts, const int* flist, int count, int* snap, cudaStream_t s);
/* speculative panel segments (panel_v2.h) */
int panel_segment_width();
size_t panel_segment_ws_bytes();
void configure_panel_kernels();
void launch_panel_chain(Front* fronts, const int* flist, int count, bool posdef, bool new_panel, const FactorParams& prm,
      cudaStream_t s);
void launch_panel_tiles(Front* fronts, const RowTile* work, int nwork, bool posdef, const FactorParams& prm, cudaStream_t s);
void launch_seg_commit(Front* fronts, const RowTile* work, int nwork, bool posdef, cudaStream_t s);
int assemble_cols_per_cta();
int scatter_chunk();

/* ---- DMMA update kernels (gemm_dmma.cu) ---- */
enum UpdateMode { UPD_INNER = 0, UPD_OUTER = 1, UPD_CONTRIB = 2, UPD_EXPLICIT = 3,
                  UPD_SEG = 4 };   // UPD_SEG: the panel's columns right of an accepted speculative segment, K = the segment
void launch_update(Front* fronts, const MatTile* work, int nwork, UpdateMode mode,
      bool big_tiles, cudaStream_t s, int max_ctas = 0,      // max_ctas: cap on the persistent grid (0 = all SMs, < 0 = one CTA per tile)
      const int4* xregs = nullptr);                         // UPD_EXPLICIT (large tiles only): {front, k0, k1, c_lo} per region
int device_sm_count();
int update_tile_size(bool big_tiles);
int inner_tile_size(bool big_tiles);
void configure_update_kernels();   // per device: opt in to > 48 KB dynamic shared memory

/* ---- solve (solve_kernels.cu) ---- */
/* SolveFront: solve_types.h */
int solve_block();
void launch_transpose_rhs(double* x, int ldx, double* xt, int n, int nr, bool to_xt, cudaStream_t s);
int solve_rhs_chunk(int nrhs);
int solve_max_chunk();
void configure_solve_kernels();
void launch_fwd_level(const SolveFront* fronts, const RowTile* work, int nwork, int nsteps,
      bool posdef, int nr, double* x, int ldx, double* ywork, cudaStream_t s,
      unsigned int* bar = nullptr);      // bar: zeroable device counter -> one cooperative launch per level
void launch_fwd_flush(const SolveFront* fronts, int first, int count, int nrhs, double* x, int ldx,
      const double* ywork, cudaStream_t s);
void launch_diag_solve(const SolveFront* fronts, int first, int count, int nrhs, double* x, int ldx,
      cudaStream_t s);
void launch_bwd_level(const SolveFront* fronts, int first, int count, const RowTile* work, int nwork,
      const int* wbeg, int nsteps, bool posdef, int nr, double* x, int ldx, double* pbuf,
      cudaStream_t s, unsigned int* bar = nullptr);

/* wide sweeps for levels of large fronts (solve_wide.h; opt-in) */
int solve_wide_block();
/* inverses of the 32 x 32 diagonal blocks of the fronts that carry SolveFront::Linv; work = {front, block} */
void launch_build_linv(const SolveFront* fronts, const int2* work, int nwork, bool posdef, cudaStream_t s);
/* Streams and events of the look-ahead inside a wide sweep: the G work on the rows of the next block stays on the
 * sweep's stream, the rest runs on the two far streams beside the following T kernels (solve_wide.h: SW_NEAR / SW_FAR).
 * The backward sweep then needs TWO accumulators per front (pbuf of 2 x count x 256 x nr doubles).
 * half: no front of the level eliminates more than 128 columns (the T kernels then keep 128 rows in shared memory). */
struct SolveAux {
   cudaStream_t far[2] = {nullptr, nullptr};
   cudaEvent_t evT[4] = {nullptr, nullptr, nullptr, nullptr}, evF[4] = {nullptr, nullptr, nullptr, nullptr};
   void create();
   void destroy();
};
void launch_fwd_level_wide(const SolveFront* fronts, int first, int count, const RowTile* work, int nwork,
      int nblk, bool posdef, int nr, double* x, double* ywork, cudaStream_t s, SolveAux* aux = nullptr, bool half = false);
void launch_bwd_level_wide(const SolveFront* fronts, int first, int count, const RowTile* work, int nwork,
      const int* wbeg, int nblk, bool posdef, int nr, double* x, double* pbuf, cudaStream_t s, SolveAux* aux = nullptr, bool half = false);

} // namespace b200
