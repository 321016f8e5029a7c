/* Speculative panel segments: shared definitions (host + device).
 *
 * The step-by-step path (factor_kernels.cu) spends five dependent launches on every
 * 32 columns: diagonal block, apply + threshold test, commit, in-panel update, swap.
 * On the large fronts at the top of the tree that chain -- not the FP64 tensor work --
 * is the critical path (measured: 134 us per 32 columns, 1.1 ms per 256-column panel).
 * APP pivoting (reference CPU engine, ldlt_app.cxx) chooses pivots inside the diagonal
 * block only and tests the rows below a posteriori, so a segment of CW = 128 columns is
 * factorised in three launches instead of twenty (kernels: factor_kernels.cu):
 *
 *   chain  ONE CTA per front factorises the CW x CW diagonal block of the segment in
 *          shared memory: four 32 x 32 blocks, each by one warp (diag_warp.cuh, the same
 *          pivoting rules), followed by the triangular solve, D^-1 scaling and threshold
 *          test of the rows of the diagonal block below it and a DMMA trailing update.
 *          Nothing is written to the front; L11, L11*D, D^-1, the block permutations and
 *          the inverses of the four diagonal 32 x 32 blocks of L11 go to a workspace
 *          (SegWS).  Any failed pivot / zero pivot gives up (ok = 0).
 *   tiles  every 128-row tile of the rows below reads its CW columns ONCE into shared
 *          memory and runs the four block steps locally on the FP64 tensor cores: left-looking
 *          update with the earlier blocks, multiplication by the inverse diagonal block,
 *          D^-1 scaling and threshold test; writes L, L*D and a backup of the originals.
 *          No communication between tiles.
 *   commit if every tile passed: diagonal block, D, perm and the row permutation of the
 *          earlier columns are written and the state machine advances by CW columns;
 *          otherwise the tiles restore their rows and the step-by-step path redoes the
 *          segment (failed pivots are rare: 23 delays in 10^6 columns on the benchmark).
 */
#pragma once
#include <cmath>
#include <cstddef>

#ifdef __CUDACC__
#define PV_FN __device__ __forceinline__
#else
#define PV_FN inline
#endif

namespace b200 {

constexpr int CW = 128;            // segment width (columns per chain / tiles / commit round)
constexpr int PV_BS = 32;          // block width inside a segment
constexpr int PV_NB = CW / PV_BS;

/* Per-front workspace of a segment (device global memory). */
struct SegWS {
   double l11[CW * CW];                    // lower factor of the diagonal block (unit diagonal; Cholesky: with its diagonal), final row order, column-major ld = CW
   double ld11[CW * CW];                   // (L11 D)(i, k), i > k, row i in the order it had when column k was eliminated
   double invl[PV_NB][PV_BS * PV_BS];      // inverse of the j-th diagonal 32 x 32 block of L11, column-major ld = 32
   double dinv[2 * CW];                    // D^-1, reference CPU layout (block_ldlt.hxx:375-406); Cholesky: 1 / l_jj
   int lperm[CW];                          // per 32-block: position jb + i holds old position jb + lperm[jb + i]
   int ok;                                 // the chain kernel factorised the whole diagonal block without a failed pivot
   int pad_[3];
};

/* D^-1 of a 32-column block as three coefficient vectors: w_j = c0 y_j + c1 y_{j+1} + c2 y_{j-1} */
PV_FN void pv_dinv_coeffs(const double* d, int j, double inf, double& v0, double& v1, double& v2) {
   v1 = 0.0; v2 = 0.0;
   if (d[2 * j] == inf) { v0 = d[2 * j + 1]; v2 = d[2 * j - 1]; }                          // second column of a 2x2
   else if (j + 1 < PV_BS && d[2 * j + 2] == inf) { v0 = d[2 * j]; v1 = d[2 * j + 1]; }   // first column of a 2x2
   else v0 = d[2 * j];
}

} // namespace b200
