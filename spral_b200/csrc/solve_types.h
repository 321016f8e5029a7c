/* Plain types shared by the solve kernels and their host emulation (no CUDA headers). */
#pragma once

namespace b200 {

constexpr int RT = 128;   // row tile of the panel and solve kernels

/* One tile of work for the panel / solve kernels: front index and row-tile index. */
struct RowTile { int front; int tile; };

/* Immutable view of a factorised front for the solves.  Row i of the front is entry
 * idx(i) of x: perm[i] for i < n (eliminated and delayed columns), rows[n0 + i - n] below. */
struct SolveFront {
   const double* L; const double* D; const int* perm; const int* rows;
   int ldl, m, n, n0, m0, nelim;
   /* fronts of the wide sweeps' critical path (256+ eliminated columns): the inverses of the 32 x 32 diagonal blocks of
    * L, block b at Linv + 1024 b, row-major, zero above the diagonal (k_build_linv); *linv_bad != 0: an inverse came out
    * large or not finite, the T kernels then substitute as they do for the fronts without inverses (Linv == nullptr) */
   const double* Linv = nullptr; const int* linv_bad = nullptr;
};

} // namespace b200
