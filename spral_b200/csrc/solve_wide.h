/* "Wide" triangular sweeps (default; SPRAL_B200_SOLVE_WIDE=0 selects the 32-column kernels): the
 * column blocks that need a device-wide synchronisation are SWB = 256 columns wide instead of 32.
 *
 * The level kernels of solve_kernels.cu advance all fronts of a level 32 columns
 * per launch (or per grid barrier); a front with 16 000 eliminated columns needs
 * 511 such steps per sweep, each bound by the latency of its dependent loads.
 * Here a block of 256 columns costs two launches:
 *   forward   T: ONE CTA per front solves the 256 x 256 triangle in 8 sub-steps of
 *                32 columns that are separated by CTA barriers only; every thread
 *                owns a row and keeps the next 32-column slab of its row in
 *                registers (prefetched one sub-step ahead), the rows of the
 *                sub-block publish the 32 x 32 diagonal block to shared memory;
 *             G: all (front, row-tile) CTAs subtract L(rows below, block) * y.
 *   backward  G: L(rows below, block)^T x per row tile, added (atomics) into one
 *                256 x nrhs accumulator per front; T: ONE CTA per front takes the
 *                accumulator (and clears it), solves the transposed triangle,
 *                sub-blocks last to first, every thread owning a column.
 * The G kernels are the bandwidth part (they read all of L below the diagonal blocks once per
 * sweep): 256 threads per 128-row tile, >= 16 independent loads in flight per thread.
 * Same mathematics and data as the 32-column kernels (NumericSubtree.hxx:286-418
 * of the reference CPU engine: gather, trsv/trsm + gemv/gemm, scatter); sums are
 * taken over up to 256 terms before they are applied, so results differ from the
 * narrow path by rounding only.
 *
 * The bodies are written against a context (tid, sync, shfl, atomic_add) and
 * compiled twice: by nvcc for the kernels in solve_kernels.cu and by g++ for
 * tests/c/solve_wide_emu.cpp, which runs them on host threads against a plain
 * gather / substitute / scatter implementation.
 */
#pragma once
#include <cstddef>
#include "solve_types.h"

#ifdef __CUDACC__
#define SW_FN __device__ __forceinline__
#else
#define SW_FN inline
#endif

namespace b200 {

constexpr int SWB = 256;     // wide block: columns per T/G pair
constexpr int SSB = 32;      // sub-block inside T
constexpr int SW_TT = 256;   // threads of a T kernel CTA (one per row / column of the block)
constexpr int SW_LK = SSB + 1;

#define SW_XI(g, k) ((size_t)(g) * NR + (k))

SW_FN int sw_min(int a, int b) { return a < b ? a : b; }
SW_FN int sw_row_index(const SolveFront& f, int i) {
   return (i < f.n ? f.perm[i] : f.rows[f.n0 + i - f.n]) - 1;
}

/* shared memory of the T kernels: xs/vs [SWB * NR], lkk [SSB * SW_LK] doubles */
/* row stride of the right-hand-side blocks in shared memory: not a multiple of 32 doubles (threads that own
 * consecutive rows would all hit one bank with 32 right-hand sides) and even, so that a row stays 16-byte
 * aligned and the broadcast reads of another row's values vectorise */
template <int NR> constexpr int sw_xld() { return NR == 1 ? 1 : (NR >= 16 ? NR + 4 : NR + 2); }
/* with 16+ right-hand sides the update of the rows below a sub-block runs on the FP64 tensor cores (cx.mma, m8n8k4 with
 * M = right-hand side): the 32-column slab of L goes through shared memory, k-major with a stride = 4 mod 16 doubles
 * (bank-conflict-free fragment loads; so is the stride NR + 4 of the right-hand sides) */
template <int NR> constexpr bool sw_T_mma() { return NR >= 16; }
/* HB = rows of the block a T CTA keeps in shared memory: SWB, or 128 on the levels whose fronts eliminate at most 128
 * columns (one block; the CTA then has 128 threads): 62 KB instead of 116 KB of shared memory with 16 right-hand sides
 * and half the registers, three CTAs per SM instead of one -- those levels have thousands of fronts */
constexpr int SW_LLD = SWB + 4;
template <int NR, int HB = SWB> constexpr size_t sw_T_smem_doubles() {
   return (size_t)HB * sw_xld<NR>() + (size_t)SSB * SW_LK + (sw_T_mma<NR>() ? (size_t)SSB * (HB + 4) : 0);
}

/* xs(rows of the 8-row groups of this warp that lie in [lo, hi)) -= slab * xs(sub-block rows): the tensor-core form of
 *   for k: s = xs[t][k]; for j: s -= cur[j] * xs[jb + j][k]
 * ls[j * LLD + t] = cur[j] of thread t (0 for the threads that do not take part). */
template <int NR, int LLD, class Ctx>
SW_FN void sw_mma_update(Ctx& cx, double* xs, const double* ls, int jb, int lo, int hi) {
   constexpr int XLD = sw_xld<NR>();
   const int lane = cx.tid() & 31, warp = cx.tid() >> 5;
   for (int i = 0; i < 4; ++i) {
      const int rg = 32 * warp + 8 * i;
      if (rg + 8 <= lo || rg >= hi) continue;
      double acc[NR / 8][2];
      #pragma unroll
      for (int j = 0; j < NR / 8; ++j) { acc[j][0] = 0.0; acc[j][1] = 0.0; }
      #pragma unroll
      for (int kk = 0; kk < SSB; kk += 4) {
         const double bfr = ls[(size_t)(kk + (lane & 3)) * LLD + rg + (lane >> 2)];
         #pragma unroll
         for (int j = 0; j < NR / 8; ++j) {
            const double afr = xs[(size_t)(jb + kk + (lane & 3)) * XLD + 8 * j + (lane >> 2)];
            cx.mma(acc[j][0], acc[j][1], afr, bfr);
         }
      }
      #pragma unroll
      for (int j = 0; j < NR / 8; ++j)
         #pragma unroll
         for (int e = 0; e < 2; ++e) xs[(size_t)(rg + 2 * (lane & 3) + e) * XLD + 8 * j + (lane >> 2)] -= acc[j][e];
   }
}
/* Look-ahead (two streams, solve_kernels.cu): the G work of a block is split into the rows of the NEXT block of the
 * front ("near": the only rows the next T kernel waits for) and the rest ("far": runs beside the following T kernels).
 * part 0 = all rows below the block, 1 = near, 2 = far; [rlo, rhi) are the rows of the front a G kernel may touch. */
enum { SW_ALL = 0, SW_NEAR = 1, SW_FAR = 2 };
SW_FN void sw_part_rows(const SolveFront& f, int kb, int w, int part, int& rlo, int& rhi) {
   const int below = kb + w;
   const int ne = below < f.nelim ? sw_min(below + SWB, f.nelim) : below;       // end of the next block's rows
   rlo = part == SW_FAR ? ne : below;
   rhi = part == SW_NEAR ? ne : f.m;
}
constexpr int SW_GT = 256;   // threads of a G kernel CTA
/* forward G: ys [SWB * NR] + the partial sums of the second column half [RT * NR] doubles */
template <int NR> constexpr size_t sw_fG_smem_doubles() { return (size_t)(SWB / 2) * NR + (size_t)RT * NR; }
/* backward G: nothing (registers and shuffles only) */
template <int NR> constexpr size_t sw_bG_smem_doubles() { return 1; }

/* ---- forward, T: y(block) = L(block, block)^-1 x(block); y -> ywork ---------------- */
/* NR right-hand sides of the NRT that share a row of x / ywork (the caller offsets the pointers to the first one):
 * the right-hand sides are independent, so a block of 64 is solved by four CTAs of 16 side by side. */
template <int NR, int NRT, bool POSDEF, int HB, class Ctx>
SW_FN void fwd_wide_T_h(Ctx& cx, const SolveFront& f, int blk, const double* x, double* ywork, double* smem) {
   const int kb = blk * SWB;
   if (kb >= f.nelim) return;
   constexpr int XLD = sw_xld<NR>();
   constexpr int LLD = HB + 4;
   double* xs = smem;
   double* lkk = smem + (size_t)HB * XLD;
   double* ls = lkk + (size_t)SSB * SW_LK;          // (tensor-core variant only)
   const int w = sw_min(SWB, f.nelim - kb);
   const int t = cx.tid(), lane = t & 31, warp = t >> 5;
   const size_t ldl = (size_t)f.ldl;
   const bool arow = t < w;
   const int g = arow ? f.perm[kb + t] - 1 : -1;
   if (t < HB) {
      #pragma unroll
      for (int k = 0; k < NR; ++k) xs[(size_t)t * XLD + k] = arow ? x[(size_t)g * NRT + k] : 0.0;
   }
   const double* Lrow = f.L + (size_t)(kb + t) + (size_t)kb * ldl;     // row kb+t of the block, from column kb
   /* inverse diagonal blocks (solve_types.h): the 32-step substitution of a sub-block becomes a 32-term product.  Four
    * entries of the block per thread (256 threads) or eight (128), loaded one sub-step ahead. */
   constexpr int TT = HB < SW_TT ? HB : SW_TT, IPT = SSB * SSB / TT;
   const bool use_inv = f.Linv != nullptr && *f.linv_bad == 0;
   const double* inv_blk = use_inv ? f.Linv + (size_t)(kb / SSB) * (SSB * SSB) + (size_t)t * IPT : nullptr;
   double icur[IPT], inxt[IPT];
   #pragma unroll
   for (int q = 0; q < IPT; ++q) icur[q] = use_inv ? inv_blk[q] : 0.0;
   double cur[SSB], nxt[SSB];
   {
      const int wd0 = sw_min(SSB, w);
      #pragma unroll
      for (int j = 0; j < SSB; ++j) cur[j] = (arow && j < wd0) ? Lrow[(size_t)j * ldl] : 0.0;
   }
   constexpr int NWARP = (HB < SW_TT ? HB : SW_TT) / 32;      // the CTA has min(HB, 256) threads
   constexpr int NRW = (NR + NWARP - 1) / NWARP;
   for (int jb = 0; jb < w; jb += SSB) {
      const int wd = sw_min(SSB, w - jb);
      if (use_inv) {                               // everybody publishes its entries of the inverse block
         #pragma unroll
         for (int q = 0; q < IPT; ++q) { const int e = t * IPT + q; lkk[(e / SSB) * SW_LK + e % SSB] = icur[q]; }
      } else if (t >= jb && t < jb + SSB) {        // the rows of the sub-block publish its diagonal block
         const int i = t - jb;
         #pragma unroll
         for (int j = 0; j < SSB; ++j) lkk[i * SW_LK + j] = (i < wd && j < wd && i >= j) ? cur[j] : 0.0;
      }
      const int jn = jb + SSB;                    // next slab of the rows below this sub-block
      if (use_inv && jn < w) {
         #pragma unroll
         for (int q = 0; q < IPT; ++q) inxt[q] = inv_blk[(size_t)(jn / SSB) * (SSB * SSB) + q];
      }
      if (sw_T_mma<NR>()) {
         #pragma unroll
         for (int j = 0; j < SSB; ++j) if (t < HB) ls[(size_t)j * LLD + t] = (arow && t >= jn) ? cur[j] : 0.0;
      }
      if (jn < w) {
         const int wdn = sw_min(SSB, w - jn);
         #pragma unroll
         for (int j = 0; j < SSB; ++j) nxt[j] = (arow && t >= jn && j < wdn) ? Lrow[(size_t)(jn + j) * ldl] : 0.0;
      } else {
         #pragma unroll
         for (int j = 0; j < SSB; ++j) nxt[j] = 0.0;
      }
      cx.sync();
      {  /* forward substitution: lanes are the rows of the sub-block, the right-hand sides are dealt to the warps */
         double v[NRW];
         if (use_inv) {                            // y = Linv_bb r: four partial sums per right-hand side
            #pragma unroll
            for (int q = 0; q < NRW; ++q) {
               const int k = warp * NRW + q;
               double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
               if (k < NR) {
                  #pragma unroll
                  for (int j = 0; j < SSB; j += 4) {
                     s0 += lkk[lane * SW_LK + j] * xs[(size_t)(jb + j) * XLD + k];
                     s1 += lkk[lane * SW_LK + j + 1] * xs[(size_t)(jb + j + 1) * XLD + k];
                     s2 += lkk[lane * SW_LK + j + 2] * xs[(size_t)(jb + j + 2) * XLD + k];
                     s3 += lkk[lane * SW_LK + j + 3] * xs[(size_t)(jb + j + 3) * XLD + k];
                  }
               }
               v[q] = (s0 + s1) + (s2 + s3);
            }
            cx.sync_warp();                        // every lane has read the old values of the sub-block's rows
         } else {
         #pragma unroll
         for (int q = 0; q < NRW; ++q) { const int k = warp * NRW + q; v[q] = (k < NR) ? xs[(size_t)(jb + lane) * XLD + k] : 0.0; }
         #pragma unroll 4
         for (int j = 0; j < wd; ++j) {
            const double l = (lane > j && lane < wd) ? lkk[lane * SW_LK + j] : 0.0;
            const double dj = POSDEF ? lkk[j * SW_LK + j] : 1.0;
            #pragma unroll
            for (int q = 0; q < NRW; ++q) {
               double yj = cx.shfl(v[q], j);
               if (POSDEF) { yj /= dj; if (lane == j) v[q] = yj; }
               v[q] -= l * yj;
            }
         }
         }
         #pragma unroll
         for (int q = 0; q < NRW; ++q) { const int k = warp * NRW + q; if (k < NR) xs[(size_t)(jb + lane) * XLD + k] = v[q]; }
      }
      cx.sync();
      if (sw_T_mma<NR>()) {
         if (jn < w) sw_mma_update<(NR >= 16 ? NR : 16), LLD>(cx, xs, ls, jb, jn, w);
         cx.sync();                                // the slab is re-written at the top of the next sub-step
      } else if (arow && t >= jn) {               // rows of the block below the sub-block
         #pragma unroll
         for (int k = 0; k < NR; ++k) {
            double s = xs[(size_t)t * XLD + k];
            #pragma unroll
            for (int j = 0; j < SSB; ++j) s -= cur[j] * xs[(size_t)(jb + j) * XLD + k];
            xs[(size_t)t * XLD + k] = s;
         }
      }
      #pragma unroll
      for (int j = 0; j < SSB; ++j) cur[j] = nxt[j];
      #pragma unroll
      for (int q = 0; q < IPT; ++q) icur[q] = inxt[q];
   }
   cx.sync();
   if (arow) {
      #pragma unroll
      for (int k = 0; k < NR; ++k) ywork[(size_t)g * NRT + k] = xs[(size_t)t * XLD + k];
   }
}
template <int NR, int NRT, bool POSDEF, class Ctx>
SW_FN void fwd_wide_T(Ctx& cx, const SolveFront& f, int blk, const double* x, double* ywork, double* smem) {
   fwd_wide_T_h<NR, NRT, POSDEF, SWB>(cx, f, blk, x, ywork, smem);
}

/* ---- forward, G: x(rows below the block) -= L(rows, block) * y --------------------- */
/* SW_GT = 256 threads per (row tile, column half `ch` of the block): thread (row r0 + (t & 127), quarter h = t >> 7)
 * sums 64 columns of its row, eight independent loads at a time; the quarters meet in shared memory, the halves in x
 * (atomics).  Two CTAs per tile: twice as many bytes in flight per SM on the levels that have few tiles. */
constexpr int SW_FSPLIT = 2;
constexpr int SW_NSPLIT = 16;   // column split of the near launches (forward); the backward ones split 4 ways
/* FS = CTAs that share the block's columns (near launches use SW_NSPLIT: a few tiles on the critical path of the sweep,
 * each thread then has ONE batch of loads instead of a dependent chain of them) */
template <int NR, class Ctx, int FS = SW_FSPLIT>
SW_FN void fwd_wide_G(Ctx& cx, const SolveFront& f, int tile, int blk, int ch, double* x, const double* ywork, double* smem,
      int rpart = SW_ALL) {
   const int kb = blk * SWB;
   if (kb >= f.nelim) return;
   const int w = sw_min(SWB, f.nelim - kb);
   int rlo, rhi;
   sw_part_rows(f, kb, w, rpart, rlo, rhi);
   if (rpart == SW_NEAR) tile += rlo / RT;           // near launches count the tiles from the first row below the block
   const int r0 = tile * RT;
   if (r0 + RT <= rlo || r0 >= rhi) return;
   constexpr int CH = SWB / FS;                     // columns of this CTA
   const int c0 = ch * CH;
   if (c0 >= w) return;
   const int wc = sw_min(CH, w - c0);
   double* ys = smem;
   double* part = smem + (size_t)CH * NR;
   const int t = cx.tid();
   for (int e = t; e < wc; e += SW_GT) {
      const int g = f.perm[kb + c0 + e] - 1;
      #pragma unroll
      for (int k = 0; k < NR; ++k) ys[(size_t)e * NR + k] = ywork[SW_XI(g, k)];
   }
   cx.sync();
   const int rl = t & (RT - 1), h = t / RT;
   const int r = r0 + rl;
   const bool active = r >= rlo && r < rhi;
   double acc[NR];
   #pragma unroll
   for (int k = 0; k < NR; ++k) acc[k] = 0.0;
   if (active) {
      const size_t ldl = (size_t)f.ldl;
      const double* Lr = f.L + r + (size_t)(kb + c0) * ldl;
      const int jend = sw_min(wc, h * (CH / 2) + CH / 2);
      constexpr int NL = 8;                        // independent loads in flight per thread
      for (int j0 = h * (CH / 2); j0 < jend; j0 += NL) {
         double l[NL];
         #pragma unroll
         for (int q = 0; q < NL; ++q) l[q] = (j0 + q < jend) ? Lr[(size_t)(j0 + q) * ldl] : 0.0;
         #pragma unroll
         for (int q = 0; q < NL; ++q) {
            const int j = (j0 + q < jend) ? j0 + q : j0;
            #pragma unroll
            for (int k = 0; k < NR; ++k) acc[k] += l[q] * ys[(size_t)j * NR + k];
         }
      }
   }
   if (h == 1) {
      #pragma unroll
      for (int k = 0; k < NR; ++k) part[(size_t)rl * NR + k] = acc[k];
   }
   cx.sync();
   if (h == 0 && active) {
      const int g = sw_row_index(f, r);
      /* several CTAs (column halves, sibling fronts) add into a row of x */
      #pragma unroll
      for (int k = 0; k < NR; ++k) cx.atomic_add(&x[SW_XI(g, k)], -(acc[k] + part[(size_t)rl * NR + k]));
   }
}

/* block handled by front f at backward step s (last block first) */
SW_FN int sw_bwd_block(const SolveFront& f, int step) {
   const int nblk = (f.nelim + SWB - 1) / SWB;
   return nblk - 1 - step;
}

/* ---- backward, G: acc(j, k) += sum over the tile's rows r below the block of L(r, kb+j) x(r, k) ---- */
/* SW_GT = 256 threads = 8 warps; warp v owns the columns [32 v, 32 v + 32) of the block, lane l the rows r0 + l + 32 q
 * (q < 4) of the tile: every load is a coalesced 256-byte row segment of one column, 4 x CG of them in flight per
 * thread.  The lane sums of CG columns are reduced with shuffles and added to the front's accumulator `acc`
 * (SWB x NR doubles, zero when the step starts; several tiles add into it: atomics). */
template <int NR, class Ctx>
SW_FN void bwd_wide_G(Ctx& cx, const SolveFront& f, int tileidx, int step, const double* x, double* acc, double* /*smem*/,
      int part = SW_ALL, int cpart = 0, int ncpart = 1) {
   const int b = sw_bwd_block(f, step);
   if (b < 0) return;
   const int kb = b * SWB;
   const int w = sw_min(SWB, f.nelim - kb);
   int rlo, rhi;
   sw_part_rows(f, kb, w, part, rlo, rhi);
   if (part == SW_NEAR) tileidx += rlo / RT;
   const int r0 = tileidx * RT;
   if (r0 + RT <= rlo || r0 >= rhi) return;         // no row of this tile in the part below the block
   const int t = cx.tid(), lane = t & 31, warp = t >> 5;
   const size_t ldl = (size_t)f.ldl;
   double xv[4][NR];
   bool act[4];
   #pragma unroll
   for (int q = 0; q < 4; ++q) {
      const int r = r0 + lane + 32 * q;
      act[q] = r >= rlo && r < rhi;
      const int g = act[q] ? sw_row_index(f, r) : 0;
      #pragma unroll
      for (int k = 0; k < NR; ++k) xv[q][k] = act[q] ? x[SW_XI(g, k)] : 0.0;
   }
   constexpr int CG = (NR <= 2) ? 8 : 4;           // columns per batch: CG * NR running sums per thread
   const int c0 = warp * 32;
   const int cw = 32 / ncpart;                     // near launches: ncpart CTAs share a warp's 32 columns
   for (int cb = cpart * cw; cb < (cpart + 1) * cw && c0 + cb < w; cb += CG) {
      double l[CG][4];
      #pragma unroll
      for (int c = 0; c < CG; ++c) {
         const bool cin = c0 + cb + c < w;
         const double* Lc = f.L + (size_t)r0 + lane + (size_t)(kb + c0 + cb + (cin ? c : 0)) * ldl;
         #pragma unroll
         for (int q = 0; q < 4; ++q) l[c][q] = (cin && act[q]) ? Lc[32 * q] : 0.0;
      }
      #pragma unroll
      for (int c = 0; c < CG; ++c) {
         #pragma unroll
         for (int k = 0; k < NR; ++k) {
            double sum = 0.0;
            #pragma unroll
            for (int q = 0; q < 4; ++q) sum += l[c][q] * xv[q][k];
            #pragma unroll
            for (int off = 16; off > 0; off >>= 1) sum += cx.shfl_xor(sum, off);
            /* every lane holds the total; lane (c NR + k) mod 32 adds it */
            if (lane == ((c * NR + k) & 31) && c0 + cb + c < w) cx.atomic_add(&acc[(size_t)(c0 + cb + c) * NR + k], sum);
         }
      }
   }
}

/* ---- backward, T: x(block) = L(block, block)^-T (x(block) - accumulator); the accumulator is cleared ---- */
template <int NR, int NRT, bool POSDEF, int HB, class Ctx>
SW_FN void bwd_wide_T_h(Ctx& cx, const SolveFront& f, int step, double* x, double* pb, double* smem) {
   const int b = sw_bwd_block(f, step);
   if (b < 0) return;
   constexpr int XLD = sw_xld<NR>();
   constexpr int LLD = HB + 4;
   double* vs = smem;
   double* lkk = smem + (size_t)HB * XLD;
   double* ls = lkk + (size_t)SSB * SW_LK;          // (tensor-core variant only)
   const int kb = b * SWB;
   const int w = sw_min(SWB, f.nelim - kb);
   const int t = cx.tid(), lane = t & 31, warp = t >> 5;
   const bool acol = t < w;
   const int g = acol ? f.perm[kb + t] - 1 : -1;
   {  /* pb: what the G kernel accumulated for this block; cleared for the next step.  All loads before the stores
       * (x and pb are distinct, but the compiler cannot know it and would wait for every store) */
      double xv[NR], pv[NR];
      #pragma unroll
      for (int k = 0; k < NR; ++k) { xv[k] = acol ? x[(size_t)g * NRT + k] : 0.0; pv[k] = pb[(size_t)t * NRT + k]; }
      #pragma unroll
      for (int k = 0; k < NR; ++k) {
         pb[(size_t)t * NRT + k] = 0.0;
         if (t < HB) vs[(size_t)t * XLD + k] = acol ? xv[k] - pv[k] : 0.0;
      }
   }
   const size_t ldl = (size_t)f.ldl;
   const double* Lcol = f.L + (size_t)kb + (size_t)(kb + t) * ldl;     // column kb+t of the block, from row kb
   const int jbl = ((w - 1) / SSB) * SSB;           // last sub-block: the first one to be solved
   constexpr int TT = HB < SW_TT ? HB : SW_TT, IPT = SSB * SSB / TT;      // inverse diagonal blocks: see fwd_wide_T_h
   const bool use_inv = f.Linv != nullptr && *f.linv_bad == 0;
   const double* inv_blk = use_inv ? f.Linv + (size_t)(kb / SSB) * (SSB * SSB) + (size_t)t * IPT : nullptr;
   double icur[IPT], inxt[IPT];
   #pragma unroll
   for (int q = 0; q < IPT; ++q) icur[q] = use_inv ? inv_blk[(size_t)(jbl / SSB) * (SSB * SSB) + q] : 0.0;
   double cur[SSB], nxt[SSB];
   #pragma unroll
   for (int i = 0; i < SSB; ++i) cur[i] = (acol && t < jbl + SSB && jbl + i < w) ? Lcol[jbl + i] : 0.0;
   constexpr int NWARP = (HB < SW_TT ? HB : SW_TT) / 32;      // the CTA has min(HB, 256) threads
   constexpr int NRW = (NR + NWARP - 1) / NWARP;
   for (int jb = jbl; jb >= 0; jb -= SSB) {
      const int wd = sw_min(SSB, w - jb);
      if (use_inv) {
         #pragma unroll
         for (int q = 0; q < IPT; ++q) { const int e = t * IPT + q; lkk[(e / SSB) * SW_LK + e % SSB] = icur[q]; }
      } else if (t >= jb && t < jb + SSB) {       // the columns of the sub-block publish its diagonal block
         const int j = t - jb;
         #pragma unroll
         for (int i = 0; i < SSB; ++i) lkk[i * SW_LK + j] = (i < wd && j < wd && i >= j) ? cur[i] : 0.0;
      }
      const int jp = jb - SSB;                    // rows of the previous sub-block, for the columns up to its end
      if (use_inv && jp >= 0) {
         #pragma unroll
         for (int q = 0; q < IPT; ++q) inxt[q] = inv_blk[(size_t)(jp / SSB) * (SSB * SSB) + q];
      }
      if (sw_T_mma<NR>()) {
         #pragma unroll
         for (int i = 0; i < SSB; ++i) if (t < HB) ls[(size_t)i * LLD + t] = (acol && t < jb) ? cur[i] : 0.0;
      }
      if (jp >= 0) {
         #pragma unroll
         for (int i = 0; i < SSB; ++i) nxt[i] = (acol && t < jp + SSB) ? Lcol[jp + i] : 0.0;
      } else {
         #pragma unroll
         for (int i = 0; i < SSB; ++i) nxt[i] = 0.0;
      }
      cx.sync();
      {  /* transposed substitution: lanes are the columns of the sub-block */
         double v[NRW];
         if (use_inv) {                            // z = Linv_bb^T r
            #pragma unroll
            for (int q = 0; q < NRW; ++q) {
               const int k = warp * NRW + q;
               double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
               if (k < NR) {
                  #pragma unroll
                  for (int i = 0; i < SSB; i += 4) {
                     s0 += lkk[i * SW_LK + lane] * vs[(size_t)(jb + i) * XLD + k];
                     s1 += lkk[(i + 1) * SW_LK + lane] * vs[(size_t)(jb + i + 1) * XLD + k];
                     s2 += lkk[(i + 2) * SW_LK + lane] * vs[(size_t)(jb + i + 2) * XLD + k];
                     s3 += lkk[(i + 3) * SW_LK + lane] * vs[(size_t)(jb + i + 3) * XLD + k];
                  }
               }
               v[q] = (s0 + s1) + (s2 + s3);
            }
            cx.sync_warp();
         } else {
         #pragma unroll
         for (int q = 0; q < NRW; ++q) { const int k = warp * NRW + q; v[q] = (k < NR) ? vs[(size_t)(jb + lane) * XLD + k] : 0.0; }
         #pragma unroll 4
         for (int j = wd - 1; j >= 0; --j) {
            const double l = (lane < j) ? lkk[j * SW_LK + lane] : 0.0;
            const double dj = POSDEF ? lkk[j * SW_LK + j] : 1.0;
            #pragma unroll
            for (int q = 0; q < NRW; ++q) {
               double zj = cx.shfl(v[q], j);
               if (POSDEF) { zj /= dj; if (lane == j) v[q] = zj; }
               v[q] -= l * zj;
            }
         }
         }
         #pragma unroll
         for (int q = 0; q < NRW; ++q) { const int k = warp * NRW + q; if (k < NR) vs[(size_t)(jb + lane) * XLD + k] = v[q]; }
      }
      cx.sync();
      if (sw_T_mma<NR>()) {
         if (jb > 0) sw_mma_update<(NR >= 16 ? NR : 16), LLD>(cx, vs, ls, jb, 0, jb);
         cx.sync();
      } else if (acol && t < jb) {                // columns of the block left of the sub-block
         #pragma unroll
         for (int k = 0; k < NR; ++k) {
            double s = vs[(size_t)t * XLD + k];
            #pragma unroll
            for (int i = 0; i < SSB; ++i) s -= cur[i] * vs[(size_t)(jb + i) * XLD + k];
            vs[(size_t)t * XLD + k] = s;
         }
      }
      #pragma unroll
      for (int i = 0; i < SSB; ++i) cur[i] = nxt[i];
      #pragma unroll
      for (int q = 0; q < IPT; ++q) icur[q] = inxt[q];
   }
   cx.sync();
   if (acol) {
      #pragma unroll
      for (int k = 0; k < NR; ++k) x[(size_t)g * NRT + k] = vs[(size_t)t * XLD + k];
   }
}
template <int NR, int NRT, bool POSDEF, class Ctx>
SW_FN void bwd_wide_T(Ctx& cx, const SolveFront& f, int step, double* x, double* pb, double* smem) {
   bwd_wide_T_h<NR, NRT, POSDEF, SWB>(cx, f, step, x, pb, smem);
}

} // namespace b200
