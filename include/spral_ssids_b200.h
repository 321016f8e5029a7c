/* spral_ssids_b200.h -- C ABI of the B200-native SSIDS numeric engine.
 *
 * This header is the drop-in boundary.  Every entry point below names the
 * reference interface it replaces (paths relative to the ralna/spral tree).
 *
 * Layer 1 (subtree plug-in ABI): the same shape as the reference CPU part's
 *   ABI `spral_ssids_cpu_*` (src/ssids/cpu/subtree.f90:42-184, definitions in
 *   src/ssids/cpu/SymbolicSubtree.cxx:11-27, src/ssids/cpu/NumericSubtree.cxx:29-262)
 *   with a leading `device` argument in the symbolic constructor, as taken by
 *   `construct_gpu_symbolic_subtree` (src/ssids/gpu/subtree.f90:76-90).
 *   A Fortran `type, extends(symbolic_subtree_base)` binds to it 1:1
 *   (see INTEGRATION.md and spral_b200/fortran/gpu_subtree_b200.f90).
 *
 * Layer 0 (device bring-up): replaces driver/cuda_helper_gpu.f90:9-43.
 *
 * Layer 2 (host analyse helpers): C restatement of the Fortran-only analyse
 *   pieces that produce the hot path's inputs (src/core_analyse.f90,
 *   src/ssids/anal.F90); used by the C implementation of the spral_ssids.h
 *   subset (Layer 3) and by the tests.
 *
 * Layer 3: the factor/solve subset of include/spral_ssids.h, same prototypes,
 *   exported under the same names by libspral_ssids_b200.so.
 *
 * All index arrays hold Fortran 1-based VALUES (as in the reference) unless
 * stated.  No exceptions cross this boundary; errors are reported through
 * `flag` fields using the reference's codes (src/ssids/datatypes.f90:25-59).
 */
#ifndef SPRAL_SSIDS_B200_H
#define SPRAL_SSIDS_B200_H

#include <stdbool.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------ */
/* Shared types                                                              */
/* ------------------------------------------------------------------------ */

/* Interoperable options; identical layout to `cpu_factor_options`
 * (src/ssids/cpu/cpu_iface.hxx:24-34 == src/ssids/cpu/cpu_iface.f90:21-31). */
struct spral_ssids_b200_options {
   int print_level;
   bool action;
   double small;
   double u;
   double multiplier;
   int64_t small_subtree_threshold;
   int cpu_block_size;
   int pivot_method;          /* 1 APP_AGGRESSIVE, 2 APP_BLOCK, 3 TPP */
   int failed_pivot_method;   /* 1 TPP, 2 PASS */
};

/* Interoperable statistics; identical layout to `ThreadStats`
 * (src/ssids/cpu/ThreadStats.hxx:48-62 == cpu_factor_stats, cpu_iface.f90:39-51).
 * cuda_error is appended (reference: inform%cuda_error, gpu/subtree.f90:455-459). */
struct spral_ssids_b200_stats {
   int flag;
   int num_delay;
   int64_t num_factor;
   int64_t num_flops;
   int num_neg;
   int num_two;
   int num_zero;
   int maxfront;
   int maxsupernode;
   int not_first_pass;
   int not_second_pass;
   int cuda_error;            /* raw cudaError_t when flag == -51 */
};

/* Flag values (src/ssids/datatypes.f90:25-59, src/ssids/cpu/ThreadStats.hxx:17-25) */
#define SPRAL_SSIDS_SUCCESS                 0
#define SPRAL_SSIDS_ERROR_SINGULAR        (-5)
#define SPRAL_SSIDS_ERROR_NOT_POS_DEF     (-6)
#define SPRAL_SSIDS_ERROR_ALLOCATION     (-50)
#define SPRAL_SSIDS_ERROR_CUDA_UNKNOWN   (-51)
#define SPRAL_SSIDS_WARNING_FACT_SINGULAR   7

/* C image of the Fortran `contrib_type` (src/ssids/contrib.f90:19-33): a
 * contribution block handed from a child part to its parent part.
 * `device` is an extension: -1 when val/rlist/delay_* are host pointers (the
 * reference's convention), otherwise the CUDA device ordinal owning them so a
 * parent part on another B200 can pull them over NVLink (P2P copy). */
struct spral_ssids_b200_contrib {
   volatile int ready;        /* set last by the producer (fkeep.F90:163-169) */
   int n;                     /* order of the block */
   const double* val;         /* n x n, lower triangle valid, ld = ldval */
   int ldval;
   const int* rlist;          /* n global (pivot order, 1-based) row indices */
   int ndelay;
   const int* delay_perm;     /* ndelay */
   const double* delay_val;   /* lddelay x ndelay */
   int lddelay;
   int owner;                 /* 0 = reference CPU part, 1 = this engine */
   bool posdef;
   void* owner_ptr;           /* numeric subtree that owns the memory */
   int device;                /* -1 host, >=0 CUDA device of the pointers */
};

/* ------------------------------------------------------------------------ */
/* Layer 0: device bring-up (replaces driver/cuda_helper_gpu.f90:9-43)       */
/* ------------------------------------------------------------------------ */

/* Counts CUDA devices and warms each one up (context + first allocation).
 * Returns 0 or the raw cudaError_t; *cnt = 0 when no device is usable. */
int spral_ssids_b200_cuda_init(int* cnt);

/* ------------------------------------------------------------------------ */
/* Layer 1: subtree plug-in ABI                                              */
/* ------------------------------------------------------------------------ */

/* replaces spral_ssids_cpu_create_symbolic_subtree (SymbolicSubtree.cxx:11-22)
 * / construct_gpu_symbolic_subtree (gpu/subtree.f90:76-160).
 * The part consists of nodes sa..en-1; contrib_idx[i] is the node (1-based,
 * global) that incoming contribution i assembles into.  sptr/sparent/rptr/
 * rlist/nptr/nlist are copied to the device; the caller may free them. */
void* spral_ssids_gpu_create_symbolic_subtree(
      int device, int n, int sa, int en, const int* sptr, const int* sparent,
      const int64_t* rptr, const int* rlist, const int64_t* nptr,
      const int64_t* nlist, int ncontrib, const int* contrib_idx,
      const struct spral_ssids_b200_options* options);

/* replaces spral_ssids_cpu_destroy_symbolic_subtree (SymbolicSubtree.cxx:24-27) */
void spral_ssids_gpu_destroy_symbolic_subtree(void* symbolic_subtree);

/* replaces spral_ssids_cpu_create_num_subtree_dbl (NumericSubtree.cxx:29-59)
 * / gpu_symbolic_subtree%factor (gpu/subtree.f90:304-460): numeric
 * factorisation of the part.  aval/scaling may be host or device pointers
 * (scaling may be NULL).  child_contrib[i] points to a
 * spral_ssids_b200_contrib.  Always returns an object (destroy it); errors
 * are in stats->flag. */
void* spral_ssids_gpu_create_num_subtree_dbl(
      bool posdef, const void* symbolic_subtree, const double* aval,
      const double* scaling, void** child_contrib,
      const struct spral_ssids_b200_options* options,
      struct spral_ssids_b200_stats* stats);

/* replaces spral_ssids_cpu_destroy_num_subtree_dbl (NumericSubtree.cxx:61-72) */
void spral_ssids_gpu_destroy_num_subtree_dbl(bool posdef, void* numeric_subtree);

/* replace spral_ssids_cpu_subtree_solve_{fwd,diag,diag_bwd,bwd}_dbl
 * (NumericSubtree.cxx:75-178) / gpu solve_* (gpu/subtree.f90:555-700).
 * x is ldx x nrhs in pivot order, updated in place; host or device pointer.
 * Return 0 or a negative flag. */
int spral_ssids_gpu_subtree_solve_fwd_dbl(bool posdef, const void* numeric_subtree,
      int nrhs, double* x, int ldx);
int spral_ssids_gpu_subtree_solve_diag_dbl(bool posdef, const void* numeric_subtree,
      int nrhs, double* x, int ldx);
int spral_ssids_gpu_subtree_solve_diag_bwd_dbl(bool posdef, const void* numeric_subtree,
      int nrhs, double* x, int ldx);
int spral_ssids_gpu_subtree_solve_bwd_dbl(bool posdef, const void* numeric_subtree,
      int nrhs, double* x, int ldx);

/* replaces spral_ssids_cpu_subtree_enquire_dbl (NumericSubtree.cxx:181-200):
 * posdef: d gets the diagonal of L per eliminated column; indef: piv_order
 * (may be NULL) and d (2 per pivot, may be NULL) in the format of
 * NumericSubtree.hxx:424-470.  Host pointers. */
void spral_ssids_gpu_subtree_enquire_dbl(bool posdef, const void* numeric_subtree,
      int* piv_order, double* d);

/* replaces spral_ssids_cpu_subtree_alter_dbl (NumericSubtree.cxx:203-215) */
void spral_ssids_gpu_subtree_alter_dbl(bool posdef, void* numeric_subtree,
      const double* d);

/* replaces spral_ssids_cpu_subtree_get_contrib_dbl (NumericSubtree.cxx:218-245)
 * / gpu get_contrib (gpu/subtree.f90:522-538): HOST copies of the root's
 * contribution block (reference convention). */
void spral_ssids_gpu_subtree_get_contrib_dbl(bool posdef, void* numeric_subtree,
      int* n, const double** val, int* ldval, const int** rlist, int* ndelay,
      const int** delay_perm, const double** delay_val, int* lddelay);

/* Extension: device-resident view of the same contribution block (no D2H);
 * *device receives the owning CUDA device. */
void spral_ssids_gpu_subtree_get_contrib_device_dbl(bool posdef, void* numeric_subtree,
      int* n, const double** val, int* ldval, const int** rlist, int* ndelay,
      const int** delay_perm, const double** delay_val, int* lddelay, int* device);

/* replaces spral_ssids_cpu_subtree_free_contrib_dbl (NumericSubtree.cxx:248-262) */
void spral_ssids_gpu_subtree_free_contrib_dbl(bool posdef, void* numeric_subtree);

/* C restatement of spral_ssids_contrib_get_data / spral_ssids_contrib_free_dbl
 * (src/ssids/contrib.f90:37-76, src/ssids/contrib_free.f90:34-48) acting on
 * struct spral_ssids_b200_contrib; only used when no Fortran host exists. */
void spral_ssids_b200_contrib_fill(struct spral_ssids_b200_contrib* c,
      bool posdef, void* numeric_subtree, bool device_resident);

/* One process per GPU (torchrun): cross-process hand-over of a part's root
 * contribution block, replacing transfer_contrib's D2H copy
 * (src/ssids/gpu/factor.f90:155-221).  The producer packs
 * [val n*n | delay_val (ndelay+n)*ndelay | delay_perm ndelay ints] into one
 * device block it keeps owning and returns its 64-byte CUDA IPC handle; the
 * consumer pulls the block over NVLink into its own memory (`device` = the
 * consumer's CUDA device; the calls may come from any host thread).  Return 0
 * or the raw cudaError_t. */
int spral_ssids_gpu_subtree_export_contrib_ipc(void* numeric_subtree, unsigned char* handle,
      int* n, int* ndelay, int64_t* bytes, void** device_block);
int spral_ssids_b200_ipc_pull(int device, const unsigned char* handle, int64_t bytes, void* dst);
int spral_ssids_b200_copy_to_host(void* dst, const void* src, int64_t bytes);
void* spral_ssids_b200_device_alloc(int device, int64_t bytes);
void spral_ssids_b200_device_free(void* p);

/* Introspection used by the parity tests ("bit-exact extend-add index maps
 * and tree scheduling"): copies back the device-side maps built by the
 * symbolic constructor.  rlist_direct has rptr[en]-rptr[sa] entries
 * (gpu/subtree.f90:204-234); level_ptr has *num_levels+1 entries, level_list
 * en-sa entries of 1-based local node indices, level 1 = deepest
 * (gpu/factor.f90:824-879).  Any pointer may be NULL. */
void spral_ssids_gpu_symbolic_get_maps(const void* symbolic_subtree,
      int* rlist_direct, int* num_levels, int* level_ptr, int* level_list);

/* Measurements of the factorisation that built this object (bench.py):
 * [0] host wall time of the call (ms), [1] device time (ms, CUDA events on the
 * factorisation's stream), [2] ms spent in the Schur-complement (DMMA) launches,
 * [3] their algorithmic flops, [4] their count ([2]-[4] only with profiling on),
 * [5] host ms spent in the per-level synchronisations, [6] kernels launched. */
void spral_ssids_gpu_subtree_get_timings(const void* numeric_subtree, double* ms, int n);

/* Switches event-bracketing of the Schur-complement launches on or off. */
void spral_ssids_b200_set_profile(int on);

/* Measured FP64 tensor-pipe (DMMA) peak of `device` in TFLOP/s: a register
 * resident mma.sync.m8n8k4.f64 issue loop; the roofline denominator. */
double spral_ssids_b200_fp64_peak_tflops(int device);

/* ------------------------------------------------------------------------ */
/* Layer 2: host analyse helpers (C restatement of Fortran-only inputs)      */
/* ------------------------------------------------------------------------ */

struct spral_ssids_b200_analysis;   /* opaque: holds sptr,sparent,rptr,rlist,nptr,nlist,parts */

/* METIS NodeND ordering of a lower-triangular CSC pattern (1-based ptr/row);
 * restates metis_order (src/metis5_wrapper.F90:102-182).  order[i] = pivot
 * position (1-based) of variable i+1.  Returns 0 or a negative flag. */
int spral_ssids_b200_metis_order(int n, const int64_t* ptr, const int* row, int* order);

/* Symbolic analysis for a given order: restates expand_pattern
 * (anal.F90:42-85), basic_analyse (core_analyse.f90:38-156), build_map
 * (anal.F90:1137-1239) and find_subtree_partition (anal.F90:289-464) for a
 * topology of one NUMA region with `ngpu` GPUs.  order is updated to the
 * final (postordered, supernode-contiguous) pivot order. */
struct spral_ssids_b200_analysis* spral_ssids_b200_analyse(
      int n, const int64_t* ptr, const int* row, int* order, int nemin,
      int ngpu, int64_t min_gpu_work, float max_load_inbalance, float gpu_perf_coeff,
      int* flag);
void spral_ssids_b200_analysis_free(struct spral_ssids_b200_analysis*);

/* Replaces the subtree partition (akeep%part, exec_loc, contrib_ptr / contrib_idx / contrib_dest of anal.F90:418-454)
 * by a caller-chosen one: part[0..nparts] = 1-based first nodes (part[nparts] = nnodes + 1), every part connected and
 * with its last node as only exit.  Returns 0, -1 if invalid. */
int spral_ssids_b200_analysis_set_partition(struct spral_ssids_b200_analysis*, int nparts, const int* part,
      const int* exec_loc);

/* Accessors (pointers stay owned by the analysis object). */
struct spral_ssids_b200_analysis_view {
   int n, nnodes, nparts;
   const int* sptr;          /* nnodes+1 */
   const int* sparent;       /* nnodes */
   const int64_t* rptr;      /* nnodes+1 */
   const int* rlist;         /* rptr[nnodes]-1 */
   const int64_t* nptr;      /* nnodes+1 */
   const int64_t* nlist;     /* 2*nz */
   const int* invp;          /* n: invp[p-1] = variable eliminated p-th */
   const int* part;          /* nparts+1 */
   const int* exec_loc;      /* nparts */
   const int* contrib_ptr;   /* nparts+3 */
   const int* contrib_idx;   /* nparts */
   const int* contrib_dest;  /* nparts */
   int64_t num_factor, num_flops;
   int maxfront, maxsupernode, maxdepth;
};
void spral_ssids_b200_analysis_get(const struct spral_ssids_b200_analysis*,
      struct spral_ssids_b200_analysis_view* view);

/* ------------------------------------------------------------------------ */
/* Layer 2b: scaling pre-processing on the host (options%scaling of          */
/* ssids_factor, src/ssids/ssids.f90:861-1028); scaling.cpp                  */
/* ------------------------------------------------------------------------ */

/* hungarian_scale_sym (src/scaling.f90:134-170, hungarian_wrapper :597-801): MC64-type
 * matching-based scaling of the symmetric matrix given by its lower triangle (1-based
 * CSC).  scaling[n]; match[n] (may be NULL): 1-based column matched to row i, negative
 * for the variables outside the matching of a structurally singular matrix.  Returns
 * inform%flag: 0, 1 (WARNING_SINGULAR, only with scale_if_singular != 0), -2
 * (ERROR_SINGULAR: scaling = 1).  *matched = size of the matching. */
int spral_ssids_b200_hungarian_scale_sym(int n, const int64_t* ptr, const int* row, const double* val,
      double* scaling, int* match, int scale_if_singular, int* matched);

/* auction_scale_sym (src/scaling.f90:269-309; options%scaling = 2): scaling from an approximate matching
 * found by the auction algorithm.  opts: NULL or {int max_iterations; int max_unchanged[3]; float
 * min_proportion[3]; float eps_initial;} as auction_options (:33-38).  match[n] (may be NULL): 1-based
 * column matched to row i, 0 if none.  Returns 0. */
int spral_ssids_b200_auction_scale_sym(int n, const int64_t* ptr, const int* row, const double* val,
      double* scaling, int* match, const void* opts, int* matched, int* iterations);

/* match_order_metis (src/match_order.f90:51-208; options%ordering = 2): matching-based ordering.  order[n]:
 * 1-based pivot position of every variable, the two variables of a matched 2-cycle consecutive;
 * scaling[n]: the matching-based scaling (options%scaling = 3 uses it at factor time).  Returns 0, 1
 * (structurally singular) or a negative flag. */
int spral_ssids_b200_match_order_metis(int n, const int64_t* ptr, const int* row, const double* val,
      int* order, double* scaling);

/* equilib_scale_sym (src/scaling.f90:381-416, inf_norm_equilib_sym :480-521);
 * equilib_options defaults: max_iterations = 10, tol = 1e-8.  Returns 0. */
int spral_ssids_b200_equilib_scale_sym(int n, const int64_t* ptr, const int* row, const double* val,
      double* scaling, int max_iterations, double tol, int* iterations);

#ifdef __cplusplus
}
#endif
#endif /* SPRAL_SSIDS_B200_H */
