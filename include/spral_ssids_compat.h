/* spral_ssids_compat.h -- the factor/solve subset of SPRAL's C interface
 * (reference: include/spral_ssids.h:15-127, implemented there by the Fortran
 * shims of interfaces/C/ssids.f90) served by the B200 engine.
 *
 * Same function names, argument lists and struct layouts as the reference
 * header, so a C program written against spral_ssids.h links against
 * libspral_ssids_b200.so unchanged.  Implemented in
 * spral_b200/csrc/ssids_capi.cpp on top of the subtree ABI
 * (spral_ssids_b200.h) and the restated analyse phase.
 *
 * All 15 functions of the reference header are provided (analyse, analyse_coord,
 * analyse_ptr32, analyse_topology -- the topology argument is ignored --, factor,
 * factor_ptr32, solve, solve1, enquire_posdef, enquire_indef, alter, free*), and
 * options.scaling = 0 (user vector), 1 (Hungarian), 2 (auction), 3 (from the
 * matching-based ordering), 4 (equilibration); options.ordering = 0, 1, 2.
 * Accepted and ignored: options.pivot_method / failed_pivot_method (the engine has
 * one pivoting strategy: APP inside 32 x 32 blocks with an a-posteriori threshold
 * test, failed columns retried in later passes, then delayed), nstream,
 * gpu_perf_coeff, the print units; inform.stat is always 0.
 */
#ifndef SPRAL_SSIDS_COMPAT_H
#define SPRAL_SSIDS_COMPAT_H
#include <stdbool.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* layout of the reference struct (include/spral_ssids.h:15-36) */
struct spral_ssids_options {
   int array_base;            /* 0 or 1: base of ptr/row/order */
   int print_level, unit_diagnostics, unit_error, unit_warning;
   int ordering;              /* 0 user, 1 METIS */
   int nemin;
   bool ignore_numa, use_gpu;
   int64_t min_gpu_work;
   float max_load_inbalance, gpu_perf_coeff;
   int scaling;
   int64_t small_subtree_threshold;
   int cpu_block_size;
   bool action;
   int pivot_method;
   double small, u;
   char unused[80];
};

/* layout of the reference struct (include/spral_ssids.h:38-57) */
struct spral_ssids_inform {
   int flag, matrix_dup, matrix_missing_diag, matrix_outrange, matrix_rank;
   int maxdepth, maxfront, num_delay;
   int64_t num_factor, num_flops;
   int num_neg, num_sup, num_two, stat, cuda_error, cublas_error, maxsupernode;
   char unused[76];
};

/* layout of the reference struct (include/spral_ssids.h:59-63) */
struct spral_numa_region {
   int nproc;
   int ngpu;
   int* gpus;
};

void spral_ssids_default_options(struct spral_ssids_options* options);
void spral_ssids_analyse(bool check, int n, int* order, const int64_t* ptr, const int* row,
      const double* val, void** akeep, const struct spral_ssids_options* options,
      struct spral_ssids_inform* inform);
/* the topology is accepted and ignored: this interface drives ONE device (the current one); the
 * subtree partition over several GPUs is the one-process-per-GPU driver's (spral_b200/dist.py) */
void spral_ssids_analyse_topology(bool check, int n, int* order, const int64_t* ptr, const int* row,
      const double* val, void** akeep, const struct spral_ssids_options* options,
      struct spral_ssids_inform* inform, int nregions, const struct spral_numa_region* regions);
void spral_ssids_analyse_ptr32(bool check, int n, int* order, const int* ptr, const int* row,
      const double* val, void** akeep, const struct spral_ssids_options* options,
      struct spral_ssids_inform* inform);
void spral_ssids_analyse_coord(int n, int* order, int64_t ne, const int* row, const int* col,
      const double* val, void** akeep, const struct spral_ssids_options* options,
      struct spral_ssids_inform* inform);
void spral_ssids_factor(bool posdef, const int64_t* ptr, const int* row, const double* val,
      double* scale, void* akeep, void** fkeep, const struct spral_ssids_options* options,
      struct spral_ssids_inform* inform);
void spral_ssids_factor_ptr32(bool posdef, const int* ptr, const int* row, const double* val,
      double* scale, void* akeep, void** fkeep, const struct spral_ssids_options* options,
      struct spral_ssids_inform* inform);
void spral_ssids_solve1(int job, double* x1, void* akeep, void* fkeep,
      const struct spral_ssids_options* options, struct spral_ssids_inform* inform);
void spral_ssids_solve(int job, int nrhs, double* x, int ldx, void* akeep, void* fkeep,
      const struct spral_ssids_options* options, struct spral_ssids_inform* inform);
int spral_ssids_free_akeep(void** akeep);
int spral_ssids_free_fkeep(void** fkeep);
int spral_ssids_free(void** akeep, void** fkeep);
void spral_ssids_enquire_posdef(const void* akeep, const void* fkeep,
      const struct spral_ssids_options* options, struct spral_ssids_inform* inform, double* d);
void spral_ssids_enquire_indef(const void* akeep, const void* fkeep,
      const struct spral_ssids_options* options, struct spral_ssids_inform* inform,
      int* piv_order, double* d);
void spral_ssids_alter(const double* d, const void* akeep, void* fkeep,
      const struct spral_ssids_options* options, struct spral_ssids_inform* inform);

#ifdef __cplusplus
}
#endif
#endif
