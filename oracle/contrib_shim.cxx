/* TEST INFRASTRUCTURE (oracle build) -- not part of the product path.
 *
 * C restatement of the two Fortran routines the reference CPU engine calls to
 * read and release a contribution block handed over from another part:
 *   spral_ssids_contrib_get_data   reference: src/ssids/contrib.f90:37-76
 *   spral_ssids_contrib_free_dbl   reference: src/ssids/contrib_free.f90:17-48
 * (prototypes: src/ssids/contrib.h:16-21).  They act on the C image of the
 * Fortran contrib_type, struct spral_ssids_b200_contrib.
 *
 * Also holds oracle_factor(): the reference numeric constructor must run
 * inside an OpenMP parallel region's `single` so that it sizes its workspaces
 * for all threads (reference: src/ssids/cpu/NumericSubtree.hxx:76-81; the
 * Fortran caller does this in src/ssids/fkeep.F90:104-179).
 */
#include <omp.h>
#include <sched.h>
#include <cstdio>
#include <cstdlib>
#include "spral_ssids_b200.h"

extern "C" {
void spral_ssids_cpu_subtree_free_contrib_dbl(bool posdef, void* subtree);
void* spral_ssids_cpu_create_num_subtree_dbl(bool posdef, const void* symb,
      const double* aval, const double* scaling, void** child_contrib,
      const void* options, void* stats);
void oracle_blas_single_thread(void);

static void (*gpu_free_contrib_hook)(bool, void*) = nullptr;

/* Lets the harness route owner==1 blocks back to the engine that owns them. */
void oracle_set_gpu_free_contrib(void (*fn)(bool, void*)) { gpu_free_contrib_hook = fn; }

/* follows contrib.f90:37-76 */
void spral_ssids_contrib_get_data(const void* const contrib, int* const n,
      const double** const val, int* const ldval, const int** const rlist,
      int* const ndelay, const int** const delay_perm,
      const double** const delay_val, int* const lddelay) {
   if(!contrib) return;
   auto* c = static_cast<const spral_ssids_b200_contrib*>(contrib);
   while(!c->ready) {
      #pragma omp taskyield
      sched_yield();
   }
   *n = c->n;
   *val = c->val;
   *ldval = c->ldval;
   *rlist = c->rlist;
   *ndelay = c->ndelay;
   if(c->delay_val) { *delay_perm = c->delay_perm; *delay_val = c->delay_val; }
   else             { *delay_perm = nullptr;       *delay_val = nullptr; }
   *lddelay = c->lddelay;
}

/* follows contrib_free.f90:17-31,34-48 */
void spral_ssids_contrib_free_dbl(void* const contrib) {
   if(!contrib) return;
   auto* c = static_cast<spral_ssids_b200_contrib*>(contrib);
   switch(c->owner) {
   case 0: spral_ssids_cpu_subtree_free_contrib_dbl(c->posdef, c->owner_ptr); break;
   case 2: break; /* memory owned by the test harness (a block received from another process) */
   case 1:
      if(gpu_free_contrib_hook) { gpu_free_contrib_hook(c->posdef, c->owner_ptr); break; }
      /* fallthrough */
   default:
      fprintf(stderr, "Unrecognised contrib owner %d\n", c->owner);
      abort();
   }
}

/* Runs the reference numeric constructor with `nthreads` OpenMP threads. */
void* oracle_factor(bool posdef, const void* symb, const double* aval,
      const double* scaling, void** child_contrib, const void* options,
      void* stats, int nthreads) {
   void* result = nullptr;
   oracle_blas_single_thread();
   if(nthreads < 1) nthreads = omp_get_max_threads();
   #pragma omp parallel num_threads(nthreads) default(shared)
   {
      #pragma omp single
      result = spral_ssids_cpu_create_num_subtree_dbl(posdef, symb, aval,
            scaling, child_contrib, options, stats);
   }
   return result;
}

int oracle_cancellation_enabled(void) { return omp_get_cancellation(); }
int oracle_max_threads(void) { return omp_get_max_threads(); }
} /* extern "C" */
