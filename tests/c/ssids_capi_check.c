/* C client of the spral_ssids.h-compatible interface served by the B200 engine.
 * The first case is the workflow of the reference's C example (5x5 indefinite
 * matrix with solution 1..5, examples/C/ssids.c); the others exercise data
 * checking, 0-based input, ptr32, several right-hand sides and the posdef path.
 * Prints "CAPI OK" and returns 0 when every check holds. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "spral_ssids_compat.h"

static int fails = 0;
#define CHECK(cond, msg) do { if (!(cond)) { printf("FAIL: %s (line %d)\n", msg, __LINE__); fails++; } } while (0)

static void case_example(void) {
   struct spral_ssids_options opt;
   struct spral_ssids_inform inf;
   void *akeep = NULL, *fkeep = NULL;
   spral_ssids_default_options(&opt);
   opt.array_base = 1;
   int n = 5;
   int64_t ptr[] = {1, 3, 6, 8, 9, 10};
   int row[] = {1, 2, 2, 3, 5, 3, 4, 4, 5};
   double val[] = {2.0, 1.0, 4.0, 1.0, 1.0, 3.0, 2.0, -1.0, 2.0};
   double x[] = {4.0, 17.0, 19.0, 2.0, 12.0};
   spral_ssids_analyse(true, n, NULL, ptr, row, NULL, &akeep, &opt, &inf);
   CHECK(inf.flag == 0, "analyse flag");
   CHECK(inf.num_factor == 15 && inf.num_flops == 55, "analyse prediction");
   spral_ssids_factor(false, NULL, NULL, val, NULL, akeep, &fkeep, &opt, &inf);
   CHECK(inf.flag == 0, "factor flag");
   CHECK(inf.num_neg == 1 && inf.matrix_rank == 5, "inertia");
   spral_ssids_solve1(0, x, akeep, fkeep, &opt, &inf);
   CHECK(inf.flag == 0, "solve flag");
   for (int i = 0; i < n; ++i) CHECK(fabs(x[i] - (i + 1.0)) < 1e-12, "solution 1..5");
   int piv[5]; double d[10];
   spral_ssids_enquire_indef(akeep, fkeep, &opt, &inf, piv, d);
   CHECK(inf.flag == 0, "enquire flag");
   int seen[6] = {0};
   for (int i = 0; i < n; ++i) { int p = abs(piv[i]); CHECK(p >= 1 && p <= 5 && !seen[p], "piv_order is a permutation"); if (p >= 1 && p <= 5) seen[p] = 1; }
   printf("solution: %g %g %g %g %g  piv_order: %d %d %d %d %d\n", x[0], x[1], x[2], x[3], x[4],
          piv[0], piv[1], piv[2], piv[3], piv[4]);
   /* job 1 + job 4 == job 0; alter: doubling D^-1 doubles the solution */
   double y[] = {4.0, 17.0, 19.0, 2.0, 12.0};
   spral_ssids_solve1(1, y, akeep, fkeep, &opt, &inf);
   spral_ssids_solve1(4, y, akeep, fkeep, &opt, &inf);
   for (int i = 0; i < n; ++i) CHECK(fabs(y[i] - (i + 1.0)) < 1e-12, "job 1 then 4");
   for (int i = 0; i < 10; ++i) d[i] *= 2.0;
   spral_ssids_alter(d, akeep, fkeep, &opt, &inf);
   double z[] = {4.0, 17.0, 19.0, 2.0, 12.0};
   spral_ssids_solve1(0, z, akeep, fkeep, &opt, &inf);
   for (int i = 0; i < n; ++i) CHECK(fabs(z[i] - 2.0 * (i + 1.0)) < 1e-11, "alter");
   double dd[5];
   spral_ssids_enquire_posdef(akeep, fkeep, &opt, &inf, dd);
   CHECK(inf.flag == -13, "enquire_posdef on an indefinite factorisation -> NOT_LLT");
   spral_ssids_solve1(7, z, akeep, fkeep, &opt, &inf);
   CHECK(inf.flag == -11, "job out of range");
   CHECK(spral_ssids_free(&akeep, &fkeep) == 0 && akeep == NULL && fkeep == NULL, "free");
}

/* same matrix, 0-based, ptr32, entries given in the UPPER triangle, one duplicate
 * (split value) and one out-of-range entry: data checking must repair all of it */
static void case_checking(void) {
   struct spral_ssids_options opt;
   struct spral_ssids_inform inf;
   void *akeep = NULL, *fkeep = NULL;
   spral_ssids_default_options(&opt);
   int n = 5;
   /* columns: 0:{0,1} 1:{1,2,4} 2:{2,3} 3:{3} 4:{4}; column 1's (2,1) entry is given as
    * (1,2) in column 2; diagonal (1,1)=4 is split 1.5+2.5; (9,3) is out of range */
   int ptr[] = {0, 2, 5, 8, 10, 11};
   int row[] = {0, 1,   1, 4, 1,   2, 3, 1,   3, 9,   4};
   double val[] = {2.0, 1.0,   1.5, 1.0, 2.5,   3.0, 2.0, 1.0,   -1.0, 77.0,   2.0};
   double x[2 * 5] = {4.0, 17.0, 19.0, 2.0, 12.0,   8.0, 34.0, 38.0, 4.0, 24.0};
   spral_ssids_analyse_ptr32(true, n, NULL, ptr, row, NULL, &akeep, &opt, &inf);
   CHECK(inf.flag == 3, "warning: duplicates and out-of-range");
   CHECK(inf.matrix_dup == 1 && inf.matrix_outrange == 1, "counts of repaired entries");
   spral_ssids_factor(false, NULL, NULL, val, NULL, akeep, &fkeep, &opt, &inf);
   CHECK(inf.flag >= 0 && inf.num_neg == 1, "factor of the repaired matrix");
   spral_ssids_solve(0, 2, x, 5, akeep, fkeep, &opt, &inf);
   for (int i = 0; i < n; ++i) {
      CHECK(fabs(x[i] - (i + 1.0)) < 1e-12, "rhs 1");
      CHECK(fabs(x[5 + i] - 2.0 * (i + 1.0)) < 1e-12, "rhs 2");
   }
   spral_ssids_free(&akeep, &fkeep);
}

/* 2-D 5-point Laplacian 30x30, positive definite, user ordering = natural, user scaling */
static void case_posdef(void) {
   struct spral_ssids_options opt;
   struct spral_ssids_inform inf;
   void *akeep = NULL, *fkeep = NULL;
   spral_ssids_default_options(&opt);
   opt.array_base = 1;
   int g = 30, n = g * g;
   int64_t *ptr = malloc((n + 1) * sizeof(int64_t));
   int *row = malloc(3 * n * sizeof(int));
   double *val = malloc(3 * n * sizeof(double));
   int64_t ne = 0;
   for (int j = 0; j < n; ++j) {
      ptr[j] = ne + 1;
      row[ne] = j + 1; val[ne++] = 4.0;
      if ((j % g) != g - 1) { row[ne] = j + 2; val[ne++] = -1.0; }
      if (j + g < n) { row[ne] = j + g + 1; val[ne++] = -1.0; }
   }
   ptr[n] = ne + 1;
   double *x = malloc(n * sizeof(double)), *b = calloc(n, sizeof(double)), *scale = malloc(n * sizeof(double));
   for (int j = 0; j < n; ++j)       /* b = A * ones */
      for (int64_t k = ptr[j] - 1; k < ptr[j + 1] - 1; ++k) {
         int i = row[k] - 1;
         b[i] += val[k];
         if (i != j) b[j] += val[k];
      }
   for (int i = 0; i < n; ++i) { x[i] = b[i]; scale[i] = 0.5; }
   spral_ssids_analyse(false, n, NULL, ptr, row, NULL, &akeep, &opt, &inf);
   CHECK(inf.flag == 0 && inf.num_sup > 0, "analyse");
   spral_ssids_factor(true, NULL, NULL, val, scale, akeep, &fkeep, &opt, &inf);
   CHECK(inf.flag == 0 && inf.matrix_rank == n, "posdef factor with user scaling");
   spral_ssids_solve1(0, x, akeep, fkeep, &opt, &inf);
   double err = 0;
   for (int i = 0; i < n; ++i) err = fmax(err, fabs(x[i] - 1.0));
   CHECK(err < 1e-11, "posdef solution");
   double *d = malloc(n * sizeof(double));
   spral_ssids_enquire_posdef(akeep, fkeep, &opt, &inf, d);
   int pos = 1;
   for (int i = 0; i < n; ++i) pos = pos && d[i] > 0;
   CHECK(inf.flag == 0 && pos, "enquire_posdef: positive diagonal of L");
   /* not positive definite -> -6 */
   val[0] = -4.0;
   spral_ssids_factor(true, NULL, NULL, val, NULL, akeep, &fkeep, &opt, &inf);
   CHECK(inf.flag == -6, "not positive definite");
   spral_ssids_free(&akeep, &fkeep);
   free(ptr); free(row); free(val); free(x); free(b); free(scale); free(d);
}

int main(void) {
   case_example();
   case_checking();
   case_posdef();
   if (fails) { printf("CAPI FAILED: %d checks\n", fails); return 1; }
   printf("CAPI OK\n");
   return 0;
}
