/* C client of the spral_ssids.h-compatible interface: coordinate input (ssids_analyse_coord), the
 * matching-based ordering (options.ordering = 2) with its saved scaling (options.scaling = 3), the
 * computed scalings (options.scaling = 1, 4) and factor_ptr32, on the 5x5 matrix of the reference's C
 * example (examples/C/ssids.c: solution 1..5).  Prints "CAPI COORD OK" and returns 0 when every check holds. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "spral_ssids_compat.h"

static int fails = 0;
#define CHECK(cond, msg) do { if (!(cond)) { printf("FAIL: %s (line %d)\n", msg, __LINE__); fails++; } } while (0)

/* lower triangle of the example: (1,1)=2 (2,1)=1 (2,2)=4 (3,2)=1 (5,2)=1 (3,3)=3 (4,3)=2 (4,4)=-1 (5,5)=2 */
static void solve_and_check(void* akeep, void* fkeep, struct spral_ssids_options* opt, const char* what) {
   struct spral_ssids_inform inf;
   double x[] = {4.0, 17.0, 19.0, 2.0, 12.0};
   spral_ssids_solve1(0, x, akeep, fkeep, opt, &inf);
   CHECK(inf.flag == 0, what);
   for (int i = 0; i < 5; ++i) CHECK(fabs(x[i] - (i + 1.0)) < 1e-11, what);
}

int main(void) {
   struct spral_ssids_options opt;
   struct spral_ssids_inform inf;
   void *akeep = NULL, *fkeep = NULL;
   /* shuffled triplets, 1-based; (2,3) given in the UPPER triangle, (2,2) split in two duplicates,
    * one entry out of range */
   int row[] = {5, 1, 2, 4, 2, 3, 2, 5, 4, 2, 7};
   int col[] = {5, 1, 3, 4, 2, 3, 1, 2, 3, 2, 1};
   double val[] = {2.0, 2.0, 1.0, -1.0, 1.5, 3.0, 1.0, 1.0, 2.0, 2.5, 9.0};
   int64_t ne = 11;

   spral_ssids_default_options(&opt);
   opt.array_base = 1;
   spral_ssids_analyse_coord(5, NULL, ne, row, col, NULL, &akeep, &opt, &inf);
   CHECK(inf.flag == 3, "coord: duplicates and out-of-range entries -> warning 3");
   CHECK(inf.matrix_dup == 1 && inf.matrix_outrange == 1, "coord: counts");
   CHECK(inf.num_factor == 15 && inf.num_flops == 55, "coord: analyse prediction");
   spral_ssids_factor(false, NULL, NULL, val, NULL, akeep, &fkeep, &opt, &inf);
   CHECK(inf.flag >= 0 && inf.num_neg == 1 && inf.matrix_rank == 5, "coord: factor");
   solve_and_check(akeep, fkeep, &opt, "coord: solution 1..5");

   /* computed scalings */
   double scale[5];
   opt.scaling = 1;
   spral_ssids_factor(false, NULL, NULL, val, scale, akeep, &fkeep, &opt, &inf);
   CHECK(inf.flag >= 0 && inf.num_neg == 1, "scaling = 1 (matching-based)");
   for (int i = 0; i < 5; ++i) CHECK(scale[i] > 0 && isfinite(scale[i]), "scaling returned to the caller");
   solve_and_check(akeep, fkeep, &opt, "scaling = 1: solution");
   opt.scaling = 4;
   spral_ssids_factor_ptr32(false, NULL, NULL, val, NULL, akeep, &fkeep, &opt, &inf);
   CHECK(inf.flag >= 0 && inf.num_neg == 1, "scaling = 4 (equilibration), factor_ptr32");
   solve_and_check(akeep, fkeep, &opt, "scaling = 4: solution");
   opt.scaling = 3;
   spral_ssids_factor(false, NULL, NULL, val, NULL, akeep, &fkeep, &opt, &inf);
   CHECK(inf.flag == -15, "scaling = 3 without a matching-based ordering -> NO_SAVED_SCALING");
   spral_ssids_free(&akeep, &fkeep);

   /* matching-based ordering needs the values */
   opt.ordering = 2;
   opt.scaling = 3;
   int order[5];
   spral_ssids_analyse_coord(5, order, ne, row, col, NULL, &akeep, &opt, &inf);
   CHECK(inf.flag == -9, "ordering = 2 without values -> ERROR_VAL");
   spral_ssids_analyse_coord(5, order, ne, row, col, val, &akeep, &opt, &inf);
   CHECK(inf.flag == 3, "ordering = 2: analyse");
   int seen[6] = {0};
   for (int i = 0; i < 5; ++i) { CHECK(order[i] >= 1 && order[i] <= 5 && !seen[order[i]], "order is a permutation"); if (order[i] >= 1 && order[i] <= 5) seen[order[i]] = 1; }
   spral_ssids_factor(false, NULL, NULL, val, NULL, akeep, &fkeep, &opt, &inf);
   CHECK(inf.flag >= 0 && inf.num_neg == 1 && inf.matrix_rank == 5, "ordering = 2, scaling = 3: factor");
   solve_and_check(akeep, fkeep, &opt, "ordering = 2: solution");
   CHECK(spral_ssids_free(&akeep, &fkeep) == 0, "free");

   if (fails) { printf("CAPI COORD FAILED: %d checks\n", fails); return 1; }
   printf("CAPI COORD OK\n");
   return 0;
}
