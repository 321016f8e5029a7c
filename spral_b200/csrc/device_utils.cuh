/* Small device helpers shared by the kernels of the B200 SSIDS engine. */
#pragma once
#include <math_constants.h>
#include "engine.h"

namespace b200 {

/* Number of columns of the current block column that are accepted: the first
 * failing column, moved back by one if that would split a 2x2 pivot (the
 * second column of a 2x2 carries +Inf in dinv[2*j], block_ldlt.hxx:403-406). */
__device__ __forceinline__ int calc_ne(const Front* f) {
   int ne = f->first_fail;
   if (ne > 0 && ne < f->bs && isinf(f->ws->dinv[2 * ne])) ne--;
   return ne;
}

} // namespace b200
