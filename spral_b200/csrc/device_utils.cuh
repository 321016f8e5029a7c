/* Small device helpers shared by the kernels of the B200 SSIDS engine. */
#pragma once
#include <math_constants.h>
#include "engine.h"
#include "pivot_state.h"

namespace b200 {

/* calc_ne(): pivot_state.h */

} // namespace b200
