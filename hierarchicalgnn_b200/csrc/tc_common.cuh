// tcgen05 / TMEM / mbarrier PTX wrappers (sm_100a). Filled in with the fused
// tensor-core edge step.
#pragma once
#include "common.cuh"
