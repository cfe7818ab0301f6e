// Instantiations of the persistent kernel for one model (own translation unit: built in parallel).
// Two tile granularities: ITEMS = 7 (3584 parents per scan sub-chunk) and ITEMS = 9 (4608; keeps a
// 4096-particle filter of the batched mode in ONE sub-chunk); ITEMS = 2 for small filters.
#include "pf_config.h"
#include "pf_kernels.cuh"

typedef void (*KernelFn)(const pfb200::DeviceProblem);

template <int ITEMS>
static KernelFn pick(int dtype, bool general) {
  using namespace pfb200;
  if (dtype == 0)
    return general ? pf_persistent_kernel<kModelLGSS, double, PFB_BLOCK, ITEMS, PFB_MINB, true>
                   : pf_persistent_kernel<kModelLGSS, double, PFB_BLOCK, ITEMS, PFB_MINB, false>;
  return general ? pf_persistent_kernel<kModelLGSS, float, PFB_BLOCK, ITEMS, PFB_MINB, true>
                 : pf_persistent_kernel<kModelLGSS, float, PFB_BLOCK, ITEMS, PFB_MINB, false>;
}

KernelFn pfb200_kernel_lgss(int dtype, bool general, int items) {
  if (items == PFB_ITEMS_SMALL) return pick<PFB_ITEMS_SMALL>(dtype, general);
  return items == PFB_ITEMS_ALT ? pick<PFB_ITEMS_ALT>(dtype, general) : pick<PFB_ITEMS>(dtype, general);
}
