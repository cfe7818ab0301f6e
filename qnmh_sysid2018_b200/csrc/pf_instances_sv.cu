// Instantiations of the persistent kernel for one model (own translation unit: built in parallel).
#include "pf_config.h"
#include "pf_kernels.cuh"

typedef void (*KernelFn)(const pfb200::DeviceProblem);

KernelFn pfb200_kernel_sv(int dtype, bool general) {
  using namespace pfb200;
  if (dtype == 0)
    return general ? pf_persistent_kernel<kModelSV, double, PFB_BLOCK, PFB_ITEMS, PFB_MINB, true>
                   : pf_persistent_kernel<kModelSV, double, PFB_BLOCK, PFB_ITEMS, PFB_MINB, false>;
  return general ? pf_persistent_kernel<kModelSV, float, PFB_BLOCK, PFB_ITEMS, PFB_MINB, true>
                 : pf_persistent_kernel<kModelSV, float, PFB_BLOCK, PFB_ITEMS, PFB_MINB, false>;
}
