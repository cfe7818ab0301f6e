// ONE (model, dtype, flavour) slice of the persistent kernel's instantiations: compiled 16 times with
// -DPFB_INST_MODEL=0..3 -DPFB_INST_DTYPE=0|1 -DPFB_INST_GENERAL=0|1 into separate objects (built in
// parallel; see Makefile / __graft_entry__.build()).  Each object exposes one getter that picks the tile
// granularity: ITEMS = 7 (3584 parents per scan sub-chunk), 9 (4608: a 4096-particle filter of the batched
// mode in ONE sub-chunk) or 2 (small filters).
#include <type_traits>

#include "pf_config.h"
#include "pf_kernels.cuh"
#include "pf_cluster.cuh"

#if !defined(PFB_INST_MODEL) || !defined(PFB_INST_DTYPE) || !defined(PFB_INST_GENERAL)
#error "compile with -DPFB_INST_MODEL=<0..3> -DPFB_INST_DTYPE=<0|1> -DPFB_INST_GENERAL=<0|1>"
#endif

typedef void (*KernelFn)(const pfb200::DeviceProblem);
#define PFB_CAT2(a, b, c, d) a##_##b##_##c##_##d
#define PFB_CAT(a, b, c, d) PFB_CAT2(a, b, c, d)

KernelFn PFB_CAT(pfb200_kernel, PFB_INST_MODEL, PFB_INST_DTYPE, PFB_INST_GENERAL)(int items) {
  using namespace pfb200;
  using Real = std::conditional<PFB_INST_DTYPE == 0, double, float>::type;
  constexpr bool kGeneral = PFB_INST_GENERAL != 0;
  if (items == PFB_ITEMS_SMALL) return pf_persistent_kernel<PFB_INST_MODEL, Real, PFB_BLOCK, PFB_ITEMS_SMALL, PFB_MINB, kGeneral>;
  if (items == PFB_ITEMS_ALT) return pf_persistent_kernel<PFB_INST_MODEL, Real, PFB_BLOCK, PFB_ITEMS_ALT, PFB_MINB, kGeneral>;
  return pf_persistent_kernel<PFB_INST_MODEL, Real, PFB_BLOCK, PFB_ITEMS, PFB_MINB, kGeneral>;
}

// the cluster kernel (small and medium filters on one thread-block cluster) of this (model, dtype, flavour);
// the fully adapted filter stays on the persistent kernel
KernelFn PFB_CAT(pfb200_cluster_kernel, PFB_INST_MODEL, PFB_INST_DTYPE, PFB_INST_GENERAL)() {
  using namespace pfb200;
  using Real = std::conditional<PFB_INST_DTYPE == 0, double, float>::type;
#if PFB_INST_MODEL == 3
  return nullptr;
#else
  return pf_cluster_kernel<PFB_INST_MODEL, Real, PFB_INST_GENERAL != 0>;
#endif
}
