// deform_inst.cu — compiled once per feature set (-DRZ_FEAT=<bits>) so the kernel
// instantiations build in parallel.  Exposes one lookup function per feature set.
#include "deform_kernel.cuh"
#include "kernel_table.h"

#ifndef RZ_FEAT
#error "compile with -DRZ_FEAT=<feature bits>"
#endif

namespace rz {

#define RZ_CAT2(a, b) a##b
#define RZ_CAT(a, b) RZ_CAT2(a, b)

template <int I, int NT, bool ST>
static KernelEntry entry() {
  KernelEntry e;
  e.fn = reinterpret_cast<const void*>(&deform_kernel<I, NT, ST, RZ_FEAT>);
  e.I = I; e.NT = NT; e.staged = ST; e.feat = RZ_FEAT;
  return e;
}

// FEAT == 0 (the plain BDEF path) gets every launch shape; feature sets get a reduced list.
KernelEntry RZ_CAT(lookup_feat_, RZ_FEAT)(int I, int NT, bool staged) {
#define RZ_TRY(i, nt, st) if (I == i && NT == nt && staged == st) return entry<i, nt, st>();
#if RZ_FEAT == 0
  RZ_TRY(1, 256, false) RZ_TRY(2, 256, false) RZ_TRY(4, 256, false) RZ_TRY(8, 256, false)
  RZ_TRY(1, 512, false) RZ_TRY(2, 512, false) RZ_TRY(4, 512, false) RZ_TRY(8, 512, false)
  RZ_TRY(1, 256, true)  RZ_TRY(2, 256, true)  RZ_TRY(4, 256, true)  RZ_TRY(8, 256, true)
  RZ_TRY(1, 512, true)  RZ_TRY(2, 512, true)  RZ_TRY(4, 512, true)  RZ_TRY(8, 512, true)
#else
  RZ_TRY(1, 256, false) RZ_TRY(2, 256, false) RZ_TRY(4, 256, false)
  RZ_TRY(1, 512, true)  RZ_TRY(2, 512, true)  RZ_TRY(4, 512, true)
#endif
#undef RZ_TRY
  KernelEntry none;
  none.fn = nullptr; none.I = 0; none.NT = 0; none.staged = false; none.feat = RZ_FEAT;
  return none;
}

}  // namespace rz
