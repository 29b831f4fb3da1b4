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

template <int I, int NT, int MINB, int SB, int NB>
static KernelEntry entry() {
  KernelEntry e;
  e.fn = reinterpret_cast<const void*>(&deform_kernel<I, NT, MINB, RZ_FEAT, SB, NB>);
  e.I = I; e.NT = NT; e.MINB = MINB; e.SB = SB; e.NB = NB; e.feat = RZ_FEAT;
  return e;
}

// MINB <= 0 matches the first compiled entry with the requested I and NT
KernelEntry RZ_CAT(lookup_feat_, RZ_FEAT)(int I, int NT, int MINB) {
#define RZ_TRY(i, nt, mb, sb, nb) if (I == i && NT == nt && (MINB <= 0 || MINB == mb)) return entry<i, nt, mb, sb, nb>();
#if RZ_FEAT == 0
  RZ_SHAPES_FULL(RZ_TRY)
#elif RZ_FEAT_IS_MID(RZ_FEAT)
  RZ_SHAPES_MID(RZ_TRY)
#else
  RZ_SHAPES_LITE(RZ_TRY)
#endif
#undef RZ_TRY
  KernelEntry none;
  none.fn = nullptr; none.I = 0; none.NT = 0; none.MINB = 0; none.SB = 0; none.NB = 0; none.feat = RZ_FEAT;
  return none;
}

}  // namespace rz
