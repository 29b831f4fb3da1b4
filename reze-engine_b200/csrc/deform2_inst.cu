// deform2_inst.cu — instantiations of the two-vertices-per-lane plain-path kernel (deform2_kernel.cuh) and their lookup.
#include "deform2_kernel.cuh"
#include "kernel_table.h"

namespace rz {

template <int I, int NT, int MINB, int SB, int NBUF>
static KernelEntry entry2() {
  KernelEntry e;
  e.fn = reinterpret_cast<const void*>(&deform2_kernel<I, NT, MINB, SB, NBUF>);
  e.I = I; e.NT = NT; e.MINB = MINB; e.SB = SB; e.NB = NBUF; e.feat = 0;
  return e;
}

// MINB <= 0 matches the first compiled entry with the requested I and NT; SB <= 0 any sub-batch
KernelEntry lookup_v2(int I, int NT, int MINB, int SB) {
#define RZ_TRY(i, nt, mb, sb, nbuf) if (I == i && NT == nt && (MINB <= 0 || MINB == mb) && (SB <= 0 || SB == sb)) return entry2<i, nt, mb, sb, nbuf>();
  RZ_SHAPES_V2(RZ_TRY)
#undef RZ_TRY
  KernelEntry none;
  none.fn = nullptr; none.I = 0; none.NT = 0; none.MINB = 0; none.SB = 0; none.NB = 0; none.feat = 0;
  return none;
}

}  // namespace rz
