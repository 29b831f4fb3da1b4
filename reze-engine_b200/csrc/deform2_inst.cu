// deform2_inst.cu — instantiations of the two-vertices-per-lane kernel (deform2_kernel.cuh) and their lookup.
// Compiled once per output layout (-DRZ_OUT2=<OUT2_*>) so the layouts build in parallel.
#include "deform2_kernel.cuh"
#include "kernel_table.h"

#ifndef RZ_OUT2
#error "compile with -DRZ_OUT2=<output layout>"
#endif

namespace rz {

#define RZ_CAT2(a, b) a##b
#define RZ_CAT(a, b) RZ_CAT2(a, b)

template <int I, int NT, int MINB, int SB, int NBUF>
static KernelEntry entry2() {
  KernelEntry e;
  e.fn = reinterpret_cast<const void*>(&deform2_kernel<I, NT, MINB, SB, NBUF, RZ_OUT2>);
  e.I = I; e.NT = NT; e.MINB = MINB; e.SB = SB; e.NB = NBUF; e.feat = RZ_OUT2;
  return e;
}

// MINB <= 0 matches the first compiled entry with the requested I and NT; SB <= 0 any sub-batch
KernelEntry RZ_CAT(lookup_v2_out, RZ_OUT2)(int I, int NT, int MINB, int SB) {
#define RZ_TRY(i, nt, mb, sb, nbuf) if (I == i && NT == nt && (MINB <= 0 || MINB == mb) && (SB <= 0 || SB == sb)) return entry2<i, nt, mb, sb, nbuf>();
#if RZ_OUT2 == 0
  RZ_SHAPES_V2(RZ_TRY)
#else
  RZ_SHAPES_V2_LITE(RZ_TRY)
#endif
#undef RZ_TRY
  KernelEntry none;
  none.fn = nullptr; none.I = 0; none.NT = 0; none.MINB = 0; none.SB = 0; none.NB = 0; none.feat = RZ_OUT2;
  return none;
}

}  // namespace rz
