// kernel_table.h — lookup of compiled deform_kernel instantiations.
#pragma once
namespace rz {
struct KernelEntry {
  const void* fn;   // __global__ function pointer (nullptr = shape not compiled for this feature set)
  int I, NT, MINB;  // instances per group, compute threads (CTA = NT + 32), CTAs/SM the registers are sized for
  int feat;
};
// launch shapes (I, NT, MINB).  FULL: the plain BDEF path; LITE: every other feature set.
#define RZ_SHAPES_FULL(X) \
  X(1, 256, 4) X(2, 256, 3) X(2, 256, 2) X(3, 256, 2) X(4, 256, 1) \
  X(1, 512, 2) X(2, 512, 2) X(2, 512, 1) X(3, 512, 1) X(4, 512, 1) \
  X(2, 768, 1) X(3, 768, 1) X(2, 1024, 1) X(3, 1024, 1)
#define RZ_SHAPES_LITE(X) X(1, 256, 2) X(2, 256, 2) X(2, 512, 1) X(4, 512, 1)
// feature sets compiled (bit meaning: deform_kernel.cuh FEAT_*); keep in sync with build.py
#define RZ_FEAT_LIST(X) X(0) X(1) X(3) X(4) X(7) X(16) X(19) X(8) X(11) X(15) X(24) X(27)
#define RZ_DECL(f) KernelEntry lookup_feat_##f(int I, int NT, int MINB);
RZ_FEAT_LIST(RZ_DECL)
#undef RZ_DECL
}  // namespace rz
