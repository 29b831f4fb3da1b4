// kernel_table.h — lookup of compiled deform_kernel instantiations.
#pragma once
namespace rz {
struct KernelEntry {
  const void* fn;   // __global__ function pointer (nullptr = shape not compiled for this feature set)
  int I, NT;
  bool staged;
  int feat;
};
// feature sets compiled (bit meaning: deform_kernel.cuh FEAT_*); keep in sync with build.py
#define RZ_FEAT_LIST(X) X(0) X(1) X(3) X(4) X(7) X(16) X(19) X(8) X(11) X(15) X(24) X(27)
#define RZ_DECL(f) KernelEntry lookup_feat_##f(int I, int NT, bool staged);
RZ_FEAT_LIST(RZ_DECL)
#undef RZ_DECL
}  // namespace rz
