// kernel_table.h — lookup of compiled deform_kernel instantiations.
#pragma once
namespace rz {
struct KernelEntry {
  const void* fn;   // __global__ function pointer (nullptr = shape not compiled for this feature set)
  int I, NT, MINB;  // instances per group, threads per CTA, CTAs/SM the registers are sized for
  int SB, NB;       // gather sub-batch, staging buffers per warp
  int feat;
};
// launch shapes (I, NT, MINB).  FULL: the plain BDEF path; LITE: every other feature set.
// X(I, NT, MINB, SB, NB)
#define RZ_SHAPES_FULL(X) \
  X(1, 256, 4, 1, 2) X(2, 256, 3, 2, 2) X(2, 256, 2, 2, 2) X(3, 256, 2, 3, 2) X(4, 256, 1, 4, 2) X(6, 256, 1, 3, 2) \
  X(1, 512, 2, 1, 2) X(2, 512, 2, 2, 2) X(2, 512, 1, 2, 2) X(3, 512, 1, 3, 2) X(4, 512, 1, 4, 2) X(6, 512, 1, 3, 1) \
  X(2, 768, 1, 2, 2) X(3, 768, 1, 3, 2) X(4, 768, 1, 2, 1) X(2, 1024, 1, 2, 2) X(3, 1024, 1, 3, 2)
#define RZ_SHAPES_LITE(X) X(1, 256, 2, 1, 2) X(2, 256, 2, 2, 2) X(2, 512, 1, 2, 2) X(4, 512, 1, 4, 2)
// MID: feature sets without morph / SDEF / global palette (bounds, positions only, outline, interleaved): LITE + the two
// wide single-buffered shapes the plain path prefers
#define RZ_SHAPES_MID(X) RZ_SHAPES_LITE(X) X(6, 512, 1, 3, 1) X(4, 768, 1, 2, 1)
#define RZ_FEAT_IS_MID(f) ((f) != 0 && ((f) & (1 | 2 | 8)) == 0)
// feature sets compiled (bit meaning: deform_kernel.cuh FEAT_*); keep in sync with build.py
// Every combination of the public flags resolves to one of these (rze_b200.cu resolve_feat picks the smallest superset;
// extras of a superset are inert): per output layout {planar, positions only 16, outline 32, interleaved 64} x palette
// {shared, global 8} there is the bare set and the "everything" set (morph 1 + SDEF 2 + bounds 4), plus the common ones.
#define RZ_FEAT_LIST(X) X(0) X(1) X(3) X(4) X(7) X(16) X(19) X(20) X(23) X(8) X(11) X(15) X(24) X(31) X(32) X(39) X(40) X(47) \
  X(64) X(71) X(72) X(79)
// two-vertices-per-lane plain path (deform2_kernel.cuh): X(I, NT, MINB, SB, NBUF)
#define RZ_SHAPES_V2(X) \
  X(6, 384, 1, 2, 2) X(6, 512, 1, 1, 2) X(6, 512, 1, 1, 3) X(6, 256, 1, 3, 2) X(6, 256, 1, 2, 2) X(5, 512, 1, 1, 2) \
  X(4, 384, 1, 2, 2) X(4, 512, 1, 2, 2) X(4, 512, 1, 1, 2) X(4, 256, 1, 2, 2) X(4, 256, 1, 4, 1) \
  X(3, 256, 2, 3, 1) X(3, 256, 1, 3, 2) X(3, 512, 1, 1, 2) X(2, 256, 2, 2, 1) X(2, 256, 1, 2, 2) X(2, 512, 1, 1, 2) X(1, 256, 2, 1, 2)
// the other output layouts (positions only / outline hull / interleaved): fewer shapes
#define RZ_SHAPES_V2_LITE(X) \
  X(6, 512, 1, 1, 2) X(4, 512, 1, 2, 2) X(4, 512, 1, 1, 2) X(4, 384, 1, 2, 2) X(4, 256, 1, 2, 2) X(3, 512, 1, 1, 2) X(2, 512, 1, 1, 2) \
  X(2, 256, 2, 2, 1) X(2, 256, 1, 2, 2) X(1, 256, 2, 1, 2)
KernelEntry lookup_v2_out0(int I, int NT, int MINB, int SB);   // OUT2_PLANAR
KernelEntry lookup_v2_out1(int I, int NT, int MINB, int SB);   // OUT2_NONRM
KernelEntry lookup_v2_out2(int I, int NT, int MINB, int SB);   // OUT2_HULL
KernelEntry lookup_v2_out3(int I, int NT, int MINB, int SB);   // OUT2_ILV
KernelEntry lookup_v2_out4(int I, int NT, int MINB, int SB);   // OUT2_BOUNDS
#define RZ_DECL(f) KernelEntry lookup_feat_##f(int I, int NT, int MINB);
RZ_FEAT_LIST(RZ_DECL)
#undef RZ_DECL
}  // namespace rz
