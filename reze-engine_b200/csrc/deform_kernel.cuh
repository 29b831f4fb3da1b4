// deform_kernel.cuh — the fused morph + BDEF1/2/4/SDEF skinning kernel for sm_100a.
//
// Replaces the per-vertex blend the reference runs inside three WGSL vertex shaders
// (engine.ts:245-276 main, 431-463 outline, 692-715 depth-only) and materialises the
// skinned stream once per frame for K independent character instances.
//
// Work decomposition (B200: 148 SMs, 227 KB smem/SM, HBM-bound on the 24 B/vertex write):
//   work item = (group of I instances) x (chunk of vertex tiles); persistent CTAs pull items from
//   an atomic counter.  Per item the I bone palettes (B x 48 B each, 3x4 skin matrices in "pair layout")
//   are staged into shared memory with one cp.async.bulk (TMA bulk copy, mbarrier complete_tx) per
//   instance.  Every thread keeps ONE vertex (pos, normal, 4 joints, 4 pre-normalised weights: 52 B,
//   float4-vectorised, L2-resident, prefetched one pass ahead) in registers and evaluates it for the
//   I instances, so the static mesh is read once per I outputs.
//   Results go to a WARP-PRIVATE, double-buffered shared-memory staging area laid out exactly like 32
//   vertices of the output planes and leave with cp.async.bulk shared->global (TMA bulk store, L2
//   evict-first) issued by one elected lane: no CTA-wide barrier and no cross-warp dependency inside an
//   item.  The palette gather is the SM-side bottleneck (48 B per influence through the 128 B/clk
//   shared-memory pipe), so the influence loop is specialised per warp on the warp-maximum influence
//   count, lanes and influence slots are arranged at load time so that aligned lane pairs read the same
//   rows (lane_plan.h), and nothing branches per lane.
//   Morphs: lane-interleaved rows per warp, accumulated before the blend.  SDEF: a dense second phase
//   over (vertex, instance) pairs.  Fused consumers: AABB, outline hull plane, interleaved stream.
//   No tensor cores: the work is a gather of 3x4 mat-vecs.  DESIGN.md section 4 has the details.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace rz {

constexpr int kTile = 256;          // vertices per preprocessing tile (permutation unit)
constexpr int kMorphPF = 4;         // morph entries of a vertex fetched one pass ahead
constexpr int kMorphBatch = 8;      // further entries per L2 round trip
constexpr int kRowF4 = 3;           // float4 per bone.  PAIR LAYOUT of the 3x4 skin matrix m[row][col] (48 B):
                                    //   A = (m00,m10,m01,m11)  B = (m02,m12,m03,m13)  C = (m20,m21,m22,m23)
                                    // rows 0/1 are interleaved so that (x,y) of a transformed point come out of one FFMA2

// meta word layout
constexpr uint32_t kMetaSlotMask = 0x3FFu;   // bits 0-9  : slot inside the tile
constexpr int      kMetaNinfShift = 10;      // bits 10-12: 1 + index of last non-zero weight
constexpr uint32_t kMetaValid = 1u << 13;
constexpr uint32_t kMetaMorph = 1u << 14;
constexpr uint32_t kMetaSdef  = 1u << 15;

enum : int { FEAT_MORPH = 1, FEAT_SDEF = 2, FEAT_BOUNDS = 4, FEAT_GPAL = 8, FEAT_NONRM = 16,
              FEAT_HULL = 32,   // third output plane: outline hull pos' + n^' * edgeScale[v]   (engine.ts:458-461)
              FEAT_ILV = 64 };  // one interleaved 32-byte [pos, nrm, uv] stream per instance   (engine.ts:340-347 vertex stride)

struct DeformParams {
  const float4* __restrict__ rec0;     // [Vp] px,py,pz, w0      (weights already normalised like engine.ts:255-258)
  const float4* __restrict__ rec1;     // [Vp] nx,ny,nz, w1
  const float4* __restrict__ rec2;     // [Vp] w2, w3, joints01 bits, joints23 bits
  const uint32_t* __restrict__ meta;   // [Vp] slot | ninf | flags
  const uint2*  __restrict__ mrange;   // [Vp/32] per warp: (first entry, depth) into ments
  const float4* __restrict__ ments;    // lane-interleaved per warp: entry u of lane l at first + u*32 + l = (dx,dy,dz, morph id bits); padding = zeros
  const uint32_t* __restrict__ sdefIdx;// [Vp] word of lane l = descriptor of the l-th SDEF vertex of its warp: table index | output slot << 24 (~0u: none)
  const float4* __restrict__ sdefTab;  // [nSdef*3]: (C.xyz,c0.x) (c0.yz,c1.xy) (c1.z, w0, w1, palette rows j0 | j1 << 16)
  const float*  __restrict__ skin;     // [P][B][12]
  const float4* __restrict__ quat;     // [P][B] rotation of each skin matrix as a quaternion (SDEF only; skin_quats_kernel)
  const uint32_t* __restrict__ inst2pal; // [K] or nullptr (identity)
  const float*  __restrict__ mweights; // dense [K][Mpad]
  const float* __restrict__ edge;      // [Vp] outline offset of the lane's vertex = material edgeSize * 0.01 (FEAT_HULL)
  const float2* __restrict__ uv;       // [Vp] texture coordinates of the lane's vertex, passed through (FEAT_ILV)
  float* __restrict__ out;
  float* __restrict__ bounds;          // [K][6] as ordered ints, or nullptr
  unsigned long long instStrideF;      // floats between instances
  unsigned long long nrmOffF;          // floats from pos plane to normal plane
  unsigned long long hullOffF;         // floats from pos plane to outline-hull plane (FEAT_HULL)
  uint32_t V, B, nTiles, K0, Kcount, Mpad;
  uint32_t nGroups, nChunks, tilesPerChunk;   // instance groups; chunks per (fine) instance group; tiles per chunk
  uint32_t nCoarse, nItems;                   // the first nCoarse instance groups are ONE item each (long items first, short ones
                                              // level the tail: rze_b200.cu pick_items); nItems = nCoarse + (nGroups - nCoarse) * nChunks
  uint32_t packedMeta;                 // 1: meta lives in the spare bits of the joint words (B <= 4096), no separate load
  uint32_t posStride, rowStride;       // palette addressing: chunk r of palette row `pos` sits at pos*posStride + r*rowStride bytes
  const uint32_t* __restrict__ chunkTab; // [nChunks+1] tile boundaries of cost-balanced chunks, or nullptr (uniform tilesPerChunk)
  uint32_t* counter;
};

// ------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}"
      ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
// global -> shared bulk copy (TMA, 1-D), completion on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar), "l"(pol) : "memory");
}
// shared -> global bulk store (TMA, 1-D), tracked by bulk groups
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src_smem, uint32_t bytes, uint64_t pol) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
               ::"l"(dst), "r"(src_smem), "r"(bytes), "l"(pol) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// exactly one lane of a converged warp (lets the compiler keep TMA operands in uniform registers)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ float4 ldg_el(const float4* p, uint64_t pol) {   // read-only, keep in L2 (mesh is re-read by every group)
  float4 r;
  asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ uint32_t ldg_el(const uint32_t* p, uint64_t pol) {
  uint32_t r;
  asm volatile("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ uint2 ldg_el(const uint2* p, uint64_t pol) {
  uint2 r;
  asm volatile("ld.global.nc.L2::cache_hint.v2.u32 {%0,%1}, [%2], %3;" : "=r"(r.x), "=r"(r.y) : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ float4 lds128(uint32_t a) {
  float4 r;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(a));
  return r;
}
__device__ __forceinline__ float lds32(uint32_t a) {
  float r;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(r) : "r"(a));
  return r;
}
__device__ __forceinline__ void sts3(uint32_t a, float x, float y, float z) {
  asm volatile("st.shared.f32 [%0], %1;\n\tst.shared.f32 [%0+4], %2;\n\tst.shared.f32 [%0+8], %3;" ::"r"(a), "f"(x), "f"(y), "f"(z) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t a, float x, float y, float z, float w) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ void st_cs(float* p, float v) { asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory"); }

// Blackwell packed FP32 (FFMA2 / FMUL2: two fp32 lanes per instruction).  The blend of 3x4 matrices is element-wise,
// so a float4 row costs 2 issue slots instead of 4.
__device__ __forceinline__ float4 f4_scale(float4 a, float2 s) {
  const float2 lo = __fmul2_rn(make_float2(a.x, a.y), s), hi = __fmul2_rn(make_float2(a.z, a.w), s);
  return make_float4(lo.x, lo.y, hi.x, hi.y);
}
__device__ __forceinline__ float4 f4_fma(float4 a, float2 s, float4 c) {
  const float2 lo = __ffma2_rn(make_float2(a.x, a.y), s, make_float2(c.x, c.y));
  const float2 hi = __ffma2_rn(make_float2(a.z, a.w), s, make_float2(c.z, c.w));
  return make_float4(lo.x, lo.y, hi.x, hi.y);
}

// order-preserving float <-> int for atomic min/max
__device__ __forceinline__ int f2ord(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7FFFFFFF; }

// ---- SDEF helpers: same formulas as the reference's math (math.ts:406-448 toQuatFromArray, 156-189 slerp,
//      352-384 fromQuat) applied to the rotation part of two skin matrices.
struct Q4 { float x, y, z, w; };
__device__ __forceinline__ Q4 quat_from_rows(float4 r0, float4 r1, float4 r2) {
  const float m00 = r0.x, m01 = r0.y, m02 = r0.z;
  const float m10 = r1.x, m11 = r1.y, m12 = r1.z;
  const float m20 = r2.x, m21 = r2.y, m22 = r2.z;
  const float trace = m00 + m11 + m22;
  float x, y, z, w;
  if (trace > 0.f) {
    const float s = sqrtf(trace + 1.0f) * 2.f;
    w = 0.25f * s; x = (m21 - m12) / s; y = (m02 - m20) / s; z = (m10 - m01) / s;
  } else if (m00 > m11 && m00 > m22) {
    const float s = sqrtf(1.0f + m00 - m11 - m22) * 2.f;
    w = (m21 - m12) / s; x = 0.25f * s; y = (m01 + m10) / s; z = (m02 + m20) / s;
  } else if (m11 > m22) {
    const float s = sqrtf(1.0f + m11 - m00 - m22) * 2.f;
    w = (m02 - m20) / s; x = (m01 + m10) / s; y = 0.25f * s; z = (m12 + m21) / s;
  } else {
    const float s = sqrtf(1.0f + m22 - m00 - m11) * 2.f;
    w = (m10 - m01) / s; x = (m02 + m20) / s; y = (m12 + m21) / s; z = 0.25f * s;
  }
  const float inv = 1.0f / sqrtf(x * x + y * y + z * z + w * w);
  return Q4{x * inv, y * inv, z * inv, w * inv};
}
// sin(y) for 0 <= y <= pi/2: y * (1 + z*P(z)), z = y^2, Taylor through y^13 (truncation 7e-10 at pi/2)
__device__ __forceinline__ float sin_0_halfpi(float y) {
  const float z = y * y;
  float p = 1.6059044e-10f;
  p = fmaf(p, z, -2.5052108e-08f);
  p = fmaf(p, z, 2.7557319e-06f);
  p = fmaf(p, z, -1.9841270e-04f);
  p = fmaf(p, z, 8.3333333e-03f);
  p = fmaf(p, z, -1.6666667e-01f);
  return fmaf(y * z, p, y);
}
__device__ __forceinline__ Q4 quat_slerp(Q4 a, Q4 b, float t) {
  float c = a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
  if (c < 0.f) { c = -c; b.x = -b.x; b.y = -b.y; b.z = -b.z; b.w = -b.w; }
  if (c > 0.9995f) {
    const float x = a.x + t * (b.x - a.x), y = a.y + t * (b.y - a.y), z = a.z + t * (b.z - a.z), w = a.w + t * (b.w - a.w);
    const float inv = rsqrtf(x * x + y * y + z * z + w * w);        // 2 ulp, far inside the parity tolerance
    return Q4{x * inv, y * inv, z * inv, w * inv};
  }
  // every sine argument lies in [0, pi/2] here (0 <= c <= 0.9995, 0 <= t <= 1).  __sinf is not an option: near the 0.9995
  // threshold th0 ~ 0.03 rad and its absolute error (2^-21.4) would be a 1e-5 RELATIVE error of s0/s1 (measured: 2.2e-5 on
  // config 3), outside the parity tolerance; sin_0_halfpi keeps ~1e-7 relative without sinf's range-reduction slow path.
  const float th0 = acosf(c), rs = __fdividef(1.0f, sin_0_halfpi(th0)), th = th0 * t;
  const float s0 = sin_0_halfpi(th0 - th) * rs, s1 = sin_0_halfpi(th) * rs;
  return Q4{s0 * a.x + s1 * b.x, s0 * a.y + s1 * b.y, s0 * a.z + s1 * b.z, s0 * a.w + s1 * b.w};
}

template <int N> struct IntC { static constexpr int value = N; };

// per-thread vertex record
struct VRec {
  float4 r0, r1, r2;
  uint32_t meta;
  uint32_t sdesc;   // SDEF: descriptor of the lane-th SDEF vertex of this warp (~0u: none)
  uint2 mr;         // MORPH: (first entry, depth) of this warp's lane-interleaved morph entries (warp-uniform)
  float edge;       // HULL: outline offset along the skinned normal
  float2 uv;        // ILV: texture coordinates (passed through)
};

// shared-memory control block at the start of dynamic smem; data starts at byte kCtrlBytes
struct Ctrl {
  unsigned long long palBar;
  uint32_t item;
  uint32_t pad;
};
constexpr uint32_t kCtrlBytes = 64;

// ------------------------------------------------------------------ the kernel
// I    : instances evaluated per vertex pass (palettes resident in shared memory)
// NT   : threads per CTA (multiple of 256)
// MINB : CTAs per SM the register budget is sized for
// FEAT : FEAT_* bit set
// SB   : instances whose gathers are in flight together (sub-batch; divides I; bounds the register footprint)
// NB   : staging buffers per warp (2 = double-buffered; 1 when the palettes of a wide group leave no room)
template <int I, int NT, int MINB, int FEAT, int SB = I, int NB = 2>
__global__ void __launch_bounds__(NT, MINB) deform_kernel(const DeformParams prm) {
  static_assert(I % SB == 0, "sub-batch must divide the group");
  constexpr int kStageBufs = NB;
  constexpr bool MORPH = (FEAT & FEAT_MORPH) != 0;
  constexpr bool SDEF = (FEAT & FEAT_SDEF) != 0;
  constexpr bool BOUNDS = (FEAT & FEAT_BOUNDS) != 0;
  constexpr bool GPAL = (FEAT & FEAT_GPAL) != 0;      // palette too large for smem: gather from global/L1
  constexpr bool NRM = (FEAT & FEAT_NONRM) == 0;
  constexpr bool HULL = (FEAT & FEAT_HULL) != 0;
  constexpr bool ILV = (FEAT & FEAT_ILV) != 0;
  static_assert(!(HULL && !NRM) && !(ILV && !NRM) && !(HULL && ILV), "hull / interleaved output need normals; one layout at a time");
  constexpr int PLANES = ILV ? 1 : (NRM ? 2 : 1) + (HULL ? 1 : 0);
  constexpr uint32_t kVtxB = ILV ? 32u : 12u;          // bytes per vertex of a staging plane
  constexpr uint32_t kPlaneB = 32 * kVtxB;             // bytes of one warp-private staging plane (32 vertices)
  constexpr uint32_t kNrmO = ILV ? 12u : kPlaneB;      // normal of a vertex relative to its position
  constexpr uint32_t kHullO = 2 * kPlaneB;             // outline hull position relative to its position (HULL)
  // SDEF: the dense phase addresses I consecutive lanes to the SAME row / slot of I different instances, so every per-instance
  // region (palette, quaternions, staging) is skewed by one 16-byte bank group per instance: otherwise those lanes collide
  // I-fold in one bank group (ncu on config 3: 39 % of the shared-memory wavefronts were bank conflicts)
  constexpr uint32_t kSkew = SDEF ? 16u : 0u;
  constexpr uint32_t kInstB = PLANES * kPlaneB + kSkew; // one instance of one warp
  constexpr uint32_t kBufB = I * kInstB - kSkew;       // one staging buffer of one warp (no skew after the last instance)

  extern __shared__ __align__(128) unsigned char smem_raw[];
  const uint32_t sbase = smem_u32(smem_raw);
  Ctrl* ctrl = reinterpret_cast<Ctrl*>(smem_raw);
  const uint32_t palBar = sbase + (uint32_t)offsetof(Ctrl, palBar);
  const uint32_t B = prm.B;
  const uint32_t sPal = sbase + kCtrlBytes;
  const uint32_t palBytes = GPAL ? 0u : B * 48u;
  const uint32_t palStride = GPAL ? 0u : palBytes + kSkew;
  const uint32_t sMw = sPal + (uint32_t)I * palStride - (GPAL ? 0u : kSkew);
  const uint32_t mwBytes = MORPH ? prm.Mpad * 4u : 0u;
  const uint32_t sQuat = sMw + (uint32_t)I * mwBytes;               // [I][B] float4 quaternions (SDEF, palette in smem)
  const uint32_t quatBytes = (SDEF && !GPAL) ? B * 16u : 0u;
  const uint32_t quatStride = (SDEF && !GPAL) ? quatBytes + kSkew : 0u;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform by construction: TMA operands stay in uniform registers
  // warp-private staging: [kStageBufs][I][PLANES][32*3 floats], laid out exactly like 32 vertices of the output planes
  const uint32_t sStageW = sQuat + (uint32_t)I * quatStride - ((SDEF && !GPAL) ? kSkew : 0u) + (uint32_t)warp * (kStageBufs * kBufB);

  if (tid == 0) {
    mbar_init(palBar, 1);
    fence_mbar_init();
  }
  __syncthreads();

  uint32_t palPhase = 0;
  uint32_t sbuf = 0;                                   // which of the warp's two staging buffers the next pass fills
  const uint64_t polFirst = policy_evict_first();
  const uint64_t polLast = policy_evict_last();

  for (;;) {
    if (tid == 0) ctrl->item = atomicAdd(prm.counter, 1u);
    __syncthreads();
    const uint32_t item = *reinterpret_cast<volatile uint32_t*>(&ctrl->item);
    if (item >= prm.nItems) break;
    uint32_t g = item, tile0 = 0, tile1 = prm.nTiles;
    if (item >= prm.nCoarse) {
      const uint32_t r = item - prm.nCoarse, gq = r / prm.nChunks, chunk = r - gq * prm.nChunks;
      g = prm.nCoarse + gq;
      tile0 = prm.chunkTab ? __ldg(prm.chunkTab + chunk) : chunk * prm.tilesPerChunk;
      tile1 = prm.chunkTab ? __ldg(prm.chunkTab + chunk + 1) : min(prm.nTiles, tile0 + prm.tilesPerChunk);
    }
    const uint32_t kBase = prm.K0 + g * I;                       // first instance of this group
    const uint32_t nInst = min((uint32_t)I, prm.K0 + prm.Kcount - kBase);
    float* const outItem = prm.out + (size_t)kBase * prm.instStrideF;   // pos plane of the group's first instance

    // ---- stage the palettes (+ morph weights) of the I instances: TMA bulk copies on one mbarrier
    const float* gpal[I];
#pragma unroll
    for (int i = 0; i < I; ++i) {
      const uint32_t k = kBase + min((uint32_t)i, nInst - 1);    // a partial group re-evaluates its last instance (never stored)
      const uint32_t pidx = prm.inst2pal ? __ldg(prm.inst2pal + k) : k;
      gpal[i] = prm.skin + (size_t)pidx * B * 12;
    }
    if (!GPAL || MORPH) {
      if (tid == 0) {
        mbar_expect_tx(palBar, (uint32_t)I * (palBytes + mwBytes + quatBytes));
#pragma unroll
        for (int i = 0; i < I; ++i) {
          if (!GPAL) bulk_g2s(sPal + (uint32_t)i * palStride, gpal[i], palBytes, palBar, polLast);
          if (SDEF && !GPAL)   // same palette index as gpal[i]: (gpal[i] - skin) / 12 floats = palette * B rows
            bulk_g2s(sQuat + (uint32_t)i * quatStride, prm.quat + (size_t)(gpal[i] - prm.skin) / 12, quatBytes, palBar, polLast);
          if (MORPH) {
            const uint32_t k = kBase + min((uint32_t)i, nInst - 1);
            bulk_g2s(sMw + (uint32_t)i * mwBytes, prm.mweights + (size_t)k * prm.Mpad, mwBytes, palBar, polLast);
          }
        }
      }
    }

    float bmin[BOUNDS ? I : 1][3], bmax[BOUNDS ? I : 1][3];
    if (BOUNDS) {
#pragma unroll
      for (int i = 0; i < I; ++i)
#pragma unroll
        for (int c = 0; c < 3; ++c) { bmin[i][c] = 3.4e38f; bmax[i][c] = -3.4e38f; }
    }

    // vertex record of a pass (the next pass is prefetched while the current one is evaluated)
    auto load_rec = [&](uint32_t t) -> VRec {
      VRec v;
      const uint32_t p = t * kTile + tid;
      if (t + tid / kTile < tile1) {
        v.r0 = ldg_el(prm.rec0 + p, polLast);
        v.r1 = ldg_el(prm.rec1 + p, polLast);
        v.r2 = ldg_el(prm.rec2 + p, polLast);
        v.meta = prm.packedMeta ? 0u : ldg_el(prm.meta + p, polLast);
        v.sdesc = SDEF ? ldg_el(prm.sdefIdx + p, polLast) : ~0u;
        v.mr = MORPH ? ldg_el(prm.mrange + p / 32u, polLast) : make_uint2(0u, 0u);
        v.edge = HULL ? __ldg(prm.edge + p) : 0.f;
        v.uv = ILV ? __ldg(prm.uv + p) : make_float2(0.f, 0.f);
      } else {                                                  // the far tiles of a wide pass fall off the chunk
        v.r0 = make_float4(0.f, 0.f, 0.f, 1.f);
        v.r1 = make_float4(0.f, 0.f, 0.f, 0.f);
        v.r2 = make_float4(0.f, 0.f, 0.f, 0.f);
        v.meta = (uint32_t)lane | (1u << kMetaNinfShift);
        v.sdesc = ~0u;
        v.mr = make_uint2(0u, 0u);
        v.edge = 0.f;
        v.uv = make_float2(0.f, 0.f);
      }
      return v;
    };
    VRec cur = load_rec(tile0);
    // first kMorphPF morph entries of a record's vertex; fetched one pass ahead so that the dependent chain
    // record -> entries -> weights costs one L2 round trip less per pass
    float4 pf[MORPH ? kMorphPF : 1];
    // (warps without morph rows -- ~90 % of a PMX mesh -- skip the prefetch altogether: mr.y is warp-uniform)
    auto load_pf = [&](const VRec& r) {
      if (r.mr.y == 0u) return;
#pragma unroll
      for (int u = 0; u < kMorphPF; ++u)
        pf[u] = ((uint32_t)u < r.mr.y) ? ldg_el(prm.ments + r.mr.x + (uint32_t)u * 32u + (uint32_t)lane, polLast) : make_float4(0.f, 0.f, 0.f, 0.f);
    };

    if (!GPAL || MORPH) mbar_wait(palBar, palPhase);
    palPhase ^= 1u;
    if (MORPH) load_pf(cur);

    for (uint32_t t = tile0; t < tile1; t += NT / kTile) {
      const VRec v = cur;
      if (t + NT / kTile < tile1) cur = load_rec(t + NT / kTile);
      if (t + (uint32_t)warp / (kTile / 32) >= tile1) continue;       // warp-uniform: this warp has no tile in the pass

      uint32_t jb01 = __float_as_uint(v.r2.z), jb23 = __float_as_uint(v.r2.w);
      uint32_t meta = v.meta;
      if (prm.packedMeta) {
        // 12-bit palette rows + 11 meta bits share the two joint words: saves one LDG (one LSU wavefront) per vertex
        const uint32_t m11 = (jb01 >> 24) | ((jb23 >> 24) << 8);
        meta = (m11 & 31u) | (((m11 >> 5) & 7u) << kMetaNinfShift) | ((m11 & 0x100u) ? kMetaValid : 0u) |
               ((m11 & 0x200u) ? kMetaMorph : 0u) | ((m11 & 0x400u) ? kMetaSdef : 0u);
        jb01 = (jb01 & 0xFFFu) | (((jb01 >> 12) & 0xFFFu) << 16);
        jb23 = (jb23 & 0xFFFu) | (((jb23 >> 12) & 0xFFFu) << 16);
      }
      const bool valid = (meta & kMetaValid) != 0;
      const uint32_t slot = meta & 31u;                           // position among the warp's 32 output vertices
      const int ninf = (int)((meta >> kMetaNinfShift) & 7u);
      const int nmax = __reduce_max_sync(0xffffffffu, ninf);
      const float w0 = v.r0.w, w1 = v.r1.w, w2 = v.r2.x, w3 = v.r2.y;
      const uint32_t pS = prm.posStride, rS = prm.rowStride, rS2 = 2u * rS;
      const uint32_t j0 = (jb01 & 0xFFFFu) * pS, j1 = (jb01 >> 16) * pS;
      const uint32_t j2 = (jb23 & 0xFFFFu) * pS, j3 = (jb23 >> 16) * pS;
      const float vnx = v.r1.x, vny = v.r1.y, vnz = v.r1.z;
      const uint32_t warpVtx0 = t * kTile + (uint32_t)warp * 32u;        // first output vertex of this warp (uniform)
      const uint32_t nWarpVerts = min(32u, prm.V - min(prm.V, warpVtx0));
      const uint32_t nAligned = ILV ? nWarpVerts : (nWarpVerts & ~3u);   // bulk sizes must be multiples of 16 B

      // ---- morph accumulation: p~ = p + sum_m w[k][m] * delta_m[v]   (model space, before skinning)
      float px[MORPH ? I : 1], py[MORPH ? I : 1], pz[MORPH ? I : 1];
#pragma unroll
      for (int i = 0; i < (MORPH ? I : 1); ++i) { px[i] = v.r0.x; py[i] = v.r0.y; pz[i] = v.r0.z; }
      if (MORPH) {
        const uint2 mr = v.mr;                                     // prefetched with the record
        if (mr.y) {
          auto apply = [&](const float4 d) {
            const uint32_t m = __float_as_uint(d.w);               // padded entries: delta 0, morph 0
#pragma unroll
            for (int i = 0; i < I; ++i) {
              const float wgt = lds32(sMw + ((uint32_t)i * prm.Mpad + m) * 4u);
              px[i] = fmaf(wgt, d.x, px[i]);
              py[i] = fmaf(wgt, d.y, py[i]);
              pz[i] = fmaf(wgt, d.z, pz[i]);
            }
          };
#pragma unroll
          for (int u = 0; u < kMorphPF; ++u) apply(pf[u]);      // (pf is only rewritten at the end of this pass, for the next one)
          // the rest kMorphBatch depth steps at a time so their L2 latencies overlap; the bound is warp-uniform
          const float4* ep = prm.ments + mr.x + (uint32_t)lane;
          for (uint32_t u0 = kMorphPF; u0 < mr.y; u0 += kMorphBatch) {
            float4 d[kMorphBatch];
#pragma unroll
            for (int u = 0; u < kMorphBatch; ++u) d[u] = (u0 + u < mr.y) ? ldg_el(ep + (size_t)(u0 + u) * 32u, polLast) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int u = 0; u < kMorphBatch; ++u) apply(d[u]);
          }
        }
      }

      // ---- SDEF runs as a DENSE second phase (after the linear pass, below): the warp's SDEF vertices x I instances are
      // spread over the lanes as (vertex, instance) pairs, so the spherical blend costs one pass per 32 pairs instead of
      // one predicated pass per instance whenever any lane is SDEF.  Owner lanes only park (p~, n) in their staging slots.
      // The pair -> table mapping is static: lane l's descriptor word describes the l-th SDEF vertex of this warp, and
      // the records of round 0 are fetched here, ahead of the palette gathers, so the L2 latency overlaps the linear pass.
      const bool isSdef = SDEF && (meta & kMetaSdef);
      uint32_t sdCount = 0, sdW = ~0u;
      float4 sT0 = make_float4(0.f, 0.f, 0.f, 0.f), sT1 = sT0, sT2 = sT0;
      if (SDEF) {
        sdCount = (uint32_t)__popc(__ballot_sync(0xffffffffu, v.sdesc != ~0u));
        if (sdCount) {
          const uint32_t r = (uint32_t)lane / (uint32_t)I;
          sdW = __shfl_sync(0xffffffffu, v.sdesc, (int)r);
          if (r < sdCount) {
            const float4* e = prm.sdefTab + (size_t)(sdW & 0xFFFFFFu) * 3;
            sT0 = __ldg(e); sT1 = __ldg(e + 1); sT2 = __ldg(e + 2);
          }
        }
      }

      // the bulk stores issued from this buffer two passes ago must have finished reading it
      if (elect_one()) bulk_wait_read<kStageBufs - 1>();
      __syncwarp();
      const uint32_t stg = sStageW + sbuf * kBufB;
      const uint32_t stgLane = stg + slot * kVtxB;

      // splats for the packed mat-vec (once per vertex unless morphing moves the position per instance)
      const float2 nx2 = make_float2(vnx, vnx), ny2 = make_float2(vny, vny), nz2 = make_float2(vnz, vnz);
      const float2 w0_2 = make_float2(w0, w0), w1_2 = make_float2(w1, w1), w2_2 = make_float2(w2, w2), w3_2 = make_float2(w3, w3);

      auto body = [&](auto NM, const int i0) {
        constexpr int NV = decltype(NM)::value;                   // 0: rigid warp with unit weights, else warp-max influence count
        constexpr int NMAX = NV == 0 ? 1 : NV;
        // ---- phase 1: palette gathers of influences 0/1 for all I instances, issued back to back (latency overlaps).
        // No predication: a lane whose weight for influence k is zero carries (set at load time, rze_b200.cu) the joint of
        // an ACTIVE lane of its own warp, so its gather costs no extra shared-memory wavefront (same 16-byte chunk is
        // broadcast) and its FFMA adds an exact 0.
        float4 a0[SB], a1[SB], a2[SB], b0[SB], b1[SB], b2[SB];
#pragma unroll
        for (int ii = 0; ii < SB; ++ii) {
          const int i = i0 + ii;
          if (GPAL) {
            const float4* pal = reinterpret_cast<const float4*>(gpal[i]);
            a0[ii] = pal[j0 / 16]; a1[ii] = pal[(j0 + rS) / 16]; a2[ii] = pal[(j0 + rS2) / 16];
            if (NMAX > 1) { b0[ii] = pal[j1 / 16]; b1[ii] = pal[(j1 + rS) / 16]; b2[ii] = pal[(j1 + rS2) / 16]; }
          } else {
            const uint32_t pb = sPal + (uint32_t)i * palStride;
            a0[ii] = lds128(pb + j0); a1[ii] = lds128(pb + j0 + rS); a2[ii] = lds128(pb + j0 + rS2);
            if (NMAX > 1) { b0[ii] = lds128(pb + j1); b1[ii] = lds128(pb + j1 + rS); b2[ii] = lds128(pb + j1 + rS2); }
          }
        }
        // ---- phase 2: blend + transform + staging
#pragma unroll
        for (int ii = 0; ii < SB; ++ii) {
          const int i = i0 + ii;
          const float qx = px[MORPH ? i : 0], qy = py[MORPH ? i : 0], qz = pz[MORPH ? i : 0];
          float ox, oy, oz, nx = 0.f, ny = 0.f, nz = 0.f;
          {
            // ---- linear blend: M = sum_i w_i M_i (zero weights contribute exactly 0), then one packed mat-vec each
            float4 mA, mB, mC;
            if (NV == 0) { mA = a0[ii]; mB = a1[ii]; mC = a2[ii]; }          // every lane has the single weight 1.0
            else { mA = f4_scale(a0[ii], w0_2); mB = f4_scale(a1[ii], w0_2); mC = f4_scale(a2[ii], w0_2); }
            if (NMAX > 1) { mA = f4_fma(b0[ii], w1_2, mA); mB = f4_fma(b1[ii], w1_2, mB); mC = f4_fma(b2[ii], w1_2, mC); }
            if (NMAX > 2) {
              float4 c0, c1, c2;
              if (GPAL) { const float4* pal = reinterpret_cast<const float4*>(gpal[i]); c0 = pal[j2 / 16]; c1 = pal[(j2 + rS) / 16]; c2 = pal[(j2 + rS2) / 16]; }
              else { const uint32_t pb = sPal + (uint32_t)i * palStride; c0 = lds128(pb + j2); c1 = lds128(pb + j2 + rS); c2 = lds128(pb + j2 + rS2); }
              mA = f4_fma(c0, w2_2, mA); mB = f4_fma(c1, w2_2, mB); mC = f4_fma(c2, w2_2, mC);
            }
            if (NMAX > 3) {
              float4 c0, c1, c2;
              if (GPAL) { const float4* pal = reinterpret_cast<const float4*>(gpal[i]); c0 = pal[j3 / 16]; c1 = pal[(j3 + rS) / 16]; c2 = pal[(j3 + rS2) / 16]; }
              else { const uint32_t pb = sPal + (uint32_t)i * palStride; c0 = lds128(pb + j3); c1 = lds128(pb + j3 + rS); c2 = lds128(pb + j3 + rS2); }
              mA = f4_fma(c0, w3_2, mA); mB = f4_fma(c1, w3_2, mB); mC = f4_fma(c2, w3_2, mC);
            }
            // (ox,oy) = (m00,m10)*qx + (m01,m11)*qy + (m02,m12)*qz + (m03,m13);  oz = row 2 . (q,1)
            const float2 qx2 = make_float2(qx, qx), qy2 = make_float2(qy, qy), qz2 = make_float2(qz, qz);
            const float2 oxy = __ffma2_rn(make_float2(mA.x, mA.y), qx2,
                               __ffma2_rn(make_float2(mA.z, mA.w), qy2, __ffma2_rn(make_float2(mB.x, mB.y), qz2, make_float2(mB.z, mB.w))));
            ox = oxy.x; oy = oxy.y;
            oz = fmaf(mC.x, qx, fmaf(mC.y, qy, fmaf(mC.z, qz, mC.w)));
            if (NRM) {
              const float2 nxy = __ffma2_rn(make_float2(mA.x, mA.y), nx2,
                                 __ffma2_rn(make_float2(mA.z, mA.w), ny2, __fmul2_rn(make_float2(mB.x, mB.y), nz2)));
              nx = nxy.x; ny = nxy.y;
              nz = fmaf(mC.x, vnx, fmaf(mC.y, vny, mC.z * vnz));
            }
          }
          if (NRM) {
            const float l2 = fmaf(nx, nx, fmaf(ny, ny, nz * nz));
            const float rl = l2 > 0.f ? rsqrtf(l2) : 0.f;        // normalize(0) := 0 (SURVEY 8c edge case)
            nx *= rl; ny *= rl; nz *= rl;
          }
          if (SDEF) {
            if (isSdef) { ox = qx; oy = qy; oz = qz; nx = vnx; ny = vny; nz = vnz; }   // parked for the dense phase
          }
          if (BOUNDS) {
            if (valid && !isSdef) {
              bmin[i][0] = fminf(bmin[i][0], ox); bmax[i][0] = fmaxf(bmax[i][0], ox);
              bmin[i][1] = fminf(bmin[i][1], oy); bmax[i][1] = fmaxf(bmax[i][1], oy);
              bmin[i][2] = fminf(bmin[i][2], oz); bmax[i][2] = fmaxf(bmax[i][2], oz);
            }
          }
          {
            const uint32_t sa = stgLane + (uint32_t)i * kInstB;
            if (ILV) {
              sts128(sa, ox, oy, oz, nx);
              sts128(sa + 16u, ny, nz, v.uv.x, v.uv.y);
            } else {
              sts3(sa, ox, oy, oz);
              if (NRM) sts3(sa + kNrmO, nx, ny, nz);
              if (HULL) {
                // MMD inverted hull: pos' + n^' * edgeSize * 0.01 (engine.ts:458-461); SDEF owners park the offset instead
                if (SDEF && isSdef) sts3(sa + kHullO, v.edge, 0.f, 0.f);
                else sts3(sa + kHullO, fmaf(nx, v.edge, ox), fmaf(ny, v.edge, oy), fmaf(nz, v.edge, oz));
              }
            }
          }
        }
      };
      const bool unitW = __all_sync(0xffffffffu, w0 == 1.0f);
#pragma unroll
      for (int i0 = 0; i0 < I; i0 += SB) {
        switch (nmax) {
          case 1: if (unitW) body(IntC<0>{}, i0); else body(IntC<1>{}, i0); break;
          case 2: body(IntC<2>{}, i0); break;
          case 3: body(IntC<3>{}, i0); break;
          default: body(IntC<4>{}, i0); break;
        }
      }

      // ---- SDEF dense phase: pos' = R(q)·(p~ - C) + w0·(M0·c0) + w1·(M1·c1), n' = normalize(R(q)·n),
      //      q = slerp(quat(M0), quat(M1), w1).  quat(M) comes from the per-palette table (skin_quats_kernel: same
      //      math.ts:406-448 formula, evaluated once per bone instead of once per vertex-instance).
      if (SDEF) {
        if (sdCount) {                                            // warp-uniform
          __syncwarp();                                           // the owners' parked (p~, n) are visible
          for (uint32_t base = 0; base < sdCount * (uint32_t)I; base += 32u) {
            const uint32_t pp = base + (uint32_t)lane;
            const uint32_t r = pp / (uint32_t)I, i = pp - r * (uint32_t)I;
            uint32_t w = sdW;
            float4 t0 = sT0, t1 = sT1, t2 = sT2;
            if (base) {                                           // rare: more than 32/I SDEF vertices in this warp
              w = __shfl_sync(0xffffffffu, v.sdesc, (int)(r & 31u));
              if (r < sdCount) {
                const float4* e = prm.sdefTab + (size_t)(w & 0xFFFFFFu) * 3;
                t0 = __ldg(e); t1 = __ldg(e + 1); t2 = __ldg(e + 2);
              }
            }
            if (r < sdCount) {
              const uint32_t sa = stg + i * kInstB + (w >> 24) * kVtxB;
              const float qx = lds32(sa), qy = lds32(sa + 4), qz = lds32(sa + 8);
              const uint32_t jj = __float_as_uint(t2.w);
              const uint32_t r0p = jj & 0xFFFFu, r1p = jj >> 16;                 // palette rows of the two bones
              const float sw0 = t2.y, sw1 = t2.z;
              float4 a0, a1, a2, b0, b1, b2, qa, qb;
              if (GPAL) {
                const uint32_t k = kBase + min(i, nInst - 1);
                const uint32_t pidx = prm.inst2pal ? __ldg(prm.inst2pal + k) : k;
                const float4* pal = reinterpret_cast<const float4*>(prm.skin + (size_t)pidx * B * 12);
                a0 = pal[(r0p * pS) / 16]; a1 = pal[(r0p * pS + rS) / 16]; a2 = pal[(r0p * pS + rS2) / 16];
                b0 = pal[(r1p * pS) / 16]; b1 = pal[(r1p * pS + rS) / 16]; b2 = pal[(r1p * pS + rS2) / 16];
                qa = __ldg(prm.quat + (size_t)pidx * B + r0p); qb = __ldg(prm.quat + (size_t)pidx * B + r1p);
              } else {
                const uint32_t pb = sPal + i * palStride, qbase = sQuat + i * quatStride;
                a0 = lds128(pb + r0p * pS); a1 = lds128(pb + r0p * pS + rS); a2 = lds128(pb + r0p * pS + rS2);
                b0 = lds128(pb + r1p * pS); b1 = lds128(pb + r1p * pS + rS); b2 = lds128(pb + r1p * pS + rS2);
                qa = lds128(qbase + r0p * 16u); qb = lds128(qbase + r1p * 16u);
              }
              // un-pair the rows (kRowF4): r* = rows of M0, s* = rows of M1
              const float4 r0 = make_float4(a0.x, a0.z, a1.x, a1.z), r1 = make_float4(a0.y, a0.w, a1.y, a1.w), r2 = a2;
              const float4 s0 = make_float4(b0.x, b0.z, b1.x, b1.z), s1 = make_float4(b0.y, b0.w, b1.y, b1.w), s2 = b2;
              const Q4 q = quat_slerp(Q4{qa.x, qa.y, qa.z, qa.w}, Q4{qb.x, qb.y, qb.z, qb.w}, sw1);
              const float x2 = q.x + q.x, y2 = q.y + q.y, z2 = q.z + q.z;
              const float xx = q.x * x2, xy = q.x * y2, xz = q.x * z2, yy = q.y * y2, yz = q.y * z2, zz = q.z * z2;
              const float wx = q.w * x2, wy = q.w * y2, wz = q.w * z2;
              const float R00 = 1.f - (yy + zz), R01 = xy - wz, R02 = xz + wy;
              const float R10 = xy + wz, R11 = 1.f - (xx + zz), R12 = yz - wx;
              const float R20 = xz - wy, R21 = yz + wx, R22 = 1.f - (xx + yy);
              const float dx = qx - t0.x, dy = qy - t0.y, dz = qz - t0.z;
              const float c0x = t0.w, c0y = t1.x, c0z = t1.y, c1x = t1.z, c1y = t1.w, c1z = t2.x;
              const float e0x = fmaf(r0.x, c0x, fmaf(r0.y, c0y, fmaf(r0.z, c0z, r0.w)));
              const float e0y = fmaf(r1.x, c0x, fmaf(r1.y, c0y, fmaf(r1.z, c0z, r1.w)));
              const float e0z = fmaf(r2.x, c0x, fmaf(r2.y, c0y, fmaf(r2.z, c0z, r2.w)));
              const float e1x = fmaf(s0.x, c1x, fmaf(s0.y, c1y, fmaf(s0.z, c1z, s0.w)));
              const float e1y = fmaf(s1.x, c1x, fmaf(s1.y, c1y, fmaf(s1.z, c1z, s1.w)));
              const float e1z = fmaf(s2.x, c1x, fmaf(s2.y, c1y, fmaf(s2.z, c1z, s2.w)));
              const float ox = fmaf(R00, dx, fmaf(R01, dy, R02 * dz)) + sw0 * e0x + sw1 * e1x;
              const float oy = fmaf(R10, dx, fmaf(R11, dy, R12 * dz)) + sw0 * e0y + sw1 * e1y;
              const float oz = fmaf(R20, dx, fmaf(R21, dy, R22 * dz)) + sw0 * e0z + sw1 * e1z;
              sts3(sa, ox, oy, oz);
              if (NRM) {
                const float vx = lds32(sa + kNrmO), vy = lds32(sa + kNrmO + 4), vz = lds32(sa + kNrmO + 8);
                float nx = fmaf(R00, vx, fmaf(R01, vy, R02 * vz));
                float ny = fmaf(R10, vx, fmaf(R11, vy, R12 * vz));
                float nz = fmaf(R20, vx, fmaf(R21, vy, R22 * vz));
                const float l2 = fmaf(nx, nx, fmaf(ny, ny, nz * nz));
                const float rl = l2 > 0.f ? rsqrtf(l2) : 0.f;
                nx *= rl; ny *= rl; nz *= rl;
                sts3(sa + kNrmO, nx, ny, nz);
                if (HULL) {
                  const float eg = lds32(sa + kHullO);              // parked by the owner lane
                  sts3(sa + kHullO, fmaf(nx, eg, ox), fmaf(ny, eg, oy), fmaf(nz, eg, oz));
                }
              }
              if (BOUNDS) {
#pragma unroll
                for (int ii = 0; ii < I; ++ii)
                  if ((uint32_t)ii == i) {
                    bmin[ii][0] = fminf(bmin[ii][0], ox); bmax[ii][0] = fmaxf(bmax[ii][0], ox);
                    bmin[ii][1] = fminf(bmin[ii][1], oy); bmax[ii][1] = fmaxf(bmax[ii][1], oy);
                    bmin[ii][2] = fminf(bmin[ii][2], oz); bmax[ii][2] = fmaxf(bmax[ii][2], oz);
                  }
              }
            }
          }
        }
      }

      // ---- drain: this warp's 32 vertices x I instances leave through the TMA (one elected lane)
      fence_proxy_async();
      __syncwarp();
      if (elect_one()) {
        if (nAligned) {
          float* dst = outItem + (size_t)warpVtx0 * (kVtxB / 4u);
#pragma unroll
          for (int i = 0; i < I; ++i) {
            if ((uint32_t)i < nInst) {
              bulk_s2g(dst, stg + (uint32_t)i * kInstB, nAligned * kVtxB, polFirst);
              if (NRM && !ILV) bulk_s2g(dst + prm.nrmOffF, stg + (uint32_t)i * kInstB + kPlaneB, nAligned * 12u, polFirst);
              if (HULL) bulk_s2g(dst + prm.hullOffF, stg + (uint32_t)i * kInstB + kHullO, nAligned * 12u, polFirst);
            }
            dst += prm.instStrideF;
          }
        }
        bulk_commit();
      }
      if (nAligned != nWarpVerts) {
        // cold path (at most one warp of the whole mesh): the ragged tail (< 4 vertices) leaves with plain stores
        const uint32_t nf = (nWarpVerts - nAligned) * 3u;
        if ((uint32_t)lane < nf) {
          const uint32_t o = nAligned * 3u + (uint32_t)lane;
          for (uint32_t i = 0; i < nInst; ++i) {
            float* dst = outItem + (size_t)i * prm.instStrideF + (size_t)warpVtx0 * 3 + o;
            st_cs(dst, lds32(stg + i * kInstB + o * 4u));
            if (NRM) st_cs(dst + prm.nrmOffF, lds32(stg + i * kInstB + kPlaneB + o * 4u));
            if (HULL) st_cs(dst + prm.hullOffF, lds32(stg + i * kInstB + kHullO + o * 4u));
          }
        }
        __syncwarp();
      }
      if (kStageBufs > 1) sbuf ^= 1u;
      if (MORPH) {
        if (t + NT / kTile < tile1) load_pf(cur);                  // cur's record has landed by now
      }
    }

    if (BOUNDS) {
      // per-item reduction: warp shuffle, then one atomic per warp per bound (ordered-int encoding)
#pragma unroll
      for (int i = 0; i < I; ++i) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float lo = bmin[i][c], hi = bmax[i][c];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
          }
          if (lane == 0 && (uint32_t)i < nInst) {
            int* bp = reinterpret_cast<int*>(prm.bounds) + (size_t)(kBase + i) * 6;
            atomicMin(bp + c, f2ord(lo));
            atomicMax(bp + 3 + c, f2ord(hi));
          }
        }
      }
    }
    __syncthreads();   // everyone is done with this item's palettes before they are overwritten
  }

  // staging must outlive the TMA reads.  Bulk groups are tracked per THREAD and elect.sync only promises a deterministic
  // leader, not lane 0: every lane waits (a lane that committed no group returns at once)
  bulk_wait0();
}

}  // namespace rz
