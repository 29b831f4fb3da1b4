// deform_kernel.cuh — the fused morph + BDEF1/2/4/SDEF skinning kernel for sm_100a.
//
// Replaces the per-vertex blend the reference runs inside three WGSL vertex shaders
// (engine.ts:245-276 main, 431-463 outline, 692-715 depth-only) and materialises the
// skinned stream once per frame for K independent character instances.
//
// Work decomposition (B200: 148 SMs, 227 KB smem, HBM-bound on the 24 B/vertex write):
//   work item = (group of I instances) x (chunk of vertex tiles); persistent CTAs pull
//   items from an atomic counter.  Per item the I bone palettes (B x 48 B each, 3x4
//   row-major skin matrices) are staged into shared memory with one cp.async.bulk
//   (TMA bulk copy, mbarrier complete_tx) per instance; each thread then keeps ONE
//   vertex (pos, normal, 4 joints, 4 weights = 40 B, float4-vectorised, L2-resident)
//   in registers and evaluates it for the I instances, so the static mesh is read once
//   per I outputs.  Results are written either directly (st.global.cs) or staged in
//   shared memory in final layout and drained with cp.async.bulk shared->global
//   (TMA bulk store, double-buffered, evict-first), which keeps the LSU free for the
//   palette gathers.  No tensor cores: the work is a gather of 3x4 mat-vecs.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rz {

constexpr int kTile = 256;          // vertices per preprocessing tile (permutation unit)
constexpr int kRowF4 = 3;           // float4 per bone (3x4 row-major)

// meta word (rec1.w) layout
constexpr uint32_t kMetaSlotMask = 0x3FFu;   // bits 0-9  : slot inside the tile
constexpr int      kMetaNinfShift = 10;      // bits 10-12: 1 + index of last non-zero weight
constexpr uint32_t kMetaValid = 1u << 13;
constexpr uint32_t kMetaMorph = 1u << 14;
constexpr uint32_t kMetaSdef  = 1u << 15;

enum : int { FEAT_MORPH = 1, FEAT_SDEF = 2, FEAT_BOUNDS = 4, FEAT_GPAL = 8, FEAT_NONRM = 16 };

struct DeformParams {
  const float4* __restrict__ rec0;     // [Vp] px,py,pz, weights(u8x4 bits)
  const float4* __restrict__ rec1;     // [Vp] nx,ny,nz, meta bits
  const uint2*  __restrict__ joints;   // [Vp] 4 x u16
  const uint2*  __restrict__ mrange;   // [Vp] (first entry, count) into ments
  const float4* __restrict__ ments;    // [nnz] dx,dy,dz, morph id bits
  const uint32_t* __restrict__ sdefIdx;// [Vp] index into sdefTab (valid when kMetaSdef)
  const float4* __restrict__ sdefTab;  // [nSdef*3]: (C.xyz,c0.x) (c0.yz,c1.xy) (c1.z,0,0,0)
  const float*  __restrict__ skin;     // [P][B][12]
  const uint32_t* __restrict__ inst2pal; // [K] or nullptr (identity)
  const float*  __restrict__ mweights; // dense [K][Mpad]
  float* __restrict__ out;
  float* __restrict__ bounds;          // [K][6] as ordered ints, or nullptr
  unsigned long long instStrideF;      // floats between instances
  unsigned long long nrmOffF;          // floats from pos plane to normal plane
  uint32_t V, B, nTiles, K0, Kcount, Mpad;
  uint32_t nGroups, nChunks, tilesPerChunk;
  uint32_t* counter;
};

// ------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}"
      ::"r"(smem_u32(bar)), "r"(parity) : "memory");
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
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      ::"r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
// shared -> global bulk store (TMA, 1-D), tracked by bulk groups
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src_smem, uint32_t bytes, uint64_t pol) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
               ::"l"(dst), "r"(smem_u32(src_smem)), "r"(bytes), "l"(pol) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ float4 ldg_el(const float4* p, uint64_t pol) {   // read-only, keep in L2 (mesh is re-read by every group)
  float4 r;
  asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ uint2 ldg_el(const uint2* p, uint64_t pol) {
  uint2 r;
  asm volatile("ld.global.nc.L2::cache_hint.v2.u32 {%0,%1}, [%2], %3;" : "=r"(r.x), "=r"(r.y) : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ void st_cs(float* p, float v) { asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory"); }

__device__ __forceinline__ float4 f4_scale(float4 a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
__device__ __forceinline__ float4 f4_fma(float4 a, float s, float4 c) {
  return make_float4(fmaf(a.x, s, c.x), fmaf(a.y, s, c.y), fmaf(a.z, s, c.z), fmaf(a.w, s, c.w));
}

// order-preserving float <-> int for atomic min/max
__device__ __forceinline__ int f2ord(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7FFFFFFF; }

// ---- SDEF helpers: same formulas as the reference's math (math.ts:406-448 toQuatFromArray, 156-189 slerp,
//      352-384 fromQuat) applied to the rotation part of two skin matrices.
struct Q4 { float x, y, z, w; };
__device__ __forceinline__ Q4 quat_from_rows(float4 r0, float4 r1, float4 r2) {
  // r? are matrix rows: m[row][col]; reference names mRC
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
__device__ __forceinline__ Q4 quat_slerp(Q4 a, Q4 b, float t) {
  float c = a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
  if (c < 0.f) { c = -c; b.x = -b.x; b.y = -b.y; b.z = -b.z; b.w = -b.w; }
  if (c > 0.9995f) {
    const float x = a.x + t * (b.x - a.x), y = a.y + t * (b.y - a.y), z = a.z + t * (b.z - a.z), w = a.w + t * (b.w - a.w);
    const float inv = 1.0f / sqrtf(x * x + y * y + z * z + w * w);
    return Q4{x * inv, y * inv, z * inv, w * inv};
  }
  const float th0 = acosf(c), s = sinf(th0), th = th0 * t;
  const float s0 = sinf(th0 - th) / s, s1 = sinf(th) / s;
  return Q4{s0 * a.x + s1 * b.x, s0 * a.y + s1 * b.y, s0 * a.z + s1 * b.z, s0 * a.w + s1 * b.w};
}

// ------------------------------------------------------------------ the kernel
// I      : instances evaluated per vertex pass (register-tiled)
// NT     : threads per CTA (256 or 512); one vertex per thread per pass
// STAGED : smem-staged TMA bulk stores instead of direct register stores
// FEAT   : FEAT_* bit set
template <int I, int NT, bool STAGED, int FEAT>
__global__ void __launch_bounds__(NT, 1) deform_kernel(const DeformParams prm) {
  constexpr bool MORPH = (FEAT & FEAT_MORPH) != 0;
  constexpr bool SDEF = (FEAT & FEAT_SDEF) != 0;
  constexpr bool BOUNDS = (FEAT & FEAT_BOUNDS) != 0;
  constexpr bool GPAL = (FEAT & FEAT_GPAL) != 0;      // palette too large for smem: gather from global/L1
  constexpr bool NRM = (FEAT & FEAT_NONRM) == 0;
  constexpr int PLANES = NRM ? 2 : 1;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  volatile uint32_t* s_item = reinterpret_cast<volatile uint32_t*>(smem_raw + 8);
  unsigned char* sp = smem_raw + 16;
  float4* s_pal = reinterpret_cast<float4*>(sp);
  if (!GPAL) sp += (size_t)I * prm.B * kRowF4 * sizeof(float4);
  float* s_mw = reinterpret_cast<float*>(sp);
  if (MORPH) sp += (size_t)I * prm.Mpad * sizeof(float);
  float* s_stage = reinterpret_cast<float*>(sp);      // [2][I][PLANES][NT*3]
  constexpr int kStagePlaneF = NT * 3;
  constexpr int kStageBufF = I * PLANES * kStagePlaneF;

  const int tid = threadIdx.x;
  const uint32_t B = prm.B;

  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncthreads();

  uint32_t phase = 0;
  uint32_t stageBuf = 0;
  const uint64_t polFirst = policy_evict_first();
  const uint64_t polLast = policy_evict_last();

  for (;;) {
    if (tid == 0) *s_item = atomicAdd(prm.counter, 1u);
    __syncthreads();
    const uint32_t item = *s_item;
    if (item >= prm.nGroups * prm.nChunks) break;
    const uint32_t g = item / prm.nChunks;
    const uint32_t chunk = item - g * prm.nChunks;
    const uint32_t kBase = prm.K0 + g * I;                       // first instance of this group
    const uint32_t nInst = min((uint32_t)I, prm.K0 + prm.Kcount - kBase);
    const uint32_t tile0 = chunk * prm.tilesPerChunk;
    const uint32_t tile1 = min(prm.nTiles, tile0 + prm.tilesPerChunk);

    // ---- stage the palettes (+ morph weights) of the I instances: TMA bulk copies on one mbarrier
    const float* gpal[I];
#pragma unroll
    for (int i = 0; i < I; ++i) {
      const uint32_t k = kBase + min((uint32_t)i, nInst - 1);    // clamp: unused lanes of a partial group recompute the last one
      const uint32_t pidx = prm.inst2pal ? __ldg(prm.inst2pal + k) : k;
      gpal[i] = prm.skin + (size_t)pidx * B * 12;
    }
    if (!GPAL || MORPH) {
      if (tid == 0) {
        const uint32_t palBytes = GPAL ? 0u : B * 48u;
        const uint32_t mwBytes = MORPH ? prm.Mpad * 4u : 0u;
        mbar_expect_tx(bar, (uint32_t)I * (palBytes + mwBytes));
#pragma unroll
        for (int i = 0; i < I; ++i) {
          if (!GPAL) bulk_g2s(s_pal + (size_t)i * B * kRowF4, gpal[i], palBytes, bar, polLast);
          if (MORPH) {
            const uint32_t k = kBase + min((uint32_t)i, nInst - 1);
            bulk_g2s(s_mw + (size_t)i * prm.Mpad, prm.mweights + (size_t)k * prm.Mpad, mwBytes, bar, polLast);
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

    bool waited = false;
    for (uint32_t t = tile0; t < tile1; t += NT / kTile) {
      const uint32_t p = t * kTile + tid;                        // processing-order index
      const uint32_t tileOfThread = t + tid / kTile;
      const bool inRange = tileOfThread < tile1;                 // second half of a 512-thread pass may fall off the chunk
      float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0;
      uint2 jj = make_uint2(0u, 0u);
      if (inRange) {
        r0 = ldg_el(prm.rec0 + p, polLast);
        r1 = ldg_el(prm.rec1 + p, polLast);
        jj = ldg_el(prm.joints + p, polLast);
      } else {
        r0.w = __uint_as_float(255u);
      }
      const uint32_t meta = __float_as_uint(r1.w);
      const uint32_t wb = __float_as_uint(r0.w);
      const bool valid = inRange && (meta & kMetaValid);
      const uint32_t slot = meta & kMetaSlotMask;
      const int ninf = inRange ? (int)((meta >> kMetaNinfShift) & 7u) : 1;

      // weights exactly as the vertex shader derives them (engine.ts:255-258): unorm8 -> f32, renormalise
      float w0 = (float)(wb & 255u) / 255.0f, w1 = (float)((wb >> 8) & 255u) / 255.0f;
      float w2 = (float)((wb >> 16) & 255u) / 255.0f, w3 = (float)(wb >> 24) / 255.0f;
      {
        const float wsum = w0 + w1 + w2 + w3;
        if (wsum > 0.0001f) {
          const float inv = 1.0f / wsum;
          w0 *= inv; w1 *= inv; w2 *= inv; w3 *= inv;
        } else {
          w0 = 1.f; w1 = 0.f; w2 = 0.f; w3 = 0.f;
        }
      }
      const uint32_t j0 = (jj.x & 0xFFFFu) * kRowF4, j1 = (jj.x >> 16) * kRowF4;
      const uint32_t j2 = (jj.y & 0xFFFFu) * kRowF4, j3 = (jj.y >> 16) * kRowF4;

      // ---- morph accumulation: p~ = p + sum_m w[k][m] * delta_m[v]   (model space, before skinning)
      float px[I], py[I], pz[I];
#pragma unroll
      for (int i = 0; i < I; ++i) { px[i] = r0.x; py[i] = r0.y; pz[i] = r0.z; }

      if (!waited) {   // palettes / weights must have landed before the first gather of this item
        if (!GPAL || MORPH) mbar_wait(bar, phase);
        phase ^= 1u;
        waited = true;
      }

      if (MORPH) {
        if (meta & kMetaMorph) {
          const uint2 mr = __ldg(prm.mrange + p);
          for (uint32_t e = mr.x; e < mr.x + mr.y; ++e) {
            const float4 d = __ldg(prm.ments + e);
            const uint32_t m = __float_as_uint(d.w);
#pragma unroll
            for (int i = 0; i < I; ++i) {
              const float wgt = s_mw[(size_t)i * prm.Mpad + m];
              px[i] = fmaf(wgt, d.x, px[i]);
              py[i] = fmaf(wgt, d.y, py[i]);
              pz[i] = fmaf(wgt, d.z, pz[i]);
            }
          }
        }
      }

      float ox[I], oy[I], oz[I], nx[I], ny[I], nz[I];
      const bool isSdef = SDEF && (meta & kMetaSdef);
      if (!isSdef) {
#pragma unroll
        for (int i = 0; i < I; ++i) {
          const float4* pal = GPAL ? reinterpret_cast<const float4*>(gpal[i]) : (s_pal + (size_t)i * B * kRowF4);
          float4 m0 = f4_scale(pal[j0], w0), m1 = f4_scale(pal[j0 + 1], w0), m2 = f4_scale(pal[j0 + 2], w0);
          if (ninf > 1) {
            m0 = f4_fma(pal[j1], w1, m0); m1 = f4_fma(pal[j1 + 1], w1, m1); m2 = f4_fma(pal[j1 + 2], w1, m2);
            if (ninf > 2) {
              m0 = f4_fma(pal[j2], w2, m0); m1 = f4_fma(pal[j2 + 1], w2, m1); m2 = f4_fma(pal[j2 + 2], w2, m2);
              if (ninf > 3) {
                m0 = f4_fma(pal[j3], w3, m0); m1 = f4_fma(pal[j3 + 1], w3, m1); m2 = f4_fma(pal[j3 + 2], w3, m2);
              }
            }
          }
          ox[i] = fmaf(m0.x, px[i], fmaf(m0.y, py[i], fmaf(m0.z, pz[i], m0.w)));
          oy[i] = fmaf(m1.x, px[i], fmaf(m1.y, py[i], fmaf(m1.z, pz[i], m1.w)));
          oz[i] = fmaf(m2.x, px[i], fmaf(m2.y, py[i], fmaf(m2.z, pz[i], m2.w)));
          if (NRM) {
            const float ax = fmaf(m0.x, r1.x, fmaf(m0.y, r1.y, m0.z * r1.z));
            const float ay = fmaf(m1.x, r1.x, fmaf(m1.y, r1.y, m1.z * r1.z));
            const float az = fmaf(m2.x, r1.x, fmaf(m2.y, r1.y, m2.z * r1.z));
            const float l2 = fmaf(ax, ax, fmaf(ay, ay, az * az));
            const float rl = l2 > 0.f ? rsqrtf(l2) : 0.f;        // normalize(0) := 0 (oracle convention)
            nx[i] = ax * rl; ny[i] = ay * rl; nz[i] = az * rl;
          }
        }
      } else {
        // ---- SDEF: spherical blend of the two bone rotations around C
        const uint32_t si = __ldg(prm.sdefIdx + p);
        const float4 t0 = __ldg(prm.sdefTab + (size_t)si * 3), t1 = __ldg(prm.sdefTab + (size_t)si * 3 + 1),
                     t2 = __ldg(prm.sdefTab + (size_t)si * 3 + 2);
        const float Cx = t0.x, Cy = t0.y, Cz = t0.z;
        const float c0x = t0.w, c0y = t1.x, c0z = t1.y, c1x = t1.z, c1y = t1.w, c1z = t2.x;
#pragma unroll
        for (int i = 0; i < I; ++i) {
          const float4* pal = GPAL ? reinterpret_cast<const float4*>(gpal[i]) : (s_pal + (size_t)i * B * kRowF4);
          const float4 a0 = pal[j0], a1 = pal[j0 + 1], a2 = pal[j0 + 2];
          const float4 b0 = pal[j1], b1 = pal[j1 + 1], b2 = pal[j1 + 2];
          const Q4 q = quat_slerp(quat_from_rows(a0, a1, a2), quat_from_rows(b0, b1, b2), w1);
          const float x2 = q.x + q.x, y2 = q.y + q.y, z2 = q.z + q.z;
          const float xx = q.x * x2, xy = q.x * y2, xz = q.x * z2, yy = q.y * y2, yz = q.y * z2, zz = q.z * z2;
          const float wx = q.w * x2, wy = q.w * y2, wz = q.w * z2;
          const float R00 = 1.f - (yy + zz), R01 = xy - wz, R02 = xz + wy;
          const float R10 = xy + wz, R11 = 1.f - (xx + zz), R12 = yz - wx;
          const float R20 = xz - wy, R21 = yz + wx, R22 = 1.f - (xx + yy);
          const float dx = px[i] - Cx, dy = py[i] - Cy, dz = pz[i] - Cz;
          const float e0x = fmaf(a0.x, c0x, fmaf(a0.y, c0y, fmaf(a0.z, c0z, a0.w)));
          const float e0y = fmaf(a1.x, c0x, fmaf(a1.y, c0y, fmaf(a1.z, c0z, a1.w)));
          const float e0z = fmaf(a2.x, c0x, fmaf(a2.y, c0y, fmaf(a2.z, c0z, a2.w)));
          const float e1x = fmaf(b0.x, c1x, fmaf(b0.y, c1y, fmaf(b0.z, c1z, b0.w)));
          const float e1y = fmaf(b1.x, c1x, fmaf(b1.y, c1y, fmaf(b1.z, c1z, b1.w)));
          const float e1z = fmaf(b2.x, c1x, fmaf(b2.y, c1y, fmaf(b2.z, c1z, b2.w)));
          ox[i] = fmaf(R00, dx, fmaf(R01, dy, R02 * dz)) + w0 * e0x + w1 * e1x;
          oy[i] = fmaf(R10, dx, fmaf(R11, dy, R12 * dz)) + w0 * e0y + w1 * e1y;
          oz[i] = fmaf(R20, dx, fmaf(R21, dy, R22 * dz)) + w0 * e0z + w1 * e1z;
          if (NRM) {
            const float ax = fmaf(R00, r1.x, fmaf(R01, r1.y, R02 * r1.z));
            const float ay = fmaf(R10, r1.x, fmaf(R11, r1.y, R12 * r1.z));
            const float az = fmaf(R20, r1.x, fmaf(R21, r1.y, R22 * r1.z));
            const float l2 = fmaf(ax, ax, fmaf(ay, ay, az * az));
            const float rl = l2 > 0.f ? rsqrtf(l2) : 0.f;
            nx[i] = ax * rl; ny[i] = ay * rl; nz[i] = az * rl;
          }
        }
      }

      if (BOUNDS && valid) {
#pragma unroll
        for (int i = 0; i < I; ++i) {
          bmin[i][0] = fminf(bmin[i][0], ox[i]); bmax[i][0] = fmaxf(bmax[i][0], ox[i]);
          bmin[i][1] = fminf(bmin[i][1], oy[i]); bmax[i][1] = fmaxf(bmax[i][1], oy[i]);
          bmin[i][2] = fminf(bmin[i][2], oz[i]); bmax[i][2] = fmaxf(bmax[i][2], oz[i]);
        }
      }

      // ---- write out
      const uint32_t passBase = t * kTile;                        // first vertex id of this pass
      const uint32_t vertsHere = min((uint32_t)NT, prm.V - passBase);
      const uint32_t vertsInChunk = min(vertsHere, (tile1 - t) * (uint32_t)kTile);
      const uint32_t vid = tileOfThread * kTile + slot;
      const bool bulkOk = STAGED && ((vertsInChunk * 3u) & 3u) == 0u;
      if (bulkOk) {
        float* st = s_stage + (size_t)stageBuf * kStageBufF;
        const int so = ((tid / kTile) * kTile + (int)slot) * 3;
#pragma unroll
        for (int i = 0; i < I; ++i) {
          float* ps = st + (size_t)(i * PLANES) * kStagePlaneF + so;
          ps[0] = ox[i]; ps[1] = oy[i]; ps[2] = oz[i];
          if (NRM) {
            float* ns = ps + kStagePlaneF;
            ns[0] = nx[i]; ns[1] = ny[i]; ns[2] = nz[i];
          }
        }
        fence_proxy_async();
        if (tid == 0) bulk_wait_read0();        // the previous pass' bulk stores have released the other buffer
        __syncthreads();
        if (tid == 0) {
          const uint32_t bytes = vertsInChunk * 12u;
#pragma unroll
          for (int i = 0; i < I; ++i) {
            if ((uint32_t)i < nInst) {
              float* dst = prm.out + (size_t)(kBase + i) * prm.instStrideF + (size_t)passBase * 3;
              bulk_s2g(dst, st + (size_t)(i * PLANES) * kStagePlaneF, bytes, polFirst);
              if (NRM) bulk_s2g(dst + prm.nrmOffF, st + (size_t)(i * PLANES + 1) * kStagePlaneF, bytes, polFirst);
            }
          }
          bulk_commit();
        }
        stageBuf ^= 1u;
      } else if (valid) {
#pragma unroll
        for (int i = 0; i < I; ++i) {
          if ((uint32_t)i < nInst) {
            float* dst = prm.out + (size_t)(kBase + i) * prm.instStrideF + (size_t)vid * 3;
            st_cs(dst, ox[i]); st_cs(dst + 1, oy[i]); st_cs(dst + 2, oz[i]);
            if (NRM) {
              float* dn = dst + prm.nrmOffF;
              st_cs(dn, nx[i]); st_cs(dn + 1, ny[i]); st_cs(dn + 2, nz[i]);
            }
          }
        }
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
          if ((tid & 31) == 0 && (uint32_t)i < nInst) {
            int* bp = reinterpret_cast<int*>(prm.bounds) + (size_t)(kBase + i) * 6;
            atomicMin(bp + c, f2ord(lo));
            atomicMax(bp + 3 + c, f2ord(hi));
          }
        }
      }
    }
    __syncthreads();   // everyone is done with this item's palettes before they are overwritten
  }

  if (STAGED && tid == 0) bulk_wait0();
}

}  // namespace rz
