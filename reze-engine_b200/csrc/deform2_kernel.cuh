// deform2_kernel.cuh — the plain BDEF1/2/4 path with TWO vertices per lane (sm_100a).
//
// Same job as deform_kernel<.., FEAT = 0> (deform_kernel.cuh: the vertex-shader blend engine.ts:245-276, materialised
// once per frame for K instances), restructured around what bounds that kernel on B200: the shared-memory pipe.  Per
// 32 vertex-instances the one-vertex-per-lane kernel spends ~20 of its ~31 LSU cycles on palette-row gathers (one
// LDS.128 x3 per influence slot per warp) and ~100 issue slots, a third of them per-pass overhead; HBM needs 33 cycles for
// the same 768 bytes at the measured peak, so neither pipe has slack and the kernel follows the SM clock (1.81 ms at
// 1965 MHz, 2.05 ms once the power manager settles at ~1750 MHz).  Here a lane evaluates two vertices that share ONE
// list of <= 4 palette rows (lane_plan2.h pairs them at load time inside windows of 64 consecutive vertices, and packs
// aligned lane pairs onto the same rows): a warp covers 64 vertices with the gather instructions, the record loads, the
// unpacking and the TMA issue of 32.  Blend weights are per vertex (zero where a row is not the vertex' own bone), so
// every vertex still sums exactly its own (bone, weight) terms.
//
// Work item = (group of I instances) x (chunk of vertex groups); persistent CTAs pull items from an atomic counter; the
// I palettes are staged with TMA bulk copies; every warp walks its own vertex groups with no CTA-wide barrier inside an
// item.  Results are staged in warp-private shared memory laid out like 64 vertices of the output planes and leave as
// cp.async.bulk shared->global stores, one commit group per sub-batch of SB instances; the sub-batches rotate through NBUF
// small staging buffers, so the stores of one sub-batch drain while the next one is computed.
#pragma once
#include "deform_kernel.cuh"

namespace rz {

constexpr int kRec2Planes = 7;   // float4 planes of the two-vertex lane record (SoA: plane q of lane L at rec[q*lanes + L])
// q0 = (pA.xyz, wA0)  q1 = (nA.xyz, wA1)  q2 = (pB.xyz, wB0)  q3 = (nB.xyz, wB1)  q4 = (wA2, wA3, wB2, wB3)
// q5 = (row0 | row1 << 16,  row2 | row3 << 16,  meta,  first output vertex of the group)   [bit patterns]
// q6 = (uA, vA, uB, vB) texture coordinates (interleaved output) or (edgeA, edgeB, 0, 0) outline offsets (hull output);
//      only read by those two layouts
// meta: bits 0-5 slotA, 6-11 slotB (output position inside the group's 64-vertex staging), 12-14 slots used by this lane,
//       15 vertex A real, 16 vertex B real, 17-23 vertices the group covers (0..64)
constexpr int kM2SlotB = 6, kM2N = 12, kM2Cnt = 17;
constexpr uint32_t kM2HasA = 1u << 15, kM2HasB = 1u << 16;

// output layouts of the two-vertex kernel (the feature sets without morphs / SDEF; the AABB with the planar layout only)
enum : int { OUT2_PLANAR = 0,   // position plane + normal plane                        (engine.ts:245-276)
             OUT2_NONRM = 1,    // positions only: the depth-only blend                 (engine.ts:692-715)
             OUT2_HULL = 2,     // + outline hull plane pos' + n^' * edge               (engine.ts:431-463, 458-461)
             OUT2_ILV = 3,      // one interleaved 32-byte [pos, nrm, uv] stream        (engine.ts:340-347)
             OUT2_BOUNDS = 4 }; // planar + per-instance AABB of the skinned positions  

struct Deform2Params {
  const float4* __restrict__ rec;        // [6][lanes]
  const float* __restrict__ skin;        // [P][B][12], pair layout (deform_kernel.cuh kRowF4)
  const uint32_t* __restrict__ inst2pal; // [K] or nullptr (identity)
  float* __restrict__ out;
  float* __restrict__ bounds;            // [K][6] as ordered ints (OUT2_BOUNDS), else unused
  unsigned long long instStrideF;        // floats between instances
  unsigned long long nrmOffF;            // floats from the position plane to the normal plane
  unsigned long long hullOffF;           // floats from the position plane to the outline-hull plane (OUT2_HULL)
  uint32_t lanes;                        // vertex groups * 32
  uint32_t nVG;                          // vertex groups (32 lanes each)
  uint32_t V, B, K0, Kcount;
  // work items: the first nCoarse instance groups are ONE item each (all vertex groups), every later instance group is cut
  // into nChunks items of vgPerChunk vertex groups -- long items first, short ones to level the tail (rze_b200.cu pick_items)
  uint32_t nGroups, nCoarse, nChunks, vgPerChunk, nItems;
  uint32_t* counter;
};

struct Rec2 { float4 q0, q1, q2, q3, q4, q5, q6; };

// I: instances per palette stage; NT: threads per CTA; MINB: CTAs/SM the registers are sized for;
// SB: instances per sub-batch (gathers in flight together, one store commit group); divides I
// NBUF: staging buffers per warp, each holding ONE sub-batch; sub-batches rotate through them, so the stores of up to
//       NBUF-1 sub-batches drain while the next one is computed -- and the staging footprint no longer grows with I
// OUT : output layout (OUT2_*)
template <int I, int NT, int MINB, int SB, int NBUF, int OUT = OUT2_PLANAR>
__global__ void __launch_bounds__(NT, MINB) deform2_kernel(const Deform2Params prm) {
  static_assert(I % SB == 0, "sub-batch must divide the group");
  constexpr int W = NT / 32;
  constexpr int NSB = I / SB;                          // sub-batches = store commit groups per pass
  constexpr bool ILV = OUT == OUT2_ILV, HULL = OUT == OUT2_HULL, NRM = OUT != OUT2_NONRM, BOUNDS = OUT == OUT2_BOUNDS;
  constexpr uint32_t kVtxB = ILV ? 32u : 12u;          // bytes per vertex of a staging plane
  constexpr uint32_t kPlaneB = 64u * kVtxB;            // one staging plane of a warp: 64 vertices
  constexpr uint32_t kInstB = (ILV ? 1u : (NRM ? 2u : 1u) + (HULL ? 1u : 0u)) * kPlaneB;   // the planes of one instance
  constexpr uint32_t kHullO = 2u * kPlaneB;            // hull position relative to the position (OUT2_HULL)
  constexpr uint32_t kBufB = (uint32_t)SB * kInstB;    // one staging buffer: a sub-batch

  extern __shared__ __align__(128) unsigned char smem_raw[];
  const uint32_t sbase = smem_u32(smem_raw);
  Ctrl* ctrl = reinterpret_cast<Ctrl*>(smem_raw);
  const uint32_t palBar = sbase + (uint32_t)offsetof(Ctrl, palBar);
  const uint32_t B = prm.B;
  const uint32_t sPal = sbase + kCtrlBytes;
  const uint32_t palBytes = B * 48u;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform by construction (TMA operands in uniform registers)
  const uint32_t sStageW = sPal + (uint32_t)I * palBytes + (uint32_t)warp * (NBUF * kBufB);

  if (tid == 0) {
    mbar_init(palBar, 1);
    fence_mbar_init();
  }
  __syncthreads();

  uint32_t palPhase = 0;
  uint32_t sbuf = 0;                                   // staging buffer the next sub-batch fills
  const uint64_t polFirst = policy_evict_first();
  const uint64_t polLast = policy_evict_last();

  for (;;) {
    if (tid == 0) ctrl->item = atomicAdd(prm.counter, 1u);
    __syncthreads();
    const uint32_t item = *reinterpret_cast<volatile uint32_t*>(&ctrl->item);
    if (item >= prm.nItems) break;
    uint32_t g = item, vg0 = 0, vg1 = prm.nVG;
    if (item >= prm.nCoarse) {
      const uint32_t r = item - prm.nCoarse, gq = r / prm.nChunks;
      g = prm.nCoarse + gq;
      vg0 = (r - gq * prm.nChunks) * prm.vgPerChunk;
      vg1 = min(prm.nVG, vg0 + prm.vgPerChunk);
    }
    const uint32_t kBase = prm.K0 + g * I;
    const uint32_t nInst = min((uint32_t)I, prm.K0 + prm.Kcount - kBase);
    float* const outItem = prm.out + (size_t)kBase * prm.instStrideF;
    // AABB: running min / max of the lane's vertices per instance, in registers over the whole item (the sub-batch loop is
    // unrolled for this layout so that the instance index is static)
    float bmin[BOUNDS ? I : 1][3], bmax[BOUNDS ? I : 1][3];
    if (BOUNDS) {
#pragma unroll
      for (int i = 0; i < I; ++i)
#pragma unroll
        for (int c = 0; c < 3; ++c) { bmin[i][c] = 3.4e38f; bmax[i][c] = -3.4e38f; }
    }

    // ---- stage the I palettes: TMA bulk copies on one mbarrier (a partial group re-stages its last palette, never stored)
    if (tid == 0) {
      mbar_expect_tx(palBar, (uint32_t)I * palBytes);
#pragma unroll
      for (int i = 0; i < I; ++i) {
        const uint32_t k = kBase + min((uint32_t)i, nInst - 1);
        const uint32_t pidx = prm.inst2pal ? __ldg(prm.inst2pal + k) : k;
        bulk_g2s(sPal + (uint32_t)i * palBytes, prm.skin + (size_t)pidx * B * 12, palBytes, palBar, polLast);
      }
    }

    auto load_rec = [&](uint32_t vg) -> Rec2 {
      Rec2 r;
      const float4* p = prm.rec + (size_t)vg * 32u + (uint32_t)lane;
      r.q0 = ldg_el(p, polLast);
      r.q1 = ldg_el(p + prm.lanes, polLast);
      r.q2 = ldg_el(p + 2u * (size_t)prm.lanes, polLast);
      r.q3 = ldg_el(p + 3u * (size_t)prm.lanes, polLast);
      r.q4 = ldg_el(p + 4u * (size_t)prm.lanes, polLast);
      r.q5 = ldg_el(p + 5u * (size_t)prm.lanes, polLast);
      r.q6 = (ILV || HULL) ? ldg_el(p + 6u * (size_t)prm.lanes, polLast) : make_float4(0.f, 0.f, 0.f, 0.f);
      return r;
    };
    uint32_t vg = vg0 + (uint32_t)warp;
    Rec2 cur;
    if (vg < vg1) cur = load_rec(vg);

    mbar_wait(palBar, palPhase);
    palPhase ^= 1u;

    for (; vg < vg1; vg += W) {                                  // warp-private loop: no CTA-wide barrier inside an item
      const Rec2 v = cur;
      if (vg + W < vg1) cur = load_rec(vg + W);                   // next pass' record lands while this one is evaluated

      const uint32_t j01 = __float_as_uint(v.q5.x), j23 = __float_as_uint(v.q5.y), meta = __float_as_uint(v.q5.z);
      const uint32_t first = __shfl_sync(0xffffffffu, __float_as_uint(v.q5.w), 0);     // group constants: uniform registers
      const uint32_t cnt = __shfl_sync(0xffffffffu, (meta >> kM2Cnt) & 127u, 0);
      const uint32_t j0 = (j01 & 0xFFFFu) * 48u, j1 = (j01 >> 16) * 48u, j2 = (j23 & 0xFFFFu) * 48u, j3 = (j23 >> 16) * 48u;
      const uint32_t oA = (meta & 63u) * kVtxB, oB = ((meta >> kM2SlotB) & 63u) * kVtxB;   // the lane's two staging slots
      const int nmax = __reduce_max_sync(0xffffffffu, (int)((meta >> kM2N) & 7u));
      // unit weights everywhere (rigid window): the blend is the row itself
      const bool unitW = __all_sync(0xffffffffu, (v.q0.w == 1.0f || !(meta & kM2HasA)) && (v.q2.w == 1.0f || !(meta & kM2HasB)));
      const float2 wA0 = make_float2(v.q0.w, v.q0.w), wA1 = make_float2(v.q1.w, v.q1.w), wA2 = make_float2(v.q4.x, v.q4.x), wA3 = make_float2(v.q4.y, v.q4.y);
      const float2 wB0 = make_float2(v.q2.w, v.q2.w), wB1 = make_float2(v.q3.w, v.q3.w), wB2 = make_float2(v.q4.z, v.q4.z), wB3 = make_float2(v.q4.w, v.q4.w);
      const uint32_t nAligned = ILV ? cnt : (cnt & ~3u);          // bulk sizes are multiples of 16 bytes

      // mat-vec of one vertex with its blended matrix (pair layout: (x,y) out of packed FFMA2), normalise, stage.
      // ex0 / ex1: the vertex' texture coordinates (ILV) or its outline offset (HULL)
      auto emit = [&](const float4 mA, const float4 mB, const float4 mC, const float px, const float py, const float pz,
                      const float nxi, const float nyi, const float nzi, const float ex0, const float ex1, const uint32_t sa, const bool sw,
                      const int inst, const bool real) {
        const float2 qx2 = make_float2(px, px), qy2 = make_float2(py, py), qz2 = make_float2(pz, pz);
        const float2 oxy = __ffma2_rn(make_float2(mA.x, mA.y), qx2,
                           __ffma2_rn(make_float2(mA.z, mA.w), qy2, __ffma2_rn(make_float2(mB.x, mB.y), qz2, make_float2(mB.z, mB.w))));
        const float oz = fmaf(mC.x, px, fmaf(mC.y, py, fmaf(mC.z, pz, mC.w)));
        if (BOUNDS) {
          if (real) {                                             // (padding lanes and a fallback group's B side carry no vertex)
            bmin[inst][0] = fminf(bmin[inst][0], oxy.x); bmax[inst][0] = fmaxf(bmax[inst][0], oxy.x);
            bmin[inst][1] = fminf(bmin[inst][1], oxy.y); bmax[inst][1] = fmaxf(bmax[inst][1], oxy.y);
            bmin[inst][2] = fminf(bmin[inst][2], oz); bmax[inst][2] = fmaxf(bmax[inst][2], oz);
          }
        }
        float nx = 0.f, ny = 0.f, nz = 0.f;
        if (NRM) {
          const float2 nx2 = make_float2(nxi, nxi), ny2 = make_float2(nyi, nyi), nz2 = make_float2(nzi, nzi);
          const float2 nxy = __ffma2_rn(make_float2(mA.x, mA.y), nx2,
                             __ffma2_rn(make_float2(mA.z, mA.w), ny2, __fmul2_rn(make_float2(mB.x, mB.y), nz2)));
          nx = nxy.x; ny = nxy.y; nz = fmaf(mC.x, nxi, fmaf(mC.y, nyi, mC.z * nzi));
          const float l2 = fmaf(nx, nx, fmaf(ny, ny, nz * nz));
          const float rl = l2 > 0.f ? rsqrtf(l2) : 0.f;             // normalize(0) := 0 (SURVEY 8c edge case)
          nx *= rl; ny *= rl; nz *= rl;
        }
        if (ILV) {
          // 32-byte records: the two 16-byte halves of a record are written in opposite order by every second group of four
          // slots, so the 8 lanes of a quarter-warp whose slots are distinct modulo 8 hit 8 distinct 16-byte bank groups in
          // each of the two store instructions (in plain order slots s and s + 4 collide: 2-way conflicts on every store)
          // (sw = bit 2 of the slot)
          const float4 h0 = make_float4(oxy.x, oxy.y, oz, nx), h1 = make_float4(ny, nz, ex0, ex1);
          const float4 f = sw ? h1 : h0, g = sw ? h0 : h1;
          sts128(sa + (sw ? 16u : 0u), f.x, f.y, f.z, f.w);
          sts128(sa + (sw ? 0u : 16u), g.x, g.y, g.z, g.w);
        } else {
          sts3(sa, oxy.x, oxy.y, oz);
          if (NRM) sts3(sa + kPlaneB, nx, ny, nz);
          if (HULL) sts3(sa + kHullO, fmaf(nx, ex0, oxy.x), fmaf(ny, ex0, oxy.y), fmaf(nz, ex0, oz));   // engine.ts:458-461
        }
      };

      auto body = [&](auto NM, const int i0, const uint32_t stg) {
        constexpr int NV = decltype(NM)::value;                   // 0: rigid window (unit weights), else warp-max slot count
        constexpr int NMAX = NV == 0 ? 1 : NV;
        // ---- phase 1: rows of slots 0/1 for the whole sub-batch, issued back to back (their latencies overlap).  A lane
        // whose weight for a slot is zero carries the row of a neighbouring lane (lane_plan.h), so nothing is predicated.
        float4 a0[SB], a1[SB], a2[SB], b0[SB], b1[SB], b2[SB];
#pragma unroll
        for (int ii = 0; ii < SB; ++ii) {
          const uint32_t pb = sPal + (uint32_t)(i0 + ii) * palBytes;
          a0[ii] = lds128(pb + j0); a1[ii] = lds128(pb + j0 + 16u); a2[ii] = lds128(pb + j0 + 32u);
          if (NMAX > 1) { b0[ii] = lds128(pb + j1); b1[ii] = lds128(pb + j1 + 16u); b2[ii] = lds128(pb + j1 + 32u); }
        }
        // ---- phase 2: per instance, blend the rows once for each of the lane's two vertices, transform, stage
#pragma unroll
        for (int ii = 0; ii < SB; ++ii) {
          const uint32_t pb = sPal + (uint32_t)(i0 + ii) * palBytes;
          const uint32_t so = stg + (uint32_t)ii * kInstB;
          float4 c0, c1, c2, d0, d1, d2;
          if (NMAX > 2) { c0 = lds128(pb + j2); c1 = lds128(pb + j2 + 16u); c2 = lds128(pb + j2 + 32u); }
          if (NMAX > 3) { d0 = lds128(pb + j3); d1 = lds128(pb + j3 + 16u); d2 = lds128(pb + j3 + 32u); }
          {
            float4 mA, mB, mC;
            if (NV == 0) { mA = a0[ii]; mB = a1[ii]; mC = a2[ii]; }
            else { mA = f4_scale(a0[ii], wA0); mB = f4_scale(a1[ii], wA0); mC = f4_scale(a2[ii], wA0); }
            if (NMAX > 1) { mA = f4_fma(b0[ii], wA1, mA); mB = f4_fma(b1[ii], wA1, mB); mC = f4_fma(b2[ii], wA1, mC); }
            if (NMAX > 2) { mA = f4_fma(c0, wA2, mA); mB = f4_fma(c1, wA2, mB); mC = f4_fma(c2, wA2, mC); }
            if (NMAX > 3) { mA = f4_fma(d0, wA3, mA); mB = f4_fma(d1, wA3, mB); mC = f4_fma(d2, wA3, mC); }
            emit(mA, mB, mC, v.q0.x, v.q0.y, v.q0.z, v.q1.x, v.q1.y, v.q1.z, v.q6.x, ILV ? v.q6.y : 0.f, so + oA, (meta & 4u) != 0u,
                 BOUNDS ? i0 + ii : 0, (meta & kM2HasA) != 0u);
          }
          {   // (a fallback group's lanes carry no second vertex: their B side blends zeros into a slot that is never drained --
              //  cheaper than a branch that would keep the two independent dependency chains from interleaving)
            float4 mA, mB, mC;
            if (NV == 0) { mA = a0[ii]; mB = a1[ii]; mC = a2[ii]; }
            else { mA = f4_scale(a0[ii], wB0); mB = f4_scale(a1[ii], wB0); mC = f4_scale(a2[ii], wB0); }
            if (NMAX > 1) { mA = f4_fma(b0[ii], wB1, mA); mB = f4_fma(b1[ii], wB1, mB); mC = f4_fma(b2[ii], wB1, mC); }
            if (NMAX > 2) { mA = f4_fma(c0, wB2, mA); mB = f4_fma(c1, wB2, mB); mC = f4_fma(c2, wB2, mC); }
            if (NMAX > 3) { mA = f4_fma(d0, wB3, mA); mB = f4_fma(d1, wB3, mB); mC = f4_fma(d2, wB3, mC); }
            emit(mA, mB, mC, v.q2.x, v.q2.y, v.q2.z, v.q3.x, v.q3.y, v.q3.z, ILV ? v.q6.z : v.q6.y, ILV ? v.q6.w : 0.f, so + oB, (meta & (4u << kM2SlotB)) != 0u,
                 BOUNDS ? i0 + ii : 0, (meta & kM2HasB) != 0u);
          }
        }
      };

      float* const dst0 = outItem + (size_t)first * (kVtxB / 4u);
      auto sub_batch = [&](const int sbi) {
        const int i0 = sbi * SB;
        // the stores issued from this staging buffer NBUF sub-batches ago must have finished reading it: all but the
        // NBUF-1 most recent commit groups are awaited
        if (elect_one()) bulk_wait_read<NBUF - 1>();
        __syncwarp();
        const uint32_t stg = sStageW + sbuf * kBufB;
        switch (nmax) {
          case 1: if (unitW) body(IntC<0>{}, i0, stg); else body(IntC<1>{}, i0, stg); break;
          case 2: body(IntC<2>{}, i0, stg); break;
          case 3: body(IntC<3>{}, i0, stg); break;
          default: body(IntC<4>{}, i0, stg); break;
        }
        // ---- drain this sub-batch: the warp's 64 vertices x SB instances leave through the TMA (one elected lane)
        fence_proxy_async();
        __syncwarp();
        if (elect_one()) {
          if (nAligned) {
#pragma unroll
            for (int ii = 0; ii < SB; ++ii) {
              const int i = i0 + ii;
              if ((uint32_t)i < nInst) {
                float* dst = dst0 + (size_t)i * prm.instStrideF;
                bulk_s2g(dst, stg + (uint32_t)ii * kInstB, nAligned * kVtxB, polFirst);
                if (NRM && !ILV) bulk_s2g(dst + prm.nrmOffF, stg + (uint32_t)ii * kInstB + kPlaneB, nAligned * 12u, polFirst);
                if (HULL) bulk_s2g(dst + prm.hullOffF, stg + (uint32_t)ii * kInstB + kHullO, nAligned * 12u, polFirst);
              }
            }
          }
          bulk_commit();
        }
        if (nAligned != cnt) {
          // cold path (the mesh's last group only): the ragged tail (< 4 vertices) leaves with plain stores
          __syncwarp();
          const uint32_t nf = (cnt - nAligned) * 3u;
          if ((uint32_t)lane < nf) {
            const uint32_t o = nAligned * 3u + (uint32_t)lane;
            for (int ii = 0; ii < SB; ++ii) {
              const uint32_t i = (uint32_t)(i0 + ii);
              if (i >= nInst) break;
              float* dst = dst0 + (size_t)i * prm.instStrideF + o;
              st_cs(dst, lds32(stg + (uint32_t)ii * kInstB + o * 4u));
              if (NRM) st_cs(dst + prm.nrmOffF, lds32(stg + (uint32_t)ii * kInstB + kPlaneB + o * 4u));
              if (HULL) st_cs(dst + prm.hullOffF, lds32(stg + (uint32_t)ii * kInstB + kHullO + o * 4u));
            }
          }
          __syncwarp();
        }
        sbuf = (sbuf + 1u == (uint32_t)NBUF) ? 0u : sbuf + 1u;
      };
      if (BOUNDS) {
#pragma unroll
        for (int sbi = 0; sbi < NSB; ++sbi) sub_batch(sbi);       // static instance indices for the register accumulators
      } else {
#pragma unroll 1
        for (int sbi = 0; sbi < NSB; ++sbi) sub_batch(sbi);       // (not unrolled: one copy of the five bodies keeps the I-cache warm)
      }
    }
    if (BOUNDS) {
      // per-item reduction: warp shuffle, then one atomic per warp per bound (ordered-int encoding), as deform_kernel does
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
  bulk_wait0();        // staging must outlive the TMA reads (bulk groups are per thread: every lane waits)
}

}  // namespace rz
