// aux_kernels.cuh — small per-frame helper kernels (included by rze_b200.cu only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rz {

// ------------------------------------------------------------------ skin = world * invBind
// Replaces the reference's only compute shader (engine.ts:920-929).  One thread per (palette, bone);
// keeps rows 0..2 of the column-major product (row 3 never reaches the blend's outputs, engine.ts:260-272).
__global__ void skin_matrices_kernel(const float4* __restrict__ world, const float4* __restrict__ invBind,
                                     float4* __restrict__ skin, const uint32_t* __restrict__ bonePos, uint32_t P, uint32_t B, uint32_t soa) {
  const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= P * B) return;
  const uint32_t b = idx % B;
  const float4 w0 = world[(size_t)idx * 4], w1 = world[(size_t)idx * 4 + 1], w2 = world[(size_t)idx * 4 + 2],
               w3 = world[(size_t)idx * 4 + 3];                      // columns of world
  float r[3][4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const float4 ib = __ldg(invBind + (size_t)b * 4 + c);            // column c of invBind
    r[0][c] = fmaf(w3.x, ib.w, fmaf(w2.x, ib.z, fmaf(w1.x, ib.y, w0.x * ib.x)));
    r[1][c] = fmaf(w3.y, ib.w, fmaf(w2.y, ib.z, fmaf(w1.y, ib.y, w0.y * ib.x)));
    r[2][c] = fmaf(w3.z, ib.w, fmaf(w2.z, ib.z, fmaf(w1.z, ib.y, w0.z * ib.x)));
  }
  const size_t row = (size_t)(idx - b) + __ldg(bonePos + b);         // bank-aware palette permutation (rze_b200.cu rebuild_tables)
  // pair layout (deform_kernel.cuh kRowF4): rows 0/1 interleaved, row 2 as is
  const float4 cA = make_float4(r[0][0], r[1][0], r[0][1], r[1][1]);
  const float4 cB = make_float4(r[0][2], r[1][2], r[0][3], r[1][3]);
  const float4 cC = make_float4(r[2][0], r[2][1], r[2][2], r[2][3]);
  if (soa) {          // [P][3][B] float4: chunk r of all bones contiguous (8 consecutive palette rows = one 128-byte line)
    const size_t pbase = (size_t)(idx - b) * 3, pos = __ldg(bonePos + b);
    skin[pbase + pos] = cA; skin[pbase + B + pos] = cB; skin[pbase + 2 * (size_t)B + pos] = cC;
  } else {            // [P][B][3] float4
    skin[row * 3 + 0] = cA; skin[row * 3 + 1] = cB; skin[row * 3 + 2] = cC;
  }
}

// (A one-CTA-per-palette variant with coalesced loads and stores through shared memory runs 52.7 -> 38.9 us on its own but
//  makes the frame SLOWER, 1.870 -> 1.894 ms: its 64 KB of shared memory per CTA cannot share an SM with the deform kernel's
//  221 KB, so the tails of the two kernels stop overlapping.  This shared-memory-free kernel stays.)
// quat[p][row] = rotation part of skin[p][row] as a quaternion (math.ts:406-448 Mat4.toQuatFromArray, via
// deform_kernel.cuh quat_from_rows).  Feeds the SDEF dense phase: the deform kernel slerps these instead of converting
// two matrices per SDEF vertex-instance.  One thread per (palette, palette row).
__global__ void skin_quats_kernel(const float4* __restrict__ skin, float4* __restrict__ quat, uint32_t P, uint32_t B, uint32_t soa) {
  const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= P * B) return;
  const uint32_t row = idx % B;
  float4 cA, cB, cC;
  if (soa) { const size_t pb = (size_t)(idx - row) * 3; cA = skin[pb + row]; cB = skin[pb + B + row]; cC = skin[pb + 2 * (size_t)B + row]; }
  else { cA = skin[(size_t)idx * 3]; cB = skin[(size_t)idx * 3 + 1]; cC = skin[(size_t)idx * 3 + 2]; }
  // un-pair (deform_kernel.cuh kRowF4)
  const Q4 q = quat_from_rows(make_float4(cA.x, cA.z, cB.x, cB.z), make_float4(cA.y, cA.w, cB.y, cB.w), cC);
  quat[idx] = make_float4(q.x, q.y, q.z, q.w);
}

// dense per-instance morph weights: dense[k][m] = 0, dense[k][activeIds[a]] += w[k][a]
__global__ void morph_weights_kernel(const float* __restrict__ w, const uint32_t* __restrict__ ids, float* __restrict__ dense,
                                     uint32_t K, uint32_t Mact, uint32_t Mpad) {
  const uint32_t k = blockIdx.x;
  if (k >= K) return;
  for (uint32_t m = threadIdx.x; m < Mpad; m += blockDim.x) dense[(size_t)k * Mpad + m] = 0.f;
  __syncthreads();
  for (uint32_t a = threadIdx.x; a < Mact; a += blockDim.x) atomicAdd(&dense[(size_t)k * Mpad + ids[a]], w[(size_t)k * Mact + a]);
}

// ---------------------------------------------------------------------------------------------------------------
// GPU pose evaluation: replaces Model.evaluatePose (model.ts:325-328) + the palette upload for crowds.
//   tween  : q = slerp(start, target, easeInOut(clamp((now-start)/max(1,dur))))            model.ts:158-194, math.ts:2-4,156-189
//   local  : R = fromQuat(append) * fromQuat(q), append = slerp(identity, +/-q[appendParent], |ratio|)   model.ts:356-386
//   world  : world[b] = world[parent] * (T(bind) * R)      parents first (level order)     model.ts:397-414
//   skin   : world * invBind, rows 0..2, written straight into the deform kernel's palette layout
// One CTA per palette; bones of one hierarchy level are evaluated in parallel, world matrices live in shared memory
// as 3x4 affine (the reference's products never leave the affine group: bottom row stays 0,0,0,1).
struct PoseSkeleton {
  const int32_t* __restrict__ parent;        // [B]
  const float* __restrict__ bindT;           // [B][3]
  const int32_t* __restrict__ appendParent;  // [B] (-1: none)
  const float* __restrict__ appendRatio;     // [B] clamped to [-1,1]; 0 when no append rotation
  const uint32_t* __restrict__ levelBones;   // bones sorted by depth
  const uint32_t* __restrict__ levelStart;   // [nLevels+1]
  uint32_t nLevels, B;
};
struct PoseTweens {
  const float4* __restrict__ start;   // [B] xyzw
  const float4* __restrict__ target;  // [B]
  const float4* __restrict__ rest;    // [B] local rotation of bones without an active tween
  const float* __restrict__ startMs;  // [B]
  const float* __restrict__ durMs;    // [B]
  const uint8_t* __restrict__ active; // [B]
};

// Keyframe tracks of one animation clip: bone b owns keys keyStart[b] .. keyStart[b+1]-1 (times ascending, ms).
// Evaluation = what the reference's playback produces when its timers fire on time (engine.ts:1451-1553 + the tween
// rule model.ts:158-194): a key at t = 0 is applied instantly, bones without one start from identity; key i is reached
// by a tween from key i-1 over [t(i-1), t(i)] with quadratic ease + slerp; the last key is held.
struct PoseTracks {
  const uint32_t* __restrict__ keyStart;   // [B+1]
  const float* __restrict__ keyMs;         // [nKeys]
  const float4* __restrict__ keyQ;         // [nKeys] normalised xyzw
};

__device__ __forceinline__ float4 q_slerp(float4 a, float4 b, float t) {   // math.ts:156-189
  float c = a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
  if (c < 0.f) { c = -c; b.x = -b.x; b.y = -b.y; b.z = -b.z; b.w = -b.w; }
  if (c > 0.9995f) {
    const float x = a.x + t * (b.x - a.x), y = a.y + t * (b.y - a.y), z = a.z + t * (b.z - a.z), w = a.w + t * (b.w - a.w);
    const float inv = 1.0f / sqrtf(x * x + y * y + z * z + w * w);
    return make_float4(x * inv, y * inv, z * inv, w * inv);
  }
  const float th0 = acosf(c), s = sinf(th0), th = th0 * t;
  const float s0 = sinf(th0 - th) / s, s1 = sinf(th) / s;
  return make_float4(s0 * a.x + s1 * b.x, s0 * a.y + s1 * b.y, s0 * a.z + s1 * b.z, s0 * a.w + s1 * b.w);
}
__device__ __forceinline__ void q_to_rows(float4 q, float r[3][3]) {     // math.ts:352-384 (row r, column c)
  const float x2 = q.x + q.x, y2 = q.y + q.y, z2 = q.z + q.z;
  const float xx = q.x * x2, xy = q.x * y2, xz = q.x * z2, yy = q.y * y2, yz = q.y * z2, zz = q.z * z2;
  const float wx = q.w * x2, wy = q.w * y2, wz = q.w * z2;
  r[0][0] = 1.f - (yy + zz); r[0][1] = xy - wz; r[0][2] = xz + wy;
  r[1][0] = xy + wz; r[1][1] = 1.f - (xx + zz); r[1][2] = yz - wx;
  r[2][0] = xz - wy; r[2][1] = yz + wx; r[2][2] = 1.f - (xx + yy);
}

__device__ __forceinline__ float ease_in_out(float t) {          // math.ts:2-4
  const float u = -2.0f * t + 2.0f;
  return t < 0.5f ? 2.0f * t * t : 1.0f - (u * u) * 0.5f;
}

// local rotation of bone b of pose p.  MODE 0: uploaded; MODE 1: shared tween table at nowMs[p]; MODE 2: keyframe tracks at nowMs[p]
template <int MODE>
__device__ __forceinline__ float4 eval_rotation(uint32_t b, uint32_t p, uint32_t B, const PoseTweens& tw, const PoseTracks& tr,
                                                const float4* __restrict__ localRot, const float* __restrict__ nowMs) {
  if (MODE == 0) return localRot[(size_t)p * B + b];
  if (MODE == 1) {
    if (!tw.active[b]) return tw.rest[b];
    const float dur = fmaxf(1.0f, tw.durMs[b]);
    const float t = fminf(1.0f, fmaxf(0.0f, (nowMs[p] - tw.startMs[b]) / dur));
    return q_slerp(tw.start[b], tw.target[b], ease_in_out(t));
  }
  const uint32_t k0 = tr.keyStart[b], k1 = tr.keyStart[b + 1];
  if (k0 == k1) return tw.rest[b];                               // bone not animated by the clip
  const float now = nowMs[p];
  float tPrev = 0.0f;
  float4 qPrev = make_float4(0.f, 0.f, 0.f, 1.f);                // reset to identity unless a key sits at t = 0
  for (uint32_t k = k0; k < k1; ++k) {
    const float tk = tr.keyMs[k];
    const float4 qk = tr.keyQ[k];
    if (tk <= 0.0f) { qPrev = qk; tPrev = 0.0f; continue; }
    if (now <= tPrev) return qPrev;
    if (now < tk) {
      const float t = fminf(1.0f, fmaxf(0.0f, (now - tPrev) / fmaxf(1.0f, tk - tPrev)));
      return q_slerp(qPrev, qk, ease_in_out(t));
    }
    qPrev = qk; tPrev = tk;
  }
  return qPrev;
}

// MODE 0: local rotations uploaded ([P][B] float4);  MODE 1: rotations from the shared tween table at time nowMs[p];  MODE 2: tracks
template <int MODE>
__global__ void pose_kernel(PoseSkeleton sk, PoseTweens tw, PoseTracks tr, const float4* __restrict__ localRot, const float* __restrict__ nowMs,
                            const float4* __restrict__ invBind, const uint32_t* __restrict__ bonePos, float4* __restrict__ skin,
                            float4* __restrict__ worldOut /* optional [P][B][3] rows, may be null */, uint32_t soa) {
  extern __shared__ float4 s_world[];            // [B][3] rows of the 3x4 world matrices
  float4* s_q = s_world + (size_t)sk.B * 3;      // [B] local rotations (needed by append children)
  const uint32_t p = blockIdx.x, B = sk.B;
  for (uint32_t b = threadIdx.x; b < B; b += blockDim.x) s_q[b] = eval_rotation<MODE>(b, p, B, tw, tr, localRot, nowMs);
  __syncthreads();
  for (uint32_t L = 0; L < sk.nLevels; ++L) {
    const uint32_t l0 = sk.levelStart[L], l1 = sk.levelStart[L + 1];
    for (uint32_t i = l0 + threadIdx.x; i < l1; i += blockDim.x) {
      const uint32_t b = sk.levelBones[i];
      float R[3][3];
      q_to_rows(s_q[b], R);
      const int32_t ap = sk.appendParent[b];
      const float ratio = sk.appendRatio[b];
      if (ap >= 0 && fabsf(ratio) > 1e-6f) {
        float4 a = s_q[ap];
        if (ratio < 0.f) { a.x = -a.x; a.y = -a.y; a.z = -a.z; }
        const float4 aq = q_slerp(make_float4(0.f, 0.f, 0.f, 1.f), a, fabsf(ratio));
        float A[3][3], T[3][3];
        q_to_rows(aq, A);
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int c = 0; c < 3; ++c) T[r][c] = A[r][0] * R[0][c] + A[r][1] * R[1][c] + A[r][2] * R[2][c];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int c = 0; c < 3; ++c) R[r][c] = T[r][c];
      }
      const float tx = sk.bindT[(size_t)b * 3], ty = sk.bindT[(size_t)b * 3 + 1], tz = sk.bindT[(size_t)b * 3 + 2];
      float W[3][4];
      const int32_t par = sk.parent[b];
      if (par >= 0) {
        const float4 p0 = s_world[(size_t)par * 3], p1 = s_world[(size_t)par * 3 + 1], p2 = s_world[(size_t)par * 3 + 2];
        const float P[3][4] = {{p0.x, p0.y, p0.z, p0.w}, {p1.x, p1.y, p1.z, p1.w}, {p2.x, p2.y, p2.z, p2.w}};
#pragma unroll
        for (int r = 0; r < 3; ++r) {
#pragma unroll
          for (int c = 0; c < 3; ++c) W[r][c] = P[r][0] * R[0][c] + P[r][1] * R[1][c] + P[r][2] * R[2][c];
          W[r][3] = P[r][0] * tx + P[r][1] * ty + P[r][2] * tz + P[r][3];
        }
      } else {
#pragma unroll
        for (int r = 0; r < 3; ++r) { W[r][0] = R[r][0]; W[r][1] = R[r][1]; W[r][2] = R[r][2]; }
        W[0][3] = tx; W[1][3] = ty; W[2][3] = tz;
      }
#pragma unroll
      for (int r = 0; r < 3; ++r) s_world[(size_t)b * 3 + r] = make_float4(W[r][0], W[r][1], W[r][2], W[r][3]);
      if (worldOut)
#pragma unroll
        for (int r = 0; r < 3; ++r) worldOut[((size_t)p * B + b) * 3 + r] = make_float4(W[r][0], W[r][1], W[r][2], W[r][3]);
      // skin = world * invBind (rows 0..2); invBind column-major mat4
      float S[3][4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float4 ib = __ldg(invBind + (size_t)b * 4 + c);
#pragma unroll
        for (int r = 0; r < 3; ++r) S[r][c] = W[r][0] * ib.x + W[r][1] * ib.y + W[r][2] * ib.z + W[r][3] * ib.w;
      }
      const float4 cA = make_float4(S[0][0], S[1][0], S[0][1], S[1][1]);
      const float4 cB = make_float4(S[0][2], S[1][2], S[0][3], S[1][3]);
      const float4 cC = make_float4(S[2][0], S[2][1], S[2][2], S[2][3]);
      const size_t pos = __ldg(bonePos + b);
      if (soa) {
        const size_t pb = (size_t)p * B * 3;
        skin[pb + pos] = cA; skin[pb + B + pos] = cB; skin[pb + 2 * (size_t)B + pos] = cC;
      } else {
        const size_t row = (size_t)p * B + pos;
        skin[row * 3] = cA; skin[row * 3 + 1] = cB; skin[row * 3 + 2] = cC;
      }
    }
    __syncthreads();
  }
}

// Same evaluation without the per-level barriers: every thread composes the local transforms of its bone's ancestor
// chain root -> bone (the association order of the reference's recursion, model.ts:405-411), so one CTA needs only
// two barriers per palette.  ~depth x 36 FMA per bone instead of 36, but no 40-level latency chain: 4x faster for crowds.
template <int MODE>
__global__ void pose_chain_kernel(PoseSkeleton sk, PoseTweens tw, PoseTracks tr, const uint32_t* __restrict__ chainStart, const uint32_t* __restrict__ chainBones,
                                  const float4* __restrict__ localRot, const float* __restrict__ nowMs, const float4* __restrict__ invBind,
                                  const uint32_t* __restrict__ bonePos, float4* __restrict__ skin, uint32_t soa) {
  extern __shared__ float4 s_loc[];              // [B][3] rows of the local 3x4 transforms T(bind) * R
  float4* s_q = s_loc + (size_t)sk.B * 3;        // [B] local rotations
  const uint32_t p = blockIdx.x, B = sk.B;
  for (uint32_t b = threadIdx.x; b < B; b += blockDim.x) s_q[b] = eval_rotation<MODE>(b, p, B, tw, tr, localRot, nowMs);
  __syncthreads();
  for (uint32_t b = threadIdx.x; b < B; b += blockDim.x) {
    float R[3][3];
    q_to_rows(s_q[b], R);
    const int32_t ap = sk.appendParent[b];
    const float ratio = sk.appendRatio[b];
    if (ap >= 0 && fabsf(ratio) > 1e-6f) {
      float4 a = s_q[ap];
      if (ratio < 0.f) { a.x = -a.x; a.y = -a.y; a.z = -a.z; }
      const float4 aq = q_slerp(make_float4(0.f, 0.f, 0.f, 1.f), a, fabsf(ratio));
      float A[3][3], T[3][3];
      q_to_rows(aq, A);
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) T[r][c] = A[r][0] * R[0][c] + A[r][1] * R[1][c] + A[r][2] * R[2][c];
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) R[r][c] = T[r][c];
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) s_loc[(size_t)b * 3 + r] = make_float4(R[r][0], R[r][1], R[r][2], sk.bindT[(size_t)b * 3 + r]);
  }
  __syncthreads();
  for (uint32_t b = threadIdx.x; b < B; b += blockDim.x) {
    const uint32_t c0 = chainStart[b], c1 = chainStart[b + 1];
    uint32_t a = chainBones[c0];
    float4 w0 = s_loc[(size_t)a * 3], w1 = s_loc[(size_t)a * 3 + 1], w2 = s_loc[(size_t)a * 3 + 2];
    for (uint32_t i = c0 + 1; i < c1; ++i) {
      a = chainBones[i];
      const float4 l0 = s_loc[(size_t)a * 3], l1 = s_loc[(size_t)a * 3 + 1], l2 = s_loc[(size_t)a * 3 + 2];
      // W <- W * L  (affine 3x4): same term order as the level kernel
      const float4 n0 = make_float4(w0.x * l0.x + w0.y * l1.x + w0.z * l2.x, w0.x * l0.y + w0.y * l1.y + w0.z * l2.y,
                                    w0.x * l0.z + w0.y * l1.z + w0.z * l2.z, w0.x * l0.w + w0.y * l1.w + w0.z * l2.w + w0.w);
      const float4 n1 = make_float4(w1.x * l0.x + w1.y * l1.x + w1.z * l2.x, w1.x * l0.y + w1.y * l1.y + w1.z * l2.y,
                                    w1.x * l0.z + w1.y * l1.z + w1.z * l2.z, w1.x * l0.w + w1.y * l1.w + w1.z * l2.w + w1.w);
      const float4 n2 = make_float4(w2.x * l0.x + w2.y * l1.x + w2.z * l2.x, w2.x * l0.y + w2.y * l1.y + w2.z * l2.y,
                                    w2.x * l0.z + w2.y * l1.z + w2.z * l2.z, w2.x * l0.w + w2.y * l1.w + w2.z * l2.w + w2.w);
      w0 = n0; w1 = n1; w2 = n2;
    }
    const float W[3][4] = {{w0.x, w0.y, w0.z, w0.w}, {w1.x, w1.y, w1.z, w1.w}, {w2.x, w2.y, w2.z, w2.w}};
    float S[3][4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float4 ib = __ldg(invBind + (size_t)b * 4 + c);
#pragma unroll
      for (int r = 0; r < 3; ++r) S[r][c] = W[r][0] * ib.x + W[r][1] * ib.y + W[r][2] * ib.z + W[r][3] * ib.w;
    }
    const float4 cA = make_float4(S[0][0], S[1][0], S[0][1], S[1][1]);
    const float4 cB = make_float4(S[0][2], S[1][2], S[0][3], S[1][3]);
    const float4 cC = make_float4(S[2][0], S[2][1], S[2][2], S[2][3]);
    const size_t pos = __ldg(bonePos + b);
    if (soa) {
      const size_t pb = (size_t)p * B * 3;
      skin[pb + pos] = cA; skin[pb + B + pos] = cB; skin[pb + 2 * (size_t)B + pos] = cC;
    } else {
      const size_t row = (size_t)p * B + pos;
      skin[row * 3] = cA; skin[row * 3 + 1] = cB; skin[row * 3 + 2] = cC;
    }
  }
}

// Same evaluation by POINTER JUMPING: the chain walk above multiplies depth x 36 FMA per bone (a 23-level average on the
// benchmark skeleton makes the kernel instruction-bound: 0.17 ms for 4096 poses x 512 bones); here every bone holds the
// product of its local transforms over a window of ancestors that doubles each round,
//     W[b] <- W[anc(b)] * W[b],  anc(b) <- anc(anc(b)),
// so ceil(log2(depth)) rounds of one 3x4 product + one barrier finish the whole skeleton (6 rounds for depth 44).  The
// products associate differently from the reference's recursion (model.ts:405-411), a <= 1e-6 relative difference of the
// matrices, inside the 1e-5 tolerance of the path.  MODE 1 (shared tween table) also takes the per-bone slerp constants
// (angle, 1/sin, hemisphere sign: functions of the two keys only) from `twAux`, computed once in rz_set_tweens, so a
// (pose, bone) rotation costs two polynomial sines instead of acos + three sines.
// Shared memory, one bone per thread (B <= blockDim.x): W [B][3] float4, the local rotations [B] float4, anc [B] int =
// 68 B per bone (34 KB at B = 512: under the 48 KB that needs no opt-in); more bones than threads: two ping-pong copies of W
// and of anc, the rotations alias the second W copy (104 B per bone).
template <int MODE>
__global__ void pose_jump_kernel(PoseSkeleton sk, PoseTweens tw, PoseTracks tr, const float4* __restrict__ twAux,
                                 const float4* __restrict__ localRot, const float* __restrict__ nowMs, const float4* __restrict__ invBind,
                                 const float4* __restrict__ invBindSoA /* [4][B]: column c of bone b at c*B + b (coalesced) */,
                                 const uint32_t* __restrict__ bonePos, float4* __restrict__ skin, uint32_t rounds, uint32_t soa) {
  extern __shared__ float4 s_w[];
  const uint32_t p = blockIdx.x, B = sk.B;
  const bool onePerThread = B <= blockDim.x;
  float4* wA = s_w;
  float4* wB = s_w + (size_t)B * 3;                    // (ping-pong copy: only when a thread owns several bones)
  float4* s_q = wB;                                    // [B]; aliases wB, dead before the first round writes it
  int32_t* ancA = reinterpret_cast<int32_t*>(s_w + (size_t)B * (onePerThread ? 4 : 6));
  int32_t* ancB = ancA + B;
  for (uint32_t b = threadIdx.x; b < B; b += blockDim.x) {
    float4 q;
    if (MODE == 1) {
      if (!tw.active[b]) {
        q = tw.rest[b];
      } else {
        const float dur = fmaxf(1.0f, tw.durMs[b]);
        const float t = fminf(1.0f, fmaxf(0.0f, (nowMs[p] - tw.startMs[b]) / dur));
        const float e = ease_in_out(t);
        const float4 a = tw.start[b], ax = twAux[b];   // ax = (theta0, 1/sin(theta0), lerp branch flag, hemisphere sign)
        float4 bq = tw.target[b];
        bq.x *= ax.w; bq.y *= ax.w; bq.z *= ax.w; bq.w *= ax.w;
        if (ax.z != 0.f) {                             // cos > 0.9995: lerp + normalise (math.ts:166-175)
          const float x = a.x + e * (bq.x - a.x), y = a.y + e * (bq.y - a.y), z = a.z + e * (bq.z - a.z), w = a.w + e * (bq.w - a.w);
          const float inv = rsqrtf(x * x + y * y + z * z + w * w);
          q = make_float4(x * inv, y * inv, z * inv, w * inv);
        } else {
          const float th = ax.x * e;
          const float s0 = sin_0_halfpi(ax.x - th) * ax.y, s1 = sin_0_halfpi(th) * ax.y;
          q = make_float4(s0 * a.x + s1 * bq.x, s0 * a.y + s1 * bq.y, s0 * a.z + s1 * bq.z, s0 * a.w + s1 * bq.w);
        }
      }
    } else {
      q = eval_rotation<MODE>(b, p, B, tw, tr, localRot, nowMs);
    }
    s_q[b] = q;
  }
  __syncthreads();
  // local transforms T(bind) * R (R = append * local, model.ts:356-386) into registers, then into wA (s_q stays intact: it
  // lives in wB)
  for (uint32_t b = threadIdx.x; b < B; b += blockDim.x) {
    float R[3][3];
    q_to_rows(s_q[b], R);
    const int32_t ap = sk.appendParent[b];
    const float ratio = sk.appendRatio[b];
    if (ap >= 0 && fabsf(ratio) > 1e-6f) {
      float4 a = s_q[ap];
      if (ratio < 0.f) { a.x = -a.x; a.y = -a.y; a.z = -a.z; }
      const float4 aq = q_slerp(make_float4(0.f, 0.f, 0.f, 1.f), a, fabsf(ratio));
      float A[3][3], T[3][3];
      q_to_rows(aq, A);
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) T[r][c] = A[r][0] * R[0][c] + A[r][1] * R[1][c] + A[r][2] * R[2][c];
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) R[r][c] = T[r][c];
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) wA[(size_t)b * 3 + r] = make_float4(R[r][0], R[r][1], R[r][2], sk.bindT[(size_t)b * 3 + r]);
    ancA[b] = sk.parent[b];
  }
  __syncthreads();
  if (onePerThread) {
    // one bone per thread: the own window stays in registers, a round reads the ancestor's window (48 B) and, after a
    // barrier, publishes the product in place (48 B); bones whose window has reached the root drop out of both
    const uint32_t b = threadIdx.x;
    const bool mine = b < B;
    float4 w0 = make_float4(0.f, 0.f, 0.f, 0.f), w1 = w0, w2 = w0;
    int32_t a = -1;
    if (mine) { w0 = wA[(size_t)b * 3]; w1 = wA[(size_t)b * 3 + 1]; w2 = wA[(size_t)b * 3 + 2]; a = ancA[b]; }
    for (uint32_t r = 0; r < rounds; ++r) {
      const bool act = a >= 0;
      int32_t na = -1;
      if (act) {
        const float4 u0 = wA[(size_t)a * 3], u1 = wA[(size_t)a * 3 + 1], u2 = wA[(size_t)a * 3 + 2];
        na = ancA[a];
        const float4 n0 = make_float4(u0.x * w0.x + u0.y * w1.x + u0.z * w2.x, u0.x * w0.y + u0.y * w1.y + u0.z * w2.y,
                                      u0.x * w0.z + u0.y * w1.z + u0.z * w2.z, u0.x * w0.w + u0.y * w1.w + u0.z * w2.w + u0.w);
        const float4 n1 = make_float4(u1.x * w0.x + u1.y * w1.x + u1.z * w2.x, u1.x * w0.y + u1.y * w1.y + u1.z * w2.y,
                                      u1.x * w0.z + u1.y * w1.z + u1.z * w2.z, u1.x * w0.w + u1.y * w1.w + u1.z * w2.w + u1.w);
        const float4 n2 = make_float4(u2.x * w0.x + u2.y * w1.x + u2.z * w2.x, u2.x * w0.y + u2.y * w1.y + u2.z * w2.y,
                                      u2.x * w0.z + u2.y * w1.z + u2.z * w2.z, u2.x * w0.w + u2.y * w1.w + u2.z * w2.w + u2.w);
        w0 = n0; w1 = n1; w2 = n2;
      }
      if (!__syncthreads_or(act)) break;                 // (uniform) every window has reached its root
      if (act) {
        wA[(size_t)b * 3] = w0; wA[(size_t)b * 3 + 1] = w1; wA[(size_t)b * 3 + 2] = w2;
        ancA[b] = na;
        a = na;
      }
      __syncthreads();
    }
    // skin = W * invBind, assembled in shared memory in palette order (over W: nobody reads the windows any more, the
    // loop ended on a barrier) and written out as one contiguous B x 48-byte block: 3 fully coalesced 16-byte stores per thread instead of 3 scattered ones (32 lines each)
    if (mine) {
      const float W[3][4] = {{w0.x, w0.y, w0.z, w0.w}, {w1.x, w1.y, w1.z, w1.w}, {w2.x, w2.y, w2.z, w2.w}};
      float S[3][4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float4 ib = __ldg(invBindSoA + (size_t)c * B + b);
#pragma unroll
        for (int r = 0; r < 3; ++r) S[r][c] = W[r][0] * ib.x + W[r][1] * ib.y + W[r][2] * ib.z + W[r][3] * ib.w;
      }
      const float4 cA = make_float4(S[0][0], S[1][0], S[0][1], S[1][1]);
      const float4 cB = make_float4(S[0][2], S[1][2], S[0][3], S[1][3]);
      const float4 cC = make_float4(S[2][0], S[2][1], S[2][2], S[2][3]);
      const uint32_t pos = __ldg(bonePos + b);
      if (soa) { wA[pos] = cA; wA[B + pos] = cB; wA[2 * (size_t)B + pos] = cC; }
      else { wA[(size_t)pos * 3] = cA; wA[(size_t)pos * 3 + 1] = cB; wA[(size_t)pos * 3 + 2] = cC; }
    }
    __syncthreads();
    float4* dst = skin + (size_t)p * B * 3;
    for (uint32_t i = threadIdx.x; i < B * 3; i += blockDim.x) dst[i] = wA[i];
    return;
  }
  float4* wc = wA;
  float4* wn = wB;
  int32_t* ac = ancA;
  int32_t* an = ancB;
  for (uint32_t r = 0; r < rounds; ++r) {
    for (uint32_t b = threadIdx.x; b < B; b += blockDim.x) {
      const int32_t a = ac[b];
      float4 w0 = wc[(size_t)b * 3], w1 = wc[(size_t)b * 3 + 1], w2 = wc[(size_t)b * 3 + 2];
      int32_t na = -1;
      if (a >= 0) {
        // W <- W[a] * W[b] (affine 3x4): rows of the ancestor window times the own window
        const float4 u0 = wc[(size_t)a * 3], u1 = wc[(size_t)a * 3 + 1], u2 = wc[(size_t)a * 3 + 2];
        const float4 n0 = make_float4(u0.x * w0.x + u0.y * w1.x + u0.z * w2.x, u0.x * w0.y + u0.y * w1.y + u0.z * w2.y,
                                      u0.x * w0.z + u0.y * w1.z + u0.z * w2.z, u0.x * w0.w + u0.y * w1.w + u0.z * w2.w + u0.w);
        const float4 n1 = make_float4(u1.x * w0.x + u1.y * w1.x + u1.z * w2.x, u1.x * w0.y + u1.y * w1.y + u1.z * w2.y,
                                      u1.x * w0.z + u1.y * w1.z + u1.z * w2.z, u1.x * w0.w + u1.y * w1.w + u1.z * w2.w + u1.w);
        const float4 n2 = make_float4(u2.x * w0.x + u2.y * w1.x + u2.z * w2.x, u2.x * w0.y + u2.y * w1.y + u2.z * w2.y,
                                      u2.x * w0.z + u2.y * w1.z + u2.z * w2.z, u2.x * w0.w + u2.y * w1.w + u2.z * w2.w + u2.w);
        w0 = n0; w1 = n1; w2 = n2;
        na = ac[a];
      }
      wn[(size_t)b * 3] = w0; wn[(size_t)b * 3 + 1] = w1; wn[(size_t)b * 3 + 2] = w2;
      an[b] = na;
    }
    __syncthreads();
    float4* tw_ = wc; wc = wn; wn = tw_;
    int32_t* ta = ac; ac = an; an = ta;
  }
  for (uint32_t b = threadIdx.x; b < B; b += blockDim.x) {
    const float4 w0 = wc[(size_t)b * 3], w1 = wc[(size_t)b * 3 + 1], w2 = wc[(size_t)b * 3 + 2];
    const float W[3][4] = {{w0.x, w0.y, w0.z, w0.w}, {w1.x, w1.y, w1.z, w1.w}, {w2.x, w2.y, w2.z, w2.w}};
    float S[3][4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float4 ib = __ldg(invBind + (size_t)b * 4 + c);
#pragma unroll
      for (int r = 0; r < 3; ++r) S[r][c] = W[r][0] * ib.x + W[r][1] * ib.y + W[r][2] * ib.z + W[r][3] * ib.w;
    }
    const float4 cA = make_float4(S[0][0], S[1][0], S[0][1], S[1][1]);
    const float4 cB = make_float4(S[0][2], S[1][2], S[0][3], S[1][3]);
    const float4 cC = make_float4(S[2][0], S[2][1], S[2][2], S[2][3]);
    const size_t pos = __ldg(bonePos + b);
    if (soa) {
      const size_t pb = (size_t)p * B * 3;
      skin[pb + pos] = cA; skin[pb + B + pos] = cB; skin[pb + 2 * (size_t)B + pos] = cC;
    } else {
      const size_t row = (size_t)p * B + pos;
      skin[row * 3] = cA; skin[row * 3 + 1] = cB; skin[row * 3 + 2] = cC;
    }
  }
}

// Physics -> bone feedback (physics.ts:714-751): for every bone driven by dynamic rigid bodies,
//   boneWorld = fromPositionRotation(bodyPos, bodyRot) x bodyOffsetMatrixInverse,      skin = boneWorld x invBind,
// written over the palette row that the pose evaluation produced (children are NOT re-evaluated, exactly like the
// reference's in-place edit).  Bodies of one bone are applied in index order and a result whose [0] or [15] is NaN or
// >= 1e6 in magnitude is skipped, so the last VALID body wins.  One thread per (palette, driven bone).
//   boneList [nBones] driven bones; bodyStart [nBones+1] into bodyIds (ascending body index); offInv [nBodies][16] col-major;
//   posQuat [P][nBodies][7] = x,y,z, qx,qy,qz,qw.
__global__ void apply_bodies_kernel(const uint32_t* __restrict__ boneList, const uint32_t* __restrict__ bodyStart,
                                    const uint32_t* __restrict__ bodyIds, const float* __restrict__ offInv, const float* __restrict__ posQuat,
                                    const float4* __restrict__ invBind, const uint32_t* __restrict__ bonePos, float4* __restrict__ skin,
                                    uint32_t nBones, uint32_t nBodies, uint32_t P, uint32_t B, uint32_t soa) {
  const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= P * nBones) return;
  const uint32_t p = idx / nBones, li = idx - p * nBones, bone = boneList[li];
  float W[4][4];                                        // W[r][c]
  bool have = false;
  for (uint32_t k = bodyStart[li]; k < bodyStart[li + 1]; ++k) {
    const uint32_t body = bodyIds[k];
    const float* pq = posQuat + ((size_t)p * nBodies + body) * 7;
    float R[3][3];
    q_to_rows(make_float4(pq[3], pq[4], pq[5], pq[6]), R);           // math.ts:352-384 via fromPositionRotation (387-393)
    const float* o = offInv + (size_t)body * 16;                        // column-major: o[c*4 + r]
    float N[4][4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
#pragma unroll
      for (int r = 0; r < 3; ++r) N[r][c] = R[r][0] * o[c * 4] + R[r][1] * o[c * 4 + 1] + R[r][2] * o[c * 4 + 2] + pq[r] * o[c * 4 + 3];
      N[3][c] = o[c * 4 + 3];                                           // bottom row of the node matrix is (0,0,0,1)
    }
    const bool ok = !isnan(N[0][0]) && !isnan(N[3][3]) && fabsf(N[0][0]) < 1e6f && fabsf(N[3][3]) < 1e6f;
    if (ok) {
      have = true;
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) W[r][c] = N[r][c];
    }
  }
  if (!have) return;
  float S[3][4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const float4 ib = __ldg(invBind + (size_t)bone * 4 + c);
#pragma unroll
    for (int r = 0; r < 3; ++r) S[r][c] = W[r][0] * ib.x + W[r][1] * ib.y + W[r][2] * ib.z + W[r][3] * ib.w;
  }
  const float4 cA = make_float4(S[0][0], S[1][0], S[0][1], S[1][1]);
  const float4 cB = make_float4(S[0][2], S[1][2], S[0][3], S[1][3]);
  const float4 cC = make_float4(S[2][0], S[2][1], S[2][2], S[2][3]);
  const size_t pos = __ldg(bonePos + bone);
  if (soa) {
    const size_t pb = (size_t)p * B * 3;
    skin[pb + pos] = cA; skin[pb + B + pos] = cB; skin[pb + 2 * (size_t)B + pos] = cC;
  } else {
    const size_t row = (size_t)p * B + pos;
    skin[row * 3] = cA; skin[row * 3 + 1] = cB; skin[row * 3 + 2] = cC;
  }
}

__global__ void bounds_reset_kernel(int* b, uint32_t n6) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n6) b[i] = (i % 6) < 3 ? 0x7F7FFFFF : (int)(0x7F7FFFFF ^ 0x7FFFFFFF) | (int)0x80000000;
}

}  // namespace rz
