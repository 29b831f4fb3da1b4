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

// dense per-instance morph weights: dense[k][m] = 0, dense[k][activeIds[a]] += w[k][a]
__global__ void morph_weights_kernel(const float* __restrict__ w, const uint32_t* __restrict__ ids, float* __restrict__ dense,
                                     uint32_t K, uint32_t Mact, uint32_t Mpad) {
  const uint32_t k = blockIdx.x;
  if (k >= K) return;
  for (uint32_t m = threadIdx.x; m < Mpad; m += blockDim.x) dense[(size_t)k * Mpad + m] = 0.f;
  __syncthreads();
  for (uint32_t a = threadIdx.x; a < Mact; a += blockDim.x) atomicAdd(&dense[(size_t)k * Mpad + ids[a]], w[(size_t)k * Mact + a]);
}

__global__ void bounds_reset_kernel(int* b, uint32_t n6) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n6) b[i] = (i % 6) < 3 ? 0x7F7FFFFF : (int)(0x7F7FFFFF ^ 0x7FFFFFFF) | (int)0x80000000;
}

}  // namespace rz
