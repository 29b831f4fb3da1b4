// ubench_lds2.cu — LDS.128 cost for palette-like gathers on sm_100a: d distinct palette rows per warp instruction,
// (a) 48-byte rows (AoS) with / without 16-byte bank-group collisions, (b) [3][B] float4 (SoA by chunk) with the
// rows inside one aligned group of 8 / spread.  Calibrates the layout choice of the bone palette in shared memory.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <stdint.h>
__device__ __forceinline__ float4 lds128(uint32_t a) {
  float4 r;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(a) : "memory");
  return r;
}
// layout 0: address = row*48 + r*16 ; layout 1: address = r*512*16 + row*16   (512 rows)
__global__ void k(const uint32_t* __restrict__ rowOfLane, float* out, long long* cyc, int iters, int layout) {
  extern __shared__ float4 sm[];
  for (int i = threadIdx.x; i < 512 * 3; i += blockDim.x) sm[i] = make_float4(i, 1, 2, 3);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const uint32_t sb = (uint32_t)__cvta_generic_to_shared(sm);
  const uint32_t pS = layout ? 16u : 48u, rS = layout ? 512u * 16u : 16u;
  uint32_t row = rowOfLane[lane];
  float acc = 0.f;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const uint32_t base = sb + ((row + 64u * ((it + u) & 1)) & 511u) * pS;   // +64 rows keeps bank group and 8-group alignment
      float4 a = lds128(base), b = lds128(base + rS), c = lds128(base + 2 * rS);
      acc += (a.x + a.y + a.z + a.w) + (b.x + b.y + b.z + b.w) + (c.x + c.y + c.z + c.w);
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  float* out; long long* cyc; uint32_t* d_rows;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 64); cudaMalloc(&d_rows, 128);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 512 * 48);
  const int iters = 1000, threads = 512;
  srand(7);
  auto run = [&](const char* name, std::vector<uint32_t> rows, int layout) {
    cudaMemcpy(d_rows, rows.data(), 128, cudaMemcpyHostToDevice);
    for (int rep = 0; rep < 2; ++rep) { k<<<1, threads, 512 * 48>>>(d_rows, out, cyc, iters, layout); cudaDeviceSynchronize(); }
    long long h = 0; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-74s %6.2f cycles per LDS.128  [%s]\n", name, (double)h / (iters * 24.0) / (threads / 32), cudaGetErrorString(cudaGetLastError()));
  };
  for (int d : {1, 2, 3, 4, 6, 8, 16}) {
    for (int mode = 0; mode < 5; ++mode) {
      std::vector<uint32_t> bones(d);
      int layout = mode >= 3;
      for (int i = 0; i < d; ++i) {
        if (mode == 0) bones[i] = (i % 8) + 8 * (1 + (rand() % 6)) + 64 * (i / 8);           // AoS, distinct bank groups, far apart
        else if (mode == 1) bones[i] = 8 * (1 + i);                                         // AoS, all same bank group
        else if (mode == 2) bones[i] = ((i / 2) % 8) + 8 * (1 + i);                          // AoS, pairwise colliding
        else if (mode == 3) bones[i] = 16 + i;                                              // SoA, consecutive rows (one or two 128 B lines)
        else bones[i] = (i % 8) + 8 * (1 + 2 * i);                                          // SoA, each row in a different line, distinct slot
      }
      std::vector<uint32_t> rows(32);
      for (int l = 0; l < 32; ++l) rows[l] = bones[l * d / 32];                             // contiguous lane runs (lanes are sorted by bone)
      char name[160];
      const char* mn[] = {"AoS 48B rows, distinct bank groups", "AoS 48B rows, all in one bank group", "AoS 48B rows, pairwise colliding",
                          "SoA [3][B], consecutive rows (same 128B line)", "SoA [3][B], rows in different lines"};
      snprintf(name, sizeof name, "d=%2d  %s", d, mn[mode]);
      run(name, rows, layout);
    }
  }
  return 0;
}
