// ubench_lds.cu — how many cycles does a warp-wide LDS.128 cost for different address patterns on sm_100a?
// (decides whether warp-/quarter-uniform bone gathers are cheaper than scattered ones).  nvcc -arch=sm_100a, run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

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

// mode: 0 all lanes same 16B; 1 each quarter-warp same 16B (4 distinct, distinct bank groups); 2 all lanes distinct, conflict-free
//       3 8 distinct 16B chunks spread over lanes round-robin (lane%8), distinct bank groups; 4 pairs of lanes share (16 distinct)
//       5 LDS.32 all same; 6 LDS.32 distinct conflict-free; 7: quarter-warps same chunk but all four quarters the SAME bank group (different rows)
template <int MODE>
__global__ void k(float* out, long long* cyc, int iters, int nwarps_active) {
  extern __shared__ float4 sm[];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = make_float4(i, 1, 2, 3);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  uint32_t base = (uint32_t)__cvta_generic_to_shared(sm);
  uint32_t off;
  if (MODE == 0 || MODE == 5) off = 0;
  else if (MODE == 1) off = (lane / 8) * 16;
  else if (MODE == 2) off = lane * 16;
  else if (MODE == 3) off = (lane % 8) * 16;
  else if (MODE == 4) off = (lane / 2) * 16;
  else if (MODE == 6) off = lane * 4;
  else off = (lane / 8) * 128;   // 7
  float acc = 0.f;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const uint32_t a = base + off + ((it * 16 + u) & 63) * 512;
      if (MODE == 5 || MODE == 6) acc += lds32(a);
      else { float4 v = lds128(a); acc += v.x + v.w; }
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name) {
  float* out; long long* cyc;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1024);
  const int iters = 2000, threads = 512;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  k<MODE><<<1, threads, 65536>>>(out, cyc, iters, 0);
  cudaDeviceSynchronize();
  k<MODE><<<1, threads, 65536>>>(out, cyc, iters, 0);
  cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const double per_warp_instr = (double)h / (iters * 16.0) / (threads / 32);   // cycles per warp-level load with 16 warps contending
  printf("%-60s %8.3f cycles per warp-instruction (SM-wide throughput, 16 warps)\n", name, per_warp_instr);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0>("LDS.128 all 32 lanes same address");
  run<1>("LDS.128 quarter-warp uniform (4 distinct chunks, 4 bank groups)");
  run<7>("LDS.128 quarter-warp uniform, all quarters same bank group");
  run<3>("LDS.128 8 distinct chunks, lane%8 (every quarter sees all 8)");
  run<4>("LDS.128 16 distinct chunks (pairs share)");
  run<2>("LDS.128 32 distinct chunks, conflict-free");
  run<5>("LDS.32 all same address");
  run<6>("LDS.32 32 distinct, conflict-free");
  return 0;
}
