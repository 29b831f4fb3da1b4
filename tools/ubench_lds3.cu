// ubench_lds3.cu — cost of fetching one 48-byte palette row per lane with LDS.128 x3 / LDS.64 x6 / LDS.32 x12 on sm_100a,
// for lane->row patterns like the deform kernel's (lanes sorted by bone: runs of equal rows, arbitrary lengths).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <stdint.h>
template <int W>
__global__ void k(const uint32_t* __restrict__ rowOfLane, float* out, long long* cyc, int iters) {
  extern __shared__ float4 sm[];
  for (int i = threadIdx.x; i < 1024 * 3; i += blockDim.x) sm[i] = make_float4(i, 1, 2, 3);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const uint32_t sb = (uint32_t)__cvta_generic_to_shared(sm);
  const uint32_t row = rowOfLane[lane];
  float acc = 0.f;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    // a different (but equally patterned) set of rows every iteration so nothing can be CSE'd; +8 rows keeps bank groups
    const uint32_t base = sb + ((row + 8u * (uint32_t)(it & 63)) & 1023u) * 48u;
    if (W == 16) {
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        float4 v;
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(base + r * 16) : "memory");
        acc += v.x + v.y + v.z + v.w;
      }
    } else if (W == 8) {
#pragma unroll
      for (int r = 0; r < 6; ++r) {
        float2 v;
        asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(base + r * 8) : "memory");
        acc += v.x + v.y;
      }
    } else {
#pragma unroll
      for (int r = 0; r < 12; ++r) {
        float v;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(base + r * 4) : "memory");
        acc += v;
      }
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
template <int W>
double run(uint32_t* d_rows, float* out, long long* cyc, const std::vector<uint32_t>& rows) {
  const int iters = 4000, threads = 512;
  cudaMemcpy(d_rows, rows.data(), 128, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(k<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, 1024 * 48);
  for (int rep = 0; rep < 2; ++rep) { k<W><<<1, threads, 1024 * 48>>>(d_rows, out, cyc, iters); cudaDeviceSynchronize(); }
  long long h = 0; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  return (double)h / iters / (threads / 32);     // SM cycles per warp to fetch one 48-byte row per lane
}
int main() {
  float* out; long long* cyc; uint32_t* d_rows;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 64); cudaMalloc(&d_rows, 128);
  srand(3);
  struct Pat { const char* name; std::vector<uint32_t> rows; };
  std::vector<Pat> pats;
  auto runs = [&](std::vector<int> lens, bool distinctGroups) {
    std::vector<uint32_t> rows; int i = 0;
    for (int len : lens) { uint32_t r = distinctGroups ? (uint32_t)(i % 8) + 16u * (uint32_t)(1 + i) : (uint32_t)(8 * (3 + 2 * i)); for (int t = 0; t < len; ++t) rows.push_back(r); ++i; }
    rows.resize(32, rows.back()); return rows;
  };
  pats.push_back({"1 row (warp-uniform)", runs({32}, true)});
  pats.push_back({"4 rows, runs 8/8/8/8, distinct bank groups", runs({8, 8, 8, 8}, true)});
  pats.push_back({"4 rows, runs 11/9/7/5, distinct bank groups", runs({11, 9, 7, 5}, true)});
  pats.push_back({"4 rows, runs 11/9/7/5, SAME bank group", runs({11, 9, 7, 5}, false)});
  pats.push_back({"6 rows, runs 9/7/6/5/3/2, distinct bank groups", runs({9, 7, 6, 5, 3, 2}, true)});
  pats.push_back({"8 rows, runs of 4, distinct bank groups", runs({4, 4, 4, 4, 4, 4, 4, 4}, true)});
  { std::vector<uint32_t> r(32); for (int l = 0; l < 32; ++l) r[l] = (uint32_t)(l % 8) + 16u * (uint32_t)(1 + l); pats.push_back({"32 distinct rows", r}); }
  printf("SM cycles per warp to fetch one 48-byte palette row per lane (16 warps resident, shared-memory pipe saturated)\n");
  printf("%-52s %10s %10s %10s\n", "lane -> row pattern", "3xLDS.128", "6xLDS.64", "12xLDS.32");
  for (auto& p : pats)
    printf("%-52s %10.2f %10.2f %10.2f\n", p.name, run<16>(d_rows, out, cyc, p.rows), run<8>(d_rows, out, cyc, p.rows), run<4>(d_rows, out, cyc, p.rows));
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
