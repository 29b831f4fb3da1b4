// ubench_tma_streams.cu — what limits warp-private cp.async.bulk stores: the op size, or the number of output streams a
// CTA feeds at once?  Emulates the deform kernels' store pattern: a CTA of W warps walks "passes"; in pass p warp w owns chunk
// (p * W + w) of EACH of NS streams (stream = one output plane of one instance, 1 MiB apart and more), CH bytes per chunk;
// chunks of a pass are adjacent within a stream (W * CH contiguous bytes per stream per pass), CTAs own disjoint streams.
// planar pos+nrm, I = 4: NS = 8, CH = 768.   interleaved, I = 4: NS = 4, CH = 2048.   hull: NS = 12, CH = 768.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
template <int NS, int CH>
__global__ void k(char* out, size_t streamBytes, int passes) {
  extern __shared__ __align__(128) unsigned char sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
  constexpr int S = NS * CH;                                   // bytes staged per warp per pass
  const uint32_t base = smem_u32(sm) + warp * (2 * S);
  char* cta = out + (size_t)blockIdx.x * NS * streamBytes;
  uint32_t buf = 0;
  for (int p = 0; p < passes; ++p) {
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    __syncwarp();
    const uint32_t sb = base + buf * S;
    for (int o = lane * 4; o < S; o += 128) asm volatile("st.shared.f32 [%0], %1;" ::"r"(sb + o), "f"((float)p) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
#pragma unroll
      for (int s = 0; s < NS; ++s) bulk_s2g(cta + (size_t)s * streamBytes + ((size_t)p * W + warp) * CH, sb + s * CH, CH);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    buf ^= 1;
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
template <int NS, int CH>
void run(int warps, const char* what) {
  const int ctas = 148, passes = 600 * 768 / CH * 8 / NS;      // ~equal bytes per configuration
  const size_t streamBytes = ((size_t)passes * warps * CH + (1 << 20) - 1) >> 20 << 20;
  const size_t total = (size_t)ctas * NS * passes * warps * CH;
  char* out; cudaMalloc(&out, (size_t)ctas * NS * streamBytes);
  const int smem = warps * 2 * NS * CH;
  cudaFuncSetAttribute(k<NS, CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<NS, CH><<<ctas, warps * 32, smem>>>(out, streamBytes, passes);
  cudaEventRecord(e0);
  k<NS, CH><<<ctas, warps * 32, smem>>>(out, streamBytes, passes);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("%-34s streams/CTA=%2d chunk=%4d B warps=%2d  total %.2f GB  %.3f ms  %.0f GB/s  (%s)\n", what, NS, CH, warps, total / 1e9, ms,
         total / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}
int main() {
  for (int rep = 0; rep < 3; ++rep) {
    run<8, 768>(8, "planar I=4, 768 B ops");
    run<8, 1536>(8, "planar I=4, 1536 B ops");
    run<8, 768>(16, "planar I=4, 768 B ops");
    run<4, 2048>(8, "interleaved I=4");
    run<6, 1536>(8, "hull as 6x1536");
    run<12, 768>(8, "hull I=4");
  }
  return 0;
}
