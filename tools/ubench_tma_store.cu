// ubench_tma_store.cu — sustained shared->global cp.async.bulk store bandwidth vs. op size on sm_100a.
// Each warp owns a double-buffered staging region; per iteration: 32 lanes write S bytes to smem, fence.proxy.async,
// one lane issues `nops` bulk stores of S/nops bytes, commit, wait_group.read 1.  Output range >> L2.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
template <int OPB>   // bytes per bulk op
__global__ void k(float* out, size_t bytesPerWarp, int iters) {
  extern __shared__ __align__(128) unsigned char sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int S = 3072;                      // bytes staged per warp per iteration (fixed work), in S/OPB ops
  const uint32_t base = smem_u32(sm) + warp * (2 * S);
  const size_t gw = (size_t)blockIdx.x * (blockDim.x >> 5) + warp;
  char* dst = reinterpret_cast<char*>(out) + gw * bytesPerWarp;
  uint32_t buf = 0;
  for (int it = 0; it < iters; ++it) {
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    __syncwarp();
    const uint32_t sb = base + buf * S;
    for (int o = lane * 4; o < S; o += 128) asm volatile("st.shared.f32 [%0], %1;" ::"r"(sb + o), "f"((float)it) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
#pragma unroll
      for (int o = 0; o < S; o += OPB) bulk_s2g(dst + (size_t)it * S + o, sb + o, OPB);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    buf ^= 1;
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
template <int OPB>
void run(int warpsPerCta, int ctasPerSm) {
  const int iters = 400, sms = 148;
  const int ctas = sms * ctasPerSm;
  const size_t bytesPerWarp = (size_t)iters * 3072;
  const size_t total = (size_t)ctas * warpsPerCta * bytesPerWarp;
  float* out; cudaMalloc(&out, total);
  const int smem = warpsPerCta * 2 * 3072;
  cudaFuncSetAttribute(k<OPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<OPB><<<ctas, warpsPerCta * 32, smem>>>(out, bytesPerWarp, iters);
  cudaEventRecord(e0);
  k<OPB><<<ctas, warpsPerCta * 32, smem>>>(out, bytesPerWarp, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("op=%5d B  warps/CTA=%2d CTAs/SM=%d  total %.2f GB  %.3f ms  %.0f GB/s  (%s)\n", OPB, warpsPerCta, ctasPerSm, total / 1e9, ms,
         total / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}
int main() {
  for (int w : {8, 16}) for (int c : {1, 2, 3}) { if (w * c > 48) continue; run<384>(w, c); run<768>(w, c); run<1536>(w, c); run<3072>(w, c); }
  return 0;
}
