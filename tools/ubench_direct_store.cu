// ubench_direct_store.cu — can registers -> global stores replace the shared-memory staging + TMA bulk store of the deform
// kernel?  A lane that owns 4 (or 8) CONSECUTIVE output vertices holds 48 (96) contiguous bytes per plane and can write them
// with 3 (6) st.global.v4 — no STS, no TMA read of the staging buffer (12 of the kernel's ~32 shared-memory cycles per
// warp-instance).  The catch: within one instruction the 32 lanes write 16 B out of every 48 (96) B, i.e. partial sectors
// that only complete over the 3 (6) instructions.  This measures what that pattern sustains against HBM, next to fully
// coalesced STG.128 and to the bulk-store baseline (ubench_tma_store.cu, same per-warp streaming layout, output >> L2).
//   mode 0: coalesced      lane l, instr c -> base + c*512 + l*16
//   mode 1: 48-byte chunks lane l, instr c -> base + l*48  + c*16      (4 vertices x 12 B per lane)
//   mode 2: 96-byte chunks lane l, instr c -> base + l*96  + c*16      (8 vertices per lane, whole 32-byte sectors per lane)
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
__device__ __forceinline__ void stg_cs(void* p, float4 v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void stg_ef(void* p, float4 v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}
template <int MODE, int HINT>
__global__ void k(float* out, size_t bytesPerWarp, int iters) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int S = MODE == 2 ? 3072 : 1536;               // bytes per warp per iteration
  constexpr int NI = S / 512;                              // STG.128 per lane per iteration
  const size_t gw = (size_t)blockIdx.x * (blockDim.x >> 5) + warp;
  char* dst = reinterpret_cast<char*>(out) + gw * bytesPerWarp;
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  for (int it = 0; it < iters; ++it) {
    char* b = dst + (size_t)it * S;
#pragma unroll
    for (int c = 0; c < NI; ++c) {
      char* p = MODE == 0 ? b + c * 512 + lane * 16 : b + lane * (S / 32) + c * 16;
      const float4 v = make_float4((float)it, (float)c, (float)lane, 1.f);
      if (HINT) stg_ef(p, v, pol); else stg_cs(p, v);
    }
  }
}
template <int MODE, int HINT>
void run(int warpsPerCta, int ctasPerSm) {
  const int sms = 148, ctas = sms * ctasPerSm;
  constexpr int S = MODE == 2 ? 3072 : 1536;
  const int iters = MODE == 2 ? 400 : 800;
  const size_t bytesPerWarp = (size_t)iters * S;
  const size_t total = (size_t)ctas * warpsPerCta * bytesPerWarp;
  float* out; cudaMalloc(&out, total);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE, HINT><<<ctas, warpsPerCta * 32>>>(out, bytesPerWarp, iters);
  cudaEventRecord(e0);
  k<MODE, HINT><<<ctas, warpsPerCta * 32>>>(out, bytesPerWarp, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  static const char* names[] = {"coalesced STG.128", "48 B per lane (3 STG.128)", "96 B per lane (6 STG.128)"};
  printf("%-28s %s  warps/CTA=%2d CTAs/SM=%d  total %.2f GB  %.3f ms  %.0f GB/s  (%s)\n", names[MODE], HINT ? "L2::evict_first" : "st.cs          ",
         warpsPerCta, ctasPerSm, total / 1e9, ms, total / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}
int main() {
  for (int w : {8, 16}) for (int c : {1, 2}) {
    run<0, 0>(w, c); run<1, 0>(w, c); run<2, 0>(w, c);
    run<0, 1>(w, c); run<1, 1>(w, c); run<2, 1>(w, c);
  }
  return 0;
}
