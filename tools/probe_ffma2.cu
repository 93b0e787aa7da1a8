// probe_ffma2.cu -- does fma.rn.f32x2 (SASS FFMA2) double fp32 throughput on sm_100a, or is it cracked into two FFMAs?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/probe_ffma2 tools/probe_ffma2.cu && tools/bin/probe_ffma2
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t pk(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }

template <int MODE>
__global__ void k(float* out, int iters, float s) {
  float x = threadIdx.x * 1e-3f;
  if (MODE == 0) {  // 16 independent scalar chains = 16 fp32 FMAs per iteration
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = fma1(a[i], s, x);
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) r += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  } else {  // 8 independent packed chains = 16 fp32 FMAs per iteration
    uint64_t a[8];
    const uint64_t ss = pk(s, s), xx = pk(x, x);
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = pk(x + i, x - i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = fma2(a[i], ss, xx);
    }
    uint64_t r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(static_cast<uint32_t>(r ^ (r >> 32)));
  }
}

int main() {
  float* out;
  cudaMalloc(&out, 148 * 1024 * sizeof(float));
  const int iters = 20000;
  for (int warps : {4, 8, 16, 32}) {
    for (int mode = 0; mode < 2; ++mode) {
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        if (mode == 0) k<0><<<148, warps * 32>>>(out, iters, 1.0001f); else k<1><<<148, warps * 32>>>(out, iters, 1.0001f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
      }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      const double fma_per_sm = double(iters) * 16 * warps * 32;
      printf("warps/SM %2d  %s: %.3f ms  -> %.1f fp32 FMA / cycle / SM at 1.9 GHz\n", warps, mode ? "FFMA2" : "FFMA ", ms,
             fma_per_sm / (ms * 1e-3 * 1.9e9));
    }
  }
  return 0;
}
