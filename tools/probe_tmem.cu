// probe_tmem.cu — microbenchmarks behind the epilogue design of the fused MLP kernels (B200):
//   (1) TMEM read throughput of tcgen05.ld 32x32b (.x16 / .x32 / .x64) for 4 / 8 / 16 reader warps,
//   (2) the cost of a relu+pack+swizzled-smem-store epilogue pass on top of it,
//   (3) back-to-back 128x128x128 SS MMAs (issue -> commit -> wait) latency.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/probe_tmem tools/probe_tmem.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../modulus_b200/csrc/mgn_tc.cuh"
#include "../modulus_b200/csrc/mgn_tile.cuh"
using namespace mgn;
using namespace mgn::tile;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2);} } while (0)

__device__ __forceinline__ void tmem_ld64(uint32_t taddr, uint32_t (&v)[64]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
      "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
      "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]),
        "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]),
        "=r"(v[40]), "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]),
        "=r"(v[48]), "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]),
        "=r"(v[56]), "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
      : "r"(taddr) : "memory");
}

// mode 0: ld only (x16), 1: ld only (x32), 2: ld only (x64), 3: x32 + relu/pack/store to swizzled smem,
// 4: x32 with both halves issued before the wait + relu/pack/store, 5: smem read-modify-write only (no TMEM)
__global__ void __launch_bounds__(512, 1) tmem_read_kernel(int mode, int n_warps, int iters, long long* out, float* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) tmem_alloc(&slot, 128);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = slot;
  float acc = 0.f;
  long long t0 = 0, t1 = 0;
  if (warp < n_warps) {
    const int q = warp & 3, part = warp >> 2, nparts = n_warps >> 2;  // column split among warps of a lane quarter
    const int cols = 128 / nparts, c0 = part * cols;
    const int row = q * 32 + lane;
    const uint32_t t = tmem + (static_cast<uint32_t>(q * 32) << 16) + c0;
    asm volatile("bar.sync 1, %0;" ::"r"(n_warps * 32));
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (mode == 0) {
        for (int c = 0; c < cols; c += 16) { uint32_t v[16]; tmem_ld16(t + c, v); tmem_ld_wait(); for (int j = 0; j < 16; ++j) acc += __uint_as_float(v[j]); }
      } else if (mode == 1) {
        for (int c = 0; c < cols; c += 32) { uint32_t v[32]; tmem_ld32(t + c, v); tmem_ld_wait(); for (int j = 0; j < 32; ++j) acc += __uint_as_float(v[j]); }
      } else if (mode == 2) {
        for (int c = 0; c < cols; c += 64) { uint32_t v[64]; tmem_ld64(t + c, v); tmem_ld_wait(); for (int j = 0; j < 64; ++j) acc += __uint_as_float(v[j]); }
      } else if (mode == 3) {
        for (int c = 0; c < cols; c += 32) {
          uint32_t v[32]; tmem_ld32(t + c, v); tmem_ld_wait();
          uint32_t o[16];
          for (int j = 0; j < 16; ++j) o[j] = pack_bf16x2(fmaxf(__uint_as_float(v[2*j]) + 0.5f, 0.f), fmaxf(__uint_as_float(v[2*j+1]) + 0.25f, 0.f));
          row_store32p(smem, row, c0 + c, o);
        }
      } else if (mode == 4) {
        if (cols == 64) {
          uint32_t v[32], w[32]; tmem_ld32(t, v); tmem_ld32(t + 32, w); tmem_ld_wait();
          uint32_t o[16];
          for (int j = 0; j < 16; ++j) o[j] = pack_bf16x2(fmaxf(__uint_as_float(v[2*j]) + 0.5f, 0.f), fmaxf(__uint_as_float(v[2*j+1]) + 0.25f, 0.f));
          row_store32p(smem, row, c0, o);
          for (int j = 0; j < 16; ++j) o[j] = pack_bf16x2(fmaxf(__uint_as_float(w[2*j]) + 0.5f, 0.f), fmaxf(__uint_as_float(w[2*j+1]) + 0.25f, 0.f));
          row_store32p(smem, row, c0 + 32, o);
        }
      } else if (mode == 5) {
        for (int c = 0; c < cols; c += 32) {
          uint32_t h[16]; row_load32p(smem, row, c0 + c, h);
          for (int j = 0; j < 16; ++j) h[j] = pack_bf16x2(bf_lo(h[j]) + 1.f, bf_hi(h[j]) + 1.f);
          row_store32p(smem, row, c0 + c, h);
        }
      }
      asm volatile("bar.sync 1, %0;" ::"r"(n_warps * 32));
    }
    t1 = clock64();
  }
  if (tid == 0) out[blockIdx.x] = (t1 - t0) / iters;
  if (acc == 123.456f) sink[0] = acc;
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

// back-to-back GEMM latency: issue 8 MMAs (128x128x16 each) + commit + wait, repeated
__global__ void __launch_bounds__(128, 1) mma_chain_kernel(int iters, int n_mma, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t slot;
  __shared__ uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 4 * kPB / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc(&slot, 128);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = slot;
  if (tid == 0) {
    const uint32_t aA = smem_u32(smem), aB = aA + 2 * kPB;
    const uint32_t idesc = umma_idesc_bf16(128, 128, 0, 0);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      for (int k = 0; k < n_mma; ++k)
        umma_ss(tmem, umma_desc_kmajor(aA + ((k & 7) >> 2) * kPB, k & 3), umma_desc_kmajor(aB + ((k & 7) >> 2) * kPB, k & 3), idesc, k != 0);
      umma_commit(&bar);
      while (!mbar_try_wait(&bar, it & 1)) {}
      tc_fence_after_sync();
    }
    out[blockIdx.x] = (clock64() - t0) / iters;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

int main() {
  long long* d_out; float* d_sink;
  CK(cudaMalloc(&d_out, 148 * 8)); CK(cudaMalloc(&d_sink, 4));
  CK(cudaFuncSetAttribute(tmem_read_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * kPB));
  CK(cudaFuncSetAttribute(mma_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * kPB));
  const char* names[] = {"ld.x16 only", "ld.x32 only", "ld.x64 only", "ld.x32 + relu/pack/STS", "2x ld.x32 then wait + relu/pack/STS", "smem RMW only"};
  for (int mode = 0; mode < 6; ++mode)
    for (int nw = 4; nw <= 16; nw *= 2) {
      if (mode == 4 && nw != 8) continue;
      if (mode == 2 && nw == 16) continue;
      tmem_read_kernel<<<148, 512, 2 * kPB>>>(mode, nw, 200, d_out, d_sink);
      CK(cudaDeviceSynchronize());
      long long h[148]; CK(cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost));
      printf("%-40s warps=%2d : %lld cycles per 128x128 fp32 accumulator pass (64 KB)\n", names[mode], nw, h[0]);
    }
  for (int n_mma = 8; n_mma <= 32; n_mma *= 2) {
    mma_chain_kernel<<<148, 128, 4 * kPB>>>(200, n_mma, d_out);
    CK(cudaDeviceSynchronize());
    long long h[148]; CK(cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost));
    printf("MMA chain: %2d x (128x128x16) SS MMAs + commit + wait : %lld cycles\n", n_mma, h[0]);
  }
  return 0;
}
