// probe_tf32.cu — hardware check for fp32 callers on the tensor cores (verdict item 6):
// tcgen05.mma kind::tf32 on fp32 operands split as x = hi + lo (hi = tf32(x), lo = tf32(x - hi)).
//   mode 0: D = A_hi B_hi^T                                   (1 x TF32: ~2^-11 per operand)
//   mode 1: D = A_lo B_hi^T + A_hi B_lo^T + A_hi B_hi^T       (3 x TF32: lo*lo dropped, ~2^-21)
//   mode 2: the same MMA stream repeated R times, timed with clock64 (cycles per 128x128x8 tf32 instruction)
//   mode 3: bf16 MMA stream (128x128x16 per instruction) timed the same way, for the ratio
// One CTA, D = 128 x 128, K = 64; operands K-major in 128-byte-swizzle panels of 32 fp32 columns (the layout arithmetic is
// the bf16 one: a k-step is 32 bytes either way).  Errors are reported against a float64 product of the fp32 inputs, next to
// the error of a plain fp32 dot product -- the bar an "exact fp32" path is held to.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o probe_tf32 probe_tf32.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../modulus_b200/csrc/mgn_tc.cuh"

using namespace mgn;

#define CK(x)                                                                           \
  do {                                                                                  \
    cudaError_t e_ = (x);                                                               \
    if (e_ != cudaSuccess) {                                                            \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);  \
      exit(2);                                                                          \
    }                                                                                   \
  } while (0)

constexpr int T = 128;                // rows of A, rows of B (= columns of D)
constexpr int K = 64;                 // fp32 reduction length: 2 panels of 32 columns
constexpr int PANEL_BYTES = T * 128;  // 128 rows x 128 B

// kind::tf32 instruction descriptor: c_format = 1 (f32), a_format = b_format = 2 (tf32), K-major operands
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}

__device__ __forceinline__ void umma_ss_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}

__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

__global__ void __launch_bounds__(128, 1)
probe_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int mode, int reps,
             long long* cycles, int* err) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sAh = smem;                     // 2 panels each
  uint8_t* sAl = smem + 2 * PANEL_BYTES;
  uint8_t* sBh = smem + 4 * PANEL_BYTES;
  uint8_t* sBl = smem + 6 * PANEL_BYTES;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;

  // thread t stages row t of A and of B: 16 chunks of 4 floats, hi and lo parts
  for (int c = 0; c < K / 4; ++c) {
    const int panel = c >> 3, ch = c & 7;
    const float4 a = *reinterpret_cast<const float4*>(A + tid * K + c * 4);
    const float4 b = *reinterpret_cast<const float4*>(B + tid * K + c * 4);
    float4 ah, al, bh, bl;
    ah.x = to_tf32(a.x); ah.y = to_tf32(a.y); ah.z = to_tf32(a.z); ah.w = to_tf32(a.w);
    al.x = to_tf32(a.x - ah.x); al.y = to_tf32(a.y - ah.y); al.z = to_tf32(a.z - ah.z); al.w = to_tf32(a.w - ah.w);
    bh.x = to_tf32(b.x); bh.y = to_tf32(b.y); bh.z = to_tf32(b.z); bh.w = to_tf32(b.w);
    bl.x = to_tf32(b.x - bh.x); bl.y = to_tf32(b.y - bh.y); bl.z = to_tf32(b.z - bh.z); bl.w = to_tf32(b.w - bh.w);
    const uint32_t off = panel * PANEL_BYTES + sw128_offset(tid, ch);
    *reinterpret_cast<float4*>(sAh + off) = ah;
    *reinterpret_cast<float4*>(sAl + off) = al;
    *reinterpret_cast<float4*>(sBh + off) = bh;
    *reinterpret_cast<float4*>(sBl + off) = bl;
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 128);
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;

  long long t0 = 0;
  if (tid == 0) {
    const uint32_t ah = smem_u32(sAh), al = smem_u32(sAl), bh = smem_u32(sBh), bl = smem_u32(sBl);
    const uint32_t id32 = idesc_tf32(128, 128), id16 = umma_idesc_bf16(128, 128, 0, 0);
    t0 = clock64();
    if (mode == 0 || mode == 1) {
      uint32_t acc = 0;
      if (mode == 1) {  // small terms first
        for (int j = 0; j < K / 8; ++j, acc = 1)
          umma_ss_tf32(tmem, umma_desc_kmajor(al + (j >> 2) * PANEL_BYTES, j & 3), umma_desc_kmajor(bh + (j >> 2) * PANEL_BYTES, j & 3), id32, acc);
        for (int j = 0; j < K / 8; ++j)
          umma_ss_tf32(tmem, umma_desc_kmajor(ah + (j >> 2) * PANEL_BYTES, j & 3), umma_desc_kmajor(bl + (j >> 2) * PANEL_BYTES, j & 3), id32, 1);
      }
      for (int j = 0; j < K / 8; ++j, acc = 1)
        umma_ss_tf32(tmem, umma_desc_kmajor(ah + (j >> 2) * PANEL_BYTES, j & 3), umma_desc_kmajor(bh + (j >> 2) * PANEL_BYTES, j & 3), id32, acc);
    } else if (mode == 2) {
      for (int r = 0; r < reps; ++r)
        for (int j = 0; j < K / 8; ++j)
          umma_ss_tf32(tmem, umma_desc_kmajor(ah + (j >> 2) * PANEL_BYTES, j & 3), umma_desc_kmajor(bh + (j >> 2) * PANEL_BYTES, j & 3), id32, (r | j) != 0);
    } else {
      for (int r = 0; r < reps; ++r)  // the same bytes read as bf16 pairs: values are meaningless, the timing is not
        for (int j = 0; j < K / 8; ++j)
          umma_ss(tmem, umma_desc_kmajor(ah + (j >> 2) * PANEL_BYTES, j & 3), umma_desc_kmajor(bh + (j >> 2) * PANEL_BYTES, j & 3), id16, (r | j) != 0);
    }
    umma_commit(&bar);
  }
  __syncwarp();
  if (!mbar_wait(&bar, 0)) {
    if (tid == 0) *err = 1;
  }
  if (tid == 0) *cycles = clock64() - t0;
  tc_fence_after_sync();

  const uint32_t lane_addr = tmem + (static_cast<uint32_t>(warp * 32) << 16);
  for (int g = 0; g < 4; ++g) {
    uint32_t v[32];
    tmem_ld32(lane_addr + g * 32, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) D[tid * T + g * 32 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

int main() {
  std::vector<float> hA(T * K), hB(T * K), hD(T * T);
  srand(4321);
  for (int i = 0; i < T * K; ++i) {
    hA[i] = (rand() / (float)RAND_MAX) * 2.f - 1.f;
    hB[i] = (rand() / (float)RAND_MAX) * 2.f - 1.f;
  }
  std::vector<double> ref(T * T);
  std::vector<float> ref32(T * T);
  double scale = 0;  // sum |a||b| per element: the natural error scale of a dot product
  for (int m = 0; m < T; ++m)
    for (int n = 0; n < T; ++n) {
      double s = 0, sa = 0;
      float s32 = 0.f;
      for (int k = 0; k < K; ++k) {
        s += (double)hA[m * K + k] * hB[n * K + k];
        sa += fabs((double)hA[m * K + k] * hB[n * K + k]);
        s32 += hA[m * K + k] * hB[n * K + k];
      }
      ref[m * T + n] = s;
      ref32[m * T + n] = s32;
      scale += sa;
    }
  scale /= (double)T * T;
  double e32 = 0;
  for (int i = 0; i < T * T; ++i) e32 = fmax(e32, fabs(ref32[i] - ref[i]));
  printf("K = %d, mean sum|a b| = %.3f; plain fp32 dot product: max |err| = %.3g (%.3g of the scale)\n", K, scale, e32, e32 / scale);

  float *dA, *dB, *dD;
  int* dErr;
  long long* dCyc;
  CK(cudaMalloc(&dA, T * K * 4));
  CK(cudaMalloc(&dB, T * K * 4));
  CK(cudaMalloc(&dD, T * T * 4));
  CK(cudaMalloc(&dErr, 4));
  CK(cudaMalloc(&dCyc, 8));
  CK(cudaMemcpy(dA, hA.data(), T * K * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), T * K * 4, cudaMemcpyHostToDevice));
  const int smem_bytes = 8 * PANEL_BYTES + 1024;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  const char* names[] = {"1 x TF32 (hi hi)", "3 x TF32 (lo hi + hi lo + hi hi)", "tf32 MMA stream", "bf16 MMA stream"};
  const int reps = 64;
  for (int mode = 0; mode < 4; ++mode) {
    CK(cudaMemset(dD, 0xFF, T * T * 4));
    CK(cudaMemset(dErr, 0, 4));
    probe_kernel<<<1, 128, smem_bytes>>>(dA, dB, dD, mode, reps, dCyc, dErr);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("%-36s : CUDA ERROR %s\n", names[mode], cudaGetErrorString(e));
      return 3;
    }
    int herr = 0;
    long long cyc = 0;
    CK(cudaMemcpy(&herr, dErr, 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&cyc, dCyc, 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hD.data(), dD, T * T * 4, cudaMemcpyDeviceToHost));
    if (mode < 2) {
      double maxerr = 0;
      for (int i = 0; i < T * T; ++i) {
        const double d = fabs((double)hD[i] - ref[i]);
        if (d > maxerr || d != d) maxerr = d;
      }
      printf("%-36s : max |err| = %.3g (%.3g of the scale = 2^%.1f; %.1f x the plain fp32 dot product) timeout=%d  D[0,0..2]=%g %g %g ref=%g %g %g\n",
             names[mode], maxerr, maxerr / scale, log2(maxerr / scale), maxerr / e32, herr, hD[0], hD[1], hD[2], ref[0], ref[1], ref[2]);
    } else {
      const double per = (double)cyc / (reps * (K / 8));
      printf("%-36s : %lld cycles for %d instructions = %.1f cycles per instruction (128x128x%d) timeout=%d\n", names[mode], cyc,
             reps * (K / 8), per, mode == 2 ? 8 : 16, herr);
    }
  }
  return 0;
}
