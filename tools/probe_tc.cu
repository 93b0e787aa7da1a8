// probe_tc.cu — hardware check of the UMMA operand layouts used by the fused kernels.
// One CTA, 128x128x128 bf16 GEMM in four operand modes:
//   0: D = A  * B^T  (A K-major smem, B K-major smem)       forward  X * W^T
//   1: D = A^T* B    (A MN-major smem, B MN-major smem)     wgrad    dY^T * X
//   2: D = A  * B    (A K-major smem, B MN-major smem)      dgrad    dY * W
//   3: D = A  * B^T  (A in TMEM as packed bf16, B K-major)  chained  H * W^T
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o probe_tc probe_tc.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include <cuda_runtime.h>
#include "../modulus_b200/csrc/mgn_tc.cuh"

using namespace mgn;

#define CK(x)                                                                    \
  do {                                                                           \
    cudaError_t e_ = (x);                                                        \
    if (e_ != cudaSuccess) {                                                     \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                   \
    }                                                                            \
  } while (0)

constexpr int T = 128;
constexpr int PANEL_BYTES = T * 128;  // 128 rows x 128 B

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ B,
             float* __restrict__ D, int mode, uint32_t mn_lbo, uint32_t mn_sbo, int* err) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                    // 2 panels
  uint8_t* sB = smem + 2 * PANEL_BYTES;  // 2 panels
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5;

  // stage A, B: thread t copies row t (16 chunks of 16 B)
  for (int c = 0; c < 16; ++c) {
    const int panel = c >> 3, ch = c & 7;
    *reinterpret_cast<uint4*>(sA + panel * PANEL_BYTES + sw128_offset(tid, ch)) =
        *reinterpret_cast<const uint4*>(A + tid * T + c * 8);
    *reinterpret_cast<uint4*>(sB + panel * PANEL_BYTES + sw128_offset(tid, ch)) =
        *reinterpret_cast<const uint4*>(B + tid * T + c * 8);
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 256);
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  const uint32_t tmem_a = tmem + 128;  // columns 128.. hold packed bf16 A (mode 3)

  if (mode == 3) {
    // thread = row; pack own row of A into 64 columns
    const uint32_t lane_addr = tmem_a + (static_cast<uint32_t>(warp * 32) << 16);
    for (int g = 0; g < 4; ++g) {
      uint32_t v[16];
      for (int j = 0; j < 16; ++j)
        v[j] = *reinterpret_cast<const uint32_t*>(A + tid * T + g * 32 + j * 2);
      tmem_st16(lane_addr + g * 16, v);
    }
    tmem_st_wait();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
  }

  if (tid == 0) {
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
    if (mode == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, 128, 0, 0);
      for (int j = 0; j < 8; ++j)
        umma_ss(tmem, umma_desc_kmajor(a0 + (j >> 2) * PANEL_BYTES, j & 3),
                umma_desc_kmajor(b0 + (j >> 2) * PANEL_BYTES, j & 3), idesc, j > 0);
    } else if (mode == 1) {
      const uint32_t idesc = umma_idesc_bf16(128, 128, 1, 1);
      for (int j = 0; j < 8; ++j)
        umma_ss(tmem, umma_smem_desc(a0 + j * 2048, mn_lbo, mn_sbo),
                umma_smem_desc(b0 + j * 2048, mn_lbo, mn_sbo), idesc, j > 0);
    } else if (mode == 2) {
      const uint32_t idesc = umma_idesc_bf16(128, 128, 0, 1);
      for (int j = 0; j < 8; ++j)
        umma_ss(tmem, umma_desc_kmajor(a0 + (j >> 2) * PANEL_BYTES, j & 3),
                umma_smem_desc(b0 + j * 2048, mn_lbo, mn_sbo), idesc, j > 0);
    } else {
      const uint32_t idesc = umma_idesc_bf16(128, 128, 0, 0);
      for (int j = 0; j < 8; ++j)
        umma_ts(tmem, tmem_a + j * 8, umma_desc_kmajor(b0 + (j >> 2) * PANEL_BYTES, j & 3),
                idesc, j > 0);
    }
    umma_commit(&bar);
  }
  __syncwarp();
  if (!mbar_wait(&bar, 0)) {
    if (tid == 0) *err = 1;
  }
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
  if (warp == 0) tmem_dealloc(tmem, 256);
}

static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

int main() {
  std::vector<__nv_bfloat16> hA(T * T), hB(T * T);
  std::vector<float> fA(T * T), fB(T * T);
  srand(1234);
  for (int i = 0; i < T * T; ++i) {
    fA[i] = bf(((rand() % 17) - 8) / 8.0f);
    fB[i] = bf(((rand() % 17) - 8) / 8.0f);
    hA[i] = __float2bfloat16(fA[i]);
    hB[i] = __float2bfloat16(fB[i]);
  }
  __nv_bfloat16 *dA, *dB;
  float* dD;
  int* dErr;
  CK(cudaMalloc(&dA, T * T * 2));
  CK(cudaMalloc(&dB, T * T * 2));
  CK(cudaMalloc(&dD, T * T * 4));
  CK(cudaMalloc(&dErr, 4));
  CK(cudaMemcpy(dA, hA.data(), T * T * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), T * T * 2, cudaMemcpyHostToDevice));
  const int smem_bytes = 4 * PANEL_BYTES + 1024;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));

  struct Case { int mode; uint32_t lbo, sbo; const char* name; };
  Case cases[] = {
      {0, 0, 0, "mode0 NT  K-major/K-major"},
      {1, 16384, 1024, "mode1 TN  MN/MN  lbo=panel sbo=1024"},
      {1, 1024, 16384, "mode1 TN  MN/MN  lbo=1024 sbo=panel (swapped)"},
      {2, 16384, 1024, "mode2 NN  K/MN   lbo=panel sbo=1024"},
      {2, 1024, 16384, "mode2 NN  K/MN   lbo=1024 sbo=panel (swapped)"},
      {3, 0, 0, "mode3 TS  A in TMEM (packed bf16), B K-major"},
  };
  std::vector<float> hD(T * T), ref(T * T);
  int n_ok = 0;
  for (const Case& c : cases) {
    for (int m = 0; m < T; ++m)
      for (int n = 0; n < T; ++n) {
        float s = 0.f;
        for (int k = 0; k < T; ++k) {
          float a = (c.mode == 1) ? fA[k * T + m] : fA[m * T + k];
          float b = (c.mode == 0 || c.mode == 3) ? fB[n * T + k] : fB[k * T + n];
          s += a * b;
        }
        ref[m * T + n] = s;
      }
    CK(cudaMemset(dD, 0xFF, T * T * 4));
    CK(cudaMemset(dErr, 0, 4));
    probe_kernel<<<1, 128, smem_bytes>>>(dA, dB, dD, c.mode, c.lbo, c.sbo, dErr);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("%-55s : CUDA ERROR %s\n", c.name, cudaGetErrorString(e));
      return 3;  // context is dead
    }
    int herr = 0;
    CK(cudaMemcpy(&herr, dErr, 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hD.data(), dD, T * T * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0;
    int bad = 0;
    for (int i = 0; i < T * T; ++i) {
      double d = fabs((double)hD[i] - ref[i]);
      if (!(d <= 1e-3)) ++bad;
      if (d > maxerr || d != d) maxerr = d;
    }
    printf("%-55s : %s  maxerr=%g bad=%d timeout=%d  D[0,0..3]=%g %g %g %g ref=%g %g %g %g\n", c.name,
           (bad == 0 && !herr) ? "PASS" : "FAIL", maxerr, bad, herr, hD[0], hD[1], hD[2], hD[3],
           ref[0], ref[1], ref[2], ref[3]);
    if (bad == 0 && !herr) ++n_ok;
  }
  printf("probe_tc: %d/%d cases pass\n", n_ok, (int)(sizeof(cases) / sizeof(cases[0])));
  return 0;
}
