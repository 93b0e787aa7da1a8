// mgn_gemm_f32_tc.cu — fp32 nn.Linear products on the tensor cores with fp32 accuracy (3 x TF32 operand split).
//
//     out[M, N] = act( x[M, K] W[N, K]^T + bias[N] )        x, W, bias, out fp32;  K % 32 == 0, N % 128 == 0
//
// = the nn.Linear products of MeshGraphMLP (models/gnn_layers/mesh_graph_mlp.py:142-168, 200-203) for fp32 callers -- the
// reference's default precision (meshgraphnet.py:128-150; its tests switch TF32 off, test_meshgraphnet_snmg.py:31) -- which the
// reference leaves to cuBLAS SGEMM.  `tcgen05.mma kind::tf32` alone keeps 11 bits per operand; here every operand is split as
//     v = hi + lo,   hi = tf32(v),   lo = tf32(v - hi)
// and the three products  x_lo W_hi^T + x_hi W_lo^T + x_hi W_hi^T  are accumulated in TMEM (the dropped x_lo W_lo^T term is
// 2^-22 relative): the large hi x hi sum in one accumulator, the two small cross terms in a second one, added in the epilogue.
// Measured on hardware (tools/probe_tf32.cu, tools/probe_gemm_f32_tc.cu, profiles/r02_probe_tf32.txt,
// r02_probe_gemm_f32_tc.txt): as close to the float64 product as the exact-fp32 SIMT kernels are (worst element 1.8 x / 0.9 x
// theirs at K = 384 / 128; with ONE accumulator 5 x: the tensor core's accumulate loses a little at every one of the 3 K / 8
// steps), a 128x128x8 TF32 instruction issues at the rate of a 128x128x16 bf16 one, and at the c2 shape the forward product
// runs 6.3 x and the data gradient 3.7 x faster than the SIMT kernels.
//
// Pipeline (persistent, one CTA per SM, 128-row tiles of x, one 128-column block of N per launch): warp 0 issues the MMAs,
// warp 1 the TMA loads, warps 2-5 split the x chunk in shared memory (thread = row: raw fp32 in, hi written in place, lo into
// the neighbouring panel), warps 6-9 drain the accumulator (bias, activation, 16-byte global stores).  A stage holds ONE
// 32-column K chunk: x (raw -> hi), x lo, W hi, W lo = four 16 KB panels in the 128-byte-swizzle layout (32 fp32 columns per
// 128-byte row: the descriptor arithmetic of the bf16 kernels carries over, a k-step is 32 bytes).  The weight arrives
// already split: mgn_split_weight_tf32 writes the [2N, K] image (hi rows, then lo rows; optionally of the transposed weight,
// which makes the same kernel compute the data gradient g_x = g_y W).  Two accumulator pairs in TMEM (hi x hi | cross terms):
// the epilogue of tile i runs under the main loop of tile i + 1.
#include "mgn_common.cuh"
#include "mgn_tc.cuh"
#include "mgn_tile.cuh"
#include "mgn_tma.cuh"

namespace mgn {
namespace f32tc {

using namespace tile;
constexpr int kSplitWarps = 4;
constexpr int kEpiWarps = 4;
constexpr int kThreads = 32 * (2 + kSplitWarps + kEpiWarps);
constexpr int kH = 128;
constexpr int kStages = 3;
constexpr int kStageBytes = 4 * kPB;  // x hi | x lo | W hi | W lo
constexpr int kKc = 32;               // fp32 columns per chunk = one 128-byte panel row

struct Params {
  long long M;
  int kc;        // K / 32
  int act;       // MGN_ACT_NONE or MGN_ACT_RELU
  int n0;        // first output column of this launch (row of the hi image)
  int lo_row0;   // row of the lo image that matches n0 (N + n0)
  const float* bias;  // already offset by n0 (nullable)
  float* out;         // already offset by n0
  long long ld_out;
  int* status;
  alignas(64) CUtensorMap m_x, m_w;
};

enum { B_FULL = 0, B_SPLIT = kStages, B_EMPTY = 2 * kStages, B_ACCFULL = 3 * kStages, B_ACCFREE = 3 * kStages + 2, B_NUM = 3 * kStages + 4 };
constexpr int kSmemBytes = kStages * kStageBytes + kH * 4 + 16 * 8 + 16;

// kind::tf32 instruction descriptor: c_format = 1 (f32), a_format = b_format = 2 (tf32), both operands K-major
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

__global__ void __launch_bounds__(kThreads, 1) gemm_f32_tc_kernel(const __grid_constant__ Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if ((smem_u32(smem) & 1023u) != 0) {
    if (tid == 0 && p.status) atomicOr(p.status, 2);
    return;
  }
  uint8_t* sStage = smem;
  float* sBias = reinterpret_cast<float*>(smem + kStages * kStageBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sBias) + kH * 4);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(bars) + 16 * 8);

  for (int i = tid; i < kH; i += kThreads) sBias[i] = p.bias ? p.bias[i] : 0.f;
  if (tid == 0) {
    for (int b = 0; b < B_NUM; ++b) {
      int cnt = 1;
      if (b >= B_SPLIT && b < B_SPLIT + kStages) cnt = kSplitWarps;
      if (b >= B_ACCFREE) cnt = kEpiWarps;
      mbar_init(&bars[b], cnt);
    }
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const long long n_tiles = (p.M + kRows - 1) / kRows;
  const int n_my = static_cast<int>((n_tiles > blockIdx.x) ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0);
  const int kc = p.kc;
  bool timed_out = false;

  if (warp == 0) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      const uint32_t aS = smem_u32(sStage);
      const uint32_t idesc = idesc_tf32(128, 128);
      int sc = 0;
      for (int it = 0; it < n_my && !timed_out; ++it) {
        const int b = it & 1, ua = it >> 1;
        // accumulator b is free once the epilogue has drained its previous tile
        if (it >= 2 && !wait_clk(&bars[B_ACCFREE + b], (ua - 1) & 1)) { timed_out = true; break; }
        // two accumulators per tile: the hi x hi product and the two small cross terms.  The tensor core's fp32 accumulate
        // loses a little at every step; kept apart, the small terms neither suffer from nor add to the rounding of the large
        // sum, which then takes K / 8 steps instead of 3 K / 8 (measured: tools/probe_gemm_f32_tc.cu)
        const uint32_t tAcc = tmem + b * 2 * kH, tSide = tAcc + kH;
        for (int c = 0; c < kc; ++c, ++sc) {
          const int s = sc % kStages;
          if (!wait_clk(&bars[B_SPLIT + s], (sc / kStages) & 1)) { timed_out = true; break; }
          tc_fence_after_sync();
          const uint32_t aXh = aS + s * kStageBytes, aXl = aXh + kPB, aWh = aXh + 2 * kPB, aWl = aXh + 3 * kPB;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_ss_tf32(tSide, umma_desc_kmajor(aXl, kk), umma_desc_kmajor(aWh, kk), idesc, (c | kk) != 0);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) umma_ss_tf32(tSide, umma_desc_kmajor(aXh, kk), umma_desc_kmajor(aWl, kk), idesc, 1);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_ss_tf32(tAcc, umma_desc_kmajor(aXh, kk), umma_desc_kmajor(aWh, kk), idesc, (c | kk) != 0);
          umma_commit(&bars[B_EMPTY + s]);  // the stage is free once these MMAs have read it
        }
        if (timed_out) break;
        umma_commit(&bars[B_ACCFULL + b]);
      }
    }
  } else if (warp == 1) {
    // =========================== loader (TMA) ===========================
    if (lane == 0) {
      int sc = 0;
      for (int it = 0; it < n_my && !timed_out; ++it) {
        const long long row0 = (static_cast<long long>(blockIdx.x) + static_cast<long long>(it) * gridDim.x) * kRows;
        for (int c = 0; c < kc; ++c, ++sc) {
          const int s = sc % kStages;
          if (sc >= kStages && !wait_clk(&bars[B_EMPTY + s], ((sc / kStages) & 1) ^ 1)) { timed_out = true; break; }
          mbar_arrive_expect_tx(&bars[B_FULL + s], static_cast<uint32_t>(3 * kPB));
          const uint32_t dst = smem_u32(sStage) + s * kStageBytes;
          tma_load_2d(dst, &p.m_x, c * kKc, static_cast<int>(row0), &bars[B_FULL + s]);
          tma_load_2d(dst + 2 * kPB, &p.m_w, c * kKc, p.n0, &bars[B_FULL + s]);
          tma_load_2d(dst + 3 * kPB, &p.m_w, c * kKc, p.lo_row0, &bars[B_FULL + s]);
        }
      }
    }
  } else if (warp < 2 + kSplitWarps) {
    // =========================== splitters: x chunk -> hi (in place) and lo ===========================
    const int row = tid - 64;
    int sc = 0;
    for (int it = 0; it < n_my && !timed_out; ++it) {
      for (int c = 0; c < kc; ++c, ++sc) {
        const int s = sc % kStages;
        if (!__all_sync(0xffffffffu, wait_clk(&bars[B_FULL + s], (sc / kStages) & 1))) { timed_out = true; break; }
        uint8_t* pX = sStage + s * kStageBytes;
        uint8_t* pL = pX + kPB;
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          const uint32_t off = sw128_offset(row, ch);
          const float4 v = *reinterpret_cast<const float4*>(pX + off);
          float4 h, l;
          h.x = to_tf32(v.x); h.y = to_tf32(v.y); h.z = to_tf32(v.z); h.w = to_tf32(v.w);
          l.x = to_tf32(v.x - h.x); l.y = to_tf32(v.y - h.y); l.z = to_tf32(v.z - h.z); l.w = to_tf32(v.w - h.w);
          *reinterpret_cast<float4*>(pX + off) = h;
          *reinterpret_cast<float4*>(pL + off) = l;
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[B_SPLIT + s]);
      }
    }
  } else {
    // =========================== epilogue (4 warps: thread = row = TMEM lane) ===========================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    for (int it = 0; it < n_my && !timed_out; ++it) {
      const int b = it & 1, ua = it >> 1;
      const long long grow = (static_cast<long long>(blockIdx.x) + static_cast<long long>(it) * gridDim.x) * kRows + row;
      if (!__all_sync(0xffffffffu, wait_clk(&bars[B_ACCFULL + b], ua & 1))) { timed_out = true; break; }
      tc_fence_after_sync();
      const uint32_t t_acc = tmem + b * 2 * kH + (static_cast<uint32_t>(q * 32) << 16);
      float* orow = p.out + grow * p.ld_out;
#pragma unroll 1
      for (int g = 0; g < 4; ++g) {
        uint32_t v[32], u[32];
        tmem_ld32(t_acc + 32 * g, v);
        tmem_ld32(t_acc + kH + 32 * g, u);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(u[j]));
        if (grow < p.M) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 o;
            o.x = __uint_as_float(v[4 * j]) + sBias[32 * g + 4 * j];
            o.y = __uint_as_float(v[4 * j + 1]) + sBias[32 * g + 4 * j + 1];
            o.z = __uint_as_float(v[4 * j + 2]) + sBias[32 * g + 4 * j + 2];
            o.w = __uint_as_float(v[4 * j + 3]) + sBias[32 * g + 4 * j + 3];
            if (p.act == MGN_ACT_RELU) {
              o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
            }
            *reinterpret_cast<float4*>(orow + 32 * g + 4 * j) = o;
          }
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[B_ACCFREE + b]);
    }
  }
  if (timed_out && p.status != nullptr) atomicOr(p.status, 1);
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// fp32 weight [rows, cols] (row stride ld) -> split image [2 N, K] dense: hi rows 0..N-1, lo rows N..2N-1, where
// (N, K) = (rows, cols), or (cols, rows) for the transposed weight
__global__ void split_weight_kernel(const float* __restrict__ w, long long rows, long long cols, long long ld,
                                    float* __restrict__ out, int transpose) {
  const long long n = rows * cols;
  const long long N = transpose ? cols : rows, K = transpose ? rows : cols;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / K, c = i - r * K;  // output coordinates (row of the image, k)
    const float v = transpose ? __ldg(w + c * ld + r) : __ldg(w + r * ld + c);
    const float h = to_tf32(v);
    out[i] = h;
    out[N * K + i] = to_tf32(v - h);
  }
}

// tensor map over an fp32 table [rows, cols] with a row stride of ld elements: box = {32 columns, box_rows}, 128-byte swizzle
static inline int make_f32_map(CUtensorMap* m, const void* base, long long rows, long long ld, long long cols, int box_rows) {
  TmaEncodeTiledFn enc = tma_encoder();
  if (enc == nullptr) return -100;
  if (rows <= 0) rows = 1;
  const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  const cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * 4};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(kKc), static_cast<cuuint32_t>(box_rows)};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -101;
}

}  // namespace f32tc
}  // namespace mgn

using namespace mgn;

extern "C" int mgn_split_weight_tf32(const float* w, int64_t rows, int64_t cols, int64_t ld, float* out, int transpose,
                                     mgn_stream_t stream) {
  MGN_CHECK_ARG(rows >= 0 && cols >= 0 && ld >= cols);
  if (rows == 0 || cols == 0) return MGN_OK;
  MGN_CHECK_ARG(w && out);
  const long long n = rows * cols;
  const int grid = static_cast<int>((n + 255) / 256 < 4096 ? (n + 255) / 256 : 4096);
  f32tc::split_weight_kernel<<<grid, 256, 0, MGN_ST(as_stream(stream))>>>(w, rows, cols, ld, out, transpose);
  return mgn_launch_status();
}

extern "C" int mgn_linear_f32_tc(const float* x, int64_t ld_x, int64_t M, int64_t K, const float* w_split, int64_t N,
                                 const float* bias, int act, float* out, int64_t ld_out, int* status, mgn_stream_t stream) {
  MGN_CHECK_ARG(M >= 0 && K > 0 && N > 0 && K % f32tc::kKc == 0 && N % f32tc::kH == 0 && ld_x >= K && ld_out >= N);
  MGN_CHECK_ARG(act == MGN_ACT_NONE || act == MGN_ACT_RELU);
  MGN_CHECK_ARG(2 * N < (int64_t(1) << 31) && M < (int64_t(1) << 31));
  if (M == 0) return MGN_OK;
  MGN_CHECK_ARG(x && w_split && out && ld_x % 4 == 0 && ld_out % 4 == 0);
  for (const void* q : {static_cast<const void*>(x), static_cast<const void*>(w_split), static_cast<const void*>(out)})
    MGN_CHECK_ARG((reinterpret_cast<uintptr_t>(q) & 15) == 0);
  const long long n_tiles = (M + tile::kRows - 1) / tile::kRows;
  const int grid = static_cast<int>(n_tiles < num_sms() ? n_tiles : num_sms());
  cudaStream_t st = as_stream(stream);
  static PerDeviceFlag configured_flag;
  bool& configured = configured_flag.get();
  if (!configured) {
    cudaError_t ce = cudaFuncSetAttribute(f32tc::gemm_f32_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, f32tc::kSmemBytes);
    if (ce != cudaSuccess) return static_cast<int>(ce);
    configured = true;
  }
  f32tc::Params p{};
  p.M = M;
  p.kc = static_cast<int>(K / f32tc::kKc);
  p.act = act;
  p.ld_out = ld_out;
  p.status = status;
  int e = f32tc::make_f32_map(&p.m_x, x, M, ld_x, K, 128);
  e |= f32tc::make_f32_map(&p.m_w, w_split, 2 * N, K, K, 128);
  if (e != 0) return MGN_EINVAL;
  for (int64_t n0 = 0; n0 < N; n0 += f32tc::kH) {  // one 128-column block of N per launch (x is re-read per block)
    p.n0 = static_cast<int>(n0);
    p.lo_row0 = static_cast<int>(N + n0);
    p.bias = bias ? bias + n0 : nullptr;
    p.out = out + n0;
    f32tc::gemm_f32_tc_kernel<<<grid, f32tc::kThreads, f32tc::kSmemBytes, MGN_ST(st)>>>(p);
    const int rc = mgn_launch_status();
    if (rc != MGN_OK) return rc;
  }
  return MGN_OK;
}
