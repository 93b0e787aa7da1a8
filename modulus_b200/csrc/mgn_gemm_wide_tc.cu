// mgn_gemm_wide_tc.cu — K-looped tcgen05 GEMM for the MLP widths the fused hidden-128 kernels do not cover
// (GraphCast: hidden 512, AeroGraphNet: 256-wide encoders / decoder; any activation, the concatenated first layers with
// K = 768 / 1536):
//
//     out[M, N] = act( x[M, K] Wb[N, K]^T + bias[N] )        x, Wb, out bf16, fp32 accumulate in TMEM
//                 K % 64 == 0, N % 128 == 0; one launch covers up to four 128-column blocks of N
//
// = the nn.Linear products of MeshGraphMLP (models/gnn_layers/mesh_graph_mlp.py:142-168, 200-203) that the reference
// leaves to cuBLAS; here they are what `ops.MLPFn` calls for bf16 activations whose widths are multiples of 128 (forward
// and the data gradient g_x = g_y W, which is the same product with the transposed weight image).
//
// Pipeline (persistent, one CTA per SM): warp 0 issues the MMAs, warp 1 the TMA loads, eight epilogue warps drain the
// accumulators.  A stage holds ONE 64-column K chunk: the x panel of the row tile plus the matching K chunk of every
// 128-row block of Wb (1 + nb panels of 16 KB).  All nb accumulators of a row tile live in TMEM at once (nb x 128
// columns), so x is read exactly once per launch; Wb (<= 1.5 MB) is re-streamed per row tile from L2.  The weight image
// is bf16 in global memory (mgn_cast_weight_bf16 converts the optimizer's fp32 tensor, optionally transposed, once per
// call): operands enter shared memory only through TMA in the 128-byte-swizzle panel layout the UMMA descriptors read.
#include "mgn_common.cuh"
#include "mgn_tc.cuh"
#include "mgn_tile.cuh"
#include "mgn_tma.cuh"

namespace mgn {
namespace wide {

using namespace tile;
constexpr int kEpiWarps = 8;
constexpr int kThreads = 32 * (2 + kEpiWarps);
constexpr int kH = 128;
constexpr int kMaxStages = 4;
constexpr int kMaxNb = 4;

struct Params {
  long long M;
  int kc;        // K / 64
  int nb;        // N / 128 of this launch (<= 4)
  int n_stages;  // ring depth (host: what fits next to the output staging tile)
  int act;       // MGN_ACT_NONE or MGN_ACT_RELU (others run as a separate elementwise pass over the stored pre-activation)
  const float* bias;
  int* status;
  alignas(64) CUtensorMap m_x, m_w, m_out;
};

// shared memory: [n_stages][(1 + nb) panels] | output staging tile (2 panels) | bias [nb * 128] fp32 | barriers | tmem slot
__host__ __device__ inline int stage_bytes(int nb) { return (1 + nb) * kPB; }
__host__ __device__ inline int smem_bytes(int nb, int n_stages) {
  return n_stages * stage_bytes(nb) + 2 * kPB + kMaxNb * kH * 4 + 32 * 8 + 16;
}
enum { B_FULL = 0, B_EMPTY = kMaxStages, B_ACCFULL = 2 * kMaxStages, B_ACCFREE = 2 * kMaxStages + 1, B_NUM = 2 * kMaxStages + 2 };

__global__ void __launch_bounds__(kThreads, 1) gemm_wide_tc_kernel(const __grid_constant__ Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if ((smem_u32(smem) & 1023u) != 0) {
    if (tid == 0 && p.status) atomicOr(p.status, 2);
    return;
  }
  const int nb = p.nb, S = p.n_stages, kc = p.kc;
  const int sbytes = stage_bytes(nb);
  uint8_t* sStage = smem;
  uint8_t* sOut = smem + S * sbytes;
  float* sBias = reinterpret_cast<float*>(sOut + 2 * kPB);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sBias) + kMaxNb * kH * 4);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(bars) + 32 * 8);

  for (int i = tid; i < nb * kH; i += kThreads) sBias[i] = p.bias ? p.bias[i] : 0.f;
  if (tid == 0) {
    for (int b = 0; b < B_NUM; ++b) mbar_init(&bars[b], b == B_ACCFREE ? kEpiWarps : 1);
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
  bool timed_out = false;

  if (warp == 0) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      const uint32_t aS = smem_u32(sStage);
      const uint32_t idesc = umma_idesc_bf16(128, 128, 0, 0);
      int sc = 0;
      for (int it = 0; it < n_my && !timed_out; ++it) {
        // the single accumulator set is free once the epilogue has drained the previous tile
        if (it > 0 && !wait_clk(&bars[B_ACCFREE], (it - 1) & 1)) { timed_out = true; break; }
        for (int c = 0; c < kc; ++c, ++sc) {
          const int s = sc % S;
          if (!wait_clk(&bars[B_FULL + s], (sc / S) & 1)) { timed_out = true; break; }
          tc_fence_after_sync();
          const uint32_t aX = aS + s * sbytes;
          for (int n = 0; n < nb; ++n) {
            const uint32_t aW = aX + (1 + n) * kPB;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_ss(tmem + n * kH, umma_desc_kmajor(aX, kk), umma_desc_kmajor(aW, kk), idesc, (c | kk) != 0);
          }
          umma_commit(&bars[B_EMPTY + s]);  // the stage is free once these MMAs have read it
        }
        umma_commit(&bars[B_ACCFULL]);
      }
    }
  } else if (warp == 1) {
    // =========================== loader (TMA) ===========================
    if (lane == 0) {
      int sc = 0;
      for (int it = 0; it < n_my && !timed_out; ++it) {
        const long long row0 = (static_cast<long long>(blockIdx.x) + static_cast<long long>(it) * gridDim.x) * kRows;
        for (int c = 0; c < kc; ++c, ++sc) {
          const int s = sc % S;
          if (sc >= S && !wait_clk(&bars[B_EMPTY + s], ((sc / S) & 1) ^ 1)) { timed_out = true; break; }
          mbar_arrive_expect_tx(&bars[B_FULL + s], static_cast<uint32_t>(sbytes));
          const uint32_t dst = smem_u32(sStage) + s * sbytes;
          tma_load_2d(dst, &p.m_x, c * 64, static_cast<int>(row0), &bars[B_FULL + s]);
          for (int n = 0; n < nb; ++n) tma_load_2d(dst + (1 + n) * kPB, &p.m_w, c * 64, n * kH, &bars[B_FULL + s]);
        }
      }
    }
  } else {
    // =========================== epilogue (8 warps) ===========================
    const int q = warp & 3;
    const int ch = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int c0 = ch * 64;
    const bool leader = warp == 2 && lane == 0;
    for (int it = 0; it < n_my && !timed_out; ++it) {
      const long long row0 = (static_cast<long long>(blockIdx.x) + static_cast<long long>(it) * gridDim.x) * kRows;
      if (!__all_sync(0xffffffffu, wait_clk(&bars[B_ACCFULL], it & 1))) { timed_out = true; break; }
      tc_fence_after_sync();
      for (int n = 0; n < nb; ++n) {
        if (leader) tma_store_wait_read();  // the previous result block has left the staging tile
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const uint32_t t_acc = tmem + n * kH + (static_cast<uint32_t>(q * 32) << 16) + c0;
        const float* bias = sBias + n * kH + c0;
#pragma unroll 1
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t v[32];
          tmem_ld32(t_acc + 32 * hh, v);
          tmem_ld_wait();
          uint32_t o[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const uint32_t w = f2_to_bf16x2(f2_add(f2_packu(v[2 * j], v[2 * j + 1]), f2_ld(bias + 32 * hh + 2 * j)));
            o[j] = p.act == MGN_ACT_RELU ? relu_bf16x2(w) : w;
          }
          row_store32p(sOut, row, c0 + 32 * hh, o);
        }
        fence_proxy_async_smem();
        asm volatile("bar.sync 2, 256;" ::: "memory");
        if (leader) {
          tma_store_2d(&p.m_out, smem_u32(sOut), n * kH, static_cast<int>(row0));
          tma_store_2d(&p.m_out, smem_u32(sOut) + kPB, n * kH + 64, static_cast<int>(row0));
          tma_store_commit();
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[B_ACCFREE]);
    }
    if (leader) tma_store_wait_all();
  }
  if (timed_out && p.status != nullptr) atomicOr(p.status, 1);
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// fp32 [rows, cols] (row stride ld) -> bf16 [rows, cols] or, transposed, bf16 [cols, rows] (both dense)
__global__ void cast_weight_kernel(const float* __restrict__ w, long long rows, long long cols, long long ld,
                                   bf16* __restrict__ out, int transpose) {
  const long long n = rows * cols;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    if (!transpose) {
      const long long r = i / cols, c = i - r * cols;
      out[i] = __float2bfloat16_rn(__ldg(w + r * ld + c));
    } else {  // i runs over the OUTPUT [cols, rows]: coalesced stores, strided (L2-resident) loads
      const long long c = i / rows, r = i - c * rows;
      out[i] = __float2bfloat16_rn(__ldg(w + r * ld + c));
    }
  }
}

}  // namespace wide
}  // namespace mgn

using namespace mgn;

extern "C" int mgn_cast_weight_bf16(const float* w, int64_t rows, int64_t cols, int64_t ld, void* out, int transpose,
                                    mgn_stream_t stream) {
  MGN_CHECK_ARG(rows >= 0 && cols >= 0 && ld >= cols);
  if (rows == 0 || cols == 0) return MGN_OK;
  MGN_CHECK_ARG(w && out);
  const long long n = rows * cols;
  const int grid = static_cast<int>((n + 255) / 256 < 4096 ? (n + 255) / 256 : 4096);
  wide::cast_weight_kernel<<<grid, 256, 0, MGN_ST(as_stream(stream))>>>(w, rows, cols, ld, static_cast<bf16*>(out), transpose);
  return mgn_launch_status();
}

extern "C" int mgn_gemm_bf16_tc(const void* x, int64_t ld_x, int64_t M, int64_t K, const void* w_bf16, int64_t ld_w,
                                int64_t N, const float* bias, int act, void* out, int64_t ld_out, int* status,
                                mgn_stream_t stream) {
  MGN_CHECK_ARG(M >= 0 && K > 0 && N > 0 && K % 64 == 0 && N % wide::kH == 0 && ld_x >= K && ld_w >= K && ld_out >= N);
  MGN_CHECK_ARG(act == MGN_ACT_NONE || act == MGN_ACT_RELU);
  if (M == 0) return MGN_OK;
  MGN_CHECK_ARG(x && w_bf16 && out && ld_x % 8 == 0 && ld_w % 8 == 0 && ld_out % 8 == 0);
  for (const void* q : {x, w_bf16, static_cast<const void*>(out)}) MGN_CHECK_ARG((reinterpret_cast<uintptr_t>(q) & 15) == 0);
  const long long n_tiles = (M + tile::kRows - 1) / tile::kRows;
  const int grid = static_cast<int>(n_tiles < num_sms() ? n_tiles : num_sms());
  cudaStream_t st = as_stream(stream);
  static PerDeviceFlag configured_flag;
  bool& configured = configured_flag.get();
  if (!configured) {
    cudaError_t ce = cudaFuncSetAttribute(wide::gemm_wide_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (ce != cudaSuccess) return static_cast<int>(ce);
    configured = true;
  }
  for (int64_t n0 = 0; n0 < N; n0 += wide::kMaxNb * wide::kH) {  // up to four 128-column blocks of N per launch
    const int nb = static_cast<int>((N - n0 < wide::kMaxNb * wide::kH ? N - n0 : wide::kMaxNb * wide::kH) / wide::kH);
    wide::Params p{};
    p.M = M;
    p.kc = static_cast<int>(K / 64);
    p.nb = nb;
    int stages = wide::kMaxStages;
    while (stages > 1 && wide::smem_bytes(nb, stages) > 227 * 1024) --stages;
    p.n_stages = stages;
    p.act = act;
    p.bias = bias ? bias + n0 : nullptr;
    p.status = status;
    const bf16* wb = static_cast<const bf16*>(w_bf16) + n0 * ld_w;
    bf16* ob = static_cast<bf16*>(out) + n0;
    int e = tma_make_rows_map(&p.m_x, x, M, ld_x, 128, static_cast<int>(K));
    e |= tma_make_rows_map(&p.m_w, wb, nb * wide::kH, ld_w, 128, static_cast<int>(K));
    e |= tma_make_rows_map(&p.m_out, ob, M, ld_out, 128, nb * wide::kH);
    if (e != 0) return MGN_EINVAL;
    wide::gemm_wide_tc_kernel<<<grid, wide::kThreads, wide::smem_bytes(nb, stages), MGN_ST(st)>>>(p);
    const int rc = mgn_launch_status();
    if (rc != MGN_OK) return rc;
  }
  return MGN_OK;
}
