// mgn_node_gemm_tc.cu — node-level plain GEMMs of the fused MeshGraphNet path on tcgen05 / TMEM with TMA staging.
//
//     out[M, 128 nb] = x[M, 128 kb] W[128 nb, 128 kb]^T (+ residual[M, 128], nb == 1)        kb * nb <= 3
//
// Two shapes matter: the per-layer projection table P = nfeat [W1e_src ; W1e_dst ; W1n_node]^T  (kb = 1, nb = 3) and
// the node-feature gradient g_nfeat += T Wp (kb = 3, nb = 1) -- the products the reference leaves to cuBLAS through
// autograd for the node-row column blocks of the first Linear (models/gnn_layers/mesh_graph_mlp.py:142-168; the
// lin_src / lin_dst products of MeshGraphEdgeMLPSum, :396-405).  Both are HBM-bound (one pass over x, one over out),
// so the kernel is a plain persistent pipeline: warp 0 issues MMAs, warp 1 issues the TMA row-tile loads into a ring
// of three 32 KB stages (one stage = one 128-column K block of one tile), eight epilogue warps drain a ring of four
// TMEM accumulators through one 32 KB staging tile that leaves by TMA store (the residual tile arrives there by TMA
// first and is updated in place).  Weights are converted fp32 -> bf16 once per CTA from the optimizer's tensors.
#include "mgn_common.cuh"
#include "mgn_tc.cuh"
#include "mgn_tile.cuh"
#include "mgn_tma.cuh"

namespace mgn {
namespace ng {

using namespace tile;
constexpr int kStages = 3;
constexpr int kAccs = 4;
constexpr int kEpiWarps = 8;
constexpr int kThreads = 32 * (2 + kEpiWarps);
constexpr int kH = 128;

struct Params {
  long long M;
  const float* w;
  long long ld_w;
  int kb, nb;
  int has_res;
  int* status;
  alignas(64) CUtensorMap m_x, m_out, m_res;
};

struct Smem {
  static constexpr int kW = 0;                        // [nb][kb] blocks of 2 panels
  static constexpr int kStage = 6 * kPB;              // kStages x 2 panels
  static constexpr int kOut = kStage + kStages * 2 * kPB;
  static constexpr int kBars = kOut + 2 * kPB;
  static constexpr int kTmemSlot = kBars + 32 * 8;
  static constexpr int kTotal = kTmemSlot + 16;
};
enum { B_FULL = 0, B_EMPTY = kStages, B_ACCFULL = 2 * kStages, B_ACCFREE = 2 * kStages + kAccs, B_RES = 2 * kStages + 2 * kAccs,
       B_NUM = 2 * kStages + 2 * kAccs + 1 };

__global__ void __launch_bounds__(kThreads, 1) node_gemm_tc_kernel(const __grid_constant__ Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if ((smem_u32(smem) & 1023u) != 0) {
    if (tid == 0 && p.status) atomicOr(p.status, 2);
    return;
  }
  uint8_t* sW = smem + Smem::kW;
  uint8_t* sStage = smem + Smem::kStage;
  uint8_t* sOut = smem + Smem::kOut;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Smem::kBars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + Smem::kTmemSlot);
  const int kb = p.kb, nb = p.nb;

  for (int n = 0; n < nb; ++n)
    for (int k = 0; k < kb; ++k)
      stage_weight_ld(sW + (n * kb + k) * 2 * kPB, p.w + static_cast<long long>(n) * kH * p.ld_w + k * kH, p.ld_w, kH, kH, 2,
                      tid, kThreads);
  if (tid == 0) {
    for (int b = 0; b < B_NUM; ++b) mbar_init(&bars[b], (b >= B_ACCFREE && b < B_RES) ? kEpiWarps : 1);
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
      const uint32_t aW = smem_u32(sW), aS = smem_u32(sStage);
      const uint32_t idesc = umma_idesc_bf16(128, 128, 0, 0);
      int sc = 0, jb = 0;
      for (int it = 0; it < n_my && !timed_out; ++it) {
        for (int n = 0; n < nb && !timed_out; ++n, ++jb) {
          const int a = jb % kAccs;
          if (jb >= kAccs && !wait_clk(&bars[B_ACCFREE + a], ((jb / kAccs) & 1) ^ 1)) { timed_out = true; break; }
          for (int k = 0; k < kb; ++k) {
            const int s = (sc + k) % kStages;
            if (n == 0 && !wait_clk(&bars[B_FULL + s], ((sc + k) / kStages) & 1)) { timed_out = true; break; }
            tc_fence_after_sync();
            const uint32_t aA = aS + s * 2 * kPB, aB = aW + (n * kb + k) * 2 * kPB;
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              umma_ss(tmem + a * 128, umma_desc_kmajor(aA + (kk >> 2) * kPB, kk & 3), umma_desc_kmajor(aB + (kk >> 2) * kPB, kk & 3),
                      idesc, (k | kk) != 0);
          }
          umma_commit(&bars[B_ACCFULL + a]);
        }
        for (int k = 0; k < kb; ++k) umma_commit(&bars[B_EMPTY + (sc + k) % kStages]);  // stages free once these MMAs are done
        sc += kb;
      }
    }
  } else if (warp == 1) {
    // =========================== loader (TMA) ===========================
    if (lane == 0) {
      int sc = 0;
      for (int it = 0; it < n_my && !timed_out; ++it) {
        const long long row0 = (static_cast<long long>(blockIdx.x) + static_cast<long long>(it) * gridDim.x) * kRows;
        for (int k = 0; k < kb; ++k, ++sc) {
          const int s = sc % kStages;
          if (sc >= kStages && !wait_clk(&bars[B_EMPTY + s], ((sc / kStages) & 1) ^ 1)) { timed_out = true; break; }
          mbar_arrive_expect_tx(&bars[B_FULL + s], 2 * kPB);
          const uint32_t dst = smem_u32(sStage) + s * 2 * kPB;
          tma_load_2d(dst, &p.m_x, k * kH, static_cast<int>(row0), &bars[B_FULL + s]);
          tma_load_2d(dst + kPB, &p.m_x, k * kH + 64, static_cast<int>(row0), &bars[B_FULL + s]);
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
    int jb = 0;
    for (int it = 0; it < n_my && !timed_out; ++it) {
      const long long row0 = (static_cast<long long>(blockIdx.x) + static_cast<long long>(it) * gridDim.x) * kRows;
      for (int n = 0; n < nb; ++n, ++jb) {
        const int a = jb % kAccs;
        if (leader) {
          tma_store_wait_read();  // the previous result tile has left the staging buffer
          if (p.has_res) {
            mbar_arrive_expect_tx(&bars[B_RES], 2 * kPB);
            tma_load_2d(smem_u32(sOut), &p.m_res, 0, static_cast<int>(row0), &bars[B_RES]);
            tma_load_2d(smem_u32(sOut) + kPB, &p.m_res, 64, static_cast<int>(row0), &bars[B_RES]);
          }
        }
        bool ok = true;
        if (p.has_res) ok = wait_clk(&bars[B_RES], jb & 1);
        else asm volatile("bar.sync 1, 256;" ::: "memory");
        ok = ok && wait_clk(&bars[B_ACCFULL + a], (jb / kAccs) & 1);
        if (!__all_sync(0xffffffffu, ok)) { timed_out = true; break; }
        tc_fence_after_sync();
        const uint32_t t_acc = tmem + a * 128 + (static_cast<uint32_t>(q * 32) << 16) + c0;
#pragma unroll 1
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t v[32];
          tmem_ld32(t_acc + 32 * hh, v);
          uint32_t r[16];
          if (p.has_res) row_load32p(sOut, row, c0 + 32 * hh, r);
          tmem_ld_wait();
          uint32_t o[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float y0 = __uint_as_float(v[2 * j]), y1 = __uint_as_float(v[2 * j + 1]);
            if (p.has_res) {
              y0 += bf_lo(r[j]);
              y1 += bf_hi(r[j]);
            }
            o[j] = pack_bf16x2(y0, y1);
          }
          row_store32p(sOut, row, c0 + 32 * hh, o);
        }
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[B_ACCFREE + a]);
        fence_proxy_async_smem();
        asm volatile("bar.sync 2, 256;" ::: "memory");
        if (leader) {
          tma_store_2d(&p.m_out, smem_u32(sOut), n * kH, static_cast<int>(row0));
          tma_store_2d(&p.m_out, smem_u32(sOut) + kPB, n * kH + 64, static_cast<int>(row0));
          tma_store_commit();
        }
      }
    }
    if (leader) tma_store_wait_all();
  }
  if (timed_out && p.status != nullptr) atomicOr(p.status, 1);
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace ng
}  // namespace mgn

using namespace mgn;

extern "C" int mgn_node_gemm_tc(const void* x, int64_t ld_x, int kb, int64_t M, const float* w, int64_t ld_w, int nb,
                                const void* residual, void* out, int64_t ld_out, int* status, mgn_stream_t stream) {
  MGN_CHECK_ARG(M >= 0 && w && kb >= 1 && nb >= 1 && kb * nb <= 3 && ld_w >= 128 * kb && ld_x >= 128 * kb &&
                ld_out >= 128 * nb);
  MGN_CHECK_ARG(residual == nullptr || nb == 1);
  if (M == 0) return MGN_OK;
  MGN_CHECK_ARG(x && out && ld_x % 8 == 0 && ld_out % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
                (reinterpret_cast<uintptr_t>(out) & 15) == 0);
  if (residual) MGN_CHECK_ARG((reinterpret_cast<uintptr_t>(residual) & 15) == 0);
  ng::Params p{};
  p.M = M;
  p.w = w;
  p.ld_w = ld_w;
  p.kb = kb;
  p.nb = nb;
  p.has_res = residual != nullptr;
  p.status = status;
  int e = tma_make_rows_map(&p.m_x, x, M, ld_x, 128, 128 * kb);
  e |= tma_make_rows_map(&p.m_out, out, M, ld_out, 128, 128 * nb);
  if (residual) e |= tma_make_rows_map(&p.m_res, residual, M, 128, 128, 128);
  if (e != 0) return MGN_EINVAL;
  static PerDeviceFlag configured_flag;
  bool& configured = configured_flag.get();
  if (!configured) {
    cudaError_t ce = cudaFuncSetAttribute(ng::node_gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ng::Smem::kTotal);
    if (ce != cudaSuccess) return static_cast<int>(ce);
    configured = true;
  }
  const long long n_tiles = (M + tile::kRows - 1) / tile::kRows;
  const int grid = static_cast<int>(n_tiles < num_sms() ? n_tiles : num_sms());
  ng::node_gemm_tc_kernel<<<grid, ng::kThreads, ng::Smem::kTotal, MGN_ST(as_stream(stream))>>>(p);
  return mgn_launch_status();
}
