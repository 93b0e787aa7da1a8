// mgn_wgrad_tc.cu — node-level weight-gradient GEMM on tcgen05:  out[128*JB, 128] = G[M, 128*JB]^T X[M, 128]
//
// Used for the weight gradients of the first-Linear column blocks that act on per-node rows
// (fused.py: g_Wp = T^T nfeat with T = [csr_sum(g_z1) | csc_sum(g_z1) | g_z1_node]); the reference gets
// the same numbers from autograd's addmm backward over the materialised [E, 3H] concat
// (physicsnemo/models/gnn_layers/mesh_graph_mlp.py:267-275).
//
// One persistent CTA per SM; the reduction dimension is the row index, so both operands are read
// MN-major straight from the row-major tiles staged by cp.async (no transpose pass), and the JB
// accumulators stay in TMEM for the whole kernel.  Per-CTA fp32 partials + fixed-order second stage.
#include "mgn_common.cuh"
#include "mgn_tc.cuh"
#include "mgn_reduce.cuh"

namespace mgn {
namespace wg {

constexpr int kPB = 16384;
constexpr int kThreads = 160;  // warp 0: MMA issuer / TMEM owner, warps 1-4: movers (+ final drain)
constexpr int kStages = 3;
constexpr int kH = 128;

struct Params {
  const bf16* g;
  long long ld_g;
  int jb;
  const bf16* x;
  long long ld_x;
  long long M;
  float* partials;  // [grid][jb * 128 * 128]
  int* status;
};

struct Smem {
  static constexpr int kRing = 0;                         // kStages x (G: 2 panels, X: 2 panels)
  static constexpr int kBars = kStages * 4 * kPB;
  static constexpr int kTmemSlot = kBars + 8 * 8;
  static constexpr int kTotal = kTmemSlot + 16;
};

__device__ __forceinline__ bool wait_clk(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return true;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 400000000LL) return false;
  }
  return true;
}

__global__ void __launch_bounds__(kThreads, 1) wgrad_tc_kernel(const Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if ((smem_u32(smem) & 1023u) != 0) {
    if (tid == 0 && p.status) atomicOr(p.status, 2);
    return;
  }
  uint8_t* ring = smem + Smem::kRing;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + Smem::kBars);
  uint64_t* empty = full + kStages;
  uint64_t* done = empty + kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + Smem::kTmemSlot);
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], 4);
      mbar_init(&empty[s], 1);
    }
    mbar_init(done, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;

  const long long n_tiles = (p.M + 127) / 128;
  const int n_my = static_cast<int>((n_tiles > blockIdx.x) ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0);
  const int n_items = n_my * p.jb;
  bool timed_out = false;

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t id_tn = umma_idesc_bf16(128, 128, 1, 1);
      const uint32_t ring_addr = smem_u32(ring);
      for (int i = 0; i < n_items; ++i) {
        const int slot = i % kStages;
        const int it = i / p.jb, j = i - it * p.jb;
        if (!wait_clk(&full[slot], (i / kStages) & 1)) {
          timed_out = true;
          break;
        }
        tc_fence_after_sync();
        const uint32_t ag = ring_addr + slot * 4 * kPB, ax = ag + 2 * kPB;
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_ss(tmem + j * 128, umma_desc_mnmajor(ag, k, kPB), umma_desc_mnmajor(ax, k, kPB), id_tn, (it | k) != 0);
        umma_commit(&empty[slot]);
      }
      umma_commit(done);
    }
  } else {
    const int mt = tid - 32;
    const int chunk = mt & 15, rsub = mt >> 4;
    const uint32_t ring_addr = smem_u32(ring);
    int prev_slot = -1;
    for (int i = 0; i < n_items; ++i) {
      const int slot = i % kStages;
      const int it = i / p.jb, j = i - it * p.jb;
      const long long row0 = (static_cast<long long>(blockIdx.x) + static_cast<long long>(it) * gridDim.x) * 128;
      const bool ok = __all_sync(0xffffffffu, wait_clk(&empty[slot], ((i / kStages) & 1) ^ 1));
      if (!ok) {
        timed_out = true;
        break;
      }
      const uint32_t sg = ring_addr + slot * 4 * kPB, sx = sg + 2 * kPB;
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        const int row = r * 8 + rsub;
        const long long grow = row0 + row;
        const bool v = grow < p.M;
        const long long gr = v ? grow : 0;
        const uint32_t off = (chunk >> 3) * kPB + sw128_offset(row, chunk & 7);
        cp_async16_zfill(sg + off, p.g + gr * p.ld_g + j * kH + chunk * 8, v);
        cp_async16_zfill(sx + off, p.x + gr * p.ld_x + chunk * 8, v);
      }
      cp_async_commit();
      if (prev_slot >= 0) {
        cp_async_wait<1>();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[prev_slot]);
      }
      prev_slot = slot;
    }
    if (prev_slot >= 0 && !timed_out) {
      cp_async_wait<0>();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[prev_slot]);
    }
    cp_async_wait<0>();
    // ---- drain the accumulators once every MMA has completed
    const bool ok = __all_sync(0xffffffffu, wait_clk(done, 0));
    if (!ok) timed_out = true;
    tc_fence_after_sync();
    if (ok && n_my > 0) {
      const int q = warp & 3;
      const int row = q * 32 + lane;
      float* part = p.partials + static_cast<long long>(blockIdx.x) * p.jb * kH * kH;
      for (int j = 0; j < p.jb; ++j) {
        const uint32_t t = tmem + j * 128 + (static_cast<uint32_t>(q * 32) << 16);
        float* dst = part + (j * kH + row) * kH;
#pragma unroll 1
        for (int g = 0; g < 4; ++g) {
          uint32_t v[32];
          tmem_ld32(t + g * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int u = 0; u < 8; ++u)
            reinterpret_cast<float4*>(dst + g * 32)[u] =
                make_float4(__uint_as_float(v[4 * u]), __uint_as_float(v[4 * u + 1]), __uint_as_float(v[4 * u + 2]),
                            __uint_as_float(v[4 * u + 3]));
        }
      }
    }
  }
  if (timed_out && p.status) atomicOr(p.status, 1);
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace wg
}  // namespace mgn

using namespace mgn;

static int wg_grid(int64_t M) {
  const long long n_tiles = (M + 127) / 128;
  return static_cast<int>(n_tiles < num_sms() ? n_tiles : num_sms());
}

extern "C" size_t mgn_wgrad_tc_workspace_bytes(int64_t M, int n_blocks) {
  if (M <= 0) return 0;
  return static_cast<size_t>(wg_grid(M)) * n_blocks * 128 * 128 * sizeof(float);
}

extern "C" int mgn_wgrad_tc(const void* g, int64_t ld_g, int n_blocks, const void* x, int64_t ld_x, int64_t M,
                            float* out, int64_t ld_out, void* workspace, size_t workspace_bytes, int* status,
                            mgn_stream_t stream) {
  MGN_CHECK_ARG(M > 0 && g && x && out && n_blocks >= 1 && n_blocks <= 3 && ld_g >= 128 * n_blocks && ld_x >= 128 &&
                ld_out >= 128);
  MGN_CHECK_ARG(ld_g % 8 == 0 && ld_x % 8 == 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0 &&
                (reinterpret_cast<uintptr_t>(x) & 15) == 0);
  if (workspace == nullptr || workspace_bytes < mgn_wgrad_tc_workspace_bytes(M, n_blocks)) return MGN_EWORKSPACE;
  static PerDeviceFlag configured_flag;
  bool& configured = configured_flag.get();
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(wg::wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, wg::Smem::kTotal);
    if (e != cudaSuccess) return static_cast<int>(e);
    configured = true;
  }
  wg::Params p{};
  p.g = static_cast<const bf16*>(g);
  p.ld_g = ld_g;
  p.jb = n_blocks;
  p.x = static_cast<const bf16*>(x);
  p.ld_x = ld_x;
  p.M = M;
  p.partials = static_cast<float*>(workspace);
  p.status = status;
  cudaStream_t st = as_stream(stream);
  const int grid = wg_grid(M);
  wg::wgrad_tc_kernel<<<grid, wg::kThreads, wg::Smem::kTotal, MGN_ST(st)>>>(p);
  int rc = mgn_launch_status();
  if (rc != MGN_OK) return rc;
  ReduceParams rp{};
  rp.partials = p.partials;
  rp.stride = static_cast<long long>(n_blocks) * 128 * 128;
  rp.n_parts = grid;
  rp.n_seg = 1;
  rp.seg[0] = ReduceSeg{out, ld_out, 128 * n_blocks, 128, 0, 128};
  reduce_cta_partials_kernel<<<dim3(64, 1), 256, 0, MGN_ST(st)>>>(rp);
  return mgn_launch_status();
}
