// mgn_edge_bwd2_tc.cu — MeshEdgeBlock backward with the first hidden activation kept from the forward pass.
//
// The generic fused backward (mgn_mlp_bwd_tc.cu) recomputes the whole forward chain per tile: for the edge block that
// is the gather of both projection rows, GEMM1 and the widest epilogue pass, and the gathers are what the tile
// boundary waits for.  The third-generation forward (mgn_edge_fwd3_tc.cu) leaves h1 = relu(z1) in global memory
// (+ E*H*2 bytes each way, far below what the HBM has to spare at this kernel's pace), so here a tile starts from a
// TMA load of h1:
//
//   recompute       h2 = relu(h1 W2^T + b2) ; y = h2 W3^T + b3
//   LayerNorm bwd   g_out = go1 (+ go2 rows) ; g_y = rstd (ghat - mean(ghat) - xhat mean(ghat xhat)), ghat = g_out gamma
//   layer 3         gW3 += g_y^T h2 ; g_z2 = (g_y W3) * (h2 > 0)
//   layer 2         gW2 += g_z2^T h1 ; g_z1 = (g_z2 W2) * (h1 > 0)
//   layer 1         gW1a += g_z1^T efeat ; g_efeat = g_z1 W1a + g_out        (g_z1 also leaves: its CSC / CSR sums are
//                                                                              the gradient of the projection rows)
//
// = the backward of physicsnemo/models/gnn_layers/mesh_edge_block.py:88-96 (autograd over cuBLAS / ATen there).
// No gathered operand except the go2 rows, no b1, no source / destination tables.  Four 32 KB tile buffers: buffer 0
// = go2 rows -> g_y -> efeat tile; buffer 1 = X (go1 -> g_out -> g_efeat); H1 and H2 swap the other two from tile to
// tile, so that the next tile's h1 streams in right after the layer-2 MMAs; its incoming-gradient rows follow the
// layer-1 MMAs (go2) and the g_efeat store (go1).
// Roles as in the generic kernel: warp 0 MMA issuer, warps 1-4 reducers (go2 gather, column sums), warps 5-12
// epilogue, warp 13 loader (TMA).
// Epilogue passes read the accumulator as two 32-column halves.  -DMGN_BWD2_PIPE16 selects four software-pipelined 16-column chunks
// instead (tmem_pass64): fewer cycles per tile but not faster in wall time on a power-capped B200 (profiles/r02_ab_epilogue.md).
#ifndef MGN_BWD2_PIPE16
#define MGN_NO_PIPE16
#endif
#include "mgn_common.cuh"
#include "mgn_tc.cuh"
#include "mgn_reduce.cuh"
#include "mgn_tile.cuh"
#include "mgn_tma.cuh"
#include "mgn_agg.cuh"

#ifndef MGN_BWD2_AGG_AHEAD
#define MGN_BWD2_AGG_AHEAD 1
#endif
#ifndef MGN_BWD2_AGG_SIDE
#define MGN_BWD2_AGG_SIDE 4
#endif
namespace mgn {
namespace bwd2 {

using namespace tile;
constexpr int kEpiWarps = 8;
constexpr int kLoaderWarp = 5 + kEpiWarps;
// -DMGN_BWD2_AGG_HELPERS=2: warps 14-15 take a third of the fused destination sums (12 segment lanes instead of 8: two
// rounds per mesh tile instead of three).  Measured SLOWER (3.03 -> 3.26 ms per launch at c3, profiles/r02_fwd3_movers.md):
// the kernel is short of shared-memory bandwidth, not of threads.  Off by default.
#ifndef MGN_BWD2_AGG_HELPERS
#define MGN_BWD2_AGG_HELPERS 0
#endif
constexpr int kAggHelpers = MGN_BWD2_AGG_HELPERS;
constexpr int kAggLanes = 8 + 2 * kAggHelpers;
constexpr int kThreads = 32 * (kLoaderWarp + 1 + kAggHelpers);
constexpr int kH = 128;

struct Params {
  const bf16* h1;       // [M,128]
  RowSrc go1, go2;      // incoming gradient rows, summed (go2 optional, gathered)
  long long M;
  const float *w1, *w2, *b2, *w3, *b3, *gamma;
  long long ld_w1;
  float eps;
  // fused destination sums of the g_z1 rows (mgn_agg.cuh; kAgg instances): CSC offsets, destination of every row
  // (ascending), output table [n_dst,128] with row stride ld_agg, boundary records
  const int32_t* seg_off;
  const int32_t* seg_id;
  bf16* agg;
  long long ld_agg;
  float* agg_part;
  int32_t* agg_part_v;
  float* partials;      // [gridDim.x][Part::kTotal]
  long long part_floats;
  int* status;
  long long* timing;
  alignas(64) CUtensorMap m_a, m_h1, m_go1, m_ga, m_gz1;
};

struct Part {
  static constexpr int kW1 = 0;
  static constexpr int kW2 = kH * kH;
  static constexpr int kW3 = kW2 + kH * kH;
  static constexpr int kB1 = kW3 + kH * kH;
  static constexpr int kB2 = kB1 + kH;
  static constexpr int kB3 = kB2 + kH;
  static constexpr int kGamma = kB3 + kH;
  static constexpr int kBeta = kGamma + kH;
  static constexpr int kTotal = kBeta + kH;
};

struct Smem {
  static constexpr int kW1 = 0;
  static constexpr int kW2 = 2 * kPB;
  static constexpr int kW3 = 4 * kPB;
  static constexpr int kBuf = 6 * kPB;        // 4 tile buffers x 2 panels
  static constexpr int kPar = kBuf + 8 * kPB;  // b2, b3, gamma
  static constexpr int kBars = kPar + 4 * kH * 4;
  static constexpr int kTmemSlot = kBars + 24 * 8;
  static constexpr int kTiming = kTmemSlot + 16;
  static constexpr int kTotal = kTiming + 3 * 16 * 8;
};

// B_H1L: h1 tile landed (loader, tx).  B_GO: incoming-gradient rows staged (4 reducer warps + loader).  B_A2: efeat
// tile staged for the layer-1 weight gradient (loader, tx).  B_MMA + k: GEMM2, GEMM3, dgrad3, dgrad2, dgrad1, wgrad1
// complete; B_W3 / B_W2: weight-gradient MMAs of layer 3 / 2 complete.  B_E + k: E2, E3, E4, E5, E6 complete.
// B_CS + k: column sums after E3 / E4 / E5 done.  B_ST / B_XF: the g_z1 / g_A result tile has left its buffer.
// B_DR: the last pass has READ the accumulator (its arithmetic and stores still run while the next tile's GEMM2 starts).
enum { B_H1L = 0, B_GO = 1, B_A2 = 2, B_MMA = 3, B_W3 = 9, B_W2 = 10, B_E = 11, B_CS = 16, B_ST = 19, B_XF = 20, B_DR = 21, B_NUM = 22 };
enum { R_A = -1, R_X = 0, R_H1 = 1, R_H2 = 2 };

__device__ __forceinline__ bool bf_pos_lo(uint32_t w) { return static_cast<int32_t>(w << 16) > 0; }
__device__ __forceinline__ bool bf_pos_hi(uint32_t w) { return static_cast<int32_t>(w & 0xFFFF0000u) > 0; }

__device__ __forceinline__ void tma_tile(uint8_t* buf, const CUtensorMap* map, long long row0, uint64_t* bar) {
  const uint32_t dst = smem_u32(buf);
  tma_load_2d(dst, map, 0, static_cast<int>(row0), bar);
  tma_load_2d(dst + kPB, map, 64, static_cast<int>(row0), bar);
}

// kAddGout: the block's residual runs over the layer-1 input rows (edge block: g_a = g_z1 W1a + g_out); false for the
// node block.  Compile-time: a run-time flag in the last epilogue pass costs the edge instance 9 % through spills.
// kAgg: the reducer warps also sum the g_z1 tile by destination segment (they have the slack for it in this kernel:
// no projection-row gathers).
// Register budget: __launch_bounds__(448, 1) makes ptxas budget for 512 threads (128 registers); 14 warps x 144 x 32 also
// fit the register file, -DMGN_MAXNREG=144 asks for that instead (A/B switch).
#ifdef MGN_MAXNREG
#define MGN_BWD2_BOUNDS __maxnreg__(MGN_MAXNREG)
#else
#define MGN_BWD2_BOUNDS __launch_bounds__(kThreads, 1)
#endif
template <bool kAddGout, bool kAgg>
__global__ void MGN_BWD2_BOUNDS edge_bwd2_kernel(const __grid_constant__ Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if ((smem_u32(smem) & 1023u) != 0) {
    if (tid == 0 && p.status) atomicOr(p.status, 2);
    return;
  }
  uint8_t* sW1 = smem + Smem::kW1;
  uint8_t* sW2 = smem + Smem::kW2;
  uint8_t* sW3 = smem + Smem::kW3;
  uint8_t* buf0 = smem + Smem::kBuf;
  float* sPar = reinterpret_cast<float*>(smem + Smem::kPar);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Smem::kBars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + Smem::kTmemSlot);
// buffer 0 = A, buffer 1 = X; H1 / H2 swap buffers 2 and 3 from tile to tile (the next tile's h1 streams into this
// tile's H2 buffer)
#define MGN_BUF(role, it) (buf0 + ((role) < 0 ? 0 : ((role) == R_X ? 1 : ((role) == R_H1 ? 2 + ((it) & 1) : 3 - ((it) & 1)))) * (2 * kPB))
  const bool has_go2 = p.go2.tab != nullptr;
  constexpr bool has_ln = true;
  constexpr bool need_ga = true;
  (void)has_ln;
  (void)need_ga;

  stage_weight_ld(sW1, p.w1, p.ld_w1, kH, kH, 2, tid, kThreads);
  stage_weight_ld(sW2, p.w2, kH, kH, kH, 2, tid, kThreads);
  stage_weight_ld(sW3, p.w3, kH, kH, kH, 2, tid, kThreads);
  for (int i = tid; i < kH; i += kThreads) {
    sPar[kH + i] = p.b2 ? p.b2[i] : 0.f;
    sPar[2 * kH + i] = p.b3 ? p.b3[i] : 0.f;
    sPar[3 * kH + i] = p.gamma[i];
  }
  if (tid == 0) {
    mbar_init(&bars[B_H1L], 1);
    mbar_init(&bars[B_GO], 5);
    mbar_init(&bars[B_A2], 1);
    for (int b = B_MMA; b < B_E; ++b) mbar_init(&bars[b], 1);
    for (int b = B_E; b < B_CS; ++b) mbar_init(&bars[b], kEpiWarps);
    for (int b = B_CS; b < B_ST; ++b) mbar_init(&bars[b], (kAgg && b == B_CS + 2) ? 4 + kAggHelpers : 4);
    mbar_init(&bars[B_ST], 1);
    mbar_init(&bars[B_XF], 1);
    mbar_init(&bars[B_DR], kEpiWarps);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tAcc = tmem, tW1 = tmem + 128, tW2 = tmem + 256, tW3 = tmem + 384;
#ifdef MGN_DEBUG_HOOKS
  long long dbg_c0 = 0, dbg_g0 = 0;  // per-CTA cycles and nanoseconds of the tile loop -> timing[96 + 4 * cta ...]
  if (p.timing != nullptr && tid == 0) {
    dbg_c0 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(dbg_g0));
  }
#endif

  const long long n_tiles = (p.M + kRows - 1) / kRows;
  const int n_my = static_cast<int>((n_tiles > blockIdx.x) ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0);
  bool timed_out = false;
  const bool tm_on = p.timing != nullptr && blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == kLoaderWarp || warp == 5);
  long long* tm = reinterpret_cast<long long*>(smem + Smem::kTiming) + (warp == 0 ? 0 : (warp == kLoaderWarp ? 16 : 32));
  if (tm_on)
    for (int i = 0; i < 16; ++i) tm[i] = 0;
  long long tlast = clock64();
#define MGN_T(i)                      \
  if (tm_on) {                        \
    const long long t_ = clock64();   \
    tm[i] += t_ - tlast;              \
    tlast = t_;                       \
  }
  float* scratch = reinterpret_cast<float*>(MGN_BUF(R_H2, n_my > 0 ? n_my - 1 : 0));

  if (warp == 0) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      const uint32_t aA = smem_u32(buf0);
      const uint32_t aW1 = smem_u32(sW1), aW2 = smem_u32(sW2), aW3 = smem_u32(sW3);
      const uint32_t id_nt = umma_idesc_bf16(128, 128, 0, 0);   // D = A(K-major) * B(K-major)^T
      const uint32_t id_tn = umma_idesc_bf16(128, 128, 1, 1);   // D = A(MN)^T * B(MN)          (wgrad)
      const uint32_t id_nn = umma_idesc_bf16(128, 128, 0, 1);   // D = A(K-major) * B(MN)       (dgrad)
      for (int it = 0; it < n_my; ++it) {
        const uint32_t par = it & 1;
        const uint32_t aH1 = smem_u32(MGN_BUF(R_H1, it));
        const uint32_t aH2 = smem_u32(MGN_BUF(R_H2, it));
#define MGN_W(b, ph)                         \
  if (!wait_clk(&bars[b], ph)) {             \
    timed_out = true;                        \
    break;                                   \
  }
        // ---- GEMM2: acc = h1 W2^T   (h1 streamed in during the previous tile)
        MGN_W(B_H1L, par);
        if (it > 0) MGN_W(B_DR, par ^ 1);  // the previous tile's last epilogue pass has read the accumulator
        MGN_T(0);
        tc_fence_after_sync();
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_ss(tAcc, umma_desc_kmajor(aH1 + (k >> 2) * kPB, k & 3), umma_desc_kmajor(aW2 + (k >> 2) * kPB, k & 3), id_nt, k != 0);
        umma_commit(&bars[B_MMA + 0]);
        MGN_T(1);
        // ---- GEMM3: acc = h2 W3^T
        MGN_W(B_E + 0, par);
        MGN_T(2);
        tc_fence_after_sync();
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_ss(tAcc, umma_desc_kmajor(aH2 + (k >> 2) * kPB, k & 3), umma_desc_kmajor(aW3 + (k >> 2) * kPB, k & 3), id_nt, k != 0);
        umma_commit(&bars[B_MMA + 1]);
        MGN_T(3);
        // ---- layer 3: acc = g_y W3 ; gW3 += g_y^T h2          (g_y in the A buffer)
        MGN_W(B_E + 1, par);
        MGN_T(4);
        tc_fence_after_sync();
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_ss(tAcc, umma_desc_kmajor(aA + (k >> 2) * kPB, k & 3), umma_desc_mnmajor(aW3, k, kPB), id_nn, k != 0);
        umma_commit(&bars[B_MMA + 2]);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          umma_ss(tW3, umma_desc_mnmajor(aA, j, kPB), umma_desc_mnmajor(aH2, j, kPB), id_tn, (it | j) != 0);
        umma_commit(&bars[B_W3]);
        MGN_T(5);
        // ---- layer 2: acc = g_z2 W2 ; gW2 += g_z2^T h1        (g_z2 in the H2 buffer)
        MGN_W(B_E + 2, par);
        MGN_T(6);
        tc_fence_after_sync();
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_ss(tAcc, umma_desc_kmajor(aH2 + (k >> 2) * kPB, k & 3), umma_desc_mnmajor(aW2, k, kPB), id_nn, k != 0);
        umma_commit(&bars[B_MMA + 3]);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          umma_ss(tW2, umma_desc_mnmajor(aH2, j, kPB), umma_desc_mnmajor(aH1, j, kPB), id_tn, (it | j) != 0);
        umma_commit(&bars[B_W2]);
        MGN_T(7);
        // ---- layer 1: acc = g_z1 W1a (epilogue may start on it at once) ; gW1a += g_z1^T efeat
        MGN_W(B_E + 3, par);
        MGN_T(8);
        MGN_W(B_A2, par);  // (also orders every reducer's column sums of X before E6 rewrites X)
        tc_fence_after_sync();
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_ss(tAcc, umma_desc_kmajor(aH1 + (k >> 2) * kPB, k & 3), umma_desc_mnmajor(aW1, k, kPB), id_nn, k != 0);
        umma_commit(&bars[B_MMA + 4]);
        MGN_T(9);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          umma_ss(tW1, umma_desc_mnmajor(aH1, j, kPB), umma_desc_mnmajor(aA, j, kPB), id_tn, (it | j) != 0);
        umma_commit(&bars[B_MMA + 5]);
        MGN_T(10);
#undef MGN_W
      }
    }
  } else if (warp <= 4) {
    // =========================== reducers ===========================
    const int mt = tid - 32;
#define MGN_W(b, ph)                                                      \
  {                                                                       \
    const bool ok_ = __all_sync(0xffffffffu, wait_clk(&bars[b], ph));     \
    if (!ok_) {                                                           \
      timed_out = true;                                                   \
      break;                                                              \
    }                                                                     \
  }
#define MGN_PUBLISH(b)        \
  fence_proxy_async_smem();   \
  __syncwarp();               \
  if (lane == 0) mbar_arrive(&bars[b]);
    float cs_b1[8] = {0, 0, 0, 0, 0, 0, 0, 0}, cs_b2[8] = {0, 0, 0, 0, 0, 0, 0, 0}, cs_b3[8] = {0, 0, 0, 0, 0, 0, 0, 0},
          cs_beta[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const int rsub_m = mt >> 4;
    const bool go1_gathered = p.go1.idx != nullptr;
    int32_t r_go[16];
    // gathered parts of a tile's incoming gradient: go1 by rows -> X, go2 rows -> A
    auto stage_go = [&](long long r0, uint8_t* bXn, uint8_t* bAn) {
      if (go1_gathered) {
        fetch_row_ids(p.go1.idx, r0, p.M, rsub_m, r_go);
        stage_rows_async(bXn, p.go1, r_go, r0, p.M, mt);
      }
      if (has_go2) {
        if (!(go1_gathered && p.go2.idx == p.go1.idx)) fetch_row_ids(p.go2.idx, r0, p.M, rsub_m, r_go);
        stage_rows_async(bAn, p.go2, r_go, r0, p.M, mt);
      }
      cp_async_commit();
      cp_async_wait<0>();
    };
    if (n_my > 0) {
      stage_go(static_cast<long long>(blockIdx.x) * kRows, MGN_BUF(R_X, 0), MGN_BUF(R_A, 0));
      MGN_PUBLISH(B_GO);
    }
    for (int it = 0; it < n_my; ++it) {
      const uint32_t par = it & 1;
      const bool more = it + 1 < n_my;
      const long long row0 = (static_cast<long long>(blockIdx.x) + static_cast<long long>(it) * gridDim.x) * kRows;
      const long long row0n = row0 + static_cast<long long>(gridDim.x) * kRows;
      uint8_t* bA = MGN_BUF(R_A, it);
      uint8_t* bX = MGN_BUF(R_X, it);
      uint8_t* bH1 = MGN_BUF(R_H1, it);
      uint8_t* bH2 = MGN_BUF(R_H2, it);
      agg::TileSegs ts{};  // index loads of the tile's destination sums, issued before anything is waited for
      if (kAgg) ts = agg::tile_segments_begin<kAggLanes>(row0, p.M, p.seg_off, p.seg_id, mt);
      // after E3: X = g_out (summed), A = g_y
      MGN_W(B_E + 1, par);
      colsum_tile(bX, mt, cs_beta);
      colsum_tile(bA, mt, cs_b3);
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[B_CS + 0]);
      MGN_W(B_E + 2, par);
      colsum_tile(bH2, mt, cs_b2);
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[B_CS + 1]);
      MGN_W(B_E + 3, par);
      colsum_tile(bH1, mt, cs_b1);
      // (one segment at a time here: these warps also carry 32 column sums, the side-by-side form of the forward spills)
      if (kAgg) agg::tile_segment_sum<MGN_BWD2_AGG_AHEAD, MGN_BWD2_AGG_SIDE, kAggLanes>(bH1, row0, ts, p.seg_off, p.agg, p.ld_agg, p.agg_part, p.agg_part_v, mt);
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[B_CS + 2]);
      if (more) {  // next tile's gathered gradient rows: A is free after the layer-1 MMAs, X once g_efeat has left
        MGN_W(B_MMA + 5, par);
        if (go1_gathered) MGN_W(B_XF, par);
        stage_go(row0n, bX, bA);
        MGN_PUBLISH(B_GO);
      }
    }
#undef MGN_W
    asm volatile("bar.sync 10, 384;" ::: "memory");  // reducers + epilogue: nobody reads a tile buffer any more
    {
      const int chunk = mt & 15, rsub = mt >> 4;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        scratch[(0 * 8 + rsub) * kH + chunk * 8 + j] = cs_b1[j];
        scratch[(1 * 8 + rsub) * kH + chunk * 8 + j] = cs_b2[j];
        scratch[(2 * 8 + rsub) * kH + chunk * 8 + j] = cs_b3[j];
        scratch[(3 * 8 + rsub) * kH + chunk * 8 + j] = cs_beta[j];
      }
    }
  } else if (warp > kLoaderWarp) {
    // =========================== helpers of the fused destination sums ===========================
    if (kAgg) {
      const int mt = 128 + (tid - 32 * (kLoaderWarp + 1));  // segment lanes 8 .. kAggLanes - 1
      for (int it = 0; it < n_my; ++it) {
        const long long row0 = (static_cast<long long>(blockIdx.x) + static_cast<long long>(it) * gridDim.x) * kRows;
        const agg::TileSegs ts = agg::tile_segments_begin<kAggLanes>(row0, p.M, p.seg_off, p.seg_id, mt);
        if (!__all_sync(0xffffffffu, wait_clk(&bars[B_E + 3], it & 1))) {
          timed_out = true;
          break;
        }
        agg::tile_segment_sum<MGN_BWD2_AGG_AHEAD, MGN_BWD2_AGG_SIDE, kAggLanes>(MGN_BUF(R_H1, it), row0, ts, p.seg_off, p.agg, p.ld_agg,
                                                                              p.agg_part, p.agg_part_v, mt);
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[B_CS + 2]);
      }
    }
  } else if (warp == kLoaderWarp) {
    // =========================== loader (TMA) ===========================
    if (lane == 0) {
      const bool go1_tma = p.go1.idx == nullptr;
#define MGN_W(b, ph)             \
  if (!wait_clk(&bars[b], ph)) { \
    timed_out = true;            \
    break;                       \
  }
      if (n_my > 0) {
        const long long row00 = static_cast<long long>(blockIdx.x) * kRows;
        mbar_arrive_expect_tx(&bars[B_H1L], 2 * kPB);
        tma_tile(MGN_BUF(R_H1, 0), &p.m_h1, row00, &bars[B_H1L]);
        if (go1_tma) {
          mbar_arrive_expect_tx(&bars[B_GO], 2 * kPB);
          tma_tile(MGN_BUF(R_X, 0), &p.m_go1, row00, &bars[B_GO]);
        } else {
          mbar_arrive(&bars[B_GO]);
        }
      }
      for (int it = 0; it < n_my; ++it) {
        const uint32_t par = it & 1;
        const bool more = it + 1 < n_my;
        const long long row0 = (static_cast<long long>(blockIdx.x) + static_cast<long long>(it) * gridDim.x) * kRows;
        const long long row0n = row0 + static_cast<long long>(gridDim.x) * kRows;
        uint8_t* bA = MGN_BUF(R_A, it);
        uint8_t* bX = MGN_BUF(R_X, it);
        uint8_t* bH1 = MGN_BUF(R_H1, it);
        uint8_t* bH2 = MGN_BUF(R_H2, it);
        MGN_T(0);
        // efeat tile for the layer-1 weight gradient, once the layer-3 MMAs and the column sums are done with g_y
        MGN_W(B_W3, par);
        MGN_W(B_CS + 0, par);
        MGN_T(1);
        mbar_arrive_expect_tx(&bars[B_A2], 2 * kPB);
        tma_tile(bA, &p.m_a, row0, &bars[B_A2]);
        if (more) {  // next tile's h1 -> this tile's H2 buffer (free after the layer-2 MMAs and its column sums)
          MGN_W(B_W2, par);
          MGN_W(B_CS + 1, par);
          MGN_T(2);
          mbar_arrive_expect_tx(&bars[B_H1L], 2 * kPB);
          tma_tile(bH2, &p.m_h1, row0n, &bars[B_H1L]);
        }
        // g_z1 tile (H1) -> global
        MGN_W(B_E + 3, par);
        MGN_T(3);
        tma_store_2d(&p.m_gz1, smem_u32(bH1), 0, static_cast<int>(row0));
        tma_store_2d(&p.m_gz1, smem_u32(bH1) + kPB, 64, static_cast<int>(row0));
        tma_store_commit();
        tma_store_wait_read();
        mbar_arrive(&bars[B_ST]);
        MGN_T(4);
        // g_efeat tile (X) -> global, then the next tile's dense incoming gradient into the same buffer
        MGN_W(B_E + 4, par);
        MGN_T(6);
        tma_store_2d(&p.m_ga, smem_u32(bX), 0, static_cast<int>(row0));
        tma_store_2d(&p.m_ga, smem_u32(bX) + kPB, 64, static_cast<int>(row0));
        tma_store_commit();
        tma_store_wait_read();
        mbar_arrive(&bars[B_XF]);
        if (more) {
          if (go1_tma) {
            mbar_arrive_expect_tx(&bars[B_GO], 2 * kPB);
            tma_tile(bX, &p.m_go1, row0n, &bars[B_GO]);
          } else {
            mbar_arrive(&bars[B_GO]);
          }
        }
        MGN_T(7);
      }
#undef MGN_W
      tma_store_wait_all();
    }
  } else {
    // =========================== epilogue (8 warps) ===========================
    const int q = warp & 3;
    const int ch = (warp - 5) >> 2;
    const int row = q * 32 + lane;
    const int c0 = ch * 64;
    const uint32_t t_acc = tAcc + (static_cast<uint32_t>(q * 32) << 16) + c0;
    const float* b2 = sPar + kH + c0;
    const float* b3 = sPar + 2 * kH + c0;
    const float* gam = sPar + 3 * kH + c0;
    const uint32_t xch_own = ch * kPB + sw128_offset(row, 0);
    const uint32_t xch_other = (ch ^ 1) * kPB + sw128_offset(row, 0);
    float gg[4] = {0.f, 0.f, 0.f, 0.f};
#define MGN_W(b, ph)                                                      \
  {                                                                       \
    const bool ok_ = __all_sync(0xffffffffu, wait_clk(&bars[b], ph));     \
    if (!ok_) {                                                           \
      timed_out = true;                                                   \
      break;                                                              \
    }                                                                     \
  }
#define MGN_EPI_DONE(b)       \
  fence_proxy_async_smem();   \
  tc_fence_before_sync();     \
  __syncwarp();               \
  if (lane == 0) mbar_arrive(&bars[b]);
#define MGN_ROW_SYNC() asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory")
    for (int it = 0; it < n_my; ++it) {
      const uint32_t par = it & 1;
      uint8_t* bA = MGN_BUF(R_A, it);
      uint8_t* bX = MGN_BUF(R_X, it);
      uint8_t* bH1 = MGN_BUF(R_H1, it);
      uint8_t* bH2 = MGN_BUF(R_H2, it);
      // ---- E2: h2 = relu(acc + b2) -> H2   (H2 is the previous tile's H1: its layer-1 MMAs, its g_z1 store and its
      //      column sums must be done with it)
      MGN_W(B_MMA + 0, par);
      if (it > 0) {
        MGN_W(B_MMA + 5, par ^ 1);
        MGN_W(B_ST, par ^ 1);
        MGN_W(B_CS + 2, par ^ 1);
      }
      MGN_T(0);
      tc_fence_after_sync();
#ifdef MGN_NO_PIPE16
#pragma unroll 1
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t v[32];
        tmem_ld32(t_acc + 32 * hh, v);
        tmem_ld_wait();
        uint32_t o[16];
#pragma unroll
        for (int j = 0; j < 16; ++j)  // relu(round(x)) == round(relu(x)): FADD2 + F2FP + HMNMX2 per pair
          o[j] = relu_bf16x2(f2_to_bf16x2(f2_add(f2_packu(v[2 * j], v[2 * j + 1]), f2_ld(b2 + 32 * hh + 2 * j))));
        row_store32p(bH2, row, c0 + 32 * hh, o);
      }
#else
      tmem_pass64(t_acc, [&](int i, const uint32_t(&v)[16]) {
        uint32_t o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)  // relu(round(x)) == round(relu(x)): FADD2 + F2FP + HMNMX2 per pair
          o[j] = relu_bf16x2(f2_to_bf16x2(f2_add(f2_packu(v[2 * j], v[2 * j + 1]), f2_ld(b2 + 16 * i + 2 * j))));
        row_store16p(bH2, row, c0 + 16 * i, o);
      });
#endif
      MGN_EPI_DONE(B_E + 0);
      MGN_T(1);
      // ---- E3: LayerNorm backward: g_out = go1 (+ go2) -> X ; g_y -> A
      MGN_W(B_MMA + 1, par);
      MGN_T(2);
      MGN_W(B_GO, par);
      MGN_T(3);
      tc_fence_after_sync();
      {
        // g_out = go1 (+ go2) -> X; LayerNorm statistics of y = acc + b3 and of ghat = g_out * gamma.  All fp32 arithmetic
        // runs two lanes per instruction (f2_*); the go1 + go2 sum is one packed bf16 add per pair.
        float s_y, s_yy, s_g, s_gy;
#ifdef MGN_NO_PIPE16
        {
          uint64_t sy2 = 0ull, syy2 = 0ull, sg2 = 0ull, sgy2 = 0ull;
#pragma unroll 1
          for (int hh = 0; hh < 2; ++hh) {
            const int cc = c0 + 32 * hh;
            uint32_t v[32];
            tmem_ld32(t_acc + 32 * hh, v);
            uint32_t go[16];
            row_load32p(bX, row, cc, go);
            if (has_go2) {
              uint32_t g2[16];
              row_load32p(bA, row, cc, g2);
#pragma unroll
              for (int j = 0; j < 16; ++j) go[j] = add_bf16x2(go[j], g2[j]);
              row_store32p(bX, row, cc, go);
            }
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const uint64_t y2 = f2_add(f2_packu(v[2 * j], v[2 * j + 1]), f2_ld(b3 + 32 * hh + 2 * j));
              const uint64_t gh2 = f2_mul(f2_from_bf16x2(go[j]), f2_ld(gam + 32 * hh + 2 * j));
              sy2 = f2_add(sy2, y2);
              syy2 = f2_fma(y2, y2, syy2);
              sg2 = f2_add(sg2, gh2);
              sgy2 = f2_fma(gh2, y2, sgy2);
            }
          }
          s_y = f2_lo(sy2) + f2_hi(sy2);
          s_yy = f2_lo(syy2) + f2_hi(syy2);
          s_g = f2_lo(sg2) + f2_hi(sg2);
          s_gy = f2_lo(sgy2) + f2_hi(sgy2);
        }
#else
        {
          uint64_t sy2 = 0ull, syy2 = 0ull, sg2 = 0ull, sgy2 = 0ull;
          tmem_pass64(t_acc, [&](int i, const uint32_t(&v)[16]) {
            const int cc = c0 + 16 * i;
            uint32_t go[8];
            row_load16p(bX, row, cc, go);
            if (has_go2) {
              uint32_t g2[8];
              row_load16p(bA, row, cc, g2);
#pragma unroll
              for (int j = 0; j < 8; ++j) go[j] = add_bf16x2(go[j], g2[j]);
              row_store16p(bX, row, cc, go);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const uint64_t y2 = f2_add(f2_packu(v[2 * j], v[2 * j + 1]), f2_ld(b3 + 16 * i + 2 * j));
              const uint64_t gh2 = f2_mul(f2_from_bf16x2(go[j]), f2_ld(gam + 16 * i + 2 * j));
              sy2 = f2_add(sy2, y2);
              syy2 = f2_fma(y2, y2, syy2);
              sg2 = f2_add(sg2, gh2);
              sgy2 = f2_fma(gh2, y2, sgy2);
            }
          });
          s_y = f2_lo(sy2) + f2_hi(sy2);
          s_yy = f2_lo(syy2) + f2_hi(syy2);
          s_g = f2_lo(sg2) + f2_hi(sg2);
          s_gy = f2_lo(sgy2) + f2_hi(sgy2);
        }
#endif
        *reinterpret_cast<float4*>(bA + xch_own) = make_float4(s_y, s_yy, s_g, s_gy);
        MGN_ROW_SYNC();
        {
          const float4 t = *reinterpret_cast<const float4*>(bA + xch_other);
          s_y += t.x;
          s_yy += t.y;
          s_g += t.z;
          s_gy += t.w;
        }
        MGN_ROW_SYNC();  // both halves have read the exchange before g_y overwrites the A buffer
        const float mu = s_y * (1.f / kH);
        const float var = fmaxf(s_yy * (1.f / kH) - mu * mu, 0.f);
        const float rstd = rsqrtf(var + p.eps);
        const float m1 = s_g * (1.f / kH);
        const float m2 = (s_gy - mu * s_g) * rstd * (1.f / kH);  // mean(ghat * xhat)
        // g_y = rstd (ghat - m1 - xhat m2) with xhat = rstd (y - mu), expanded so that each element is two FMAs:
        //   g_y = rstd * ghat - k2 * y + k0,   k2 = rstd^2 m2,   k0 = k2 mu - rstd m1;   xhat = rstd * y - rstd mu
        const float k2 = rstd * rstd * m2;
        const uint64_t A2 = f2_splat(rstd), NK2 = f2_splat(-k2), K0 = f2_splat(fmaf(k2, mu, -rstd * m1)),
                       NAMU = f2_splat(-rstd * mu);
#ifdef MGN_NO_PIPE16
#pragma unroll 1
        for (int hh = 0; hh < 2; ++hh) {
          const int cc = c0 + 32 * hh;
          uint32_t v[32];
          tmem_ld32(t_acc + 32 * hh, v);
          uint32_t go[16];
          row_load32p(bX, row, cc, go);
          tmem_ld_wait();
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            float t[16];
            uint32_t o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int c = 16 * h + 2 * j;  // column within this 32-column half
              const uint64_t y2 = f2_add(f2_packu(v[c], v[c + 1]), f2_ld(b3 + 32 * hh + c));
              const uint64_t g2 = f2_from_bf16x2(go[c >> 1]);
              const uint64_t gh2 = f2_mul(g2, f2_ld(gam + 32 * hh + c));
              o[j] = f2_to_bf16x2(f2_fma(A2, gh2, f2_fma(NK2, y2, K0)));
              const uint64_t tt = f2_mul(g2, f2_fma(A2, y2, NAMU));  // gamma-gradient contribution g_out * xhat
              t[2 * j] = f2_lo(tt);
              t[2 * j + 1] = f2_hi(tt);
            }
            uint8_t* base = bA + ch * kPB;
            const int c8 = 4 * hh + 2 * h;
            *reinterpret_cast<uint4*>(base + sw128_offset(row, c8)) = make_uint4(o[0], o[1], o[2], o[3]);
            *reinterpret_cast<uint4*>(base + sw128_offset(row, c8 + 1)) = make_uint4(o[4], o[5], o[6], o[7]);
            const float cs = warp_colsum16(t, lane);
            if (hh == 0) gg[h] += cs;
            else gg[2 + h] += cs;
          }
        }
#else
        tmem_pass64(t_acc, [&](int i, const uint32_t(&v)[16]) {
          uint32_t go[8];
          row_load16p(bX, row, c0 + 16 * i, go);
          float t[16];
          uint32_t o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint64_t y2 = f2_add(f2_packu(v[2 * j], v[2 * j + 1]), f2_ld(b3 + 16 * i + 2 * j));
            const uint64_t g2 = f2_from_bf16x2(go[j]);
            const uint64_t gh2 = f2_mul(g2, f2_ld(gam + 16 * i + 2 * j));
            o[j] = f2_to_bf16x2(f2_fma(A2, gh2, f2_fma(NK2, y2, K0)));
            const uint64_t tt = f2_mul(g2, f2_fma(A2, y2, NAMU));  // gamma-gradient contribution g_out * xhat
            t[2 * j] = f2_lo(tt);
            t[2 * j + 1] = f2_hi(tt);
          }
          row_store16p(bA, row, c0 + 16 * i, o);
          gg[i] += warp_colsum16(t, lane);
        });
#endif
      }
      MGN_EPI_DONE(B_E + 1);
      MGN_T(4);
      // ---- E4: g_z2 = acc * (h2 > 0), in place in H2
      MGN_W(B_MMA + 2, par);
      tc_fence_after_sync();
      {
        // (the layer's weight-gradient MMAs still read this buffer: compute into registers, store once they are done)
#ifdef MGN_NO_PIPE16
        uint32_t hq[2][16];
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t v[32];
          tmem_ld32(t_acc + 32 * hh, v);
          row_load32p(bH2, row, c0 + 32 * hh, hq[hh]);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j)  // F2FP + HSET2 + LOP3 per pair
            hq[hh][j] = mask_pos_bf16x2(pack_bf16x2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])), hq[hh][j]);
        }
        MGN_W(B_W3, par);
        row_store32p(bH2, row, c0, hq[0]);
        row_store32p(bH2, row, c0 + 32, hq[1]);
#else
        uint32_t hq[4][8];
        tmem_pass64(t_acc, [&](int i, const uint32_t(&v)[16]) {
          row_load16p(bH2, row, c0 + 16 * i, hq[i]);
#pragma unroll
          for (int j = 0; j < 8; ++j)  // F2FP + HSET2 + LOP3 per pair
            hq[i][j] = mask_pos_bf16x2(pack_bf16x2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])), hq[i][j]);
        });
        MGN_W(B_W3, par);
#pragma unroll
        for (int i = 0; i < 4; ++i) row_store16p(bH2, row, c0 + 16 * i, hq[i]);
#endif
      }
      MGN_EPI_DONE(B_E + 2);
      MGN_T(5);
      // ---- E5: g_z1 = acc * (h1 > 0), in place in H1
      MGN_W(B_MMA + 3, par);
      tc_fence_after_sync();
      {
        // (the layer's weight-gradient MMAs still read this buffer: compute into registers, store once they are done)
#ifdef MGN_NO_PIPE16
        uint32_t hq[2][16];
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t v[32];
          tmem_ld32(t_acc + 32 * hh, v);
          row_load32p(bH1, row, c0 + 32 * hh, hq[hh]);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j)  // F2FP + HSET2 + LOP3 per pair
            hq[hh][j] = mask_pos_bf16x2(pack_bf16x2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])), hq[hh][j]);
        }
        MGN_W(B_W2, par);
        row_store32p(bH1, row, c0, hq[0]);
        row_store32p(bH1, row, c0 + 32, hq[1]);
#else
        uint32_t hq[4][8];
        tmem_pass64(t_acc, [&](int i, const uint32_t(&v)[16]) {
          row_load16p(bH1, row, c0 + 16 * i, hq[i]);
#pragma unroll
          for (int j = 0; j < 8; ++j)  // F2FP + HSET2 + LOP3 per pair
            hq[i][j] = mask_pos_bf16x2(pack_bf16x2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])), hq[i][j]);
        });
        MGN_W(B_W2, par);
#pragma unroll
        for (int i = 0; i < 4; ++i) row_store16p(bH1, row, c0 + 16 * i, hq[i]);
#endif
      }
      MGN_EPI_DONE(B_E + 3);
      MGN_T(6);
      // ---- E6: g_A = acc (+ g_out), in place in X (runs while the layer-1 weight-gradient MMAs execute)
      MGN_W(B_MMA + 4, par);
      tc_fence_after_sync();
      {
#ifdef MGN_NO_PIPE16
#pragma unroll 1
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t v[32];
          tmem_ld32(t_acc + 32 * hh, v);
          uint32_t go[16];
          if (kAddGout) row_load32p(bX, row, c0 + 32 * hh, go);
          tmem_ld_wait();
          if (hh == 1) {  // accumulator drained: the next tile's GEMM2 may overwrite it
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[B_DR]);
          }
#pragma unroll
          for (int j = 0; j < 16; ++j)
            go[j] = kAddGout ? f2_to_bf16x2(f2_add(f2_packu(v[2 * j], v[2 * j + 1]), f2_from_bf16x2(go[j])))
                             : pack_bf16x2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
          row_store32p(bX, row, c0 + 32 * hh, go);
        }
      #else
        tmem_pass64(t_acc, [&](int i, const uint32_t(&v)[16]) {
          uint32_t go[8];
          if (kAddGout) row_load16p(bX, row, c0 + 16 * i, go);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            go[j] = kAddGout ? f2_to_bf16x2(f2_add(f2_packu(v[2 * j], v[2 * j + 1]), f2_from_bf16x2(go[j])))
                             : pack_bf16x2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
          row_store16p(bX, row, c0 + 16 * i, go);
        });
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[B_DR]);
#endif
      }
      MGN_EPI_DONE(B_E + 4);
      MGN_T(7);
    }
#undef MGN_W
    asm volatile("bar.sync 10, 384;" ::: "memory");  // reducers + epilogue: nobody reads a tile buffer any more
    if (lane < 16) {
#pragma unroll
      for (int g = 0; g < 4; ++g) scratch[4 * 8 * kH + q * kH + c0 + g * 16 + lane] = gg[g];
    }
  }
  if (tm_on) {
    const int role = warp == 0 ? 0 : (warp == kLoaderWarp ? 1 : 2);
    for (int i = 0; i < 16; ++i) p.timing[role * 32 + i] = tm[i];
  }
  if (timed_out && p.status != nullptr) atomicOr(p.status, 1);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
#ifdef MGN_DEBUG_HOOKS
  if (p.timing != nullptr && tid == 0) {
    long long g1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
    p.timing[96 + 4 * blockIdx.x] = clock64() - dbg_c0;
    p.timing[96 + 4 * blockIdx.x + 1] = g1 - dbg_g0;
    p.timing[96 + 4 * blockIdx.x + 2] = n_my;
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    p.timing[96 + 4 * blockIdx.x + 3] = smid;
  }
#endif

  // ---------------- write this CTA's partial gradients ----------------
  float* part = p.partials + static_cast<long long>(blockIdx.x) * p.part_floats;
  if (warp >= 5 && warp < kLoaderWarp) {
    const int q = warp & 3;
    const int ch = (warp - 5) >> 2;
    const int row = q * 32 + lane;  // TMEM lane = output-feature row of the weight gradient
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
#pragma unroll 1
    for (int w = 0; w < 3; ++w) {
      const uint32_t t = (w == 0 ? tW1 : (w == 1 ? tW2 : tW3)) + lane_off;
      const int ncol = kH;
      float* dst = part + (w == 0 ? Part::kW1 : (w == 1 ? Part::kW2 : Part::kW3)) + row * ncol;
#pragma unroll 1
      for (int g = ch * (ncol / 64); g < (ch + 1) * (ncol / 64); ++g) {  // this warp's half of the columns
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
  if (tid < kH) {
    float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int k = 0; k < 4; ++k) s[k] += scratch[(k * 8 + r) * kH + tid];
    }
    float sg = 0.f;
#pragma unroll
    for (int w = 0; w < 4; ++w) sg += scratch[4 * 8 * kH + w * kH + tid];
    part[Part::kB1 + tid] = s[0];
    part[Part::kB2 + tid] = s[1];
    part[Part::kB3 + tid] = s[2];
    part[Part::kBeta + tid] = s[3];
    part[Part::kGamma + tid] = sg;
  }
  tc_fence_before_sync();
  __syncthreads();  // every tcgen05.ld of the dump above has completed
  if (warp == 0) tmem_dealloc(tmem, 512);
}


}  // namespace bwd2
}  // namespace mgn

using namespace mgn;

static int bwd2_grid(int64_t M) {
  const long long n_tiles = (M + bwd2::kRows - 1) / bwd2::kRows;
  return static_cast<int>(n_tiles < num_sms() ? n_tiles : num_sms());
}

#ifdef MGN_DEBUG_HOOKS
static long long* g_bwd2_timing = nullptr;
extern "C" int mgn_debug_set_edge_bwd2_timing(void* dev_buf) {
  g_bwd2_timing = static_cast<long long*>(dev_buf);
  return MGN_OK;
}
#else
static constexpr long long* g_bwd2_timing = nullptr;
#endif

extern "C" size_t mgn_edge_block_bwd_tc_workspace_bytes(int64_t n_edges) {
  if (n_edges <= 0) return 0;
  return static_cast<size_t>(bwd2_grid(n_edges)) * bwd2::Part::kTotal * sizeof(float);
}

extern "C" int mgn_edge_block_bwd_tc(const void* efeat, const void* h1, const void* go1, const int32_t* go1_idx,
                                     const void* go2, const int32_t* go2_idx, int64_t n_edges, const float* w1,
                                     int64_t ld_w1, const float* w2, const float* b2, const float* w3, const float* b3,
                                     const float* gamma, float eps, int add_gout, void* g_efeat, void* g_z1, int64_t g_z1_ld,
                                     float* g_w1, int64_t ld_gw1, float* g_b1, float* g_w2, float* g_b2, float* g_w3,
                                     float* g_b3, float* g_gamma, float* g_beta, void* workspace, size_t workspace_bytes,
                                     const int32_t* csc_offsets, const int32_t* dst_idx, int64_t n_dst, void* gz1_agg,
                                     int64_t ld_agg, void* agg_workspace, size_t agg_workspace_bytes, int* status,
                                     mgn_stream_t stream) {
  const int64_t M = n_edges;
  MGN_CHECK_ARG(M >= 0 && w1 && w2 && w3 && gamma && ld_w1 >= bwd2::kH);
  if (M == 0) return MGN_OK;
  MGN_CHECK_ARG(efeat && h1 && go1 && g_efeat && g_z1 && workspace);
  MGN_CHECK_ARG(g_z1_ld >= bwd2::kH && g_z1_ld % 8 == 0);
  for (const void* q : {efeat, h1, go1, go2, static_cast<const void*>(g_efeat), static_cast<const void*>(g_z1)})
    MGN_CHECK_ARG((reinterpret_cast<uintptr_t>(q) & 15) == 0);
  if (workspace_bytes < mgn_edge_block_bwd_tc_workspace_bytes(M)) return MGN_EWORKSPACE;
  bwd2::Params p{};
  p.h1 = static_cast<const bf16*>(h1);
  p.go1 = tile::RowSrc{static_cast<const bf16*>(go1), go1_idx, bwd2::kH, 0};
  p.go2 = tile::RowSrc{static_cast<const bf16*>(go2), go2_idx, bwd2::kH, 0};
  MGN_CHECK_ARG(go2 == nullptr || go2_idx != nullptr);
  p.M = M;
  p.w1 = w1; p.w2 = w2; p.b2 = b2; p.w3 = w3; p.b3 = b3; p.gamma = gamma;
  p.ld_w1 = ld_w1;
  p.eps = eps;
  p.partials = static_cast<float*>(workspace);
  p.part_floats = bwd2::Part::kTotal;
  p.status = status;
  p.timing = g_bwd2_timing;
  int e = tma_make_rows_map(&p.m_a, efeat, M, bwd2::kH, 128);
  e |= tma_make_rows_map(&p.m_h1, h1, M, bwd2::kH, 128);
  if (go1_idx == nullptr) e |= tma_make_rows_map(&p.m_go1, go1, M, bwd2::kH, 128);
  e |= tma_make_rows_map(&p.m_ga, g_efeat, M, bwd2::kH, 128);
  e |= tma_make_rows_map(&p.m_gz1, g_z1, M, g_z1_ld, 128);
  if (e != 0) return MGN_EINVAL;
  const bool with_agg = csc_offsets != nullptr;
  if (with_agg) {  // rows are CSC-ordered edges with destinations dst_idx
    MGN_CHECK_ARG(add_gout && dst_idx && gz1_agg && n_dst > 0 && ld_agg >= bwd2::kH && ld_agg % 8 == 0 &&
                  (reinterpret_cast<uintptr_t>(gz1_agg) & 15) == 0 && agg_workspace != nullptr);
    if (agg_workspace_bytes < agg::workspace_bytes(M)) return MGN_EWORKSPACE;
    const long long n_tiles_ = (M + bwd2::kRows - 1) / bwd2::kRows;
    p.seg_off = csc_offsets;
    p.seg_id = dst_idx;
    p.agg = static_cast<bf16*>(gz1_agg);
    p.ld_agg = ld_agg;
    p.agg_part = static_cast<float*>(agg_workspace);
    p.agg_part_v = reinterpret_cast<int32_t*>(p.agg_part + 2 * n_tiles_ * bwd2::kH);
  }
  static PerDeviceFlag configured_flag;
  bool& configured = configured_flag.get();
  if (!configured) {
    cudaError_t ce = cudaFuncSetAttribute(bwd2::edge_bwd2_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bwd2::Smem::kTotal);
    if (ce == cudaSuccess)
      ce = cudaFuncSetAttribute(bwd2::edge_bwd2_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bwd2::Smem::kTotal);
    if (ce == cudaSuccess)
      ce = cudaFuncSetAttribute(bwd2::edge_bwd2_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bwd2::Smem::kTotal);
    if (ce != cudaSuccess) return static_cast<int>(ce);
    configured = true;
  }
  cudaStream_t st = as_stream(stream);
  const int grid = bwd2_grid(M);
  if (with_agg) bwd2::edge_bwd2_kernel<true, true><<<grid, bwd2::kThreads, bwd2::Smem::kTotal, MGN_ST(st)>>>(p);
  else if (add_gout) bwd2::edge_bwd2_kernel<true, false><<<grid, bwd2::kThreads, bwd2::Smem::kTotal, MGN_ST(st)>>>(p);
  else bwd2::edge_bwd2_kernel<false, false><<<grid, bwd2::kThreads, bwd2::Smem::kTotal, MGN_ST(st)>>>(p);
  int rc = mgn_launch_status();
  if (rc != MGN_OK) return rc;
  ReduceParams rp{};
  rp.partials = p.partials;
  rp.stride = p.part_floats;
  rp.n_parts = grid;
  using PT = bwd2::Part;
  int ns = 0;
  rp.seg[ns++] = ReduceSeg{g_w1, ld_gw1, bwd2::kH, bwd2::kH, PT::kW1, bwd2::kH};
  rp.seg[ns++] = ReduceSeg{g_w2, bwd2::kH, bwd2::kH, bwd2::kH, PT::kW2, bwd2::kH};
  rp.seg[ns++] = ReduceSeg{g_w3, bwd2::kH, bwd2::kH, bwd2::kH, PT::kW3, bwd2::kH};
  rp.seg[ns++] = ReduceSeg{g_b1, bwd2::kH, 1, bwd2::kH, PT::kB1, bwd2::kH};
  rp.seg[ns++] = ReduceSeg{g_b2, bwd2::kH, 1, bwd2::kH, PT::kB2, bwd2::kH};
  rp.seg[ns++] = ReduceSeg{g_b3, bwd2::kH, 1, bwd2::kH, PT::kB3, bwd2::kH};
  rp.seg[ns++] = ReduceSeg{g_gamma, bwd2::kH, 1, bwd2::kH, PT::kGamma, bwd2::kH};
  rp.seg[ns++] = ReduceSeg{g_beta, bwd2::kH, 1, bwd2::kH, PT::kBeta, bwd2::kH};
  rp.n_seg = ns;
  reduce_cta_partials_kernel<<<dim3(64, ns), 256, 0, MGN_ST(st)>>>(rp);
  rc = mgn_launch_status();
  if (rc != MGN_OK || !with_agg) return rc;
  const long long n_rec = 2 * ((M + bwd2::kRows - 1) / bwd2::kRows);
  agg::agg_fixup_kernel<<<static_cast<unsigned>((n_rec * 32 + 255) / 256), 256, 0, MGN_ST(st)>>>(
      p.agg_part, p.agg_part_v, n_rec, p.agg, p.ld_agg, n_dst);
  return mgn_launch_status();
}
