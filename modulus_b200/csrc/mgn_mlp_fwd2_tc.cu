// mgn_mlp_fwd2_tc.cu — fused MeshGraphMLP forward, second generation (tcgen05 / TMEM, bf16, hidden 128, ReLU).
//
//     z1  = A W1^T + b1 + G1[r] + G2[r]        A: 128-row tile of layer-input rows, G*: additive gathered rows
//     h1  = relu(z1) ; h2 = relu(h1 W2^T + b2) ; y = h2 W3^T + b3
//     out = [LayerNorm(y) * gamma + beta] [+ residual]
//
// = MeshEdgeBlock.forward / MeshNodeBlock.forward / encoder+decoder MeshGraphMLP of the reference
// (physicsnemo/models/gnn_layers/mesh_edge_block.py:88-96, mesh_node_block.py:82-92, mesh_graph_mlp.py:142-203)
// with the first Linear split per input block (see modulus_b200/fused.py).
//
// Data movement rules learnt from the first-generation kernel (profiles/r01_*):
//   * every global->shared byte moves by cp.async (16-byte, L1-bypassing): with > 200 KB of shared memory the L1
//     is a few KB and LDG-based staging collapses to a handful of lines in flight;
//   * row indices are fetched one tile ahead (the L1TEX queue returns loads in order behind the gathers);
//   * epilogue threads (thread = row = TMEM lane) touch TMEM and shared memory only; the output tile is written
//     in place over the A tile (the residual is read from shared memory, never re-fetched) and leaves the SM
//     through coalesced 16-byte stores issued by the mover warps;
//   * hidden activations go back to TMEM as packed bf16 and feed the next GEMM as the A operand.
//
// Warp roles (416 threads): warp 0 = MMA issuer + TMEM owner, warps 1-4 = movers, warps 5-12 = epilogue
// (two warps per TMEM lane quarter, 64 columns each; LayerNorm row sums are exchanged through spare TMEM columns).
// Shared memory: W1, W2, W3 (96 KB, converted fp32 -> bf16 once per CTA from the optimizer's tensors),
// two A buffers (the next tile's rows stream in while the current tile computes), one G1 and one G2 buffer.
#include "mgn_common.cuh"
#include "mgn_tc.cuh"
#include "mgn_tile.cuh"
#include "mgn_tma.cuh"
#include "mgn_agg.cuh"
#include "mgn_edge_fwd3.h"

namespace mgn {
namespace fwd2 {

using namespace tile;
constexpr int kLoaderWarp = 13;  // TMA: dense A tiles in, result tiles out
constexpr int kThreads = 32 * (kLoaderWarp + 1);
constexpr int kH = 128;

struct Params {
  RowSrc a;             // layer-1 input rows [*,128]                      (KP == 2)
  const void* small_x;  // raw [M, small_in] features zero-padded to K=64   (KP == 1)
  int small_in;
  int small_is_f32;
  RowSrc g1, g2;        // additive rows of layer 1 (tab == nullptr: absent)
  RowSrc res;           // residual rows [*,128] (tab == nullptr: none); res_is_a: the A rows themselves
  int res_is_a;
  int single;           // 1: out = A W1^T + b3 (+ residual): one GEMM (node-level projections)
  long long M;
  const float *w1, *b1, *w2, *b2, *w3, *b3, *gamma, *beta;
  long long ld_w1;
  int k1_true;
  int n_out;
  float eps;
  bf16* out;
  long long ld_out;
  bf16* h1_out;  // [M,128] relu(z1), kept for mgn_edge_block_bwd_tc (nullptr: not stored)
  int* status;
  long long* timing;
  int tma_a, tma_out;  // dense A tiles arrive / result tiles leave through the loader warp (tensor maps below)
  // fused aggregate_and_concat sum (models/gnn_layers/utils.py:337-378): the output rows are CSC-ordered edges, so a
  // destination's rows are contiguous; the mover warps sum each segment from the result tile in shared memory.
  // Segments wholly inside a tile go straight to agg; a tile's first and last segment go to fp32 records that
  // agg_fixup_kernel combines in tile order (deterministic)
  const int32_t* seg_off;  // [n_seg + 1] CSC offsets (nullptr: no aggregation)
  const int32_t* seg_id;   // [M] destination of every row, ascending
  long long n_seg;
  bf16* agg;               // [n_seg, 128], row stride ld_agg
  long long ld_agg;
  float* agg_part;         // [2 * n_tiles][128]
  int32_t* agg_part_v;     // [2 * n_tiles] destination id of each record, -1: empty
  long long agg_row_base;  // CSC position of this launch's row 0
  long long agg_rec_base;  // record index (in tiles) of this launch's first tile
  alignas(64) CUtensorMap m_a, m_out;
};

// B_IN: next tile staged (four mover warps: gathered / small rows; loader: dense A tile by TMA);
// B_STD: the result tile of a tile has left its A buffer
// B_AGG: the mover warps have finished reading a result tile for the fused aggregation
// B_RES: the residual rows of a tile (node block: + nfeat) are staged in the G2 buffer.  They get their own barrier: the
// buffer is busy until the END of the previous tile, and gating B_IN on it made every tile's first GEMM wait for a gather
// issued after the previous tile's last epilogue (node forward 13.8 k cycles per tile, 4.6 k of them in that wait).
enum { B_IN = 0, B_M1 = 1, B_M2 = 2, B_M3 = 3, B_H1 = 4, B_H2 = 5, B_OUT = 6, B_STD = 7, B_AGG = 8, B_RES = 9, B_NUM = 10 };

template <int KP>
struct Smem {
  static constexpr int kW1 = 0;
  static constexpr int kW2 = KP * kPB;
  static constexpr int kW3 = kW2 + 2 * kPB;
  static constexpr int kA = kW3 + 2 * kPB;   // 2 buffers x 2 panels
  static constexpr int kG1 = kA + 4 * kPB;
  static constexpr int kG2 = kG1 + 2 * kPB;
  static constexpr int kPar = kG2 + 2 * kPB;  // b1, b2, b3, gamma, beta
  static constexpr int kBars = kPar + 5 * kH * 4;
  static constexpr int kTmemSlot = kBars + 10 * 8;
  static constexpr int kTiming = kTmemSlot + 16;  // 3 roles x 8 x int64
  static constexpr int kTotal = kTiming + 3 * 8 * 8;
};

#define MGN_T(i)                      \
  if (tm_on) {                        \
    const long long t_ = clock64();   \
    tm[i] += t_ - tlast;              \
    tlast = t_;                       \
  }

template <int KP>
__global__ void __launch_bounds__(kThreads, 1) mlp3_fwd2_tc_kernel(const __grid_constant__ Params p) {
  using L = Smem<KP>;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if ((smem_u32(smem) & 1023u) != 0) {
    if (tid == 0 && p.status) atomicOr(p.status, 2);
    return;
  }
  uint8_t* sW1 = smem + L::kW1;
  uint8_t* sW2 = smem + L::kW2;
  uint8_t* sW3 = smem + L::kW3;
  uint8_t* bA0 = smem + L::kA;
  uint8_t* bG1 = smem + L::kG1;
  uint8_t* bG2 = smem + L::kG2;
  float* sPar = reinterpret_cast<float*>(smem + L::kPar);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kBars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::kTmemSlot);

  const bool has_ln = p.gamma != nullptr;
  const bool has_g1 = p.g1.tab != nullptr;
  const bool has_g2 = p.g2.tab != nullptr;
  const bool res_g2 = p.res.tab != nullptr && !p.res_is_a;  // residual rows travel in the G2 buffer
  const bool direct_out = p.n_out < kH;                      // narrow outputs (decoder) are stored by the epilogue

  // ---------------- one-time setup ----------------
  stage_weight_ld(sW1, p.w1, p.ld_w1, kH, p.k1_true, KP, tid, kThreads);
  if (!p.single) {
    stage_weight_ld(sW2, p.w2, kH, kH, kH, 2, tid, kThreads);
    stage_weight_ld(sW3, p.w3, kH, p.n_out, kH, 2, tid, kThreads);
  }
  for (int i = tid; i < kH; i += kThreads) {
    sPar[i] = p.b1 ? p.b1[i] : 0.f;
    sPar[kH + i] = p.b2 ? p.b2[i] : 0.f;
    sPar[2 * kH + i] = (p.b3 && i < p.n_out) ? p.b3[i] : 0.f;
    sPar[3 * kH + i] = has_ln ? p.gamma[i] : 1.f;
    sPar[4 * kH + i] = (has_ln && p.beta) ? p.beta[i] : 0.f;
  }
  if (tid == 0) {
    mbar_init(&bars[B_IN], 5);
    mbar_init(&bars[B_STD], 1);
    mbar_init(&bars[B_AGG], 4);
    mbar_init(&bars[B_RES], 4);
    mbar_init(&bars[B_M1], 1);
    mbar_init(&bars[B_M2], 1);
    mbar_init(&bars[B_M3], 1);
    mbar_init(&bars[B_H1], 8);
    mbar_init(&bars[B_H2], 8);
    mbar_init(&bars[B_OUT], 8);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  // two accumulators (tile parity): the next tile's first GEMM runs while this tile's last epilogue drains the other one
  const uint32_t tAcc0 = tmem, tH = tmem + 256, tX = tmem + 320;

  const long long n_tiles = (p.M + kRows - 1) / kRows;
  const int n_my = static_cast<int>((n_tiles > blockIdx.x) ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0);
  bool timed_out = false;
  const bool tm_on = p.timing != nullptr && blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 1 || warp == 5);
  long long* tm = reinterpret_cast<long long*>(smem + L::kTiming) + (warp == 0 ? 0 : (warp == 1 ? 8 : 16));
  if (tm_on) {
    for (int i = 0; i < 8; ++i) tm[i] = 0;
  }
  long long tlast = clock64();

  if (warp == 0) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      const uint32_t aA0 = smem_u32(bA0);
      const uint32_t aW1 = smem_u32(sW1), aW2 = smem_u32(sW2), aW3 = smem_u32(sW3);
      const uint32_t idesc = umma_idesc_bf16(128, 128, 0, 0);
      for (int it = 0; it < n_my; ++it) {
        const uint32_t par = it & 1;
        const uint32_t aA = aA0 + (it & 1) * 2 * kPB;
#define MGN_W(b, ph)             \
  if (!wait_clk(&bars[b], ph)) { \
    timed_out = true;            \
    break;                       \
  }
        const uint32_t tAcc = tAcc0 + par * 128;
        if (p.single || it == 0) {  // (otherwise this tile's first GEMM was issued during the previous tile)
          MGN_W(B_IN, par);
          if (it > 0) MGN_W(B_OUT, par ^ 1);  // single-GEMM mode: the tile two back has drained this accumulator
          MGN_T(0);
          tc_fence_after_sync();
#pragma unroll
          for (int k = 0; k < KP * 4; ++k)
            umma_ss(tAcc, umma_desc_kmajor(aA + (k >> 2) * kPB, k & 3), umma_desc_kmajor(aW1 + (k >> 2) * kPB, k & 3), idesc,
                    k != 0);
          umma_commit(&bars[B_M1]);
        }
        MGN_T(1);
        if (p.single) continue;
        MGN_W(B_H1, par);
        MGN_T(2);
        tc_fence_after_sync();
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_ts(tAcc, tH + k * 8, umma_desc_kmajor(aW2 + (k >> 2) * kPB, k & 3), idesc, k != 0);
        umma_commit(&bars[B_M2]);
        MGN_W(B_H2, par);
        MGN_T(3);
        tc_fence_after_sync();
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_ts(tAcc, tH + k * 8, umma_desc_kmajor(aW3 + (k >> 2) * kPB, k & 3), idesc, k != 0);
        umma_commit(&bars[B_M3]);
        MGN_T(4);
        if (it + 1 < n_my) {
          // next tile's first GEMM into the other accumulator (drained by E3 of the tile before this one, which every
          // epilogue warp finished before it signalled B_H2 of this tile); it overlaps this tile's E3
          MGN_W(B_IN, par ^ 1);
          tc_fence_after_sync();
          const uint32_t aAn = aA0 + ((it + 1) & 1) * 2 * kPB;
          const uint32_t tAccN = tAcc0 + (par ^ 1) * 128;
#pragma unroll
          for (int k = 0; k < KP * 4; ++k)
            umma_ss(tAccN, umma_desc_kmajor(aAn + (k >> 2) * kPB, k & 3), umma_desc_kmajor(aW1 + (k >> 2) * kPB, k & 3), idesc,
                    k != 0);
          umma_commit(&bars[B_M1]);
        }
#undef MGN_W
      }
    }
  } else if (warp <= 4) {
    // =========================== movers ===========================
    const int mt = tid - 32;
    const int rsub_m = mt >> 4;
#define MGN_W(b, ph)                                                      \
  {                                                                       \
    const bool ok_ = __all_sync(0xffffffffu, wait_clk(&bars[b], ph));     \
    if (!ok_) {                                                           \
      timed_out = true;                                                   \
      break;                                                              \
    }                                                                     \
  }
#define MGN_MOVER_SYNC() asm volatile("bar.sync 1, 128;" ::: "memory")
    int32_t r_a[16], r_g1[16], r_g2[16];
    const RowSrc& g2src = res_g2 ? p.res : p.g2;
    const bool use_g2buf = has_g2 || res_g2;
    const long long stride = static_cast<long long>(gridDim.x) * kRows;
    long long row0 = static_cast<long long>(blockIdx.x) * kRows;
    // prologue: tile 0
    if (n_my > 0) {
      if (KP == 2) {
        if (!p.tma_a) {
          fetch_row_ids(p.a.idx, row0, p.M, rsub_m, r_a);
          stage_rows_async(bA0, p.a, r_a, row0, p.M, mt);
        }
      } else {
        stage_small(bA0, p.small_x, p.small_in, p.small_is_f32, row0, p.M, mt);
      }
      if (has_g1) {
        fetch_row_ids(p.g1.idx, row0, p.M, rsub_m, r_g1);
        stage_rows_async(bG1, p.g1, r_g1, row0, p.M, mt);
      }
      if (use_g2buf) {
        fetch_row_ids(g2src.idx, row0, p.M, rsub_m, r_g2);
        stage_rows_async(bG2, g2src, r_g2, row0, p.M, mt);
      }
      cp_async_commit();
      if (n_my > 1) {  // row ids of tile 1
        if (KP == 2 && !p.tma_a) fetch_row_ids(p.a.idx, row0 + stride, p.M, rsub_m, r_a);
        if (has_g1) fetch_row_ids(p.g1.idx, row0 + stride, p.M, rsub_m, r_g1);
        if (use_g2buf) fetch_row_ids(g2src.idx, row0 + stride, p.M, rsub_m, r_g2);
      }
      cp_async_wait<0>();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[B_IN]);
      if (res_g2 && lane == 0) mbar_arrive(&bars[B_RES]);
    }
    for (int it = 0; it < n_my; ++it) {
      const uint32_t par = it & 1;
      const bool more = it + 1 < n_my;
      const long long row1 = row0 + stride;  // next tile of this CTA
      uint8_t* bAcur = bA0 + (it & 1) * 2 * kPB;
      uint8_t* bAnext = bA0 + ((it + 1) & 1) * 2 * kPB;
      // next tile's A rows stream into the other A buffer right away (its previous output was stored last iteration)
      if (more && !p.tma_a) {
        if (p.tma_out && it > 0) MGN_W(B_STD, par ^ 1);  // the loader's store of tile it - 1 has left that buffer
        if (KP == 2) stage_rows_async(bAnext, p.a, r_a, row1, p.M, mt);
        else stage_small(bAnext, p.small_x, p.small_in, p.small_is_f32, row1, p.M, mt);
      }
      MGN_T(0);
      // the additive-row buffers are free once the layer-1 epilogue has consumed them
      if (!p.single) MGN_W(B_H1, par);
      MGN_T(1);
      // h1 of this tile (left in the G1 buffer by E1) -> global, with the row mapping of the gather that follows
      if (p.h1_out != nullptr) store_rows(bG1, p.h1_out, kH, row0, p.M, mt);
      if (more && has_g1) stage_rows_async(bG1, p.g1, r_g1, row1, p.M, mt);
      if (more && use_g2buf && !res_g2) stage_rows_async(bG2, g2src, r_g2, row1, p.M, mt);
      // (single-GEMM mode has no layer-1 epilogue to order against: publishing early could run two barrier phases
      //  ahead of a waiter, which a parity wait cannot distinguish)
      const bool late_publish = p.single;
      if (more && !late_publish) {  // next tile is complete: publish it now, long before this tile's output is due
        cp_async_commit();
        cp_async_wait<0>();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[B_IN]);
      }
      // output tile (written in place over the A tile) -> global, coalesced
      MGN_W(B_OUT, par);
      MGN_T(2);
      if (more && res_g2) {  // the residual rows in bG2 were needed until now: the next tile's follow, on their own barrier
        stage_rows_async(bG2, g2src, r_g2, row1, p.M, mt);
        cp_async_commit();
        cp_async_wait<0>();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[B_RES]);
      }
      if (more && late_publish) {
        cp_async_commit();
        cp_async_wait<0>();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[B_IN]);
      }
      if (!direct_out && !p.tma_out) store_rows(bAcur, p.out, p.ld_out, row0, p.M, mt);
      if (p.seg_off != nullptr) {  // segmented sum of the result tile by destination
        const agg::TileSegs ts = agg::tile_segments_begin(row0, p.M, p.seg_off, p.seg_id, mt);
        agg::tile_segment_sum(bAcur, row0, ts, p.seg_off, p.agg, p.ld_agg, p.agg_part, p.agg_part_v, mt, p.agg_row_base,
                              p.agg_rec_base);
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[B_AGG]);
      }
      if (more && it + 2 < n_my) {  // row ids two tiles ahead of the running one
        if (KP == 2 && !p.tma_a) fetch_row_ids(p.a.idx, row1 + stride, p.M, rsub_m, r_a);
        if (has_g1) fetch_row_ids(p.g1.idx, row1 + stride, p.M, rsub_m, r_g1);
        if (use_g2buf) fetch_row_ids(g2src.idx, row1 + stride, p.M, rsub_m, r_g2);
      }
      MGN_T(3);
      MGN_MOVER_SYNC();  // all movers have finished reading bAcur before it is refilled (tile it + 2)
      MGN_T(4);
      row0 = row1;
    }
#undef MGN_W
  } else if (warp == kLoaderWarp) {
    // =========================== loader (TMA) ===========================
    if (lane == 0) {
      const long long stride = static_cast<long long>(gridDim.x) * kRows;
      long long row0 = static_cast<long long>(blockIdx.x) * kRows;
      auto load_a = [&](uint8_t* buf, long long r0) {
        if (p.tma_a) {
          mbar_arrive_expect_tx(&bars[B_IN], 2 * kPB);
          tma_load_2d(smem_u32(buf), &p.m_a, 0, static_cast<int>(r0), &bars[B_IN]);
          tma_load_2d(smem_u32(buf) + kPB, &p.m_a, 64, static_cast<int>(r0), &bars[B_IN]);
        } else {
          mbar_arrive(&bars[B_IN]);
        }
      };
      if (n_my > 0) load_a(bA0, row0);
      for (int it = 0; it < n_my; ++it) {
        const uint32_t par = it & 1;
        uint8_t* bAcur = bA0 + (it & 1) * 2 * kPB;
        uint8_t* bAnext = bA0 + ((it + 1) & 1) * 2 * kPB;
        if (it + 1 < n_my) {
          // (this tile's B_IN phase must be complete before an arrival may count for the next tile's)
          if (!wait_clk(&bars[B_IN], par)) { timed_out = true; break; }
          if (p.seg_off != nullptr && it > 0 && !wait_clk(&bars[B_AGG], par ^ 1)) { timed_out = true; break; }
          load_a(bAnext, row0 + stride);  // that buffer's previous result tile left it during the last iteration
        }
        if (!wait_clk(&bars[B_OUT], par)) { timed_out = true; break; }
        if (p.tma_out) {
          tma_store_2d(&p.m_out, smem_u32(bAcur), 0, static_cast<int>(row0));
          tma_store_2d(&p.m_out, smem_u32(bAcur) + kPB, 64, static_cast<int>(row0));
          tma_store_commit();
          tma_store_wait_read();
        }
        mbar_arrive(&bars[B_STD]);
        row0 += stride;
      }
      tma_store_wait_all();
    }
  } else {
    // =========================== epilogue (8 warps) ===========================
    const int q = warp & 3;
    const int ch = (warp - 5) >> 2;
    const int row = q * 32 + lane;
    const int c0 = ch * 64;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t t_acc0 = tAcc0 + lane_off + c0;
    const uint32_t t_h = tH + lane_off + ch * 32;
    const uint32_t t_x = tX + lane_off;
    const float* b1 = sPar + c0;
    const float* b2 = sPar + kH + c0;
    const float* b3 = sPar + 2 * kH + c0;
    const float* gam = sPar + 3 * kH + c0;
    const float* bet = sPar + 4 * kH + c0;
#define MGN_W(b, ph)                                                      \
  {                                                                       \
    const bool ok_ = __all_sync(0xffffffffu, wait_clk(&bars[b], ph));     \
    if (!ok_) {                                                           \
      timed_out = true;                                                   \
      break;                                                              \
    }                                                                     \
  }
// the two warps that share tile rows (same TMEM lane quarter) synchronise on their own named barrier
#define MGN_ROW_SYNC()                                            \
  tc_fence_before_sync();                                         \
  asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory");       \
  tc_fence_after_sync()
    for (int it = 0; it < n_my; ++it) {
      const uint32_t par = it & 1;
      const uint32_t t_acc = t_acc0 + par * 128;
      uint8_t* bAcur = bA0 + (it & 1) * 2 * kPB;
      const long long grow = (static_cast<long long>(blockIdx.x) + static_cast<long long>(it) * gridDim.x) * kRows + row;
      // ---- E1: h1 = relu(acc + b1 + G1 + G2) -> TMEM (packed bf16)
      MGN_W(B_M1, par);
      MGN_W(B_IN, par);
      MGN_T(0);
      tc_fence_after_sync();
      if (!p.single) {
#pragma unroll 1
      for (int hh = 0; hh < 2; ++hh) {
        const int cc = c0 + 32 * hh;
        uint32_t v[32];
        tmem_ld32(t_acc + 32 * hh, v);
        uint32_t ga[16], gb[16];
        if (has_g1) row_load32p(bG1, row, cc, ga);
        if (has_g2) row_load32p(bG2, row, cc, gb);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {  // two fp32 lanes per instruction (mgn_tile.cuh); relu after the bf16 rounding
          uint64_t z = f2_add(f2_packu(v[2 * j], v[2 * j + 1]), f2_ld(b1 + 32 * hh + 2 * j));
          if (has_g1) z = f2_add(z, f2_from_bf16x2(ga[j]));
          if (has_g2) z = f2_add(z, f2_from_bf16x2(gb[j]));
          pk[j] = relu_bf16x2(f2_to_bf16x2(z));
        }
        tmem_st16(t_h + 16 * hh, pk);
        if (p.h1_out != nullptr) row_store32p(bG1, row, cc, pk);  // over this thread's own consumed G1 span
      }
      tmem_st_wait();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[B_H1]);
      MGN_T(1);
      // ---- E2: h2 = relu(acc + b2) -> TMEM
      MGN_W(B_M2, par);
      MGN_T(2);
      tc_fence_after_sync();
#pragma unroll 1
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t v[32];
        tmem_ld32(t_acc + 32 * hh, v);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 16; ++j)
          pk[j] = relu_bf16x2(f2_to_bf16x2(f2_add(f2_packu(v[2 * j], v[2 * j + 1]), f2_ld(b2 + 32 * hh + 2 * j))));
        tmem_st16(t_h + 16 * hh, pk);
      }
      tmem_st_wait();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[B_H2]);
      MGN_T(3);
      // ---- E3: y = acc + b3 ; LayerNorm ; + residual ; -> output tile (in place over the A tile) or global
      MGN_W(B_M3, par);
      }  // !single
      MGN_T(4);
      tc_fence_after_sync();
      float mu = 0.f, rstd = 1.f;
      if (has_ln) {
        // one pass over the accumulator for both row sums (fp32), exchanged between the two column halves of a row
        // through spare TMEM columns; only the two warps that share the rows synchronise
        float s, ss;
        {
          uint64_t s2 = 0ull, ss2 = 0ull;
#pragma unroll 1
          for (int hh = 0; hh < 2; ++hh) {
            uint32_t v[32];
            tmem_ld32(t_acc + 32 * hh, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const uint64_t y2 = f2_add(f2_packu(v[2 * j], v[2 * j + 1]), f2_ld(b3 + 32 * hh + 2 * j));
              s2 = f2_add(s2, y2);
              ss2 = f2_fma(y2, y2, ss2);
            }
          }
          s = f2_lo(s2) + f2_hi(s2);
          ss = f2_lo(ss2) + f2_hi(ss2);
        }
        MGN_T(6);
        tmem_st2(t_x + ch * 2, __float_as_uint(s), __float_as_uint(ss));
        tmem_st_wait();
        MGN_ROW_SYNC();
        uint32_t o0, o1;
        tmem_ld2(t_x + (ch ^ 1) * 2, o0, o1);
        tmem_ld_wait();
        MGN_ROW_SYNC();  // the partner has read this tile's sums before the next tile overwrites them
        mu = (s + __uint_as_float(o0)) * (1.f / kH);
        const float var = fmaxf((ss + __uint_as_float(o1)) * (1.f / kH) - mu * mu, 0.f);
        rstd = rsqrtf(var + p.eps);
        MGN_T(7);
      }
      const uint8_t* rbuf = p.res_is_a ? bAcur : bG2;
      const bool has_res = p.res.tab != nullptr || p.res_is_a;
      if (res_g2) MGN_W(B_RES, par);
#pragma unroll 1
      for (int hh = 0; hh < 2; ++hh) {
        const int cc = c0 + 32 * hh;
        uint32_t v[32];
        tmem_ld32(t_acc + 32 * hh, v);
        uint32_t r[16];
        if (has_res) row_load32p(rbuf, row, cc, r);
        tmem_ld_wait();
        if (!direct_out) {
          uint32_t o[16];
          const uint64_t NMU = f2_splat(-mu), RS = f2_splat(rstd);
#pragma unroll
          for (int j = 0; j < 16; ++j) {  // ((y - mu) rstd) gamma + beta (+ residual), packed fp32 pairs
            const int c = 32 * hh + 2 * j;
            uint64_t y2 = f2_add(f2_packu(v[2 * j], v[2 * j + 1]), f2_ld(b3 + c));
            if (has_ln) y2 = f2_fma(f2_mul(f2_add(y2, NMU), RS), f2_ld(gam + c), f2_ld(bet + c));
            if (has_res) y2 = f2_add(y2, f2_from_bf16x2(r[j]));
            o[j] = f2_to_bf16x2(y2);
          }
          row_store32p(bAcur, row, cc, o);
        } else if (grow < p.M) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float y = __uint_as_float(v[j]) + b3[32 * hh + j];
            if (has_ln) y = (y - mu) * rstd * gam[32 * hh + j] + bet[32 * hh + j];
            if (has_res) y += (j & 1) ? bf_hi(r[j >> 1]) : bf_lo(r[j >> 1]);
            if (cc + j < p.n_out) p.out[grow * p.ld_out + cc + j] = __float2bfloat16_rn(y);
          }
        }
      }
      fence_proxy_async_smem();  // the result tile leaves through the async proxy (TMA store)
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[B_OUT]);
      MGN_T(5);
    }
#undef MGN_W
  }
  if (tm_on) {
    const int role = warp == 0 ? 0 : (warp == 1 ? 1 : 2);
    for (int i = 0; i < 8; ++i) p.timing[role * 32 + i] = tm[i];
  }
  if (timed_out && p.status != nullptr) atomicOr(p.status, 1);
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

#ifdef MGN_DEBUG_HOOKS
static long long* g_timing = nullptr;
#else
static constexpr long long* g_timing = nullptr;
#endif

template <int KP>
static int launch(Params& p, cudaStream_t st) {
  using L = Smem<KP>;
  static PerDeviceFlag configured_flag;
  bool& configured = configured_flag.get();
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(mlp3_fwd2_tc_kernel<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal);
    if (e != cudaSuccess) return static_cast<int>(e);
    configured = true;
  }
  const long long n_tiles = (p.M + kRows - 1) / kRows;
  const int grid = static_cast<int>(n_tiles < num_sms() ? n_tiles : num_sms());
  p.timing = g_timing;
  p.tma_a = KP == 2 && p.a.idx == nullptr;
  p.tma_out = p.n_out == kH;
  if (p.tma_a && tma_make_rows_map(&p.m_a, p.a.tab + p.a.col0, p.M, p.a.ld, 128) != 0) return MGN_EINVAL;
  if (p.tma_out && tma_make_rows_map(&p.m_out, p.out, p.M, p.ld_out, 128) != 0) return MGN_EINVAL;
  mlp3_fwd2_tc_kernel<KP><<<grid, kThreads, L::kTotal, MGN_ST(st)>>>(p);
  return mgn_launch_status();
}

}  // namespace fwd2
}  // namespace mgn

using namespace mgn;

#ifdef MGN_DEBUG_HOOKS
extern "C" int mgn_debug_set_fwd2_timing(void* dev_buf) {
  edge_fwd3_set_timing(static_cast<long long*>(dev_buf));  // (the edge-block kernel fills the first 6 slots)
  fwd2::g_timing = static_cast<long long*>(dev_buf);
  return MGN_OK;
}
#endif

extern "C" size_t mgn_mlp3_fwd2_agg_workspace_bytes(int64_t M) { return agg::workspace_bytes(M); }

struct AggArgs {
  const int32_t* seg_off = nullptr;
  int64_t n_seg = 0;
  void* agg = nullptr;
  int64_t ld_agg = 0;
  void* workspace = nullptr;
  size_t workspace_bytes = 0;
  // several launches over consecutive row ranges of one edge table: position of this launch's row 0, total rows of
  // the table (sizes the shared record array), index of this launch's first tile record, run the fix-up after it
  int64_t row_base = 0;
  int64_t total_tiles = 0;
  int64_t rec_base = 0;
  int fixup = 1;
  void* h1_out = nullptr;  // edge form on the third-generation kernel: keep relu(z1) for the backward pass
};

static int fwd2_run(const void* a_tab, const int32_t* a_idx, const void* small_x, int small_in,
                                int small_is_f32, const void* g1_tab, const int32_t* g1_idx, int64_t g1_ld,
                                int64_t g1_col0, const void* g2_tab, const int32_t* g2_idx, int64_t g2_ld,
                                int64_t g2_col0, const void* res_tab, int res_is_a, int64_t M, const float* w1,
                                int64_t ld_w1, const float* b1, const float* w2, const float* b2, const float* w3,
                                const float* b3, const float* gamma, const float* beta, int n_out, float eps, void* out,
                                int64_t ld_out, int* status, mgn_stream_t stream, const AggArgs& ag) {
  MGN_CHECK_ARG(M >= 0 && w1 && w2 && w3 && n_out >= 1 && n_out <= fwd2::kH && ld_out >= n_out);
  MGN_CHECK_ARG(gamma == nullptr || n_out == fwd2::kH);
  if (M == 0) return MGN_OK;
  MGN_CHECK_ARG(out != nullptr);
  fwd2::Params p{};
  p.a = tile::RowSrc{static_cast<const bf16*>(a_tab), a_idx, fwd2::kH, 0};
  p.small_x = small_x;
  p.small_in = small_in;
  p.small_is_f32 = small_is_f32;
  p.g1 = tile::RowSrc{static_cast<const bf16*>(g1_tab), g1_idx, g1_ld, g1_col0};
  p.g2 = tile::RowSrc{static_cast<const bf16*>(g2_tab), g2_idx, g2_ld, g2_col0};
  p.res = tile::RowSrc{static_cast<const bf16*>(res_tab), nullptr, fwd2::kH, 0};
  p.res_is_a = res_is_a;
  p.single = 0;
  if (g1_tab) MGN_CHECK_ARG(g1_ld % 8 == 0 && g1_col0 % 8 == 0 && (reinterpret_cast<uintptr_t>(g1_tab) & 15) == 0);
  if (g2_tab) MGN_CHECK_ARG(g1_tab && g2_ld % 8 == 0 && g2_col0 % 8 == 0 && (reinterpret_cast<uintptr_t>(g2_tab) & 15) == 0);
  // the residual either is the A tile itself or rides in the (then unused) second additive-row buffer
  MGN_CHECK_ARG(!(res_is_a && (small_in > 0 || res_tab != nullptr)));
  MGN_CHECK_ARG(!(res_tab != nullptr && g2_tab != nullptr));
  if (res_tab) MGN_CHECK_ARG((reinterpret_cast<uintptr_t>(res_tab) & 15) == 0 && n_out == fwd2::kH);
  if (n_out == fwd2::kH) MGN_CHECK_ARG(ld_out % 8 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0);
  p.M = M;
  p.w1 = w1; p.b1 = b1; p.w2 = w2; p.b2 = b2; p.w3 = w3; p.b3 = b3; p.gamma = gamma; p.beta = beta;
  p.ld_w1 = ld_w1;
  p.n_out = n_out;
  p.eps = eps;
  p.out = static_cast<bf16*>(out);
  p.ld_out = ld_out;
  p.status = status;
  cudaStream_t st = as_stream(stream);
  const long long n_tiles = (M + tile::kRows - 1) / tile::kRows;
  if (ag.seg_off != nullptr) {  // fused aggregation: rows are CSC-ordered edges, g2_idx is their destination
    MGN_CHECK_ARG(g2_idx != nullptr && ag.agg != nullptr && ag.n_seg > 0 && ag.ld_agg >= fwd2::kH && ag.ld_agg % 8 == 0 &&
                  (reinterpret_cast<uintptr_t>(ag.agg) & 15) == 0 && n_out == fwd2::kH && ag.workspace != nullptr);
    const long long tot_tiles = ag.total_tiles > 0 ? ag.total_tiles : n_tiles;
    MGN_CHECK_ARG(ag.rec_base >= 0 && ag.rec_base + n_tiles <= tot_tiles && ag.row_base >= 0);
    if (ag.workspace_bytes < static_cast<size_t>(2 * tot_tiles) * (fwd2::kH * sizeof(float) + sizeof(int32_t)))
      return MGN_EWORKSPACE;
    p.seg_off = ag.seg_off;
    p.seg_id = g2_idx;
    p.n_seg = ag.n_seg;
    p.agg = static_cast<bf16*>(ag.agg);
    p.ld_agg = ag.ld_agg;
    p.agg_part = static_cast<float*>(ag.workspace);
    p.agg_part_v = reinterpret_cast<int32_t*>(p.agg_part + 2 * tot_tiles * fwd2::kH);
    p.agg_row_base = ag.row_base;
    p.agg_rec_base = ag.rec_base;
  }
  int rc;
  // the edge block proper (gathered source / destination projections, LayerNorm, residual = input, destination sums)
  // runs on the two-tiles-in-flight kernel (mgn_edge_fwd3_tc.cu)
  const bool edge3 = ag.seg_off != nullptr && small_in <= 0 && a_tab != nullptr && a_idx == nullptr &&
                     g1_tab != nullptr && g1_idx != nullptr && g2_tab != nullptr && g2_idx != nullptr && res_is_a &&
                     gamma != nullptr && n_out == fwd2::kH && ld_out == fwd2::kH && ld_w1 >= fwd2::kH &&
                     (reinterpret_cast<uintptr_t>(a_tab) & 15) == 0;
  if (!edge3 && ag.h1_out != nullptr) {  // second-generation kernel: multi-GEMM forms with a 128-wide hidden layer
    MGN_CHECK_ARG((reinterpret_cast<uintptr_t>(ag.h1_out) & 15) == 0);
    p.h1_out = static_cast<bf16*>(ag.h1_out);
  }
  if (edge3) {
    MGN_CHECK_ARG(ag.h1_out == nullptr || (reinterpret_cast<uintptr_t>(ag.h1_out) & 15) == 0);
    fwd3::Args x{};
    x.a = static_cast<const bf16*>(a_tab);
    x.M = M;
    x.g1_tab = static_cast<const bf16*>(g1_tab);
    x.g1_idx = g1_idx;
    x.g1_ld = g1_ld;
    x.g1_col0 = g1_col0;
    x.g2_tab = static_cast<const bf16*>(g2_tab);
    x.g2_idx = g2_idx;
    x.g2_ld = g2_ld;
    x.g2_col0 = g2_col0;
    x.w1 = w1; x.b1 = b1; x.w2 = w2; x.b2 = b2; x.w3 = w3; x.b3 = b3; x.gamma = gamma; x.beta = beta;
    x.ld_w1 = ld_w1;
    x.eps = eps;
    x.out = static_cast<bf16*>(out);
    x.h1_out = static_cast<bf16*>(ag.h1_out);
    x.seg_off = p.seg_off;
    x.agg = p.agg;
    x.ld_agg = p.ld_agg;
    x.agg_part = p.agg_part;
    x.agg_part_v = p.agg_part_v;
    x.agg_row_base = p.agg_row_base;
    x.agg_rec_base = p.agg_rec_base;
    x.status = status;
    rc = edge_fwd3_launch(x, st);
  } else if (small_in > 0) {
    MGN_CHECK_ARG(small_x != nullptr && small_in <= 64 && ld_w1 >= small_in && g1_tab == nullptr);
    p.k1_true = small_in;
    rc = fwd2::launch<1>(p, st);
  } else {
    MGN_CHECK_ARG(a_tab != nullptr && ld_w1 >= fwd2::kH && (reinterpret_cast<uintptr_t>(a_tab) & 15) == 0);
    p.k1_true = fwd2::kH;
    rc = fwd2::launch<2>(p, st);
  }
  if (rc != MGN_OK || ag.seg_off == nullptr || !ag.fixup) return rc;
  const long long n_rec = 2 * (ag.total_tiles > 0 ? ag.total_tiles : n_tiles);
  agg::agg_fixup_kernel<<<static_cast<unsigned>((n_rec * 32 + 255) / 256), 256, 0, MGN_ST(st)>>>(
      p.agg_part, p.agg_part_v, n_rec, p.agg, p.ld_agg, p.n_seg);
  return mgn_launch_status();
}

extern "C" int mgn_mlp3_fwd2_tc(const void* a_tab, const int32_t* a_idx, const void* small_x, int small_in,
                                int small_is_f32, const void* g1_tab, const int32_t* g1_idx, int64_t g1_ld,
                                int64_t g1_col0, const void* g2_tab, const int32_t* g2_idx, int64_t g2_ld,
                                int64_t g2_col0, const void* res_tab, int res_is_a, int64_t M, const float* w1,
                                int64_t ld_w1, const float* b1, const float* w2, const float* b2, const float* w3,
                                const float* b3, const float* gamma, const float* beta, int n_out, float eps, void* out,
                                int64_t ld_out, int* status, mgn_stream_t stream) {
  return fwd2_run(a_tab, a_idx, small_x, small_in, small_is_f32, g1_tab, g1_idx, g1_ld, g1_col0, g2_tab, g2_idx, g2_ld, g2_col0,
                  res_tab, res_is_a, M, w1, ld_w1, b1, w2, b2, w3, b3, gamma, beta, n_out, eps, out, ld_out, status, stream,
                  AggArgs{});
}

extern "C" int mgn_edge_block_fwd_tc(const void* efeat, const void* p_src, const int32_t* src_idx, int64_t p_src_ld,
                                     int64_t p_src_col0, const void* p_dst, const int32_t* dst_idx, int64_t p_dst_ld,
                                     int64_t p_dst_col0, int64_t n_edges, const float* w1, int64_t ld_w1, const float* b1,
                                     const float* w2, const float* b2, const float* w3, const float* b3,
                                     const float* gamma, const float* beta, float eps, void* efeat_out, void* h1_out,
                                     const int32_t* csc_offsets, int64_t n_dst, void* agg, int64_t ld_agg,
                                     void* workspace, size_t workspace_bytes, int* status, mgn_stream_t stream) {
  MGN_CHECK_ARG(efeat && p_src && src_idx && p_dst && dst_idx && csc_offsets && agg && efeat_out);
  AggArgs ag;
  ag.h1_out = h1_out;
  ag.seg_off = csc_offsets;
  ag.n_seg = n_dst;
  ag.agg = agg;
  ag.ld_agg = ld_agg;
  ag.workspace = workspace;
  ag.workspace_bytes = workspace_bytes;
  return fwd2_run(efeat, nullptr, nullptr, 0, 0, p_src, src_idx, p_src_ld, p_src_col0, p_dst, dst_idx, p_dst_ld, p_dst_col0,
                  nullptr, 1, n_edges, w1, ld_w1, b1, w2, b2, w3, b3, gamma, beta, fwd2::kH, eps, efeat_out, fwd2::kH, status,
                  stream, ag);
}

// One of several launches over consecutive row ranges [row_base, row_base + n_rows) of one CSC-ordered edge table
// (the partitioned path runs the interior edges while the halo exchange is in flight, then the boundary runs):
// pointers / indices are those of the range, total_tiles = sum over the launches of ceil(rows / 128), rec_base = that sum
// over the ranges before this one.  mgn_agg_fixup finishes the destination sums once every range has run.
extern "C" int mgn_edge_block_fwd_part_tc(const void* efeat, const void* p_src, const int32_t* src_idx, int64_t p_src_ld,
                                          int64_t p_src_col0, const void* p_dst, const int32_t* dst_idx, int64_t p_dst_ld,
                                          int64_t p_dst_col0, int64_t n_rows, const float* w1, int64_t ld_w1,
                                          const float* b1, const float* w2, const float* b2, const float* w3,
                                          const float* b3, const float* gamma, const float* beta, float eps,
                                          void* efeat_out, void* h1_out, const int32_t* csc_offsets, int64_t n_dst, void* agg,
                                          int64_t ld_agg, void* workspace, size_t workspace_bytes, int64_t row_base,
                                          int64_t total_tiles, int64_t rec_base, int* status, mgn_stream_t stream) {
  if (n_rows == 0) return MGN_OK;
  MGN_CHECK_ARG(efeat && p_src && src_idx && p_dst && dst_idx && csc_offsets && agg && efeat_out && total_tiles > 0);
  AggArgs ag;
  ag.h1_out = h1_out;
  ag.seg_off = csc_offsets;
  ag.n_seg = n_dst;
  ag.agg = agg;
  ag.ld_agg = ld_agg;
  ag.workspace = workspace;
  ag.workspace_bytes = workspace_bytes;
  ag.row_base = row_base;
  ag.total_tiles = total_tiles;
  ag.rec_base = rec_base;
  ag.fixup = 0;
  return fwd2_run(efeat, nullptr, nullptr, 0, 0, p_src, src_idx, p_src_ld, p_src_col0, p_dst, dst_idx, p_dst_ld, p_dst_col0,
                  nullptr, 1, n_rows, w1, ld_w1, b1, w2, b2, w3, b3, gamma, beta, fwd2::kH, eps, efeat_out, fwd2::kH, status,
                  stream, ag);
}

/* MeshNodeBlock.forward (mesh_node_block.py:82-92) with the first Linear split per input block:
 *   nfeat_out[v] = nfeat[v] + LN(MLP(agg[v] W1a^T + P_node[v] + b1)),  P_node = columns [p_col0, p_col0+128) of p_tab;
 * h1_out (nullable) receives relu(z1) for mgn_edge_block_bwd_tc (add_gout = 0). */
extern "C" int mgn_node_block_fwd_tc(const void* agg, const void* p_tab, int64_t p_ld, int64_t p_col0, const void* nfeat,
                                     int64_t n_nodes, const float* w1, int64_t ld_w1, const float* b1, const float* w2,
                                     const float* b2, const float* w3, const float* b3, const float* gamma,
                                     const float* beta, float eps, void* nfeat_out, void* h1_out, int* status,
                                     mgn_stream_t stream) {
  MGN_CHECK_ARG(agg && p_tab && nfeat && nfeat_out);
#ifndef MGN_NODE_FWD2
  // two-tiles-in-flight kernel, node form (mgn_edge_fwd3_tc.cu, kNode); -DMGN_NODE_FWD2 keeps the second-generation kernel
  const auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  if (n_nodes > 0 && gamma != nullptr && ld_w1 >= fwd2::kH && p_ld % 8 == 0 && p_col0 % 8 == 0 && al16(agg) && al16(p_tab) &&
      al16(nfeat) && al16(nfeat_out) && al16(h1_out) && w1 && w2 && w3) {
    fwd3::Args x{};
    x.a = static_cast<const bf16*>(agg);
    x.M = n_nodes;
    x.g2_tab = static_cast<const bf16*>(p_tab);
    x.g2_ld = p_ld;
    x.g2_col0 = p_col0;
    x.w1 = w1; x.b1 = b1; x.w2 = w2; x.b2 = b2; x.w3 = w3; x.b3 = b3; x.gamma = gamma; x.beta = beta;
    x.ld_w1 = ld_w1;
    x.eps = eps;
    x.res = static_cast<const bf16*>(nfeat);
    x.out = static_cast<bf16*>(nfeat_out);
    x.h1_out = static_cast<bf16*>(h1_out);
    x.status = status;
    return edge_fwd3_launch(x, as_stream(stream));
  }
#endif
  AggArgs ag;
  ag.h1_out = h1_out;
  return fwd2_run(agg, nullptr, nullptr, 0, 0, p_tab, nullptr, p_ld, p_col0, nullptr, nullptr, 0, 0, nfeat, 0, n_nodes, w1,
                  ld_w1, b1, w2, b2, w3, b3, gamma, beta, fwd2::kH, eps, nfeat_out, fwd2::kH, status, stream, ag);
}

extern "C" int mgn_agg_fixup(void* workspace, int64_t total_tiles, void* agg, int64_t ld_agg, int64_t n_dst,
                             mgn_stream_t stream) {
  MGN_CHECK_ARG(workspace && agg && total_tiles > 0 && n_dst > 0 && ld_agg >= fwd2::kH);
  float* part = static_cast<float*>(workspace);
  int32_t* part_v = reinterpret_cast<int32_t*>(part + 2 * total_tiles * fwd2::kH);
  const long long n_rec = 2 * total_tiles;
  agg::agg_fixup_kernel<<<static_cast<unsigned>((n_rec * 32 + 255) / 256), 256, 0, MGN_ST(as_stream(stream))>>>(
      part, part_v, n_rec, static_cast<bf16*>(agg), ld_agg, n_dst);
  return mgn_launch_status();
}

// out[M,128] (row stride ld_out) = x[M,128] (row stride ld_x) W^T + bias (+ residual[M,128]); W [128, >=128] fp32
// with row stride ld_w.  Node-level projections of the fused path (P = nfeat Wp^T, g_nfeat += T Wp).
extern "C" int mgn_linear128_tc(const void* x, int64_t ld_x, int64_t M, const float* w, int64_t ld_w, const float* bias,
                                const void* residual, void* out, int64_t ld_out, int* status, mgn_stream_t stream) {
  MGN_CHECK_ARG(M >= 0 && w && ld_w >= fwd2::kH && ld_x >= fwd2::kH && ld_out >= fwd2::kH);
  if (M == 0) return MGN_OK;
  MGN_CHECK_ARG(x && out && ld_x % 8 == 0 && ld_out % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
                (reinterpret_cast<uintptr_t>(out) & 15) == 0);
  if (residual) MGN_CHECK_ARG((reinterpret_cast<uintptr_t>(residual) & 15) == 0);
  fwd2::Params p{};
  p.a = tile::RowSrc{static_cast<const bf16*>(x), nullptr, ld_x, 0};
  p.res = tile::RowSrc{static_cast<const bf16*>(residual), nullptr, fwd2::kH, 0};
  p.res_is_a = 0;
  p.single = 1;
  p.M = M;
  p.w1 = w;
  p.ld_w1 = ld_w;
  p.k1_true = fwd2::kH;
  // W2 / W3 are not used by the single-GEMM mode; the staging code still reads 128x128 floats from them
  p.w2 = w;
  p.w3 = w;
  p.b3 = bias;
  p.n_out = fwd2::kH;
  p.out = static_cast<bf16*>(out);
  p.ld_out = ld_out;
  p.status = status;
  return fwd2::launch<2>(p, as_stream(stream));
}
