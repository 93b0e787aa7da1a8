// mgn_mlp_bwd_tc.cu — fused MeshGraphMLP backward on tcgen05 / TMEM (bf16 storage, hidden 128, ReLU).
//
// One persistent CTA per SM.  For every 128-row tile the kernel RECOMPUTES the forward hidden
// activations (nothing but the layer inputs is kept from the forward pass) and then walks the
// chain backwards, all GEMMs on the tensor cores with operands staged in shared memory:
//
//   fwd recompute   z1 = A W1^T + G + b1 ; h1 = relu(z1)        (G = optional additive gathered rows)
//                   h2 = relu(h1 W2^T + b2) ; y = h2 W3^T + b3
//   LayerNorm bwd   g_y  = rstd * (ghat - mean(ghat) - xhat * mean(ghat * xhat)),  ghat = g_out * gamma
//   layer 3         gW3 += g_y^T  h2      g_h2 = g_y  W3     g_z2 = g_h2 * (h2 > 0)
//   layer 2         gW2 += g_z2^T h1      g_h1 = g_z2 W2     g_z1 = g_h1 * (h1 > 0)
//   layer 1         gW1 += g_z1^T A       g_A  = g_z1 W1 (+ g_out for the residual connection)
//
// The three weight-gradient accumulators stay resident in TMEM (3 x 128 columns) for the whole life
// of the CTA and are written once, as per-CTA fp32 partials that a second kernel sums in a fixed
// order (deterministic, no atomics).  One smem copy of each operand serves every GEMM that touches
// it: the same swizzled [rows][64]-bf16 panels are read K-major (forward / dgrad A operand) and
// MN-major (wgrad operands, dgrad B operand) -- layouts verified on hardware by tools/probe_tc.cu.
//
// This is the backward of MeshEdgeBlock / MeshNodeBlock / the encoder+decoder MeshGraphMLPs of the
// reference (physicsnemo/models/gnn_layers/mesh_graph_mlp.py:142-203, mesh_edge_block.py:88-96,
// mesh_node_block.py:82-92), which the reference leaves to autograd over cuBLAS + ATen kernels.
//
// Warp roles (288 threads): warp 0 = MMA issuer (one lane) + TMEM owner; warps 1-4 = movers
// (stage tiles global->smem with 128-bit coalesced loads incl. the row gathers, write result
// tiles smem->global coalesced, bias-gradient column sums); warps 5-8 = epilogue (thread = tile
// row = TMEM lane).
#include "mgn_common.cuh"
#include "mgn_tc.cuh"
#include "mgn_reduce.cuh"
#include "mgn_tile.cuh"
#include "mgn_tma.cuh"

namespace mgn {

namespace bwd {

using namespace tile;
constexpr int kEpiWarps = 8;                        // two per TMEM lane quarter, 64 columns each
constexpr int kLoaderWarp = 5 + kEpiWarps;            // warp 13: TMA loads / stores
constexpr int kThreads = 32 * (kLoaderWarp + 1);     // warp 0: MMA, warps 1-4: reducers, warps 5-12: epilogue
constexpr int kH = 128;

struct Params {
  RowSrc a;             // layer-1 input rows [*,128]                       (KP == 2)
  const void* small_x;  // raw [M, small_in] features zero-padded to K=64    (KP == 1)
  int small_in;
  int small_is_f32;
  RowSrc g1, g2;        // additive rows of layer 1 (tab == nullptr: absent)
  RowSrc go1, go2;      // incoming gradient rows, summed (go2 optional)
  int go_small;         // go1 is a dense [M, n_out] bf16 matrix with n_out < 128
  long long M;
  const float *w1, *b1, *w2, *b2, *w3, *b3, *gamma;
  long long ld_w1;
  int k1_true;
  int n_out;
  float eps;
  bf16* g_a;            // [M,128] gradient w.r.t. the layer-1 input rows (nullable)
  int add_gout;         // g_a += g_out (residual connection on the A rows)
  bf16* g_z1;           // [M,128] gradient w.r.t. the layer-1 pre-activation (nullable), row stride g_z1_ld
  long long g_z1_ld;
  float* partials;      // [gridDim.x][part_floats]
  long long part_floats;
  int* status;
  long long* timing;    // debug: [3 roles][32] cycle counters of CTA 0 (nullable)
  // tensor maps (mgn_tma.cuh): a row source without idx uses a {64 x 128} box map, one with idx a {64 x 1} gather map
  // whose row extent is kOobRow, so that the index kOobRow (rows past M) reads as zeros
  alignas(64) CUtensorMap m_a, m_g1, m_g2, m_go1, m_go2, m_ga, m_gz1;
};
constexpr int kOobRow = 1 << 30;

enum { kStatusTimeout = 1, kStatusSmem = 2 };

#define MGN_T(i)                      \
  if (tm_on) {                        \
    const long long t_ = clock64();   \
    tm[i] += t_ - tlast;              \
    tlast = t_;                       \
  }  // tm lives in shared memory: the counters must not cost registers

// barriers.  B_A / B_G: layer-1 input rows / additive rows of a tile are staged (published one tile ahead);
// B_GO: incoming gradient rows staged; B_A2: layer-1 input rows staged again for the weight gradient;
// B_MMA1 + k: k-th group of MMAs of the tile has completed; B_E1 + k: k-th epilogue phase has completed.
// B_CS + k: the reducer warps have finished the k-th column-sum pass (its buffers may be overwritten).
// B_ST: the g_z1 result tile has left the H1 buffer.  B_W3 / B_W2: the weight-gradient MMAs of layer 3 / 2 have
// completed (B_MMA1 + 3 / + 4 fire as soon as the data-gradient MMAs issued before them have).
enum { B_A = 0, B_G = 1, B_GO = 2, B_A2 = 3, B_MMA1 = 4, B_E1 = 11, B_CS = 17, B_ST = 20, B_W3 = 21, B_W2 = 22, B_NUM = 23 };

// loader warp: stage one 128-row tile (two panels) of a row source -- two box loads (lane 0) or 64 gather4 loads
// (lane l: rows 4l .. 4l+3, both panels); 2 * kPB bytes complete on `bar` either way
__device__ __forceinline__ void tma_stage_rows(uint8_t* buf, const CUtensorMap* map, const int32_t* __restrict__ idx,
                                               long long row0, long long M, uint64_t* bar, int lane) {
  const uint32_t dst = smem_u32(buf);
  if (idx == nullptr) {
    if (lane == 0) {
      tma_load_2d(dst, map, 0, static_cast<int>(row0), bar);
      tma_load_2d(dst + kPB, map, 64, static_cast<int>(row0), bar);
    }
  } else {
    int r[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long grow = row0 + 4 * lane + j;
      r[j] = grow < M ? __ldg(idx + grow) : kOobRow;
    }
    tma_gather4(dst + 4 * lane * 128, map, 0, r[0], r[1], r[2], r[3], bar);
    tma_gather4(dst + kPB + 4 * lane * 128, map, 64, r[0], r[1], r[2], r[3], bar);
  }
}

// Four 32 KB tile buffers.  Buffer 0 always holds the layer-1 input (A -> go2 rows -> g_y -> A again -> next tile's
// A); the other three rotate roles from tile to tile so that the next tile's gathered rows stream in while this tile
// is still in its backward half: the next tile's G1 takes over this tile's H2 buffer (free after the layer-2 MMAs),
// its G2 this tile's H1 buffer (free after the layer-1 MMAs) and its H1 this tile's X buffer (once g_A has left).
enum { R_A = 0, R_X = 1, R_H1 = 2, R_H2 = 3 };

template <int KP>
struct Smem {
  static constexpr int kW1 = 0;
  static constexpr int kW2 = KP * kPB;
  static constexpr int kW3 = kW2 + 2 * kPB;
  static constexpr int kBuf = kW3 + 2 * kPB;   // 4 tile buffers x 2 panels
  static constexpr int kPar = kBuf + 8 * kPB;  // b1, b2, b3, gamma
  static constexpr int kBars = kPar + 4 * kH * 4;
  static constexpr int kTmemSlot = kBars + 24 * 8;
  static constexpr int kTiming = kTmemSlot + 16;  // 3 roles x 16 x int64 (debug)
  static constexpr int kTotal = kTiming + 3 * 16 * 8;
};

// per-CTA partial layout (floats)
template <int KP>
struct Part {
  static constexpr int kW1 = 0;
  static constexpr int kW2 = kH * 64 * KP;
  static constexpr int kW3 = kW2 + kH * kH;
  static constexpr int kB1 = kW3 + kH * kH;
  static constexpr int kB2 = kB1 + kH;
  static constexpr int kB3 = kB2 + kH;
  static constexpr int kGamma = kB3 + kH;
  static constexpr int kBeta = kGamma + kH;
  static constexpr int kTotal = kBeta + kH;
};

__device__ __forceinline__ bool bf_pos_lo(uint32_t w) { return static_cast<int32_t>(w << 16) > 0; }
__device__ __forceinline__ bool bf_pos_hi(uint32_t w) { return static_cast<int32_t>(w & 0xFFFF0000u) > 0; }

template <int KP>
__global__ void __launch_bounds__(kThreads, 1) mlp3_bwd_tc_kernel(const __grid_constant__ Params p) {
  using L = Smem<KP>;
  using PT = Part<KP>;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if ((smem_u32(smem) & 1023u) != 0) {  // uniform: the swizzled panels need 1024-byte alignment
    if (tid == 0 && p.status) atomicOr(p.status, kStatusSmem);
    return;
  }
  uint8_t* sW1 = smem + L::kW1;
  uint8_t* sW2 = smem + L::kW2;
  uint8_t* sW3 = smem + L::kW3;
  uint8_t* buf0 = smem + L::kBuf;
  float* sPar = reinterpret_cast<float*>(smem + L::kPar);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kBars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::kTmemSlot);
#define MGN_BUF(role, it) (buf0 + ((role) == R_A ? 0 : 1 + (((role) - 1 + 2 * ((it) % 3)) % 3)) * (2 * kPB))

  const bool has_ln = p.gamma != nullptr;
  const bool has_g = p.g1.tab != nullptr;
  const bool has_g2 = p.g2.tab != nullptr;
  const bool has_go2 = p.go2.tab != nullptr;
  const bool need_ga = p.g_a != nullptr;
  constexpr int N1 = 64 * KP;  // width of the layer-1 input

  // ---------------- one-time setup ----------------
  stage_weight_ld(sW1, p.w1, p.ld_w1, kH, p.k1_true, KP, tid, kThreads);
  stage_weight_ld(sW2, p.w2, kH, kH, kH, 2, tid, kThreads);
  stage_weight_ld(sW3, p.w3, kH, p.n_out, kH, 2, tid, kThreads);
  for (int i = tid; i < kH; i += kThreads) {
    sPar[i] = p.b1 ? p.b1[i] : 0.f;
    sPar[kH + i] = p.b2 ? p.b2[i] : 0.f;
    sPar[2 * kH + i] = (p.b3 && i < p.n_out) ? p.b3[i] : 0.f;
    sPar[3 * kH + i] = has_ln ? p.gamma[i] : 1.f;
  }
  if (tid == 0) {
    // B_A / B_A2: one arrive.expect_tx of the loader (TMA) or the four reducer warps (KP == 1: plain stores)
    mbar_init(&bars[B_A], KP == 2 ? 1 : 4);
    mbar_init(&bars[B_A2], KP == 2 ? 1 : 4);
    // B_GO / B_G: the four reducer warps (cp.async row gathers) + the loader (TMA part / result tile has left)
    mbar_init(&bars[B_GO], 5);
    mbar_init(&bars[B_G], 5);
    for (int b = B_MMA1; b < B_ST; ++b) mbar_init(&bars[b], b < B_E1 ? 1 : (b < B_CS ? kEpiWarps : 4));
    mbar_init(&bars[B_ST], 1);
    mbar_init(&bars[B_W3], 1);
    mbar_init(&bars[B_W2], 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tAcc = tmem, tW1 = tmem + 128, tW2 = tmem + 256, tW3 = tmem + 384;

  const long long n_tiles = (p.M + kRows - 1) / kRows;
  const int n_my = static_cast<int>((n_tiles > blockIdx.x) ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0);
  bool timed_out = false;
  const bool tm_on = p.timing != nullptr && blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == kLoaderWarp || warp == 5);
  long long* tm = reinterpret_cast<long long*>(smem + L::kTiming) + (warp == 0 ? 0 : (warp == kLoaderWarp ? 16 : 32));
  if (tm_on) {
    for (int i = 0; i < 16; ++i) tm[i] = 0;
  }
  long long tlast = clock64();

  // per-CTA reduction scratch (bias / beta column sums of the movers, gamma sums of the epilogue): the H2 buffer
  // of this CTA's last tile, which nobody touches after that tile's layer-2 MMAs
  float* scratch = reinterpret_cast<float*>(MGN_BUF(R_H2, n_my > 0 ? n_my - 1 : 0));

  if (warp == 0) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      const uint32_t a0 = smem_u32(buf0);
      const uint32_t aW1 = smem_u32(sW1), aW2 = smem_u32(sW2), aW3 = smem_u32(sW3);
      const uint32_t id_nt = umma_idesc_bf16(128, 128, 0, 0);   // D = A(K-major) * B(K-major)^T
      const uint32_t id_tn = umma_idesc_bf16(128, 128, 1, 1);   // D = A(MN)^T * B(MN)          (wgrad)
      const uint32_t id_nn = umma_idesc_bf16(128, 128, 0, 1);   // D = A(K-major) * B(MN)       (dgrad)
      const uint32_t id_tn1 = umma_idesc_bf16(128, N1, 1, 1);
      const uint32_t id_nn1 = umma_idesc_bf16(128, N1, 0, 1);
      for (int it = 0; it < n_my; ++it) {
        const uint32_t par = it & 1;
        const uint32_t aA = a0;
        const uint32_t aH1 = smem_u32(MGN_BUF(R_H1, it));
        const uint32_t aH2 = smem_u32(MGN_BUF(R_H2, it));
#define MGN_W(b, ph)                         \
  if (!wait_clk(&bars[b], ph)) {             \
    timed_out = true;                        \
    break;                                   \
  }
        // ---- GEMM1: acc = A W1^T   (A was staged during the previous tile)
        MGN_W(B_A, par);
        if (it > 0) MGN_W(B_E1 + 5, par ^ 1);  // the previous tile's last epilogue has drained the accumulator
        MGN_T(0);
        tc_fence_after_sync();
#pragma unroll
        for (int k = 0; k < KP * 4; ++k)
          umma_ss(tAcc, umma_desc_kmajor(aA + (k >> 2) * kPB, k & 3), umma_desc_kmajor(aW1 + (k >> 2) * kPB, k & 3),
                  id_nt, k != 0);
        umma_commit(&bars[B_MMA1 + 0]);
        MGN_T(1);
        // ---- GEMM2: acc = h1 W2^T
        MGN_W(B_E1 + 0, par);
        MGN_T(2);
        tc_fence_after_sync();
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_ss(tAcc, umma_desc_kmajor(aH1 + (k >> 2) * kPB, k & 3), umma_desc_kmajor(aW2 + (k >> 2) * kPB, k & 3),
                  id_nt, k != 0);
        umma_commit(&bars[B_MMA1 + 1]);
        MGN_T(3);
        // ---- GEMM3: acc = h2 W3^T
        MGN_W(B_E1 + 1, par);
        MGN_T(4);
        tc_fence_after_sync();
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_ss(tAcc, umma_desc_kmajor(aH2 + (k >> 2) * kPB, k & 3), umma_desc_kmajor(aW3 + (k >> 2) * kPB, k & 3),
                  id_nt, k != 0);
        umma_commit(&bars[B_MMA1 + 2]);
        MGN_T(5);
        // ---- layer 3: gW3 += g_y^T h2 ; acc = g_y W3          (g_y in the A buffer)
        MGN_W(B_E1 + 2, par);
        MGN_T(6);
        tc_fence_after_sync();
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_ss(tAcc, umma_desc_kmajor(aA + (k >> 2) * kPB, k & 3), umma_desc_mnmajor(aW3, k, kPB), id_nn, k != 0);
        umma_commit(&bars[B_MMA1 + 3]);  // E4 may read the accumulator while the weight-gradient MMAs run
#pragma unroll
        for (int j = 0; j < 8; ++j)
          umma_ss(tW3, umma_desc_mnmajor(aA, j, kPB), umma_desc_mnmajor(aH2, j, kPB), id_tn, (it | j) != 0);
        umma_commit(&bars[B_W3]);
        MGN_T(7);
        // ---- layer 2: gW2 += g_z2^T h1 ; acc = g_z2 W2        (g_z2 in the H2 buffer)
        MGN_W(B_E1 + 3, par);
        MGN_T(8);
        tc_fence_after_sync();
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_ss(tAcc, umma_desc_kmajor(aH2 + (k >> 2) * kPB, k & 3), umma_desc_mnmajor(aW2, k, kPB), id_nn, k != 0);
        umma_commit(&bars[B_MMA1 + 4]);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          umma_ss(tW2, umma_desc_mnmajor(aH2, j, kPB), umma_desc_mnmajor(aH1, j, kPB), id_tn, (it | j) != 0);
        umma_commit(&bars[B_W2]);
        MGN_T(9);
        // ---- layer 1: acc = g_z1 W1 (epilogue may start on it at once) ; gW1 += g_z1^T A
        //      (g_z1 in the H1 buffer, A re-staged in the A buffer)
        MGN_W(B_E1 + 4, par);
        MGN_T(10);
        // (waiting for the re-staged A here too orders every mover's column sums of X before E6 rewrites X)
        MGN_W(B_A2, par);
        tc_fence_after_sync();
        if (need_ga) {
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma_ss(tAcc, umma_desc_kmajor(aH1 + (k >> 2) * kPB, k & 3), umma_desc_mnmajor(aW1, k, kPB), id_nn1,
                    k != 0);
        }
        umma_commit(&bars[B_MMA1 + 5]);
        MGN_T(11);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          umma_ss(tW1, umma_desc_mnmajor(aH1, j, kPB), umma_desc_mnmajor(aA, j, kPB), id_tn1, (it | j) != 0);
        umma_commit(&bars[B_MMA1 + 6]);
        MGN_T(12);
#undef MGN_W
      }
    }
  } else if (warp <= 4) {
    // =========================== reducers ===========================
    // bias / beta gradients = column sums of the gradient tiles, read from shared memory while the main chain runs;
    // with raw small-width inputs (KP == 1) or a narrow incoming gradient they also stage those by plain stores
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
#define MGN_MOVER_SYNC() asm volatile("bar.sync 1, 128;" ::: "memory")
    // running column sums (fixed columns per thread)
    float cs_b1[8] = {0, 0, 0, 0, 0, 0, 0, 0}, cs_b2[8] = {0, 0, 0, 0, 0, 0, 0, 0}, cs_b3[8] = {0, 0, 0, 0, 0, 0, 0, 0},
          cs_beta[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const int rsub_m = mt >> 4;
    const bool go2_shares_g2 = has_go2 && p.go2.idx == p.g2.idx && p.go2.idx != nullptr;
    const bool go1_gathered = !p.go_small && p.go1.idx != nullptr;
    int32_t r_g1[16], r_g2[16], r_tmp[16];
    if (n_my > 0) {  // prologue: tile 0
      const long long row00 = static_cast<long long>(blockIdx.x) * kRows;
      if (KP == 1) {
        stage_small(MGN_BUF(R_A, 0), p.small_x, p.small_in, p.small_is_f32, row00, p.M, mt);
        MGN_PUBLISH(B_A);
      }
      fetch_row_ids(has_g ? p.g1.idx : nullptr, row00, p.M, rsub_m, r_g1);
      fetch_row_ids(has_g2 ? p.g2.idx : (has_go2 ? p.go2.idx : nullptr), row00, p.M, rsub_m, r_g2);
      if (has_g) stage_rows_async(MGN_BUF(R_X, 0), p.g1, r_g1, row00, p.M, mt);
      if (has_g2) stage_rows_async(MGN_BUF(R_H2, 0), p.g2, r_g2, row00, p.M, mt);
      cp_async_commit();
      cp_async_wait<0>();
      MGN_PUBLISH(B_G);
    }
    for (int it = 0; it < n_my; ++it) {
      const uint32_t par = it & 1;
      const bool more = it + 1 < n_my;
      const long long row0 = (static_cast<long long>(blockIdx.x) + static_cast<long long>(it) * gridDim.x) * kRows;
      const long long row0n = row0 + static_cast<long long>(gridDim.x) * kRows;  // next tile of this CTA
      uint8_t* bA = MGN_BUF(R_A, it);
      uint8_t* bX = MGN_BUF(R_X, it);
      uint8_t* bH1 = MGN_BUF(R_H1, it);
      uint8_t* bH2 = MGN_BUF(R_H2, it);
      if (more && has_g) fetch_row_ids(p.g1.idx, row0n, p.M, rsub_m, r_g1);  // used at the end of this tile
      MGN_W(B_E1 + 0, par);
      if (!p.go_small) {  // gathered parts of the incoming gradient (go1 by rows -> X, go2 -> A); summed in E3
        if (go1_gathered) {
          fetch_row_ids(p.go1.idx, row0, p.M, rsub_m, r_tmp);
          stage_rows_async(bX, p.go1, r_tmp, row0, p.M, mt);
        }
        if (has_go2) {
          if (!go2_shares_g2) fetch_row_ids(p.go2.idx, row0, p.M, rsub_m, r_g2);
          stage_rows_async(bA, p.go2, r_g2, row0, p.M, mt);
        }
        cp_async_commit();
        if (more && (has_g2 || go2_shares_g2))
          fetch_row_ids(has_g2 ? p.g2.idx : p.go2.idx, row0n, p.M, rsub_m, r_g2);  // used at the end of this tile
        cp_async_wait<0>();
        MGN_PUBLISH(B_GO);
      } else {  // narrow incoming gradient (decoder): zero-padded to 128 columns -> X
        const int chunk = mt & 15, rsub = mt >> 4;
        for (int i = 0; i < 16; ++i) {
          const int row = i * 8 + rsub;
          const long long grow = row0 + row;
          float f[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int col = chunk * 8 + q;
            f[q] = (grow < p.M && col < p.n_out) ? __bfloat162float(p.go1.tab[grow * p.go1.ld + col]) : 0.f;
          }
          uint4 v4;
          v4.x = pack_bf16x2(f[0], f[1]);
          v4.y = pack_bf16x2(f[2], f[3]);
          v4.z = pack_bf16x2(f[4], f[5]);
          v4.w = pack_bf16x2(f[6], f[7]);
          *reinterpret_cast<uint4*>(bX + (chunk >> 3) * kPB + sw128_offset(row, chunk & 7)) = v4;
        }
        MGN_PUBLISH(B_GO);
      }
      // after E3: X = g_out (summed), A = g_y
      MGN_W(B_E1 + 2, par);
      colsum_tile(bX, mt, cs_beta);
      colsum_tile(bA, mt, cs_b3);
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[B_CS + 0]);
      if (KP == 1) {  // re-stage the layer-1 input once the layer-3 MMAs have consumed g_y
        MGN_W(B_W3, par);
        MGN_MOVER_SYNC();
        stage_small(bA, p.small_x, p.small_in, p.small_is_f32, row0, p.M, mt);
        MGN_PUBLISH(B_A2);
      }
      MGN_W(B_E1 + 3, par);
      colsum_tile(bH2, mt, cs_b2);
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[B_CS + 1]);
      if (more && has_g) {  // next tile's G1 rows -> this tile's H2 buffer (free after the layer-2 MMAs)
        MGN_W(B_W2, par);
        stage_rows_async(bH2, p.g1, r_g1, row0n, p.M, mt);
        cp_async_commit();
      }
      MGN_W(B_E1 + 4, par);
      colsum_tile(bH1, mt, cs_b1);
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[B_CS + 2]);
      if (more) {  // next tile's G2 rows -> this tile's H1 buffer (free after the layer-1 MMAs and the g_z1 store)
        MGN_W(B_MMA1 + 6, par);
        if (KP == 1) {  // ... and its raw input -> the A buffer
          MGN_MOVER_SYNC();
          stage_small(bA, p.small_x, p.small_in, p.small_is_f32, row0n, p.M, mt);
          MGN_PUBLISH(B_A);
        }
        MGN_W(B_ST, par);
        if (has_g2) stage_rows_async(bH1, p.g2, r_g2, row0n, p.M, mt);
        cp_async_commit();
        cp_async_wait<0>();
        MGN_PUBLISH(B_G);
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
  } else if (warp == kLoaderWarp) {
    // =========================== loader (TMA) ===========================
    // every bulk tensor load / gather / store of the CTA.  A tile's rows are requested as early as their buffer is
    // free: the next tile's A right after this tile's layer-2 MMAs, its additive rows after the layer-1 MMAs.
#define MGN_W(b, ph)                                                      \
  {                                                                       \
    const bool ok_ = __all_sync(0xffffffffu, wait_clk(&bars[b], ph));     \
    if (!ok_) {                                                           \
      timed_out = true;                                                   \
      break;                                                              \
    }                                                                     \
  }
    const bool go1_tma = !p.go_small && p.go1.idx == nullptr;
    if (n_my > 0) {  // prologue: tile 0
      const long long row00 = static_cast<long long>(blockIdx.x) * kRows;
      if (lane == 0) {
        if (KP == 2) mbar_arrive_expect_tx(&bars[B_A], 2 * kPB);
        mbar_arrive(&bars[B_G]);  // (no earlier result tile to wait for)
      }
      __syncwarp();
      if (KP == 2) tma_stage_rows(MGN_BUF(R_A, 0), &p.m_a, nullptr, row00, p.M, &bars[B_A], lane);
    }
    for (int it = 0; it < n_my; ++it) {
      const uint32_t par = it & 1;
      const bool more = it + 1 < n_my;
      const long long row0 = (static_cast<long long>(blockIdx.x) + static_cast<long long>(it) * gridDim.x) * kRows;
      const long long row0n = row0 + static_cast<long long>(gridDim.x) * kRows;  // next tile of this CTA
      uint8_t* bA = MGN_BUF(R_A, it);
      uint8_t* bX = MGN_BUF(R_X, it);
      uint8_t* bH1 = MGN_BUF(R_H1, it);
      MGN_T(0);
      // dense incoming gradient once E1 has consumed the additive rows: go1 -> X
      MGN_W(B_E1 + 0, par);
      MGN_T(1);
      if (lane == 0) {
        if (go1_tma) mbar_arrive_expect_tx(&bars[B_GO], 2 * kPB);
        else mbar_arrive(&bars[B_GO]);
      }
      __syncwarp();
      if (go1_tma) tma_stage_rows(bX, &p.m_go1, nullptr, row0, p.M, &bars[B_GO], lane);
      MGN_T(2);
      if (KP == 2) {
        // layer-1 input again (for its weight gradient) once the layer-3 MMAs and the column sums are done with g_y
        MGN_W(B_W3, par);
        MGN_W(B_CS + 0, par);
        MGN_T(3);
        if (lane == 0) mbar_arrive_expect_tx(&bars[B_A2], 2 * kPB);
        __syncwarp();
        tma_stage_rows(bA, &p.m_a, nullptr, row0, p.M, &bars[B_A2], lane);
        MGN_T(4);
      }
      // g_z1 tile (H1) -> global
      MGN_W(B_E1 + 4, par);
      MGN_T(7);
      if (lane == 0) {
        if (p.g_z1 != nullptr) {
          tma_store_2d(&p.m_gz1, smem_u32(bH1), 0, static_cast<int>(row0));
          tma_store_2d(&p.m_gz1, smem_u32(bH1) + kPB, 64, static_cast<int>(row0));
          tma_store_commit();
          tma_store_wait_read();
        }
        mbar_arrive(&bars[B_ST]);
      }
      __syncwarp();
      MGN_T(5);
      if (KP == 2 && more) {  // next tile's layer-1 input -> the A buffer, as soon as the layer-1 MMAs are done with it
        MGN_W(B_MMA1 + 6, par);
        MGN_T(6);
        if (lane == 0) mbar_arrive_expect_tx(&bars[B_A], 2 * kPB);
        __syncwarp();
        tma_stage_rows(bA, &p.m_a, nullptr, row0n, p.M, &bars[B_A], lane);
      }
      MGN_T(8);
      MGN_W(B_E1 + 5, par);
      MGN_T(9);
      if (lane == 0) {
        if (need_ga) {  // g_A tile (X) -> global
          tma_store_2d(&p.m_ga, smem_u32(bX), 0, static_cast<int>(row0));
          tma_store_2d(&p.m_ga, smem_u32(bX) + kPB, 64, static_cast<int>(row0));
          tma_store_commit();
          tma_store_wait_read();
        }
        if (more) mbar_arrive(&bars[B_G]);  // X is the next tile's H1: its E1 may write there now
      }
      __syncwarp();
      MGN_T(10);
    }
#undef MGN_W
    if (lane == 0) tma_store_wait_all();
  } else {
    // =========================== epilogue (8 warps) ===========================
    // two warps per TMEM lane quarter: warp (q, ch) owns tile rows [32q, 32q+32) (thread = row = TMEM lane)
    // and columns [64 ch, 64 ch + 64) = panel ch of every tile buffer, processed as two 32-column halves
    const int q = warp & 3;
    const int ch = (warp - 5) >> 2;
    const int row = q * 32 + lane;
    const int c0 = ch * 64;
    const uint32_t t_acc = tAcc + (static_cast<uint32_t>(q * 32) << 16) + c0;
    const float* b1 = sPar + c0;
    const float* b2 = sPar + kH + c0;
    const float* b3 = sPar + 2 * kH + c0;
    const float* gam = sPar + 3 * kH + c0;
    // LayerNorm row-sum exchange slot of (row, column half): the first 16 bytes of this thread's own 128-byte span
    // of the A buffer (only this thread reads that span in E3, so it may overwrite it as soon as it has)
    const uint32_t xch_own = ch * kPB + sw128_offset(row, 0);
    const uint32_t xch_other = (ch ^ 1) * kPB + sw128_offset(row, 0);
    float gg[4] = {0.f, 0.f, 0.f, 0.f};  // gamma gradient: lane (< 16) holds column c0 + 16 g + lane over this warp's rows
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
// the two warps that share tile rows (same TMEM lane quarter) synchronise on their own named barrier
#define MGN_ROW_SYNC() asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory")
    for (int it = 0; it < n_my; ++it) {
      const uint32_t par = it & 1;
      uint8_t* bA = MGN_BUF(R_A, it);
      uint8_t* bX = MGN_BUF(R_X, it);
      uint8_t* bH1 = MGN_BUF(R_H1, it);
      uint8_t* bH2 = MGN_BUF(R_H2, it);
      // ---- E1: h1 = relu(acc + b1 + g1 rows + g2 rows) -> H1
      MGN_W(B_MMA1 + 0, par);
      MGN_W(B_G, par);
      MGN_T(0);
      tc_fence_after_sync();
#pragma unroll 1
      for (int hh = 0; hh < 2; ++hh) {
        const int cc = c0 + 32 * hh;
        uint32_t v[32];
        tmem_ld32(t_acc + 32 * hh, v);
        uint32_t ga[16], gb[16];
        if (has_g) row_load32p(bX, row, cc, ga);
        if (has_g2) row_load32p(bH2, row, cc, gb);
        tmem_ld_wait();
        uint32_t o[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {  // two fp32 lanes per instruction (mgn_tile.cuh); relu after the bf16 rounding
          uint64_t z = f2_add(f2_packu(v[2 * j], v[2 * j + 1]), f2_ld(b1 + 32 * hh + 2 * j));
          if (has_g) z = f2_add(z, f2_from_bf16x2(ga[j]));
          if (has_g2) z = f2_add(z, f2_from_bf16x2(gb[j]));
          o[j] = relu_bf16x2(f2_to_bf16x2(z));
        }
        row_store32p(bH1, row, cc, o);
      }
      MGN_EPI_DONE(B_E1 + 0);
      MGN_T(1);
      // ---- E2: h2 = relu(acc + b2) -> H2
      MGN_W(B_MMA1 + 1, par);
      MGN_T(2);
      tc_fence_after_sync();
#pragma unroll 1
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t v[32];
        tmem_ld32(t_acc + 32 * hh, v);
        tmem_ld_wait();
        uint32_t o[16];
#pragma unroll
        for (int j = 0; j < 16; ++j)
          o[j] = relu_bf16x2(f2_to_bf16x2(f2_add(f2_packu(v[2 * j], v[2 * j + 1]), f2_ld(b2 + 32 * hh + 2 * j))));
        row_store32p(bH2, row, c0 + 32 * hh, o);
      }
      MGN_EPI_DONE(B_E1 + 1);
      MGN_T(3);
      // ---- E3: LayerNorm backward: g_out = go1 (+ go2) -> X ; g_y -> A
      MGN_W(B_MMA1 + 2, par);
      MGN_T(4);
      MGN_W(B_GO, par);
      MGN_T(5);
      tc_fence_after_sync();
      if (has_ln) {
        float s_y, s_yy, s_g, s_gy;
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
        // g_y = rstd ghat - k2 y + k0, xhat = rstd y - rstd mu (two FMAs per element; see mgn_edge_bwd2_tc.cu)
        const float k2 = rstd * rstd * m2;
        const uint64_t A2 = f2_splat(rstd), NK2 = f2_splat(-k2), K0 = f2_splat(fmaf(k2, mu, -rstd * m1)),
                       NAMU = f2_splat(-rstd * mu);
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
      } else {
#pragma unroll 1
        for (int hh = 0; hh < 2; ++hh) {
          const int cc = c0 + 32 * hh;
          uint32_t go[16];
          row_load32p(bX, row, cc, go);
          if (has_go2) {
            uint32_t g2[16];
            row_load32p(bA, row, cc, g2);
#pragma unroll
            for (int j = 0; j < 16; ++j) go[j] = add_bf16x2(go[j], g2[j]);
            row_store32p(bX, row, cc, go);
          }
          row_store32p(bA, row, cc, go);
        }
      }
      MGN_EPI_DONE(B_E1 + 2);
      MGN_T(6);
      // ---- E4: g_z2 = acc * (h2 > 0), in place in H2
      MGN_W(B_MMA1 + 3, par);
      MGN_T(7);
      tc_fence_after_sync();
      {
        // (the layer's weight-gradient MMAs still read this buffer: compute into registers, store once they are done)
        uint32_t hq[2][16];
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t v[32];
          tmem_ld32(t_acc + 32 * hh, v);
          row_load32p(bH2, row, c0 + 32 * hh, hq[hh]);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j)
            hq[hh][j] = mask_pos_bf16x2(pack_bf16x2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])), hq[hh][j]);
        }
        MGN_W(B_W3, par);
        row_store32p(bH2, row, c0, hq[0]);
        row_store32p(bH2, row, c0 + 32, hq[1]);
      }
      MGN_EPI_DONE(B_E1 + 3);
      MGN_T(8);
      // ---- E5: g_z1 = acc * (h1 > 0), in place in H1
      MGN_W(B_MMA1 + 4, par);
      MGN_T(9);
      tc_fence_after_sync();
      {
        // (the layer's weight-gradient MMAs still read this buffer: compute into registers, store once they are done)
        uint32_t hq[2][16];
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t v[32];
          tmem_ld32(t_acc + 32 * hh, v);
          row_load32p(bH1, row, c0 + 32 * hh, hq[hh]);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j)
            hq[hh][j] = mask_pos_bf16x2(pack_bf16x2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])), hq[hh][j]);
        }
        MGN_W(B_W2, par);
        row_store32p(bH1, row, c0, hq[0]);
        row_store32p(bH1, row, c0 + 32, hq[1]);
      }
      MGN_EPI_DONE(B_E1 + 4);
      MGN_T(10);
      // ---- E6: g_A = acc (+ g_out), in place in X (runs while the layer-1 weight-gradient MMAs execute)
      MGN_W(B_MMA1 + 5, par);
      MGN_T(11);
      tc_fence_after_sync();
      if (need_ga) {
#pragma unroll 1
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t v[32];
          tmem_ld32(t_acc + 32 * hh, v);
          uint32_t go[16];
          if (p.add_gout) row_load32p(bX, row, c0 + 32 * hh, go);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j)
            go[j] = p.add_gout ? f2_to_bf16x2(f2_add(f2_packu(v[2 * j], v[2 * j + 1]), f2_from_bf16x2(go[j])))
                               : pack_bf16x2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
          row_store32p(bX, row, c0 + 32 * hh, go);
        }
      }
      MGN_EPI_DONE(B_E1 + 5);
      MGN_T(12);
    }
#undef MGN_W
    asm volatile("bar.sync 10, 384;" ::: "memory");  // movers + epilogue: nobody reads a tile buffer any more
    if (lane < 16) {
#pragma unroll
      for (int g = 0; g < 4; ++g) scratch[4 * 8 * kH + q * kH + c0 + g * 16 + lane] = gg[g];
    }
  }
  if (tm_on) {
    const int role = warp == 0 ? 0 : (warp == kLoaderWarp ? 1 : 2);
    for (int i = 0; i < 16; ++i) p.timing[role * 32 + i] = tm[i];
  }

  if (timed_out && p.status != nullptr) atomicOr(p.status, kStatusTimeout);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();

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
      const int ncol = (w == 0) ? N1 : kH;
      float* dst = part + (w == 0 ? PT::kW1 : (w == 1 ? PT::kW2 : PT::kW3)) + row * ncol;
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
    part[PT::kB1 + tid] = s[0];
    part[PT::kB2 + tid] = s[1];
    part[PT::kB3 + tid] = s[2];
    part[PT::kBeta + tid] = s[3];
    part[PT::kGamma + tid] = sg;
  }
  tc_fence_before_sync();
  __syncthreads();  // every tcgen05.ld of the dump above has completed
  if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int KP>
static int launch(Params& p, int grid, cudaStream_t st) {
  using L = Smem<KP>;
  static PerDeviceFlag configured_flag;
  bool& configured = configured_flag.get();
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(mlp3_bwd_tc_kernel<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal);
    if (e != cudaSuccess) return static_cast<int>(e);
    configured = true;
  }
  p.part_floats = Part<KP>::kTotal;
  mlp3_bwd_tc_kernel<KP><<<grid, kThreads, L::kTotal, MGN_ST(st)>>>(p);
  return mgn_launch_status();
}

}  // namespace bwd
}  // namespace mgn

using namespace mgn;

static int bwd_grid(int64_t M) {
  const long long n_tiles = (M + bwd::kRows - 1) / bwd::kRows;
  return static_cast<int>(n_tiles < num_sms() ? n_tiles : num_sms());
}

#ifdef MGN_DEBUG_HOOKS
static long long* g_bwd_timing = nullptr;
/* debug hook: device buffer of 96 int64 that CTA 0 of the next backward launches fills with per-phase cycles */
extern "C" int mgn_debug_set_bwd_timing(void* dev_buf) {
  g_bwd_timing = static_cast<long long*>(dev_buf);
  return MGN_OK;
}
#else
static constexpr long long* g_bwd_timing = nullptr;
#endif

extern "C" size_t mgn_mlp3_bwd_tc_workspace_bytes(int64_t M) {
  if (M <= 0) return 0;
  return static_cast<size_t>(bwd_grid(M)) * bwd::Part<2>::kTotal * sizeof(float);
}

extern "C" int mgn_mlp3_bwd_tc(const void* a_tab, const int32_t* a_idx, const void* small_x, int small_in,
                               int small_is_f32, const void* g1_tab, const int32_t* g1_idx, int64_t g1_ld,
                               int64_t g1_col0, const void* g2_tab, const int32_t* g2_idx, int64_t g2_ld,
                               int64_t g2_col0, const void* go1, const int32_t* go1_idx, const void* go2,
                               const int32_t* go2_idx, int64_t M,
                               const float* w1, int64_t ld_w1, const float* b1, const float* w2, const float* b2,
                               const float* w3, const float* b3, const float* gamma, int n_out, float eps, void* g_a,
                               int add_gout, void* g_z1, int64_t g_z1_ld, float* g_w1, int64_t ld_gw1, float* g_b1, float* g_w2,
                               float* g_b2, float* g_w3, float* g_b3, float* g_gamma, float* g_beta, void* workspace,
                               size_t workspace_bytes, int* status, mgn_stream_t stream) {
  MGN_CHECK_ARG(M >= 0 && w1 && w2 && w3 && n_out >= 1 && n_out <= bwd::kH && go1 != nullptr);
  MGN_CHECK_ARG(gamma == nullptr || n_out == bwd::kH);
  if (M == 0) return MGN_OK;  // caller zero-fills gradients of an empty batch
  MGN_CHECK_ARG(workspace != nullptr);
  if (workspace_bytes < mgn_mlp3_bwd_tc_workspace_bytes(M)) return MGN_EWORKSPACE;
  bwd::Params p{};
  p.a = bwd::RowSrc{static_cast<const bf16*>(a_tab), a_idx, bwd::kH, 0};
  p.small_x = small_x;
  p.small_in = small_in;
  p.small_is_f32 = small_is_f32;
  p.g1 = bwd::RowSrc{static_cast<const bf16*>(g1_tab), g1_idx, g1_ld, g1_col0};
  p.g2 = bwd::RowSrc{static_cast<const bf16*>(g2_tab), g2_idx, g2_ld, g2_col0};
  p.go1 = bwd::RowSrc{static_cast<const bf16*>(go1), go1_idx, n_out, 0};
  p.go2 = bwd::RowSrc{static_cast<const bf16*>(go2), go2_idx, bwd::kH, 0};
  p.go_small = n_out < bwd::kH;
  MGN_CHECK_ARG(!(p.go_small && (go2 != nullptr || go1_idx != nullptr)));
  if (g1_tab) MGN_CHECK_ARG(g1_ld % 8 == 0 && g1_col0 % 8 == 0 && (reinterpret_cast<uintptr_t>(g1_tab) & 15) == 0);
  if (g2_tab) MGN_CHECK_ARG(g1_tab && g2_ld % 8 == 0 && g2_col0 % 8 == 0 && (reinterpret_cast<uintptr_t>(g2_tab) & 15) == 0);
  p.M = M;
  p.w1 = w1; p.b1 = b1; p.w2 = w2; p.b2 = b2; p.w3 = w3; p.b3 = b3; p.gamma = gamma;
  p.ld_w1 = ld_w1;
  p.n_out = n_out;
  p.eps = eps;
  p.g_a = static_cast<bf16*>(g_a);
  p.add_gout = add_gout;
  p.g_z1 = static_cast<bf16*>(g_z1);
  p.g_z1_ld = g_z1_ld > 0 ? g_z1_ld : bwd::kH;
  if (g_z1) MGN_CHECK_ARG(p.g_z1_ld % 8 == 0 && (reinterpret_cast<uintptr_t>(g_z1) & 15) == 0);
  p.partials = static_cast<float*>(workspace);
  p.status = status;
  p.timing = g_bwd_timing;
  cudaStream_t st = as_stream(stream);
  const int grid = bwd_grid(M);
  int rc;
  int n1;
  {  // tensor maps of every table the loader warp touches
    auto mk = [&](CUtensorMap* m, const bwd::RowSrc& s) -> int {
      if (s.tab == nullptr) return 0;
      return tma_make_rows_map(m, s.tab + s.col0, s.idx ? bwd::kOobRow : M, s.ld, s.idx ? 1 : 128);
    };
    int e = 0;
    if (small_in <= 0) e |= mk(&p.m_a, p.a);
    e |= mk(&p.m_g1, p.g1);
    e |= mk(&p.m_g2, p.g2);
    if (!p.go_small) e |= mk(&p.m_go1, p.go1);
    e |= mk(&p.m_go2, p.go2);
    if (g_a) e |= tma_make_rows_map(&p.m_ga, g_a, M, bwd::kH, 128);
    if (g_z1) e |= tma_make_rows_map(&p.m_gz1, g_z1, M, p.g_z1_ld, 128);
    if (e != 0) return MGN_EINVAL;
  }
  if (small_in > 0) {
    MGN_CHECK_ARG(small_x != nullptr && small_in <= 64 && g_a == nullptr && ld_w1 >= small_in);
    p.k1_true = small_in;
    n1 = small_in;
    rc = bwd::launch<1>(p, grid, st);
  } else {
    MGN_CHECK_ARG(a_tab != nullptr && ld_w1 >= bwd::kH && (reinterpret_cast<uintptr_t>(a_tab) & 15) == 0);
    p.k1_true = bwd::kH;
    n1 = bwd::kH;
    rc = bwd::launch<2>(p, grid, st);
  }
  if (rc != MGN_OK) return rc;
  ReduceParams rp{};
  rp.partials = p.partials;
  rp.stride = p.part_floats;
  rp.n_parts = grid;
  const int kp = small_in > 0 ? 1 : 2;
  const int oW2 = bwd::kH * 64 * kp, oW3 = oW2 + bwd::kH * bwd::kH, oB1 = oW3 + bwd::kH * bwd::kH;
  int ns = 0;
  // gW1: the TMEM accumulator is [128][64*kp]; only the first n1 columns are real
  rp.seg[ns++] = ReduceSeg{g_w1, ld_gw1, bwd::kH, n1, 0, 64 * kp};
  rp.seg[ns++] = ReduceSeg{g_w2, bwd::kH, bwd::kH, bwd::kH, oW2, bwd::kH};
  rp.seg[ns++] = ReduceSeg{g_w3, bwd::kH, n_out, bwd::kH, oW3, bwd::kH};
  rp.seg[ns++] = ReduceSeg{g_b1, bwd::kH, 1, bwd::kH, oB1, bwd::kH};
  rp.seg[ns++] = ReduceSeg{g_b2, bwd::kH, 1, bwd::kH, oB1 + bwd::kH, bwd::kH};
  rp.seg[ns++] = ReduceSeg{g_b3, bwd::kH, 1, n_out, oB1 + 2 * bwd::kH, bwd::kH};
  rp.seg[ns++] = ReduceSeg{g_gamma, bwd::kH, 1, bwd::kH, oB1 + 3 * bwd::kH, bwd::kH};
  rp.seg[ns++] = ReduceSeg{g_beta, bwd::kH, 1, bwd::kH, oB1 + 4 * bwd::kH, bwd::kH};
  rp.n_seg = ns;
  reduce_cta_partials_kernel<<<dim3(64, ns), 256, 0, MGN_ST(st)>>>(rp);
  return mgn_launch_status();
}
