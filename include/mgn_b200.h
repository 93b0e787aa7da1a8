/* mgn_b200.h — C ABI of libmgn_b200.so: the B200 (sm_100a) MeshGraphNet message-passing path.
 *
 * Drop-in boundary for NVIDIA PhysicsNeMo's MeshGraphNet hot path.  Every entry point takes
 * plain device pointers, sizes, a dtype enum and a CUDA stream; no torch types.  The library
 * never allocates user-visible memory (outputs and workspaces are caller-provided), keeps no
 * mutable global state (the only statics are write-once per-device caches: SM count, kernel attributes, and the
 * launch counter), never synchronises the host and is CUDA-graph capturable.  Profiling hooks exist only in builds
 * made with -DMGN_DEBUG_HOOKS (include/mgn_b200_debug.h); the product library does not export them.
 *
 * Return value of every function: 0 = ok, <0 = argument error (MGN_E*), >0 = cudaError_t.
 *
 * Conventions
 *   - feature tables are row-major [rows, D]; dtype is MGN_F32 or MGN_BF16 (fp32 accumulate)
 *   - weights / biases / LayerNorm affine parameters are ALWAYS fp32 and are read in place
 *     (nn.Linear layout [out, in]) -- the optimizer's tensors, no shadow copies
 *   - graph indices are int32; a graph is a CSC (offsets[n_dst+1], indices[E] = source id of
 *     each in-edge), edge e of the CSC is row e of every edge-feature table
 *   - "reference" citations are paths inside NVIDIA/modulus (physicsnemo 1.1.0a0)
 */
#ifndef MGN_B200_H_
#define MGN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* mgn_stream_t; /* cudaStream_t */

enum { MGN_F32 = 0, MGN_BF16 = 1 };

enum {
  MGN_OK = 0,
  MGN_EINVAL = -1,       /* bad argument (null pointer, negative size, unknown enum) */
  MGN_EUNSUPPORTED = -2, /* shape/dtype combination this kernel family does not cover */
  MGN_EALIGN = -3,       /* pointer or leading dimension not 16-byte aligned */
  MGN_EWORKSPACE = -4    /* workspace too small */
};

/* activation ids (reference: physicsnemo/models/layers/activations.py:173-199) */
enum {
  MGN_ACT_NONE = 0,
  MGN_ACT_RELU = 1,
  MGN_ACT_SILU = 2,
  MGN_ACT_TANH = 3,
  MGN_ACT_SIGMOID = 4,
  MGN_ACT_GELU = 5,
  MGN_ACT_LEAKY_RELU = 6,
  MGN_ACT_ELU = 7
};

int mgn_version(void);
/* sha256 of the sources + flags this binary was built from (modulus_b200/build.py); the loader refuses a library whose
 * digest differs from the sources next to it, so a stale .so can never be bound to a newer header. */
const char* mgn_build_digest(void);
const char* mgn_error_string(int code);
/* kernels launched by this library since it was loaded (all streams, all entry points) */
int64_t mgn_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * Graph plan: CSC -> CSR transpose (the structure the backward scatter walks).
 * Replaces: cugraph-ops StaticCSC/BipartiteCSC(reverse_graph_bwd=True) built in
 *           models/gnn_layers/graph.py:347-445 and DGL's reverse-graph formats used by
 *           apply_edges/update_all backward (gnn_layers/utils.py:141-147, 370-372).
 *   csc_dst[e]      destination of CSC edge e (expanded offsets; graph.py:462-471)
 *   csr_offsets[u]  start of source u's out-edge list           [n_src+1]
 *   csr_eids[j]     CSC edge position, ascending within a source [E]   (stable => bit-exact)
 * ---------------------------------------------------------------------------------------- */
size_t mgn_csr_workspace_bytes(int64_t n_src, int64_t n_dst, int64_t n_edges);
int mgn_csr_from_csc(const int32_t* offsets, const int32_t* indices, int64_t n_src, int64_t n_dst,
                     int64_t n_edges, int32_t* csc_dst, int32_t* csr_offsets, int32_t* csr_eids,
                     void* workspace, size_t workspace_bytes, mgn_stream_t stream);

/* Stable grouping of n items by integer key in [0, n_keys): offsets[n_keys+1] and ids[n] (item
 * positions, ascending inside each key).  COO -> CSC is group_by(dst) and COO -> CSR is
 * group_by(src); replaces DGL's adj_tensors("csc") used by CuGraphCSC.from_dgl
 * (models/gnn_layers/graph.py:143-193) for graphs that arrive in edge-id order. */
size_t mgn_group_by_key_workspace_bytes(int64_t n_keys);
int mgn_group_by_key(const int32_t* keys, int64_t n, int64_t n_keys, int32_t* offsets, int32_t* ids,
                     void* workspace, size_t workspace_bytes, mgn_stream_t stream);

/* out[j] = s for j in [offsets[s], offsets[s+1])  (destination id of every CSC edge,
 * graph.py:462-471 repeat_interleave) */
int mgn_expand_offsets(const int32_t* offsets, int64_t n_segments, int32_t* out, mgn_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Operator seam (models/gnn_layers/utils.py)
 * ---------------------------------------------------------------------------------------- */
/* concat_efeat forward, utils.py:151-229 / concat_message_function :94-109:
 *   out[e] = [ efeat[e] | src_feat[src_idx[e]] | dst_feat[dst_idx[e]] ]      out: [E, De+Ds+Dd] */
int mgn_concat_efeat_fwd(int dtype, const void* efeat, int64_t De, const void* src_feat, int64_t Ds,
                         const void* dst_feat, int64_t Dd, const int32_t* src_idx,
                         const int32_t* dst_idx, int64_t n_edges, void* out, mgn_stream_t stream);

/* sum_efeat forward, utils.py:232-334:  out[e] = efeat[e] + src_feat[src_idx[e]] + dst_feat[dst_idx[e]] */
int mgn_sum_efeat_fwd(int dtype, const void* efeat, const void* src_feat, const void* dst_feat,
                      int64_t D, const int32_t* src_idx, const int32_t* dst_idx, int64_t n_edges,
                      void* out, mgn_stream_t stream);

/* Deterministic, atomic-free segmented sum (the reduction inside aggregate_and_concat,
 * utils.py:337-427, and the backward of both gathers):
 *   out[s, out_col0 : out_col0+D] (+)= scale_s * sum_{j in [offsets[s], offsets[s+1])}
 *                                       in[ eids ? eids[j] : j , in_col0 : in_col0+D ]
 * scale_s = 1 (sum) or 1/max(len_s,1) (mean).  accumulate != 0 adds to the existing out. */
int mgn_segment_sum(int dtype, const void* in, int64_t ld_in, int64_t in_col0, int64_t D,
                    const int32_t* offsets, const int32_t* eids, int64_t n_segments, void* out,
                    int64_t ld_out, int64_t out_col0, int mean, int accumulate,
                    mgn_stream_t stream);
/* mgn_segment_sum with hub handling for skewed-degree graphs: segments longer than 512 rows are split into 2048-row
 * chunks, each summed by a whole CTA into an fp32 partial, partials combined in chunk order (bit-reproducible).
 * n_rows = rows of `in` covered by the offsets; workspace >= mgn_segment_sum_workspace_bytes(n_rows, D). */
size_t mgn_segment_sum_workspace_bytes(int64_t n_rows, int64_t D);
int mgn_segment_sum_balanced(int dtype, const void* in, int64_t ld_in, int64_t in_col0, int64_t D,
                             const int32_t* offsets, const int32_t* eids, int64_t n_segments, void* out,
                             int64_t ld_out, int64_t out_col0, int mean, int accumulate, int64_t n_rows,
                             void* workspace, size_t workspace_bytes, mgn_stream_t stream);

/* Halo exchange over peer memory (NVLink / NVSwitch), replacing gather + all-to-all + copy-in of the reference's
 * indexed_all_to_all_v (physicsnemo/distributed/utils.py:541-765) for ranks of one node:
 *   rows [seg_begin[r], seg_begin[r+1]) of the send list (row i = bytes [col0_bytes, col0_bytes + row_bytes) of table row
 *   idx ? idx[i] : i) are stored at dst_base[r] + (i - seg_begin[r]) * dst_ld_bytes, dst_base[r] being a PEER-MAPPED address
 *   in rank r's receive buffer; when every store of the launch is done, `epoch` is written (system-scope release) to
 *   flag_addr[r] (peer-mapped address of this rank's flag slot in r's memory; 0 = no flag).  seg_begin / dst_base /
 *   flag_addr are HOST arrays of n_peers+1 / n_peers / n_peers entries (n_peers <= 16); counter = one zeroed device int.
 * mgn_halo_wait: blocks the stream until flags[r] >= epoch for every r in need_mask (system-scope acquire); a peer that
 * does not arrive within ~10 s sets bit 8 of *status instead of hanging the device. */
int mgn_halo_push(const void* tab, int64_t ld_bytes, int64_t col0_bytes, int64_t row_bytes, const int32_t* idx,
                  int64_t n_rows, int n_peers, const int64_t* seg_begin, const int64_t* dst_base, int64_t dst_ld_bytes,
                  const int64_t* flag_addr, int epoch, void* counter, mgn_stream_t stream);
int mgn_halo_wait(const void* flags, int need_mask, int epoch, int* status, mgn_stream_t stream);

/* Row gather into a column slice (backward of the segmented sum, halo packing):
 *   out[r, out_col0 : out_col0+D] = scale_r * in[ idx ? idx[r] : r , in_col0 : in_col0+D ]
 * inv_deg_offsets != NULL: scale_r = 1/max(deg(idx[r]),1) with deg from those CSC offsets. */
int mgn_gather_rows(int dtype, const void* in, int64_t ld_in, int64_t in_col0, int64_t D,
                    const int32_t* idx, int64_t n_rows, void* out, int64_t ld_out,
                    int64_t out_col0, const int32_t* inv_deg_offsets, mgn_stream_t stream);

/* out[idx[r]] += in[r] for r in order, deterministic because idx_sorted_perm lists rows grouped by
 * destination (backward of the halo gather; replaces index_add_ in distributed/utils.py:688-705).
 * Implemented on top of mgn_segment_sum; see modulus_b200/distributed. */

/* ------------------------------------------------------------------------------------------
 * Dense pieces of MeshGraphMLP (models/gnn_layers/mesh_graph_mlp.py:142-203), fp32-accurate
 * SIMT path used for fp32 models and as the on-device cross-check of the tensor-core path.
 * ---------------------------------------------------------------------------------------- */
/* y = x W^T + b ; h = act(y).  y_pre may be NULL (not stored); h may alias nothing. */
int mgn_linear_fwd(int dtype, const void* x, int64_t ldx, int64_t M, int64_t K, const float* w,
                   const float* b, int64_t N, int act, void* y_pre, void* h, int64_t ldh,
                   mgn_stream_t stream);
/* y = act(x) elementwise (the leading activation of MeshGraphEdgeMLPSum, mesh_graph_mlp.py:352) */
int mgn_act_fwd(int dtype, const void* x, int act, void* y, int64_t n, mgn_stream_t stream);
/* g_y = g_h * act'(y_pre)  (ReLU may pass the post-activation as y_pre) */
int mgn_act_bwd(int dtype, const void* g_h, const void* y_pre, int act, void* g_y, int64_t n,
                mgn_stream_t stream);
/* g_x[M,K] = g_y[M,N] W[N,K] */
int mgn_linear_bwd_data(int dtype, const void* g_y, int64_t M, int64_t N, const float* w, int64_t K,
                        void* g_x, int64_t ldgx, mgn_stream_t stream);
/* g_w[N,K] = g_y^T x ; g_b[N] = colsum(g_y)   (fp32 outputs, deterministic two-stage split-M) */
size_t mgn_linear_bwd_weight_workspace_bytes(int64_t M, int64_t N, int64_t K);
int mgn_linear_bwd_weight(int dtype, const void* g_y, const void* x, int64_t ldx, int64_t M, int64_t N,
                          int64_t K, float* g_w, float* g_b, void* workspace, size_t workspace_bytes,
                          mgn_stream_t stream);
/* out = LayerNorm(x) * gamma + beta (+ residual); eps as nn.LayerNorm (1e-5); mean/rstd fp32 [M] */
int mgn_layernorm_fwd(int dtype, const void* x, int64_t M, int64_t D, const float* gamma,
                      const float* beta, float eps, const void* residual, void* out, float* mean,
                      float* rstd, mgn_stream_t stream);
size_t mgn_layernorm_bwd_workspace_bytes(int64_t M, int64_t D);
int mgn_layernorm_bwd(int dtype, const void* g_out, const void* x, const float* mean,
                      const float* rstd, const float* gamma, int64_t M, int64_t D, void* g_x,
                      float* g_gamma, float* g_beta, void* workspace, size_t workspace_bytes,
                      mgn_stream_t stream);
/* out = a + b (elementwise, n elements) */
int mgn_add(int dtype, const void* a, const void* b, void* out, int64_t n, mgn_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Fused tensor-core path (bf16 storage, hidden width 128, ReLU): tcgen05.mma + TMEM.
 * ---------------------------------------------------------------------------------------- */
/* Fused MeshGraphMLP forward over 128-row tiles with the first Linear split the way the reference's "concat trick" does
 * (MeshGraphEdgeMLPSum, mesh_graph_mlp.py:278-458: per-node partial products gathered and summed per edge; all operand rows staged by cp.async, output
 * tile leaves through coalesced stores, next tile prefetched while the current one computes):
 *     z1  = A W1^T + b1 + g1_tab[g1_idx[r]] + g2_tab[g2_idx[r]]     (A = a_tab[a_idx[r]] or raw small_x)
 *     out = [LayerNorm]( W3 relu(W2 relu(z1) + b2) + b3 ) + residual
 * residual: res_is_a != 0 -> the A rows themselves (MeshEdgeBlock: efeat' = mlp(...) + efeat,
 * mesh_edge_block.py:95); res_tab != NULL -> rows of another [M,128] table (MeshNodeBlock: + nfeat,
 * mesh_node_block.py:91; excludes g2_tab). */
int mgn_mlp3_fwd2_tc(const void* a_tab, const int32_t* a_idx, const void* small_x, int small_in,
                     int small_is_f32, const void* g1_tab, const int32_t* g1_idx, int64_t g1_ld,
                     int64_t g1_col0, const void* g2_tab, const int32_t* g2_idx, int64_t g2_ld,
                     int64_t g2_col0, const void* res_tab, int res_is_a, int64_t M, const float* w1,
                     int64_t ld_w1, const float* b1, const float* w2, const float* b2, const float* w3,
                     const float* b3, const float* gamma, const float* beta, int n_out, float eps, void* out,
                     int64_t ld_out, int* status, mgn_stream_t stream);

/* Fused MeshGraphMLP backward over 128-row tiles (forward hiddens are recomputed, nothing but the
 * layer inputs is read back):
 *     z1 = A W1^T + G + b1,  A = a_tab[a_idx[r]] (or raw small_x),  G = g1_tab[g1_idx[r]] (+ g2_tab[g2_idx[r]])
 *     g_out = go1[go1_idx[r]] (+ go2[go2_idx[r]])             incoming gradient of the block output (idx NULL = r)
 *   outputs
 *     g_a   [M,128] bf16 = dL/dA (+ g_out when add_gout: the residual connection)      (nullable)
 *     g_z1  [M,128] bf16 (row stride g_z1_ld) = dL/dz1 (= dL/dG rows; reduced per node by mgn_segment_sum) (nullable)
 *     g_w1 [128, k1] (row stride ld_gw1), g_w2, g_w3 [n_out,128], g_b1..3, g_gamma, g_beta   fp32, nullable
 * Replaces the autograd backward of MeshEdgeBlock.forward (mesh_edge_block.py:88-96), MeshNodeBlock.forward
 * (mesh_node_block.py:82-92) and MeshGraphMLP.forward (mesh_graph_mlp.py:200-203): cuBLAS dgrad/wgrad GEMMs,
 * native_layer_norm_backward, threshold_backward and the index_add_ scatter of the gathered rows.
 * w1 is read with row stride ld_w1 (a column block of the [128, 3*128] first Linear).  gamma == NULL:
 * no LayerNorm (decoder; then n_out may be < 128 and go1 is a dense [M, n_out] bf16 matrix). */
size_t mgn_mlp3_bwd_tc_workspace_bytes(int64_t M);
int mgn_mlp3_bwd_tc(const void* a_tab, const int32_t* a_idx, const void* small_x, int small_in,
                    int small_is_f32, const void* g1_tab, const int32_t* g1_idx, int64_t g1_ld,
                    int64_t g1_col0, const void* g2_tab, const int32_t* g2_idx, int64_t g2_ld,
                    int64_t g2_col0, const void* go1, const int32_t* go1_idx, const void* go2,
                    const int32_t* go2_idx, int64_t M,
                    const float* w1, int64_t ld_w1, const float* b1, const float* w2, const float* b2,
                    const float* w3, const float* b3, const float* gamma, int n_out, float eps, void* g_a,
                    int add_gout, void* g_z1, int64_t g_z1_ld, float* g_w1, int64_t ld_gw1, float* g_b1, float* g_w2,
                    float* g_b2, float* g_w3, float* g_b3, float* g_gamma, float* g_beta, void* workspace,
                    size_t workspace_bytes, int* status, mgn_stream_t stream);

/* MeshEdgeBlock.forward + the "sum" aggregation of the following MeshNodeBlock in one pass (mesh_edge_block.py:88-96,
 * utils.py:337-378): efeat_out[e] = efeat[e] + LN(MLP(efeat[e] W1a^T + P_src[src[e]] + P_dst[dst[e]])) for CSC-ordered
 * edges, and agg[v] = sum of efeat_out over the incoming edges of v, taken from the result tiles while they are still
 * in shared memory (fixed summation order; nodes without incoming edges get zeros).  h1_out (nullable, [n_edges,128]
 * bf16) receives the first hidden activation relu(z1) for mgn_edge_block_bwd_tc. */
size_t mgn_mlp3_fwd2_agg_workspace_bytes(int64_t n_edges);
int mgn_edge_block_fwd_tc(const void* efeat, const void* p_src, const int32_t* src_idx, int64_t p_src_ld,
                          int64_t p_src_col0, const void* p_dst, const int32_t* dst_idx, int64_t p_dst_ld,
                          int64_t p_dst_col0, int64_t n_edges, const float* w1, int64_t ld_w1, const float* b1,
                          const float* w2, const float* b2, const float* w3, const float* b3,
                          const float* gamma, const float* beta, float eps, void* efeat_out, void* h1_out,
                          const int32_t* csc_offsets, int64_t n_dst, void* agg, int64_t ld_agg,
                          void* workspace, size_t workspace_bytes, int* status, mgn_stream_t stream);

/* The same over one of several consecutive row ranges of the edge table (partitioned graphs run the interior range while
 * the halo exchange is in flight): pointers / indices are those of the range, row_base its CSC position, total_tiles the
 * sum of ceil(rows/128) over all ranges, rec_base that sum over the ranges before this one; the workspace
 * (2 * total_tiles * 516 bytes) is shared and mgn_agg_fixup completes agg once every range has run. */
int mgn_edge_block_fwd_part_tc(const void* efeat, const void* p_src, const int32_t* src_idx, int64_t p_src_ld,
                               int64_t p_src_col0, const void* p_dst, const int32_t* dst_idx, int64_t p_dst_ld,
                               int64_t p_dst_col0, int64_t n_rows, const float* w1, int64_t ld_w1,
                               const float* b1, const float* w2, const float* b2, const float* w3,
                               const float* b3, const float* gamma, const float* beta, float eps,
                               void* efeat_out, void* h1_out, const int32_t* csc_offsets, int64_t n_dst, void* agg,
                               int64_t ld_agg, void* workspace, size_t workspace_bytes, int64_t row_base,
                               int64_t total_tiles, int64_t rec_base, int* status, mgn_stream_t stream);
int mgn_agg_fixup(void* workspace, int64_t total_tiles, void* agg, int64_t ld_agg, int64_t n_dst,
                  mgn_stream_t stream);
/* MeshNodeBlock.forward (mesh_node_block.py:82-92) with the first Linear split per input block:
 *   nfeat_out[v] = nfeat[v] + LN(MLP(agg[v] W1a^T + P_node[v] + b1)),  P_node = columns [p_col0, p_col0+128) of p_tab;
 * h1_out (nullable, [n_nodes,128] bf16) receives relu(z1) for mgn_edge_block_bwd_tc with add_gout = 0. */
int mgn_node_block_fwd_tc(const void* agg, const void* p_tab, int64_t p_ld, int64_t p_col0, const void* nfeat,
                          int64_t n_nodes, const float* w1, int64_t ld_w1, const float* b1, const float* w2,
                          const float* b2, const float* w3, const float* b3, const float* gamma,
                          const float* beta, float eps, void* nfeat_out, void* h1_out, int* status,
                          mgn_stream_t stream);

/* MeshEdgeBlock backward from the first hidden activation h1 = relu(z1) that mgn_edge_block_fwd_tc stored (no gather
 * of the projection rows, no GEMM1 recompute): go1 (dense, or gathered by go1_idx) + go2[go2_idx] = gradient of the
 * block output; writes g_efeat [n_edges,128], g_z1 [n_edges,128] (row stride g_z1_ld; its CSC / CSR sums are the
 * gradients of the destination / source projection rows) and the fp32 parameter gradients of W1[:, :128] (row stride
 * ld_gw1), b1, W2, b2, W3, b3, gamma, beta (deterministic per-CTA partials + ordered reduction).  add_gout = 1: the
 * block's residual runs over the layer-1 input rows (edge block: g_efeat = g_z1 W1a + g_out); 0 for the node block,
 * whose layer-1 input is the aggregate (efeat := agg, h1 from mgn_node_block_fwd_tc, g_efeat := g_agg).
 * csc_offsets (nullable; edge block only): also write gz1_agg[v] = sum of g_z1 over the incoming edges of v (rows are
 * CSC-ordered edges with destinations dst_idx; workspace mgn_mlp3_fwd2_agg_workspace_bytes(n_edges)) -- the gradient
 * of the destination projection rows, taken from the g_z1 tiles while they are in shared memory. */
size_t mgn_edge_block_bwd_tc_workspace_bytes(int64_t n_edges);
int mgn_edge_block_bwd_tc(const void* efeat, const void* h1, const void* go1, const int32_t* go1_idx,
                          const void* go2, const int32_t* go2_idx, int64_t n_edges, const float* w1,
                          int64_t ld_w1, const float* w2, const float* b2, const float* w3, const float* b3,
                          const float* gamma, float eps, int add_gout, void* g_efeat, void* g_z1, int64_t g_z1_ld,
                          float* g_w1, int64_t ld_gw1, float* g_b1, float* g_w2, float* g_b2, float* g_w3,
                          float* g_b3, float* g_gamma, float* g_beta, void* workspace, size_t workspace_bytes,
                          const int32_t* csc_offsets, const int32_t* dst_idx, int64_t n_dst, void* gz1_agg,
                          int64_t ld_agg, void* agg_workspace, size_t agg_workspace_bytes, int* status,
                          mgn_stream_t stream);

/* Node-level plain GEMMs of the fused path (bf16 rows, fp32 weights read in place):
 *   mgn_node_gemm_tc / mgn_linear128_tc : out = x W^T (+ residual) over [M,128] column blocks
 *   mgn_wgrad_tc  : out[128*n_blocks, 128] (fp32, row stride ld_out) = G[M, 128*n_blocks]^T X[M,128]
 * They replace the cuBLAS GEMMs autograd runs for the node-row column blocks of the first Linear
 * (mesh_graph_mlp.py:142-168 / the lin_src, lin_dst products of MeshGraphEdgeMLPSum :396-405). */
/* K-looped tensor-core GEMM for MLP widths beyond 128 (GraphCast 512, AeroGraphNet 256; first layers with K = 768 / 1536):
 *   out[M,N] (bf16, row stride ld_out) = act( x[M,K] (bf16, ld_x) Wb[N,K]^T (bf16 image, ld_w) + bias[N] (fp32, nullable) )
 *   K % 64 == 0, N % 128 == 0, act = MGN_ACT_NONE | MGN_ACT_RELU; fp32 accumulation in TMEM, operands staged by TMA.
 * mgn_cast_weight_bf16 makes the bf16 image of an fp32 nn.Linear weight [rows, cols] (row stride ld): transpose == 0 ->
 * [rows, cols], transpose != 0 -> [cols, rows] (the operand of the data gradient g_x = g_y W).
 * Replaces the cuBLAS GEMMs behind nn.Linear in MeshGraphMLP (mesh_graph_mlp.py:142-168, 200-203) for bf16 activations. */
int mgn_cast_weight_bf16(const float* w, int64_t rows, int64_t cols, int64_t ld, void* out, int transpose,
                         mgn_stream_t stream);
int mgn_gemm_bf16_tc(const void* x, int64_t ld_x, int64_t M, int64_t K, const void* w_bf16, int64_t ld_w, int64_t N,
                     const float* bias, int act, void* out, int64_t ld_out, int* status, mgn_stream_t stream);
/* fp32 nn.Linear on the tensor cores with fp32 accuracy (3 x TF32 operand split, tcgen05.mma kind::tf32; tools/probe_tf32.cu):
 *   out[M,N] (fp32, row stride ld_out) = act( x[M,K] (fp32, ld_x) W[N,K]^T + bias[N] (nullable) ),  K % 32 == 0, N % 128 == 0,
 *   act = MGN_ACT_NONE | MGN_ACT_RELU.  w_split is the [2N, K] image written by mgn_split_weight_tf32 from the fp32 nn.Linear
 *   weight [rows, cols] (row stride ld): hi = tf32(w) in rows 0..N-1, lo = tf32(w - hi) in rows N..2N-1; transpose != 0 splits
 *   the transposed weight ((N, K) = (cols, rows): the operand of the data gradient g_x = g_y W).
 * Replaces the cuBLAS SGEMMs behind nn.Linear in MeshGraphMLP (mesh_graph_mlp.py:142-168, 200-203) for fp32 callers, the
 * reference's default precision (meshgraphnet.py:128-150).  Reached through modulus_b200.ops.linear_f32_tc; the fp32 MODEL path
 * still calls the exact-fp32 SIMT kernels below (mgn_linear_fwd / mgn_linear_bwd_data). */
int mgn_split_weight_tf32(const float* w, int64_t rows, int64_t cols, int64_t ld, float* out, int transpose,
                          mgn_stream_t stream);
int mgn_linear_f32_tc(const float* x, int64_t ld_x, int64_t M, int64_t K, const float* w_split, int64_t N,
                      const float* bias, int act, float* out, int64_t ld_out, int* status, mgn_stream_t stream);
/* out[M,128] (row stride ld_out) = x[M,128] (row stride ld_x) W^T + bias (+ residual[M,128]); second-generation
 * pipeline (cp.async staging, coalesced stores).  Wider products are issued per 128-column block. */
int mgn_linear128_tc(const void* x, int64_t ld_x, int64_t M, const float* w, int64_t ld_w, const float* bias,
                     const void* residual, void* out, int64_t ld_out, int* status, mgn_stream_t stream);
/* out[M, 128 nb] (row stride ld_out) = x[M, 128 kb] (row stride ld_x) W[128 nb, 128 kb]^T (+ residual[M,128], nb == 1),
 * kb * nb <= 3: TMA-staged persistent tcgen05 pipeline, one pass over x and one over out (the projection table
 * P = nfeat Wp^T with kb = 1, nb = 3 and the node-feature gradient g_nfeat + T Wp with kb = 3, nb = 1). */
int mgn_node_gemm_tc(const void* x, int64_t ld_x, int kb, int64_t M, const float* w, int64_t ld_w, int nb,
                     const void* residual, void* out, int64_t ld_out, int* status, mgn_stream_t stream);
size_t mgn_wgrad_tc_workspace_bytes(int64_t M, int n_blocks);
int mgn_wgrad_tc(const void* g, int64_t ld_g, int n_blocks, const void* x, int64_t ld_x, int64_t M,
                 float* out, int64_t ld_out, void* workspace, size_t workspace_bytes, int* status,
                 mgn_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * The steps either side of the path.
 * ---------------------------------------------------------------------------------------- */
/* Multi-tensor Adam: one launch (plus a one-thread step-counter tick) updates all n tensors.
 * Replaces: apex.optimizers.FusedAdam / torch.optim.Adam at examples/cfd/vortex_shedding_mgn/train.py:111-123,
 *           and GradScaler.step's skip-on-inf (:161-163) through found_inf / inv_scale.
 * params / grads / exp_avg / exp_avg_sq: DEVICE arrays of n fp32 device pointers (grads[i] may be NULL: tensor
 * skipped); numel: DEVICE int64[n]; chunk c covers elements [chunk_start[c], chunk_start[c] + chunk_elems) of tensor
 * chunk_tensor[c] (chunk_elems a multiple of 4).  step: DEVICE float counter, incremented by this call before the
 * update (bias corrections 1 - beta^step in double); lr_dev / inv_scale / found_inf: optional DEVICE scalars
 * (learning rate override, gradient multiplier 1/loss_scale, skip flag).  adamw != 0: decoupled weight decay,
 * else L2 added to the gradient (torch.optim.Adam).  Update order and rounding follow torch's single-tensor Adam.
 * No host synchronisation, CUDA-graph capturable. */
int mgn_adam_multi_step(void* const* params, const void* const* grads, void* const* exp_avg,
                        void* const* exp_avg_sq, const int64_t* numel, const int32_t* chunk_tensor,
                        const int64_t* chunk_start, int64_t n_chunks, int64_t chunk_elems, float lr,
                        float beta1, float beta2, float eps, float weight_decay, int adamw, float* step,
                        const float* lr_dev, const float* inv_scale, const float* found_inf,
                        mgn_stream_t stream);
/* out[e, 0:dim] = pos[src[e]] - pos[dst[e]], out[e, dim] = its Euclidean norm, then (x - mu[k]) / sd[k] per column
 * when mu / sd (fp32 [dim+1], optional) are given.  pos fp32 [n, dim], dim in {2, 3}; out fp32 [E, dim+1].
 * Replaces: VortexSheddingDataset.add_edge_features + normalize_edge
 *           (datapipes/gnn/vortex_shedding_dataset.py:324-349). */
int mgn_edge_features(const float* pos, int dim, const int32_t* src, const int32_t* dst, int64_t n_edges,
                      const float* mu, const float* sd, float* out, mgn_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MGN_B200_H_ */
