// mgn_edge_fwd3.h — launch interface of the two-tiles-in-flight edge forward kernel (mgn_edge_fwd3_tc.cu), used by
// the mgn_edge_block_fwd*_tc entry points in mgn_mlp_fwd2_tc.cu.
#pragma once
#include "mgn_common.cuh"

namespace mgn {
namespace fwd3 {
struct Args {
  const bf16* a;  // efeat [M,128], dense rows (also the residual)
  long long M;
  const bf16* g1_tab;  // source projections, gathered by g1_idx into shared memory
  const int32_t* g1_idx;
  long long g1_ld, g1_col0;
  const bf16* g2_tab;  // destination projections, read directly by g2_idx (= destination of every row, ascending)
  const int32_t* g2_idx;
  long long g2_ld, g2_col0;
  const float *w1, *b1, *w2, *b2, *w3, *b3, *gamma, *beta;
  long long ld_w1;
  float eps;
  const bf16* res;  // node form (nullptr: edge form): residual rows [M,128]; then g1_* are unused, g2_idx is ignored (own row)
  bf16* out;  // [M,128]
  bf16* h1_out;  // [M,128] first hidden activation relu(z1), kept for the backward pass (nullptr: not stored)
  // fused destination sums (mgn_agg.cuh); seg_off == nullptr: off
  const int32_t* seg_off;
  bf16* agg;
  long long ld_agg;
  float* agg_part;
  int32_t* agg_part_v;
  long long agg_row_base, agg_rec_base;
  int* status;
};
}  // namespace fwd3
int edge_fwd3_launch(const fwd3::Args& args, cudaStream_t st);
void edge_fwd3_set_timing(long long* buf);  // debug: 6 x int64 per-phase cycles of CTA 0 (nullptr: off)
}  // namespace mgn
