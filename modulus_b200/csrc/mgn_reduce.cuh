// mgn_reduce.cuh — deterministic second stage of the per-CTA partial reductions used by the
// tensor-core backward / wgrad kernels: every output element is the sum of one fp32 value per CTA,
// added in CTA order (no atomics, bit-reproducible).
#pragma once
#include "mgn_common.cuh"

namespace mgn {

struct ReduceSeg {
  float* dst;
  long long ld_dst;
  int rows, cols;
  int src_off;
  int src_ld;
};
struct ReduceParams {
  const float* partials;
  long long stride;
  int n_parts;
  int n_seg;
  ReduceSeg seg[8];
};

static __global__ void reduce_cta_partials_kernel(const ReduceParams rp) {
  const ReduceSeg sg = rp.seg[blockIdx.y];
  if (sg.dst == nullptr) return;
  const int n = sg.rows * sg.cols;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int r = i / sg.cols, c = i - r * sg.cols;
    const float* src = rp.partials + sg.src_off + r * sg.src_ld + c;
    // added in CTA order; the loads are independent, 16 of them in flight per thread (one per 200 KB-strided partial)
    float s = 0.f;
    int k = 0;
    for (; k + 16 <= rp.n_parts; k += 16) {
      float v[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) v[u] = __ldg(src + static_cast<long long>(k + u) * rp.stride);
#pragma unroll
      for (int u = 0; u < 16; ++u) s += v[u];
    }
    for (; k < rp.n_parts; ++k) s += __ldg(src + static_cast<long long>(k) * rp.stride);
    sg.dst[r * sg.ld_dst + c] = s;
  }
}

}  // namespace mgn
