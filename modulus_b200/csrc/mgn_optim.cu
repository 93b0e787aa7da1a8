// mgn_optim.cu — the two steps either side of the message-passing path:
//   * multi-tensor Adam: ONE launch updates every parameter tensor of the model (263 tensors for the
//     default MeshGraphNet) from a device-resident pointer table.  Replaces apex FusedAdam /
//     torch.optim.Adam of examples/cfd/vortex_shedding_mgn/train.py:111-123.
//   * edge features of a mesh graph from node coordinates: relative displacement and its norm, optionally
//     normalised (datapipes/gnn/vortex_shedding_dataset.py:324-349).
// Both are HBM-bound elementwise passes: 16-byte accesses where the tensors allow it, nothing staged.
#include "mgn_common.cuh"

namespace mgn {
namespace {

constexpr int kAdamThreads = 256;

// one thread: advance the step counter unless the step is skipped (GradScaler found an inf)
__global__ void adam_tick_kernel(float* step, const float* found_inf) {
  if (found_inf != nullptr && *found_inf != 0.f) return;
  *step += 1.f;
}

struct AdamScalars {
  float lr, beta1, beta2, eps, weight_decay;
  int adamw;
};

__device__ __forceinline__ void adam_update(float& p, float g, float& m, float& v, const AdamScalars& s,
                                            float step_size, float inv_bc2_sqrt, float inv_scale) {
  g *= inv_scale;
  if (s.weight_decay != 0.f) {
    if (s.adamw)
      p *= 1.f - s.lr * s.weight_decay;  // decoupled decay (AdamW / apex adam_w_mode=True)
    else
      g += s.weight_decay * p;  // L2 (torch.optim.Adam)
  }
  m = m + (g - m) * (1.f - s.beta1);             // exp_avg.lerp_(grad, 1 - beta1)
  v = v * s.beta2 + (1.f - s.beta2) * g * g;     // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
  const float denom = sqrtf(v) * inv_bc2_sqrt + s.eps;
  p = p - step_size * (m / denom);
}

// grid = chunks; chunk c covers elements [chunk_start[c], chunk_start[c] + chunk_elems) of tensor chunk_tensor[c]
__global__ void __launch_bounds__(kAdamThreads)
adam_multi_kernel(float* const* __restrict__ params, const float* const* __restrict__ grads,
                  float* const* __restrict__ exp_avg, float* const* __restrict__ exp_avg_sq,
                  const int64_t* __restrict__ numel, const int32_t* __restrict__ chunk_tensor,
                  const int64_t* __restrict__ chunk_start, int64_t chunk_elems, AdamScalars s,
                  const float* __restrict__ step, const float* __restrict__ lr_dev,
                  const float* __restrict__ inv_scale_dev, const float* __restrict__ found_inf) {
  if (found_inf != nullptr && *found_inf != 0.f) return;  // skipped step: parameters and moments untouched
  __shared__ float sh[2];
  if (threadIdx.x == 0) {
    const double t = static_cast<double>(*step);
    const double bc1 = 1.0 - pow(static_cast<double>(s.beta1), t);
    const double bc2 = 1.0 - pow(static_cast<double>(s.beta2), t);
    const float lr = lr_dev != nullptr ? *lr_dev : s.lr;
    sh[0] = static_cast<float>(static_cast<double>(lr) / bc1);
    sh[1] = static_cast<float>(1.0 / sqrt(bc2));
  }
  __syncthreads();
  const float step_size = sh[0], inv_bc2_sqrt = sh[1];
  if (lr_dev != nullptr) s.lr = *lr_dev;
  const float inv_scale = inv_scale_dev != nullptr ? *inv_scale_dev : 1.f;

  const int t = chunk_tensor[blockIdx.x];
  const int64_t start = chunk_start[blockIdx.x];
  const int64_t n = numel[t];
  const int64_t end = start + chunk_elems < n ? start + chunk_elems : n;
  float* p = params[t];
  const float* g = grads[t];
  float* m = exp_avg[t];
  float* v = exp_avg_sq[t];
  if (g == nullptr) return;  // parameter without a gradient this step

  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v)) & 15) == 0;  // chunk starts are multiples of 4 elements
  int64_t i = start;
  if (vec) {
    const int64_t end4 = start + ((end - start) & ~int64_t(3));
    for (i = start + 4 * threadIdx.x; i < end4; i += 4 * kAdamThreads) {
      float4 pv = *reinterpret_cast<float4*>(p + i);
      const float4 gv = *reinterpret_cast<const float4*>(g + i);
      float4 mv = *reinterpret_cast<float4*>(m + i);
      float4 vv = *reinterpret_cast<float4*>(v + i);
      adam_update(pv.x, gv.x, mv.x, vv.x, s, step_size, inv_bc2_sqrt, inv_scale);
      adam_update(pv.y, gv.y, mv.y, vv.y, s, step_size, inv_bc2_sqrt, inv_scale);
      adam_update(pv.z, gv.z, mv.z, vv.z, s, step_size, inv_bc2_sqrt, inv_scale);
      adam_update(pv.w, gv.w, mv.w, vv.w, s, step_size, inv_bc2_sqrt, inv_scale);
      *reinterpret_cast<float4*>(p + i) = pv;
      *reinterpret_cast<float4*>(m + i) = mv;
      *reinterpret_cast<float4*>(v + i) = vv;
    }
    i = end4;
  }
  for (i += threadIdx.x; i < end; i += kAdamThreads) {
    float pv = p[i], mv = m[i], vv = v[i];
    adam_update(pv, g[i], mv, vv, s, step_size, inv_bc2_sqrt, inv_scale);
    p[i] = pv;
    m[i] = mv;
    v[i] = vv;
  }
}

// one thread per edge row: out[e] = (pos[src[e]] - pos[dst[e]], |.|), then (x - mu) / std per column
template <int DIM>
__global__ void edge_features_kernel(const float* __restrict__ pos, const int32_t* __restrict__ src,
                                     const int32_t* __restrict__ dst, int64_t n_edges,
                                     const float* __restrict__ mu, const float* __restrict__ sd,
                                     float* __restrict__ out) {
  const int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= n_edges) return;
  const float* a = pos + int64_t(src[e]) * DIM;
  const float* b = pos + int64_t(dst[e]) * DIM;
  float d[DIM + 1];
  float ss = 0.f;
#pragma unroll
  for (int k = 0; k < DIM; ++k) {
    d[k] = a[k] - b[k];
    ss += d[k] * d[k];
  }
  d[DIM] = sqrtf(ss);
#pragma unroll
  for (int k = 0; k <= DIM; ++k) {
    if (mu != nullptr) d[k] -= mu[k];
    if (sd != nullptr) d[k] /= sd[k];
  }
  if constexpr (DIM == 3) {  // a 3-D row is exactly 16 bytes: one 128-bit store per edge, fully coalesced
    *reinterpret_cast<float4*>(out + e * 4) = make_float4(d[0], d[1], d[2], d[3]);
  } else {
#pragma unroll
    for (int k = 0; k <= DIM; ++k) out[e * (DIM + 1) + k] = d[k];
  }
}

}  // namespace
}  // namespace mgn

using namespace mgn;

extern "C" int mgn_adam_multi_step(void* const* params, const void* const* grads, void* const* exp_avg,
                                   void* const* exp_avg_sq, const int64_t* numel, const int32_t* chunk_tensor,
                                   const int64_t* chunk_start, int64_t n_chunks, int64_t chunk_elems, float lr,
                                   float beta1, float beta2, float eps, float weight_decay, int adamw, float* step,
                                   const float* lr_dev, const float* inv_scale, const float* found_inf,
                                   mgn_stream_t stream) {
  MGN_CHECK_ARG(n_chunks >= 0 && chunk_elems > 0 && (chunk_elems & 3) == 0 && step != nullptr);
  MGN_CHECK_ARG(n_chunks < (int64_t(1) << 31));
  cudaStream_t st = as_stream(stream);
  adam_tick_kernel<<<1, 1, 0, MGN_ST(st)>>>(step, found_inf);
  if (n_chunks == 0) return mgn_launch_status();
  MGN_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && numel && chunk_tensor && chunk_start);
  AdamScalars s{lr, beta1, beta2, eps, weight_decay, adamw};
  adam_multi_kernel<<<static_cast<unsigned>(n_chunks), kAdamThreads, 0, MGN_ST(st)>>>(
      reinterpret_cast<float* const*>(params), reinterpret_cast<const float* const*>(grads),
      reinterpret_cast<float* const*>(exp_avg), reinterpret_cast<float* const*>(exp_avg_sq), numel, chunk_tensor,
      chunk_start, chunk_elems, s, step, lr_dev, inv_scale, found_inf);
  return mgn_launch_status();
}

extern "C" int mgn_edge_features(const float* pos, int dim, const int32_t* src, const int32_t* dst, int64_t n_edges,
                                 const float* mu, const float* sd, float* out, mgn_stream_t stream) {
  MGN_CHECK_ARG(n_edges >= 0 && (dim == 2 || dim == 3));
  if (n_edges == 0) return MGN_OK;
  MGN_CHECK_ARG(pos && src && dst && out);
  if (dim == 3 && (reinterpret_cast<uintptr_t>(out) & 15) != 0) return MGN_EALIGN;
  cudaStream_t st = as_stream(stream);
  const unsigned grid = static_cast<unsigned>((n_edges + 255) / 256);
  if (dim == 2)
    edge_features_kernel<2><<<grid, 256, 0, MGN_ST(st)>>>(pos, src, dst, n_edges, mu, sd, out);
  else
    edge_features_kernel<3><<<grid, 256, 0, MGN_ST(st)>>>(pos, src, dst, n_edges, mu, sd, out);
  return mgn_launch_status();
}
