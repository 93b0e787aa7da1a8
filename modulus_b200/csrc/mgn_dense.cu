// mgn_dense.cu — fp32-accurate SIMT implementation of the dense pieces of MeshGraphMLP
// (Linear / activation / LayerNorm, forward and backward).  This is the path fp32 models run
// on (the reference tests disable TF32, test_meshgraphnet_snmg.py:31, so the 1e-4 parity
// bar needs true fp32 FMA accumulation) and the on-device cross-check for the tcgen05 path.
// Activations may be fp32 or bf16 (fp32 accumulate); parameters are fp32, read in place.
#include "mgn_common.cuh"

namespace mgn {

// ---------------------------------------------------------------------------------------
// generic strided SIMT GEMM:  C[i,j] = sum_r A(i,r) * B(r,j)
//   A(i,r) = A[i*sa_i + r*sa_r]   B(r,j) = B[r*sb_r + j*sb_j]
// 64x64x16 tiles, 256 threads, 4x4 outputs per thread, optional split over r (blockIdx.z)
// ---------------------------------------------------------------------------------------
constexpr int GB_M = 64, GB_N = 64, GB_K = 16;

template <typename TA, typename TB, typename TC>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const TA* __restrict__ A, int64_t sa_i, int64_t sa_r, const TB* __restrict__ B, int64_t sb_r,
                 int64_t sb_j, int64_t I, int64_t J, int64_t R, int64_t r_chunk, TC* __restrict__ acc_out,
                 int64_t ldc) {
  __shared__ float As[GB_K][GB_M + 4];
  __shared__ float Bs[GB_K][GB_N + 4];
  const int t = threadIdx.x;
  const int ti = t / 16, tj = t % 16;  // 16x16 threads, 4x4 each
  // grid.x is a linear tile id, column tile fastest: CTAs sharing a row tile of A are adjacent, and the row-tile
  // count is not bound by the 65 535 limit of grid.y (6 M-edge tables have 94 k row tiles)
  const int64_t nj = (J + GB_N - 1) / GB_N;
  const int64_t i0 = (static_cast<int64_t>(blockIdx.x) / nj) * GB_M;
  const int64_t j0 = (static_cast<int64_t>(blockIdx.x) % nj) * GB_N;
  const int64_t r_begin = static_cast<int64_t>(blockIdx.z) * r_chunk;
  const int64_t r_end = min(R, r_begin + r_chunk);
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;

  const bool a_r_contig = (sa_r == 1);
  const bool b_j_contig = (sb_j == 1);
  // register-staged double buffering: the global loads of step r0 + GB_K are issued before the FMAs of step r0, so
  // their latency overlaps the arithmetic instead of being exposed once per step
  float pa[4], pb[4];
  auto gload = [&](int64_t r0) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int idx = t + u * 256;
      int ii, rr;
      if (a_r_contig) { rr = idx % GB_K; ii = idx / GB_K; } else { ii = idx % GB_M; rr = idx / GB_M; }
      const int64_t gi = i0 + ii, gr = r0 + rr;
      pa[u] = (gi < I && gr < r_end) ? Num<TA>::to_f(A[gi * sa_i + gr * sa_r]) : 0.f;
      int jj, rb;
      if (b_j_contig) { jj = idx % GB_N; rb = idx / GB_N; } else { rb = idx % GB_K; jj = idx / GB_K; }
      const int64_t gj = j0 + jj, grb = r0 + rb;
      pb[u] = (gj < J && grb < r_end) ? Num<TB>::to_f(B[grb * sb_r + gj * sb_j]) : 0.f;
    }
  };
  if (r_begin < r_end) gload(r_begin);
  for (int64_t r0 = r_begin; r0 < r_end; r0 += GB_K) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int idx = t + u * 256;
      int ii, rr;
      if (a_r_contig) { rr = idx % GB_K; ii = idx / GB_K; } else { ii = idx % GB_M; rr = idx / GB_M; }
      As[rr][ii] = pa[u];
      int jj, rb;
      if (b_j_contig) { jj = idx % GB_N; rb = idx / GB_N; } else { rb = idx % GB_K; jj = idx / GB_K; }
      Bs[rb][jj] = pb[u];
    }
    __syncthreads();
    if (r0 + GB_K < r_end) gload(r0 + GB_K);
#pragma unroll
    for (int k = 0; k < GB_K; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) a[u] = As[k][ti * 4 + u];
#pragma unroll
      for (int u = 0; u < 4; ++u) b[u] = Bs[k][tj * 4 + u];
#pragma unroll
      for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) acc[x][y] = fmaf(a[x], b[y], acc[x][y]);
    }
    __syncthreads();
  }
  TC* out = acc_out + static_cast<int64_t>(blockIdx.z) * I * ldc;
#pragma unroll
  for (int x = 0; x < 4; ++x) {
    const int64_t gi = i0 + ti * 4 + x;
    if (gi >= I) continue;
#pragma unroll
    for (int y = 0; y < 4; ++y) {
      const int64_t gj = j0 + tj * 4 + y;
      if (gj < J) out[gi * ldc + gj] = Num<TC>::from_f(acc[x][y]);
    }
  }
}

// linear forward fused variant: same mainloop, epilogue adds bias, stores pre-activation (optional)
// and activation in the activation dtype
template <typename T>
__global__ void __launch_bounds__(256)
linear_fwd_kernel(const T* __restrict__ X, int64_t ldx, const float* __restrict__ W, const float* __restrict__ bias,
                  int64_t M, int64_t N, int64_t K, int act, T* __restrict__ Ypre, T* __restrict__ H, int64_t ldh) {
  __shared__ float As[GB_K][GB_M + 4];
  __shared__ float Bs[GB_K][GB_N + 4];
  const int t = threadIdx.x;
  const int ti = t / 16, tj = t % 16;
  const int64_t nj = (N + GB_N - 1) / GB_N;  // linear tile id on grid.x, column tile fastest (see gemm_simt_kernel)
  const int64_t i0 = (static_cast<int64_t>(blockIdx.x) / nj) * GB_M;
  const int64_t j0 = (static_cast<int64_t>(blockIdx.x) % nj) * GB_N;
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  float pa[4], pb[4];  // next step's operands, loaded while this step's FMAs run
  auto gload = [&](int64_t r0) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int idx = t + u * 256;
      const int rr = idx % GB_K, ii = idx / GB_K;
      const int64_t gi = i0 + ii, gr = r0 + rr;
      pa[u] = (gi < M && gr < K) ? Num<T>::to_f(X[gi * ldx + gr]) : 0.f;
      const int64_t gj = j0 + ii;
      pb[u] = (gj < N && gr < K) ? W[gj * K + gr] : 0.f;
    }
  };
  if (K > 0) gload(0);
  for (int64_t r0 = 0; r0 < K; r0 += GB_K) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int idx = t + u * 256;
      const int rr = idx % GB_K, ii = idx / GB_K;
      As[rr][ii] = pa[u];
      Bs[rr][ii] = pb[u];
    }
    __syncthreads();
    if (r0 + GB_K < K) gload(r0 + GB_K);
#pragma unroll
    for (int k = 0; k < GB_K; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) a[u] = As[k][ti * 4 + u];
#pragma unroll
      for (int u = 0; u < 4; ++u) b[u] = Bs[k][tj * 4 + u];
#pragma unroll
      for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) acc[x][y] = fmaf(a[x], b[y], acc[x][y]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int x = 0; x < 4; ++x) {
    const int64_t gi = i0 + ti * 4 + x;
    if (gi >= M) continue;
#pragma unroll
    for (int y = 0; y < 4; ++y) {
      const int64_t gj = j0 + tj * 4 + y;
      if (gj >= N) continue;
      float v = acc[x][y] + (bias ? bias[gj] : 0.f);
      if (Ypre) Ypre[gi * N + gj] = Num<T>::from_f(v);
      if (H) H[gi * ldh + gj] = Num<T>::from_f(act_fwd(act, v));
    }
  }
}

// out[i] = sum_s partial[s][i]  (fixed order), optional cast
template <typename TO>
__global__ void reduce_partials_kernel(const float* __restrict__ partial, int64_t n, int splits, TO* __restrict__ out) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    float s = 0.f;
    for (int k = 0; k < splits; ++k) s += partial[static_cast<int64_t>(k) * n + i];
    out[i] = Num<TO>::from_f(s);
  }
}

// column sums of a [M,N] matrix, stage 1: block b sums rows [b*rows_per_block, ...)
template <typename T>
__global__ void colsum_partial_kernel(const T* __restrict__ X, int64_t M, int64_t N, int64_t rows_per_block,
                                      float* __restrict__ partial) {
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * rows_per_block;
  const int64_t r1 = min(M, r0 + rows_per_block);
  for (int64_t c = threadIdx.x; c < N; c += blockDim.x) {
    float s = 0.f;
    for (int64_t r = r0; r < r1; ++r) s += Num<T>::to_f(X[r * N + c]);
    partial[static_cast<int64_t>(blockIdx.x) * N + c] = s;
  }
}

template <typename T>
__global__ void act_bwd_kernel(const T* __restrict__ g_h, const T* __restrict__ y_pre, int act, T* __restrict__ g_y,
                               int64_t n) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    g_y[i] = Num<T>::from_f(Num<T>::to_f(g_h[i]) * act_grad(act, Num<T>::to_f(y_pre[i])));
}

template <typename T>
__global__ void act_fwd_kernel(const T* __restrict__ x, int act, T* __restrict__ y, int64_t n) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    y[i] = Num<T>::from_f(act_fwd(act, Num<T>::to_f(x[i])));
}

template <typename T>
__global__ void add_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out, int64_t n) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = Num<T>::from_f(Num<T>::to_f(a[i]) + Num<T>::to_f(b[i]));
}

// ---------------------------------------------------------------------------------------
// LayerNorm: warp per row, fp32 statistics, biased variance, eps inside the sqrt
// (torch.nn.LayerNorm semantics, mesh_graph_mlp.py:165-166)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename T>
__global__ void __launch_bounds__(256)
layernorm_fwd_kernel(const T* __restrict__ X, int64_t M, int D, const float* __restrict__ gamma,
                     const float* __restrict__ beta, float eps, const T* __restrict__ residual, T* __restrict__ out,
                     float* __restrict__ mean_out, float* __restrict__ rstd_out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t r = warp; r < M; r += nwarps) {
    const T* x = X + r * D;
    float s = 0.f;
    for (int c = lane; c < D; c += 32) s += Num<T>::to_f(x[c]);
    const float mu = warp_sum(s) / D;
    float q = 0.f;
    for (int c = lane; c < D; c += 32) {
      const float d = Num<T>::to_f(x[c]) - mu;
      q += d * d;
    }
    const float rstd = rsqrtf(warp_sum(q) / D + eps);
    for (int c = lane; c < D; c += 32) {
      float v = (Num<T>::to_f(x[c]) - mu) * rstd;
      v = v * (gamma ? gamma[c] : 1.f) + (beta ? beta[c] : 0.f);
      if (residual) v += Num<T>::to_f(residual[r * D + c]);
      out[r * D + c] = Num<T>::from_f(v);
    }
    if (lane == 0) {
      if (mean_out) mean_out[r] = mu;
      if (rstd_out) rstd_out[r] = rstd;
    }
  }
}

constexpr int LN_MAXC = 32;  // D <= 1024

// stage 1: g_x per row + per-block partial sums of d_gamma / d_beta (fixed order)
template <typename T>
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const T* __restrict__ G, const T* __restrict__ X, const float* __restrict__ mean,
                     const float* __restrict__ rstd, const float* __restrict__ gamma, int64_t M, int D,
                     int64_t rows_per_block, T* __restrict__ GX, float* __restrict__ partial) {
  __shared__ float red[8][64];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * rows_per_block;
  const int64_t r1 = min(M, r0 + rows_per_block);
  float dg[LN_MAXC], db[LN_MAXC];
#pragma unroll
  for (int k = 0; k < LN_MAXC; ++k) dg[k] = db[k] = 0.f;
  for (int64_t r = r0 + w; r < r1; r += 8) {
    const float mu = mean[r], rs = rstd[r];
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int k = 0; k < LN_MAXC; ++k) {
      const int c = lane + 32 * k;
      if (c < D) {
        const float g = Num<T>::to_f(G[r * D + c]);
        const float xh = (Num<T>::to_f(X[r * D + c]) - mu) * rs;
        const float dxh = g * (gamma ? gamma[c] : 1.f);
        c1 += dxh;
        c2 += dxh * xh;
        dg[k] += g * xh;
        db[k] += g;
      }
    }
    c1 = warp_sum(c1) / D;
    c2 = warp_sum(c2) / D;
#pragma unroll
    for (int k = 0; k < LN_MAXC; ++k) {
      const int c = lane + 32 * k;
      if (c < D) {
        const float g = Num<T>::to_f(G[r * D + c]);
        const float xh = (Num<T>::to_f(X[r * D + c]) - mu) * rs;
        const float dxh = g * (gamma ? gamma[c] : 1.f);
        GX[r * D + c] = Num<T>::from_f(rs * (dxh - c1 - xh * c2));
      }
    }
  }
  // block reduction over the 8 warps, column block by column block
  float* pg = partial + static_cast<int64_t>(blockIdx.x) * 2 * D;
  for (int k = 0; k < LN_MAXC; ++k) {
    if (32 * k >= D) break;
    red[w][lane] = dg[k];
    red[w][32 + lane] = db[k];
    __syncthreads();
    if (w == 0) {
      float sg = 0.f, sb = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        sg += red[q][lane];
        sb += red[q][32 + lane];
      }
      const int c = lane + 32 * k;
      if (c < D) {
        pg[c] = sg;
        pg[D + c] = sb;
      }
    }
    __syncthreads();
  }
}

// final stage of the LayerNorm parameter gradients: sum block partials [nb][gamma(D)|beta(D)]
__global__ void ln_reduce2_kernel(const float* __restrict__ partial, int64_t D, int nb, float* __restrict__ gg,
                                  float* __restrict__ gb) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= 2 * D) return;
  float s = 0.f;
  for (int k = 0; k < nb; ++k) s += partial[static_cast<int64_t>(k) * 2 * D + i];
  if (i < D) gg[i] = s; else gb[i - D] = s;
}

static inline int ew_grid(int64_t n) {
  int64_t b = (n + 255) / 256;
  const int64_t cap = static_cast<int64_t>(num_sms()) * 16;
  return static_cast<int>(b < 1 ? 1 : (b > cap ? cap : b));
}

template <typename T>
static int linear_fwd_t(const void* x, int64_t ldx, int64_t M, int64_t K, const float* w, const float* b, int64_t N,
                        int act, void* y_pre, void* h, int64_t ldh, cudaStream_t st) {
  if (M == 0 || N == 0) return MGN_OK;
  const int64_t n_tiles = ((N + GB_N - 1) / GB_N) * ((M + GB_M - 1) / GB_M);
  if (n_tiles > 0x7fffffffLL) return MGN_EUNSUPPORTED;
  dim3 grid(static_cast<unsigned>(n_tiles), 1, 1);
  linear_fwd_kernel<T><<<grid, 256, 0, MGN_ST(st)>>>(static_cast<const T*>(x), ldx, w, b, M, N, K, act,
                                             static_cast<T*>(y_pre), static_cast<T*>(h), ldh);
  return mgn_launch_status();
}

}  // namespace mgn

using namespace mgn;

extern "C" int mgn_linear_fwd(int dtype, const void* x, int64_t ldx, int64_t M, int64_t K, const float* w,
                              const float* b, int64_t N, int act, void* y_pre, void* h, int64_t ldh,
                              mgn_stream_t stream) {
  MGN_CHECK_ARG(M >= 0 && K >= 0 && N >= 0 && ldx >= K && ldh >= N);
  if (M == 0) return MGN_OK;
  MGN_CHECK_ARG(x && w && (y_pre || h));
  MGN_CHECK_ARG(act >= MGN_ACT_NONE && act <= MGN_ACT_ELU);
  if (dtype == MGN_F32) return linear_fwd_t<float>(x, ldx, M, K, w, b, N, act, y_pre, h, ldh, as_stream(stream));
  if (dtype == MGN_BF16) return linear_fwd_t<bf16>(x, ldx, M, K, w, b, N, act, y_pre, h, ldh, as_stream(stream));
  return MGN_EINVAL;
}

extern "C" int mgn_act_bwd(int dtype, const void* g_h, const void* y_pre, int act, void* g_y, int64_t n,
                           mgn_stream_t stream) {
  MGN_CHECK_ARG(n >= 0);
  if (n == 0) return MGN_OK;
  MGN_CHECK_ARG(g_h && y_pre && g_y && act >= MGN_ACT_NONE && act <= MGN_ACT_ELU);
  cudaStream_t st = as_stream(stream);
  if (dtype == MGN_F32)
    act_bwd_kernel<float><<<ew_grid(n), 256, 0, MGN_ST(st)>>>(static_cast<const float*>(g_h), static_cast<const float*>(y_pre),
                                                      act, static_cast<float*>(g_y), n);
  else if (dtype == MGN_BF16)
    act_bwd_kernel<bf16><<<ew_grid(n), 256, 0, MGN_ST(st)>>>(static_cast<const bf16*>(g_h), static_cast<const bf16*>(y_pre),
                                                     act, static_cast<bf16*>(g_y), n);
  else
    return MGN_EINVAL;
  return mgn_launch_status();
}

extern "C" int mgn_act_fwd(int dtype, const void* x, int act, void* y, int64_t n, mgn_stream_t stream) {
  MGN_CHECK_ARG(n >= 0);
  if (n == 0) return MGN_OK;
  MGN_CHECK_ARG(x && y && act >= MGN_ACT_NONE && act <= MGN_ACT_ELU);
  cudaStream_t st = as_stream(stream);
  if (dtype == MGN_F32)
    act_fwd_kernel<float><<<ew_grid(n), 256, 0, MGN_ST(st)>>>(static_cast<const float*>(x), act, static_cast<float*>(y), n);
  else if (dtype == MGN_BF16)
    act_fwd_kernel<bf16><<<ew_grid(n), 256, 0, MGN_ST(st)>>>(static_cast<const bf16*>(x), act, static_cast<bf16*>(y), n);
  else
    return MGN_EINVAL;
  return mgn_launch_status();
}

extern "C" int mgn_add(int dtype, const void* a, const void* b, void* out, int64_t n, mgn_stream_t stream) {
  MGN_CHECK_ARG(n >= 0);
  if (n == 0) return MGN_OK;
  MGN_CHECK_ARG(a && b && out);
  cudaStream_t st = as_stream(stream);
  if (dtype == MGN_F32)
    add_kernel<float><<<ew_grid(n), 256, 0, MGN_ST(st)>>>(static_cast<const float*>(a), static_cast<const float*>(b),
                                                  static_cast<float*>(out), n);
  else if (dtype == MGN_BF16)
    add_kernel<bf16><<<ew_grid(n), 256, 0, MGN_ST(st)>>>(static_cast<const bf16*>(a), static_cast<const bf16*>(b),
                                                 static_cast<bf16*>(out), n);
  else
    return MGN_EINVAL;
  return mgn_launch_status();
}

// g_x[M,K] = g_y[M,N] * W[N,K]
extern "C" int mgn_linear_bwd_data(int dtype, const void* g_y, int64_t M, int64_t N, const float* w, int64_t K,
                                   void* g_x, int64_t ldgx, mgn_stream_t stream) {
  MGN_CHECK_ARG(M >= 0 && N >= 0 && K >= 0 && ldgx >= K);
  if (M == 0 || K == 0) return MGN_OK;
  MGN_CHECK_ARG(g_y && w && g_x);
  cudaStream_t st = as_stream(stream);
  const int64_t n_tiles = ((K + GB_N - 1) / GB_N) * ((M + GB_M - 1) / GB_M);
  if (n_tiles > 0x7fffffffLL) return MGN_EUNSUPPORTED;
  dim3 grid(static_cast<unsigned>(n_tiles), 1, 1);
  // C[m,k] = sum_n g_y[m,n] W[n,k]:  A(i=m, r=n) = g_y[m*N+n],  B(r=n, j=k) = W[n*K+k]
  if (dtype == MGN_F32)
    gemm_simt_kernel<float, float, float><<<grid, 256, 0, MGN_ST(st)>>>(static_cast<const float*>(g_y), N, 1, w, K, 1, M, K,
                                                                N, N, static_cast<float*>(g_x), ldgx);
  else if (dtype == MGN_BF16)
    gemm_simt_kernel<bf16, float, bf16><<<grid, 256, 0, MGN_ST(st)>>>(static_cast<const bf16*>(g_y), N, 1, w, K, 1, M, K, N,
                                                              N, static_cast<bf16*>(g_x), ldgx);
  else
    return MGN_EINVAL;
  return mgn_launch_status();
}

// Split count of the weight-gradient reduction over the M rows.  The mainloop is not software-pipelined, so the
// latency of each 16-row step is hidden by other CTAs only: aim at ~1024 CTAs in flight (about 7 per SM) whatever
// the [N, K] tile count, with at least 64 rows per CTA.  (One split per 4096 rows left a 10k-edge mesh with 36 CTAs
// on 148 SMs: 0.55 ms per call, 81 % of the fp32 step at the vortex-shedding size.)
static int64_t wgrad_splits(int64_t M, int64_t N, int64_t K) {
  const int64_t tiles = ((K + GB_N - 1) / GB_N) * ((N + GB_M - 1) / GB_M);
  int64_t splits = (1024 + tiles - 1) / (tiles > 0 ? tiles : 1);
  const int64_t by_rows = (M + 63) / 64;
  if (splits > by_rows) splits = by_rows;
  if (splits > 256) splits = 256;
  if (splits < 1) splits = 1;
  return splits;
}

extern "C" size_t mgn_linear_bwd_weight_workspace_bytes(int64_t M, int64_t N, int64_t K) {
  return static_cast<size_t>(wgrad_splits(M, N, K) * (N * K + N) * sizeof(float));
}

extern "C" int mgn_linear_bwd_weight(int dtype, const void* g_y, const void* x, int64_t ldx, int64_t M, int64_t N,
                                     int64_t K, float* g_w, float* g_b, void* workspace, size_t workspace_bytes,
                                     mgn_stream_t stream) {
  MGN_CHECK_ARG(M >= 0 && N >= 0 && K >= 0 && ldx >= K);
  MGN_CHECK_ARG(g_w && workspace);
  if (workspace_bytes < mgn_linear_bwd_weight_workspace_bytes(M, N, K)) return MGN_EWORKSPACE;
  cudaStream_t st = as_stream(stream);
  if (M == 0) {
    cudaMemsetAsync(g_w, 0, N * K * sizeof(float), st);
    if (g_b) cudaMemsetAsync(g_b, 0, N * sizeof(float), st);
    return mgn_launch_status();
  }
  MGN_CHECK_ARG(g_y && x);
  int64_t splits = wgrad_splits(M, N, K);
  const int64_t r_chunk = ((M + splits - 1) / splits + GB_K - 1) / GB_K * GB_K;
  splits = (M + r_chunk - 1) / r_chunk;
  float* part_w = static_cast<float*>(workspace);
  float* part_b = part_w + splits * N * K;
  dim3 grid(static_cast<unsigned>(((K + GB_N - 1) / GB_N) * ((N + GB_M - 1) / GB_M)), 1, static_cast<unsigned>(splits));
  // C[n,k] = sum_m g_y[m,n] x[m,k]:  A(i=n, r=m) = g_y[m*N+n],  B(r=m, j=k) = x[m*ldx+k]
  if (dtype == MGN_F32)
    gemm_simt_kernel<float, float, float><<<grid, 256, 0, MGN_ST(st)>>>(static_cast<const float*>(g_y), 1, N,
                                                         static_cast<const float*>(x), ldx, 1, N, K, M, r_chunk,
                                                         part_w, K);
  else if (dtype == MGN_BF16)
    gemm_simt_kernel<bf16, bf16, float><<<grid, 256, 0, MGN_ST(st)>>>(static_cast<const bf16*>(g_y), 1, N,
                                                       static_cast<const bf16*>(x), ldx, 1, N, K, M, r_chunk, part_w,
                                                       K);
  else
    return MGN_EINVAL;
  reduce_partials_kernel<float><<<ew_grid(N * K), 256, 0, MGN_ST(st)>>>(part_w, N * K, static_cast<int>(splits), g_w);
  if (g_b) {
    if (dtype == MGN_F32)
      colsum_partial_kernel<float><<<static_cast<unsigned>(splits), 128, 0, MGN_ST(st)>>>(static_cast<const float*>(g_y), M, N,
                                                                                  r_chunk, part_b);
    else
      colsum_partial_kernel<bf16><<<static_cast<unsigned>(splits), 128, 0, MGN_ST(st)>>>(static_cast<const bf16*>(g_y), M, N,
                                                                                 r_chunk, part_b);
    reduce_partials_kernel<float><<<ew_grid(N), 256, 0, MGN_ST(st)>>>(part_b, N, static_cast<int>(splits), g_b);
  }
  return mgn_launch_status();
}

extern "C" int mgn_layernorm_fwd(int dtype, const void* x, int64_t M, int64_t D, const float* gamma,
                                 const float* beta, float eps, const void* residual, void* out, float* mean,
                                 float* rstd, mgn_stream_t stream) {
  MGN_CHECK_ARG(M >= 0 && D > 0);
  if (M == 0) return MGN_OK;
  MGN_CHECK_ARG(x && out);
  cudaStream_t st = as_stream(stream);
  const int grid = ew_grid(M * 32);
  if (dtype == MGN_F32)
    layernorm_fwd_kernel<float><<<grid, 256, 0, MGN_ST(st)>>>(static_cast<const float*>(x), M, static_cast<int>(D), gamma,
                                                      beta, eps, static_cast<const float*>(residual),
                                                      static_cast<float*>(out), mean, rstd);
  else if (dtype == MGN_BF16)
    layernorm_fwd_kernel<bf16><<<grid, 256, 0, MGN_ST(st)>>>(static_cast<const bf16*>(x), M, static_cast<int>(D), gamma, beta,
                                                     eps, static_cast<const bf16*>(residual), static_cast<bf16*>(out),
                                                     mean, rstd);
  else
    return MGN_EINVAL;
  return mgn_launch_status();
}

// 32 rows per block (4 per warp) until the grid reaches 4 blocks per SM: each row is a dependent load -> warp
// reduction -> store chain, so small inputs need many short blocks (256 rows per block cost 0.12 ms at 11k rows)
static inline int64_t ln_bwd_blocks(int64_t M) {
  int64_t nb = (M + 31) / 32;
  const int64_t cap = static_cast<int64_t>(num_sms()) * 4;
  if (nb > cap) nb = cap;
  if (nb < 1) nb = 1;
  return nb;
}

extern "C" size_t mgn_layernorm_bwd_workspace_bytes(int64_t M, int64_t D) {
  // sized for the largest grid this library would ever pick (independent of the device)
  int64_t nb = (M + 31) / 32;
  if (nb > 4096) nb = 4096;
  if (nb < 1) nb = 1;
  return static_cast<size_t>(nb * 2 * D * sizeof(float));
}

extern "C" int mgn_layernorm_bwd(int dtype, const void* g_out, const void* x, const float* mean, const float* rstd,
                                 const float* gamma, int64_t M, int64_t D, void* g_x, float* g_gamma, float* g_beta,
                                 void* workspace, size_t workspace_bytes, mgn_stream_t stream) {
  MGN_CHECK_ARG(M >= 0 && D > 0);
  if (D > 32 * LN_MAXC) return MGN_EUNSUPPORTED;
  MGN_CHECK_ARG(g_gamma && g_beta && workspace);
  cudaStream_t st = as_stream(stream);
  if (M == 0) {
    cudaMemsetAsync(g_gamma, 0, D * sizeof(float), st);
    cudaMemsetAsync(g_beta, 0, D * sizeof(float), st);
    return mgn_launch_status();
  }
  MGN_CHECK_ARG(g_out && x && mean && rstd && g_x);
  int64_t nb = ln_bwd_blocks(M);
  if (static_cast<size_t>(nb * 2 * D * sizeof(float)) > workspace_bytes) return MGN_EWORKSPACE;
  const int64_t rows_per_block = (M + nb - 1) / nb;
  nb = (M + rows_per_block - 1) / rows_per_block;
  float* partial = static_cast<float*>(workspace);
  if (dtype == MGN_F32)
    layernorm_bwd_kernel<float><<<static_cast<unsigned>(nb), 256, 0, MGN_ST(st)>>>(
        static_cast<const float*>(g_out), static_cast<const float*>(x), mean, rstd, gamma, M, static_cast<int>(D),
        rows_per_block, static_cast<float*>(g_x), partial);
  else if (dtype == MGN_BF16)
    layernorm_bwd_kernel<bf16><<<static_cast<unsigned>(nb), 256, 0, MGN_ST(st)>>>(
        static_cast<const bf16*>(g_out), static_cast<const bf16*>(x), mean, rstd, gamma, M, static_cast<int>(D),
        rows_per_block, static_cast<bf16*>(g_x), partial);
  else
    return MGN_EINVAL;
  ln_reduce2_kernel<<<static_cast<unsigned>((2 * D + 255) / 256), 256, 0, MGN_ST(st)>>>(partial, D, static_cast<int>(nb), g_gamma,
                                                                             g_beta);
  return mgn_launch_status();
}
