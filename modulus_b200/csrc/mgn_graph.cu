// mgn_graph.cu — graph plan construction on the device: CSC -> (csc_dst, CSR).
//
// The CSR lists, for every source node, the CSC positions of its out-edges in ascending
// order.  Ascending order makes the result unique (bit-exact against a stable argsort by
// source, oracle/mgn_oracle.py:csr_from_csc) even though the fill uses integer atomics.
#include "mgn_common.cuh"

namespace mgn {

constexpr int kScanBlock = 1024;

// one warp per segment: out[j] = s for j in [offsets[s], offsets[s+1])
__global__ void expand_offsets_kernel(const int32_t* __restrict__ offsets, int64_t n_seg,
                                      int32_t* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t v = warp; v < n_seg; v += nwarps) {
    const int32_t b = offsets[v], e = offsets[v + 1];
    for (int32_t j = b + lane; j < e; j += 32) out[j] = static_cast<int32_t>(v);
  }
}

// histogram of keys (integer atomics: the result does not depend on the order)
__global__ void count_keys_kernel(const int32_t* __restrict__ keys, int64_t n, int32_t* __restrict__ count) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < n; e += stride)
    atomicAdd(&count[keys[e]], 1);
}

// exclusive scan, phase 1: per-block scan of kScanBlock elements (in place), block total out
__global__ void scan_blocks_kernel(int32_t* __restrict__ data, int64_t n, int32_t* __restrict__ block_sums) {
  __shared__ int32_t warp_tot[32];
  const int64_t i = static_cast<int64_t>(blockIdx.x) * kScanBlock + threadIdx.x;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int32_t v = i < n ? data[i] : 0;
  int32_t x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) warp_tot[w] = x;
  __syncthreads();
  if (w == 0) {
    int32_t t = warp_tot[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int32_t y = __shfl_up_sync(0xffffffffu, t, o);
      if (lane >= o) t += y;
    }
    warp_tot[lane] = t;
  }
  __syncthreads();
  const int32_t prefix = (w > 0 ? warp_tot[w - 1] : 0) + x - v;  // exclusive
  if (i < n) data[i] = prefix;
  if (threadIdx.x == kScanBlock - 1) block_sums[blockIdx.x] = prefix + v;
}

// phase 2: one block scans the block totals (exclusive, in place)
__global__ void scan_sums_kernel(int32_t* __restrict__ sums, int64_t nb) {
  __shared__ int32_t warp_tot[32];
  __shared__ int32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int64_t base = 0; base < nb; base += kScanBlock) {
    const int64_t i = base + threadIdx.x;
    int32_t v = i < nb ? sums[i] : 0;
    int32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) warp_tot[w] = x;
    __syncthreads();
    if (w == 0) {
      int32_t t = warp_tot[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int32_t y = __shfl_up_sync(0xffffffffu, t, o);
        if (lane >= o) t += y;
      }
      warp_tot[lane] = t;
    }
    __syncthreads();
    const int32_t c = carry;
    if (i < nb) sums[i] = c + (w > 0 ? warp_tot[w - 1] : 0) + x - v;
    __syncthreads();
    if (threadIdx.x == kScanBlock - 1) carry = c + warp_tot[31];
    __syncthreads();
  }
}

// phase 3: add block prefix
__global__ void scan_add_kernel(int32_t* __restrict__ data, int64_t n, const int32_t* __restrict__ sums) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * kScanBlock + threadIdx.x;
  if (i < n) data[i] += sums[blockIdx.x];
}

// fill CSR edge lists (arbitrary order inside a segment; sorted afterwards)
__global__ void csr_fill_kernel(const int32_t* __restrict__ indices, int64_t n_edges,
                                const int32_t* __restrict__ csr_offsets, int32_t* __restrict__ cursor,
                                int32_t* __restrict__ csr_eids) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < n_edges; e += stride) {
    const int32_t s = indices[e];
    const int32_t p = csr_offsets[s] + atomicAdd(&cursor[s], 1);
    csr_eids[p] = static_cast<int32_t>(e);
  }
}

constexpr int kShortSeg = 48;

// thread per source: insertion-sort short segments, queue long ones
__global__ void csr_sort_short_kernel(const int32_t* __restrict__ csr_offsets, int64_t n_src,
                                      int32_t* __restrict__ csr_eids, int32_t* __restrict__ long_list,
                                      int32_t* __restrict__ long_count) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t u = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; u < n_src; u += stride) {
    const int32_t b = csr_offsets[u], e = csr_offsets[u + 1];
    const int32_t len = e - b;
    if (len <= 1) continue;
    if (len > kShortSeg) {
      long_list[atomicAdd(long_count, 1)] = static_cast<int32_t>(u);
      continue;
    }
    int32_t* a = csr_eids + b;
    for (int32_t i = 1; i < len; ++i) {
      const int32_t key = a[i];
      int32_t j = i - 1;
      while (j >= 0 && a[j] > key) {
        a[j + 1] = a[j];
        --j;
      }
      a[j + 1] = key;
    }
  }
}

// block per long segment: bitonic sort in global memory (virtual padding to a power of two)
__global__ void csr_sort_long_kernel(const int32_t* __restrict__ csr_offsets, int32_t* __restrict__ csr_eids,
                                     const int32_t* __restrict__ long_list, const int32_t* __restrict__ long_count) {
  const int32_t n_long = *long_count;
  for (int32_t li = blockIdx.x; li < n_long; li += gridDim.x) {
    const int32_t u = long_list[li];
    const int32_t b = csr_offsets[u];
    const int32_t len = csr_offsets[u + 1] - b;
    int32_t* a = csr_eids + b;
    int32_t p2 = 1;
    while (p2 < len) p2 <<= 1;
    // ascending-comparator-only bitonic network (mirror step, then half-cleaners): slots
    // >= len behave as +inf, so comparators touching them are no-ops and can be skipped.
    for (int32_t k = 2; k <= p2; k <<= 1) {
      for (int32_t j = k >> 1; j > 0; j >>= 1) {
        const bool mirror = (j == (k >> 1));
        for (int32_t i = threadIdx.x; i < p2; i += blockDim.x) {
          const int32_t l = mirror ? (i ^ (k - 1)) : (i ^ j);
          if (l > i && l < len) {
            const int32_t x = a[i], y = a[l];
            if (x > y) {
              a[i] = y;
              a[l] = x;
            }
          }
        }
        __syncthreads();
      }
    }
  }
}

}  // namespace mgn

using namespace mgn;

extern "C" size_t mgn_group_by_key_workspace_bytes(int64_t n_keys) {
  const int64_t nb = (n_keys + 1 + kScanBlock - 1) / kScanBlock;
  return static_cast<size_t>((2 * (n_keys + 1) + nb + 64) * sizeof(int32_t));
}

extern "C" int mgn_group_by_key(const int32_t* keys, int64_t n, int64_t n_keys, int32_t* offsets, int32_t* ids,
                                void* workspace, size_t workspace_bytes, mgn_stream_t stream) {
  MGN_CHECK_ARG(offsets && workspace && n >= 0 && n_keys >= 0 && n < (int64_t(1) << 31));
  MGN_CHECK_ARG(n == 0 || (keys && ids));
  if (workspace_bytes < mgn_group_by_key_workspace_bytes(n_keys)) return MGN_EWORKSPACE;
  cudaStream_t st = as_stream(stream);
  const int64_t n1 = n_keys + 1;
  const int64_t nb = (n1 + kScanBlock - 1) / kScanBlock;
  int32_t* cursor = static_cast<int32_t*>(workspace);
  int32_t* long_list = cursor + n1;
  int32_t* block_sums = long_list + n1;
  int32_t* long_count = block_sums + nb + 1;
  cudaMemsetAsync(offsets, 0, n1 * sizeof(int32_t), st);
  cudaMemsetAsync(cursor, 0, n1 * sizeof(int32_t), st);
  cudaMemsetAsync(long_count, 0, sizeof(int32_t), st);
  const int grid = num_sms() * 8;
  if (n > 0) count_keys_kernel<<<grid, 256, 0, MGN_ST(st)>>>(keys, n, offsets);
  // offsets holds counts[0..n_keys) and 0 at [n_keys]; exclusive scan over n_keys+1 entries
  scan_blocks_kernel<<<static_cast<unsigned>(nb), kScanBlock, 0, MGN_ST(st)>>>(offsets, n1, block_sums);
  scan_sums_kernel<<<1, kScanBlock, 0, MGN_ST(st)>>>(block_sums, nb);
  scan_add_kernel<<<static_cast<unsigned>(nb), kScanBlock, 0, MGN_ST(st)>>>(offsets, n1, block_sums);
  if (n > 0) {
    csr_fill_kernel<<<grid, 256, 0, MGN_ST(st)>>>(keys, n, offsets, cursor, ids);
    csr_sort_short_kernel<<<grid, 256, 0, MGN_ST(st)>>>(offsets, n_keys, ids, long_list, long_count);
    csr_sort_long_kernel<<<num_sms() * 2, 256, 0, MGN_ST(st)>>>(offsets, ids, long_list, long_count);
  }
  return mgn_launch_status();
}

extern "C" int mgn_expand_offsets(const int32_t* offsets, int64_t n_segments, int32_t* out, mgn_stream_t stream) {
  MGN_CHECK_ARG(n_segments >= 0);
  if (n_segments == 0) return MGN_OK;
  MGN_CHECK_ARG(offsets && out);
  expand_offsets_kernel<<<num_sms() * 8, 256, 0, MGN_ST(as_stream(stream))>>>(offsets, n_segments, out);
  return mgn_launch_status();
}

extern "C" size_t mgn_csr_workspace_bytes(int64_t n_src, int64_t n_dst, int64_t n_edges) {
  (void)n_dst;
  (void)n_edges;
  return mgn_group_by_key_workspace_bytes(n_src);
}

extern "C" int mgn_csr_from_csc(const int32_t* offsets, const int32_t* indices, int64_t n_src,
                                int64_t n_dst, int64_t n_edges, int32_t* csc_dst, int32_t* csr_offsets,
                                int32_t* csr_eids, void* workspace, size_t workspace_bytes,
                                mgn_stream_t stream) {
  MGN_CHECK_ARG(offsets && csr_offsets && workspace);
  MGN_CHECK_ARG(n_src >= 0 && n_dst >= 0 && n_edges >= 0 && n_edges < (int64_t(1) << 31));
  MGN_CHECK_ARG(n_edges == 0 || (indices && csc_dst && csr_eids));
  int rc = mgn_expand_offsets(offsets, n_dst, csc_dst, stream);
  if (rc != MGN_OK) return rc;
  return mgn_group_by_key(indices, n_edges, n_src, csr_offsets, csr_eids, workspace, workspace_bytes, stream);
}
