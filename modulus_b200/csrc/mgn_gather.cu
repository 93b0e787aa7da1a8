// mgn_gather.cu — the HBM-bound halves of the operator seam (gnn_layers/utils.py):
// two-sided row gather (+concat / +sum), deterministic atomic-free segmented sum over a
// CSC or CSR structure, and a generic indexed row gather.  All 128-bit vectorised with a
// scalar fallback for feature widths that are not multiples of 16 bytes.
#include "mgn_common.cuh"
#include "mgn_tc.cuh"
#include "mgn_tile.cuh"

namespace mgn {
using tile::f2_add;
using tile::f2_from_bf16x2;
using tile::f2_hi;
using tile::f2_lo;
using tile::f2_packu;

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <typename T>
__device__ __forceinline__ uint4 ldg16(const T* p) {
  return __ldg(reinterpret_cast<const uint4*>(p));
}
__device__ __forceinline__ uint4 ldg16(const char* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

// fp32 sums of one 16-byte chunk, kept as packed pairs (FADD2: half the issue slots of scalar adds, same rounding)
template <typename T> struct Acc16;
template <> struct Acc16<bf16> {
  uint64_t a[4];
  __device__ __forceinline__ void zero() { a[0] = a[1] = a[2] = a[3] = 0ull; }
  __device__ __forceinline__ void add(const uint4& r) {
    a[0] = f2_add(a[0], f2_from_bf16x2(r.x));
    a[1] = f2_add(a[1], f2_from_bf16x2(r.y));
    a[2] = f2_add(a[2], f2_from_bf16x2(r.z));
    a[3] = f2_add(a[3], f2_from_bf16x2(r.w));
  }
  __device__ __forceinline__ void get(float (&f)[8]) const {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      f[2 * i] = f2_lo(a[i]);
      f[2 * i + 1] = f2_hi(a[i]);
    }
  }
};
template <> struct Acc16<float> {
  uint64_t a[2];
  __device__ __forceinline__ void zero() { a[0] = a[1] = 0ull; }
  __device__ __forceinline__ void add(const uint4& r) {
    a[0] = f2_add(a[0], f2_packu(r.x, r.y));
    a[1] = f2_add(a[1], f2_packu(r.z, r.w));
  }
  __device__ __forceinline__ void get(float (&f)[4]) const {
    f[0] = f2_lo(a[0]);
    f[1] = f2_hi(a[0]);
    f[2] = f2_lo(a[1]);
    f[3] = f2_hi(a[1]);
  }
};

// ---------------------------------------------------------------------------------------
// concat_efeat forward: out[e] = [efeat[e] | src_feat[src[e]] | dst_feat[dst[e]]]
// one warp owns kRows consecutive edges; lanes walk the 16-byte chunks of the output row
// ---------------------------------------------------------------------------------------
template <typename T, int kRows>
__global__ void __launch_bounds__(256)
concat_efeat_vec_kernel(const T* __restrict__ efeat, int ce, const T* __restrict__ sfeat, int cs,
                        const T* __restrict__ dfeat, int cd, const int32_t* __restrict__ src,
                        const int32_t* __restrict__ dst, int64_t E, T* __restrict__ out) {
  constexpr int V = Num<T>::kVec;
  const int lane = threadIdx.x & 31;
  const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  const int cpr = ce + cs + cd;
  for (int64_t e0 = warp * kRows; e0 < E; e0 += nwarps * kRows) {
    int32_t s[kRows], d[kRows];
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      const int64_t e = e0 + r;
      s[r] = e < E ? __ldg(src + e) : 0;
      d[r] = e < E ? __ldg(dst + e) : 0;
    }
    for (int c = lane; c < cpr; c += 32) {
      uint4 v[kRows];
#pragma unroll
      for (int r = 0; r < kRows; ++r) {
        const int64_t e = e0 + r;
        if (e < E) {
          const T* p;
          if (c < ce) p = efeat + (e * ce + c) * V;
          else if (c < ce + cs) p = sfeat + (static_cast<int64_t>(s[r]) * cs + (c - ce)) * V;
          else p = dfeat + (static_cast<int64_t>(d[r]) * cd + (c - ce - cs)) * V;
          v[r] = ldg16(p);
        }
      }
#pragma unroll
      for (int r = 0; r < kRows; ++r) {
        const int64_t e = e0 + r;
        if (e < E) *reinterpret_cast<uint4*>(out + (e * cpr + c) * V) = v[r];
      }
    }
  }
}

template <typename T>
__global__ void concat_efeat_scalar_kernel(const T* __restrict__ efeat, int De, const T* __restrict__ sfeat,
                                           int Ds, const T* __restrict__ dfeat, int Dd,
                                           const int32_t* __restrict__ src, const int32_t* __restrict__ dst,
                                           int64_t E, T* __restrict__ out) {
  const int D = De + Ds + Dd;
  const int64_t total = E * D;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t e = i / D;
    const int c = static_cast<int>(i - e * D);
    T v;
    if (c < De) v = efeat[e * De + c];
    else if (c < De + Ds) v = sfeat[static_cast<int64_t>(src[e]) * Ds + (c - De)];
    else v = dfeat[static_cast<int64_t>(dst[e]) * Dd + (c - De - Ds)];
    out[i] = v;
  }
}

// ---------------------------------------------------------------------------------------
// sum_efeat forward: out[e] = efeat[e] + src_feat[src[e]] + dst_feat[dst[e]]  (fp32 add)
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
sum_efeat_vec_kernel(const T* __restrict__ efeat, const T* __restrict__ sfeat, const T* __restrict__ dfeat,
                     int chunks, const int32_t* __restrict__ src, const int32_t* __restrict__ dst, int64_t E,
                     T* __restrict__ out) {
  constexpr int V = Num<T>::kVec;
  const int64_t total = E * chunks;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t e = i / chunks;
    const int c = static_cast<int>(i - e * chunks);
    Vec16<T> a, b, d;
    a.raw = ldg16(efeat + (e * chunks + c) * V);
    b.raw = ldg16(sfeat + (static_cast<int64_t>(__ldg(src + e)) * chunks + c) * V);
    d.raw = ldg16(dfeat + (static_cast<int64_t>(__ldg(dst + e)) * chunks + c) * V);
    float fa[V], fb[V], fd[V];
    a.unpack(fa); b.unpack(fb); d.unpack(fd);
#pragma unroll
    for (int k = 0; k < V; ++k) fa[k] = fa[k] + fb[k] + fd[k];
    a.pack(fa);
    *reinterpret_cast<uint4*>(out + (e * chunks + c) * V) = a.raw;
  }
}

template <typename T>
__global__ void sum_efeat_scalar_kernel(const T* __restrict__ efeat, const T* __restrict__ sfeat,
                                        const T* __restrict__ dfeat, int D, const int32_t* __restrict__ src,
                                        const int32_t* __restrict__ dst, int64_t E, T* __restrict__ out) {
  const int64_t total = E * D;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t e = i / D;
    const int c = static_cast<int>(i - e * D);
    const float v = Num<T>::to_f(efeat[i]) + Num<T>::to_f(sfeat[static_cast<int64_t>(src[e]) * D + c]) +
                    Num<T>::to_f(dfeat[static_cast<int64_t>(dst[e]) * D + c]);
    out[i] = Num<T>::from_f(v);
  }
}

// ---------------------------------------------------------------------------------------
// Long segments (hub nodes of skewed-degree graphs).  A warp that meets a segment longer than kLongSeg rows does not
// walk it: it appends ceil(len / kLongChunk) chunk entries to a worklist (one atomicAdd reserves a contiguous run, so
// the order of the chunks inside a segment is fixed); segment_long_partial_kernel sums each chunk with a whole CTA
// into an fp32 partial row, segment_long_combine_kernel adds the partials of a segment in chunk order.  Which run of
// the worklist a segment lands in depends on timing, its value does not: results stay bit-reproducible.
// ---------------------------------------------------------------------------------------
#ifndef MGN_LONG_SEG
#define MGN_LONG_SEG 64
#endif
constexpr int kLongSeg = MGN_LONG_SEG;  // (A/B switch; modulus_b200/ops.py LONG_SEGMENT mirrors the default)
constexpr int kLongChunk = 2048;
struct LongEntry {
  int32_t seg, chunk, n_chunks, pad;
};
struct LongList {
  int32_t* counter;    // [4] (counter, capacity overflow flag)
  LongEntry* entries;  // [capacity]
  int32_t capacity;
};

__device__ __forceinline__ void push_long_segment(const LongList& ll, int64_t s, int32_t len) {
  const int32_t n = (len + kLongChunk - 1) / kLongChunk;
  const int32_t base = atomicAdd(ll.counter, n);
  if (base + n > ll.capacity) {
    atomicExch(ll.counter + 1, 1);
    return;
  }
  for (int32_t k = 0; k < n; ++k) ll.entries[base + k] = LongEntry{static_cast<int32_t>(s), k, n, 0};
}

template <typename T, bool kEids>
__global__ void __launch_bounds__(256)
segment_long_partial_kernel(const T* __restrict__ in, int64_t ld_in, int64_t in_col0, int chunks,
                            const int32_t* __restrict__ offsets, const int32_t* __restrict__ eids, LongList ll,
                            float* __restrict__ partials, int64_t D) {
  constexpr int V = Num<T>::kVec;
  const int n_entries = min(*ll.counter, ll.capacity);
  __shared__ float red[8][32 * V];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const char* colp = reinterpret_cast<const char*>(in + in_col0);
  const uint32_t ld_b = static_cast<uint32_t>(ld_in * static_cast<int64_t>(sizeof(T)));
  for (int en = blockIdx.x; en < n_entries; en += gridDim.x) {
    const LongEntry le = ll.entries[en];
    const int32_t b = __ldg(offsets + le.seg) + le.chunk * kLongChunk;
    const int32_t e = min(__ldg(offsets + le.seg + 1), b + kLongChunk);
    // lane = (row slot, 16-byte column chunk): a row narrower than 32 chunks (H = 128 bf16: 16) puts 32 / G rows into one
    // warp-wide load instead of idling half the lanes; 8 loads per lane are in flight (fixed row -> lane assignment and
    // addition order: bit-reproducible)
    const int G = chunks <= 4 ? 4 : (chunks <= 8 ? 8 : (chunks <= 16 ? 16 : 32));
    const int R = 32 / G;
    const int sub = lane / G, cl = lane % G;
    for (int cb = 0; cb < chunks; cb += G) {
      const int c = cb + cl;
      const bool active = c < chunks;
      Acc16<T> acc;
      acc.zero();
      const int rows_per_pass = 8 * R;  // 8 warps x R row slots
      const int n_fly = G == 32 ? 4 : 8;  // wide rows already fill the memory pipe with 4 loads per lane (measured)
      for (int32_t j = b + warp * R + sub; j < e; j += n_fly * rows_per_pass) {
        // row ids first, then every row load back to back: a row id fetched between two row loads makes the next row
        // load wait on the scoreboard the previous one holds, i.e. one round trip per row instead of one per pass
        int32_t rr[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int32_t jj = j + rows_per_pass * u;
          rr[u] = (active && u < n_fly && jj < e) ? (kEids ? __ldg(eids + jj) : jj) : -1;
        }
        uint4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (rr[u] >= 0) v[u] = ldg16(colp + static_cast<uint64_t>(static_cast<uint32_t>(rr[u])) * ld_b + static_cast<uint32_t>(c) * 16u);
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (rr[u] >= 0) acc.add(v[u]);
      }
      float accf[V];
      acc.get(accf);
#pragma unroll
      for (int k = 0; k < V; ++k) red[warp][lane * V + k] = accf[k];
      __syncthreads();
      for (int i = threadIdx.x; i < G * V; i += blockDim.x) {
        const int col = cb * V + i;
        if (col < D) {
          float sum = 0.f;
          for (int w = 0; w < 8; ++w)
            for (int r = 0; r < R; ++r) sum += red[w][r * G * V + i];
          partials[static_cast<int64_t>(en) * D + col] = sum;
        }
      }
      __syncthreads();
    }
  }
}

// One CTA per hub: warp w adds the partial rows w, w + 8, w + 16, ... of the segment (four rows x up to four 32-column
// groups in flight per lane: a hub of a million rows has 500 partial rows, which one thread per column would walk as
// 500 dependent L2 round trips), the eight warp sums are then added in warp order.  Fixed assignment, fixed order.
template <typename T>
__global__ void __launch_bounds__(256)
segment_long_combine_kernel(LongList ll, const float* __restrict__ partials, int64_t D, const int32_t* __restrict__ offsets,
                            T* __restrict__ out, int64_t ld_out, int64_t out_col0, int mean, int accumulate) {
  const int n_entries = min(*ll.counter, ll.capacity);
  __shared__ float red[8][128];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int en = blockIdx.x; en < n_entries; en += gridDim.x) {
    const LongEntry le = ll.entries[en];
    if (le.chunk != 0) continue;
    const float scale = mean ? 1.f / static_cast<float>(max(offsets[le.seg + 1] - offsets[le.seg], 1)) : 1.f;
    for (int64_t c0 = 0; c0 < D; c0 += 128) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      const float* base = partials + static_cast<int64_t>(en) * D + c0 + lane;
      int k = warp;
      for (; k + 24 < le.n_chunks; k += 32) {
        float v[4][4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int j = 0; j < 4; ++j)
            v[u][j] = (c0 + lane + 32 * j < D) ? base[static_cast<int64_t>(k + 8 * u) * D + 32 * j] : 0.f;
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[j] += v[u][j];
      }
      for (; k < le.n_chunks; k += 8)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (c0 + lane + 32 * j < D) acc[j] += base[static_cast<int64_t>(k) * D + 32 * j];
#pragma unroll
      for (int j = 0; j < 4; ++j) red[warp][lane + 32 * j] = acc[j];
      __syncthreads();
      if (threadIdx.x < 128 && c0 + threadIdx.x < D) {
        float sum = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) sum += red[w][threadIdx.x];
        T* o = out + static_cast<int64_t>(le.seg) * ld_out + out_col0 + c0 + threadIdx.x;
        const float prev = accumulate ? Num<T>::to_f(*o) : 0.f;
        *o = Num<T>::from_f(prev + sum * scale);
      }
      __syncthreads();
    }
  }
}

// ---------------------------------------------------------------------------------------
// Segmented sum.  One warp per segment; the warp is split into 32/G row groups of G lanes
// (G = lanes needed for one row's 16-byte chunks), each group sums every (32/G)-th row in
// fp32, then the groups are combined with a fixed xor-shuffle tree: no atomics, run-to-run
// bit-identical.  Rows of a CSC segment are contiguous (eids == nullptr); CSR segments go
// through the edge-id indirection.
// ---------------------------------------------------------------------------------------
template <typename T, int G>
__global__ void __launch_bounds__(256)
segment_sum_vec_kernel(const T* __restrict__ in, int64_t ld_in, int64_t in_col0, int chunks,
                       const int32_t* __restrict__ offsets, const int32_t* __restrict__ eids, int64_t n_seg,
                       T* __restrict__ out, int64_t ld_out, int64_t out_col0, int mean, int accumulate, LongList ll) {
  constexpr int V = Num<T>::kVec;
  constexpr int R = 32 / G;
  const int lane = threadIdx.x & 31;
  const int g = lane / G, l = lane % G;
  const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t s = warp; s < n_seg; s += nwarps) {
    const int32_t b = __ldg(offsets + s), e = __ldg(offsets + s + 1);
    if (ll.counter != nullptr && e - b > kLongSeg) {  // hub: handed to the long-segment kernels
      if (lane == 0) push_long_segment(ll, s, e - b);
      continue;
    }
    const float scale = mean ? 1.f / static_cast<float>(max(e - b, 1)) : 1.f;
    for (int cb = 0; cb < chunks; cb += G) {
      const int c = cb + l;
      const bool active = c < chunks;
      float acc[V];
#pragma unroll
      for (int k = 0; k < V; ++k) acc[k] = 0.f;
      // up to 4 rows in flight per lane group, also for short segments (mesh in-degrees are ~6: a predicated batch
      // keeps every row of the segment in flight at once instead of one dependent load per iteration)
      for (int32_t j = b + g; j < e; j += 4 * R) {
        uint4 v[4];
        bool on[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          on[u] = active && (j + u * R < e);
          if (on[u]) {
            const int64_t row = eids ? __ldg(eids + j + u * R) : (j + u * R);
            v[u] = ldg16(in + row * ld_in + in_col0 + static_cast<int64_t>(c) * V);
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (on[u]) {
            Vec16<T> t;
            t.raw = v[u];
            float f[V];
            t.unpack(f);
#pragma unroll
            for (int k = 0; k < V; ++k) acc[k] += f[k];
          }
        }
      }
#pragma unroll
      for (int off = G; off < 32; off <<= 1) {
#pragma unroll
        for (int k = 0; k < V; ++k) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], off);
      }
      if (g == 0 && active) {
        T* o = out + s * ld_out + out_col0 + static_cast<int64_t>(c) * V;
        Vec16<T> t;
        if (accumulate) {
          t.raw = *reinterpret_cast<const uint4*>(o);
          float f[V];
          t.unpack(f);
#pragma unroll
          for (int k = 0; k < V; ++k) acc[k] = f[k] + acc[k] * scale;
        } else {
#pragma unroll
          for (int k = 0; k < V; ++k) acc[k] *= scale;
        }
        t.pack(acc);
        *reinterpret_cast<uint4*>(o) = t.raw;
      }
    }
  }
}

// Same sum, same order, more bytes in flight: when one lane group covers a whole row (chunks == G) a warp takes S
// consecutive segments at a time and issues the first U rows per lane group of ALL of them before consuming any
// (mesh in-degrees are ~6: with G = 16 that is the whole segment), i.e. S*U 16-byte loads in flight per lane
// instead of one dependent load per short segment.  Longer segments finish in the tail loop.
// The three dependent loads of a group (segment bounds -> row ids -> rows) are software-pipelined across the groups a warp
// visits: while the rows of group k are in flight the row ids of group k + 1 and the bounds of group k + 2 are fetched, so
// the only exposed round trips are the two of the prologue (measured at c3: CSR sum 0.50 -> see profiles/r02_segment_sum.md).
// The order of additions inside a segment is unchanged.
template <typename T, int G, int S, int U>
__global__ void __launch_bounds__(256, 2)  // two blocks per SM: the bytes in flight are what this kernel lives on
segment_sum_batch_kernel(const T* __restrict__ in, int64_t ld_in, int64_t in_col0, const int32_t* __restrict__ offsets,
                         const int32_t* __restrict__ eids, int64_t n_seg, T* __restrict__ out, int64_t ld_out,
                         int64_t out_col0, int mean, int accumulate, LongList ll) {
  constexpr int V = Num<T>::kVec;
  constexpr int R = 32 / G;
  const int lane = threadIdx.x & 31;
  const int g = lane / G, c = lane % G;
  const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t stride = ((static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5) * S;
  const T* col = in + in_col0 + static_cast<int64_t>(c) * V;

  auto load_bounds = [&](int64_t s0, int32_t(&b)[S], int32_t(&e)[S]) {
#pragma unroll
    for (int i = 0; i < S; ++i) {
      const bool valid = s0 + i < n_seg;
      b[i] = valid ? __ldg(offsets + s0 + i) : 0;
      e[i] = valid ? __ldg(offsets + s0 + i + 1) : 0;
    }
  };
  // row ids of the first U rows per lane group of every segment of a group (-1: none); hubs are handed to the
  // long-segment kernels and contribute nothing here
  auto load_row_ids = [&](const int32_t(&b)[S], const int32_t(&e)[S], int32_t(&rows)[S][U]) {
#pragma unroll
    for (int i = 0; i < S; ++i) {
      const bool skip = ll.counter != nullptr && e[i] - b[i] > kLongSeg;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int32_t j = b[i] + g + u * R;
        rows[i][u] = (j < e[i] && !skip) ? (eids ? __ldg(eids + j) : j) : -1;
      }
    }
  };

  int64_t s0 = warp * S;
  if (s0 >= n_seg) return;
  int32_t b[S], e[S], rows[S][U], bn[S], en[S];
  load_bounds(s0, b, e);
  load_bounds(s0 + stride, bn, en);
  load_row_ids(b, e, rows);
  for (; s0 < n_seg; s0 += stride) {
    uint4 v[S][U];
#pragma unroll
    for (int i = 0; i < S; ++i)
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (rows[i][u] >= 0) v[i][u] = ldg16(col + static_cast<int64_t>(rows[i][u]) * ld_in);
    // next group's row ids (its bounds were requested one iteration ago) and the bounds of the group after it
    int32_t rows_n[S][U], bnn[S], enn[S];
    load_row_ids(bn, en, rows_n);
    load_bounds(s0 + 2 * stride, bnn, enn);
#pragma unroll
    for (int i = 0; i < S; ++i) {
      const bool is_long = ll.counter != nullptr && e[i] - b[i] > kLongSeg;
      if (is_long && lane == 0) push_long_segment(ll, s0 + i, e[i] - b[i]);
      float acc[V];
#pragma unroll
      for (int k = 0; k < V; ++k) acc[k] = 0.f;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (rows[i][u] >= 0) {
          Vec16<T> t;
          t.raw = v[i][u];
          float f[V];
          t.unpack(f);
#pragma unroll
          for (int k = 0; k < V; ++k) acc[k] += f[k];
        }
      }
      if (!is_long) {
        // longer segments (uniform per warp): four rows per lane group in flight, added in the same order as before
        int32_t j = b[i] + g + U * R;
        for (; j + 3 * R < e[i]; j += 4 * R) {
          int64_t rr[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) rr[u] = eids ? static_cast<int64_t>(__ldg(eids + j + u * R)) : static_cast<int64_t>(j + u * R);
          uint4 vv[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) vv[u] = ldg16(col + rr[u] * ld_in);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            Vec16<T> t;
            t.raw = vv[u];
            float f[V];
            t.unpack(f);
#pragma unroll
            for (int k = 0; k < V; ++k) acc[k] += f[k];
          }
        }
        for (; j < e[i]; j += R) {
          const int64_t row = eids ? __ldg(eids + j) : j;
          Vec16<T> t;
          t.raw = ldg16(col + row * ld_in);
          float f[V];
          t.unpack(f);
#pragma unroll
          for (int k = 0; k < V; ++k) acc[k] += f[k];
        }
      }
#pragma unroll
      for (int off = G; off < 32; off <<= 1) {
#pragma unroll
        for (int k = 0; k < V; ++k) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], off);
      }
      if (g == 0 && s0 + i < n_seg && !is_long) {
        const float scale = mean ? 1.f / static_cast<float>(max(e[i] - b[i], 1)) : 1.f;
        T* o = out + (s0 + i) * ld_out + out_col0 + static_cast<int64_t>(c) * V;
        Vec16<T> t;
        if (accumulate) {
          t.raw = *reinterpret_cast<const uint4*>(o);
          float f[V];
          t.unpack(f);
#pragma unroll
          for (int k = 0; k < V; ++k) acc[k] = f[k] + acc[k] * scale;
        } else {
#pragma unroll
          for (int k = 0; k < V; ++k) acc[k] *= scale;
        }
        t.pack(acc);
        *reinterpret_cast<uint4*>(o) = t.raw;
      }
    }
#pragma unroll
    for (int i = 0; i < S; ++i) {
      b[i] = bn[i];
      e[i] = en[i];
      bn[i] = bnn[i];
      en[i] = enn[i];
#pragma unroll
      for (int u = 0; u < U; ++u) rows[i][u] = rows_n[i][u];
    }
  }
}

// ---------------------------------------------------------------------------------------
// Segmented sum, third form (the default): G lanes own one SEGMENT (not one row group of a segment), so a warp
// works on 32 / G segments side by side and nothing is exchanged between lanes -- the xor-shuffle tree and the
// per-row-group bookkeeping of the forms above were a third of their instructions, and at mesh degrees (~6 rows of
// 256 bytes) those kernels were bound by instruction issue and fetch, not by HBM (ncu: 50 % issue slots busy at 19 %
// of the warps, `no_instruction` stalls; 1 M EMPTY segments cost 271 us).  Per lane S x U 16-byte loads are in flight
// (the first U rows of S segments); bounds of the group after next and row ids of the next group are fetched while
// the rows of this one are in flight.  Rows are added in ascending order with packed fp32 adds: bit-reproducible,
// and the same order as the sums fused into the edge kernels (mgn_agg.cuh).  Column blocks beyond 32 chunks go to
// blockIdx.y.  kEids is a template parameter: a possible row-id load between two row loads makes ptxas put both on
// one scoreboard and the row loads then complete one at a time.
// ---------------------------------------------------------------------------------------
template <typename T, int G, int S, int U, bool kEids>
__global__ void __launch_bounds__(256, 2)
segment_sum_sub_kernel(const T* __restrict__ in, int64_t ld_in, int64_t in_col0, int chunks, const int32_t* __restrict__ offsets,
                       const int32_t* __restrict__ eids, int64_t n_seg, T* __restrict__ out, int64_t ld_out,
                       int64_t out_col0, int mean, int accumulate, LongList ll) {
  constexpr int V = Num<T>::kVec;
  constexpr int R = 32 / G;  // segments side by side in a warp
#ifdef MGN_SEG_TAIL
  constexpr int kTail = MGN_SEG_TAIL;
#else
  constexpr int kTail = 8;
#endif
  const int lane = threadIdx.x & 31;
  const int sub = lane / G;
  const int c = static_cast<int>(blockIdx.y) * G + lane % G;
  const bool active = c < chunks;
  const bool leader = lane % G == 0 && blockIdx.y == 0;
  const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t stride = ((static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5) * (S * R);
  const char* colp = reinterpret_cast<const char*>(in + in_col0 + static_cast<int64_t>(c) * V);
  const uint32_t ld_b = static_cast<uint32_t>(ld_in * static_cast<int64_t>(sizeof(T)));
  auto row_ptr = [&](int32_t row) { return colp + static_cast<uint64_t>(static_cast<uint32_t>(row)) * ld_b; };

  // segment i of a group: s0 + i * R + sub (neighbouring lane groups write neighbouring output rows)
  auto load_bounds = [&](int64_t s0, int32_t(&b)[S], int32_t(&e)[S]) {
#pragma unroll
    for (int i = 0; i < S; ++i) {
      const int64_t s = s0 + i * R + sub;
      const bool valid = s < n_seg && active;
      b[i] = valid ? __ldg(offsets + s) : 0;
      e[i] = valid ? __ldg(offsets + s + 1) : 0;
    }
  };
  auto load_row_ids = [&](const int32_t(&b)[S], const int32_t(&e)[S], int32_t(&rows)[S][U]) {
#pragma unroll
    for (int i = 0; i < S; ++i) {
      const bool skip = ll.counter != nullptr && e[i] - b[i] > kLongSeg;  // hubs: long-segment kernels
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int32_t j = b[i] + u;
        rows[i][u] = (j < e[i] && !skip) ? (kEids ? __ldg(eids + j) : j) : -1;
      }
    }
  };

  int64_t s0 = warp * (S * R);
  if (s0 >= n_seg) return;
  int32_t b[S], e[S], rows[S][U], bn[S], en[S];
  load_bounds(s0, b, e);
  load_bounds(s0 + stride, bn, en);
  load_row_ids(b, e, rows);
  for (; s0 < n_seg; s0 += stride) {
    uint4 v[S][U];
#pragma unroll
    for (int i = 0; i < S; ++i)
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (rows[i][u] >= 0) v[i][u] = ldg16(row_ptr(rows[i][u]));
    int32_t rows_n[S][U], bnn[S], enn[S];
    load_row_ids(bn, en, rows_n);
    load_bounds(s0 + 2 * stride, bnn, enn);
#pragma unroll
    for (int i = 0; i < S; ++i) {
      const int64_t s = s0 + i * R + sub;
      const int32_t len = e[i] - b[i];
      const bool is_long = ll.counter != nullptr && len > kLongSeg;
      if (is_long && leader) push_long_segment(ll, s, len);
      Acc16<T> acc;
      acc.zero();
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (rows[i][u] >= 0) acc.add(v[i][u]);
      if (len > U && !is_long) {  // longer segments: kTail rows in flight
        int32_t j = b[i] + U;
        for (; j + kTail - 1 < e[i]; j += kTail) {
          int32_t rr[kTail];
#pragma unroll
          for (int u = 0; u < kTail; ++u) rr[u] = kEids ? __ldg(eids + j + u) : j + u;
          uint4 vv[kTail];
#pragma unroll
          for (int u = 0; u < kTail; ++u) vv[u] = ldg16(row_ptr(rr[u]));
#pragma unroll
          for (int u = 0; u < kTail; ++u) acc.add(vv[u]);
        }
        for (; j < e[i]; ++j) acc.add(ldg16(row_ptr(kEids ? __ldg(eids + j) : j)));
      }
      if (active && s < n_seg && !is_long) {
        float f[V];
        acc.get(f);
        const float scale = mean ? 1.f / static_cast<float>(max(len, 1)) : 1.f;
        T* o = out + s * ld_out + out_col0 + static_cast<int64_t>(c) * V;
        Vec16<T> t;
        if (accumulate) {
          t.raw = *reinterpret_cast<const uint4*>(o);
          float p[V];
          t.unpack(p);
#pragma unroll
          for (int k = 0; k < V; ++k) f[k] = p[k] + f[k] * scale;
        } else if (mean) {
#pragma unroll
          for (int k = 0; k < V; ++k) f[k] *= scale;
        }
        t.pack(f);
        *reinterpret_cast<uint4*>(o) = t.raw;
      }
    }
#pragma unroll
    for (int i = 0; i < S; ++i) {
      b[i] = bn[i];
      e[i] = en[i];
      bn[i] = bnn[i];
      en[i] = enn[i];
#pragma unroll
      for (int u = 0; u < U; ++u) rows[i][u] = rows_n[i][u];
    }
  }
}

template <typename T>
__global__ void segment_sum_scalar_kernel(const T* __restrict__ in, int64_t ld_in, int64_t in_col0, int D,
                                          const int32_t* __restrict__ offsets, const int32_t* __restrict__ eids,
                                          int64_t n_seg, T* __restrict__ out, int64_t ld_out, int64_t out_col0,
                                          int mean, int accumulate) {
  const int64_t total = n_seg * D;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t s = i / D;
    const int c = static_cast<int>(i - s * D);
    const int32_t b = offsets[s], e = offsets[s + 1];
    float acc = 0.f;
    for (int32_t j = b; j < e; ++j) {
      const int64_t row = eids ? eids[j] : j;
      acc += Num<T>::to_f(in[row * ld_in + in_col0 + c]);
    }
    if (mean) acc /= static_cast<float>(max(e - b, 1));
    T* o = out + s * ld_out + out_col0 + c;
    if (accumulate) acc += Num<T>::to_f(*o);
    *o = Num<T>::from_f(acc);
  }
}

// ---------------------------------------------------------------------------------------
// generic indexed row gather into a column slice
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
gather_rows_vec_kernel(const T* __restrict__ in, int64_t ld_in, int64_t in_col0, int chunks,
                       const int32_t* __restrict__ idx, int64_t n_rows, T* __restrict__ out, int64_t ld_out,
                       int64_t out_col0, const int32_t* __restrict__ deg_offsets) {
  constexpr int V = Num<T>::kVec;
  const int64_t total = n_rows * chunks;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t r = i / chunks;
    const int c = static_cast<int>(i - r * chunks);
    const int64_t row = idx ? __ldg(idx + r) : r;
    Vec16<T> t;
    t.raw = ldg16(in + row * ld_in + in_col0 + static_cast<int64_t>(c) * V);
    if (deg_offsets) {
      const float sc = 1.f / static_cast<float>(max(__ldg(deg_offsets + row + 1) - __ldg(deg_offsets + row), 1));
      float f[V];
      t.unpack(f);
#pragma unroll
      for (int k = 0; k < V; ++k) f[k] *= sc;
      t.pack(f);
    }
    *reinterpret_cast<uint4*>(out + r * ld_out + out_col0 + static_cast<int64_t>(c) * V) = t.raw;
  }
}

template <typename T>
__global__ void gather_rows_scalar_kernel(const T* __restrict__ in, int64_t ld_in, int64_t in_col0, int D,
                                          const int32_t* __restrict__ idx, int64_t n_rows, T* __restrict__ out,
                                          int64_t ld_out, int64_t out_col0, const int32_t* __restrict__ deg_offsets) {
  const int64_t total = n_rows * D;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t r = i / D;
    const int c = static_cast<int>(i - r * D);
    const int64_t row = idx ? idx[r] : r;
    float v = Num<T>::to_f(in[row * ld_in + in_col0 + c]);
    if (deg_offsets) v /= static_cast<float>(max(deg_offsets[row + 1] - deg_offsets[row], 1));
    out[r * ld_out + out_col0 + c] = Num<T>::from_f(v);
  }
}

static inline int grid_for(int64_t work_items, int threads_per_item_block = 256, int max_waves = 16) {
  int64_t blocks = (work_items + threads_per_item_block - 1) / threads_per_item_block;
  const int64_t cap = static_cast<int64_t>(num_sms()) * max_waves;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

template <typename T>
static int concat_efeat_fwd_t(const void* efeat, int64_t De, const void* sfeat, int64_t Ds, const void* dfeat,
                              int64_t Dd, const int32_t* src, const int32_t* dst, int64_t E, void* out,
                              cudaStream_t st) {
  constexpr int V = Num<T>::kVec;
  if (E == 0) return MGN_OK;
  const bool vec = De % V == 0 && Ds % V == 0 && Dd % V == 0 && aligned16(efeat) && aligned16(sfeat) &&
                   aligned16(dfeat) && aligned16(out);
  if (vec) {
    constexpr int kRows = 4;
    const int64_t warps = (E + kRows - 1) / kRows;
    concat_efeat_vec_kernel<T, kRows><<<grid_for(warps * 32), 256, 0, MGN_ST(st)>>>(
        static_cast<const T*>(efeat), static_cast<int>(De / V), static_cast<const T*>(sfeat),
        static_cast<int>(Ds / V), static_cast<const T*>(dfeat), static_cast<int>(Dd / V), src, dst, E,
        static_cast<T*>(out));
  } else {
    concat_efeat_scalar_kernel<T><<<grid_for(E * (De + Ds + Dd)), 256, 0, MGN_ST(st)>>>(
        static_cast<const T*>(efeat), static_cast<int>(De), static_cast<const T*>(sfeat), static_cast<int>(Ds),
        static_cast<const T*>(dfeat), static_cast<int>(Dd), src, dst, E, static_cast<T*>(out));
  }
  return mgn_launch_status();
}

template <typename T>
static int sum_efeat_fwd_t(const void* efeat, const void* sfeat, const void* dfeat, int64_t D, const int32_t* src,
                           const int32_t* dst, int64_t E, void* out, cudaStream_t st) {
  constexpr int V = Num<T>::kVec;
  if (E == 0) return MGN_OK;
  const bool vec = D % V == 0 && aligned16(efeat) && aligned16(sfeat) && aligned16(dfeat) && aligned16(out);
  if (vec)
    sum_efeat_vec_kernel<T><<<grid_for(E * (D / V)), 256, 0, MGN_ST(st)>>>(
        static_cast<const T*>(efeat), static_cast<const T*>(sfeat), static_cast<const T*>(dfeat),
        static_cast<int>(D / V), src, dst, E, static_cast<T*>(out));
  else
    sum_efeat_scalar_kernel<T><<<grid_for(E * D), 256, 0, MGN_ST(st)>>>(
        static_cast<const T*>(efeat), static_cast<const T*>(sfeat), static_cast<const T*>(dfeat),
        static_cast<int>(D), src, dst, E, static_cast<T*>(out));
  return mgn_launch_status();
}

static size_t long_capacity(int64_t n_rows) { return static_cast<size_t>(n_rows / kLongChunk + n_rows / kLongSeg + 16); }

template <typename T>
static int segment_sum_t(const void* in, int64_t ld_in, int64_t in_col0, int64_t D, const int32_t* offsets,
                         const int32_t* eids, int64_t n_seg, void* out, int64_t ld_out, int64_t out_col0, int mean,
                         int accumulate, cudaStream_t st, int64_t n_rows = 0, void* workspace = nullptr) {
  constexpr int V = Num<T>::kVec;
  if (n_seg == 0 || D == 0) return MGN_OK;
  const bool vec = D % V == 0 && ld_in % V == 0 && in_col0 % V == 0 && ld_out % V == 0 && out_col0 % V == 0 &&
                   aligned16(in) && aligned16(out);
  const T* i_ = static_cast<const T*>(in);
  T* o_ = static_cast<T*>(out);
  if (vec) {
    const int chunks = static_cast<int>(D / V);
    const int grid = grid_for(n_seg * 32);
    LongList ll{nullptr, nullptr, 0};
    float* partials = nullptr;
    if (workspace != nullptr) {  // balanced variant: hubs go through the worklist
      const size_t cap = long_capacity(n_rows);
      ll.counter = static_cast<int32_t*>(workspace);
      ll.entries = reinterpret_cast<LongEntry*>(static_cast<char*>(workspace) + 16);
      ll.capacity = static_cast<int32_t>(cap);
      partials = reinterpret_cast<float*>(static_cast<char*>(workspace) + 16 + cap * sizeof(LongEntry));
      cudaError_t ce = cudaMemsetAsync(ll.counter, 0, 16, st);
      if (ce != cudaSuccess) return static_cast<int>(ce);
    }
#ifdef MGN_SEG_OLD  // A/B: the row-group forms
#define MGN_SEG(G)                                                                                     \
  segment_sum_vec_kernel<T, G><<<grid, 256, 0, MGN_ST(st)>>>(i_, ld_in, in_col0, chunks, offsets, eids, n_seg, \
                                                     o_, ld_out, out_col0, mean, accumulate, ll)
    if (chunks == 16)
      segment_sum_batch_kernel<T, 16, 4, 3><<<grid_for((n_seg + 3) / 4 * 32), 256, 0, MGN_ST(st)>>>(
          i_, ld_in, in_col0, offsets, eids, n_seg, o_, ld_out, out_col0, mean, accumulate, ll);
    else if (chunks <= 4) MGN_SEG(4);
    else if (chunks <= 8) MGN_SEG(8);
    else if (chunks <= 16) MGN_SEG(16);
    else MGN_SEG(32);
#undef MGN_SEG
#else
    if (ld_in * static_cast<int64_t>(sizeof(T)) >= (int64_t{1} << 32)) return MGN_EINVAL;
    constexpr int kS = 2, kU = 6;
#define MGN_SUB(G)                                                                                              \
  do {                                                                                                          \
    const int64_t warps = (n_seg + kS * (32 / G) - 1) / (kS * (32 / G));                                        \
    const dim3 grid_(grid_for(warps * 32), (chunks + G - 1) / G);                                               \
    if (eids)                                                                                                   \
      segment_sum_sub_kernel<T, G, kS, kU, true><<<grid_, 256, 0, MGN_ST(st)>>>(                                \
          i_, ld_in, in_col0, chunks, offsets, eids, n_seg, o_, ld_out, out_col0, mean, accumulate, ll);        \
    else                                                                                                        \
      segment_sum_sub_kernel<T, G, kS, kU, false><<<grid_, 256, 0, MGN_ST(st)>>>(                               \
          i_, ld_in, in_col0, chunks, offsets, eids, n_seg, o_, ld_out, out_col0, mean, accumulate, ll);        \
  } while (0)
    if (chunks <= 4) MGN_SUB(4);
    else if (chunks <= 8) MGN_SUB(8);
    else if (chunks <= 16) MGN_SUB(16);
    else MGN_SUB(32);
#undef MGN_SUB
    (void)grid;
#endif
    int rc = mgn_launch_status();
    if (rc != MGN_OK || ll.counter == nullptr) return rc;
    const int lgrid = ll.capacity < 4 * num_sms() ? ll.capacity : 4 * num_sms();
    if (eids) segment_long_partial_kernel<T, true><<<lgrid, 256, 0, MGN_ST(st)>>>(i_, ld_in, in_col0, chunks, offsets, eids, ll, partials, D);
    else segment_long_partial_kernel<T, false><<<lgrid, 256, 0, MGN_ST(st)>>>(i_, ld_in, in_col0, chunks, offsets, eids, ll, partials, D);
    segment_long_combine_kernel<T><<<lgrid, 256, 0, MGN_ST(st)>>>(ll, partials, D, offsets, o_, ld_out, out_col0, mean,
                                                             accumulate);
  } else {
    segment_sum_scalar_kernel<T><<<grid_for(n_seg * D), 256, 0, MGN_ST(st)>>>(i_, ld_in, in_col0, static_cast<int>(D),
                                                                      offsets, eids, n_seg, o_, ld_out, out_col0,
                                                                      mean, accumulate);
  }
  return mgn_launch_status();
}

template <typename T>
static int gather_rows_t(const void* in, int64_t ld_in, int64_t in_col0, int64_t D, const int32_t* idx,
                         int64_t n_rows, void* out, int64_t ld_out, int64_t out_col0, const int32_t* deg_offsets,
                         cudaStream_t st) {
  constexpr int V = Num<T>::kVec;
  if (n_rows == 0 || D == 0) return MGN_OK;
  const bool vec = D % V == 0 && ld_in % V == 0 && in_col0 % V == 0 && ld_out % V == 0 && out_col0 % V == 0 &&
                   aligned16(in) && aligned16(out);
  if (vec)
    gather_rows_vec_kernel<T><<<grid_for(n_rows * (D / V)), 256, 0, MGN_ST(st)>>>(
        static_cast<const T*>(in), ld_in, in_col0, static_cast<int>(D / V), idx, n_rows, static_cast<T*>(out),
        ld_out, out_col0, deg_offsets);
  else
    gather_rows_scalar_kernel<T><<<grid_for(n_rows * D), 256, 0, MGN_ST(st)>>>(
        static_cast<const T*>(in), ld_in, in_col0, static_cast<int>(D), idx, n_rows, static_cast<T*>(out), ld_out,
        out_col0, deg_offsets);
  return mgn_launch_status();
}

}  // namespace mgn

using namespace mgn;

extern "C" int mgn_concat_efeat_fwd(int dtype, const void* efeat, int64_t De, const void* src_feat, int64_t Ds,
                                    const void* dst_feat, int64_t Dd, const int32_t* src_idx,
                                    const int32_t* dst_idx, int64_t n_edges, void* out, mgn_stream_t stream) {
  MGN_CHECK_ARG(n_edges >= 0 && De >= 0 && Ds >= 0 && Dd >= 0);
  if (n_edges == 0) return MGN_OK;
  MGN_CHECK_ARG(efeat && src_feat && dst_feat && src_idx && dst_idx && out);
  if (dtype == MGN_F32)
    return concat_efeat_fwd_t<float>(efeat, De, src_feat, Ds, dst_feat, Dd, src_idx, dst_idx, n_edges, out,
                                     as_stream(stream));
  if (dtype == MGN_BF16)
    return concat_efeat_fwd_t<bf16>(efeat, De, src_feat, Ds, dst_feat, Dd, src_idx, dst_idx, n_edges, out,
                                    as_stream(stream));
  return MGN_EINVAL;
}

extern "C" int mgn_sum_efeat_fwd(int dtype, const void* efeat, const void* src_feat, const void* dst_feat,
                                 int64_t D, const int32_t* src_idx, const int32_t* dst_idx, int64_t n_edges,
                                 void* out, mgn_stream_t stream) {
  MGN_CHECK_ARG(n_edges >= 0 && D >= 0);
  if (n_edges == 0) return MGN_OK;
  MGN_CHECK_ARG(efeat && src_feat && dst_feat && src_idx && dst_idx && out);
  if (dtype == MGN_F32)
    return sum_efeat_fwd_t<float>(efeat, src_feat, dst_feat, D, src_idx, dst_idx, n_edges, out, as_stream(stream));
  if (dtype == MGN_BF16)
    return sum_efeat_fwd_t<bf16>(efeat, src_feat, dst_feat, D, src_idx, dst_idx, n_edges, out, as_stream(stream));
  return MGN_EINVAL;
}

extern "C" int mgn_segment_sum(int dtype, const void* in, int64_t ld_in, int64_t in_col0, int64_t D,
                               const int32_t* offsets, const int32_t* eids, int64_t n_segments, void* out,
                               int64_t ld_out, int64_t out_col0, int mean, int accumulate, mgn_stream_t stream) {
  MGN_CHECK_ARG(n_segments >= 0 && D >= 0 && ld_in >= 0 && ld_out >= 0 && in_col0 >= 0 && out_col0 >= 0);
  if (n_segments == 0) return MGN_OK;
  MGN_CHECK_ARG(offsets && out);
  if (dtype == MGN_F32)
    return segment_sum_t<float>(in, ld_in, in_col0, D, offsets, eids, n_segments, out, ld_out, out_col0, mean,
                                accumulate, as_stream(stream));
  if (dtype == MGN_BF16)
    return segment_sum_t<bf16>(in, ld_in, in_col0, D, offsets, eids, n_segments, out, ld_out, out_col0, mean,
                               accumulate, as_stream(stream));
  return MGN_EINVAL;
}

/* mgn_segment_sum with hub handling: segments longer than 256 rows are split into 2048-row chunks summed by whole
 * CTAs and combined in chunk order (skewed-degree graphs; same results run to run).  n_rows = rows of `in` that the
 * offsets cover (sizes the worklist), workspace >= mgn_segment_sum_workspace_bytes(n_rows, D). */
extern "C" size_t mgn_segment_sum_workspace_bytes(int64_t n_rows, int64_t D) {
  if (n_rows <= 0) return 16;
  return 16 + long_capacity(n_rows) * (sizeof(LongEntry) + static_cast<size_t>(D) * sizeof(float));
}

extern "C" int mgn_segment_sum_balanced(int dtype, const void* in, int64_t ld_in, int64_t in_col0, int64_t D,
                                        const int32_t* offsets, const int32_t* eids, int64_t n_segments, void* out,
                                        int64_t ld_out, int64_t out_col0, int mean, int accumulate, int64_t n_rows,
                                        void* workspace, size_t workspace_bytes, mgn_stream_t stream) {
  MGN_CHECK_ARG(n_segments >= 0 && D >= 0 && ld_in >= 0 && ld_out >= 0 && in_col0 >= 0 && out_col0 >= 0 && n_rows >= 0);
  if (n_segments == 0) return MGN_OK;
  MGN_CHECK_ARG(offsets && out && workspace);
  if (workspace_bytes < mgn_segment_sum_workspace_bytes(n_rows, D)) return MGN_EWORKSPACE;
  if (dtype == MGN_F32)
    return segment_sum_t<float>(in, ld_in, in_col0, D, offsets, eids, n_segments, out, ld_out, out_col0, mean,
                                accumulate, as_stream(stream), n_rows, workspace);
  if (dtype == MGN_BF16)
    return segment_sum_t<bf16>(in, ld_in, in_col0, D, offsets, eids, n_segments, out, ld_out, out_col0, mean,
                               accumulate, as_stream(stream), n_rows, workspace);
  return MGN_EINVAL;
}

extern "C" int mgn_gather_rows(int dtype, const void* in, int64_t ld_in, int64_t in_col0, int64_t D,
                               const int32_t* idx, int64_t n_rows, void* out, int64_t ld_out, int64_t out_col0,
                               const int32_t* inv_deg_offsets, mgn_stream_t stream) {
  MGN_CHECK_ARG(n_rows >= 0 && D >= 0 && ld_in >= 0 && ld_out >= 0 && in_col0 >= 0 && out_col0 >= 0);
  if (n_rows == 0) return MGN_OK;
  MGN_CHECK_ARG(in && out);
  if (dtype == MGN_F32)
    return gather_rows_t<float>(in, ld_in, in_col0, D, idx, n_rows, out, ld_out, out_col0, inv_deg_offsets,
                                as_stream(stream));
  if (dtype == MGN_BF16)
    return gather_rows_t<bf16>(in, ld_in, in_col0, D, idx, n_rows, out, ld_out, out_col0, inv_deg_offsets,
                               as_stream(stream));
  return MGN_EINVAL;
}
