// mgn_agg.cuh — destination sums fused into the edge kernels (aggregate_and_concat "sum" of the reference,
// models/gnn_layers/utils.py:337-378, and its transpose in backward).
//
// Edge rows are in CSC order, so the rows of one destination are contiguous.  While a 128-row result tile is still in
// shared memory four warps sum each destination segment of the tile (fp32, rows in ascending order).  Segments that
// lie wholly inside the tile -- empty ones included -- are written straight to the [n_dst, 128] table; the tile's
// first and last segment may continue in a neighbouring tile, so they go to fp32 records (two per tile) that
// agg_fixup_kernel combines in tile order: no atomics, bit-reproducible.
#pragma once
#include "mgn_common.cuh"
#include "mgn_tc.cuh"
#include "mgn_tile.cuh"

namespace mgn {
namespace agg {

using namespace tile;
constexpr int kH = 128;
#ifdef MGN_AGG_FLY
constexpr int kAggFly = MGN_AGG_FLY;  // (A/B switch)
#else
constexpr int kAggFly = 4;
#endif
#ifdef MGN_AGG_SIDE
constexpr int kAggSide = MGN_AGG_SIDE;  // (A/B switch)
#else
constexpr int kAggSide = 2;  // rows in flight per segment while kSegAhead segments are summed side by side
#endif

static inline size_t workspace_bytes(int64_t M) {
  if (M <= 0) return 0;
  const size_t n_tiles = static_cast<size_t>((M + kRows - 1) / kRows);
  return 2 * n_tiles * (kH * sizeof(float) + sizeof(int32_t));
}

// 128 threads (mt = 0..127): thread = (16-byte column chunk, segment lane); buf = result tile (two swizzled panels)
// row0 / M / seg_id are relative to the rows this launch covers; row_base = position of its row 0 in the CSC order
// (seg_off holds global positions), rec_base = index of its first tile's records (several launches over consecutive
// row ranges share one record array in row order and one fix-up).
//
// The index loads (first / last destination of the tile, then the bounds of this thread's segments) are a chain of
// two L2 round trips: `tile_segments_begin` issues them EARLY -- the callers run it before they wait for the tile.
#ifdef MGN_SEG_AHEAD
constexpr int kSegAhead = MGN_SEG_AHEAD;  // (A/B switch)
#else
constexpr int kSegAhead = 3;
#endif
// rounds of segment bounds fetched ahead per thread (a mesh tile has ~22 segments = 3 rounds of 8)
struct TileSegs {
  int v_first, v_last, nrows;
  int32_t ob[kSegAhead], oe[kSegAhead];  // bounds of this thread's first segments: v_first + segment lane + 8 j
};

// stage A: first / last destination of the tile; stage B (needs A's values): bounds of this thread's segments.  A
// caller that runs A one tile before B never waits for the first round trip (the warp issues in order: B's address
// arithmetic would otherwise stall on A's loads), and fetching the bounds of ALL rounds up front takes the third
// round trip -- once per round in the summing loop -- off the loop.
__device__ __forceinline__ TileSegs tile_segments_ids(long long row0, long long M, const int32_t* __restrict__ seg_id) {
  TileSegs ts;
  const long long rem = M - row0;
  ts.nrows = rem < kRows ? static_cast<int>(rem) : kRows;
  ts.v_first = __ldg(seg_id + row0);
  ts.v_last = __ldg(seg_id + row0 + ts.nrows - 1);
#pragma unroll
  for (int j = 0; j < kSegAhead; ++j) ts.ob[j] = ts.oe[j] = 0;
  return ts;
}
// (kLanes = segment lanes = threads / 16: 8 for the four mover warps, 12 when two more warps help)
template <int kLanes = 8>
__device__ __forceinline__ void tile_segments_bounds(TileSegs& ts, const int32_t* __restrict__ seg_off, int mt) {
#pragma unroll
  for (int j = 0; j < kSegAhead; ++j) {
    const int v = ts.v_first + (mt >> 4) + kLanes * j;
    if (v <= ts.v_last) {
      ts.ob[j] = __ldg(seg_off + v);
      ts.oe[j] = __ldg(seg_off + v + 1);
    }
  }
}
template <int kLanes = 8>
__device__ __forceinline__ TileSegs tile_segments_begin(long long row0, long long M, const int32_t* __restrict__ seg_off,
                                                        const int32_t* __restrict__ seg_id, int mt) {
  TileSegs ts = tile_segments_ids(row0, M, seg_id);
  tile_segments_bounds<kLanes>(ts, seg_off, mt);
  return ts;
}

// kA rounds of this thread's segments are summed side by side (kS rows of each in flight), the rest one at a time.
template <int kA = kSegAhead, int kS = kAggSide, int kLanes = 8>
__device__ __forceinline__ void tile_segment_sum(const uint8_t* buf, long long row0, const TileSegs& ts,
                                                 const int32_t* __restrict__ seg_off, bf16* __restrict__ out,
                                                 long long ld_out, float* __restrict__ part, int32_t* __restrict__ part_v,
                                                 int mt, long long row_base = 0, long long rec_base = 0) {
  const int nrows = ts.nrows, v_first = ts.v_first, v_last = ts.v_last;
  const int chunk = mt & 15, sl = mt >> 4;
  const long long tile = rec_base + row0 / kRows;
  const uint8_t* col = buf + (chunk >> 3) * kPB;
  const long long g0 = row_base + row0;  // global position of the tile's first row
  auto put_segment = [&](int v, const uint64_t (&acc)[4]) {
    if (v == v_first || v == v_last) {
      const long long rec = tile * 2 + ((v == v_last && v != v_first) ? 1 : 0);
      float4* d = reinterpret_cast<float4*>(part + rec * kH + chunk * 8);
      d[0] = make_float4(f2_lo(acc[0]), f2_hi(acc[0]), f2_lo(acc[1]), f2_hi(acc[1]));
      d[1] = make_float4(f2_lo(acc[2]), f2_hi(acc[2]), f2_lo(acc[3]), f2_hi(acc[3]));
      if (chunk == 0) part_v[rec] = v;
    } else {
      *reinterpret_cast<uint4*>(out + static_cast<long long>(v) * ld_out + chunk * 8) =
          make_uint4(f2_to_bf16x2(acc[0]), f2_to_bf16x2(acc[1]), f2_to_bf16x2(acc[2]), f2_to_bf16x2(acc[3]));
    }
  };
  auto add_row = [&](uint64_t (&acc)[4], const uint4& t) {
    acc[0] = f2_add(acc[0], f2_from_bf16x2(t.x));
    acc[1] = f2_add(acc[1], f2_from_bf16x2(t.y));
    acc[2] = f2_add(acc[2], f2_from_bf16x2(t.z));
    acc[3] = f2_add(acc[3], f2_from_bf16x2(t.w));
  };
  auto row_ld = [&](int r) { return *reinterpret_cast<const uint4*>(col + sw128_offset(r, chunk & 7)); };
  // The thread's first kA segments side by side: kS rows of EACH in flight per pass (a mesh tile is ~22
  // segments of ~6 rows = 3 per thread: two round trips to shared memory for the whole tile instead of two per segment,
  // one segment after the other).  Rows past a segment's end read as +0: every sum is still taken in ascending row order.
  {
    int b[kA], e[kA];
    uint64_t acc[kA][4];
    int longest = 0;
#pragma unroll
    for (int j = 0; j < kA; ++j) {
      const long long ob = ts.ob[j], oe = ts.oe[j];
      const bool on = v_first + sl + kLanes * j <= v_last;
      b[j] = on ? static_cast<int>((ob > g0 ? ob : g0) - g0) : 0;
      e[j] = on ? static_cast<int>((oe < g0 + nrows ? oe : g0 + nrows) - g0) : 0;
      longest = max(longest, e[j] - b[j]);
      acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0ull;
    }
    for (int base = 0; base < longest; base += kS) {
      uint4 t[kA][kS];
#pragma unroll
      for (int j = 0; j < kA; ++j)
#pragma unroll
        for (int u = 0; u < kS; ++u)
          t[j][u] = b[j] + base + u < e[j] ? row_ld(b[j] + base + u) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
      for (int j = 0; j < kA; ++j)
#pragma unroll
        for (int u = 0; u < kS; ++u) add_row(acc[j], t[j][u]);
    }
#pragma unroll
    for (int j = 0; j < kA; ++j)
      if (v_first + sl + kLanes * j <= v_last) put_segment(v_first + sl + kLanes * j, acc[j]);
  }
  int v = v_first + sl + kLanes * kA;
  if (v <= v_last) {  // tiles of many short (or empty) segments: one at a time, bounds fetched a round ahead
    long long ob = __ldg(seg_off + v), oe = __ldg(seg_off + v + 1);
    for (; v <= v_last; v += kLanes) {
      const int b = static_cast<int>((ob > g0 ? ob : g0) - g0);
      const int e = static_cast<int>((oe < g0 + nrows ? oe : g0 + nrows) - g0);
      if (v + kLanes <= v_last) {
        ob = __ldg(seg_off + v + kLanes);
        oe = __ldg(seg_off + v + kLanes + 1);
      }
      uint64_t acc[4] = {0ull, 0ull, 0ull, 0ull};
      for (int r = b; r < e; r += kAggFly) {
        uint4 t[kAggFly];
#pragma unroll
        for (int u = 0; u < kAggFly; ++u) t[u] = r + u < e ? row_ld(r + u) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int u = 0; u < kAggFly; ++u) add_row(acc, t[u]);
      }
      put_segment(v, acc);
    }
  }
  if (v_first == v_last && mt == 0) part_v[tile * 2 + 1] = -1;
}

// Combine the per-tile boundary records of the fused aggregation (see Params::seg_off): one warp per record; the
// first record of a run of equal destination ids sums the run in record (= tile) order, writes the row, and
// zero-fills the destination rows that fall between two tiles (nodes without incoming edges).
static __global__ void __launch_bounds__(256) agg_fixup_kernel(const float* __restrict__ part, const int32_t* __restrict__ part_v,
                                                        long long n_rec, bf16* __restrict__ agg, long long ld_agg,
                                                        long long n_seg) {
  const long long r = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= n_rec) return;
  const int v = part_v[r];
  if (v < 0) return;
  long long q = r - 1;
  while (q >= 0 && part_v[q] < 0) --q;
  const int vp = q >= 0 ? part_v[q] : -1;
  if (vp == v) return;  // not the head of its run
  const uint2 zero = make_uint2(0u, 0u);
  // (only a tile's FIRST record can have unwritten rows before it: everything between a tile's first and last segment
  //  was written by the main kernel, empty segments included)
  if ((r & 1) == 0)
    for (long long g = static_cast<long long>(vp) + 1; g < v; ++g) *reinterpret_cast<uint2*>(agg + g * ld_agg + lane * 4) = zero;
  float4 acc = reinterpret_cast<const float4*>(part + r * kH)[lane];
  long long k = r + 1;
  for (; k < n_rec; ++k) {
    const int vk = part_v[k];
    if (vk == v) {
      const float4 t = reinterpret_cast<const float4*>(part + k * kH)[lane];
      acc.x += t.x;
      acc.y += t.y;
      acc.z += t.z;
      acc.w += t.w;
    } else if (vk >= 0) {
      break;
    }
  }
  *reinterpret_cast<uint2*>(agg + static_cast<long long>(v) * ld_agg + lane * 4) =
      make_uint2(pack_bf16x2(acc.x, acc.y), pack_bf16x2(acc.z, acc.w));
  if (k >= n_rec)  // last run: trailing nodes without incoming edges
    for (long long g = static_cast<long long>(v) + 1; g < n_seg; ++g) *reinterpret_cast<uint2*>(agg + g * ld_agg + lane * 4) = zero;
}

}  // namespace agg
}  // namespace mgn
