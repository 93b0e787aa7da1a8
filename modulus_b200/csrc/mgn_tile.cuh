// mgn_tile.cuh — tile staging / epilogue helpers shared by the fused tensor-core MLP kernels.
//
// A tile is 128 rows of bf16 features held in shared memory as 64-column "panels" in the 128-byte
// swizzled layout UMMA descriptors read (mgn_tc.cuh).  Movers fill tiles with cp.async (L1-bypassing,
// fully asynchronous 16-byte gathers), epilogue threads own one tile row each (= one TMEM lane).
#pragma once
#include "mgn_common.cuh"
#include "mgn_tc.cuh"

namespace mgn {
namespace tile {

constexpr int kRows = 128;   // tile rows
constexpr int kPB = 16384;   // bytes per panel: 128 rows x 64 bf16

struct RowSrc {  // row r of the tile source lives at tab[(idx ? idx[r] : r) * ld + col0 ...]
  const bf16* tab;
  const int32_t* idx;
  long long ld;
  long long col0;
};

__device__ __forceinline__ bool wait_clk(uint64_t* bar, uint32_t parity) {
#ifdef MGN_WAIT_HINT
  // variant under test: hinted try_wait, clock tested once per 256 iterations (~6 instead of ~16 instructions per turn)
  if (mbar_try_wait_hint(bar, parity)) return true;
  const long long t0h = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait_hint(bar, parity)) {
    if ((++spins & 255u) == 0 && clock64() - t0h > 400000000LL) return false;
  }
  return true;
#endif
  if (mbar_try_wait(bar, parity)) return true;
  const long long t0 = clock64();
#ifdef MGN_WAIT_LEAN
  // variant under test: plain polling, clock tested once per 64 turns (~5 instead of ~16 instructions per turn)
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 63u) == 0 && clock64() - t0 > 400000000LL) return false;
  }
  return true;
#endif
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 400000000LL) return false;  // ~0.2 s: a wrong descriptor must not hang the box
#ifdef MGN_WAIT_SLEEP
    __nanosleep(MGN_WAIT_SLEEP);  // second A/B switch: back-off between polls (-DMGN_WAIT_SLEEP=ns), off by default
#endif
  }
  return true;
}

__device__ __forceinline__ uint4 ldg128(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

__device__ __forceinline__ uint4 add_bf16x8(uint4 a, uint4 b) {
  const uint32_t x[4] = {a.x, a.y, a.z, a.w}, y[4] = {b.x, b.y, b.z, b.w};
  uint32_t r[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 fa = unpack_bf16x2(x[i]), fb = unpack_bf16x2(y[i]);
    r[i] = pack_bf16x2(fa.x + fb.x, fa.y + fb.y);
  }
  return make_uint4(r[0], r[1], r[2], r[3]);
}

// fp32 [n_rows, k_true] (row stride ld) -> bf16 K-major SW128 panels [n_panels][128][64], zero padded
__device__ __forceinline__ void stage_weight_ld(uint8_t* dst, const float* __restrict__ w, long long ld, int n_rows,
                                                int k_true, int n_panels, int tid, int nthreads) {
  const int per_row = n_panels * 8;
  for (int item = tid; item < kRows * per_row; item += nthreads) {
    const int row = item / per_row;
    const int rem = item - row * per_row;
    const int panel = rem >> 3, chunk = rem & 7;
    const int k0 = panel * 64 + chunk * 8;
    float f[8];
    const float* src = w + row * ld + k0;
    if (row < n_rows && k0 + 8 <= k_true && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {  // two 16-byte loads
      const float4 a = __ldg(reinterpret_cast<const float4*>(src)), b = __ldg(reinterpret_cast<const float4*>(src) + 1);
      f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
      f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = (row < n_rows && (k0 + j) < k_true) ? __ldg(src + j) : 0.f;
    }
    uint4 v;
    v.x = pack_bf16x2(f[0], f[1]);
    v.y = pack_bf16x2(f[2], f[3]);
    v.z = pack_bf16x2(f[4], f[5]);
    v.w = pack_bf16x2(f[6], f[7]);
    *reinterpret_cast<uint4*>(dst + panel * kPB + sw128_offset(row, chunk)) = v;
  }
}

// movers: stage 2 panels of rows [row0, row0+128): v = s1[row] (+ s2[row]); rows >= M are zero.
// Row indices are fetched first, then all row loads of a batch are in flight together (a dependent
// idx -> row chain per row would serialise ~1 us round trips).
__device__ __forceinline__ void stage_rows_sum(uint8_t* buf, const RowSrc& s1, const RowSrc& s2, long long row0,
                                               long long M, int mt) {
  const int chunk = mt & 15, rsub = mt >> 4;
  const bool two = s2.tab != nullptr;
  int32_t r1[16], r2[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const long long grow = row0 + i * 8 + rsub;
    const long long gc = grow < M ? grow : M - 1;
    r1[i] = s1.idx ? __ldg(s1.idx + gc) : static_cast<int32_t>(gc);
    r2[i] = (two && s2.idx) ? __ldg(s2.idx + gc) : static_cast<int32_t>(gc);
  }
#pragma unroll
  for (int base = 0; base < 16; base += 8) {
    uint4 v1[8], v2[8];
#pragma unroll
    for (int u = 0; u < 8; ++u)
      v1[u] = ldg128(s1.tab + static_cast<long long>(r1[base + u]) * s1.ld + s1.col0 + chunk * 8);
    if (two) {
#pragma unroll
      for (int u = 0; u < 8; ++u)
        v2[u] = ldg128(s2.tab + static_cast<long long>(r2[base + u]) * s2.ld + s2.col0 + chunk * 8);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int row = (base + u) * 8 + rsub;
      uint4 v = two ? add_bf16x8(v1[u], v2[u]) : v1[u];
      if (row0 + row >= M) v = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(buf + (chunk >> 3) * kPB + sw128_offset(row, chunk & 7)) = v;
    }
  }
}

// movers: row ids of this thread's 16 rows (rows i*8 + rsub) of the tile starting at row0.  Issued EARLY (one
// tile ahead): the L1TEX queue returns loads in order, so an index load issued behind a batch of cp.async
// gathers only comes back after them -- a dependent idx -> gather chain costs a full memory round trip each.
__device__ __forceinline__ void fetch_row_ids(const int32_t* __restrict__ idx, long long row0, long long M, int rsub,
                                              int32_t (&r)[16]) {
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const long long grow = row0 + i * 8 + rsub;
    const long long gc = grow < M ? grow : M - 1;
    r[i] = idx ? __ldg(idx + gc) : static_cast<int32_t>(gc);
  }
}

// movers: stage 2 panels asynchronously (cp.async.cg: L1-bypassing 16-byte requests, zero fill past M) from the
// rows r[]; the caller commits / waits before publishing the tile
__device__ __forceinline__ void stage_rows_async(uint8_t* buf, const RowSrc& s, const int32_t (&r)[16], long long row0,
                                                 long long M, int mt) {
  const int chunk = mt & 15, rsub = mt >> 4;
  const uint32_t base = smem_u32(buf) + (chunk >> 3) * kPB;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int row = i * 8 + rsub;
    cp_async16_zfill(base + sw128_offset(row, chunk & 7), s.tab + static_cast<long long>(r[i]) * s.ld + s.col0 + chunk * 8,
                     row0 + row < M);
  }
}

// movers: raw [M, n_in] features (fp32 or bf16), zero padded to one 64-column panel
__device__ __forceinline__ void stage_small(uint8_t* buf, const void* x, int n_in, int is_f32, long long row0,
                                            long long M, int mt) {
  const int chunk = mt & 7, rsub = mt >> 3;
#pragma unroll 4
  for (int i = 0; i < 8; ++i) {
    const int row = i * 16 + rsub;
    const long long grow = row0 + row;
    float f[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int col = chunk * 8 + q;
      float v = 0.f;
      if (grow < M && col < n_in)
        v = is_f32 ? __ldg(static_cast<const float*>(x) + grow * n_in + col)
                   : __bfloat162float(static_cast<const bf16*>(x)[grow * n_in + col]);
      f[q] = v;
    }
    uint4 v4;
    v4.x = pack_bf16x2(f[0], f[1]);
    v4.y = pack_bf16x2(f[2], f[3]);
    v4.z = pack_bf16x2(f[4], f[5]);
    v4.w = pack_bf16x2(f[6], f[7]);
    *reinterpret_cast<uint4*>(buf + sw128_offset(row, chunk)) = v4;
  }
}

// movers: smem tile (2 panels) -> global rows [row0, ...) of a dense [M,128] bf16 matrix, coalesced
__device__ __forceinline__ void store_rows(const uint8_t* buf, bf16* dst, long long ld, long long row0, long long M,
                                           int mt) {
  const int chunk = mt & 15, rsub = mt >> 4;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int row = i * 8 + rsub;
    const long long grow = row0 + row;
    if (grow < M) {
      const uint4 v = *reinterpret_cast<const uint4*>(buf + (chunk >> 3) * kPB + sw128_offset(row, chunk & 7));
      *reinterpret_cast<uint4*>(dst + grow * ld + chunk * 8) = v;
    }
  }
}

// movers: acc[j] += sum over this thread's rows of tile[row][chunk*8 + j]
__device__ __forceinline__ void colsum_tile(const uint8_t* buf, int mt, float (&acc)[8]) {
  const int chunk = mt & 15, rsub = mt >> 4;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int row = i * 8 + rsub;
    const uint4 v = *reinterpret_cast<const uint4*>(buf + (chunk >> 3) * kPB + sw128_offset(row, chunk & 7));
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = unpack_bf16x2(w[e]);
      acc[2 * e] += f.x;
      acc[2 * e + 1] += f.y;
    }
  }
}

// epilogue: 32 bf16 of this thread's row (column group g) <-> registers
__device__ __forceinline__ void row_load32(const uint8_t* buf, int row, int g, float (&f)[32]) {
  const uint8_t* base = buf + (g >> 1) * kPB;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const uint4 v = *reinterpret_cast<const uint4*>(base + sw128_offset(row, (g & 1) * 4 + u));
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 t = unpack_bf16x2(w[e]);
      f[u * 8 + 2 * e] = t.x;
      f[u * 8 + 2 * e + 1] = t.y;
    }
  }
}
__device__ __forceinline__ void row_store32(uint8_t* buf, int row, int g, const float (&f)[32]) {
  uint8_t* base = buf + (g >> 1) * kPB;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    uint4 v;
    v.x = pack_bf16x2(f[u * 8 + 0], f[u * 8 + 1]);
    v.y = pack_bf16x2(f[u * 8 + 2], f[u * 8 + 3]);
    v.z = pack_bf16x2(f[u * 8 + 4], f[u * 8 + 5]);
    v.w = pack_bf16x2(f[u * 8 + 6], f[u * 8 + 7]);
    *reinterpret_cast<uint4*>(base + sw128_offset(row, (g & 1) * 4 + u)) = v;
  }
}

// 16 bf16 of this thread's row starting at column col (multiple of 16) <-> registers
__device__ __forceinline__ void row_load16(const uint8_t* buf, int row, int col, float (&f)[16]) {
  const uint8_t* base = buf + (col >> 6) * kPB;
  const int c8 = (col & 63) >> 3;
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const uint4 v = *reinterpret_cast<const uint4*>(base + sw128_offset(row, c8 + u));
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 t = unpack_bf16x2(w[e]);
      f[u * 8 + 2 * e] = t.x;
      f[u * 8 + 2 * e + 1] = t.y;
    }
  }
}
__device__ __forceinline__ void row_store16(uint8_t* buf, int row, int col, const float (&f)[16]) {
  uint8_t* base = buf + (col >> 6) * kPB;
  const int c8 = (col & 63) >> 3;
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    uint4 v;
    v.x = pack_bf16x2(f[u * 8 + 0], f[u * 8 + 1]);
    v.y = pack_bf16x2(f[u * 8 + 2], f[u * 8 + 3]);
    v.z = pack_bf16x2(f[u * 8 + 4], f[u * 8 + 5]);
    v.w = pack_bf16x2(f[u * 8 + 6], f[u * 8 + 7]);
    *reinterpret_cast<uint4*>(base + sw128_offset(row, c8 + u)) = v;
  }
}

// ---- packed variants (32 bf16 of this thread's row at column col, a multiple of 32, as 16 bf16x2 words) ----
__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }
__device__ __forceinline__ void row_load32p(const uint8_t* buf, int row, int col, uint32_t (&w)[16]) {
  const uint8_t* base = buf + (col >> 6) * kPB;
  const int c8 = (col & 63) >> 3;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const uint4 v = *reinterpret_cast<const uint4*>(base + sw128_offset(row, c8 + u));
    w[4 * u] = v.x;
    w[4 * u + 1] = v.y;
    w[4 * u + 2] = v.z;
    w[4 * u + 3] = v.w;
  }
}
__device__ __forceinline__ void row_store32p(uint8_t* buf, int row, int col, const uint32_t (&w)[16]) {
  uint8_t* base = buf + (col >> 6) * kPB;
  const int c8 = (col & 63) >> 3;
#pragma unroll
  for (int u = 0; u < 4; ++u)
    *reinterpret_cast<uint4*>(base + sw128_offset(row, c8 + u)) = make_uint4(w[4 * u], w[4 * u + 1], w[4 * u + 2], w[4 * u + 3]);
}

// ---- two fp32 lanes per instruction (sm_100a: add / mul / fma .f32x2 = SASS FADD2 / FMUL2 / FFMA2).  The fp32 pipe does
// not get faster (tools/probe_ffma2.cu: 128 FMA / cycle / SM either way) but every pair costs ONE issue slot, and the
// epilogue passes are issue-bound.  A pair is an even-aligned 64-bit register pair: packing is free when the compiler can
// place the two halves next to each other (tcgen05.ld outputs, unpacked bf16x2 words).
__device__ __forceinline__ uint64_t f2_pack(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ uint64_t f2_packu(uint32_t lo, uint32_t hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
__device__ __forceinline__ uint64_t f2_splat(float x) { return f2_pack(x, x); }
__device__ __forceinline__ float f2_lo(uint64_t v) { return __uint_as_float(static_cast<uint32_t>(v)); }
__device__ __forceinline__ float f2_hi(uint64_t v) { return __uint_as_float(static_cast<uint32_t>(v >> 32)); }
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t f2_mul(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
// bf16x2 word -> fp32 pair (exact), fp32 pair -> bf16x2 word (round to nearest even, one F2FP)
__device__ __forceinline__ uint64_t f2_from_bf16x2(uint32_t w) { return f2_packu(w << 16, w & 0xFFFF0000u); }
__device__ __forceinline__ uint32_t f2_to_bf16x2(uint64_t v) { return pack_bf16x2(f2_lo(v), f2_hi(v)); }
__device__ __forceinline__ uint64_t f2_ld(const float* p) { return *reinterpret_cast<const uint64_t*>(p); }  // 8-byte aligned

// ---- packed bf16x2 helpers (one instruction each: HMNMX2 / HSET2 + LOP3 / HFMA2)
__device__ __forceinline__ uint32_t relu_bf16x2(uint32_t w) {
  const __nv_bfloat162 z = __floats2bfloat162_rn(0.f, 0.f);
  __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&w), z);
  return *reinterpret_cast<uint32_t*>(&r);
}
// keep the halves of `w` whose counterpart in `h` is > 0 (ReLU mask from the stored activation), zero the others
__device__ __forceinline__ uint32_t mask_pos_bf16x2(uint32_t w, uint32_t h) {
  const __nv_bfloat162 z = __floats2bfloat162_rn(0.f, 0.f);
  return w & __hgt2_mask(*reinterpret_cast<__nv_bfloat162*>(&h), z);
}
// a + b per half with ONE rounding (= bf16(exact sum), what rounding the fp32 sum gives except in double-rounding ties)
__device__ __forceinline__ uint32_t add_bf16x2(uint32_t a, uint32_t b) {
  __nv_bfloat162 r = __hadd2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
  return *reinterpret_cast<uint32_t*>(&r);
}

// 16 bf16 of this thread's row at column col (a multiple of 16) as 8 bf16x2 words
__device__ __forceinline__ void row_load16p(const uint8_t* buf, int row, int col, uint32_t (&w)[8]) {
  const uint8_t* base = buf + (col >> 6) * kPB;
  const int c8 = (col & 63) >> 3;
  const uint4 a = *reinterpret_cast<const uint4*>(base + sw128_offset(row, c8));
  const uint4 b = *reinterpret_cast<const uint4*>(base + sw128_offset(row, c8 + 1));
  w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w;
  w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
}
__device__ __forceinline__ void row_store16p(uint8_t* buf, int row, int col, const uint32_t (&w)[8]) {
  uint8_t* base = buf + (col >> 6) * kPB;
  const int c8 = (col & 63) >> 3;
  *reinterpret_cast<uint4*>(base + sw128_offset(row, c8)) = make_uint4(w[0], w[1], w[2], w[3]);
  *reinterpret_cast<uint4*>(base + sw128_offset(row, c8 + 1)) = make_uint4(w[4], w[5], w[6], w[7]);
}

// One pass over this thread's 64 accumulator columns in four 16-column chunks, f(chunk, v[16]).  The tcgen05.ld of chunk
// i + 1 is in flight while chunk i is processed: tcgen05.wait::ld waits for ALL outstanding loads, so the next load is
// issued right after the wait and before the arithmetic of the chunk that just arrived.
template <typename F>
__device__ __forceinline__ void tmem_pass64(uint32_t taddr, F&& f) {
  uint32_t va[16], vb[16];
  tmem_ld16(taddr, va);
  tmem_ld_wait();
  tmem_ld16(taddr + 16, vb);
  f(0, va);
  tmem_ld_wait();
  tmem_ld16(taddr + 32, va);
  f(1, vb);
  tmem_ld_wait();
  tmem_ld16(taddr + 48, vb);
  f(2, va);
  tmem_ld_wait();
  f(3, vb);
}

// warp transpose-reduce of 16 columns: on return lane L holds the sum over all 32 lanes of their v[L & 15]
__device__ __forceinline__ float warp_colsum16(float (&v)[16], int lane) {
#pragma unroll
  for (int off = 8; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = up ? v[i] : v[i + off];
      const float keep = up ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 16);
}

// warp transpose-reduce: on return lane L holds sum over the 32 lanes of their v[L]   (31 shuffles)
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = up ? v[i] : v[i + off];
      const float keep = up ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}


}  // namespace tile
}  // namespace mgn
