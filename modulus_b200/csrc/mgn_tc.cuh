// mgn_tc.cuh — sm_100a tensor-core plumbing used by the fused MeshGraphNet kernels:
// tcgen05.mma / TMEM allocation / tcgen05.ld|st / mbarrier / proxy fences, plus the
// shared-memory operand layout (128-byte swizzle) and its UMMA descriptors.
//
// Everything here is inline PTX; there is no library dependency.  Layout facts
// (checked on hardware by tools/probe_tc.cu):
//
//  * An operand tile is stored as 64-column bf16 "panels": panel = [rows][64] bf16,
//    one row = 128 B, 8 rows = one 1024-B swizzle atom, 16-B chunk c of row r lives at
//    chunk position (c ^ (r & 7)).  Panels must be 1024-B aligned.
//  * The SAME bytes serve as a K-major operand (rows = M/N index, columns = K) and
//    as an MN-major operand (rows = K index, columns = M/N index); only the
//    descriptor differs.  That is what lets one smem copy of an activation tile feed
//    the forward GEMM, the dgrad GEMM and the wgrad GEMM.
#pragma once
#include <cstdint>
#include <cuda_bf16.h>

namespace mgn {

// ----------------------------------------------------------------------------------
// shared-memory addressing
// ----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

constexpr int kPanelCols = 64;          // bf16 columns per panel (128 B per row)
constexpr int kPanelRowBytes = 128;

// byte offset of (row, 16-byte chunk) inside one SW128 panel
__device__ __forceinline__ uint32_t sw128_offset(int row, int chunk) {
  return static_cast<uint32_t>(row) * kPanelRowBytes +
         (static_cast<uint32_t>(chunk ^ (row & 7)) << 4);
}

// ----------------------------------------------------------------------------------
// UMMA descriptors
// ----------------------------------------------------------------------------------
// Shared-memory matrix descriptor (64 bit):
//  [0,14)  start address >> 4      [16,30) leading byte offset >> 4
//  [32,46) stride byte offset >> 4 [46,48) version = 1 (sm_100)
//  [61,64) layout type (2 = SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t saddr, uint32_t lbo_bytes,
                                                   uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// K-major operand: rows are the M (or N) index, the K extent of one MMA (16 bf16 =
// 32 B) lies inside a 128-B row.  SBO = distance between 8-row groups.
__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t panel_addr, int k16_in_panel) {
  return umma_smem_desc(panel_addr + k16_in_panel * 32, 16, 1024);
}

// MN-major operand: rows are the K index; one MMA consumes 16 rows (2 atoms, SBO apart);
// the M/N extent beyond 64 columns continues in the next panel (LBO apart).
__device__ __forceinline__ uint64_t umma_desc_mnmajor(uint32_t panel_addr, int k16_row_block,
                                                      uint32_t panel_stride_bytes) {
  return umma_smem_desc(panel_addr + k16_row_block * 2048, panel_stride_bytes, 1024);
}

// Instruction descriptor for kind::f16, bf16 x bf16 -> fp32.
//  [4,6) c_format=1 (f32)  [7,10) a_format=1 (bf16)  [10,13) b_format=1 (bf16)
//  [15] a_major (1 = MN)   [16] b_major (1 = MN)   [17,23) N>>3   [24,29) M>>4
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(a_mn) << 15) |
         (static_cast<uint32_t>(b_mn) << 16) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}

// ----------------------------------------------------------------------------------
// tcgen05.mma (single thread issues on behalf of the CTA)
// ----------------------------------------------------------------------------------
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                        uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// A operand read from TMEM (lane = row, two bf16 per 32-bit column), B from smem.
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc,
                                        uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Arrive on an mbarrier once every MMA issued so far by this thread has completed.
// (implies tcgen05.fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void tc_fence_before_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// make generic-proxy smem writes (st.shared / cp.async) visible to the async proxy (UMMA)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ----------------------------------------------------------------------------------
// TMEM allocation (one full warp), 32 <= cols <= 512, power of two
// ----------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_dst)),
               "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols)
               : "memory");
}

// ----------------------------------------------------------------------------------
// TMEM <-> registers.  Warp w may only touch lanes [32*(w%4), 32*(w%4)+32): thread t of
// the warp owns lane 32*(w%4)+t, registers are consecutive 32-bit columns.
// ----------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
        "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]),
        "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
      "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]),
      "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
      : "memory");
}
__device__ __forceinline__ void tmem_st2(uint32_t taddr, uint32_t a, uint32_t b) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void tmem_ld2(uint32_t taddr, uint32_t& a, uint32_t& b) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// ----------------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)
               : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok;
}
#ifdef MGN_WAIT_HINT
// A/B switch for round 2 (profiles/r01_issue_slots.md): try_wait with a suspend-time hint of MGN_WAIT_HINT ns, so the
// hardware parks a waiting warp instead of returning to the polling loop every ~200 cycles.  Off by default: build
// with MGN_NVCC_EXTRA="-DMGN_WAIT_HINT=20000" python -m modulus_b200.build to try it.
__device__ __forceinline__ uint32_t mbar_try_wait_hint(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(static_cast<uint32_t>(MGN_WAIT_HINT))
      : "memory");
  return ok;
}
#endif
// Bounded wait: a wrong descriptor must never hang the GPU box.  Returns false on
// timeout (callers flag an error and bail out).
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
#ifdef MGN_UNBOUNDED_WAIT
  while (!mbar_try_wait(bar, parity)) {
  }
  return true;
#else
  for (uint32_t it = 0; it < (1u << 20); ++it)
    if (mbar_try_wait(bar, parity)) return true;
  return false;
#endif
}

// ----------------------------------------------------------------------------------
// cp.async (LDGSTS), 16 bytes, L2-only caching (streamed rows)
// ----------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(uint32_t saddr, const void* gptr) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async16_zfill(uint32_t saddr, const void* gptr, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(saddr), "l"(gptr), "r"(sz)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() {
  asm volatile("cp.async.commit_group;" ::: "memory");
}
// every cp.async this thread has issued so far arrives on `bar` when it lands (the barrier's expected count includes
// this arrival: .noinc); the thread itself does not wait
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ----------------------------------------------------------------------------------
// small numeric helpers
// ----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t v) {
  __nv_bfloat162 h = *reinterpret_cast<__nv_bfloat162*>(&v);
  return __bfloat1622float2(h);
}

}  // namespace mgn
