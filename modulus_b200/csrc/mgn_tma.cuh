// mgn_tma.cuh — Tensor Memory Accelerator plumbing of the fused MeshGraphNet kernels (sm_100a).
//
// Feature tables are row-major bf16 [rows, ld]; a tile buffer in shared memory is two 64-column panels in the
// 128-byte-swizzle layout (mgn_tc.cuh).  One tensor map per table describes a 128-column window of it with a
// {64 columns x R rows} box and CU_TENSOR_MAP_SWIZZLE_128B, so that
//   * a dense tile is two bulk tensor loads (R = 128 rows each), rows past the end of the table arrive as zeros,
//   * a gathered tile (the concat_efeat / halo row gathers of the reference, models/gnn_layers/utils.py:94-148) is
//     32 x 2 `tile::gather4` loads (R = 1, four row indices per instruction),
//   * a result tile leaves through two bulk tensor stores, clipped at the end of the table,
// all issued by a single warp and completing on an mbarrier transaction count -- no LSU instruction slots, no
// register staging.  The driver entry point is looked up at run time (no link-time dependency on libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace mgn {

typedef CUresult (*TmaEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                     const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                     CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline TmaEncodeTiledFn tma_encoder() {
  static TmaEncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<TmaEncodeTiledFn>(p);
  }
  return fn;
}

// Tensor map over columns [0, cols) of a bf16 table starting at `base` (16-byte aligned) with `rows` rows and a row
// stride of `ld` elements (multiple of 8): box = {64 columns, box_rows}.  Returns 0 or a negative error.
static inline int tma_make_rows_map(CUtensorMap* m, const void* base, long long rows, long long ld, int box_rows,
                                    int cols = 128) {
  TmaEncodeTiledFn enc = tma_encoder();
  if (enc == nullptr) return -100;
  if (rows <= 0) rows = 1;
  const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  const cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * 2};
  const cuuint32_t box[2] = {64, static_cast<cuuint32_t>(box_rows)};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -101;
}

#ifdef __CUDACC__
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(
                   static_cast<uint32_t>(__cvta_generic_to_shared(bar))),
               "r"(bytes)
               : "memory");
}
// global -> shared, one {64 x box_rows} box at (col, row); completes `box bytes` on the mbarrier
__device__ __forceinline__ void tma_load_2d(uint32_t dst_smem, const CUtensorMap* map, int col, int row, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          dst_smem),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(static_cast<uint32_t>(__cvta_generic_to_shared(bar))), "r"(col), "r"(row)
      : "memory");
}
// global -> shared, four rows r0..r3 of 64 columns starting at col, written as four consecutive 128-byte smem rows
__device__ __forceinline__ void tma_gather4(uint32_t dst_smem, const CUtensorMap* map, int col, int r0, int r1, int r2,
                                            int r3, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, "
      "%5, %6, %7}], [%2];" ::"r"(dst_smem),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(static_cast<uint32_t>(__cvta_generic_to_shared(bar))), "r"(col),
      "r"(r0), "r"(r1), "r"(r2), "r"(r3)
      : "memory");
}
// global -> L2 only: the box will be requested by a later tma_load_2d (hides the HBM round trip)
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int col, int row) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(map)),
               "r"(col), "r"(row)
               : "memory");
}
// shared -> global, one {64 x box_rows} box at (col, row), clipped at the table bounds (bulk async-group)
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src_smem, int col, int row) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(src_smem), "r"(col), "r"(row)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all committed stores have finished READING shared memory (their buffers may be overwritten)
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// all committed stores are complete
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
#endif

}  // namespace mgn
