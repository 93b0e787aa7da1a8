// probe_tma.cu — hardware check of the TMA forms used by the fused kernels (mgn_tma.cuh):
//   dense {64 x 128} box loads, tile::gather4 row gathers (incl. out-of-range row -> zeros) and box stores,
//   all against the 128-byte-swizzled panel layout the UMMA descriptors read (sw128_offset).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/bin/probe_tma tools/probe_tma.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../modulus_b200/csrc/mgn_tc.cuh"
#include "../modulus_b200/csrc/mgn_tma.cuh"
using namespace mgn;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2);} } while (0)
constexpr int kPB = 16384;
struct Maps { alignas(64) CUtensorMap dense; alignas(64) CUtensorMap gat; alignas(64) CUtensorMap out; };

__global__ void __launch_bounds__(128, 1) probe(const __grid_constant__ Maps m, const int* idx, int row0, int n_rows_table,
                                                uint16_t* dump_dense, uint16_t* dump_gat, int* err) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint8_t* bD = smem;
  uint8_t* bG = smem + 2 * kPB;
  if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_fence_init(); }
  __syncthreads();
  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(&bar[0], 2 * kPB);
      tma_load_2d(smem_u32(bD), &m.dense, 0, row0, &bar[0]);
      tma_load_2d(smem_u32(bD) + kPB, &m.dense, 64, row0, &bar[0]);
      mbar_arrive_expect_tx(&bar[1], 2 * kPB);
    }
    __syncwarp();
    int r[4];
    for (int j = 0; j < 4; ++j) r[j] = idx[4 * lane + j];
    tma_gather4(smem_u32(bG) + 4 * lane * 128, &m.gat, 0, r[0], r[1], r[2], r[3], &bar[1]);
    tma_gather4(smem_u32(bG) + kPB + 4 * lane * 128, &m.gat, 64, r[0], r[1], r[2], r[3], &bar[1]);
  }
  if (!mbar_wait(&bar[0], 0) || !mbar_wait(&bar[1], 0)) { if (tid == 0) *err = 1; return; }
  // de-swizzle with the kernels' own addressing and dump
  for (int i = tid; i < 128 * 128; i += 128) {
    const int row = i >> 7, col = i & 127;
    const uint32_t off = (col >> 6) * kPB + sw128_offset(row, (col & 63) >> 3) + (col & 7) * 2;
    dump_dense[i] = *reinterpret_cast<uint16_t*>(bD + off);
    dump_gat[i] = *reinterpret_cast<uint16_t*>(bG + off);
  }
  __syncthreads();
  // store the gathered tile to `out` rows [row0, row0+128)
  fence_proxy_async_smem();
  __syncthreads();
  if (tid == 0) {
    tma_store_2d(&m.out, smem_u32(bG), 0, row0);
    tma_store_2d(&m.out, smem_u32(bG) + kPB, 64, row0);
    tma_store_commit();
    tma_store_wait_all();
  }
}

int main() {
  const int R = 1000, LD = 384, COL0 = 128, row0 = 896;  // tile rows 896..1023: rows >= 1000 are out of range
  std::vector<uint16_t> h(static_cast<size_t>(R) * LD);
  for (int r = 0; r < R; ++r) for (int c = 0; c < LD; ++c) h[static_cast<size_t>(r) * LD + c] = static_cast<uint16_t>((r * 7 + c * 13) & 0x7fff);
  std::vector<int> hidx(128);
  for (int i = 0; i < 128; ++i) hidx[i] = (i * 37 + 5) % R;
  hidx[17] = R;       // out of range -> zeros expected
  hidx[126] = R + 5;  // out of range -> zeros expected
  uint16_t *d_tab, *d_out, *d_dd, *d_dg; int *d_idx, *d_err;
  CK(cudaMalloc(&d_tab, h.size() * 2)); CK(cudaMalloc(&d_out, static_cast<size_t>(R) * 128 * 2));
  CK(cudaMalloc(&d_dd, 128 * 128 * 2)); CK(cudaMalloc(&d_dg, 128 * 128 * 2)); CK(cudaMalloc(&d_idx, 512)); CK(cudaMalloc(&d_err, 4));
  CK(cudaMemcpy(d_tab, h.data(), h.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_idx, hidx.data(), 512, cudaMemcpyHostToDevice));
  CK(cudaMemset(d_out, 0xff, static_cast<size_t>(R) * 128 * 2)); CK(cudaMemset(d_err, 0, 4));
  Maps m;
  int rc = tma_make_rows_map(&m.dense, d_tab + COL0, R, LD, 128); if (rc) { printf("map dense rc=%d\n", rc); return 1; }
  rc = tma_make_rows_map(&m.gat, d_tab + COL0, R, LD, 1); if (rc) { printf("map gather rc=%d\n", rc); return 1; }
  rc = tma_make_rows_map(&m.out, d_out, R, 128, 128); if (rc) { printf("map out rc=%d\n", rc); return 1; }
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * kPB));
  probe<<<1, 128, 4 * kPB>>>(m, d_idx, row0, R, d_dd, d_dg, d_err);
  CK(cudaDeviceSynchronize());
  int err; CK(cudaMemcpy(&err, d_err, 4, cudaMemcpyDeviceToHost));
  std::vector<uint16_t> dd(128 * 128), dg(128 * 128), out(static_cast<size_t>(R) * 128);
  CK(cudaMemcpy(dd.data(), d_dd, dd.size() * 2, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(dg.data(), d_dg, dg.size() * 2, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(out.data(), d_out, out.size() * 2, cudaMemcpyDeviceToHost));
  int bad_d = 0, bad_g = 0, bad_s = 0;
  for (int r = 0; r < 128; ++r) for (int c = 0; c < 128; ++c) {
    const int gr = row0 + r;
    const uint16_t want_d = gr < R ? h[static_cast<size_t>(gr) * LD + COL0 + c] : 0;
    const uint16_t want_g = hidx[r] < R ? h[static_cast<size_t>(hidx[r]) * LD + COL0 + c] : 0;
    bad_d += dd[r * 128 + c] != want_d; bad_g += dg[r * 128 + c] != want_g;
    if (gr < R) bad_s += out[static_cast<size_t>(gr) * 128 + c] != want_g;
  }
  int untouched = 0; for (int r = 0; r < row0; ++r) untouched += out[static_cast<size_t>(r) * 128] == 0xffff;
  printf("timeout=%d dense mismatches=%d gather4 mismatches=%d store mismatches=%d rows below tile untouched=%d/%d\n", err, bad_d, bad_g, bad_s, untouched, row0);
  printf(bad_d + bad_g + bad_s + err == 0 && untouched == row0 ? "TMA PROBE OK\n" : "TMA PROBE FAILED\n");
  return 0;
}
