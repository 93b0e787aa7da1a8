// mgn_halo.cu — halo rows pushed straight into the peers' memory over NVLink / NVSwitch.
//
// The partitioned MeshGraphNet exchanges, per layer and direction, the source-projection rows (forward) or their
// gradients (backward) of the halo nodes (physicsnemo/models/gnn_layers/distributed_graph.py:999-1011 ->
// distributed/utils.py indexed_all_to_all_v = gather + NCCL all-to-all + copy-in).  Here ONE launch does all of it:
// every thread reads 16 bytes of a row that some peer needs and stores them at that row's final place in the PEER's
// receive buffer (a peer-mapped address of a symmetric allocation), the last CTA to finish publishes an epoch number in
// every destination's flag slot (system-scope release), and the receiver waits for the epochs of the ranks it expects rows
// from (mgn_halo_wait, system-scope acquire) right before the launches that read the rows.  No packing buffer, no
// collective call, no copy-in.  Bounded waits: a peer that never arrives sets a status bit instead of hanging the GPU.
#include "mgn_common.cuh"

namespace mgn {
namespace halo {

constexpr int kMaxPeers = 16;

struct PushParams {
  const char* tab;  // rows to send: bytes [col0_b, col0_b + row_b) of row (idx ? idx[i] : i), row stride ld_b
  long long ld_b, col0_b;
  int row_b;  // bytes per row (multiple of 16)
  const int32_t* idx;
  long long n_rows;
  int n_peers;
  long long seg_begin[kMaxPeers + 1];          // rows [seg_begin[r], seg_begin[r + 1]) go to peer r
  unsigned long long dst_base[kMaxPeers];      // peer-mapped address of the first of those rows in r's buffer
  long long dst_ld_b;                          // row stride there
  unsigned long long flag_addr[kMaxPeers];     // peer-mapped address of this rank's flag slot in r's memory (0: none)
  int epoch;
  unsigned int* counter;  // zero between launches (the last CTA resets it)
};

__device__ __forceinline__ void st_release_sys(int* p, int v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_sys(const int* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(256) halo_push_kernel(const PushParams p) {
  const int chunks = p.row_b >> 4;
  const long long total = p.n_rows * chunks;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < total; t += stride) {
    const long long row = t / chunks;
    const int c = static_cast<int>(t - row * chunks);
    int r = 0;
    while (r + 1 < p.n_peers && row >= p.seg_begin[r + 1]) ++r;
    const long long src_row = p.idx ? static_cast<long long>(__ldg(p.idx + row)) : row;
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(p.tab + src_row * p.ld_b + p.col0_b) + c);
    char* dst = reinterpret_cast<char*>(p.dst_base[r]) + (row - p.seg_begin[r]) * p.dst_ld_b;
    reinterpret_cast<uint4*>(dst)[c] = v;
  }
  // release: this CTA's peer stores are ordered before its count; the last CTA's flags after every count
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned done = atomicAdd(p.counter, 1u);
    if (done == gridDim.x - 1) {
      *p.counter = 0u;
      __threadfence_system();
      for (int r = 0; r < p.n_peers; ++r)
        if (p.flag_addr[r] != 0ull) st_release_sys(reinterpret_cast<int*>(p.flag_addr[r]), p.epoch);
    }
  }
}

// one thread per expected source rank; ~10 s at 2 GHz before it gives up
__global__ void halo_wait_kernel(const int* flags, unsigned need_mask, int epoch, int* status) {
  const int r = threadIdx.x;
  if (r >= kMaxPeers || !((need_mask >> r) & 1u)) return;
  const long long t0 = clock64();
  while (ld_acquire_sys(flags + r) - epoch < 0) {
    if (clock64() - t0 > 20000000000LL) {
      if (status != nullptr) atomicOr(status, 8);
      break;
    }
    __nanosleep(100);
  }
}

}  // namespace halo
}  // namespace mgn

using namespace mgn;

extern "C" int mgn_halo_push(const void* tab, int64_t ld_bytes, int64_t col0_bytes, int64_t row_bytes, const int32_t* idx,
                             int64_t n_rows, int n_peers, const int64_t* seg_begin, const int64_t* dst_base,
                             int64_t dst_ld_bytes, const int64_t* flag_addr, int epoch, void* counter, mgn_stream_t stream) {
  MGN_CHECK_ARG(n_rows >= 0 && n_peers > 0 && n_peers <= halo::kMaxPeers && seg_begin && dst_base && flag_addr && counter);
  MGN_CHECK_ARG(row_bytes > 0 && row_bytes % 16 == 0 && ld_bytes % 16 == 0 && col0_bytes % 16 == 0 && dst_ld_bytes % 16 == 0);
  MGN_CHECK_ARG(n_rows == 0 || (tab != nullptr && (reinterpret_cast<uintptr_t>(tab) & 15) == 0));
  halo::PushParams p{};
  p.tab = static_cast<const char*>(tab);
  p.ld_b = ld_bytes;
  p.col0_b = col0_bytes;
  p.row_b = static_cast<int>(row_bytes);
  p.idx = idx;
  p.n_rows = n_rows;
  p.n_peers = n_peers;
  for (int r = 0; r <= n_peers; ++r) p.seg_begin[r] = seg_begin[r];
  MGN_CHECK_ARG(p.seg_begin[0] == 0 && p.seg_begin[n_peers] == n_rows);
  for (int r = 0; r < n_peers; ++r) {
    MGN_CHECK_ARG(p.seg_begin[r + 1] >= p.seg_begin[r]);
    p.dst_base[r] = static_cast<unsigned long long>(dst_base[r]);
    p.flag_addr[r] = static_cast<unsigned long long>(flag_addr[r]);
    MGN_CHECK_ARG(p.seg_begin[r + 1] == p.seg_begin[r] || (p.dst_base[r] != 0ull && (p.dst_base[r] & 15ull) == 0));
  }
  p.dst_ld_b = dst_ld_bytes;
  p.epoch = epoch;
  p.counter = static_cast<unsigned int*>(counter);
  const long long total = n_rows * (row_bytes / 16);
  long long blocks = (total + 255) / 256;
  const long long cap = 8LL * num_sms();
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;  // (a rank with nothing to send still publishes its epoch)
  halo::halo_push_kernel<<<static_cast<unsigned>(blocks), 256, 0, MGN_ST(as_stream(stream))>>>(p);
  return mgn_launch_status();
}

extern "C" int mgn_halo_wait(const void* flags, int need_mask, int epoch, int* status, mgn_stream_t stream) {
  MGN_CHECK_ARG(flags != nullptr);
  if (need_mask == 0) return MGN_OK;
  halo::halo_wait_kernel<<<1, 32, 0, MGN_ST(as_stream(stream))>>>(static_cast<const int*>(flags),
                                                                  static_cast<unsigned>(need_mask), epoch, status);
  return mgn_launch_status();
}
