// mgn_common.cuh — shared declarations for the MeshGraphNet hot-path library (sm_100a).
#pragma once
#include <cstdint>
#include <cstddef>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "../../include/mgn_b200.h"

namespace mgn {

typedef __nv_bfloat16 bf16;

#define MGN_CHECK_ARG(cond) \
  do {                      \
    if (!(cond)) return MGN_EINVAL; \
  } while (0)

// every extern "C" entry returns 0, a negative argument error, or a positive cudaError_t
static inline int mgn_launch_status() {
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? MGN_OK : static_cast<int>(e);
}

static inline cudaStream_t as_stream(mgn_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// number of kernels launched by this library since load (bench.py reports the delta over its timed
// region as "gpu_launches"); every launch site passes its stream through MGN_ST()
extern unsigned long long g_mgn_launches;
static inline cudaStream_t count_launch(cudaStream_t s) {
  __atomic_fetch_add(&g_mgn_launches, 1ull, __ATOMIC_RELAXED);
  return s;
}
#define MGN_ST(s) ::mgn::count_launch(s)

// Per-device caches: the library may be called for several devices of one process (the caller selects the device,
// e.g. `with torch.cuda.device(t.device)`); function attributes and the SM count belong to the CURRENT device.
constexpr int kMaxDevices = 64;
static inline int current_device_slot() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev >= 0 && dev < kMaxDevices) ? dev : 0;
}
struct PerDeviceFlag {  // `static PerDeviceFlag f; bool& done = f.get();`
  bool done[kMaxDevices] = {};
  bool& get() { return done[current_device_slot()]; }
};

static inline int num_sms() {
  static int n[kMaxDevices] = {};
  const int slot = current_device_slot();
  if (n[slot] == 0) {
    int dev = 0, v = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    n[slot] = v > 0 ? v : 148;
  }
  return n[slot];
}

template <typename T> struct Num;
template <> struct Num<float> {
  static __device__ __forceinline__ float to_f(float v) { return v; }
  static __device__ __forceinline__ float from_f(float v) { return v; }
  static constexpr int kVec = 4;  // elements per 16 bytes
};
template <> struct Num<bf16> {
  static __device__ __forceinline__ float to_f(bf16 v) { return __bfloat162float(v); }
  static __device__ __forceinline__ bf16 from_f(float v) { return __float2bfloat16_rn(v); }
  static constexpr int kVec = 8;
};

// 16-byte vector of T with fp32 pack/unpack
template <typename T> struct Vec16;
template <> struct Vec16<float> {
  static constexpr int N = 4;
  uint4 raw;
  __device__ __forceinline__ void unpack(float (&f)[4]) const {
    f[0] = __uint_as_float(raw.x); f[1] = __uint_as_float(raw.y);
    f[2] = __uint_as_float(raw.z); f[3] = __uint_as_float(raw.w);
  }
  __device__ __forceinline__ void pack(const float (&f)[4]) {
    raw.x = __float_as_uint(f[0]); raw.y = __float_as_uint(f[1]);
    raw.z = __float_as_uint(f[2]); raw.w = __float_as_uint(f[3]);
  }
};
template <> struct Vec16<bf16> {
  static constexpr int N = 8;
  uint4 raw;
  __device__ __forceinline__ void unpack(float (&f)[8]) const {
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
    }
  }
  __device__ __forceinline__ void pack(const float (&f)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    raw = make_uint4(w[0], w[1], w[2], w[3]);
  }
};

// activation functions of the reference's get_activation table that the MGN recipes use
// (physicsnemo/models/layers/activations.py:173-199)
__device__ __forceinline__ float act_fwd(int act, float x) {
  switch (act) {
    case MGN_ACT_RELU: return x > 0.f ? x : 0.f;
    case MGN_ACT_SILU: return x / (1.f + expf(-x));
    case MGN_ACT_TANH: return tanhf(x);
    case MGN_ACT_SIGMOID: return 1.f / (1.f + expf(-x));
    case MGN_ACT_GELU: return 0.5f * x * (1.f + erff(x * 0.70710678118654752f));
    case MGN_ACT_LEAKY_RELU: return x > 0.f ? x : 0.1f * x;  // slope of the reference table (activations.py:175)
    case MGN_ACT_ELU: return x > 0.f ? x : (expf(x) - 1.f);
    default: return x;
  }
}
// derivative w.r.t. the pre-activation x
__device__ __forceinline__ float act_grad(int act, float x) {
  switch (act) {
    case MGN_ACT_RELU: return x > 0.f ? 1.f : 0.f;
    case MGN_ACT_SILU: {
      const float s = 1.f / (1.f + expf(-x));
      return s * (1.f + x * (1.f - s));
    }
    case MGN_ACT_TANH: {
      const float t = tanhf(x);
      return 1.f - t * t;
    }
    case MGN_ACT_SIGMOID: {
      const float s = 1.f / (1.f + expf(-x));
      return s * (1.f - s);
    }
    case MGN_ACT_GELU:
      return 0.5f * (1.f + erff(x * 0.70710678118654752f)) +
             x * 0.3989422804014327f * expf(-0.5f * x * x);
    case MGN_ACT_LEAKY_RELU: return x > 0.f ? 1.f : 0.1f;
    case MGN_ACT_ELU: return x > 0.f ? 1.f : expf(x);
    default: return 1.f;
  }
}

}  // namespace mgn
