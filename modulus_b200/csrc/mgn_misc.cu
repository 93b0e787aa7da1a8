// mgn_misc.cu — version / error strings of the C ABI.
#include "mgn_common.cuh"

namespace mgn {
unsigned long long g_mgn_launches = 0;
}

extern "C" int mgn_version(void) { return 200; }

#ifndef MGN_BUILD_DIGEST
#define MGN_BUILD_DIGEST "MGNDIGEST:unknown"
#endif
// "MGNDIGEST:<sha256>": the marker lets build.py find the digest in the file without loading it
extern "C" const char* mgn_build_digest(void) { return MGN_BUILD_DIGEST + 10; }

extern "C" int64_t mgn_launch_count(void) {
  return static_cast<int64_t>(__atomic_load_n(&mgn::g_mgn_launches, __ATOMIC_RELAXED));
}

extern "C" const char* mgn_error_string(int code) {
  switch (code) {
    case MGN_OK: return "ok";
    case MGN_EINVAL: return "invalid argument";
    case MGN_EUNSUPPORTED: return "unsupported shape or dtype for this kernel family";
    case MGN_EALIGN: return "pointer or leading dimension not 16-byte aligned";
    case MGN_EWORKSPACE: return "workspace too small";
    default: return code > 0 ? cudaGetErrorString(static_cast<cudaError_t>(code)) : "unknown error";
  }
}
