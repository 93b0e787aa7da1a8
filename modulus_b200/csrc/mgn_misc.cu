// mgn_misc.cu — version / error strings of the C ABI.
#include "mgn_common.cuh"

extern "C" int mgn_version(void) { return 100; }

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
