// Library-level entry points: ABI version, error strings, device check.
#include "common.cuh"

extern "C" int sl_abi_version(void) { return SL_ABI_VERSION; }

extern "C" const char* sl_error_string(int code) {
  switch (code) {
    case SL_OK: return "ok";
    case SL_EINVAL: return "segland_b200: size/shape argument out of the supported range";
    case SL_ENULL: return "segland_b200: required pointer is NULL";
    case SL_EALIGN: return "segland_b200: pointer or extent violates the required alignment";
    case SL_EUNSUPPORTED: return "segland_b200: device is not sm_100 (B200); no other target is built";
    default: break;
  }
  if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
  return "segland_b200: unknown error code";
}

extern "C" int sl_check_device(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return static_cast<int>(e);
  int major = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) return static_cast<int>(e);
  return major == 10 ? SL_OK : SL_EUNSUPPORTED;
}
