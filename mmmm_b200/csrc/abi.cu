// C-ABI housekeeping: version, error strings, device check.
#include "common.cuh"

namespace vex {
thread_local int g_last_cuda_error = 0;
}

extern "C" int vex_abi_version(void) { return VEX_ABI_VERSION; }

extern "C" const char* vex_error_string(int code) {
  switch (code) {
    case VEX_OK: return "ok";
    case VEX_E_INVALID: return "invalid argument";
    case VEX_E_UNSUPPORTED: return "unsupported shape";
    case VEX_E_CUDA: return "CUDA error (see vex_last_cuda_error)";
    case VEX_E_NO_DEVICE: return "no sm_100 device";
    default: return "unknown error";
  }
}

extern "C" int vex_last_cuda_error(void) { return vex::g_last_cuda_error; }

extern "C" int vex_device_check(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return VEX_E_NO_DEVICE;
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return VEX_E_NO_DEVICE;
  return major == 10 ? VEX_OK : VEX_E_NO_DEVICE;
}
