// api.cu — library-level entry points of libb200ltx (version, last error, device probe).
#include "api_internal.h"

extern "C" const char* b200_last_error(void) { return b200::last_error_buf(); }
extern "C" int b200_version(void) { return 100; }  // 0.1.0

// 0 when the current device can run the kernels (compute capability 10.x), negative otherwise.
extern "C" int b200_device_check(void) {
  int dev = 0, major = 0, minor = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return b200::arg_error("device_check: no CUDA device", -2);
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10) return b200::arg_error("device_check: kernels are built for sm_100a only", -3);
  return 0;
}
