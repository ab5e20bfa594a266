// api_internal.h — error plumbing shared by the C-ABI entry points of libb200ltx.
// Contract (include/b200ltx.h): 0 = success, negative = argument/shape/alignment violation
// (nothing was launched), positive = cudaError_t of the launch.  Never throws, never aborts.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

namespace b200 {

inline char* last_error_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}
inline int arg_error(const char* msg, int code = -1) {
  snprintf(last_error_buf(), 512, "%s", msg);
  return code;
}
inline int launch_status(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(last_error_buf(), 512, "%s: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

}  // namespace b200
