// Shared by every translation unit of libsam2b200.so: error codes + last-error string.
// The C ABI never throws: entry points return 0 or a negative code and record a message that
// sam2b200_last_error() returns (thread-local).
#pragma once

#include <cuda_runtime.h>
#include <stdio.h>

#define SAM2B200_OK 0
#define SAM2B200_ERR_INVALID (-1)
#define SAM2B200_ERR_CUDA (-2)
#define SAM2B200_ERR_UNSUPPORTED (-3)
#define SAM2B200_ERR_DRIVER (-4)

namespace sam2b200 {

char* last_error_buffer();  // defined in abi.cu (thread-local, 512 bytes)
void count_launches(int n);  // defined in abi.cu: kernels launched by this library (process-wide)

inline int fail(int code, const char* msg) {
  snprintf(last_error_buffer(), 512, "%s", msg);
  return code;
}

inline int check_launch(const char* what, int n_kernels = 1) {
  count_launches(n_kernels);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(last_error_buffer(), 512, "%s: %s", what, cudaGetErrorString(e));
    return SAM2B200_ERR_CUDA;
  }
  return SAM2B200_OK;
}

}  // namespace sam2b200
