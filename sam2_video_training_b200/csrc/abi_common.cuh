// Shared by every translation unit of libsam2b200.so: error codes + last-error string.
// The C ABI never throws: entry points return 0 or a negative code and record a message that
// sam2b200_last_error() returns (thread-local).
#pragma once

#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

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

// Launch with the programmatic-dependent-launch attribute (see sm100.cuh pdl_wait / pdl_launch_dependents): the kernel may begin its
// set-up while the previous kernel of the stream drains.  Measured on the cfg2 step: worth 0.15 ms for the GEMM family (csrc/gemm.cu, on
// by default, SAM2B200_NO_PDL=1 switches it off), but NOT for the LayerNorm passes and mlp_dh (40.56 / 40.94 vs 40.50 / 40.46 ms with
// the attribute on those too: their early-launched CTAs take SM slots from the side stream) -- those use it only with SAM2B200_PDL_ALL=1.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  static const bool pdl = getenv("SAM2B200_PDL_ALL") != nullptr && getenv("SAM2B200_NO_PDL") == nullptr;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace sam2b200
