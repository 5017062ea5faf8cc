// Library-level entry points of libsam2b200.so (version, last error, device probe).
#include "abi_common.cuh"

namespace sam2b200 {
char* last_error_buffer() {
  static thread_local char buf[512] = {0};
  return buf;
}
}  // namespace sam2b200

extern "C" {

int sam2b200_version(void) { return 100; }  // 0.1.0

const char* sam2b200_last_error(void) { return sam2b200::last_error_buffer(); }

// 0 when device `dev` is an sm_100 part (the only target this library is built for).
int sam2b200_check_device(int dev) {
  cudaDeviceProp p;
  cudaError_t e = cudaGetDeviceProperties(&p, dev);
  if (e != cudaSuccess) return sam2b200::fail(SAM2B200_ERR_CUDA, cudaGetErrorString(e));
  if (p.major != 10) {
    snprintf(sam2b200::last_error_buffer(), 512,
             "device %d is sm_%d%d; libsam2b200 is built for sm_100a only", dev, p.major, p.minor);
    return SAM2B200_ERR_UNSUPPORTED;
  }
  return SAM2B200_OK;
}

}  // extern "C"
