// Library-level entry points of libsam2b200.so (version, last error, device probe).
#include <atomic>

#include "abi_common.cuh"

namespace sam2b200 {
char* last_error_buffer() {
  static thread_local char buf[512] = {0};
  return buf;
}
static std::atomic<long long> g_launches{0};
void count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace sam2b200

extern "C" {

// Number of CUDA kernels this library has launched in this process (bench.py's gpu_launches).
long long sam2b200_launch_count(void) { return sam2b200::g_launches.load(std::memory_order_relaxed); }

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
