// Micro-probe (GPU box): cycles per tcgen05.mma for the instruction shapes the attention kernels
// use -- A from TMEM vs SMEM, N = 64 / 128 / 256, K-major vs MN-major B -- issued back to back by one
// thread per CTA, on 1 and on 148 CTAs.  Operand contents are irrelevant (uninitialised memory).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_probe mma_probe.cu && ./mma_probe
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include "sm100.cuh"

using namespace sm100;

struct Result { long long cycles; };

// mode: 0 = TS K-major B, 1 = SS K-major B, 2 = TS MN-major B (N = 256 slabs)
__global__ void __launch_bounds__(128, 1) probe_kernel(int mode, int n, int count, int tma_like_writes, Result* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&tmem_base, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base;
  if (warp == 1) {
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc_bf16(128, n, 0, mode == 2 ? 1 : 0);
    const uint32_t abase = smem_u32(smem);              // 64 KB A region (SS)
    const uint32_t bbase = smem_u32(smem + 65536);      // B tiles
    const uint32_t a_lo = desc_lo_sw128(abase, 16), b_lo = desc_lo_sw128(bbase, mode == 2 ? 8192 : 16);
    const uint32_t slab = (uint32_t)(n * 128) >> 4;
    long long t0 = clock64();
    for (int i = 0; i < count; i += 16) {
      if (leader) {
#pragma unroll
        for (int ks = 0; ks < 16; ++ks) {
          if (mode == 0) {
            umma_ts_lohi(tmem + 0, tmem + 256 + ks * 8, b_lo + (ks >> 2) * slab + (ks & 3) * 2, kDescHiSw128_1024, idesc, 1);
          } else if (mode == 1) {
            umma_ss_lohi(tmem + 0, a_lo + (ks >> 2) * 1024 + (ks & 3) * 2, b_lo + (ks >> 2) * slab + (ks & 3) * 2,
                         kDescHiSw128_1024, idesc, 1);
          } else {
            umma_ts_lohi(tmem + 0, tmem + 256 + (ks & 3) * 8, b_lo + (ks & 3) * 128, kDescHiSw128_1024, idesc, 1);
          }
        }
      }
      __syncwarp();
    }
    if (leader) umma_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (leader) out[blockIdx.x].cycles = t1 - t0;
  } else if (warp >= 2 && tma_like_writes) {
    // background shared-memory write traffic (st.shared 16 B per thread per iteration) while the MMAs run
    uint4* dst = reinterpret_cast<uint4*>(smem + 131072) + (threadIdx.x - 64);
    for (int i = 0; i < count * 4; ++i) { dst[(i & 31) * 64] = make_uint4(i, i, i, i); }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
  Result* d;
  cudaMalloc(&d, 148 * sizeof(Result));
  const size_t smem = 200 * 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int count = 4096;
  const char* names[3] = {"TS  K-major B", "SS  K-major B", "TS MN-major B"};
  for (int grid : {1, 148}) {
    for (int mode = 0; mode < 3; ++mode) {
      for (int n : {64, 128, 256}) {
        if (mode == 2 && n != 256) continue;
        for (int w = 0; w < 2; ++w) {
          probe_kernel<<<grid, 128, smem>>>(mode, n, count, w, d);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          Result h[148];
          cudaMemcpy(h, d, grid * sizeof(Result), cudaMemcpyDeviceToHost);
          long long mx = 0;
          for (int i = 0; i < grid; ++i) mx = h[i].cycles > mx ? h[i].cycles : mx;
          const double per = (double)mx / count;
          const double ideal = 128.0 * n / 256.0;
          printf("grid %3d  %s  M=128 N=%3d K=16  smem-writes=%d : %7.1f cycles/MMA (ideal %5.1f) -> %5.1f %% of tensor peak\n", grid,
                 names[mode], n, w, per, ideal, 100.0 * ideal / per);
        }
      }
    }
  }
  return 0;
}
