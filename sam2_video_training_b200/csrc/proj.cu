// Input projections of RoPEAttention with bias and axial RoPE fused into the GEMM epilogue (transformer.py:277-279 +
// :296-302, position_encoding.py:212-239):
//
//     Y[R, Nout] = X[R, K] . W[Nout, K]^T + bias,   leading `rope_chunks` x 128 output columns rotated per row
//
// One launch replaces addmm (x3 for the stacked q/k/v of the self-attention) + the RoPE pass: the rotation happens in
// registers on the fp32 accumulator, before the single rounding to bf16 -- no un-rotated q / k ever reaches HBM.
// K = 256 (queries / self-attention, X = LayerNorm output) or 64 (memory-bank keys and values, X = bf16(memory + pos)).
// Nout is a multiple of 128 and is split over up to three contiguous [R, 256] outputs (q | k | v).
//
// Blackwell mapping (the GEMM skeleton of csrc/mlp.cu): one CTA per 128 rows; X tile by TMA -> TMEM as the A operand;
// W streamed in [128 x K] K-major chunks (TMA, 2-stage ring); 128 x 128 x K tcgen05 GEMMs into double-buffered TMEM
// accumulators; epilogue warps add the bias, rotate, stage their 32 x 64 block in the 128-byte-swizzle box layout and
// issue their own TMA stores.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "abi_common.cuh"
#include "attn_kernels.cuh"
#include "tma_desc.cuh"

namespace proj {

using namespace sm100;
using attn::kBoxBytes;
using attn::kSlabBytes;

constexpr int kBlockM = 128;
constexpr int kBlockN = 128;
constexpr int kMaxK = 256;
constexpr int kWTileBytes = kBlockN * kMaxK * 2;     // 64 KB: up to four [128 rows x 128 B] slabs
constexpr int kStageBytes = kBlockM * kBlockN * 2;   // 32 KB: output staging, two [128 rows x 128 B] slabs
constexpr int kThreads = 320;
constexpr int kEpiWarps = 8;
constexpr uint32_t kColA = 0, kColAcc0 = 128, kColAcc1 = 256;

struct Shared {
  alignas(1024) uint8_t w_tiles[2][kWTileBytes];     // stage 1 also stages the X tile at start
  alignas(1024) uint8_t stage[2][kStageBytes];
  alignas(8) uint64_t w_full[2];
  uint64_t w_empty[2];
  uint64_t acc_full[2];
  uint64_t acc_free[2];
  uint64_t a_full;
  uint64_t a_ready;
  float bias[768];                                   // fp32 copy of the bias (Nout <= 768)
  uint32_t tmem_base;
};

struct Params {
  int K;                        // 256 or 64
  int n_chunks;                 // Nout / 128
  int rows;                     // R
  const __nv_bfloat16* bias;    // [Nout] or nullptr
  int rope_chunks;              // leading chunks (of 128 output columns) that are rotated; 0 = none
  const float2* table;          // [period, 128] (cos, sin)
  int rows_per_item;            // L: row r belongs to position r % L of its batch item
  int n_rope_rows;              // positions [0, n_rope_rows) are rotated (object-pointer keys are not)
  int period;                   // table row = position % period
};

__global__ void __launch_bounds__(kThreads, 1)
proj_kernel(const __grid_constant__ CUtensorMap map_x,      // X [R, K] bf16, box 64 x 128
            const __grid_constant__ CUtensorMap map_w,      // W [Nout, K] bf16, box 64 x 128
            const __grid_constant__ CUtensorMap map_o0,     // outputs [R, 256] bf16, box 64 x 32 (store)
            const __grid_constant__ CUtensorMap map_o1, const __grid_constant__ CUtensorMap map_o2, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  Shared& sh = *reinterpret_cast<Shared*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row_tile = blockIdx.x;
  const int nc = p.n_chunks;
  const int kslabs = p.K >> 6;                       // 64-column slabs of the contraction dimension (4 or 1)

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sh.w_full[i], 1); mbar_init(&sh.w_empty[i], 1);
      mbar_init(&sh.acc_full[i], 1); mbar_init(&sh.acc_free[i], kEpiWarps * 32);
    }
    mbar_init(&sh.a_full, 1);
    mbar_init(&sh.a_ready, kEpiWarps * 32);
    fence_barrier_init();
  }
  if (warp == 8 && lane == 0) { prefetch_tmap(&map_x); prefetch_tmap(&map_w); }
  if (warp == 0 && lane == 0) { prefetch_tmap(&map_o0); prefetch_tmap(&map_o1); prefetch_tmap(&map_o2); }
  if (warp == 9) { tmem_alloc(&sh.tmem_base, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sh.tmem_base;

  if (warp == 8) {
    // ===================== TMA producer =====================
    const bool leader = elect_one();
    if (leader) {
      mbar_arrive_expect_tx(&sh.a_full, kslabs * kSlabBytes);
      for (int c = 0; c < kslabs; ++c) tma_load_3d(&sh.w_tiles[1][c * kSlabBytes], &map_x, &sh.a_full, c * 64, row_tile * kBlockM, 0);
    }
    __syncwarp();
    for (int j = 0; j < nc; ++j) {
      const int s = j & 1;
      if (j == 1) mbar_wait(&sh.a_ready, 0);          // first use of the stage that staged the X tile
      mbar_wait(&sh.w_empty[s], ((j >> 1) & 1) ^ 1);
      if (leader) {
        mbar_arrive_expect_tx(&sh.w_full[s], kslabs * kSlabBytes);
        for (int c = 0; c < kslabs; ++c) tma_load_3d(&sh.w_tiles[s][c * kSlabBytes], &map_w, &sh.w_full[s], c * 64, j * kBlockN, 0);
      }
      __syncwarp();
    }
  } else if (warp == 9) {
    // ===================== MMA issuer =====================
    const bool leader = elect_one();
    constexpr uint32_t idesc = make_idesc_bf16(kBlockM, kBlockN, 0, 0);       // A (TMEM, K-major) . W^T, W K-major
    const uint32_t w_lo0 = desc_lo_sw128(smem_u32(&sh.w_tiles[0][0]), 16);    // K-major: LBO unused
    const int ksteps = p.K >> 4;
    mbar_wait(&sh.a_ready, 0);
    tc_fence_after();
    for (int j = 0; j < nc; ++j) {
      const int s = j & 1;
      const uint32_t ph = (j >> 1) & 1;
      mbar_wait(&sh.w_full[s], ph);
      mbar_wait(&sh.acc_free[s], ph ^ 1);
      tc_fence_after();
      if (leader) {
        const uint32_t wlo = w_lo0 + s * (kWTileBytes >> 4);
        const uint32_t d = tmem + (s ? kColAcc1 : kColAcc0);
        for (int ks = 0; ks < ksteps; ++ks)       // K-major SW128: slab ks / 4 (16 KB), 32 B per k-step inside the 128 B row
          umma_ts_lohi(d, tmem + kColA + ks * 8, wlo + (ks >> 2) * (kSlabBytes >> 4) + (ks & 3) * 2, kDescHiSw128_1024, idesc,
                       ks > 0);
        umma_commit(&sh.w_empty[s]);
        umma_commit(&sh.acc_full[s]);
      }
      __syncwarp();
    }
  } else {
    // ===================== epilogue warps (0..7) =====================
    const int quarter = warp & 3;
    const int half = warp >> 2;                   // which 64 of the chunk's 128 columns
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = tmem + (uint32_t(quarter * 32) << 16);
    mbar_wait(&sh.a_full, 0);
    {   // X tile: shared -> registers -> TMEM (bf16 pairs); half h moves slabs 2h, 2h+1 (K = 256) or slab 0 (K = 64, h = 0)
      const int c_begin = (kslabs == 4) ? half * 2 : 0, c_end = (kslabs == 4) ? half * 2 + 2 : (half == 0 ? 1 : 0);
      for (int c = c_begin; c < c_end; ++c) {
        const uint32_t base = smem_u32(&sh.w_tiles[1][0]) + c * kSlabBytes + row * 128;
        uint32_t r[32];
#pragma unroll
        for (int v = 0; v < 8; ++v) {
          const uint4 u = lds128(base + ((v ^ (row & 7)) << 4));
          r[4 * v + 0] = u.x; r[4 * v + 1] = u.y; r[4 * v + 2] = u.z; r[4 * v + 3] = u.w;
        }
        SAM2B200_TMEM_ST32(lane_addr + kColA + c * 32, r);
      }
      tmem_wait_st();
    }
    tc_fence_before();
    mbar_arrive(&sh.a_ready);

    const int row0 = row_tile * kBlockM + quarter * 32;
    const long long grow = (long long)row0 + lane;                     // global row of this thread
    const int pos = (int)(grow % p.rows_per_item);                     // position inside the batch item
    const bool rot_row = p.rope_chunks > 0 && pos < p.n_rope_rows;
    const float2* trow = p.table + (long long)(pos % p.period) * 128;
    // bias -> shared memory once (fp32), read back with broadcast loads in the epilogue
    for (int i = threadIdx.x; i < nc * kBlockN; i += kEpiWarps * 32) sh.bias[i] = p.bias ? __bfloat162float(p.bias[i]) : 0.f;
    asm volatile("bar.sync 5, 256;" ::: "memory");
    for (int j = 0; j < nc; ++j) {
      const int s = j & 1;
      const int col0 = j * kBlockN + half * 64;                        // first output column of this thread's 64
      // the (cos, sin) row is L2-resident but ~1 us away: fetch it before waiting for the accumulator
      const bool rotate = rot_row && j < p.rope_chunks;
      float2 cs[32];
      if (rotate) {
        const float4* src = reinterpret_cast<const float4*>(trow + ((col0 & 255) >> 1));
#pragma unroll
        for (int i = 0; i < 16; ++i) { const float4 f = __ldg(src + i); cs[2 * i] = make_float2(f.x, f.y); cs[2 * i + 1] = make_float2(f.z, f.w); }
      }
      mbar_wait(&sh.acc_full[s], (j >> 1) & 1);
      tc_fence_after();
      // this warp's staging region is private: only its own previous store into this buffer (2 chunks ago) must have been read
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      __syncwarp();
      const uint32_t srow = smem_u32(&sh.stage[s][0]) + half * kSlabBytes + row * 128;
#pragma unroll
      for (int sb = 0; sb < 2; ++sb) {            // two sub-blocks of 32 columns
        uint32_t acc[32];
        SAM2B200_TMEM_LD32(lane_addr + (s ? kColAcc1 : kColAcc0) + half * 64 + sb * 32, acc);
        tmem_wait_ld();
        if (sb == 1) { tc_fence_before(); mbar_arrive(&sh.acc_free[s]); }
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(acc[i]) + sh.bias[col0 + sb * 32 + i];
        if (rotate) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {          // (re, im) = (v[2i], v[2i+1]) times (cos + i sin)
            const float2 c2 = cs[sb * 16 + i];
            const float re = v[2 * i] * c2.x - v[2 * i + 1] * c2.y;
            const float im = v[2 * i] * c2.y + v[2 * i + 1] * c2.x;
            v[2 * i] = re; v[2 * i + 1] = im;
          }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
          sts128(srow + (((sb * 4 + q) ^ (row & 7)) << 4), pack_bf16(v[8 * q + 0], v[8 * q + 1]), pack_bf16(v[8 * q + 2], v[8 * q + 3]),
                 pack_bf16(v[8 * q + 4], v[8 * q + 5]), pack_bf16(v[8 * q + 6], v[8 * q + 7]));
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        const int which = j >> 1;                                       // output tensor: 256 columns = 2 chunks each
        const CUtensorMap* mo = which == 0 ? &map_o0 : (which == 1 ? &map_o1 : &map_o2);
        if (row0 < p.rows)
          tma_store_3d(mo, &sh.stage[s][half * kSlabBytes + quarter * kBoxBytes], (j & 1) * kBlockN + half * 64, row0, 0);
        tma_store_commit();
      }
    }
    if (lane == 0) tma_store_wait_read();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem, 512);
}

}  // namespace proj

namespace {

int make_bf16_matrix_map(CUtensorMap* map, const void* base, long long rows, long long cols, int box_rows) {
  sam2b200::PFN_encodeTiled enc = sam2b200::get_encode_tiled();
  if (!enc) return sam2b200::fail(SAM2B200_ERR_DRIVER, "cuTensorMapEncodeTiled entry point not found");
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, 1};
  cuuint64_t strides[2] = {(cuuint64_t)cols * 2, (cuuint64_t)rows * (cuuint64_t)cols * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(sam2b200::last_error_buffer(), 512, "cuTensorMapEncodeTiled (proj) failed (%d) rows=%lld cols=%lld", (int)r, rows, cols);
    return SAM2B200_ERR_DRIVER;
  }
  return SAM2B200_OK;
}

}  // namespace

extern "C" {

// x: [R, K] bf16 (K = 256 or 64); w: [Nout, K] bf16; bias: [Nout] bf16 or NULL; out0 / out1 / out2: [R, 256] bf16
// contiguous, receiving output columns [0,256) / [256,512) / [512,768) (Nout = 256 * number of outputs).  The first
// rope_cols output columns (a multiple of 256) are rotated with table [period, 128] (cos, sin) for rows whose position
// (row % rows_per_item) is < n_rope_rows; table row = position % period.
int sam2b200_proj_rope(const void* x, const void* w, const void* bias, void* out0, void* out1, void* out2, long long R, int K,
                       int Nout, int rope_cols, const float* table, int rows_per_item, int n_rope_rows, int period,
                       cudaStream_t stream) {
  const int n_out = Nout / 256;
  if (!x || !w || !out0 || R <= 0 || R > 0x7fffffffLL - 256 || (K != 256 && K != 64) || Nout <= 0 || (Nout % 256) || n_out > 3 ||
      (n_out > 1 && !out1) || (n_out > 2 && !out2) || rope_cols < 0 || (rope_cols % 256) || rope_cols > Nout ||
      (rope_cols > 0 && (!table || rows_per_item <= 0 || n_rope_rows < 0 || period <= 0)) ||
      ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(out0) |
        reinterpret_cast<uintptr_t>(out1) | reinterpret_cast<uintptr_t>(out2) | reinterpret_cast<uintptr_t>(bias)) & 15))
    return sam2b200::fail(SAM2B200_ERR_INVALID, "proj_rope: bad arguments");
  CUtensorMap map_x, map_w, map_o[3];
  int rc;
  if ((rc = make_bf16_matrix_map(&map_x, x, R, K, 128))) return rc;
  if ((rc = make_bf16_matrix_map(&map_w, w, Nout, K, 128))) return rc;
  void* outs[3] = {out0, out1 ? out1 : out0, out2 ? out2 : out0};
  for (int i = 0; i < 3; ++i)
    if ((rc = make_bf16_matrix_map(&map_o[i], outs[i], R, 256, 32))) return rc;
  const size_t smem = sizeof(proj::Shared) + 1024;
  cudaError_t e = cudaFuncSetAttribute(proj::proj_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return sam2b200::fail(SAM2B200_ERR_CUDA, cudaGetErrorString(e));
  proj::Params p{};
  p.K = K; p.n_chunks = Nout / proj::kBlockN; p.rows = (int)R; p.bias = static_cast<const __nv_bfloat16*>(bias);
  p.rope_chunks = rope_cols / proj::kBlockN; p.table = reinterpret_cast<const float2*>(table);
  p.rows_per_item = rows_per_item > 0 ? rows_per_item : 1; p.n_rope_rows = n_rope_rows; p.period = period > 0 ? period : 1;
  const unsigned grid = (unsigned)((R + proj::kBlockM - 1) / proj::kBlockM);
  proj::proj_kernel<<<grid, proj::kThreads, smem, stream>>>(map_x, map_w, map_o[0], map_o[1], map_o[2], p);
  return sam2b200::check_launch("proj_rope");
}

}  // extern "C"
