// MLP backward of MemoryAttentionLayer, first half (memory_attention.py:95-98: tgt2 = linear2(dropout(relu(linear1(x))))):
//
//     dh[R, F] = (dm[R, 256] . W2[256, F]) o (h[R, F] > 0) * scale           (F = dim_feedforward = 2048)
//
// i.e. the input gradient of linear2 with the ReLU (and hidden-dropout) backward fused into the GEMM epilogue.  The
// cuBLAS GEMM + separate mask pass it replaces moves 545 MB per layer and frame at the cfg2 shape (R = 32 256):
// GEMM writes dh (132 MB), the mask pass reads dh + h and writes dh again; this kernel reads dm + h and writes dh once
// (281 MB) -- HBM-bound, ~2.8x less time on the critical path of the backward (cuBLASLt's DRELU epilogue was measured
// at 473 us, profiles/r1_cublaslt_epilogue_probe.txt).
//
// Blackwell mapping (same building blocks as attn_kernels.cuh): one CTA per 128 rows; the 128 x 256 dm tile arrives by
// TMA and lives in TENSOR MEMORY as the bf16 A operand; W2 is streamed in [256 x 128] chunks (MN-major B operand, TMA,
// 2-stage ring); each chunk is one 128 x 128 x 256 tcgen05 GEMM into a double-buffered TMEM accumulator; the matching
// 128 x 128 tile of h arrives by TMA in the 128-byte-swizzle box layout, every epilogue warp masks its 32 x 64 block IN
// PLACE in that shared-memory tile and hands it to the TMA unit as the output box (no separate staging buffer).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "abi_common.cuh"
#include "attn_kernels.cuh"
#include "tma_desc.cuh"

namespace mlp {

using namespace sm100;
using attn::kBoxBytes;
using attn::kSlabBytes;

constexpr int kBlockM = 128;                     // rows per CTA (TMEM lanes)
constexpr int kK = 256;                          // d_model: contraction length
constexpr int kBlockN = 128;                     // columns of dh per chunk
constexpr int kBTileBytes = kK * kBlockN * 2;    // 64 KB: two [256 rows x 128 B] slabs
constexpr int kBSlabBytes = kK * 128;            // 32 KB
constexpr int kHTileBytes = kBlockM * kBlockN * 2;   // 32 KB: two [128 rows x 128 B] slabs
constexpr int kThreads = 320;                    // warps 0-7 epilogue, 8 TMA producer, 9 MMA issuer
constexpr int kEpiWarps = 8;
constexpr uint32_t kColA = 0;                    // 128 TMEM columns: dm tile (bf16 pairs)
constexpr uint32_t kColAcc0 = 128, kColAcc1 = 256;   // 2 x 128 fp32 accumulator columns

struct Shared {
  alignas(1024) uint8_t b_tiles[2][kBTileBytes];     // W2 chunks (stage 1 also stages the dm tile at start)
  alignas(1024) uint8_t h_tiles[2][kHTileBytes];     // h tile in, masked dh tile out (in place)
  alignas(8) uint64_t b_full[2];
  uint64_t b_empty[2];
  uint64_t acc_full[2];
  uint64_t acc_free[2];
  uint64_t h_full[2];
  uint64_t h_free[2];
  uint64_t a_full;
  uint64_t a_ready;
  uint32_t tmem_base;
};

struct Params {
  int n_chunks;       // F / 128
  int rows;           // R
  float scale;        // 1 / (1 - p) of the hidden dropout (1 if none)
  float* dbias;       // [F] fp32 or nullptr: column sums of dh (the bias gradient of linear1) are ADDED here (fp32 atomics)
};

__global__ void __launch_bounds__(kThreads, 1)
dh_kernel(const __grid_constant__ CUtensorMap map_dm,     // dm  [R, 256]  bf16, box 64 x 128
          const __grid_constant__ CUtensorMap map_w2,     // W2  [256, F]  bf16, box 64 x 256
          const __grid_constant__ CUtensorMap map_h,      // h   [R, F]    bf16, box 64 x 128 (load)
          const __grid_constant__ CUtensorMap map_dh,     // dh  [R, F]    bf16, box 64 x 32  (store)
          const Params p) {
  extern __shared__ uint8_t smem_raw[];
  Shared& sh = *reinterpret_cast<Shared*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row_tile = blockIdx.x;
  const int nc = p.n_chunks;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sh.b_full[i], 1); mbar_init(&sh.b_empty[i], 1);
      mbar_init(&sh.acc_full[i], 1); mbar_init(&sh.acc_free[i], kEpiWarps * 32);
      mbar_init(&sh.h_full[i], 1); mbar_init(&sh.h_free[i], kEpiWarps);
    }
    mbar_init(&sh.a_full, 1);
    mbar_init(&sh.a_ready, kEpiWarps * 32);
    fence_barrier_init();
  }
  if (warp == 8 && lane == 0) { prefetch_tmap(&map_dm); prefetch_tmap(&map_w2); prefetch_tmap(&map_h); }
  if (warp == 0 && lane == 0) prefetch_tmap(&map_dh);
  if (warp == 9) { tmem_alloc(&sh.tmem_base, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_launch_dependents();      // programmatic dependent launch: nothing above touched global memory
  pdl_wait();
  const uint32_t tmem = sh.tmem_base;

  if (warp == 8) {
    // ===================== TMA producer =====================
    const bool leader = elect_one();
    if (leader) {   // dm tile -> ring stage 1 (4 slabs of [128 rows x 128 B] = 64 KB), moved to TMEM by the epilogue warps
      mbar_arrive_expect_tx(&sh.a_full, 4 * kSlabBytes);
#pragma unroll
      for (int c = 0; c < 4; ++c) tma_load_3d(&sh.b_tiles[1][c * kSlabBytes], &map_dm, &sh.a_full, c * 64, row_tile * kBlockM, 0);
    }
    __syncwarp();
    for (int j = 0; j < nc; ++j) {
      const int s = j & 1;
      const uint32_t ph = (j >> 1) & 1;
      if (j == 1) mbar_wait(&sh.a_ready, 0);          // first use of the stage that staged the dm tile
      mbar_wait(&sh.b_empty[s], ph ^ 1);
      if (leader) {
        mbar_arrive_expect_tx(&sh.b_full[s], kBTileBytes);
#pragma unroll
        for (int c = 0; c < 2; ++c) tma_load_3d(&sh.b_tiles[s][c * kBSlabBytes], &map_w2, &sh.b_full[s], j * kBlockN + c * 64, 0, 0);
      }
      __syncwarp();
      mbar_wait(&sh.h_free[s], ph ^ 1);               // the previous output box of this buffer has been read by the TMA unit
      if (leader) {
        mbar_arrive_expect_tx(&sh.h_full[s], kHTileBytes);
#pragma unroll
        for (int c = 0; c < 2; ++c)
          tma_load_3d(&sh.h_tiles[s][c * kSlabBytes], &map_h, &sh.h_full[s], j * kBlockN + c * 64, row_tile * kBlockM, 0);
      }
      __syncwarp();
    }
  } else if (warp == 9) {
    // ===================== MMA issuer =====================
    const bool leader = elect_one();
    constexpr uint32_t idesc = make_idesc_bf16(kBlockM, kBlockN, 0, 1);          // A (TMEM, K-major) . B (SMEM, MN-major)
    const uint32_t b_lo0 = desc_lo_sw128(smem_u32(&sh.b_tiles[0][0]), kBSlabBytes);   // MN-major: LBO = slab stride
    mbar_wait(&sh.a_ready, 0);
    tc_fence_after();
    for (int j = 0; j < nc; ++j) {
      const int s = j & 1;
      const uint32_t ph = (j >> 1) & 1;
      mbar_wait(&sh.b_full[s], ph);
      mbar_wait(&sh.acc_free[s], ph ^ 1);
      tc_fence_after();
      if (leader) {
        const uint32_t blo = b_lo0 + s * (kBTileBytes >> 4);
        const uint32_t d = tmem + (s ? kColAcc1 : kColAcc0);
#pragma unroll
        for (int ks = 0; ks < kK / 16; ++ks)      // MN-major SW128: 16 k-rows = 2 groups of 8 rows = 2048 B per k-step
          umma_ts_lohi(d, tmem + kColA + ks * 8, blo + ks * (2048 >> 4), kDescHiSw128_1024, idesc, ks > 0);
        umma_commit(&sh.b_empty[s]);
        umma_commit(&sh.acc_full[s]);
      }
      __syncwarp();
    }
  } else {
    // ===================== epilogue warps (0..7) =====================
    const int quarter = warp & 3;                 // TMEM lane quarter
    const int half = warp >> 2;                   // which 64 of the chunk's 128 columns = which slab of the h tile
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = tmem + (uint32_t(quarter * 32) << 16);
    mbar_wait(&sh.a_full, 0);
    attn::stage_to_tmem_half(smem_u32(&sh.b_tiles[1][0]) + half * 2 * kSlabBytes, row, lane_addr + kColA, half);
    tc_fence_before();
    mbar_arrive(&sh.a_ready);
    const int row0 = row_tile * kBlockM + quarter * 32;
    const float scale = p.scale;
    for (int j = 0; j < nc; ++j) {
      const int s = j & 1;
      const uint32_t ph = (j >> 1) & 1;
      mbar_wait(&sh.acc_full[s], ph);
      tc_fence_after();
      uint32_t acc[64];
      SAM2B200_TMEM_LD32(lane_addr + (s ? kColAcc1 : kColAcc0) + half * 64, acc);
      SAM2B200_TMEM_LD32(lane_addr + (s ? kColAcc1 : kColAcc0) + half * 64 + 32, (acc + 32));
      tmem_wait_ld();
      tc_fence_before();
      mbar_arrive(&sh.acc_free[s]);               // the accumulator may be overwritten by chunk j + 2
      mbar_wait(&sh.h_full[s], ph);
      // this thread's 128 bytes: row `row` of slab `half` (64 bf16), 16-byte chunks XOR-swizzled with (row & 7)
      const uint32_t hrow = smem_u32(&sh.h_tiles[s][0]) + half * kSlabBytes + row * 128;
      const bool row_in = (row0 + lane) < p.rows;     // rows beyond the tensor are clipped by the store, not by the column sums
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const uint32_t addr = hrow + ((q ^ (row & 7)) << 4);
        const uint4 hv = lds128(addr);
        const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w};
        uint32_t o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 hf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&hw[e]));
          const float v0 = hf.x > 0.f ? __uint_as_float(acc[q * 8 + 2 * e]) * scale : 0.f;
          const float v1 = hf.y > 0.f ? __uint_as_float(acc[q * 8 + 2 * e + 1]) * scale : 0.f;
          o[e] = pack_bf16(v0, v1);
          // keep the masked values (as the bf16 numbers the weight-gradient GEMM will see) for the bias gradient
          const float2 r2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&o[e]));
          acc[q * 8 + 2 * e] = __float_as_uint(row_in ? r2.x : 0.f);
          acc[q * 8 + 2 * e + 1] = __float_as_uint(row_in ? r2.y : 0.f);
        }
        sts128(addr, o[0], o[1], o[2], o[3]);
      }
      if (p.dbias != nullptr) {   // bias gradient of linear1: column sums of this warp's 32 x 64 block (31-shuffle butterfly per 32 columns)
        const float* mv = reinterpret_cast<const float*>(acc);
        const float c0s = attn::warp_column_sum32(mv, lane);
        const float c1s = attn::warp_column_sum32(mv + 32, lane);
        atomicAdd(p.dbias + j * kBlockN + half * 64 + lane, c0s);
        atomicAdd(p.dbias + j * kBlockN + half * 64 + 32 + lane, c1s);
      }
      // the warp's box: 32 rows x 64 columns = rows quarter*32 .. +31 of slab `half`
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        if (row0 < p.rows)                        // rows beyond the tensor inside the box are clipped by the TMA unit
          tma_store_3d(&map_dh, &sh.h_tiles[s][half * kSlabBytes + quarter * kBoxBytes], j * kBlockN + half * 64, row0, 0);
        tma_store_commit();
        tma_store_wait_read();                    // the TMA unit has read the box (4 KB): the buffer can take the h tile of chunk j + 2
        mbar_arrive(&sh.h_free[s]);
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem, 512);
}

}  // namespace mlp

namespace {

// row-major bf16 matrix [rows, cols] with leading dimension ld (elements) as a 3-D tensor map (cols, rows, 1) with a
// (64 columns x box_rows) box, 128-byte swizzle
int make_matrix_map(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld, int box_rows) {
  sam2b200::PFN_encodeTiled enc = sam2b200::get_encode_tiled();
  if (!enc) return sam2b200::fail(SAM2B200_ERR_DRIVER, "cuTensorMapEncodeTiled entry point not found");
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, 1};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)rows * (cuuint64_t)ld * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(sam2b200::last_error_buffer(), 512, "cuTensorMapEncodeTiled (matrix) failed (%d) rows=%lld cols=%lld", (int)r, rows, cols);
    return SAM2B200_ERR_DRIVER;
  }
  return SAM2B200_OK;
}

}  // namespace

extern "C" {

// dh[R, F] = (dm[R, 256] . W2[256, F]) o (h[R, F] > 0) * scale; all bf16 row-major contiguous, F a multiple of 128.
// dbias (optional): [F] fp32, += column sums of dh -- the bias gradient of linear1 from the tile in registers (no extra pass).
int sam2b200_mlp_dh(const void* dm, const void* w2, const void* h, void* dh, float* dbias, long long R, int F, float scale,
                    cudaStream_t stream) {
  if (!dm || !w2 || !h || !dh || R <= 0 || F <= 0 || (F % mlp::kBlockN) || R > 0x7fffffffLL - 256 ||
      ((reinterpret_cast<uintptr_t>(dm) | reinterpret_cast<uintptr_t>(w2) | reinterpret_cast<uintptr_t>(h) |
        reinterpret_cast<uintptr_t>(dh)) & 15))
    return sam2b200::fail(SAM2B200_ERR_INVALID, "mlp_dh: bad arguments (F must be a multiple of 128, 16-byte aligned tensors)");
  CUtensorMap map_dm, map_w2, map_h, map_dh;
  int rc;
  if ((rc = make_matrix_map(&map_dm, dm, R, 256, 256, 128))) return rc;
  if ((rc = make_matrix_map(&map_w2, w2, 256, F, F, 256))) return rc;
  if ((rc = make_matrix_map(&map_h, h, R, F, F, 128))) return rc;
  if ((rc = make_matrix_map(&map_dh, dh, R, F, F, 32))) return rc;
  const size_t smem = sizeof(mlp::Shared) + 1024;
  cudaError_t e = cudaFuncSetAttribute(mlp::dh_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return sam2b200::fail(SAM2B200_ERR_CUDA, cudaGetErrorString(e));
  mlp::Params p{F / mlp::kBlockN, (int)R, scale, dbias};
  const unsigned grid = (unsigned)((R + mlp::kBlockM - 1) / mlp::kBlockM);
  if (sam2b200::launch_pdl(mlp::dh_kernel, dim3(grid), dim3(mlp::kThreads), smem, stream, map_dm, map_w2, map_h, map_dh, p) != cudaSuccess)
    return sam2b200::fail(SAM2B200_ERR_CUDA, "mlp_dh: launch failed");
  return sam2b200::check_launch("mlp_dh");
}

}  // extern "C"
