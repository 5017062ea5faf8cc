// Dense GEMMs around the attention core of MemoryAttentionLayer that have no LayerNorm in front of them (those are
// csrc/lnproj.cu) and no ReLU mask behind them (csrc/mlp.cu):
//
//     C[R, Nout] (bf16) = epi( A[R, K] (bf16) . B  + bias )          Nout = 64 | 256 x (1..8),  K a multiple of 64
//
// and, behind the stand-alone LayerNorm pass (glue.cu ln_fwd), the projections of the pre-norm block heads with their epilogues:
// q|k|v / q with bias + axial RoPE on the fp32 accumulator (transformer.py:277-302), linear1 with bias + ReLU (+ hidden dropout)
// (memory_attention.py:95-97) -- C split over up to three [R, out_width] outputs.
//
//   B layout 0 ("NT"): B = W[BN, K] row-major, C = A W^T  -- forward projections: linear2 (memory_attention.py:97-98, K = 2048) and
//                      the memory-key projection k_proj (transformer.py:278, K = 64) with the axial rotation of
//                      position_encoding.py:212-239 applied on the fp32 accumulator (rows >= n_rope_rows of an item -- the object
//                      pointers, transformer.py:296-302 -- are not rotated);
//   B layout 1 ("NN"): B = W[K, BN] row-major, C = A W    -- input gradients dX = dY W of linear1 (K = 2048), of the stacked
//                      q|k|v projection (K = 768), of out_proj / q_proj (K = 256) and of the folded Wo Wv (BN = 64).
//
// Blackwell mapping: RESIDENT CTAs (one per SM) walk the 128-row tiles of C; A and B k-slices of 64 arrive by TMA in a 4-stage
// ring shared by consecutive tiles (the producer runs ahead into the next tile while the current one drains); both operands are
// read from shared memory (SS-mode tcgen05.mma, M = 128, N = BN, K = 16 per instruction), the B tile K-major (NT) or MN-major
// (NN) straight from the TMA boxes -- no transposed weight copy exists anywhere; two BN-column fp32 accumulators in tensor memory
// alternate between tiles, so the epilogue of tile i (tcgen05.ld -> bias / rotation -> bf16 -> swizzled staging box -> TMA store
// per warp) overlaps the MMAs of tile i + 1.  HBM-bound at every shape of the stack: algorithmic bytes = 2 (R K + K BN + R BN).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "abi_common.cuh"
#include "dropout.cuh"
#include "sm100.cuh"
#include "tma_desc.cuh"

namespace gemm {

using namespace sm100;

constexpr int kBlockM = 128;                 // rows per tile (TMEM lanes)
constexpr int kBlockK = 64;                  // contraction slice per ring stage (one 128-byte swizzle row)
constexpr int kStages = 4;
constexpr int kThreads = 320;                // warps 0-7 epilogue, 8 TMA producer, 9 MMA issuer
constexpr int kEpiWarps = 8;
constexpr int kATileBytes = kBlockM * 128;   // 16 KB: [128 rows x 128 B]
constexpr int kBoxBytes = 32 * 128;          // one warp's output box: 32 rows x 64 bf16 columns

template <int BN>
struct Shared {
  alignas(1024) uint8_t a_tiles[kStages][kATileBytes];
  alignas(1024) uint8_t b_tiles[kStages][BN * 128];          // NT: [BN rows x 128 B]; NN: BN / 64 slabs of [64 k-rows x 128 B]
  alignas(1024) uint8_t stage[kEpiWarps][kBoxBytes];         // per-warp output staging (128-byte swizzle box layout)
  alignas(8) uint64_t full[kStages];
  uint64_t empty[kStages];
  uint64_t acc_full[2];
  uint64_t acc_free[2];
  float bias[BN];                                            // the current column block's bias slice
  uint32_t tmem_base;
};

struct Params {
  int rows;                 // R
  int n_tiles;              // ceil(R / 128) * n_col_blocks; tile t = (row tile t / n_col_blocks, column block t % n_col_blocks)
  int n_col_blocks;         // Nout / BN
  int blocks_per_out;       // output tensor of column block cb = cb / blocks_per_out, its first column (cb % blocks_per_out) * BN
  int ksteps;               // K / 64
  int nout;                 // Nout (dropout element index = row * Nout + column)
  int rope_blocks;          // leading column blocks (of 256) that are rotated
  int rope_w;               // sqrt(period) if the table is axial over a square grid (row = x part | y part), else 0
  int relu;                 // max(., 0) on the biased accumulator, then drop_out
  sam2b200::Dropout drop_out;
  const float* bias;        // [Nout] fp32 or nullptr
  const float2* table;      // RoPE: [period, 128] (cos, sin) or nullptr
  int rows_per_item;        // row r is position r % rows_per_item of its batch item
  int n_rope_rows;          // positions [0, n_rope_rows) are rotated
  int period;               // table row = position % period
  const float* dot_rows;    // BN = 64 only: [R, 64] fp32 or nullptr -> dot_out[r] = sum_c bf16(C[r, c]) * dot_rows[r, c]
  float* dot_out;           // [R] fp32 (the attention backward's Delta = rowsum(dO' o out64) from the GEMM that produces dO')
};

template <int BN, int B_MN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap map_a,      // A [R, K]: box 64 x 128
            const __grid_constant__ CUtensorMap map_b,      // NT: W [BN, K], box 64 x BN;  NN: W [K, BN], box 64 x 64
            const __grid_constant__ CUtensorMap map_c0,     // outputs [R, out_width]: box 64 x 32 (store)
            const __grid_constant__ CUtensorMap map_c1, const __grid_constant__ CUtensorMap map_c2, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  Shared<BN>& sh = *reinterpret_cast<Shared<BN>*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t kBBytes = BN * 128;
  constexpr uint32_t kTmemCols = 2 * BN < 32 ? 32 : 2 * BN;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(&sh.full[i], 1); mbar_init(&sh.empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&sh.acc_full[i], 1); mbar_init(&sh.acc_free[i], kEpiWarps * 32); }
    fence_barrier_init();
  }
  if (warp == 8 && lane == 0) { prefetch_tmap(&map_a); prefetch_tmap(&map_b); }
  if (warp == 0 && lane == 0) { prefetch_tmap(&map_c0); prefetch_tmap(&map_c1); prefetch_tmap(&map_c2); }
  if (warp == 9) { tmem_alloc(&sh.tmem_base, kTmemCols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sh.tmem_base;
  const int ksteps = p.ksteps;
  const int ncb = p.n_col_blocks;

  if (warp == 8) {
    // ===================== TMA producer: one (A, B) k-slice per ring slot, running across tile boundaries =====================
    const bool leader = elect_one();
    uint32_t slot = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      const int row_tile = tile / ncb, cb = tile - row_tile * ncb;
      for (int ks = 0; ks < ksteps; ++ks, ++slot) {
        const int s = slot % kStages;
        mbar_wait(&sh.empty[s], ((slot / kStages) & 1) ^ 1);
        if (leader) {
          mbar_arrive_expect_tx(&sh.full[s], kATileBytes + kBBytes);
          tma_load_3d(&sh.a_tiles[s][0], &map_a, &sh.full[s], ks * kBlockK, row_tile * kBlockM, 0);
          if (B_MN) {
#pragma unroll
            for (int c = 0; c < BN / 64; ++c)
              tma_load_3d(&sh.b_tiles[s][c * 8192], &map_b, &sh.full[s], cb * BN + c * 64, ks * kBlockK, 0);
          } else {
            tma_load_3d(&sh.b_tiles[s][0], &map_b, &sh.full[s], ks * kBlockK, cb * BN, 0);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 9) {
    // ===================== MMA issuer =====================
    const bool leader = elect_one();
    constexpr uint32_t idesc = make_idesc_bf16(kBlockM, BN, 0, B_MN);
    const uint32_t a_lo0 = desc_lo_sw128(smem_u32(&sh.a_tiles[0][0]), 16);                 // K-major: LBO unused
    const uint32_t b_lo0 = desc_lo_sw128(smem_u32(&sh.b_tiles[0][0]), B_MN ? 8192 : 16);   // MN-major: LBO = 64-column slab stride
    constexpr uint32_t b_kstep = B_MN ? (2048 >> 4) : 2;    // 16 k-rows x 128 B (MN-major) | 32 B inside the row (K-major)
    uint32_t slot = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
      const int ab = it & 1;
      mbar_wait(&sh.acc_free[ab], ((it >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d = tmem + ab * BN;
      for (int ks = 0; ks < ksteps; ++ks, ++slot) {
        const int s = slot % kStages;
        mbar_wait(&sh.full[s], (slot / kStages) & 1);
        tc_fence_after();
        if (leader) {
          const uint32_t alo = a_lo0 + s * (kATileBytes >> 4);
          const uint32_t blo = b_lo0 + s * (kBBytes >> 4);
#pragma unroll
          for (int k16 = 0; k16 < 4; ++k16)
            umma_ss_lohi(d, alo + k16 * 2, blo + k16 * b_kstep, kDescHiSw128_1024, idesc, (ks | k16) != 0);
          umma_commit(&sh.empty[s]);
          if (ks == ksteps - 1) umma_commit(&sh.acc_full[ab]);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue warps (0..7): warp (quarter, half) owns rows quarter*32.. of column chunks c = half, half + 2 =====================
    const int quarter = warp & 3;
    const int half = warp >> 2;
    const uint32_t lane_addr = tmem + (uint32_t(quarter * 32) << 16);
    const uint32_t srow = smem_u32(&sh.stage[warp][0]) + lane * 128;
    constexpr int kChunks = BN / 64;
    const bool drop_on = p.drop_out.seed != nullptr;
    const uint32_t dkey = drop_on ? sam2b200::dropout_key(*p.drop_out.seed, p.drop_out.site) : 0u;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
      const int ab = it & 1;
      const int row_tile = tile / ncb, cb = tile - row_tile * ncb;
      const int row0 = row_tile * kBlockM + quarter * 32;
      const int which = cb / p.blocks_per_out;
      const int ocol0 = (cb - which * p.blocks_per_out) * BN;         // first column of this block inside its output tensor
      const CUtensorMap* mo = which == 0 ? &map_c0 : (which == 1 ? &map_c1 : &map_c2);
      // this column block's bias slice -> shared memory (broadcast reads below); with one column block it is loaded once
      if (it == 0 || ncb > 1) {
        if (it > 0) asm volatile("bar.sync 5, 256;" ::: "memory");      // every warp has finished reading the previous slice
        if ((int)threadIdx.x < BN) sh.bias[threadIdx.x] = p.bias ? __ldg(p.bias + cb * BN + threadIdx.x) : 0.f;
        asm volatile("bar.sync 5, 256;" ::: "memory");
      }
      const float2* trow_x = nullptr;             // (cos, sin) rows of this thread's position: columns [0, 128) | [128, 256) of the head
      const float2* trow_y = nullptr;
      if (BN == 256 && cb < p.rope_blocks) {
        const int pos = (int)(((long long)row0 + lane) % p.rows_per_item);
        if (pos < p.n_rope_rows) {
          const int tpos = pos % p.period;
          // axial table over a square grid: the first 64 pairs depend on x = pos % w only, the last 64 on y = pos / w only --
          // a tile touches w + 128 / w distinct half rows instead of 128 (L1 resident)
          trow_x = p.table + (long long)(p.rope_w > 0 ? tpos % p.rope_w : tpos) * 128;
          trow_y = p.table + (long long)(p.rope_w > 0 ? tpos - tpos % p.rope_w : tpos) * 128;
        }
      }
      mbar_wait(&sh.acc_full[ab], (it >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int c = half; c < kChunks; c += 2) {
        // the warp's previous box must have been read by the TMA unit before its staging buffer is rewritten
        if (lane == 0) tma_store_wait_read();
        __syncwarp();
        float dot = 0.f;
        const float2* trow = (c < 2) ? trow_x : trow_y;
#pragma unroll
        for (int sb = 0; sb < 2; ++sb) {          // two sub-blocks of 32 columns
          uint32_t acc[32];
          SAM2B200_TMEM_LD32(lane_addr + ab * BN + c * 64 + sb * 32, acc);
          float4 cs[16];
          if (BN == 256 && trow != nullptr) {     // the (cos, sin) pairs of this row's 32 columns, fetched under the TMEM load
            const float4* src = reinterpret_cast<const float4*>(trow + c * 32 + sb * 16);
#pragma unroll
            for (int i = 0; i < 8; ++i) cs[i] = __ldg(src + i);
          }
          tmem_wait_ld();
          if (sb == 1 && c + 2 >= kChunks) { tc_fence_before(); mbar_arrive(&sh.acc_free[ab]); }   // this thread's last read of the accumulator
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(acc[i]) + sh.bias[c * 64 + sb * 32 + i];
          if (BN == 256 && trow != nullptr) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {         // (re, im) = (v[2j], v[2j+1]) times (cos + i sin), two pairs per float4
              const float r0 = v[4 * i] * cs[i].x - v[4 * i + 1] * cs[i].y, i0 = v[4 * i] * cs[i].y + v[4 * i + 1] * cs[i].x;
              const float r1 = v[4 * i + 2] * cs[i].z - v[4 * i + 3] * cs[i].w, i1 = v[4 * i + 2] * cs[i].w + v[4 * i + 3] * cs[i].z;
              v[4 * i] = r0; v[4 * i + 1] = i0; v[4 * i + 2] = r1; v[4 * i + 3] = i1;
            }
          }
          if (p.relu) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
            if (drop_on) {
              const uint32_t idx = (uint32_t)(((long long)row0 + lane) * p.nout + cb * BN + c * 64 + sb * 32);
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = sam2b200::dropout_keep(dkey, idx + i, p.drop_out.thresh) ? v[i] * p.drop_out.inv_keep : 0.f;
            }
          }
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) pk[i] = pack_bf16(v[2 * i], v[2 * i + 1]);
          if (BN == 64 && p.dot_rows != nullptr && row0 + lane < p.rows) {      // row dot product with the values the consumer will read (bf16)
            const float4* orow = reinterpret_cast<const float4*>(p.dot_rows + (long long)(row0 + lane) * 64 + sb * 32);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 o4 = __ldg(orow + i);
              const float2 lo = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&pk[2 * i]));
              const float2 hi = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&pk[2 * i + 1]));
              dot += lo.x * o4.x + lo.y * o4.y + hi.x * o4.z + hi.y * o4.w;
            }
          }
#pragma unroll
          for (int q = 0; q < 4; ++q)
            sts128(srow + (((sb * 4 + q) ^ (lane & 7)) << 4), pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
        }
        if (BN == 64 && p.dot_out != nullptr && row0 + lane < p.rows) p.dot_out[row0 + lane] = dot;
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          if (row0 < p.rows) tma_store_3d(mo, &sh.stage[warp][0], ocol0 + c * 64, row0, 0);   // rows beyond R are clipped by the TMA unit
          tma_store_commit();
        }
      }
      if (kChunks == 1 && half == 1) {            // BN = 64: the upper four warps only release the accumulator
        tc_fence_before();
        mbar_arrive(&sh.acc_free[ab]);
      }
    }
    if (lane == 0) tma_store_wait_read();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem, kTmemCols);
}

}  // namespace gemm

namespace {

// row-major bf16 matrix [rows, cols] with leading dimension ld (elements) as a 3-D tensor map (cols, rows, 1), box 64 columns x
// box_rows, 128-byte swizzle; out-of-range rows / columns load as zeros and are clipped on store
int make_map(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld, int box_rows) {
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
    snprintf(sam2b200::last_error_buffer(), 512, "cuTensorMapEncodeTiled (gemm) failed (%d) rows=%lld cols=%lld ld=%lld", (int)r, rows, cols, ld);
    return SAM2B200_ERR_DRIVER;
  }
  return SAM2B200_OK;
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

template <int BN, int B_MN>
int launch(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap* mc, const gemm::Params& p, cudaStream_t stream) {
  const size_t smem = sizeof(gemm::Shared<BN>) + 1024;
  static_assert(sizeof(gemm::Shared<BN>) + 1024 <= 227 * 1024, "shared memory of one CTA");
  cudaError_t e = cudaFuncSetAttribute(gemm::gemm_kernel<BN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return sam2b200::fail(SAM2B200_ERR_CUDA, cudaGetErrorString(e));
  const unsigned grid = (unsigned)(p.n_tiles < sm_count() ? p.n_tiles : sm_count());
  gemm::gemm_kernel<BN, B_MN><<<grid, gemm::kThreads, smem, stream>>>(ma, mb, mc[0], mc[1], mc[2], p);
  return sam2b200::check_launch("gemm");
}

bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace

extern "C" {

// C[R, Nout] = epi(a[R, K] . B + bias), C split over n_out = Nout / out_width <= 3 bf16 tensors [R, out_width] (row stride ldc):
// Nout = 64 (one output) or a multiple of 256 up to 2048, out_width a multiple of 256 (or 64), K % 64 == 0.
// b_layout 0: b = W[Nout, K] (row stride ldb), C = a W^T;  b_layout 1: b = W[K, Nout] (row stride ldb), C = a W.  bias: [Nout] fp32 | NULL.
// rope_cols (a multiple of 256, table != NULL): the leading output columns are rotated with table [period, 128] (cos, sin) -- pair
// index = (column mod 256) / 2 -- for rows whose position (row % rows_per_item) is < n_rope_rows, table row = position % period.
// relu != 0: max(., 0) on the biased accumulator, then (drop_p > 0) inverted dropout with element index row * Nout + column.
// dot_rows / dot_out (Nout = 64, both or neither): dot_out[r] = sum_c bf16(C[r, c]) * dot_rows[r, c], dot_rows [R, 64] fp32 contiguous.
int sam2b200_gemm_ex(void* out0, void* out1, void* out2, int out_width, long long ldc, const void* a, long long lda, const void* b,
                     long long ldb, int b_layout, long long R, int K, int Nout, const float* bias, int rope_cols, const float* table,
                     int rows_per_item, int n_rope_rows, int period, int relu, float drop_p, const unsigned long long* drop_seed,
                     unsigned drop_site, const float* dot_rows, float* dot_out, cudaStream_t stream) {
  const int bn = Nout == 64 ? 64 : 256;
  const int n_out = out_width > 0 ? Nout / out_width : 0;
  if (!out0 || !a || !b || R <= 0 || R > 0x7fffffffLL - 256 || K <= 0 || (K % 64) || Nout <= 0 || (Nout % bn) || Nout > 2048 ||
      out_width <= 0 || (out_width % bn) || n_out * out_width != Nout || n_out > 3 || (n_out > 1 && !out1) || (n_out > 2 && !out2) ||
      (b_layout != 0 && b_layout != 1) || lda < K || ldc < out_width || ldb < (b_layout ? Nout : K) || ((lda | ldb | ldc) & 7) ||
      rope_cols < 0 || (rope_cols % 256) || rope_cols > Nout || (rope_cols > 0 && (bn != 256 || !table || rows_per_item <= 0 || n_rope_rows < 0 || period <= 0)) ||
      drop_p < 0.f || drop_p >= 1.f || (drop_p > 0.f && drop_seed && R * (long long)Nout >= (1LL << 32)) ||
      ((dot_rows != nullptr) != (dot_out != nullptr)) || (dot_rows && Nout != 64) ||
      !al16(out0) || !al16(out1) || !al16(out2) || !al16(a) || !al16(b) || !al16(bias) || !al16(table) || !al16(dot_rows))
    return sam2b200::fail(SAM2B200_ERR_INVALID, "gemm: bad arguments (Nout = 64 | k x 256 <= 2048, K % 64 == 0, 16-byte aligned tensors, strides % 8 == 0)");
  CUtensorMap ma, mb, mc[3];
  int rc;
  if ((rc = make_map(&ma, a, R, K, lda, 128))) return rc;
  if (b_layout == 0) { if ((rc = make_map(&mb, b, Nout, K, ldb, bn))) return rc; }
  else               { if ((rc = make_map(&mb, b, K, Nout, ldb, 64))) return rc; }
  void* outs[3] = {out0, out1 ? out1 : out0, out2 ? out2 : out0};
  for (int i = 0; i < 3; ++i)
    if ((rc = make_map(&mc[i], outs[i], R, out_width, ldc, 32))) return rc;
  gemm::Params p{};
  p.rows = (int)R; p.n_col_blocks = Nout / bn; p.n_tiles = (int)((R + gemm::kBlockM - 1) / gemm::kBlockM) * p.n_col_blocks;
  p.blocks_per_out = out_width / bn; p.ksteps = K / gemm::kBlockK; p.nout = Nout; p.bias = bias;
  p.rope_blocks = rope_cols / 256; p.table = reinterpret_cast<const float2*>(table); p.rows_per_item = rows_per_item > 0 ? rows_per_item : 1;
  p.n_rope_rows = n_rope_rows; p.period = period > 0 ? period : 1;
  if (rope_cols > 0) { int w = (int)(sqrt((double)p.period) + 0.5); p.rope_w = (w * w == p.period) ? w : 0; }
  p.relu = relu; p.drop_out = sam2b200::make_dropout(drop_seed, drop_site, drop_p);
  p.dot_rows = dot_rows; p.dot_out = dot_out;
  if (bn == 256) return b_layout ? launch<256, 1>(ma, mb, mc, p, stream) : launch<256, 0>(ma, mb, mc, p, stream);
  return b_layout ? launch<64, 1>(ma, mb, mc, p, stream) : launch<64, 0>(ma, mb, mc, p, stream);
}

// One output of width No = 256 | 64: c[R, No] = a . B (+ bias); table != NULL rotates the whole output (the memory-key projection).
int sam2b200_gemm(void* c, long long ldc, const void* a, long long lda, const void* b, long long ldb, int b_layout, long long R, int K,
                  int No, const float* bias, const float* table, int rows_per_item, int n_rope_rows, int period, const float* dot_rows, float* dot_out,
                  cudaStream_t stream) {
  if (No != 256 && No != 64) return sam2b200::fail(SAM2B200_ERR_INVALID, "gemm: No must be 256 or 64");
  return sam2b200_gemm_ex(c, nullptr, nullptr, No, ldc, a, lda, b, ldb, b_layout, R, K, No, bias, table ? No : 0, table, rows_per_item,
                          n_rope_rows, period, 0, 0.f, nullptr, 0, dot_rows, dot_out, stream);
}

}  // extern "C"
