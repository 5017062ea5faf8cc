// Weight-gradient GEMMs of the MemoryAttention backward as ONE split-K tcgen05 kernel that accumulates IN PLACE:
//
//     C[Mo, No] (fp32, += ) = A[R, Mo]^T . B[R, No]        A = output gradient of a linear layer, B = its input, bf16
//
// (dW = dY^T X for nn.Linear: memory_attention.py:97 linear1/linear2, transformer.py:213-216 q/k/v/out projections.)
// R is the long dimension (B * N = 32 256 rows at cfg2, B * M = 227 360 for the memory-key projection) and the outputs
// are small (256 x 64 ... 2048 x 256), so the work is split over R: every CTA owns a [256 x NT] output tile and a slice of
// the rows, and adds its partial tile to C with vector fp32 reductions (red.global.add.v4.f32) -- which is exactly the
// "accumulate into the gradient bucket" semantics the backward needs (beta = 1), without the workspace + splitKreduce pass
// cuBLAS runs for the same shapes (profiles/r2_ncu_launches_bench_cfg2.txt: 7 split-K GEMMs + 7 reduce kernels per layer).
//
// Blackwell mapping: both operands are MN-major for this product (the contraction runs over rows, which are the slow
// dimension of both matrices) -- the [64 rows x 64 columns] TMA boxes (128-byte swizzle) are consumed directly as
// MN-major UMMA operands (same descriptors as the P.V product of the attention kernels, csrc/attn_kernels.cuh).
//   warp 8  TMA producer: per 64-row slab 4 boxes of A (256 output rows) + NT / 64 boxes of B, 3- or 4-stage ring
//   warp 9  tcgen05.mma issuer: two [128 x NT] fp32 accumulators in tensor memory (2 x NT columns), 4 k-steps per slab each
//   warps 0-7  epilogue: TMEM -> registers -> red.global.add.v4.f32 (each thread owns 32 contiguous floats of a C row)
//
// BIAS GRADIENT (dbias[Mo] += column sums of A = A^T . 1): the MN-major B operand is 64-column boxes one LBO apart, so a constant
// box of ones behind the last B box turns the column sum into 16 more accumulator columns of the SAME instructions (N = NT + 16;
// only the CTAs of output column tile 0 do it).  The attention backward's gradient epilogues then no longer need the
// 31-shuffle butterfly + atomics per 32 x 32 block (1.4 us of a 5 us epilogue, profiles/r2_timeline_epilogue.txt).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "abi_common.cuh"
#include "sm100.cuh"
#include "tma_desc.cuh"

namespace wgrad {

using namespace sm100;

// Output tile per CTA: [TM x NT] with TM = 256 (two 128-row accumulators) for the wide outputs -- every B slab is shared by
// two MMAs, which halves the L2 -> shared-memory traffic -- or TM = NT = 128 for the [256 x 256] / [768 x 256] outputs, where the
// fp32 reductions dominate: 148 CTAs x 256 KB of partial tiles (39 MB of red traffic, ~26 us) shrink to 148 x 64 KB.
constexpr int kSlab = 64;            // contraction rows per pipeline stage
constexpr int kBoxBytes = kSlab * 128;   // one [64 rows x 64 cols] bf16 box: 8 KB
constexpr int kThreads = 320;
constexpr int kEpiWarps = 8;
constexpr int kMaxStages = 4;

template <int TM, int NT>
struct Shared {
  static constexpr bool kBias = NT <= 128;                                         // room for 16 more accumulator columns
  static constexpr int kLoadBytes = (TM / 64 + NT / 64) * kBoxBytes;               // 64 KB (256 x 256) | 40 KB (256 x 64) | 32 KB (128 x 128)
  static constexpr int kStageBytes = kLoadBytes + (kBias ? kBoxBytes : 0);         // + the box of ones
  static constexpr int kStages = (kStageBytes >= 65536) ? 3 : 4;
  static constexpr int kAccStride = (NT == 64) ? 128 : ((TM == 256 && NT == 128) ? 256 : NT);   // TMEM columns between the two accumulators (NT = 128 carries 16 bias columns)
  alignas(1024) uint8_t tiles[kStages][kStageBytes];     // [A box 0..3 | B box 0..NT/64-1 | ones]
  alignas(8) uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t acc_done;
  uint64_t ones_ready;
  uint32_t tmem_base;
};

struct Params {
  float* c;                 // [Mo, ldc] fp32, accumulated into
  long long ldc;
  long long rows;           // R
  long long rows_per_split; // multiple of 64
  int n_tiles_m, n_tiles_n; // output tiles
  float* dbias;             // [Mo] fp32 or nullptr: += column sums of A
  float* c2;                // two-operand form (TM = 256, NT = 128): output columns 64..127 go to c2 [Mo, ldc2] (B2's product), else nullptr
  long long ldc2;
};

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <int TM, int NT>
__global__ void __launch_bounds__(kThreads, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap map_a,   // A [R, Mo] bf16, box 64 cols x 64 rows
             const __grid_constant__ CUtensorMap map_b,   // B [R, No] bf16, box 64 cols x 64 rows
             const __grid_constant__ CUtensorMap map_b2,  // two-operand form: second B [R, 64] (columns 64..127 of the tile), else = map_b
             const Params p) {
  using Sh = Shared<TM, NT>;
  constexpr int kTileM = TM;
  constexpr int kAcc = TM / 128;
  extern __shared__ uint8_t smem_raw[];
  Sh& sh = *reinterpret_cast<Sh*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x, split = blockIdx.y;
  const int tm = tile / p.n_tiles_n, tn = tile % p.n_tiles_n;
  const long long r_begin = (long long)split * p.rows_per_split;
  const long long r_end = min(p.rows, r_begin + p.rows_per_split);
  const int nslab = (int)((r_end - r_begin + kSlab - 1) / kSlab);      // >= 1 by construction of the grid
  constexpr int kStages = Sh::kStages;
  constexpr int kAccStride = Sh::kAccStride;
  constexpr uint32_t kTmemCols = (Sh::kBias && !(TM == 256 && NT == 128)) ? 256 : 512;     // 2 x 128 (NT = 64) | 144 (NT = 128) | 2 x 256
  const bool do_bias = Sh::kBias && p.dbias != nullptr && tn == 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&sh.full[s], 1); mbar_init(&sh.empty[s], 1); }
    mbar_init(&sh.acc_done, 1);
    mbar_init(&sh.ones_ready, kEpiWarps * 32);
    fence_barrier_init();
  }
  if (warp == 8 && lane == 0) { prefetch_tmap(&map_a); prefetch_tmap(&map_b); prefetch_tmap(&map_b2); }
  if (warp == 9) { tmem_alloc(&sh.tmem_base, kTmemCols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sh.tmem_base;

  if (warp == 8) {
    // ===================== TMA producer =====================
    const bool leader = elect_one();
    for (int j = 0; j < nslab; ++j) {
      const int s = j % kStages;
      mbar_wait(&sh.empty[s], ((j / kStages) & 1) ^ 1);
      if (leader) {
        const int row0 = (int)(r_begin + (long long)j * kSlab);
        mbar_arrive_expect_tx(&sh.full[s], Sh::kLoadBytes);
#pragma unroll
        for (int c = 0; c < kTileM / 64; ++c)
          tma_load_3d(&sh.tiles[s][c * kBoxBytes], &map_a, &sh.full[s], tm * kTileM + c * 64, row0, 0);
#pragma unroll
        for (int c = 0; c < NT / 64; ++c) {
          if (p.c2 != nullptr)   // two operands of width 64: box c comes from operand c
            tma_load_3d(&sh.tiles[s][(kTileM / 64 + c) * kBoxBytes], c == 0 ? &map_b : &map_b2, &sh.full[s], 0, row0, 0);
          else
            tma_load_3d(&sh.tiles[s][(kTileM / 64 + c) * kBoxBytes], &map_b, &sh.full[s], tn * NT + c * 64, row0, 0);
        }
      }
      __syncwarp();
    }
  } else if (warp == 9) {
    // ===================== MMA issuer =====================
    const bool leader = elect_one();
    constexpr uint32_t idesc_plain = make_idesc_bf16(128, NT, 1, 1);        // A^T and B both MN-major
    constexpr uint32_t idesc_ones = make_idesc_bf16(128, Sh::kBias ? NT + 16 : NT, 1, 1);
    const uint32_t idesc = do_bias ? idesc_ones : idesc_plain;
    if (do_bias) mbar_wait(&sh.ones_ready, 0);
    const uint32_t base_lo = desc_lo_sw128(smem_u32(&sh.tiles[0][0]), kBoxBytes);   // LBO = stride between 64-column boxes
    for (int j = 0; j < nslab; ++j) {
      const int s = j % kStages;
      mbar_wait(&sh.full[s], (j / kStages) & 1);
      tc_fence_after();
      if (leader) {
        const uint32_t st = base_lo + s * (Sh::kStageBytes >> 4);
        const uint32_t b_lo = st + ((kTileM / 64) * kBoxBytes >> 4);
#pragma unroll
        for (int h = 0; h < kAcc; ++h) {
          const uint32_t a_lo = st + (h * 2 * kBoxBytes >> 4);                // rows h * 128 .. of the output tile: boxes 2h, 2h + 1
#pragma unroll
          for (int ks = 0; ks < kSlab / 16; ++ks)                             // 16 contraction rows = 2 groups of 8 rows = 2048 B
            umma_ss_lohi(tmem + h * kAccStride, a_lo + ks * (2048 >> 4), b_lo + ks * (2048 >> 4), kDescHiSw128_1024, idesc, (j > 0) || (ks > 0));
        }
        umma_commit(&sh.empty[s]);
        if (j + 1 >= nslab) umma_commit(&sh.acc_done);
      }
      __syncwarp();
    }
  } else {
    // ===================== epilogue: partial tile -> C with vector reductions =====================
    const int quarter = warp & 3, half = warp >> 2;
    const uint32_t lane_addr = tmem + (uint32_t(quarter * 32) << 16);
    if (do_bias) {   // the constant operand: bf16 ones in the box behind the B boxes of every stage (2 x 16 bytes per thread and stage)
#pragma unroll
      for (int s = 0; s < kStages; ++s) {
        const uint32_t box = smem_u32(&sh.tiles[s][Sh::kLoadBytes]) + threadIdx.x * 16;
        sts128(box, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
        sts128(box + 4096, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
      }
      fence_proxy_async();
      mbar_arrive(&sh.ones_ready);
    }
    mbar_wait(&sh.acc_done, 0);
    tc_fence_after();
    constexpr int kColsPerWarp = NT / 2;                 // this warp's share of the tile's columns
#pragma unroll
    for (int h = 0; h < kAcc; ++h) {
      const long long crow = (long long)tm * kTileM + h * 128 + quarter * 32 + lane;
      float* dst = (p.c2 != nullptr && half == 1) ? p.c2 + crow * p.ldc2        // two-operand form: this warp's 64 columns are the second product
                                                  : p.c + crow * p.ldc + (long long)tn * NT + half * kColsPerWarp;
#pragma unroll 1
      for (int cc = 0; cc < kColsPerWarp / 32; ++cc) {
        uint32_t o[32];
        SAM2B200_TMEM_LD32(lane_addr + h * kAccStride + half * kColsPerWarp + cc * 32, o);
        tmem_wait_ld();
#pragma unroll
        for (int k = 0; k < 8; ++k)
          red_add_v4(dst + cc * 32 + 4 * k, __uint_as_float(o[4 * k]), __uint_as_float(o[4 * k + 1]), __uint_as_float(o[4 * k + 2]),
                     __uint_as_float(o[4 * k + 3]));
      }
      if (do_bias && half == 0) {     // columns NT .. NT + 15 all hold sum_r A[r, row]
        uint32_t o[16];
        SAM2B200_TMEM_LD16(lane_addr + h * kAccStride + NT, o);
        tmem_wait_ld();
        atomicAdd(p.dbias + crow, __uint_as_float(o[0]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem, kTmemCols);
}

}  // namespace wgrad

namespace {

int make_rows_map(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld) {
  sam2b200::PFN_encodeTiled enc = sam2b200::get_encode_tiled();
  if (!enc) return sam2b200::fail(SAM2B200_ERR_DRIVER, "cuTensorMapEncodeTiled entry point not found");
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, 1};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)rows * (cuuint64_t)ld * 2};
  cuuint32_t box[3] = {64, 64, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(sam2b200::last_error_buffer(), 512, "cuTensorMapEncodeTiled (wgrad) failed (%d) rows=%lld cols=%lld ld=%lld", (int)r, rows, cols, ld);
    return SAM2B200_ERR_DRIVER;
  }
  return SAM2B200_OK;
}

template <int TM, int NT>
int launch(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mb2, const wgrad::Params& p, int splits, cudaStream_t stream) {
  const size_t smem = sizeof(wgrad::Shared<TM, NT>) + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(wgrad::wgrad_kernel<TM, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return sam2b200::fail(SAM2B200_ERR_CUDA, cudaGetErrorString(e));
    attr_set = true;
  }
  dim3 grid((unsigned)(p.n_tiles_m * p.n_tiles_n), (unsigned)splits, 1);
  wgrad::wgrad_kernel<TM, NT><<<grid, wgrad::kThreads, smem, stream>>>(ma, mb, mb2, p);
  return sam2b200::check_launch("wgrad");
}

}  // namespace

extern "C" {

// c [Mo, ldc] fp32 += a[R, Mo]^T . b[R, No]; a: bf16, row stride lda elements; b: bf16, row stride ldb.  Mo a multiple of 256,
// No = 64 or a multiple of 256.  Partial tiles are added with fp32 reductions (order not fixed: results are reproducible to
// fp32 round-off only, like every split-K scheme with atomics).  dbias (nullable): [Mo] fp32 += column sums of a (the bias
// gradient of the same linear layer) from the same instructions; only for the [256 x 64] / [128 x 128] tile shapes
// (No = 64, or No = 256 with Mo <= 768).
int sam2b200_wgrad(float* c, long long ldc, const void* a, long long lda, const void* b, long long ldb, long long R, int Mo, int No,
                   float* dbias, cudaStream_t stream) {
  if (!c || !a || !b || R <= 0 || R >= (1LL << 31) || Mo <= 0 || (Mo % 256) || !(No == 64 || (No > 0 && No % 256 == 0)) || ldc < No ||
      lda < Mo || ldb < No || (lda % 8) || (ldb % 8) || (ldc % 4) || (reinterpret_cast<uintptr_t>(c) & 15) ||
      (reinterpret_cast<uintptr_t>(a) & 15) || (reinterpret_cast<uintptr_t>(b) & 15))
    return sam2b200::fail(SAM2B200_ERR_INVALID, "wgrad: bad arguments (Mo % 256 == 0, No = 64 or a multiple of 256, 16-byte aligned rows)");
  CUtensorMap ma, mb;
  int rc;
  if ((rc = make_rows_map(&ma, a, R, Mo, lda))) return rc;
  if ((rc = make_rows_map(&mb, b, R, No, ldb))) return rc;
  // tile shape: [256 x 64] for the 64-wide outputs, [128 x 128] when the output is at most 768 x 256 (reduction traffic
  // dominates), [256 x 256] for the [256 x 2048] / [2048 x 256] MLP weights (operand traffic dominates)
  const bool small = No == 256 && Mo <= 768;
  if (dbias && !(No == 64 || small)) return sam2b200::fail(SAM2B200_ERR_INVALID, "wgrad: dbias needs No = 64 or (No = 256 and Mo <= 768)");
  const int tm_sz = small ? 128 : 256;
  const int nt = (No == 64) ? 64 : (small ? 128 : 256);
  wgrad::Params p{};
  p.c = c; p.ldc = ldc; p.rows = R; p.dbias = dbias;
  p.n_tiles_m = Mo / tm_sz; p.n_tiles_n = No / nt;
  const int tiles = p.n_tiles_m * p.n_tiles_n;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long slabs = (R + wgrad::kSlab - 1) / wgrad::kSlab;
  long long splits = sms / tiles;                          // one wave of CTAs (never a second, nearly empty one)
  if (splits > slabs / 4) splits = slabs / 4;              // >= 4 slabs per CTA: the prologue / epilogue must amortise
  if (splits < 1) splits = 1;
  long long slabs_per = (slabs + splits - 1) / splits;
  splits = (slabs + slabs_per - 1) / slabs_per;            // no empty split
  p.rows_per_split = slabs_per * wgrad::kSlab;
  if (nt == 64) return launch<256, 64>(ma, mb, mb, p, (int)splits, stream);
  if (small) return launch<128, 128>(ma, mb, mb, p, (int)splits, stream);
  return launch<256, 256>(ma, mb, mb, p, (int)splits, stream);
}

// Two products that share the long operand in ONE pass over it:  c [256, ldc] += a^T . b   and   c2 [256, ldc2] += a^T . b2, with a [R, 256],
// b and b2 [R, 64] (bf16; row strides lda / ldb / ldb2).  The cross-attention key projection's weight gradient (b = the key source) and
// the per-segment sums of the key gradient (b2 = the one-hot segment indicator of the packed bank: its position tensors' gradients,
// fused_stack.segment_indicator) both contract dk [B M, 256] over its rows.  dbias (nullable, [256]) += column sums of a.
int sam2b200_wgrad2(float* c, long long ldc, float* c2, long long ldc2, const void* a, long long lda, const void* b, long long ldb,
                    const void* b2, long long ldb2, long long R, float* dbias, cudaStream_t stream) {
  if (!c || !c2 || !a || !b || !b2 || R <= 0 || R >= (1LL << 31) || ldc < 64 || ldc2 < 64 || lda < 256 || ldb < 64 || ldb2 < 64 ||
      (lda % 8) || (ldb % 8) || (ldb2 % 8) || (ldc % 4) || (ldc2 % 4) ||
      ((reinterpret_cast<uintptr_t>(c) | reinterpret_cast<uintptr_t>(c2) | reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) |
        reinterpret_cast<uintptr_t>(b2)) & 15))
    return sam2b200::fail(SAM2B200_ERR_INVALID, "wgrad2: bad arguments (a [R, 256], b / b2 [R, 64], 16-byte aligned rows)");
  CUtensorMap ma, mb, mb2;
  int rc;
  if ((rc = make_rows_map(&ma, a, R, 256, lda))) return rc;
  if ((rc = make_rows_map(&mb, b, R, 64, ldb))) return rc;
  if ((rc = make_rows_map(&mb2, b2, R, 64, ldb2))) return rc;
  wgrad::Params p{};
  p.c = c; p.ldc = ldc; p.c2 = c2; p.ldc2 = ldc2; p.rows = R; p.dbias = dbias; p.n_tiles_m = 1; p.n_tiles_n = 1;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long slabs = (R + wgrad::kSlab - 1) / wgrad::kSlab;
  long long splits = sms;
  if (splits > slabs / 4) splits = slabs / 4;
  if (splits < 1) splits = 1;
  long long slabs_per = (slabs + splits - 1) / splits;
  splits = (slabs + slabs_per - 1) / slabs_per;
  p.rows_per_split = slabs_per * wgrad::kSlab;
  return launch<256, 128>(ma, mb, mb2, p, (int)splits, stream);
}

}  // extern "C"
