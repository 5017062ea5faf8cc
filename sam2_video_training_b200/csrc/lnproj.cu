// Pre-norm block head of MemoryAttentionLayer in ONE kernel (memory_attention.py:58-64, 66-81, 95-97 + transformer.py:277-302):
//
//     x'      = x + dropout(res)                         residual stream (fp32), res = output of the previous branch (bf16)
//     y       = LayerNorm(x') * gamma + beta             bf16, kept for the weight gradients of the projection
//     OUT     = epi( y . W[Nout, 256]^T + bias )         epi = axial RoPE on the leading outputs (q | k) or ReLU (+ hidden dropout)
//
// It replaces, per call, ln_fwd + 1..3 cuBLAS addmm + 0..2 RoPE passes (11 launches per layer, 3 now): the normalised
// activations go from registers to the tensor core without a round trip through HBM, q / k never exist un-rotated
// ("RoPE applied on load" of the north star: the rotation acts on the fp32 accumulator, before the one rounding to bf16).
//   self-attention head : W = [Wq | Wk | Wv] (768 x 256), RoPE on q and k, three [R, 256] outputs
//   cross-attention head: W = Wq (256 x 256), RoPE, one output
//   MLP head            : W = W1 (2048 x 256), bias + ReLU (+ dropout), one [R, 2048] output
//
// Blackwell mapping: one CTA per 128 rows, TWO CTAs PER SM (104 KB of shared memory and 256 tensor-memory columns each):
// while one CTA is in its HBM-bound LayerNorm prologue the other one is in its tensor-core phase, and all 252 row tiles of
// the cfg2 step are resident at once (no wave tail).  First version (one CTA per SM, 128-column chunks): the prologue alone
// took twice as long as the stand-alone ln_fwd kernel (8 warps x 4 rows in flight per SM) -- profiles/r2_lnproj_bench.txt.
//   warps 0-7  LayerNorm prologue -- one warp per row, four rows in flight per warp (128-bit loads), the bf16 result
//              written straight into shared memory in the K-major 128-byte-swizzle UMMA layout and moved to TENSOR MEMORY
//              (tcgen05.st) as the A operand; the same warps are the epilogue afterwards (bias, rotation / ReLU on the fp32
//              accumulator, bf16 staging in the swizzled box layout, one TMA store per 32 rows and chunk)
//   warp 8     TMA producer: W streamed in [64 x 256] K-major chunks (L2 resident), 2-stage ring (the ring first stages X)
//   warp 9     tcgen05.mma issuer: 128 x 64 x 256 per chunk, A from TMEM, double-buffered TMEM accumulators
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "abi_common.cuh"
#include "attn_kernels.cuh"
#include "tma_desc.cuh"

namespace lnproj {

using namespace sm100;
using attn::kBoxBytes;
using attn::kSlabBytes;

constexpr int kBlockM = 128;
constexpr int kBlockN = 64;
constexpr int kK = 256;
constexpr int kWSlabBytes = kBlockN * 128;           // 8 KB: one [64 rows x 128 B] slab of a W chunk
constexpr int kWTileBytes = kBlockN * kK * 2;        // 32 KB: four such slabs
constexpr int kStageBytes = kBlockM * kBlockN * 2;   // 16 KB: output staging, one [128 rows x 128 B] slab
constexpr int kThreads = 320;
constexpr int kEpiWarps = 8;
constexpr int kMaxNout = 2048;
constexpr int kTmemCols = 256;
constexpr uint32_t kColA = 0, kColAcc0 = 128, kColAcc1 = 192;
static_assert(2 * kWTileBytes == 4 * kSlabBytes, "the two W stages together stage the [128 x 256] X tile");

struct Shared {
  alignas(1024) uint8_t w_tiles[2][kWTileBytes];     // both stages together hold the normalised X tile first (four 16 KB slabs)
  alignas(1024) uint8_t stage[2][kStageBytes];
  alignas(8) uint64_t w_full[2];
  uint64_t w_empty[2];
  uint64_t acc_full[2];
  uint64_t acc_free[2];
  uint64_t a_ready;
  alignas(16) float bias[kMaxNout];
  alignas(16) float gamma[kK];
  alignas(16) float beta[kK];
  uint32_t tmem_base;
};
static_assert(2 * (sizeof(Shared) + 1024) <= 227 * 1024, "two CTAs per SM");

struct Params {
  // ---- LayerNorm prologue
  const float* x;                 // [R, 256] fp32 residual stream
  const __nv_bfloat16* res;       // [R, 256] bf16 branch output added first, or nullptr
  float* x_out;                   // [R, 256] fp32 x + dropout(res) (required iff res)
  const float* gamma;
  const float* beta;
  __nv_bfloat16* y_out;           // [R, 256] bf16 LayerNorm output (kept for the backward), or nullptr
  float* mean;
  float* rstd;
  float eps;
  sam2b200::Dropout drop_res;     // dropout on res; element index = row * 256 + column
  long long rows;
  // ---- GEMM + epilogue
  int n_chunks;                   // Nout / 64
  const __nv_bfloat16* bias;      // [Nout] or nullptr
  int chunks_per_out;             // output tensor of chunk j = j / chunks_per_out, its column (j % chunks_per_out) * 64
  int rope_chunks;                // leading 128-column groups (two chunks each) that are rotated (pair index = (column mod 256) / 2)
  const float2* table;            // [period, 128] (cos, sin)
  int rows_per_item, n_rope_rows, period, rope_w;
  int relu;                       // max(., 0) on the biased accumulator
  sam2b200::Dropout drop_out;     // dropout after the ReLU; element index = row * Nout + column
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ void unpack8(const uint4 u, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 t = __bfloat1622float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}

__global__ void __launch_bounds__(kThreads, 2)
lnproj_kernel(const __grid_constant__ CUtensorMap map_w,      // W [Nout, 256] bf16, box 64 x 64
              const __grid_constant__ CUtensorMap map_o0,     // outputs [R, out_width] bf16, box 64 x 32 (store)
              const __grid_constant__ CUtensorMap map_o1, const __grid_constant__ CUtensorMap map_o2, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  Shared& sh = *reinterpret_cast<Shared*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row_tile = blockIdx.x;
  const int nc = p.n_chunks;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sh.w_full[i], 1); mbar_init(&sh.w_empty[i], 1);
      mbar_init(&sh.acc_full[i], 1); mbar_init(&sh.acc_free[i], kEpiWarps * 32);
    }
    mbar_init(&sh.a_ready, kEpiWarps * 32);
    fence_barrier_init();
  }
  if (warp == 8 && lane == 0) prefetch_tmap(&map_w);
  if (warp == 0 && lane == 0) { prefetch_tmap(&map_o0); prefetch_tmap(&map_o1); prefetch_tmap(&map_o2); }
  if (warp == 9) { tmem_alloc(&sh.tmem_base, kTmemCols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sh.tmem_base;

  if (warp == 8) {
    // ===================== TMA producer: W chunks =====================
    const bool leader = elect_one();
    mbar_wait(&sh.a_ready, 0);                        // the ring staged the normalised X tile until it reached tensor memory
    for (int j = 0; j < nc; ++j) {
      const int s = j & 1;
      mbar_wait(&sh.w_empty[s], ((j >> 1) & 1) ^ 1);
      if (leader) {
        mbar_arrive_expect_tx(&sh.w_full[s], kWTileBytes);
#pragma unroll
        for (int c = 0; c < 4; ++c) tma_load_3d(&sh.w_tiles[s][c * kWSlabBytes], &map_w, &sh.w_full[s], c * 64, j * kBlockN, 0);
      }
      __syncwarp();
    }
  } else if (warp == 9) {
    // ===================== MMA issuer =====================
    const bool leader = elect_one();
    constexpr uint32_t idesc = make_idesc_bf16(kBlockM, kBlockN, 0, 0);       // A (TMEM, K-major) . W^T, W K-major
    const uint32_t w_lo0 = desc_lo_sw128(smem_u32(&sh.w_tiles[0][0]), 16);
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
#pragma unroll
        for (int ks = 0; ks < kK / 16; ++ks)      // K-major SW128: slab ks / 4 (8 KB), 32 B per k-step inside the 128 B row
          umma_ts_lohi(d, tmem + kColA + ks * 8, wlo + (ks >> 2) * (kWSlabBytes >> 4) + (ks & 3) * 2, kDescHiSw128_1024, idesc, ks > 0);
        umma_commit(&sh.w_empty[s]);
        umma_commit(&sh.acc_full[s]);
      }
      __syncwarp();
    }
  } else {
    // ===================== warps 0..7: LayerNorm prologue, then epilogue =====================
    {
      // bias -> shared memory (fp32), read back with broadcast loads in the epilogue
      for (int i = threadIdx.x; i < nc * kBlockN; i += kEpiWarps * 32) sh.bias[i] = p.bias ? __bfloat162float(p.bias[i]) : 0.f;
      if (threadIdx.x < kK) { sh.gamma[threadIdx.x] = p.gamma[threadIdx.x]; sh.beta[threadIdx.x] = p.beta[threadIdx.x]; }
      asm volatile("bar.sync 5, 256;" ::: "memory");
      const bool drop_on = p.drop_res.seed != nullptr && p.res != nullptr;
      const uint32_t key = drop_on ? sam2b200::dropout_key(*p.drop_res.seed, p.drop_res.site) : 0u;
      // this lane's 8 columns are one 16-byte chunk of slab lane / 8
      const uint32_t xs = smem_u32(&sh.w_tiles[0][0]) + (lane >> 3) * kSlabBytes;
      constexpr int kU = 4;                                   // rows in flight per warp
#pragma unroll 1
      for (int it = 0; it < 16 / kU; ++it) {
        float v[kU][8];
        uint4 rr[kU];
        long long grow[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) {                        // all loads first
          const int r = warp * 16 + it * kU + u;
          grow[u] = (long long)row_tile * kBlockM + r;
          if (grow[u] < p.rows) {
            const float4* xp = reinterpret_cast<const float4*>(p.x + grow[u] * kK + lane * 8);
            const float4 a = xp[0], b = xp[1];
            v[u][0] = a.x; v[u][1] = a.y; v[u][2] = a.z; v[u][3] = a.w; v[u][4] = b.x; v[u][5] = b.y; v[u][6] = b.z; v[u][7] = b.w;
            if (p.res != nullptr) rr[u] = *reinterpret_cast<const uint4*>(p.res + grow[u] * kK + lane * 8);
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[u][i] = 0.f;
            rr[u] = make_uint4(0u, 0u, 0u, 0u);
          }
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) {
          const int r = warp * 16 + it * kU + u;
          const bool ok = grow[u] < p.rows;
          if (p.res != nullptr) {
            float rf[8];
            unpack8(rr[u], rf);
            if (drop_on) {
              const uint32_t idx = (uint32_t)(grow[u] * kK + lane * 8);
#pragma unroll
              for (int i = 0; i < 8; ++i) rf[i] = sam2b200::dropout_keep(key, idx + i, p.drop_res.thresh) ? rf[i] * p.drop_res.inv_keep : 0.f;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) v[u][i] += rf[i];
            if (ok) {
              float4* op = reinterpret_cast<float4*>(p.x_out + grow[u] * kK + lane * 8);
              op[0] = make_float4(v[u][0], v[u][1], v[u][2], v[u][3]);
              op[1] = make_float4(v[u][4], v[u][5], v[u][6], v[u][7]);
            }
          }
          float s = 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i) s += v[u][i];
          const float mu = warp_sum(s) * (1.0f / kK);
          float q = 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i) { const float d = v[u][i] - mu; q += d * d; }
          const float rs = rsqrtf(warp_sum(q) * (1.0f / kK) + p.eps);
          if (ok && lane == 0) { p.mean[grow[u]] = mu; p.rstd[grow[u]] = rs; }
          float y[8];
          {
            const float4 g0 = *reinterpret_cast<const float4*>(&sh.gamma[lane * 8]), g1 = *reinterpret_cast<const float4*>(&sh.gamma[lane * 8 + 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&sh.beta[lane * 8]), b1 = *reinterpret_cast<const float4*>(&sh.beta[lane * 8 + 4]);
            const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i) y[i] = ok ? (v[u][i] - mu) * rs * gg[i] + bb[i] : 0.f;
          }
          const uint32_t w0 = pack_bf16(y[0], y[1]), w1 = pack_bf16(y[2], y[3]), w2 = pack_bf16(y[4], y[5]), w3 = pack_bf16(y[6], y[7]);
          if (ok && p.y_out != nullptr) *reinterpret_cast<uint4*>(p.y_out + grow[u] * kK + lane * 8) = make_uint4(w0, w1, w2, w3);
          sts128(xs + r * 128 + (((lane & 7) ^ (r & 7)) << 4), w0, w1, w2, w3);   // K-major SW128: chunk (lane % 8) ^ (row % 8)
        }
      }
    }
    asm volatile("bar.sync 5, 256;" ::: "memory");     // the whole X tile (and the bias) is in shared memory
    const int quarter = warp & 3;
    const int half = warp >> 2;                   // which 32 of a chunk's 64 columns; which two slabs of X
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = tmem + (uint32_t(quarter * 32) << 16);
    {   // X tile: shared -> registers -> TMEM (bf16 pairs); half h moves slabs 2h, 2h + 1
#pragma unroll
      for (int c = half * 2; c < half * 2 + 2; ++c) {
        const uint32_t base = smem_u32(&sh.w_tiles[0][0]) + c * kSlabBytes + row * 128;
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
    fence_proxy_async();      // the staging slabs are handed to the TMA producer (W chunk 1 lands there)
    tc_fence_before();
    mbar_arrive(&sh.a_ready);

    const int row0 = row_tile * kBlockM + quarter * 32;
    const long long grow = (long long)row0 + lane;                     // global row of this thread
    const int pos = (int)(grow % p.rows_per_item);                     // position inside the batch item
    const bool rot_row = p.rope_chunks > 0 && pos < p.n_rope_rows;
    const int tpos = pos % p.period;
    const bool drop_out_on = p.drop_out.seed != nullptr;
    const uint32_t okey = drop_out_on ? sam2b200::dropout_key(*p.drop_out.seed, p.drop_out.site) : 0u;
    const int nout = nc * kBlockN;
    for (int j = 0; j < nc; ++j) {
      const int s = j & 1;
      const int col0 = j * kBlockN + half * 32;                        // first output column of this thread's 32
      // the (cos, sin) row is L2 / L1 resident but far away: fetch it before waiting for the accumulator
      const bool rotate = rot_row && (j >> 1) < p.rope_chunks;
      const float4* tsrc = nullptr;
      if (rotate) {
        const int c256 = col0 & 255;                                   // column inside the 256-wide head
        const int trow = p.rope_w > 0 ? (c256 < 128 ? tpos % p.rope_w : tpos - tpos % p.rope_w) : tpos;
        tsrc = reinterpret_cast<const float4*>(p.table + (long long)trow * 128 + (c256 >> 1));
#ifdef SAM2B200_DEBUG_LNPROJ
        if (threadIdx.x == 0 && blockIdx.x == 0 && j == 0)
          printf("lnproj dbg: table %p tsrc %p trow %d c256 %d period %d w %d rope_chunks %d nc %d\n", (const void*)p.table, (const void*)tsrc, trow,
                 c256, p.period, p.rope_w, p.rope_chunks, nc);
        if (trow < 0 || trow >= p.period || c256 < 0 || c256 > 224) __trap();
#endif
      }
      mbar_wait(&sh.acc_full[s], (j >> 1) & 1);
      tc_fence_after();
      // the staging slab of this buffer was handed to the TMA unit two chunks ago by the half-0 warp of this row quarter
      if (half == 0 && lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");
      const uint32_t srow = smem_u32(&sh.stage[s][0]) + row * 128;
      uint32_t acc[32];
      SAM2B200_TMEM_LD32(lane_addr + (s ? kColAcc1 : kColAcc0) + half * 32, acc);
      tmem_wait_ld();
      tc_fence_before();
      mbar_arrive(&sh.acc_free[s]);
      float v[32];
      {
        const float4* bsrc = reinterpret_cast<const float4*>(&sh.bias[col0]);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 bq = bsrc[i];
          v[4 * i] = __uint_as_float(acc[4 * i]) + bq.x; v[4 * i + 1] = __uint_as_float(acc[4 * i + 1]) + bq.y;
          v[4 * i + 2] = __uint_as_float(acc[4 * i + 2]) + bq.z; v[4 * i + 3] = __uint_as_float(acc[4 * i + 3]) + bq.w;
        }
      }
      if (rotate) {                             // (cos, sin) rows: w + 128 / w distinct rows per CTA (axial addressing), L1 resident
#pragma unroll
        for (int i = 0; i < 8; ++i) {           // (re, im) = (v[2k], v[2k+1]) times (cos + i sin), two pairs per 16-byte load
#ifdef SAM2B200_DEBUG_LNPROJ_NOLOAD
          const float4 f = make_float4(1.f, 0.f, 1.f, 0.f);
#else
          const float4 f = __ldg(tsrc + i);
#endif
          const float re0 = v[4 * i] * f.x - v[4 * i + 1] * f.y, im0 = v[4 * i] * f.y + v[4 * i + 1] * f.x;
          const float re1 = v[4 * i + 2] * f.z - v[4 * i + 3] * f.w, im1 = v[4 * i + 2] * f.w + v[4 * i + 3] * f.z;
          v[4 * i] = re0; v[4 * i + 1] = im0; v[4 * i + 2] = re1; v[4 * i + 3] = im1;
        }
      }
      if (p.relu) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
        if (drop_out_on) {
          const uint32_t idx = (uint32_t)(grow * nout + col0);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = sam2b200::dropout_keep(okey, idx + i, p.drop_out.thresh) ? v[i] * p.drop_out.inv_keep : 0.f;
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)               // this thread's 64 bytes = 16-byte chunks 4 half .. 4 half + 3 of its 128-byte row
        sts128(srow + (((half * 4 + q) ^ (row & 7)) << 4), pack_bf16(v[8 * q + 0], v[8 * q + 1]), pack_bf16(v[8 * q + 2], v[8 * q + 3]),
               pack_bf16(v[8 * q + 4], v[8 * q + 5]), pack_bf16(v[8 * q + 6], v[8 * q + 7]));
      fence_proxy_async();
      asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");     // both halves of the 32 x 64 box are staged
      if (half == 0 && lane == 0) {
        const int cpo = p.chunks_per_out;
        const int which = j / cpo;
        const CUtensorMap* mo = which == 0 ? &map_o0 : (which == 1 ? &map_o1 : &map_o2);
        if (row0 < p.rows)
          tma_store_3d(mo, &sh.stage[s][quarter * kBoxBytes], (j % cpo) * kBlockN, row0, 0);
        tma_store_commit();
      }
    }
    if (half == 0 && lane == 0) tma_store_wait_read();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem, kTmemCols);
}

}  // namespace lnproj

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
    snprintf(sam2b200::last_error_buffer(), 512, "cuTensorMapEncodeTiled (ln_proj) failed (%d) rows=%lld cols=%lld", (int)r, rows, cols);
    return SAM2B200_ERR_DRIVER;
  }
  return SAM2B200_OK;
}

bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace

extern "C" {

// x [R, 256] fp32; res [R, 256] bf16 or NULL (then x_out is ignored); gamma / beta [256] fp32; y_out [R, 256] bf16 or NULL;
// mean / rstd [R] fp32; w [Nout, 256] bf16; bias [Nout] bf16 or NULL.  Outputs: n_out tensors [R, out_width] bf16
// (n_out * out_width = Nout; out_width a multiple of 128; n_out <= 3).  rope_cols (a multiple of 128, <= Nout): the leading
// output columns are rotated with `table` [period, 128] (cos, sin) -- pair index = (column mod 256) / 2 -- for rows whose
// position (row mod rows_per_item) is < n_rope_rows.  relu != 0: ReLU on the biased accumulator, then (drop_out_p > 0)
// inverted dropout with element index row * Nout + column.  drop_res_p > 0: inverted dropout on res, index row * 256 + column.
int sam2b200_ln_proj(const float* x, const void* res, float* x_out, const float* gamma, const float* beta, void* y_out,
                     float* mean, float* rstd, long long R, float eps, const void* w, const void* bias, int Nout, void* out0,
                     void* out1, void* out2, int out_width, int rope_cols, const float* table, int rows_per_item,
                     int n_rope_rows, int period, int relu, float drop_res_p, const unsigned long long* drop_res_seed,
                     unsigned drop_res_site, float drop_out_p, const unsigned long long* drop_out_seed, unsigned drop_out_site,
                     cudaStream_t stream) {
  const int n_out = out_width > 0 ? Nout / out_width : 0;
  if (!x || !gamma || !beta || !mean || !rstd || !w || !out0 || R <= 0 || R > 0x7fffffffLL - 256 || Nout <= 0 ||
      Nout > lnproj::kMaxNout || (Nout % 128) || out_width <= 0 || (out_width % 128) || n_out * out_width != Nout || n_out > 3 ||
      (n_out > 1 && !out1) || (n_out > 2 && !out2) || (res && !x_out) || rope_cols < 0 || (rope_cols % 128) || rope_cols > Nout ||
      (rope_cols > 0 && (!table || rows_per_item <= 0 || n_rope_rows < 0 || period <= 0)) || drop_res_p < 0.f || drop_res_p >= 1.f ||
      drop_out_p < 0.f || drop_out_p >= 1.f || (drop_out_p > 0.f && drop_out_seed && R * Nout >= (1LL << 32)) ||
      !al16(x) || !al16(res) || !al16(x_out) || !al16(y_out) || !al16(w) || !al16(bias) || !al16(out0) || !al16(out1) || !al16(out2) ||
      !al16(gamma) || !al16(beta))
    return sam2b200::fail(SAM2B200_ERR_INVALID, "ln_proj: bad arguments");
  CUtensorMap map_w, map_o[3];
  int rc;
  if ((rc = make_bf16_matrix_map(&map_w, w, Nout, 256, lnproj::kBlockN))) return rc;
  void* outs[3] = {out0, out1 ? out1 : out0, out2 ? out2 : out0};
  for (int i = 0; i < 3; ++i)
    if ((rc = make_bf16_matrix_map(&map_o[i], outs[i], R, out_width, 32))) return rc;
  const size_t smem = sizeof(lnproj::Shared) + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(lnproj::lnproj_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess)   // two CTAs per SM need the full shared-memory carve-out
      e = cudaFuncSetAttribute(lnproj::lnproj_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return sam2b200::fail(SAM2B200_ERR_CUDA, cudaGetErrorString(e));
    attr_set = true;
  }
  lnproj::Params p{};
  p.x = x; p.res = static_cast<const __nv_bfloat16*>(res); p.x_out = x_out; p.gamma = gamma; p.beta = beta;
  p.y_out = static_cast<__nv_bfloat16*>(y_out); p.mean = mean; p.rstd = rstd; p.eps = eps; p.rows = R;
  p.drop_res = sam2b200::make_dropout(drop_res_seed, drop_res_site, drop_res_p);
  p.n_chunks = Nout / lnproj::kBlockN; p.bias = static_cast<const __nv_bfloat16*>(bias);
  p.chunks_per_out = out_width / lnproj::kBlockN;
  p.rope_chunks = rope_cols / 128;   // in units of 128 columns (two chunks)
  p.table = reinterpret_cast<const float2*>(table);
  p.rows_per_item = rows_per_item > 0 ? rows_per_item : 1; p.n_rope_rows = n_rope_rows; p.period = period > 0 ? period : 1;
  {
    int wdt = 0;
    if (rope_cols > 0) { wdt = (int)(sqrt((double)p.period) + 0.5); if (wdt * wdt != p.period) wdt = 0; }
    p.rope_w = wdt;
  }
  p.relu = relu;
  p.drop_out = sam2b200::make_dropout(drop_out_seed, drop_out_site, drop_out_p);
  const unsigned grid = (unsigned)((R + lnproj::kBlockM - 1) / lnproj::kBlockM);
  lnproj::lnproj_kernel<<<grid, lnproj::kThreads, smem, stream>>>(map_w, map_o[0], map_o[1], map_o[2], p);
  return sam2b200::check_launch("ln_proj");
}

}  // extern "C"
