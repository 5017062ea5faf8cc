// sm_100a flash-style attention kernels for SAM2 memory attention: ONE head, head_dim 256,
// bf16 operands (q, k already rotated by the axial RoPE pre-pass, attn.cu), fp32 accumulation.
//
// Two kernel templates cover forward and backward ("Attention kernels"):
//
//   two_gemm_kernel<MODE>   scores = A . X^T  ->  P = f(scores)  ->  ACC += P . Y
//     MODE_FWD : A = Q tile (128 queries), X = K tiles, Y = V tiles, online softmax, O = ACC / l
//     MODE_DV  : A = K tile (128 keys),    X = Q tiles, Y = dO tiles, P^T = exp2(S^T c - LSE2[q])
//
//   three_gemm_kernel<MODE> S = A1 . X^T, dP = A2 . Y^T, dS = P o (dP - Delta), ACC += dS . X
//     MODE_DQ  : A1 = Q, A2 = dO, X = K tiles, Y = V tiles   (LSE2/Delta per row)
//     MODE_DK  : A1 = K, A2 = V,  X = Q tiles, Y = dO tiles  (LSE2/Delta per column)
//
// Hardware mapping (identical in every kernel):
//   * the fixed 128 x 256 operand A lives in TENSOR MEMORY as the bf16 A-operand of tcgen05.mma
//     (.kind::f16, A from TMEM): it arrives by TMA in the last (still unused) ring stage and is moved
//     shared -> registers -> tensor memory (tcgen05.st) once;
//   * results leave through shared memory too: every epilogue warp stages its 32 accumulator rows in the
//     (by then idle) ring in the 128-byte-swizzle box layout and issues its own TMA bulk-tensor stores --
//     fully coalesced writes, rows beyond the tensor clipped by the TMA unit, no block-level barrier;
//   * streamed 64 x 256 tiles arrive by TMA (cp.async.bulk.tensor, 128-byte swizzle) into a
//     shared-memory ring; the SAME tile is consumed K-major (scores GEMM, contraction over the
//     256 features) and MN-major (accumulate GEMM, contraction over the 64 rows);
//   * fp32 accumulators (scores 128 x 64, ACC 128 x 256) live in tensor memory; the
//     probability tile is written back to tensor memory as bf16 (tcgen05.st) and fed to the
//     second GEMM as its A operand, so it never touches shared memory;
//   * warp roles (320 threads): warps 0-7 = softmax / epilogue -- TWO warps per 32-lane TMEM
//     quarter (warp w and w+4 own lanes 32*(w%4)..+31), each owning half of the tile's 64 columns,
//     so every SM sub-partition has two warps to overlap tcgen05.ld / MUFU / FMA latencies;
//     warp 8 = TMA producer, warp 9 = tcgen05.mma issuer + TMEM allocator.
#pragma once

#include "dropout.cuh"
#include "sm100.cuh"

namespace attn {

using namespace sm100;

constexpr int kD = 256;            // head dim
constexpr int kBlockM = 128;       // rows of the fixed operand (TMEM lanes)
constexpr int kBlockN = 64;        // rows of a streamed tile
constexpr int kHalfN = kBlockN / 2;
constexpr int kStages = 3;
constexpr int kTileBytes = kBlockN * kD * 2;        // 32 KB
constexpr int kChunkBytes = kBlockN * 128;          // one 64-column slab of a tile: 8 KB
constexpr int kNumSoftmaxWarps = 8;
constexpr int kNumSoftmaxThreads = kNumSoftmaxWarps * 32;
constexpr int kProducerWarp = 8;
constexpr int kMmaWarp = 9;
constexpr int kThreads = 320;

// tensor-memory column map (512 columns allocated)
constexpr uint32_t kColAcc = 0;      // 256 fp32 columns: ACC (O / dV)
constexpr uint32_t kColA = 256;      // 128 columns: fixed operand A, bf16 pairs
constexpr uint32_t kColS0 = 384;     // 64 fp32 columns: scores buffer 0
constexpr uint32_t kColS1 = 448;     // scores buffer 1
// Inside a scores buffer the softmax warp that owns columns [32h, 32h+32) overwrites ITS OWN first 16
// columns with the 32 bf16 probabilities (16 packed columns): P lives at +0..15 (keys 0-31) and
// +32..47 (keys 32-63).  k-step ks (16 keys = 8 packed columns) of the accumulate GEMM reads:
__device__ __forceinline__ uint32_t p_col_of_kstep(int ks) { return (ks < 2) ? ks * 8 : 32 + (ks - 2) * 8; }

enum { MODE_FWD = 0, MODE_DV = 1 };

// Optional per-CTA phase timeline (debugging / profiling aid; nullptr in production): 8 x u64 per CTA =
// {smid, t_entry, t_setup_done, t_operand_in_tmem, t_first_scores, t_loop_done, t_epilogue_done, nt}, %globaltimer ns.
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ unsigned smid() {
  unsigned r;
  asm volatile("mov.u32 %0, %smid;" : "=r"(r));
  return r;
}
#define SAM2B200_STAMP(buf, slot)                                                                     \
  do {                                                                                                \
    if ((buf) != nullptr && threadIdx.x == 0)                                                         \
      (buf)[((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 8 + (slot)] = gtimer(); \
  } while (0)

// Optional fused epilogue for gradient outputs: conjugate axial rotation + bf16 store with a row stride.
struct GradOut {
  int is_bf16;                 // element type of the output tensor map (bf16 or fp32)
  float* bias_grad;            // [256] fp32 or nullptr: column sums of the (rotated-back, scaled) gradient rows are
                               // ADDED here with red.global.add (the bias gradient of the projection that made
                               // this operand) -- saves a separate pass over the gradient tensor
  const float2* rope_table;    // [period, 128] (cos, sin) or nullptr = no rotation
  int rope_rows;               // rows [0, rope_rows) of every batch item are un-rotated (conjugate)
  int rope_period;             // table row = row % rope_period
  int rope_w;                  // > 0: AXIAL table of a rope_w x rope_w grid (period = rope_w^2): the first 64 (cos, sin) pairs of
                               // row p depend only on x = p mod w, the last 64 only on y = p div w, so they are read from rows x and
                               // y * w -- the 128 rows of a CTA then touch w + 128 / w table rows (16 KB at w = 24, L1-resident)
                               // instead of 128 (128 KB from L2, ~1.5 us per CTA); 0 = address the full row
  int rope_smem;               // > 0 (bytes, set by the host when rope_w > 0 and the kernel's shared memory has room): the CTA stages
                               // those w + 128 / w table rows in shared memory while its operands are in flight (rope_stage), and
                               // the epilogue reads them from there.  The L2 round trip per 32-column chunk was the largest single
                               // item of the gradient epilogue (profiles/r2_timeline_epilogue.txt: 3.6 us with rotation, 2.0 without).
};

// Shared-memory copy of the axial rotation table for the kBlockM rows of a CTA:
//   X part  rope_w rows x 64 (cos, sin) pairs, row stride 528 B (lanes read consecutive rows: conflict-free 16-byte loads)
//   Y part  one 512 B row per "slot" = run of rows with the same y: slot(r) = (r + x0) / w, x0 = (first row of the CTA) mod w
constexpr int kRopeXStride = 528;
__host__ __device__ constexpr int rope_y_bytes(int w) { return ((128 - 1 + w - 1) / w + 1) * 512; }
__host__ __device__ constexpr int rope_smem_bytes(int w) { return w * kRopeXStride + rope_y_bytes(w); }   // persistent kernels: + one more Y part

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// X part: the same for every CTA / item.  tid in [0, 256).
__device__ __forceinline__ void rope_stage_x(const GradOut& g, uint32_t area, int tid) {
  const int w = g.rope_w;
  for (int idx = tid; idx < w * 32; idx += 256)
    cp_async16(area + (idx >> 5) * kRopeXStride + (idx & 31) * 16, g.rope_table + (size_t)(idx >> 5) * 128 + (idx & 31) * 2);
}
// Y part for the 128 rows starting at row `cta_row0` of a batch item, into `ybase` (up to 128 / w + 2 rows of 512 B).
__device__ __forceinline__ void rope_stage_y(const GradOut& g, uint32_t ybase, int cta_row0, int tid) {
  const int w = g.rope_w;
  const int pos0 = cta_row0 % g.rope_period;
  const int x0 = pos0 % w, y0 = pos0 / w;
  const int ny = (kBlockM - 1 + x0) / w + 1;
  for (int idx = tid; idx < ny * 32; idx += 256) {
    const int yr = (y0 + (idx >> 5)) % w;
    cp_async16(ybase + (idx >> 5) * 512 + (idx & 31) * 16, g.rope_table + (size_t)(yr * w) * 128 + 64 + (idx & 31) * 2);
  }
}
// This thread's 64 pairs (row `row` of the CTA, columns [128 half, 128 half + 128)): shared-memory address of the first one.
__device__ __forceinline__ uint32_t rope_tab_addr(const GradOut& g, uint32_t area, uint32_t ybase, int cta_row0, int row, int half) {
  const int w = g.rope_w;
  const int x0 = (cta_row0 % g.rope_period) % w;
  return half == 0 ? area + ((x0 + row) % w) * kRopeXStride : ybase + ((x0 + row) / w) * 512;
}
__device__ __forceinline__ void load_table_chunk_smem(uint32_t tab, bool rotate, int chunk /*0..3 within the half*/, float2* t) {
  if (rotate) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint4 f = lds128(tab + chunk * 128 + i * 16);
      t[2 * i] = make_float2(__uint_as_float(f.x), __uint_as_float(f.y));
      t[2 * i + 1] = make_float2(__uint_as_float(f.z), __uint_as_float(f.w));
    }
  }
}

struct TwoGemmParams {
  int La;                      // rows of a per batch (N in FWD, M in DV)
  int Lx;                      // streamed length per batch (M in FWD, N in DV)
  float scale_log2;            // softmax scale * log2(e)
  // FWD outputs
  int has_out_f32;             // FWD, nsplit == 1: also store the fp32 copy of out (map_o32; kept for Delta = rowsum(dO o O))
  float* lse2;                 // [B, La] log2-domain LSE (FWD: output; DV: input, length Lx)
  float* part_acc;             // [nsplit, B, La, 256] fp32 un-normalised partials (nsplit > 1)
  float* part_ml;              // [nsplit, B, La, 2]  (m_ref * c, l)
  GradOut gout;                // DV output (dV)
  sam2b200::Dropout drop;      // attention-probability dropout (transformer.py:304-306); element index (b N + q) M + key
  int tiles_per_split;
  unsigned long long* dbg;     // optional timeline buffer
  int n_items, n_atiles;       // persistent kernels: items = (a block, batch) pairs, a blocks per batch
  void* out_small;             // DY = 64 forward: out' [B, La, 64] bf16
  float* out_small_f32;        //                  and its fp32 copy (optional)
  float* rowsum_drop;          // DY = 64 forward with dropout: [B, La] row sums of the dropped, re-scaled probabilities
                               // (the factor of the value bias: out = out' Wv^T + rowsum bv)
  // PROJ (fused output projection, transformer.py:308-309): proj = O . W^T + proj_bias (+ rowsum * proj_rank1), bf16 [B, La, 256]
  const float* proj_bias;      // [256] fp32
  const float* proj_rank1;     // [256] fp32 or nullptr (DY = 64 with dropout: Wo bv, multiplied by the row sum)
};

struct SharedStorage {
  alignas(1024) uint8_t x_tiles[kStages][kTileBytes];
  alignas(1024) uint8_t y_tiles[kStages][kTileBytes];
  alignas(8) uint64_t x_full[kStages];
  uint64_t x_empty[kStages];
  uint64_t y_full[kStages];
  uint64_t y_empty[kStages];
  uint64_t s_full[2];
  uint64_t p_ready[2];
  uint64_t acc_done;
  uint64_t a_full;             // TMA: fixed operand landed in ring stage kStages-1
  uint64_t a_ready;            // softmax warps: operand moved to tensor memory (stage kStages-1 is free again)
  float colvec[2][kBlockN];    // DV: LSE2 of the tile's columns
  float xchg[2][2][kBlockM];   // [buffer][half][row]: row-max exchange between the two halves of a row
  float lsum[2][kBlockM];      // [half][row]: row-sum exchange in the epilogue
  uint64_t loop_done;          // PROJ: every MMA of the main loop has completed (the ring is free for the projection weight)
  uint64_t w_full;             // PROJ: projection weight landed in shared memory
  uint64_t o_ready;            // PROJ: the normalised bf16 output tile is in tensor memory (A operand of the projection)
  uint64_t proj_done;          // PROJ: projection MMAs completed
  uint32_t tmem_base;
};
// PROJ tensor-memory columns.  DY = 64: the bf16 output operand goes to the unused columns behind the 64-column accumulator,
// the [128 x 256] projection result over the (then dead) Q operand + score buffers.  DY = 256: the bf16 output operand replaces
// the Q operand, the result overwrites the accumulator once every warp has read it.
constexpr uint32_t kColProjA64 = 64, kColProjD64 = 256;
constexpr int kProjW64Offset = 8192;      // DY = 64: the weight halves live behind the 8 KB a Y stage uses

// The fixed operand is staged by TMA as four [128 rows x 128 B] slabs (64 features each, 128-byte swizzle).
// Each of the two warps of a lane quarter moves HALF of its rows' 256 features (slabs 2*half, 2*half+1, which
// sit in `region`): 16-byte shared loads (conflict-free: the swizzle spreads 8 consecutive rows over all banks)
// -> registers -> tcgen05.st as packed bf16 pairs.
constexpr int kSlabBytes = kBlockM * 128;           // 16 KB
__device__ __forceinline__ void stage_to_tmem_half(uint32_t region, int row, uint32_t taddr, int half) {
#pragma unroll
  for (int cc = 0; cc < 2; ++cc) {
    const uint32_t base = region + cc * kSlabBytes + row * 128;
    uint32_t r[32];
#pragma unroll
    for (int v = 0; v < 8; ++v) {
      const uint4 u = lds128(base + ((v ^ (row & 7)) << 4));
      r[4 * v + 0] = u.x; r[4 * v + 1] = u.y; r[4 * v + 2] = u.z; r[4 * v + 3] = u.w;
    }
    SAM2B200_TMEM_ST32(taddr + (half * 2 + cc) * 32, r);
  }
  tmem_wait_st();
}

__device__ __forceinline__ void pair_barrier(int quarter) {
  asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");
}

// ---- TMA-store epilogue staging.  A warp owns 32 accumulator rows x 128 columns.  bf16: two boxes of
// [32 rows x 64 cols] (4 KB each); fp32: four boxes of [32 rows x 32 cols] (4 KB each); a box row is 128 B and
// its 16-byte chunks are XOR-swizzled with (row & 7), the layout CU_TENSOR_MAP_SWIZZLE_128B expects.
constexpr int kBoxBytes = 32 * 128;
__device__ __forceinline__ void stage_chunk_bf16(uint32_t stage, int lane, int i /*chunk 0..3 of the warp's 128 cols*/,
                                                 const float* v) {
  const uint32_t base = stage + (i >> 1) * kBoxBytes + lane * 128;
#pragma unroll
  for (int q4 = 0; q4 < 4; ++q4)
    sts128(base + ((((i & 1) * 4 + q4) ^ (lane & 7)) << 4), pack_bf16(v[8 * q4 + 0], v[8 * q4 + 1]),
           pack_bf16(v[8 * q4 + 2], v[8 * q4 + 3]), pack_bf16(v[8 * q4 + 4], v[8 * q4 + 5]),
           pack_bf16(v[8 * q4 + 6], v[8 * q4 + 7]));
}
__device__ __forceinline__ void stage_chunk_f32(uint32_t stage, int lane, int i, const float* v) {
  const uint32_t base = stage + i * kBoxBytes + lane * 128;
#pragma unroll
  for (int q4 = 0; q4 < 8; ++q4)
    sts128(base + ((q4 ^ (lane & 7)) << 4), __float_as_uint(v[4 * q4]), __float_as_uint(v[4 * q4 + 1]),
           __float_as_uint(v[4 * q4 + 2]), __float_as_uint(v[4 * q4 + 3]));
}
// Make the warp's staged box visible to the async proxy and let one lane hand it to the TMA unit.
__device__ __forceinline__ void store_box(const CUtensorMap* map, uint32_t box_smem, int lane, int c0, int row0, int b,
                                          bool in_range) {
  fence_proxy_async();
  __syncwarp();
  if (lane == 0 && in_range) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 :
                 : "l"(reinterpret_cast<uint64_t>(map)), "r"(box_smem), "r"(c0), "r"(row0), "r"(b)
                 : "memory");
    tma_store_commit();
  }
}

// Gradient epilogue of one warp (32 rows x 128 columns = 4 chunks of 32 columns): TMEM -> registers -> scale ->
// conjugate axial rotation -> shared-memory staging -> TMA store.  The (cos, sin) table rows are L2-resident but
// ~1 us away: the caller issues the loads of chunk 0 BEFORE waiting for the last MMA (table_prefetch), the loads
// for chunk i+1 are issued before chunk i is processed, and the TMEM load of chunk i+1 is in flight meanwhile.
__device__ __forceinline__ void load_table_chunk(const GradOut& g, bool rotate, int row_in_batch, int col0, float2* t) {
  if (rotate) {
    const int pos = row_in_batch % g.rope_period;
    const int trow = g.rope_w > 0 ? (col0 < 128 ? pos % g.rope_w : pos - pos % g.rope_w) : pos;
    const float4* src = reinterpret_cast<const float4*>(g.rope_table + (long long)trow * 128 + (col0 >> 1));
#pragma unroll
    for (int i = 0; i < 8; ++i) { const float4 f = __ldg(src + i); t[2 * i] = make_float2(f.x, f.y); t[2 * i + 1] = make_float2(f.z, f.w); }
  }
}

// Column sums of a [32 rows (lanes) x 32 columns (v)] block: a 5-step butterfly in which every lane keeps the half
// of the columns its lane bit selects; lane L ends up with the sum of column L.  31 shuffles per block.
__device__ __forceinline__ float warp_column_sum32(const float* v, int lane) {
  float a[16], b[8], c[4], d[2];
  {
    const bool hi = lane & 16;
#pragma unroll
    for (int i = 0; i < 16; ++i) { const float keep = hi ? v[i + 16] : v[i], send = hi ? v[i] : v[i + 16]; a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16); }
  }
  {
    const bool hi = lane & 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) { const float keep = hi ? a[i + 8] : a[i], send = hi ? a[i] : a[i + 8]; b[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8); }
  }
  {
    const bool hi = lane & 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float keep = hi ? b[i + 4] : b[i], send = hi ? b[i] : b[i + 4]; c[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4); }
  }
  {
    const bool hi = lane & 2;
#pragma unroll
    for (int i = 0; i < 2; ++i) { const float keep = hi ? c[i + 2] : c[i], send = hi ? c[i] : c[i + 2]; d[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2); }
  }
  const bool hi = lane & 1;
  const float keep = hi ? d[1] : d[0], send = hi ? d[0] : d[1];
  return keep + __shfl_xor_sync(0xffffffffu, send, 1);
}

// acc_addr: TMEM address of column 0 of this thread's accumulator row; half selects columns [128h, 128h+128).
// stage: this warp's staging area (8 KB bf16 / 16 KB fp32); row0 = first row (in the batch item) of the warp.
__device__ __forceinline__ void grad_epilogue(const GradOut& g, const CUtensorMap* map, uint32_t stage, uint32_t acc_addr,
                                              int half, int lane, int row0, int La, int b, float scale, bool rotate,
                                              float2* tcur /* chunk 0 of the table, already loaded (tab == 0) */,
                                              unsigned long long* dbg = nullptr /* SAM2B200_EPI_TIMELINE builds only */,
                                              uint32_t tab = 0 /* rope_tab_addr: the staged table, 0 = read the global table */) {
  const int row_in_batch = row0 + lane;
  if (tab != 0) load_table_chunk_smem(tab, rotate, 0, tcur);
  float2 tnext[16];
  uint32_t ocur[32], onext[32];
  const int c0 = half * 4;
  SAM2B200_TMEM_LD32(acc_addr + c0 * 32, ocur);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int cc = c0 + i;
    tmem_wait_ld();
    if (i + 1 < 4) {
      SAM2B200_TMEM_LD32(acc_addr + (cc + 1) * 32, onext);
      if (tab != 0) load_table_chunk_smem(tab, rotate, i + 1, tnext);
      else load_table_chunk(g, rotate, row_in_batch, (cc + 1) * 32, tnext);
    }
    float v[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) v[k] = __uint_as_float(ocur[k]) * scale;
    if (rotate) {
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const float2 cs = tcur[k];
        const float re = v[2 * k] * cs.x + v[2 * k + 1] * cs.y;      // multiply by conj(cos + i sin)
        const float im = v[2 * k + 1] * cs.x - v[2 * k] * cs.y;
        v[2 * k] = re; v[2 * k + 1] = im;
      }
    }
    if (g.bias_grad != nullptr) {
      if (row_in_batch >= La) {
#pragma unroll
        for (int k = 0; k < 32; ++k) v[k] = 0.f;        // rows beyond the tensor are clipped by the store, not by the sum
      }
      const float cs = warp_column_sum32(v, lane);
      atomicAdd(g.bias_grad + cc * 32 + lane, cs);
    }
    if (g.is_bf16) {
      stage_chunk_bf16(stage, lane, i, v);
      if (i & 1) store_box(map, stage + (i >> 1) * kBoxBytes, lane, half * 128 + (i >> 1) * 64, row0, b, row0 < La);
    } else {
      stage_chunk_f32(stage, lane, i, v);
      store_box(map, stage + i * kBoxBytes, lane, half * 128 + i * 32, row0, b, row0 < La);
    }
    if (i + 1 < 4) {
#pragma unroll
      for (int k = 0; k < 32; ++k) ocur[k] = onext[k];
#pragma unroll
      for (int k = 0; k < 16; ++k) tcur[k] = tnext[k];
    }
#ifdef SAM2B200_EPI_TIMELINE
    if (i == 0) SAM2B200_STAMP(dbg, 2);
    if (i == 1) SAM2B200_STAMP(dbg, 3);
    if (i == 3) SAM2B200_STAMP(dbg, 4);
#endif
  }
  if (lane == 0) tma_store_wait_read();
#ifdef SAM2B200_EPI_TIMELINE
  SAM2B200_STAMP(dbg, 7);
#endif
}

// PROJ epilogue, phase 2: this warp's 32 rows x 128 columns of the projection result (TMEM) + bias (+ rowsum * rank-1 vector)
// -> bf16 -> staging (two 64-column boxes) -> TMA store.
__device__ __forceinline__ void proj_store(const TwoGemmParams& p, const CUtensorMap* map_p, uint32_t stage, uint32_t d_addr /*column 0 of this half*/,
                                           int half, int lane, int row0, int b, float rowsum) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint32_t o[32];
    SAM2B200_TMEM_LD32(d_addr + i * 32, o);
    tmem_wait_ld();
    const int col0 = half * 128 + i * 32;
    float v[32];
    const float4* bp = reinterpret_cast<const float4*>(p.proj_bias + col0);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float4 bq = __ldg(bp + k);
      v[4 * k] = __uint_as_float(o[4 * k]) + bq.x; v[4 * k + 1] = __uint_as_float(o[4 * k + 1]) + bq.y;
      v[4 * k + 2] = __uint_as_float(o[4 * k + 2]) + bq.z; v[4 * k + 3] = __uint_as_float(o[4 * k + 3]) + bq.w;
    }
    if (p.proj_rank1 != nullptr) {
      const float4* rp = reinterpret_cast<const float4*>(p.proj_rank1 + col0);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float4 rq = __ldg(rp + k);
        v[4 * k] = fmaf(rowsum, rq.x, v[4 * k]); v[4 * k + 1] = fmaf(rowsum, rq.y, v[4 * k + 1]);
        v[4 * k + 2] = fmaf(rowsum, rq.z, v[4 * k + 2]); v[4 * k + 3] = fmaf(rowsum, rq.w, v[4 * k + 3]);
      }
    }
    stage_chunk_bf16(stage, lane, i, v);
    if (i & 1) store_box(map_p, stage + (i >> 1) * kBoxBytes, lane, half * 128 + (i >> 1) * 64, row0, b, row0 < p.La);
  }
  if (lane == 0) tma_store_wait_read();
}

// DROP: attention-probability dropout compiled in or out (a run-time branch in the softmax loops costs registers and
// scheduling freedom even when never taken -- measured on the pair kernel).
// DY: width of the streamed Y operand and of the accumulator.  256 = the projected values; 64 (forward only) = the raw
// 64-d memory features of the cross-attention: out' = softmax(.) mem, the value projection is applied to the [N, 64] result
// afterwards (softmax rows sum to 1, so out = out' Wv^T + bv exactly) -- 4x fewer PV FLOPs and no [B, M, 256] V tensor.
// PROJ (forward, no split-KV): the OUTPUT PROJECTION of the attention module (transformer.py:308-309; for the raw-memory
// cross-attention the folded Wo Wv) runs in this kernel's epilogue: the normalised output tile goes back to tensor memory as
// bf16, one more tcgen05 GEMM against the projection weight staged in the idle tile ring, bias added on the fp32
// accumulator, the projected [128 x 256] tile leaves through map_p.  The un-projected output is still written (the backward
// needs it for the weight gradient and Delta).
template <int MODE, bool DROP, int DY = kD, bool PROJ = false>
__global__ void __launch_bounds__(kThreads, 1)
two_gemm_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_y,
                const __grid_constant__ CUtensorMap map_a,    // fixed operand [B, La, 256] bf16, box 64 x 128
                const __grid_constant__ CUtensorMap map_o,    // FWD: out bf16 (box 64 x 32); DV: dV (bf16 64 x 32 | fp32 32 x 32)
                const __grid_constant__ CUtensorMap map_o32,  // FWD: fp32 copy of out (box 32 x 32); PROJ: the weight [256, DY], box 64 x 128
                const __grid_constant__ CUtensorMap map_p,    // PROJ: projected output [B, La, 256] bf16, box 64 x 32
                const TwoGemmParams p) {
  static_assert(DY == kD || (DY == 64 && MODE == MODE_FWD), "narrow Y: forward only");
  static_assert(!PROJ || MODE == MODE_FWD, "fused projection: forward only");
  extern __shared__ uint8_t smem_raw[];
  SharedStorage& sh = *reinterpret_cast<SharedStorage*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int a_tile = blockIdx.x;       // which 128-row block of the fixed operand
  const int b = blockIdx.y;            // batch (object)
  const int split = blockIdx.z;
  const int nsplit = gridDim.z;

  const int total_tiles = (p.Lx + kBlockN - 1) / kBlockN;
  const int t_begin = split * p.tiles_per_split;
  const int t_end = min(total_tiles, t_begin + p.tiles_per_split);
  const int nt = t_end - t_begin;      // >= 1 by construction of the grid
  if (p.dbg != nullptr && threadIdx.x == 0) {
    unsigned long long* e = p.dbg + ((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 8;
    e[0] = smid(); e[7] = nt;
  }
  SAM2B200_STAMP(p.dbg, 1);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&sh.x_full[s], 1); mbar_init(&sh.x_empty[s], 1);
      mbar_init(&sh.y_full[s], 1); mbar_init(&sh.y_empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) { mbar_init(&sh.s_full[i], 1); mbar_init(&sh.p_ready[i], kNumSoftmaxThreads); }
    mbar_init(&sh.acc_done, 1);
    mbar_init(&sh.a_full, 1);
    mbar_init(&sh.a_ready, kNumSoftmaxThreads);
    if (PROJ) {
      mbar_init(&sh.loop_done, 1); mbar_init(&sh.w_full, 1); mbar_init(&sh.o_ready, kNumSoftmaxThreads); mbar_init(&sh.proj_done, 1);
    }
    fence_barrier_init();
  }
  if (warp == kProducerWarp && lane == 0) { prefetch_tmap(&map_a); prefetch_tmap(&map_x); prefetch_tmap(&map_y); if (PROJ) prefetch_tmap(&map_o32); }
  if (warp == 0 && lane == 0) { prefetch_tmap(&map_o); if (MODE == MODE_FWD) prefetch_tmap(&map_o32); if (PROJ) prefetch_tmap(&map_p); }
  if (warp == kMmaWarp) { tmem_alloc(&sh.tmem_base, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sh.tmem_base;
  SAM2B200_STAMP(p.dbg, 2);

  if (warp == kProducerWarp) {
    // ===================== TMA producer (converged warp, one elected issuer) =====================
    const bool leader = elect_one();
    if (leader) {   // fixed operand -> the last ring stage (slabs 0,1 in its X buffer, slabs 2,3 in its Y buffer)
      mbar_arrive_expect_tx(&sh.a_full, 4 * kSlabBytes);
#pragma unroll
      for (int c = 0; c < 4; ++c)
        tma_load_3d((c < 2 ? &sh.x_tiles[kStages - 1][0] : &sh.y_tiles[kStages - 1][0]) + (c & 1) * kSlabBytes, &map_a,
                    &sh.a_full, c * 64, a_tile * kBlockM, b);
      if (PROJ && DY == 64) {   // folded weight [256, 64]: two [128 x 128 B] halves behind the 8 KB the narrow Y tiles use (never touched by the ring)
        mbar_arrive_expect_tx(&sh.w_full, 2 * kSlabBytes);
        tma_load_3d(&sh.y_tiles[0][kProjW64Offset], &map_o32, &sh.w_full, 0, 0, 0);
        tma_load_3d(&sh.y_tiles[1][kProjW64Offset], &map_o32, &sh.w_full, 0, kBlockM, 0);
      }
    }
    __syncwarp();
    for (int j = 0; j < nt; ++j) {
      const int s = j % kStages;
      const uint32_t ph = (j / kStages) & 1;
      const int row0 = (t_begin + j) * kBlockN;
      if (j == kStages - 1) mbar_wait(&sh.a_ready, 0);   // first use of the stage that staged the fixed operand
      mbar_wait(&sh.x_empty[s], ph ^ 1);
      if (leader) {
        mbar_arrive_expect_tx(&sh.x_full[s], kTileBytes);
#pragma unroll
        for (int c = 0; c < 4; ++c)
          tma_load_3d(&sh.x_tiles[s][c * kChunkBytes], &map_x, &sh.x_full[s], c * 64, row0, b);
      }
      __syncwarp();
      mbar_wait(&sh.y_empty[s], ph ^ 1);
      if (leader) {
        mbar_arrive_expect_tx(&sh.y_full[s], kBlockN * DY * 2);
#pragma unroll
        for (int c = 0; c < DY / 64; ++c)
          tma_load_3d(&sh.y_tiles[s][c * kChunkBytes], &map_y, &sh.y_full[s], c * 64, row0, b);
      }
      __syncwarp();
    }
    if (PROJ && DY == kD) {
      // Wo [256, 256]: four K slabs of [256 rows x 128 B] = 32 KB each, into Y stages 0..2 and X stage 2 once the whole ring is idle
      // (X stages 0, 1 are the epilogue's staging area); overlaps the normalisation / store of the un-projected output
      mbar_wait(&sh.loop_done, 0);
      if (leader) {
        mbar_arrive_expect_tx(&sh.w_full, 8 * kSlabBytes);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint8_t* dst = (c < 3) ? &sh.y_tiles[c][0] : &sh.x_tiles[2][0];
          tma_load_3d(dst, &map_o32, &sh.w_full, c * 64, 0, 0);
          tma_load_3d(dst + kSlabBytes, &map_o32, &sh.w_full, c * 64, kBlockM, 0);
        }
      }
      __syncwarp();
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer =====================
    // The whole warp stays converged (all lanes poll the mbarriers); ONE elected lane issues.  Descriptors are
    // `base + compile-time offset` (one uniform add per tcgen05.mma): with a rebuilt descriptor and a
    // lane==0 branch the issue loop cost ~87 cycles per MMA and starved the tensor pipe (see
    // profiles/r1_mma_probe.txt); like this a 128x64x16 MMA issues every ~37 cycles.
    const bool leader = elect_one();
    constexpr uint32_t idesc_s = make_idesc_bf16(kBlockM, kBlockN, 0, 0);   // A(TMEM) . X^T, X K-major
    constexpr uint32_t idesc_acc = make_idesc_bf16(kBlockM, DY, 0, 1);      // P(TMEM) . Y,  Y MN-major
    const uint32_t x_lo0 = desc_lo_sw128(smem_u32(&sh.x_tiles[0][0]), 16);           // K-major: LBO unused
    const uint32_t y_lo0 = desc_lo_sw128(smem_u32(&sh.y_tiles[0][0]), kChunkBytes);  // MN-major: LBO = slab stride
    auto issue_scores = [&](int t) {
      const int s = t % kStages;
      mbar_wait(&sh.x_full[s], (t / kStages) & 1);
      tc_fence_after();
      if (leader) {
        const uint32_t xlo = x_lo0 + s * (kTileBytes >> 4);
        const uint32_t d = tmem + ((t & 1) ? kColS1 : kColS0);
#pragma unroll
        for (int ks = 0; ks < kD / 16; ++ks)   // K-major SW128: slab ks/4, 32 B per k-step inside the 128 B row
          umma_ts_lohi(d, tmem + kColA + ks * 8, xlo + (ks >> 2) * (kChunkBytes >> 4) + (ks & 3) * 2,
                       kDescHiSw128_1024, idesc_s, ks > 0);
        umma_commit(&sh.x_empty[s]);
        umma_commit(&sh.s_full[t & 1]);
      }
      __syncwarp();
    };
    mbar_wait(&sh.a_ready, 0);
    tc_fence_after();
    issue_scores(0);
    if (nt > 1) issue_scores(1);
    for (int j = 0; j < nt; ++j) {
      const int s = j % kStages;
      mbar_wait(&sh.p_ready[j & 1], (j >> 1) & 1);
      mbar_wait(&sh.y_full[s], (j / kStages) & 1);
      tc_fence_after();
      if (leader) {
        const uint32_t ylo = y_lo0 + s * (kTileBytes >> 4);
        const uint32_t pa = tmem + ((j & 1) ? kColS1 : kColS0);
#pragma unroll
        for (int ks = 0; ks < kBlockN / 16; ++ks)   // MN-major SW128: 16 rows = 2 groups of 8 rows = 2048 B per k-step
          umma_ts_lohi(tmem + kColAcc, pa + p_col_of_kstep(ks), ylo + ks * (2048 >> 4), kDescHiSw128_1024, idesc_acc,
                       (j > 0) || (ks > 0));
        umma_commit(&sh.y_empty[s]);
        umma_commit(&sh.acc_done);
        if (PROJ && j + 1 >= nt) umma_commit(&sh.loop_done);
      }
      __syncwarp();
      if (j + 2 < nt) issue_scores(j + 2);
    }
    if (PROJ) {
      // ---- output projection: D[128 x 256] = O_bf16 (TMEM) . W^T (shared memory, K-major)
      mbar_wait(&sh.w_full, 0);
      mbar_wait(&sh.o_ready, 0);
      tc_fence_after();
      if (leader) {
        if (DY == 64) {
          constexpr uint32_t idesc_p = make_idesc_bf16(kBlockM, 128, 0, 0);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const uint32_t wlo = desc_lo_sw128(smem_u32(&sh.y_tiles[h][kProjW64Offset]), 16);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              umma_ts_lohi(tmem + kColProjD64 + h * 128, tmem + kColProjA64 + ks * 8, wlo + ks * 2, kDescHiSw128_1024, idesc_p, ks > 0);
          }
        } else {
          constexpr uint32_t idesc_p = make_idesc_bf16(kBlockM, kD, 0, 0);
#pragma unroll
          for (int ks = 0; ks < kD / 16; ++ks) {
            const int c = ks >> 2;
            const uint32_t wlo = desc_lo_sw128(smem_u32((c < 3) ? &sh.y_tiles[c][0] : &sh.x_tiles[2][0]), 16);
            umma_ts_lohi(tmem + kColAcc, tmem + kColA + ks * 8, wlo + (ks & 3) * 2, kDescHiSw128_1024, idesc_p, ks > 0);
          }
        }
        umma_commit(&sh.proj_done);
      }
      __syncwarp();
    }
  } else {
    // ===================== softmax / epilogue warps (0..7) =====================
    const int quarter = warp & 3;                        // TMEM lane quarter this warp may access
    const int half = warp >> 2;                          // which 32 of the tile's 64 columns
    const int row = quarter * 32 + lane;                 // TMEM lane == accumulator row
    const uint32_t lane_addr = tmem + (uint32_t(quarter * 32) << 16);
    const long long a_row_idx = (long long)a_tile * kBlockM + row;
    const bool row_valid = a_row_idx < p.La;
    {
      mbar_wait(&sh.a_full, 0);
      stage_to_tmem_half(smem_u32(half ? &sh.y_tiles[kStages - 1][0] : &sh.x_tiles[kStages - 1][0]), row, lane_addr + kColA, half);
      tc_fence_before();
      mbar_arrive(&sh.a_ready);
      SAM2B200_STAMP(p.dbg, 3);
    }
    // this warp's epilogue staging area inside the (then idle) tile ring, and its first row in the batch item
    const int row0 = a_tile * kBlockM + quarter * 32;
    // (PROJ at full width: only X stages 0, 1 -- 8 KB per warp -- the rest of the ring receives the projection weight)
    const uint32_t stage = smem_u32(&sh.x_tiles[0][0]) +
                           warp * ((PROJ && DY == kD) ? 2 * kBoxBytes : ((MODE == MODE_FWD) ? 6 * kBoxBytes : 4 * kBoxBytes));
    const bool rotate = false;   // dV is never rotated
    float2 tcur[16];

    const float c = p.scale_log2;
    constexpr bool drop_on = DROP;
    const uint32_t drop_key = drop_on ? sam2b200::dropout_key(*p.drop.seed, p.drop.site) : 0u;
    float m_ref = -INFINITY;   // running (lazily updated) row max of the raw scores -- identical in both halves
    float l = 0.f;             // this half's running sum of exp2((s - m_ref) c)
    float lk = 0.f;            // DY = 64 with dropout: the same sum over the KEPT probabilities only

    // DV: per-column LSE2 (columns are queries) is staged through shared memory ONE TILE AHEAD, so the
    // global-load latency is off the per-tile critical path.
    float cv_next = INFINITY;
    if (MODE == MODE_DV && threadIdx.x < kBlockN) {
      const int col = t_begin * kBlockN + threadIdx.x;
      cv_next = (col < p.Lx) ? p.lse2[(long long)b * p.Lx + col] : INFINITY;
    }
    for (int j = 0; j < nt; ++j) {
      const int t = t_begin + j;
      const uint32_t sbuf = lane_addr + ((j & 1) ? kColS1 : kColS0);
      if (MODE == MODE_DV) {
        if (threadIdx.x < kBlockN) {
          sh.colvec[j & 1][threadIdx.x] = cv_next;
          const int col = (t + 1) * kBlockN + threadIdx.x;
          cv_next = (j + 1 < nt && col < p.Lx) ? p.lse2[(long long)b * p.Lx + col] : INFINITY;
        }
        asm volatile("bar.sync 5, 256;" ::: "memory");
      }
      mbar_wait(&sh.s_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      if (j == 0) SAM2B200_STAMP(p.dbg, 4);
      uint32_t r0[32];
      SAM2B200_TMEM_LD32(sbuf + half * kHalfN, r0);
      tmem_wait_ld();
      float sv[kHalfN];
#pragma unroll
      for (int i = 0; i < kHalfN; ++i) sv[i] = __uint_as_float(r0[i]);

      uint32_t pk[16];
      bool acc_synced = false;
      if (MODE == MODE_FWD) {
        const int ncols = p.Lx - t * kBlockN - half * kHalfN;   // valid columns of this half-tile
        if (ncols < kHalfN) {
#pragma unroll
          for (int i = 0; i < kHalfN; ++i) if (i >= ncols) sv[i] = -INFINITY;
        }
        float mx = sv[0];
#pragma unroll
        for (int i = 1; i < kHalfN; ++i) mx = fmaxf(mx, sv[i]);
        // combine the two halves' row maxima (both warps of the pair must take the same decision)
        sh.xchg[j & 1][half][row] = mx;
        pair_barrier(quarter);
        mx = fmaxf(mx, sh.xchg[j & 1][half ^ 1][row]);
        // Lazy rescale: keep the stale reference max unless the true max grew by > 2^8 in the
        // exp2 domain (P then stays < 256, exact enough in bf16/fp32); first tile just adopts it.
        const bool grow = (mx - m_ref) * c > 8.0f;
        if (j == 0) {
          m_ref = mx;
        } else if (__any_sync(0xffffffffu, grow)) {
          const float m_new = grow ? mx : m_ref;
          const float f = ex2((m_ref - m_new) * c);      // 1.0 for rows that do not move
          mbar_wait(&sh.acc_done, (j - 1) & 1);          // PV[j-1] has landed in ACC
          acc_synced = true;
          tc_fence_after();
#pragma unroll 1
          for (int cc = half * (DY / 64); cc < (half + 1) * (DY / 64); ++cc) {   // each half rescales its half of the ACC columns
            uint32_t o[32];
            SAM2B200_TMEM_LD32(lane_addr + kColAcc + cc * 32, o);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
            SAM2B200_TMEM_ST32(lane_addr + kColAcc + cc * 32, o);
          }
          tmem_wait_st();
          l *= f;
          lk *= f;
          m_ref = m_new;
        }
        const float mc = m_ref * c;
        float sum0 = 0.f, sum1 = 0.f;
        // dropout acts on the normalised probabilities: the row sum l keeps every term, only the PV operand is masked
        const uint32_t didx = (uint32_t)(((long long)b * p.La + a_row_idx) * p.Lx) + (uint32_t)(t * kBlockN + half * kHalfN);
#pragma unroll
        for (int i = 0; i < kHalfN; i += 2) {
          float e0 = ex2(fmaf(sv[i], c, -mc));
          float e1 = ex2(fmaf(sv[i + 1], c, -mc));
          sum0 += e0; sum1 += e1;
          if (drop_on) {
            e0 = sam2b200::dropout_keep(drop_key, didx + i, p.drop.thresh) ? e0 : 0.f;
            e1 = sam2b200::dropout_keep(drop_key, didx + i + 1, p.drop.thresh) ? e1 : 0.f;
            if (DY == 64) lk += e0 + e1;
          }
          pk[i >> 1] = pack_bf16(e0, e1);
        }
        l += sum0 + sum1;
      } else {
        const float* cv = &sh.colvec[j & 1][half * kHalfN];
        // rows are keys, columns queries: index (b N + q) M + key advances by M per column
        const uint32_t didx = (uint32_t)(((long long)b * p.Lx + (t * kBlockN + half * kHalfN)) * p.La + a_row_idx);
#pragma unroll
        for (int i = 0; i < kHalfN; i += 2) {
          float e0 = ex2(fmaf(sv[i], c, -cv[i]));
          float e1 = ex2(fmaf(sv[i + 1], c, -cv[i + 1]));
          if (drop_on) {
            e0 = sam2b200::dropout_keep(drop_key, didx + (uint32_t)i * (uint32_t)p.La, p.drop.thresh) ? e0 : 0.f;
            e1 = sam2b200::dropout_keep(drop_key, didx + (uint32_t)(i + 1) * (uint32_t)p.La, p.drop.thresh) ? e1 : 0.f;
          }
          pk[i >> 1] = pack_bf16(e0, e1);
        }
      }
      SAM2B200_TMEM_ST16(sbuf + half * kHalfN, pk);       // P (bf16 pairs) over this warp's OWN first 16 columns
      tmem_wait_st();
      // Observe EVERY phase of acc_done, in order: a parity wait is only unambiguous while the
      // waiter is at most one phase behind.  PV[j-1] was issued right after QK[j], so by now it has
      // (almost always) completed and this wait is free; PV[j] cannot complete before our arrive.
      if (j > 0 && !acc_synced) mbar_wait(&sh.acc_done, (j - 1) & 1);
      tc_fence_before();
      mbar_arrive(&sh.p_ready[j & 1]);
    }

    // ---------------- epilogue: each half stores its 128 of the 256 output columns ----------------
    mbar_wait(&sh.acc_done, (nt - 1) & 1);
    tc_fence_after();
    SAM2B200_STAMP(p.dbg, 5);
    if (MODE == MODE_FWD) {
      // total row sum = sum of the two halves' partial sums
      sh.lsum[half][row] = l;
      pair_barrier(quarter);
      l += sh.lsum[half ^ 1][row];
      if (DY == 64) {
        // narrow accumulator: this warp holds 32 rows x 32 of the 64 output columns -> plain 16-byte global stores
        const float inv_l = p.drop.inv_keep / l;
        if (drop_on) {          // row sum of the dropped, re-scaled probabilities (both halves; xchg is idle by now)
          sh.xchg[0][half][row] = lk;
          pair_barrier(quarter);
          lk += sh.xchg[0][half ^ 1][row];
          if (row_valid && half == 0 && p.rowsum_drop != nullptr) p.rowsum_drop[(long long)b * p.La + a_row_idx] = lk * inv_l;
        }
        uint32_t o[32];
        SAM2B200_TMEM_LD32(lane_addr + kColAcc + half * 32, o);
        tmem_wait_ld();
        float v[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) v[k] = __uint_as_float(o[k]) * inv_l;
        if (PROJ) {   // the bf16 output tile (what out' holds) back to tensor memory: A operand of the projection
          uint32_t pk2[16];
#pragma unroll
          for (int k = 0; k < 16; ++k) pk2[k] = pack_bf16(v[2 * k], v[2 * k + 1]);
          SAM2B200_TMEM_ST16(lane_addr + kColProjA64 + half * 16, pk2);
          tmem_wait_st();
          tc_fence_before();
          mbar_arrive(&sh.o_ready);
        }
        if (row_valid) {
          const long long off = ((long long)b * p.La + a_row_idx) * 64 + half * 32;
          uint4* o16 = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out_small) + off);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            o16[k] = make_uint4(pack_bf16(v[8 * k], v[8 * k + 1]), pack_bf16(v[8 * k + 2], v[8 * k + 3]),
                                pack_bf16(v[8 * k + 4], v[8 * k + 5]), pack_bf16(v[8 * k + 6], v[8 * k + 7]));
          if (p.out_small_f32 != nullptr) {
            float4* o32 = reinterpret_cast<float4*>(p.out_small_f32 + off);
#pragma unroll
            for (int k = 0; k < 8; ++k) o32[k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
          }
          if (half == 0) p.lse2[(long long)b * p.La + a_row_idx] = fmaf(m_ref, c, log2f(l));
        }
        if (PROJ) {
          mbar_wait(&sh.proj_done, 0);
          tc_fence_after();
          proj_store(p, &map_p, stage, lane_addr + kColProjD64 + half * 128, half, lane, row0, b, drop_on ? lk * inv_l : 0.f);
        }
      } else if (nsplit == 1) {
        // out (bf16, two 64-column boxes) and optionally its fp32 copy (four 32-column boxes) leave through this
        // warp's staging area: [bf16 box 0 | bf16 box 1 | fp32 box 0..3]
        const float inv_l = p.drop.inv_keep / l;    // inverted dropout: kept probabilities are scaled by 1 / (1 - p)
        uint32_t ocur[32], onext[32];
        SAM2B200_TMEM_LD32(lane_addr + kColAcc + half * 128, ocur);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          tmem_wait_ld();
          if (i + 1 < 4) SAM2B200_TMEM_LD32(lane_addr + kColAcc + half * 128 + (i + 1) * 32, onext);
          float v[32];
#pragma unroll
          for (int k = 0; k < 32; ++k) v[k] = __uint_as_float(ocur[k]) * inv_l;
          stage_chunk_bf16(stage, lane, i, v);
          if (PROJ) {
            // fp32 copy straight from registers (128 contiguous bytes per thread); the bf16 tile back to tensor memory over the
            // dead Q operand: A operand of the projection
            if (p.out_small_f32 != nullptr && row_valid) {
              float4* o32p = reinterpret_cast<float4*>(p.out_small_f32 + ((long long)b * p.La + a_row_idx) * kD + half * 128 + i * 32);
#pragma unroll
              for (int k = 0; k < 8; ++k) o32p[k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
            }
            uint32_t pk2[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) pk2[k] = pack_bf16(v[2 * k], v[2 * k + 1]);
            SAM2B200_TMEM_ST16(lane_addr + kColA + half * 64 + i * 16, pk2);
          } else if (p.has_out_f32) {
            stage_chunk_f32(stage + 2 * kBoxBytes, lane, i, v);
          }
          if (i & 1) store_box(&map_o, stage + (i >> 1) * kBoxBytes, lane, half * 128 + (i >> 1) * 64, row0, b, row0 < p.La);
          if (!PROJ && p.has_out_f32) store_box(&map_o32, stage + (2 + i) * kBoxBytes, lane, half * 128 + i * 32, row0, b, row0 < p.La);
          if (i + 1 < 4) {
#pragma unroll
            for (int k = 0; k < 32; ++k) ocur[k] = onext[k];
          }
        }
        if (PROJ) {
          tmem_wait_st();
          tc_fence_before();
          mbar_arrive(&sh.o_ready);      // every warp has read its part of the accumulator: the projection may overwrite it
        }
        if (row_valid && half == 0) p.lse2[(long long)b * p.La + a_row_idx] = fmaf(m_ref, c, log2f(l));
        if (lane == 0) tma_store_wait_read();
        if (PROJ) {
          __syncwarp();
          mbar_wait(&sh.proj_done, 0);
          tc_fence_after();
          proj_store(p, &map_p, stage, lane_addr + kColAcc + half * 128, half, lane, row0, b, 0.f);
        }
      } else {
        const long long prow = ((long long)split * gridDim.y + b) * p.La + a_row_idx;
        float* orow = p.part_acc + prow * kD;
#pragma unroll 1
        for (int cc = half * 4; cc < half * 4 + 4; ++cc) {
          uint32_t o[32];
          SAM2B200_TMEM_LD32(lane_addr + kColAcc + cc * 32, o);
          tmem_wait_ld();
          if (row_valid) {
#pragma unroll
            for (int v = 0; v < 8; ++v)
              *reinterpret_cast<float4*>(orow + cc * 32 + v * 4) =
                  make_float4(__uint_as_float(o[4 * v]) * p.drop.inv_keep, __uint_as_float(o[4 * v + 1]) * p.drop.inv_keep,
                              __uint_as_float(o[4 * v + 2]) * p.drop.inv_keep, __uint_as_float(o[4 * v + 3]) * p.drop.inv_keep);
          }
        }
        if (row_valid && half == 0) { p.part_ml[prow * 2] = m_ref * c; p.part_ml[prow * 2 + 1] = l; }
      }
    } else {
      grad_epilogue(p.gout, &map_o, stage, lane_addr + kColAcc, half, lane, row0, p.La, b, p.drop.inv_keep, rotate, tcur);
    }
  }

  tc_fence_before();
  __syncthreads();
  SAM2B200_STAMP(p.dbg, 6);
  if (warp == kMmaWarp) tmem_dealloc(tmem, 512);
}

// =====================================================================================
// three_gemm_kernel: S = A1 . X^T, dP = A2 . Y^T, dS = P o (dP - Delta), ACC += dS . X
// =====================================================================================
enum { MODE_DQ = 0, MODE_DK = 1 };

constexpr int kStages3 = 2;
constexpr int kA2Bytes = kBlockM * kD * 2;          // 64 KB, four [128 x 128 B] slabs
constexpr int kA2ChunkBytes = kBlockM * 128;        // 16 KB
// tensor-memory column map
constexpr uint32_t k3ColAcc = 0;     // 256: dQ / dK accumulator
constexpr uint32_t k3ColA1 = 256;    // 128: fixed operand A1 (bf16 pairs)
constexpr uint32_t k3ColS = 384;     // 64 : S
constexpr uint32_t k3ColDP = 448;    // 64 : dP, then dS (bf16 pairs; same own-column placement as P above)

struct ThreeGemmParams {
  int La;
  int Lx;                      // streamed length (M in DQ, N in DK)
  float scale_log2;            // scale * log2(e)
  float scale;                 // softmax scale (applied to ACC in the epilogue)
  const float* lse2;           // [B, N]   log2-domain LSE of the forward
  const float* delta;          // [B, N]   rowsum(dO o O)
  GradOut gout;                // dQ / dK
  sam2b200::Dropout drop;      // attention-probability dropout: dP is masked and scaled like P was in the forward
  unsigned long long* dbg;     // optional timeline buffer
  int n_items, n_atiles;       // persistent kernel: items = (a block, batch) pairs, a blocks per batch
  const float* dp_bias;        // raw-memory path with dropout: [B, N] per-query constant dO . bv added to dP before the mask
};

struct SharedStorage3 {
  alignas(1024) uint8_t a2[kA2Bytes];
  alignas(1024) uint8_t x_tiles[kStages3][kTileBytes];
  alignas(1024) uint8_t y_tiles[kStages3][kTileBytes];
  alignas(8) uint64_t x_full[kStages3];
  uint64_t x_empty[kStages3];
  uint64_t y_full[kStages3];
  uint64_t y_empty[kStages3];
  uint64_t a2_full;
  uint64_t a1_full;            // TMA: A1 landed in ring stage kStages3-1
  uint64_t a1_ready;           // compute warps: A1 moved to tensor memory
  uint64_t s_full;
  uint64_t s_free;
  uint64_t dp_full;
  uint64_t ds_ready;
  uint64_t acc_done;
  uint64_t a2_empty;           // persistent kernel: the last dP MMA of an item has read A2
  float col_lse[2][kBlockN];
  float col_delta[2][kBlockN];
  uint32_t tmem_base;
};

// DY: width of the A2 / Y operands of the dP GEMM: 256 = projected values (A2 = dO or V), 64 = the raw memory features
// with A2 / Y = dO' = dO Wv or mem (dP = dO V^T = dO' mem^T up to a per-row constant that dP - Delta cancels).
template <int MODE, bool DROP, int DY = kD>
__global__ void __launch_bounds__(kThreads, 1)
three_gemm_kernel(const __grid_constant__ CUtensorMap map_a2, const __grid_constant__ CUtensorMap map_x,
                  const __grid_constant__ CUtensorMap map_y,
                  const __grid_constant__ CUtensorMap map_a1,   // A1 [B, La, 256] bf16, box 64 x 128
                  const __grid_constant__ CUtensorMap map_g,    // dQ / dK (bf16 box 64 x 32 | fp32 box 32 x 32)
                  const ThreeGemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  SharedStorage3& sh = *reinterpret_cast<SharedStorage3*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int a_tile = blockIdx.x;
  const int b = blockIdx.y;
  const int nt = (p.Lx + kBlockN - 1) / kBlockN;
  if (p.dbg != nullptr && threadIdx.x == 0) {
    unsigned long long* e = p.dbg + ((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 8;
    e[0] = smid(); e[7] = nt;
  }
  SAM2B200_STAMP(p.dbg, 1);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages3; ++s) {
      mbar_init(&sh.x_full[s], 1); mbar_init(&sh.x_empty[s], 1);
      mbar_init(&sh.y_full[s], 1); mbar_init(&sh.y_empty[s], 1);
    }
    mbar_init(&sh.a2_full, 1);
    mbar_init(&sh.a1_full, 1);
    mbar_init(&sh.a1_ready, kNumSoftmaxThreads);
    mbar_init(&sh.s_full, 1);
    mbar_init(&sh.s_free, kNumSoftmaxThreads);
    mbar_init(&sh.dp_full, 1);
    mbar_init(&sh.ds_ready, kNumSoftmaxThreads);
    mbar_init(&sh.acc_done, 1);
    fence_barrier_init();
  }
  if (warp == kProducerWarp && lane == 0) { prefetch_tmap(&map_a1); prefetch_tmap(&map_a2); prefetch_tmap(&map_x); prefetch_tmap(&map_y); }
  if (warp == 0 && lane == 0) prefetch_tmap(&map_g);
  if (warp == kMmaWarp) { tmem_alloc(&sh.tmem_base, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sh.tmem_base;
  SAM2B200_STAMP(p.dbg, 2);

  if (warp == kProducerWarp) {
    const bool leader = elect_one();
    if (leader) {
      mbar_arrive_expect_tx(&sh.a1_full, 4 * kSlabBytes);   // A1 -> the last ring stage (needed first: it goes to TMEM)
#pragma unroll
      for (int c = 0; c < 4; ++c)
        tma_load_3d((c < 2 ? &sh.x_tiles[kStages3 - 1][0] : &sh.y_tiles[kStages3 - 1][0]) + (c & 1) * kSlabBytes, &map_a1,
                    &sh.a1_full, c * 64, a_tile * kBlockM, b);
      mbar_arrive_expect_tx(&sh.a2_full, kBlockM * DY * 2);
#pragma unroll
      for (int c = 0; c < DY / 64; ++c)
        tma_load_3d(&sh.a2[c * kA2ChunkBytes], &map_a2, &sh.a2_full, c * 64, a_tile * kBlockM, b);
    }
    __syncwarp();
    for (int j = 0; j < nt; ++j) {
      const int s = j % kStages3;
      const uint32_t ph = (j / kStages3) & 1;
      const int row0 = j * kBlockN;
      if (j == kStages3 - 1) mbar_wait(&sh.a1_ready, 0);   // first use of the stage that staged A1
      mbar_wait(&sh.x_empty[s], ph ^ 1);
      if (leader) {
        mbar_arrive_expect_tx(&sh.x_full[s], kTileBytes);
#pragma unroll
        for (int c = 0; c < 4; ++c)
          tma_load_3d(&sh.x_tiles[s][c * kChunkBytes], &map_x, &sh.x_full[s], c * 64, row0, b);
      }
      __syncwarp();
      mbar_wait(&sh.y_empty[s], ph ^ 1);
      if (leader) {
        mbar_arrive_expect_tx(&sh.y_full[s], kBlockN * DY * 2);
#pragma unroll
        for (int c = 0; c < DY / 64; ++c)
          tma_load_3d(&sh.y_tiles[s][c * kChunkBytes], &map_y, &sh.y_full[s], c * 64, row0, b);
      }
      __syncwarp();
    }
  } else if (warp == kMmaWarp) {
    const bool leader = elect_one();    // converged warp, one elected issuer, add-only descriptors (see two_gemm_kernel)
    constexpr uint32_t idesc_s = make_idesc_bf16(kBlockM, kBlockN, 0, 0);
    constexpr uint32_t idesc_acc = make_idesc_bf16(kBlockM, kD, 0, 1);
    const uint32_t a2_lo = desc_lo_sw128(smem_u32(&sh.a2[0]), 16);
    const uint32_t x_lo0 = desc_lo_sw128(smem_u32(&sh.x_tiles[0][0]), 16);            // K-major view of X
    const uint32_t xm_lo0 = desc_lo_sw128(smem_u32(&sh.x_tiles[0][0]), kChunkBytes);  // MN-major view of X
    const uint32_t y_lo0 = desc_lo_sw128(smem_u32(&sh.y_tiles[0][0]), 16);
    auto issue_s = [&](int t) {        // S[t] = A1(TMEM) . X[t]^T
      const int s = t % kStages3;
      mbar_wait(&sh.x_full[s], (t / kStages3) & 1);
      tc_fence_after();
      if (leader) {
        const uint32_t xlo = x_lo0 + s * (kTileBytes >> 4);
#pragma unroll
        for (int ks = 0; ks < kD / 16; ++ks)
          umma_ts_lohi(tmem + k3ColS, tmem + k3ColA1 + ks * 8, xlo + (ks >> 2) * (kChunkBytes >> 4) + (ks & 3) * 2,
                       kDescHiSw128_1024, idesc_s, ks > 0);
        umma_commit(&sh.s_full);
      }
      __syncwarp();
    };
    auto issue_dp = [&](int t) {       // dP[t] = A2(SMEM) . Y[t]^T
      const int s = t % kStages3;
      mbar_wait(&sh.y_full[s], (t / kStages3) & 1);
      tc_fence_after();
      if (leader) {
        const uint32_t ylo = y_lo0 + s * (kTileBytes >> 4);
#pragma unroll
        for (int ks = 0; ks < DY / 16; ++ks)
          umma_ss_lohi(tmem + k3ColDP, a2_lo + (ks >> 2) * (kA2ChunkBytes >> 4) + (ks & 3) * 2,
                       ylo + (ks >> 2) * (kChunkBytes >> 4) + (ks & 3) * 2, kDescHiSw128_1024, idesc_s, ks > 0);
        umma_commit(&sh.y_empty[s]);
        umma_commit(&sh.dp_full);
      }
      __syncwarp();
    };
    mbar_wait(&sh.a1_ready, 0);
    mbar_wait(&sh.a2_full, 0);
    tc_fence_after();
    issue_s(0);
    issue_dp(0);
    for (int j = 0; j < nt; ++j) {
      const int s = j % kStages3;
      if (j + 1 < nt) {                 // S[j+1] as soon as the compute warps have read S[j]
        mbar_wait(&sh.s_free, j & 1);
        tc_fence_after();
        issue_s(j + 1);
      }
      mbar_wait(&sh.ds_ready, j & 1);
      tc_fence_after();
      if (leader) {
        const uint32_t xlo = xm_lo0 + s * (kTileBytes >> 4);
#pragma unroll
        for (int ks = 0; ks < kBlockN / 16; ++ks)
          umma_ts_lohi(tmem + k3ColAcc, tmem + k3ColDP + p_col_of_kstep(ks), xlo + ks * (2048 >> 4), kDescHiSw128_1024,
                       idesc_acc, (j > 0) || (ks > 0));
        umma_commit(&sh.x_empty[s]);
        if (j + 1 >= nt) umma_commit(&sh.acc_done);
      }
      __syncwarp();
      if (j + 1 < nt) issue_dp(j + 1);
    }
  } else {
    const int quarter = warp & 3;
    const int half = warp >> 2;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = tmem + (uint32_t(quarter * 32) << 16);
    const long long a_row_idx = (long long)a_tile * kBlockM + row;
    const bool row_valid = a_row_idx < p.La;
    uint32_t tab = 0;
    if (p.gout.rope_smem > 0) {     // rotation table of this CTA's rows -> shared memory, under the operand fetch
      const uint32_t area = smem_u32(&sh) + (uint32_t)sizeof(SharedStorage3);
      const uint32_t ybase = area + p.gout.rope_w * kRopeXStride;
      rope_stage_x(p.gout, area, threadIdx.x);
      rope_stage_y(p.gout, ybase, a_tile * kBlockM, threadIdx.x);
      cp_async_commit();
      tab = rope_tab_addr(p.gout, area, ybase, a_tile * kBlockM, row, half);
    }
    {
      mbar_wait(&sh.a1_full, 0);
      stage_to_tmem_half(smem_u32(half ? &sh.y_tiles[kStages3 - 1][0] : &sh.x_tiles[kStages3 - 1][0]), row, lane_addr + k3ColA1, half);
      tc_fence_before();
      mbar_arrive(&sh.a1_ready);
      SAM2B200_STAMP(p.dbg, 3);
    }
    const int row0 = a_tile * kBlockM + quarter * 32;
    const uint32_t stage = smem_u32(&sh.a2[0]) + warp * (4 * kBoxBytes);   // epilogue staging in the (then idle) operand buffers
    const bool rotate = p.gout.rope_table != nullptr && (row0 + lane) < p.gout.rope_rows;
    const float c = p.scale_log2;
    constexpr bool drop_on = DROP;
    const uint32_t drop_key = drop_on ? sam2b200::dropout_key(*p.drop.seed, p.drop.site) : 0u;
    float row_lse = 0.f, row_delta = 0.f;
    if (MODE == MODE_DQ && row_valid) {
      row_lse = p.lse2[(long long)b * p.La + a_row_idx];
      row_delta = p.delta[(long long)b * p.La + a_row_idx];
    }
    float lse_next = INFINITY, delta_next = 0.f;   // DK: per-column vectors staged one tile ahead
    if (MODE == MODE_DK && threadIdx.x < kBlockN && (int)threadIdx.x < p.Lx) {
      lse_next = p.lse2[(long long)b * p.Lx + threadIdx.x];
      delta_next = p.delta[(long long)b * p.Lx + threadIdx.x];
    }
    for (int j = 0; j < nt; ++j) {
      if (MODE == MODE_DK) {
        if (threadIdx.x < kBlockN) {
          sh.col_lse[j & 1][threadIdx.x] = lse_next;
          sh.col_delta[j & 1][threadIdx.x] = delta_next;
          const int col = (j + 1) * kBlockN + threadIdx.x;
          const bool ok = (j + 1 < nt) && col < p.Lx;
          lse_next = ok ? p.lse2[(long long)b * p.Lx + col] : INFINITY;
          delta_next = ok ? p.delta[(long long)b * p.Lx + col] : 0.f;
        }
        asm volatile("bar.sync 5, 256;" ::: "memory");
      }
      mbar_wait(&sh.s_full, j & 1);
      tc_fence_after();
      if (j == 0) SAM2B200_STAMP(p.dbg, 4);
      float pv[kHalfN];
      {
        uint32_t r0[32];
        SAM2B200_TMEM_LD32(lane_addr + k3ColS + half * kHalfN, r0);
        tmem_wait_ld();
        tc_fence_before();
        mbar_arrive(&sh.s_free);          // S region may be overwritten by S[j+1]
        const int ncols = p.Lx - j * kBlockN - half * kHalfN;
#pragma unroll
        for (int i = 0; i < kHalfN; ++i) {
          const float sraw = __uint_as_float(r0[i]);
          float e;
          if (MODE == MODE_DQ) e = (i < ncols) ? ex2(fmaf(sraw, c, -row_lse)) : 0.f;
          else e = ex2(fmaf(sraw, c, -sh.col_lse[j & 1][half * kHalfN + i]));
          pv[i] = e;
        }
      }
      mbar_wait(&sh.dp_full, j & 1);
      tc_fence_after();
      uint32_t pk[16];
      {
        uint32_t r0[32];
        SAM2B200_TMEM_LD32(lane_addr + k3ColDP + half * kHalfN, r0);
        tmem_wait_ld();
        // element (query q, key k) has dropout index (b N + q) M + k: DQ rows are queries, DK rows are keys
        const uint32_t didx = (MODE == MODE_DQ)
            ? (uint32_t)(((long long)b * p.La + a_row_idx) * p.Lx) + (uint32_t)(j * kBlockN + half * kHalfN)
            : (uint32_t)(((long long)b * p.Lx + (j * kBlockN + half * kHalfN)) * p.La + a_row_idx);
        const uint32_t dstep = (MODE == MODE_DQ) ? 1u : (uint32_t)p.La;
#pragma unroll
        for (int i = 0; i < kHalfN; i += 2) {
          float d0 = __uint_as_float(r0[i]);
          float d1 = __uint_as_float(r0[i + 1]);
          if (drop_on) {
            d0 = sam2b200::dropout_keep(drop_key, didx + (uint32_t)i * dstep, p.drop.thresh) ? d0 * p.drop.inv_keep : 0.f;
            d1 = sam2b200::dropout_keep(drop_key, didx + (uint32_t)(i + 1) * dstep, p.drop.thresh) ? d1 * p.drop.inv_keep : 0.f;
          }
          const float dl0 = (MODE == MODE_DQ) ? row_delta : sh.col_delta[j & 1][half * kHalfN + i];
          const float dl1 = (MODE == MODE_DQ) ? row_delta : sh.col_delta[j & 1][half * kHalfN + i + 1];
          pk[i >> 1] = pack_bf16(pv[i] * (d0 - dl0), pv[i + 1] * (d1 - dl1));
        }
      }
      SAM2B200_TMEM_ST16(lane_addr + k3ColDP + half * kHalfN, pk);   // dS over this warp's own dP columns
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(&sh.ds_ready);
    }
    float2 tcur[16];
    if (tab == 0) load_table_chunk(p.gout, rotate, row0 + lane, half * 128, tcur);   // in flight while the last MMAs drain
    else { cp_async_wait_all(); asm volatile("bar.sync 6, 256;" ::: "memory"); }       // the staged table is complete and visible
    mbar_wait(&sh.acc_done, 0);
    tc_fence_after();
    SAM2B200_STAMP(p.dbg, 5);
    grad_epilogue(p.gout, &map_g, stage, lane_addr + k3ColAcc, half, lane, row0, p.La, b, p.scale, rotate, tcur, nullptr, tab);
  }

  tc_fence_before();
  __syncthreads();
  SAM2B200_STAMP(p.dbg, 6);
  if (warp == kMmaWarp) tmem_dealloc(tmem, 512);
}

}  // namespace attn
