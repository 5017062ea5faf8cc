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
// Blackwell mapping: RESIDENT CTAs (one per SM) walk 128-row x BN-column tiles of C (BN = 256; 128 for rotated outputs -- the x and
// y halves of a head; 64); A k-slices of 64 arrive by TMA in a ring shared by consecutive tiles (the producer runs ahead into the next
// tile while the current one drains; ring depth, staging boxes and the table cache are laid out per problem by the host).  K <= 256: the CTA keeps ONE column block whose weight slices are loaded once and stay
// in shared memory (a first version re-streamed them per tile and was bound by L2 -> SM traffic: 387 MB for the linear1 head at cfg2);
// K > 256: weight slices travel with the A slices and, with more row tiles than SMs, two row tiles (two accumulators) share every
// weight slice.  Both operands are read from shared memory (SS-mode tcgen05.mma, M = 128, N = BN,
// K = 16 per instruction), the B tile K-major (NT) or MN-major (NN) straight from the TMA boxes -- no transposed weight copy exists
// anywhere; two BN-column fp32 accumulators in tensor memory alternate between tiles, so the epilogue of tile i (tcgen05.ld -> bias /
// rotation / ReLU -> bf16 -> swizzled staging box -> TMA store per warp) overlaps the MMAs of tile i + 1.  The x half of the axial
// rotation table (<= 64 rows x 64 pairs) is cached in shared memory with a padded row stride: a warp's 32 rows have up to 32 different
// x positions, i.e. 32 different L1 lines per load instruction when read from global memory; the y half (<= 3 rows per warp) is read
// through L1.  128-column blocks run TWO epilogue groups of 8 warps on four accumulators (the drain of a 128 x 128 accumulator is a
// latency chain of ~1.5 us per warp).  The kernel is launched with programmatic dependent launch: its set-up overlaps the tail of the
// kernel in front of it.  HBM-bound at every shape of the stack: algorithmic bytes = 2 (R K + K Nout + R Nout); measured it is bound
// by the accumulator drain (profiles/r2_timeline_gemm.txt).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#include "abi_common.cuh"
#include "dropout.cuh"
#include "sm100.cuh"
#include "tma_desc.cuh"

namespace gemm {

using namespace sm100;

constexpr int kBlockM = 128;                 // rows per accumulator (TMEM lanes)
constexpr int kBlockK = 64;                  // contraction slice per ring slot (one 128-byte swizzle row)
constexpr int kMaxSlots = 12;
constexpr int kEpiWarps = 8;                 // epilogue warps per group: warp (quarter, half) = TMEM lane quarter x column half
// EG epilogue groups of 8 warps (EG = 2, BN = 128 only: four 128-column accumulators, group g drains items g, g + 2, ...), then one
// TMA producer warp and one MMA issuer warp
__host__ __device__ constexpr int threads_for(int eg) { return (eg * kEpiWarps + 2) * 32; }
constexpr int kATileBytes = kBlockM * 128;   // 16 KB: [128 rows x 128 B]
constexpr int kBoxBytes = 32 * 128;          // one warp's output box: 32 rows x 64 bf16 columns
constexpr int kXStride = 528;                // bytes per cached table row: 64 (cos, sin) pairs + 16 B pad (conflict-free 16-byte reads)
constexpr int kXRows = 64;                   // cached x rows (grids up to 64 x 64 = 1024 px; larger grids read the table from L1 / L2)
constexpr int kCtrlBytes = 2048;

struct Ctrl {                                // first 2 KB of the (1024-aligned) dynamic shared memory
  uint64_t full[kMaxSlots];
  uint64_t empty[kMaxSlots];
  uint64_t b_full[4];                        // resident-weight mode: slice ks of the column block has landed
  uint64_t acc_full[4];
  uint64_t acc_free[4];
  float bias[256];                           // the current column block's bias slice
  uint32_t tmem_base;
};
static_assert(sizeof(Ctrl) <= kCtrlBytes, "control block");

struct Params {
  int rows;                 // R
  int n_row_tiles;          // ceil(R / 128)
  int n_col_blocks;         // Nout / BN
  int blocks_per_out;       // output tensor of column block cb = cb / blocks_per_out, its first column (cb % blocks_per_out) * BN
  int ksteps;               // K / 64
  int nout;                 // Nout (dropout element index = row * Nout + column)
  int rope_blocks;          // BN = 128: leading column blocks that are rotated (two per 256-wide head)
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
  // shared-memory layout chosen by the host (byte offsets from the 1024-aligned base; every region 1024-aligned but the x cache)
  int wres;                 // 1: K <= 256, the column block's weight slices stay in the B region; 0: B slices travel with the A slices
  int n_slots;              // ring slots (each: MT A tiles of 16 KB, + one B slice when !wres)
  int stage_bufs;           // staging boxes per epilogue warp (1 | 2)
  uint32_t off_a, off_b, off_stage, off_x;
  unsigned long long* dbg;  // optional per-CTA phase timeline (32 x u64 per CTA, %globaltimer ns; sam2b200_gemm_debug_timeline), else nullptr
};

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// Work items of one CTA; an item = MT consecutive row tiles x one column block.  Resident weights: the CTA keeps ONE column block and
// walks items blockIdx.x / ncb, + gridDim.x / ncb, ... (gridDim.x is a multiple of ncb).  Streamed weights: items t = blockIdx.x,
// + gridDim.x, ... with the column block fastest (CTAs that run together share the A tiles in L2).
struct Schedule {
  int ncb, n_groups, first, step, cb_fixed;
  bool wres;
  __device__ Schedule(const Params& p, int mt) {
    ncb = p.n_col_blocks; n_groups = (p.n_row_tiles + mt - 1) / mt; wres = p.wres != 0;
    if (wres) { cb_fixed = blockIdx.x % ncb; first = blockIdx.x / ncb; step = gridDim.x / ncb; }
    else      { cb_fixed = 0; first = blockIdx.x; step = gridDim.x; }
  }
  __device__ bool item(int i, int& group, int& cb) const {
    const int t = first + i * step;
    if (wres) { group = t; cb = cb_fixed; return t < n_groups; }
    group = t / ncb; cb = t - group * ncb;
    return group < n_groups;
  }
};

template <int BN, int B_MN, int MT, int EG>
__global__ void __launch_bounds__(threads_for(EG), 1)
gemm_kernel(const __grid_constant__ CUtensorMap map_a,      // A [R, K]: box 64 x 128
            const __grid_constant__ CUtensorMap map_b,      // NT: W [Nout, K], box 64 x BN;  NN: W [K, Nout], box 64 x 64
            const __grid_constant__ CUtensorMap map_c0,     // outputs [R, out_width]: box 64 x 32 (store)
            const __grid_constant__ CUtensorMap map_c1, const __grid_constant__ CUtensorMap map_c2, const Params p) {
  static_assert(MT == 1 || BN == 256, "two accumulators per item: BN = 256 only (2 x 256 TMEM columns)");
  static_assert(EG == 1 || (BN == 128 && MT == 1), "two epilogue groups: BN = 128 only (4 x 128 TMEM columns)");
  constexpr int kProdWarp = EG * kEpiWarps, kMmaWarp = kProdWarp + 1;
  constexpr int kAcc = 2 * EG;                              // accumulators in tensor memory (MT = 1)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  Ctrl& sh = *reinterpret_cast<Ctrl*>(base);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t kBBytes = BN * 128;
  constexpr uint32_t kABytes = MT * kATileBytes;            // A bytes per ring slot
  constexpr uint32_t kTmemCols = (MT == 2 ? 2 : kAcc) * BN;
  const uint32_t a_base = smem_u32(base + p.off_a), b_base = smem_u32(base + p.off_b);
  const int nslots = p.n_slots;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kMaxSlots; ++i) { mbar_init(&sh.full[i], 1); mbar_init(&sh.empty[i], 1); }
    for (int i = 0; i < 4; ++i) mbar_init(&sh.b_full[i], 1);
    for (int i = 0; i < 4; ++i) { mbar_init(&sh.acc_full[i], 1); mbar_init(&sh.acc_free[i], kEpiWarps * 32); }
    fence_barrier_init();
  }
  if (warp == kProdWarp && lane == 0) { prefetch_tmap(&map_a); prefetch_tmap(&map_b); }
  if (warp == 0 && lane == 0) { prefetch_tmap(&map_c0); prefetch_tmap(&map_c1); prefetch_tmap(&map_c2); }
  if (warp == kMmaWarp) { tmem_alloc(&sh.tmem_base, kTmemCols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // programmatic dependent launch: everything above touched no global memory; the kernel behind this one may start its own set-up
  // now, and this one waits here for the kernel in front of it (whose output is an operand of this GEMM) to be complete
  pdl_launch_dependents();
  pdl_wait();
  const uint32_t tmem = sh.tmem_base;
  const int ksteps = p.ksteps;
  const Schedule sched(p, MT);
  unsigned long long* dbg = p.dbg ? p.dbg + (size_t)blockIdx.x * 32 : nullptr;
  if (dbg && threadIdx.x == 0) dbg[0] = gtime();

  if (warp == kProdWarp) {
    // ===================== TMA producer =====================
    const bool leader = elect_one();
    auto load_b = [&](uint32_t dst, uint64_t* bar, int ks, int cb) {
      if (B_MN) {
#pragma unroll
        for (int c = 0; c < BN / 64; ++c) tma_load_3d_addr(dst + c * 8192, &map_b, bar, cb * BN + c * 64, ks * kBlockK, 0);
      } else {
        tma_load_3d_addr(dst, &map_b, bar, ks * kBlockK, cb * BN, 0);
      }
    };
    if (sched.wres && leader) {               // the column block's weights: once, slice ks -> B slot ks
      for (int ks = 0; ks < ksteps; ++ks) {
        mbar_arrive_expect_tx(&sh.b_full[ks], kBBytes);
        load_b(b_base + ks * kBBytes, &sh.b_full[ks], ks, sched.cb_fixed);
      }
    }
    __syncwarp();
    int s = 0;
    uint32_t ph = 0;
    int group, cb;
    for (int i = 0; sched.item(i, group, cb); ++i) {
      for (int ks = 0; ks < ksteps; ++ks) {
        mbar_wait(&sh.empty[s], ph ^ 1);
        if (leader) {
          mbar_arrive_expect_tx(&sh.full[s], kABytes + (sched.wres ? 0u : kBBytes));
#pragma unroll
          for (int m = 0; m < MT; ++m)          // a row tile beyond R loads as zeros (its stores are skipped)
            tma_load_3d_addr(a_base + s * kABytes + m * kATileBytes, &map_a, &sh.full[s], ks * kBlockK, (group * MT + m) * kBlockM, 0);
          if (!sched.wres) load_b(b_base + s * kBBytes, &sh.full[s], ks, cb);
        }
        __syncwarp();
        if (++s == nslots) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer =====================
    const bool leader = elect_one();
    constexpr uint32_t idesc = make_idesc_bf16(kBlockM, BN, 0, B_MN);
    const uint32_t a_lo0 = desc_lo_sw128(a_base, 16);                   // K-major: LBO unused
    const uint32_t b_lo0 = desc_lo_sw128(b_base, B_MN ? 8192 : 16);     // MN-major: LBO = 64-column slab stride
    constexpr uint32_t b_kstep = B_MN ? (2048 >> 4) : 2;    // 16 k-rows x 128 B (MN-major) | 32 B inside the row (K-major)
    int s = 0;
    uint32_t ph = 0;
    int group, cb;
    for (int it = 0; sched.item(it, group, cb); ++it) {
      // MT = 1: accumulators alternate between items; MT = 2: both belong to the item (no overlap with the previous drain)
      const int ab = MT == 1 ? (it % kAcc) : 0;
      mbar_wait(&sh.acc_free[ab], (MT == 1 ? ((it / kAcc) & 1) : (it & 1)) ^ 1);
      tc_fence_after();
      if (dbg && leader && it < 4) dbg[2 + 4 * it] = gtime();        // accumulator free: this item's MMAs may be issued
      const uint32_t d = tmem + ab * BN;
      for (int ks = 0; ks < ksteps; ++ks) {
        mbar_wait(&sh.full[s], ph);
        if (sched.wres && it == 0) mbar_wait(&sh.b_full[ks], 0);
        tc_fence_after();
        if (leader) {
          const uint32_t alo = a_lo0 + s * (kABytes >> 4);
          const uint32_t blo = b_lo0 + (sched.wres ? ks : s) * (kBBytes >> 4);
#pragma unroll
          for (int m = 0; m < MT; ++m)
#pragma unroll
            for (int k16 = 0; k16 < 4; ++k16)
              umma_ss_lohi(d + m * BN, alo + m * (kATileBytes >> 4) + k16 * 2, blo + k16 * b_kstep, kDescHiSw128_1024, idesc, (ks | k16) != 0);
          umma_commit(&sh.empty[s]);
          if (ks == ksteps - 1) { umma_commit(&sh.acc_full[ab]); if (dbg && it < 4) dbg[3 + 4 * it] = gtime(); }   // last slice had landed, all MMAs issued
        }
        __syncwarp();
        if (++s == nslots) { s = 0; ph ^= 1; }
      }
    }
  } else {
    // ===================== epilogue warps (0..7): warp (quarter, half) owns rows quarter*32.. of column chunks c = half, half + 2 =====================
    const int group = warp / kEpiWarps;                     // epilogue group (EG = 2: items group, group + 2, ...)
    const int gtid = threadIdx.x - group * kEpiWarps * 32;  // thread index inside the group
    const int quarter = warp & 3;
    const int half = (warp >> 2) & 1;
    const uint32_t lane_addr = tmem + (uint32_t(quarter * 32) << 16);
    const int nbufs = p.stage_bufs;
    const uint32_t stage0 = smem_u32(base + p.off_stage) + warp * nbufs * kBoxBytes;
    uint32_t nstore = 0;
    constexpr int kChunks = BN / 64;
    constexpr int kMyChunks = (kChunks + 1) / 2;            // chunks per warp (BN = 64: the upper four warps have none)
    const bool drop_on = p.drop_out.seed != nullptr;
    const uint32_t dkey = drop_on ? sam2b200::dropout_key(*p.drop_out.seed, p.drop_out.site) : 0u;
    // x half of the axial rotation table -> shared memory, once per CTA (rows x = 0 .. w-1, 64 pairs each)
    const bool xcached = BN == 128 && p.rope_blocks > 0 && p.rope_w > 0 && p.rope_w <= kXRows;
    const uint32_t xbase = smem_u32(base + p.off_x);
    if (BN == 128 && xcached) {
      const float4* src = reinterpret_cast<const float4*>(p.table);
      for (int idx = threadIdx.x; idx < p.rope_w * 32; idx += EG * kEpiWarps * 32) {
        const int x = idx >> 5, q = idx & 31;
        *reinterpret_cast<float4*>(base + p.off_x + x * kXStride + q * 16) = __ldg(src + x * 64 + q);
      }
      asm volatile("bar.sync 7, %0;" ::"r"(EG * kEpiWarps * 32) : "memory");
    }
    float* bias_s = sh.bias + group * BN;               // this group's copy of the bias slice
    int rgroup, cb;
    for (int it = group; sched.item(it, rgroup, cb); it += EG) {
      const int ab = MT == 1 ? (it % kAcc) : 0;
      const int which = cb / p.blocks_per_out;
      const int ocol0 = (cb - which * p.blocks_per_out) * BN;         // first column of this block inside its output tensor
      const CUtensorMap* mo = which == 0 ? &map_c0 : (which == 1 ? &map_c1 : &map_c2);
      // this column block's bias slice -> shared memory (broadcast reads below); with a fixed column block it is loaded once
      if (it == group || (!sched.wres && sched.ncb > 1)) {
        if (it != group) asm volatile("bar.sync %0, 256;" ::"r"(5 + group) : "memory");      // every warp of the group has finished reading the previous slice
        if (gtid < BN) bias_s[gtid] = p.bias ? __ldg(p.bias + cb * BN + gtid) : 0.f;
        asm volatile("bar.sync %0, 256;" ::"r"(5 + group) : "memory");
      }
      mbar_wait(&sh.acc_full[ab], MT == 1 ? ((it / kAcc) & 1) : (it & 1));
      tc_fence_after();
      if (dbg && threadIdx.x == 0 && it < 4) dbg[4 + 4 * it] = gtime();   // MMAs of this item complete
      if (kMyChunks * 2 > kChunks && half == 1) {   // BN = 64: the upper four warps only release the accumulator
        tc_fence_before();
        mbar_arrive(&sh.acc_free[ab]);
        continue;
      }
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        const int row0 = (rgroup * MT + m) * kBlockM + quarter * 32;
        const uint32_t acc_col = (MT == 1 ? ab : m) * BN;
        // rotation (BN = 128): block cb covers head columns [0, 128) (the x half of the axial table) if cb is even, else [128, 256) (y half)
        bool rotate = false;
        const float4* tsrc = nullptr;               // this thread's (cos, sin) pairs for the block's 128 columns: 64 pairs = 32 float4
        uint32_t xs_addr = 0;
        if (BN == 128 && cb < p.rope_blocks) {
          const int pos = (int)((unsigned)(row0 + lane) % (unsigned)p.rows_per_item);
          if (pos < p.n_rope_rows) {
            rotate = true;
            const int tpos = pos % p.period;
            const bool xhalf = (cb & 1) == 0;
            if (xhalf && xcached) xs_addr = xbase + (tpos % p.rope_w) * kXStride;
            else {
              const int trow = p.rope_w > 0 ? (xhalf ? tpos % p.rope_w : tpos - tpos % p.rope_w) : tpos;
              tsrc = reinterpret_cast<const float4*>(p.table + (long long)trow * 128 + (xhalf ? 0 : 64));
            }
          }
        }
#pragma unroll
        for (int ci = 0; ci < kMyChunks; ++ci) {
          const int c = half + 2 * ci;
          float dot = 0.f;
          // staging: the box handed to the TMA unit `nbufs` chunks ago must have been read before its buffer is rewritten
          const uint32_t sbuf = stage0 + ((nbufs == 2) ? (nstore & 1) : 0) * kBoxBytes;
          if (lane == 0) {
            if (nbufs == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            else tma_store_wait_read();
          }
          __syncwarp();
          const uint32_t srow = sbuf + lane * 128;
#pragma unroll
          for (int sb = 0; sb < 2; ++sb) {          // two sub-blocks of 32 columns
            uint32_t acc[32];
            SAM2B200_TMEM_LD32(lane_addr + acc_col + c * 64 + sb * 32, acc);
            float4 cs[8];
            if (BN == 128 && rotate) {              // the 16 (cos, sin) pairs of this row's 32 columns, fetched under the TMEM load
              if (tsrc == nullptr) {
#pragma unroll
                for (int i = 0; i < 8; ++i) { const uint4 u = lds128(xs_addr + (c * 16 + sb * 8 + i) * 16);
                  cs[i] = make_float4(__uint_as_float(u.x), __uint_as_float(u.y), __uint_as_float(u.z), __uint_as_float(u.w)); }
              } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) cs[i] = __ldg(tsrc + c * 16 + sb * 8 + i);
              }
            }
            tmem_wait_ld();
            if (sb == 1 && ci == kMyChunks - 1 && m == MT - 1) { tc_fence_before(); mbar_arrive(&sh.acc_free[ab]); }   // this thread's last read of the accumulator(s)
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(acc[i]) + bias_s[c * 64 + sb * 32 + i];
            if (BN == 128 && rotate) {
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
            if (row0 < p.rows) tma_store_3d_addr(mo, sbuf, ocol0 + c * 64, row0, 0);   // rows beyond R are clipped by the TMA unit
            tma_store_commit();
          }
          ++nstore;
        }
      }
      if (dbg && threadIdx.x == 0 && it < 4) dbg[5 + 4 * it] = gtime();   // warp 0: last box of this item handed to the TMA unit
      if (dbg && threadIdx.x == 0) dbg[31] = (unsigned long long)(it + 1);
    }
    if (dbg && threadIdx.x == 0) dbg[30] = gtime();
    if (lane == 0) tma_store_wait_read();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem, kTmemCols);
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

unsigned long long* g_dbg = nullptr;      // sam2b200_gemm_debug_timeline: consecutive launches append 32 x u64 per CTA
size_t g_dbg_cap = 0, g_dbg_used = 0;

constexpr int kSmemBudget = 227 * 1024 - 1024 - gemm::kCtrlBytes;      // after the alignment slack and the control block

// Shared-memory layout + grid for one problem (host only; also behind sam2b200_gemm_plan, which the CPU tests sweep).
// Regions (1024-aligned): control block | B | A ring | staging | x cache.
int plan(gemm::Params& p, int BN, int MT, int EG, int sms, unsigned* grid_out, size_t* smem_out) {
  const int b_slice = BN * 128, a_slot = MT * gemm::kATileBytes;
  const int x_bytes = (BN == 128 && p.rope_blocks > 0 && p.rope_w > 0 && p.rope_w <= gemm::kXRows) ? ((p.rope_w * gemm::kXStride + 1023) & ~1023) : 0;
  int stage_bytes = EG * gemm::kEpiWarps * gemm::kBoxBytes;
  p.stage_bufs = 1;
  int left;
  if (p.wres) {
    left = kSmemBudget - p.ksteps * b_slice - stage_bytes - x_bytes;
    p.n_slots = left / a_slot;
    // more than two row tiles of A in flight buy nothing: a second staging box per warp instead
    if (p.n_slots > 2 * p.ksteps + 2 && left - stage_bytes >= (2 * p.ksteps + 2) * a_slot) { p.stage_bufs = 2; stage_bytes *= 2; left -= stage_bytes / 2; p.n_slots = left / a_slot; }
  } else {
    left = kSmemBudget - stage_bytes - x_bytes;
    p.n_slots = left / (a_slot + b_slice);
  }
  if (p.n_slots > gemm::kMaxSlots) p.n_slots = gemm::kMaxSlots;
  if (p.n_slots < 2) return sam2b200::fail(SAM2B200_ERR_UNSUPPORTED, "gemm: shared memory layout does not fit");
  const int b_bytes = (p.wres ? p.ksteps : p.n_slots) * b_slice;
  p.off_b = gemm::kCtrlBytes; p.off_a = p.off_b + b_bytes; p.off_stage = p.off_a + p.n_slots * a_slot; p.off_x = p.off_stage + stage_bytes;
  const size_t smem = (size_t)p.off_x + x_bytes + 1024;
  if (smem > 227 * 1024) return sam2b200::fail(SAM2B200_ERR_UNSUPPORTED, "gemm: shared memory layout does not fit");
  const int ncb = p.n_col_blocks, groups = (p.n_row_tiles + MT - 1) / MT;
  unsigned grid;
  if (p.wres) {                               // resident weights: a CTA keeps one column block, grid = CTAs per block x blocks
    int per_cb = sms / ncb < 1 ? 1 : sms / ncb;
    if (per_cb > groups) per_cb = groups;
    grid = (unsigned)(per_cb * ncb);
  } else {
    const long long items = (long long)groups * ncb;
    grid = (unsigned)(items < sms ? items : sms);
  }
  *grid_out = grid; *smem_out = smem;
  return SAM2B200_OK;
}

// Kernel variant for a problem: column-block width, row tiles per item, epilogue groups (see the header comment).
void choose_variant(int K, int Nout, int rope_cols, int n_row_tiles, int sms, int* bn, int* mt, int* eg) {
  // rotated outputs: 128-column blocks = the x / y halves of a head.  SAM2B200_GEMM_BN128=1 (experiment, NOT faster: the linear1 head
  // takes 46.7 instead of 37.3 us -- twice the A traffic from L2 outweighs the second epilogue group): 128-column blocks with two
  // epilogue groups for every resident-weight problem wider than 256
  static const bool bn128_all = getenv("SAM2B200_GEMM_BN128") != nullptr;
  static const bool eg1 = getenv("SAM2B200_GEMM_EG1") != nullptr;
  const bool wres = K / gemm::kBlockK <= 4;
  *bn = Nout == 64 ? 64 : ((rope_cols > 0 || (bn128_all && K <= 256 && Nout > 256)) ? 128 : 256);
  // streamed weights and more row tiles than SMs: two row tiles (two accumulators) per item share every weight slice -- the kernel
  // is bound by L2 -> SM traffic otherwise (1 MB of weights per 128 rows at K = 2048)
  *mt = (*bn == 256 && !wres && n_row_tiles > sms) ? 2 : 1;
  // two epilogue groups (16 warps, four accumulators) for 128-column blocks with resident weights: the drain of a 128 x 128 accumulator
  // is a latency chain of ~1.5 us per warp, twice the MMA time -- SAM2B200_GEMM_EG1=1 keeps one group (A/B)
  *eg = (*bn == 128 && wres && !eg1) ? 2 : 1;
}

template <int BN, int B_MN, int MT, int EG = 1>
int launch(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap* mc, gemm::Params p, cudaStream_t stream) {
  unsigned grid;
  size_t smem;
  if (int rc = plan(p, BN, MT, EG, sm_count(), &grid, &smem)) return rc;
  cudaError_t e = cudaFuncSetAttribute(gemm::gemm_kernel<BN, B_MN, MT, EG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return sam2b200::fail(SAM2B200_ERR_CUDA, cudaGetErrorString(e));
  if (g_dbg && g_dbg_used + (size_t)grid * 32 <= g_dbg_cap) { p.dbg = g_dbg + g_dbg_used; g_dbg_used += (size_t)grid * 32; }
  static const bool pdl = getenv("SAM2B200_NO_PDL") == nullptr;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(gemm::threads_for(EG)); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  cudaError_t le = cudaLaunchKernelEx(&cfg, gemm::gemm_kernel<BN, B_MN, MT, EG>, ma, mb, mc[0], mc[1], mc[2], p);
  if (le != cudaSuccess) return sam2b200::fail(SAM2B200_ERR_CUDA, cudaGetErrorString(le));
  return sam2b200::check_launch("gemm");
}

bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace

extern "C" {

// Debug aid: per-CTA %globaltimer phase stamps of the following sam2b200_gemm* launches into buf (n_u64 entries, 32 per CTA:
// [0] start, per item i < 4: [2+4i] accumulator free, [3+4i] MMAs issued, [4+4i] MMAs complete, [5+4i] last store issued, [30] end,
// [31] tiles of the CTA).  buf = NULL switches it off; returns the number of entries written since the last call.
long long sam2b200_gemm_debug_timeline(void* buf, long long n_u64) {
  const long long used = (long long)g_dbg_used;
  g_dbg = static_cast<unsigned long long*>(buf); g_dbg_cap = buf ? (size_t)n_u64 : 0; g_dbg_used = 0;
  return used;
}

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
  int bn, mt, eg;
  choose_variant(K, Nout, rope_cols, (int)((R + gemm::kBlockM - 1) / gemm::kBlockM), sm_count(), &bn, &mt, &eg);
  const int n_out = out_width > 0 ? Nout / out_width : 0;
  if (!out0 || !a || !b || R <= 0 || R > 0x7fffffffLL - 256 || K <= 0 || (K % 64) || Nout <= 0 || (Nout % bn) || Nout > 2048 ||
      out_width <= 0 || (out_width % bn) || n_out * out_width != Nout || n_out > 3 || (n_out > 1 && !out1) || (n_out > 2 && !out2) ||
      (b_layout != 0 && b_layout != 1) || lda < K || ldc < out_width || ldb < (b_layout ? Nout : K) || ((lda | ldb | ldc) & 7) ||
      rope_cols < 0 || (rope_cols % 256) || rope_cols > Nout || (rope_cols > 0 && (!table || rows_per_item <= 0 || n_rope_rows < 0 || period <= 0)) ||
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
  p.rows = (int)R; p.n_col_blocks = Nout / bn; p.n_row_tiles = (int)((R + gemm::kBlockM - 1) / gemm::kBlockM);
  p.blocks_per_out = out_width / bn; p.ksteps = K / gemm::kBlockK; p.nout = Nout; p.bias = bias;
  p.rope_blocks = rope_cols / 128; p.table = reinterpret_cast<const float2*>(table); p.rows_per_item = rows_per_item > 0 ? rows_per_item : 1;
  p.n_rope_rows = n_rope_rows; p.period = period > 0 ? period : 1;
  if (rope_cols > 0) { int w = (int)(sqrt((double)p.period) + 0.5); p.rope_w = (w * w == p.period) ? w : 0; }
  p.relu = relu; p.drop_out = sam2b200::make_dropout(drop_seed, drop_site, drop_p);
  p.dot_rows = dot_rows; p.dot_out = dot_out;
  p.wres = p.ksteps <= 4;
  if (bn == 256) {
    if (mt == 2) return b_layout ? launch<256, 1, 2>(ma, mb, mc, p, stream) : launch<256, 0, 2>(ma, mb, mc, p, stream);
    return b_layout ? launch<256, 1, 1>(ma, mb, mc, p, stream) : launch<256, 0, 1>(ma, mb, mc, p, stream);
  }
  if (bn == 128) {
    if (eg == 2) return b_layout ? launch<128, 1, 1, 2>(ma, mb, mc, p, stream) : launch<128, 0, 1, 2>(ma, mb, mc, p, stream);
    return b_layout ? launch<128, 1, 1>(ma, mb, mc, p, stream) : launch<128, 0, 1>(ma, mb, mc, p, stream);
  }
  return b_layout ? launch<64, 1, 1>(ma, mb, mc, p, stream) : launch<64, 0, 1>(ma, mb, mc, p, stream);
}

// Host-only: the kernel variant and shared-memory layout sam2b200_gemm_ex would use for a problem on a GPU with `sms` SMs (no device
// needed).  out[8] = {BN, row tiles per item, epilogue groups, ring slots, staging boxes per warp, grid, resident weights (0 | 1),
// dynamic shared memory bytes}.  Returns 0, or the error sam2b200_gemm_ex would return for the layout.
int sam2b200_gemm_plan(long long R, int K, int Nout, int rope_cols, int period, int sms, long long* out) {
  if (!out || R <= 0 || K <= 0 || (K % 64) || Nout <= 0 || sms <= 0 || !(Nout == 64 || Nout % 256 == 0) || Nout > 2048 || (rope_cols % 256) || rope_cols > Nout)
    return sam2b200::fail(SAM2B200_ERR_INVALID, "gemm_plan: bad arguments");
  int bn, mt, eg;
  const int n_row_tiles = (int)((R + gemm::kBlockM - 1) / gemm::kBlockM);
  choose_variant(K, Nout, rope_cols, n_row_tiles, sms, &bn, &mt, &eg);
  gemm::Params p{};
  p.rows = (int)R; p.n_col_blocks = Nout / bn; p.n_row_tiles = n_row_tiles; p.ksteps = K / gemm::kBlockK; p.rope_blocks = rope_cols / 128;
  if (rope_cols > 0 && period > 0) { int w = (int)(sqrt((double)period) + 0.5); p.rope_w = (w * w == period) ? w : 0; }
  p.wres = p.ksteps <= 4;
  unsigned grid;
  size_t smem;
  if (int rc = plan(p, bn, mt, eg, sms, &grid, &smem)) return rc;
  out[0] = bn; out[1] = mt; out[2] = eg; out[3] = p.n_slots; out[4] = p.stage_bufs; out[5] = grid; out[6] = p.wres; out[7] = (long long)smem;
  return SAM2B200_OK;
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
