// Backward kernels of the raw-memory cross-attention with TWO SOFTMAX GROUPS working on alternate tiles.
//
// Round-1 evidence (profiles/r1_ncu_attn_v64_cfg2_cross.csv, scripts/timeline_v64.py): with S and dP double-buffered the
// MMA warp never waits for operands, yet a tile still takes ~1800 clk against ~1150 clk of tensor-pipe work, issue slots
// 25-32 % busy -- the eight softmax warps run ONE dependent chain per tile
//     wait S -> tcgen05.ld -> 32 x (FMA, EX2) -> wait dP -> tcgen05.ld -> 32 x (sub, mul, pack) -> tcgen05.st -> arrive
// whose latencies (two TMEM loads, the MUFU queue, the store) two warps per scheduler cannot hide.
// Here sixteen softmax warps form two groups of eight: group g owns score / dP buffer g and the tiles j = g (mod 2).
// Both chains are in flight at once (four warps per scheduler), the MMA warp alternates between the groups'
// `ds_ready` barriers exactly as before (tcgen05.mma executes in issue order, see attn_v64_kernels.cuh), and all sixteen
// warps drain the accumulator in the epilogue (64 columns each instead of 128).
//     TMEM   ACC 256 | S0 64 | S1 64 | dP0 64 | dP1 64           (unchanged)
//     SMEM   A1 64 KB | 3-stage ring of X tiles (96 KB) | Y tiles (24 KB) | A2 16 KB      (unchanged)
//     warps  0-7 softmax group 0, 8-15 softmax group 1, 16 TMA producer, 17 MMA issuer   (576 threads, <= 112 registers)
// Register budget: a thread touches its 32 columns of a tile in two passes of 16 (S and dP loaded together), the
// epilogue works on 32-column chunks with a single (cos, sin) buffer.
// Numerics: the same products in the same order as three_gemm_v64_kernel -- results are bit-identical.
#pragma once

#include "attn_v64_kernels.cuh"

namespace attn {

constexpr int kX2SoftmaxWarps = 16;
constexpr int kX2ProducerWarp = 16;
constexpr int kX2MmaWarp = 17;
constexpr int kX2Threads = 18 * 32;
constexpr int kX2GroupThreads = 256;

struct SharedStorageV64x2 {
  alignas(1024) uint8_t a1[kA2Bytes];                         // 64 KB; epilogue staging (with the ring behind it) once the MMAs are done
  alignas(1024) uint8_t x_tiles[kV64Stages][kTileBytes];
  alignas(1024) uint8_t y_tiles[kV64Stages][kV64YBytes];
  alignas(1024) uint8_t a2[kV64A2Bytes];
  alignas(8) uint64_t x_full[kV64Stages];
  uint64_t x_empty[kV64Stages];
  uint64_t y_full[kV64Stages];
  uint64_t y_empty[kV64Stages];
  uint64_t a_full;
  uint64_t s_full[2];
  uint64_t dp_full[2];
  uint64_t ds_ready[2];
  uint64_t acc_done;
  float col_lse[2][2][kBlockN];      // [group][use & 1][column of the tile]
  float col_delta[2][2][kBlockN];
  float col_bias[2][2][kBlockN];
  uint32_t tmem_base;
};
static_assert(offsetof(SharedStorageV64x2, x_tiles) == kA2Bytes, "epilogue staging spans a1 + the X ring");

// Gradient epilogue of one warp for 64 accumulator columns [col0, col0 + 64) of its 32 rows (two chunks of 32):
// TMEM -> registers -> scale -> conjugate axial rotation -> (bias-gradient column sums) -> staging -> TMA store.
// tbuf: the (cos, sin) pairs of chunk 0, loaded by the caller before it waited for the last MMA.
__device__ __forceinline__ void grad_epilogue64(const GradOut& g, const CUtensorMap* map, uint32_t stage, uint32_t lane_addr_acc,
                                                int col0, int lane, int row0, int La, int b, float scale, bool rotate,
                                                float2* tbuf) {
  const int row_in_batch = row0 + lane;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int cc = (col0 >> 5) + i;                    // 32-column chunk of the 256-wide row
    uint32_t o[32];
    SAM2B200_TMEM_LD32(lane_addr_acc + cc * 32, o);
    tmem_wait_ld();
    float v[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) v[k] = __uint_as_float(o[k]) * scale;
    if (rotate) {
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const float2 cs = tbuf[k];
        const float re = v[2 * k] * cs.x + v[2 * k + 1] * cs.y;      // multiply by conj(cos + i sin)
        const float im = v[2 * k + 1] * cs.x - v[2 * k] * cs.y;
        v[2 * k] = re; v[2 * k + 1] = im;
      }
    }
    if (i == 0) load_table_chunk(g, rotate, row_in_batch, (cc + 1) * 32, tbuf);   // chunk 1's pairs: in flight under the work below
    if (g.bias_grad != nullptr) {
      if (row_in_batch >= La) {
#pragma unroll
        for (int k = 0; k < 32; ++k) v[k] = 0.f;        // rows beyond the tensor are clipped by the store, not by the sum
      }
      const float cs = warp_column_sum32(v, lane);
      atomicAdd(g.bias_grad + cc * 32 + lane, cs);
    }
    if (g.is_bf16) {
      stage_chunk_bf16(stage, lane, i, v);                                       // one [32 rows x 64 cols] box
      if (i == 1) store_box(map, stage, lane, col0, row0, b, row0 < La);
    } else {
      stage_chunk_f32(stage, lane, i, v);                                        // two [32 rows x 32 cols] boxes
      store_box(map, stage + i * kBoxBytes, lane, col0 + i * 32, row0, b, row0 < La);
    }
  }
  if (lane == 0) tma_store_wait_read();
}

template <int MODE, bool DROP = false>
__global__ void __launch_bounds__(kX2Threads, 1)
three_gemm_v64x2_kernel(const __grid_constant__ CUtensorMap map_a2,   // [B, La, 64] bf16, box 64 x 128
                        const __grid_constant__ CUtensorMap map_x,    // [B, Lx, 256] bf16, box 64 x 64
                        const __grid_constant__ CUtensorMap map_y,    // [B, Lx, 64] bf16, box 64 x 64
                        const __grid_constant__ CUtensorMap map_a1,   // [B, La, 256] bf16, box 64 x 128
                        const __grid_constant__ CUtensorMap map_g,    // dQ / dK
                        const ThreeGemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  SharedStorageV64x2& sh = *reinterpret_cast<SharedStorageV64x2*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int a_tile = blockIdx.x;
  const int b = blockIdx.y;
  const int nt = (p.Lx + kBlockN - 1) / kBlockN;
  unsigned long long* dbg = p.dbg ? p.dbg + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 8 : nullptr;
  if (dbg != nullptr && threadIdx.x == 0) { dbg[0] = smid(); dbg[7] = nt; dbg[1] = gtimer(); }

  if (threadIdx.x == 0) {
    for (int s = 0; s < kV64Stages; ++s) {
      mbar_init(&sh.x_full[s], 1); mbar_init(&sh.x_empty[s], 1);
      mbar_init(&sh.y_full[s], 1); mbar_init(&sh.y_empty[s], 1);
    }
    mbar_init(&sh.a_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sh.s_full[i], 1); mbar_init(&sh.dp_full[i], 1); mbar_init(&sh.ds_ready[i], kX2GroupThreads);
    }
    mbar_init(&sh.acc_done, 1);
    fence_barrier_init();
  }
  if (warp == kX2ProducerWarp && lane == 0) { prefetch_tmap(&map_a1); prefetch_tmap(&map_a2); prefetch_tmap(&map_x); prefetch_tmap(&map_y); }
  if (warp == 0 && lane == 0) prefetch_tmap(&map_g);
  if (warp == kX2MmaWarp) { tmem_alloc(&sh.tmem_base, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sh.tmem_base;
  if (dbg != nullptr && threadIdx.x == 0) dbg[2] = gtimer();

  if (warp == kX2ProducerWarp) {
    // ===================== TMA producer =====================
    const bool leader = elect_one();
    if (leader) {
      mbar_arrive_expect_tx(&sh.a_full, kA2Bytes + kV64A2Bytes);
#pragma unroll
      for (int c = 0; c < 4; ++c)
        tma_load_3d(&sh.a1[c * kA2ChunkBytes], &map_a1, &sh.a_full, c * 64, a_tile * kBlockM, b);
      tma_load_3d(&sh.a2[0], &map_a2, &sh.a_full, 0, a_tile * kBlockM, b);
    }
    __syncwarp();
    for (int j = 0; j < nt; ++j) {
      const int s = j % kV64Stages;
      const uint32_t ph = (j / kV64Stages) & 1;
      const int row0 = j * kBlockN;
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
        mbar_arrive_expect_tx(&sh.y_full[s], kV64YBytes);
        tma_load_3d(&sh.y_tiles[s][0], &map_y, &sh.y_full[s], 0, row0, b);
      }
      __syncwarp();
    }
  } else if (warp == kX2MmaWarp) {
    // ===================== MMA issuer (one elected lane; the warp stays converged) =====================
    const bool leader = elect_one();
    constexpr uint32_t idesc_s = make_idesc_bf16(kBlockM, kBlockN, 0, 0);
    constexpr uint32_t idesc_acc = make_idesc_bf16(kBlockM, kD, 0, 1);
    const uint32_t a1_lo = desc_lo_sw128(smem_u32(&sh.a1[0]), 16);
    const uint32_t a2_lo = desc_lo_sw128(smem_u32(&sh.a2[0]), 16);
    const uint32_t x_lo0 = desc_lo_sw128(smem_u32(&sh.x_tiles[0][0]), 16);            // K-major view of X
    const uint32_t xm_lo0 = desc_lo_sw128(smem_u32(&sh.x_tiles[0][0]), kChunkBytes);  // MN-major view of X
    const uint32_t y_lo0 = desc_lo_sw128(smem_u32(&sh.y_tiles[0][0]), 16);
    auto issue_s_dp = [&](int t) {     // S[t] = A1 . X[t]^T  and  dP[t] = A2 . Y[t]^T into buffer t & 1
      const int s = t % kV64Stages;
      const uint32_t ph = (t / kV64Stages) & 1;
      mbar_wait(&sh.x_full[s], ph);
      mbar_wait(&sh.y_full[s], ph);
      tc_fence_after();
      if (leader) {
        const uint32_t xlo = x_lo0 + s * (kTileBytes >> 4);
        const uint32_t ylo = y_lo0 + s * (kV64YBytes >> 4);
        const uint32_t ds = tmem + ((t & 1) ? kVColS1 : kVColS0);
        const uint32_t dd = tmem + ((t & 1) ? kVColDP1 : kVColDP0);
#pragma unroll
        for (int ks = 0; ks < kD / 16; ++ks)
          umma_ss_lohi(ds, a1_lo + (ks >> 2) * (kA2ChunkBytes >> 4) + (ks & 3) * 2,
                       xlo + (ks >> 2) * (kChunkBytes >> 4) + (ks & 3) * 2, kDescHiSw128_1024, idesc_s, ks > 0);
        umma_commit(&sh.s_full[t & 1]);
#pragma unroll
        for (int ks = 0; ks < 64 / 16; ++ks)
          umma_ss_lohi(dd, a2_lo + ks * 2, ylo + ks * 2, kDescHiSw128_1024, idesc_s, ks > 0);
        umma_commit(&sh.y_empty[s]);
        umma_commit(&sh.dp_full[t & 1]);
      }
      __syncwarp();
    };
    mbar_wait(&sh.a_full, 0);
    tc_fence_after();
    if (dbg != nullptr && lane == 0) dbg[3] = gtimer();
    issue_s_dp(0);
    if (nt > 1) issue_s_dp(1);
    for (int j = 0; j < nt; ++j) {
      const int s = j % kV64Stages;
      mbar_wait(&sh.ds_ready[j & 1], (j >> 1) & 1);
      tc_fence_after();
      if (leader) {
        const uint32_t xlo = xm_lo0 + s * (kTileBytes >> 4);
        const uint32_t da = tmem + ((j & 1) ? kVColDP1 : kVColDP0);
#pragma unroll
        for (int ks = 0; ks < kBlockN / 16; ++ks)
          umma_ts_lohi(tmem + kVColAcc, da + p_col_of_kstep(ks), xlo + ks * (2048 >> 4), kDescHiSw128_1024, idesc_acc,
                       (j > 0) || (ks > 0));
        umma_commit(&sh.x_empty[s]);
        if (j + 1 >= nt) umma_commit(&sh.acc_done);
      }
      __syncwarp();
      if (j + 2 < nt) issue_s_dp(j + 2);   // overwrites S(j) (read before ds_ready(j)) and dS(j) (read by the MMAs just issued)
    }
  } else {
    // ===================== softmax groups (warps 0-7: even tiles, 8-15: odd tiles), then all 16 drain ACC ==========
    const int group = warp >> 3;
    const int quarter = warp & 3;                          // TMEM lane quarter (hardware: warp id mod 4)
    const int half = (warp >> 2) & 1;                      // which 32 of the tile's 64 columns
    const int gtid = threadIdx.x & (kX2GroupThreads - 1);
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = tmem + (uint32_t(quarter * 32) << 16);
    const long long a_row_idx = (long long)a_tile * kBlockM + row;
    const bool row_valid = a_row_idx < p.La;
    const float c = p.scale_log2;
    float row_lse = 0.f, row_delta = 0.f, row_bias = 0.f;
    if (MODE == MODE_DQ && row_valid) {
      row_lse = p.lse2[(long long)b * p.La + a_row_idx];
      row_delta = p.delta[(long long)b * p.La + a_row_idx];
      if (DROP) row_bias = p.dp_bias[(long long)b * p.La + a_row_idx];
    }
    // DK: the per-column (= per-query) vectors of this group's NEXT tile, fetched one tile ahead by 64 threads of the group
    float lse_next = INFINITY, delta_next = 0.f, bias_next = 0.f;
    if (MODE == MODE_DK && gtid < kBlockN) {
      const int col = group * kBlockN + gtid;
      if (group < nt && col < p.Lx) {
        lse_next = p.lse2[(long long)b * p.Lx + col];
        delta_next = p.delta[(long long)b * p.Lx + col];
        if (DROP) bias_next = p.dp_bias[(long long)b * p.Lx + col];
      }
    }
    const uint32_t drop_key = DROP ? sam2b200::dropout_key(*p.drop.seed, p.drop.site) : 0u;
    const uint32_t sbase = lane_addr + (group ? kVColS1 : kVColS0) + half * kHalfN;
    const uint32_t dbase = lane_addr + (group ? kVColDP1 : kVColDP0) + half * kHalfN;
    for (int j = group; j < nt; j += 2) {
      const int use = j >> 1;
      if (MODE == MODE_DK) {
        if (gtid < kBlockN) {
          sh.col_lse[group][use & 1][gtid] = lse_next;
          sh.col_delta[group][use & 1][gtid] = delta_next;
          if (DROP) sh.col_bias[group][use & 1][gtid] = bias_next;
          const int col = (j + 2) * kBlockN + gtid;
          const bool ok = (j + 2 < nt) && col < p.Lx;
          lse_next = ok ? p.lse2[(long long)b * p.Lx + col] : INFINITY;
          delta_next = ok ? p.delta[(long long)b * p.Lx + col] : 0.f;
          if (DROP) bias_next = ok ? p.dp_bias[(long long)b * p.Lx + col] : 0.f;
        }
        asm volatile("bar.sync %0, 256;" ::"r"(5 + group) : "memory");
      }
      mbar_wait(&sh.s_full[group], use & 1);
      mbar_wait(&sh.dp_full[group], use & 1);
      tc_fence_after();
      if (dbg != nullptr && j == 0 && threadIdx.x == 0) dbg[4] = gtimer();
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t sr[16], dr[16];
        SAM2B200_TMEM_LD16(sbase + h * 16, sr);
        SAM2B200_TMEM_LD16(dbase + h * 16, dr);
        tmem_wait_ld();
        const int col_in_tile = half * kHalfN + h * 16;
        const int ncols = p.Lx - j * kBlockN - col_in_tile;       // DQ: valid columns of this 16-chunk
        // element (query q, key k) has dropout index (b N + q) M + k: DQ rows are queries, DK rows are keys
        const uint32_t didx = (MODE == MODE_DQ)
            ? (uint32_t)(((long long)b * p.La + a_row_idx) * p.Lx) + (uint32_t)(j * kBlockN + col_in_tile)
            : (uint32_t)(((long long)b * p.Lx + (j * kBlockN + col_in_tile)) * p.La + a_row_idx);
        const uint32_t dstep = (MODE == MODE_DQ) ? 1u : (uint32_t)p.La;
        const float* cl = &sh.col_lse[group][use & 1][col_in_tile];
        const float* cd = &sh.col_delta[group][use & 1][col_in_tile];
        const float* cb = &sh.col_bias[group][use & 1][col_in_tile];
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
          float p0, p1;
          if (MODE == MODE_DQ) {
            p0 = (i < ncols) ? ex2(fmaf(__uint_as_float(sr[i]), c, -row_lse)) : 0.f;
            p1 = (i + 1 < ncols) ? ex2(fmaf(__uint_as_float(sr[i + 1]), c, -row_lse)) : 0.f;
          } else {
            p0 = ex2(fmaf(__uint_as_float(sr[i]), c, -cl[i]));
            p1 = ex2(fmaf(__uint_as_float(sr[i + 1]), c, -cl[i + 1]));
          }
          float d0 = __uint_as_float(dr[i]), d1 = __uint_as_float(dr[i + 1]);
          if (DROP) {
            const float cb0 = (MODE == MODE_DQ) ? row_bias : cb[i];
            const float cb1 = (MODE == MODE_DQ) ? row_bias : cb[i + 1];
            d0 = sam2b200::dropout_keep(drop_key, didx + (uint32_t)i * dstep, p.drop.thresh) ? (d0 + cb0) * p.drop.inv_keep : 0.f;
            d1 = sam2b200::dropout_keep(drop_key, didx + (uint32_t)(i + 1) * dstep, p.drop.thresh) ? (d1 + cb1) * p.drop.inv_keep : 0.f;
          }
          const float dl0 = (MODE == MODE_DQ) ? row_delta : cd[i];
          const float dl1 = (MODE == MODE_DQ) ? row_delta : cd[i + 1];
          pk[i >> 1] = pack_bf16(p0 * (d0 - dl0), p1 * (d1 - dl1));
        }
        // dS (bf16 pairs) over this warp's own dP columns: chunk h -> packed columns [8h, 8h + 8) (already read above)
        SAM2B200_TMEM_ST8(dbase + h * 8, pk);
      }
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(&sh.ds_ready[group]);
    }
    // ---------------- epilogue: sixteen warps, 32 rows x 64 columns each ----------------
    const int part = warp >> 2;                            // columns [64 part, 64 part + 64)
    const int row0 = a_tile * kBlockM + quarter * 32;
    const bool rotate = p.gout.rope_table != nullptr && (row0 + lane) < p.gout.rope_rows;
    const uint32_t stage = smem_u32(&sh.a1[0]) + warp * (2 * kBoxBytes);   // a1 + the ring behind it are idle by now
    float2 tbuf[16];
    load_table_chunk(p.gout, rotate, row0 + lane, part * 64, tbuf);       // in flight while the last MMAs drain
    mbar_wait(&sh.acc_done, 0);
    tc_fence_after();
    if (dbg != nullptr && threadIdx.x == 0) dbg[5] = gtimer();
    grad_epilogue64(p.gout, &map_g, stage, lane_addr + kVColAcc, part * 64, lane, row0, p.La, b, p.scale, rotate, tbuf);
  }

  tc_fence_before();
  __syncthreads();
  if (dbg != nullptr && threadIdx.x == 0) dbg[6] = gtimer();
  if (warp == kX2MmaWarp) tmem_dealloc(tmem, 512);
}

}  // namespace attn
