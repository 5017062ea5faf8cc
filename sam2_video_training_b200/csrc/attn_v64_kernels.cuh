// Backward kernels of the raw-memory cross-attention (sam2b200_attn_bwd_v64) with DOUBLE-BUFFERED score / dP tiles.
//
// With the dP GEMM at K = 64 the generic three_gemm_kernel is no longer limited by the tensor pipe but by its
// single-buffered chain  S MMA -> softmax reads S -> ... -> dS written -> accumulate MMA -> dP MMA  (four barrier hops per
// tile: 1.55 us per tile for ~1.0 us of MMA work).  Here the TMEM-resident operand A1 moves to shared memory (the narrow
// A2 / Y operands freed 72 KB of it) and the 128 freed TMEM columns double-buffer S and dP:
//     TMEM   ACC 256 | S0 64 | S1 64 | dP0 64 | dP1 64
//     SMEM   A1 64 KB (4 K-major slabs) | A2 16 KB | 3-stage ring of X tiles (32 KB) + Y tiles (8 KB)
// S(j+2) and dP(j+2) are issued right after the accumulate MMA of tile j, so the MMAs of the next two tiles run while the
// softmax warps work on tile j: per tile the softmax warps wait only for data that has been ready for a whole tile time.
// tcgen05.mma executes in issue order, which is what lets dP(j+2) overwrite the buffer the accumulate MMA of tile j
// (issued just before) reads dS from, and S(j+2) overwrite S(j) once ds_ready(j) has been observed.
//   MODE_DQ: A1 = Q block, A2 = dO' = dO Wv block, X = K tiles, Y = memory tiles   (LSE2 / Delta per row)
//   MODE_DK: A1 = K block, A2 = memory block,    X = Q tiles, Y = dO' tiles        (LSE2 / Delta per column)
// Numerics: the same products in the same order as three_gemm_kernel<MODE, false, 64>.
#pragma once

#include "attn_kernels.cuh"

namespace attn {

constexpr int kV64Stages = 3;
constexpr int kV64YBytes = kBlockN * 64 * 2;        // 8 KB: one 64-column slab of a streamed tile
constexpr int kV64A2Bytes = kBlockM * 64 * 2;       // 16 KB
constexpr uint32_t kVColAcc = 0;
constexpr uint32_t kVColS0 = 256, kVColS1 = 320, kVColDP0 = 384, kVColDP1 = 448;

struct SharedStorageV64 {
  alignas(1024) uint8_t a1[kA2Bytes];                         // 64 KB; epilogue staging once the MMAs are done
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
  float col_lse[2][kBlockN];
  float col_delta[2][kBlockN];
  float col_bias[2][kBlockN];
  uint32_t tmem_base;
};

// DROP: attention-probability dropout (compile-time).  dP = keep ? (dO' mem^T + dO . bv) / (1 - p) : 0 -- under dropout the
// rows of the probability matrix no longer sum to 1, so the per-query constant dO . bv (p.dp_bias) does not cancel.
template <int MODE, bool DROP = false>
__global__ void __launch_bounds__(kThreads, 1)
three_gemm_v64_kernel(const __grid_constant__ CUtensorMap map_a2,   // [B, La, 64] bf16, box 64 x 128
                      const __grid_constant__ CUtensorMap map_x,    // [B, Lx, 256] bf16, box 64 x 64
                      const __grid_constant__ CUtensorMap map_y,    // [B, Lx, 64] bf16, box 64 x 64
                      const __grid_constant__ CUtensorMap map_a1,   // [B, La, 256] bf16, box 64 x 128
                      const __grid_constant__ CUtensorMap map_g,    // dQ / dK
                      const ThreeGemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  SharedStorageV64& sh = *reinterpret_cast<SharedStorageV64*>(
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
    for (int s = 0; s < kV64Stages; ++s) {
      mbar_init(&sh.x_full[s], 1); mbar_init(&sh.x_empty[s], 1);
      mbar_init(&sh.y_full[s], 1); mbar_init(&sh.y_empty[s], 1);
    }
    mbar_init(&sh.a_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sh.s_full[i], 1); mbar_init(&sh.dp_full[i], 1); mbar_init(&sh.ds_ready[i], kNumSoftmaxThreads);
    }
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
#ifndef SAM2B200_EPI_TIMELINE
  SAM2B200_STAMP(p.dbg, 2);
#endif

  if (warp == kProducerWarp) {
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
  } else if (warp == kMmaWarp) {
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
#ifndef SAM2B200_EPI_TIMELINE
    if (p.dbg != nullptr && lane == 0)
      p.dbg[((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 8 + 3] = gtimer();
#endif
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
    const int quarter = warp & 3;
    const int half = warp >> 2;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = tmem + (uint32_t(quarter * 32) << 16);
    const long long a_row_idx = (long long)a_tile * kBlockM + row;
    const bool row_valid = a_row_idx < p.La;
    const int row0 = a_tile * kBlockM + quarter * 32;
    uint32_t tab = 0;
    if (p.gout.rope_smem > 0) {     // rotation table of this CTA's rows -> shared memory, under the operand fetch
      const uint32_t area = smem_u32(&sh) + (uint32_t)sizeof(SharedStorageV64);
      const uint32_t ybase = area + p.gout.rope_w * kRopeXStride;
      rope_stage_x(p.gout, area, threadIdx.x);
      rope_stage_y(p.gout, ybase, a_tile * kBlockM, threadIdx.x);
      cp_async_commit();
      tab = rope_tab_addr(p.gout, area, ybase, a_tile * kBlockM, row, half);
    }
    const uint32_t stage = smem_u32(&sh.a1[0]) + warp * (4 * kBoxBytes);   // a1 + the ring behind it are idle at the epilogue
    const bool rotate = p.gout.rope_table != nullptr && (row0 + lane) < p.gout.rope_rows;
    const float c = p.scale_log2;
    float row_lse = 0.f, row_delta = 0.f, row_bias = 0.f;
    if (MODE == MODE_DQ && row_valid) {
      row_lse = p.lse2[(long long)b * p.La + a_row_idx];
      row_delta = p.delta[(long long)b * p.La + a_row_idx];
      if (DROP) row_bias = p.dp_bias[(long long)b * p.La + a_row_idx];
    }
    float lse_next = INFINITY, delta_next = 0.f, bias_next = 0.f;
    if (MODE == MODE_DK && threadIdx.x < kBlockN && (int)threadIdx.x < p.Lx) {
      lse_next = p.lse2[(long long)b * p.Lx + threadIdx.x];
      delta_next = p.delta[(long long)b * p.Lx + threadIdx.x];
      if (DROP) bias_next = p.dp_bias[(long long)b * p.Lx + threadIdx.x];
    }
    const uint32_t drop_key = DROP ? sam2b200::dropout_key(*p.drop.seed, p.drop.site) : 0u;
    for (int j = 0; j < nt; ++j) {
      if (MODE == MODE_DK) {
        if (threadIdx.x < kBlockN) {
          sh.col_lse[j & 1][threadIdx.x] = lse_next;
          sh.col_delta[j & 1][threadIdx.x] = delta_next;
          if (DROP) sh.col_bias[j & 1][threadIdx.x] = bias_next;
          const int col = (j + 1) * kBlockN + threadIdx.x;
          const bool ok = (j + 1 < nt) && col < p.Lx;
          lse_next = ok ? p.lse2[(long long)b * p.Lx + col] : INFINITY;
          delta_next = ok ? p.delta[(long long)b * p.Lx + col] : 0.f;
          if (DROP) bias_next = ok ? p.dp_bias[(long long)b * p.Lx + col] : 0.f;
        }
        asm volatile("bar.sync 5, 256;" ::: "memory");
      }
      const uint32_t sbuf = lane_addr + ((j & 1) ? kVColS1 : kVColS0) + half * kHalfN;
      const uint32_t dbuf = lane_addr + ((j & 1) ? kVColDP1 : kVColDP0) + half * kHalfN;
      mbar_wait(&sh.s_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
#ifndef SAM2B200_EPI_TIMELINE
      if (j == 0) SAM2B200_STAMP(p.dbg, 4);
#endif
      float pv[kHalfN];
      {
        uint32_t r0[32];
        SAM2B200_TMEM_LD32(sbuf, r0);
        tmem_wait_ld();
        const int ncols = p.Lx - j * kBlockN - half * kHalfN;
#pragma unroll
        for (int i = 0; i < kHalfN; ++i) {
          const float sraw = __uint_as_float(r0[i]);
          if (MODE == MODE_DQ) pv[i] = (i < ncols) ? ex2(fmaf(sraw, c, -row_lse)) : 0.f;
          else pv[i] = ex2(fmaf(sraw, c, -sh.col_lse[j & 1][half * kHalfN + i]));
        }
      }
      mbar_wait(&sh.dp_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      uint32_t pk[16];
      {
        uint32_t r0[32];
        SAM2B200_TMEM_LD32(dbuf, r0);
        tmem_wait_ld();
        // element (query q, key k) has dropout index (b N + q) M + k: DQ rows are queries, DK rows are keys
        const uint32_t didx = (MODE == MODE_DQ)
            ? (uint32_t)(((long long)b * p.La + a_row_idx) * p.Lx) + (uint32_t)(j * kBlockN + half * kHalfN)
            : (uint32_t)(((long long)b * p.Lx + (j * kBlockN + half * kHalfN)) * p.La + a_row_idx);
        const uint32_t dstep = (MODE == MODE_DQ) ? 1u : (uint32_t)p.La;
#pragma unroll
        for (int i = 0; i < kHalfN; i += 2) {
          const float dl0 = (MODE == MODE_DQ) ? row_delta : sh.col_delta[j & 1][half * kHalfN + i];
          const float dl1 = (MODE == MODE_DQ) ? row_delta : sh.col_delta[j & 1][half * kHalfN + i + 1];
          float d0 = __uint_as_float(r0[i]), d1 = __uint_as_float(r0[i + 1]);
          if (DROP) {
            const float cb0 = (MODE == MODE_DQ) ? row_bias : sh.col_bias[j & 1][half * kHalfN + i];
            const float cb1 = (MODE == MODE_DQ) ? row_bias : sh.col_bias[j & 1][half * kHalfN + i + 1];
            d0 = sam2b200::dropout_keep(drop_key, didx + (uint32_t)i * dstep, p.drop.thresh) ? (d0 + cb0) * p.drop.inv_keep : 0.f;
            d1 = sam2b200::dropout_keep(drop_key, didx + (uint32_t)(i + 1) * dstep, p.drop.thresh) ? (d1 + cb1) * p.drop.inv_keep : 0.f;
          }
          pk[i >> 1] = pack_bf16(pv[i] * (d0 - dl0), pv[i + 1] * (d1 - dl1));
        }
      }
      SAM2B200_TMEM_ST16(dbuf, pk);          // dS over this warp's own dP columns
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(&sh.ds_ready[j & 1]);
    }
    float2 tcur[16];
    if (tab == 0) load_table_chunk(p.gout, rotate, row0 + lane, half * 128, tcur);
    else { cp_async_wait_all(); asm volatile("bar.sync 6, 256;" ::: "memory"); }       // the staged table is complete and visible
    mbar_wait(&sh.acc_done, 0);
    tc_fence_after();
    SAM2B200_STAMP(p.dbg, 5);
    grad_epilogue(p.gout, &map_g, stage, lane_addr + kVColAcc, half, lane, row0, p.La, b, p.scale, rotate, tcur, p.dbg, tab);
  }

  tc_fence_before();
  __syncthreads();
  SAM2B200_STAMP(p.dbg, 6);
  if (warp == kMmaWarp) tmem_dealloc(tmem, 512);
}

}  // namespace attn
