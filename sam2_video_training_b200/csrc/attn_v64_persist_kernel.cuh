// PERSISTENT version of the raw-memory cross-attention backward kernels (three_gemm_v64_kernel, attn_v64_kernels.cuh).
//
// With N = 576 queries a dK CTA runs only 9 tiles: of its 16.8 us (profiles/r2_timeline_v64_cfg2_axial_table.txt,
// r2_timeline_epilogue.txt) the loop takes 9.7, the epilogue 2.4 and the rest is per-CTA fixed cost that has nothing to do
// with the item: launch gap 0.9, barrier / TMEM set-up 0.3, operand fetch (A1 64 KB + A2 16 KB) 1.4, first tile + first
// S GEMM 0.8.  One resident CTA per SM that walks over (row block, object) items removes the first two and hides the other two
// behind the previous item's epilogue -- which works here (and did not for the 256-d kernels, profiles/r1_persistent_key_side.txt)
// because both fixed operands live in SHARED memory (SS-mode MMAs), so nothing has to pass through the softmax warps:
//   * the epilogue stages its output tiles in two stages of the (idle) tile ring instead of the A1 buffer, so the producer
//     re-loads A1 / A2 for the next item the moment the last MMA of this item has completed (acc_done),
//   * the ring stage the staging does not use receives the next item's tile 0 at the same time, the other two stages follow when
//     the TMA stores have read the staging (stage_free),
//   * the MMA warp issues S / dP of the next item's tiles 0 and 1 while the softmax warps are still draining the accumulator;
//     the first accumulate MMA of the next item waits for ds_ready of its tile 0, which every softmax thread signals only after
//     it has left the epilogue (program order), so the accumulator is never overwritten early.
// Tile counters run across items (g = it * nt + j): ring stage g % 3, S / dP / column-vector buffer g & 1.
// Numerics: the same instructions on the same data as three_gemm_v64_kernel -- results are bit-identical.
// Requirements (host): bf16 gradients (8 KB of staging per warp), nt >= 3.
#pragma once

#include "attn_v64_kernels.cuh"

namespace attn {

struct SharedStorageV64P {
  alignas(1024) uint8_t a1[kA2Bytes];
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
  uint64_t stage_free;       // the epilogue's TMA stores have read the staging stages
  float col_lse[2][kBlockN];
  float col_delta[2][kBlockN];
  float col_bias[2][kBlockN];
  uint32_t tmem_base;
};

template <int MODE, bool DROP = false>
__global__ void __launch_bounds__(kThreads, 1)
three_gemm_v64_persistent_kernel(const __grid_constant__ CUtensorMap map_a2,   // [B, La, 64] bf16, box 64 x 128
                                 const __grid_constant__ CUtensorMap map_x,    // [B, Lx, 256] bf16, box 64 x 64
                                 const __grid_constant__ CUtensorMap map_y,    // [B, Lx, 64] bf16, box 64 x 64
                                 const __grid_constant__ CUtensorMap map_a1,   // [B, La, 256] bf16, box 64 x 128
                                 const __grid_constant__ CUtensorMap map_g,    // dQ / dK (bf16)
                                 const ThreeGemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  SharedStorageV64P& sh = *reinterpret_cast<SharedStorageV64P*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nt = (p.Lx + kBlockN - 1) / kBlockN;
  const int n_my = (p.n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // items blockIdx.x, + gridDim.x, ...

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
    mbar_init(&sh.stage_free, kNumSoftmaxThreads);
    fence_barrier_init();
  }
  if (warp == kProducerWarp && lane == 0) { prefetch_tmap(&map_a1); prefetch_tmap(&map_a2); prefetch_tmap(&map_x); prefetch_tmap(&map_y); }
  if (warp == 0 && lane == 0) prefetch_tmap(&map_g);
  if (warp == kMmaWarp) { tmem_alloc(&sh.tmem_base, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sh.tmem_base;

  if (warp == kProducerWarp) {
    const bool leader = elect_one();
    for (int it = 0; it < n_my; ++it) {
      const int item = (int)blockIdx.x + it * (int)gridDim.x;
      const int a_tile = item % p.n_atiles, b = item / p.n_atiles;
      if (it > 0) mbar_wait(&sh.acc_done, (it - 1) & 1);     // every MMA of the previous item has completed: A1 / A2 are free
      if (leader) {
        mbar_arrive_expect_tx(&sh.a_full, kA2Bytes + kV64A2Bytes);
#pragma unroll
        for (int c = 0; c < 4; ++c)
          tma_load_3d(&sh.a1[c * kA2ChunkBytes], &map_a1, &sh.a_full, c * 64, a_tile * kBlockM, b);
        tma_load_3d(&sh.a2[0], &map_a2, &sh.a_full, 0, a_tile * kBlockM, b);
      }
      __syncwarp();
      for (int j = 0; j < nt; ++j) {
        const int g = it * nt + j;
        const int s = g % kV64Stages;
        const uint32_t ph = (g / kV64Stages) & 1;
        const int row0 = j * kBlockN;
        mbar_wait(&sh.x_empty[s], ph ^ 1);
        // tiles 1 and 2 of an item land in the stages the previous item's epilogue used as staging
        if (it > 0 && (j == 1 || j == 2)) mbar_wait(&sh.stage_free, (it - 1) & 1);
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
    auto issue_s_dp = [&](int g) {     // S = A1 . X^T  and  dP = A2 . Y^T of global tile g into buffer g & 1
      const int s = g % kV64Stages;
      const uint32_t ph = (g / kV64Stages) & 1;
      mbar_wait(&sh.x_full[s], ph);
      mbar_wait(&sh.y_full[s], ph);
      tc_fence_after();
      if (leader) {
        const uint32_t xlo = x_lo0 + s * (kTileBytes >> 4);
        const uint32_t ylo = y_lo0 + s * (kV64YBytes >> 4);
        const uint32_t ds = tmem + ((g & 1) ? kVColS1 : kVColS0);
        const uint32_t dd = tmem + ((g & 1) ? kVColDP1 : kVColDP0);
#pragma unroll
        for (int ks = 0; ks < kD / 16; ++ks)
          umma_ss_lohi(ds, a1_lo + (ks >> 2) * (kA2ChunkBytes >> 4) + (ks & 3) * 2,
                       xlo + (ks >> 2) * (kChunkBytes >> 4) + (ks & 3) * 2, kDescHiSw128_1024, idesc_s, ks > 0);
        umma_commit(&sh.s_full[g & 1]);
#pragma unroll
        for (int ks = 0; ks < 64 / 16; ++ks)
          umma_ss_lohi(dd, a2_lo + ks * 2, ylo + ks * 2, kDescHiSw128_1024, idesc_s, ks > 0);
        umma_commit(&sh.y_empty[s]);
        umma_commit(&sh.dp_full[g & 1]);
      }
      __syncwarp();
    };
    for (int it = 0; it < n_my; ++it) {
      const int g0 = it * nt;
      mbar_wait(&sh.a_full, it & 1);
      tc_fence_after();
      issue_s_dp(g0);
      issue_s_dp(g0 + 1);
      for (int j = 0; j < nt; ++j) {
        const int g = g0 + j;
        const int s = g % kV64Stages;
        mbar_wait(&sh.ds_ready[g & 1], (g >> 1) & 1);
        tc_fence_after();
        if (leader) {
          const uint32_t xlo = xm_lo0 + s * (kTileBytes >> 4);
          const uint32_t da = tmem + ((g & 1) ? kVColDP1 : kVColDP0);
#pragma unroll
          for (int ks = 0; ks < kBlockN / 16; ++ks)
            umma_ts_lohi(tmem + kVColAcc, da + p_col_of_kstep(ks), xlo + ks * (2048 >> 4), kDescHiSw128_1024, idesc_acc,
                         (j > 0) || (ks > 0));
          umma_commit(&sh.x_empty[s]);
          if (j + 1 >= nt) umma_commit(&sh.acc_done);
        }
        __syncwarp();
        if (j + 2 < nt) issue_s_dp(g + 2);
      }
    }
  } else {
    const int quarter = warp & 3;
    const int half = warp >> 2;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = tmem + (uint32_t(quarter * 32) << 16);
    const float c = p.scale_log2;
    const uint32_t drop_key = DROP ? sam2b200::dropout_key(*p.drop.seed, p.drop.site) : 0u;
    // rotation table in shared memory: the X part once, the Y part of each item's rows into one of two buffers (item parity)
    const bool tab_on = p.gout.rope_smem > 0;
    const uint32_t tab_area = smem_u32(&sh) + (uint32_t)sizeof(SharedStorageV64P);
    const uint32_t tab_y0 = tab_area + p.gout.rope_w * kRopeXStride;
    if (tab_on) rope_stage_x(p.gout, tab_area, threadIdx.x);
    for (int it = 0; it < n_my; ++it) {
      const int item = (int)blockIdx.x + it * (int)gridDim.x;
      const int a_tile = item % p.n_atiles, b = item / p.n_atiles;
      const int g0 = it * nt;
      const long long a_row_idx = (long long)a_tile * kBlockM + row;
      const bool row_valid = a_row_idx < p.La;
      const int row0 = a_tile * kBlockM + quarter * 32;
      const bool rotate = p.gout.rope_table != nullptr && (row0 + lane) < p.gout.rope_rows;
      uint32_t tab = 0;
      if (tab_on) {
        const uint32_t ybase = tab_y0 + (it & 1) * rope_y_bytes(p.gout.rope_w);
        rope_stage_y(p.gout, ybase, a_tile * kBlockM, threadIdx.x);
        cp_async_commit();
        tab = rope_tab_addr(p.gout, tab_area, ybase, a_tile * kBlockM, row, half);
      }
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
      for (int j = 0; j < nt; ++j) {
        const int g = g0 + j;
        const int cb = g & 1;
        if (MODE == MODE_DK) {
          if (threadIdx.x < kBlockN) {
            sh.col_lse[cb][threadIdx.x] = lse_next;
            sh.col_delta[cb][threadIdx.x] = delta_next;
            if (DROP) sh.col_bias[cb][threadIdx.x] = bias_next;
            const int col = (j + 1) * kBlockN + threadIdx.x;
            const bool ok = (j + 1 < nt) && col < p.Lx;
            lse_next = ok ? p.lse2[(long long)b * p.Lx + col] : INFINITY;
            delta_next = ok ? p.delta[(long long)b * p.Lx + col] : 0.f;
            if (DROP) bias_next = ok ? p.dp_bias[(long long)b * p.Lx + col] : 0.f;
          }
          asm volatile("bar.sync 5, 256;" ::: "memory");
        }
        const uint32_t sbuf = lane_addr + (cb ? kVColS1 : kVColS0) + half * kHalfN;
        const uint32_t dbuf = lane_addr + (cb ? kVColDP1 : kVColDP0) + half * kHalfN;
        mbar_wait(&sh.s_full[cb], (g >> 1) & 1);
        tc_fence_after();
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
            else pv[i] = ex2(fmaf(sraw, c, -sh.col_lse[cb][half * kHalfN + i]));
          }
        }
        mbar_wait(&sh.dp_full[cb], (g >> 1) & 1);
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
            const float dl0 = (MODE == MODE_DQ) ? row_delta : sh.col_delta[cb][half * kHalfN + i];
            const float dl1 = (MODE == MODE_DQ) ? row_delta : sh.col_delta[cb][half * kHalfN + i + 1];
            float d0 = __uint_as_float(r0[i]), d1 = __uint_as_float(r0[i + 1]);
            if (DROP) {
              const float cb0 = (MODE == MODE_DQ) ? row_bias : sh.col_bias[cb][half * kHalfN + i];
              const float cb1 = (MODE == MODE_DQ) ? row_bias : sh.col_bias[cb][half * kHalfN + i + 1];
              d0 = sam2b200::dropout_keep(drop_key, didx + (uint32_t)i * dstep, p.drop.thresh) ? (d0 + cb0) * p.drop.inv_keep : 0.f;
              d1 = sam2b200::dropout_keep(drop_key, didx + (uint32_t)(i + 1) * dstep, p.drop.thresh) ? (d1 + cb1) * p.drop.inv_keep : 0.f;
            }
            pk[i >> 1] = pack_bf16(pv[i] * (d0 - dl0), pv[i + 1] * (d1 - dl1));
          }
        }
        SAM2B200_TMEM_ST16(dbuf, pk);          // dS over this warp's own dP columns
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(&sh.ds_ready[cb]);
      }
      // ---- epilogue: staging in the two ring stages the NEXT tile (global index g0 + nt) does not use
      const int s_next = (g0 + nt) % kV64Stages;
      const int s_stage = (s_next + 1 + (warp >> 2)) % kV64Stages;          // warps 0-3 -> s_next + 1, warps 4-7 -> s_next + 2
      const uint32_t stage = smem_u32(&sh.x_tiles[s_stage][0]) + (warp & 3) * (2 * kBoxBytes);
      float2 tcur[16];
      if (tab == 0) load_table_chunk(p.gout, rotate, row0 + lane, half * 128, tcur);
      else { cp_async_wait_all(); asm volatile("bar.sync 6, 256;" ::: "memory"); }       // the staged table is complete and visible
      mbar_wait(&sh.acc_done, it & 1);
      tc_fence_after();
      grad_epilogue(p.gout, &map_g, stage, lane_addr + kVColAcc, half, lane, row0, p.La, b, p.scale, rotate, tcur, nullptr, tab);
      __syncwarp();                        // lane 0 has waited for the TMA stores to finish reading the staging
      tc_fence_before();                   // this thread's accumulator reads are ordered before the arrivals that let the next item's MMAs start
      mbar_arrive(&sh.stage_free);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem, 512);
}

}  // namespace attn
