// Persistent key-side backward kernels for SHORT query loops (N < 1024: cfg1 / cfg2, 9 tiles of 64 queries per key block).
//
// With one CTA per (key block, object) the dV kernel spends 44 % of a CTA's life outside its main loop at N = 576
// (profiles: setup 0.26 us, fixed operand -> TMEM 1.95 us, first scores 0.54 us, epilogue 2.66 us, inter-CTA gap 0.8 us
// against an 8.06 us main loop).  Here a CTA stays resident and walks a static list of (key block, object) items.  The
// shared-memory ring is addressed by a running SLOT counter and every item owns nt + 2 consecutive slots:
//     [ fixed operand A | tile 0 | ... | tile nt-1 | epilogue staging ]
// so the producer warp streams the NEXT item's key block (and its first tile) into the ring while the softmax warps
// are still draining the current accumulator through the staging slot -- barrier set-up, TMEM allocation, tensor-map
// fetch and the TMA latency of the fixed operand are paid once per CTA instead of once per item.
// Every slot completes exactly one phase of x_full / y_full / x_empty / y_empty of its ring stage, whatever its kind:
//   A slot     full <- TMA (slabs 0,1 -> X buffer, slabs 2,3 -> Y buffer), empty <- relayed by the MMA warp once all
//              softmax threads have moved the operand to tensor memory (a_ready);
//   tile slot  full <- TMA, empty <- tcgen05.commit of the MMAs that read it (as in the one-shot kernels);
//   EPI slot   full <- plain arrive of the producer (the stage is free), empty <- relayed by the MMA warp at the next
//              item's a_ready (every lane 0 has passed cp.async.bulk.wait_group.read by then).
// Numerics are identical to two_gemm_kernel<MODE_DV> (same tiles, same instruction order per item).
#pragma once

#include "attn_kernels.cuh"

namespace attn {

__device__ __forceinline__ int ring_stage(int r) { return r % kStages; }
__device__ __forceinline__ uint32_t ring_parity(int r) { return (uint32_t)(r / kStages) & 1u; }

// grid: min(n_items, #SMs) CTAs; item = blockIdx.x + it * gridDim.x -> (key block = item % n_atiles, object = item / n_atiles)
template <bool DROP>
__global__ void __launch_bounds__(kThreads, 1)
dv_persistent_kernel(const __grid_constant__ CUtensorMap map_x,    // Q  [B, N, 256] bf16, box 64 x 64
                     const __grid_constant__ CUtensorMap map_y,    // dO [B, N, 256] bf16, box 64 x 64
                     const __grid_constant__ CUtensorMap map_a,    // K  [B, M, 256] bf16, box 64 x 128
                     const __grid_constant__ CUtensorMap map_o,    // dV bf16, box 64 x 32
                     const TwoGemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  SharedStorage& sh = *reinterpret_cast<SharedStorage*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nt = (p.Lx + kBlockN - 1) / kBlockN;
  const int slots = nt + 2;
  const int n_my = (p.n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&sh.x_full[s], 1); mbar_init(&sh.x_empty[s], 1);
      mbar_init(&sh.y_full[s], 1); mbar_init(&sh.y_empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) { mbar_init(&sh.s_full[i], 1); mbar_init(&sh.p_ready[i], kNumSoftmaxThreads); }
    mbar_init(&sh.acc_done, 1);
    mbar_init(&sh.a_full, 1);
    mbar_init(&sh.a_ready, kNumSoftmaxThreads);
    fence_barrier_init();
  }
  if (warp == kProducerWarp && lane == 0) { prefetch_tmap(&map_a); prefetch_tmap(&map_x); prefetch_tmap(&map_y); }
  if (warp == 0 && lane == 0) prefetch_tmap(&map_o);
  if (warp == kMmaWarp) { tmem_alloc(&sh.tmem_base, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sh.tmem_base;

  if (warp == kProducerWarp) {
    // ===================== TMA producer =====================
    const bool leader = elect_one();
    for (int it = 0; it < n_my; ++it) {
      const int item = (int)blockIdx.x + it * (int)gridDim.x;
      const int a_tile = item % p.n_atiles, b = item / p.n_atiles;
      int r = it * slots;
      {   // A slot: the key block, two slabs per buffer
        const int s = ring_stage(r);
        const uint32_t ph = ring_parity(r);
        mbar_wait(&sh.x_empty[s], ph ^ 1);
        mbar_wait(&sh.y_empty[s], ph ^ 1);
        if (leader) {
          mbar_arrive_expect_tx(&sh.x_full[s], 2 * kSlabBytes);
          tma_load_3d(&sh.x_tiles[s][0], &map_a, &sh.x_full[s], 0, a_tile * kBlockM, b);
          tma_load_3d(&sh.x_tiles[s][kSlabBytes], &map_a, &sh.x_full[s], 64, a_tile * kBlockM, b);
          mbar_arrive_expect_tx(&sh.y_full[s], 2 * kSlabBytes);
          tma_load_3d(&sh.y_tiles[s][0], &map_a, &sh.y_full[s], 128, a_tile * kBlockM, b);
          tma_load_3d(&sh.y_tiles[s][kSlabBytes], &map_a, &sh.y_full[s], 192, a_tile * kBlockM, b);
        }
        __syncwarp();
      }
      for (int j = 0; j < nt; ++j) {
        r = it * slots + 1 + j;
        const int s = ring_stage(r);
        const uint32_t ph = ring_parity(r);
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
          mbar_arrive_expect_tx(&sh.y_full[s], kTileBytes);
#pragma unroll
          for (int c = 0; c < 4; ++c)
            tma_load_3d(&sh.y_tiles[s][c * kChunkBytes], &map_y, &sh.y_full[s], c * 64, row0, b);
        }
        __syncwarp();
      }
      {   // EPI slot: hand a free stage to the epilogue warps
        r = it * slots + nt + 1;
        const int s = ring_stage(r);
        const uint32_t ph = ring_parity(r);
        mbar_wait(&sh.x_empty[s], ph ^ 1);
        mbar_wait(&sh.y_empty[s], ph ^ 1);
        if (leader) { mbar_arrive(&sh.x_full[s]); mbar_arrive(&sh.y_full[s]); }
        __syncwarp();
      }
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer =====================
    const bool leader = elect_one();
    constexpr uint32_t idesc_s = make_idesc_bf16(kBlockM, kBlockN, 0, 0);
    constexpr uint32_t idesc_acc = make_idesc_bf16(kBlockM, kD, 0, 1);
    const uint32_t x_lo0 = desc_lo_sw128(smem_u32(&sh.x_tiles[0][0]), 16);
    const uint32_t y_lo0 = desc_lo_sw128(smem_u32(&sh.y_tiles[0][0]), kChunkBytes);
    int tc0 = 0;
    for (int it = 0; it < n_my; ++it) {
      const int rA = it * slots;
      auto issue_scores = [&](int j) {
        const int r = rA + 1 + j;
        const int s = ring_stage(r);
        const int tc = tc0 + j;
        mbar_wait(&sh.x_full[s], ring_parity(r));
        tc_fence_after();
        if (leader) {
          const uint32_t xlo = x_lo0 + s * (kTileBytes >> 4);
          const uint32_t d = tmem + ((tc & 1) ? kColS1 : kColS0);
#pragma unroll
          for (int ks = 0; ks < kD / 16; ++ks)
            umma_ts_lohi(d, tmem + kColA + ks * 8, xlo + (ks >> 2) * (kChunkBytes >> 4) + (ks & 3) * 2,
                         kDescHiSw128_1024, idesc_s, ks > 0);
          umma_commit(&sh.x_empty[s]);
          umma_commit(&sh.s_full[tc & 1]);
        }
        __syncwarp();
      };
      mbar_wait(&sh.a_ready, it & 1);
      tc_fence_after();
      if (leader) {   // release the A slot of this item and the EPI slot of the previous one
        mbar_arrive(&sh.x_empty[ring_stage(rA)]);
        mbar_arrive(&sh.y_empty[ring_stage(rA)]);
        if (it > 0) {
          mbar_arrive(&sh.x_empty[ring_stage(rA - 1)]);
          mbar_arrive(&sh.y_empty[ring_stage(rA - 1)]);
        }
      }
      __syncwarp();
      issue_scores(0);
      if (nt > 1) issue_scores(1);
      for (int j = 0; j < nt; ++j) {
        const int r = rA + 1 + j;
        const int s = ring_stage(r);
        const int tc = tc0 + j;
        mbar_wait(&sh.p_ready[tc & 1], (tc >> 1) & 1);
        mbar_wait(&sh.y_full[s], ring_parity(r));
        tc_fence_after();
        if (leader) {
          const uint32_t ylo = y_lo0 + s * (kTileBytes >> 4);
          const uint32_t pa = tmem + ((tc & 1) ? kColS1 : kColS0);
#pragma unroll
          for (int ks = 0; ks < kBlockN / 16; ++ks)
            umma_ts_lohi(tmem + kColAcc, pa + p_col_of_kstep(ks), ylo + ks * (2048 >> 4), kDescHiSw128_1024, idesc_acc,
                         (j > 0) || (ks > 0));
          umma_commit(&sh.y_empty[s]);
          umma_commit(&sh.acc_done);
        }
        __syncwarp();
        if (j + 2 < nt) issue_scores(j + 2);
      }
      tc0 += nt;
    }
  } else {
    // ===================== softmax / epilogue warps (0..7) =====================
    const int quarter = warp & 3;
    const int half = warp >> 2;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = tmem + (uint32_t(quarter * 32) << 16);
    const float c = p.scale_log2;
    constexpr bool drop_on = DROP;
    const uint32_t drop_key = drop_on ? sam2b200::dropout_key(*p.drop.seed, p.drop.site) : 0u;
    float2 tcur[16];
    int tc0 = 0;
    for (int it = 0; it < n_my; ++it) {
      const int item = (int)blockIdx.x + it * (int)gridDim.x;
      const int a_tile = item % p.n_atiles, b = item / p.n_atiles;
      const int rA = it * slots, rE = rA + nt + 1;
      const int sA = ring_stage(rA), sE = ring_stage(rE);
      const long long a_row_idx = (long long)a_tile * kBlockM + row;
      // first column vector (LSE2 of the queries of tile 0) before anything blocks
      float cv_next = INFINITY;
      if (threadIdx.x < kBlockN) cv_next = (threadIdx.x < p.Lx) ? p.lse2[(long long)b * p.Lx + threadIdx.x] : INFINITY;
      // fixed operand: shared memory -> tensor memory
      mbar_wait(half ? &sh.y_full[sA] : &sh.x_full[sA], ring_parity(rA));
      stage_to_tmem_half(smem_u32(half ? &sh.y_tiles[sA][0] : &sh.x_tiles[sA][0]), row, lane_addr + kColA, half);
      tc_fence_before();
      mbar_arrive(&sh.a_ready);

      for (int j = 0; j < nt; ++j) {
        const int tc = tc0 + j;
        const uint32_t sbuf = lane_addr + ((tc & 1) ? kColS1 : kColS0);
        if (threadIdx.x < kBlockN) {
          sh.colvec[j & 1][threadIdx.x] = cv_next;
          const int col = (j + 1) * kBlockN + threadIdx.x;
          cv_next = (j + 1 < nt && col < p.Lx) ? p.lse2[(long long)b * p.Lx + col] : INFINITY;
        }
        asm volatile("bar.sync 5, 256;" ::: "memory");
        mbar_wait(&sh.s_full[tc & 1], (tc >> 1) & 1);
        tc_fence_after();
        uint32_t r0[32];
        SAM2B200_TMEM_LD32(sbuf + half * kHalfN, r0);
        tmem_wait_ld();
        const float* cv = &sh.colvec[j & 1][half * kHalfN];
        const uint32_t didx = (uint32_t)(((long long)b * p.Lx + (j * kBlockN + half * kHalfN)) * p.La + a_row_idx);
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < kHalfN; i += 2) {
          float e0 = ex2(fmaf(__uint_as_float(r0[i]), c, -cv[i]));
          float e1 = ex2(fmaf(__uint_as_float(r0[i + 1]), c, -cv[i + 1]));
          if (drop_on) {
            e0 = sam2b200::dropout_keep(drop_key, didx + (uint32_t)i * (uint32_t)p.La, p.drop.thresh) ? e0 : 0.f;
            e1 = sam2b200::dropout_keep(drop_key, didx + (uint32_t)(i + 1) * (uint32_t)p.La, p.drop.thresh) ? e1 : 0.f;
          }
          pk[i >> 1] = pack_bf16(e0, e1);
        }
        SAM2B200_TMEM_ST16(sbuf + half * kHalfN, pk);
        tmem_wait_st();
        if (j > 0) mbar_wait(&sh.acc_done, (tc - 1) & 1);     // observe every phase of acc_done, in order
        tc_fence_before();
        mbar_arrive(&sh.p_ready[tc & 1]);
      }

      // epilogue through the EPI slot: half 0 stages in its X buffer, half 1 in its Y buffer (8 KB per warp)
      mbar_wait(&sh.acc_done, (tc0 + nt - 1) & 1);
      tc_fence_after();
      mbar_wait(half ? &sh.y_full[sE] : &sh.x_full[sE], ring_parity(rE));
      const uint32_t stage = smem_u32(half ? &sh.y_tiles[sE][0] : &sh.x_tiles[sE][0]) + quarter * 2 * kBoxBytes;
      const int row0 = a_tile * kBlockM + quarter * 32;
      grad_epilogue(p.gout, &map_o, stage, lane_addr + kColAcc, half, lane, row0, p.La, b, p.drop.inv_keep, false, tcur);
      tc0 += nt;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem, 512);
}

__device__ __forceinline__ int ring3_stage(int r) { return r % kStages3; }
__device__ __forceinline__ uint32_t ring3_parity(int r) { return (uint32_t)(r / kStages3) & 1u; }

// dK = scale * dS^T Q, persistent over (key block, object) items.  A1 = K block (tensor memory, for S^T = K Q^T),
// A2 = V block (its own 64 KB buffer, for dP^T = V dO^T), X = Q tiles, Y = dO tiles.  Ring slots per item:
// [ A1 | tile 0 .. tile nt-1 | epilogue staging ] over the two ring stages; A2 of the NEXT item is fetched as soon as the
// last dP MMA of the current item has completed (a2_empty), i.e. under the last softmax pass and the epilogue.
template <bool DROP, int DY = kD>     // DY = 64: raw-memory cross-attention (A2 = memory block, Y = dO Wv tiles)
__global__ void __launch_bounds__(kThreads, 1)
dk_persistent_kernel(const __grid_constant__ CUtensorMap map_a2,   // V  [B, M, 256] bf16, box 64 x 128
                     const __grid_constant__ CUtensorMap map_x,    // Q  box 64 x 64
                     const __grid_constant__ CUtensorMap map_y,    // dO box 64 x 64
                     const __grid_constant__ CUtensorMap map_a1,   // K  box 64 x 128
                     const __grid_constant__ CUtensorMap map_g,    // dK bf16, box 64 x 32
                     const ThreeGemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  SharedStorage3& sh = *reinterpret_cast<SharedStorage3*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nt = (p.Lx + kBlockN - 1) / kBlockN;
  const int slots = nt + 2;
  const int n_my = (p.n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages3; ++s) {
      mbar_init(&sh.x_full[s], 1); mbar_init(&sh.x_empty[s], 1);
      mbar_init(&sh.y_full[s], 1); mbar_init(&sh.y_empty[s], 1);
    }
    mbar_init(&sh.a2_full, 1);
    mbar_init(&sh.a2_empty, 1);
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

  if (warp == kProducerWarp) {
    const bool leader = elect_one();
    for (int it = 0; it < n_my; ++it) {
      const int item = (int)blockIdx.x + it * (int)gridDim.x;
      const int a_tile = item % p.n_atiles, b = item / p.n_atiles;
      int r = it * slots;
      {   // A1 slot
        const int s = ring3_stage(r);
        const uint32_t ph = ring3_parity(r);
        mbar_wait(&sh.x_empty[s], ph ^ 1);
        mbar_wait(&sh.y_empty[s], ph ^ 1);
        if (leader) {
          mbar_arrive_expect_tx(&sh.x_full[s], 2 * kSlabBytes);
          tma_load_3d(&sh.x_tiles[s][0], &map_a1, &sh.x_full[s], 0, a_tile * kBlockM, b);
          tma_load_3d(&sh.x_tiles[s][kSlabBytes], &map_a1, &sh.x_full[s], 64, a_tile * kBlockM, b);
          mbar_arrive_expect_tx(&sh.y_full[s], 2 * kSlabBytes);
          tma_load_3d(&sh.y_tiles[s][0], &map_a1, &sh.y_full[s], 128, a_tile * kBlockM, b);
          tma_load_3d(&sh.y_tiles[s][kSlabBytes], &map_a1, &sh.y_full[s], 192, a_tile * kBlockM, b);
        }
        __syncwarp();
      }
      mbar_wait(&sh.a2_empty, (uint32_t)(it & 1) ^ 1u);   // the previous item's dP MMAs are done with A2
      if (leader) {
        mbar_arrive_expect_tx(&sh.a2_full, kBlockM * DY * 2);
#pragma unroll
        for (int c = 0; c < DY / 64; ++c)
          tma_load_3d(&sh.a2[c * kA2ChunkBytes], &map_a2, &sh.a2_full, c * 64, a_tile * kBlockM, b);
      }
      __syncwarp();
      for (int j = 0; j < nt; ++j) {
        r = it * slots + 1 + j;
        const int s = ring3_stage(r);
        const uint32_t ph = ring3_parity(r);
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
          mbar_arrive_expect_tx(&sh.y_full[s], kBlockN * DY * 2);
#pragma unroll
          for (int c = 0; c < DY / 64; ++c)
            tma_load_3d(&sh.y_tiles[s][c * kChunkBytes], &map_y, &sh.y_full[s], c * 64, row0, b);
        }
        __syncwarp();
      }
      {   // EPI slot
        r = it * slots + nt + 1;
        const int s = ring3_stage(r);
        const uint32_t ph = ring3_parity(r);
        mbar_wait(&sh.x_empty[s], ph ^ 1);
        mbar_wait(&sh.y_empty[s], ph ^ 1);
        if (leader) { mbar_arrive(&sh.x_full[s]); mbar_arrive(&sh.y_full[s]); }
        __syncwarp();
      }
    }
  } else if (warp == kMmaWarp) {
    const bool leader = elect_one();
    constexpr uint32_t idesc_s = make_idesc_bf16(kBlockM, kBlockN, 0, 0);
    constexpr uint32_t idesc_acc = make_idesc_bf16(kBlockM, kD, 0, 1);
    const uint32_t a2_lo = desc_lo_sw128(smem_u32(&sh.a2[0]), 16);
    const uint32_t x_lo0 = desc_lo_sw128(smem_u32(&sh.x_tiles[0][0]), 16);
    const uint32_t xm_lo0 = desc_lo_sw128(smem_u32(&sh.x_tiles[0][0]), kChunkBytes);
    const uint32_t y_lo0 = desc_lo_sw128(smem_u32(&sh.y_tiles[0][0]), 16);
    int tc0 = 0;
    for (int it = 0; it < n_my; ++it) {
      const int rA = it * slots;
      auto issue_s = [&](int j) {        // S^T[j] = A1(TMEM) . X[j]^T
        const int r = rA + 1 + j;
        const int s = ring3_stage(r);
        mbar_wait(&sh.x_full[s], ring3_parity(r));
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
      auto issue_dp = [&](int j) {       // dP^T[j] = A2(SMEM) . Y[j]^T
        const int r = rA + 1 + j;
        const int s = ring3_stage(r);
        mbar_wait(&sh.y_full[s], ring3_parity(r));
        tc_fence_after();
        if (leader) {
          const uint32_t ylo = y_lo0 + s * (kTileBytes >> 4);
#pragma unroll
          for (int ks = 0; ks < DY / 16; ++ks)
            umma_ss_lohi(tmem + k3ColDP, a2_lo + (ks >> 2) * (kA2ChunkBytes >> 4) + (ks & 3) * 2,
                         ylo + (ks >> 2) * (kChunkBytes >> 4) + (ks & 3) * 2, kDescHiSw128_1024, idesc_s, ks > 0);
          umma_commit(&sh.y_empty[s]);
          umma_commit(&sh.dp_full);
          if (j + 1 >= nt) umma_commit(&sh.a2_empty);
        }
        __syncwarp();
      };
      mbar_wait(&sh.a1_ready, it & 1);
      if (leader) {   // release the A1 slot of this item and the EPI slot of the previous one
        mbar_arrive(&sh.x_empty[ring3_stage(rA)]);
        mbar_arrive(&sh.y_empty[ring3_stage(rA)]);
        if (it > 0) {
          mbar_arrive(&sh.x_empty[ring3_stage(rA - 1)]);
          mbar_arrive(&sh.y_empty[ring3_stage(rA - 1)]);
        }
      }
      __syncwarp();
      mbar_wait(&sh.a2_full, it & 1);
      tc_fence_after();
      issue_s(0);
      issue_dp(0);
      for (int j = 0; j < nt; ++j) {
        const int r = rA + 1 + j;
        const int s = ring3_stage(r);
        const int tc = tc0 + j;
        if (j + 1 < nt) {
          mbar_wait(&sh.s_free, tc & 1);
          tc_fence_after();
          issue_s(j + 1);
        }
        mbar_wait(&sh.ds_ready, tc & 1);
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
      // the last tile's S was read (s_free) but nobody waited for that phase: consume it so the parities stay in step
      mbar_wait(&sh.s_free, (tc0 + nt - 1) & 1);
      tc0 += nt;
    }
  } else {
    const int quarter = warp & 3;
    const int half = warp >> 2;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = tmem + (uint32_t(quarter * 32) << 16);
    const float c = p.scale_log2;
    constexpr bool drop_on = DROP;
    const uint32_t drop_key = drop_on ? sam2b200::dropout_key(*p.drop.seed, p.drop.site) : 0u;
    int tc0 = 0;
    // rotation table in shared memory: the X part once, the Y part of each item's rows into one of two buffers (item parity) --
    // the buffer an item overwrites was last read two epilogues ago, and every warp passes the barrier in front of each epilogue
    const bool tab_on = p.gout.rope_smem > 0;
    const uint32_t tab_area = smem_u32(&sh) + (uint32_t)sizeof(SharedStorage3);
    const uint32_t tab_y0 = tab_area + p.gout.rope_w * kRopeXStride;
    if (tab_on) rope_stage_x(p.gout, tab_area, threadIdx.x);
    for (int it = 0; it < n_my; ++it) {
      const int item = (int)blockIdx.x + it * (int)gridDim.x;
      const int a_tile = item % p.n_atiles, b = item / p.n_atiles;
      const int rA = it * slots, rE = rA + nt + 1;
      const int sA = ring3_stage(rA), sE = ring3_stage(rE);
      const long long a_row_idx = (long long)a_tile * kBlockM + row;
      uint32_t tab = 0;
      if (tab_on) {
        const uint32_t ybase = tab_y0 + (it & 1) * rope_y_bytes(p.gout.rope_w);
        rope_stage_y(p.gout, ybase, a_tile * kBlockM, threadIdx.x);
        cp_async_commit();
        tab = rope_tab_addr(p.gout, tab_area, ybase, a_tile * kBlockM, row, half);
      }
      float lse_next = INFINITY, delta_next = 0.f;   // per-column vectors staged one tile ahead
      if (threadIdx.x < kBlockN && (int)threadIdx.x < p.Lx) {
        lse_next = p.lse2[(long long)b * p.Lx + threadIdx.x];
        delta_next = p.delta[(long long)b * p.Lx + threadIdx.x];
      }
      mbar_wait(half ? &sh.y_full[sA] : &sh.x_full[sA], ring3_parity(rA));
      stage_to_tmem_half(smem_u32(half ? &sh.y_tiles[sA][0] : &sh.x_tiles[sA][0]), row, lane_addr + k3ColA1, half);
      tc_fence_before();
      mbar_arrive(&sh.a1_ready);
      const int row0 = a_tile * kBlockM + quarter * 32;
      const bool rotate = p.gout.rope_table != nullptr && (row0 + lane) < p.gout.rope_rows;

      for (int j = 0; j < nt; ++j) {
        const int tc = tc0 + j;
        if (threadIdx.x < kBlockN) {
          sh.col_lse[j & 1][threadIdx.x] = lse_next;
          sh.col_delta[j & 1][threadIdx.x] = delta_next;
          const int col = (j + 1) * kBlockN + threadIdx.x;
          const bool ok = (j + 1 < nt) && col < p.Lx;
          lse_next = ok ? p.lse2[(long long)b * p.Lx + col] : INFINITY;
          delta_next = ok ? p.delta[(long long)b * p.Lx + col] : 0.f;
        }
        asm volatile("bar.sync 5, 256;" ::: "memory");
        mbar_wait(&sh.s_full, tc & 1);
        tc_fence_after();
        float pv[kHalfN];
        {
          uint32_t r0[32];
          SAM2B200_TMEM_LD32(lane_addr + k3ColS + half * kHalfN, r0);
          tmem_wait_ld();
          tc_fence_before();
          mbar_arrive(&sh.s_free);
#pragma unroll
          for (int i = 0; i < kHalfN; ++i)
            pv[i] = ex2(fmaf(__uint_as_float(r0[i]), c, -sh.col_lse[j & 1][half * kHalfN + i]));
        }
        mbar_wait(&sh.dp_full, tc & 1);
        tc_fence_after();
        uint32_t pk[16];
        {
          uint32_t r0[32];
          SAM2B200_TMEM_LD32(lane_addr + k3ColDP + half * kHalfN, r0);
          tmem_wait_ld();
          const uint32_t didx = (uint32_t)(((long long)b * p.Lx + (j * kBlockN + half * kHalfN)) * p.La + a_row_idx);
          const uint32_t dstep = (uint32_t)p.La;
#pragma unroll
          for (int i = 0; i < kHalfN; i += 2) {
            float d0 = __uint_as_float(r0[i]);
            float d1 = __uint_as_float(r0[i + 1]);
            if (drop_on) {
              d0 = sam2b200::dropout_keep(drop_key, didx + (uint32_t)i * dstep, p.drop.thresh) ? d0 * p.drop.inv_keep : 0.f;
              d1 = sam2b200::dropout_keep(drop_key, didx + (uint32_t)(i + 1) * dstep, p.drop.thresh) ? d1 * p.drop.inv_keep : 0.f;
            }
            const float dl0 = sh.col_delta[j & 1][half * kHalfN + i];
            const float dl1 = sh.col_delta[j & 1][half * kHalfN + i + 1];
            pk[i >> 1] = pack_bf16(pv[i] * (d0 - dl0), pv[i + 1] * (d1 - dl1));
          }
        }
        SAM2B200_TMEM_ST16(lane_addr + k3ColDP + half * kHalfN, pk);
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(&sh.ds_ready);
      }
      float2 tcur[16];
      if (tab == 0) load_table_chunk(p.gout, rotate, row0 + lane, half * 128, tcur);
      else { cp_async_wait_all(); asm volatile("bar.sync 6, 256;" ::: "memory"); }     // the staged table is complete and visible
      mbar_wait(&sh.acc_done, it & 1);
      tc_fence_after();
      mbar_wait(half ? &sh.y_full[sE] : &sh.x_full[sE], ring3_parity(rE));
      const uint32_t stage = smem_u32(half ? &sh.y_tiles[sE][0] : &sh.x_tiles[sE][0]) + quarter * 2 * kBoxBytes;
      grad_epilogue(p.gout, &map_g, stage, lane_addr + k3ColAcc, half, lane, row0, p.La, b, p.scale, rotate, tcur, nullptr, tab);
      tc0 += nt;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem, 512);
}

}  // namespace attn
