// Key-side backward of the attention core as ONE kernel on CTA PAIRS (thread-block cluster of 2):
//
//   rank 0 ("V side"):  S^T  = K . Q^T      ->  P^T  = exp2(S^T c - LSE2[q])     ->  dV += P^T . dO
//   rank 1 ("K side"):  dP^T = V . dO^T     ->  dS^T = P^T o (dP^T - Delta[q])   ->  dK += dS^T . Q
//
// Both CTAs own the same 128-key block and stream the same (Q, dO) tiles; the probability tile P^T that the K side
// needs is computed ONCE, by the V side, and shipped to the partner's shared memory through distributed shared
// memory (st.shared::cluster + a cluster-scope mbarrier).  The separate dV and dK kernels (attn_kernels.cuh) execute
// 2 + 3 = 5 GEMM units per (key block, query tile) because the dK kernel has to recompute S^T; the pair executes
// 2 + 2 = 4, evenly split over the two SMs.  (A single CTA cannot hold both accumulators: dV and dK are
// 2 x 256 fp32 columns = all of tensor memory.)
//
// Everything else -- fixed operand by TMA -> SMEM -> TMEM, 3-stage TMA tile ring, score / probability tiles in TMEM,
// TMA-store epilogues with fused conjugate RoPE, scale and bias gradients -- is the machinery of two_gemm_kernel.
//
// The exchange is asynchronous: the V side stages its tile in its own shared memory and ONE thread per lane quarter
// issues cp.async.bulk.shared::cluster (4 KB), which completes on an mbarrier of the partner; the partner hands buffers
// back with a RELAXED remote arrive (4 buffers deep).  Measured on B200 (profiles/r1_pair_kernel_experiment.txt):
// synchronous st.shared::cluster + release.cluster arrives in the compute warps cost +1.5 us per tile; the release.cluster
// hand-back alone +0.3 us per tile.  The cluster launch and its two barriers cost ~3.5 us per pair, so the pair is used
// for long query loops only (N >= 1024, attn.cu): -8 % backward time at N = 4096, -3 % at N = 1024, +6 % at N = 576.
#pragma once

#include "attn_kernels.cuh"

namespace attn {

struct PairParams {
  int Lk;                      // keys per batch item (rows of K, V, dK, dV)
  int Lq;                      // queries per batch item (rows of Q, dO; streamed)
  float scale_log2;            // softmax scale * log2(e)
  float scale;                 // softmax scale (applied to dK in the epilogue)
  const float* lse2;           // [B, Lq] log2-domain LSE of the forward
  const float* delta;          // [B, Lq] rowsum(dO o O)
  GradOut gout_v, gout_k;
  sam2b200::Dropout drop;      // attention-probability dropout; element index (b Lq + q) Lk + key
};

constexpr int kPairStages = 2;     // tile ring (a 2-GEMM loop needs its next tiles ~1.5 tile times ahead: two stages suffice)
constexpr int kPBufs = 4;          // P^T exchange depth: the V side may run 4 tiles ahead of the K side's reads, which
                                   // covers the copy + remote-arrive round trip (2 buffers throttled it to ~1.4 us / tile)
struct PairShared {
  alignas(1024) uint8_t x_tiles[kPairStages][kTileBytes];
  alignas(1024) uint8_t y_tiles[kPairStages][kTileBytes];
  // P^T tiles [128 keys][64 queries] bf16, double-buffered, 16-byte chunks XOR (row & 7): rank 0 stages its tile here
  // and a bulk async copy (cp.async.bulk.shared::cluster) moves it into the same buffer of rank 1
  alignas(1024) uint8_t p_buf[kPBufs][kBlockM * 128];
  alignas(8) uint64_t x_full[kPairStages];
  uint64_t x_empty[kPairStages];
  uint64_t y_full[kPairStages];
  uint64_t y_empty[kPairStages];
  uint64_t s_full[2];
  uint64_t p_ready[2];
  uint64_t acc_done;
  uint64_t a_full;
  uint64_t a_ready;
  uint64_t p_full[kPBufs];          // live in rank 1: P^T of tile j has landed in p_buf[j & 1] (16 KB of complete_tx from rank 0's copies)
  uint64_t p_free[kPBufs];     // live in rank 0: rank 1 has read p_buf[j % kPBufs] (one arrival per rank-1 warp)
  float colvec[2][kBlockN];    // per-column vector of the tile: LSE2 (rank 0) / Delta (rank 1)
  uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address -> shared::cluster address of the same variable in CTA `rank`
__device__ __forceinline__ uint32_t map_to_rank(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster128(uint32_t cluster_addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(cluster_addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// Bulk asynchronous copy own shared memory -> shared memory of another CTA of the cluster; the bytes complete on an
// mbarrier of the destination CTA (both destination addresses are shared::cluster addresses from map_to_rank).
__device__ __forceinline__ void bulk_copy_to_cluster(uint32_t dst_cluster_addr, uint32_t src_cta_addr, uint32_t bytes,
                                                     uint32_t dst_cluster_bar) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :
               : "r"(dst_cluster_addr), "r"(src_cta_addr), "r"(bytes), "r"(dst_cluster_bar)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_bar_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar_addr) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote_relaxed(uint32_t cluster_bar_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait_cluster(bar, parity)) {
    if (++spins > SAM2B200_SPIN_LIMIT) {
      printf("sam2b200: cluster mbarrier timeout block (%d,%d) thread %d bar %p parity %u\n", blockIdx.x, blockIdx.y,
             threadIdx.x, bar, parity);
      __trap();
    }
  }
}

// DROP is a compile-time switch: with the dropout code merely branched around at run time the no-dropout kernel lost
// 27 % (917 -> 1169 us for the key side at cfg4: the masked copy `pm` and the second tcgen05.st site cost registers and
// scheduling freedom in the V-side loop even when never executed).
template <bool DROP>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
kv_pair_kernel(const __grid_constant__ CUtensorMap map_q64, const __grid_constant__ CUtensorMap map_do64,
               const __grid_constant__ CUtensorMap map_k128, const __grid_constant__ CUtensorMap map_v128,
               const __grid_constant__ CUtensorMap map_dv, const __grid_constant__ CUtensorMap map_dk,
               const PairParams p) {
  extern __shared__ uint8_t smem_raw[];
  PairShared& sh = *reinterpret_cast<PairShared*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();          // 0: V side, 1: K side
  const bool kside = rank != 0;
  const int a_tile = blockIdx.x >> 1;               // 128-key block shared by the pair
  const int b = blockIdx.y;
  const int nt = (p.Lq + kBlockN - 1) / kBlockN;
  // scores GEMM streams X (K-major view), accumulate GEMM streams Y (MN-major view)
  const CUtensorMap* map_x = kside ? &map_do64 : &map_q64;
  const CUtensorMap* map_y = kside ? &map_q64 : &map_do64;
  const CUtensorMap* map_a = kside ? &map_v128 : &map_k128;
  const CUtensorMap* map_o = kside ? &map_dk : &map_dv;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kPairStages; ++s) {
      mbar_init(&sh.x_full[s], 1); mbar_init(&sh.x_empty[s], 1);
      mbar_init(&sh.y_full[s], 1); mbar_init(&sh.y_empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) { mbar_init(&sh.s_full[i], 1); mbar_init(&sh.p_ready[i], kNumSoftmaxThreads); }
    mbar_init(&sh.acc_done, 1);
    mbar_init(&sh.a_full, 1);
    mbar_init(&sh.a_ready, kNumSoftmaxThreads);
    for (int i = 0; i < kPBufs; ++i) { mbar_init(&sh.p_full[i], 1); mbar_init(&sh.p_free[i], kNumSoftmaxWarps); }
    fence_barrier_init();
  }
  if (warp == kProducerWarp && lane == 0) { prefetch_tmap(map_a); prefetch_tmap(map_x); prefetch_tmap(map_y); }
  if (warp == 0 && lane == 0) prefetch_tmap(map_o);
  if (warp == kMmaWarp) { tmem_alloc(&sh.tmem_base, 512); tmem_relinquish(); }
  tc_fence_before();
  cluster_sync_all();                                // barrier inits of BOTH CTAs visible before any remote arrive
  tc_fence_after();
  const uint32_t tmem = sh.tmem_base;

  if (warp == kProducerWarp) {
    // ===================== TMA producer =====================
    const bool leader = elect_one();
    if (leader) {
      mbar_arrive_expect_tx(&sh.a_full, 4 * kSlabBytes);
#pragma unroll
      for (int c = 0; c < 4; ++c)
        tma_load_3d((c < 2 ? &sh.x_tiles[kPairStages - 1][0] : &sh.y_tiles[kPairStages - 1][0]) + (c & 1) * kSlabBytes, map_a,
                    &sh.a_full, c * 64, a_tile * kBlockM, b);
    }
    __syncwarp();
    for (int j = 0; j < nt; ++j) {
      const int s = j % kPairStages;
      const uint32_t ph = (j / kPairStages) & 1;
      const int row0 = j * kBlockN;
      if (j == kPairStages - 1) mbar_wait(&sh.a_ready, 0);
      mbar_wait(&sh.x_empty[s], ph ^ 1);
      if (leader) {
        mbar_arrive_expect_tx(&sh.x_full[s], kTileBytes);
#pragma unroll
        for (int c = 0; c < 4; ++c) tma_load_3d(&sh.x_tiles[s][c * kChunkBytes], map_x, &sh.x_full[s], c * 64, row0, b);
      }
      __syncwarp();
      mbar_wait(&sh.y_empty[s], ph ^ 1);
      if (leader) {
        mbar_arrive_expect_tx(&sh.y_full[s], kTileBytes);
#pragma unroll
        for (int c = 0; c < 4; ++c) tma_load_3d(&sh.y_tiles[s][c * kChunkBytes], map_y, &sh.y_full[s], c * 64, row0, b);
      }
      __syncwarp();
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer (identical on both sides) =====================
    const bool leader = elect_one();
    constexpr uint32_t idesc_s = make_idesc_bf16(kBlockM, kBlockN, 0, 0);
    constexpr uint32_t idesc_acc = make_idesc_bf16(kBlockM, kD, 0, 1);
    const uint32_t x_lo0 = desc_lo_sw128(smem_u32(&sh.x_tiles[0][0]), 16);
    const uint32_t y_lo0 = desc_lo_sw128(smem_u32(&sh.y_tiles[0][0]), kChunkBytes);
    auto issue_scores = [&](int t) {
      const int s = t % kPairStages;
      mbar_wait(&sh.x_full[s], (t / kPairStages) & 1);
      tc_fence_after();
      if (leader) {
        const uint32_t xlo = x_lo0 + s * (kTileBytes >> 4);
        const uint32_t d = tmem + ((t & 1) ? kColS1 : kColS0);
#pragma unroll
        for (int ks = 0; ks < kD / 16; ++ks)
          umma_ts_lohi(d, tmem + kColA + ks * 8, xlo + (ks >> 2) * (kChunkBytes >> 4) + (ks & 3) * 2, kDescHiSw128_1024,
                       idesc_s, ks > 0);
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
      const int s = j % kPairStages;
      mbar_wait(&sh.p_ready[j & 1], (j >> 1) & 1);
      mbar_wait(&sh.y_full[s], (j / kPairStages) & 1);
      tc_fence_after();
      if (leader) {
        const uint32_t ylo = y_lo0 + s * (kTileBytes >> 4);
        const uint32_t pa = tmem + ((j & 1) ? kColS1 : kColS0);
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
  } else {
    // ===================== compute / epilogue warps (0..7) =====================
    const int quarter = warp & 3;
    const int half = warp >> 2;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = tmem + (uint32_t(quarter * 32) << 16);
    {
      mbar_wait(&sh.a_full, 0);
      stage_to_tmem_half(smem_u32(half ? &sh.y_tiles[kPairStages - 1][0] : &sh.x_tiles[kPairStages - 1][0]), row, lane_addr + kColA, half);
      tc_fence_before();
      mbar_arrive(&sh.a_ready);
    }
    const int row0 = a_tile * kBlockM + quarter * 32;
    const uint32_t stage = smem_u32(&sh.x_tiles[0][0]) + warp * (4 * kBoxBytes);
    const GradOut& gout = kside ? p.gout_k : p.gout_v;
    const bool rotate = gout.rope_table != nullptr && (row0 + lane) < gout.rope_rows;
    const float c = p.scale_log2;
    constexpr bool drop_on = DROP;
    const uint32_t drop_key = drop_on ? sam2b200::dropout_key(*p.drop.seed, p.drop.site) : 0u;
    const long long key_row = (long long)a_tile * kBlockM + row;
    // this thread's 64 bytes of the P^T exchange tile (row `row`, 16-byte chunks 4*half .. 4*half+3, XOR-swizzled)
    const uint32_t p_local = smem_u32(&sh.p_buf[0][0]) + row * 128;
    const uint32_t p_remote = map_to_rank(p_local, 1);                       // rank 0 writes into rank 1's buffers
    const uint32_t p_full_remote = map_to_rank(smem_u32(&sh.p_full[0]), 1);  // rank 0 -> rank 1
    const uint32_t p_free_remote = map_to_rank(smem_u32(&sh.p_free[0]), 0);  // rank 1 -> rank 0
    constexpr uint32_t kPBufBytes = kBlockM * 128;

    const float* colsrc = kside ? p.delta : p.lse2;
    const float col_oob = kside ? 0.f : INFINITY;
    float cv_next = col_oob;
    if (threadIdx.x < kBlockN && (int)threadIdx.x < p.Lq) cv_next = colsrc[(long long)b * p.Lq + threadIdx.x];
    for (int j = 0; j < nt; ++j) {
      const uint32_t sbuf = lane_addr + ((j & 1) ? kColS1 : kColS0);
      if (threadIdx.x < kBlockN) {
        sh.colvec[j & 1][threadIdx.x] = cv_next;
        const int col = (j + 1) * kBlockN + threadIdx.x;
        cv_next = (j + 1 < nt && col < p.Lq) ? colsrc[(long long)b * p.Lq + col] : col_oob;
      }
      if (kside && threadIdx.x == 0) mbar_arrive_expect_tx(&sh.p_full[j % kPBufs], kPBufBytes);   // arm: 16 KB from the partner
      asm volatile("bar.sync 5, 256;" ::: "memory");
      mbar_wait(&sh.s_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      uint32_t r0[32];
      SAM2B200_TMEM_LD32(sbuf + half * kHalfN, r0);
      tmem_wait_ld();
      const float* cv = &sh.colvec[j & 1][half * kHalfN];
      uint32_t pk[16];
      if (!kside) {
        // ---- V side: P^T = exp2(S^T c - LSE2[q]); keep it (TMEM, A operand of dV += P^T dO) and ship it to the partner
        const uint32_t didx = (uint32_t)(((long long)b * p.Lq + (j * kBlockN + half * kHalfN)) * p.Lk + key_row);
        uint32_t pm[16];     // dropout-masked copy for this side's dV GEMM; the partner needs the un-masked P^T
#pragma unroll
        for (int i = 0; i < kHalfN; i += 2) {
          const float e0 = ex2(fmaf(__uint_as_float(r0[i]), c, -cv[i]));
          const float e1 = ex2(fmaf(__uint_as_float(r0[i + 1]), c, -cv[i + 1]));
          pk[i >> 1] = pack_bf16(e0, e1);
          if (drop_on)
            pm[i >> 1] = pack_bf16(sam2b200::dropout_keep(drop_key, didx + (uint32_t)i * (uint32_t)p.Lk, p.drop.thresh) ? e0 : 0.f,
                                   sam2b200::dropout_keep(drop_key, didx + (uint32_t)(i + 1) * (uint32_t)p.Lk, p.drop.thresh) ? e1 : 0.f);
        }
        if (drop_on) { SAM2B200_TMEM_ST16(sbuf + half * kHalfN, pm); } else { SAM2B200_TMEM_ST16(sbuf + half * kHalfN, pk); }
        tmem_wait_st();
        // own pipeline first: the dV GEMM of this tile must not wait for the partner
        if (j > 0) mbar_wait(&sh.acc_done, (j - 1) & 1);
        tc_fence_before();
        mbar_arrive(&sh.p_ready[j & 1]);
        // ship P^T asynchronously: stage the 64 bytes of this thread locally, then ONE bulk copy per lane quarter
        // (32 rows x 128 B = 4 KB contiguous) into the partner's buffer, completing on the partner's mbarrier
        const int pb = j % kPBufs;
        if (j >= kPBufs) mbar_wait(&sh.p_free[pb], ((j - kPBufs) / kPBufs) & 1);   // partner has read the tile that used this buffer
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4)
          sts128(p_local + pb * kPBufBytes + (((half * 4 + q4) ^ (row & 7)) << 4), pk[4 * q4], pk[4 * q4 + 1],
                 pk[4 * q4 + 2], pk[4 * q4 + 3]);
        fence_proxy_async();
        pair_barrier(quarter);                                           // both halves of the quarter's 32 rows are staged
        if (half == 0 && lane == 0)
          bulk_copy_to_cluster(p_remote + pb * kPBufBytes - row * 128 + quarter * 32 * 128,
                               p_local + pb * kPBufBytes - row * 128 + quarter * 32 * 128, 32 * 128, p_full_remote + pb * 8);
        continue;
      } else {
        // ---- K side: dS^T = P^T o (dP^T - Delta[q]) with P^T from the partner
        const int pb = j % kPBufs;
        mbar_wait(&sh.p_full[pb], (j / kPBufs) & 1);
        uint4 pv[4];
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) pv[q4] = lds128(p_local + pb * kPBufBytes + (((half * 4 + q4) ^ (row & 7)) << 4));
        const uint32_t pw[16] = {pv[0].x, pv[0].y, pv[0].z, pv[0].w, pv[1].x, pv[1].y, pv[1].z, pv[1].w,
                                 pv[2].x, pv[2].y, pv[2].z, pv[2].w, pv[3].x, pv[3].y, pv[3].z, pv[3].w};
        const uint32_t didx = (uint32_t)(((long long)b * p.Lq + (j * kBlockN + half * kHalfN)) * p.Lk + key_row);
#pragma unroll
        for (int i = 0; i < kHalfN; i += 2) {
          const __nv_bfloat162 pp = *reinterpret_cast<const __nv_bfloat162*>(&pw[i >> 1]);
          const float2 pf = __bfloat1622float2(pp);
          float d0 = __uint_as_float(r0[i]), d1 = __uint_as_float(r0[i + 1]);
          if (drop_on) {
            d0 = sam2b200::dropout_keep(drop_key, didx + (uint32_t)i * (uint32_t)p.Lk, p.drop.thresh) ? d0 * p.drop.inv_keep : 0.f;
            d1 = sam2b200::dropout_keep(drop_key, didx + (uint32_t)(i + 1) * (uint32_t)p.Lk, p.drop.thresh) ? d1 * p.drop.inv_keep : 0.f;
          }
          pk[i >> 1] = pack_bf16(pf.x * (d0 - cv[i]), pf.y * (d1 - cv[i + 1]));
        }
        SAM2B200_TMEM_ST16(sbuf + half * kHalfN, pk);
        // Hand the buffer back.  RELAXED on purpose: a release.cluster arrive here (8 per tile) measured +0.3 us per tile
        // (it drains the warp's outstanding memory operations).  The P^T values have been consumed by the arithmetic
        // above (in-order issue: the shared loads have completed), and the partner can overwrite the buffer only after
        // this arrive has crossed the cluster AND its next bulk copy has been issued and delivered.
        __syncwarp();
        if (lane == 0) mbar_arrive_remote_relaxed(p_free_remote + pb * 8);
        tmem_wait_st();
      }
      if (j > 0) mbar_wait(&sh.acc_done, (j - 1) & 1);   // observe every phase of acc_done in order (see two_gemm_kernel)
      tc_fence_before();
      mbar_arrive(&sh.p_ready[j & 1]);
    }

    float2 tcur[16];
    load_table_chunk(gout, rotate, row0 + lane, half * 128, tcur);
    mbar_wait(&sh.acc_done, (nt - 1) & 1);
    tc_fence_after();
    grad_epilogue(gout, map_o, stage, lane_addr + kColAcc, half, lane, row0, p.Lk, b, kside ? p.scale : p.drop.inv_keep, rotate, tcur);
  }

  tc_fence_before();
  cluster_sync_all();          // neither CTA may exit while its partner can still write / arrive into its shared memory
  if (warp == kMmaWarp) tmem_dealloc(tmem, 512);
}

}  // namespace attn
