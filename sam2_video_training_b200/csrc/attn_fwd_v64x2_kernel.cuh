// Forward of the raw-memory cross-attention with TWO INDEPENDENT ONLINE-SOFTMAX STREAMS per CTA.
// EXPERIMENT (sam2b200_debug_set_variant key 2 = 1), NOT THE DEFAULT: measured on B200 it runs exactly as fast as the single-stream
// kernel (profiles/r2_fwd_two_softmax_streams_experiment.txt: cfg2 0.108 vs 0.109 ms, cfg4 0.659 vs 0.633 ms), i.e. the
// premise below was wrong -- the forward is paced by the 20 MMAs per tile (16 of shape 128x64x16 at ~37 clk issue cadence
// + 4), not by the softmax chain.  Kept as the parity-tested record of that measurement.
//
// two_gemm_kernel<FWD, ., 64> is bound by its softmax chain, not by the tensor pipe: per 64-key tile the MMAs take
// 512 (S = Q K^T, A from tensor memory) + 128 (P . mem, 64 columns) = 640 clk, but one tile takes ~1200 clk
// (profiles/r2_ncu_kernels_cfg2.csv: tensor pipe 43 %): wait S -> tcgen05.ld -> row max -> pair exchange -> 32 x (FMA, EX2)
// -> pack -> tcgen05.st -> arrive, one dependent chain for eight warps.  (The backward kernels are different: there the
// SS-mode MMAs' shared-memory reads set the pace, and a second softmax group bought nothing --
// profiles/r2_two_softmax_groups_experiment.txt.)
// The narrow accumulator (64 columns) leaves room in tensor memory for TWO of everything:
//     TMEM   ACC0 64 | ACC1 64 | Q 128 | S[group 0] 2 x 64 | S[group 1] 2 x 64                      = 512 columns
// Sixteen softmax warps form two groups; group g owns the tiles j = g (mod 2), its own running (max, sum) per row, its own
// accumulator ACC_g and two score buffers.  The MMA warp runs the score GEMMs four tiles ahead and alternates the two
// groups' P . mem products.  At the end the two partial results are merged like the partials of a split-KV forward:
//     m = max(m0, m1),  O = (O0 2^(m0 - m) + O1 2^(m1 - m)) / (l0 2^(m0 - m) + l1 2^(m1 - m)).
// PROJ: the folded output projection out_proj(v_proj(.)) in the epilogue, exactly as in two_gemm_kernel<.., 64, true>.
// Numerics: the same per-tile arithmetic; rows are normalised by the same sums in a different association (fp32).
#pragma once

#include "attn_kernels.cuh"

namespace attn {

constexpr int kF2Threads = 18 * 32;
constexpr int kF2Producer = 16, kF2Mma = 17;
constexpr uint32_t kF2ColAcc0 = 0, kF2ColAcc1 = 64, kF2ColA = 128, kF2ColS = 256;     // S buffer (g, u) at kF2ColS + (2 g + u) * 64
constexpr uint32_t kF2ColProjA = 128, kF2ColProjD = 256;                                // after the loop: over the dead Q / score columns
// Rings: the score GEMMs run FOUR tiles ahead of the P . mem products and one producer thread loads K_j, mem_j, K_j+1, ... in
// order, so K_j+4 can only be requested once mem_j+3's stage is free, i.e. once P . mem of tile j + 3 - kF2YStages has
// completed.  With three stages of each (the single-stream kernel's ring) that is tile j itself -- issued a moment ago -- and
// every tile then pays a full TMA latency on the MMA warp (measured: 2 x slower than the single-stream kernel).  Four K
// stages (reuse distance = the look-ahead) and six memory stages (8 KB each) take the producer off the critical path.
constexpr int kF2XStages = 4, kF2YStages = 6;
constexpr int kF2YBytes = kBlockN * 64 * 2;                                              // 8 KB: one [64 x 64] bf16 memory tile

struct SharedStorageF2 {
  alignas(1024) uint8_t x_tiles[kF2XStages][kTileBytes];   // stages 2, 3 stage Q first
  alignas(1024) uint8_t y_tiles[kF2YStages][kF2YBytes];
  alignas(1024) uint8_t w_tile[2 * kSlabBytes];             // PROJ: the folded weight, two [128 x 64] halves
  alignas(8) uint64_t x_full[kF2XStages];
  uint64_t x_empty[kF2XStages];
  uint64_t y_full[kF2YStages];
  uint64_t y_empty[kF2YStages];
  uint64_t s_full[4];
  uint64_t p_ready[4];
  uint64_t acc_done[2];
  uint64_t a_full, a_ready;
  uint64_t w_full, o_ready, proj_done;
  float xchg[2][2][2][kBlockM];     // [group][use][half][row]: row-max exchange between the two halves of a row
  float lsum[2][2][kBlockM];        // [group][half][row]
  float gstat[3][kBlockM];          // group 1 -> group 0: (m_ref * c, l, kept-probability sum)
  uint32_t tmem_base;
};

template <bool DROP, bool PROJ>
__global__ void __launch_bounds__(kF2Threads, 1)
fwd_v64x2_kernel(const __grid_constant__ CUtensorMap map_x,    // K [B, M, 256] bf16, box 64 x 64
                 const __grid_constant__ CUtensorMap map_y,    // mem [B, M, 64] bf16, box 64 x 64
                 const __grid_constant__ CUtensorMap map_a,    // Q [B, N, 256] bf16, box 64 x 128
                 const __grid_constant__ CUtensorMap map_w,    // PROJ: folded weight [256, 64] bf16, box 64 x 128
                 const __grid_constant__ CUtensorMap map_p,    // PROJ: projected output [B, N, 256] bf16, box 64 x 32
                 const TwoGemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  SharedStorageF2& sh = *reinterpret_cast<SharedStorageF2*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int a_tile = blockIdx.x, b = blockIdx.y;
  const int nt = (p.Lx + kBlockN - 1) / kBlockN;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kF2XStages; ++s) { mbar_init(&sh.x_full[s], 1); mbar_init(&sh.x_empty[s], 1); }
    for (int s = 0; s < kF2YStages; ++s) { mbar_init(&sh.y_full[s], 1); mbar_init(&sh.y_empty[s], 1); }
    for (int i = 0; i < 4; ++i) { mbar_init(&sh.s_full[i], 1); mbar_init(&sh.p_ready[i], kNumSoftmaxThreads); }
    for (int i = 0; i < 2; ++i) mbar_init(&sh.acc_done[i], 1);
    mbar_init(&sh.a_full, 1);
    mbar_init(&sh.a_ready, 2 * kNumSoftmaxThreads);
    mbar_init(&sh.w_full, 1); mbar_init(&sh.o_ready, kNumSoftmaxThreads); mbar_init(&sh.proj_done, 1);
    fence_barrier_init();
  }
  if (warp == kF2Producer && lane == 0) { prefetch_tmap(&map_a); prefetch_tmap(&map_x); prefetch_tmap(&map_y); if (PROJ) prefetch_tmap(&map_w); }
  if (warp == 0 && lane == 0 && PROJ) prefetch_tmap(&map_p);
  if (warp == kF2Mma) { tmem_alloc(&sh.tmem_base, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sh.tmem_base;

  if (warp == kF2Producer) {
    // ===================== TMA producer =====================
    const bool leader = elect_one();
    if (leader) {   // Q -> K stages 2 (slabs 0, 1) and 3 (slabs 2, 3)
      mbar_arrive_expect_tx(&sh.a_full, 4 * kSlabBytes);
#pragma unroll
      for (int c = 0; c < 4; ++c)
        tma_load_3d(&sh.x_tiles[2 + (c >> 1)][0] + (c & 1) * kSlabBytes, &map_a, &sh.a_full, c * 64, a_tile * kBlockM, b);
      if (PROJ) {
        mbar_arrive_expect_tx(&sh.w_full, 2 * kSlabBytes);
        tma_load_3d(&sh.w_tile[0], &map_w, &sh.w_full, 0, 0, 0);
        tma_load_3d(&sh.w_tile[kSlabBytes], &map_w, &sh.w_full, 0, kBlockM, 0);
      }
    }
    __syncwarp();
    for (int j = 0; j < nt; ++j) {
      const int sx = j % kF2XStages, sy = j % kF2YStages;
      const int row0 = j * kBlockN;
      if (j == 2) mbar_wait(&sh.a_ready, 0);   // first use of the stages that staged Q
      mbar_wait(&sh.x_empty[sx], ((j / kF2XStages) & 1) ^ 1);
      if (leader) {
        mbar_arrive_expect_tx(&sh.x_full[sx], kTileBytes);
#pragma unroll
        for (int c = 0; c < 4; ++c) tma_load_3d(&sh.x_tiles[sx][c * kChunkBytes], &map_x, &sh.x_full[sx], c * 64, row0, b);
      }
      __syncwarp();
      mbar_wait(&sh.y_empty[sy], ((j / kF2YStages) & 1) ^ 1);
      if (leader) {
        mbar_arrive_expect_tx(&sh.y_full[sy], kF2YBytes);
        tma_load_3d(&sh.y_tiles[sy][0], &map_y, &sh.y_full[sy], 0, row0, b);
      }
      __syncwarp();
    }
  } else if (warp == kF2Mma) {
    // ===================== MMA issuer =====================
    const bool leader = elect_one();
    constexpr uint32_t idesc_s = make_idesc_bf16(kBlockM, kBlockN, 0, 0);
    constexpr uint32_t idesc_acc = make_idesc_bf16(kBlockM, 64, 0, 1);
    const uint32_t x_lo0 = desc_lo_sw128(smem_u32(&sh.x_tiles[0][0]), 16);
    const uint32_t y_lo0 = desc_lo_sw128(smem_u32(&sh.y_tiles[0][0]), kChunkBytes);
    auto sbuf_col = [&](int t) { return kF2ColS + (uint32_t)(((t & 1) * 2 + ((t >> 1) & 1)) * 64); };   // tile t -> (group, use) buffer
    auto issue_scores = [&](int t) {
      const int s = t % kF2XStages;
      mbar_wait(&sh.x_full[s], (t / kF2XStages) & 1);
      tc_fence_after();
      if (leader) {
        const uint32_t xlo = x_lo0 + s * (kTileBytes >> 4);
        const uint32_t d = tmem + sbuf_col(t);
#pragma unroll
        for (int ks = 0; ks < kD / 16; ++ks)
          umma_ts_lohi(d, tmem + kF2ColA + ks * 8, xlo + (ks >> 2) * (kChunkBytes >> 4) + (ks & 3) * 2, kDescHiSw128_1024, idesc_s, ks > 0);
        umma_commit(&sh.x_empty[s]);
        umma_commit(&sh.s_full[(t & 1) * 2 + ((t >> 1) & 1)]);
      }
      __syncwarp();
    };
    mbar_wait(&sh.a_ready, 0);
    tc_fence_after();
    for (int t = 0; t < 4 && t < nt; ++t) {
      // the Q-staging stages (2, 3) are reloaded by the producer only after a_ready, which has been observed
      issue_scores(t);
    }
    for (int j = 0; j < nt; ++j) {
      const int s = j % kF2YStages;
      const int g = j & 1, buf = g * 2 + ((j >> 1) & 1);
      mbar_wait(&sh.p_ready[buf], (j >> 2) & 1);
      mbar_wait(&sh.y_full[s], (j / kF2YStages) & 1);
      tc_fence_after();
      if (leader) {
        const uint32_t ylo = y_lo0 + s * (kF2YBytes >> 4);
        const uint32_t pa = tmem + sbuf_col(j);
#pragma unroll
        for (int ks = 0; ks < kBlockN / 16; ++ks)
          umma_ts_lohi(tmem + (g ? kF2ColAcc1 : kF2ColAcc0), pa + p_col_of_kstep(ks), ylo + ks * (2048 >> 4), kDescHiSw128_1024, idesc_acc,
                       (j > 1) || (ks > 0));
        umma_commit(&sh.y_empty[s]);
        umma_commit(&sh.acc_done[g]);
      }
      __syncwarp();
      if (j + 4 < nt) issue_scores(j + 4);     // overwrites the score buffer whose probabilities the MMAs just issued read
    }
    if (PROJ) {
      mbar_wait(&sh.w_full, 0);
      mbar_wait(&sh.o_ready, 0);
      tc_fence_after();
      if (leader) {
        constexpr uint32_t idesc_p = make_idesc_bf16(kBlockM, 128, 0, 0);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const uint32_t wlo = desc_lo_sw128(smem_u32(&sh.w_tile[h * kSlabBytes]), 16);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_ts_lohi(tmem + kF2ColProjD + h * 128, tmem + kF2ColProjA + ks * 8, wlo + ks * 2, kDescHiSw128_1024, idesc_p, ks > 0);
        }
        umma_commit(&sh.proj_done);
      }
      __syncwarp();
    }
  } else {
    // ===================== softmax groups: warps 0-7 even tiles, warps 8-15 odd tiles =====================
    const int group = warp >> 3;
    const int quarter = warp & 3;
    const int half = (warp >> 2) & 1;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = tmem + (uint32_t(quarter * 32) << 16);
    const long long a_row_idx = (long long)a_tile * kBlockM + row;
    const bool row_valid = a_row_idx < p.La;
    {   // Q: shared -> registers -> TMEM; warp (quarter, part = warp / 4) moves slab `part` (64 features) of its 32 rows
      const int part = warp >> 2;
      mbar_wait(&sh.a_full, 0);
      const uint32_t region = smem_u32(&sh.x_tiles[2 + (part >> 1)][0]) + (part & 1) * kSlabBytes + row * 128;
      uint32_t r[32];
#pragma unroll
      for (int v = 0; v < 8; ++v) {
        const uint4 u = lds128(region + ((v ^ (row & 7)) << 4));
        r[4 * v + 0] = u.x; r[4 * v + 1] = u.y; r[4 * v + 2] = u.z; r[4 * v + 3] = u.w;
      }
      SAM2B200_TMEM_ST32(lane_addr + kF2ColA + part * 32, r);
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(&sh.a_ready);
    }
    const float c = p.scale_log2;
    const uint32_t drop_key = DROP ? sam2b200::dropout_key(*p.drop.seed, p.drop.site) : 0u;
    float m_ref = -INFINITY, l = 0.f, lk = 0.f;
    const uint32_t acc_col = group ? kF2ColAcc1 : kF2ColAcc0;
    int k = 0;                                              // index of the tile within this group
    for (int j = group; j < nt; j += 2, ++k) {
      const int use = k & 1, buf = group * 2 + use;
      const uint32_t sbuf = lane_addr + kF2ColS + buf * 64;
      mbar_wait(&sh.s_full[buf], (k >> 1) & 1);
      tc_fence_after();
      uint32_t r0[32];
      SAM2B200_TMEM_LD32(sbuf + half * kHalfN, r0);
      tmem_wait_ld();
      float sv[kHalfN];
#pragma unroll
      for (int i = 0; i < kHalfN; ++i) sv[i] = __uint_as_float(r0[i]);
      const int ncols = p.Lx - j * kBlockN - half * kHalfN;
      if (ncols < kHalfN) {
#pragma unroll
        for (int i = 0; i < kHalfN; ++i) if (i >= ncols) sv[i] = -INFINITY;
      }
      float mx = sv[0];
#pragma unroll
      for (int i = 1; i < kHalfN; ++i) mx = fmaxf(mx, sv[i]);
      sh.xchg[group][use][half][row] = mx;
      asm volatile("bar.sync %0, 64;" ::"r"(1 + group * 4 + quarter) : "memory");
      mx = fmaxf(mx, sh.xchg[group][use][half ^ 1][row]);
      bool acc_synced = false;
      const bool grow = (mx - m_ref) * c > 8.0f;
      if (k == 0) {
        m_ref = mx;
      } else if (__any_sync(0xffffffffu, grow)) {
        const float m_new = grow ? mx : m_ref;
        const float f = ex2((m_ref - m_new) * c);
        mbar_wait(&sh.acc_done[group], (k - 1) & 1);       // this group's previous P . mem has landed
        acc_synced = true;
        tc_fence_after();
        uint32_t o[32];
        SAM2B200_TMEM_LD32(lane_addr + acc_col + half * 32, o);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
        SAM2B200_TMEM_ST32(lane_addr + acc_col + half * 32, o);
        tmem_wait_st();
        l *= f; lk *= f;
        m_ref = m_new;
      }
      const float mc = m_ref * c;
      float sum0 = 0.f, sum1 = 0.f;
      const uint32_t didx = (uint32_t)(((long long)b * p.La + a_row_idx) * p.Lx) + (uint32_t)(j * kBlockN + half * kHalfN);
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < kHalfN; i += 2) {
        float e0 = ex2(fmaf(sv[i], c, -mc));
        float e1 = ex2(fmaf(sv[i + 1], c, -mc));
        sum0 += e0; sum1 += e1;
        if (DROP) {
          e0 = sam2b200::dropout_keep(drop_key, didx + i, p.drop.thresh) ? e0 : 0.f;
          e1 = sam2b200::dropout_keep(drop_key, didx + i + 1, p.drop.thresh) ? e1 : 0.f;
          lk += e0 + e1;
        }
        pk[i >> 1] = pack_bf16(e0, e1);
      }
      l += sum0 + sum1;
      SAM2B200_TMEM_ST16(sbuf + half * kHalfN, pk);
      tmem_wait_st();
      if (k > 0 && !acc_synced) mbar_wait(&sh.acc_done[group], (k - 1) & 1);   // observe every phase in order
      tc_fence_before();
      mbar_arrive(&sh.p_ready[buf]);
    }
    // ---------------- epilogue ----------------
    if (k > 0) { mbar_wait(&sh.acc_done[group], (k - 1) & 1); tc_fence_after(); }
    sh.lsum[group][half][row] = l;
    asm volatile("bar.sync %0, 64;" ::"r"(1 + group * 4 + quarter) : "memory");
    l += sh.lsum[group][half ^ 1][row];
    if (DROP) {
      sh.xchg[group][0][half][row] = lk;
      asm volatile("bar.sync %0, 64;" ::"r"(1 + group * 4 + quarter) : "memory");
      lk += sh.xchg[group][0][half ^ 1][row];
    }
    if (group == 1 && half == 0) { sh.gstat[0][row] = (k > 0) ? m_ref * c : -INFINITY; sh.gstat[1][row] = l; sh.gstat[2][row] = lk; }
    asm volatile("bar.sync 9, 512;" ::: "memory");          // both groups are done with the loop; group 1's statistics are visible
    if (group == 0) {
      const float m0 = m_ref * c, m1 = sh.gstat[0][row];
      const float mm = fmaxf(m0, m1);
      const float f0 = ex2(m0 - mm), f1 = (m1 == -INFINITY) ? 0.f : ex2(m1 - mm);
      const float lt = l * f0 + sh.gstat[1][row] * f1;
      const float lkt = lk * f0 + sh.gstat[2][row] * f1;
      const float inv_l = p.drop.inv_keep / lt;
      const float w0 = f0 * inv_l, w1 = f1 * inv_l;
      uint32_t o0[32], o1[32];
      SAM2B200_TMEM_LD32(lane_addr + kF2ColAcc0 + half * 32, o0);
      SAM2B200_TMEM_LD32(lane_addr + kF2ColAcc1 + half * 32, o1);
      tmem_wait_ld();
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = (nt > 1) ? fmaf(__uint_as_float(o1[i]), w1, __uint_as_float(o0[i]) * w0) : __uint_as_float(o0[i]) * w0;
      if (PROJ) {
        uint32_t pk2[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) pk2[i] = pack_bf16(v[2 * i], v[2 * i + 1]);
        SAM2B200_TMEM_ST16(lane_addr + kF2ColProjA + half * 16, pk2);
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(&sh.o_ready);
      }
      if (row_valid) {
        const long long off = ((long long)b * p.La + a_row_idx) * 64 + half * 32;
        uint4* o16 = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out_small) + off);
#pragma unroll
        for (int i = 0; i < 4; ++i)
          o16[i] = make_uint4(pack_bf16(v[8 * i], v[8 * i + 1]), pack_bf16(v[8 * i + 2], v[8 * i + 3]), pack_bf16(v[8 * i + 4], v[8 * i + 5]),
                              pack_bf16(v[8 * i + 6], v[8 * i + 7]));
        if (p.out_small_f32 != nullptr) {
          float4* o32 = reinterpret_cast<float4*>(p.out_small_f32 + off);
#pragma unroll
          for (int i = 0; i < 8; ++i) o32[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        }
        if (half == 0) {
          p.lse2[(long long)b * p.La + a_row_idx] = mm + log2f(lt);
          if (DROP && p.rowsum_drop != nullptr) p.rowsum_drop[(long long)b * p.La + a_row_idx] = lkt * inv_l;
        }
      }
      if (PROJ) {
        const int row0 = a_tile * kBlockM + quarter * 32;
        const uint32_t stage = smem_u32(&sh.x_tiles[0][0]) + warp * (2 * kBoxBytes);     // the ring is idle: 8 KB per warp
        mbar_wait(&sh.proj_done, 0);
        tc_fence_after();
        proj_store(p, &map_p, stage, lane_addr + kF2ColProjD + half * 128, half, lane, row0, b, DROP ? lkt * inv_l : 0.f);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kF2Mma) tmem_dealloc(tmem, 512);
}

}  // namespace attn
