// Fused per-frame multi-object mask loss for sm_100a (HBM-bound).
//
// Replaces, for the shapes the SAM2 training wrapper produces (one step per frame, one mask per
// channel), the reference's ~40 ATen kernels per frame:
//   sam2_video/model/losses.py:20-34   dice_loss
//   sam2_video/model/losses.py:37-57   sigmoid_focal_loss
//   sam2_video/model/losses.py:60-76   iou_loss
//   sam2_video/model/losses.py:143-238 MultiStepMultiMasksAndIous._update_losses (valid filter)
//   sam2_video/model/losses.py:308-372 BCECategoryLoss.forward
//
// Forward: ONE launch, one pass over logits (fp32, 4 B/px) + targets (u8, 1 B/px) for all frames x
// channels.  Every warp instruction reads 512 contiguous bytes of logits (lane-interleaved 128-bit
// loads) and 128 contiguous bytes of targets; 20 independent loads are in flight per thread before any
// math.  Per-thread register accumulation, warp-shuffle + shared-memory block reduction, one partial
// record per block (no floating-point atomics -> deterministic).  The LAST block of a channel (ticket
// counter) folds that channel's records in fp64 in a fixed order, and the last channel of the launch
// does the valid-channel filter, Nv, the dice / focal / IoU algebra and the sum over channels and
// frames -- there is no second kernel.
// Backward: one pass reading logits + targets and writing dlogits (9 B/px), using the per-channel
// sums of the forward (dice needs full-image sums, so it cannot be fused into the forward pass).
//
// Arithmetic.  With z = (t ? -x : x) / T_logit:  q = sigmoid(z) = 1 - p_t and softplus(z) = BCE(x, t), so
//   focal = alpha_t * softplus(z) * q^gamma,   p = t ? 1 - q : q,
//   d focal / dx = sgn * alpha_t * q^gamma * (q + gamma * softplus(z) * (1 - q)),  sgn = t ? -1 : +1,
// i.e. 3 SFU operations per pixel (ex2, rcp, lg2) and ~20 issue slots: at 5 B/px the forward needs
// 4.7 px/clk/SM to saturate HBM and the SFU pipe delivers 5.3.
#include <cuda_runtime.h>
#include <stdint.h>

#include "abi_common.cuh"
#include "loss_math.cuh"

namespace {
using namespace lossmath;

constexpr int kThreads = 256;
constexpr int kPxPerThreadIter = 16;                 // 4 x (float4 + u32) per thread per iteration
// Work per block, measured on B200 (scripts/loss_sweep.sh, profiles/r1_loss_sweep.txt): the forward pays a
// fixed per-block cost (block reduction, fence, ticket), so it takes 16384 px per block in 4 rounds of
// 8 loads per thread; the backward has no epilogue and is fastest with the smallest blocks (4096 px).
#ifndef SAM2B200_LOSS_FWD_ROUNDS
#define SAM2B200_LOSS_FWD_ROUNDS 4
#endif
#ifndef SAM2B200_LOSS_BWD_ROUNDS
#define SAM2B200_LOSS_BWD_ROUNDS 1
#endif
constexpr int kItersPerRound = 1;                    // loads of one round are all issued before its math
constexpr int kIterPx = kThreads * kPxPerThreadIter;                 // 4096
constexpr int kFwdRounds = SAM2B200_LOSS_FWD_ROUNDS;
constexpr int kBwdRounds = SAM2B200_LOSS_BWD_ROUNDS;
constexpr int kFwdChunk = kIterPx * kItersPerRound * kFwdRounds;     // px per forward block
constexpr int kBwdChunk = kIterPx * kItersPerRound * kBwdRounds;     // px per backward block
static_assert(kFwdChunk / kThreads <= 255 && kBwdChunk / kThreads <= 255, "8-bit per-thread counters");
constexpr int kMaxFrames = 128;                      // frames per launch (pointer tables in the kernel parameters: 2 x 1 KB + 1 KB for the backward)
constexpr int kNumSums = 6;                          // focal|bce, p*t, p, t, inter, union
constexpr int kRec = 8;                              // floats per block record

struct FramePtrs {
  const float* logits[kMaxFrames];
  const uint8_t* targets[kMaxFrames];   // frame f: [C, HW] bytes (frames of several clips may come from different tensors)
};
struct FrameOutPtrs {
  float* dlogits[kMaxFrames];
};

struct FwdParams {
  const float* pos_weight;     // [C] or null (BCE)
  const float* iou_pred;       // [tt, C] (multistep)
  float* records;              // [tt*C][nblk][kRec]
  int* chan_ticket;            // [tt*C], zero on entry
  int* done_ticket;            // [1], zero on entry
  float* chan_sums;            // [tt*C][6]
  int* n_valid;                // [tt]
  float* losses;               // [4]
  long long HW;
  int tt, C, frame0;
  float inv_temp, alpha, gamma;
  int iou_l1, reduction_mean, accumulate, vec_ok;
};

// grid: (blocks_per_channel, tt*C)
template <int MODE, bool G2, bool UNIT_T>
__global__ void __launch_bounds__(kThreads)
mask_loss_fwd_kernel(const __grid_constant__ FramePtrs fp, const __grid_constant__ FwdParams P) {
  const int fc = blockIdx.y;
  const int f = fc / P.C, c = fc % P.C;
  const long long HW = P.HW;
  const float* __restrict__ x = fp.logits[f] + (long long)c * HW;
  const uint8_t* __restrict__ tg = fp.targets[f] + (long long)c * HW;
  const long long begin = (long long)blockIdx.x * kFwdChunk;
  const long long end = (begin + kFwdChunk < HW) ? (begin + kFwdChunk) : HW;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  Acc<MODE, G2, UNIT_T> acc;
  acc.init();
  if (P.vec_ok && end - begin == kFwdChunk) {
#pragma unroll 1
    for (int r = 0; r < kFwdRounds; ++r) {
      float4 xv[kItersPerRound][4];
      uint32_t tv[kItersPerRound][4];
#pragma unroll
      for (int it = 0; it < kItersPerRound; ++it) {
        // warp `warp` owns 512 consecutive px of this iteration; load j covers 128 of them (512 B of logits)
        const long long base = begin + (long long)(r * kItersPerRound + it) * kIterPx + warp * 512 + lane * 4;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          tv[it][j] = ldg_u32(tg + base + j * 128);
          xv[it][j] = ldg_f4(x + base + j * 128);
        }
      }
#pragma unroll
      for (int it = 0; it < kItersPerRound; ++it) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t tw = tv[it][j];
          acc.add(xv[it][j].x, (tw & 0x000000ffu) != 0u, P.inv_temp, P.gamma);
          acc.add(xv[it][j].y, (tw & 0x0000ff00u) != 0u, P.inv_temp, P.gamma);
          acc.add(xv[it][j].z, (tw & 0x00ff0000u) != 0u, P.inv_temp, P.gamma);
          acc.add(xv[it][j].w, (tw & 0xff000000u) != 0u, P.inv_temp, P.gamma);
        }
      }
    }
  } else {
    // ragged tail / unaligned rows: <= 64 px per thread, scalar loads
    for (long long i = begin + threadIdx.x; i < end; i += kThreads) acc.add(x[i], tg[i] != 0, P.inv_temp, P.gamma);
  }

  // ---- block record: [f0, f1, a, bq, T, P, I, 0] ----
  // every thread parks its 7 values in shared memory ([slot][thread]: conflict-free); warp w then folds slot w
  // (8 values per lane in a fixed order + one shuffle tree) -- ~4x fewer instructions than 7 shuffle trees per warp.
  __shared__ float red[kRec][kThreads];
  __shared__ double tot[kRec];
  __shared__ int s_flag;
  red[0][threadIdx.x] = acc.f0;
  red[1][threadIdx.x] = acc.f1;
  red[2][threadIdx.x] = (MODE == 0) ? acc.a : 0.f;
  red[3][threadIdx.x] = (MODE == 0) ? acc.bq : 0.f;
  red[4][threadIdx.x] = (float)(acc.cnt & 0xffu);
  red[5][threadIdx.x] = (float)((acc.cnt >> 8) & 0xffu);
  red[6][threadIdx.x] = (float)(acc.cnt >> 16);
  __syncthreads();
  const int nblk = gridDim.x;
  {
    float v = 0.f;
    if (warp < 7) {
#pragma unroll
      for (int i = 0; i < kThreads / 32; ++i) v += red[warp][i * 32 + lane];
      v = warp_sum(v);
    }
    if (lane == 0) {
      P.records[((long long)fc * nblk + blockIdx.x) * kRec + warp] = v;
      __threadfence();
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) s_flag = (atomicAdd(&P.chan_ticket[fc], 1) == nblk - 1);
  __syncthreads();
  if (!s_flag) return;

  // ---- last block of this channel: fold its records (fixed order, fp64) ----
  __threadfence();
  {
    const float* rec = P.records + (long long)fc * nblk * kRec;
    double s = 0.0;
    for (int b = lane; b < nblk; b += 32) s += (double)__ldcg(rec + (long long)b * kRec + warp);   // warp w folds slot w
    s = warp_sum_d(s);
    if (lane == 0) tot[warp] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float* cs = P.chan_sums + (long long)fc * kNumSums;
    const double T = tot[4];
    if (MODE == 0) {
      const double F = (P.alpha >= 0.f) ? (double)P.alpha * tot[1] + (1.0 - (double)P.alpha) * tot[0] : tot[0] + tot[1];
      cs[0] = (float)F;
      cs[1] = (float)(T - tot[2]);                      // sum p t   = sum_{t=1} (1 - q)
      cs[2] = (float)(tot[3] + T - 2.0 * tot[2]);       // sum p     = sum_{t=0} q + sum_{t=1} (1 - q)
      cs[3] = (float)T;
      cs[4] = (float)tot[6];
      cs[5] = (float)(tot[5] + T - tot[6]);             // |pred or gt|
    } else {
      const double pw = (P.pos_weight != nullptr) ? (double)P.pos_weight[c] : 1.0;
      cs[0] = (float)(tot[0] + pw * tot[1]);
      cs[1] = 0.f; cs[2] = 0.f; cs[3] = (float)T; cs[4] = 0.f; cs[5] = 0.f;
    }
    P.chan_ticket[fc] = 0;                               // leave the ticket region zeroed for the next launch
    __threadfence();
    s_flag = (atomicAdd(P.done_ticket, 1) == P.tt * P.C - 1);
  }
  __syncthreads();
  if (!s_flag) return;

  // ---- last channel of the launch: valid filter, Nv, loss algebra, sums over channels and frames ----
  __threadfence();
  __shared__ int nv_sh[kMaxFrames];
  __shared__ double part[3][kThreads / 32];
  const int n = P.tt * P.C;
  for (int i = threadIdx.x; i < P.tt; i += kThreads) nv_sh[i] = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += kThreads)
    if (__ldcg(P.chan_sums + (long long)i * kNumSums + 3) > 0.f) atomicAdd(&nv_sh[i / P.C], 1);
  __syncthreads();
  double lm = 0.0, ld = 0.0, li = 0.0;
  if (MODE == 0) {
    for (int i = threadIdx.x; i < n; i += kThreads) {
      const float* s = P.chan_sums + (long long)i * kNumSums;
      const double s0 = __ldcg(s), s1 = __ldcg(s + 1), s2 = __ldcg(s + 2), s3 = __ldcg(s + 3), s4 = __ldcg(s + 4), s5 = __ldcg(s + 5);
      const int nv = nv_sh[i / P.C];
      if (!(s3 > 0.0) || nv == 0) continue;             // no foreground: filtered out (losses.py:153-159)
      lm += s0 / (double)HW / nv;
      ld += (1.0 - (2.0 * s1 + 1.0) / (s2 + s3 + 1.0)) / nv;
      const double dd = (double)P.iou_pred[i] - s4 / fmax(s5, 1.0);
      li += (P.iou_l1 ? fabs(dd) : dd * dd) / nv;
    }
  } else {
    // per frame: sum_{valid c} S0 / (mean ? Nv * HW : 1); a frame with no valid channel is NaN under "mean"
    // (the reference takes the mean of an empty selection, losses.py:365)
    for (int i = threadIdx.x; i < n; i += kThreads) {
      const float* s = P.chan_sums + (long long)i * kNumSums;
      const double s0 = __ldcg(s), s3 = __ldcg(s + 3);
      const int nv = nv_sh[i / P.C];
      if (s3 > 0.0) lm += P.reduction_mean ? s0 / ((double)nv * (double)HW) : s0;
    }
    if (P.reduction_mean)
      for (int i = threadIdx.x; i < P.tt; i += kThreads)
        if (nv_sh[i] == 0) lm += __longlong_as_double(0x7ff8000000000000LL);
  }
  lm = warp_sum_d(lm); ld = warp_sum_d(ld); li = warp_sum_d(li);
  if (lane == 0) { part[0][warp] = lm; part[1][warp] = ld; part[2][warp] = li; }
  __syncthreads();
  for (int i = threadIdx.x; i < P.tt; i += kThreads) P.n_valid[i] = nv_sh[i];
  if (threadIdx.x == 0) {
    *P.done_ticket = 0;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    for (int w = 0; w < kThreads / 32; ++w) { a0 += part[0][w]; a1 += part[1][w]; a2 += part[2][w]; }
    if (P.accumulate) {
      P.losses[0] += (float)a0;
      if (MODE == 0) { P.losses[1] += (float)a1; P.losses[2] += (float)a2; }
    } else {
      P.losses[0] = (float)a0;
      if (MODE == 0) { P.losses[1] = (float)a1; P.losses[2] = (float)a2; P.losses[3] = 0.f; }
    }
  }
}

struct BwdParams {
  const float* pos_weight;
  const float* chan_sums;      // [tt*C][6]
  const int* n_valid;          // [tt]
  const float* gout;           // [3] d/d(loss_mask, loss_dice, loss_iou)  or [1] (BCE)
  const float* iou_pred;       // [tt*C]
  float* diou;                 // [tt*C]
  const float* raw_coef;       // [tt*C][3] or null.  Non-null ("functional" backward): per-channel coefficients
                               // (d/d sum_focal, d/d sum_pt + d/d sum_p, d/d sum_p) given directly, no valid filter
  long long HW;
  int C, frame0;
  float inv_temp, alpha, gamma;
  int iou_l1, reduction_mean, vec_ok;
};

// Backward.  grid: (blocks_per_channel, tt*C).
template <int MODE, bool G2, bool UNIT_T>
__global__ void __launch_bounds__(kThreads)
mask_loss_bwd_kernel(const __grid_constant__ FramePtrs fp, const __grid_constant__ FrameOutPtrs op,
                     const __grid_constant__ BwdParams P) {
  const int fc = blockIdx.y;
  const int f = fc / P.C, c = fc % P.C;
  const long long HW = P.HW;
  const float* __restrict__ x = fp.logits[f] + (long long)c * HW;
  float* __restrict__ dx = op.dlogits[f] + (long long)c * HW;
  const uint8_t* __restrict__ tg = fp.targets[f] + (long long)c * HW;
  const bool raw = P.raw_coef != nullptr;
  const float* s = raw ? nullptr : P.chan_sums + (long long)fc * kNumSums;
  const int nv = raw ? 1 : P.n_valid[f];
  const bool valid = raw || (s[3] > 0.f && nv > 0);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  // per-channel coefficients: grad = kf_t * q^gamma * (q + gamma softplus (1-q)) + dc_t * q (1-q)      (MODE 0)
  //                           grad = kb_t * q                                                            (MODE 1)
  float kf_fg = 0.f, kf_bg = 0.f, dc_fg = 0.f, dc_bg = 0.f;
  if (raw) {
    const float* cf = P.raw_coef + (long long)fc * 3;
    kf_fg = -cf[0] * P.inv_temp * ((P.alpha >= 0.f) ? P.alpha : 1.0f);
    kf_bg = cf[0] * P.inv_temp * ((P.alpha >= 0.f) ? (1.0f - P.alpha) : 1.0f);
    dc_fg = cf[1] * P.inv_temp;
    dc_bg = cf[2] * P.inv_temp;
  } else if (valid) {
    if (MODE == 0) {
      const float inv_nv_t = P.inv_temp / (float)nv;
      const float kf = P.gout[0] * inv_nv_t / (float)HW;
      kf_fg = -kf * ((P.alpha >= 0.f) ? P.alpha : 1.0f);
      kf_bg = kf * ((P.alpha >= 0.f) ? (1.0f - P.alpha) : 1.0f);
      // d dice / dx = (dice_b - dice_a t) p (1-p),  dice_a = 2 kd / (D+1),  dice_b = kd (Nn+1) / (D+1)^2
      const float dp1 = s[2] + s[3] + 1.0f;
      const float nn1 = 2.0f * s[1] + 1.0f;
      const float kd = P.gout[1] * inv_nv_t;
      dc_bg = kd * nn1 / (dp1 * dp1);
      dc_fg = dc_bg - kd * 2.0f / dp1;
    } else {
      const float pw = (P.pos_weight != nullptr) ? P.pos_weight[c] : 1.0f;
      const float kb = P.gout[0] * P.inv_temp * (P.reduction_mean ? 1.0f / ((float)nv * (float)HW) : 1.0f);
      kf_fg = -kb * pw;
      kf_bg = kb;
    }
  }
  if (MODE == 0 && !raw && blockIdx.x == 0 && threadIdx.x == 0) {
    float g = 0.f;
    if (valid) {
      const float actual = s[4] / fmaxf(s[5], 1.0f);
      const float d = P.iou_pred[fc] - actual;
      g = (P.iou_l1 ? (float)((d > 0.f) - (d < 0.f)) : 2.0f * d) * P.gout[2] / (float)nv;
    }
    P.diou[fc] = g;
  }

  auto grad = [&](float xraw, bool t) -> float {
    float d, sp;
    softplus_terms<UNIT_T>(t ? -xraw : xraw, P.inv_temp, d, sp);
    const float q = rcp_approx(d);
    const float kf = t ? kf_fg : kf_bg;
    if (MODE == 0) {
      const float omq = 1.0f - q;
      float u, qg;
      if (G2) { u = fmaf(sp + sp, omq, q); qg = q * q; }
      else { u = fmaf(P.gamma * sp, omq, q); qg = (P.gamma == 0.f) ? 1.0f : __powf(q, P.gamma); }
      return fmaf(t ? dc_fg : dc_bg, q * omq, (u * qg) * kf);
    } else {
      return kf * q;
    }
  };

  const long long begin = (long long)blockIdx.x * kBwdChunk;
  const long long end = (begin + kBwdChunk < HW) ? (begin + kBwdChunk) : HW;
  if (P.vec_ok && end - begin == kBwdChunk) {
    if (!valid) {  // masked-out channel: write zeros without reading
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int it = 0; it < kItersPerRound * kBwdRounds; ++it) {
        const long long base = begin + (long long)it * kIterPx + warp * 512 + lane * 4;
#pragma unroll
        for (int j = 0; j < 4; ++j) stg_f4(dx + base + j * 128, z);
      }
      return;
    }
#pragma unroll 1
    for (int r = 0; r < kBwdRounds; ++r) {
      float4 xv[kItersPerRound][4];
      uint32_t tv[kItersPerRound][4];
#pragma unroll
      for (int it = 0; it < kItersPerRound; ++it) {
        const long long base = begin + (long long)(r * kItersPerRound + it) * kIterPx + warp * 512 + lane * 4;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          tv[it][j] = ldg_u32(tg + base + j * 128);
          xv[it][j] = ldg_f4(x + base + j * 128);
        }
      }
#pragma unroll
      for (int it = 0; it < kItersPerRound; ++it) {
        const long long base = begin + (long long)(r * kItersPerRound + it) * kIterPx + warp * 512 + lane * 4;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t tw = tv[it][j];
          float4 o;
          o.x = grad(xv[it][j].x, (tw & 0x000000ffu) != 0u);
          o.y = grad(xv[it][j].y, (tw & 0x0000ff00u) != 0u);
          o.z = grad(xv[it][j].z, (tw & 0x00ff0000u) != 0u);
          o.w = grad(xv[it][j].w, (tw & 0xff000000u) != 0u);
          stg_f4(dx + base + j * 128, o);
        }
      }
    }
  } else {
    for (long long i = begin + threadIdx.x; i < end; i += kThreads) dx[i] = valid ? grad(x[i], tg[i] != 0) : 0.f;
  }
}

int fwd_blocks_per_channel(long long HW) { return (int)((HW + kFwdChunk - 1) / kFwdChunk); }
int bwd_blocks_per_channel(long long HW) { return (int)((HW + kBwdChunk - 1) / kBwdChunk); }

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
bool aligned4(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 3u) == 0; }

size_t records_bytes(int tt, int C, long long HW) {
  size_t b = (size_t)tt * C * fwd_blocks_per_channel(HW) * kRec * sizeof(float);
  return (b + 255) & ~size_t(255);
}
// The ticket region sits at the START of the workspace and has a FIXED size (65535 channel tickets + 1), so that a
// workspace shared by calls of different shapes never has records of one call where another expects zeroed tickets.
constexpr int kMaxChannelsPerLaunch = 65535;
constexpr size_t kTicketBytes = ((size_t)(kMaxChannelsPerLaunch + 1) * sizeof(int) + 255) & ~size_t(255);

}  // namespace

extern "C" {

// Workspace the forward needs: per-block records + the ticket counters of the in-kernel finalisation.
size_t sam2b200_mask_loss_workspace_bytes(int T, int C, long long HW) {
  int tt = T < kMaxFrames ? T : kMaxFrames;
  return kTicketBytes + records_bytes(tt, C, HW);
}

// targets: one [T, C, HW] tensor (targets_base) or one pointer per frame (target_ptrs[f] -> [C, HW]); exactly one is non-null
static int mask_loss_fwd_impl(const float* const* logits, const uint8_t* targets_base, const uint8_t* const* target_ptrs, const float* iou_pred,
                              const float* pos_weight, void* workspace, float* chan_sums, int* n_valid,
                              float* losses, int T, int C, long long HW, int mode, float alpha,
                              float gamma, float inv_temp, int iou_l1, int reduction_mean,
                              cudaStream_t stream) {
  // mode bit 8 (SAM2B200_LOSS_TICKETS_ZEROED): the caller guarantees that the ticket region of `workspace` is zero
  // (zero-initialised once and since then only used by this function, which leaves it zeroed) -> no memset node.
  const bool tickets_zeroed = (mode & 0x100) != 0;
  mode &= 0xff;
  if (T <= 0 || C <= 0 || HW <= 0 || !logits || (!targets_base) == (!target_ptrs) || !workspace || !chan_sums || !n_valid ||
      !losses || (mode == 0 && !iou_pred) || (mode != 0 && mode != 1))
    return sam2b200::fail(SAM2B200_ERR_INVALID, "mask_loss_fwd: bad arguments");
  const int nblk = fwd_blocks_per_channel(HW);
  int launches = 0;
  for (int f0 = 0; f0 < T; f0 += kMaxFrames) {
    const int tt = (T - f0 < kMaxFrames) ? (T - f0) : kMaxFrames;
    if ((long long)tt * C > kMaxChannelsPerLaunch) return sam2b200::fail(SAM2B200_ERR_INVALID, "mask_loss_fwd: frames x channels per launch > 65535");
    FramePtrs fp;
    int vec_ok = (HW % 4 == 0);
    for (int f = 0; f < tt; ++f) {
      fp.logits[f] = logits[f0 + f];
      fp.targets[f] = target_ptrs ? target_ptrs[f0 + f] : targets_base + (size_t)(f0 + f) * C * HW;
      if (!fp.logits[f] || !fp.targets[f]) return sam2b200::fail(SAM2B200_ERR_INVALID, "mask_loss_fwd: null frame");
      vec_ok = vec_ok && aligned16(fp.logits[f]) && aligned4(fp.targets[f]);
    }
    FwdParams P;
    P.pos_weight = pos_weight; P.iou_pred = iou_pred ? iou_pred + (size_t)f0 * C : nullptr;
    P.chan_ticket = static_cast<int*>(workspace);
    P.done_ticket = P.chan_ticket + kMaxChannelsPerLaunch;
    P.records = reinterpret_cast<float*>(static_cast<char*>(workspace) + kTicketBytes);
    P.chan_sums = chan_sums + (size_t)f0 * C * kNumSums;
    P.n_valid = n_valid + f0; P.losses = losses; P.HW = HW; P.tt = tt; P.C = C; P.frame0 = f0;
    P.inv_temp = inv_temp; P.alpha = alpha; P.gamma = gamma; P.iou_l1 = iou_l1; P.reduction_mean = reduction_mean;
    P.accumulate = f0 > 0; P.vec_ok = vec_ok;
    if (!tickets_zeroed) {
      cudaError_t e = cudaMemsetAsync(P.chan_ticket, 0, (size_t)tt * C * sizeof(int), stream);
      if (e == cudaSuccess) e = cudaMemsetAsync(P.done_ticket, 0, sizeof(int), stream);
      if (e != cudaSuccess) return sam2b200::fail(SAM2B200_ERR_CUDA, "mask_loss_fwd: cudaMemsetAsync failed");
    }
    dim3 grid(nblk, tt * C);
    const bool unit_t = inv_temp == 1.0f;
    if (mode == 0) {
      if (gamma == 2.0f && unit_t) mask_loss_fwd_kernel<0, true, true><<<grid, kThreads, 0, stream>>>(fp, P);
      else if (gamma == 2.0f) mask_loss_fwd_kernel<0, true, false><<<grid, kThreads, 0, stream>>>(fp, P);
      else mask_loss_fwd_kernel<0, false, false><<<grid, kThreads, 0, stream>>>(fp, P);
    } else {
      if (unit_t) mask_loss_fwd_kernel<1, true, true><<<grid, kThreads, 0, stream>>>(fp, P);
      else mask_loss_fwd_kernel<1, true, false><<<grid, kThreads, 0, stream>>>(fp, P);
    }
    ++launches;
  }
  return sam2b200::check_launch("mask_loss_fwd", launches);
}

int sam2b200_mask_loss_fwd(const float* const* logits, const uint8_t* targets, const float* iou_pred,
                           const float* pos_weight, void* workspace, float* chan_sums, int* n_valid,
                           float* losses, int T, int C, long long HW, int mode, float alpha,
                           float gamma, float inv_temp, int iou_l1, int reduction_mean,
                           cudaStream_t stream) {
  return mask_loss_fwd_impl(logits, targets, nullptr, iou_pred, pos_weight, workspace, chan_sums, n_valid, losses, T, C, HW, mode, alpha,
                            gamma, inv_temp, iou_l1, reduction_mean, stream);
}

// The same with one target pointer per frame (target_ptrs[f] -> [C, HW] bytes): the frames of SEVERAL clips -- whose targets live in
// different tensors -- go through one launch (up to 128 frames per launch); the result is the sum over all frames, i.e. over the clips.
int sam2b200_mask_loss_fwd_frames(const float* const* logits, const uint8_t* const* target_ptrs, const float* iou_pred,
                                  const float* pos_weight, void* workspace, float* chan_sums, int* n_valid,
                                  float* losses, int T, int C, long long HW, int mode, float alpha,
                                  float gamma, float inv_temp, int iou_l1, int reduction_mean,
                                  cudaStream_t stream) {
  return mask_loss_fwd_impl(logits, nullptr, target_ptrs, iou_pred, pos_weight, workspace, chan_sums, n_valid, losses, T, C, HW, mode, alpha,
                            gamma, inv_temp, iou_l1, reduction_mean, stream);
}

static int mask_loss_bwd_impl(const float* const* logits, float* const* dlogits, const uint8_t* targets_base,
                              const uint8_t* const* target_ptrs, const float* iou_pred, const float* pos_weight, const float* chan_sums,
                              const int* n_valid, const float* grad_losses, float* diou, const float* raw_coef, int T, int C,
                              long long HW, int mode, float alpha, float gamma, float inv_temp,
                              int iou_l1, int reduction_mean, cudaStream_t stream) {
  const int nblk = bwd_blocks_per_channel(HW);
  int launches = 0;
  for (int f0 = 0; f0 < T; f0 += kMaxFrames) {
    const int tt = (T - f0 < kMaxFrames) ? (T - f0) : kMaxFrames;
    FramePtrs fp;
    FrameOutPtrs op;
    int vec_ok = (HW % 4 == 0);
    for (int f = 0; f < tt; ++f) {
      fp.logits[f] = logits[f0 + f];
      fp.targets[f] = target_ptrs ? target_ptrs[f0 + f] : targets_base + (size_t)(f0 + f) * C * HW;
      op.dlogits[f] = dlogits[f0 + f];
      if (!fp.logits[f] || !op.dlogits[f] || !fp.targets[f])
        return sam2b200::fail(SAM2B200_ERR_INVALID, "mask_loss_bwd: null frame");
      vec_ok = vec_ok && aligned16(fp.logits[f]) && aligned16(op.dlogits[f]) && aligned4(fp.targets[f]);
    }
    BwdParams P;
    P.pos_weight = pos_weight; P.chan_sums = chan_sums ? chan_sums + (size_t)f0 * C * kNumSums : nullptr;
    P.n_valid = n_valid ? n_valid + f0 : nullptr; P.gout = grad_losses;
    P.iou_pred = iou_pred ? iou_pred + (size_t)f0 * C : nullptr;
    P.diou = diou ? diou + (size_t)f0 * C : nullptr;
    P.raw_coef = raw_coef ? raw_coef + (size_t)f0 * C * 3 : nullptr;
    P.HW = HW; P.C = C; P.frame0 = f0; P.inv_temp = inv_temp; P.alpha = alpha; P.gamma = gamma;
    P.iou_l1 = iou_l1; P.reduction_mean = reduction_mean; P.vec_ok = vec_ok;
    dim3 grid(nblk, tt * C);
    const bool unit_t = inv_temp == 1.0f;
    if (mode == 0) {
      if (gamma == 2.0f && unit_t) mask_loss_bwd_kernel<0, true, true><<<grid, kThreads, 0, stream>>>(fp, op, P);
      else if (gamma == 2.0f) mask_loss_bwd_kernel<0, true, false><<<grid, kThreads, 0, stream>>>(fp, op, P);
      else mask_loss_bwd_kernel<0, false, false><<<grid, kThreads, 0, stream>>>(fp, op, P);
    } else {
      if (unit_t) mask_loss_bwd_kernel<1, true, true><<<grid, kThreads, 0, stream>>>(fp, op, P);
      else mask_loss_bwd_kernel<1, true, false><<<grid, kThreads, 0, stream>>>(fp, op, P);
    }
    ++launches;
  }
  return sam2b200::check_launch("mask_loss_bwd", launches);
}

int sam2b200_mask_loss_bwd(const float* const* logits, float* const* dlogits, const uint8_t* targets,
                           const float* iou_pred, const float* pos_weight, const float* chan_sums,
                           const int* n_valid, const float* grad_losses, float* diou, int T, int C,
                           long long HW, int mode, float alpha, float gamma, float inv_temp,
                           int iou_l1, int reduction_mean, cudaStream_t stream) {
  if (T <= 0 || C <= 0 || HW <= 0 || !logits || !dlogits || !targets || !chan_sums || !n_valid ||
      !grad_losses || (mode == 0 && (!iou_pred || !diou)) || (mode != 0 && mode != 1))
    return sam2b200::fail(SAM2B200_ERR_INVALID, "mask_loss_bwd: bad arguments");
  return mask_loss_bwd_impl(logits, dlogits, targets, nullptr, iou_pred, pos_weight, chan_sums, n_valid, grad_losses, diou, nullptr, T, C,
                            HW, mode, alpha, gamma, inv_temp, iou_l1, reduction_mean, stream);
}

// Backward of sam2b200_mask_loss_fwd_frames (one target pointer per frame).
int sam2b200_mask_loss_bwd_frames(const float* const* logits, float* const* dlogits, const uint8_t* const* target_ptrs,
                                  const float* iou_pred, const float* pos_weight, const float* chan_sums,
                                  const int* n_valid, const float* grad_losses, float* diou, int T, int C,
                                  long long HW, int mode, float alpha, float gamma, float inv_temp,
                                  int iou_l1, int reduction_mean, cudaStream_t stream) {
  if (T <= 0 || C <= 0 || HW <= 0 || !logits || !dlogits || !target_ptrs || !chan_sums || !n_valid ||
      !grad_losses || (mode == 0 && (!iou_pred || !diou)) || (mode != 0 && mode != 1))
    return sam2b200::fail(SAM2B200_ERR_INVALID, "mask_loss_bwd_frames: bad arguments");
  return mask_loss_bwd_impl(logits, dlogits, nullptr, target_ptrs, iou_pred, pos_weight, chan_sums, n_valid, grad_losses, diou, nullptr, T, C,
                            HW, mode, alpha, gamma, inv_temp, iou_l1, reduction_mean, stream);
}

// Backward of the per-channel sums themselves (functional forms dice_loss / sigmoid_focal_loss, losses.py:20-57):
// dlogits = coef[c][0] * d(sum focal_c)/dx + (t ? coef[c][1] : coef[c][2]) * p (1 - p), coef: [T, C, 3] fp32 with
// coef[.][1] = d/d(sum p t) + d/d(sum p), coef[.][2] = d/d(sum p).  No valid-channel filter.
int sam2b200_mask_loss_bwd_coef(const float* const* logits, float* const* dlogits, const uint8_t* targets,
                                const float* coef, int T, int C, long long HW, float alpha, float gamma, float inv_temp,
                                cudaStream_t stream) {
  if (T <= 0 || C <= 0 || HW <= 0 || !logits || !dlogits || !targets || !coef)
    return sam2b200::fail(SAM2B200_ERR_INVALID, "mask_loss_bwd_coef: bad arguments");
  return mask_loss_bwd_impl(logits, dlogits, targets, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, coef, T, C, HW, 0,
                            alpha, gamma, inv_temp, 0, 1, stream);
}

}  // extern "C"
