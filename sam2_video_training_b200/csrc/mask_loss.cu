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
// Forward: ONE pass over logits (fp32, 4 B/px) + targets (u8, 1 B/px) for all frames x channels,
// 128-bit loads, per-thread register accumulation of the six per-channel sums, warp-shuffle +
// shared-memory block reduction, one partial row per block (no atomics, deterministic), then a
// tiny finalize kernel (valid filter, Nv, dice/focal/IoU algebra, sum over channels and frames).
// Backward: one pass reading logits + targets and writing dlogits (9 B/px), using the per-channel
// sums of the forward (dice needs full-image sums, so it cannot be fused into the forward pass).
#include <cuda_runtime.h>
#include <stdint.h>

#include "abi_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kPxPerThreadIter = 16;                 // 4 x float4 + 1 x uint4 per iteration
constexpr int kItersPerBlock = 2;                    // -> 8192 px per block
constexpr int kChunk = kThreads * kPxPerThreadIter * kItersPerBlock;
constexpr int kMaxFrames = 64;                       // frames per launch (pointer table in params)
constexpr int kNumSums = 6;                          // focal|bce, p*t, p, t, inter, union

struct FramePtrs {
  const float* logits[kMaxFrames];
};
struct FrameOutPtrs {
  float* dlogits[kMaxFrames];
};

__device__ __forceinline__ float4 ldg_f4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ uint4 ldg_u4(const uint8_t* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_f4(float* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x),
               "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

struct PixelTerms {
  float p;    // sigmoid(x)
  float ce;   // BCE-with-logits(x, t), unweighted
  float sp;   // softplus(-x)
};

// One exp, one log, one reciprocal per pixel (3 SFU ops), overflow-free for any x.
__device__ __forceinline__ PixelTerms pixel_terms(float x, float t) {
  PixelTerms r;
  float e = __expf(-fabsf(x));
  float inv = __frcp_rn(1.0f + e);
  r.p = (x >= 0.f) ? inv : e * inv;
  float l1p = __logf(1.0f + e);
  r.sp = fmaxf(-x, 0.f) + l1p;               // softplus(-x)
  r.ce = fmaf(1.0f - t, x, r.sp);            // (1-t) x + softplus(-x)
  return r;
}

template <int MODE>  // 0 = focal+dice+iou sums, 1 = BCE category sums
struct Acc {
  float s[kNumSums];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int i = 0; i < kNumSums; ++i) s[i] = 0.f;
  }
  __device__ __forceinline__ void add(float xraw, float t, float inv_temp, float alpha, float gamma,
                                      float pos_w) {
    float x = xraw * inv_temp;
    PixelTerms pt = pixel_terms(x, t);
    if (MODE == 0) {
      float q = (t != 0.f) ? (1.0f - pt.p) : pt.p;           // 1 - p_t
      float mod = (gamma == 2.0f) ? q * q : ((gamma == 0.f) ? 1.0f : __powf(q, gamma));
      float fl = pt.ce * mod;
      if (alpha >= 0.f) fl *= (t != 0.f) ? alpha : (1.0f - alpha);
      s[0] += fl;
      s[1] += pt.p * t;
      s[2] += pt.p;
      s[3] += t;
      bool pr = x > 0.f, gt = t > 0.f;
      s[4] += (pr && gt) ? 1.f : 0.f;
      s[5] += (pr || gt) ? 1.f : 0.f;
    } else {
      // (1-t) x + (1 + (pw-1) t) softplus(-x)
      float w = fmaf(pos_w - 1.0f, t, 1.0f);
      s[0] += fmaf(1.0f - t, x, w * pt.sp);
      s[3] += t;
    }
  }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// grid: (blocks_per_channel, T*C).  partials: [T*C][blocks_per_channel][kNumSums]
template <int MODE>
__global__ void __launch_bounds__(kThreads)
mask_loss_fwd_kernel(const __grid_constant__ FramePtrs fp, const uint8_t* __restrict__ targets,
                     const float* __restrict__ pos_weight, float* __restrict__ partials, int C,
                     long long HW, int frame0, float inv_temp, float alpha, float gamma, int vec_ok) {
  const int fc = blockIdx.y;
  const int f = fc / C, c = fc % C;
  const float* __restrict__ x = fp.logits[f] + (long long)c * HW;
  const uint8_t* __restrict__ tg = targets + ((long long)(frame0 + f) * C + c) * HW;
  const float pw = (MODE == 1 && pos_weight != nullptr) ? pos_weight[c] : 1.0f;
  const long long begin = (long long)blockIdx.x * kChunk;
  const long long end = (begin + kChunk < HW) ? (begin + kChunk) : HW;

  Acc<MODE> acc;
  acc.init();
  if (vec_ok && end - begin == kChunk) {
    // full chunk, 16B-aligned rows: all loads of the block issued before any math
    float4 xv[kItersPerBlock][4];
    uint4 tv[kItersPerBlock];
#pragma unroll
    for (int it = 0; it < kItersPerBlock; ++it) {
      // thread owns 16 consecutive px; a warp covers 512 px = 2 KB of logits per iteration
      long long base = begin + ((long long)it * kThreads + threadIdx.x) * kPxPerThreadIter;
      tv[it] = ldg_u4(tg + base);
#pragma unroll
      for (int j = 0; j < 4; ++j) xv[it][j] = ldg_f4(x + base + 4 * j);
    }
#pragma unroll
    for (int it = 0; it < kItersPerBlock; ++it) {
      const uint32_t tw[4] = {tv[it].x, tv[it].y, tv[it].z, tv[it].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float xs[4] = {xv[it][j].x, xv[it][j].y, xv[it][j].z, xv[it][j].w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float t = ((tw[j] >> (8 * k)) & 0xffu) ? 1.0f : 0.0f;
          acc.add(xs[k], t, inv_temp, alpha, gamma, pw);
        }
      }
    }
  } else {
    for (long long i = begin + threadIdx.x; i < end; i += kThreads) {
      float t = tg[i] ? 1.0f : 0.0f;
      acc.add(x[i], t, inv_temp, alpha, gamma, pw);
    }
  }

  __shared__ float red[kThreads / 32][kNumSums];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < kNumSums; ++i) {
    float v = warp_sum(acc.s[i]);
    if (lane == 0) red[warp][i] = v;
  }
  __syncthreads();
  if (threadIdx.x < kNumSums) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) v += red[w][threadIdx.x];
    partials[((long long)fc * gridDim.x + blockIdx.x) * kNumSums + threadIdx.x] = v;
  }
}

// One block.  chan_sums: [T*C][6] (fp32, summed in fp64 from the partials in a fixed order),
// n_valid: [T], losses: [4] = loss_mask, loss_dice, loss_iou, loss_class (frames accumulated by
// the caller passing accumulate=1 for the 2nd.. launch of a long clip).
__global__ void mask_loss_finalize_kernel(const float* __restrict__ partials, int nblk, int T, int C,
                                          long long HW, const float* __restrict__ iou_pred,
                                          int iou_l1, float* __restrict__ chan_sums,
                                          int* __restrict__ n_valid, float* __restrict__ losses,
                                          int accumulate) {
  extern __shared__ double sh[];  // [T*C][6]
  const int n = T * C;
  for (int i = threadIdx.x; i < n * kNumSums; i += blockDim.x) {
    const int fc = i / kNumSums, k = i % kNumSums;
    double s = 0.0;
    for (int b = 0; b < nblk; ++b) s += (double)partials[((long long)fc * nblk + b) * kNumSums + k];
    sh[i] = s;
    chan_sums[i] = (float)s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double lm = 0.0, ld = 0.0, li = 0.0;
    for (int f = 0; f < T; ++f) {
      int nv = 0;
      for (int c = 0; c < C; ++c) nv += sh[(f * C + c) * kNumSums + 3] > 0.0;
      n_valid[f] = nv;
      if (nv == 0) continue;  // host raises ValueError("No valid masks") (losses.py:153-161)
      for (int c = 0; c < C; ++c) {
        const double* s = &sh[(f * C + c) * kNumSums];
        if (!(s[3] > 0.0)) continue;
        lm += s[0] / (double)HW / nv;
        ld += (1.0 - (2.0 * s[1] + 1.0) / (s[2] + s[3] + 1.0)) / nv;
        double actual = s[4] / fmax(s[5], 1.0);
        double d = (double)iou_pred[f * C + c] - actual;
        li += (iou_l1 ? fabs(d) : d * d) / nv;
      }
    }
    if (accumulate) {
      losses[0] += (float)lm; losses[1] += (float)ld; losses[2] += (float)li;
    } else {
      losses[0] = (float)lm; losses[1] = (float)ld; losses[2] = (float)li; losses[3] = 0.f;
    }
  }
}

// BCE finalize: losses[0] (+)= sum_f [ sum_{valid c} S0 / (reduction_mean ? Nv*HW : 1) ]
__global__ void bce_loss_finalize_kernel(const float* __restrict__ partials, int nblk, int T, int C,
                                         long long HW, int reduction_mean,
                                         float* __restrict__ chan_sums, int* __restrict__ n_valid,
                                         float* __restrict__ losses, int accumulate) {
  extern __shared__ double sh[];
  const int n = T * C;
  for (int i = threadIdx.x; i < n * kNumSums; i += blockDim.x) {
    const int fc = i / kNumSums, k = i % kNumSums;
    double s = 0.0;
    for (int b = 0; b < nblk; ++b) s += (double)partials[((long long)fc * nblk + b) * kNumSums + k];
    sh[i] = s;
    chan_sums[i] = (float)s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int f = 0; f < T; ++f) {
      int nv = 0;
      double s0 = 0.0;
      for (int c = 0; c < C; ++c) {
        const double* s = &sh[(f * C + c) * kNumSums];
        if (s[3] > 0.0) { nv += 1; s0 += s[0]; }
      }
      n_valid[f] = nv;
      // mean over an empty selection is NaN in the reference (losses.py:365 with no valid channel)
      tot += reduction_mean ? s0 / ((double)nv * (double)HW) : s0;
    }
    if (accumulate) losses[0] += (float)tot; else losses[0] = (float)tot;
  }
}

// Backward.  grid: (blocks_per_channel, T*C).  gout: [3] = d/d loss_mask, d/d loss_dice, d/d loss_iou
// (MODE 0) or [1] = d/d (sum over frames of per-frame loss) (MODE 1).
template <int MODE>
__global__ void __launch_bounds__(kThreads)
mask_loss_bwd_kernel(const __grid_constant__ FramePtrs fp, const __grid_constant__ FrameOutPtrs op,
                     const uint8_t* __restrict__ targets, const float* __restrict__ pos_weight,
                     const float* __restrict__ chan_sums, const int* __restrict__ n_valid,
                     const float* __restrict__ gout, const float* __restrict__ iou_pred,
                     float* __restrict__ diou, int C, long long HW, int frame0, float inv_temp,
                     float alpha, float gamma, int iou_l1, int reduction_mean, int vec_ok) {
  const int fc = blockIdx.y;
  const int f = fc / C, c = fc % C;
  const float* __restrict__ x = fp.logits[f] + (long long)c * HW;
  float* __restrict__ dx = op.dlogits[f] + (long long)c * HW;
  const uint8_t* __restrict__ tg = targets + ((long long)(frame0 + f) * C + c) * HW;
  const float* s = chan_sums + (long long)fc * kNumSums;
  const int nv = n_valid[f];
  const bool valid = s[3] > 0.f && nv > 0;
  const float pw = (MODE == 1 && pos_weight != nullptr) ? pos_weight[c] : 1.0f;

  float k_focal = 0.f, dice_a = 0.f, dice_b = 0.f, k_bce = 0.f;
  if (valid) {
    if (MODE == 0) {
      const float inv_nv_t = inv_temp / (float)nv;
      k_focal = gout[0] * inv_nv_t / (float)HW;
      // d dice / dx = -(2 t (D+1) - (Nn+1)) / (D+1)^2 * p (1-p) = (dice_b - dice_a * t) p (1-p)
      const float dp1 = s[2] + s[3] + 1.0f;
      const float nn1 = 2.0f * s[1] + 1.0f;
      const float kd = gout[1] * inv_nv_t;
      dice_a = kd * 2.0f / dp1;
      dice_b = kd * nn1 / (dp1 * dp1);
    } else {
      k_bce = gout[0] * inv_temp * (reduction_mean ? 1.0f / ((float)nv * (float)HW) : 1.0f);
    }
  }
  if (MODE == 0 && blockIdx.x == 0 && threadIdx.x == 0) {
    float g = 0.f;
    if (valid) {
      float actual = s[4] / fmaxf(s[5], 1.0f);
      float d = iou_pred[fc] - actual;
      g = (iou_l1 ? ((d > 0.f) - (d < 0.f)) : 2.0f * d) * gout[2] / (float)nv;
    }
    diou[fc] = g;
  }

  auto grad = [&](float xraw, float t) -> float {
    if (!valid) return 0.f;
    float xs = xraw * inv_temp;
    PixelTerms pt = pixel_terms(xs, t);
    if (MODE == 0) {
      float q = (t != 0.f) ? (1.0f - pt.p) : pt.p;
      float pq = pt.p * (1.0f - pt.p);
      float sgn = (t != 0.f) ? -1.0f : 1.0f;  // d q / dx = (1 - 2t) p (1-p)
      float qg, qg1;                          // q^gamma, gamma * q^(gamma-1)
      if (gamma == 2.0f) { qg = q * q; qg1 = 2.0f * q; }
      else if (gamma == 0.f) { qg = 1.0f; qg1 = 0.f; }
      else { qg1 = gamma * __powf(q, gamma - 1.0f); qg = __powf(q, gamma); }
      float df = (pt.p - t) * qg + pt.ce * qg1 * sgn * pq;
      if (alpha >= 0.f) df *= (t != 0.f) ? alpha : (1.0f - alpha);
      return k_focal * df + (dice_b - dice_a * t) * pq;
    } else {
      float w = fmaf(pw - 1.0f, t, 1.0f);
      return k_bce * ((1.0f - t) - w * (1.0f - pt.p));
    }
  };

  const long long begin = (long long)blockIdx.x * kChunk;
  const long long end = (begin + kChunk < HW) ? (begin + kChunk) : HW;
  if (vec_ok && end - begin == kChunk) {
    if (!valid) {  // masked-out channel: write zeros without reading
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int it = 0; it < kItersPerBlock; ++it) {
        long long base = begin + ((long long)it * kThreads + threadIdx.x) * kPxPerThreadIter;
#pragma unroll
        for (int j = 0; j < 4; ++j) stg_f4(dx + base + 4 * j, z);
      }
      return;
    }
    float4 xv[kItersPerBlock][4];
    uint4 tv[kItersPerBlock];
#pragma unroll
    for (int it = 0; it < kItersPerBlock; ++it) {
      long long base = begin + ((long long)it * kThreads + threadIdx.x) * kPxPerThreadIter;
      tv[it] = ldg_u4(tg + base);
#pragma unroll
      for (int j = 0; j < 4; ++j) xv[it][j] = ldg_f4(x + base + 4 * j);
    }
#pragma unroll
    for (int it = 0; it < kItersPerBlock; ++it) {
      long long base = begin + ((long long)it * kThreads + threadIdx.x) * kPxPerThreadIter;
      const uint32_t tw[4] = {tv[it].x, tv[it].y, tv[it].z, tv[it].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float4 o;
        o.x = grad(xv[it][j].x, ((tw[j] >> 0) & 0xffu) ? 1.f : 0.f);
        o.y = grad(xv[it][j].y, ((tw[j] >> 8) & 0xffu) ? 1.f : 0.f);
        o.z = grad(xv[it][j].z, ((tw[j] >> 16) & 0xffu) ? 1.f : 0.f);
        o.w = grad(xv[it][j].w, ((tw[j] >> 24) & 0xffu) ? 1.f : 0.f);
        stg_f4(dx + base + 4 * j, o);
      }
    }
  } else {
    for (long long i = begin + threadIdx.x; i < end; i += kThreads)
      dx[i] = grad(x[i], tg[i] ? 1.0f : 0.0f);
  }
}

int blocks_per_channel(long long HW) { return (int)((HW + kChunk - 1) / kChunk); }

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace

extern "C" {

// Workspace (floats) the forward needs for its per-block partial sums.
size_t sam2b200_mask_loss_workspace_bytes(int T, int C, long long HW) {
  int tt = T < kMaxFrames ? T : kMaxFrames;
  return (size_t)tt * C * blocks_per_channel(HW) * kNumSums * sizeof(float);
}

int sam2b200_mask_loss_fwd(const float* const* logits, const uint8_t* targets, const float* iou_pred,
                           const float* pos_weight, void* workspace, float* chan_sums, int* n_valid,
                           float* losses, int T, int C, long long HW, int mode, float alpha,
                           float gamma, float inv_temp, int iou_l1, int reduction_mean,
                           cudaStream_t stream) {
  if (T <= 0 || C <= 0 || HW <= 0 || !logits || !targets || !workspace || !chan_sums || !n_valid ||
      !losses || (mode == 0 && !iou_pred) || (mode != 0 && mode != 1))
    return sam2b200::fail(SAM2B200_ERR_INVALID, "mask_loss_fwd: bad arguments");
  const int nblk = blocks_per_channel(HW);
  for (int f0 = 0; f0 < T; f0 += kMaxFrames) {
    const int tt = (T - f0 < kMaxFrames) ? (T - f0) : kMaxFrames;
    FramePtrs fp;
    int vec_ok = (HW % 16 == 0) && aligned16(targets);
    for (int f = 0; f < tt; ++f) {
      fp.logits[f] = logits[f0 + f];
      if (!fp.logits[f]) return sam2b200::fail(SAM2B200_ERR_INVALID, "mask_loss_fwd: null frame");
      vec_ok = vec_ok && aligned16(fp.logits[f]);
    }
    dim3 grid(nblk, tt * C);
    float* part = static_cast<float*>(workspace);
    if (mode == 0)
      mask_loss_fwd_kernel<0><<<grid, kThreads, 0, stream>>>(fp, targets, nullptr, part, C, HW, f0,
                                                             inv_temp, alpha, gamma, vec_ok);
    else
      mask_loss_fwd_kernel<1><<<grid, kThreads, 0, stream>>>(fp, targets, pos_weight, part, C, HW,
                                                             f0, inv_temp, alpha, gamma, vec_ok);
    const size_t sh = (size_t)tt * C * kNumSums * sizeof(double);
    if (sh > 48 * 1024) return sam2b200::fail(SAM2B200_ERR_INVALID, "mask_loss_fwd: T*C too large");
    if (mode == 0)
      mask_loss_finalize_kernel<<<1, 256, sh, stream>>>(
          part, nblk, tt, C, HW, iou_pred + (size_t)f0 * C, iou_l1,
          chan_sums + (size_t)f0 * C * kNumSums, n_valid + f0, losses, f0 > 0);
    else
      bce_loss_finalize_kernel<<<1, 256, sh, stream>>>(part, nblk, tt, C, HW, reduction_mean,
                                                       chan_sums + (size_t)f0 * C * kNumSums,
                                                       n_valid + f0, losses, f0 > 0);
  }
  return sam2b200::check_launch("mask_loss_fwd", 2 * ((T + kMaxFrames - 1) / kMaxFrames));
}

int sam2b200_mask_loss_bwd(const float* const* logits, float* const* dlogits, const uint8_t* targets,
                           const float* iou_pred, const float* pos_weight, const float* chan_sums,
                           const int* n_valid, const float* grad_losses, float* diou, int T, int C,
                           long long HW, int mode, float alpha, float gamma, float inv_temp,
                           int iou_l1, int reduction_mean, cudaStream_t stream) {
  if (T <= 0 || C <= 0 || HW <= 0 || !logits || !dlogits || !targets || !chan_sums || !n_valid ||
      !grad_losses || (mode == 0 && (!iou_pred || !diou)) || (mode != 0 && mode != 1))
    return sam2b200::fail(SAM2B200_ERR_INVALID, "mask_loss_bwd: bad arguments");
  const int nblk = blocks_per_channel(HW);
  for (int f0 = 0; f0 < T; f0 += kMaxFrames) {
    const int tt = (T - f0 < kMaxFrames) ? (T - f0) : kMaxFrames;
    FramePtrs fp;
    FrameOutPtrs op;
    int vec_ok = (HW % 16 == 0) && aligned16(targets);
    for (int f = 0; f < tt; ++f) {
      fp.logits[f] = logits[f0 + f];
      op.dlogits[f] = dlogits[f0 + f];
      if (!fp.logits[f] || !op.dlogits[f])
        return sam2b200::fail(SAM2B200_ERR_INVALID, "mask_loss_bwd: null frame");
      vec_ok = vec_ok && aligned16(fp.logits[f]) && aligned16(op.dlogits[f]);
    }
    dim3 grid(nblk, tt * C);
    if (mode == 0)
      mask_loss_bwd_kernel<0><<<grid, kThreads, 0, stream>>>(
          fp, op, targets, nullptr, chan_sums + (size_t)f0 * C * kNumSums, n_valid + f0, grad_losses,
          iou_pred + (size_t)f0 * C, diou + (size_t)f0 * C, C, HW, f0, inv_temp, alpha, gamma,
          iou_l1, reduction_mean, vec_ok);
    else
      mask_loss_bwd_kernel<1><<<grid, kThreads, 0, stream>>>(
          fp, op, targets, pos_weight, chan_sums + (size_t)f0 * C * kNumSums, n_valid + f0,
          grad_losses, nullptr, nullptr, C, HW, f0, inv_temp, alpha, gamma, iou_l1, reduction_mean,
          vec_ok);
  }
  return sam2b200::check_launch("mask_loss_bwd", (T + kMaxFrames - 1) / kMaxFrames);
}

}  // extern "C"
