// Producer side of the mask loss fused into it for sm_100a.
//
// Replaces, per frame, the chain the reference runs between the mask decoder and the loss value:
//   sam2_video/model/modeling/sam2_base.py:393-399   F.interpolate(low_res.float(), (S, S), "bilinear", align_corners=False)
//   sam2_video/utils/masks.py:102-116                _grouped_max        (pixel-wise max over the objects of a category)
//   sam2_video/utils/masks.py:92-100,118-145         area weights sum(sigmoid(high-res logits)), weighted IoU average
//   sam2_video/model/losses.py:20-76,143-238         focal + dice + IoU with the valid-channel filter
// The high-resolution logits ([n_obj, 1, S, S] fp32 per frame, then [C, 1, S, S] after the merge) are never
// materialised: the forward reads the low-res logits (4 B per 16 high-res px per object) and the targets (1 B/px),
// up-samples in registers from a shared-memory patch, takes the max over the category's objects and feeds the same
// per-pixel accumulator as mask_loss.cu; the backward recomputes the up-sampling, routes d loss / d merged-logit to
// the arg-max object (first maximal index, like torch.max), adds the gradient that reaches the logits through the
// (not detached) area weights of the IoU average, and applies the ADJOINT of the bilinear filter as a gather: each
// low-res pixel sums its 8 x 8 high-res neighbourhood from a shared-memory tile, so every output element is written
// by exactly one thread (no atomics, no memset, deterministic).
//
// 4x bilinear, align_corners=False: high-res pixel p reads low-res taps i0 = ((p + 2) >> 2) - 1 and i0 + 1 (both clamped
// to [0, s-1]) with weight lambda = ((p + 2) & 3) / 4 + 1/8 on the second.  An aligned group of 4 pixels 4j..4j+3 therefore
// needs the low-res columns j-1, j, j+1 with weights (.375,.625) (.125,.875) (.875,.125) (.625,.375).
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "abi_common.cuh"
#include "loss_math.cuh"

namespace {
using namespace lossmath;

constexpr int kThreads = 256;
constexpr int kMaxFrames = 64;
constexpr int kRec = 8;
constexpr int kNumSums = 6;
constexpr int kLowRows = 8, kLowCols = 32;           // low-res tile of a block = 32 x 128 high-res px
constexpr int kPatchRows = kLowRows + 2;             // low rows Y0-1 .. Y0+8
constexpr int kPatchCols = kLowCols + 4;             // low cols X0-2 .. X0+33
constexpr int kHiRows = 4 * kLowRows + 4;            // backward: high rows 4 Y0 - 2 .. 4 Y0 + 33
constexpr int kGroups = kLowCols + 2;                // backward: aligned 4-px groups X0-1 .. X0+32
constexpr int kItems = kHiRows * kGroups;            // 1224 (row, group) items of 4 px
constexpr int kItemsPerThread = (kItems + kThreads - 1) / kThreads;   // 5

struct LowPtrs { const float* p[kMaxFrames]; };
struct LowOutPtrs { float* p[kMaxFrames]; };

// patch[r][c] = L[clamp(Y0 - 1 + r)][clamp(X0 - 2 + c)]: the clamps implement the border rule of the filter
__device__ __forceinline__ void load_patch(float (*patch)[kPatchCols], const float* __restrict__ L, int s, int Y0, int X0) {
  for (int e = threadIdx.x; e < kPatchRows * kPatchCols; e += kThreads) {
    const int r = e / kPatchCols, c = e % kPatchCols;
    const int y = min(max(Y0 - 1 + r, 0), s - 1), x = min(max(X0 - 2 + c, 0), s - 1);
    patch[r][c] = __ldg(L + (long long)y * s + x);
  }
}

// the 4 up-sampled values of high-res row `py`, aligned group `j` (global indices; tile origin Y0, X0)
__device__ __forceinline__ void upsample4(const float (*patch)[kPatchCols], int py, int j, int Y0, int X0, float u[4]) {
  const int i0 = ((py + 2) >> 2) - 1;
  const float ly = (float)((py + 2) & 3) * 0.25f + 0.125f;
  const int r0 = i0 - (Y0 - 1), cj = j - (X0 - 2);
  const float a0 = patch[r0][cj - 1], b0 = patch[r0][cj], c0 = patch[r0][cj + 1];
  const float a1 = patch[r0 + 1][cj - 1], b1 = patch[r0 + 1][cj], c1 = patch[r0 + 1][cj + 1];
  const float h00 = fmaf(0.375f, a0, 0.625f * b0), h01 = fmaf(0.125f, a0, 0.875f * b0);
  const float h02 = fmaf(0.125f, c0, 0.875f * b0), h03 = fmaf(0.375f, c0, 0.625f * b0);
  const float h10 = fmaf(0.375f, a1, 0.625f * b1), h11 = fmaf(0.125f, a1, 0.875f * b1);
  const float h12 = fmaf(0.125f, c1, 0.875f * b1), h13 = fmaf(0.375f, c1, 0.625f * b1);
  const float my = 1.0f - ly;
  u[0] = fmaf(ly, h10, my * h00);
  u[1] = fmaf(ly, h11, my * h01);
  u[2] = fmaf(ly, h12, my * h02);
  u[3] = fmaf(ly, h13, my * h03);
}

__device__ __forceinline__ float sigmoid_fast(float x) {
  return rcp_approx(1.0f + ex2_approx(fmaxf(x, -80.0f) * -kLog2e));
}

struct FwdParams {
  const uint8_t* targets;      // [tt, C, S, S]
  const int* group_offsets;    // [C + 1]
  const int* group_members;    // [n_obj]
  float* loss_rec;             // [tt*C][ntiles][kRec]
  float* area_rec;             // [tt*n_obj][ntiles]
  int C, n_obj, s, tiles_x;
  float inv_temp, gamma;
};

// grid: (tiles, C, tt)
template <bool G2, bool UNIT_T>
__global__ void __launch_bounds__(kThreads)
merged_loss_fwd_kernel(const __grid_constant__ LowPtrs lp, const __grid_constant__ FwdParams P) {
  __shared__ float patch[kPatchRows][kPatchCols];
  __shared__ float warp_area[kThreads / 32];
  __shared__ float red[kRec][kThreads];
  const int tile = blockIdx.x, c = blockIdx.y, f = blockIdx.z;
  const int s = P.s, S = 4 * s;
  const int Y0 = (tile / P.tiles_x) * kLowRows, X0 = (tile % P.tiles_x) * kLowCols;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j = X0 + lane;
  const bool col_ok = j < s;
  const int m0 = P.group_offsets[c], m1 = P.group_offsets[c + 1];
  const int ntiles = gridDim.x;

  float x[4][4];
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int e = 0; e < 4; ++e) x[k][e] = (m1 > m0) ? -CUDART_INF_F : 0.0f;   // empty category: zeros (masks.py:111-112)

  for (int m = m0; m < m1; ++m) {
    const int obj = P.group_members[m];
    load_patch(patch, lp.p[f] + (long long)obj * s * s, s, Y0, X0);
    __syncthreads();
    float area = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int py = 4 * Y0 + warp + 8 * k;
      if (col_ok && py < S) {
        float u[4];
        upsample4(patch, py, j, Y0, X0, u);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          x[k][e] = fmaxf(x[k][e], u[e]);
          area += sigmoid_fast(u[e]);
        }
      }
    }
    area = warp_sum(area);
    if (lane == 0) warp_area[warp] = area;
    __syncthreads();
    if (threadIdx.x == 0) {
      float a = 0.f;
#pragma unroll
      for (int w = 0; w < kThreads / 32; ++w) a += warp_area[w];
      P.area_rec[((long long)f * P.n_obj + obj) * ntiles + tile] = a;
    }
  }

  Acc<0, G2, UNIT_T> acc;
  acc.init();
  const uint8_t* __restrict__ tg = P.targets + ((long long)f * P.C + c) * S * S;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int py = 4 * Y0 + warp + 8 * k;
    if (col_ok && py < S) {
      const uint32_t tw = ldg_u32(tg + (long long)py * S + 4 * j);
      acc.add(x[k][0], (tw & 0x000000ffu) != 0u, P.inv_temp, P.gamma);
      acc.add(x[k][1], (tw & 0x0000ff00u) != 0u, P.inv_temp, P.gamma);
      acc.add(x[k][2], (tw & 0x00ff0000u) != 0u, P.inv_temp, P.gamma);
      acc.add(x[k][3], (tw & 0xff000000u) != 0u, P.inv_temp, P.gamma);
    }
  }
  red[0][threadIdx.x] = acc.f0;
  red[1][threadIdx.x] = acc.f1;
  red[2][threadIdx.x] = acc.a;
  red[3][threadIdx.x] = acc.bq;
  red[4][threadIdx.x] = (float)(acc.cnt & 0xffu);
  red[5][threadIdx.x] = (float)((acc.cnt >> 8) & 0xffu);
  red[6][threadIdx.x] = (float)(acc.cnt >> 16);
  __syncthreads();
  if (warp < 7) {
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < kThreads / 32; ++i) v += red[warp][i * 32 + lane];
    v = warp_sum(v);
    if (lane == 0) P.loss_rec[(((long long)f * P.C + c) * ntiles + tile) * kRec + warp] = v;
  }
}

struct FoldParams {
  const float* loss_rec;
  const float* area_rec;
  const int* group_offsets;
  const int* group_members;
  const float* obj_iou;        // [tt, n_obj]
  float* chan_sums;            // [tt*C][6]
  float* obj_area;             // [tt, n_obj]
  float* cat_iou;              // [tt*C]
  float* cat_w;                // [tt*C]
  int C, n_obj, ntiles;
  float alpha;
};

// grid: tt*C.  Folds the per-tile records of one (frame, category) in fp64, fixed order; merges the IoU predictions.
__global__ void __launch_bounds__(kThreads) merged_loss_fold_kernel(const __grid_constant__ FoldParams P) {
  __shared__ double tot[kRec];
  const int fc = blockIdx.x, f = fc / P.C, c = fc % P.C;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp < 7) {
    const float* rec = P.loss_rec + (long long)fc * P.ntiles * kRec;
    double sacc = 0.0;
    for (int b = lane; b < P.ntiles; b += 32) sacc += (double)rec[(long long)b * kRec + warp];
    sacc = warp_sum_d(sacc);
    if (lane == 0) tot[warp] = sacc;
  }
  const int m0 = P.group_offsets[c], m1 = P.group_offsets[c + 1];
  for (int m = m0 + warp; m < m1; m += kThreads / 32) {
    const int obj = P.group_members[m];
    const float* rec = P.area_rec + ((long long)f * P.n_obj + obj) * P.ntiles;
    double sacc = 0.0;
    for (int b = lane; b < P.ntiles; b += 32) sacc += (double)rec[b];
    sacc = warp_sum_d(sacc);
    if (lane == 0) P.obj_area[(long long)f * P.n_obj + obj] = (float)sacc;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float* cs = P.chan_sums + (long long)fc * kNumSums;
    const double T = tot[4];
    const double F = (P.alpha >= 0.f) ? (double)P.alpha * tot[1] + (1.0 - (double)P.alpha) * tot[0] : tot[0] + tot[1];
    cs[0] = (float)F;
    cs[1] = (float)(T - tot[2]);
    cs[2] = (float)(tot[3] + T - 2.0 * tot[2]);
    cs[3] = (float)T;
    cs[4] = (float)tot[6];
    cs[5] = (float)(tot[5] + T - tot[6]);
    // masks.py:118-145: weighted average with the (fp32) area weights; all-zero weights -> plain mean; no objects -> 0
    float w = 0.f, num = 0.f, mean = 0.f;
    for (int m = m0; m < m1; ++m) {
      const int obj = P.group_members[m];
      const float a = P.obj_area[(long long)f * P.n_obj + obj], q = P.obj_iou[(long long)f * P.n_obj + obj];
      w += a; num = fmaf(a, q, num); mean += q;
    }
    P.cat_w[fc] = w;
    P.cat_iou[fc] = (m1 == m0) ? 0.f : (w == 0.f ? mean / (float)(m1 - m0) : num / w);
  }
}

struct FinalParams {
  const float* chan_sums;
  const float* cat_iou;
  int* n_valid;
  float* losses;
  int tt, C, iou_l1;
  double HW;
};

// one block: valid-channel filter, Nv per frame, loss algebra, sums over channels and frames (losses.py:143-238)
__global__ void __launch_bounds__(kThreads) merged_loss_final_kernel(const __grid_constant__ FinalParams P) {
  __shared__ int nv_sh[kMaxFrames];
  __shared__ double part[3][kThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = P.tt * P.C;
  for (int i = threadIdx.x; i < P.tt; i += kThreads) nv_sh[i] = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += kThreads)
    if (P.chan_sums[(long long)i * kNumSums + 3] > 0.f) atomicAdd(&nv_sh[i / P.C], 1);
  __syncthreads();
  double lm = 0.0, ld = 0.0, li = 0.0;
  for (int i = threadIdx.x; i < n; i += kThreads) {
    const float* sm = P.chan_sums + (long long)i * kNumSums;
    const double s0 = sm[0], s1 = sm[1], s2 = sm[2], s3 = sm[3], s4 = sm[4], s5 = sm[5];
    const int nv = nv_sh[i / P.C];
    if (!(s3 > 0.0) || nv == 0) continue;
    lm += s0 / P.HW / nv;
    ld += (1.0 - (2.0 * s1 + 1.0) / (s2 + s3 + 1.0)) / nv;
    const double dd = (double)P.cat_iou[i] - s4 / fmax(s5, 1.0);
    li += (P.iou_l1 ? fabs(dd) : dd * dd) / nv;
  }
  lm = warp_sum_d(lm); ld = warp_sum_d(ld); li = warp_sum_d(li);
  if (lane == 0) { part[0][warp] = lm; part[1][warp] = ld; part[2][warp] = li; }
  __syncthreads();
  for (int i = threadIdx.x; i < P.tt; i += kThreads) P.n_valid[i] = nv_sh[i];
  if (threadIdx.x == 0) {
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    for (int w = 0; w < kThreads / 32; ++w) { a0 += part[0][w]; a1 += part[1][w]; a2 += part[2][w]; }
    P.losses[0] = (float)a0; P.losses[1] = (float)a1; P.losses[2] = (float)a2; P.losses[3] = 0.f;
  }
}

struct BwdParams {
  const uint8_t* targets;
  const int* group_offsets;
  const int* group_members;
  const float* obj_iou;        // [tt, n_obj]
  const float* chan_sums;      // [tt*C][6]
  const float* obj_area;       // [tt, n_obj]
  const float* cat_iou;        // [tt*C]
  const float* cat_w;          // [tt*C]
  const int* n_valid;          // [tt]
  const float* gout;           // [3]
  float* d_obj_iou;            // [tt, n_obj]
  int C, n_obj, s, tiles_x;
  float inv_temp, alpha, gamma;
  int iou_l1;
};

// grid: (tiles, C, tt)
template <bool G2, bool UNIT_T>
__global__ void __launch_bounds__(kThreads)
merged_loss_bwd_kernel(const __grid_constant__ LowPtrs lp, const __grid_constant__ LowOutPtrs op,
                       const __grid_constant__ BwdParams P) {
  __shared__ float patch[kPatchRows][kPatchCols];
  __shared__ float plane[4][kHiRows][kGroups];       // d loss / d up-sampled logit of the current object, by pixel phase
  const int tile = blockIdx.x, c = blockIdx.y, f = blockIdx.z;
  const int fc = f * P.C + c;
  const int s = P.s, S = 4 * s;
  const int Y0 = (tile / P.tiles_x) * kLowRows, X0 = (tile % P.tiles_x) * kLowCols;
  const int m0 = P.group_offsets[c], m1 = P.group_offsets[c + 1];
  if (m1 == m0) return;                               // no object feeds this category: nothing to write
  const int Yl = threadIdx.x >> 5, Xl = threadIdx.x & 31;
  const int Yo = Y0 + Yl, Xo = X0 + Xl;
  const bool out_ok = Yo < s && Xo < s;

  const float* sums = P.chan_sums + (long long)fc * kNumSums;
  const int nv = P.n_valid[f];
  const bool valid = sums[3] > 0.f && nv > 0;

  // d loss / d (merged IoU prediction of this category)
  float g_iou = 0.f;
  if (valid) {
    const float d = P.cat_iou[fc] - sums[4] / fmaxf(sums[5], 1.0f);
    g_iou = (P.iou_l1 ? (float)((d > 0.f) - (d < 0.f)) : 2.0f * d) * P.gout[2] / (float)nv;
  }
  const float W = P.cat_w[fc];
  if (tile == 0) {                                    // d / d per-object IoU predictions (masks.py:143 or the mean fallback :141)
    for (int m = m0 + threadIdx.x; m < m1; m += kThreads) {
      const int obj = P.group_members[m];
      P.d_obj_iou[(long long)f * P.n_obj + obj] =
          (W == 0.f) ? g_iou / (float)(m1 - m0) : g_iou * P.obj_area[(long long)f * P.n_obj + obj] / W;
    }
  }
  if (!valid) {                                       // filtered-out channel (losses.py:153-159): zero gradient
    if (out_ok)
      for (int m = m0; m < m1; ++m) op.p[f][((long long)P.group_members[m] * s + Yo) * s + Xo] = 0.f;
    return;
  }

  // per-channel coefficients of d loss / d merged logit (same algebra as mask_loss_bwd_kernel)
  const float inv_nv_t = P.inv_temp / (float)nv;
  const float kf = P.gout[0] * inv_nv_t / ((float)S * (float)S);
  const float kf_fg = -kf * ((P.alpha >= 0.f) ? P.alpha : 1.0f);
  const float kf_bg = kf * ((P.alpha >= 0.f) ? (1.0f - P.alpha) : 1.0f);
  const float dp1 = sums[2] + sums[3] + 1.0f, nn1 = 2.0f * sums[1] + 1.0f;
  const float kd = P.gout[1] * inv_nv_t;
  const float dc_bg = kd * nn1 / (dp1 * dp1);
  const float dc_fg = dc_bg - kd * 2.0f / dp1;
  auto grad = [&](float xraw, bool t) -> float {
    float d, sp;
    softplus_terms<UNIT_T>(t ? -xraw : xraw, P.inv_temp, d, sp);
    const float q = rcp_approx(d);
    const float omq = 1.0f - q;
    float u, qg;
    if (G2) { u = fmaf(sp + sp, omq, q); qg = q * q; }
    else { u = fmaf(P.gamma * sp, omq, q); qg = (P.gamma == 0.f) ? 1.0f : __powf(q, P.gamma); }
    return fmaf(t ? dc_fg : dc_bg, q * omq, (u * qg) * (t ? kf_fg : kf_bg));
  };

  // this thread's (high-res row, 4-px group) items
  int ipy[kItemsPerThread], ij[kItemsPerThread];
  bool iin[kItemsPerThread];
#pragma unroll
  for (int k = 0; k < kItemsPerThread; ++k) {
    const int it = threadIdx.x + k * kThreads;
    const int r = it / kGroups, g = it % kGroups;
    ipy[k] = 4 * Y0 - 2 + r;
    ij[k] = X0 - 1 + g;
    iin[k] = it < kItems && ipy[k] >= 0 && ipy[k] < S && ij[k] >= 0 && ij[k] < s;
  }

  // pass 1: merged logit and arg-max object (first maximal index) per pixel
  float x[kItemsPerThread][4];
  uint32_t arg[kItemsPerThread];
#pragma unroll
  for (int k = 0; k < kItemsPerThread; ++k) {
    arg[k] = 0u;
#pragma unroll
    for (int e = 0; e < 4; ++e) x[k][e] = -CUDART_INF_F;
  }
  for (int m = m0; m < m1; ++m) {
    load_patch(patch, lp.p[f] + (long long)P.group_members[m] * s * s, s, Y0, X0);
    __syncthreads();
    const uint32_t ml = (uint32_t)(m - m0);
#pragma unroll
    for (int k = 0; k < kItemsPerThread; ++k) {
      if (!iin[k]) continue;
      float u[4];
      upsample4(patch, ipy[k], ij[k], Y0, X0, u);
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (u[e] > x[k][e]) { x[k][e] = u[e]; arg[k] = (arg[k] & ~(0xffu << (8 * e))) | (ml << (8 * e)); }
    }
    __syncthreads();
  }
  // d loss / d merged logit (overwrites x)
  const uint8_t* __restrict__ tg = P.targets + (long long)fc * S * S;
#pragma unroll
  for (int k = 0; k < kItemsPerThread; ++k) {
    if (!iin[k]) continue;
    const uint32_t tw = ldg_u32(tg + (long long)ipy[k] * S + 4 * ij[k]);
    x[k][0] = grad(x[k][0], (tw & 0x000000ffu) != 0u);
    x[k][1] = grad(x[k][1], (tw & 0x0000ff00u) != 0u);
    x[k][2] = grad(x[k][2], (tw & 0x00ff0000u) != 0u);
    x[k][3] = grad(x[k][3], (tw & 0xff000000u) != 0u);
  }

  // adjoint filter weights of this thread's low-res output pixel (border taps fold onto the clamped index)
  float wy[8], wx[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float base = (k < 4) ? (0.125f + 0.25f * (float)k) : (0.875f - 0.25f * (float)(k - 4));
    wy[k] = base; wx[k] = base;
  }
  if (Yo == 0) { wy[2] = 1.f; wy[3] = 1.f; }
  if (Yo == s - 1) { wy[4] = 1.f; wy[5] = 1.f; }
  if (Xo == 0) { wx[2] = 1.f; wx[3] = 1.f; }
  if (Xo == s - 1) { wx[4] = 1.f; wx[5] = 1.f; }

  // pass 2: per object, d loss / d up-sampled logit -> shared planes -> gather through the adjoint filter
  for (int m = m0; m < m1; ++m) {
    const int obj = P.group_members[m];
    load_patch(patch, lp.p[f] + (long long)obj * s * s, s, Y0, X0);
    // gradient through the area weight of the IoU average: d iou_c / d w_obj = (iou_obj - iou_c) / W
    const float gi = (W == 0.f) ? 0.f : g_iou * (P.obj_iou[(long long)f * P.n_obj + obj] - P.cat_iou[fc]) / W;
    __syncthreads();
    const uint32_t ml = (uint32_t)(m - m0);
#pragma unroll
    for (int k = 0; k < kItemsPerThread; ++k) {
      const int it = threadIdx.x + k * kThreads;
      if (it >= kItems) continue;
      const int r = it / kGroups, g = it % kGroups;
      float du[4] = {0.f, 0.f, 0.f, 0.f};
      if (iin[k]) {
        float u[4];
        upsample4(patch, ipy[k], ij[k], Y0, X0, u);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float p = sigmoid_fast(u[e]);
          du[e] = fmaf(gi, p * (1.0f - p), (((arg[k] >> (8 * e)) & 0xffu) == ml) ? x[k][e] : 0.f);
        }
      }
      plane[0][r][g] = du[0]; plane[1][r][g] = du[1]; plane[2][r][g] = du[2]; plane[3][r][g] = du[3];
    }
    __syncthreads();
    if (out_ok) {
      float acc = 0.f;
#pragma unroll
      for (int ky = 0; ky < 8; ++ky) {
        const int r = 4 * Yl + ky;
        float row = wx[0] * plane[2][r][Xl];
        row = fmaf(wx[1], plane[3][r][Xl], row);
        row = fmaf(wx[2], plane[0][r][Xl + 1], row);
        row = fmaf(wx[3], plane[1][r][Xl + 1], row);
        row = fmaf(wx[4], plane[2][r][Xl + 1], row);
        row = fmaf(wx[5], plane[3][r][Xl + 1], row);
        row = fmaf(wx[6], plane[0][r][Xl + 2], row);
        row = fmaf(wx[7], plane[1][r][Xl + 2], row);
        acc = fmaf(wy[ky], row, acc);
      }
      op.p[f][((long long)obj * s + Yo) * s + Xo] = acc;
    }
  }
}

bool aligned4(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 3u) == 0; }
size_t round256(size_t b) { return (b + 255) & ~size_t(255); }
int tiles_x_of(int s) { return (s + kLowCols - 1) / kLowCols; }
int tiles_of(int s) { return tiles_x_of(s) * ((s + kLowRows - 1) / kLowRows); }

}  // namespace

extern "C" {

size_t sam2b200_merged_loss_workspace_bytes(int T, int C, int n_obj, int s) {
  if (T <= 0 || C <= 0 || n_obj <= 0 || s <= 0) return 0;
  const size_t nt = (size_t)tiles_of(s);
  return round256((size_t)T * C * nt * kRec * sizeof(float)) + round256((size_t)T * n_obj * nt * sizeof(float));
}

int sam2b200_merged_loss_fwd(const float* const* low_res, const uint8_t* targets, const float* obj_iou,
                             const int* group_offsets, const int* group_members, void* workspace, float* chan_sums,
                             float* obj_area, float* cat_iou, float* cat_w, int* n_valid, float* losses, int T, int C,
                             int n_obj, int s, float alpha, float gamma, float inv_temp, int iou_l1, cudaStream_t stream) {
  if (T <= 0 || C <= 0 || n_obj <= 0 || s <= 0 || !low_res || !targets || !obj_iou || !group_offsets || !group_members ||
      !workspace || !chan_sums || !obj_area || !cat_iou || !cat_w || !n_valid || !losses)
    return sam2b200::fail(SAM2B200_ERR_INVALID, "merged_loss_fwd: bad arguments");
  if (T > kMaxFrames || C > 65535) return sam2b200::fail(SAM2B200_ERR_UNSUPPORTED, "merged_loss_fwd: T > 64 frames or C > 65535 per call");
  if (!aligned4(targets)) return sam2b200::fail(SAM2B200_ERR_INVALID, "merged_loss_fwd: targets must be 4-byte aligned");
  LowPtrs lp;
  for (int f = 0; f < T; ++f) {
    lp.p[f] = low_res[f];
    if (!lp.p[f]) return sam2b200::fail(SAM2B200_ERR_INVALID, "merged_loss_fwd: null frame");
  }
  const int nt = tiles_of(s);
  FwdParams P;
  P.targets = targets; P.group_offsets = group_offsets; P.group_members = group_members;
  P.loss_rec = static_cast<float*>(workspace);
  P.area_rec = reinterpret_cast<float*>(static_cast<char*>(workspace) + round256((size_t)T * C * nt * kRec * sizeof(float)));
  P.C = C; P.n_obj = n_obj; P.s = s; P.tiles_x = tiles_x_of(s); P.inv_temp = inv_temp; P.gamma = gamma;
  dim3 grid(nt, C, T);
  if (gamma == 2.0f && inv_temp == 1.0f) merged_loss_fwd_kernel<true, true><<<grid, kThreads, 0, stream>>>(lp, P);
  else merged_loss_fwd_kernel<false, false><<<grid, kThreads, 0, stream>>>(lp, P);
  FoldParams Q;
  Q.loss_rec = P.loss_rec; Q.area_rec = P.area_rec; Q.group_offsets = group_offsets; Q.group_members = group_members;
  Q.obj_iou = obj_iou; Q.chan_sums = chan_sums; Q.obj_area = obj_area; Q.cat_iou = cat_iou; Q.cat_w = cat_w;
  Q.C = C; Q.n_obj = n_obj; Q.ntiles = nt; Q.alpha = alpha;
  merged_loss_fold_kernel<<<T * C, kThreads, 0, stream>>>(Q);
  FinalParams R;
  R.chan_sums = chan_sums; R.cat_iou = cat_iou; R.n_valid = n_valid; R.losses = losses; R.tt = T; R.C = C; R.iou_l1 = iou_l1;
  R.HW = 16.0 * (double)s * (double)s;
  merged_loss_final_kernel<<<1, kThreads, 0, stream>>>(R);
  return sam2b200::check_launch("merged_loss_fwd", 3);
}

int sam2b200_merged_loss_bwd(const float* const* low_res, float* const* dlow_res, const uint8_t* targets,
                             const float* obj_iou, const int* group_offsets, const int* group_members,
                             const float* chan_sums, const float* obj_area, const float* cat_iou, const float* cat_w,
                             const int* n_valid, const float* grad_losses, float* d_obj_iou, int T, int C, int n_obj, int s,
                             float alpha, float gamma, float inv_temp, int iou_l1, cudaStream_t stream) {
  if (T <= 0 || C <= 0 || n_obj <= 0 || s <= 0 || !low_res || !dlow_res || !targets || !obj_iou || !group_offsets ||
      !group_members || !chan_sums || !obj_area || !cat_iou || !cat_w || !n_valid || !grad_losses || !d_obj_iou)
    return sam2b200::fail(SAM2B200_ERR_INVALID, "merged_loss_bwd: bad arguments");
  if (T > kMaxFrames || C > 65535) return sam2b200::fail(SAM2B200_ERR_UNSUPPORTED, "merged_loss_bwd: T > 64 frames or C > 65535 per call");
  if (!aligned4(targets)) return sam2b200::fail(SAM2B200_ERR_INVALID, "merged_loss_bwd: targets must be 4-byte aligned");
  LowPtrs lp;
  LowOutPtrs op;
  for (int f = 0; f < T; ++f) {
    lp.p[f] = low_res[f];
    op.p[f] = dlow_res[f];
    if (!lp.p[f] || !op.p[f]) return sam2b200::fail(SAM2B200_ERR_INVALID, "merged_loss_bwd: null frame");
  }
  BwdParams P;
  P.targets = targets; P.group_offsets = group_offsets; P.group_members = group_members; P.obj_iou = obj_iou;
  P.chan_sums = chan_sums; P.obj_area = obj_area; P.cat_iou = cat_iou; P.cat_w = cat_w; P.n_valid = n_valid;
  P.gout = grad_losses; P.d_obj_iou = d_obj_iou; P.C = C; P.n_obj = n_obj; P.s = s; P.tiles_x = tiles_x_of(s);
  P.inv_temp = inv_temp; P.alpha = alpha; P.gamma = gamma; P.iou_l1 = iou_l1;
  dim3 grid(tiles_of(s), C, T);
  if (gamma == 2.0f && inv_temp == 1.0f) merged_loss_bwd_kernel<true, true><<<grid, kThreads, 0, stream>>>(lp, op, P);
  else merged_loss_bwd_kernel<false, false><<<grid, kThreads, 0, stream>>>(lp, op, P);
  return sam2b200::check_launch("merged_loss_bwd", 1);
}

}  // extern "C"
