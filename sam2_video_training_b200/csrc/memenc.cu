// Kernels of the SAM2 memory encoder (sam2_video/model/modeling/memory_encoder.py) that are not GEMMs, in the channels-last
// (token-major) layout the rest of the B200 path uses:
//
//   ln_gelu_fwd / ln_gelu_bwd   y = GELU(LayerNorm_C(x) * w + b) per pixel over C = 4 | 16 | 64 | 256 channels: the
//                               LayerNorm2d + GELU pair that follows every strided convolution of MaskDownSampler
//                               (memory_encoder.py:38-53; LayerNorm2d: sam2_utils.py:141-153, eps 1e-6; exact erf GELU)
//   dwconv7_fwd / _bwd_w        depth-wise 7 x 7 convolution, padding 3, of CXBlock (memory_encoder.py:84-91), [B, H, W, C]
//                               fp32; the data gradient is the same kernel with the taps flipped
//
// HBM-bound element-wise / stencil work: 128-bit accesses where the width allows, reductions for the parameter gradients in
// registers -> shared memory -> one fp32 atomic per (block, parameter).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "abi_common.cuh"

namespace {

__device__ __forceinline__ float gelu_f(float z) { return 0.5f * z * (1.0f + erff(z * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad(float z) {
  return 0.5f * (1.0f + erff(z * 0.70710678118654752f)) + z * 0.3989422804014327f * __expf(-0.5f * z * z);
}

// One pixel is handled by G = min(32, C) consecutive lanes, V = C / G channels per lane (C = 4: 8 pixels per warp ... C = 256:
// one pixel per warp, 8 channels per lane).  x, y: [P, C] fp32 (channels-last pixels).
template <int C>
struct LnShape {
  static constexpr int G = (C < 32) ? C : 32;
  static constexpr int V = C / G;
  static constexpr int PPW = 32 / G;      // pixels per warp
};

template <int C>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = LnShape<C>::G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int C, bool ACT>
__global__ void __launch_bounds__(256)
ln_gelu_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b, float* __restrict__ y,
                   long long P, float eps) {
  using S = LnShape<C>;
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long pix = warp * S::PPW + lane / S::G;
  const int c0 = (lane % S::G) * S::V;
  const bool ok = pix < P;
  float v[S::V];
#pragma unroll
  for (int i = 0; i < S::V; ++i) v[i] = ok ? x[pix * C + c0 + i] : 0.f;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < S::V; ++i) s += v[i];
  const float mu = group_sum<C>(s) * (1.0f / C);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < S::V; ++i) { const float d = v[i] - mu; q += d * d; }
  const float rs = rsqrtf(group_sum<C>(q) * (1.0f / C) + eps);
  if (ok) {
#pragma unroll
    for (int i = 0; i < S::V; ++i) {
      const float z = (v[i] - mu) * rs * w[c0 + i] + b[c0 + i];
      y[pix * C + c0 + i] = ACT ? gelu_f(z) : z;
    }
  }
}

// dx, and dw / db accumulated with atomics (one per (block, channel)).  Each warp walks pixels with a grid stride.
template <int C, bool ACT>
__global__ void __launch_bounds__(256)
ln_gelu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                   float* __restrict__ dx, float* __restrict__ dw, float* __restrict__ db, long long P, float eps) {
  using S = LnShape<C>;
  __shared__ float sw[C], sb[C];
  for (int i = threadIdx.x; i < C; i += blockDim.x) { sw[i] = 0.f; sb[i] = 0.f; }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int c0 = (lane % S::G) * S::V;
  float wv[S::V], bv[S::V], aw[S::V], ab[S::V];
#pragma unroll
  for (int i = 0; i < S::V; ++i) { wv[i] = w[c0 + i]; bv[i] = b[c0 + i]; aw[i] = 0.f; ab[i] = 0.f; }
  const long long nwarp = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; warp * S::PPW < P; warp += nwarp) {
    const long long pix = warp * S::PPW + lane / S::G;
    const bool ok = pix < P;
    float v[S::V], g[S::V];
#pragma unroll
    for (int i = 0; i < S::V; ++i) { v[i] = ok ? x[pix * C + c0 + i] : 0.f; g[i] = ok ? dy[pix * C + c0 + i] : 0.f; }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < S::V; ++i) s += v[i];
    const float mu = group_sum<C>(s) * (1.0f / C);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < S::V; ++i) { const float d = v[i] - mu; q += d * d; }
    const float rs = rsqrtf(group_sum<C>(q) * (1.0f / C) + eps);
    float xh[S::V], dxh[S::V];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < S::V; ++i) {
      xh[i] = (v[i] - mu) * rs;
      const float z = xh[i] * wv[i] + bv[i];
      const float dz = ACT ? g[i] * gelu_grad(z) : g[i];
      aw[i] += dz * xh[i];
      ab[i] += dz;
      dxh[i] = dz * wv[i];
      s1 += dxh[i];
      s2 += dxh[i] * xh[i];
    }
    s1 = group_sum<C>(s1) * (1.0f / C);
    s2 = group_sum<C>(s2) * (1.0f / C);
    if (ok) {
#pragma unroll
      for (int i = 0; i < S::V; ++i) dx[pix * C + c0 + i] = rs * (dxh[i] - s1 - xh[i] * s2);
    }
  }
#pragma unroll
  for (int i = 0; i < S::V; ++i) { atomicAdd(&sw[c0 + i], aw[i]); atomicAdd(&sb[c0 + i], ab[i]); }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) { atomicAdd(dw + i, sw[i]); atomicAdd(db + i, sb[i]); }
}

// ---- depth-wise 7 x 7, padding 3, channels-last [B, H, W, C]: one thread per channel (a warp reads 128 contiguous bytes per
// pixel) and per [kTy x kTx] output tile; the (kTy + 6) x (kTx + 6) input window is walked row by row with a register row
// buffer, so every input value is loaded once per tile: 100 loads for 784 FMAs.
constexpr int kTy = 4, kTx = 4;

template <bool FLIP>
__global__ void __launch_bounds__(256)
dwconv7_kernel(const float* __restrict__ x, const float* __restrict__ wgt /*[C, 7, 7]*/, const float* __restrict__ bias, float* __restrict__ y,
               int B, int H, int W, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int tx_n = (W + kTx - 1) / kTx, ty_n = (H + kTy - 1) / kTy;
  const int tx = blockIdx.y % tx_n, ty = (blockIdx.y / tx_n) % ty_n, bb = blockIdx.y / (tx_n * ty_n);
  if (c >= C) return;
  float wk[49];
#pragma unroll
  for (int i = 0; i < 49; ++i) wk[i] = wgt[c * 49 + (FLIP ? 48 - i : i)];
  float acc[kTy][kTx];
  const float b0 = bias ? bias[c] : 0.f;
#pragma unroll
  for (int i = 0; i < kTy; ++i)
#pragma unroll
    for (int k = 0; k < kTx; ++k) acc[i][k] = b0;
  const int x0 = tx * kTx, y0 = ty * kTy;
#pragma unroll
  for (int r = 0; r < kTy + 6; ++r) {                     // input row y0 + r - 3 contributes to output rows oy with ky = r - oy in [0, 7)
    const int iy = y0 + r - 3;
    if (iy < 0 || iy >= H) continue;
    const float* row = x + (((long long)bb * H + iy) * W) * C + c;
    float in[kTx + 6];
#pragma unroll
    for (int i = 0; i < kTx + 6; ++i) {
      const int ix = x0 + i - 3;
      in[i] = (ix >= 0 && ix < W) ? row[(long long)ix * C] : 0.f;
    }
#pragma unroll
    for (int oy = 0; oy < kTy; ++oy) {
      const int ky = r - oy;
      if (ky < 0 || ky >= 7) continue;
#pragma unroll
      for (int kx = 0; kx < 7; ++kx)
#pragma unroll
        for (int k = 0; k < kTx; ++k) acc[oy][k] = fmaf(wk[ky * 7 + kx], in[k + kx], acc[oy][k]);
    }
  }
#pragma unroll
  for (int oy = 0; oy < kTy; ++oy)
#pragma unroll
    for (int k = 0; k < kTx; ++k)
      if (y0 + oy < H && x0 + k < W) y[(((long long)bb * H + y0 + oy) * W + x0 + k) * C + c] = acc[oy][k];
}

// dw[c, ky, kx] = sum_{b, y, x} dy[b, y, x, c] x[b, y + ky - 3, x + kx - 3, c];  db[c] = sum dy.  One block per (image, kRowsPerBlock
// output rows), one thread per channel.  For every (output row, ky) the input row and the gradient row are walked in chunks of 8
// pixels held in registers (8 + 14 loads for 7 x 8 FMAs); the 49 + 1 partial sums of a block go to part[block][50][C] with plain
// coalesced stores and a second kernel folds the blocks in a fixed order (deterministic, no atomics).
constexpr int kRowsPerBlock = 4;
constexpr int kWChunk = 8;
__global__ void __launch_bounds__(256)
dwconv7_bwd_w_kernel(const float* __restrict__ dy, const float* __restrict__ x, float* __restrict__ part, int B, int H, int W, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int nrb = (H + kRowsPerBlock - 1) / kRowsPerBlock;
  const int rb = blockIdx.y % nrb, bb = blockIdx.y / nrb;
  if (c >= C) return;
  float acc[49];
#pragma unroll
  for (int i = 0; i < 49; ++i) acc[i] = 0.f;
  float accb = 0.f;
  for (int yy = rb * kRowsPerBlock; yy < min(H, (rb + 1) * kRowsPerBlock); ++yy) {
    const float* grow = dy + (((long long)bb * H + yy) * W) * C + c;
    for (int xc = 0; xc < W; xc += kWChunk) {
      float g[kWChunk];
#pragma unroll
      for (int i = 0; i < kWChunk; ++i) { g[i] = (xc + i < W) ? grow[(long long)(xc + i) * C] : 0.f; accb += g[i]; }
#pragma unroll
      for (int ky = 0; ky < 7; ++ky) {
        const int iy = yy + ky - 3;
        if (iy < 0 || iy >= H) continue;
        const float* row = x + (((long long)bb * H + iy) * W) * C + c;
        float in[kWChunk + 6];
#pragma unroll
        for (int i = 0; i < kWChunk + 6; ++i) {
          const int ix = xc + i - 3;
          in[i] = (ix >= 0 && ix < W) ? row[(long long)ix * C] : 0.f;
        }
#pragma unroll
        for (int kx = 0; kx < 7; ++kx)
#pragma unroll
          for (int i = 0; i < kWChunk; ++i) acc[ky * 7 + kx] = fmaf(g[i], in[i + kx], acc[ky * 7 + kx]);
      }
    }
  }
  float* dst = part + (long long)blockIdx.y * 50 * C + c;
#pragma unroll
  for (int i = 0; i < 49; ++i) dst[(long long)i * C] = acc[i];
  dst[49LL * C] = accb;
}

// dw[c * 49 + i] += sum_blocks part[blk][i][c], db[c] += sum_blocks part[blk][49][c]
__global__ void __launch_bounds__(256)
dwconv7_bwd_w_fold_kernel(const float* __restrict__ part, float* __restrict__ dw, float* __restrict__ db, int nblk, int C) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;      // e = i * C + c
  if (e >= 50 * C) return;
  float s = 0.f;
  for (int k = 0; k < nblk; ++k) s += part[(long long)k * 50 * C + e];
  const int i = e / C, c = e % C;
  if (i < 49) dw[c * 49 + i] += s;
  else db[c] += s;
}

template <int C>
int launch_ln(int bwd, int act, const float* a0, const float* x, const float* w, const float* b, float* o0, float* dw, float* db, long long P,
              float eps, cudaStream_t stream) {
  using S = LnShape<C>;
  const long long warps = (P + S::PPW - 1) / S::PPW;
  if (!bwd) {
    const unsigned blocks = (unsigned)((warps * 32 + 255) / 256);
    if (act) ln_gelu_fwd_kernel<C, true><<<blocks, 256, 0, stream>>>(x, w, b, o0, P, eps);
    else ln_gelu_fwd_kernel<C, false><<<blocks, 256, 0, stream>>>(x, w, b, o0, P, eps);
  } else {
    long long blocks = (warps * 32 + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (act) ln_gelu_bwd_kernel<C, true><<<(unsigned)blocks, 256, 0, stream>>>(a0, x, w, b, o0, dw, db, P, eps);
    else ln_gelu_bwd_kernel<C, false><<<(unsigned)blocks, 256, 0, stream>>>(a0, x, w, b, o0, dw, db, P, eps);
  }
  return sam2b200::check_launch("ln_gelu");
}

}  // namespace

extern "C" {

// y [P, C] = act(LayerNorm_C(x [P, C]) * w + b), act = exact GELU (act != 0) or identity; C in {4, 16, 64, 256}; fp32.
int sam2b200_ln_gelu_fwd(const float* x, const float* w, const float* b, float* y, long long P, int C, float eps, int act,
                         cudaStream_t stream) {
  if (!x || !w || !b || !y || P <= 0) return sam2b200::fail(SAM2B200_ERR_INVALID, "ln_gelu_fwd: bad arguments");
  switch (C) {
    case 4: return launch_ln<4>(0, act, nullptr, x, w, b, y, nullptr, nullptr, P, eps, stream);
    case 16: return launch_ln<16>(0, act, nullptr, x, w, b, y, nullptr, nullptr, P, eps, stream);
    case 64: return launch_ln<64>(0, act, nullptr, x, w, b, y, nullptr, nullptr, P, eps, stream);
    case 256: return launch_ln<256>(0, act, nullptr, x, w, b, y, nullptr, nullptr, P, eps, stream);
  }
  return sam2b200::fail(SAM2B200_ERR_UNSUPPORTED, "ln_gelu: C must be 4, 16, 64 or 256");
}

// dx [P, C] = d/dx, dw / db [C] += parameter gradients (fp32 atomics) for dy [P, C].
int sam2b200_ln_gelu_bwd(const float* dy, const float* x, const float* w, const float* b, float* dx, float* dw, float* db, long long P, int C,
                         float eps, int act, cudaStream_t stream) {
  if (!dy || !x || !w || !b || !dx || !dw || !db || P <= 0) return sam2b200::fail(SAM2B200_ERR_INVALID, "ln_gelu_bwd: bad arguments");
  switch (C) {
    case 4: return launch_ln<4>(1, act, dy, x, w, b, dx, dw, db, P, eps, stream);
    case 16: return launch_ln<16>(1, act, dy, x, w, b, dx, dw, db, P, eps, stream);
    case 64: return launch_ln<64>(1, act, dy, x, w, b, dx, dw, db, P, eps, stream);
    case 256: return launch_ln<256>(1, act, dy, x, w, b, dx, dw, db, P, eps, stream);
  }
  return sam2b200::fail(SAM2B200_ERR_UNSUPPORTED, "ln_gelu: C must be 4, 16, 64 or 256");
}

// y [B, H, W, C] = depth-wise 7 x 7 correlation of x with w [C, 7, 7] (+ bias [C] or NULL), padding 3.  flip != 0 uses the taps
// mirrored in both axes: the data gradient of the same convolution (call with x = dy, bias = NULL).
int sam2b200_dwconv7(const float* x, const float* w, const float* bias, float* y, int B, int H, int W, int C, int flip, cudaStream_t stream) {
  if (!x || !w || !y || B <= 0 || H <= 0 || W <= 0 || C <= 0) return sam2b200::fail(SAM2B200_ERR_INVALID, "dwconv7: bad arguments");
  const long long gy = (long long)B * ((H + kTy - 1) / kTy) * ((W + kTx - 1) / kTx);
  if (gy > 0x7fffffffLL) return sam2b200::fail(SAM2B200_ERR_INVALID, "dwconv7: grid too large");
  const int tpb = C >= 256 ? 256 : ((C + 31) / 32) * 32;
  dim3 grid((C + tpb - 1) / tpb, (unsigned)gy);
  if (flip) dwconv7_kernel<true><<<grid, tpb, 0, stream>>>(x, w, bias, y, B, H, W, C);
  else dwconv7_kernel<false><<<grid, tpb, 0, stream>>>(x, w, bias, y, B, H, W, C);
  return sam2b200::check_launch("dwconv7");
}

size_t sam2b200_dwconv7_bwd_w_workspace_bytes(int B, int H, int C) {
  return (size_t)B * ((H + kRowsPerBlock - 1) / kRowsPerBlock) * 50 * (size_t)C * sizeof(float);
}

// dw [C, 7, 7] += , db [C] += for the same convolution; workspace: sam2b200_dwconv7_bwd_w_workspace_bytes(B, H, C) bytes.
// Deterministic (per-block partial sums folded in a fixed order).
int sam2b200_dwconv7_bwd_w(const float* dy, const float* x, float* dw, float* db, void* workspace, int B, int H, int W, int C,
                           cudaStream_t stream) {
  if (!dy || !x || !dw || !db || !workspace || B <= 0 || H <= 0 || W <= 0 || C <= 0)
    return sam2b200::fail(SAM2B200_ERR_INVALID, "dwconv7_bwd_w: bad arguments");
  const int tpb = C >= 256 ? 256 : ((C + 31) / 32) * 32;
  const int nblk = B * ((H + kRowsPerBlock - 1) / kRowsPerBlock);
  dim3 grid((C + tpb - 1) / tpb, (unsigned)nblk);
  float* part = static_cast<float*>(workspace);
  dwconv7_bwd_w_kernel<<<grid, tpb, 0, stream>>>(dy, x, part, B, H, W, C);
  dwconv7_bwd_w_fold_kernel<<<(50 * C + 255) / 256, 256, 0, stream>>>(part, dw, db, nblk, C);
  return sam2b200::check_launch("dwconv7_bwd_w", 2);
}

}  // extern "C"
