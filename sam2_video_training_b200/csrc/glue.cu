// HBM-bound fused element-wise / reduction kernels around the GEMMs and attention kernels of one
// MemoryAttentionLayer (d_model = 256).  They replace, per layer, the reference's
//   nn.LayerNorm x3 + residual adds + dropout(0) + dtype casts   (memory_attention.py:58-99)
//   autograd's LayerNorm backward, bias-gradient reductions, ReLU backward
// with one pass over the data each (the ATen kernels measured 4-10x the algorithmic bytes):
//   ln_fwd        : x' = x + residual(bf16) -> fp32;  y = LN(x') -> bf16 (+ fp32); mean, rstd
//   ln_bwd        : g_out = g_in + LN'(dy);  d gamma, d beta  (two-stage deterministic reduction)
//   cast_colsum   : fp32 [R,C] -> bf16 copy + column sums (bias gradient of the consumer GEMM)
//   colsum / relu : bf16 [R,C] column sums, optionally masking by (h > 0) in place (ReLU backward)
// All kernels: 128-bit accesses, one warp per 256-wide row, no atomics.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "abi_common.cuh"
#include "dropout.cuh"

namespace {

constexpr int kD = 256;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 t = __bfloat1622float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 u;
  __nv_bfloat162 a = __floats2bfloat162_rn(f[0], f[1]), b = __floats2bfloat162_rn(f[2], f[3]);
  __nv_bfloat162 c = __floats2bfloat162_rn(f[4], f[5]), d = __floats2bfloat162_rn(f[6], f[7]);
  u.x = *reinterpret_cast<uint32_t*>(&a); u.y = *reinterpret_cast<uint32_t*>(&b);
  u.z = *reinterpret_cast<uint32_t*>(&c); u.w = *reinterpret_cast<uint32_t*>(&d);
  return u;
}

// ------------------------------------------------------------------ LayerNorm forward
// One warp per row.  x: [R,256] fp32; res: [R,256] bf16 or null; x_out: fp32 or null (x + res);
// y16 / y32: normalised output (either may be null).  out_ld_rows: if > 0, y32 is written
// transposed as [n][b] from rows ordered [b][n] (final norm back to seq-first): row r = b*Nn + n
// goes to (n*Bb + b).
__global__ void __launch_bounds__(256)
ln_fwd_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ res, float* __restrict__ x_out,
              const float* __restrict__ gamma, const float* __restrict__ beta, __nv_bfloat16* __restrict__ y16,
              float* __restrict__ y32, float* __restrict__ mean, float* __restrict__ rstd, long long rows,
              float eps, int tr_b, int tr_n, const sam2b200::Dropout drop) {
  // drop: inverted dropout on the residual branch, x' = x + dropout(res) (memory_attention.py:64,81,99);
  // element index = row * 256 + column
  // programmatic dependent launch: the next kernel of the stream may start its set-up; this one waits for the one in front of it
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* xp = reinterpret_cast<const float4*>(x + row * kD + lane * 8);
  float v[8];
  float4 a = xp[0], b = xp[1];
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  if (res != nullptr) {
    float r[8];
    unpack8(*reinterpret_cast<const uint4*>(res + row * kD + lane * 8), r);
    if (drop.seed != nullptr) {
      const uint32_t key = sam2b200::dropout_key(*drop.seed, drop.site);
      const uint32_t idx = (uint32_t)(row * kD + lane * 8);
#pragma unroll
      for (int i = 0; i < 8; ++i) r[i] = sam2b200::dropout_keep(key, idx + i, drop.thresh) ? r[i] * drop.inv_keep : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] += r[i];
    if (x_out != nullptr) {
      float4* op = reinterpret_cast<float4*>(x_out + row * kD + lane * 8);
      op[0] = make_float4(v[0], v[1], v[2], v[3]);
      op[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i];
  const float mu = warp_sum(s) * (1.0f / kD);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { const float d = v[i] - mu; q += d * d; }
  const float rs = rsqrtf(warp_sum(q) * (1.0f / kD) + eps);
  if (lane == 0) { if (mean) mean[row] = mu; if (rstd) rstd[row] = rs; }
  const float4* gp = reinterpret_cast<const float4*>(gamma + lane * 8);
  const float4* bp = reinterpret_cast<const float4*>(beta + lane * 8);
  const float4 g0 = gp[0], g1 = gp[1], b0 = bp[0], b1 = bp[1];
  const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
  const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
  float y[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) y[i] = (v[i] - mu) * rs * gg[i] + bb[i];
  if (y16 != nullptr) *reinterpret_cast<uint4*>(y16 + row * kD + lane * 8) = pack8(y);
  if (y32 != nullptr) {
    long long orow = row;
    if (tr_b > 0) { const long long bi = row / tr_n, ni = row % tr_n; orow = ni * tr_b + bi; }
    float4* op = reinterpret_cast<float4*>(y32 + orow * kD + lane * 8);
    op[0] = make_float4(y[0], y[1], y[2], y[3]);
    op[1] = make_float4(y[4], y[5], y[6], y[7]);
  }
}

// ------------------------------------------------------------------ LayerNorm backward
// dy: bf16 [R,256] (dy16) or fp32 (dy32; tr_b > 0 reads it transposed like ln_fwd writes y32).
// g_out = (g_in ? g_in : 0) + dx.  Per-block partial d gamma / d beta -> part[blk][2][256].
__global__ void __launch_bounds__(256)
ln_bwd_kernel(const __nv_bfloat16* __restrict__ dy16, const float* __restrict__ dy32, const float* __restrict__ x,
              const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
              const float* __restrict__ g_in, float* __restrict__ g_out, __nv_bfloat16* __restrict__ g16,
              float* __restrict__ part, long long rows, int tr_b, int tr_n, const sam2b200::Dropout drop) {
  // drop: the dropout that sat on the branch whose output gradient g16 is (x' = x + dropout(branch)): g_out (the
  // residual-stream gradient) is not masked, its bf16 copy for the branch is.
  // g16 (optional): bf16 copy of g_out -- the operand of the next GEMMs of the backward chain -- and a third partial
  // row with its column sums (the bias gradient of the projection whose output gradient g_out is).
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");     // programmatic dependent launch (see ln_fwd_kernel)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float4* gp = reinterpret_cast<const float4*>(gamma + lane * 8);
  const float4 g0 = gp[0], g1 = gp[1];
  const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
  float dgam[8], dbet[8], dbias[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { dgam[i] = 0.f; dbet[i] = 0.f; dbias[i] = 0.f; }
  const int nrow = (g16 != nullptr) ? 3 : 2;
  const long long warps_total = (long long)gridDim.x * (blockDim.x >> 5);
  const uint32_t drop_key = (g16 != nullptr && drop.seed != nullptr) ? sam2b200::dropout_key(*drop.seed, drop.site) : 0u;
  // Two rows per iteration: the loads of both rows (dy, x, g_in, mean, rstd) are issued before any arithmetic, which
  // doubles the bytes a warp keeps in flight (the kernel is latency bound at 16 warps per SM).
#ifndef SAM2B200_LN_BWD_UNROLL
#define SAM2B200_LN_BWD_UNROLL 2
#endif
  constexpr int kU = SAM2B200_LN_BWD_UNROLL;
  for (long long base = (long long)blockIdx.x * (blockDim.x >> 5) + warp; base < rows; base += kU * warps_total) {
    float dy[kU][8], xv[kU][8], gi[kU][8], mu[kU], rs[kU];
    bool ok[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const long long row = base + u * warps_total;
      ok[u] = row < rows;
      if (!ok[u]) continue;
      if (dy16 != nullptr) {
        unpack8(*reinterpret_cast<const uint4*>(dy16 + row * kD + lane * 8), dy[u]);
      } else {
        long long irow = row;
        if (tr_b > 0) { const long long bi = row / tr_n, ni = row % tr_n; irow = ni * tr_b + bi; }
        const float4* dp = reinterpret_cast<const float4*>(dy32 + irow * kD + lane * 8);
        const float4 a = dp[0], b = dp[1];
        dy[u][0] = a.x; dy[u][1] = a.y; dy[u][2] = a.z; dy[u][3] = a.w; dy[u][4] = b.x; dy[u][5] = b.y; dy[u][6] = b.z; dy[u][7] = b.w;
      }
      const float4* xp = reinterpret_cast<const float4*>(x + row * kD + lane * 8);
      const float4 a = xp[0], b = xp[1];
      xv[u][0] = a.x; xv[u][1] = a.y; xv[u][2] = a.z; xv[u][3] = a.w; xv[u][4] = b.x; xv[u][5] = b.y; xv[u][6] = b.z; xv[u][7] = b.w;
      if (g_in != nullptr) {
        const float4* ip = reinterpret_cast<const float4*>(g_in + row * kD + lane * 8);
        const float4 c = ip[0], d = ip[1];
        gi[u][0] = c.x; gi[u][1] = c.y; gi[u][2] = c.z; gi[u][3] = c.w; gi[u][4] = d.x; gi[u][5] = d.y; gi[u][6] = d.z; gi[u][7] = d.w;
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) gi[u][i] = 0.f;
      }
      mu[u] = mean[row]; rs[u] = rstd[row];
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      if (!ok[u]) continue;                            // warp-uniform
      const long long row = base + u * warps_total;
      float xh[8], s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        xh[i] = (xv[u][i] - mu[u]) * rs[u];
        const float dg = dy[u][i] * gg[i];
        s1 += dg; s2 += dg * xh[i];
        dgam[i] += dy[u][i] * xh[i];
        dbet[i] += dy[u][i];
      }
      s1 = warp_sum(s1) * (1.0f / kD);
      s2 = warp_sum(s2) * (1.0f / kD);
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = rs[u] * (dy[u][i] * gg[i] - s1 - xh[i] * s2) + gi[u][i];
      float4* op = reinterpret_cast<float4*>(g_out + row * kD + lane * 8);
      op[0] = make_float4(o[0], o[1], o[2], o[3]);
      op[1] = make_float4(o[4], o[5], o[6], o[7]);
      if (g16 != nullptr) {
        if (drop.seed != nullptr) {
          const uint32_t idx = (uint32_t)(row * kD + lane * 8);
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] = sam2b200::dropout_keep(drop_key, idx + i, drop.thresh) ? o[i] * drop.inv_keep : 0.f;
        }
        const uint4 uu = pack8(o);
        *reinterpret_cast<uint4*>(g16 + row * kD + lane * 8) = uu;
        float r[8];
        unpack8(uu, r);                                 // sum what the consumer GEMM will see (the bf16-rounded values)
#pragma unroll
        for (int i = 0; i < 8; ++i) dbias[i] += r[i];
      }
    }
  }
  __shared__ float sh[8][3][kD];
#pragma unroll
  for (int i = 0; i < 8; ++i) { sh[warp][0][lane * 8 + i] = dgam[i]; sh[warp][1][lane * 8 + i] = dbet[i]; sh[warp][2][lane * 8 + i] = dbias[i]; }
  __syncthreads();
  for (int i = threadIdx.x; i < nrow * kD; i += blockDim.x) {
    const int which = i / kD, c = i % kD;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += sh[w][which][c];
    part[((long long)blockIdx.x * nrow + which) * kD + c] = s;
  }
}

// out[c] += sum_blk part[blk][c]   (fixed order -> deterministic).  32 columns x 8 block-groups per CTA.
__global__ void __launch_bounds__(256)
partial_reduce_add_kernel(const float* __restrict__ part, int nblk, int width, float* __restrict__ out0,
                          float* __restrict__ out1, int split, float* __restrict__ out2 = nullptr) {
  // part: [nblk][width]; columns [0, split) go to out0, [split, 2 split) to out1, [2 split, width) to out2
  // (LN: gamma | beta | bias of the next projection)
  __shared__ float sh[8][32];
  const int cl = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  float s = 0.f;
  if (c < width)
    for (int b = g; b < nblk; b += 8) s += part[(long long)b * width + c];
  sh[g][cl] = s;
  __syncthreads();
  if (g == 0 && c < width) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += sh[i][cl];
    if (c < split) out0[c] += t; else if (c < 2 * split || out2 == nullptr) out1[c - split] += t; else out2[c - 2 * split] += t;
  }
}

// ------------------------------------------------------------------ column sums (bias gradients)
// Thread owns 8 consecutive columns; a block covers (256 / (C/8)) rows per iteration.
// MODE 0: in = fp32 [R,C], writes bf16 copy to out16 (cast) + column sums
// MODE 1: in = bf16 [R,C] (io16), masked in place by (h16 > 0) + column sums   (ReLU backward)
// MODE 2: in = bf16 [R,C] with row stride ld, plain column sums
template <int MODE>
__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ in32, __nv_bfloat16* __restrict__ io16, const __nv_bfloat16* __restrict__ h16,
              float* __restrict__ part, long long rows, int C, long long ld, float scale) {
  // scale (MODE 1): 1 / (1 - p) of the dropout that followed the ReLU -- h16 is the DROPPED activation, so (h > 0)
  // is the ReLU mask and the dropout mask at once
  const int tpr = C >> 3;                         // threads per row
  const int rpb = blockDim.x / tpr;               // rows per block iteration
  const int cgrp = threadIdx.x % tpr;
  const int rsub = threadIdx.x / tpr;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  if (rsub < rpb) {
    for (long long row = (long long)blockIdx.x * rpb + rsub; row < rows; row += (long long)gridDim.x * rpb) {
      float v[8];
      if (MODE == 0) {
        const float4* p = reinterpret_cast<const float4*>(in32 + row * ld + cgrp * 8);
        const float4 a = p[0], b = p[1];
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        *reinterpret_cast<uint4*>(io16 + row * C + cgrp * 8) = pack8(v);
        // sum what the consumer GEMM will see (the bf16-rounded values)
        float r[8];
        uint4 u = pack8(v);
        unpack8(u, r);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = r[i];
      } else {
        uint4 u = *reinterpret_cast<const uint4*>(io16 + row * ld + cgrp * 8);
        unpack8(u, v);
        if (MODE == 1) {
          float hv[8];
          unpack8(*reinterpret_cast<const uint4*>(h16 + row * ld + cgrp * 8), hv);
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = hv[i] > 0.f ? v[i] * scale : 0.f;
          *reinterpret_cast<uint4*>(io16 + row * ld + cgrp * 8) = pack8(v);
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += v[i];
    }
  }
  // reduce the rpb row-groups of the block through shared memory
  extern __shared__ float sh[];                   // [rpb][C]
  if (rsub < rpb) {
#pragma unroll
    for (int i = 0; i < 8; ++i) sh[rsub * C + cgrp * 8 + i] = acc[i];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int r = 0; r < rpb; ++r) s += sh[r * C + c];
    part[(long long)blockIdx.x * C + c] = s;
  }
}

__global__ void __launch_bounds__(256)
dropout_inplace_kernel(__nv_bfloat16* __restrict__ x, long long groups, const sam2b200::Dropout drop) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= groups) return;
  const uint32_t key = sam2b200::dropout_key(*drop.seed, drop.site);
  float v[8];
  unpack8(*reinterpret_cast<const uint4*>(x + g * 8), v);
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = sam2b200::dropout_keep(key, (uint32_t)(g * 8 + i), drop.thresh) ? v[i] * drop.inv_keep : 0.f;
  *reinterpret_cast<uint4*>(x + g * 8) = pack8(v);
}

__global__ void dropout_mask_kernel(unsigned char* __restrict__ out, long long index0, long long n, const sam2b200::Dropout drop) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const bool keep = drop.seed == nullptr || sam2b200::dropout_keep(sam2b200::dropout_key(*drop.seed, drop.site), (uint32_t)(index0 + i), drop.thresh);
  out[i] = keep ? 1 : 0;
}

// ------------------------------------------------------------------ packing between the module layout and the kernel layout
// MemoryAttention.forward receives seq-first tensors ([L, B, C], memory_attention.py:119-148); the stack works batch-first.
// out[(b * L + l), :] = a[(l * B + b), :] (+ alpha2 * a2[(l * B + b), :]) as fp32 (OutT = float) or bf16, optionally a second
// output out2 from a alone (memory: memk = bf16(memory + pos), memv = bf16(memory) in ONE pass).  One warp per row; C = 256 or 64.
// inverse = 1 maps batch-first rows back to seq-first (gradients): out[(l * B + b), :] = scale * (a[(b * L + l), :] + a2[...]).
template <typename OutT, int C>
__global__ void __launch_bounds__(256)
permute_rows_kernel(const float* __restrict__ a, const float* __restrict__ a2, float alpha2, OutT* __restrict__ out,
                    OutT* __restrict__ out2, long long rows, int B, int L, int inverse, float scale) {
  constexpr int kPerLane = C / 32;                  // 8 (C = 256) or 2 (C = 64) consecutive floats per lane
  const long long orow = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (orow >= rows) return;
  long long irow;
  if (!inverse) { const long long b = orow / L, l = orow % L; irow = l * B + b; }      // out batch-first <- in seq-first
  else { const long long l = orow / B, b = orow % B; irow = b * L + l; }               // out seq-first <- in batch-first
  float v[kPerLane], w[kPerLane], t[kPerLane];
  auto load = [&](const float* src, float* dst) {     // 2 x float4 (C = 256) or 1 x float2 (C = 64) per lane
    if constexpr (kPerLane == 8) {
      const float4* p4 = reinterpret_cast<const float4*>(src + irow * C + lane * 8);
      const float4 x0 = p4[0], x1 = p4[1];
      dst[0] = x0.x; dst[1] = x0.y; dst[2] = x0.z; dst[3] = x0.w; dst[4] = x1.x; dst[5] = x1.y; dst[6] = x1.z; dst[7] = x1.w;
    } else {
      const float2 x0 = *reinterpret_cast<const float2*>(src + irow * C + lane * 2);
      dst[0] = x0.x; dst[1] = x0.y;
    }
  };
  load(a, w);
#pragma unroll
  for (int i = 0; i < kPerLane; ++i) v[i] = w[i];
  if (a2 != nullptr) {
    load(a2, t);
#pragma unroll
    for (int i = 0; i < kPerLane; ++i) v[i] = fmaf(alpha2, t[i], v[i]);
  }
#pragma unroll
  for (int i = 0; i < kPerLane; ++i) { v[i] *= scale; w[i] *= scale; }
  auto store = [&](OutT* dst, const float* src) {
    if constexpr (sizeof(OutT) == 4 && kPerLane == 8) {
      float4* p4 = reinterpret_cast<float4*>(dst + orow * C + lane * 8);
      p4[0] = make_float4(src[0], src[1], src[2], src[3]);
      p4[1] = make_float4(src[4], src[5], src[6], src[7]);
    } else if constexpr (sizeof(OutT) == 4) {
      *reinterpret_cast<float2*>(dst + orow * C + lane * 2) = make_float2(src[0], src[1]);
    } else if constexpr (kPerLane == 8) {
      *reinterpret_cast<uint4*>(dst + orow * C + lane * 8) = pack8(src);
    } else {
      *reinterpret_cast<__nv_bfloat162*>(dst + orow * C + lane * 2) = __floats2bfloat162_rn(src[0], src[1]);
    }
  };
  store(out, v);
  if (out2 != nullptr) store(out2, w);
}

int grid_for_rows(long long rows, int rows_per_block, int blocks_per_sm = 2) {
  long long g = (rows + rows_per_block - 1) / rows_per_block;
  const long long cap = 148 * blocks_per_sm;
  return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}

// resident blocks per SM of ln_bwd_kernel (bytes in flight): measured on B200, see profiles/r1_ln_bwd_occupancy.txt
int ln_bwd_blocks_per_sm() {
  static int v = 0;
  if (v == 0) {
    const char* e = getenv("SAM2B200_LN_BWD_BLOCKS_PER_SM");
    v = e ? atoi(e) : 2;
    if (v < 1 || v > 8) v = 2;
  }
  return v;
}

// Parameter gradients of the FOLDED cross-attention projection ca = out64 (Wo Wv)^T + (Wo bv + bo)   (fused_stack.py, raw-memory path):
// given G = d/d(Wo Wv) = dca^T out64 [256, 64], g = colsum(dca) [256] and g_rs (= g without attention dropout),
//   dWo [256, 256] += G Wv^T + g_rs (x) bv      dWv [256, 64] += Wo^T G      dbo [256] += g      dbv [64 | 256] += Wo^T g_rs
// Four products of at most 256 x 256 x 64 (17 MFLOP in all): one launch instead of the nine tiny ATen / cuBLAS launches per layer and
// frame (outer product, two SIMT SGEMMs, GEMV, four adds).  Block b: row b of dWo (thread j = column j), row b of dWv (threads 0..63)
// and element b of dbo / dbv.  Wv / bv: the fp32 master weights of v_proj ([256, 64] / [256]); Wo [256, 256].
__global__ void __launch_bounds__(256)
fold_grads_kernel(const float* __restrict__ G, const float* __restrict__ g, const float* __restrict__ g_rs, const float* __restrict__ Wo,
                  const float* __restrict__ Wv, const float* __restrict__ bv, float* __restrict__ dWo, float* __restrict__ dWv,
                  float* __restrict__ dbo, float* __restrict__ dbv) {
  __shared__ float g_row[64];           // G[b, :]
  __shared__ float wo_col[256];         // Wo[:, b]
  __shared__ float red[8];
  __shared__ float part4[256];
  const int b = blockIdx.x, j = threadIdx.x;
  if (j < 64) g_row[j] = G[b * 64 + j];
  wo_col[j] = Wo[j * 256 + b];
  __syncthreads();
  // dWo[b, j] += sum_k G[b, k] Wv[j, k] + g_rs[b] bv[j]
  {
    const float4* wv = reinterpret_cast<const float4*>(Wv + j * 64);
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const float4 w = wv[k];
      acc += g_row[4 * k] * w.x + g_row[4 * k + 1] * w.y + g_row[4 * k + 2] * w.z + g_row[4 * k + 3] * w.w;
    }
    dWo[b * 256 + j] += acc + g_rs[b] * bv[j];
  }
  // dWv[b, c] += sum_i Wo[i, b] G[i, c]      (thread j: column c = j % 64, rows i = (j / 64) * 64 .. + 63; coalesced reads of G rows)
  {
    const int c = j & 63, i0 = (j >> 6) * 64;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 4
    for (int i = 0; i < 64; i += 4) {
      a0 += wo_col[i0 + i] * G[(i0 + i) * 64 + c];
      a1 += wo_col[i0 + i + 1] * G[(i0 + i + 1) * 64 + c];
      a2 += wo_col[i0 + i + 2] * G[(i0 + i + 2) * 64 + c];
      a3 += wo_col[i0 + i + 3] * G[(i0 + i + 3) * 64 + c];
    }
    part4[j] = (a0 + a1) + (a2 + a3);
  }
  __syncthreads();
  if (j < 64) dWv[b * 64 + j] += (part4[j] + part4[64 + j]) + (part4[128 + j] + part4[192 + j]);
  // dbv[b] += sum_i Wo[i, b] g_rs[i];   dbo[b] += g[b]
  float part = wo_col[j] * g_rs[j];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if ((j & 31) == 0) red[j >> 5] = part;
  __syncthreads();
  if (j == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w];
    dbv[b] += t;
    dbo[b] += g[b];
  }
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace

extern "C" {

// y = LayerNorm(x + res) over the last dim (256).  Any of res / x_out / y16 / y32 / mean / rstd may be
// NULL.  tr_b > 0: y32 is written seq-first ([n][b][256]) from batch-first rows (b*tr_n + n).
int sam2b200_ln_fwd(const float* x, const void* res_bf16, float* x_out, const float* gamma, const float* beta,
                    void* y_bf16, float* y_f32, float* mean, float* rstd, long long rows, float eps, int tr_b,
                    int tr_n, float drop_p, const unsigned long long* drop_seed, unsigned drop_site, cudaStream_t stream) {
  if (!x || !gamma || !beta || rows <= 0 || !aligned16(x) || (res_bf16 && !aligned16(res_bf16)) ||
      (x_out && !aligned16(x_out)) || (y_bf16 && !aligned16(y_bf16)) || (y_f32 && !aligned16(y_f32)) || drop_p < 0.f ||
      drop_p >= 1.f || rows * kD >= (1LL << 32))
    return sam2b200::fail(SAM2B200_ERR_INVALID, "ln_fwd: bad arguments");
  const long long blocks = (rows * 32 + 255) / 256;
  if (sam2b200::launch_pdl(ln_fwd_kernel, dim3((unsigned)blocks), dim3(256), 0, stream, x, (const __nv_bfloat16*)res_bf16, x_out, gamma, beta,
                                                       (__nv_bfloat16*)y_bf16, y_f32, mean, rstd, rows, eps, tr_b, tr_n,
                                                       sam2b200::make_dropout(drop_seed, drop_site, drop_p)) != cudaSuccess)
    return sam2b200::fail(SAM2B200_ERR_CUDA, "ln_fwd: launch failed");
  return sam2b200::check_launch("ln_fwd");
}

size_t sam2b200_ln_bwd_workspace_bytes(long long rows) {
  return (size_t)grid_for_rows(rows, 8 * 8, 8) * 3 * kD * sizeof(float);   // sized for the largest grid
}

// g_out = g_in + dLN/dx; dgamma += ..., dbeta += ... (accumulated into the given fp32 buffers).
// Exactly one of dy_bf16 / dy_f32 is non-NULL.
// g_out_bf16 / dbias (optional, both or neither): bf16 copy of g_out for the next GEMMs and dbias += its column sums.
// stages: bit 0 = the row pass (g_out, g_out_bf16, per-block partial sums in `workspace`), bit 1 = the fold of the partial sums into
// dgamma / dbeta / dbias.  The fold only feeds parameter gradients, so a caller may enqueue it on another stream (after an event
// behind the row pass) and keep the 24-CTA kernel off the critical path of the residual-stream gradient.
int sam2b200_ln_bwd_stages(const void* dy_bf16, const float* dy_f32, const float* x, const float* mean, const float* rstd,
                           const float* gamma, const float* g_in, float* g_out, void* g_out_bf16, float* dgamma, float* dbeta,
                           float* dbias, void* workspace, long long rows, int tr_b, int tr_n, float drop_p,
                           const unsigned long long* drop_seed, unsigned drop_site, int stages, cudaStream_t stream) {
  if ((!dy_bf16) == (!dy_f32) || !x || !mean || !rstd || !gamma || !g_out || !dgamma || !dbeta || !workspace ||
      rows <= 0 || (!g_out_bf16) != (!dbias) || drop_p < 0.f || drop_p >= 1.f || (stages & 3) == 0)
    return sam2b200::fail(SAM2B200_ERR_INVALID, "ln_bwd: bad arguments");
  const int nblk = grid_for_rows(rows, 8 * 8, ln_bwd_blocks_per_sm());
  const int nrow = g_out_bf16 ? 3 : 2;
  float* part = static_cast<float*>(workspace);
  int launches = 0;
  if (stages & 1) {
    if (sam2b200::launch_pdl(ln_bwd_kernel, dim3(nblk), dim3(256), 0, stream, (const __nv_bfloat16*)dy_bf16, dy_f32, x, mean, rstd, gamma, g_in, g_out,
                                            (__nv_bfloat16*)g_out_bf16, part, rows, tr_b, tr_n,
                                            sam2b200::make_dropout(drop_seed, drop_site, drop_p)) != cudaSuccess)
      return sam2b200::fail(SAM2B200_ERR_CUDA, "ln_bwd: launch failed");
    ++launches;
  }
  if (stages & 2) {
    partial_reduce_add_kernel<<<(nrow * kD + 31) / 32, 256, 0, stream>>>(part, nblk, nrow * kD, dgamma, dbeta, kD, dbias);
    ++launches;
  }
  return sam2b200::check_launch("ln_bwd", launches);
}

int sam2b200_ln_bwd(const void* dy_bf16, const float* dy_f32, const float* x, const float* mean, const float* rstd,
                    const float* gamma, const float* g_in, float* g_out, void* g_out_bf16, float* dgamma, float* dbeta,
                    float* dbias, void* workspace, long long rows, int tr_b, int tr_n, float drop_p,
                    const unsigned long long* drop_seed, unsigned drop_site, cudaStream_t stream) {
  return sam2b200_ln_bwd_stages(dy_bf16, dy_f32, x, mean, rstd, gamma, g_in, g_out, g_out_bf16, dgamma, dbeta, dbias, workspace, rows, tr_b, tr_n,
                                drop_p, drop_seed, drop_site, 3, stream);
}

size_t sam2b200_colsum_workspace_bytes(long long rows, int C) {
  const int rpb = 256 / (C / 8);
  return (size_t)grid_for_rows(rows, rpb * 16) * C * sizeof(float);
}

// mode 0: in_f32 [R,C] -> out bf16 [R,C] (cast) and colsum += column sums of the rounded values
// mode 1: io_bf16 [R,C] *= (h_bf16 > 0) in place and colsum += column sums   (ReLU backward + bias grad)
// mode 2: colsum += column sums of io_bf16 (row stride ld elements)
// C must be a multiple of 8 with C/8 dividing 256 (256, 512, 1024, 2048) or equal to 768.
int sam2b200_colsum(int mode, const float* in_f32, void* io_bf16, const void* h_bf16, float* colsum, void* workspace,
                    long long rows, int C, long long ld, float scale, cudaStream_t stream) {
  if (rows <= 0 || C <= 0 || (C % 8) || (C / 8) > 256 || !colsum || !workspace || !io_bf16 || (mode == 0 && !in_f32) ||
      (mode == 1 && !h_bf16) || mode < 0 || mode > 2)
    return sam2b200::fail(SAM2B200_ERR_INVALID, "colsum: bad arguments");
  const int tpr = C / 8;
  const int rpb = 256 / tpr;
  if (ld <= 0) ld = C;
  const int nblk = grid_for_rows(rows, rpb * 16);
  float* part = static_cast<float*>(workspace);
  const size_t sh = (size_t)rpb * C * sizeof(float);
  if (mode == 0)
    colsum_kernel<0><<<nblk, 256, sh, stream>>>(in_f32, (__nv_bfloat16*)io_bf16, nullptr, part, rows, C, ld, 1.0f);
  else if (mode == 1)
    colsum_kernel<1><<<nblk, 256, sh, stream>>>(nullptr, (__nv_bfloat16*)io_bf16, (const __nv_bfloat16*)h_bf16, part,
                                                rows, C, ld, scale);
  else
    colsum_kernel<2><<<nblk, 256, sh, stream>>>(nullptr, (__nv_bfloat16*)io_bf16, nullptr, part, rows, C, ld, 1.0f);
  partial_reduce_add_kernel<<<(C + 31) / 32, 256, 0, stream>>>(part, nblk, C, colsum, colsum, C);
  return sam2b200::check_launch("colsum", 2);
}

// Row permutation between seq-first [L, B, C] and batch-first [B, L, C] with fused add / scale / cast (C = 256 or 64):
//   inverse = 0: out[b, l] = scale * (a[l, b] + alpha2 * a2[l, b]), out2[b, l] = scale * a[l, b]   (a2, out2 optional)
//   inverse = 1: out[l, b] = scale * (a[b, l] + alpha2 * a2[b, l]), out2[l, b] = scale * a[b, l]
// out_bf16 = 1 writes bf16.  Replaces curr.float() + 0.1 * curr_pos -> transpose -> contiguous, (memory + memory_pos)
// -> transpose -> bf16 (memory_attention.py:140-148, :75-76) and the transposes of the input gradients.
int sam2b200_permute_rows(const float* a, const float* a2, float alpha2, void* out, void* out2, int out_bf16, int B, int L,
                          int C, int inverse, float scale, cudaStream_t stream) {
  if (!a || !out || B <= 0 || L <= 0 || (C != 256 && C != 64) || (out_bf16 & ~1) || !aligned16(a) || (a2 && !aligned16(a2)) ||
      !aligned16(out) || (out2 && !aligned16(out2)))
    return sam2b200::fail(SAM2B200_ERR_INVALID, "permute_rows: bad arguments (C must be 256 or 64, 16-byte aligned tensors)");
  const long long rows = (long long)B * L;
  const unsigned blocks = (unsigned)((rows * 32 + 255) / 256);
  if (C == 256 && !out_bf16)
    permute_rows_kernel<float, 256><<<blocks, 256, 0, stream>>>(a, a2, alpha2, (float*)out, (float*)out2, rows, B, L, inverse, scale);
  else if (C == 256)
    permute_rows_kernel<__nv_bfloat16, 256><<<blocks, 256, 0, stream>>>(a, a2, alpha2, (__nv_bfloat16*)out, (__nv_bfloat16*)out2, rows, B, L, inverse, scale);
  else if (!out_bf16)
    permute_rows_kernel<float, 64><<<blocks, 256, 0, stream>>>(a, a2, alpha2, (float*)out, (float*)out2, rows, B, L, inverse, scale);
  else
    permute_rows_kernel<__nv_bfloat16, 64><<<blocks, 256, 0, stream>>>(a, a2, alpha2, (__nv_bfloat16*)out, (__nv_bfloat16*)out2, rows, B, L, inverse, scale);
  return sam2b200::check_launch("permute_rows");
}

// Inverted dropout in place on a bf16 tensor of n elements (the MLP's hidden activation, memory_attention.py:97:
// linear2(dropout(relu(linear1(x))))); element index = flat index.
int sam2b200_dropout_inplace(void* x_bf16, long long n, float drop_p, const unsigned long long* drop_seed,
                             unsigned drop_site, cudaStream_t stream) {
  if (!x_bf16 || n <= 0 || (n % 8) || n >= (1LL << 32) || !aligned16(x_bf16) || drop_p < 0.f || drop_p >= 1.f || !drop_seed)
    return sam2b200::fail(SAM2B200_ERR_INVALID, "dropout_inplace: bad arguments");
  if (drop_p == 0.f) return SAM2B200_OK;
  const long long groups = n / 8;
  dropout_inplace_kernel<<<(unsigned)((groups + 255) / 256), 256, 0, stream>>>((__nv_bfloat16*)x_bf16, groups,
                                                                               sam2b200::make_dropout(drop_seed, drop_site, drop_p));
  return sam2b200::check_launch("dropout_inplace");
}

// Test aid: the keep mask (1 = kept) of elements [index0, index0 + n) of one dropout site, as bytes.
int sam2b200_dropout_mask(unsigned char* out, long long index0, long long n, float drop_p,
                          const unsigned long long* drop_seed, unsigned drop_site, cudaStream_t stream) {
  if (!out || n <= 0 || index0 < 0 || index0 + n > (1LL << 32) || drop_p < 0.f || drop_p >= 1.f || !drop_seed)
    return sam2b200::fail(SAM2B200_ERR_INVALID, "dropout_mask: bad arguments");
  dropout_mask_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(out, index0, n, sam2b200::make_dropout(drop_seed, drop_site, drop_p));
  return sam2b200::check_launch("dropout_mask");
}

// Gradients of the folded cross-attention projection (see fold_grads_kernel): all fp32, accumulated in place.  G [256, 64], g / g_rs [256]
// (g_rs = g unless attention dropout is on), Wo [256, 256], Wv [256, 64], bv [256]; dWo [256, 256], dWv [256, 64], dbo / dbv [256].
int sam2b200_fold_grads(const float* G, const float* g, const float* g_rs, const float* Wo, const float* Wv, const float* bv,
                        float* dWo, float* dWv, float* dbo, float* dbv, cudaStream_t stream) {
  if (!G || !g || !g_rs || !Wo || !Wv || !bv || !dWo || !dWv || !dbo || !dbv || !aligned16(Wv))
    return sam2b200::fail(SAM2B200_ERR_INVALID, "fold_grads: bad arguments");
  fold_grads_kernel<<<256, 256, 0, stream>>>(G, g, g_rs, Wo, Wv, bv, dWo, dWv, dbo, dbv);
  return sam2b200::check_launch("fold_grads");
}

}  // extern "C"
