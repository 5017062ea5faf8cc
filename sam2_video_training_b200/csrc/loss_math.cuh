// Per-pixel arithmetic shared by the mask-loss kernels (mask_loss.cu, merge_loss.cu): streaming loads, SFU
// approximations, the z = (t ? -x : x) softplus trick and the per-thread accumulator.  See mask_loss.cu for the algebra.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lossmath {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

__device__ __forceinline__ float4 ldg_f4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ uint32_t ldg_u32(const uint8_t* p) {
  uint32_t r;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_f4(float* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x),
               "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// z -> (d = 1 + exp(-z), softplus(z)); z is clamped below so that exp(-z) stays finite.
template <bool UNIT_T>
__device__ __forceinline__ void softplus_terms(float zraw, float inv_temp, float& d, float& sp) {
  const float z = fmaxf(UNIT_T ? zraw : zraw * inv_temp, -80.0f);
  d = 1.0f + ex2_approx(z * -kLog2e);
  sp = fmaf(lg2_approx(d), kLn2, z);       // z + log(1 + exp(-z))
}

// MODE 0 (focal + dice + IoU):  f0 / f1 = sum over background / foreground px of softplus(z) q^gamma,
//   a = sum over foreground of q, bq = sum of q, cnt = packed counters T | P << 8 | I << 16
//   (T = #foreground, P = #(x > 0), I = #(x > 0 and foreground); <= 255 px per thread).
// MODE 1 (BCE):  f0 / f1 = sum over background / foreground px of softplus(z), cnt = T.
template <int MODE, bool G2, bool UNIT_T>
struct Acc {
  float f0, f1, a, bq;
  uint32_t cnt, inc_fg, inc_bg;
  __device__ __forceinline__ void init() {
    f0 = f1 = a = bq = 0.f; cnt = 0u;
    // counter increments of a pixel with x > 0, kept in registers (opaque to the compiler, so that the
    // per-pixel update is SEL + predicated IADD instead of two re-materialised immediates + SEL + IADD)
    asm volatile("mov.u32 %0, 0x10100;" : "=r"(inc_fg));
    asm volatile("mov.u32 %0, 0x100;" : "=r"(inc_bg));
  }
  __device__ __forceinline__ void add(float x, bool t, float inv_temp, float gamma) {
    float d, sp;
    softplus_terms<UNIT_T>(t ? -x : x, inv_temp, d, sp);
    if (MODE == 0) {
      const float q = rcp_approx(d);
      float w;                                        // softplus * q^(gamma-1)
      if (G2) w = sp * q;
      else w = (gamma == 0.f) ? sp * d : sp * __powf(q, gamma - 1.0f);   // d = 1/q
      if (t) { f1 = fmaf(w, q, f1); a += q; cnt += 1u; }
      else f0 = fmaf(w, q, f0);
      bq += q;
      if (x > 0.f) cnt += t ? inc_fg : inc_bg;
    } else {
      if (t) { f1 += sp; cnt += 1u; }
      else f0 += sp;
    }
  }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace lossmath
