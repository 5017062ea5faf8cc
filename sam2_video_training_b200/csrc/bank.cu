// Memory-bank assembly for MemoryAttention in two launches (spatial frames, pointer tokens; HBM-bound gather / transpose).
//
// Replaces the data movement of SAM2Base._prepare_memory_conditioned_features
// (sam2_video/model/modeling/sam2_base.py:597-692): per selected past frame
//     feats.flatten(2).permute(2, 0, 1)                         (:602-603)
//     maskmem_pos_enc.flatten(2).permute(2, 0, 1) + tpos[...]   (:605-610)
// for the object pointers the split into C / mem_dim tokens (:666-672), and the two torch.cat calls (:691-692):
// ~3 kernels + 2 copies per frame and two concatenations in the reference, one pass here.  Which frames and pointers
// enter the bank is decided on the host (memory_bank.py), exactly like the reference.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "abi_common.cuh"

namespace {

constexpr int kMaxSlots = 40;     // spatial memory frames per call (num_maskmem - 1 + selected conditioning frames)
constexpr int kMaxPtrs = 64;      // object pointers per call
constexpr int kMd = 64;           // mem_dim

struct BankArgs {
  const void* feats[kMaxSlots];   // [B, 64, HW] fp32 | bf16
  const void* pos[kMaxSlots];     // [B, 64, HW] fp32 | bf16
  const float* tpos[kMaxSlots];   // [64] fp32: maskmem_tpos_enc row of the slot
  const void* ptrs[kMaxPtrs];     // [B, C] fp32 | bf16
};

template <typename T>
__device__ __forceinline__ float ldf(const void* p, long long i) {
  if constexpr (sizeof(T) == 4) return static_cast<const float*>(p)[i];
  else return __bfloat162float(static_cast<const __nv_bfloat16*>(p)[i]);
}

// grid: (ceil(HW / 32), B, n_slots); block 256.  Tile: 64 channels x 32 tokens through shared memory, so that reads are
// contiguous along tokens ([B, 64, HW]) and writes contiguous along channels ([M, B, 64]).
template <typename T>
__global__ void __launch_bounds__(256)
bank_spatial_kernel(const __grid_constant__ BankArgs a, float* __restrict__ memory, float* __restrict__ memory_pos, int B, int HW) {
  __shared__ float tf[kMd][33], tp[kMd][33];
  const int s = blockIdx.z, b = blockIdx.y, t0 = blockIdx.x * 32;
  const int tok = threadIdx.x & 31, c0 = threadIdx.x >> 5;
  const void* f = a.feats[s];
  const void* p = a.pos[s];
  const long long base = (long long)b * kMd * HW;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int ch = c0 + 8 * k;
    const bool ok = t0 + tok < HW;
    tf[ch][tok] = ok ? ldf<T>(f, base + (long long)ch * HW + t0 + tok) : 0.f;
    tp[ch][tok] = ok ? ldf<T>(p, base + (long long)ch * HW + t0 + tok) : 0.f;
  }
  __syncthreads();
  const int ot = threadIdx.x >> 3, och = (threadIdx.x & 7) * 8;     // 32 tokens x 8 channel groups of 8
  if (t0 + ot < HW) {
    const float* tpos = a.tpos[s];
    float vf[8], vp[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { vf[i] = tf[och + i][ot]; vp[i] = tp[och + i][ot] + tpos[och + i]; }
    const long long o = (((long long)s * HW + t0 + ot) * B + b) * kMd + och;
    *reinterpret_cast<float4*>(memory + o) = make_float4(vf[0], vf[1], vf[2], vf[3]);
    *reinterpret_cast<float4*>(memory + o + 4) = make_float4(vf[4], vf[5], vf[6], vf[7]);
    *reinterpret_cast<float4*>(memory_pos + o) = make_float4(vp[0], vp[1], vp[2], vp[3]);
    *reinterpret_cast<float4*>(memory_pos + o + 4) = make_float4(vp[4], vp[5], vp[6], vp[7]);
  }
}

// Object pointers: token (i * C/64 + c) of batch item b = ptrs[i][b, 64 c : 64 c + 64]; its position = obj_pos[i].
// One thread per output element.
template <typename T>
__global__ void bank_pointer_kernel(const __grid_constant__ BankArgs a, const float* __restrict__ obj_pos, float* __restrict__ memory,
                                    float* __restrict__ memory_pos, long long row0, int n_ptrs, int B, int C) {
  const int per = C / kMd;
  const long long total = (long long)n_ptrs * per * B * kMd;
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int ch = (int)(e % kMd);
  const int b = (int)((e / kMd) % B);
  const int tokn = (int)(e / ((long long)kMd * B));
  const int i = tokn / per, c = tokn % per;
  const long long o = ((row0 + tokn) * B + b) * kMd + ch;
  memory[o] = ldf<T>(a.ptrs[i], (long long)b * C + c * kMd + ch);
  memory_pos[o] = obj_pos ? obj_pos[(long long)i * kMd + ch] : 0.f;
}

// ---- PACKED variant: the bank is written directly in the layout the fused stack's kernels read:
// batch-first rows b * M + m, bf16, memk = feat + pos + tpos (the key source, `memory + memory_pos` of
// memory_attention.py:75-76) and memv = feat (the value source) -- no fp32 [M, B, 64] tensors, no re-pack pass.
__device__ __forceinline__ uint4 pack8_bf16(const float* v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
  __nv_bfloat162 c = __floats2bfloat162_rn(v[4], v[5]), d = __floats2bfloat162_rn(v[6], v[7]);
  return make_uint4(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b), *reinterpret_cast<uint32_t*>(&c),
                    *reinterpret_cast<uint32_t*>(&d));
}

template <typename T>
__global__ void __launch_bounds__(256)
bank_spatial_packed_kernel(const __grid_constant__ BankArgs a, __nv_bfloat16* __restrict__ memk, __nv_bfloat16* __restrict__ memv,
                           int B, int HW, long long M) {
  __shared__ float tf[kMd][33], tp[kMd][33];
  const int s = blockIdx.z, b = blockIdx.y, t0 = blockIdx.x * 32;
  const int tok = threadIdx.x & 31, c0 = threadIdx.x >> 5;
  const void* f = a.feats[s];
  const void* p = a.pos[s];
  const long long base = (long long)b * kMd * HW;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int ch = c0 + 8 * k;
    const bool ok = t0 + tok < HW;
    tf[ch][tok] = ok ? ldf<T>(f, base + (long long)ch * HW + t0 + tok) : 0.f;
    tp[ch][tok] = ok ? ldf<T>(p, base + (long long)ch * HW + t0 + tok) : 0.f;
  }
  __syncthreads();
  const int ot = threadIdx.x >> 3, och = (threadIdx.x & 7) * 8;
  if (t0 + ot < HW) {
    const float* tpos = a.tpos[s];
    float vf[8], vk[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { vf[i] = tf[och + i][ot]; vk[i] = vf[i] + (tp[och + i][ot] + tpos[och + i]); }   // same association as the unpacked path
    const long long o = ((long long)b * M + (long long)s * HW + t0 + ot) * kMd + och;
    *reinterpret_cast<uint4*>(memv + o) = pack8_bf16(vf);
    *reinterpret_cast<uint4*>(memk + o) = pack8_bf16(vk);
  }
}

template <typename T>
__global__ void bank_pointer_packed_kernel(const __grid_constant__ BankArgs a, const float* __restrict__ obj_pos,
                                           __nv_bfloat16* __restrict__ memk, __nv_bfloat16* __restrict__ memv, long long row0,
                                           int n_ptrs, int B, int C, long long M) {
  const int per = C / kMd;
  const long long total = (long long)n_ptrs * per * B * kMd;
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int ch = (int)(e % kMd);
  const int tokn = (int)((e / kMd) % ((long long)n_ptrs * per));
  const int b = (int)(e / ((long long)kMd * n_ptrs * per));
  const int i = tokn / per, c = tokn % per;
  const long long o = ((long long)b * M + row0 + tokn) * kMd + ch;
  const float v = ldf<T>(a.ptrs[i], (long long)b * C + c * kMd + ch);
  memv[o] = __float2bfloat16(v);
  memk[o] = __float2bfloat16(v + (obj_pos ? obj_pos[(long long)i * kMd + ch] : 0.f));
}

}  // namespace

extern "C" {

// Same inputs as sam2b200_bank_gather; outputs memk, memv: bf16 [B, M, 64] (M = n_slots * HW + n_ptrs * C / 64), the
// packed key / value sources of the fused MemoryAttention stack.
int sam2b200_bank_gather_packed(const void* const* feats, const void* const* pos, const float* const* tpos, int n_slots,
                                int feat_dtype, const void* const* ptrs, int n_ptrs, int ptr_dtype, const float* obj_pos,
                                void* memk, void* memv, int B, int HW, int mem_dim, int C, cudaStream_t stream) {
  if (n_slots < 0 || n_slots > kMaxSlots || n_ptrs < 0 || n_ptrs > kMaxPtrs || (n_slots == 0 && n_ptrs == 0) || !memk || !memv ||
      B <= 0 || B > 65535 || HW <= 0 || mem_dim != kMd || C <= 0 || (C % kMd) || (feat_dtype & ~1) || (ptr_dtype & ~1) ||
      (n_slots > 0 && (!feats || !pos || !tpos)) || (n_ptrs > 0 && !ptrs) ||
      (reinterpret_cast<uintptr_t>(memk) & 15) || (reinterpret_cast<uintptr_t>(memv) & 15))
    return sam2b200::fail(SAM2B200_ERR_INVALID, "bank_gather_packed: bad arguments (mem_dim must be 64, <= 40 frames, <= 64 pointers)");
  BankArgs a{};
  for (int s = 0; s < n_slots; ++s) {
    if (!feats[s] || !pos[s] || !tpos[s]) return sam2b200::fail(SAM2B200_ERR_INVALID, "bank_gather_packed: null frame pointer");
    a.feats[s] = feats[s]; a.pos[s] = pos[s]; a.tpos[s] = tpos[s];
  }
  for (int i = 0; i < n_ptrs; ++i) {
    if (!ptrs[i]) return sam2b200::fail(SAM2B200_ERR_INVALID, "bank_gather_packed: null object pointer");
    a.ptrs[i] = ptrs[i];
  }
  const long long M = (long long)n_slots * HW + (long long)n_ptrs * (C / kMd);
  __nv_bfloat16* mk = static_cast<__nv_bfloat16*>(memk);
  __nv_bfloat16* mv = static_cast<__nv_bfloat16*>(memv);
  int launches = 0;
  if (n_slots > 0) {
    dim3 grid((HW + 31) / 32, B, n_slots);
    if (feat_dtype == 0) bank_spatial_packed_kernel<float><<<grid, 256, 0, stream>>>(a, mk, mv, B, HW, M);
    else bank_spatial_packed_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(a, mk, mv, B, HW, M);
    ++launches;
  }
  if (n_ptrs > 0) {
    const long long total = (long long)n_ptrs * (C / kMd) * B * kMd;
    const unsigned blocks = (unsigned)((total + 255) / 256);
    const long long row0 = (long long)n_slots * HW;
    if (ptr_dtype == 0) bank_pointer_packed_kernel<float><<<blocks, 256, 0, stream>>>(a, obj_pos, mk, mv, row0, n_ptrs, B, C, M);
    else bank_pointer_packed_kernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>(a, obj_pos, mk, mv, row0, n_ptrs, B, C, M);
    ++launches;
  }
  return sam2b200::check_launch("bank_gather_packed", launches);
}


// feats / pos / tpos / ptrs: HOST arrays of device pointers.  dtype 0 = fp32, 1 = bf16.  memory, memory_pos:
// [n_slots * HW + n_ptrs * C / 64, B, 64] fp32, fully written.  obj_pos: [n_ptrs, 64] fp32 or NULL (zeros).
int sam2b200_bank_gather(const void* const* feats, const void* const* pos, const float* const* tpos, int n_slots,
                         int feat_dtype, const void* const* ptrs, int n_ptrs, int ptr_dtype, const float* obj_pos,
                         float* memory, float* memory_pos, int B, int HW, int mem_dim, int C, cudaStream_t stream) {
  if (n_slots < 0 || n_slots > kMaxSlots || n_ptrs < 0 || n_ptrs > kMaxPtrs || (n_slots == 0 && n_ptrs == 0) || !memory ||
      !memory_pos || B <= 0 || B > 65535 || HW <= 0 || mem_dim != kMd || C <= 0 || (C % kMd) || (feat_dtype & ~1) || (ptr_dtype & ~1) ||
      (n_slots > 0 && (!feats || !pos || !tpos)) || (n_ptrs > 0 && !ptrs) ||
      (reinterpret_cast<uintptr_t>(memory) & 15) || (reinterpret_cast<uintptr_t>(memory_pos) & 15))
    return sam2b200::fail(SAM2B200_ERR_INVALID, "bank_gather: bad arguments (mem_dim must be 64, <= 40 frames, <= 64 pointers)");
  BankArgs a{};
  for (int s = 0; s < n_slots; ++s) {
    if (!feats[s] || !pos[s] || !tpos[s]) return sam2b200::fail(SAM2B200_ERR_INVALID, "bank_gather: null frame pointer");
    a.feats[s] = feats[s]; a.pos[s] = pos[s]; a.tpos[s] = tpos[s];
  }
  for (int i = 0; i < n_ptrs; ++i) {
    if (!ptrs[i]) return sam2b200::fail(SAM2B200_ERR_INVALID, "bank_gather: null object pointer");
    a.ptrs[i] = ptrs[i];
  }
  int launches = 0;
  if (n_slots > 0) {
    dim3 grid((HW + 31) / 32, B, n_slots);
    if (feat_dtype == 0) bank_spatial_kernel<float><<<grid, 256, 0, stream>>>(a, memory, memory_pos, B, HW);
    else bank_spatial_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(a, memory, memory_pos, B, HW);
    ++launches;
  }
  if (n_ptrs > 0) {
    const long long total = (long long)n_ptrs * (C / kMd) * B * kMd;
    const unsigned blocks = (unsigned)((total + 255) / 256);
    const long long row0 = (long long)n_slots * HW;
    if (ptr_dtype == 0) bank_pointer_kernel<float><<<blocks, 256, 0, stream>>>(a, obj_pos, memory, memory_pos, row0, n_ptrs, B, C);
    else bank_pointer_kernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>(a, obj_pos, memory, memory_pos, row0, n_ptrs, B, C);
    ++launches;
  }
  return sam2b200::check_launch("bank_gather", launches);
}

}  // extern "C"
