// C-ABI entry points of the attention path + the small HBM-bound helper kernels around the
// tcgen05 kernels of attn_kernels.cuh: axial RoPE (forward / conjugate), Delta = rowsum(dO o O),
// split-KV combine.
//
// Replaces (reference, paths relative to its root):
//   sam2_video/model/modeling/position_encoding.py:212-239  apply_rotary_enc
//   sam2_video/model/modeling/sam/transformer.py:296-306    k[:, :, :num_k_rope] = ...; SDPA
// and their autograd backward.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#include "abi_common.cuh"
#include "attn_kernels.cuh"
#include "attn_pair_kernel.cuh"
#include "attn_persist_kernels.cuh"
#include "attn_v64_kernels.cuh"
#include "attn_v64x2_kernels.cuh"
#include "attn_fwd_v64x2_kernel.cuh"
#include "attn_v64_persist_kernel.cuh"
#include "tma_desc.cuh"

namespace {

// ------------------------------------------------------------------ axial RoPE
// x: [B, L, 256] (fp32 or bf16); rows [0, n_rope) of every batch are rotated with the table row
// (row mod n_tokens); rows >= n_rope are copied (object-pointer keys, transformer.py:296).
// table: [n_tokens, 128] float2 (cos, sin) built on the host side exactly like the reference's
// compute_axial_cis (position_encoding.py:192-201).  inverse = conjugate rotation (backward).
// Rotation is done in fp32 registers (position_encoding.py:218-220 upcasts as well).
template <typename TIn, typename TOut>
__global__ void rope_kernel(const TIn* __restrict__ x, TOut* __restrict__ out,
                            const float2* __restrict__ table, long long rows_total, int L, int n_rope,
                            int n_tokens, int inverse) {
  // one thread = 8 consecutive features (4 complex pairs); 32 threads = one row
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long row = gid >> 5;
  const int seg = (int)(gid & 31);
  if (row >= rows_total) return;
  const int r = (int)(row % L);
  float v[8];
  if constexpr (sizeof(TIn) == 4) {
    const float4* src = reinterpret_cast<const float4*>(x + row * 256 + seg * 8);
    float4 a = src[0], b = src[1];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
    uint4 u = *reinterpret_cast<const uint4*>(x + row * 256 + seg * 8);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
  }
  if (r < n_rope) {
    const float2* t = table + (long long)(r % n_tokens) * 128 + seg * 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 cs = t[i];
      float sn = inverse ? -cs.y : cs.y;
      float re = v[2 * i] * cs.x - v[2 * i + 1] * sn;
      float im = v[2 * i] * sn + v[2 * i + 1] * cs.x;
      v[2 * i] = re; v[2 * i + 1] = im;
    }
  }
  if constexpr (sizeof(TOut) == 4) {
    float4* dst = reinterpret_cast<float4*>(out + row * 256 + seg * 8);
    dst[0] = make_float4(v[0], v[1], v[2], v[3]);
    dst[1] = make_float4(v[4], v[5], v[6], v[7]);
  } else {
    uint4 u;
    u.x = sm100::pack_bf16(v[0], v[1]); u.y = sm100::pack_bf16(v[2], v[3]);
    u.z = sm100::pack_bf16(v[4], v[5]); u.w = sm100::pack_bf16(v[6], v[7]);
    *reinterpret_cast<uint4*>(out + row * 256 + seg * 8) = u;
  }
}

// ------------------------------------------------------------------ Delta = rowsum(dO o O)
template <bool O_F32>
__global__ void delta_kernel(const void* __restrict__ o_, const __nv_bfloat16* __restrict__ d_o,
                             float* __restrict__ delta, long long rows) {
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float ov[8];
  if (O_F32) {
    const float4* p = reinterpret_cast<const float4*>(static_cast<const float*>(o_) + row * 256 + lane * 8);
    const float4 a = p[0], b = p[1];
    ov[0] = a.x; ov[1] = a.y; ov[2] = a.z; ov[3] = a.w; ov[4] = b.x; ov[5] = b.y; ov[6] = b.z; ov[7] = b.w;
  } else {
    uint4 a = *reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(o_) + row * 256 + lane * 8);
    const __nv_bfloat162* ha = reinterpret_cast<const __nv_bfloat162*>(&a);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(ha[i]); ov[2 * i] = f.x; ov[2 * i + 1] = f.y; }
  }
  uint4 b = *reinterpret_cast<const uint4*>(d_o + row * 256 + lane * 8);
  const __nv_bfloat162* hb = reinterpret_cast<const __nv_bfloat162*>(&b);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 fb = __bfloat1622float2(hb[i]);
    s += ov[2 * i] * fb.x + ov[2 * i + 1] * fb.y;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if (lane == 0) delta[row] = s;
}

// ------------------------------------------------------------------ split-KV combine
// part_acc: [nsplit, rows, 256] fp32 un-normalised, part_ml: [nsplit, rows, 2] = (m*c, l)
__global__ void combine_kernel(const float* __restrict__ part_acc, const float* __restrict__ part_ml,
                               __nv_bfloat16* __restrict__ out, float* __restrict__ out_f32,
                               float* __restrict__ lse2, long long rows, int nsplit) {
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float mmax = -INFINITY;
  for (int z = 0; z < nsplit; ++z) mmax = fmaxf(mmax, part_ml[((long long)z * rows + row) * 2]);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  float lsum = 0.f;
  for (int z = 0; z < nsplit; ++z) {
    const float* ml = part_ml + ((long long)z * rows + row) * 2;
    const float w = exp2f(ml[0] - mmax);
    lsum += w * ml[1];
    const float4* src = reinterpret_cast<const float4*>(part_acc + ((long long)z * rows + row) * 256 + lane * 8);
    float4 a = src[0], b = src[1];
    acc[0] += w * a.x; acc[1] += w * a.y; acc[2] += w * a.z; acc[3] += w * a.w;
    acc[4] += w * b.x; acc[5] += w * b.y; acc[6] += w * b.z; acc[7] += w * b.w;
  }
  const float inv = 1.0f / lsum;
  uint4 u;
  u.x = sm100::pack_bf16(acc[0] * inv, acc[1] * inv); u.y = sm100::pack_bf16(acc[2] * inv, acc[3] * inv);
  u.z = sm100::pack_bf16(acc[4] * inv, acc[5] * inv); u.w = sm100::pack_bf16(acc[6] * inv, acc[7] * inv);
  *reinterpret_cast<uint4*>(out + row * 256 + lane * 8) = u;
  if (out_f32 != nullptr) {
    float4* fp = reinterpret_cast<float4*>(out_f32 + row * 256 + lane * 8);
    fp[0] = make_float4(acc[0] * inv, acc[1] * inv, acc[2] * inv, acc[3] * inv);
    fp[1] = make_float4(acc[4] * inv, acc[5] * inv, acc[6] * inv, acc[7] * inv);
  }
  if (lane == 0) lse2[row] = mmax + log2f(lsum);
}

constexpr float kLog2e = 1.4426950408889634f;

// Debug aid: when set (sam2b200_debug_set_timeline), every attention kernel launch records a per-CTA phase
// timeline (8 x u64 per CTA) into this device buffer; consecutive launches append.
unsigned long long* g_timeline = nullptr;
size_t g_timeline_cap = 0, g_timeline_used = 0;
unsigned long long* timeline_slice(size_t ctas) {
  if (!g_timeline || g_timeline_used + ctas * 8 > g_timeline_cap) return nullptr;
  unsigned long long* p = g_timeline + g_timeline_used;
  g_timeline_used += ctas * 8;
  return p;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

template <typename K>
int set_smem(K kernel, size_t bytes) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) return sam2b200::fail(SAM2B200_ERR_CUDA, cudaGetErrorString(e));
  return 0;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace

int g_variant[4] = {0, 0, 0, 0};

// Width of the square grid an axial RoPE table of `period` rows belongs to (0 if period is not a square, or when the
// compact addressing is switched off with sam2b200_debug_set_variant(1, 1)): see GradOut::rope_w.
int axial_w(int period) {
  if (g_variant[1] == 1 || period <= 0) return 0;
  int w = (int)(sqrt((double)period) + 0.5);
  return (w * w == period) ? w : 0;
}

// Shared-memory bytes for the staged rotation table of a gradient epilogue (GradOut::rope_smem): 0 when there is no axial
// table, when the kernel's shared memory has no room for it (w > 32 next to the raw-memory kernels' 200 KB), or when
// sam2b200_debug_set_variant(3, 1) asks for the global-memory path.
int rope_stage_bytes(const attn::GradOut& g, size_t base_smem, bool persistent) {
  if (!g.rope_table || g.rope_w <= 0 || g.rope_period != g.rope_w * g.rope_w || g_variant[3] == 1) return 0;
  const int bytes = attn::rope_smem_bytes(g.rope_w) + (persistent ? attn::rope_y_bytes(g.rope_w) : 0);
  return (base_smem + (size_t)bytes <= 232448) ? bytes : 0;      // 227 KB per CTA
}

extern "C" {

int sam2b200_debug_set_variant(int key, int value) {
  if (key < 0 || key >= 4) return -1;
  const int old = g_variant[key];
  g_variant[key] = value;
  return old;
}

// Debug aid (not part of the training path): buf = device buffer of n_u64 u64 (or NULL to switch off).
// Each subsequent attention kernel launch appends grid-size x 8 u64 {smid, t_entry, t_setup, t_operand, t_first_scores,
// t_loop_done, t_end, tiles} in %globaltimer ns.  Returns the number of u64 used so far.
long long sam2b200_debug_set_timeline(void* buf, long long n_u64) {
  long long used = (long long)g_timeline_used;
  g_timeline = static_cast<unsigned long long*>(buf);
  g_timeline_cap = buf ? (size_t)n_u64 : 0;
  g_timeline_used = 0;
  return used;
}

// in_dtype / out_dtype: 0 = fp32, 1 = bf16
int sam2b200_rope_apply(const void* x, int in_dtype, void* out, int out_dtype, const float* table, int B,
                        int L, int n_rope, int n_tokens, int inverse, cudaStream_t stream) {
  if (!x || !out || !table || B <= 0 || L <= 0 || n_rope < 0 || n_rope > L || n_tokens <= 0 ||
      (n_rope % n_tokens) != 0 || !aligned16(x) || !aligned16(out))
    return sam2b200::fail(SAM2B200_ERR_INVALID, "rope_apply: bad arguments (n_rope must be a multiple of n_tokens)");
  const long long rows = (long long)B * L;
  const int threads = 256;
  const long long blocks = (rows * 32 + threads - 1) / threads;
  const float2* t = reinterpret_cast<const float2*>(table);
  if (in_dtype == 0 && out_dtype == 1)
    rope_kernel<float, __nv_bfloat16><<<(unsigned)blocks, threads, 0, stream>>>((const float*)x, (__nv_bfloat16*)out, t, rows, L, n_rope, n_tokens, inverse);
  else if (in_dtype == 1 && out_dtype == 1)
    rope_kernel<__nv_bfloat16, __nv_bfloat16><<<(unsigned)blocks, threads, 0, stream>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)out, t, rows, L, n_rope, n_tokens, inverse);
  else if (in_dtype == 0 && out_dtype == 0)
    rope_kernel<float, float><<<(unsigned)blocks, threads, 0, stream>>>((const float*)x, (float*)out, t, rows, L, n_rope, n_tokens, inverse);
  else if (in_dtype == 1 && out_dtype == 0)
    rope_kernel<__nv_bfloat16, float><<<(unsigned)blocks, threads, 0, stream>>>((const __nv_bfloat16*)x, (float*)out, t, rows, L, n_rope, n_tokens, inverse);
  else
    return sam2b200::fail(SAM2B200_ERR_INVALID, "rope_apply: dtype must be 0 (fp32) or 1 (bf16)");
  return sam2b200::check_launch("rope_apply");
}

// Number of KV splits that fills the 148 SMs for small grids (flash-decoding style); 1 otherwise.
int sam2b200_attn_default_nsplit(int B, int N, int M) {
  const int ctas = B * ((N + attn::kBlockM - 1) / attn::kBlockM);
  const int tiles = (M + attn::kBlockN - 1) / attn::kBlockN;
  if (ctas >= 148 || tiles < 8) return 1;
  int ns = 148 / ctas;
  if (ns > tiles / 4) ns = tiles / 4;   // keep >= 4 tiles per split
  if (ns > 16) ns = 16;
  return ns < 1 ? 1 : ns;
}

size_t sam2b200_attn_fwd_workspace_bytes(int B, int N, int M, int nsplit) {
  (void)M;
  if (nsplit <= 1) return 0;
  return (size_t)nsplit * B * N * (256 + 2) * sizeof(float);
}

// q: [B, N, 256], k, v: [B, M, 256] bf16 (q, k already rotated); out: [B, N, 256] bf16;
// lse2: [B, N] fp32 = log2(sum_j exp(scale * q.k_j)).
// drop_p > 0 with a device seed: attention-probability dropout (transformer.py:304-306) -- the probabilities that
// multiply V are masked and scaled by 1 / (1 - p); the mask is a function of (*drop_seed, drop_site, b, query, key)
// that sam2b200_attn_bwd_ex regenerates (csrc/dropout.cuh).
int sam2b200_attn_fwd_ex(const void* q, const void* k, const void* v, void* out, float* out_f32, float* lse2,
                         void* workspace, size_t workspace_bytes, int B, int N, int M, float scale, int nsplit,
                         float drop_p, const unsigned long long* drop_seed, unsigned drop_site, cudaStream_t stream) {
  if (!q || !k || !v || !out || !lse2 || B <= 0 || N <= 0 || M <= 0 || B > 65535 || !aligned16(q) ||
      !aligned16(k) || !aligned16(v) || !aligned16(out) || drop_p < 0.f || drop_p >= 1.f ||
      (drop_p > 0.f && drop_seed && (long long)B * N * M >= (1LL << 32)))
    return sam2b200::fail(SAM2B200_ERR_INVALID, "attn_fwd: bad arguments (with dropout B*N*M must be < 2^32)");
  const int total_tiles = (M + attn::kBlockN - 1) / attn::kBlockN;
  if (nsplit < 1) nsplit = 1;
  if (nsplit > total_tiles) nsplit = total_tiles;
  int tiles_per_split = (total_tiles + nsplit - 1) / nsplit;
  nsplit = (total_tiles + tiles_per_split - 1) / tiles_per_split;   // no empty split
  if (nsplit > 1 && (workspace == nullptr || workspace_bytes < sam2b200_attn_fwd_workspace_bytes(B, N, M, nsplit)))
    return sam2b200::fail(SAM2B200_ERR_INVALID, "attn_fwd: workspace too small for the requested split");
  if (out_f32 && !aligned16(out_f32)) return sam2b200::fail(SAM2B200_ERR_INVALID, "attn_fwd: out_f32 must be 16-byte aligned");
  CUtensorMap map_k, map_v, map_q, map_o, map_o32;
  int rc;
  if ((rc = sam2b200::make_rows256_map(&map_k, k, B, M, attn::kBlockN))) return rc;
  if ((rc = sam2b200::make_rows256_map(&map_v, v, B, M, attn::kBlockN))) return rc;
  if ((rc = sam2b200::make_rows256_map(&map_q, q, B, N, attn::kBlockM))) return rc;
  if ((rc = sam2b200::make_out_map(&map_o, out, 1, B, N, 256, 32))) return rc;
  if (out_f32) { if ((rc = sam2b200::make_out_map(&map_o32, out_f32, 0, B, N, 256, 32))) return rc; }
  else map_o32 = map_o;
  attn::TwoGemmParams p{};
  p.La = N; p.Lx = M; p.scale_log2 = scale * kLog2e;
  p.has_out_f32 = out_f32 != nullptr; p.lse2 = lse2; p.tiles_per_split = tiles_per_split;
  p.drop = sam2b200::make_dropout(drop_seed, drop_site, drop_p);
  if (nsplit > 1) {
    p.part_acc = (float*)workspace;
    p.part_ml = p.part_acc + (size_t)nsplit * B * N * 256;
  }
  const size_t smem = sizeof(attn::SharedStorage) + 1024;
  dim3 grid((N + attn::kBlockM - 1) / attn::kBlockM, B, nsplit);
  p.dbg = timeline_slice((size_t)grid.x * grid.y * grid.z);
  if (p.drop.seed != nullptr) {
    if ((rc = set_smem(attn::two_gemm_kernel<attn::MODE_FWD, true>, smem))) return rc;
    attn::two_gemm_kernel<attn::MODE_FWD, true><<<grid, attn::kThreads, smem, stream>>>(map_k, map_v, map_q, map_o, map_o32, map_o, p);
  } else {
    if ((rc = set_smem(attn::two_gemm_kernel<attn::MODE_FWD, false>, smem))) return rc;
    attn::two_gemm_kernel<attn::MODE_FWD, false><<<grid, attn::kThreads, smem, stream>>>(map_k, map_v, map_q, map_o, map_o32, map_o, p);
  }
  if ((rc = sam2b200::check_launch("attn_fwd"))) return rc;
  if (nsplit > 1) {
    const long long rows = (long long)B * N;
    combine_kernel<<<(unsigned)((rows * 32 + 255) / 256), 256, 0, stream>>>(p.part_acc, p.part_ml, (__nv_bfloat16*)out, out_f32, lse2, rows, nsplit);
    if ((rc = sam2b200::check_launch("attn_fwd combine"))) return rc;
  }
  return SAM2B200_OK;
}

// Cross-attention forward on the RAW 64-d memory features (kv_in_dim = 64, sam2.1_hiera_t.yaml:41-50):
// out64 = softmax(scale q k^T) memv, [B, N, 64] bf16 (+ fp32 copy).  Because softmax rows sum to 1, the reference's
// softmax(.) (memv Wv^T + bv) equals out64 Wv^T + bv: the caller applies v_proj to the [B N, 64] result instead of to the
// [B M, 64] memory -- 4x fewer PV FLOPs, no [B, M, 256] value tensor.
// With attention-probability dropout (drop_p > 0, device seed) the rows of the dropped matrix no longer sum to 1:
// out = out64 Wv^T + rowsum_drop bv, rowsum_drop [B, N] = row sums of the kept, re-scaled probabilities (required then).
// q: [B, N, 256], k: [B, M, 256] (rotated), memv: [B, M, 64] bf16.
int sam2b200_attn_fwd_v64(const void* q, const void* k, const void* memv, void* out64, float* out64_f32, float* lse2,
                          float* rowsum_drop, int B, int N, int M, float scale, float drop_p,
                          const unsigned long long* drop_seed, unsigned drop_site, cudaStream_t stream) {
  if (!q || !k || !memv || !out64 || !lse2 || B <= 0 || N <= 0 || M <= 0 || B > 65535 || !aligned16(q) || !aligned16(k) ||
      !aligned16(memv) || !aligned16(out64) || (out64_f32 && !aligned16(out64_f32)) || drop_p < 0.f || drop_p >= 1.f ||
      (drop_p > 0.f && drop_seed && (!rowsum_drop || (long long)B * N * M >= (1LL << 32))))
    return sam2b200::fail(SAM2B200_ERR_INVALID, "attn_fwd_v64: bad arguments");
  CUtensorMap map_k, map_v, map_q;
  int rc;
  if ((rc = sam2b200::make_rows256_map(&map_k, k, B, M, attn::kBlockN))) return rc;
  if ((rc = sam2b200::make_rows256_map(&map_v, memv, B, M, attn::kBlockN, 64))) return rc;
  if ((rc = sam2b200::make_rows256_map(&map_q, q, B, N, attn::kBlockM))) return rc;
  attn::TwoGemmParams p{};
  p.La = N; p.Lx = M; p.scale_log2 = scale * kLog2e;
  p.lse2 = lse2; p.tiles_per_split = (M + attn::kBlockN - 1) / attn::kBlockN;
  p.drop = sam2b200::make_dropout(drop_seed, drop_site, drop_p);
  p.out_small = out64; p.out_small_f32 = out64_f32; p.rowsum_drop = rowsum_drop;
  const size_t smem = sizeof(attn::SharedStorage) + 1024;
  dim3 grid((N + attn::kBlockM - 1) / attn::kBlockM, B, 1);
  if (g_variant[2] == 1) {   // experiment: two independent online-softmax streams per CTA (attn_fwd_v64x2_kernel.cuh); same speed, not the default
    const size_t smem2 = sizeof(attn::SharedStorageF2) + 1024;
    if (p.drop.seed != nullptr) {
      if ((rc = set_smem(attn::fwd_v64x2_kernel<true, false>, smem2))) return rc;
      attn::fwd_v64x2_kernel<true, false><<<grid, attn::kF2Threads, smem2, stream>>>(map_k, map_v, map_q, map_q, map_q, p);
    } else {
      if ((rc = set_smem(attn::fwd_v64x2_kernel<false, false>, smem2))) return rc;
      attn::fwd_v64x2_kernel<false, false><<<grid, attn::kF2Threads, smem2, stream>>>(map_k, map_v, map_q, map_q, map_q, p);
    }
    return sam2b200::check_launch("attn_fwd_v64");
  }
  if (p.drop.seed != nullptr) {
    if ((rc = set_smem(attn::two_gemm_kernel<attn::MODE_FWD, true, 64>, smem))) return rc;
    attn::two_gemm_kernel<attn::MODE_FWD, true, 64><<<grid, attn::kThreads, smem, stream>>>(map_k, map_v, map_q, map_q, map_q, map_q, p);
  } else {
    if ((rc = set_smem(attn::two_gemm_kernel<attn::MODE_FWD, false, 64>, smem))) return rc;
    attn::two_gemm_kernel<attn::MODE_FWD, false, 64><<<grid, attn::kThreads, smem, stream>>>(map_k, map_v, map_q, map_q, map_q, map_q, p);
  }
  return sam2b200::check_launch("attn_fwd_v64");
}

// ---- forward with the OUTPUT PROJECTION fused into the epilogue (transformer.py:308-309; two_gemm_kernel<.., PROJ = true>) ----
// proj_out [B, N, 256] bf16 = out . w^T + bias, w [256, 256] bf16 (K-major, as nn.Linear stores it), bias [256] fp32.
// out (bf16) and out_f32 (optional) are still written: the backward needs them (weight gradient of the projection, Delta).
// No split-KV: the caller uses this entry only when B * ceil(N / 128) CTAs fill the GPU.
int sam2b200_attn_fwd_proj(const void* q, const void* k, const void* v, void* out, float* out_f32, float* lse2, const void* w,
                           const float* bias, void* proj_out, int B, int N, int M, float scale, float drop_p,
                           const unsigned long long* drop_seed, unsigned drop_site, cudaStream_t stream) {
  if (!q || !k || !v || !out || !lse2 || !w || !bias || !proj_out || B <= 0 || N <= 0 || M <= 0 || B > 65535 || !aligned16(q) ||
      !aligned16(k) || !aligned16(v) || !aligned16(out) || !aligned16(w) || !aligned16(bias) || !aligned16(proj_out) ||
      (out_f32 && !aligned16(out_f32)) || drop_p < 0.f || drop_p >= 1.f ||
      (drop_p > 0.f && drop_seed && (long long)B * N * M >= (1LL << 32)))
    return sam2b200::fail(SAM2B200_ERR_INVALID, "attn_fwd_proj: bad arguments");
  CUtensorMap map_k, map_v, map_q, map_o, map_w, map_p;
  int rc;
  if ((rc = sam2b200::make_rows256_map(&map_k, k, B, M, attn::kBlockN))) return rc;
  if ((rc = sam2b200::make_rows256_map(&map_v, v, B, M, attn::kBlockN))) return rc;
  if ((rc = sam2b200::make_rows256_map(&map_q, q, B, N, attn::kBlockM))) return rc;
  if ((rc = sam2b200::make_out_map(&map_o, out, 1, B, N, 256, 32))) return rc;
  if ((rc = sam2b200::make_rows256_map(&map_w, w, 1, 256, attn::kBlockM))) return rc;
  if ((rc = sam2b200::make_out_map(&map_p, proj_out, 1, B, N, 256, 32))) return rc;
  attn::TwoGemmParams p{};
  p.La = N; p.Lx = M; p.scale_log2 = scale * kLog2e;
  p.has_out_f32 = 0; p.out_small_f32 = out_f32; p.lse2 = lse2;
  p.tiles_per_split = (M + attn::kBlockN - 1) / attn::kBlockN;
  p.drop = sam2b200::make_dropout(drop_seed, drop_site, drop_p);
  p.proj_bias = bias;
  const size_t smem = sizeof(attn::SharedStorage) + 1024;
  dim3 grid((N + attn::kBlockM - 1) / attn::kBlockM, B, 1);
  if (p.drop.seed != nullptr) {
    if ((rc = set_smem(attn::two_gemm_kernel<attn::MODE_FWD, true, attn::kD, true>, smem))) return rc;
    attn::two_gemm_kernel<attn::MODE_FWD, true, attn::kD, true><<<grid, attn::kThreads, smem, stream>>>(map_k, map_v, map_q, map_o, map_w, map_p, p);
  } else {
    if ((rc = set_smem(attn::two_gemm_kernel<attn::MODE_FWD, false, attn::kD, true>, smem))) return rc;
    attn::two_gemm_kernel<attn::MODE_FWD, false, attn::kD, true><<<grid, attn::kThreads, smem, stream>>>(map_k, map_v, map_q, map_o, map_w, map_p, p);
  }
  return sam2b200::check_launch("attn_fwd_proj");
}

// Raw-memory cross-attention forward with the folded projection out_proj(v_proj(.)) in the epilogue:
// proj_out [B, N, 256] bf16 = out64 . w^T + bias (+ rowsum_drop * rank1), w [256, 64] bf16 = Wo Wv, bias [256] fp32 = Wo bv + bo
// (without dropout) or bo (with dropout, then rank1 [256] fp32 = Wo bv).  Other arguments as sam2b200_attn_fwd_v64.
int sam2b200_attn_fwd_v64_proj(const void* q, const void* k, const void* memv, void* out64, float* out64_f32, float* lse2,
                               float* rowsum_drop, const void* w, const float* bias, const float* rank1, void* proj_out, int B, int N,
                               int M, float scale, float drop_p, const unsigned long long* drop_seed, unsigned drop_site,
                               cudaStream_t stream) {
  if (!q || !k || !memv || !out64 || !lse2 || !w || !bias || !proj_out || B <= 0 || N <= 0 || M <= 0 || B > 65535 || !aligned16(q) ||
      !aligned16(k) || !aligned16(memv) || !aligned16(out64) || (out64_f32 && !aligned16(out64_f32)) || !aligned16(w) || !aligned16(bias) ||
      (rank1 && !aligned16(rank1)) || !aligned16(proj_out) || drop_p < 0.f || drop_p >= 1.f ||
      (drop_p > 0.f && drop_seed && (!rowsum_drop || !rank1 || (long long)B * N * M >= (1LL << 32))))
    return sam2b200::fail(SAM2B200_ERR_INVALID, "attn_fwd_v64_proj: bad arguments");
  CUtensorMap map_k, map_v, map_q, map_w, map_p;
  int rc;
  if ((rc = sam2b200::make_rows256_map(&map_k, k, B, M, attn::kBlockN))) return rc;
  if ((rc = sam2b200::make_rows256_map(&map_v, memv, B, M, attn::kBlockN, 64))) return rc;
  if ((rc = sam2b200::make_rows256_map(&map_q, q, B, N, attn::kBlockM))) return rc;
  if ((rc = sam2b200::make_rows256_map(&map_w, w, 1, 256, attn::kBlockM, 64))) return rc;
  if ((rc = sam2b200::make_out_map(&map_p, proj_out, 1, B, N, 256, 32))) return rc;
  attn::TwoGemmParams p{};
  p.La = N; p.Lx = M; p.scale_log2 = scale * kLog2e;
  p.lse2 = lse2; p.tiles_per_split = (M + attn::kBlockN - 1) / attn::kBlockN;
  p.drop = sam2b200::make_dropout(drop_seed, drop_site, drop_p);
  p.out_small = out64; p.out_small_f32 = out64_f32; p.rowsum_drop = rowsum_drop;
  p.proj_bias = bias; p.proj_rank1 = (p.drop.seed != nullptr) ? rank1 : nullptr;
  const size_t smem = sizeof(attn::SharedStorage) + 1024;
  dim3 grid((N + attn::kBlockM - 1) / attn::kBlockM, B, 1);
  if (g_variant[2] == 1) {
    const size_t smem2 = sizeof(attn::SharedStorageF2) + 1024;
    if (p.drop.seed != nullptr) {
      if ((rc = set_smem(attn::fwd_v64x2_kernel<true, true>, smem2))) return rc;
      attn::fwd_v64x2_kernel<true, true><<<grid, attn::kF2Threads, smem2, stream>>>(map_k, map_v, map_q, map_w, map_p, p);
    } else {
      if ((rc = set_smem(attn::fwd_v64x2_kernel<false, true>, smem2))) return rc;
      attn::fwd_v64x2_kernel<false, true><<<grid, attn::kF2Threads, smem2, stream>>>(map_k, map_v, map_q, map_w, map_p, p);
    }
    return sam2b200::check_launch("attn_fwd_v64_proj");
  }
  if (p.drop.seed != nullptr) {
    if ((rc = set_smem(attn::two_gemm_kernel<attn::MODE_FWD, true, 64, true>, smem))) return rc;
    attn::two_gemm_kernel<attn::MODE_FWD, true, 64, true><<<grid, attn::kThreads, smem, stream>>>(map_k, map_v, map_q, map_q, map_w, map_p, p);
  } else {
    if ((rc = set_smem(attn::two_gemm_kernel<attn::MODE_FWD, false, 64, true>, smem))) return rc;
    attn::two_gemm_kernel<attn::MODE_FWD, false, 64, true><<<grid, attn::kThreads, smem, stream>>>(map_k, map_v, map_q, map_q, map_w, map_p, p);
  }
  return sam2b200::check_launch("attn_fwd_v64_proj");
}

int sam2b200_attn_fwd(const void* q, const void* k, const void* v, void* out, float* out_f32, float* lse2,
                      void* workspace, size_t workspace_bytes, int B, int N, int M, float scale, int nsplit,
                      cudaStream_t stream) {
  return sam2b200_attn_fwd_ex(q, k, v, out, out_f32, lse2, workspace, workspace_bytes, B, N, M, scale, nsplit, 0.f, nullptr, 0,
                              stream);
}

// Backward of out = softmax(scale q k^T) v.  q, k, v, dout bf16; lse2 from the forward; the forward's
// output either as bf16 (`out`) or, preferred, its fp32 copy (`out_f32`): Delta = rowsum(dO o O) then has
// no per-row rounding bias, which matters when dP - Delta cancels (smooth / highly correlated values).
// delta: [B, N] fp32 scratch.  dq: [B, N, ldq], dk: [B, M, ldk], dv: [B, M, ldv] (fp32 if grad_dtype == 0,
// bf16 if 1), fully written in their first 256 columns.  If rope_table != NULL the conjugate axial rotation
// is applied in the epilogue: to every row of dq and to rows [0, n_rope_k) of dk (table row = row % rope_period),
// i.e. dq / dk are gradients with respect to the UN-rotated projections.
// dbias_q / dbias_k / dbias_v (optional, fp32 [256] each): the column sums of dq / dk / dv over all B x rows are ADDED
// to them inside the gradient epilogues (fp32 atomics: summation order is not fixed) -- the bias gradients of the
// q / k / v projections (transformer.py:220-222) without another pass over the gradient tensors.
int sam2b200_attn_bwd_ex(const void* q, const void* k, const void* v, const void* out, const float* out_f32,
                         const void* dout, const float* lse2, float* delta, void* dq, void* dk, void* dv,
                         int grad_dtype, int ldq, int ldk, int ldv, const float* rope_table, int rope_period,
                         int n_rope_k, int B, int N, int M, float scale, float* dbias_q, float* dbias_k, float* dbias_v,
                         int parts, float drop_p, const unsigned long long* drop_seed, unsigned drop_site,
                         cudaStream_t stream) {
  // parts: bit mask of the kernels to launch (0 = all): 1 Delta = rowsum(dO o O), 2 dV, 4 dK, 8 dQ.  dK and dQ need Delta.
  // Lets the caller put the key-side kernels of the cross-attention (whose results only feed weight / memory-bank
  // gradients) on a second stream, off the critical path of the residual-stream gradient.
  if (parts == 0) parts = 15;
  if (!q || !k || !v || (!out && !out_f32) || !dout || !lse2 || !delta || ((parts & 8) && !dq) || ((parts & 4) && !dk) ||
      ((parts & 2) && !dv) || B <= 0 || N <= 0 || M <= 0 ||
      B > 65535 || !aligned16(q) || !aligned16(k) || !aligned16(v) || !aligned16(dout) ||
      !aligned16(dq) || !aligned16(dk) || !aligned16(dv) || (grad_dtype != 0 && grad_dtype != 1) ||
      ldq < 256 || ldk < 256 || ldv < 256 || (ldq % 8) || (ldk % 8) || (ldv % 8) ||
      (rope_table && (rope_period <= 0 || n_rope_k < 0 || n_rope_k > M)) || drop_p < 0.f || drop_p >= 1.f ||
      (drop_p > 0.f && drop_seed && (long long)B * N * M >= (1LL << 32)))
    return sam2b200::fail(SAM2B200_ERR_INVALID, "attn_bwd: bad arguments");
  const sam2b200::Dropout drop = sam2b200::make_dropout(drop_seed, drop_site, drop_p);
  const bool drop_on = drop.seed != nullptr;
  int rc;
  const long long rows = (long long)B * N;
  if (parts & 1) {
    if (out_f32 != nullptr)
      delta_kernel<true><<<(unsigned)((rows * 32 + 255) / 256), 256, 0, stream>>>(out_f32, (const __nv_bfloat16*)dout, delta, rows);
    else
      delta_kernel<false><<<(unsigned)((rows * 32 + 255) / 256), 256, 0, stream>>>(out, (const __nv_bfloat16*)dout, delta, rows);
    if ((rc = sam2b200::check_launch("attn_bwd delta"))) return rc;
  }

  CUtensorMap map_q64, map_k64, map_v64, map_do64, map_do128, map_v128, map_q128, map_k128, map_dq, map_dk, map_dv;
  if ((rc = sam2b200::make_rows256_map(&map_q128, q, B, N, attn::kBlockM))) return rc;
  if ((rc = sam2b200::make_rows256_map(&map_k128, k, B, M, attn::kBlockM))) return rc;
  if ((parts & 8) && (rc = sam2b200::make_out_map(&map_dq, dq, grad_dtype, B, N, ldq, 32))) return rc;
  if ((parts & 4) && (rc = sam2b200::make_out_map(&map_dk, dk, grad_dtype, B, M, ldk, 32))) return rc;
  if ((parts & 2) && (rc = sam2b200::make_out_map(&map_dv, dv, grad_dtype, B, M, ldv, 32))) return rc;
  if ((rc = sam2b200::make_rows256_map(&map_q64, q, B, N, attn::kBlockN))) return rc;
  if ((rc = sam2b200::make_rows256_map(&map_k64, k, B, M, attn::kBlockN))) return rc;
  if ((rc = sam2b200::make_rows256_map(&map_v64, v, B, M, attn::kBlockN))) return rc;
  if ((rc = sam2b200::make_rows256_map(&map_do64, dout, B, N, attn::kBlockN))) return rc;
  if ((rc = sam2b200::make_rows256_map(&map_do128, dout, B, N, attn::kBlockM))) return rc;
  if ((rc = sam2b200::make_rows256_map(&map_v128, v, B, M, attn::kBlockM))) return rc;
  const float2* table = reinterpret_cast<const float2*>(rope_table);

  // dV and dK together on CTA pairs (attn_pair_kernel.cuh): P^T is computed once and shared through distributed shared
  // memory -- 4 GEMM units per (key block, query tile) instead of the 5 of the separate kernels below.  The cluster
  // launch + its barriers cost ~3.5 us per pair, so it only pays for long query loops: measured on B200
  // (profiles/r1_pair_kernel_experiment.txt) -8 % at N = 4096, -3 % at N = 1024, +6 % at N = 576.
  // SAM2B200_PAIR_KERNEL=1 / SAM2B200_NO_PAIR_KERNEL=1 force it on / off.
  static const bool force_pair = getenv("SAM2B200_PAIR_KERNEL") != nullptr;
  static const bool no_pair = getenv("SAM2B200_NO_PAIR_KERNEL") != nullptr;
  const bool use_pair = !no_pair && (force_pair || N >= 1024);
  // Persistent key-side kernels (attn_persist_kernels.cuh) for the short query loops the pair kernel does not cover.
  // SAM2B200_NO_PERSIST=1 switches back to one CTA per (key block, object).
  static const bool no_persist = getenv("SAM2B200_NO_PERSIST") != nullptr;
  const bool use_persist = !no_persist && N < 1024;   // long loops: 2.7 % slower than one-shot at N = 4096 (the staging slot costs a ring stage)
  if ((parts & 6) == 6 && use_pair) {
    attn::PairParams p{};
    p.Lk = M; p.Lq = N; p.scale_log2 = scale * kLog2e; p.scale = scale; p.lse2 = lse2; p.delta = delta;
    p.gout_v = attn::GradOut{grad_dtype, dbias_v, nullptr, 0, 1, 0};
    p.gout_k = attn::GradOut{grad_dtype, dbias_k, table, table ? n_rope_k : 0, table ? rope_period : 1, table ? axial_w(rope_period) : 0};
    p.drop = drop;
    const size_t smem = sizeof(attn::PairShared) + 1024;
    dim3 grid(2 * ((M + attn::kBlockM - 1) / attn::kBlockM), B, 1);
    if (drop_on) {
      if ((rc = set_smem(attn::kv_pair_kernel<true>, smem))) return rc;
      attn::kv_pair_kernel<true><<<grid, attn::kThreads, smem, stream>>>(map_q64, map_do64, map_k128, map_v128, map_dv, map_dk, p);
    } else {
      if ((rc = set_smem(attn::kv_pair_kernel<false>, smem))) return rc;
      attn::kv_pair_kernel<false><<<grid, attn::kThreads, smem, stream>>>(map_q64, map_do64, map_k128, map_v128, map_dv, map_dk, p);
    }
    if ((rc = sam2b200::check_launch("attn_bwd dV+dK pair"))) return rc;
    parts &= ~6;
  }
  // dV = P^T dO: fixed K block, stream (Q, dO) tiles
  if (parts & 2) {
    attn::TwoGemmParams p{};
    p.La = M; p.Lx = N; p.scale_log2 = scale * kLog2e;
    p.lse2 = const_cast<float*>(lse2);
    p.gout = attn::GradOut{grad_dtype, dbias_v, nullptr, 0, 1, 0};
    p.drop = drop;
    p.tiles_per_split = (N + attn::kBlockN - 1) / attn::kBlockN;
    const size_t smem = sizeof(attn::SharedStorage) + 1024;
    dim3 grid((M + attn::kBlockM - 1) / attn::kBlockM, B, 1);
    p.n_atiles = (int)grid.x;
    p.n_items = (int)(grid.x * grid.y);
    if (use_persist && grad_dtype && p.n_items > num_sms()) {
      // resident CTAs walking (key block, object) items: attn_persist_kernels.cuh
      if (drop_on) {
        if ((rc = set_smem(attn::dv_persistent_kernel<true>, smem))) return rc;
        attn::dv_persistent_kernel<true><<<num_sms(), attn::kThreads, smem, stream>>>(map_q64, map_do64, map_k128, map_dv, p);
      } else {
        if ((rc = set_smem(attn::dv_persistent_kernel<false>, smem))) return rc;
        attn::dv_persistent_kernel<false><<<num_sms(), attn::kThreads, smem, stream>>>(map_q64, map_do64, map_k128, map_dv, p);
      }
    } else {
      p.dbg = timeline_slice((size_t)grid.x * grid.y);
      if (drop_on) {
        if ((rc = set_smem(attn::two_gemm_kernel<attn::MODE_DV, true>, smem))) return rc;
        attn::two_gemm_kernel<attn::MODE_DV, true><<<grid, attn::kThreads, smem, stream>>>(map_q64, map_do64, map_k128, map_dv, map_dv, map_dv, p);
      } else {
        if ((rc = set_smem(attn::two_gemm_kernel<attn::MODE_DV, false>, smem))) return rc;
        attn::two_gemm_kernel<attn::MODE_DV, false><<<grid, attn::kThreads, smem, stream>>>(map_q64, map_do64, map_k128, map_dv, map_dv, map_dv, p);
      }
    }
    if ((rc = sam2b200::check_launch("attn_bwd dV"))) return rc;
  }
  const size_t smem3 = sizeof(attn::SharedStorage3) + 1024;
  // dK = scale * dS^T Q: fixed (K in TMEM, V in SMEM), stream (Q, dO) tiles
  if (parts & 4) {
    attn::ThreeGemmParams p{};
    p.La = M; p.Lx = N; p.scale_log2 = scale * kLog2e; p.scale = scale;
    p.lse2 = lse2; p.delta = delta;
    p.gout = attn::GradOut{grad_dtype, dbias_k, table, table ? n_rope_k : 0, table ? rope_period : 1, table ? axial_w(rope_period) : 0};
    p.drop = drop;
    dim3 grid((M + attn::kBlockM - 1) / attn::kBlockM, B, 1);
    p.n_atiles = (int)grid.x;
    p.n_items = (int)(grid.x * grid.y);
    static const bool no_persist_dk = getenv("SAM2B200_NO_PERSIST_DK") != nullptr;
    if (use_persist && !no_persist_dk && grad_dtype && p.n_items > num_sms()) {
      p.gout.rope_smem = rope_stage_bytes(p.gout, smem3, true);
      const size_t sm = smem3 + p.gout.rope_smem;
      if (drop_on) {
        if ((rc = set_smem(attn::dk_persistent_kernel<true>, sm))) return rc;
        attn::dk_persistent_kernel<true><<<num_sms(), attn::kThreads, sm, stream>>>(map_v128, map_q64, map_do64, map_k128, map_dk, p);
      } else {
        if ((rc = set_smem(attn::dk_persistent_kernel<false>, sm))) return rc;
        attn::dk_persistent_kernel<false><<<num_sms(), attn::kThreads, sm, stream>>>(map_v128, map_q64, map_do64, map_k128, map_dk, p);
      }
    } else {
      p.dbg = timeline_slice((size_t)grid.x * grid.y);
      p.gout.rope_smem = rope_stage_bytes(p.gout, smem3, false);
      const size_t sm = smem3 + p.gout.rope_smem;
      if (drop_on) {
        if ((rc = set_smem(attn::three_gemm_kernel<attn::MODE_DK, true>, sm))) return rc;
        attn::three_gemm_kernel<attn::MODE_DK, true><<<grid, attn::kThreads, sm, stream>>>(map_v128, map_q64, map_do64, map_k128, map_dk, p);
      } else {
        if ((rc = set_smem(attn::three_gemm_kernel<attn::MODE_DK, false>, sm))) return rc;
        attn::three_gemm_kernel<attn::MODE_DK, false><<<grid, attn::kThreads, sm, stream>>>(map_v128, map_q64, map_do64, map_k128, map_dk, p);
      }
    }
    if ((rc = sam2b200::check_launch("attn_bwd dK"))) return rc;
  }
  // dQ = scale * dS K: fixed (Q in TMEM, dO in SMEM), stream (K, V) tiles
  if (parts & 8) {
    attn::ThreeGemmParams p{};
    p.La = N; p.Lx = M; p.scale_log2 = scale * kLog2e; p.scale = scale;
    p.lse2 = lse2; p.delta = delta;
    p.gout = attn::GradOut{grad_dtype, dbias_q, table, table ? N : 0, table ? rope_period : 1, table ? axial_w(rope_period) : 0};
    p.drop = drop;
    dim3 grid((N + attn::kBlockM - 1) / attn::kBlockM, B, 1);
    p.dbg = timeline_slice((size_t)grid.x * grid.y);
    p.gout.rope_smem = rope_stage_bytes(p.gout, smem3, false);
    const size_t sm = smem3 + p.gout.rope_smem;
    if (drop_on) {
      if ((rc = set_smem(attn::three_gemm_kernel<attn::MODE_DQ, true>, sm))) return rc;
      attn::three_gemm_kernel<attn::MODE_DQ, true><<<grid, attn::kThreads, sm, stream>>>(map_do128, map_k64, map_v64, map_q128, map_dq, p);
    } else {
      if ((rc = set_smem(attn::three_gemm_kernel<attn::MODE_DQ, false>, sm))) return rc;
      attn::three_gemm_kernel<attn::MODE_DQ, false><<<grid, attn::kThreads, sm, stream>>>(map_do128, map_k64, map_v64, map_q128, map_dq, p);
    }
    if ((rc = sam2b200::check_launch("attn_bwd dQ"))) return rc;
  }
  return SAM2B200_OK;
}

// Backward of the 64-d-memory cross-attention (sam2b200_attn_fwd_v64).  dout64 = dO Wv ([B, N, 64] bf16, the gradient
// w.r.t. out64), delta = rowsum(dout64 o out64) ([B, N] fp32, computed by the caller): dP = dout64 memv^T differs from
// dO V^T by a per-row constant, which dP - Delta cancels.  parts: 4 = dK, 8 = dQ (there is no dV: the gradient of the
// value projection is dO^T out64, a [256, 64] GEMM on the caller's side).  Other arguments as sam2b200_attn_bwd_ex.
// With dropout (same drop_p / seed / site as the forward): dp_bias [B, N] = dO . bv per query (added to dP before the
// mask) and delta = rowsum(dout64 o out64) + dp_bias * rowsum_drop.
int sam2b200_attn_bwd_v64(const void* q, const void* k, const void* memv, const void* dout64, const float* lse2,
                          const float* delta, void* dq, void* dk, int grad_dtype, int ldq, int ldk, const float* rope_table,
                          int rope_period, int n_rope_k, int B, int N, int M, float scale, float* dbias_q, float* dbias_k,
                          int parts, const float* dp_bias, float drop_p, const unsigned long long* drop_seed,
                          unsigned drop_site, cudaStream_t stream) {
  if (!q || !k || !memv || !dout64 || !lse2 || !delta || ((parts & 8) && !dq) || ((parts & 4) && !dk) || (parts & ~12) ||
      B <= 0 || N <= 0 || M <= 0 || B > 65535 || !aligned16(q) || !aligned16(k) || !aligned16(memv) || !aligned16(dout64) ||
      (rope_table && rope_period <= 0) || ldq < 256 || ldk < 256 || drop_p < 0.f || drop_p >= 1.f ||
      (drop_p > 0.f && drop_seed && (!dp_bias || (long long)B * N * M >= (1LL << 32))))
    return sam2b200::fail(SAM2B200_ERR_INVALID, "attn_bwd_v64: bad arguments");
  int rc;
  CUtensorMap map_q64, map_k64, map_q128, map_k128, map_m64, map_m128, map_d64, map_d128, map_dq, map_dk;
  if ((rc = sam2b200::make_rows256_map(&map_q128, q, B, N, attn::kBlockM))) return rc;
  if ((rc = sam2b200::make_rows256_map(&map_k128, k, B, M, attn::kBlockM))) return rc;
  if ((rc = sam2b200::make_rows256_map(&map_q64, q, B, N, attn::kBlockN))) return rc;
  if ((rc = sam2b200::make_rows256_map(&map_k64, k, B, M, attn::kBlockN))) return rc;
  if ((rc = sam2b200::make_rows256_map(&map_m64, memv, B, M, attn::kBlockN, 64))) return rc;
  if ((rc = sam2b200::make_rows256_map(&map_m128, memv, B, M, attn::kBlockM, 64))) return rc;
  if ((rc = sam2b200::make_rows256_map(&map_d64, dout64, B, N, attn::kBlockN, 64))) return rc;
  if ((rc = sam2b200::make_rows256_map(&map_d128, dout64, B, N, attn::kBlockM, 64))) return rc;
  if ((parts & 8) && (rc = sam2b200::make_out_map(&map_dq, dq, grad_dtype, B, N, ldq, 32))) return rc;
  if ((parts & 4) && (rc = sam2b200::make_out_map(&map_dk, dk, grad_dtype, B, M, ldk, 32))) return rc;
  const float2* table = reinterpret_cast<const float2*>(rope_table);
  const size_t smem3 = sizeof(attn::SharedStorage3) + 1024;
  const sam2b200::Dropout nodrop = sam2b200::make_dropout(drop_seed, drop_site, drop_p);   // (no dropout unless a seed is given)
  const bool drop_on = nodrop.seed != nullptr;
  const size_t smemv = sizeof(attn::SharedStorageV64) + 1024;
  const size_t smemvp = sizeof(attn::SharedStorageV64P) + 1024;
  if (parts & 4) {   // dK: A1 = K block (TMEM), A2 = memory block, X = Q tiles, Y = dout64 tiles
    attn::ThreeGemmParams p{};
    p.La = M; p.Lx = N; p.scale_log2 = scale * kLog2e; p.scale = scale; p.lse2 = lse2; p.delta = delta;
    p.gout = attn::GradOut{grad_dtype, dbias_k, table, table ? n_rope_k : 0, table ? rope_period : 1, table ? axial_w(rope_period) : 0};
    p.drop = nodrop; p.dp_bias = dp_bias;
    dim3 grid((M + attn::kBlockM - 1) / attn::kBlockM, B, 1);
    p.n_atiles = (int)grid.x;
    p.n_items = (int)(grid.x * grid.y);
    static const bool no_persist = getenv("SAM2B200_NO_PERSIST") != nullptr;
    static const bool single_buf = getenv("SAM2B200_V64_SINGLE_BUFFER") != nullptr;   // A/B: the generic single-buffered kernels
    static const bool persist_dk = getenv("SAM2B200_V64_PERSIST_DK") != nullptr;
    // A/B: two softmax groups on alternate tiles.  NOT the default: measured equal to the single-group kernels (the loop is
    // bound by the SS-mode MMAs' shared-memory reads, not by softmax latency -- profiles/r2_two_softmax_groups_experiment.txt)
    static const bool x2_env = getenv("SAM2B200_V64_X2") != nullptr;
    const bool no_x2 = !(x2_env || g_variant[0] == 1);
    const size_t smemx2 = sizeof(attn::SharedStorageV64x2) + 1024;
    if (!no_x2 && !single_buf && !persist_dk) {   // two softmax groups on alternate tiles (attn_v64x2_kernels.cuh)
      p.dbg = timeline_slice((size_t)grid.x * grid.y);
      if (drop_on) {
        if ((rc = set_smem(attn::three_gemm_v64x2_kernel<attn::MODE_DK, true>, smemx2))) return rc;
        attn::three_gemm_v64x2_kernel<attn::MODE_DK, true><<<grid, attn::kX2Threads, smemx2, stream>>>(map_m128, map_q64, map_d64, map_k128, map_dk, p);
      } else {
        if ((rc = set_smem(attn::three_gemm_v64x2_kernel<attn::MODE_DK>, smemx2))) return rc;
        attn::three_gemm_v64x2_kernel<attn::MODE_DK><<<grid, attn::kX2Threads, smemx2, stream>>>(map_m128, map_q64, map_d64, map_k128, map_dk, p);
      }
    } else if (!single_buf && !persist_dk && g_variant[0] != 2 && grad_dtype == 1 && N > 2 * attn::kBlockN &&
               (N <= 12 * attn::kBlockN || g_variant[0] == 3) && p.n_items > num_sms()) {
      // resident CTAs walking over (key block, object) items: the next item's operands land under this item's epilogue
      // (attn_v64_persist_kernel.cuh); sam2b200_debug_set_variant(0, 2) = one CTA per item.  Short query loops only
      // (N <= 768: dK 204 -> 189 us at cfg2; no gain at N = 1024, 5 % slower at N = 4096 -- profiles/r2_v64_persist_ab.txt)
      p.gout.rope_smem = rope_stage_bytes(p.gout, smemvp, true);
      const size_t sm = smemvp + p.gout.rope_smem;
      if (drop_on) {
        if ((rc = set_smem(attn::three_gemm_v64_persistent_kernel<attn::MODE_DK, true>, sm))) return rc;
        attn::three_gemm_v64_persistent_kernel<attn::MODE_DK, true><<<num_sms(), attn::kThreads, sm, stream>>>(map_m128, map_q64, map_d64, map_k128, map_dk, p);
      } else {
        if ((rc = set_smem(attn::three_gemm_v64_persistent_kernel<attn::MODE_DK>, sm))) return rc;
        attn::three_gemm_v64_persistent_kernel<attn::MODE_DK><<<num_sms(), attn::kThreads, sm, stream>>>(map_m128, map_q64, map_d64, map_k128, map_dk, p);
      }
    } else if (drop_on) {
      p.gout.rope_smem = rope_stage_bytes(p.gout, smemv, false);
      if ((rc = set_smem(attn::three_gemm_v64_kernel<attn::MODE_DK, true>, smemv + p.gout.rope_smem))) return rc;
      attn::three_gemm_v64_kernel<attn::MODE_DK, true><<<grid, attn::kThreads, smemv + p.gout.rope_smem, stream>>>(map_m128, map_q64, map_d64, map_k128, map_dk, p);
    } else if (!single_buf && !persist_dk) {   // double-buffered S / dP, both fixed operands in shared memory (attn_v64_kernels.cuh)
      p.dbg = timeline_slice((size_t)grid.x * grid.y);
      p.gout.rope_smem = rope_stage_bytes(p.gout, smemv, false);
      if ((rc = set_smem(attn::three_gemm_v64_kernel<attn::MODE_DK>, smemv + p.gout.rope_smem))) return rc;
      attn::three_gemm_v64_kernel<attn::MODE_DK><<<grid, attn::kThreads, smemv + p.gout.rope_smem, stream>>>(map_m128, map_q64, map_d64, map_k128, map_dk, p);
    } else if (!no_persist && N < 1024 && grad_dtype && p.n_items > num_sms()) {   // short query loops: resident CTAs (attn_persist_kernels.cuh)
      if ((rc = set_smem(attn::dk_persistent_kernel<false, 64>, smem3))) return rc;
      attn::dk_persistent_kernel<false, 64><<<num_sms(), attn::kThreads, smem3, stream>>>(map_m128, map_q64, map_d64, map_k128, map_dk, p);
    } else {
      if ((rc = set_smem(attn::three_gemm_kernel<attn::MODE_DK, false, 64>, smem3))) return rc;
      attn::three_gemm_kernel<attn::MODE_DK, false, 64><<<grid, attn::kThreads, smem3, stream>>>(map_m128, map_q64, map_d64, map_k128, map_dk, p);
    }
    if ((rc = sam2b200::check_launch("attn_bwd_v64 dK"))) return rc;
  }
  if (parts & 8) {   // dQ: A1 = Q block (TMEM), A2 = dout64 block, X = K tiles, Y = memory tiles
    attn::ThreeGemmParams p{};
    p.La = N; p.Lx = M; p.scale_log2 = scale * kLog2e; p.scale = scale; p.lse2 = lse2; p.delta = delta;
    p.gout = attn::GradOut{grad_dtype, dbias_q, table, table ? N : 0, table ? rope_period : 1, table ? axial_w(rope_period) : 0};
    p.drop = nodrop; p.dp_bias = dp_bias;
    dim3 grid((N + attn::kBlockM - 1) / attn::kBlockM, B, 1);
    static const bool single_buf = getenv("SAM2B200_V64_SINGLE_BUFFER") != nullptr;
    static const bool x2_env = getenv("SAM2B200_V64_X2") != nullptr;
    const bool no_x2 = !(x2_env || g_variant[0] == 1);
    const size_t smemx2 = sizeof(attn::SharedStorageV64x2) + 1024;
    if (!no_x2 && !single_buf) {
      p.dbg = timeline_slice((size_t)grid.x * grid.y);
      if (drop_on) {
        if ((rc = set_smem(attn::three_gemm_v64x2_kernel<attn::MODE_DQ, true>, smemx2))) return rc;
        attn::three_gemm_v64x2_kernel<attn::MODE_DQ, true><<<grid, attn::kX2Threads, smemx2, stream>>>(map_d128, map_k64, map_m64, map_q128, map_dq, p);
      } else {
        if ((rc = set_smem(attn::three_gemm_v64x2_kernel<attn::MODE_DQ>, smemx2))) return rc;
        attn::three_gemm_v64x2_kernel<attn::MODE_DQ><<<grid, attn::kX2Threads, smemx2, stream>>>(map_d128, map_k64, map_m64, map_q128, map_dq, p);
      }
    } else if (!single_buf && g_variant[0] == 3 && grad_dtype == 1 && M > 2 * attn::kBlockN && (int)(grid.x * grid.y) > num_sms()) {
      // query side: measured equal or slower than one CTA per item at every shape (long key loops) -- only on request (variant 3, tests)
      p.n_atiles = (int)grid.x;
      p.n_items = (int)(grid.x * grid.y);
      p.gout.rope_smem = rope_stage_bytes(p.gout, smemvp, true);
      const size_t sm = smemvp + p.gout.rope_smem;
      if (drop_on) {
        if ((rc = set_smem(attn::three_gemm_v64_persistent_kernel<attn::MODE_DQ, true>, sm))) return rc;
        attn::three_gemm_v64_persistent_kernel<attn::MODE_DQ, true><<<num_sms(), attn::kThreads, sm, stream>>>(map_d128, map_k64, map_m64, map_q128, map_dq, p);
      } else {
        if ((rc = set_smem(attn::three_gemm_v64_persistent_kernel<attn::MODE_DQ>, sm))) return rc;
        attn::three_gemm_v64_persistent_kernel<attn::MODE_DQ><<<num_sms(), attn::kThreads, sm, stream>>>(map_d128, map_k64, map_m64, map_q128, map_dq, p);
      }
    } else if (drop_on) {
      p.gout.rope_smem = rope_stage_bytes(p.gout, smemv, false);
      if ((rc = set_smem(attn::three_gemm_v64_kernel<attn::MODE_DQ, true>, smemv + p.gout.rope_smem))) return rc;
      attn::three_gemm_v64_kernel<attn::MODE_DQ, true><<<grid, attn::kThreads, smemv + p.gout.rope_smem, stream>>>(map_d128, map_k64, map_m64, map_q128, map_dq, p);
    } else if (!single_buf) {
      p.dbg = timeline_slice((size_t)grid.x * grid.y);
      p.gout.rope_smem = rope_stage_bytes(p.gout, smemv, false);
      if ((rc = set_smem(attn::three_gemm_v64_kernel<attn::MODE_DQ>, smemv + p.gout.rope_smem))) return rc;
      attn::three_gemm_v64_kernel<attn::MODE_DQ><<<grid, attn::kThreads, smemv + p.gout.rope_smem, stream>>>(map_d128, map_k64, map_m64, map_q128, map_dq, p);
    } else {
      if ((rc = set_smem(attn::three_gemm_kernel<attn::MODE_DQ, false, 64>, smem3))) return rc;
      attn::three_gemm_kernel<attn::MODE_DQ, false, 64><<<grid, attn::kThreads, smem3, stream>>>(map_d128, map_k64, map_m64, map_q128, map_dq, p);
    }
    if ((rc = sam2b200::check_launch("attn_bwd_v64 dQ"))) return rc;
  }
  return SAM2B200_OK;
}

int sam2b200_attn_bwd(const void* q, const void* k, const void* v, const void* out, const float* out_f32,
                      const void* dout, const float* lse2, float* delta, void* dq, void* dk, void* dv,
                      int grad_dtype, int ldq, int ldk, int ldv, const float* rope_table, int rope_period,
                      int n_rope_k, int B, int N, int M, float scale, cudaStream_t stream) {
  return sam2b200_attn_bwd_ex(q, k, v, out, out_f32, dout, lse2, delta, dq, dk, dv, grad_dtype, ldq, ldk, ldv, rope_table,
                              rope_period, n_rope_k, B, N, M, scale, nullptr, nullptr, nullptr, 0, 0.f, nullptr, 0, stream);
}

}  // extern "C"
