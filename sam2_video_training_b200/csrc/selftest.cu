// Stand-alone GPU self-test of libsam2b200.so (no torch): the C ABI against a double-precision CPU
// restatement written here.  Used on the GPU box for fast diagnosis before the pytest parity
// suite; structured probes isolate the PV GEMM (uniform P) and the QK GEMM (one-hot V).
//   build: see sam2_video_training_b200/build.py (target `selftest`);  run: ./sam2b200_selftest [filter]
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/sam2_b200.h"

#define CK(x)                                                                                   \
  do {                                                                                          \
    cudaError_t e_ = (x);                                                                       \
    if (e_ != cudaSuccess) {                                                                    \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);           \
      exit(3);                                                                                  \
    }                                                                                           \
  } while (0)

static uint64_t g_seed = 0x1234567ULL;
static float frand() {  // U(-1, 1)
  g_seed = g_seed * 6364136223846793005ULL + 1442695040888963407ULL;
  return ((g_seed >> 40) & 0xFFFFFF) / float(1 << 23) - 1.0f;
}
static float bf16_round(float x) { return __bfloat162float(__float2bfloat16(x)); }

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  explicit DevBuf(size_t n_) : n(n_) { CK(cudaMalloc(&p, (n ? n : 1) * sizeof(T))); }
  ~DevBuf() { cudaFree(p); }
  void up(const std::vector<T>& h) { CK(cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice)); }
  std::vector<T> down() const {
    std::vector<T> h(n);
    CK(cudaMemcpy(h.data(), p, n * sizeof(T), cudaMemcpyDeviceToHost));
    return h;
  }
};

static std::vector<__nv_bfloat16> to_bf16(const std::vector<float>& v) {
  std::vector<__nv_bfloat16> o(v.size());
  for (size_t i = 0; i < v.size(); ++i) o[i] = __float2bfloat16(v[i]);
  return o;
}

struct AttnCase {
  const char* name;
  int B, N, M, nsplit;
  int probe;  // 0 random, 1 k=0 (uniform P), 2 one-hot V
  bool bwd;
};

// CPU reference in double on the bf16-rounded inputs.
static void cpu_attn(const std::vector<float>& q, const std::vector<float>& k, const std::vector<float>& v,
                     const std::vector<float>& dout, int B, int N, int M, double scale, std::vector<double>& out,
                     std::vector<double>& lse2, std::vector<double>* dq, std::vector<double>* dk,
                     std::vector<double>* dv) {
  out.assign((size_t)B * N * 256, 0.0);
  lse2.assign((size_t)B * N, 0.0);
  if (dq) { dq->assign((size_t)B * N * 256, 0.0); dk->assign((size_t)B * M * 256, 0.0); dv->assign((size_t)B * M * 256, 0.0); }
  std::vector<double> s(M), p(M), dp(M);
  for (int b = 0; b < B; ++b)
    for (int i = 0; i < N; ++i) {
      const float* qi = &q[((size_t)b * N + i) * 256];
      double mx = -1e300;
      for (int j = 0; j < M; ++j) {
        const float* kj = &k[((size_t)b * M + j) * 256];
        double a = 0;
        for (int d = 0; d < 256; ++d) a += (double)qi[d] * kj[d];
        s[j] = a * scale;
        mx = fmax(mx, s[j]);
      }
      double l = 0;
      for (int j = 0; j < M; ++j) { p[j] = exp(s[j] - mx); l += p[j]; }
      for (int j = 0; j < M; ++j) p[j] /= l;
      lse2[(size_t)b * N + i] = (mx + log(l)) / log(2.0);
      double* oi = &out[((size_t)b * N + i) * 256];
      for (int j = 0; j < M; ++j) {
        const float* vj = &v[((size_t)b * M + j) * 256];
        for (int d = 0; d < 256; ++d) oi[d] += p[j] * vj[d];
      }
      if (!dq) continue;
      const float* doi = &dout[((size_t)b * N + i) * 256];
      double delta = 0;
      for (int d = 0; d < 256; ++d) delta += (double)doi[d] * oi[d];
      for (int j = 0; j < M; ++j) {
        const float* vj = &v[((size_t)b * M + j) * 256];
        double a = 0;
        for (int d = 0; d < 256; ++d) a += (double)doi[d] * vj[d];
        dp[j] = a;
      }
      double* dqi = &(*dq)[((size_t)b * N + i) * 256];
      for (int j = 0; j < M; ++j) {
        const double ds = p[j] * (dp[j] - delta) * scale;
        const float* kj = &k[((size_t)b * M + j) * 256];
        double* dkj = &(*dk)[((size_t)b * M + j) * 256];
        double* dvj = &(*dv)[((size_t)b * M + j) * 256];
        for (int d = 0; d < 256; ++d) {
          dqi[d] += ds * kj[d];
          dkj[d] += ds * qi[d];
          dvj[d] += p[j] * doi[d];
        }
      }
    }
}

static double rel_l2(const std::vector<double>& ref, const float* got, size_t n, double* max_abs) {
  double num = 0, den = 0, ma = 0;
  for (size_t i = 0; i < n; ++i) {
    double d = (double)got[i] - ref[i];
    num += d * d; den += ref[i] * ref[i];
    ma = fmax(ma, fabs(d));
  }
  if (max_abs) *max_abs = ma;
  return sqrt(num / fmax(den, 1e-300));
}

static int run_attn(const AttnCase& c) {
  const int B = c.B, N = c.N, M = c.M;
  const float scale = 1.0f / 16.0f;
  std::vector<float> q((size_t)B * N * 256), k((size_t)B * M * 256), v((size_t)B * M * 256), dout((size_t)B * N * 256);
  for (auto& x : q) x = bf16_round(frand() * 3.0f);
  for (auto& x : k) x = bf16_round(c.probe == 1 ? 0.f : frand() * 0.9f);
  for (auto& x : v) x = bf16_round(frand());
  for (auto& x : dout) x = bf16_round(frand());
  if (c.probe == 0 && M > 70)  // a late outlier key per batch: exercises the lazy-rescale path
    for (int b = 0; b < B; ++b)
      for (int d = 0; d < 256; ++d) k[((size_t)b * M + (M * 2 / 3)) * 256 + d] = bf16_round(q[((size_t)b * N) * 256 + d] * 0.5f);
  if (c.probe == 2)
    for (size_t j = 0; j < (size_t)B * M; ++j)
      for (int d = 0; d < 256; ++d) v[j * 256 + d] = ((int)(j % M) % 256 == d) ? 1.f : 0.f;

  DevBuf<__nv_bfloat16> dq_(q.size()), dk_(k.size()), dv_(v.size()), dout_(dout.size()), dout_o(q.size());
  dq_.up(to_bf16(q)); dk_.up(to_bf16(k)); dv_.up(to_bf16(v)); dout_.up(to_bf16(dout));
  DevBuf<float> lse((size_t)B * N), o32(q.size());
  int nsplit = c.nsplit > 0 ? c.nsplit : sam2b200_attn_default_nsplit(B, N, M);
  size_t wsb = sam2b200_attn_fwd_workspace_bytes(B, N, M, nsplit);
  DevBuf<char> ws(wsb);
  int rc = sam2b200_attn_fwd(dq_.p, dk_.p, dv_.p, dout_o.p, o32.p, lse.p, wsb ? ws.p : nullptr, wsb, B, N, M, scale, nsplit, 0);
  if (rc) { printf("[FAIL] %s: attn_fwd rc=%d %s\n", c.name, rc, sam2b200_last_error()); return 1; }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("[FAIL] %s: fwd sync: %s\n", c.name, cudaGetErrorString(e)); exit(4); }

  std::vector<double> ro, rl, rdq, rdk, rdv;
  cpu_attn(q, k, v, dout, B, N, M, scale, ro, rl, c.bwd ? &rdq : nullptr, &rdk, &rdv);
  auto ho = dout_o.down();
  std::vector<float> of(ho.size());
  for (size_t i = 0; i < ho.size(); ++i) of[i] = __bfloat162float(ho[i]);
  auto hl = lse.down();
  double ma, ml;
  double eo = rel_l2(ro, of.data(), of.size(), &ma);
  double el = rel_l2(rl, hl.data(), hl.size(), &ml);
  int fail = !(eo < 1e-2) || !(ml < 2e-3);
  printf("[%s] %-28s fwd  B=%d N=%d M=%d nsplit=%d  out rel_l2=%.3e max_abs=%.3e  lse2 max_abs=%.3e\n",
         fail ? "FAIL" : " ok ", c.name, B, N, M, nsplit, eo, ma, ml);
  if (fail) {
    for (int i = 0; i < 4; ++i)
      printf("   row %d: got %.4f %.4f %.4f %.4f | ref %.4f %.4f %.4f %.4f | lse got %.4f ref %.4f\n", i * 37 % N,
             of[(size_t)(i * 37 % N) * 256], of[(size_t)(i * 37 % N) * 256 + 1], of[(size_t)(i * 37 % N) * 256 + 64],
             of[(size_t)(i * 37 % N) * 256 + 255], ro[(size_t)(i * 37 % N) * 256], ro[(size_t)(i * 37 % N) * 256 + 1],
             ro[(size_t)(i * 37 % N) * 256 + 64], ro[(size_t)(i * 37 % N) * 256 + 255], hl[i * 37 % N], rl[i * 37 % N]);
  }
  if (!c.bwd) return fail;

  // backward uses the GPU forward's out/lse (as training does)
  DevBuf<float> gq(q.size()), gk(k.size()), gv(v.size()), delta((size_t)B * N);
  rc = sam2b200_attn_bwd(dq_.p, dk_.p, dv_.p, nullptr, o32.p, dout_.p, lse.p, delta.p, gq.p, gk.p, gv.p, 0, 256, 256, 256,
                         nullptr, 0, 0, B, N, M, scale, 0);
  if (rc) { printf("[FAIL] %s: attn_bwd rc=%d %s\n", c.name, rc, sam2b200_last_error()); return 1; }
  e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("[FAIL] %s: bwd sync: %s\n", c.name, cudaGetErrorString(e)); exit(4); }
  auto hq = gq.down(); auto hk = gk.down(); auto hv = gv.down();
  double eq = rel_l2(rdq, hq.data(), hq.size(), nullptr);
  double ek = rel_l2(rdk, hk.data(), hk.size(), nullptr);
  double ev = rel_l2(rdv, hv.data(), hv.size(), nullptr);
  int bfail = !(eq < 2e-2) || !(ek < 2e-2) || !(ev < 2e-2);
  printf("[%s] %-28s bwd  dq rel_l2=%.3e dk rel_l2=%.3e dv rel_l2=%.3e\n", bfail ? "FAIL" : " ok ", c.name, eq, ek, ev);
  if (bfail) {
    printf("   dq got %.4e %.4e ref %.4e %.4e | dk got %.4e %.4e ref %.4e %.4e | dv got %.4e %.4e ref %.4e %.4e\n", hq[0], hq[300],
           rdq[0], rdq[300], hk[0], hk[300], rdk[0], rdk[300], hv[0], hv[300], rdv[0], rdv[300]);
  }
  return fail | bfail;
}

// ------------------------------------------------------------------ mask loss
static int run_loss(int T, int C, int S, int l1) {
  const long long HW = (long long)S * S;
  std::vector<float> x((size_t)T * C * HW), iou((size_t)T * C);
  std::vector<uint8_t> t((size_t)T * C * HW);
  for (auto& a : x) a = frand() * 8.f;
  for (auto& a : iou) a = 0.5f + 0.5f * frand();
  for (int f = 0; f < T; ++f)
    for (int c = 0; c < C; ++c) {
      const bool empty = (c % 4 == 3) && C > 1;
      for (long long i = 0; i < HW; ++i) {
        long long yy = i / S, xx = i % S;
        double dx = (xx - S * 0.4) / (S * 0.25), dy = (yy - S * 0.55) / (S * 0.2);
        t[((size_t)f * C + c) * HW + i] = (!empty && dx * dx + dy * dy < 1.0 + 0.1 * c) ? 1 : 0;
      }
    }
  const double alpha = 0.25, gamma = 2.0, wm = 20, wd = 1, wi = 1;
  // CPU double
  double lm = 0, ld = 0, li = 0;
  std::vector<double> gx(x.size(), 0.0), gi(iou.size(), 0.0);
  for (int f = 0; f < T; ++f) {
    int nv = 0;
    std::vector<double> sums((size_t)C * 6, 0.0);
    for (int c = 0; c < C; ++c) {
      double* s = &sums[(size_t)c * 6];
      for (long long i = 0; i < HW; ++i) {
        size_t id = ((size_t)f * C + c) * HW + i;
        double xv = x[id], tv = t[id];
        double p = 1.0 / (1.0 + exp(-xv));
        double ce = fmax(xv, 0.0) - xv * tv + log1p(exp(-fabs(xv)));
        double q = tv ? 1 - p : p;
        s[0] += (tv ? alpha : 1 - alpha) * ce * pow(q, gamma);
        s[1] += p * tv; s[2] += p; s[3] += tv;
        s[4] += (xv > 0 && tv > 0); s[5] += (xv > 0 || tv > 0);
      }
      nv += s[3] > 0;
    }
    for (int c = 0; c < C; ++c) {
      double* s = &sums[(size_t)c * 6];
      if (!(s[3] > 0)) continue;
      lm += s[0] / HW / nv;
      ld += (1 - (2 * s[1] + 1) / (s[2] + s[3] + 1)) / nv;
      double act = s[4] / fmax(s[5], 1.0), d = iou[(size_t)f * C + c] - act;
      li += (l1 ? fabs(d) : d * d) / nv;
      gi[(size_t)f * C + c] = wi * (l1 ? (d > 0) - (d < 0) : 2 * d) / nv;
      for (long long i = 0; i < HW; ++i) {
        size_t id = ((size_t)f * C + c) * HW + i;
        double xv = x[id], tv = t[id];
        double p = 1.0 / (1.0 + exp(-xv));
        double ce = fmax(xv, 0.0) - xv * tv + log1p(exp(-fabs(xv)));
        double q = tv ? 1 - p : p;
        double at = tv ? alpha : 1 - alpha;
        double df = at * ((p - tv) * q * q + ce * 2 * q * (1 - 2 * tv) * p * (1 - p));
        double D = s[2] + s[3], Nn = 2 * s[1];
        double dd = -(2 * tv * (D + 1) - (Nn + 1)) / ((D + 1) * (D + 1)) * p * (1 - p);
        gx[id] = (wm * df / HW + wd * dd) / nv;
      }
    }
  }
  DevBuf<float> dx(x.size()), diou_in(iou.size()), gdx(x.size()), gdi(iou.size()), sums((size_t)T * C * 6), losses(4), gl(3);
  DevBuf<uint8_t> dt(t.size());
  DevBuf<int> nvd(T);
  dx.up(x); diou_in.up(iou); dt.up(t);
  gl.up(std::vector<float>{(float)wm, (float)wd, (float)wi});
  size_t wsb = sam2b200_mask_loss_workspace_bytes(T, C, HW);
  DevBuf<char> ws(wsb);
  std::vector<const float*> lp(T);
  std::vector<float*> gp(T);
  for (int f = 0; f < T; ++f) { lp[f] = dx.p + (size_t)f * C * HW; gp[f] = gdx.p + (size_t)f * C * HW; }
  int rc = sam2b200_mask_loss_fwd(lp.data(), dt.p, diou_in.p, nullptr, ws.p, sums.p, nvd.p, losses.p, T, C, HW, 0, (float)alpha,
                                  (float)gamma, 1.0f, l1, 1, 0);
  if (rc) { printf("[FAIL] loss fwd rc=%d %s\n", rc, sam2b200_last_error()); return 1; }
  rc = sam2b200_mask_loss_bwd(lp.data(), gp.data(), dt.p, diou_in.p, nullptr, sums.p, nvd.p, gl.p, gdi.p, T, C, HW, 0, (float)alpha,
                              (float)gamma, 1.0f, l1, 1, 0);
  if (rc) { printf("[FAIL] loss bwd rc=%d %s\n", rc, sam2b200_last_error()); return 1; }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("[FAIL] loss sync: %s\n", cudaGetErrorString(e)); exit(4); }
  auto hl = losses.down(); auto hg = gdx.down(); auto hi = gdi.down();
  double e0 = fabs(hl[0] - lm) / fmax(fabs(lm), 1e-12), e1 = fabs(hl[1] - ld) / fmax(fabs(ld), 1e-12), e2 = fabs(hl[2] - li) / fmax(fabs(li), 1e-12);
  double eg = rel_l2(gx, hg.data(), hg.size(), nullptr), egi = rel_l2(gi, hi.data(), hi.size(), nullptr);
  int fail = !(e0 < 1e-4 && e1 < 1e-4 && e2 < 1e-4 && eg < 1e-4 && egi < 1e-5);
  printf("[%s] loss T=%d C=%d S=%d %s: rel err mask %.2e dice %.2e iou %.2e | dlogits rel_l2 %.2e diou %.2e\n", fail ? "FAIL" : " ok ",
         T, C, S, l1 ? "L1" : "MSE", e0, e1, e2, eg, egi);
  return fail;
}

int main(int argc, char** argv) {
  const char* filter = argc > 1 ? argv[1] : "";
  if (sam2b200_check_device(0)) { printf("device check: %s\n", sam2b200_last_error()); return 2; }
  int fails = 0;
  if (!*filter || !strcmp(filter, "loss")) {
    fails += run_loss(2, 3, 16, 1);
    fails += run_loss(3, 5, 40, 0);
    fails += run_loss(2, 4, 192, 1);
    fails += run_loss(1, 2, 250, 0);  // HW not a multiple of 16: scalar path
  }
  const AttnCase cases[] = {
      {"probe_uniformP_1tile", 1, 128, 64, 1, 1, false},
      {"probe_onehotV_1tile", 1, 128, 64, 1, 2, false},
      {"rand_1tile", 1, 128, 64, 1, 0, true},
      {"probe_uniformP_4tiles", 1, 128, 256, 1, 1, false},
      {"probe_onehotV_4tiles", 1, 128, 256, 1, 2, false},
      {"rand_4tiles", 1, 128, 256, 1, 0, true},
      {"ragged", 2, 200, 300, 1, 0, true},
      {"ragged_split3", 2, 200, 300, 3, 0, false},
      {"cfg1_frame1", 1, 576, 580, 0, 0, true},
      {"cfg1_frame7", 1, 576, 4060, 0, 0, true},
      {"self_attn_1024", 2, 1024, 1024, 1, 0, true},
  };
  for (const auto& c : cases) {
    if (*filter && strcmp(filter, "attn") && !strstr(c.name, filter)) continue;
    fails += run_attn(c);
  }
  printf("selftest: %d failure(s)\n", fails);
  return fails ? 1 : 0;
}
