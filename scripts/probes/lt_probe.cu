// Probe (GPU box): which cuBLASLt fused epilogues does this cuBLAS build support for the shapes of the
// MemoryAttention projections / MLP, are they right, and how fast are they?
//   1. h  = relu(y W1^T + b1)          RELU_AUX_BIAS   (bf16 out + ReLU bit mask)
//   2. dh = (dm W2) o mask, db1 = colsum(dh)   DRELU_BGRAD
//   3. dW += dY^T X, db = colsum(dY)   BGRADB, fp32 out, beta = 1
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 lt_probe.cu -o lt_probe -lcublasLt
#include <cublasLt.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)
#define LT(x) do { cublasStatus_t s = (x); if (s != CUBLAS_STATUS_SUCCESS) { printf("cublasLt error %d at %s:%d\n", (int)s, __FILE__, __LINE__); return -1; } } while (0)

static cublasLtHandle_t lt;
static void* ws;
static size_t ws_bytes = 64 << 20;

// Row-major C[M,N] (+)= op(A)[M,K] op(B)[K,N]; cuBLAS sees C^T (N x M) = op(B)^T op(A)^T.
// epilogue: bias / aux refer to the cuBLAS view: vectors of length N (rows of C^T) for BIAS / DRELU_BGRAD / BGRADA,
// length M for BGRADB.
static int gemm(const void* A, const void* B, void* C, int M, int N, int K, int lda, int ldb, int ldc, bool tA, bool tB,
                cudaDataType ctype, float alpha, float beta, cublasLtEpilogue_t epi, void* bias, cudaDataType bias_type,
                void* aux, long long aux_ld, float* ms_out) {
  cublasLtMatmulDesc_t op;
  LT(cublasLtMatmulDescCreate(&op, CUBLAS_COMPUTE_32F, CUDA_R_32F));
  cublasOperation_t opa = tB ? CUBLAS_OP_T : CUBLAS_OP_N;   // first cuBLAS operand = B buffer
  cublasOperation_t opb = tA ? CUBLAS_OP_T : CUBLAS_OP_N;   // second = A buffer
  LT(cublasLtMatmulDescSetAttribute(op, CUBLASLT_MATMUL_DESC_TRANSA, &opa, sizeof(opa)));
  LT(cublasLtMatmulDescSetAttribute(op, CUBLASLT_MATMUL_DESC_TRANSB, &opb, sizeof(opb)));
  LT(cublasLtMatmulDescSetAttribute(op, CUBLASLT_MATMUL_DESC_EPILOGUE, &epi, sizeof(epi)));
  if (bias) {
    LT(cublasLtMatmulDescSetAttribute(op, CUBLASLT_MATMUL_DESC_BIAS_POINTER, &bias, sizeof(bias)));
    LT(cublasLtMatmulDescSetAttribute(op, CUBLASLT_MATMUL_DESC_BIAS_DATA_TYPE, &bias_type, sizeof(bias_type)));
  }
  if (aux) {
    LT(cublasLtMatmulDescSetAttribute(op, CUBLASLT_MATMUL_DESC_EPILOGUE_AUX_POINTER, &aux, sizeof(aux)));
    int64_t ld = aux_ld;
    LT(cublasLtMatmulDescSetAttribute(op, CUBLASLT_MATMUL_DESC_EPILOGUE_AUX_LD, &ld, sizeof(ld)));
  }
  cublasLtMatrixLayout_t la, lb, lc;
  // first operand: B buffer, col-major view: tB ? [K x N] ld=ldb : [N x K] ld=ldb
  LT(cublasLtMatrixLayoutCreate(&la, CUDA_R_16BF, tB ? K : N, tB ? N : K, ldb));
  LT(cublasLtMatrixLayoutCreate(&lb, CUDA_R_16BF, tA ? M : K, tA ? K : M, lda));
  LT(cublasLtMatrixLayoutCreate(&lc, ctype, N, M, ldc));
  cublasLtMatmulPreference_t pref;
  LT(cublasLtMatmulPreferenceCreate(&pref));
  LT(cublasLtMatmulPreferenceSetAttribute(pref, CUBLASLT_MATMUL_PREF_MAX_WORKSPACE_BYTES, &ws_bytes, sizeof(ws_bytes)));
  cublasLtMatmulHeuristicResult_t heur[4];
  int found = 0;
  cublasStatus_t st = cublasLtMatmulAlgoGetHeuristic(lt, op, la, lb, lc, lc, pref, 4, heur, &found);
  if (st != CUBLAS_STATUS_SUCCESS || found == 0) { printf("  no algorithm (status %d, found %d)\n", (int)st, found); return -2; }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = ms_out ? 20 : 1;
  for (int w = 0; w < (ms_out ? 3 : 0); ++w)
    LT(cublasLtMatmul(lt, op, &alpha, B, la, A, lb, &beta, C, lc, C, lc, &heur[0].algo, ws, ws_bytes, 0));
  cudaEventRecord(e0);
  for (int i = 0; i < iters; ++i)
    LT(cublasLtMatmul(lt, op, &alpha, B, la, A, lb, &beta, C, lc, C, lc, &heur[0].algo, ws, ws_bytes, 0));
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  if (ms_out) { cudaEventElapsedTime(ms_out, e0, e1); *ms_out /= iters; }
  cublasLtMatmulPreferenceDestroy(pref);
  cublasLtMatrixLayoutDestroy(la); cublasLtMatrixLayoutDestroy(lb); cublasLtMatrixLayoutDestroy(lc);
  cublasLtMatmulDescDestroy(op);
  return 0;
}

static std::vector<float> rnd(size_t n, float s, unsigned seed) {
  std::vector<float> v(n);
  unsigned x = seed * 2654435761u + 12345u;
  for (size_t i = 0; i < n; ++i) { x = x * 1664525u + 1013904223u; v[i] = s * (((x >> 8) & 0xffff) / 32768.0f - 1.0f); }
  return v;
}
static __nv_bfloat16* up_bf16(const std::vector<float>& h, std::vector<float>* rounded) {
  std::vector<__nv_bfloat16> b(h.size());
  if (rounded) rounded->resize(h.size());
  for (size_t i = 0; i < h.size(); ++i) { b[i] = __float2bfloat16(h[i]); if (rounded) (*rounded)[i] = __bfloat162float(b[i]); }
  __nv_bfloat16* d; CK(cudaMalloc(&d, h.size() * 2)); CK(cudaMemcpy(d, b.data(), h.size() * 2, cudaMemcpyHostToDevice));
  return d;
}
static std::vector<float> down_bf16(const __nv_bfloat16* d, size_t n) {
  std::vector<__nv_bfloat16> b(n); CK(cudaMemcpy(b.data(), d, n * 2, cudaMemcpyDeviceToHost));
  std::vector<float> f(n); for (size_t i = 0; i < n; ++i) f[i] = __bfloat162float(b[i]); return f;
}
static double rel(const std::vector<float>& a, const std::vector<float>& b) {
  double num = 0, den = 0; for (size_t i = 0; i < a.size(); ++i) { num += (a[i] - b[i]) * (double)(a[i] - b[i]); den += (double)b[i] * b[i]; }
  return sqrt(num / (den + 1e-30));
}

int main() {
  if (cublasLtCreate(&lt) != CUBLAS_STATUS_SUCCESS) { printf("cublasLtCreate failed\n"); return 1; }
  CK(cudaMalloc(&ws, ws_bytes));
  // ---------------- correctness at a small size ----------------
  {
    const int R = 384, D = 256, F = 2048;
    std::vector<float> y, w1, w2, dm, b1 = rnd(F, 0.5f, 5);
    __nv_bfloat16* dy = up_bf16(rnd((size_t)R * D, 1.f, 1), &y);
    __nv_bfloat16* dw1 = up_bf16(rnd((size_t)F * D, 0.08f, 2), &w1);     // [F, D]
    __nv_bfloat16* dw2 = up_bf16(rnd((size_t)D * F, 0.03f, 3), &w2);     // [D, F]
    __nv_bfloat16* ddm = up_bf16(rnd((size_t)R * D, 1.f, 4), &dm);       // [R, D]
    std::vector<float> b1r; __nv_bfloat16* db1 = up_bf16(b1, &b1r);
    __nv_bfloat16 *dh, *ddh; CK(cudaMalloc(&dh, (size_t)R * F * 2)); CK(cudaMalloc(&ddh, (size_t)R * F * 2));
    const long long aux_ld = 2048;   // bits per column of the cuBLAS D (m = F), multiple of 128
    void* mask; CK(cudaMalloc(&mask, (size_t)R * aux_ld / 8)); CK(cudaMemset(mask, 0, (size_t)R * aux_ld / 8));
    printf("[1] RELU_AUX_BIAS  h[R,F] = relu(y W1^T + b1): ");
    int rc = gemm(dy, dw1, dh, R, F, D, D, D, F, false, true, CUDA_R_16BF, 1.f, 0.f, CUBLASLT_EPILOGUE_RELU_AUX_BIAS, db1,
                  CUDA_R_16BF, mask, aux_ld, nullptr);
    std::vector<float> href((size_t)R * F);
    for (int r = 0; r < R; ++r) for (int f = 0; f < F; ++f) {
      double s = b1r[f]; for (int k = 0; k < D; ++k) s += (double)y[(size_t)r * D + k] * w1[(size_t)f * D + k];
      href[(size_t)r * F + f] = s > 0 ? (float)s : 0.f;
    }
    if (rc == 0) printf("rel err %.3e\n", rel(down_bf16(dh, (size_t)R * F), href));
    printf("[2] DRELU_BGRAD    dh[R,F] = (dm W2) o (h > 0), db1 = colsum(dh): ");
    float* dbg; CK(cudaMalloc(&dbg, F * 4)); CK(cudaMemset(dbg, 0, F * 4));
    rc = gemm(ddm, dw2, ddh, R, F, D, D, F, F, false, false, CUDA_R_16BF, 1.f, 0.f, CUBLASLT_EPILOGUE_DRELU_BGRAD, dbg, CUDA_R_32F,
              mask, aux_ld, nullptr);
    if (rc != 0) {
      printf("    retry with bf16 bias-gradient type: ");
      rc = gemm(ddm, dw2, ddh, R, F, D, D, F, F, false, false, CUDA_R_16BF, 1.f, 0.f, CUBLASLT_EPILOGUE_DRELU_BGRAD, dbg, CUDA_R_16BF,
                mask, aux_ld, nullptr);
      if (rc == 0) printf("(bf16 bias grad works) ");
    }
    if (rc == 0) {
      std::vector<float> dref((size_t)R * F), bref(F, 0.f);
      for (int r = 0; r < R; ++r) for (int f = 0; f < F; ++f) {
        double s = 0; for (int k = 0; k < D; ++k) s += (double)dm[(size_t)r * D + k] * w2[(size_t)k * F + f];
        float v = href[(size_t)r * F + f] > 0 ? (float)s : 0.f;
        dref[(size_t)r * F + f] = v; bref[f] += v;
      }
      std::vector<float> bg(F); CK(cudaMemcpy(bg.data(), dbg, F * 4, cudaMemcpyDeviceToHost));
      printf("dh rel err %.3e, db1 rel err %.3e\n", rel(down_bf16(ddh, (size_t)R * F), dref), rel(bg, bref));
    }
    printf("[3] BGRADB         dW[F,D] += dh^T y (fp32, beta = 1), db = colsum(dh): ");
    float* dW; CK(cudaMalloc(&dW, (size_t)F * D * 4));
    std::vector<float> ones((size_t)F * D, 1.f); CK(cudaMemcpy(dW, ones.data(), (size_t)F * D * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dbg, 0, F * 4));
    // row-major dW[M=F, N=D] = dh^T[F,R] y[R,D]: A = dh (transA), B = y
    rc = gemm(dh, dy, dW, F, D, R, F, D, D, true, false, CUDA_R_32F, 1.f, 1.f, CUBLASLT_EPILOGUE_BGRADB, dbg, CUDA_R_32F, nullptr, 0, nullptr);
    if (rc == 0) {
      std::vector<float> hh = down_bf16(dh, (size_t)R * F), wref((size_t)F * D), bref(F, 0.f), got((size_t)F * D), bg(F);
      for (int f = 0; f < F; ++f) {
        for (int k = 0; k < D; ++k) { double s = 1.0; for (int r = 0; r < R; ++r) s += (double)hh[(size_t)r * F + f] * y[(size_t)r * D + k]; wref[(size_t)f * D + k] = (float)s; }
        for (int r = 0; r < R; ++r) bref[f] += hh[(size_t)r * F + f];
      }
      CK(cudaMemcpy(got.data(), dW, got.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(bg.data(), dbg, F * 4, cudaMemcpyDeviceToHost));
      printf("dW rel err %.3e, db rel err %.3e\n", rel(got, wref), rel(bg, bref));
    }
  }
  // ---------------- timing at the cfg2 shapes (R = 56 x 576) ----------------
  {
    const int R = 32256, D = 256, F = 2048;
    __nv_bfloat16 *y, *w1, *w2, *dm, *h, *dh, *b1; float *dW, *dbg; void* mask;
    CK(cudaMalloc(&y, (size_t)R * D * 2)); CK(cudaMalloc(&w1, (size_t)F * D * 2)); CK(cudaMalloc(&w2, (size_t)D * F * 2));
    CK(cudaMalloc(&dm, (size_t)R * D * 2)); CK(cudaMalloc(&h, (size_t)R * F * 2)); CK(cudaMalloc(&dh, (size_t)R * F * 2));
    CK(cudaMalloc(&b1, F * 2)); CK(cudaMalloc(&dW, (size_t)F * D * 4)); CK(cudaMalloc(&dbg, F * 4)); CK(cudaMalloc(&mask, (size_t)R * F / 8));
    CK(cudaMemset(y, 0, (size_t)R * D * 2)); CK(cudaMemset(w1, 0, (size_t)F * D * 2)); CK(cudaMemset(w2, 0, (size_t)D * F * 2));
    CK(cudaMemset(dm, 0, (size_t)R * D * 2)); CK(cudaMemset(b1, 0, F * 2)); CK(cudaMemset(dW, 0, (size_t)F * D * 4)); CK(cudaMemset(mask, 0xff, (size_t)R * F / 8));
    float ms;
    if (gemm(y, w1, h, R, F, D, D, D, F, false, true, CUDA_R_16BF, 1.f, 0.f, CUBLASLT_EPILOGUE_RELU_BIAS, b1, CUDA_R_16BF, nullptr, 0, &ms) == 0)
      printf("time R=%d  RELU_BIAS          %.1f us\n", R, ms * 1e3);
    if (gemm(y, w1, h, R, F, D, D, D, F, false, true, CUDA_R_16BF, 1.f, 0.f, CUBLASLT_EPILOGUE_RELU_AUX_BIAS, b1, CUDA_R_16BF, mask, F, &ms) == 0)
      printf("time R=%d  RELU_AUX_BIAS      %.1f us\n", R, ms * 1e3);
    if (gemm(dm, w2, dh, R, F, D, D, F, F, false, false, CUDA_R_16BF, 1.f, 0.f, CUBLASLT_EPILOGUE_DEFAULT, nullptr, CUDA_R_32F, nullptr, 0, &ms) == 0)
      printf("time R=%d  dh plain           %.1f us\n", R, ms * 1e3);
    if (gemm(dm, w2, dh, R, F, D, D, F, F, false, false, CUDA_R_16BF, 1.f, 0.f, CUBLASLT_EPILOGUE_DRELU_BGRAD, dbg, CUDA_R_32F, mask, F, &ms) == 0)
      printf("time R=%d  dh DRELU_BGRAD     %.1f us\n", R, ms * 1e3);
    if (gemm(dm, w2, dh, R, F, D, D, F, F, false, false, CUDA_R_16BF, 1.f, 0.f, CUBLASLT_EPILOGUE_DRELU, nullptr, CUDA_R_32F, mask, F, &ms) == 0)
      printf("time R=%d  dh DRELU           %.1f us\n", R, ms * 1e3);
    if (gemm(dm, w2, dh, R, F, D, D, F, F, false, false, CUDA_R_16BF, 1.f, 0.f, CUBLASLT_EPILOGUE_DRELU_BGRAD, dbg, CUDA_R_16BF, mask, F, &ms) == 0)
      printf("time R=%d  dh DRELU_BGRAD(bf16) %.1f us\n", R, ms * 1e3);
    if (gemm(dh, y, dW, F, D, R, F, D, D, true, false, CUDA_R_32F, 1.f, 1.f, CUBLASLT_EPILOGUE_DEFAULT, nullptr, CUDA_R_32F, nullptr, 0, &ms) == 0)
      printf("time R=%d  dW1 plain beta=1   %.1f us\n", R, ms * 1e3);
    if (gemm(dh, y, dW, F, D, R, F, D, D, true, false, CUDA_R_32F, 1.f, 1.f, CUBLASLT_EPILOGUE_BGRADB, dbg, CUDA_R_32F, nullptr, 0, &ms) == 0)
      printf("time R=%d  dW1 BGRADB beta=1  %.1f us\n", R, ms * 1e3);
    // 256 x 256 weight gradient with a 256-wide dY (projection layers): dW[256,256] += dY^T X
    if (gemm(dm, y, dW, D, D, R, D, D, D, true, false, CUDA_R_32F, 1.f, 1.f, CUBLASLT_EPILOGUE_DEFAULT, nullptr, CUDA_R_32F, nullptr, 0, &ms) == 0)
      printf("time R=%d  dW256 plain beta=1 %.1f us\n", R, ms * 1e3);
    if (gemm(dm, y, dW, D, D, R, D, D, D, true, false, CUDA_R_32F, 1.f, 1.f, CUBLASLT_EPILOGUE_BGRADB, dbg, CUDA_R_32F, nullptr, 0, &ms) == 0)
      printf("time R=%d  dW256 BGRADB       %.1f us\n", R, ms * 1e3);
    // the big cross-attention one: dWk[256, 64] += dk2^T memk with R = 56 x 4060
    const int RM = 227360;
    __nv_bfloat16 *dk2, *memk; CK(cudaMalloc(&dk2, (size_t)RM * D * 2)); CK(cudaMalloc(&memk, (size_t)RM * 64 * 2));
    CK(cudaMemset(dk2, 0, (size_t)RM * D * 2)); CK(cudaMemset(memk, 0, (size_t)RM * 64 * 2));
    if (gemm(dk2, memk, dW, D, 64, RM, D, 64, 64, true, false, CUDA_R_32F, 1.f, 1.f, CUBLASLT_EPILOGUE_DEFAULT, nullptr, CUDA_R_32F, nullptr, 0, &ms) == 0)
      printf("time RM=%d dWk plain beta=1   %.1f us\n", RM, ms * 1e3);
    if (gemm(dk2, memk, dW, D, 64, RM, D, 64, 64, true, false, CUDA_R_32F, 1.f, 1.f, CUBLASLT_EPILOGUE_BGRADB, dbg, CUDA_R_32F, nullptr, 0, &ms) == 0)
      printf("time RM=%d dWk BGRADB         %.1f us\n", RM, ms * 1e3);
  }
  return 0;
}
