// Standalone bring-up probe for b200mm_gemm_bf16 (not part of the product; used under gpurun before the Python
// stack exists). Compares against a naive fp32 CUDA reference and times the kernel with CUDA events.
//   gemm_probe a_mn b_mn M N K splits [iters]
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../include/b200mm.h"

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                         \
    }                                                                                  \
  } while (0)

__global__ void ref_gemm(const __nv_bfloat16* A, int64_t lda, int a_mn, const __nv_bfloat16* B, int64_t ldb, int b_mn, float* D,
                         int64_t M, int64_t N, int64_t K) {
  int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t m = blockIdx.y;
  if (n >= N || m >= M) return;
  float acc = 0.f;
  for (int64_t k = 0; k < K; ++k) {
    float a = __bfloat162float(a_mn ? A[k * lda + m] : A[m * lda + k]);
    float b = __bfloat162float(b_mn ? B[k * ldb + n] : B[n * ldb + k]);
    acc += a * b;
  }
  D[m * N + n] = acc;
}

static uint32_t rng_state = 12345u;
static float frand() {
  rng_state = rng_state * 1664525u + 1013904223u;
  return ((rng_state >> 8) & 0xFFFF) / 65536.0f - 0.5f;
}

int main(int argc, char** argv) {
  if (argc < 7) {
    printf("usage: gemm_probe a_mn b_mn M N K splits [iters]\n");
    return 1;
  }
  int a_mn = atoi(argv[1]), b_mn = atoi(argv[2]);
  int64_t M = atoll(argv[3]), N = atoll(argv[4]), K = atoll(argv[5]);
  int splits = atoi(argv[6]);
  int iters = argc > 7 ? atoi(argv[7]) : 0;
  if (b200mm_check_device() != 0) {
    printf("device check: %s\n", b200mm_last_error());
    return 3;
  }
  int64_t a_rows = a_mn ? K : M, a_cols = a_mn ? M : K;
  int64_t b_rows = b_mn ? K : N, b_cols = b_mn ? N : K;
  int64_t lda = (a_cols + 7) / 8 * 8, ldb = (b_cols + 7) / 8 * 8;
  std::vector<__nv_bfloat16> hA(a_rows * lda), hB(b_rows * ldb);
  for (auto& x : hA) x = __float2bfloat16(frand());
  for (auto& x : hB) x = __float2bfloat16(frand());
  __nv_bfloat16 *dA, *dB;
  float *dD, *dRef, *dWs = nullptr;
  CK(cudaMalloc(&dA, hA.size() * 2));
  CK(cudaMalloc(&dB, hB.size() * 2));
  CK(cudaMalloc(&dD, M * N * 4));
  CK(cudaMalloc(&dRef, M * N * 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0xFF, M * N * 4));
  int64_t ws_bytes = b200mm_gemm_workspace_bytes(M, N, splits);
  if (ws_bytes) CK(cudaMalloc(&dWs, ws_bytes));

  b200mm_gemm_args g = {};
  g.A = dA; g.lda = lda; g.a_mn = a_mn;
  g.B = dB; g.ldb = ldb; g.b_mn = b_mn;
  g.D = dD; g.ldd = N; g.d_f32 = 1;
  g.M = M; g.N = N; g.K = K; g.alpha = 1.f;
  g.splits = splits; g.workspace = dWs; g.workspace_bytes = ws_bytes;
  int rc = b200mm_gemm_bf16(&g, nullptr);
  if (rc) {
    printf("gemm rc=%d: %s\n", rc, b200mm_last_error());
    return 4;
  }
  CK(cudaDeviceSynchronize());
  dim3 grid((unsigned)((N + 127) / 128), (unsigned)M);
  ref_gemm<<<grid, 128>>>(dA, lda, a_mn, dB, ldb, b_mn, dRef, M, N, K);
  CK(cudaDeviceSynchronize());
  std::vector<float> hD(M * N), hR(M * N);
  CK(cudaMemcpy(hD.data(), dD, M * N * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hR.data(), dRef, M * N * 4, cudaMemcpyDeviceToHost));
  double max_abs = 0, ref_sq = 0, err_sq = 0;
  int64_t bad = 0, first_bad = -1;
  for (int64_t i = 0; i < M * N; ++i) {
    double d = (double)hD[i] - hR[i];
    if (!(fabs(d) <= 1e-2 + 1e-3 * fabs(hR[i]))) {
      if (first_bad < 0) first_bad = i;
      ++bad;
    }
    if (fabs(d) > max_abs) max_abs = fabs(d);
    ref_sq += (double)hR[i] * hR[i];
    err_sq += d * d;
  }
  printf("a_mn=%d b_mn=%d M=%lld N=%lld K=%lld splits=%d  max_abs=%.4g rel_l2=%.3g bad=%lld", a_mn, b_mn, (long long)M,
         (long long)N, (long long)K, splits, max_abs, sqrt(err_sq / (ref_sq + 1e-30)), (long long)bad);
  if (first_bad >= 0)
    printf(" first_bad=(%lld,%lld) got=%g ref=%g", (long long)(first_bad / N), (long long)(first_bad % N), hD[first_bad],
           hR[first_bad]);
  printf(" %s\n", bad == 0 ? "PASS" : "FAIL");

  if (iters > 0) {
    // bf16 output for the timing run (the production configuration)
    __nv_bfloat16* dOut;
    CK(cudaMalloc(&dOut, M * N * 2));
    g.D = dOut; g.d_f32 = 0;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) b200mm_gemm_bf16(&g, nullptr);
    CK(cudaEventRecord(e0));
    for (int i = 0; i < iters; ++i) b200mm_gemm_bf16(&g, nullptr);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= iters;
    printf("  time %.3f ms  %.1f TFLOP/s\n", ms, 2.0 * M * N * K / (ms * 1e-3) / 1e12);
  }
  return bad == 0 ? 0 : 5;
}
