// Standalone (no torch) check of the tcgen05 GEMM against a naive CUDA-core
// reference on the same bf16 inputs.  Build: see Makefile target `cuda_tests`.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../multimodal-context-reasoning_b200/csrc/gemm.h"
#include "../../multimodal-context-reasoning_b200/csrc/tma_host.h"

using namespace modcr;

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);  \
      exit(2);                                                                         \
    }                                                                                  \
  } while (0)

__host__ __device__ inline float gelu_h(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__host__ __device__ inline float gelu_grad_h(float x) {
  float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
  float pdf = 0.3989422804014327f * expf(-0.5f * x * x);
  return cdf + x * pdf;
}

// reference: one thread per output
__global__ void ref_gemm(const __nv_bfloat16* A, const __nv_bfloat16* B, float* C, int M, int N, int K, int64_t lda,
                         int64_t ldb, int a_mn, int b_mn, float alpha, const float* bias, int act,
                         const __nv_bfloat16* in, int in_op, float* pre) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  int m = blockIdx.y;
  if (n >= N || m >= M) return;
  float acc = 0.f;
  for (int k = 0; k < K; ++k) {
    float a = __bfloat162float(a_mn ? A[(int64_t)k * lda + m] : A[(int64_t)m * lda + k]);
    float b = __bfloat162float(b_mn ? B[(int64_t)k * ldb + n] : B[(int64_t)n * ldb + k]);
    acc += a * b;
  }
  float x = acc * alpha + (bias ? bias[n] : 0.f);
  if (pre) pre[(int64_t)m * N + n] = x;
  if (act == 1) x = gelu_h(x);
  if (act == 2) x = tanhf(x);
  if (in_op == 1) x += __bfloat162float(in[(int64_t)m * N + n]);
  if (in_op == 2) x *= gelu_grad_h(__bfloat162float(in[(int64_t)m * N + n]));
  C[(int64_t)m * N + n] = x;
}

static uint32_t rng = 12345;
static float frand() {
  rng = rng * 1664525u + 1013904223u;
  return ((rng >> 8) & 0xFFFF) / 65536.0f - 0.5f;
}

struct Case {
  const char* name;
  int M, N, K, a_mn, b_mn, c_fp32, accumulate, split_k, act, has_bias, has_aux, in_op;
  int reps;
};

static int run_case(const Case& c) {
  const int M = c.M, N = c.N, K = c.K;
  // stored shapes
  size_t a_rows = c.a_mn ? K : M, a_cols = c.a_mn ? M : K;
  size_t b_rows = c.b_mn ? K : N, b_cols = c.b_mn ? N : K;
  std::vector<__nv_bfloat16> hA(a_rows * a_cols), hB(b_rows * b_cols), hIn((size_t)M * N);
  std::vector<float> hBias(N), hC0((size_t)M * N);
  for (auto& v : hA) v = __float2bfloat16(frand());
  for (auto& v : hB) v = __float2bfloat16(frand() * 0.25f);
  for (auto& v : hIn) v = __float2bfloat16(frand() * 2.0f);
  for (auto& v : hBias) v = frand();
  for (auto& v : hC0) v = c.accumulate ? frand() : 0.f;
  __nv_bfloat16 *dA, *dB, *dIn, *dCb, *dAux;
  float *dBias, *dCf, *dRef, *dPre;
  CK(cudaMalloc(&dA, hA.size() * 2));
  CK(cudaMalloc(&dB, hB.size() * 2));
  CK(cudaMalloc(&dIn, hIn.size() * 2));
  CK(cudaMalloc(&dCb, (size_t)M * N * 2));
  CK(cudaMalloc(&dAux, (size_t)M * N * 2));
  CK(cudaMalloc(&dBias, N * 4));
  CK(cudaMalloc(&dCf, (size_t)M * N * 4));
  CK(cudaMalloc(&dRef, (size_t)M * N * 4));
  CK(cudaMalloc(&dPre, (size_t)M * N * 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dIn, hIn.data(), hIn.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dBias, hBias.data(), N * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dCf, hC0.data(), (size_t)M * N * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dCb, 0xFF, (size_t)M * N * 2));
  CK(cudaMemset(dAux, 0xFF, (size_t)M * N * 2));

  dim3 rg((N + 127) / 128, M);
  ref_gemm<<<rg, 128>>>(dA, dB, dRef, M, N, K, a_cols, b_cols, c.a_mn, c.b_mn, 1.0f, c.has_bias ? dBias : nullptr,
                        c.act, dIn, c.in_op, dPre);
  CK(cudaDeviceSynchronize());

  GemmDesc g;
  g.A = dA; g.B = dB; g.lda = a_cols; g.ldb = b_cols;
  g.a_mn = c.a_mn; g.b_mn = c.b_mn;
  g.C = c.c_fp32 ? (void*)dCf : (void*)dCb; g.ldc = N; g.c_fp32 = c.c_fp32;
  g.accumulate = c.accumulate; g.split_k = c.split_k;
  g.M = M; g.N = N; g.K = K; g.alpha = 1.0f;
  g.bias = c.has_bias ? dBias : nullptr; g.act = c.act;
  if (c.has_aux) { g.aux = dAux; g.ldaux = N; }
  if (c.in_op) { g.epi_in = dIn; g.ld_in = N; g.in_op = c.in_op; }
  int rc = gemm_launch(g, 0);
  if (rc) { printf("[FAIL] %s: launch rc=%d (%s)\n", c.name, rc, last_error()); return 1; }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("[FAIL] %s: kernel error %s\n", c.name, cudaGetErrorString(e)); exit(3); }

  std::vector<float> hRef((size_t)M * N), hGot((size_t)M * N), hPre((size_t)M * N);
  CK(cudaMemcpy(hRef.data(), dRef, (size_t)M * N * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hPre.data(), dPre, (size_t)M * N * 4, cudaMemcpyDeviceToHost));
  if (c.c_fp32) {
    CK(cudaMemcpy(hGot.data(), dCf, (size_t)M * N * 4, cudaMemcpyDeviceToHost));
  } else {
    std::vector<__nv_bfloat16> t((size_t)M * N);
    CK(cudaMemcpy(t.data(), dCb, (size_t)M * N * 2, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < t.size(); ++i) hGot[i] = __bfloat162float(t[i]);
  }
  double max_err = 0, max_ref = 0;
  size_t bad = 0, first_bad = (size_t)-1;
  for (size_t i = 0; i < hRef.size(); ++i) {
    float ref = hRef[i] + (c.accumulate ? hC0[i] : 0.f);
    float tol = c.c_fp32 ? (2e-3f + 1e-3f * fabsf(ref)) : (2e-2f + 1e-2f * fabsf(ref));
    float err = fabsf(hGot[i] - ref);
    if (!(err <= tol)) { if (bad == 0) first_bad = i; ++bad; }
    if (err > max_err || err != err) max_err = err;
    if (fabsf(ref) > max_ref) max_ref = fabsf(ref);
  }
  size_t bad_aux = 0;
  if (c.has_aux) {
    std::vector<__nv_bfloat16> t((size_t)M * N);
    CK(cudaMemcpy(t.data(), dAux, (size_t)M * N * 2, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < t.size(); ++i) {
      float err = fabsf(__bfloat162float(t[i]) - hPre[i]);
      if (!(err <= 2e-2f + 1e-2f * fabsf(hPre[i]))) ++bad_aux;
    }
  }
  float ms = 0;
  if (c.reps > 0) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) gemm_launch(g, 0);
    cudaEventRecord(e0);
    for (int i = 0; i < c.reps; ++i) gemm_launch(g, 0);
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= c.reps;
  }
  double tf = ms > 0 ? 2.0 * M * N * K / (ms * 1e-3) / 1e12 : 0;
  printf("[%s] %-28s M=%d N=%d K=%d max_err=%.4g max_ref=%.3g bad=%zu bad_aux=%zu", (bad || bad_aux) ? "FAIL" : "PASS",
         c.name, M, N, K, max_err, max_ref, bad, bad_aux);
  if (ms > 0) printf("  %.3f ms  %.1f TFLOP/s", ms, tf);
  printf("\n");
  if (bad) {
    size_t i = first_bad;
    printf("    first bad at (m=%zu, n=%zu): got %g ref %g\n", i / N, i % N, hGot[i],
           hRef[i] + (c.accumulate ? hC0[i] : 0.f));
    // error map per 128x64 block (first 4x8 blocks)
    for (int bm = 0; bm < 4 && bm * 128 < M; ++bm) {
      printf("    blk row %d:", bm);
      for (int bn = 0; bn < 12 && bn * 64 < N; ++bn) {
        size_t cnt = 0;
        for (int m = bm * 128; m < (bm + 1) * 128 && m < M; ++m)
          for (int n = bn * 64; n < (bn + 1) * 64 && n < N; ++n) {
            size_t k = (size_t)m * N + n;
            float ref = hRef[k] + (c.accumulate ? hC0[k] : 0.f);
            if (!(fabsf(hGot[k] - ref) <= 2e-2f + 1e-2f * fabsf(ref))) ++cnt;
          }
        printf(" %5zu", cnt);
      }
      printf("\n");
    }
  }
  cudaFree(dA); cudaFree(dB); cudaFree(dIn); cudaFree(dCb); cudaFree(dAux); cudaFree(dBias); cudaFree(dCf);
  cudaFree(dRef); cudaFree(dPre);
  return (bad || bad_aux) ? 1 : 0;
}

int main(int argc, char** argv) {
  bool perf = argc > 1 && !strcmp(argv[1], "perf");
  std::vector<Case> cases = {
      // name                      M     N     K   aMN bMN f32 acc split act bias aux in reps
      {"nt_small",                128,  256,   64,  0,  0,  0,  0,  1,  0,  0,  0,  0, 0},
      {"nt_k768",                 128,  256,  768,  0,  0,  0,  0,  1,  0,  1,  0,  0, 0},
      {"nt_cfg1_qkv",            1424, 2304,  768,  0,  0,  0,  0,  1,  0,  1,  0,  0, 0},
      {"nt_ktail_img",            400,  768, 2056,  0,  0,  0,  0,  1,  0,  1,  0,  0, 0},
      {"nt_gelu_aux",            1424, 3072,  768,  0,  0,  0,  0,  1,  1,  1,  1,  0, 0},
      {"nt_residual_in",         1424,  768, 3072,  0,  0,  0,  0,  1,  0,  1,  0,  1, 0},
      {"nt_tanh_small",             8,  768,  768,  0,  0,  0,  0,  1,  2,  1,  0,  0, 0},
      {"nn_dgrad",               1424,  768, 2304,  0,  1,  0,  0,  1,  0,  0,  0,  0, 0},
      {"nn_dgrad_gelugrad",      1424, 3072,  768,  0,  1,  0,  0,  1,  0,  0,  0,  2, 0},
      {"nn_dgrad_add",           1424,  768, 3072,  0,  1,  0,  0,  1,  0,  0,  0,  1, 0},
      {"nt_f32",                  256,  512,  768,  0,  0,  1,  0,  1,  0,  1,  0,  0, 0},
      {"tn_wgrad_f32",            768,  768, 1424,  1,  1,  1,  0,  1,  0,  0,  0,  0, 0},
      {"tn_wgrad_splitk_acc",     768, 3072, 1424,  1,  1,  1,  1,  4,  0,  0,  0,  0, 0},
      {"tn_wgrad_ragged",         768, 2056,  400,  1,  1,  1,  1,  2,  0,  0,  0,  0, 0},
  };
  if (perf) {
    cases = {
        {"perf_qkv",             22784, 2304,  768,  0,  0,  0,  0,  1,  0,  1,  0,  0, 20},
        {"perf_wo",              22784,  768,  768,  0,  0,  0,  0,  1,  0,  1,  0,  0, 20},
        {"perf_ffn1_gelu_aux",   22784, 3072,  768,  0,  0,  0,  0,  1,  1,  1,  1,  0, 20},
        {"perf_ffn2",            22784,  768, 3072,  0,  0,  0,  0,  1,  0,  1,  0,  0, 20},
        {"perf_dgrad_ffn2_gg",   22784, 3072,  768,  0,  1,  0,  0,  1,  0,  0,  0,  2, 20},
        {"perf_dgrad_ffn1_add",  22784,  768, 3072,  0,  1,  0,  0,  1,  0,  0,  0,  1, 20},
        {"perf_wgrad_ffn1",       3072,  768, 22784, 1,  1,  1,  1,  2,  0,  0,  0,  0, 20},
        {"perf_wgrad_qkv",        2304,  768, 22784, 1,  1,  1,  1,  3,  0,  0,  0,  0, 20},
        {"perf_wgrad_wo",          768,  768, 22784, 1,  1,  1,  1,  8,  0,  0,  0,  0, 20},
        {"perf_square_8k",        8192, 8192,  8192, 0,  0,  0,  0,  1,  0,  0,  0,  0, 10},
    };
  }
  int fails = 0;
  for (auto& c : cases) fails += run_case(c);
  printf("%s: %d failing case(s)\n", perf ? "PERF" : "CHECK", fails);
  return fails ? 1 : 0;
}
