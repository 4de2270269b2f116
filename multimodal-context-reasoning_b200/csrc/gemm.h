// Internal C++ interface of the tcgen05 GEMM (see gemm.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace modcr {

struct GemmDesc {
  const void* A = nullptr;   // bf16
  const void* B = nullptr;   // bf16
  void* C = nullptr;         // bf16 or fp32
  int64_t lda = 0, ldb = 0, ldc = 0;  // leading dimensions (elements) of the STORED matrices
  int a_mn = 0;              // 0: A stored [M,K]; 1: A stored [K,M]
  int b_mn = 0;              // 0: B stored [N,K]; 1: B stored [K,N]
  int c_fp32 = 0;
  int accumulate = 0;        // C += result (fp32 only; TMA reduce-add)
  int split_k = 1;           // >1 requires accumulate
  int M = 0, N = 0, K = 0;
  float alpha = 1.0f;
  const float* bias = nullptr;  // [N] fp32
  int act = 0;               // 0 none, 1 gelu(erf), 2 tanh
  void* aux = nullptr;       // optional bf16 [M,N] copy of the pre-activation
  int64_t ldaux = 0;
  const void* epi_in = nullptr;  // optional bf16 [M,N] epilogue input tile
  int64_t ld_in = 0;
  int in_op = 0;             // 1: C = f(acc) + in ; 2: C = f(acc) * gelu'(in) ;
                             // 3 (fp32 in / out): C = f(acc) + LayerNorm(in) = f(acc) + (in - mean[m]) rstd[m] gamma[n] + beta[n]
  const float* in_mean = nullptr;   // in_op 3: fp32 [M] row statistics of `in` ...
  const float* in_rstd = nullptr;
  const float* in_gamma = nullptr;  // ... and fp32 [N] affine parameters of the LayerNorm that produced the residual
  const float* in_beta = nullptr;
  float* a_colsum = nullptr; // optional fp32 [M]: += sum_k A[k,m] (A MN-major only; the bias gradient of a wgrad)
  // training-mode dropout of f(acc + bias) before the epilogue input: mask = hash(*drop_seed, drop_site, m * N + n)
  const uint32_t* drop_seed = nullptr;
  uint32_t drop_site = 0;
  float drop_p = 0.f;
};

int gemm_launch(const GemmDesc& g, cudaStream_t stream);
int gemm_split_k_heuristic(int M, int N, int K, int bn);

}  // namespace modcr
