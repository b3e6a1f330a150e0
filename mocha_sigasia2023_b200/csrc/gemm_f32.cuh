// fp32 SIMT GEMM used by the fp32 ("parity") precision mode of every dense layer on the path.
//   C[z][M,N] = epilogue( alpha * prologue(A[z])[M,K] * W[z]^T )
// A rows can be plain (row-major, lda) or gathered as a reflect-padded temporal convolution over
// a channel-last [B,T,V,C] activation (implicit GEMM, no im2col buffer).
#pragma once
#include "common.cuh"

namespace mocha {

enum GemmAct : int { ACT_NONE = 0, ACT_RELU = 1, ACT_GELU = 2, ACT_LRELU = 3 };

struct GemmParams {
  const float* A = nullptr;
  const float* W = nullptr;
  float* C = nullptr;
  int M = 0, N = 0, K = 0;
  int lda = 0, ldw = 0, ldc = 0;
  // batching: blockIdx.z = z1 * nz2 + z2
  int nz = 1, nz2 = 1;
  long long sA1 = 0, sA2 = 0, sW1 = 0, sW2 = 0, sC1 = 0, sC2 = 0;
  // implicit temporal convolution on A: rows are (b,t,v), k = tap*Cin + ci,
  // source row = (b*(T/tdiv) + reflect(t + tap - taps/2, T)/tdiv) * V + v
  int conv = 0, T = 0, V = 0, taps = 0, Cin = 0, tdiv = 1;
  int a_lrelu = 0;  // LeakyReLU(0.2) applied to A on load (pre-activation blocks)
  int w_kn = 0;     // W stored [K,N] row-major instead of [N,K]
  float alpha = 1.f;
  const float* bias = nullptr;
  int bias_period = 0;  // 0: bias[n]; p>0: bias[(row % p) * N + n]
  int act = ACT_NONE;
  const float* res = nullptr;  // residual added after the activation
  int ldr = 0;
  long long sR1 = 0, sR2 = 0;
};

// Enqueue on `stream`; returns MOCHA_OK or an error code (message via last_error()).
int gemm_f32(const GemmParams& p, cudaStream_t stream);

}  // namespace mocha
