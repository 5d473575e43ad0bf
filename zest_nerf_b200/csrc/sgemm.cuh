// Generic-stride fp32 GEMM on CUDA cores with a fused MLP epilogue (bias, gate, ReLU).
// Used by the fp32 radiance-MLP path (forward and backward); the bf16 hot path is mlp_tc.cu.
#pragma once
#include "common.cuh"

namespace zest {

struct GemmArgs {
  // C[i, j] (+)= sum_kk A(i, kk) * B(j, kk);  A(i,kk) = A[i*sa_i + kk*sa_k], B(j,kk) = B[j*sb_j + kk*sb_k]
  const float* A; int64_t sa_i, sa_k;
  const float* B; int64_t sb_j, sb_k;
  float* C; int64_t ldc;
  int64_t I; int J; int64_t K;
  const float* bias = nullptr;   // [J]   v = acc + bias[j]
  float* Z = nullptr; int64_t ldz = 0;          // optional: Z[i,j] = v (pre-gate)
  const float* gate = nullptr; int64_t ldg = 0; // optional: v *= gate[i,j]
  int relu = 0;                  // v = max(v, 0)
  int accumulate = 0;            // C += v instead of C = v (atomic when split over K)
  int splits = 1;                // grid.z split of the K range (forces accumulate, atomic)
};

int launch_gemm(const GemmArgs& a, cudaStream_t st);

}  // namespace zest
