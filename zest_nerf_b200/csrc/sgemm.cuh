// Generic-stride fp32 GEMM with a fused MLP epilogue (bias, gate, ReLU): CUDA-core kernel (sgemm.cu) and
// tcgen05 split-bf16 kernel (tc_gemm.cu).  Used by the fp32 radiance-MLP path (forward and backward); the
// bf16 inference hot path is mlp_tc.cu.
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
  // optional: rowsum[i] += sum_k A(i, k) over this call's K range (the bias gradient of a dW GEMM, A = dY^T)
  float* rowsum = nullptr;
  // optional fused gate backward (networks.py:176-180 reversed) on output columns j >= gb_col0, jj = j - gb_col0:
  //   v = gradient wrt h = relu(z * g);  m = (z * g > 0) ? v : 0;  gb_dZ[i, jj] = m * g;  gb_gG[i, jj] += m * z
  // (z = gb_Z, g = gb_G; all four [I, gb_ld]).  Those columns are not written to C.
  // gb_from_h: gb_Z holds the layer's stored OUTPUT h = relu(z * g) instead of z (the forward then never writes z): the
  // mask is h > 0 - the same fp32 product the forward rounded, so the same bits as z * g > 0 - and z = h / g where it is on
  // (g != 0 there; within 1 ulp of the z the forward would have stored).
  const float* gb_Z = nullptr; const float* gb_G = nullptr; float* gb_gG = nullptr; float* gb_dZ = nullptr;
  int64_t gb_ld = 0; int gb_col0 = 0; int gb_from_h = 0;
  // tensor-core engines only: device scratch in which B (a small matrix every row tile re-reads: the weights) is packed
  // into UMMA stage images once per call and then streamed by TMA; nullptr = stage B through registers like A
  void* b_scratch = nullptr; int64_t b_scratch_bytes = 0;
};

// z of the fused gate backward: stored as is, or recovered from the stored output h = relu(z * g) (GemmArgs::gb_from_h)
// (MUFU.RCP + multiply, <= 2 ulp: the IEEE division's ~30-instruction sequence made the 8-warp epilogue of the dX kernels
// 60 % longer, measured)
__device__ __forceinline__ float gate_bwd_z(float zz, float gg, bool on, int from_h) {
  return from_h ? (on ? __fdividef(zz, gg) : 0.f) : zz;
}

// dispatcher: the tcgen05 split-precision kernel (tc_gemm.cu) when the engine is 1 or 2 (default 2) and the operand strides
// allow it, else the exact-fp32 CUDA-core kernel (sgemm.cu)
int launch_gemm(const GemmArgs& a, cudaStream_t st);
int launch_gemm_simt(const GemmArgs& a, cudaStream_t st);
int launch_gemm_tc(const GemmArgs& a, int engine, cudaStream_t st);   // engine 1 = 3 x bf16, 2 = 3 x tf32
bool gemm_tc_supported(const GemmArgs& a);
int gemm_engine();            // 0 = SIMT fp32, 1 = tcgen05 3 x bf16, 2 = tcgen05 3 x tf32, split accumulators (default); env ZEST_GEMM=simt|bf16x3
void set_gemm_engine(int e);

}  // namespace zest
