// Generic-stride fp32 GEMM on the tcgen05 tensor cores: C[I,J] (+)= A[I,K] * B[J,K]^T with fp32 operands in
// global memory, each split on the fly into a head and a residual (a = a_hi + a_lo) and multiplied as three
// UMMAs with fp32 accumulation in TMEM:
//
//     A*B ~= a_lo*b_hi + a_hi*b_lo + a_hi*b_hi
//
//   engine 1: bf16 head + bf16 residual (16 mantissa bits), one accumulator
//   engine 2: tf32 head (round to nearest) + fp32 residual read as tf32 (21-22 mantissa bits, twice the tensor-pipe time); the
//             weight-operand GEMMs (forward, dX) keep the head product and the cross products in two accumulators
// Measured against fp64 (tests/test_gpu_parity.py), K = 256..347, one accumulator: fp32 SIMT 6e-7, 3 x tf32 2-3.5e-6,
// 3 x bf16 4-6e-6 of max|C|; unsplit K = 5000: 3.2e-6 / 3.7e-5 / 1.8e-5.  The tensor core's fp32 accumulate TRUNCATES:
// a bias per chained UMMA that grows linearly with the chain and compounds through the layers of the MLP.  Hence: long
// reductions are cut into <= 192-UMMA accumulators joined by round-to-nearest atomics, and engine 2 accumulates the
// small cross products apart from the head product.
//
// This is the training path's GEMM (fine_tune.py: forward with saved activations, dX, dW of networks.py:150-221),
// taking the same GemmArgs / fused epilogue (bias, pre-gate copy, gate, ReLU, accumulate, split-K atomics) as the
// CUDA-core sgemm_kernel it replaces; the exact-fp32 SIMT kernel stays selectable (zest_set_gemm_engine).
//
// Three kernels; K stages of 16-byte chunks per row and part (32 k as bf16, 16 k as tf32):
//   tc_gemm_packed_kernel  B is a weight matrix every row tile re-reads (forward, dX): packed once per call into the UMMA
//                          stage images and streamed by bulk copies.  Two CTAs per SM (<= 112.25 KB smem, 256 TMEM columns
//                          each: one CTA's epilogue runs under the other's main loop).  A arrives by 2-D tensor copies into
//                          a landing ring (a tenth warp), eight worker warps split it into the hi / lo images and run the
//                          epilogue, a ninth warp drives the weight copies and issues the UMMAs (converged warp, warp-
//                          uniform descriptors); mbarriers only in the main loop.  Operands that do not qualify for a
//                          tensor map are staged through registers (static register sets, up to 6 stages ahead).
//                          128 x 256 tiles, or 128 x 128 with separate head / cross accumulators (engine 2).
//   tc_gemm_rc_tma_kernel  both operands are [samples, features] activations and the reduction is long (dW): one CTA per SM
//                          owns a 256 x 256 tile (512 TMEM columns) and a K slice; both operands land by tensor copies, 16
//                          worker warps transpose + split them out of shared memory, split-K joined by vector atomics.
//   tc_gemm_kernel         everything else (dW of the narrow heads, unaligned operands): all 256 threads stage A and B
//                          through registers into the K-major no-swizzle images (core matrix = 8 rows x 16 B, the MLP
//                          kernel's layout), thread 0 issues 6 UMMAs per stage (2 K steps x 3 products) and commits to the
//                          stage's "empty" mbarrier; the bias gradient (row sums of A) rides along in the staging threads.
// Epilogue (shared): tcgen05.ld 32 columns per warp pass (lane = row), per-warp shared-memory transpose, fused ops (bias,
// Z copy, gate, ReLU, accumulate, the gate backward of the layer below, vector atomics for split-K), global accesses in
// which a warp covers 4 rows x 128 contiguous bytes, every load of a row batch issued before its first use.
#include <atomic>
#include <cstdlib>
#include <cstring>

#include <cuda.h>   // CUtensorMap and the cuTensorMapEncodeTiled prototype only: the entry point is looked up through the runtime

#include "sgemm.cuh"
#include "tc_ptx.cuh"

namespace zest {
namespace {

constexpr int GM = 128, GN = 256;
constexpr int kThreads = 256;
// A stage image = 4 K-chunks of 16 bytes per row and per part (hi / lo): 32 k as bf16 (KCH = 8 elements per chunk)
// or 16 k as tf32 (KCH = 4).  Same bytes either way.
constexpr uint32_t kAHalf = GM * 4 * 16;               // bytes of one part (hi or lo) of the A stage: 8 KB
constexpr uint32_t kBHalf = GN * 4 * 16;               // 16 KB
constexpr uint32_t kStage = 2 * kAHalf + 2 * kBHalf;   // 48 KB
constexpr int kStages = 2;
constexpr uint32_t kSmem = kStages * kStage;           // 96 KB -> two CTAs per SM
constexpr uint32_t kEpiRow = 144;                      // epilogue staging: 32 fp32 columns + 16 B pad per row
static_assert(8 * 32 * kEpiRow <= kSmem, "epilogue staging must fit in the stage buffers");

__device__ __forceinline__ void gemm_wait(uint32_t bar, uint32_t parity, int tag) {
  if (ptx::mbar_try_wait(bar, parity)) return;
  uint32_t spins = 0;
  while (!ptx::mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 21)) {  // a protocol bug must not hang the GPU
      printf("zest tc_gemm: barrier timeout tag=%d block=(%d,%d,%d) thread=%d\n", tag, blockIdx.x, blockIdx.y, blockIdx.z,
             threadIdx.x);
      __trap();
    }
  }
}

#ifndef ZEST_TF32_SPLIT
#define ZEST_TF32_SPLIT 2   // 2: head = (bits + 0x1000) & ~0x1fff (round to nearest, ties away = cvt.rna on finite values), residual left as the
                            // exact fp32 difference (the tensor core reads its top 19 bits); 0: cvt.rna on both, 1: masked residual.  Same measured
                            // GEMM and gradient errors; 3 instead of ~9 ALU instructions per element: 1.2 % of a fine-tune step (same-box A/B)
#endif
__device__ __forceinline__ uint32_t to_tf32(float a) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(a));
  return r;
}
// D[tmem] (+)= A[smem] * B[smem]^T, tf32 x tf32 -> fp32 (K = 8 per instruction)
__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// kind::tf32 instruction descriptor: D = f32, A = B = tf32 (format 2), both K-major, M = 128
__device__ __forceinline__ uint32_t idesc_tf32(int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
}

// one 16-byte K chunk of one row -> the hi image and the lo image
//   KCH = 8: 8 fp32 -> 8 bf16 heads + 8 bf16 residuals (a = hi + lo to 2^-18)
//   KCH = 4: 4 fp32 -> 4 tf32 heads (round to nearest) + 4 residuals (exact in fp32, truncated to tf32 by the tensor core: a = hi + lo to 2^-21)
template <int KCH>
__device__ __forceinline__ void split_chunk(const float* v, uint32_t (&h)[4], uint32_t (&l)[4]) {
  if constexpr (KCH == 8) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      h[e] = ptx::pack_bf16(v[2 * e], v[2 * e + 1]);   // element 2e in the low half = lower address
      l[e] = ptx::pack_bf16(v[2 * e] - ptx::bf16_lo(h[e]), v[2 * e + 1] - ptx::bf16_hi(h[e]));
    }
  } else {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
#if ZEST_TF32_SPLIT == 0
      h[e] = to_tf32(v[e]);
      l[e] = to_tf32(v[e] - __uint_as_float(h[e]));
#else
      h[e] = (__float_as_uint(v[e]) + 0x1000u) & 0xffffe000u;
      l[e] = __float_as_uint(v[e] - __uint_as_float(h[e]));
#if ZEST_TF32_SPLIT == 1
      l[e] &= 0xffffe000u;
#endif
#endif
    }
  }
}
template <int KCH>
__device__ __forceinline__ void store_chunk(uint32_t hi_addr, uint32_t lo_addr, const float* v) {
  uint32_t h[4], l[4];
  split_chunk<KCH>(v, h, l);
  ptx::st_smem_v4(hi_addr, h[0], h[1], h[2], h[3]);
  ptx::st_smem_v4(lo_addr, l[0], l[1], l[2], l[3]);
}

// ---- operand staging.  R = rows of the stage image (128 for A, 256 for B), NU = units per thread.
// k-contiguous operand (stride 1 along K): unit = 1 row x 1 chunk.  Lane -> (row = 8 (u / 32) + lane % 8, chunk = lane / 8):
// the 8 lanes of a quarter-warp store to 8 consecutive rows of one chunk (8 distinct 16-byte bank groups), and a warp load
// covers 8 rows x the stage's whole K extent.
template <int KCH, int R, int NU>
struct StageKC {
  float v[NU * KCH];
  __device__ __forceinline__ void fetch(const float* __restrict__ base, int64_t s_row, int64_t rows, int64_t r0, int64_t k0,
                                        int64_t ke, int rows_used, int tid) {
#pragma unroll
    for (int n = 0; n < NU; ++n) {
      const int u = tid + n * kThreads;
      const int row = (u >> 5) * 8 + (u & 7), c = (u >> 3) & 3;
      const int64_t r = r0 + row, k = k0 + c * KCH;
      const float* p = base + r * s_row + k;
      float* o = v + n * KCH;
      const bool in = row < rows_used && r < rows;
      if (in && k + KCH - 1 < ke && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
#pragma unroll
        for (int e = 0; e < KCH; e += 4) {
          const float4 a = __ldg(reinterpret_cast<const float4*>(p + e));
          o[e] = a.x; o[e + 1] = a.y; o[e + 2] = a.z; o[e + 3] = a.w;
        }
      } else {
#pragma unroll
        for (int e = 0; e < KCH; ++e) o[e] = (in && k + e < ke) ? __ldg(p + e) : 0.f;
      }
    }
  }
  // the two rows of a 128-row image this thread stages (NU = 2), and the running sums of what it staged for them
  static __device__ __forceinline__ int row_of(int slot, int tid) { const int u = tid + slot * kThreads; return (u >> 5) * 8 + (u & 7); }
  __device__ __forceinline__ void rowsums(float& s0, float& s1) const {
#pragma unroll
    for (int e = 0; e < KCH; ++e) { s0 += v[e]; s1 += v[(NU > 1 ? KCH : 0) + e]; }
  }
  __device__ __forceinline__ void store(uint32_t hi_base, uint32_t lo_base, int rows_used, int tid) const {
#pragma unroll
    for (int n = 0; n < NU; ++n) {
      const int u = tid + n * kThreads;
      const int row = (u >> 5) * 8 + (u & 7), c = (u >> 3) & 3;
      if (row >= rows_used) continue;
      const uint32_t off = (uint32_t)c * (R * 16) + (uint32_t)row * 16;
      store_chunk<KCH>(hi_base + off, lo_base + off, v + n * KCH);
    }
  }
};

// row-contiguous operand (stride 1 along the output dimension): unit = 2 rows x 1 chunk, one 8-byte load per k; the 32
// lanes of a load cover 256 contiguous bytes.
template <int KCH, int R, int NU>
struct StageRC {
  float v[NU * 2 * KCH];
  __device__ __forceinline__ void fetch(const float* __restrict__ base, int64_t s_k, int64_t rows, int64_t r0, int64_t k0,
                                        int64_t ke, int rows_used, int tid) {
#pragma unroll
    for (int n = 0; n < NU; ++n) {
      const int u = tid + n * kThreads;
      const int rp = u % (R / 2), c = u / (R / 2);
      const int64_t r = r0 + 2 * rp, k = k0 + c * KCH;
      const float* p = base + k * s_k + r;
      float* o = v + n * 2 * KCH;
      const bool used = 2 * rp < rows_used;
      const bool vec = used && r + 1 < rows && (s_k & 1) == 0 && (reinterpret_cast<uintptr_t>(p) & 7) == 0;
#pragma unroll
      for (int e = 0; e < KCH; ++e) {
        if (vec && k + e < ke) {
          const float2 t = __ldg(reinterpret_cast<const float2*>(p + e * s_k));
          o[e] = t.x; o[KCH + e] = t.y;
        } else {
          const bool kin = used && k + e < ke;
          o[e] = (kin && r < rows) ? __ldg(p + e * s_k) : 0.f;
          o[KCH + e] = (kin && r + 1 < rows) ? __ldg(p + e * s_k + 1) : 0.f;
        }
      }
    }
  }
  static __device__ __forceinline__ int row_of(int slot, int tid) { return 2 * (tid % (R / 2)) + slot; }
  __device__ __forceinline__ void rowsums(float& s0, float& s1) const {
#pragma unroll
    for (int e = 0; e < KCH; ++e) { s0 += v[e]; s1 += v[KCH + e]; }
  }
  __device__ __forceinline__ void store(uint32_t hi_base, uint32_t lo_base, int rows_used, int tid) const {
    // lanes 4..7 of every quarter-warp store their second row first: the eight 16-byte stores of one shared-memory
    // wavefront then land in eight distinct 16-byte bank groups
    const bool swap = (tid >> 2) & 1;
#pragma unroll
    for (int n = 0; n < NU; ++n) {
      const int u = tid + n * kThreads;
      const int rp = u % (R / 2), c = u / (R / 2);
      if (2 * rp >= rows_used) continue;
      const float* o = v + n * 2 * KCH;
      float first[KCH], second[KCH];
#pragma unroll
      for (int e = 0; e < KCH; ++e) { first[e] = swap ? o[KCH + e] : o[e]; second[e] = swap ? o[e] : o[KCH + e]; }
      const uint32_t off = (uint32_t)c * (R * 16) + (uint32_t)(2 * rp) * 16;
      const uint32_t o1 = off + (swap ? 16u : 0u), o2 = off + (swap ? 0u : 16u);
      store_chunk<KCH>(hi_base + o1, lo_base + o1, first);
      store_chunk<KCH>(hi_base + o2, lo_base + o2, second);
    }
  }
};

// Global accesses of the fused gate backward: L2-only loads / stores (nothing here is re-read through L1) and the gate
// gradient accumulated by a vector reduction (red.global.add.v4.f32: one round-to-nearest add per element per kernel, the same
// bits as load + add + store except that the reduction flushes denormals, one stream less in flight).  Measured inside a fine-tune step (tools/gemm_timeline_step.py): the
// dX kernels' CTA lifetime 56.0 k -> 50.3 k cycles; the hints alone 3 %, larger load batches nothing - the epilogue is bound
// by the memory system's throughput on this 3-read / 2-write mix (~4.5 TB/s), not by latency.
__device__ __forceinline__ float4 epi_ld(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void epi_st(float* p, float4 v) { __stcg(reinterpret_cast<float4*>(p), v); }

// ---- epilogue (8 warps).  Warp w owns TMEM lanes 32 (w % 4) .. +31 (rows) and the 32-column passes w / 4, w / 4 + 2, ...
// tcgen05.ld gives lane = row; a per-warp shared-memory transpose (the stage buffers are free by now) turns that into
// lane = (row % 4, 4 consecutive columns) so that every global access of a warp covers 4 rows x 128 contiguous bytes.
template <int CROSS = 0>   // CROSS > 0: add the cross-product accumulator that lives CROSS columns above the main one
__device__ __forceinline__ void gemm_epilogue(const GemmArgs& g, uint32_t tmem, uint32_t smem0, int warp, int lane, int64_t i0,
                                              int j0, int im, int jn, int n_mma, bool atomic, bool first_split) {
  const int q = warp & 3;
  const uint32_t stg = smem0 + (uint32_t)warp * (32 * kEpiRow);
  const int cq = (lane & 7) * 4, rsub = lane >> 3;
  for (int c0 = (warp >> 2) * 32; c0 < n_mma; c0 += 64) {
    uint32_t r[32];
    ptx::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
    if constexpr (CROSS > 0) {
      uint32_t r2[32];
      ptx::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)CROSS + (uint32_t)c0, r2);
      ptx::tc_wait_ld();
#pragma unroll
      for (int e = 0; e < 32; ++e) r[e] = __float_as_uint(__uint_as_float(r[e]) + __uint_as_float(r2[e]));
    } else {
      ptx::tc_wait_ld();
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) ptx::st_smem_v4(stg + (uint32_t)lane * kEpiRow + e * 16, r[4 * e], r[4 * e + 1], r[4 * e + 2], r[4 * e + 3]);
    __syncwarp();
    const int j = j0 + c0 + cq;                 // this lane's first column
    const int nv = g.J - j;                     // valid columns among its 4 (<= 0: none)
    float b4[4] = {0.f, 0.f, 0.f, 0.f};
    if (g.bias && first_split) {
#pragma unroll
      for (int t = 0; t < 4; ++t) if (t < nv) b4[t] = __ldg(g.bias + j + t);
    }
    // Fast paths: all four columns valid and every row's pointer 16-byte aligned (aligned first row, row strides multiples of
    // 4 floats).  Every global load of a batch of rows is issued before the first use (the epilogue is latency-bound otherwise).
    const auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    const bool in_gb = g.gb_dZ != nullptr && j + 3 >= g.gb_col0;
    const int jj = j - g.gb_col0;
    bool fast = nv >= 4 && !atomic;
    if (in_gb) fast = fast && jj >= 0 && (g.gb_ld & 3) == 0 && al16(g.gb_Z + jj) && al16(g.gb_G + jj) && al16(g.gb_gG + jj) && al16(g.gb_dZ + jj) &&
                      (!g.accumulate || ((g.ldc & 3) == 0 && al16(g.C + j)));
    else fast = fast && (g.ldc & 3) == 0 && al16(g.C + j) && (!g.Z || ((g.ldz & 3) == 0 && al16(g.Z + j))) &&
                (!g.gate || ((g.ldg & 3) == 0 && al16(g.gate + j)));
    if (fast && !in_gb) {          // forward / plain: [Z = v], v *= gate, relu, [+= C], C = v
#pragma unroll
      for (int h = 0; h < 1; ++h) {
        float4 gt4[8];
#pragma unroll
        for (int r4 = 0; r4 < 8; ++r4) {
          const int row = q * 32 + (h * 8 + r4) * 4 + rsub;
          const int64_t i = i0 + row;
          gt4[r4] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (row < im && g.gate) gt4[r4] = __ldg(reinterpret_cast<const float4*>(g.gate + i * g.ldg + j));
        }
#pragma unroll
        for (int r4 = 0; r4 < 8; ++r4) {
          const int lr = (h * 8 + r4) * 4 + rsub, row = q * 32 + lr;
          const uint4 w = ptx::ld_smem_v4(stg + (uint32_t)lr * kEpiRow + cq * 4);
          if (row >= im) continue;
          const int64_t i = i0 + row;
          float4 v = make_float4(__uint_as_float(w.x) + b4[0], __uint_as_float(w.y) + b4[1], __uint_as_float(w.z) + b4[2], __uint_as_float(w.w) + b4[3]);
          if (g.Z) *reinterpret_cast<float4*>(g.Z + i * g.ldz + j) = v;
          if (g.gate) { v.x *= gt4[r4].x; v.y *= gt4[r4].y; v.z *= gt4[r4].z; v.w *= gt4[r4].w; }
          if (g.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
          if (g.accumulate) {   // rare on this path (two dX GEMMs per pass): read at use
            const float4 o4 = *reinterpret_cast<const float4*>(g.C + i * g.ldc + j);
            v.x += o4.x; v.y += o4.y; v.z += o4.z; v.w += o4.w;
          }
          *reinterpret_cast<float4*>(g.C + i * g.ldc + j) = v;
        }
      }
      __syncwarp();
      continue;
    }
    if (fast) {                    // fused gate backward (see GemmArgs): v [+ C] -> dZ, gG
      constexpr int RG = 4;                   // row groups (of 4 rows) whose loads are in flight together (8: no faster)
#pragma unroll
      for (int h = 0; h < 8 / RG; ++h) {
        float4 zz[RG], gg[RG];
#pragma unroll
        for (int r2 = 0; r2 < RG; ++r2) {
          const int row = q * 32 + (h * RG + r2) * 4 + rsub;
          const int64_t i = i0 + row, o = i * g.gb_ld + jj;
          zz[r2] = gg[r2] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (row < im) {
            zz[r2] = epi_ld(g.gb_Z + o);
            gg[r2] = epi_ld(g.gb_G + o);
          }
        }
#pragma unroll
        for (int r2 = 0; r2 < RG; ++r2) {
          const int lr = (h * RG + r2) * 4 + rsub, row = q * 32 + lr;
          const uint4 w = ptx::ld_smem_v4(stg + (uint32_t)lr * kEpiRow + cq * 4);
          if (row >= im) continue;
          const int64_t o = (i0 + row) * g.gb_ld + jj;
          float4 o4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (g.accumulate) o4 = *reinterpret_cast<const float4*>(g.C + (i0 + row) * g.ldc + j);   // last layer only: read at use
          const float v0 = __uint_as_float(w.x) + b4[0] + o4.x, v1 = __uint_as_float(w.y) + b4[1] + o4.y;
          const float v2 = __uint_as_float(w.z) + b4[2] + o4.z, v3 = __uint_as_float(w.w) + b4[3] + o4.w;
          const int fh = g.gb_from_h;
          const bool on0 = fh ? zz[r2].x > 0.f : zz[r2].x * gg[r2].x > 0.f, on1 = fh ? zz[r2].y > 0.f : zz[r2].y * gg[r2].y > 0.f;
          const bool on2 = fh ? zz[r2].z > 0.f : zz[r2].z * gg[r2].z > 0.f, on3 = fh ? zz[r2].w > 0.f : zz[r2].w * gg[r2].w > 0.f;
          const float m0 = on0 ? v0 : 0.f, m1 = on1 ? v1 : 0.f, m2 = on2 ? v2 : 0.f, m3 = on3 ? v3 : 0.f;
          epi_st(g.gb_dZ + o, make_float4(m0 * gg[r2].x, m1 * gg[r2].y, m2 * gg[r2].z, m3 * gg[r2].w));
          atomicAdd(reinterpret_cast<float4*>(g.gb_gG + o),
                    make_float4(m0 * gate_bwd_z(zz[r2].x, gg[r2].x, on0, fh), m1 * gate_bwd_z(zz[r2].y, gg[r2].y, on1, fh),
                                m2 * gate_bwd_z(zz[r2].z, gg[r2].z, on2, fh), m3 * gate_bwd_z(zz[r2].w, gg[r2].w, on3, fh)));
        }
      }
      __syncwarp();
      continue;
    }
    // generic path: ragged edges, misaligned rows (the skip layer's [pe | h] blocks), split-K atomics
#pragma unroll 2
    for (int rr = 0; rr < 32; rr += 4) {
      const int row = q * 32 + rr + rsub;
      const uint4 w = ptx::ld_smem_v4(stg + (uint32_t)(rr + rsub) * kEpiRow + cq * 4);
      if (row >= im || nv <= 0) continue;
      const int64_t i = i0 + row;
      float v[4] = {__uint_as_float(w.x) + b4[0], __uint_as_float(w.y) + b4[1], __uint_as_float(w.z) + b4[2], __uint_as_float(w.w) + b4[3]};
      float* c = g.C + i * g.ldc + j;
      float* z = g.Z ? g.Z + i * g.ldz + j : nullptr;
      const float* gt = g.gate ? g.gate + i * g.ldg + j : nullptr;
      const bool full = nv >= 4;
      if (g.gb_dZ && j + 3 >= g.gb_col0) {   // fused gate backward (see GemmArgs); columns below gb_col0 fall through to C
        if (g.accumulate) {
#pragma unroll
          for (int t = 0; t < 4; ++t) if (t < nv) v[t] += c[t];
        }
        const int64_t o = i * g.gb_ld + (j - g.gb_col0);
        if (full && j >= g.gb_col0 && ((g.gb_ld | (int64_t)(j - g.gb_col0)) & 3) == 0 &&
            ((reinterpret_cast<uintptr_t>(g.gb_Z) | reinterpret_cast<uintptr_t>(g.gb_G) | reinterpret_cast<uintptr_t>(g.gb_gG) |
              reinterpret_cast<uintptr_t>(g.gb_dZ)) & 15) == 0) {
          const float4 zz = __ldg(reinterpret_cast<const float4*>(g.gb_Z + o));
          const float4 gg = __ldg(reinterpret_cast<const float4*>(g.gb_G + o));
          float4 acc = *reinterpret_cast<const float4*>(g.gb_gG + o);
          const int fh = g.gb_from_h;
          const bool on0 = fh ? zz.x > 0.f : zz.x * gg.x > 0.f, on1 = fh ? zz.y > 0.f : zz.y * gg.y > 0.f;
          const bool on2 = fh ? zz.z > 0.f : zz.z * gg.z > 0.f, on3 = fh ? zz.w > 0.f : zz.w * gg.w > 0.f;
          const float m0 = on0 ? v[0] : 0.f, m1 = on1 ? v[1] : 0.f, m2 = on2 ? v[2] : 0.f, m3 = on3 ? v[3] : 0.f;
          *reinterpret_cast<float4*>(g.gb_dZ + o) = make_float4(m0 * gg.x, m1 * gg.y, m2 * gg.z, m3 * gg.w);
          acc.x += m0 * gate_bwd_z(zz.x, gg.x, on0, fh); acc.y += m1 * gate_bwd_z(zz.y, gg.y, on1, fh);
          acc.z += m2 * gate_bwd_z(zz.z, gg.z, on2, fh); acc.w += m3 * gate_bwd_z(zz.w, gg.w, on3, fh);
          *reinterpret_cast<float4*>(g.gb_gG + o) = acc;
        } else {
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            if (t >= nv) continue;
            if (j + t < g.gb_col0) { c[t] = v[t]; continue; }     // (v already holds C + acc when accumulating)
            const float zz = __ldg(g.gb_Z + o + t), gg = __ldg(g.gb_G + o + t);
            const bool on = g.gb_from_h ? (zz > 0.f) : (zz * gg > 0.f);
            const float m = on ? v[t] : 0.f;
            g.gb_dZ[o + t] = m * gg;
            g.gb_gG[o + t] += m * gate_bwd_z(zz, gg, on, g.gb_from_h);
          }
        }
        continue;
      }
      if (z) {
        if (full && (reinterpret_cast<uintptr_t>(z) & 15) == 0) *reinterpret_cast<float4*>(z) = make_float4(v[0], v[1], v[2], v[3]);
        else {
#pragma unroll
          for (int t = 0; t < 4; ++t) if (t < nv) z[t] = v[t];
        }
      }
      if (gt) {
        if (full && (reinterpret_cast<uintptr_t>(gt) & 15) == 0) {
          const float4 gg = __ldg(reinterpret_cast<const float4*>(gt));
          v[0] *= gg.x; v[1] *= gg.y; v[2] *= gg.z; v[3] *= gg.w;
        } else {
#pragma unroll
          for (int t = 0; t < 4; ++t) if (t < nv) v[t] *= __ldg(gt + t);
        }
      }
      if (g.relu) {
#pragma unroll
        for (int t = 0; t < 4; ++t) v[t] = fmaxf(v[t], 0.f);
      }
      if (atomic) {
        if (full && (reinterpret_cast<uintptr_t>(c) & 15) == 0) atomicAdd(reinterpret_cast<float4*>(c), make_float4(v[0], v[1], v[2], v[3]));
        else {
#pragma unroll
          for (int t = 0; t < 4; ++t) if (t < nv) atomicAdd(c + t, v[t]);
        }
      } else if (full && (reinterpret_cast<uintptr_t>(c) & 15) == 0) {
        if (g.accumulate) {
          const float4 old = *reinterpret_cast<const float4*>(c);
          v[0] += old.x; v[1] += old.y; v[2] += old.z; v[3] += old.w;
        }
        *reinterpret_cast<float4*>(c) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int t = 0; t < 4; ++t) if (t < nv) c[t] = g.accumulate ? c[t] + v[t] : v[t];
      }
    }
    __syncwarp();   // the next pass overwrites the staging rows
  }
}

template <int KCH, bool KC, int R>
struct StagePick;
template <int KCH, int R>
struct StagePick<KCH, true, R> { using type = StageKC<KCH, R, R * 4 / kThreads>; };
template <int KCH, int R>
struct StagePick<KCH, false, R> { using type = StageRC<KCH, R, R * 2 / kThreads>; };

// KCH = 8: three bf16 UMMAs per product (K = 32 per stage); KCH = 4: three tf32 UMMAs per product (K = 16 per stage)
template <int KCH, bool AKC, bool BKC>
__global__ void __launch_bounds__(kThreads, 2) tc_gemm_kernel(GemmArgs g, int64_t kper) {
  constexpr int GK = 4 * KCH;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t empty_bar[kStages];
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t i0 = (int64_t)blockIdx.x * GM;
  const int j0 = (int)blockIdx.y * GN;
  const int64_t kb = (int64_t)blockIdx.z * kper;
  const int64_t ke = (kb + kper < g.K) ? kb + kper : g.K;
  if (kb >= ke) return;                       // empty K slice of a split (nothing to add); uniform for the CTA
  const int jn = (g.J - j0 < GN) ? g.J - j0 : GN;          // valid columns of this tile
  const int n_mma = (jn + 15) & ~15;                        // UMMA N
  const int im = (g.I - i0 < GM) ? (int)(g.I - i0) : GM;    // valid rows of this tile
  const int nkt = (int)((ke - kb + GK - 1) / GK);

  const uint32_t smem0 = ptx::smem_u32(smem_raw);
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) ptx::mbar_init(ptx::smem_u32(&empty_bar[s]), 1);
    ptx::fence_mbar_init();
  }
  if (warp == 0) { ptx::tmem_alloc(ptx::smem_u32(&s_tmem), 256); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = s_tmem;

  typename StagePick<KCH, AKC, GM>::type ra;
  typename StagePick<KCH, BKC, GN>::type rb;
  const int64_t a_s = AKC ? g.sa_i : g.sa_k, b_s = BKC ? g.sb_j : g.sb_k;
  ra.fetch(g.A, a_s, g.I, i0, kb, ke, GM, tid);            // all 128 A rows are read by the MMA: zero-fill beyond I
  rb.fetch(g.B, b_s, g.J, j0, kb, ke, n_mma, tid);
  const uint32_t idesc = KCH == 8 ? ptx::idesc_bf16(n_mma) : idesc_tf32(n_mma);
  const bool want_rowsum = g.rowsum != nullptr && blockIdx.y == 0;   // bias gradient: row sums of A (exact fp32 adds)
  float rsum0 = 0.f, rsum1 = 0.f;

  for (int kt = 0; kt < nkt; ++kt) {
    const int s = kt & 1;
    const uint32_t slot = smem0 + (uint32_t)s * kStage;
    if (kt >= kStages) gemm_wait(ptx::smem_u32(&empty_bar[s]), (uint32_t)((kt >> 1) - 1) & 1u, 1);   // UMMAs of stage kt-2 retired
    if (want_rowsum) ra.rowsums(rsum0, rsum1);
    ra.store(slot, slot + kAHalf, GM, tid);
    rb.store(slot + 2 * kAHalf, slot + 2 * kAHalf + kBHalf, n_mma, tid);
    if (kt + 1 < nkt) {
      const int64_t k0 = kb + (int64_t)(kt + 1) * GK;
      ra.fetch(g.A, a_s, g.I, i0, k0, ke, GM, tid);
      rb.fetch(g.B, b_s, g.J, j0, k0, ke, n_mma, tid);
    }
    ptx::fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      ptx::tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {   // one UMMA K step = two 16-byte chunks
        const uint64_t a_hi = ptx::smem_desc(slot + ks * (2 * GM * 16), GM * 16, 128);
        const uint64_t a_lo = ptx::smem_desc(slot + kAHalf + ks * (2 * GM * 16), GM * 16, 128);
        const uint64_t b_hi = ptx::smem_desc(slot + 2 * kAHalf + ks * (2 * GN * 16), GN * 16, 128);
        const uint64_t b_lo = ptx::smem_desc(slot + 2 * kAHalf + kBHalf + ks * (2 * GN * 16), GN * 16, 128);
        const uint32_t acc0 = (kt > 0 || ks > 0) ? 1u : 0u;
        if constexpr (KCH == 8) {
          ptx::mma_bf16_ss(tmem, a_lo, b_hi, idesc, acc0);
          ptx::mma_bf16_ss(tmem, a_hi, b_lo, idesc, 1u);
          ptx::mma_bf16_ss(tmem, a_hi, b_hi, idesc, 1u);
        } else {
          mma_tf32_ss(tmem, a_lo, b_hi, idesc, acc0);
          mma_tf32_ss(tmem, a_hi, b_lo, idesc, 1u);
          mma_tf32_ss(tmem, a_hi, b_hi, idesc, 1u);
        }
      }
      ptx::mma_commit(ptx::smem_u32(&empty_bar[s]));
    }
  }
  if (want_rowsum) {   // 4 threads (the 4 K chunks) hold partial sums of each row
    const int r0 = decltype(ra)::row_of(0, tid), r1 = decltype(ra)::row_of(1, tid);
    if (r0 < im) atomicAdd(g.rowsum + i0 + r0, rsum0);
    if (r1 < im) atomicAdd(g.rowsum + i0 + r1, rsum1);
  }
  // the last commit covers every UMMA issued before it
  gemm_wait(ptx::smem_u32(&empty_bar[(nkt - 1) & 1]), (uint32_t)((nkt - 1) >> 1) & 1u, 2);
  ptx::tc_fence_after();

  gemm_epilogue(g, tmem, smem0, warp, lane, i0, j0, im, jn, n_mma, gridDim.z > 1, blockIdx.z == 0);
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem, 256); }
}

// ---- both operands row-contiguous activations, long K (dW = dZ^T H over 524 288 samples): tensor-copy-fed variant.
// One CTA per SM owns a 256 x 256 output tile (two 128-row accumulators = all 512 TMEM columns, so both halves share one
// staging of the B operand) and one K slice.  A producer lane lands fp32 steps of 16 k x 256 features of A and of B
// (2 x 16 KB, plain [k][feature] rows: features past I / J and samples past K arrive as zeros) in a 4-slot ring; 16 worker
// warps read their (row, 8-k chunk) with eight conflict-free 4-byte shared loads - the transposition into the K-major
// images is that read pattern, no global load touches the LSU -, split to bf16 head + residual and store 16-byte chunks
// into a 3-slot image ring; the issue warp runs 6 UMMAs (2 row halves x 3 products, N = 256, K = 16) per step.
// Measured before it (tools/gemm_bench.py dW, register-staged kernel, 0.31 ms): without the UMMAs 0.30, without the loads
// 0.15, without either 0.12 - the exposed load latency of a one-stage register prefetch was half the kernel.
constexpr int kRcWorkers = 512;
struct RcRing {
  static constexpr int RM = 256;                               // rows of A (and of B) per CTA
  static constexpr int SK = 16;                                // k per step = one bf16 UMMA
  static constexpr uint32_t kRawOp = SK * RM * 4;              // one operand's fp32 step: 16 KB
  static constexpr uint32_t kRawBytes = 2 * kRawOp;            // A then B
  static constexpr uint32_t kPart = 2 * RM * 16;               // hi or lo image of one operand: 2 chunks x 256 rows x 16 B = 8 KB
  static constexpr uint32_t kImgBytes = 4 * kPart;             // A_hi | A_lo | B_hi | B_lo
  static constexpr int NI = 3, NR = 4;
  static constexpr uint32_t kRawRing = NI * kImgBytes;
  static constexpr uint32_t kBytes = kRawRing + NR * kRawBytes;
  static constexpr uint32_t kTail = 256;
  static_assert(kBytes + kTail <= 232448, "one CTA per SM: 227 KB");
  static_assert(16 * 32 * kEpiRow <= kRawRing, "epilogue staging of 16 warps must fit in the image ring");
};

__global__ void __launch_bounds__(kRcWorkers + 64, 1) tc_gemm_rc_tma_kernel(GemmArgs g, int64_t kper,
                                                                           const __grid_constant__ CUtensorMap amap,
                                                                           const __grid_constant__ CUtensorMap bmap) {
  using R = RcRing;
  constexpr int NI = R::NI, NR = R::NR;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint64_t* const img_empty = reinterpret_cast<uint64_t*>(smem_raw + R::kBytes);
  uint64_t* const img_ready = img_empty + NI;
  uint64_t* const raw_full = img_ready + NI;
  uint64_t* const raw_empty = raw_full + NR;
  uint32_t& s_tmem = *reinterpret_cast<uint32_t*>(raw_empty + NR);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t i0 = (int64_t)blockIdx.x * R::RM;
  const int j0 = (int)blockIdx.y * R::RM;
  const int64_t kb = (int64_t)blockIdx.z * kper;
  const int64_t ke = (kb + kper < g.K) ? kb + kper : g.K;
  if (kb >= ke) return;                                        // empty K slice; uniform for the CTA
  const int jn = (g.J - j0 < R::RM) ? g.J - j0 : R::RM;
  const int n_mma = (jn + 15) & ~15;
  const int im = (g.I - i0 < R::RM) ? (int)(g.I - i0) : R::RM;
  const int nkt = (int)((ke - kb + R::SK - 1) / R::SK);

  const uint32_t smem0 = ptx::smem_u32(smem_raw);
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < NI; ++s) {
      ptx::mbar_init(ptx::smem_u32(&img_empty[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&img_ready[s]), kRcWorkers / 32);
    }
#pragma unroll
    for (int s = 0; s < NR; ++s) {
      ptx::mbar_init(ptx::smem_u32(&raw_full[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&raw_empty[s]), kRcWorkers / 32);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 16) { ptx::tmem_alloc(ptx::smem_u32(&s_tmem), 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = s_tmem;

  if (warp == 16) {
    // issue warp: converged, slot numbers are immediates (loop unrolled by NI)
    const uint32_t idesc = ptx::idesc_bf16(n_mma);
    constexpr uint64_t kHi = (uint64_t)((128u >> 4) | (1u << 14)) << 32;                     // SBO = 128 B, descriptor version 1
    constexpr uint32_t kLbo = ((uint32_t)(R::RM * 16) >> 4) << 16;                            // next 16-byte K chunk: 256 rows on
    const uint32_t lo0 = (smem0 >> 4) & 0x3FFFu;
    const bool two = im > GM;                                                                 // rows 128.. exist
    for (int kt0 = 0; kt0 < nkt; kt0 += NI) {
#pragma unroll
      for (int u = 0; u < NI; ++u) {
        const int kt = kt0 + u;
        if (kt >= nkt) break;
        gemm_wait(ptx::smem_u32(&img_ready[u]), (uint32_t)(kt / NI) & 1u, 4);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          const uint32_t img = lo0 + (uint32_t)u * (R::kImgBytes >> 4);
          const uint64_t b_hi = kHi | (uint64_t)((img + ((2 * R::kPart) >> 4)) | kLbo);
          const uint64_t b_lo = kHi | (uint64_t)((img + ((3 * R::kPart) >> 4)) | kLbo);
          const uint32_t acc0 = kt > 0 ? 1u : 0u;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (h == 1 && !two) break;
            const uint64_t a_hi = kHi | (uint64_t)((img + h * ((GM * 16) >> 4)) | kLbo);
            const uint64_t a_lo = kHi | (uint64_t)((img + (R::kPart >> 4) + h * ((GM * 16) >> 4)) | kLbo);
            ptx::mma_bf16_ss(tmem + h * 256, a_lo, b_hi, idesc, acc0);
            ptx::mma_bf16_ss(tmem + h * 256, a_hi, b_lo, idesc, 1u);
            ptx::mma_bf16_ss(tmem + h * 256, a_hi, b_hi, idesc, 1u);
          }
          ptx::mma_commit(ptx::smem_u32(&img_empty[u]));
        }
        __syncwarp();
      }
    }
  } else if (warp == 17) {
    if (lane == 0) {
      ptx::prefetch_tensormap(&amap);
      ptx::prefetch_tensormap(&bmap);
      for (int kt = 0; kt < nkt; ++kt) {
        const int rs = kt % NR;
        if (kt >= NR) gemm_wait(ptx::smem_u32(&raw_empty[rs]), (uint32_t)(kt / NR - 1) & 1u, 7);
        const uint32_t bar = ptx::smem_u32(&raw_full[rs]);
        const uint32_t raw = smem0 + R::kRawRing + (uint32_t)rs * R::kRawBytes;
        const int k0 = (int)(kb + (int64_t)kt * R::SK);
        ptx::mbar_arrive_expect_tx(bar, R::kRawBytes);
        ptx::tma_load_2d(raw, &amap, (int)i0, k0, bar);
        ptx::tma_load_2d(raw + R::kRawOp, &bmap, j0, k0, bar);
      }
    }
  } else {
    // workers: thread = (row of A and of B, one of the step's two 8-k chunks)
    const int row = tid & (R::RM - 1), c = tid >> 8;
    const bool want_rowsum = g.rowsum != nullptr && blockIdx.y == 0;   // bias gradient: row sums of A (exact fp32 adds)
    float rsum = 0.f;
    const uint32_t rd = (uint32_t)(8 * c) * (R::RM * 4) + (uint32_t)row * 4;
    const uint32_t wr = (uint32_t)c * (R::RM * 16) + (uint32_t)row * 16;
    for (int kt = 0; kt < nkt; ++kt) {
      const int rs = kt % NR, si = kt % NI;
      const uint32_t raw = smem0 + R::kRawRing + (uint32_t)rs * R::kRawBytes + rd;
      const uint32_t img = smem0 + (uint32_t)si * R::kImgBytes + wr;
      gemm_wait(ptx::smem_u32(&raw_full[rs]), (uint32_t)(kt / NR) & 1u, 6);
      float a[8], b[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        a[e] = ptx::ld_smem_f32(raw + (uint32_t)e * (R::RM * 4));
        b[e] = ptx::ld_smem_f32(raw + R::kRawOp + (uint32_t)e * (R::RM * 4));
      }
      if (want_rowsum) {
#pragma unroll
        for (int e = 0; e < 8; ++e) rsum += a[e];
      }
      if (kt >= NI) gemm_wait(ptx::smem_u32(&img_empty[si]), (uint32_t)(kt / NI - 1) & 1u, 1);
      store_chunk<8>(img, img + R::kPart, a);
      store_chunk<8>(img + 2 * R::kPart, img + 3 * R::kPart, b);
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        ptx::mbar_arrive(ptx::smem_u32(&img_ready[si]));
        ptx::mbar_arrive(ptx::smem_u32(&raw_empty[rs]));   // behind the stores that consumed the landing slot (see the packed kernel)
      }
    }
    if (want_rowsum && row < im) atomicAdd(g.rowsum + i0 + row, rsum);
    gemm_wait(ptx::smem_u32(&img_empty[(nkt - 1) % NI]), (uint32_t)((nkt - 1) / NI) & 1u, 2);   // covers every UMMA issued before it
    ptx::tc_fence_after();
    const int hw = warp >> 3;                                  // warps 0-7: rows 0..127, warps 8-15: rows 128..255
    const int im_h = im - hw * GM < GM ? im - hw * GM : GM;
    if (im_h > 0)
      gemm_epilogue(g, tmem + (uint32_t)hw * 256, smem0 + (uint32_t)hw * (8 * 32 * kEpiRow), warp & 7, lane, i0 + hw * GM, j0, im_h, jn,
                    n_mma, gridDim.z > 1, blockIdx.z == 0);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 16) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem, 512); }
}

// ---- weights as the B operand (forward, dX): packed once per call into the stage images (hi | lo per TN-row tile and
// K stage, zero padded) by a tiny kernel, then streamed by the TMA engine: no registers, no thread work, stages ahead.
template <int KCH, int TN>
__global__ void gemm_pack_b_kernel(const float* __restrict__ B, int64_t sb_j, int64_t sb_k, int J, int64_t K, int nst, int64_t total,
                                   uint8_t* __restrict__ out) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int row = (int)(idx % TN), c = (int)((idx / TN) & 3);
  const int64_t sidx = idx / (4 * TN);             // tile * nst + stage
  const int st = (int)(sidx % nst);
  const int64_t j = (sidx / nst) * TN + row, k = (int64_t)st * (4 * KCH) + c * KCH;
  float v[KCH];
#pragma unroll
  for (int e = 0; e < KCH; ++e) v[e] = (j < J && k + e < K) ? __ldg(B + j * sb_j + (k + e) * sb_k) : 0.f;
  uint32_t h[4], l[4];
  split_chunk<KCH>(v, h, l);
  uint8_t* dst = out + sidx * (2 * TN * 64) + (size_t)c * (TN * 16) + (size_t)row * 16;
  *reinterpret_cast<uint4*>(dst) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(dst + TN * 64) = make_uint4(l[0], l[1], l[2], l[3]);
}

constexpr int kWorkers = 256;              // warps 0..7 stage A and run the epilogue; warp 8 drives the TMA and the UMMAs

#ifdef ZEST_GEMM_TIMELINE     // developer build: clock64 stamps of CTAs 1000..1007 of the packed kernel (tools/gemm_timeline.py)
constexpr int kTlLaunches = 512;
__device__ unsigned long long g_gemm_tl[kTlLaunches * 8 * 128];
__device__ int g_gemm_tl_launch;      // which launch of the packed kernel is running (set stream-ordered by the host)
#define GTL(slot, cond) do { if (blockIdx.x >= 1000 && blockIdx.x < 1008 && (cond)) g_gemm_tl[((size_t)g_gemm_tl_launch * 8 + (blockIdx.x - 1000)) * 128 + (slot)] = clock64(); } while (0)
int g_tl_count = 0;
int g_tl_meta[kTlLaunches][8];
#else
#define GTL(slot, cond) do { } while (0)
#endif

// DUAL = false: 128 x 256 tile, one accumulator (256 TMEM columns), 2 stages of 48 KB.
// DUAL = true:  128 x 128 tile; the head product a_hi*b_hi accumulates in TMEM columns [0,128), the two cross products in
//               [128,256), added in the epilogue in fp32; 3 stages of 32 KB.  The tensor core's fp32 accumulate truncates,
//               a bias that compounds layer by layer through the MLP's forward / dX chain; the main accumulator now sees a
//               third of the UMMAs and the cross accumulator is 2^-11 of its magnitude.
// Either way 96 KB of shared memory and 256 TMEM columns: two CTAs per SM, one's epilogue under the other's main loop.
// Ring geometry of the packed kernel.  Three rings, sized separately:
//   A ring    NSA slots of hi + lo images (16 KB), written by the workers, read by the UMMAs
//   B ring    NSB slots of packed weights (16 / 32 KB), bulk copies out of L2 issued NSB - 1 stages ahead by the issue warp
//   landing   NR slots of one fp32 A stage (8 KB) as a 2-D tensor copy lands it (TMA_A launches): a tenth warp streams the row
//             tile NR stages ahead; the workers read LDS -> split -> STS.  Without it (operand not k-contiguous / not 16-byte
//             aligned) the workers stage A through registers, up to 6 stages ahead.
// Measured on the forward layer GEMM [524288,256] x [256,256], engine 2 (tools/gemm_timeline.py, stamps per stage): with
// register staging a stage took ~2200 cycles while the loads of stage + 6 were being issued and ~1000 without - the LSU path
// (8 half-lines per LDG.128, 2 CTAs x 96 KB in flight) was the bound, not HBM (39 %), the tensor pipe (26 %) or the ALU split
// (a 3x cheaper split changed nothing).  Tensor copies take the loads off the LSU: 0.615 -> 0.440 ms.
// The ring split 2 A + 4 B (vs 3 + 3) was worth 2 % on the register path; rings above 113 KB lose the second CTA and are slower.
// Tensor-copy path, same-box A/B of the whole fine-tune step (A slots / B slots / landing slots, all 112 KB): 2 / 3 / 4 = 74.4 ms
// (shipped), 2 / 3 / 3 = 77.6, 2 / 4 / 2 = 80.4, 3 / 2 / 4 = 79.2, 2 / 2 / 6 = 79.6: both the weight copies and the landing ring
// want depth; a third A slot buys nothing.
#ifndef ZEST_GEMM_NSA_DUAL
#define ZEST_GEMM_NSA_DUAL 2
#endif
#ifndef ZEST_GEMM_NSB_DUAL
#define ZEST_GEMM_NSB_DUAL 4
#endif
template <bool DUAL, bool TMA_A>
struct PackedRing {
  static constexpr int TN = DUAL ? 128 : 256;                // output columns per CTA
  static constexpr uint32_t kBH = TN * 64;                   // bytes of one part (hi or lo) of a B stage
  static constexpr uint32_t kABytes = 2 * kAHalf, kBBytes = 2 * kBH;
  static constexpr uint32_t kRawBytes = GM * 64;             // one fp32 stage of A as the tensor copy lands it: 128 rows x 64 B
  static constexpr int NSA = DUAL ? ZEST_GEMM_NSA_DUAL : 2;
#ifndef ZEST_GEMM_NSB_TMA
#define ZEST_GEMM_NSB_TMA 3
#endif
#ifndef ZEST_GEMM_NR
#define ZEST_GEMM_NR 4
#endif
  static constexpr int NSB = DUAL ? (TMA_A ? ZEST_GEMM_NSB_TMA : ZEST_GEMM_NSB_DUAL) : 2;
  static constexpr int NR = TMA_A ? ZEST_GEMM_NR : 0;                   // fp32 landing slots of the A tensor copies
  static constexpr uint32_t kBRing = NSA * kABytes, kRawRing = kBRing + NSB * kBBytes;
  static constexpr uint32_t kBytes = kRawRing + NR * kRawBytes;
  static constexpr uint32_t kTail = 256;                     // barriers + TMEM base behind the rings
  static constexpr int gcd(int a, int b) { return b == 0 ? a : gcd(b, a % b); }
  static constexpr int U = NSA * NSB / gcd(NSA, NSB);        // the issue loop is unrolled by this: slot numbers are immediates
  static_assert(kBytes + kTail + 1024 <= 233472 / 2, "two CTAs per SM (228 KB, 1 KB reserved per CTA)");
  static_assert(8 * 32 * kEpiRow <= kRawRing, "epilogue staging must fit in the A and B rings");
};

template <int KCH, bool AKC, bool DUAL, bool TMA_A>
__global__ void __launch_bounds__(kWorkers + 64, 2) tc_gemm_packed_kernel(GemmArgs g, const uint8_t* __restrict__ bpack,
                                                                          const __grid_constant__ CUtensorMap amap) {
  static_assert(!TMA_A || (AKC && KCH == 4), "the A tensor copies land fp32 stages of 16 k");
  using Ring = PackedRing<DUAL, TMA_A>;
  constexpr int GK = 4 * KCH;
  constexpr int TN = Ring::TN;
  constexpr uint32_t kBH = Ring::kBH;
  constexpr int NSA = Ring::NSA, NSB = Ring::NSB, U = Ring::U;
  constexpr int PD = NSB - 1;                               // the weight copies run this many stages ahead of the UMMAs
  constexpr uint32_t kBRing = Ring::kBRing;                 // byte offset of the B ring
  constexpr int NR = TMA_A ? Ring::NR : 1;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // barriers and the TMEM base live behind the rings in the dynamic allocation (a static __shared__ block next to a
  // 1024-byte-aligned extern array costs 2 KB of padding, which is what decides whether two CTAs fit)
  uint64_t* const empty_bar = reinterpret_cast<uint64_t*>(smem_raw + Ring::kBytes);
  uint64_t* const full_bar = empty_bar + NSA;
  uint64_t* const aready_bar = full_bar + NSB;
  uint64_t* const raw_full = aready_bar + NSA;
  uint64_t* const raw_empty = raw_full + NR;
  uint32_t& s_tmem = *reinterpret_cast<uint32_t*>(raw_empty + NR);
  static_assert((2 * NSA + NSB + 2 * NR + 1) * 8 <= (int)Ring::kTail, "barrier block");
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // column tiles of one row tile are adjacent CTAs: the second reader of an A tile hits L2
  const int tiles_n = (g.J + TN - 1) / TN;
  const int bn = (int)(blockIdx.x % tiles_n);
  const int64_t i0 = (int64_t)(blockIdx.x / tiles_n) * GM;
  const int j0 = bn * TN;
  const int jn = (g.J - j0 < TN) ? g.J - j0 : TN;
  const int n_mma = (jn + 15) & ~15;
  const int im = (g.I - i0 < GM) ? (int)(g.I - i0) : GM;
  const int nkt = (int)((g.K + GK - 1) / GK);

  const uint32_t smem0 = ptx::smem_u32(smem_raw);
  GTL(0, tid == 0);
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < NSA; ++s) {
      ptx::mbar_init(ptx::smem_u32(&empty_bar[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&aready_bar[s]), kWorkers / 32);
    }
#pragma unroll
    for (int s = 0; s < NSB; ++s) ptx::mbar_init(ptx::smem_u32(&full_bar[s]), 1);
#pragma unroll
    for (int s = 0; s < NR; ++s) {
      ptx::mbar_init(ptx::smem_u32(&raw_full[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&raw_empty[s]), kWorkers / 32);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 8) { ptx::tmem_alloc(ptx::smem_u32(&s_tmem), 256); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = s_tmem;
  GTL(1, tid == 0);

  if (warp == 8) {
    // The issue warp stays converged: every lane polls the barriers, one elected lane issues.  Slot numbers are compile-time
    // (the stage loop is unrolled by U = lcm(NSA, NSB)), so every descriptor is a warp-uniform base plus an immediate - no
    // per-UMMA register-to-uniform waterfall (measured ~90 cycles per UMMA in the MLP kernel) on the issue path.
    const uint8_t* src = bpack + (size_t)bn * nkt * (2 * kBH);
    const uint32_t idesc = KCH == 8 ? ptx::idesc_bf16(n_mma) : idesc_tf32(n_mma);
    constexpr uint64_t kHi = (uint64_t)((128u >> 4) | (1u << 14)) << 32;                    // SBO = 128 B, descriptor version 1
    constexpr uint32_t kALbo = ((uint32_t)(GM * 16) >> 4) << 16, kBLbo = ((uint32_t)(TN * 16) >> 4) << 16;
    const uint32_t lo0 = (smem0 >> 4) & 0x3FFFu;
    const uint32_t cross = DUAL ? tmem + TN : tmem;
    auto load_b = [&](int t, int slot_i) {   // weights of stage t -> B slot t % NSB (= slot_i)
      const uint32_t bar = ptx::smem_u32(&full_bar[slot_i]);
      ptx::mbar_arrive_expect_tx(bar, 2 * kBH);
      ptx::bulk_g2s(smem0 + kBRing + (uint32_t)slot_i * Ring::kBBytes, src + (size_t)t * (2 * kBH), 2 * kBH, bar);
    };
    if (ptx::elect_one()) {
#pragma unroll
      for (int t = 0; t < PD; ++t)
        if (t < nkt) load_b(t, t);
    }
    __syncwarp();
    for (int kt0 = 0; kt0 < nkt; kt0 += U) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int kt = kt0 + u;
        if (kt >= nkt) break;
        const int sa = u % NSA, sb = u % NSB;               // kt0 is a multiple of both ring sizes
        gemm_wait(ptx::smem_u32(&aready_bar[sa]), (uint32_t)(kt / NSA) & 1u, 4);
        gemm_wait(ptx::smem_u32(&full_bar[sb]), (uint32_t)(kt / NSB) & 1u, 5);
        ptx::tc_fence_after();
        GTL(kt == 0 ? 6 : 7, lane == 0);      // first stage ready / every later stage ready (the last write = the last stage)
        GTL(16 + kt * 4 + 3, lane == 0 && kt < 24);
        if (ptx::elect_one()) {
          const uint32_t a_slot = lo0 + (uint32_t)sa * (Ring::kABytes >> 4);
          const uint32_t b_slot = lo0 + (kBRing >> 4) + (uint32_t)sb * (Ring::kBBytes >> 4);
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            const uint64_t a_hi = kHi | (uint64_t)((a_slot + ks * ((2 * GM * 16) >> 4)) | kALbo);
            const uint64_t a_lo = kHi | (uint64_t)((a_slot + (kAHalf >> 4) + ks * ((2 * GM * 16) >> 4)) | kALbo);
            const uint64_t b_hi = kHi | (uint64_t)((b_slot + ks * ((2 * TN * 16) >> 4)) | kBLbo);
            const uint64_t b_lo = kHi | (uint64_t)((b_slot + (kBH >> 4) + ks * ((2 * TN * 16) >> 4)) | kBLbo);
            const uint32_t acc0 = (kt > 0 || ks > 0) ? 1u : 0u;
            if constexpr (KCH == 8) {
              ptx::mma_bf16_ss(cross, a_lo, b_hi, idesc, acc0);
              ptx::mma_bf16_ss(cross, a_hi, b_lo, idesc, 1u);
              ptx::mma_bf16_ss(tmem, a_hi, b_hi, idesc, DUAL ? acc0 : 1u);
            } else {
              mma_tf32_ss(cross, a_lo, b_hi, idesc, acc0);
              mma_tf32_ss(cross, a_hi, b_lo, idesc, 1u);
              mma_tf32_ss(tmem, a_hi, b_hi, idesc, DUAL ? acc0 : 1u);
            }
          }
          ptx::mma_commit(ptx::smem_u32(&empty_bar[sa]));   // frees the A slot for the workers and (below) the B slot
        }
        __syncwarp();
        if (kt + PD < nkt) {   // stage kt + PD reuses the B slot of stage kt - 1: free once those UMMAs have retired
          const int pa = (u + U - 1) % NSA, pb = (u + U - 1) % NSB;
          if (kt >= 1) gemm_wait(ptx::smem_u32(&empty_bar[pa]), (uint32_t)((kt - 1) / NSA) & 1u, 3);
          if (ptx::elect_one()) load_b(kt + PD, pb);
          __syncwarp();
        }
      }
    }
  } else if (warp == 9) {
    // A producer (TMA_A launches only): one lane streams the row tile's fp32 stages into the landing ring, NR - 1 ahead of
    // the workers.  Rows past I and columns past K arrive as zeros (the map's bounds), the box always counts 8 KB.
    if constexpr (TMA_A) {
      if (lane == 0) {
        ptx::prefetch_tensormap(&amap);
        for (int kt = 0; kt < nkt; ++kt) {
          const int rs = kt % NR;
          if (kt >= NR) gemm_wait(ptx::smem_u32(&raw_empty[rs]), (uint32_t)(kt / NR - 1) & 1u, 7);
          const uint32_t bar = ptx::smem_u32(&raw_full[rs]);
          ptx::mbar_arrive_expect_tx(bar, Ring::kRawBytes);
          ptx::tma_load_2d(smem0 + Ring::kRawRing + (uint32_t)rs * Ring::kRawBytes, &amap, kt * GK, (int)i0, bar);
        }
      }
    }
  } else if constexpr (TMA_A) {
    // workers: fp32 stage out of the landing ring (64-byte rows, 16-byte chunks XOR-swizzled by the copy: chunk ^ (row / 2) % 4,
    // which makes the 8 rows a quarter-warp reads hit 8 distinct bank groups) -> hi / lo images of the A slot
    GTL(2, tid == 0);
    for (int kt = 0; kt < nkt; ++kt) {
      const int rs = kt % NR, sa = kt % NSA;
      const uint32_t raw = smem0 + Ring::kRawRing + (uint32_t)rs * Ring::kRawBytes;
      const uint32_t slot = smem0 + (uint32_t)sa * Ring::kABytes;
      gemm_wait(ptx::smem_u32(&raw_full[rs]), (uint32_t)(kt / NR) & 1u, 6);
      uint4 w[2];
#pragma unroll
      for (int n = 0; n < 2; ++n) {
        const int u = tid + n * kThreads;
        const int row = (u >> 5) * 8 + (u & 7), c = (u >> 3) & 3;
        w[n] = ptx::ld_smem_v4(raw + (uint32_t)row * 64 + (uint32_t)((c ^ ((row >> 1) & 3)) << 4));
      }
      if (kt >= NSA) gemm_wait(ptx::smem_u32(&empty_bar[sa]), (uint32_t)(kt / NSA - 1) & 1u, 1);
      GTL(16 + kt * 4 + 0, tid == 0 && kt < 24);
#pragma unroll
      for (int n = 0; n < 2; ++n) {
        const int u = tid + n * kThreads;
        const int row = (u >> 5) * 8 + (u & 7), c = (u >> 3) & 3;
        const float v[4] = {__uint_as_float(w[n].x), __uint_as_float(w[n].y), __uint_as_float(w[n].z), __uint_as_float(w[n].w)};
        const uint32_t off = (uint32_t)c * (GM * 16) + (uint32_t)row * 16;
        store_chunk<KCH>(slot + off, slot + kAHalf + off, v);
      }
      GTL(16 + kt * 4 + 1, tid == 0 && kt < 24);
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        ptx::mbar_arrive(ptx::smem_u32(&aready_bar[sa]));
        // The landing slot is released only now, behind the stores that consumed its values.  Releasing it right after
        // the LDS was a measured bug: the UMMAs' operand reads can hold the shared-memory pipe long enough for the LDS to
        // execute after the refill of the slot had started (rows of stage kt + NR in the image of stage kt).
        ptx::mbar_arrive(ptx::smem_u32(&raw_empty[rs]));
      }
      GTL(16 + kt * 4 + 2, tid == 0 && kt < 24);
    }
    GTL(3, tid == 0);
    gemm_wait(ptx::smem_u32(&empty_bar[(nkt - 1) % NSA]), (uint32_t)((nkt - 1) / NSA) & 1u, 2);
    ptx::tc_fence_after();
    GTL(4, tid == 0);
    gemm_epilogue<DUAL ? TN : 0>(g, tmem, smem0, warp, lane, i0, j0, im, jn, n_mma, false, true);
    GTL(5, tid == 0);
  } else {
    // workers: A NSET stages ahead in registers (static register sets), no CTA-wide barrier in the loop
    constexpr int NSET = KCH == 4 ? 6 : 2;     // a tf32 stage is 8 registers per thread, a bf16 stage 16
    typename StagePick<KCH, AKC, GM>::type rs[NSET];
    const int64_t a_s = AKC ? g.sa_i : g.sa_k;
#pragma unroll
    for (int u = 0; u < NSET; ++u)
      if (u < nkt) rs[u].fetch(g.A, a_s, g.I, i0, (int64_t)u * GK, g.K, GM, tid);
    auto step = [&](auto& r, int kt) {
      const int s = kt % NSA;
      const uint32_t slot = smem0 + (uint32_t)s * Ring::kABytes;
      if (kt >= NSA) gemm_wait(ptx::smem_u32(&empty_bar[s]), (uint32_t)(kt / NSA - 1) & 1u, 1);
      GTL(16 + kt * 4 + 0, tid == 0 && kt < 24);
      r.store(slot, slot + kAHalf, GM, tid);
      if (kt + NSET < nkt) r.fetch(g.A, a_s, g.I, i0, (int64_t)(kt + NSET) * GK, g.K, GM, tid);
      GTL(16 + kt * 4 + 1, tid == 0 && kt < 24);
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&aready_bar[s]));
      GTL(16 + kt * 4 + 2, tid == 0 && kt < 24);
    };
    GTL(2, tid == 0);
    for (int kt = 0; kt < nkt; kt += NSET) {
#pragma unroll
      for (int u = 0; u < NSET; ++u)
        if (kt + u < nkt) step(rs[u], kt + u);
    }
    GTL(3, tid == 0);
    gemm_wait(ptx::smem_u32(&empty_bar[(nkt - 1) % NSA]), (uint32_t)((nkt - 1) / NSA) & 1u, 2);
    ptx::tc_fence_after();
    GTL(4, tid == 0);
    gemm_epilogue<DUAL ? TN : 0>(g, tmem, smem0, warp, lane, i0, j0, im, jn, n_mma, false, true);
    GTL(5, tid == 0);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 8) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem, 256); }
  GTL(8, tid == 0);
}

// A tensor map over a k-contiguous fp32 A[I, K] (row stride lda floats): boxes of 16 k x 128 rows, 64-byte swizzle.
// False when the operand does not qualify (alignment) or the driver entry point is missing: the caller keeps the register path.
using TensorMapEncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                       const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                       CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
TensorMapEncodeFn tensor_map_encoder() {
  static const TensorMapEncodeFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<TensorMapEncodeFn>(p);
  }();
  return fn;
}
// fp32 matrix [outer, inner] with unit stride along inner and `ld` floats between outer indices -> boxes of box_inner x box_outer
bool encode_map_2d(CUtensorMap* map, const float* base, int64_t inner, int64_t outer, int64_t ld, int box_inner, int box_outer,
                   CUtensorMapSwizzle swizzle) {
  if (std::getenv("ZEST_GEMM_NO_TMA_A")) return false;           // developer A/B switch: keep the register-staged kernels
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld & 3) != 0 || ld < inner) return false;
  if (inner >= (1ll << 31) || outer >= (1ll << 31) || ld * 4 >= (1ll << 40)) return false;
  const TensorMapEncodeFn enc = tensor_map_encoder();
  if (!enc) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  const cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  const cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
bool encode_a_map(const GemmArgs& a, CUtensorMap* map) {   // k-contiguous A[I, K]: 16 k x 128 rows, 64-byte swizzle
  return a.sa_k == 1 && encode_map_2d(map, a.A, a.K, a.I, a.sa_i, 16, GM, CU_TENSOR_MAP_SWIZZLE_64B);
}

template <int KCH, bool AKC, bool DUAL, bool TMA_A>
int launch_packed_variant(const GemmArgs& a, dim3 grid, const CUtensorMap& amap, cudaStream_t st) {
  // the attribute is per device: set it on every launch (a few hundred ns) rather than once per process
  constexpr int kBytes = (int)(PackedRing<DUAL, TMA_A>::kBytes + PackedRing<DUAL, TMA_A>::kTail);
#ifdef ZEST_GEMM_TIMELINE
  {
    const int slot = g_tl_count % kTlLaunches;
    const int meta[8] = {(int)grid.x, a.Z != nullptr, a.gate != nullptr, a.gb_dZ != nullptr, a.accumulate, a.J, (int)a.K, TMA_A ? 1 : 0};
    std::memcpy(g_tl_meta[slot], meta, sizeof(meta));
    ZEST_CUDA(cudaMemcpyToSymbolAsync(g_gemm_tl_launch, &slot, sizeof(int), 0, cudaMemcpyHostToDevice, st));
    ++g_tl_count;
  }
#endif
  ZEST_CUDA(cudaFuncSetAttribute(tc_gemm_packed_kernel<KCH, AKC, DUAL, TMA_A>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBytes));
  tc_gemm_packed_kernel<KCH, AKC, DUAL, TMA_A><<<grid, kWorkers + (TMA_A ? 64 : 32), kBytes, st>>>(a, (const uint8_t*)a.b_scratch, amap);
  ZEST_LAUNCH_CHECK();
  return ZEST_OK;
}

// bytes of packed weight images a GEMM needs
template <int KCH, bool DUAL>
int64_t packed_bytes(const GemmArgs& a) {
  constexpr int TN = DUAL ? 128 : 256;
  return ((a.J + TN - 1) / TN) * ((a.K + 4 * KCH - 1) / (4 * KCH)) * (int64_t)(2 * TN * 64);
}

template <int KCH, bool DUAL>
int launch_packed(const GemmArgs& a, cudaStream_t st) {
  constexpr int GK = 4 * KCH, TN = DUAL ? 128 : 256;
  const int nst = (int)((a.K + GK - 1) / GK);
  const int64_t tiles_m = (a.I + GM - 1) / GM, tiles_n = (a.J + TN - 1) / TN;
  ZEST_CHECK_ARG(tiles_m * tiles_n < (1ll << 31), "tc gemm: shape too large for one launch");
  dim3 grid((unsigned)(tiles_m * tiles_n), 1, 1);
  const int64_t total = tiles_n * nst * 4 * TN;
  gemm_pack_b_kernel<KCH, TN><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(a.B, a.sb_j, a.sb_k, a.J, a.K, nst, total, (uint8_t*)a.b_scratch);
  ZEST_LAUNCH_CHECK();
  CUtensorMap amap;
  std::memset(&amap, 0, sizeof(amap));
  if constexpr (KCH == 4) {   // k-contiguous fp32 activations: tensor copies feed the A stages, no loads through the LSU
    if (encode_a_map(a, &amap)) return launch_packed_variant<KCH, true, DUAL, true>(a, grid, amap, st);
  }
  return a.sa_k == 1 ? launch_packed_variant<KCH, true, DUAL, false>(a, grid, amap, st)
                     : launch_packed_variant<KCH, false, DUAL, false>(a, grid, amap, st);
}

template <int KCH, bool AKC, bool BKC>
int launch_variant(const GemmArgs& a, dim3 grid, int64_t kper, cudaStream_t st) {
  ZEST_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<KCH, AKC, BKC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem));
  tc_gemm_kernel<KCH, AKC, BKC><<<grid, kThreads, kSmem, st>>>(a, kper);
  ZEST_LAUNCH_CHECK();
  return ZEST_OK;
}
// dW-shaped GEMM (both operands row-contiguous, bf16 split): tensor-copy-fed 256 x 256 tiles, one CTA per SM.  Returns
// ZEST_OK + launched = true, or launched = false when the operands do not qualify (the caller keeps the register kernel).
int launch_rc_tma(const GemmArgs& a0, bool* launched, cudaStream_t st) {
  *launched = false;
  GemmArgs a = a0;
  if (a.sa_i != 1 || a.sb_j != 1 || a.K >= (1ll << 31)) return ZEST_OK;
  CUtensorMap amap, bmap;
  if (!encode_map_2d(&amap, a.A, a.I, a.K, a.sa_k, RcRing::RM, RcRing::SK, CU_TENSOR_MAP_SWIZZLE_NONE) ||
      !encode_map_2d(&bmap, a.B, a.J, a.K, a.sb_k, RcRing::RM, RcRing::SK, CU_TENSOR_MAP_SWIZZLE_NONE))
    return ZEST_OK;
  const int64_t ti = (a.I + RcRing::RM - 1) / RcRing::RM, tj = (a.J + RcRing::RM - 1) / RcRing::RM, tiles = ti * tj;
  ZEST_CHECK_ARG(ti < (1ll << 31) && tj < 65536, "tc gemm: shape too large for one launch");
  int64_t splits = 1;
  if (a.splits > 1) {
    // <= 192 chained UMMAs per accumulator (the fp32 accumulate truncates: see launch_gemm_tc) = 1024 k per CTA, and whole
    // waves of one CTA per SM
    const int64_t min_splits = (a.K + 1023) / 1024, sms = num_sms();
    const int64_t waves = (min_splits * tiles + sms - 1) / sms;
    splits = waves * sms / tiles;
    if (splits < min_splits) splits = min_splits;
    const int64_t max_splits = (a.K + 255) / 256;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    a.accumulate = 1;
  }
  int64_t kper = (a.K + splits - 1) / splits;
  kper = (kper + 31) / 32 * 32;
  splits = (a.K + kper - 1) / kper;
  ZEST_CHECK_ARG(splits == 1 || (!a.Z && !a.gate && !a.relu && !a.gb_dZ), "tc gemm: split-K cannot fuse a non-linear epilogue");
  if (splits >= 65536) return ZEST_OK;   // grid.z limit: leave it to the register kernel's own split policy
  constexpr int kBytes = (int)(RcRing::kBytes + RcRing::kTail);
  ZEST_CUDA(cudaFuncSetAttribute(tc_gemm_rc_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBytes));
  tc_gemm_rc_tma_kernel<<<dim3((unsigned)ti, (unsigned)tj, (unsigned)splits), kRcWorkers + 64, kBytes, st>>>(a, kper, amap, bmap);
  ZEST_LAUNCH_CHECK();
  *launched = true;
  return ZEST_OK;
}

template <int KCH>
int launch_prec(const GemmArgs& a, dim3 grid, int64_t kper, cudaStream_t st) {
  const bool akc = a.sa_k == 1, bkc = a.sb_k == 1;
  if (akc && bkc) return launch_variant<KCH, true, true>(a, grid, kper, st);
  if (akc && !bkc) return launch_variant<KCH, true, false>(a, grid, kper, st);
  if (!akc && bkc) return launch_variant<KCH, false, true>(a, grid, kper, st);
  return launch_variant<KCH, false, false>(a, grid, kper, st);
}

std::atomic<int> g_engine{-1};

}  // namespace

// 0 = exact-fp32 CUDA cores, 1 = tcgen05 3 x bf16 (fastest), 2 = tcgen05 3 x tf32 with split accumulators (fp32-grade, default)
int gemm_engine() {
  int e = g_engine.load(std::memory_order_relaxed);
  if (e < 0) {
    const char* s = getenv("ZEST_GEMM");
    e = 2;
    if (s && (!strcmp(s, "simt") || !strcmp(s, "0"))) e = 0;
    else if (s && (!strcmp(s, "bf16x3") || !strcmp(s, "1"))) e = 1;
    g_engine.store(e, std::memory_order_relaxed);
  }
  return e;
}
void set_gemm_engine(int e) { g_engine.store(e < 0 ? 0 : (e > 2 ? 2 : e), std::memory_order_relaxed); }

bool gemm_tc_supported(const GemmArgs& a) {
  return (a.sa_k == 1 || a.sa_i == 1) && (a.sb_k == 1 || a.sb_j == 1) && a.K >= 1 && a.J >= 1;
}

int launch_gemm_tc(const GemmArgs& a0, int engine, cudaStream_t st) {
  GemmArgs a = a0;
  ZEST_CHECK_ARG(gemm_tc_supported(a), "tc gemm: operands need unit stride along one dimension");
  ZEST_CHECK_ARG(engine == 1 || engine == 2, "tc gemm: engine must be 1 (3 x bf16) or 2 (3 x tf32)");
  if (a.I == 0) return ZEST_OK;
  // engine 2 keeps the long split-K reductions (dW) on the bf16 split: half the UMMAs per accumulator (less truncation
  // bias, measured) and they are leaves of the graph - nothing compounds through them
  const bool bf16 = engine == 1 || a.splits > 1;
  if (bf16 && a.sa_i == 1 && a.sb_j == 1) {   // dW: tensor-copy-fed kernel (narrow A too: the landed rows past I are zeros, 0.6 % of a step faster than the register kernel)
    bool launched = false;
    const int rc = launch_rc_tma(a, &launched, st);
    if (rc != ZEST_OK || launched) return rc;
  }
  const int GK = bf16 ? 32 : 16;
  const int64_t ti = (a.I + GM - 1) / GM, tj = (a.J + GN - 1) / GN;
  ZEST_CHECK_ARG(ti < (1ll << 31) && tj < 65536, "tc gemm: shape too large for one launch");
  int64_t splits = 1;
  if (a.splits > 1) {   // split-K: the caller only says "reduce over a long K"; fill the machine (2 CTAs / SM)
    // and keep one TMEM accumulator to <= 192 UMMAs: the tensor core's fp32 accumulate truncates, a bias that grows
    // linearly with the number of UMMAs chained into one accumulator; the cross-CTA atomics round to nearest
    splits = (2 * (int64_t)num_sms() + ti * tj - 1) / (ti * tj);
    const int64_t k_cta_max = 32 * GK, min_splits = (a.K + k_cta_max - 1) / k_cta_max;
    if (splits < min_splits) splits = min_splits;
    const int64_t max_splits = (a.K + 255) / 256;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    a.accumulate = 1;
  }
  int64_t kper = (a.K + splits - 1) / splits;
  kper = (kper + GK - 1) / GK * GK;
  splits = (a.K + kper - 1) / kper;
  ZEST_CHECK_ARG(splits == 1 || (!a.Z && !a.gate && !a.relu && !a.gb_dZ), "tc gemm: split-K cannot fuse a non-linear epilogue");
  dim3 grid((unsigned)ti, (unsigned)tj, (unsigned)splits);
  if (a.b_scratch && splits == 1 && ti >= 8 && !a.rowsum) {   // B is a small matrix re-read by every row tile (weights): pack + TMA
    if (engine == 1 && packed_bytes<8, false>(a) <= a.b_scratch_bytes) return launch_packed<8, false>(a, st);
    if (engine == 2 && packed_bytes<4, true>(a) <= a.b_scratch_bytes) return launch_packed<4, true>(a, st);
  }
  return bf16 ? launch_prec<8>(a, grid, kper, st) : launch_prec<4>(a, grid, kper, st);
}

}  // namespace zest

using namespace zest;

#ifdef ZEST_GEMM_TIMELINE
// stamps of the last min(count, 512) launches of the packed kernel: out[launch % 512][cta 0..7][128], meta[launch % 512][8] =
// {grid, has Z, has gate, has fused gate backward, accumulate, J, K, tensor-copy-fed}; returns the launch count
extern "C" int zest_gemm_read_timeline(unsigned long long* host_out, int* meta_out) {
  ZEST_CUDA(cudaDeviceSynchronize());
  ZEST_CUDA(cudaMemcpyFromSymbol(host_out, g_gemm_tl, sizeof(unsigned long long) * kTlLaunches * 8 * 128));
  std::memcpy(meta_out, g_tl_meta, sizeof(g_tl_meta));
  return g_tl_count;
}
#endif

extern "C" int zest_set_gemm_engine(int engine) {
  const int prev = gemm_engine();
  set_gemm_engine(engine);
  return prev;
}

extern "C" int zest_gemm_f32(const float* A, int64_t sa_i, int64_t sa_k, const float* B, int64_t sb_j, int64_t sb_k, float* C,
                             int64_t ldc, int64_t I, int J, int64_t K, const float* bias, int accumulate, int splits,
                             int engine, void* b_scratch, int64_t b_scratch_bytes, void* stream) {
  ZEST_CHECK_ARG(A && B && C && I >= 0 && J > 0 && K > 0 && splits >= 1, "zest_gemm_f32: bad arguments");
  ZEST_CHECK_ARG(engine >= 0 && engine <= 2, "zest_gemm_f32: engine must be 0, 1 or 2");
  GemmArgs a{};
  a.A = A; a.sa_i = sa_i; a.sa_k = sa_k;
  a.B = B; a.sb_j = sb_j; a.sb_k = sb_k;
  a.C = C; a.ldc = ldc; a.I = I; a.J = J; a.K = K;
  a.bias = bias; a.accumulate = accumulate; a.splits = splits;
  a.b_scratch = b_scratch; a.b_scratch_bytes = b_scratch ? b_scratch_bytes : 0;
  if (engine >= 1) return launch_gemm_tc(a, engine, (cudaStream_t)stream);
  return launch_gemm_simt(a, (cudaStream_t)stream);
}
