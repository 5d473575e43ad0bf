// fp32 SIMT GEMM: 128x128x8 block tile, 8x8 register micro-tile, register-prefetched smem stages.
#include "sgemm.cuh"

namespace zest {

constexpr int BM = 128, BN = 128, BK = 8, TM = 8, TN = 8;

// Load a [128 x 8] operand tile (rows r0.., reduction k0..) into smem as [BK][128].
// Two thread mappings so that the 4 elements a thread fetches are contiguous in memory.
struct TileLoader {
  const float* base; int64_t s_row, s_k; int64_t rows, K;
  bool k_contig;
  __device__ __forceinline__ void fetch(int64_t r0, int64_t k0, int64_t k_end, float (&v)[4], int t) const {
    if (k_contig) {
      const int64_t r = r0 + (t >> 1);
      const int64_t k = k0 + (t & 1) * 4;
      const float* p = base + r * s_row + k;
      if (r < rows && k + 3 < k_end && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(p));
        v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = (r < rows && k + e < k_end) ? __ldg(p + e) : 0.f;
      }
    } else {
      const int64_t k = k0 + (t >> 5);
      const int64_t r = r0 + (t & 31) * 4;
      const float* p = base + k * s_k + r * s_row;
      if (k < k_end && r + 3 < rows && s_row == 1 && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(p));
        v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = (k < k_end && r + e < rows) ? __ldg(p + e * s_row) : 0.f;
      }
    }
  }
  __device__ __forceinline__ void stash(float (*s)[BM], const float (&v)[4], int t) const {
    if (k_contig) {
      const int r = t >> 1, k = (t & 1) * 4;
#pragma unroll
      for (int e = 0; e < 4; ++e) s[k + e][r] = v[e];
    } else {
      const int k = t >> 5, r = (t & 31) * 4;
      *reinterpret_cast<float4*>(&s[k][r]) = make_float4(v[0], v[1], v[2], v[3]);
    }
  }
};

__global__ void __launch_bounds__(256) sgemm_kernel(GemmArgs g) {
  __shared__ __align__(16) float As[2][BK][BM];
  __shared__ __align__(16) float Bs[2][BK][BN];
  const int t = threadIdx.x;
  const int64_t i0 = (int64_t)blockIdx.x * BM;
  const int64_t j0 = (int64_t)blockIdx.y * BN;
  const int64_t kper = (g.K + g.splits - 1) / g.splits;
  const int64_t kb = (int64_t)blockIdx.z * kper;
  const int64_t ke = (kb + kper < g.K) ? kb + kper : g.K;
  TileLoader la{g.A, g.sa_i, g.sa_k, g.I, g.K, g.sa_k == 1};
  TileLoader lb{g.B, g.sb_j, g.sb_k, (int64_t)g.J, g.K, g.sb_k == 1};
  const int tx = t & 15, ty = t >> 4;
  float acc[TM][TN];
#pragma unroll
  for (int a = 0; a < TM; ++a)
#pragma unroll
    for (int b = 0; b < TN; ++b) acc[a][b] = 0.f;

  float va[4], vb[4];
  if (kb < ke) {
    la.fetch(i0, kb, ke, va, t);
    lb.fetch(j0, kb, ke, vb, t);
    la.stash(As[0], va, t);
    lb.stash(Bs[0], vb, t);
  }
  __syncthreads();
  int buf = 0;
  for (int64_t k0 = kb; k0 < ke; k0 += BK) {
    const bool more = k0 + BK < ke;
    if (more) {
      la.fetch(i0, k0 + BK, ke, va, t);
      lb.fetch(j0, k0 + BK, ke, vb, t);
    }
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      // micro-tile = rows {4 ty .. +3, 64 + 4 ty .. +3} x columns {4 tx .. +3, 64 + 4 tx .. +3}: the 16 lanes that differ in tx
      // read 16 consecutive float4 (conflict-free); an 8-wide contiguous column run per lane is a 2-way bank conflict
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int p = 0; p < TM; ++p)
#pragma unroll
        for (int q = 0; q < TN; ++q) acc[p][q] = fmaf(a[p], b[q], acc[p][q]);
    }
    if (more) {
      la.stash(As[buf ^ 1], va, t);
      lb.stash(Bs[buf ^ 1], vb, t);
    }
    __syncthreads();
    buf ^= 1;
  }

  const bool atomic = g.splits > 1;
#pragma unroll
  for (int p = 0; p < TM; ++p) {
    const int64_t i = i0 + (p < 4 ? ty * 4 + p : 64 + ty * 4 + (p - 4));
    if (i >= g.I) continue;
#pragma unroll
    for (int q = 0; q < TN; ++q) {
      const int64_t j = j0 + (q < 4 ? tx * 4 + q : 64 + tx * 4 + (q - 4));
      if (j >= g.J) continue;
      float v = acc[p][q];
      if (g.bias && blockIdx.z == 0) v += __ldg(g.bias + j);
      if (g.Z) g.Z[i * g.ldz + j] = v;
      if (g.gate) v *= __ldg(g.gate + i * g.ldg + j);
      if (g.relu) v = fmaxf(v, 0.f);
      float* c = g.C + i * g.ldc + j;
      if (g.gb_dZ && j >= g.gb_col0) {
        if (g.accumulate) v += *c;
        const int64_t o = i * g.gb_ld + (j - g.gb_col0);
        const float zz = __ldg(g.gb_Z + o), gg = __ldg(g.gb_G + o);
        const bool on = g.gb_from_h ? (zz > 0.f) : (zz * gg > 0.f);
        const float m = on ? v : 0.f;
        g.gb_dZ[o] = m * gg;
        g.gb_gG[o] += m * gate_bwd_z(zz, gg, on, g.gb_from_h);
        continue;
      }
      if (atomic) atomicAdd(c, v);
      else if (g.accumulate) *c += v;
      else *c = v;
    }
  }
}

// rowsum[i] += sum_k A[i + k * ld], i < I (I <= 1024): each block reduces a slab of k
__global__ void rowsum_kernel(const float* __restrict__ A, int64_t ld, int64_t K, int I, float* out) {
  const int64_t per = (K + gridDim.x - 1) / gridDim.x;
  const int64_t k0 = blockIdx.x * per, k1 = (k0 + per < K) ? k0 + per : K;
  for (int i = threadIdx.x; i < I; i += blockDim.x) {
    float s = 0.f;
    for (int64_t k = k0; k < k1; ++k) s += A[k * ld + i];
    atomicAdd(out + i, s);
  }
}

int launch_gemm_simt(const GemmArgs& a, cudaStream_t st) {
  ZEST_CHECK_ARG(a.A && a.B && a.C && a.I >= 0 && a.J > 0 && a.K >= 0 && a.splits >= 1, "gemm: bad arguments");
  ZEST_CHECK_ARG(a.splits == 1 || (!a.Z && !a.gate && !a.relu && !a.gb_dZ), "gemm: split-K cannot fuse a non-linear epilogue");
  ZEST_CHECK_ARG(!a.rowsum || a.sa_i == 1, "gemm: rowsum needs A with unit stride along its rows");
  if (a.I == 0) return ZEST_OK;
  if (a.rowsum && a.K > 0) {
    const unsigned rg = (unsigned)((a.K + 511) / 512 < 1024 ? (a.K + 511) / 512 : 1024);
    rowsum_kernel<<<rg, 256, 0, st>>>(a.A, a.sa_k, a.K, (int)a.I, a.rowsum);
    ZEST_LAUNCH_CHECK();
  }
  dim3 grid((unsigned)((a.I + BM - 1) / BM), (unsigned)((a.J + BN - 1) / BN), (unsigned)a.splits);
  ZEST_CHECK_ARG(grid.y < 65536 && (a.I + BM - 1) / BM < (1ll << 31), "gemm: shape too large for one launch");
  sgemm_kernel<<<grid, 256, 0, st>>>(a);
  ZEST_LAUNCH_CHECK();
  return ZEST_OK;
}

int launch_gemm(const GemmArgs& a, cudaStream_t st) {
  const int e = gemm_engine();
  if (e >= 1 && gemm_tc_supported(a)) return launch_gemm_tc(a, e, st);
  return launch_gemm_simt(a, st);
}

}  // namespace zest
