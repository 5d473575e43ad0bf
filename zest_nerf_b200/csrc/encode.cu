// Positional encoding + MLP-input assembly for the fp32 path and the module boundary.
//
// Replaces networks.py:48-65 (Embedding.forward) and the torch.cat chain of
// renderer.py:246-297 (prepare_pts) / :300-318 (prepare_dynamic_pts):
//   x[m] = [ PE_n(ndc[m] (, t)) | feats[m] | PE_d(dirs[m / S]) ]
// Embedding layout: [v(C), sin(2^0 v)(C), cos(2^0 v)(C), sin(2^1 v)(C), ...]; 2^k * v is exact in
// fp32, sinf/cosf are the <= 1 ulp library versions (arguments reach 512 rad: no fast intrinsics).
// The bf16 tensor-core MLP does NOT use this kernel: it fuses the same encoding in its prologue.
#include "common.cuh"

namespace zest {

// One work item per (row, source value or (value, frequency) pair): an identity / feature item copies one float, a frequency
// item computes sin and cos of the same argument with one sincosf (half the range reductions of one-output-per-thread, and no
// sin / cos / copy divergence inside a warp) and writes the two outputs C (or 3) columns apart.  Same bits as sinf / cosf
// (checksums of a 524 288 x 131 encode identical); 0.427 -> 0.293 ms for that pass (tools/encode_time.py).
constexpr int kEncRows = 32;
__global__ void encode_fwd_kernel(const float* __restrict__ ndc, int ndc_ld, int has_t, float t, int nf_pts,
                                  const float* __restrict__ feats, int ldf, int F,
                                  const float* __restrict__ dirs, int nf_dir, int S, int64_t M,
                                  float* __restrict__ x, int ldx) {
  const int C = has_t ? 4 : 3;
  const int c_pe = C * (2 * nf_pts + 1);
  const int q_pe = C * (1 + nf_pts), q_dir = dirs ? 3 * (1 + nf_dir) : 0;
  const int Q = q_pe + F + q_dir;                       // items per row
  const int slab = kEncRows * Q;
  for (int64_t m0 = (int64_t)blockIdx.x * kEncRows; m0 < M; m0 += (int64_t)gridDim.x * kEncRows) {
    const int64_t ray0 = m0 / S;
    const int rem0 = (int)(m0 - ray0 * S);
    for (int e = threadIdx.x; e < slab; e += blockDim.x) {
      const int r = e / Q, q = e - r * Q;
      const int64_t m = m0 + r;
      if (m >= M) break;
      float* o = x + m * ldx;
      if (q < q_pe) {
        const int k1 = q / C, ch = q - k1 * C;           // k1 = 0: the value itself, else frequency 2^(k1 - 1)
        const float v = (ch == 3 && has_t != 2) ? t : __ldg(ndc + m * ndc_ld + ch);   // has_t == 2: 4th channel from memory
        if (k1 == 0) { o[ch] = v; continue; }
        float sn, cs;
        sincosf(v * (float)(1 << (k1 - 1)), &sn, &cs);
        o[C * (2 * k1 - 1) + ch] = sn;
        o[C * (2 * k1) + ch] = cs;
      } else if (q < q_pe + F) {
        o[c_pe + (q - q_pe)] = __ldg(feats + m * ldf + (q - q_pe));
      } else {
        const int qq = q - q_pe - F, k1 = qq / 3, ch = qq - k1 * 3;
        const float v = __ldg(dirs + (ray0 + (rem0 + r) / S) * 3 + ch);
        float* od = o + c_pe + F;
        if (k1 == 0) { od[ch] = v; continue; }
        float sn, cs;
        sincosf(v * (float)(1 << (k1 - 1)), &sn, &cs);
        od[3 * (2 * k1 - 1) + ch] = sn;
        od[3 * (2 * k1) + ch] = cs;
      }
    }
  }
}

// d/d ndc of PE: sum over blocks of gx * (1 | 2^k cos | -2^k sin); time channel has no gradient.
__global__ void encode_bwd_kernel(const float* __restrict__ ndc, int ndc_ld, int has_t, int nf_pts,
                                  const float* __restrict__ gx, int ldx, int64_t M, float* gndc, int gndc_ld,
                                  int accumulate) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int nch = has_t == 2 ? 4 : 3;   // a constant time channel (has_t == 1) has no gradient
  if (i >= M * nch) return;
  const int64_t m = i / nch;
  const int ch = (int)(i - m * nch);
  const int C = has_t ? 4 : 3;
  const float v = __ldg(ndc + m * ndc_ld + ch);
  const float* g = gx + m * ldx;
  float acc = __ldg(g + ch);
  for (int k = 0; k < nf_pts; ++k) {
    const float f = (float)(1 << k);
    float sn, cs;
    sincosf(v * f, &sn, &cs);
    acc += f * (__ldg(g + C * (1 + 2 * k) + ch) * cs - __ldg(g + C * (2 + 2 * k) + ch) * sn);
  }
  float* o = gndc + m * gndc_ld + ch;
  *o = accumulate ? (*o + acc) : acc;
}

}  // namespace zest

using namespace zest;

extern "C" int zest_encode_fwd(const float* ndc, int ndc_ld, int has_t, float t, int nf_pts, const float* feats,
                               int ldf, int F, const float* dirs, int nf_dir, int S, int64_t M, float* x,
                               int ldx, void* stream) {
  ZEST_CHECK_ARG(ndc && x && M >= 0 && ndc_ld >= 3 && nf_pts >= 0 && nf_pts <= 16 && nf_dir >= 0 && nf_dir <= 16 && S > 0,
                 "zest_encode_fwd: bad arguments");
  ZEST_CHECK_ARG(F == 0 || (feats && ldf >= F), "zest_encode_fwd: bad feats");
  ZEST_CHECK_ARG(has_t >= 0 && has_t <= 2 && (has_t != 2 || ndc_ld >= 4), "zest_encode_fwd: has_t must be 0, 1 (constant t) or 2 (4th channel in memory)");
  const int width = (has_t ? 4 : 3) * (2 * nf_pts + 1) + F + (dirs ? 3 * (2 * nf_dir + 1) : 0);
  ZEST_CHECK_ARG(ldx >= width, "zest_encode_fwd: ldx %d < row width %d", ldx, width);
  if (M == 0) return ZEST_OK;
  const int64_t blocks = (M + kEncRows - 1) / kEncRows;
  const unsigned grid = (unsigned)(blocks < (int64_t)num_sms() * 64 ? blocks : (int64_t)num_sms() * 64);
  encode_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(ndc, ndc_ld, has_t, t, nf_pts, feats, ldf, F, dirs,
                                                            nf_dir, S, M, x, ldx);
  ZEST_LAUNCH_CHECK();
  return ZEST_OK;
}

extern "C" int zest_encode_bwd(const float* ndc, int ndc_ld, int has_t, float t, int nf_pts, const float* gx,
                               int ldx, int64_t M, float* gndc, int gndc_ld, int accumulate, void* stream) {
  (void)t;
  ZEST_CHECK_ARG(ndc && gx && gndc && M >= 0 && ndc_ld >= 3 && gndc_ld >= 3 && nf_pts >= 0 && nf_pts <= 16,
                 "zest_encode_bwd: bad arguments");
  ZEST_CHECK_ARG(has_t >= 0 && has_t <= 2 && (has_t != 2 || (ndc_ld >= 4 && gndc_ld >= 4)), "zest_encode_bwd: bad has_t / strides");
  if (M == 0) return ZEST_OK;
  encode_bwd_kernel<<<(unsigned)((M * (has_t == 2 ? 4 : 3) + 255) / 256), 256, 0, (cudaStream_t)stream>>>(ndc, ndc_ld, has_t, nf_pts, gx,
                                                                                      ldx, M, gndc, gndc_ld, accumulate);
  ZEST_LAUNCH_CHECK();
  return ZEST_OK;
}
