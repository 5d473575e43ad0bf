// Per-sample core of the fused feature gather, shared by gather.cu (stand-alone kernel) and mlp_tc.cu (the
// MLP kernel's loader warps): encoding-volume trilinear sample (utils.py:433-459 -> ATen grid_sampler_3d,
// align_corners=True, zeros padding) and one source view's bilinear RGB + in-bounds mask (utils.py:461-505 ->
// projection utils.py:257-269 + ATen grid_sampler_2d, border padding).  Index arithmetic follows ATen op by op
// with explicitly rounded intrinsics (no FMA contraction) so the integer voxel / pixel corners are bit-identical
// to the reference running on CPU; the K=3 projections use the fma chain of ATen's CPU matmul.
#pragma once
#include "common.cuh"

namespace zest {

// vol: channels-last [D,Hv,Wv,8] fp32.  acc[8] = trilinear sample at ndc (nx,ny,nz); corner0 (optional) = (x0,y0,z0)
__device__ __forceinline__ void trilinear8(const float* __restrict__ vol, int D, int Hv, int Wv, float nx, float ny, float nz,
                                           float (&acc)[8], int32_t* corner0 = nullptr) {
  // utils.py:451  grid = ndc * 2 - 1.0 ; then ATen unnormalize (align_corners) per axis
  const float ix = safe_int_range(unnormalize(__fsub_rn(__fmul_rn(nx, 2.f), 1.f), Wv));
  const float iy = safe_int_range(unnormalize(__fsub_rn(__fmul_rn(ny, 2.f), 1.f), Hv));
  const float iz = safe_int_range(unnormalize(__fsub_rn(__fmul_rn(nz, 2.f), 1.f), D));
  const float fx0 = floorf(ix), fy0 = floorf(iy), fz0 = floorf(iz);
  const int x0 = (int)fx0, y0 = (int)fy0, z0 = (int)fz0;
  if (corner0) { corner0[0] = x0; corner0[1] = y0; corner0[2] = z0; }
  const float wx[2] = {(fx0 + 1.f) - ix, ix - fx0};
  const float wy[2] = {(fy0 + 1.f) - iy, iy - fy0};
  const float wz[2] = {(fz0 + 1.f) - iz, iz - fz0};
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  // issue all in-bounds corner loads first (16 independent LDG.128), then blend
  float4 c[8][2];
  float w[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int dx = k & 1, dy = (k >> 1) & 1, dz = k >> 2;
    const int x = x0 + dx, y = y0 + dy, z = z0 + dz;
    const bool ok = (unsigned)x < (unsigned)Wv && (unsigned)y < (unsigned)Hv && (unsigned)z < (unsigned)D;
    w[k] = ok ? wx[dx] * wy[dy] * wz[dz] : 0.f;
    if (ok) {
      const float4* q = reinterpret_cast<const float4*>(vol + (((int64_t)z * Hv + y) * Wv + x) * 8);
      c[k][0] = __ldg(q);
      c[k][1] = __ldg(q + 1);
    } else {
      c[k][0] = c[k][1] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    acc[0] = fmaf(w[k], c[k][0].x, acc[0]);
    acc[1] = fmaf(w[k], c[k][0].y, acc[1]);
    acc[2] = fmaf(w[k], c[k][0].z, acc[2]);
    acc[3] = fmaf(w[k], c[k][0].w, acc[3]);
    acc[4] = fmaf(w[k], c[k][1].x, acc[4]);
    acc[5] = fmaf(w[k], c[k][1].y, acc[5]);
    acc[6] = fmaf(w[k], c[k][1].z, acc[6]);
    acc[7] = fmaf(w[k], c[k][1].w, acc[7]);
  }
}

// img: one view, [H,W,4] fp32 (r,g,b,0); cm: that view's camera row (w2c 3x4 | K 3x3).  Returns (r, g, b, mask);
// corner0 (optional) = (x0, y0)
__device__ __forceinline__ float4 view_sample(const float4* __restrict__ img, int H, int W, const float* cm, float px, float py,
                                              float pz, int32_t* corner0 = nullptr) {
  const float wm1 = (float)(W - 1), hm1 = (float)(H - 1);
  // utils.py:264  pts @ R^T + T   (matmul = fma chain, then a separately rounded add)
  const float c0 = __fadd_rn(dot3(px, py, pz, cm[0], cm[1], cm[2]), cm[3]);
  const float c1 = __fadd_rn(dot3(px, py, pz, cm[4], cm[5], cm[6]), cm[7]);
  const float c2 = __fadd_rn(dot3(px, py, pz, cm[8], cm[9], cm[10]), cm[11]);
  // utils.py:268  @ K^T
  const float i0 = dot3(c0, c1, c2, cm[12], cm[13], cm[14]);
  const float i1 = dot3(c0, c1, c2, cm[15], cm[16], cm[17]);
  const float i2 = dot3(c0, c1, c2, cm[18], cm[19], cm[20]);
  // utils.py:269  (xy / z + 0.0) / (W-1, H-1) ; utils.py:487  * 2.0 - 1.0
  const float gx = __fsub_rn(__fmul_rn(__fdiv_rn(__fadd_rn(__fdiv_rn(i0, i2), 0.f), wm1), 2.f), 1.f);
  const float gy = __fsub_rn(__fmul_rn(__fdiv_rn(__fadd_rn(__fdiv_rn(i1, i2), 0.f), hm1), 2.f), 1.f);
  const float mask = (gx > -1.f && gx < 1.f && gy > -1.f && gy < 1.f) ? 1.f : 0.f;  // utils.py:496
  // ATen grid_sampler_2d, border padding: unnormalize, clip to [0, size-1], floor
  const float ix = safe_int_range(fminf(wm1, fmaxf(unnormalize(gx, W), 0.f)));
  const float iy = safe_int_range(fminf(hm1, fmaxf(unnormalize(gy, H), 0.f)));
  const float fx0 = floorf(ix), fy0 = floorf(iy);
  const int x0 = (int)fx0, y0 = (int)fy0;
  if (corner0) { corner0[0] = x0; corner0[1] = y0; }
  const float wx1 = ix - fx0, wx0 = (fx0 + 1.f) - ix;
  const float wy1 = iy - fy0, wy0 = (fy0 + 1.f) - iy;
  const bool x0ok = (unsigned)x0 < (unsigned)W, x1ok = (unsigned)(x0 + 1) < (unsigned)W;
  const bool y0ok = (unsigned)y0 < (unsigned)H, y1ok = (unsigned)(y0 + 1) < (unsigned)H;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 a = (x0ok && y0ok) ? __ldg(img + (int64_t)y0 * W + x0) : z4;
  const float4 b = (x1ok && y0ok) ? __ldg(img + (int64_t)y0 * W + x0 + 1) : z4;
  const float4 c = (x0ok && y1ok) ? __ldg(img + (int64_t)(y0 + 1) * W + x0) : z4;
  const float4 d = (x1ok && y1ok) ? __ldg(img + (int64_t)(y0 + 1) * W + x0 + 1) : z4;
  const float wa = wx0 * wy0, wb = wx1 * wy0, wc = wx0 * wy1, wd = wx1 * wy1;
  const float r = fmaf(wd, d.x, fmaf(wc, c.x, fmaf(wb, b.x, wa * a.x)));
  const float g = fmaf(wd, d.y, fmaf(wc, c.y, fmaf(wb, b.y, wa * a.y)));
  const float bl = fmaf(wd, d.z, fmaf(wc, c.z, fmaf(wb, b.z, wa * a.z)));
  return make_float4(r, g, bl, mask);
}

}  // namespace zest
