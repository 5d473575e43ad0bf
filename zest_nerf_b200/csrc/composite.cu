// Alpha compositing along the ray: one warp per ray, shuffle scans, optional early termination.
//
// Replaces renderer.py:74-89 (depth2dist), :91-113 (raw2alpha), :115-164 (raw2outputs) and
// :166-219 (raw2outputs_blending).  Samples are interleaved over lanes (sample = chunk*32 + lane)
// so every global access of a warp is one contiguous row segment; the exclusive transmittance
// product is a 5-step multiplicative __shfl_up scan per 32-sample chunk with a carried prefix.
// Backward kernels recompute the forward quantities and run the mirrored suffix-sum scan.
#include "common.cuh"

namespace zest {

constexpr int kMaxChunks = 8;  // S <= 256
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ float sigmoidf(float x) { return 1.f / (1.f + expf(-x)); }

// inclusive product scan across the warp
__device__ __forceinline__ float warp_scan_mul(float v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    float o = __shfl_up_sync(kFull, v, d);
    if (lane >= d) v *= o;
  }
  return v;
}
// inclusive suffix sum across the warp (lane i gets sum over lanes >= i)
__device__ __forceinline__ float warp_rscan_add(float v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    float o = __shfl_down_sync(kFull, v, d);
    if (lane + d < 32) v += o;
  }
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(kFull, v, d);
  return v;
}

// dist of sample s (renderer.py:81-88): (z[s+1]-z[s]) * |d|, last = 1e10 * |d|
__device__ __forceinline__ float sample_dist(const float* z, int s, int S, float c) {
  return (s + 1 < S ? __ldg(z + s + 1) - __ldg(z + s) : 1e10f) * c;
}

__global__ void __launch_bounds__(256) composite_static_fwd_kernel(
    const float* __restrict__ raw, int ld, const float* __restrict__ z, const float* __restrict__ cosang,
    const float* __restrict__ noise, int64_t R, int S, int white, float t_stop, float* rgb_map,
    float* depth_map, float* acc_map, float* weights, float* alpha_out) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= R) return;
  const float c = __ldg(cosang + r);
  const float* zr = z + r * S;
  float T = 1.f, sr = 0.f, sg = 0.f, sb = 0.f, sd = 0.f, sa = 0.f;
  int s0 = 0;
  for (; s0 < S; s0 += 32) {
    const int s = s0 + lane;
    const bool on = s < S;
    float a = 0.f, cr = 0.f, cg = 0.f, cb = 0.f, zz = 0.f;
    if (on) {
      const float* q = raw + (r * S + s) * ld;
      cr = sigmoidf(__ldg(q));
      cg = sigmoidf(__ldg(q + 1));
      cb = sigmoidf(__ldg(q + 2));
      const float sig = fmaxf(__ldg(q + 3) + (noise ? __ldg(noise + r * S + s) : 0.f), 0.f);
      a = 1.f - expf(-sig * sample_dist(zr, s, S, c));
      zz = __ldg(zr + s);
    }
    const float incl = warp_scan_mul(on ? (1.f - a + 1e-10f) : 1.f, lane);
    float excl = __shfl_up_sync(kFull, incl, 1);
    if (lane == 0) excl = 1.f;
    const float w = a * (T * excl);
    if (on) {
      if (weights) weights[r * S + s] = w;
      if (alpha_out) alpha_out[r * S + s] = a;
    }
    sr += w * cr; sg += w * cg; sb += w * cb; sd += w * zz; sa += w;
    T *= __shfl_sync(kFull, incl, 31);
    if (t_stop > 0.f && T < t_stop) { s0 += 32; break; }  // early-termination mask (warp-uniform)
  }
  for (int s = s0 + lane; s < S; s += 32) {  // masked-out tail: zero weights
    if (weights) weights[r * S + s] = 0.f;
    if (alpha_out) alpha_out[r * S + s] = 0.f;
  }
  sr = warp_sum(sr); sg = warp_sum(sg); sb = warp_sum(sb); sd = warp_sum(sd); sa = warp_sum(sa);
  if (lane == 0) {
    const float bk = white ? (1.f - sa) : 0.f;
    rgb_map[r * 3 + 0] = sr + bk;
    rgb_map[r * 3 + 1] = sg + bk;
    rgb_map[r * 3 + 2] = sb + bk;
    depth_map[r] = sd;
    if (acc_map) acc_map[r] = sa;
  }
}

__global__ void __launch_bounds__(256) composite_static_bwd_kernel(
    const float* __restrict__ raw, int ld, const float* __restrict__ z, const float* __restrict__ cosang,
    const float* __restrict__ noise, int64_t R, int S, int white, const float* __restrict__ g_rgb,
    const float* __restrict__ g_depth, const float* __restrict__ g_w, const float* __restrict__ g_alpha,
    float* g_raw, int ldg) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= R) return;
  const float c = __ldg(cosang + r);
  const float* zr = z + r * S;
  const float gr = g_rgb ? __ldg(g_rgb + r * 3) : 0.f, gg = g_rgb ? __ldg(g_rgb + r * 3 + 1) : 0.f,
              gb = g_rgb ? __ldg(g_rgb + r * 3 + 2) : 0.f, gd = g_depth ? __ldg(g_depth + r) : 0.f;
  const float gacc = white ? -(gr + gg + gb) : 0.f;
  const int nch = (S + 31) >> 5;
  float a_[kMaxChunks], T_[kMaxChunks], x_[kMaxChunks], dist_[kMaxChunks], pre_[kMaxChunks];
  float cr_[kMaxChunks], cg_[kMaxChunks], cb_[kMaxChunks];
  float T = 1.f;
#pragma unroll
  for (int j = 0; j < kMaxChunks; ++j) {
    if (j >= nch) break;
    const int s = j * 32 + lane;
    const bool on = s < S;
    float a = 0.f, pre = -1.f, dist = 0.f;
    cr_[j] = cg_[j] = cb_[j] = 0.f;
    if (on) {
      const float* q = raw + (r * S + s) * ld;
      cr_[j] = sigmoidf(__ldg(q)); cg_[j] = sigmoidf(__ldg(q + 1)); cb_[j] = sigmoidf(__ldg(q + 2));
      pre = __ldg(q + 3) + (noise ? __ldg(noise + r * S + s) : 0.f);
      dist = sample_dist(zr, s, S, c);
      a = 1.f - expf(-fmaxf(pre, 0.f) * dist);
    }
    const float x = on ? (1.f - a + 1e-10f) : 1.f;
    const float incl = warp_scan_mul(x, lane);
    float excl = __shfl_up_sync(kFull, incl, 1);
    if (lane == 0) excl = 1.f;
    a_[j] = a; x_[j] = x; dist_[j] = dist; pre_[j] = pre; T_[j] = T * excl;
    T *= __shfl_sync(kFull, incl, 31);
  }
  float carry = 0.f;  // sum_{k > i} gw_k w_k over later chunks
#pragma unroll
  for (int j = kMaxChunks - 1; j >= 0; --j) {
    if (j >= nch) continue;
    const int s = j * 32 + lane;
    const bool on = s < S;
    const float w = a_[j] * T_[j];
    const float zz = on ? __ldg(zr + s) : 0.f;
    const float gw = on ? (gr * cr_[j] + gg * cg_[j] + gb * cb_[j] + gd * zz + gacc + (g_w ? __ldg(g_w + r * S + s) : 0.f)) : 0.f;
    const float incl = warp_rscan_add(gw * w, lane);
    const float suffix = incl - gw * w + carry;  // strictly-later samples
    carry += __shfl_sync(kFull, incl, 0);
    if (on) {
      float ga = gw * T_[j] - suffix / x_[j] + (g_alpha ? __ldg(g_alpha + r * S + s) : 0.f);
      float* o = g_raw + (r * S + s) * ldg;
      o[0] = gr * w * cr_[j] * (1.f - cr_[j]);
      o[1] = gg * w * cg_[j] * (1.f - cg_[j]);
      o[2] = gb * w * cb_[j] * (1.f - cb_[j]);
      o[3] = pre_[j] > 0.f ? ga * dist_[j] * (1.f - a_[j]) : 0.f;
    }
  }
}

__global__ void __launch_bounds__(256) composite_blend_fwd_kernel(
    const float* __restrict__ raw_dy, int ld_dy, const float* __restrict__ raw_rig, int ld_rig,
    const float* __restrict__ z, const float* __restrict__ cosang, const float* __restrict__ noise,
    int64_t R, int S, float t_stop, float* rgb_map, float* depth_map, float* rgb_dy, float* depth_dy,
    float* w_dd, float* weights_dy) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= R) return;
  const float c = __ldg(cosang + r);
  const float* zr = z + r * S;
  float T = 1.f, Tf = 1.f;
  float mr = 0.f, mg = 0.f, mb = 0.f, md = 0.f, fr = 0.f, fg = 0.f, fb = 0.f, fd = 0.f, dd = 0.f;
  int s0 = 0;
  for (; s0 < S; s0 += 32) {
    const int s = s0 + lane;
    const bool on = s < S;
    float ady = 0.f, arig = 0.f, efg = 0.f, zz = 0.f;
    float dr = 0.f, dg = 0.f, db = 0.f, rr = 0.f, rg = 0.f, rb = 0.f;
    if (on) {
      const float* qd = raw_dy + (r * S + s) * ld_dy;
      const float* qr = raw_rig + (r * S + s) * ld_rig;
      dr = sigmoidf(__ldg(qd)); dg = sigmoidf(__ldg(qd + 1)); db = sigmoidf(__ldg(qd + 2));
      rr = sigmoidf(__ldg(qr)); rg = sigmoidf(__ldg(qr + 1)); rb = sigmoidf(__ldg(qr + 2));
      const float nz = noise ? __ldg(noise + r * S + s) : 0.f;
      const float dist = sample_dist(zr, s, S, c);
      const float b = __ldg(qr + 4);
      efg = 1.f - expf(-fmaxf(__ldg(qd + 3) + nz, 0.f) * dist);
      ady = efg * b;
      arig = (1.f - expf(-fmaxf(__ldg(qr + 3) + nz, 0.f) * dist)) * (1.f - b);
      zz = __ldg(zr + s);
    }
    const float incl = warp_scan_mul(on ? ((1.f - ady) * (1.f - arig) + 1e-10f) : 1.f, lane);
    const float inclf = warp_scan_mul(on ? (1.f - efg + 1e-10f) : 1.f, lane);
    float excl = __shfl_up_sync(kFull, incl, 1), exclf = __shfl_up_sync(kFull, inclf, 1);
    if (lane == 0) excl = exclf = 1.f;
    const float Ti = T * excl, wdy = Ti * ady, wrig = Ti * arig, wfg = efg * (Tf * exclf);
    mr += wdy * dr + wrig * rr; mg += wdy * dg + wrig * rg; mb += wdy * db + wrig * rb;
    md += (wdy + wrig) * zz;
    fr += wfg * dr; fg += wfg * dg; fb += wfg * db; fd += wfg * zz;
    dd += wdy;
    if (on && weights_dy) weights_dy[r * S + s] = wfg;
    T *= __shfl_sync(kFull, incl, 31);
    Tf *= __shfl_sync(kFull, inclf, 31);
    if (t_stop > 0.f && T < t_stop && Tf < t_stop) { s0 += 32; break; }
  }
  if (weights_dy)
    for (int s = s0 + lane; s < S; s += 32) weights_dy[r * S + s] = 0.f;
  mr = warp_sum(mr); mg = warp_sum(mg); mb = warp_sum(mb); md = warp_sum(md);
  fr = warp_sum(fr); fg = warp_sum(fg); fb = warp_sum(fb); fd = warp_sum(fd); dd = warp_sum(dd);
  if (lane == 0) {
    rgb_map[r * 3] = mr; rgb_map[r * 3 + 1] = mg; rgb_map[r * 3 + 2] = mb; depth_map[r] = md;
    rgb_dy[r * 3] = fr; rgb_dy[r * 3 + 1] = fg; rgb_dy[r * 3 + 2] = fb; depth_dy[r] = fd;
    w_dd[r] = dd;
  }
}

__global__ void __launch_bounds__(128) composite_blend_bwd_kernel(
    const float* __restrict__ raw_dy, int ld_dy, const float* __restrict__ raw_rig, int ld_rig,
    const float* __restrict__ z, const float* __restrict__ cosang, const float* __restrict__ noise,
    int64_t R, int S, const float* __restrict__ g_rgb, const float* __restrict__ g_depth,
    const float* __restrict__ g_rgb_dy, const float* __restrict__ g_depth_dy,
    const float* __restrict__ g_wdy, float* g_raw_dy, int ldgd, float* g_raw_rig, int ldgr) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= R) return;
  const float c = __ldg(cosang + r);
  const float* zr = z + r * S;
  float gm[3] = {0.f, 0.f, 0.f}, gf[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    if (g_rgb) gm[k] = __ldg(g_rgb + r * 3 + k);
    if (g_rgb_dy) gf[k] = __ldg(g_rgb_dy + r * 3 + k);
  }
  const float gmd = g_depth ? __ldg(g_depth + r) : 0.f, gfd = g_depth_dy ? __ldg(g_depth_dy + r) : 0.f;
  const int nch = (S + 31) >> 5;
  float T_[kMaxChunks], Tf_[kMaxChunks];
  float T = 1.f, Tf = 1.f;
  // pass 1: transmittances (per-sample terms are recomputed in pass 2 to bound registers)
#pragma unroll
  for (int j = 0; j < kMaxChunks; ++j) {
    if (j >= nch) break;
    const int s = j * 32 + lane;
    const bool on = s < S;
    float x = 1.f, xf = 1.f;
    if (on) {
      const float* qd = raw_dy + (r * S + s) * ld_dy;
      const float* qr = raw_rig + (r * S + s) * ld_rig;
      const float nz = noise ? __ldg(noise + r * S + s) : 0.f;
      const float dist = sample_dist(zr, s, S, c);
      const float b = __ldg(qr + 4);
      const float efg = 1.f - expf(-fmaxf(__ldg(qd + 3) + nz, 0.f) * dist);
      const float erig = 1.f - expf(-fmaxf(__ldg(qr + 3) + nz, 0.f) * dist);
      x = (1.f - efg * b) * (1.f - erig * (1.f - b)) + 1e-10f;
      xf = 1.f - efg + 1e-10f;
    }
    const float incl = warp_scan_mul(x, lane), inclf = warp_scan_mul(xf, lane);
    float excl = __shfl_up_sync(kFull, incl, 1), exclf = __shfl_up_sync(kFull, inclf, 1);
    if (lane == 0) excl = exclf = 1.f;
    T_[j] = T * excl; Tf_[j] = Tf * exclf;
    T *= __shfl_sync(kFull, incl, 31);
    Tf *= __shfl_sync(kFull, inclf, 31);
  }
  float carry = 0.f, carryf = 0.f;
#pragma unroll
  for (int j = kMaxChunks - 1; j >= 0; --j) {
    if (j >= nch) continue;
    const int s = j * 32 + lane;
    const bool on = s < S;
    float dc[3] = {0.f, 0.f, 0.f}, rc[3] = {0.f, 0.f, 0.f};
    float efg = 0.f, erig = 0.f, b = 0.f, dist = 0.f, pre_d = -1.f, pre_r = -1.f, zz = 0.f;
    if (on) {
      const float* qd = raw_dy + (r * S + s) * ld_dy;
      const float* qr = raw_rig + (r * S + s) * ld_rig;
#pragma unroll
      for (int k = 0; k < 3; ++k) { dc[k] = sigmoidf(__ldg(qd + k)); rc[k] = sigmoidf(__ldg(qr + k)); }
      const float nz = noise ? __ldg(noise + r * S + s) : 0.f;
      dist = sample_dist(zr, s, S, c);
      b = __ldg(qr + 4);
      pre_d = __ldg(qd + 3) + nz; pre_r = __ldg(qr + 3) + nz;
      efg = 1.f - expf(-fmaxf(pre_d, 0.f) * dist);
      erig = 1.f - expf(-fmaxf(pre_r, 0.f) * dist);
      zz = __ldg(zr + s);
    }
    const float ady = efg * b, arig = erig * (1.f - b);
    const float x = on ? ((1.f - ady) * (1.f - arig) + 1e-10f) : 1.f, xf = on ? (1.f - efg + 1e-10f) : 1.f;
    const float gwdy = gm[0] * dc[0] + gm[1] * dc[1] + gm[2] * dc[2] + gmd * zz;
    const float gwrig = gm[0] * rc[0] + gm[1] * rc[1] + gm[2] * rc[2] + gmd * zz;
    const float gwfg = gf[0] * dc[0] + gf[1] * dc[1] + gf[2] * dc[2] + gfd * zz + ((on && g_wdy) ? __ldg(g_wdy + r * S + s) : 0.f);
    const float wfg = efg * Tf_[j];
    const float tq = on ? T_[j] * (gwdy * ady + gwrig * arig) : 0.f;
    const float tf = on ? gwfg * wfg : 0.f;
    const float incl = warp_rscan_add(tq, lane), inclf = warp_rscan_add(tf, lane);
    const float suffix = incl - tq + carry, suffixf = inclf - tf + carryf;
    carry += __shfl_sync(kFull, incl, 0);
    carryf += __shfl_sync(kFull, inclf, 0);
    if (on) {
      const float gx = suffix / x;
      const float g_ady = gwdy * T_[j] - gx * (1.f - arig);
      const float g_arig = gwrig * T_[j] - gx * (1.f - ady);
      const float g_efg = g_ady * b + gwfg * Tf_[j] - suffixf / xf;
      const float g_erig = g_arig * (1.f - b);
      const float g_b = g_ady * efg - g_arig * erig;
      const float wdy = T_[j] * ady, wrig = T_[j] * arig;
      float* od = g_raw_dy + (r * S + s) * ldgd;
      float* orr = g_raw_rig + (r * S + s) * ldgr;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        od[k] = (gm[k] * wdy + gf[k] * wfg) * dc[k] * (1.f - dc[k]);
        orr[k] = gm[k] * wrig * rc[k] * (1.f - rc[k]);
      }
      od[3] = pre_d > 0.f ? g_efg * dist * (1.f - efg) : 0.f;
      orr[3] = pre_r > 0.f ? g_erig * dist * (1.f - erig) : 0.f;
      orr[4] = g_b;
    }
  }
}

}  // namespace zest

using namespace zest;

extern "C" int zest_composite_static_fwd(const float* raw, int ld_raw, const float* z, const float* cos_angle,
                                         const float* noise, int64_t R, int S, int white_bkgd, float t_stop,
                                         float* rgb_map, float* depth_map, float* acc, float* weights,
                                         float* alpha, void* stream) {
  ZEST_CHECK_ARG(raw && z && cos_angle && rgb_map && depth_map && ld_raw >= 4 && S > 0 && R >= 0,
                 "zest_composite_static_fwd: bad arguments");
  if (R == 0) return ZEST_OK;
  composite_static_fwd_kernel<<<(unsigned)((R + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
      raw, ld_raw, z, cos_angle, noise, R, S, white_bkgd, t_stop, rgb_map, depth_map, acc, weights, alpha);
  ZEST_LAUNCH_CHECK();
  return ZEST_OK;
}

extern "C" int zest_composite_static_bwd(const float* raw, int ld_raw, const float* z, const float* cos_angle,
                                         const float* noise, int64_t R, int S, int white_bkgd,
                                         const float* g_rgb_map, const float* g_depth_map,
                                         const float* g_weights, const float* g_alpha, float* g_raw,
                                         int ld_graw, void* stream) {
  ZEST_CHECK_ARG(raw && z && cos_angle && g_raw && ld_raw >= 4 && ld_graw >= 4 && S > 0 && R >= 0,
                 "zest_composite_static_bwd: bad arguments");
  ZEST_CHECK_ARG(S <= 32 * kMaxChunks, "zest_composite_static_bwd: S > %d unsupported", 32 * kMaxChunks);
  if (R == 0) return ZEST_OK;
  composite_static_bwd_kernel<<<(unsigned)((R + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
      raw, ld_raw, z, cos_angle, noise, R, S, white_bkgd, g_rgb_map, g_depth_map, g_weights, g_alpha, g_raw, ld_graw);
  ZEST_LAUNCH_CHECK();
  return ZEST_OK;
}

extern "C" int zest_composite_blend_fwd(const float* raw_dy, int ld_dy, const float* raw_rig, int ld_rig,
                                        const float* z, const float* cos_angle, const float* noise, int64_t R,
                                        int S, float t_stop, float* rgb_map, float* depth_map, float* rgb_map_dy,
                                        float* depth_map_dy, float* weights_dd, float* weights_dy, void* stream) {
  ZEST_CHECK_ARG(raw_dy && raw_rig && z && cos_angle && rgb_map && depth_map && rgb_map_dy && depth_map_dy &&
                     weights_dd && ld_dy >= 4 && ld_rig >= 5 && S > 0 && R >= 0,
                 "zest_composite_blend_fwd: bad arguments");
  if (R == 0) return ZEST_OK;
  composite_blend_fwd_kernel<<<(unsigned)((R + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
      raw_dy, ld_dy, raw_rig, ld_rig, z, cos_angle, noise, R, S, t_stop, rgb_map, depth_map, rgb_map_dy,
      depth_map_dy, weights_dd, weights_dy);
  ZEST_LAUNCH_CHECK();
  return ZEST_OK;
}

extern "C" int zest_composite_blend_bwd(const float* raw_dy, int ld_dy, const float* raw_rig, int ld_rig,
                                        const float* z, const float* cos_angle, const float* noise, int64_t R,
                                        int S, const float* g_rgb_map, const float* g_depth_map,
                                        const float* g_rgb_map_dy, const float* g_depth_map_dy,
                                        const float* g_weights_dy, float* g_raw_dy, int ld_gdy, float* g_raw_rig,
                                        int ld_grig, void* stream) {
  ZEST_CHECK_ARG(raw_dy && raw_rig && z && cos_angle && g_raw_dy && g_raw_rig && ld_dy >= 4 && ld_rig >= 5 &&
                     ld_gdy >= 4 && ld_grig >= 5 && S > 0 && R >= 0,
                 "zest_composite_blend_bwd: bad arguments");
  ZEST_CHECK_ARG(S <= 32 * kMaxChunks, "zest_composite_blend_bwd: S > %d unsupported", 32 * kMaxChunks);
  if (R == 0) return ZEST_OK;
  composite_blend_bwd_kernel<<<(unsigned)((R + 3) / 4), 128, 0, (cudaStream_t)stream>>>(
      raw_dy, ld_dy, raw_rig, ld_rig, z, cos_angle, noise, R, S, g_rgb_map, g_depth_map, g_rgb_map_dy,
      g_depth_map_dy, g_weights_dy, g_raw_dy, ld_gdy, g_raw_rig, ld_grig);
  ZEST_LAUNCH_CHECK();
  return ZEST_OK;
}
