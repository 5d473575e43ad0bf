// "Next" row f4: the training-only scene-flow reductions that re-read the per-sample [R, S, 3] tensors the
// render path returns (raw_pts_ref / _post / _prev / _pp, weights_ref_dy):
//
//   zest_sf_smooth_loss_fwd/_bwd   losses.py:142-161  compute_sf_smooth_loss  (mean |sf_s - sf_{s+1}| over the closest 95 %)
//   zest_sf_lke_loss_fwd/_bwd      losses.py:164-203  compute_sf_lke_loss     (0.5 mean (sf_fwd - sf_bwd)^2 over the closest 90 %)
//   zest_project_ndc_fwd/_bwd      utils.py:507-539   projection_from_ndc     (sum_s w p -> NDC2Euclidean -> w2c -> pixel)
//
// all through NDC2Euclidean (utils.py:507-514).  The reference runs each as 10-20 elementwise PyTorch kernels over [R, S, 3]
// temporaries plus a mean; here each is one pass (12 B per point read once, the Jacobian of NDC2Euclidean applied in
// registers on the way back).  HBM-bound streaming reductions: 24 / 36 B per sample forward.
// Loss sums are accumulated in double (per-block partial -> one atomicAdd), the mean is taken on the host side of the ABI.
#include "common.cuh"

namespace zest {
namespace {

struct Ndc2E {
  float kx, ky;   // W / (2 f), H / (2 f) applied as the reference does: ((-x) * z_e * W) / (2 f)
  float W, H, f2;
};

// utils.py:507-514: z_e = 2 / (clamp(z, -1, 0.99) - 1); x_e = -x z_e W / (2 f); y_e = -y z_e H / (2 f)   (op order kept)
__device__ __forceinline__ void ndc2e(const Ndc2E& c, float x, float y, float z, float& xe, float& ye, float& ze) {
  const float cz = fminf(fmaxf(z, -1.0f), 0.99f);
  ze = __fdiv_rn(2.f, __fsub_rn(cz, 1.f));
  xe = __fdiv_rn(__fmul_rn(__fmul_rn(-x, ze), c.W), c.f2);
  ye = __fdiv_rn(__fmul_rn(__fmul_rn(-y, ze), c.H), c.f2);
}
// (gx, gy, gz) wrt (x_e, y_e, z_e) -> gradient wrt the NDC point; torch.clamp passes the gradient on [min, max] inclusive
__device__ __forceinline__ void ndc2e_bwd(const Ndc2E& c, float x, float y, float z, float gxe, float gye, float gze, float& gx,
                                          float& gy, float& gz) {
  const float cz = fminf(fmaxf(z, -1.0f), 0.99f);
  const float d = cz - 1.f, ze = 2.f / d;
  const float dze = (z >= -1.0f && z <= 0.99f) ? -2.f / (d * d) : 0.f;
  gx = gxe * (-ze * c.W / c.f2);
  gy = gye * (-ze * c.H / c.f2);
  gz = dze * (gze - gxe * x * c.W / c.f2 - gye * y * c.H / c.f2);
}

__device__ __forceinline__ void block_sum_to(double v, double* out) {
  __shared__ double part[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) part[warp] = v;
  __syncthreads();
  if (warp == 0) {
    v = lane < (int)(blockDim.x >> 5) ? part[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) atomicAdd(out, v);
  }
}

__device__ __forceinline__ void load3(const float* p, int64_t i, float& x, float& y, float& z) {
  x = __ldg(p + 3 * i); y = __ldg(p + 3 * i + 1); z = __ldg(p + 3 * i + 2);
}
__device__ __forceinline__ float sgn(float v) { return (v > 0.f) ? 1.f : ((v < 0.f) ? -1.f : 0.f); }

// scene flow in Euclidean space at sample i: E(p1) - E(p2)
__device__ __forceinline__ void flow_at(const Ndc2E& c, const float* p1, const float* p2, int64_t i, float (&d)[3]) {
  float x, y, z, a[3], b[3];
  load3(p1, i, x, y, z); ndc2e(c, x, y, z, a[0], a[1], a[2]);
  load3(p2, i, x, y, z); ndc2e(c, x, y, z, b[0], b[1], b[2]);
#pragma unroll
  for (int k = 0; k < 3; ++k) d[k] = __fsub_rn(a[k], b[k]);
}

// one thread per (ray, sample s < n_close - 1): sum |sf_s - sf_{s+1}|
__global__ void sf_smooth_fwd_kernel(const float* __restrict__ p1, const float* __restrict__ p2, int64_t R, int S, int n_close, Ndc2E c,
                                     double* __restrict__ out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int per = n_close - 1;
  double acc = 0.0;
  if (t < R * per) {
    const int64_t r = t / per;
    const int s = (int)(t - r * per);
    float d0[3], d1[3];
    flow_at(c, p1, p2, r * S + s, d0);
    flow_at(c, p1, p2, r * S + s + 1, d1);
#pragma unroll
    for (int k = 0; k < 3; ++k) acc += (double)fabsf(__fsub_rn(d0[k], d1[k]));
  }
  block_sum_to(acc, out);
}

// one thread per (ray, sample): d loss / d sf_s = scale (sgn(sf_s - sf_{s+1}) [s < n-1] - sgn(sf_{s-1} - sf_s) [s >= 1]), s < n
__global__ void sf_smooth_bwd_kernel(const float* __restrict__ p1, const float* __restrict__ p2, int64_t R, int S, int n_close, Ndc2E c,
                                     const float* __restrict__ gout, float scale, float* __restrict__ g1, float* __restrict__ g2) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= R * S) return;
  const int s = (int)(t % S);
  float ga[3] = {0.f, 0.f, 0.f};
  if (s < n_close) {
    const float w = scale * __ldg(gout);
    float d[3], dn[3];
    flow_at(c, p1, p2, t, d);
    if (s < n_close - 1) {
      flow_at(c, p1, p2, t + 1, dn);
#pragma unroll
      for (int k = 0; k < 3; ++k) ga[k] += w * sgn(__fsub_rn(d[k], dn[k]));
    }
    if (s >= 1) {
      flow_at(c, p1, p2, t - 1, dn);
#pragma unroll
      for (int k = 0; k < 3; ++k) ga[k] -= w * sgn(__fsub_rn(dn[k], d[k]));
    }
  }
  float x, y, z, gx, gy, gz;
  if (g1) { load3(p1, t, x, y, z); ndc2e_bwd(c, x, y, z, ga[0], ga[1], ga[2], gx, gy, gz); g1[3 * t] = gx; g1[3 * t + 1] = gy; g1[3 * t + 2] = gz; }
  if (g2) { load3(p2, t, x, y, z); ndc2e_bwd(c, x, y, z, -ga[0], -ga[1], -ga[2], gx, gy, gz); g2[3 * t] = gx; g2[3 * t + 1] = gy; g2[3 * t + 2] = gz; }
}

// v = (E(post) - E(ref)) - (E(ref) - E(prev)) at one sample (losses.py:190-203, op order kept)
__device__ __forceinline__ void lke_at(const Ndc2E& c, const float* ref, const float* post, const float* prev, int64_t i, float (&v)[3]) {
  float x, y, z, a[3], b[3], d[3];
  load3(ref, i, x, y, z); ndc2e(c, x, y, z, a[0], a[1], a[2]);
  load3(post, i, x, y, z); ndc2e(c, x, y, z, b[0], b[1], b[2]);
  load3(prev, i, x, y, z); ndc2e(c, x, y, z, d[0], d[1], d[2]);
#pragma unroll
  for (int k = 0; k < 3; ++k) v[k] = __fsub_rn(__fsub_rn(b[k], a[k]), __fsub_rn(a[k], d[k]));
}

__global__ void sf_lke_fwd_kernel(const float* __restrict__ ref, const float* __restrict__ post, const float* __restrict__ prev, int64_t R,
                                  int S, int n_close, Ndc2E c, double* __restrict__ out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double acc = 0.0;
  if (t < R * n_close) {
    const int64_t r = t / n_close;
    const int s = (int)(t - r * n_close);
    float v[3];
    lke_at(c, ref, post, prev, r * S + s, v);
#pragma unroll
    for (int k = 0; k < 3; ++k) acc += (double)__fmul_rn(v[k], v[k]);
  }
  block_sum_to(acc, out);
}

__global__ void sf_lke_bwd_kernel(const float* __restrict__ ref, const float* __restrict__ post, const float* __restrict__ prev, int64_t R,
                                  int S, int n_close, Ndc2E c, const float* __restrict__ gout, float scale, float* __restrict__ g_ref,
                                  float* __restrict__ g_post, float* __restrict__ g_prev) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= R * S) return;
  const int s = (int)(t % S);
  float g[3] = {0.f, 0.f, 0.f};
  if (s < n_close) {
    float v[3];
    lke_at(c, ref, post, prev, t, v);
    const float w = scale * __ldg(gout);     // scale = 0.5 * 2 / N
#pragma unroll
    for (int k = 0; k < 3; ++k) g[k] = w * v[k];
  }
  float x, y, z, gx, gy, gz;
  if (g_ref) { load3(ref, t, x, y, z); ndc2e_bwd(c, x, y, z, -2.f * g[0], -2.f * g[1], -2.f * g[2], gx, gy, gz); g_ref[3 * t] = gx; g_ref[3 * t + 1] = gy; g_ref[3 * t + 2] = gz; }
  if (g_post) { load3(post, t, x, y, z); ndc2e_bwd(c, x, y, z, g[0], g[1], g[2], gx, gy, gz); g_post[3 * t] = gx; g_post[3 * t + 1] = gy; g_post[3 * t + 2] = gz; }
  if (g_prev) { load3(prev, t, x, y, z); ndc2e_bwd(c, x, y, z, g[0], g[1], g[2], gx, gy, gz); g_prev[3 * t] = gx; g_prev[3 * t + 1] = gy; g_prev[3 * t + 2] = gz; }
}

// ---- projection_from_ndc: warp per ray
struct Pose { float r[9], t[3]; };
// w2c: DEVICE pointer to a row-major 4x4 (or 3x4) world-to-camera matrix
__device__ __forceinline__ Pose load_pose(const float* __restrict__ w2c) {
  Pose p;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
#pragma unroll
    for (int k = 0; k < 3; ++k) p.r[3 * j + k] = __ldg(w2c + 4 * j + k);
    p.t[j] = __ldg(w2c + 4 * j + 3);
  }
  return p;
}

__device__ __forceinline__ void weighted_point(const float* __restrict__ w, const float* __restrict__ pts, int64_t ray, int S, int lane,
                                               float (&p)[3]) {
  p[0] = p[1] = p[2] = 0.f;
  for (int s = lane; s < S; s += 32) {
    const float ww = __ldg(w + ray * S + s);
    float x, y, z;
    load3(pts, ray * S + s, x, y, z);
    p[0] = fmaf(ww, x, p[0]); p[1] = fmaf(ww, y, p[1]); p[2] = fmaf(ww, z, p[2]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int k = 0; k < 3; ++k) p[k] += __shfl_xor_sync(0xffffffffu, p[k], o);
  }
}

__global__ void project_ndc_fwd_kernel(const float* __restrict__ w, const float* __restrict__ pts, int64_t R, int S, const float* __restrict__ w2c, Ndc2E c,
                                       float* __restrict__ out) {
  const int64_t ray = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (ray >= R) return;
  float p[3];
  weighted_point(w, pts, ray, S, lane, p);
  if (lane == 0) {
    const Pose pose = load_pose(w2c);
    float e[3], l[3];
    ndc2e(c, p[0], p[1], p[2], e[0], e[1], e[2]);
#pragma unroll
    for (int j = 0; j < 3; ++j) l[j] = dot3(e[0], e[1], e[2], pose.r[3 * j], pose.r[3 * j + 1], pose.r[3 * j + 2]) + pose.t[j];
    // utils.py:521-524: (x f / -z + W/2, -y f / -z + H/2)
    const float f = 0.5f * c.f2;
    out[2 * ray] = l[0] * f / -l[2] + c.W * 0.5f;
    out[2 * ray + 1] = -l[1] * f / -l[2] + c.H * 0.5f;
  }
}

__global__ void project_ndc_bwd_kernel(const float* __restrict__ w, const float* __restrict__ pts, int64_t R, int S, const float* __restrict__ w2c, Ndc2E c,
                                       const float* __restrict__ g2d, float* __restrict__ gw, float* __restrict__ gpts) {
  const int64_t ray = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (ray >= R) return;
  float p[3];
  weighted_point(w, pts, ray, S, lane, p);
  const Pose pose = load_pose(w2c);
  float e[3], l[3];
  ndc2e(c, p[0], p[1], p[2], e[0], e[1], e[2]);
#pragma unroll
  for (int j = 0; j < 3; ++j) l[j] = dot3(e[0], e[1], e[2], pose.r[3 * j], pose.r[3 * j + 1], pose.r[3 * j + 2]) + pose.t[j];
  const float f = 0.5f * c.f2;
  const float gu = __ldg(g2d + 2 * ray), gv = __ldg(g2d + 2 * ray + 1);
  // u = -f lx / lz + W/2 ; v = f ly / lz + H/2
  const float iz = 1.f / l[2];
  float gl[3] = {-gu * f * iz, gv * f * iz, (gu * f * l[0] - gv * f * l[1]) * iz * iz};
  float ge[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) ge[k] = gl[0] * pose.r[k] + gl[1] * pose.r[3 + k] + gl[2] * pose.r[6 + k];
  float gp[3];
  ndc2e_bwd(c, p[0], p[1], p[2], ge[0], ge[1], ge[2], gp[0], gp[1], gp[2]);
  for (int s = lane; s < S; s += 32) {
    const int64_t i = ray * S + s;
    const float ww = __ldg(w + i);
    float x, y, z;
    load3(pts, i, x, y, z);
    if (gw) gw[i] = gp[0] * x + gp[1] * y + gp[2] * z;
    if (gpts) { gpts[3 * i] = gp[0] * ww; gpts[3 * i + 1] = gp[1] * ww; gpts[3 * i + 2] = gp[2] * ww; }
  }
}

Ndc2E make_consts(int H, int W, float f) {
  Ndc2E c;
  c.W = (float)W; c.H = (float)H; c.f2 = 2.f * f;
  c.kx = c.W / c.f2; c.ky = c.H / c.f2;
  return c;
}

}  // namespace
}  // namespace zest

using namespace zest;

extern "C" int zest_sf_smooth_loss_fwd(const float* pts_1, const float* pts_2, int64_t R, int S, int n_close, int H, int W, float f,
                                       double* loss_sum, void* stream) {
  ZEST_CHECK_ARG(pts_1 && pts_2 && loss_sum && R >= 0 && S >= 2 && n_close >= 2 && n_close <= S && f != 0.f,
                 "zest_sf_smooth_loss_fwd: bad arguments");
  const int64_t n = R * (n_close - 1);
  if (n == 0) return ZEST_OK;
  sf_smooth_fwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(pts_1, pts_2, R, S, n_close, make_consts(H, W, f), loss_sum);
  ZEST_LAUNCH_CHECK();
  return ZEST_OK;
}

extern "C" int zest_sf_smooth_loss_bwd(const float* pts_1, const float* pts_2, int64_t R, int S, int n_close, int H, int W, float f,
                                       const float* g_loss, float* g_pts_1, float* g_pts_2, void* stream) {
  ZEST_CHECK_ARG(pts_1 && pts_2 && g_loss && R >= 0 && S >= 2 && n_close >= 2 && n_close <= S && f != 0.f,
                 "zest_sf_smooth_loss_bwd: bad arguments");
  if (R == 0) return ZEST_OK;
  const float scale = 1.f / (float)((double)R * (n_close - 1) * 3.0);
  sf_smooth_bwd_kernel<<<(unsigned)((R * S + 255) / 256), 256, 0, (cudaStream_t)stream>>>(pts_1, pts_2, R, S, n_close, make_consts(H, W, f),
                                                                                           g_loss, scale, g_pts_1, g_pts_2);
  ZEST_LAUNCH_CHECK();
  return ZEST_OK;
}

extern "C" int zest_sf_lke_loss_fwd(const float* pts_ref, const float* pts_post, const float* pts_prev, int64_t R, int S, int n_close,
                                    int H, int W, float f, double* loss_sum, void* stream) {
  ZEST_CHECK_ARG(pts_ref && pts_post && pts_prev && loss_sum && R >= 0 && S >= 1 && n_close >= 1 && n_close <= S && f != 0.f,
                 "zest_sf_lke_loss_fwd: bad arguments");
  const int64_t n = R * n_close;
  if (n == 0) return ZEST_OK;
  sf_lke_fwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(pts_ref, pts_post, pts_prev, R, S, n_close,
                                                                                    make_consts(H, W, f), loss_sum);
  ZEST_LAUNCH_CHECK();
  return ZEST_OK;
}

extern "C" int zest_sf_lke_loss_bwd(const float* pts_ref, const float* pts_post, const float* pts_prev, int64_t R, int S, int n_close,
                                    int H, int W, float f, const float* g_loss, float* g_ref, float* g_post, float* g_prev, void* stream) {
  ZEST_CHECK_ARG(pts_ref && pts_post && pts_prev && g_loss && R >= 0 && S >= 1 && n_close >= 1 && n_close <= S && f != 0.f,
                 "zest_sf_lke_loss_bwd: bad arguments");
  if (R == 0) return ZEST_OK;
  const float scale = 1.f / (float)((double)R * n_close * 3.0);
  sf_lke_bwd_kernel<<<(unsigned)((R * S + 255) / 256), 256, 0, (cudaStream_t)stream>>>(pts_ref, pts_post, pts_prev, R, S, n_close,
                                                                                        make_consts(H, W, f), g_loss, scale, g_ref, g_post, g_prev);
  ZEST_LAUNCH_CHECK();
  return ZEST_OK;
}

extern "C" int zest_project_ndc_fwd(const float* w2c, const float* weights, const float* raw_pts, int64_t R, int S, int H, int W,
                                    float f, float* pts_2d, void* stream) {
  ZEST_CHECK_ARG(w2c && weights && raw_pts && pts_2d && R >= 0 && S >= 1 && f != 0.f, "zest_project_ndc_fwd: bad arguments");
  if (R == 0) return ZEST_OK;
  project_ndc_fwd_kernel<<<(unsigned)((R * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(weights, raw_pts, R, S, w2c,
                                                                                              make_consts(H, W, f), pts_2d);
  ZEST_LAUNCH_CHECK();
  return ZEST_OK;
}

extern "C" int zest_project_ndc_bwd(const float* w2c, const float* weights, const float* raw_pts, int64_t R, int S, int H, int W,
                                    float f, const float* g_pts_2d, float* g_weights, float* g_raw_pts, void* stream) {
  ZEST_CHECK_ARG(w2c && weights && raw_pts && g_pts_2d && R >= 0 && S >= 1 && f != 0.f, "zest_project_ndc_bwd: bad arguments");
  if (R == 0) return ZEST_OK;
  project_ndc_bwd_kernel<<<(unsigned)((R * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(weights, raw_pts, R, S, w2c,
                                                                                              make_consts(H, W, f), g_pts_2d, g_weights, g_raw_pts);
  ZEST_LAUNCH_CHECK();
  return ZEST_OK;
}
