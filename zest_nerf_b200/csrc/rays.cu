// Ray builder ("next" row f1 of SURVEY.md 8f): target pixels -> the four ray tensors rendering() consumes.
//
// Replaces utils.py:133-230 (get_rays_mvs: pixel grid / given pixels -> rays_o, rays_d), utils.py:362-379
// (build_rays_base: z = near (1 - t) + far t, optional stratified jitter, pts = o + z d) and utils.py:232-288
// (get_ndc_coordinate: NDC of the reference view incl. the pad remap and its `int / Tensor` = reciprocal * int
// quirk).  Every torch op of the reference rounds separately and its K = 3 matmuls are fma chains on CPU
// (verified bit-exact); the kernel spells exactly that sequence with explicitly
// rounded intrinsics, so sample depths, world points and NDC coordinates (and therefore every voxel / pixel
// index downstream) are bit-identical to the reference's CPU path.
//
// One thread per (ray, sample).  Pure streaming writes: 28 B / sample (pts 12, ndc 12, z 4) + 12 B / ray.
#include "common.cuh"

namespace zest {

struct RayParams {
  // target pixels: ys/xs given (fp32, R each) or the row-major grid slab starting at linear pixel r0 of a W_tgt wide image
  const float* ys; const float* xs; int64_t r0; int W_tgt;
  const float* cam;               // device: [K_t 9 | c2w_t 16 | w2c_r 16 | K_r 9 | near_t far_t near_r far_r] = 54 floats

  float inv_w, inv_h;             // source W - 1, H - 1
  int pad;
  const float* t_vals;            // [S] = torch.linspace(0, 1, S)
  const float* t_rand;            // [R, S] stratified jitter or nullptr
  int64_t R; int S;
  float* pts; float* dir; float* ndc; float* z;
};

struct RayCam { float K_t[9], c2w_t[16], w2c_r[16], K_r[9], near_t, far_t, near_r, far_r; };

__device__ __forceinline__ float depth_at(const RayParams& p, const RayCam& c, int i) {   // utils.py:364  near * (1 - t) + far * t
  const float t = __ldg(p.t_vals + i);
  return __fadd_rn(__fmul_rn(c.near_t, __fsub_rn(1.f, t)), __fmul_rn(c.far_t, t));
}

__global__ void build_rays_kernel(const __grid_constant__ RayParams p) {
  __shared__ RayCam c;   // the cameras are device data (stream-ordered with whatever produced them): one copy per block
  if (threadIdx.x < 54) reinterpret_cast<float*>(&c)[threadIdx.x] = __ldg(p.cam + threadIdx.x);
  __syncthreads();
  __shared__ __align__(16) float s_pts[256 * 3], s_ndc[256 * 3];   // this block's 256 samples, written back as float4
  const int64_t m0 = (int64_t)blockIdx.x * blockDim.x;
  const int64_t total = p.R * p.S;
  const int64_t m = m0 + threadIdx.x;
  const bool live = m < total;
  if (live) {
  const int64_t r = m / p.S;
  const int s = (int)(m - r * p.S);
  float ys, xs;
  if (p.ys) { ys = __ldg(p.ys + r); xs = __ldg(p.xs + r); }
  else { const int64_t k = p.r0 + r; ys = (float)(k / p.W_tgt); xs = (float)(k % p.W_tgt); }
  // A1 (utils.py:215-223): dirs_cam = ((x - cx) / fx, (y - cy) / fy, 1) ; rays_d = dirs_cam @ c2w[:3,:3]^T
  const float dx = __fdiv_rn(__fsub_rn(xs, c.K_t[2]), c.K_t[0]), dy = __fdiv_rn(__fsub_rn(ys, c.K_t[5]), c.K_t[4]);
  float d[3], o[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    d[j] = dot3(dx, dy, 1.f, c.c2w_t[4 * j], c.c2w_t[4 * j + 1], c.c2w_t[4 * j + 2]);
    o[j] = c.c2w_t[4 * j + 3];
  }
  if (s == 0) { p.dir[r * 3] = d[0]; p.dir[r * 3 + 1] = d[1]; p.dir[r * 3 + 2] = d[2]; }
  // A2 (utils.py:362-377)
  float z = depth_at(p, c, s);
  if (p.t_rand) {
    const float lower = s == 0 ? z : __fmul_rn(0.5f, __fadd_rn(z, depth_at(p, c, s - 1)));
    const float upper = s == p.S - 1 ? z : __fmul_rn(0.5f, __fadd_rn(depth_at(p, c, s + 1), z));
    z = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), __ldg(p.t_rand + m)));
  }
  p.z[m] = z;
  // A3 (utils.py:379): pts = o + z * d
  float w[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) { w[j] = __fadd_rn(o[j], __fmul_rn(z, d[j])); s_pts[threadIdx.x * 3 + j] = w[j]; }
  // A4 (utils.py:257-285)
  float pc[3], q[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) pc[j] = __fadd_rn(dot3(w[0], w[1], w[2], c.w2c_r[4 * j], c.w2c_r[4 * j + 1], c.w2c_r[4 * j + 2]), c.w2c_r[4 * j + 3]);
#pragma unroll
  for (int j = 0; j < 3; ++j) q[j] = dot3(pc[0], pc[1], pc[2], c.K_r[3 * j], c.K_r[3 * j + 1], c.K_r[3 * j + 2]);
  float u = __fdiv_rn(__fadd_rn(__fdiv_rn(q[0], q[2]), 0.f), p.inv_w);
  float v = __fdiv_rn(__fadd_rn(__fdiv_rn(q[1], q[2]), 0.f), p.inv_h);
  const float zn = __fdiv_rn(__fsub_rn(q[2], c.near_r), __fsub_rn(c.far_r, c.near_r));
  if (p.pad > 0) {
    const float Wf = __fdiv_rn(__fadd_rn(p.inv_w, 1.f), 4.f), Hf = __fdiv_rn(__fadd_rn(p.inv_h, 1.f), 4.f);
    const float Wp = __fadd_rn(Wf, (float)(2 * p.pad)), Hp = __fadd_rn(Hf, (float)(2 * p.pad));
    // `pad / tensor` dispatches to reciprocal() * pad in the reference (quirk C2)
    v = __fadd_rn(__fdiv_rn(__fmul_rn(v, Hf), Hp), __fmul_rn(__frcp_rn(Hp), (float)p.pad));
    u = __fadd_rn(__fdiv_rn(__fmul_rn(u, Wf), Wp), __fmul_rn(__frcp_rn(Wp), (float)p.pad));
  }
  s_ndc[threadIdx.x * 3] = u; s_ndc[threadIdx.x * 3 + 1] = v; s_ndc[threadIdx.x * 3 + 2] = zn;
  }
  __syncthreads();
  // 256 samples x 3 floats = 192 float4 per tensor, contiguous in global memory (m0 * 3 floats is 16-byte aligned)
  const int64_t n_here = (total - m0 < 256 ? total - m0 : 256) * 3;
  for (int i = threadIdx.x; i < 192; i += 256) {
    if ((int64_t)(i + 1) * 4 <= n_here) {
      reinterpret_cast<float4*>(p.pts + m0 * 3)[i] = reinterpret_cast<const float4*>(s_pts)[i];
      reinterpret_cast<float4*>(p.ndc + m0 * 3)[i] = reinterpret_cast<const float4*>(s_ndc)[i];
    } else {
      for (int j = i * 4; j < n_here; ++j) { p.pts[m0 * 3 + j] = s_pts[j]; p.ndc[m0 * 3 + j] = s_ndc[j]; }
    }
  }
}

}  // namespace zest

using namespace zest;

extern "C" int zest_build_rays(const float* ys, const float* xs, int64_t r0, int W_tgt, const float* cam, int W_src, int H_src,
                               int pad, const float* t_vals, const float* t_rand, int64_t R, int S, float* rays_pts,
                               float* rays_dir, float* rays_ndc, float* depth, void* stream) {
  ZEST_CHECK_ARG(cam && t_vals && rays_pts && rays_dir && rays_ndc && depth, "zest_build_rays: null argument");
  ZEST_CHECK_ARG((ys == nullptr) == (xs == nullptr), "zest_build_rays: give both ys and xs or neither");
  ZEST_CHECK_ARG(R >= 0 && S > 0 && W_src > 1 && H_src > 1 && pad >= 0 && (ys || W_tgt > 0), "zest_build_rays: bad sizes");
  if (R == 0) return ZEST_OK;
  RayParams p{};
  p.ys = ys; p.xs = xs; p.r0 = r0; p.W_tgt = W_tgt; p.cam = cam;
  p.inv_w = (float)(W_src - 1); p.inv_h = (float)(H_src - 1); p.pad = pad;
  p.t_vals = t_vals; p.t_rand = t_rand; p.R = R; p.S = S;
  p.pts = rays_pts; p.dir = rays_dir; p.ndc = rays_ndc; p.z = depth;
  const int64_t n = R * S;
  ZEST_CHECK_ARG((n + 255) / 256 < (1ll << 31), "zest_build_rays: too many samples for one launch");
  build_rays_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p);
  ZEST_LAUNCH_CHECK();
  return ZEST_OK;
}
