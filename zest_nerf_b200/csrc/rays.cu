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
  float K_t[9], c2w_t[16];        // target intrinsics, camera-to-world
  float w2c_r[16], K_r[9];        // reference view (NDC frame)
  float near_t, far_t, near_r, far_r;
  float inv_w, inv_h;             // source W - 1, H - 1
  int pad;
  const float* t_vals;            // [S] = torch.linspace(0, 1, S)
  const float* t_rand;            // [R, S] stratified jitter or nullptr
  int64_t R; int S;
  float* pts; float* dir; float* ndc; float* z;
};

__device__ __forceinline__ float depth_at(const RayParams& p, int i) {   // utils.py:364  near * (1 - t) + far * t
  const float t = __ldg(p.t_vals + i);
  return __fadd_rn(__fmul_rn(p.near_t, __fsub_rn(1.f, t)), __fmul_rn(p.far_t, t));
}

__global__ void build_rays_kernel(const __grid_constant__ RayParams p) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= p.R * p.S) return;
  const int64_t r = m / p.S;
  const int s = (int)(m - r * p.S);
  float ys, xs;
  if (p.ys) { ys = __ldg(p.ys + r); xs = __ldg(p.xs + r); }
  else { const int64_t k = p.r0 + r; ys = (float)(k / p.W_tgt); xs = (float)(k % p.W_tgt); }
  // A1 (utils.py:215-223): dirs_cam = ((x - cx) / fx, (y - cy) / fy, 1) ; rays_d = dirs_cam @ c2w[:3,:3]^T
  const float dx = __fdiv_rn(__fsub_rn(xs, p.K_t[2]), p.K_t[0]), dy = __fdiv_rn(__fsub_rn(ys, p.K_t[5]), p.K_t[4]);
  float d[3], o[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    d[j] = dot3(dx, dy, 1.f, p.c2w_t[4 * j], p.c2w_t[4 * j + 1], p.c2w_t[4 * j + 2]);
    o[j] = p.c2w_t[4 * j + 3];
  }
  if (s == 0) { p.dir[r * 3] = d[0]; p.dir[r * 3 + 1] = d[1]; p.dir[r * 3 + 2] = d[2]; }
  // A2 (utils.py:362-377)
  float z = depth_at(p, s);
  if (p.t_rand) {
    const float lower = s == 0 ? z : __fmul_rn(0.5f, __fadd_rn(z, depth_at(p, s - 1)));
    const float upper = s == p.S - 1 ? z : __fmul_rn(0.5f, __fadd_rn(depth_at(p, s + 1), z));
    z = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), __ldg(p.t_rand + m)));
  }
  p.z[m] = z;
  // A3 (utils.py:379): pts = o + z * d
  float w[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) { w[j] = __fadd_rn(o[j], __fmul_rn(z, d[j])); p.pts[m * 3 + j] = w[j]; }
  // A4 (utils.py:257-285)
  float c[3], q[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) c[j] = __fadd_rn(dot3(w[0], w[1], w[2], p.w2c_r[4 * j], p.w2c_r[4 * j + 1], p.w2c_r[4 * j + 2]), p.w2c_r[4 * j + 3]);
#pragma unroll
  for (int j = 0; j < 3; ++j) q[j] = dot3(c[0], c[1], c[2], p.K_r[3 * j], p.K_r[3 * j + 1], p.K_r[3 * j + 2]);
  float u = __fdiv_rn(__fadd_rn(__fdiv_rn(q[0], q[2]), 0.f), p.inv_w);
  float v = __fdiv_rn(__fadd_rn(__fdiv_rn(q[1], q[2]), 0.f), p.inv_h);
  const float zn = __fdiv_rn(__fsub_rn(q[2], p.near_r), __fsub_rn(p.far_r, p.near_r));
  if (p.pad > 0) {
    const float Wf = __fdiv_rn(__fadd_rn(p.inv_w, 1.f), 4.f), Hf = __fdiv_rn(__fadd_rn(p.inv_h, 1.f), 4.f);
    const float Wp = __fadd_rn(Wf, (float)(2 * p.pad)), Hp = __fadd_rn(Hf, (float)(2 * p.pad));
    // `pad / tensor` dispatches to reciprocal() * pad in the reference (quirk C2)
    v = __fadd_rn(__fdiv_rn(__fmul_rn(v, Hf), Hp), __fmul_rn(__frcp_rn(Hp), (float)p.pad));
    u = __fadd_rn(__fdiv_rn(__fmul_rn(u, Wf), Wp), __fmul_rn(__frcp_rn(Wp), (float)p.pad));
  }
  p.ndc[m * 3] = u; p.ndc[m * 3 + 1] = v; p.ndc[m * 3 + 2] = zn;
}

}  // namespace zest

using namespace zest;

extern "C" int zest_build_rays(const float* ys, const float* xs, int64_t r0, int W_tgt, const float* K_tgt, const float* c2w_tgt,
                               const float* w2c_ref, const float* K_ref, float near_t, float far_t, float near_r, float far_r,
                               int W_src, int H_src, int pad, const float* t_vals, const float* t_rand, int64_t R, int S,
                               float* rays_pts, float* rays_dir, float* rays_ndc, float* depth, void* stream) {
  ZEST_CHECK_ARG(K_tgt && c2w_tgt && w2c_ref && K_ref && t_vals && rays_pts && rays_dir && rays_ndc && depth, "zest_build_rays: null argument");
  ZEST_CHECK_ARG((ys == nullptr) == (xs == nullptr), "zest_build_rays: give both ys and xs or neither");
  ZEST_CHECK_ARG(R >= 0 && S > 0 && W_src > 1 && H_src > 1 && pad >= 0 && (ys || W_tgt > 0), "zest_build_rays: bad sizes");
  if (R == 0) return ZEST_OK;
  RayParams p{};
  p.ys = ys; p.xs = xs; p.r0 = r0; p.W_tgt = W_tgt;
  for (int i = 0; i < 9; ++i) { p.K_t[i] = K_tgt[i]; p.K_r[i] = K_ref[i]; }
  for (int i = 0; i < 16; ++i) { p.c2w_t[i] = c2w_tgt[i]; p.w2c_r[i] = w2c_ref[i]; }
  p.near_t = near_t; p.far_t = far_t; p.near_r = near_r; p.far_r = far_r;
  p.inv_w = (float)(W_src - 1); p.inv_h = (float)(H_src - 1); p.pad = pad;
  p.t_vals = t_vals; p.t_rand = t_rand; p.R = R; p.S = S;
  p.pts = rays_pts; p.dir = rays_dir; p.ndc = rays_ndc; p.z = depth;
  const int64_t n = R * S;
  ZEST_CHECK_ARG((n + 255) / 256 < (1ll << 31), "zest_build_rays: too many samples for one launch");
  build_rays_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p);
  ZEST_LAUNCH_CHECK();
  return ZEST_OK;
}
