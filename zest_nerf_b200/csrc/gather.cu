// Fused feature gather: encoding-volume trilinear sample + per-view bilinear RGB gather + mask.
//
// Replaces utils.py:433-459 (index_point_feature -> F.grid_sample 5-D), utils.py:461-505
// (build_color_volume -> projection utils.py:257-269 + F.grid_sample 4-D, border) and the concat of
// renderer.py:51-72.  Index arithmetic follows ATen's grid sampler op by op with explicitly
// rounded intrinsics (no FMA contraction) so the integer voxel / pixel corners are bit-identical
// to the reference running on CPU; the K=3 projections use the fma chain of ATen's CPU matmul.
//
// Layout: volumes are repacked once per frame to channels-last [D,H,W,8] fp32 so that one voxel
// corner is exactly one 32-byte sector (2 x LDG.128); images to [V,H,W,4] fp32 (1 x LDG.128 per
// corner).  Thread mapping: lane -> ray, warp -> sample slot, so the 32 lanes of a warp touch 32
// horizontally adjacent target pixels at the same depth index: their voxels / source pixels are
// neighbours in memory (4 target pixels per voxel in x) and share cache lines in L1/L2.
#include "gather_core.cuh"

namespace zest {

__global__ void pack_volume_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t n_vox) {
  // src [8][n_vox] -> dst [n_vox][8]; each thread handles one voxel: 8 coalesced reads (one per
  // channel plane), two 128-bit writes.
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_vox) return;
  float v[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) v[c] = __ldg(src + c * n_vox + i);
  float4* o = reinterpret_cast<float4*>(dst + i * 8);
  o[0] = make_float4(v[0], v[1], v[2], v[3]);
  o[1] = make_float4(v[4], v[5], v[6], v[7]);
}

__global__ void unpack_volume_grad_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t n_vox) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_vox) return;
  const float4* s = reinterpret_cast<const float4*>(src + i * 8);
  float4 a = s[0], b = s[1];
  float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
  for (int c = 0; c < 8; ++c) dst[c * n_vox + i] += v[c];
}

__global__ void pack_images_kernel(const float* __restrict__ src, float* __restrict__ dst, int V, int64_t hw) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)V * hw) return;
  int64_t v = i / hw, p = i - v * hw;
  const float* s = src + v * 3 * hw + p;
  reinterpret_cast<float4*>(dst)[i] = make_float4(__ldg(s), __ldg(s + hw), __ldg(s + 2 * hw), 0.f);
}

struct GatherParams {
  const float* pts;
  const float* ndc;
  int ndc_ld;
  int64_t R;
  int S;
  const float* vol;
  int D, Hv, Wv;
  const float* img;
  int V, H, W;
  const float* cams;
  float* feats;
  int ldf;
  int32_t* vox_idx;
  int32_t* pix_idx;
};

constexpr int kGatherWarps = 8;

template <bool HAS_VOL, bool HAS_IMG>
__global__ void __launch_bounds__(32 * kGatherWarps) gather_fwd_kernel(GatherParams p) {
  extern __shared__ float s_cams[];  // [V][24]
  if (HAS_IMG) {
    for (int i = threadIdx.x; i < p.V * 24; i += blockDim.x) s_cams[i] = __ldg(p.cams + i);
    __syncthreads();
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int s_blocks = (p.S + kGatherWarps - 1) / kGatherWarps;
  const int64_t ray = (int64_t)(blockIdx.x / s_blocks) * 32 + lane;
  const int s = (blockIdx.x % s_blocks) * kGatherWarps + warp;
  if (ray >= p.R || s >= p.S) return;
  const int64_t m = ray * p.S + s;
  float* out = p.feats + m * p.ldf;

  if (HAS_VOL) {
    const float* n = p.ndc + m * p.ndc_ld;
    float acc[8];
    trilinear8(p.vol, p.D, p.Hv, p.Wv, __ldg(n), __ldg(n + 1), __ldg(n + 2), acc, p.vox_idx ? p.vox_idx + m * 3 : nullptr);
    if ((p.ldf & 3) == 0) {
      reinterpret_cast<float4*>(out)[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
      reinterpret_cast<float4*>(out)[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) out[k] = acc[k];
    }
  }

  if (HAS_IMG) {
    const float* q = p.pts + m * 3;
    const float px = __ldg(q), py = __ldg(q + 1), pz = __ldg(q + 2);
    for (int v = 0; v < p.V; ++v) {
      const float4 f = view_sample(reinterpret_cast<const float4*>(p.img) + (int64_t)v * p.H * p.W, p.H, p.W, s_cams + v * 24,
                                   px, py, pz, p.pix_idx ? p.pix_idx + (m * p.V + v) * 2 : nullptr);
      if ((p.ldf & 3) == 0) {
        reinterpret_cast<float4*>(out + 8)[v] = f;
      } else {
        out[8 + 4 * v] = f.x;
        out[9 + 4 * v] = f.y;
        out[10 + 4 * v] = f.z;
        out[11 + 4 * v] = f.w;
      }
    }
  }
}

// Backward of the trilinear part.  One thread per sample: scatter w * g into the 8 corners
// (128-bit vector atomics: 2 per corner) and reduce d out / d coord over corners and channels.
__global__ void gather_bwd_kernel(const float* __restrict__ ndc, int ndc_ld, int64_t M,
                                  const float* __restrict__ vol, int D, int Hv, int Wv,
                                  const float* __restrict__ gfeats, int ldf, float* gvol,
                                  float* gndc, int gndc_ld) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const float* n = ndc + m * ndc_ld;
  const float ix = safe_int_range(unnormalize(__fsub_rn(__fmul_rn(__ldg(n), 2.f), 1.f), Wv));
  const float iy = safe_int_range(unnormalize(__fsub_rn(__fmul_rn(__ldg(n + 1), 2.f), 1.f), Hv));
  const float iz = safe_int_range(unnormalize(__fsub_rn(__fmul_rn(__ldg(n + 2), 2.f), 1.f), D));
  const float fx0 = floorf(ix), fy0 = floorf(iy), fz0 = floorf(iz);
  const int x0 = (int)fx0, y0 = (int)fy0, z0 = (int)fz0;
  const float wx[2] = {(fx0 + 1.f) - ix, ix - fx0};
  const float wy[2] = {(fy0 + 1.f) - iy, iy - fy0};
  const float wz[2] = {(fz0 + 1.f) - iz, iz - fz0};
  float g[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) g[k] = __ldg(gfeats + m * ldf + k);
  float gix = 0.f, giy = 0.f, giz = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int dx = k & 1, dy = (k >> 1) & 1, dz = k >> 2;
    const int x = x0 + dx, y = y0 + dy, z = z0 + dz;
    if (!((unsigned)x < (unsigned)Wv && (unsigned)y < (unsigned)Hv && (unsigned)z < (unsigned)D)) continue;
    const int64_t off = (((int64_t)z * Hv + y) * Wv + x) * 8;
    const float w = wx[dx] * wy[dy] * wz[dz];
    if (gvol) {
      float4* gv = reinterpret_cast<float4*>(gvol + off);
      atomicAdd(gv, make_float4(w * g[0], w * g[1], w * g[2], w * g[3]));
      atomicAdd(gv + 1, make_float4(w * g[4], w * g[5], w * g[6], w * g[7]));
    }
    if (gndc) {
      const float4* q = reinterpret_cast<const float4*>(vol + off);
      const float4 a = __ldg(q), b = __ldg(q + 1);
      const float dot = a.x * g[0] + a.y * g[1] + a.z * g[2] + a.w * g[3] + b.x * g[4] + b.y * g[5] +
                        b.z * g[6] + b.w * g[7];
      gix += (dx ? 1.f : -1.f) * wy[dy] * wz[dz] * dot;
      giy += (dy ? 1.f : -1.f) * wx[dx] * wz[dz] * dot;
      giz += (dz ? 1.f : -1.f) * wx[dx] * wy[dy] * dot;
    }
  }
  if (gndc) {
    // d i / d ndc = (size - 1): the *2 of `ndc*2-1` cancels the /2 of the un-normalisation
    gndc[m * gndc_ld + 0] = gix * (float)(Wv - 1);
    gndc[m * gndc_ld + 1] = giy * (float)(Hv - 1);
    gndc[m * gndc_ld + 2] = giz * (float)(D - 1);
  }
}

// per ray: cos_angle = |d| (renderer.py:604), dirs = (d / |d|) @ R^T (renderer.py:258,46)
__global__ void dirfeat_kernel(const float* __restrict__ rays_dir, int64_t R, const float* __restrict__ cam,
                               float* cos_angle, float* dirs) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const float dx = rays_dir[r * 3], dy = rays_dir[r * 3 + 1], dz = rays_dir[r * 3 + 2];
  const float c = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
  const float nx = __fdiv_rn(dx, c), ny = __fdiv_rn(dy, c), nz = __fdiv_rn(dz, c);
  cos_angle[r] = c;
  dirs[r * 3 + 0] = dot3(nx, ny, nz, cam[0], cam[1], cam[2]);
  dirs[r * 3 + 1] = dot3(nx, ny, nz, cam[4], cam[5], cam[6]);
  dirs[r * 3 + 2] = dot3(nx, ny, nz, cam[8], cam[9], cam[10]);
}

}  // namespace zest

using namespace zest;

extern "C" int zest_pack_volume(const float* src, float* dst, int D, int H, int W, void* stream) {
  ZEST_CHECK_ARG(src && dst && D > 0 && H > 0 && W > 0, "zest_pack_volume: bad arguments");
  const int64_t n = (int64_t)D * H * W;
  pack_volume_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, dst, n);
  ZEST_LAUNCH_CHECK();
  return ZEST_OK;
}

extern "C" int zest_unpack_volume_grad(const float* src, float* dst, int D, int H, int W, void* stream) {
  ZEST_CHECK_ARG(src && dst && D > 0 && H > 0 && W > 0, "zest_unpack_volume_grad: bad arguments");
  const int64_t n = (int64_t)D * H * W;
  unpack_volume_grad_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, dst, n);
  ZEST_LAUNCH_CHECK();
  return ZEST_OK;
}

extern "C" int zest_pack_images(const float* src, float* dst, int V, int H, int W, void* stream) {
  ZEST_CHECK_ARG(src && dst && V > 0 && H > 0 && W > 0, "zest_pack_images: bad arguments");
  const int64_t n = (int64_t)V * H * W;
  pack_images_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, dst, V, (int64_t)H * W);
  ZEST_LAUNCH_CHECK();
  return ZEST_OK;
}

extern "C" int zest_gather_fwd(const float* rays_pts, const float* rays_ndc, int ndc_ld, int64_t R, int S,
                               const float* vol_cl, int D, int Hv, int Wv, const float* img_cl, int V,
                               int H, int W, const float* cams, float* feats, int ldf,
                               int32_t* vox_idx, int32_t* pix_idx, void* stream) {
  ZEST_CHECK_ARG(R >= 0 && S > 0 && feats, "zest_gather_fwd: bad sizes");
  ZEST_CHECK_ARG(vol_cl || img_cl, "zest_gather_fwd: neither a volume nor images given");
  ZEST_CHECK_ARG(!vol_cl || (rays_ndc && ndc_ld >= 3 && D > 0 && Hv > 0 && Wv > 0), "zest_gather_fwd: bad volume arguments");
  ZEST_CHECK_ARG(!img_cl || (rays_pts && cams && V > 0 && V <= 64 && H > 0 && W > 0), "zest_gather_fwd: bad image arguments");
  ZEST_CHECK_ARG(ldf >= (img_cl ? 8 + 4 * V : 8), "zest_gather_fwd: ldf too small");
  if (R == 0) return ZEST_OK;
  GatherParams p{rays_pts, rays_ndc, ndc_ld, R, S, vol_cl, D, Hv, Wv, img_cl, V, H, W, cams, feats, ldf, vox_idx, pix_idx};
  const int64_t blocks = ((R + 31) / 32) * ((S + kGatherWarps - 1) / kGatherWarps);
  ZEST_CHECK_ARG(blocks < (1ll << 31), "zest_gather_fwd: too many samples for one launch");
  const size_t smem = img_cl ? (size_t)V * 24 * sizeof(float) : 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (vol_cl && img_cl)
    gather_fwd_kernel<true, true><<<(unsigned)blocks, 32 * kGatherWarps, smem, st>>>(p);
  else if (vol_cl)
    gather_fwd_kernel<true, false><<<(unsigned)blocks, 32 * kGatherWarps, smem, st>>>(p);
  else
    gather_fwd_kernel<false, true><<<(unsigned)blocks, 32 * kGatherWarps, smem, st>>>(p);
  ZEST_LAUNCH_CHECK();
  return ZEST_OK;
}

extern "C" int zest_gather_bwd(const float* rays_ndc, int ndc_ld, int64_t M, const float* vol_cl, int D,
                               int Hv, int Wv, const float* gfeats, int ldf, float* gvol_cl, float* gndc,
                               int gndc_ld, void* stream) {
  ZEST_CHECK_ARG(rays_ndc && vol_cl && gfeats && M >= 0 && ndc_ld >= 3 && ldf >= 8, "zest_gather_bwd: bad arguments");
  ZEST_CHECK_ARG(!gndc || gndc_ld >= 3, "zest_gather_bwd: gndc_ld too small");
  if (M == 0 || (!gvol_cl && !gndc)) return ZEST_OK;
  gather_bwd_kernel<<<(unsigned)((M + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      rays_ndc, ndc_ld, M, vol_cl, D, Hv, Wv, gfeats, ldf, gvol_cl, gndc, gndc_ld);
  ZEST_LAUNCH_CHECK();
  return ZEST_OK;
}

extern "C" int zest_dirfeat_fwd(const float* rays_dir, int64_t R, const float* cam_ref, float* cos_angle,
                                float* dirs, void* stream) {
  ZEST_CHECK_ARG(rays_dir && cam_ref && cos_angle && dirs && R >= 0, "zest_dirfeat_fwd: bad arguments");
  if (R == 0) return ZEST_OK;
  dirfeat_kernel<<<(unsigned)((R + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rays_dir, R, cam_ref, cos_angle, dirs);
  ZEST_LAUNCH_CHECK();
  return ZEST_OK;
}
