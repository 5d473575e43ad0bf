// "Next" row f3, first half: the plane-sweep cost volume that feeds the encoding-volume CNN -
// networks.py:1077-1140 MVSNet.build_volume_cost + utils.py:49-99 homo_warp, in one pass.
//
// The reference materialises, per source view, the [B, 3, D*Hp*Wp] homography grid, a [C, D, Hp, Wp] warped feature volume
// (grid_sample), its square, two running sums, the masks and finally the variance: ~30 PyTorch kernels and ~10 volume-sized
// round trips.  Here one thread owns one voxel (d, y, x) of the padded reference frustum: it projects the voxel into every
// source view (R (x - pad, y - pad, 1) + T / depth, utils.py:77-89), gathers the 4 bilinear corners of the feature map (laid
// out as planes of channel quads; the maps are a few MB and stay in L1 / L2), keeps sum and sum of
// squares in registers and writes the image channels, the variance and the in-frustum masks exactly once.
// HBM-bound on its output: (3 V + C + V) x 4 bytes per voxel.  The backward (wrt the feature maps) recomputes the taps per
// voxel and scatters with vector atomics.
#include "common.cuh"

namespace zest {
namespace {

constexpr int kMaxSrc = 9;   // source views besides the reference (num_keyframes = 10)

struct Tap {              // bilinear footprint of one voxel in one source view (zeros padding, align_corners=True)
  int off[4];             // pixel offsets (y * W + x) of nw, ne, sw, se; -1 = outside
  float w[4];
  float mask;             // -1 < grid < 1 on both axes (networks.py:1123-1126)
};

// ATen grid_sampler_2d (bilinear, zeros, align_corners=True) at normalised (gx, gy): GridSampler.cuh:23-31,220-227
__device__ __forceinline__ void make_tap(float gx, float gy, int H, int W, Tap& t) {
  t.mask = (gx > -1.f && gx < 1.f && gy > -1.f && gy < 1.f) ? 1.f : 0.f;
  const float ix = safe_int_range(unnormalize(gx, W)), iy = safe_int_range(unnormalize(gy, H));
  const float fx = floorf(ix), fy = floorf(iy);
  const int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
  const float wx1 = ix - fx, wy1 = iy - fy;
  const float wx0 = (fx + 1.f) - ix, wy0 = (fy + 1.f) - iy;
  t.w[0] = wx0 * wy0; t.w[1] = wx1 * wy0; t.w[2] = wx0 * wy1; t.w[3] = wx1 * wy1;
  const bool xin0 = x0 >= 0 && x0 < W, xin1 = x1 >= 0 && x1 < W, yin0 = y0 >= 0 && y0 < H, yin1 = y1 >= 0 && y1 < H;
  t.off[0] = (xin0 && yin0) ? y0 * W + x0 : -1;
  t.off[1] = (xin1 && yin0) ? y0 * W + x1 : -1;
  t.off[2] = (xin0 && yin1) ? y1 * W + x0 : -1;
  t.off[3] = (xin1 && yin1) ? y1 * W + x1 : -1;
}

__device__ __forceinline__ float4 tap4(const float* __restrict__ base, int stride, const Tap& t) {   // 4 consecutive channels
  float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (t.off[k] < 0) continue;
    const float4 v = __ldg(reinterpret_cast<const float4*>(base + (int64_t)t.off[k] * stride));
    o.x += v.x * t.w[k]; o.y += v.y * t.w[k]; o.z += v.z * t.w[k]; o.w += v.w * t.w[k];
  }
  return o;
}

// proj [V - 1][12] (device): rows of src_proj @ ref_proj_inv, staged once per block
__device__ __forceinline__ void load_proj(const float* __restrict__ proj, int nsrc, float* s_proj) {
  for (int i = threadIdx.x; i < nsrc * 12; i += blockDim.x) s_proj[i] = __ldg(proj + i);
  __syncthreads();
}

// utils.py:77-89: src = R (x, y, 1) + T / depth; uv = src.xy / src.z; grid = uv / ((size - 1) / 2) - 1
__device__ __forceinline__ void sweep_tap(const float* __restrict__ pm, float gx0, float gy0, float dep, int H, int W, Tap& t) {
  float s[3];
#pragma unroll
  for (int j = 0; j < 3; ++j)
    s[j] = __fadd_rn(dot3(gx0, gy0, 1.f, pm[4 * j], pm[4 * j + 1], pm[4 * j + 2]), __fdiv_rn(pm[4 * j + 3], dep));
  const float u = __fdiv_rn(s[0], s[2]), w = __fdiv_rn(s[1], s[2]);
  const float gx = __fsub_rn(__fdiv_rn(u, (float)(W - 1) / 2.f), 1.f), gy = __fsub_rn(__fdiv_rn(w, (float)(H - 1) / 2.f), 1.f);
  make_tap(gx, gy, H, W, t);
}

// feats_cl [V, C / 4, H, W, 4]: channel quads as planes of float4 pixels - the 32 lanes of a warp (x-adjacent voxels) then read
// 32 neighbouring float4 of one plane (4 lines) instead of one float4 out of 32 different 128-byte pixels.
// imgs_cl [V, H, W, 4] (r, g, b, 0) at feature resolution, depth [D].
// Output, 9 + C channels whatever V is (networks.py:1101: `torch.empty((B, 9 + 32, ...))`; the warped images of source
// views beyond the second land in channels >= 9 and are overwritten by the variance, `:1138`):
//   CL = false: img_feat [9 + C, D, Hp, Wp] (the reference's layout), in_masks [V, D, Hp, Wp]
//   CL = true:  img_feat [D, Hp, Wp, cpad] channels-last (cpad >= 9 + C, pad channels zero) for the 3-D CNN, 128-bit stores
// The views are walked one after the other with the running sum / sum of squares of all C channels in registers, so any
// number of source views costs the same registers (num_keyframes = 10 -> 9 source views).
template <int CQ, bool CL>
__global__ void __launch_bounds__(256) cost_volume_kernel(const float* __restrict__ feats_cl, const float* __restrict__ imgs_cl,
                                                          const float* __restrict__ proj, const float* __restrict__ depth, int nsrc,
                                                          int H, int W, int D, int pad, int cpad, float* __restrict__ img_feat,
                                                          float* __restrict__ in_masks) {
  __shared__ float s_proj[kMaxSrc * 12];
  // channels-last output: every warp assembles its 32 voxels x cpad floats here and stores them as one contiguous run
  __shared__ __align__(16) float s_out[CL ? 8 * 32 * (12 + 4 * CQ) : 4];
  load_proj(proj, nsrc, s_proj);
  constexpr int C = 4 * CQ;
  const int Hp = H + 2 * pad, Wp = W + 2 * pad;
  const int64_t plane = (int64_t)Hp * Wp, vol = plane * D;
  const int64_t idx_raw = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (!CL && idx_raw >= vol) return;
  const bool live = idx_raw < vol;
  const int64_t idx = live ? idx_raw : vol - 1;      // CL: a dead lane recomputes the last voxel and only helps with the stores
  const int d = (int)(idx / plane);
  const int rem = (int)(idx - (int64_t)d * plane);
  const int y = rem / Wp, x = rem - y * Wp;
  // utils.py:66-75: pixel grid of the padded reference frame, shifted by -pad
  const float gx0 = (float)x - (float)pad, gy0 = (float)y - (float)pad;
  const float dep = __ldg(depth + d);
  const bool inside = x >= pad && x < W + pad && y >= pad && y < H + pad;     // F.pad(ref_feats, pad) is zero outside
  const int ref_off = inside ? (y - pad) * W + (x - pad) : -1;
  const int64_t vstride = (int64_t)CQ * H * W * 4;

  float4 sum[CQ], sq[CQ];
#pragma unroll
  for (int q = 0; q < CQ; ++q) {
    float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
    if (inside) f = __ldg(reinterpret_cast<const float4*>(feats_cl + (int64_t)q * H * W * 4 + (int64_t)ref_off * 4));
    sum[q] = f; sq[q] = make_float4(f.x * f.x, f.y * f.y, f.z * f.z, f.w * f.w);
  }
  // image channels: reference view (zero outside the unpadded window; the reference leaves that border uninitialised,
  // networks.py:1101-1103), then the first two warped source views
  float img9[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (inside) {
    const float4 c0 = __ldg(reinterpret_cast<const float4*>(imgs_cl + (int64_t)ref_off * 4));
    img9[0] = c0.x; img9[1] = c0.y; img9[2] = c0.z;
  }
  float cnt = 1.f;
  if (!CL && in_masks) in_masks[idx] = 1.f;
  for (int v = 0; v < nsrc; ++v) {
    Tap tap;
    sweep_tap(s_proj + 12 * v, gx0, gy0, dep, H, W, tap);
    cnt += tap.mask;
    if (!CL && in_masks) in_masks[(int64_t)(v + 1) * vol + idx] = tap.mask;
    if (v < 2) {
      const float4 cv = tap4(imgs_cl + (int64_t)(v + 1) * H * W * 4, 4, tap);
      if (v == 0) { img9[3] = cv.x; img9[4] = cv.y; img9[5] = cv.z; }     // static indices: img9 stays in registers
      else { img9[6] = cv.x; img9[7] = cv.y; img9[8] = cv.z; }
    }
    const float* fv = feats_cl + (int64_t)(v + 1) * vstride;
#pragma unroll
    for (int q = 0; q < CQ; ++q) {
      const float4 g = tap4(fv + (int64_t)q * H * W * 4, 4, tap);
      sum[q].x += g.x; sum[q].y += g.y; sum[q].z += g.z; sum[q].w += g.w;
      sq[q].x += g.x * g.x; sq[q].y += g.y * g.y; sq[q].z += g.z * g.z; sq[q].w += g.w * g.w;
    }
  }
  const float inv = __fdiv_rn(1.0f, cnt);     // networks.py:1137
  // variance of the feature channels over the views (networks.py:1105-1138)
  float var[C];
#pragma unroll
  for (int q = 0; q < CQ; ++q) {
    const float mx = sum[q].x * inv, my = sum[q].y * inv, mz = sum[q].z * inv, mw = sum[q].w * inv;
    var[4 * q + 0] = __fsub_rn(__fmul_rn(sq[q].x, inv), __fmul_rn(mx, mx));
    var[4 * q + 1] = __fsub_rn(__fmul_rn(sq[q].y, inv), __fmul_rn(my, my));
    var[4 * q + 2] = __fsub_rn(__fmul_rn(sq[q].z, inv), __fmul_rn(mz, mz));
    var[4 * q + 3] = __fsub_rn(__fmul_rn(sq[q].w, inv), __fmul_rn(mw, mw));
  }
  if (CL) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cq = cpad >> 2;
    float4* mine = reinterpret_cast<float4*>(s_out) + (warp * 32 + lane) * cq;
    float row[12 + C];     // [9 image channels | C variances | zero pad], cpad / 4 vectors
#pragma unroll
    for (int i = 0; i < 9; ++i) row[i] = img9[i];
#pragma unroll
    for (int i = 0; i < C; ++i) row[9 + i] = var[i];
    row[9 + C] = 0.f; row[10 + C] = 0.f; row[11 + C] = 0.f;
#pragma unroll
    for (int i = 0; i < (12 + C) / 4; ++i)
      if (i < cq) mine[i] = make_float4(row[4 * i], row[4 * i + 1], row[4 * i + 2], row[4 * i + 3]);
    __syncwarp();
    const int64_t warp_base = idx_raw - lane;                       // first voxel of this warp
    const int64_t n_live = vol - warp_base < 32 ? vol - warp_base : 32;
    const float4* src = reinterpret_cast<const float4*>(s_out) + warp * 32 * cq;
    float4* dst = reinterpret_cast<float4*>(img_feat) + warp_base * cq;
    for (int e = lane; e < (int)n_live * cq; e += 32) dst[e] = src[e];  // 32 lanes x 16 B contiguous per instruction
  } else {
#pragma unroll
    for (int i = 0; i < 9; ++i) img_feat[(int64_t)i * vol + idx] = img9[i];
#pragma unroll
    for (int i = 0; i < C; ++i) img_feat[(int64_t)(9 + i) * vol + idx] = var[i];
  }
}

// Backward wrt the feature maps (the images, projections and depths are data).  var_c = sq inv - (sum inv)^2 with
// sum = f_ref + sum_v g_v, sq = f_ref^2 + sum_v g_v^2  =>  d var / d x = 2 inv (x - sum inv) for x in {f_ref, g_v}; g_v is the
// bilinear tap, so its gradient scatters to the 4 corners with the tap weights.  One thread per voxel recomputes the taps and
// issues 128-bit vector atomics into the gradient maps (same planes-of-quads layout as the features).
__device__ __forceinline__ void scatter4(float* __restrict__ base, const Tap& t, float4 g) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (t.off[k] < 0) continue;
    atomicAdd(reinterpret_cast<float4*>(base + (int64_t)t.off[k] * 4), make_float4(g.x * t.w[k], g.y * t.w[k], g.z * t.w[k], g.w * t.w[k]));
  }
}

template <int NSRC>
__global__ void __launch_bounds__(256) cost_volume_bwd_kernel(const float* __restrict__ feats_cl, const float* __restrict__ proj,
                                                              const float* __restrict__ depth, int C, int H, int W, int D, int pad,
                                                              const float* __restrict__ g_var, float* __restrict__ g_feats_cl) {
  __shared__ float s_proj[kMaxSrc * 12];
  load_proj(proj, NSRC, s_proj);
  const int Hp = H + 2 * pad, Wp = W + 2 * pad;
  const int64_t plane = (int64_t)Hp * Wp, vol = plane * D;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= vol) return;
  const int d = (int)(idx / plane);
  const int rem = (int)(idx - (int64_t)d * plane);
  const int y = rem / Wp, x = rem - y * Wp;
  const float gx0 = (float)x - (float)pad, gy0 = (float)y - (float)pad;
  const float dep = __ldg(depth + d);
  const bool inside = x >= pad && x < W + pad && y >= pad && y < H + pad;
  const int ref_off = inside ? (y - pad) * W + (x - pad) : -1;
  Tap tap[NSRC];
  float cnt = 1.f;
#pragma unroll
  for (int v = 0; v < NSRC; ++v) {
    sweep_tap(s_proj + 12 * v, gx0, gy0, dep, H, W, tap[v]);
    cnt += tap[v].mask;
  }
  const float inv = __fdiv_rn(1.0f, cnt);
  const int64_t vstride = (int64_t)(C >> 2) * H * W * 4;
  for (int c = 0; c < C; c += 4) {
    const float* plane0 = feats_cl + (int64_t)(c >> 2) * H * W * 4;
    float* gplane0 = g_feats_cl + (int64_t)(c >> 2) * H * W * 4;
    float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
    if (inside) f = __ldg(reinterpret_cast<const float4*>(plane0 + (int64_t)ref_off * 4));
    float4 g[NSRC];
    float4 sum = f;
#pragma unroll
    for (int v = 0; v < NSRC; ++v) {
      g[v] = tap4(plane0 + (v + 1) * vstride, 4, tap[v]);
      sum.x += g[v].x; sum.y += g[v].y; sum.z += g[v].z; sum.w += g[v].w;
    }
    const float4 go = make_float4(__ldg(g_var + (int64_t)(c + 0) * vol + idx), __ldg(g_var + (int64_t)(c + 1) * vol + idx),
                                  __ldg(g_var + (int64_t)(c + 2) * vol + idx), __ldg(g_var + (int64_t)(c + 3) * vol + idx));
    const float k2 = 2.f * inv;
    const float4 mean = make_float4(sum.x * inv, sum.y * inv, sum.z * inv, sum.w * inv);
    if (inside)
      atomicAdd(reinterpret_cast<float4*>(gplane0 + (int64_t)ref_off * 4),
                make_float4(k2 * go.x * (f.x - mean.x), k2 * go.y * (f.y - mean.y), k2 * go.z * (f.z - mean.z), k2 * go.w * (f.w - mean.w)));
#pragma unroll
    for (int v = 0; v < NSRC; ++v)
      scatter4(gplane0 + (v + 1) * vstride, tap[v],
               make_float4(k2 * go.x * (g[v].x - mean.x), k2 * go.y * (g[v].y - mean.y), k2 * go.z * (g[v].z - mean.z), k2 * go.w * (g[v].w - mean.w)));
  }
}

}  // namespace
}  // namespace zest

using namespace zest;

extern "C" int zest_cost_volume_fwd(const float* feats_cl, const float* imgs_cl, const float* proj, const float* depth, int V, int C,
                                    int H, int W, int D, int pad, float* img_feat, float* in_masks, int channels_last, int cpad,
                                    void* stream) {
  ZEST_CHECK_ARG(feats_cl && imgs_cl && proj && depth && img_feat, "zest_cost_volume_fwd: null argument");
  ZEST_CHECK_ARG(V >= 2 && V - 1 <= kMaxSrc && C > 0 && (C % 4) == 0 && C <= 32 && H > 1 && W > 1 && D > 0 && pad >= 0,
                 "zest_cost_volume_fwd: unsupported shape (V=%d C=%d H=%d W=%d D=%d pad=%d)", V, C, H, W, D, pad);
  ZEST_CHECK_ARG(!channels_last || (cpad >= 9 + C && (cpad % 4) == 0 && cpad <= 12 + C), "zest_cost_volume_fwd: cpad must be a multiple of 4 in [9 + C, 12 + C]");
  const int64_t vol = (int64_t)D * (H + 2 * pad) * (W + 2 * pad);
  const unsigned grid = (unsigned)((vol + 255) / 256);
  cudaStream_t st = (cudaStream_t)stream;
#define ZEST_CV(CQ)                                                                                                                  \
  case CQ:                                                                                                                           \
    if (channels_last) cost_volume_kernel<CQ, true><<<grid, 256, 0, st>>>(feats_cl, imgs_cl, proj, depth, V - 1, H, W, D, pad, cpad, img_feat, in_masks); \
    else cost_volume_kernel<CQ, false><<<grid, 256, 0, st>>>(feats_cl, imgs_cl, proj, depth, V - 1, H, W, D, pad, cpad, img_feat, in_masks);              \
    break;
  switch (C / 4) {
    ZEST_CV(1) ZEST_CV(2) ZEST_CV(4) ZEST_CV(8)
    default:
      set_error("zest_cost_volume_fwd: %d feature channels not instantiated (4, 8, 16, 32)", C);
      return ZEST_E_ARG;
  }
#undef ZEST_CV
  ZEST_LAUNCH_CHECK();
  return ZEST_OK;
}

extern "C" int zest_cost_volume_bwd(const float* feats_cl, const float* proj, const float* depth, int V, int C, int H, int W, int D,
                                    int pad, const float* g_var, float* g_feats_cl, void* stream) {
  ZEST_CHECK_ARG(feats_cl && proj && depth && g_var && g_feats_cl, "zest_cost_volume_bwd: null argument");
  ZEST_CHECK_ARG(V >= 2 && V - 1 <= 4 && C > 0 && (C % 4) == 0 && H > 1 && W > 1 && D > 0 && pad >= 0,
                 "zest_cost_volume_bwd: unsupported shape (V=%d C=%d H=%d W=%d D=%d pad=%d; at most 4 source views)", V, C, H, W, D, pad);
  const int64_t vol = (int64_t)D * (H + 2 * pad) * (W + 2 * pad);
  const unsigned grid = (unsigned)((vol + 255) / 256);
  cudaStream_t st = (cudaStream_t)stream;
  switch (V - 1) {
    case 1: cost_volume_bwd_kernel<1><<<grid, 256, 0, st>>>(feats_cl, proj, depth, C, H, W, D, pad, g_var, g_feats_cl); break;
    case 2: cost_volume_bwd_kernel<2><<<grid, 256, 0, st>>>(feats_cl, proj, depth, C, H, W, D, pad, g_var, g_feats_cl); break;
    case 3: cost_volume_bwd_kernel<3><<<grid, 256, 0, st>>>(feats_cl, proj, depth, C, H, W, D, pad, g_var, g_feats_cl); break;
    default: cost_volume_bwd_kernel<4><<<grid, 256, 0, st>>>(feats_cl, proj, depth, C, H, W, D, pad, g_var, g_feats_cl); break;
  }
  ZEST_LAUNCH_CHECK();
  return ZEST_OK;
}
