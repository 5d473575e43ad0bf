// "Next" row f3, second half: the encoding-volume CNNs - networks.py:935-1059 FeatureNet (2-D FPN trunk) and
// CostRegNet (3-D U-Net) with their InPlaceABN layers (batch norm + leaky ReLU 0.01) - forward, fp32, channels-last.
//
// Layout: every activation is channels-last ([N|D, H, W, C], C a multiple of 4): one voxel's channels are one or a few
// 16-byte vectors, the cost volume arrives that way from the plane-sweep kernel and the last layer leaves the encoding
// volume as [D, Hv, Wv, 8] - exactly what the ray-path gather reads (no NCDHW round trip, no repack).
//
// conv_cl_kernel: direct convolution on the CUDA cores in exact fp32 (the reference's CPU arithmetic; cuDNN would use
// TF32).  One thread owns 8 consecutive output positions along W x 8 output channels (64 accumulators): a weight vector
// fetched from shared memory (broadcast) is reused by 8 positions, an input vector by 8 channels, and along a row the 3 taps
// of a 3 x 3 (x 3) stencil reuse a sliding window of inputs - 768 FMAs per 10 global + 24 shared 128-bit loads, so the
// kernel is FMA-issue bound, not load bound.  Inputs come straight through L1/L2 (a warp's footprint is a few rows).
// The per-channel batch statistics InPlaceABN needs (the reference runs the encoders in train() mode even for
// validation, networks.py:626) are reduced in the same kernel: warp shuffle -> shared -> one double atomicAdd per block
// and channel.  bn_act_cl_kernel then normalises, applies leaky ReLU and (U-Net skips) adds the already-activated skip
// tensor in one pass.  Transposed convolutions (k 3, s 2, p 1, output_padding 1) are gathered per output parity class.
// Tried and removed (round 2, same-box A/B): staging the 3 x 18 x 66 input halo of a 16 x 64 output tile through shared
// memory four channels at a time (cp.async, conflict-free padded rows, two blocks per SM).  It cut conv0's L2 traffic from
// 22x to 3.5x the input, but its eleven load -> barrier -> compute phases per tile were not hidden by the second block:
// 1.67 ms against 1.45 ms for the L1-path kernel below.  Also tried and removed: the cost volume as planes of channel quads with
// the lanes of a warp on consecutive voxels (every load 512 contiguous bytes = 4 cache lines instead of 32: ncu shows conv0
// bound by L1 tag lookups, 31 sectors per request, 13 % L1 hit rate) - without the sliding register window it needs 2.4x the
// load instructions and came out at 1.80 ms.  And: 256-bit loads (LDG.E.256, 8 channels per request, 48-channel records) in the
// sliding window - conv0 alone 1.32 ms under ncu, but 255 registers per thread and a 9 % fatter cost volume made the whole forward
// slower in the graph replay (4.0 vs 3.6 ms).
#include <cstdlib>

#include "common.cuh"

namespace zest {
namespace {

constexpr int kCo = 8;     // output channels per thread / per weight tile
constexpr int kConvThreads = 128;

// w [cout, cin, kd, kh, kw] (Conv) or [cin, cout, kd, kh, kw] (ConvTranspose) -> packed [cout / 8][taps][cin_pad][8]
__global__ void conv_pack_weights_kernel(const float* __restrict__ w, int cout, int cin, int taps, int transposed, int cin_pad,
                                         float* __restrict__ packed) {
  const int64_t total = (int64_t)(cout / kCo) * taps * cin_pad * kCo;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(i % kCo);
    const int c = (int)((i / kCo) % cin_pad);
    const int t = (int)((i / ((int64_t)kCo * cin_pad)) % taps);
    const int ct = (int)(i / ((int64_t)kCo * cin_pad * taps));
    const int co = ct * kCo + j;
    float v = 0.f;
    if (c < cin) v = transposed ? __ldg(w + ((int64_t)c * cout + co) * taps + t) : __ldg(w + ((int64_t)co * cin + c) * taps + t);
    packed[i] = v;
  }
}

struct ConvParams {
  const float* x; const float* w; const float* bias; float* y; double* stats;
  int N, H, W, cin;          // input [N (= D for 3-D), H, W, cin]; cin % 4 == 0
  int No, Ho, Wo, cout;      // output; cout % 8 == 0
  int pd, ph, pw;            // zero padding per dimension
};

__device__ __forceinline__ void fma_vox(float (&acc)[kCo], const float4 in, const float4* __restrict__ wq) {
  // wq: 4 input channels x 8 output channels = 8 float4 in shared memory (same address for the whole warp: broadcast)
  const float iv[4] = {in.x, in.y, in.z, in.w};
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const float4 w0 = wq[2 * c], w1 = wq[2 * c + 1];
    acc[0] = fmaf(iv[c], w0.x, acc[0]); acc[1] = fmaf(iv[c], w0.y, acc[1]); acc[2] = fmaf(iv[c], w0.z, acc[2]); acc[3] = fmaf(iv[c], w0.w, acc[3]);
    acc[4] = fmaf(iv[c], w1.x, acc[4]); acc[5] = fmaf(iv[c], w1.y, acc[5]); acc[6] = fmaf(iv[c], w1.z, acc[6]); acc[7] = fmaf(iv[c], w1.w, acc[7]);
  }
}

// block-wide reduction of per-thread channel sums / sums of squares into the global double accumulators
__device__ __forceinline__ void stats_reduce(const float (&s1)[kCo], const float (&s2)[kCo], double* __restrict__ stats, int cout, int co0) {
  __shared__ float red[kConvThreads / 32][2 * kCo];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < kCo; ++j) {
    float a = s1[j], b = s2[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
    if (lane == 0) { red[warp][j] = a; red[warp][kCo + j] = b; }
  }
  __syncthreads();
  if (threadIdx.x < 2 * kCo) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < kConvThreads / 32; ++w) t += (double)red[w][threadIdx.x];
    const int j = threadIdx.x % kCo, which = threadIdx.x / kCo;
    atomicAdd(stats + which * cout + co0 + j, t);
  }
}

template <int KD, int KH, int KW, int S, int VPT>
__global__ void __launch_bounds__(kConvThreads) conv_cl_kernel(const ConvParams p) {
  constexpr int kVox = VPT;
  extern __shared__ __align__(16) float w_s[];      // [taps][cin][8] of this block's output-channel tile
  constexpr int taps = KD * KH * KW;
  const int co0 = blockIdx.y * kCo;
  {
    const float4* src = reinterpret_cast<const float4*>(p.w + (int64_t)blockIdx.y * taps * p.cin * kCo);
    float4* dst = reinterpret_cast<float4*>(w_s);
    for (int i = threadIdx.x; i < taps * p.cin * kCo / 4; i += kConvThreads) dst[i] = __ldg(src + i);
  }
  __syncthreads();
  const int wblocks = (p.Wo + kVox - 1) / kVox;
  const int64_t items = (int64_t)p.No * p.Ho * wblocks;
  const int64_t item = (int64_t)blockIdx.x * kConvThreads + threadIdx.x;
  float acc[kVox][kCo];
#pragma unroll
  for (int v = 0; v < kVox; ++v)
#pragma unroll
    for (int j = 0; j < kCo; ++j) acc[v][j] = 0.f;
  const bool active = item < items;
  int on = 0, oy = 0, ox0 = 0;
  if (active) {
    const int xb = (int)(item % wblocks);
    const int64_t r = item / wblocks;
    oy = (int)(r % p.Ho); on = (int)(r / p.Ho);
    ox0 = xb * kVox;
    const int cq = p.cin >> 2;
    for (int kd = 0; kd < KD; ++kd) {
      const int iz = on * (KD > 1 ? S : 1) - p.pd + kd;
      if (iz < 0 || iz >= p.N) continue;
      for (int kh = 0; kh < KH; ++kh) {
        const int iy = oy * S - p.ph + kh;
        if (iy < 0 || iy >= p.H) continue;
        const float4* row = reinterpret_cast<const float4*>(p.x + ((int64_t)iz * p.H + iy) * p.W * p.cin);
        const float4* wrow = reinterpret_cast<const float4*>(w_s) + (int64_t)((kd * KH + kh) * KW) * p.cin * 2;
        const int ixb = ox0 * S - p.pw;
        if (S == 1 && KW == 3) {
          // sliding window: 10 input positions feed 8 outputs x 3 taps
          for (int q = 0; q < cq; ++q) {
            float4 in[kVox + 2];
#pragma unroll
            for (int i = 0; i < kVox + 2; ++i) {
              const int ix = ixb + i;
              in[i] = (ix >= 0 && ix < p.W) ? __ldg(row + (int64_t)ix * cq + q) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
              const float4* wq = wrow + ((int64_t)kw * p.cin + 4 * q) * 2;
#pragma unroll
              for (int v = 0; v < kVox; ++v) fma_vox(acc[v], in[v + kw], wq);
            }
          }
        } else if (S == 2 && KW == 3) {
          // stride 2: 2 VPT + 1 input positions feed VPT outputs x 3 taps (output v reads positions 2 v + kw)
          for (int q = 0; q < cq; ++q) {
            float4 in[2 * kVox + 1];
#pragma unroll
            for (int i = 0; i < 2 * kVox + 1; ++i) {
              const int ix = ixb + i;
              in[i] = (ix >= 0 && ix < p.W) ? __ldg(row + (int64_t)ix * cq + q) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
              const float4* wq = wrow + ((int64_t)kw * p.cin + 4 * q) * 2;
#pragma unroll
              for (int v = 0; v < kVox; ++v) fma_vox(acc[v], in[2 * v + kw], wq);
            }
          }
        } else {
          for (int kw = 0; kw < KW; ++kw) {
            for (int q = 0; q < cq; ++q) {
              const float4* wq = wrow + ((int64_t)kw * p.cin + 4 * q) * 2;
#pragma unroll
              for (int v = 0; v < kVox; ++v) {
                const int ix = ixb + v * S + kw;
                const float4 in = (ix >= 0 && ix < p.W) ? __ldg(row + (int64_t)ix * cq + q) : make_float4(0.f, 0.f, 0.f, 0.f);
                fma_vox(acc[v], in, wq);
              }
            }
          }
        }
      }
    }
  }
  float s1[kCo], s2[kCo];
#pragma unroll
  for (int j = 0; j < kCo; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
  if (active) {
    float b[kCo];
#pragma unroll
    for (int j = 0; j < kCo; ++j) b[j] = p.bias ? __ldg(p.bias + co0 + j) : 0.f;
    float* out = p.y + (((int64_t)on * p.Ho + oy) * p.Wo + ox0) * p.cout + co0;
#pragma unroll
    for (int v = 0; v < kVox; ++v) {
      if (ox0 + v >= p.Wo) break;
#pragma unroll
      for (int j = 0; j < kCo; ++j) { acc[v][j] += b[j]; s1[j] += acc[v][j]; s2[j] = fmaf(acc[v][j], acc[v][j], s2[j]); }
      reinterpret_cast<float4*>(out + (int64_t)v * p.cout)[0] = make_float4(acc[v][0], acc[v][1], acc[v][2], acc[v][3]);
      reinterpret_cast<float4*>(out + (int64_t)v * p.cout)[1] = make_float4(acc[v][4], acc[v][5], acc[v][6], acc[v][7]);
    }
  }
  if (p.stats) stats_reduce(s1, s2, p.stats, p.cout, co0);
}

// ConvTranspose3d(k = 3, stride 2, padding 1, output_padding 1): out = 2 x in per dimension; output o gets input i through
// tap k when o = 2 i - 1 + k: an even o has one tap per dimension (k = 1), an odd o two (k = 0, 2).  One thread = VPT outputs of
// one parity class (pz, py, px) x 8 channels; the work items are ordered class-major, so a warp walks ONE tap set (1 to 8
// taps) instead of the masked union of all 27.
template <int VPT>
__global__ void __launch_bounds__(kConvThreads) convt3_cl_kernel(const ConvParams p) {
  extern __shared__ __align__(16) float w_s[];
  const int co0 = blockIdx.y * kCo;
  {
    const float4* src = reinterpret_cast<const float4*>(p.w + (int64_t)blockIdx.y * 27 * p.cin * kCo);
    float4* dst = reinterpret_cast<float4*>(w_s);
    for (int i = threadIdx.x; i < 27 * p.cin * kCo / 4; i += kConvThreads) dst[i] = __ldg(src + i);
  }
  __syncthreads();
  const int wblocks = (p.W + VPT - 1) / VPT;           // blocks of VPT same-parity outputs = VPT input columns
  const int64_t per_class = (int64_t)p.N * p.H * wblocks;
  const int64_t item = (int64_t)blockIdx.x * kConvThreads + threadIdx.x;
  float acc[VPT][kCo];
#pragma unroll
  for (int v = 0; v < VPT; ++v)
#pragma unroll
    for (int j = 0; j < kCo; ++j) acc[v][j] = 0.f;
  const bool active = item < 8 * per_class;
  int oz = 0, oy = 0, px = 0, xb = 0;
  if (active) {
    const int cls = (int)(item / per_class);
    int64_t r = item - (int64_t)cls * per_class;
    xb = (int)(r % wblocks); r /= wblocks;
    const int hy = (int)(r % p.H), hz = (int)(r / p.H);
    px = cls & 1;
    oy = 2 * hy + ((cls >> 1) & 1); oz = 2 * hz + ((cls >> 2) & 1);
    const int cq = p.cin >> 2;
    for (int kd = 0; kd < 3; ++kd) {
      if (((oz + 1 - kd) & 1) != 0) continue;
      const int iz = (oz + 1 - kd) >> 1;
      if (iz < 0 || iz >= p.N) continue;
      for (int kh = 0; kh < 3; ++kh) {
        if (((oy + 1 - kh) & 1) != 0) continue;
        const int iy = (oy + 1 - kh) >> 1;
        if (iy < 0 || iy >= p.H) continue;
        const float4* row = reinterpret_cast<const float4*>(p.x + ((int64_t)iz * p.H + iy) * p.W * p.cin);
        for (int kw = 0; kw < 3; ++kw) {
          if (((px + 1 - kw) & 1) != 0) continue;
          const int dx = (px + 1 - kw) >> 1;        // ix = (ox + 1 - kw) / 2 = VPT xb + j + dx
          const float4* wrow = reinterpret_cast<const float4*>(w_s) + (int64_t)((kd * 3 + kh) * 3 + kw) * p.cin * 2;
          for (int q = 0; q < cq; ++q) {
            const float4* wq = wrow + (int64_t)(4 * q) * 2;
#pragma unroll
            for (int v = 0; v < VPT; ++v) {
              const int ix = xb * VPT + v + dx;
              const float4 in = (ix >= 0 && ix < p.W) ? __ldg(row + (int64_t)ix * cq + q) : make_float4(0.f, 0.f, 0.f, 0.f);
              fma_vox(acc[v], in, wq);
            }
          }
        }
      }
    }
  }
  float s1[kCo], s2[kCo];
#pragma unroll
  for (int j = 0; j < kCo; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
  if (active) {
#pragma unroll
    for (int v = 0; v < VPT; ++v) {
      const int ox = 2 * (xb * VPT + v) + px;
      if (ox >= p.Wo) break;
#pragma unroll
      for (int j = 0; j < kCo; ++j) { s1[j] += acc[v][j]; s2[j] = fmaf(acc[v][j], acc[v][j], s2[j]); }
      float* out = p.y + (((int64_t)oz * p.Ho + oy) * p.Wo + ox) * p.cout + co0;
      reinterpret_cast<float4*>(out)[0] = make_float4(acc[v][0], acc[v][1], acc[v][2], acc[v][3]);
      reinterpret_cast<float4*>(out)[1] = make_float4(acc[v][4], acc[v][5], acc[v][6], acc[v][7]);
    }
  }
  if (p.stats) stats_reduce(s1, s2, p.stats, p.cout, co0);
}

// InPlaceABN forward: y = leaky_relu(gamma (x - mean) / sqrt(var + eps) + beta, slope) (+ skip), channels-last.
// training != 0: batch statistics from the conv kernel's double sums (biased variance, as batch norm normalises with);
// block 0 also updates the running statistics (momentum, unbiased variance) like the reference's train()-mode forward.
__global__ void __launch_bounds__(256) bn_act_cl_kernel(const float* __restrict__ x, int64_t n, int C, const double* __restrict__ stats,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        float* running_mean, float* running_var, float eps, float momentum, float slope,
                                                        int training, const float* __restrict__ skip, float* __restrict__ y) {
  __shared__ float sc[64], sh[64];
  if (threadIdx.x < C) {
    const int c = threadIdx.x;
    float mean, var;
    if (training) {
      const double m = stats[c] / (double)n;
      double v = stats[C + c] / (double)n - m * m;
      if (v < 0.0) v = 0.0;
      mean = (float)m; var = (float)v;
      if (blockIdx.x == 0 && running_mean && running_var) {
        const double unbiased = n > 1 ? v * (double)n / (double)(n - 1) : v;
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
      }
    } else {
      mean = running_mean[c]; var = running_var[c];
    }
    const float inv = 1.0f / sqrtf(var + eps);
    const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
    sc[c] = g * inv; sh[c] = b - mean * g * inv;
  }
  __syncthreads();
  const int cq = C >> 2;
  const int64_t total = n * cq;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % cq) * 4;
    float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    float o[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float t = fmaf(o[k], sc[c + k], sh[c + k]);
      o[k] = t > 0.f ? t : t * slope;
    }
    if (skip) {
      const float4 s = __ldg(reinterpret_cast<const float4*>(skip) + i);
      o[0] += s.x; o[1] += s.y; o[2] += s.z; o[3] += s.w;
    }
    reinterpret_cast<float4*>(y)[i] = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// F.interpolate(mode='bilinear', align_corners=False) of packed images [V, H, W, 4] -> [V, h, w, 4]
// (ATen upsample_bilinear2d: src = scale (dst + 0.5) - 0.5 clamped at 0, scale = in / out)
__global__ void resize_bilinear_cl_kernel(const float4* __restrict__ x, int V, int H, int W, int h, int w, float4* __restrict__ y) {
  const int64_t total = (int64_t)V * h * w;
  const float sy = (float)H / (float)h, sx = (float)W / (float)w;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ox = (int)(i % w), oy = (int)((i / w) % h), v = (int)(i / ((int64_t)w * h));
    float fy = sy * ((float)oy + 0.5f) - 0.5f, fx = sx * ((float)ox + 0.5f) - 0.5f;
    fy = fy < 0.f ? 0.f : fy; fx = fx < 0.f ? 0.f : fx;
    const int y0 = (int)fy, x0 = (int)fx;
    const int y1 = y0 + (y0 < H - 1 ? 1 : 0), x1 = x0 + (x0 < W - 1 ? 1 : 0);
    const float ly1 = fy - (float)y0, lx1 = fx - (float)x0, ly0 = 1.f - ly1, lx0 = 1.f - lx1;
    const float4* img = x + (int64_t)v * H * W;
    const float4 a = __ldg(img + (int64_t)y0 * W + x0), b = __ldg(img + (int64_t)y0 * W + x1);
    const float4 c = __ldg(img + (int64_t)y1 * W + x0), d = __ldg(img + (int64_t)y1 * W + x1);
    float4 o;
    o.x = ly0 * (lx0 * a.x + lx1 * b.x) + ly1 * (lx0 * c.x + lx1 * d.x);
    o.y = ly0 * (lx0 * a.y + lx1 * b.y) + ly1 * (lx0 * c.y + lx1 * d.y);
    o.z = ly0 * (lx0 * a.z + lx1 * b.z) + ly1 * (lx0 * c.z + lx1 * d.z);
    o.w = 0.f;
    y[i] = o;
  }
}

template <int KD, int KH, int KW, int S, int VPT>
int launch_conv_vpt(const ConvParams& p, cudaStream_t st) {
  const size_t smem = (size_t)KD * KH * KW * p.cin * kCo * sizeof(float);
  ZEST_CHECK_ARG(smem <= 200 * 1024, "zest_conv_cl_fwd: weight tile of %zu bytes does not fit shared memory", smem);
  ZEST_CUDA(cudaFuncSetAttribute(conv_cl_kernel<KD, KH, KW, S, VPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t items = (int64_t)p.No * p.Ho * ((p.Wo + VPT - 1) / VPT);
  dim3 grid((unsigned)((items + kConvThreads - 1) / kConvThreads), (unsigned)(p.cout / kCo));
  conv_cl_kernel<KD, KH, KW, S, VPT><<<grid, kConvThreads, smem, st>>>(p);
  ZEST_LAUNCH_CHECK();
  return ZEST_OK;
}

// 8 outputs per thread amortise the weight fetches best, but the coarse levels of the U-Net have too few outputs to fill
// 148 SMs that way (CostRegNet.conv5: 48 blocks): below ~2 blocks per SM the layer runs with 2 outputs per thread instead.
// outputs per thread: the largest of 8 / 4 / 2 that still gives ~2 blocks per SM (ZEST_CONV_VPT=8|4|2 forces one: A/B runs)
static int pick_vpt(int64_t items_w1, int Wo, int cout_tiles, int min_blocks_per_sm = 2) {
  static const int forced = getenv("ZEST_CONV_VPT") ? atoi(getenv("ZEST_CONV_VPT")) : 0;
  if (forced == 8 || forced == 4 || forced == 2) return forced;
  for (int vpt = 8; vpt > 2; vpt >>= 1) {
    const int64_t blocks = (items_w1 * ((Wo + vpt - 1) / vpt) + kConvThreads - 1) / kConvThreads * cout_tiles;
    if (blocks >= min_blocks_per_sm * (int64_t)num_sms()) return vpt;
  }
  return 2;
}

template <int KD, int KH, int KW, int S>
int launch_conv(const ConvParams& p, cudaStream_t st) {
  switch (pick_vpt((int64_t)p.No * p.Ho, p.Wo, p.cout / kCo)) {
    case 8: return launch_conv_vpt<KD, KH, KW, S, 8>(p, st);
    case 4: return launch_conv_vpt<KD, KH, KW, S, 4>(p, st);
    default: return launch_conv_vpt<KD, KH, KW, S, 2>(p, st);
  }
}

}  // namespace
}  // namespace zest

using namespace zest;

extern "C" int zest_conv_pack_weights(const float* w, int cout, int cin, int kd, int kh, int kw, int transposed, int cin_pad, float* packed,
                                      void* stream) {
  ZEST_CHECK_ARG(w && packed && cout > 0 && (cout % kCo) == 0 && cin > 0 && cin_pad >= cin && (cin_pad % 4) == 0 && kd > 0 && kh > 0 && kw > 0,
                 "zest_conv_pack_weights: bad arguments (cout %d must be a multiple of 8, cin_pad %d a multiple of 4)", cout, cin_pad);
  const int64_t total = (int64_t)cout * kd * kh * kw * cin_pad;
  conv_pack_weights_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(w, cout, cin, kd * kh * kw, transposed, cin_pad, packed);
  ZEST_LAUNCH_CHECK();
  return ZEST_OK;
}

extern "C" int zest_conv_cl_fwd(const float* x, int N, int H, int W, int cin, const float* wpacked, const float* bias, int cout, int kd, int kh,
                                int kw, int stride, float* y, double* stats, void* stream) {
  ZEST_CHECK_ARG(x && wpacked && y && N > 0 && H > 0 && W > 0 && cin > 0 && (cin % 4) == 0 && cout > 0 && (cout % kCo) == 0,
                 "zest_conv_cl_fwd: bad arguments (cin %d must be a multiple of 4, cout %d of 8)", cin, cout);
  ZEST_CHECK_ARG(stride == 1 || stride == 2, "zest_conv_cl_fwd: stride must be 1 or 2");
  ConvParams p{};
  p.x = x; p.w = wpacked; p.bias = bias; p.y = y; p.stats = stats;
  p.N = N; p.H = H; p.W = W; p.cin = cin; p.cout = cout;
  p.pd = kd / 2; p.ph = kh / 2; p.pw = kw / 2;                       // "same" padding: every conv of both nets
  p.No = kd > 1 ? (N + 2 * p.pd - kd) / stride + 1 : N;
  p.Ho = (H + 2 * p.ph - kh) / stride + 1;
  p.Wo = (W + 2 * p.pw - kw) / stride + 1;
  cudaStream_t st = (cudaStream_t)stream;
  if (stats) ZEST_CUDA(cudaMemsetAsync(stats, 0, 2 * (size_t)cout * sizeof(double), st));
  const int key = kd * 1000 + kh * 100 + kw * 10 + stride;
  switch (key) {
    case 3331: return launch_conv<3, 3, 3, 1>(p, st);
    case 3332: return launch_conv<3, 3, 3, 2>(p, st);
    case 1331: return launch_conv<1, 3, 3, 1>(p, st);
    case 1552: return launch_conv<1, 5, 5, 2>(p, st);
    case 1111: return launch_conv<1, 1, 1, 1>(p, st);
    default:
      set_error("zest_conv_cl_fwd: kernel %dx%dx%d stride %d is not instantiated", kd, kh, kw, stride);
      return ZEST_E_ARG;
  }
}

extern "C" int zest_convt3_cl_fwd(const float* x, int D, int H, int W, int cin, const float* wpacked, int cout, float* y, double* stats,
                                  void* stream) {
  ZEST_CHECK_ARG(x && wpacked && y && D > 0 && H > 0 && W > 0 && cin > 0 && (cin % 4) == 0 && cout > 0 && (cout % kCo) == 0,
                 "zest_convt3_cl_fwd: bad arguments");
  ConvParams p{};
  p.x = x; p.w = wpacked; p.y = y; p.stats = stats;
  p.N = D; p.H = H; p.W = W; p.cin = cin; p.cout = cout;
  p.No = 2 * D; p.Ho = 2 * H; p.Wo = 2 * W;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = (size_t)27 * cin * kCo * sizeof(float);
  ZEST_CHECK_ARG(smem <= 200 * 1024, "zest_convt3_cl_fwd: weight tile does not fit shared memory");
  if (stats) ZEST_CUDA(cudaMemsetAsync(stats, 0, 2 * (size_t)cout * sizeof(double), st));
#define ZEST_CONVT(VPT)                                                                                              \
  do {                                                                                                             \
    ZEST_CUDA(cudaFuncSetAttribute(convt3_cl_kernel<VPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    const int64_t items = (int64_t)8 * D * H * ((W + VPT - 1) / VPT);                                              \
    dim3 grid((unsigned)((items + kConvThreads - 1) / kConvThreads), (unsigned)(cout / kCo));                      \
    convt3_cl_kernel<VPT><<<grid, kConvThreads, smem, st>>>(p);                                                    \
  } while (0)
  // the transposed convolutions have 1 - 8 taps per output and little reuse to lose: they prefer more, smaller threads
  switch (pick_vpt((int64_t)8 * D * H, W, cout / kCo, 8)) {
    case 8: ZEST_CONVT(8); break;
    case 4: ZEST_CONVT(4); break;
    default: ZEST_CONVT(2); break;
  }
#undef ZEST_CONVT
  ZEST_LAUNCH_CHECK();
  return ZEST_OK;
}

extern "C" int zest_bn_act_cl(const float* x, int64_t n, int C, const double* stats, const float* gamma, const float* beta,
                              float* running_mean, float* running_var, float eps, float momentum, float slope, int training,
                              const float* skip, float* y, void* stream) {
  ZEST_CHECK_ARG(x && y && n > 0 && C > 0 && C <= 64 && (C % 4) == 0, "zest_bn_act_cl: bad arguments (C %d must be a multiple of 4, <= 64)", C);
  ZEST_CHECK_ARG(training ? stats != nullptr : (running_mean && running_var), "zest_bn_act_cl: training needs batch sums, eval needs running statistics");
  const int64_t total = n * (C / 4);
  const int64_t blocks = (total + 255) / 256;
  const unsigned grid = (unsigned)(blocks < (int64_t)num_sms() * 16 ? blocks : (int64_t)num_sms() * 16);
  bn_act_cl_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, n, C, stats, gamma, beta, running_mean, running_var, eps, momentum, slope,
                                                           training, skip, y);
  ZEST_LAUNCH_CHECK();
  return ZEST_OK;
}

extern "C" int zest_resize_bilinear_cl(const float* x, int V, int H, int W, int h, int w, float* y, void* stream) {
  ZEST_CHECK_ARG(x && y && V > 0 && H > 0 && W > 0 && h > 0 && w > 0, "zest_resize_bilinear_cl: bad arguments");
  const int64_t total = (int64_t)V * h * w;
  resize_bilinear_cl_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(x), V, H, W, h, w,
                                                                                          reinterpret_cast<float4*>(y));
  ZEST_LAUNCH_CHECK();
  return ZEST_OK;
}
