// fp32-grade radiance MLP, layer by layer: forward and backward of networks.py:150-221 (Renderer.forward, v0).
//
// This is the precision-reference path ("fp32 MLP" of the parity bar: RGB/depth <= 2e-3 max-abs)
// and the training path (fine_tune.py): it keeps every activation when train=1 so the backward
// can run layer by layer.  Every layer is one GEMM with a fused epilogue (launch_gemm, sgemm.cuh): on the
// tensor cores with split-precision operands by default (tc_gemm.cu), on CUDA cores in exact fp32 on request
// (sgemm.cu).  The inference hot path is the bf16 tcgen05 kernel in mlp_tc.cu.
//
//   g   = pts_bias(feat)                                   networks.py:174
//   h_i = relu(L_i(h_{i-1}) * g), i < depth; [pe | h] after layer `skip`   :176-182
//   sigma = alpha_linear(h); feat = feature_linear(h)      :195-198
//   v = relu(views_linears[0]([feat | dirpe])); rgb = rgb_linear(v)   :199-207
//   static+sf: b = sigmoid(w_linear(h)); dynamic: sf = tanh(sf_linear(h)), prob = sigmoid(prob_linear(h))  :184-191
#include <vector>

#include "net.cuh"
#include "sgemm.cuh"

namespace zest {

int in_layer(const zest_net* n, int layer) {
  if (layer == 0) return n->in_pts;
  if (layer == n->skip + 1) return n->width + n->in_pts;
  return n->width;
}

static inline int64_t up4(int64_t v) { return (v + 3) & ~int64_t(3); }

// Workspace plan: offsets in floats, every row-major [M, ld] block starts 16-byte aligned.
struct F32Plan {
  int W, P, F, Cv, D, skip, ns;
  int ldX5, ldVX, ldXF;
  int64_t G, X5, XF, H[16], Z[16], VX, V128, SH, RGB;       // forward
  int64_t gHa, gHb, dZ, gG, gX5, gVX, gV128, gSH, gRGB;      // backward scratch (train only); dZ / gHb ping-pong as dZ_i
  int64_t gWS;                                                // stacked small heads' dW [16, W] + db [16] (train only)
  int64_t PACK;                                               // weight-operand stage images of the tensor-core GEMM
  static constexpr int64_t kPackFloats = 512 * 1024;          // 2 MiB >= 2 column tiles x 22 K stages x 32 KiB
  int64_t total;
  F32Plan(const zest_net* n, int64_t M, bool train) {
    W = n->width; P = n->in_pts; F = n->in_feat; Cv = n->in_views; D = n->depth; skip = n->skip; ns = n->n_small;
    ldX5 = (int)up4(P + W); ldVX = (int)up4(W + Cv); ldXF = (int)up4(F);
    int64_t o = 0;
    auto take = [&](int64_t per_row) { int64_t r = o; o += up4(M * per_row); return r; };
    G = take(W); X5 = take(ldX5); XF = take(ldXF);
    if (train) {
      // the backward needs z only where the layer's output is not a 16-byte-aligned [M, W] block of its own: the skip layer
      // (its h lives behind pe in X5); everywhere else z is recovered from h (GemmArgs::gb_from_h) and never stored
      for (int i = 0; i < D; ++i) { H[i] = (i == skip) ? -1 : take(W); Z[i] = (i == skip) ? take(W) : -1; }
    } else {
      const int64_t a = take(W), b = take(W);
      for (int i = 0; i < D; ++i) { H[i] = (i == skip) ? -1 : ((i & 1) ? b : a); Z[i] = -1; }
    }
    VX = take(ldVX); V128 = take(W / 2); SH = take(16); RGB = take(4);
    if (train) {
      gHa = take(W); gHb = take(W); dZ = take(W); gG = take(W); gX5 = take(ldX5); gVX = take(ldVX);
      gV128 = take(W / 2); gSH = take(16); gRGB = take(4);
      gWS = o; o += up4(16 * (int64_t)W + 16);
    }
    PACK = o; o += kPackFloats;
    total = o;
  }
};

// scratch of the GEMM in flight (stream-ordered reuse: one pack kernel + one GEMM at a time per stream)
static thread_local void* t_pack = nullptr;
struct PackScope {
  PackScope(float* ws, const F32Plan& p) { t_pack = ws + p.PACK; }
  ~PackScope() { t_pack = nullptr; }
};

// rows [r0, r0 + n) of the stacked heads' dW / db -> += into one head's gradient tensors (zero-initialised by the caller)
__global__ void heads_scatter_kernel(const float* __restrict__ gws, int W, int r0, int n, float* __restrict__ gW,
                                     float* __restrict__ gb) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (gW && i < n * W) gW[i] += gws[r0 * W + i];
  if (gb && i < n) gb[i] += gws[16 * W + r0 + i];
}

__global__ void finalize_fwd_kernel(const float* __restrict__ rgb, const float* __restrict__ sh, int kind,
                                    int64_t M, float* __restrict__ raw, int out_ch) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  float* o = raw + m * out_ch;
  o[0] = rgb[m * 4]; o[1] = rgb[m * 4 + 1]; o[2] = rgb[m * 4 + 2];
  const float* s = sh + m * 16;
  o[3] = s[0];
  if (kind == 1) {
    o[4] = 1.f / (1.f + expf(-s[1]));
  } else if (kind == 2) {
#pragma unroll
    for (int k = 0; k < 6; ++k) o[4 + k] = tanhf(s[1 + k]);
    o[10] = 1.f / (1.f + expf(-s[7]));
    o[11] = 1.f / (1.f + expf(-s[8]));
  }
}

__global__ void finalize_bwd_kernel(const float* __restrict__ graw, const float* __restrict__ sh, int kind,
                                    int64_t M, int out_ch, float* __restrict__ grgb, float* __restrict__ gsh) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const float* g = graw + m * out_ch;
  const float* s = sh + m * 16;
  grgb[m * 4] = g[0]; grgb[m * 4 + 1] = g[1]; grgb[m * 4 + 2] = g[2]; grgb[m * 4 + 3] = 0.f;
  float* o = gsh + m * 16;
#pragma unroll
  for (int k = 0; k < 16; ++k) o[k] = 0.f;
  o[0] = g[3];
  if (kind == 1) {
    const float y = 1.f / (1.f + expf(-s[1]));
    o[1] = g[4] * y * (1.f - y);
  } else if (kind == 2) {
#pragma unroll
    for (int k = 0; k < 6; ++k) { const float y = tanhf(s[1 + k]); o[1 + k] = g[4 + k] * (1.f - y * y); }
#pragma unroll
    for (int k = 0; k < 2; ++k) { const float y = 1.f / (1.f + expf(-s[7 + k])); o[7 + k] = g[10 + k] * y * (1.f - y); }
  }
}

__global__ void relu_mask_kernel(float* __restrict__ g, const float* __restrict__ y, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && !(y[i] > 0.f)) g[i] = 0.f;
}

static int copy2d(float* dst, int64_t ldd, const float* src, int64_t lds, int cols, int64_t M, cudaStream_t st) {
  if (cols <= 0 || M <= 0) return ZEST_OK;
  ZEST_CUDA(cudaMemcpy2DAsync(dst, ldd * sizeof(float), src, lds * sizeof(float), (size_t)cols * sizeof(float),
                              (size_t)M, cudaMemcpyDeviceToDevice, st));
  return ZEST_OK;
}


// y = x @ W^T + b  with W [J, K] row-major (nn.Linear)
static GemmArgs linear(const float* x, int64_t ldx, const float* W, int J, int64_t K, const float* b, float* y,
                       int64_t ldy, int64_t M) {
  GemmArgs a{};
  a.A = x; a.sa_i = ldx; a.sa_k = 1;
  a.B = W; a.sb_j = K; a.sb_k = 1;
  a.C = y; a.ldc = ldy; a.I = M; a.J = J; a.K = K; a.bias = b;
  a.b_scratch = t_pack; a.b_scratch_bytes = t_pack ? F32Plan::kPackFloats * (int64_t)sizeof(float) : 0;
  return a;
}
// gx = gy @ W  with W [N, J] row-major: reduce over N
static GemmArgs linear_bwd_x(const float* gy, int64_t ldgy, const float* W, int N, int J, float* gx, int64_t ldgx,
                             int64_t M, int accumulate) {
  GemmArgs a{};
  a.A = gy; a.sa_i = ldgy; a.sa_k = 1;
  a.B = W; a.sb_j = 1; a.sb_k = J;
  a.C = gx; a.ldc = ldgx; a.I = M; a.J = J; a.K = N; a.accumulate = accumulate;
  a.b_scratch = t_pack; a.b_scratch_bytes = t_pack ? F32Plan::kPackFloats * (int64_t)sizeof(float) : 0;
  return a;
}
// gW [N, J] += gy^T @ x : reduce over the M rows, split across grid.z;  gb [N] += column sums of gy (bias gradient)
static GemmArgs linear_bwd_w(const float* gy, int64_t ldgy, int N, const float* x, int64_t ldx, int J, float* gW,
                             int64_t M, float* gb) {
  GemmArgs a{};
  a.A = gy; a.sa_i = 1; a.sa_k = ldgy;
  a.B = x; a.sb_j = 1; a.sb_k = ldx;
  a.C = gW; a.ldc = J; a.I = N; a.J = J; a.K = M; a.accumulate = 1; a.rowsum = gb;
  int64_t s = M / 2048;
  a.splits = (int)(s < 2 ? 2 : (s > 256 ? 256 : s));
  return a;
}

}  // namespace zest

using namespace zest;

extern "C" zest_net* zest_net_create(int kind, int in_pts, int in_feat, int in_views, int width, int depth, int skip) {
  if (kind < 0 || kind > 2 || in_pts <= 0 || in_feat <= 0 || in_views <= 0 || width <= 0 || (width & 1) ||
      depth < 2 || depth > 16 || skip < 0 || skip >= depth - 1) {
    set_error("zest_net_create: unsupported configuration (kind=%d in_pts=%d in_feat=%d in_views=%d width=%d depth=%d skip=%d)",
              kind, in_pts, in_feat, in_views, width, depth, skip);
    return nullptr;
  }
  zest_net* n = new zest_net();
  n->kind = kind; n->in_pts = in_pts; n->in_feat = in_feat; n->in_views = in_views;
  n->width = width; n->depth = depth; n->skip = skip;
  n->out_ch = kind == 0 ? 4 : (kind == 1 ? 5 : 12);
  n->n_small = kind == 0 ? 1 : (kind == 1 ? 2 : 9);
  n->packed = false; n->f32 = nullptr; n->tc_blob = nullptr; n->tc_bias = nullptr; n->tc_plan_host = nullptr; n->tc_bytes = 0; n->tc_desc_dev = nullptr; n->tc_dirty = true; n->tc_counters = nullptr; n->tc_counter_next = 0;
  // blob layout; the stacked small heads keep alpha first so column 0 of SH is sigma
  int64_t o = 0;
  int pi = 0;
  auto put = [&](int64_t numel) { int64_t r = o; n->param_off[pi] = o; n->param_numel[pi] = numel; ++pi; o += numel; return r; };
  const int W = width;
  for (int i = 0; i < depth; ++i) { n->w_pts[i] = put((int64_t)W * in_layer(n, i)); n->b_pts[i] = put(W); }
  n->w_gate = put((int64_t)W * in_feat); n->b_gate = put(W);
  n->w_feat = put((int64_t)W * W); n->b_feat = put(W);
  // alpha w/b are params (2*depth+4, +5); the extra heads come last in parameter order but are
  // laid out so that [alpha.w ; extra.w] and [alpha.b ; extra.b] are contiguous matrices.
  const int64_t small_w = o; o += (int64_t)n->n_small * W;
  const int64_t small_b = o; o += n->n_small;
  n->w_small = small_w; n->b_small = small_b;
  n->param_off[pi] = small_w; n->param_numel[pi] = W; ++pi;      // alpha_linear.weight
  n->param_off[pi] = small_b; n->param_numel[pi] = 1; ++pi;      // alpha_linear.bias
  n->w_views = put((int64_t)(W / 2) * (W + in_views)); n->b_views = put(W / 2);
  n->w_rgb = put((int64_t)3 * (W / 2)); n->b_rgb = put(3);
  if (kind == 1) {
    n->param_off[pi] = small_w + W; n->param_numel[pi] = W; ++pi;       // w_linear.weight
    n->param_off[pi] = small_b + 1; n->param_numel[pi] = 1; ++pi;       // w_linear.bias
  } else if (kind == 2) {
    n->param_off[pi] = small_w + W; n->param_numel[pi] = 6 * (int64_t)W; ++pi;       // sf_linear.weight
    n->param_off[pi] = small_b + 1; n->param_numel[pi] = 6; ++pi;
    n->param_off[pi] = small_w + 7 * (int64_t)W; n->param_numel[pi] = 2 * (int64_t)W; ++pi;  // prob_linear.weight
    n->param_off[pi] = small_b + 7; n->param_numel[pi] = 2; ++pi;
  }
  n->n_params = pi;
  n->f32_floats = o;
  if (cudaMalloc(&n->f32, (size_t)o * sizeof(float)) != cudaSuccess) {
    set_error("zest_net_create: cudaMalloc of %lld bytes failed", (long long)(o * 4));
    delete n;
    return nullptr;
  }
  return n;
}

extern "C" void zest_net_destroy(zest_net* net) {
  if (!net) return;
  tc_free(net);
  if (net->f32) cudaFree(net->f32);
  delete net;
}

extern "C" int zest_net_out_channels(const zest_net* net) { return net ? net->out_ch : ZEST_E_ARG; }
extern "C" int zest_net_num_params(const zest_net* net) { return net ? net->n_params : ZEST_E_ARG; }

extern "C" int zest_net_pack(zest_net* net, const float* const* params, int n_params, void* stream) {
  ZEST_CHECK_ARG(net && params, "zest_net_pack: null argument");
  ZEST_CHECK_ARG(n_params == net->n_params, "zest_net_pack: expected %d parameter tensors, got %d", net->n_params, n_params);
  cudaStream_t st = (cudaStream_t)stream;
  for (int i = 0; i < n_params; ++i) {
    ZEST_CHECK_ARG(params[i], "zest_net_pack: parameter %d is null", i);
    ZEST_CUDA(cudaMemcpyAsync(net->f32 + net->param_off[i], params[i], (size_t)net->param_numel[i] * sizeof(float),
                              cudaMemcpyDeviceToDevice, st));
  }
  net->tc_dirty = true;   // the bf16 tensor-core image is rebuilt by the next inference launch (mlp_tc.cu), not here
  net->packed = true;
  return ZEST_OK;
}

extern "C" int64_t zest_mlp_f32_workspace(const zest_net* net, int64_t M, int train) {
  if (!net || M < 0) return ZEST_E_ARG;
  return F32Plan(net, M, train != 0).total * (int64_t)sizeof(float);
}

extern "C" int zest_mlp_fwd_f32(const zest_net* net, const float* x, int ldx, int64_t M, float* raw, void* workspace,
                                int train, void* stream) {
  ZEST_CHECK_ARG(net && x && raw && workspace && M >= 0, "zest_mlp_fwd_f32: null argument");
  if (!net->packed) { set_error("zest_mlp_fwd_f32: net not packed"); return ZEST_E_STATE; }
  ZEST_CHECK_ARG(ldx >= net->in_pts + net->in_feat + net->in_views, "zest_mlp_fwd_f32: ldx too small");
  if (M == 0) return ZEST_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const F32Plan p(net, M, train != 0);
  float* ws = (float*)workspace;
  const PackScope pack_scope(ws, p);
  const float* w = net->f32;
  const int W = p.W, P = p.P;
  float* G = ws + p.G;
  float* X5 = ws + p.X5;
  float* XF = ws + p.XF;
  // The caller's rows are [pe | feat | views] back to back (110 floats: neither the feat block nor the rows are 16-byte
  // aligned), which keeps the tensor-core GEMMs off their tensor-copy operand path.  pe is copied into the skip buffer
  // anyway; feat gets an aligned copy of its own (M x 20 floats), and layer 0 / the gate read those.
  ZEST_TRY(copy2d(X5, p.ldX5, x, ldx, P, M, st));
  ZEST_TRY(copy2d(XF, p.ldXF, x + P, ldx, p.F, M, st));
  { GemmArgs a = linear(XF, p.ldXF, w + net->w_gate, W, p.F, w + net->b_gate, G, W, M); ZEST_TRY(launch_gemm(a, st)); }
  const float* in = X5; int64_t ld_in = p.ldX5;
  for (int i = 0; i < p.D; ++i) {
    float* out = (i == p.skip) ? X5 + P : ws + p.H[i];
    const int64_t ld_out = (i == p.skip) ? p.ldX5 : W;
    GemmArgs a = linear(in, ld_in, w + net->w_pts[i], W, in_layer(net, i), w + net->b_pts[i], out, ld_out, M);
    a.gate = G; a.ldg = W; a.relu = 1;
    if (train && p.Z[i] >= 0) { a.Z = ws + p.Z[i]; a.ldz = W; }
    ZEST_TRY(launch_gemm(a, st));
    if (i == p.skip) { in = X5; ld_in = p.ldX5; } else { in = out; ld_in = ld_out; }
  }
  float* VX = ws + p.VX; float* V128 = ws + p.V128; float* SH = ws + p.SH; float* RGB = ws + p.RGB;
  { GemmArgs a = linear(in, ld_in, w + net->w_feat, W, W, w + net->b_feat, VX, p.ldVX, M); ZEST_TRY(launch_gemm(a, st)); }
  ZEST_TRY(copy2d(VX + W, p.ldVX, x + P + p.F, ldx, p.Cv, M, st));
  { GemmArgs a = linear(in, ld_in, w + net->w_small, p.ns, W, w + net->b_small, SH, 16, M); ZEST_TRY(launch_gemm(a, st)); }
  { GemmArgs a = linear(VX, p.ldVX, w + net->w_views, W / 2, W + p.Cv, w + net->b_views, V128, W / 2, M); a.relu = 1; ZEST_TRY(launch_gemm(a, st)); }
  { GemmArgs a = linear(V128, W / 2, w + net->w_rgb, 3, W / 2, w + net->b_rgb, RGB, 4, M); ZEST_TRY(launch_gemm(a, st)); }
  finalize_fwd_kernel<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(RGB, SH, net->kind, M, raw, net->out_ch);
  ZEST_LAUNCH_CHECK();
  return ZEST_OK;
}

extern "C" int zest_mlp_bwd_f32(const zest_net* net, const float* x, int ldx, int64_t M, const float* graw,
                                void* workspace, float* gx, float* const* gparams, int n_params, void* stream) {
  ZEST_CHECK_ARG(net && x && graw && workspace && M >= 0, "zest_mlp_bwd_f32: null argument");
  ZEST_CHECK_ARG(!gparams || n_params == net->n_params, "zest_mlp_bwd_f32: expected %d gradient tensors", net->n_params);
  if (!net->packed) { set_error("zest_mlp_bwd_f32: net not packed"); return ZEST_E_STATE; }
  if (M == 0) return ZEST_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const F32Plan p(net, M, true);
  float* ws = (float*)workspace;
  const PackScope pack_scope(ws, p);
  const float* w = net->f32;
  const int W = p.W, P = p.P, D = p.D;
  auto gp = [&](int idx) -> float* { return gparams ? gparams[idx] : nullptr; };
  const int I_GATE = 2 * D, I_FEAT = 2 * D + 2, I_ALPHA = 2 * D + 4, I_VIEWS = 2 * D + 6, I_RGB = 2 * D + 8, I_EXTRA = 2 * D + 10;
  float* G = ws + p.G; float* X5 = ws + p.X5; float* VX = ws + p.VX; float* V128 = ws + p.V128; float* SH = ws + p.SH;
  float* gHa = ws + p.gHa; float* gHb = ws + p.gHb; float* dZ = ws + p.dZ; float* gG = ws + p.gG; float* gX5 = ws + p.gX5;
  float* gVX = ws + p.gVX; float* gV128 = ws + p.gV128; float* gSH = ws + p.gSH; float* gRGB = ws + p.gRGB;
  const unsigned gm = (unsigned)((M + 255) / 256);

  finalize_bwd_kernel<<<gm, 256, 0, st>>>(graw, SH, net->kind, M, net->out_ch, gRGB, gSH);
  ZEST_LAUNCH_CHECK();
  // rgb_linear
  if (gp(I_RGB)) ZEST_TRY(launch_gemm(linear_bwd_w(gRGB, 4, 3, V128, W / 2, W / 2, gp(I_RGB), M, gp(I_RGB + 1)), st));
  ZEST_TRY(launch_gemm(linear_bwd_x(gRGB, 4, w + net->w_rgb, 3, W / 2, gV128, W / 2, M, 0), st));
  relu_mask_kernel<<<(unsigned)((M * (W / 2) + 255) / 256), 256, 0, st>>>(gV128, V128, M * (W / 2));
  ZEST_LAUNCH_CHECK();
  // views_linears[0]
  if (gp(I_VIEWS)) ZEST_TRY(launch_gemm(linear_bwd_w(gV128, W / 2, W / 2, VX, p.ldVX, W + p.Cv, gp(I_VIEWS), M, gp(I_VIEWS + 1)), st));
  ZEST_TRY(launch_gemm(linear_bwd_x(gV128, W / 2, w + net->w_views, W / 2, W + p.Cv, gVX, p.ldVX, M, 0), st));
  // last hidden activation
  const int last = D - 1;
  const float* Hl = ws + p.H[last];   // zest_net_create guarantees skip < depth - 1
  const int64_t ldHl = W;
  // feature_linear and the stacked small heads feed gH of the last layer
  if (gp(I_FEAT)) ZEST_TRY(launch_gemm(linear_bwd_w(gVX, p.ldVX, W, Hl, ldHl, W, gp(I_FEAT), M, gp(I_FEAT + 1)), st));
  ZEST_TRY(launch_gemm(linear_bwd_x(gVX, p.ldVX, w + net->w_feat, W, W, gHa, W, M, 0), st));
  // The small heads (alpha; + blending; + scene flow, disocclusion) are one stacked [ns, W] weight in the forward and one
  // dW GEMM here (each used to stream the last hidden activation on its own): [ns, W] + [ns] into scratch, then scattered.
  if (gp(I_ALPHA) || gp(I_EXTRA) || (net->kind == 2 && gp(I_EXTRA + 2))) {
    float* gWS = ws + p.gWS;
    ZEST_CUDA(cudaMemsetAsync(gWS, 0, (size_t)(16 * W + 16) * sizeof(float), st));
    ZEST_TRY(launch_gemm(linear_bwd_w(gSH, 16, p.ns, Hl, ldHl, W, gWS, M, gWS + 16 * W), st));
    auto scatter = [&](int r0, int n, int idx) {
      if (!gp(idx) && !gp(idx + 1)) return;
      heads_scatter_kernel<<<(unsigned)((n * W + 255) / 256), 256, 0, st>>>(gWS, W, r0, n, gp(idx), gp(idx + 1));
    };
    scatter(0, 1, I_ALPHA);
    if (net->kind == 1) scatter(1, 1, I_EXTRA);
    if (net->kind == 2) { scatter(1, 6, I_EXTRA); scatter(7, 2, I_EXTRA + 2); }
    ZEST_LAUNCH_CHECK();
  }
  // Every GEMM that produces the gradient wrt a hidden activation h_i = relu(z_i * g) applies that layer's gate backward in
  // its epilogue (GemmArgs::gb_*): it writes dZ_i = gH_i 1[z_i g > 0] g and accumulates gG += gH_i 1[..] z_i; gH_i itself
  // never reaches memory.  dZ ping-pongs between two buffers (a GEMM cannot overwrite its own A operand).
  ZEST_CUDA(cudaMemsetAsync(gG, 0, (size_t)M * W * sizeof(float), st));
  float* dZ_cur = dZ;
  float* dZ_next = gHb;
  auto fuse_gate = [&](GemmArgs& a, int layer, int col0, float* out) {
    a.gb_from_h = p.Z[layer] < 0;
    a.gb_Z = ws + (a.gb_from_h ? p.H[layer] : p.Z[layer]); a.gb_G = G; a.gb_gG = gG; a.gb_dZ = out; a.gb_ld = W; a.gb_col0 = col0;
  };
  {
    GemmArgs a = linear_bwd_x(gSH, 16, w + net->w_small, p.ns, W, gHa, W, M, 1);   // + feature_linear's part already in gHa
    fuse_gate(a, last, 0, dZ_cur);
    ZEST_TRY(launch_gemm(a, st));
  }
  for (int i = D - 1; i >= 0; --i) {
    // layer input
    const float* in; int64_t ld_in; const int K = in_layer(net, i);
    if (i == 0) { in = X5; ld_in = p.ldX5; }          // the aligned copy of pe (forward)
    else if (i == p.skip + 1) { in = X5; ld_in = p.ldX5; }
    else { in = (i - 1 == p.skip) ? X5 + P : ws + p.H[i - 1]; ld_in = (i - 1 == p.skip) ? p.ldX5 : W; }
    if (gp(2 * i)) ZEST_TRY(launch_gemm(linear_bwd_w(dZ_cur, W, W, in, ld_in, K, gp(2 * i), M, gp(2 * i + 1)), st));
    if (i == 0) {
      if (gx) ZEST_TRY(launch_gemm(linear_bwd_x(dZ_cur, W, w + net->w_pts[0], W, P, gx, ldx, M, 1), st));   // + the skip's part
    } else if (i == p.skip + 1) {
      // input = [pe | h_skip]: columns < P are d/d pe (kept in gX5, added to gx below), the rest is gH of layer `skip`
      GemmArgs a = linear_bwd_x(dZ_cur, W, w + net->w_pts[i], W, K, gX5, p.ldX5, M, 0);
      fuse_gate(a, i - 1, P, dZ_next);
      ZEST_TRY(launch_gemm(a, st));
      if (gx) ZEST_TRY(copy2d(gx, ldx, gX5, p.ldX5, P, M, st));
      float* t = dZ_cur; dZ_cur = dZ_next; dZ_next = t;
    } else {
      GemmArgs a = linear_bwd_x(dZ_cur, W, w + net->w_pts[i], W, K, gHa, W, M, 0);   // C unused: every column is fused
      fuse_gate(a, i - 1, 0, dZ_next);
      ZEST_TRY(launch_gemm(a, st));
      float* t = dZ_cur; dZ_cur = dZ_next; dZ_next = t;
    }
  }
  // gate (pts_bias)
  if (gp(I_GATE)) ZEST_TRY(launch_gemm(linear_bwd_w(gG, W, W, ws + p.XF, p.ldXF, p.F, gp(I_GATE), M, gp(I_GATE + 1)), st));
  if (gx) {
    ZEST_TRY(launch_gemm(linear_bwd_x(gG, W, w + net->w_gate, W, p.F, gx + P, ldx, M, 0), st));
    ZEST_TRY(copy2d(gx + P + p.F, ldx, gVX + W, p.ldVX, p.Cv, M, st));
  }
  return ZEST_OK;
}
