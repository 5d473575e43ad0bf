/*
 * zest_b200.h -- C ABI of the B200-native ZeST-NeRF per-ray rendering path.
 *
 * The reference (violetamenendez/zest-nerf) is pure Python on stock PyTorch ops; it has no
 * FFI of its own.  This header is the boundary a maintainer binds with ctypes (see
 * INTEGRATION.md) to replace, stage by stage, the PyTorch calls on the hot path:
 *
 *   zest_pack_volume / zest_pack_images   layout repack done once per frame (no reference
 *                                         counterpart: grid_sample reads NCDHW / NCHW directly)
 *   zest_gather_fwd / _bwd                utils.py:433-459 index_point_feature (grid_sample 3-D)
 *                                         + utils.py:461-505 build_color_volume (grid_sample 2-D,
 *                                         projection utils.py:257-269) + renderer.py:51-72
 *   zest_dirfeat_fwd                      renderer.py:604 (cos_angle) + :34-49,256-258
 *   zest_encode_fwd / _bwd                networks.py:48-65 Embedding + renderer.py:246-297 concat
 *   zest_net_create / _destroy / _pack    networks.py:73-132 Renderer parameters (packed copies)
 *   zest_mlp_fwd_f32 / zest_mlp_bwd_f32   networks.py:150-221 Renderer.forward, fp32 CUDA cores
 *   zest_mlp_fwd_tc                       same, bf16 tcgen05/TMEM tensor cores, PE fused in prologue
 *   zest_composite_static_fwd / _bwd      renderer.py:74-164 depth2dist + raw2alpha + raw2outputs
 *   zest_composite_blend_fwd / _bwd       renderer.py:166-219 raw2outputs_blending
 *   zest_gather_mlp_fwd_tc                renderer.py:51-72 + 246-318 + 237-242 -> networks.py:150-221 in ONE
 *                                         launch (gather + PE + tensor-core MLP): the inference hot path
 *   zest_build_rays                       utils.py:133-230 get_rays_mvs + :290-394 build_rays_base +
 *                                         :232-288 get_ndc_coordinate ("next" row f1)
 *   zest_sf_smooth_loss_fwd / _bwd        losses.py:142-161 compute_sf_smooth_loss            ("next" row f4)
 *   zest_sf_lke_loss_fwd / _bwd           losses.py:164-203 compute_sf_lke_loss
 *   zest_project_ndc_fwd / _bwd           utils.py:516-539 projection_from_ndc (+ :507-514 NDC2Euclidean)
 *   zest_cost_volume_fwd / _bwd           networks.py:1077-1140 MVSNet.build_volume_cost + utils.py:49-99 homo_warp
 *                                         ("next" row f3, first half)
 *   zest_conv_cl_fwd / zest_convt3_cl_fwd / zest_bn_act_cl / zest_resize_bilinear_cl / zest_conv_pack_weights
 *                                         networks.py:935-1059 FeatureNet + CostRegNet with InPlaceABN ("next" row f3,
 *                                         second half), :1142-1238 MVSNet.forward
 *
 * Conventions
 *   - Every pointer is a DEVICE pointer into memory owned by the caller (PyTorch); the library
 *     never frees or retains caller memory past a call.  The only library-owned state is the
 *     opaque zest_net handle (packed weights).
 *   - All entry points are asynchronous on the cudaStream_t passed in (as void*), perform no
 *     hidden synchronisation and keep no global mutable state except the last-error string.
 *   - Return value: 0 = ok, negative = error (ZEST_E_*); zest_last_error() gives the message of
 *     the last failure on the calling thread.  The Python wrapper raises RuntimeError.
 *   - M = number of samples (rays * samples-per-ray), row-major [M, C] tensors unless stated.
 *   - There is no CPU fallback anywhere behind this ABI.
 */
#ifndef ZEST_B200_H
#define ZEST_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ZEST_OK 0
#define ZEST_E_ARG (-1)      /* bad argument / unsupported shape */
#define ZEST_E_CUDA (-2)     /* a CUDA runtime call or launch failed */
#define ZEST_E_STATE (-3)    /* handle not packed / wrong net kind */

typedef struct zest_net zest_net; /* opaque: packed parameters of one radiance MLP */

const char* zest_last_error(void);
int zest_version(void);

/* ---- layout repack (once per frame) ------------------------------------------------------ */
/* NCDHW fp32 [C=8, D, H, W] -> channels-last [D, H, W, 8] fp32 (one 32-byte sector per voxel). */
int zest_pack_volume(const float* vol_ncdhw, float* vol_cl, int D, int H, int W, void* stream);
/* [V, 3, H, W] fp32 -> [V, H, W, 4] fp32 (r,g,b,0): one 128-bit load per bilinear corner. */
int zest_pack_images(const float* img_nchw, float* img_cl, int V, int H, int W, void* stream);
/* gradient of zest_pack_volume: channels-last grad -> NCDHW grad (accumulates: dst += src). */
int zest_unpack_volume_grad(const float* gvol_cl, float* gvol_ncdhw, int D, int H, int W, void* stream);

/* ---- fused feature gather ---------------------------------------------------------------- */
/* cams: [V, 24] fp32 per view: w2c rows 0..2 (12 floats, row-major 3x4), K (9 floats), 3 pad.
 * M = R * S samples (R rays, S samples per ray; the kernel maps lanes to rays, warps to samples).
 * feats [M, ldf]: cols 0..7 trilinear(vol, ndc), cols 8+4v.. = (r,g,b,mask) of view v.
 * ndc has row stride ndc_ld floats (3, or 4 when a time channel is interleaved).
 * vol_cl may be NULL (colour only) and img_cl may be NULL (volume only).
 * vox_idx (optional, [M,3] int32: floor(ix,iy,iz)) and pix_idx (optional, [M,V,2] int32:
 * floor of the border-clipped (ix,iy)) expose the integer indices for bit-exact parity tests. */
int zest_gather_fwd(const float* rays_pts, const float* rays_ndc, int ndc_ld, int64_t R, int S,
                    const float* vol_cl, int D, int Hv, int Wv,
                    const float* img_cl, int V, int H, int W, const float* cams,
                    float* feats, int ldf, int32_t* vox_idx, int32_t* pix_idx, void* stream);
/* backward of the trilinear part: gfeats[M, ldf] cols 0..7 -> gvol_cl (atomic scatter-add, may be
 * NULL) and gndc [M, gndc_ld] (d/d ndc through the interpolation weights, may be NULL; written,
 * not accumulated).  Images and world points are data: no gradient. */
int zest_gather_bwd(const float* rays_ndc, int ndc_ld, int64_t M, const float* vol_cl,
                    int D, int Hv, int Wv, const float* gfeats, int ldf,
                    float* gvol_cl, float* gndc, int gndc_ld, void* stream);

/* ---- direction feature ------------------------------------------------------------------- */
/* per ray: c = |d|, dirs = (d / c) @ R^T with R = w2c_ref[:3,:3] (row-major 3x3 passed as the
 * first 12-float block of a cams row).  cos_angle [R], dirs [R,3]. */
int zest_dirfeat_fwd(const float* rays_dir, int64_t R, const float* cam_ref, float* cos_angle,
                     float* dirs, void* stream);

/* ---- positional encoding + input assembly (fp32 path / module boundary) ------------------ */
/* x [M, ldx] = [ PE_{nf_pts}(pts[M, c_pts (+ time t appended if has_t)]) | feats[M, F] |
 *               PE_{nf_dir}(dirs[ray of row]) ], ray of row = row / S.
 * has_t: 0 = 3 channels; 1 = 4 channels, the 4th is the constant t (renderer.py:300-318);
 *        2 = 4 channels read from memory (stand-alone Embedding(4, N).forward, networks.py:48-65). */
int zest_encode_fwd(const float* ndc, int ndc_ld, int has_t, float t, int nf_pts,
                    const float* feats, int ldf, int F, const float* dirs, int nf_dir, int S,
                    int64_t M, float* x, int ldx, void* stream);
/* gndc[M, gndc_ld] (+)= d loss / d ndc through PE(pts) given gx[M, ldx] (first 3 channels; all 4 when has_t == 2). */
int zest_encode_bwd(const float* ndc, int ndc_ld, int has_t, float t, int nf_pts, const float* gx,
                    int ldx, int64_t M, float* gndc, int gndc_ld, int accumulate, void* stream);

/* ---- radiance MLP ------------------------------------------------------------------------ */
/* kind: 0 = plain (4 outputs), 1 = static + scene-flow (5: rgb,sigma,blend_w),
 *       2 = dynamic (12: rgb,sigma,sf_prev3,sf_post3,prob_prev,prob_post). */
zest_net* zest_net_create(int kind, int in_pts, int in_feat, int in_views, int width, int depth,
                          int skip);
void zest_net_destroy(zest_net* net);
int zest_net_out_channels(const zest_net* net);
int zest_net_num_params(const zest_net* net);
/* params: device pointers to fp32 tensors in nn.Linear layout ([out,in] weight, [out] bias), in
 * this order: pts_linears[0..depth-1] (w,b each), pts_bias (w,b), feature_linear, alpha_linear,
 * views_linears[0], rgb_linear, then kind 1: w_linear; kind 2: sf_linear, prob_linear.
 * Builds the packed fp32 and bf16 (UMMA smem-image) copies the kernels read. */
int zest_net_pack(zest_net* net, const float* const* params, int n_params, void* stream);

/* workspace size in bytes for M rows of the fp32 path (forward; `train` keeps activations). */
int64_t zest_mlp_f32_workspace(const zest_net* net, int64_t M, int train);
/* x [M, ldx] assembled input (pe | feat | views) -> raw [M, out_ch] fp32. */
int zest_mlp_fwd_f32(const zest_net* net, const float* x, int ldx, int64_t M, float* raw,
                     void* workspace, int train, void* stream);
/* backward given the forward's workspace (train=1): graw [M,out_ch] -> gx [M, ldx] (may be NULL)
 * and gparams (same order/shape as params; accumulated, fp32). */
int zest_mlp_bwd_f32(const zest_net* net, const float* x, int ldx, int64_t M, const float* graw,
                     void* workspace, float* gx, float* const* gparams, int n_params,
                     void* stream);

/* bf16 tensor-core path (tcgen05.mma, fp32 accumulate in TMEM).  PE fused in the prologue:
 * inputs are the un-encoded ndc (+ constant t), the gathered feats and the per-ray dirs. */
int zest_mlp_fwd_tc(const zest_net* net, const float* ndc, int ndc_ld, int has_t, float t,
                    const float* feats, int ldf, const float* dirs, int S, int64_t M, float* raw,
                    void* stream);
/* The fused hot path: the same kernel with the feature gather inside.  Replaces gen_pts_feats
 * (renderer.py:51-72: index_point_feature + build_color_volume) + prepare_pts (renderer.py:246-297:
 * Embedding + cat) + run_network (renderer.py:237-242 -> networks.py:150-221) in one launch: loader
 * warps sample the channels-last volume [D,Hv,Wv,8] and the views [V,H,W,4] for the NEXT 128-sample
 * tile (same arithmetic as zest_gather_fwd, bit-identical indices) while the current tile is in the
 * tensor pipe.  feats_out (optional, may be NULL) receives the gathered features in fp32
 * ([M, ldfo]: the reference's `input_feat`).  Needs net.in_feat == 8 + 4 V, V <= 14. */
int zest_gather_mlp_fwd_tc(const zest_net* net, const float* rays_pts, const float* rays_ndc,
                           int ndc_ld, int has_t, float t, const float* vol_cl, int D, int Hv, int Wv,
                           const float* img_cl, int V, int H, int W, const float* cams,
                           const float* dirs, int S, int64_t M, float* feats_out, int ldfo,
                           float* raw, void* stream);
/* same kernel fed with an already assembled x (module boundary MVSNeRF.forward(x)). */
int zest_mlp_fwd_tc_x(const zest_net* net, const float* x, int ldx, int64_t M, float* raw,
                      void* stream);

/* ---- ray builder ("next" row f1) ---------------------------------------------------------------
 * Replaces utils.get_rays_mvs (utils.py:133-230) + build_rays_base (utils.py:290-394: sample depths,
 * world points) + get_ndc_coordinate (utils.py:232-288) for one slab of target pixels.
 * cam: DEVICE table of 54 floats (row-major): K_tgt 3x3 | c2w_tgt 4x4 | w2c_ref 4x4 | K_ref 3x3 |
 * near_tgt far_tgt near_ref far_ref - stream-ordered with whatever produced the pose, no host sync.
 * ys/xs: R target pixel coordinates (fp32) or both NULL = the row-major grid slab [r0, r0 + R) of a
 * W_tgt wide image.  t_vals: [S] = torch.linspace(0, 1, S).  t_rand: [R, S] stratified jitter or
 * NULL.  W_src/H_src: size of the SOURCE views (NDC normalisation, utils.py:318).  Outputs:
 * rays_pts [R,S,3], rays_dir [R,3], rays_ndc [R,S,3], depth [R,S] - bit-identical to the
 * reference's CPU path (separately rounded ops, fma-chain K = 3 matmuls, reciprocal*pad quirk). */
int zest_build_rays(const float* ys, const float* xs, int64_t r0, int W_tgt, const float* cam,
                    int W_src, int H_src, int pad, const float* t_vals, const float* t_rand, int64_t R,
                    int S, float* rays_pts, float* rays_dir, float* rays_ndc, float* depth, void* stream);

/* ---- scene-flow reductions of the training step ("next" row f4) -----------------------------------
 * All three go through utils.NDC2Euclidean (utils.py:507-514) with the image size H, W and focal length f.
 * pts_* are the [R, S, 3] NDC sample positions the render path returns (raw_pts_ref / _post / _prev / _pp).
 * The forward entry points ADD the un-normalised sum to *loss_sum (double, device; the caller zeroes it and divides by
 * the element count); the backward ones take the upstream gradient of the MEAN loss as a device scalar g_loss and write
 * (not accumulate) the point gradients; any of them may be NULL.
 * zest_sf_smooth_loss: losses.py:142-161 compute_sf_smooth_loss - mean |sf_s - sf_{s+1}| over samples s < n_close - 1 of
 *   sf = E(pts_1) - E(pts_2), n_close = int(0.95 S); element count R (n_close - 1) 3.
 * zest_sf_lke_loss: losses.py:164-203 compute_sf_lke_loss - 0.5 mean ((E(post) - E(ref)) - (E(ref) - E(prev)))^2 over
 *   s < n_close = int(0.9 S); loss_sum receives the sum of squares; element count R n_close 3. */
int zest_sf_smooth_loss_fwd(const float* pts_1, const float* pts_2, int64_t R, int S, int n_close, int H, int W,
                            float f, double* loss_sum, void* stream);
int zest_sf_smooth_loss_bwd(const float* pts_1, const float* pts_2, int64_t R, int S, int n_close, int H, int W,
                            float f, const float* g_loss, float* g_pts_1, float* g_pts_2, void* stream);
int zest_sf_lke_loss_fwd(const float* pts_ref, const float* pts_post, const float* pts_prev, int64_t R, int S,
                         int n_close, int H, int W, float f, double* loss_sum, void* stream);
int zest_sf_lke_loss_bwd(const float* pts_ref, const float* pts_post, const float* pts_prev, int64_t R, int S,
                         int n_close, int H, int W, float f, const float* g_loss, float* g_ref, float* g_post,
                         float* g_prev, void* stream);
/* utils.py:516-539 projection_from_ndc: pts_2d[r] = pixel of w2c * E(sum_s weights[r,s] raw_pts[r,s,:]).  w2c: device
 * pointer to a row-major 4x4 world-to-camera matrix.  Backward: g_pts_2d [R,2] -> g_weights [R,S], g_raw_pts [R,S,3]. */
int zest_project_ndc_fwd(const float* w2c, const float* weights, const float* raw_pts, int64_t R, int S, int H, int W,
                         float f, float* pts_2d, void* stream);
int zest_project_ndc_bwd(const float* w2c, const float* weights, const float* raw_pts, int64_t R, int S, int H, int W,
                         float f, const float* g_pts_2d, float* g_weights, float* g_raw_pts, void* stream);

/* ---- frame driver: peer copy ---------------------------------------------------------------
 * cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault) on `stream`: `src` may be another GPU's buffer mapped through CUDA IPC
 * (driver.FrameRenderer's "ipc" transport pulls the source rank's packed frame slot with it). */
int zest_memcpy_async(void* dst, const void* src, int64_t bytes, void* stream);

/* ---- plane-sweep cost volume ("next" row f3, first half) ---------------------------------------------
 * networks.py:1077-1140 MVSNet.build_volume_cost + utils.py:49-99 homo_warp in one pass.
 * feats_cl [V, C/4, H, W, 4] feature maps as planes of channel quads (view 0 = reference, C in {4, 8, 16, 32}), imgs_cl [V, H, W, 4]
 * the images at feature resolution (r, g, b, 0), proj [(V - 1), 12] (DEVICE) = rows of the 3x4 "src_proj @ ref_proj_inv" of
 * every source view (V - 1 <= 9), depth [D] plane depths.  Hp = H + 2 pad, Wp = W + 2 pad.  The volume has 9 + C channels
 * whatever V is, like the reference's (`torch.empty((B, 9 + 32, ...))`, networks.py:1101; the warped images of source views
 * beyond the second fall on channels the variance overwrites, :1138): reference image (zero outside the unpadded window),
 * warped images of source views 1 and 2, variance of the features over all V views.
 *   channels_last == 0: img_feat [9 + C, D, Hp, Wp] (the reference's layout), in_masks [V, D, Hp, Wp] (= 1 for the reference,
 *                       -1 < grid < 1 for the source views; may be NULL)
 *   channels_last != 0: img_feat [D, Hp, Wp, cpad] with 9 + C <= cpad <= 12 + C, cpad % 4 == 0, pad channels zero - the input
 *                       layout of zest_conv_cl_fwd; in_masks is not written */
int zest_cost_volume_fwd(const float* feats_cl, const float* imgs_cl, const float* proj, const float* depth,
                         int V, int C, int H, int W, int D, int pad, float* img_feat, float* in_masks,
                         int channels_last, int cpad, void* stream);
/* Backward wrt the feature maps: g_var = gradient of the C variance channels ([C, D, Hp, Wp], i.e. img_feat + 9 channels),
 * g_feats_cl [V, C/4, H, W, 4] is ACCUMULATED into (zero it first).  Images, projections and depths are data.  V - 1 <= 4. */
int zest_cost_volume_bwd(const float* feats_cl, const float* proj, const float* depth, int V, int C, int H, int W,
                         int D, int pad, const float* g_var, float* g_feats_cl, void* stream);

/* ---- encoding-volume CNNs ("next" row f3, second half): networks.py:935-1059 FeatureNet / CostRegNet ------------
 * Channels-last fp32 activations [N (images) or D (planes), H, W, C].  Weights are repacked once per parameter version:
 * w = nn.Conv2d/3d weight [cout, cin, kd, kh, kw] (transposed == 0) or nn.ConvTranspose3d weight [cin, cout, 3, 3, 3]
 * (transposed != 0) -> packed [cout / 8][kd kh kw][cin_pad][8], cin_pad % 4 == 0 (extra input channels get zero weights). */
int zest_conv_pack_weights(const float* w, int cout, int cin, int kd, int kh, int kw, int transposed, int cin_pad,
                           float* packed, void* stream);
/* y [No, Ho, Wo, cout] = conv(x [N, H, W, cin]) (+ bias), "same" padding k / 2, stride 1 or 2 (kd == 1: 2-D convolution of
 * N images; kd == 3: 3-D over N planes).  Instantiated: 3x3x3 s1/s2, 1x3x3 s1, 1x5x5 s2, 1x1x1 s1.  stats (optional): double
 * [2 cout], zeroed and filled with per-channel sum and sum of squares of y (the batch statistics InPlaceABN normalises with). */
int zest_conv_cl_fwd(const float* x, int N, int H, int W, int cin, const float* wpacked, const float* bias, int cout,
                     int kd, int kh, int kw, int stride, float* y, double* stats, void* stream);
/* nn.ConvTranspose3d(cin, cout, 3, padding=1, output_padding=1, stride=2, bias=False): y [2 D, 2 H, 2 W, cout]. */
int zest_convt3_cl_fwd(const float* x, int D, int H, int W, int cin, const float* wpacked, int cout, float* y,
                       double* stats, void* stream);
/* InPlaceABN forward over n rows of C channels: y = leaky_relu(bn(x), slope) (+ skip, an already activated tensor of the same
 * shape: the U-Net additions networks.py:1049-1055).  training != 0: batch statistics from `stats` (zest_conv_cl_fwd), running
 * statistics updated with `momentum` (unbiased variance) when given; training == 0: running statistics.  y may alias x. */
int zest_bn_act_cl(const float* x, int64_t n, int C, const double* stats, const float* gamma, const float* beta,
                   float* running_mean, float* running_var, float eps, float momentum, float slope, int training,
                   const float* skip, float* y, void* stream);
/* F.interpolate(mode="bilinear", align_corners=False) of packed images [V, H, W, 4] -> [V, h, w, 4] (networks.py:1102). */
int zest_resize_bilinear_cl(const float* x, int V, int H, int W, int h, int w, float* y, void* stream);

/* ---- alpha compositing (warp per ray) ---------------------------------------------------- */
/* raw [R*S, ld_raw] (rgb_raw 3, sigma_raw 1, ...), z [R,S], cos_angle [R], noise [R,S] or NULL
 * (already scaled by raw_noise_std).  Outputs: rgb_map [R,3], depth_map [R], acc [R] (optional),
 * weights [R,S] (optional), alpha [R,S] (optional).  t_stop > 0 enables the early-termination
 * mask: once transmittance < t_stop the remaining samples are skipped except the tail sample
 * (whose dist is 1e10); 0 reproduces the reference exactly. */
int zest_composite_static_fwd(const float* raw, int ld_raw, const float* z, const float* cos_angle,
                              const float* noise, int64_t R, int S, int white_bkgd, float t_stop,
                              float* rgb_map, float* depth_map, float* acc, float* weights,
                              float* alpha, void* stream);
int zest_composite_static_bwd(const float* raw, int ld_raw, const float* z, const float* cos_angle,
                              const float* noise, int64_t R, int S, int white_bkgd,
                              const float* g_rgb_map, const float* g_depth_map,
                              const float* g_weights, const float* g_alpha, float* g_raw,
                              int ld_graw, void* stream);
/* blended static+dynamic composite.  raw_dy [R*S, ld_dy], raw_rig [R*S, ld_rig] (blend weight b
 * in column 4 of raw_rig, already sigmoid-ed by the net). Outputs: rgb_map_ref [R,3],
 * depth_map_ref [R], rgb_map_ref_dy [R,3], depth_map_ref_dy [R], weights_map_dd [R],
 * weights_ref_dy [R,S] (optional, the dynamic-only weights). */
int zest_composite_blend_fwd(const float* raw_dy, int ld_dy, const float* raw_rig, int ld_rig,
                             const float* z, const float* cos_angle, const float* noise, int64_t R,
                             int S, float t_stop, float* rgb_map, float* depth_map, float* rgb_map_dy,
                             float* depth_map_dy, float* weights_dd, float* weights_dy,
                             void* stream);
int zest_composite_blend_bwd(const float* raw_dy, int ld_dy, const float* raw_rig, int ld_rig,
                             const float* z, const float* cos_angle, const float* noise, int64_t R,
                             int S, const float* g_rgb_map, const float* g_depth_map,
                             const float* g_rgb_map_dy, const float* g_depth_map_dy,
                             const float* g_weights_dy, float* g_raw_dy, int ld_gdy,
                             float* g_raw_rig, int ld_grig, void* stream);

/* ---- diagnostics -------------------------------------------------------------------------- */
/* Runs D[128,N] = A[128,K] * B[N,K]^T through the tcgen05 path with the library's own smem
 * layouts (A, B bf16 row-major in global; D fp32 row-major).  Unit-test hook for the UMMA
 * descriptors; N multiple of 16 <= 256, K multiple of 16 <= 256.  variant 0 = A from shared
 * memory (SS form); 1 = A as bf16 pairs in TMEM written with tcgen05.st (TS form, K % 32 == 0):
 * the way the MLP kernel keeps its layer activations. */
int zest_tc_selftest(const uint16_t* A, const uint16_t* B, float* D, int N, int K, int variant,
                     void* stream);
/* Debug: device buffer of 4096 uint64 that builds with -DZEST_TC_TIMELINE fill with
 * (tag << 48 | clock64) records of the MLP kernel's CTA 0 (no-op in normal builds). */
int zest_tc_set_timeline(unsigned long long* buf);
/* Characterisation probe: one CTA issues reps x 16 back-to-back UMMAs (M = 128, N, K = 16; ts = A from
 * TMEM), optionally under TMEM-load traffic from its other warps.  out[0] = cycles to completion. */
int zest_tc_rate_probe(int N, int reps, int ts, int ld_traffic, long long* out, void* stream);
/* GEMM engine of the fp32 MLP path (zest_mlp_fwd_f32 / _bwd_f32).  The tensor-core engines split every fp32 operand
 * into head + residual on the fly and issue three UMMAs per product (fp32 accumulate in TMEM):
 *   2 (default) = tcgen05 3 x tf32 (22 operand bits) with the head product and the cross products in separate TMEM
 *       accumulators (the tensor core's fp32 accumulate truncates; this keeps the bias at the fp32 level): gradients of
 *       the 4096-ray fine-tune batch agree with reference autograd to 8e-4 rel L2, like the CUDA-core engine (8e-4);
 *   1 = tcgen05 3 x bf16 (16 operand bits, one accumulator): 1.4x faster, 3.6e-3 on the same gradients;
 *   0 = exact-fp32 CUDA-core kernel (3.8x slower than 1).
 * Returns the previous engine.  Env ZEST_GEMM=simt|bf16x3 selects 0|1 at load. */
int zest_set_gemm_engine(int engine);
/* Unit-test hook for all engines: C[I,J] (+)= sum_k A[i*sa_i + k*sa_k] * B[j*sb_j + k*sb_k] (+ bias[j]); splits > 1
 * splits the K range over CTAs (atomic accumulation into C, which the caller zeroes or pre-fills).  b_scratch
 * (optional, device, b_scratch_bytes >= ceil(J/256) * ceil(K/16) * 32 KiB) lets the tensor-core engines pack B into
 * UMMA stage images once and stream it by TMA (the MLP path does this for its weight operands). */
int zest_gemm_f32(const float* A, int64_t sa_i, int64_t sa_k, const float* B, int64_t sb_j, int64_t sb_k,
                  float* C, int64_t ldc, int64_t I, int J, int64_t K, const float* bias, int accumulate,
                  int splits, int engine, void* b_scratch, int64_t b_scratch_bytes, void* stream);
/* How many kernels this library has launched since load (bench.py's gpu_launches). */
int64_t zest_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* ZEST_B200_H */
