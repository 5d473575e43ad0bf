// Packed parameters of one radiance MLP (networks.py:73-132 Renderer, v0, use_viewdirs).
#pragma once
#include <atomic>

#include "common.cuh"

struct zest_net {
  int kind;      // 0 plain, 1 static + blend weight, 2 dynamic (scene flow + probs)
  int in_pts, in_feat, in_views, width, depth, skip;
  int out_ch;    // 4 / 5 / 12
  int n_small;   // rows of the stacked small heads [alpha | w] or [alpha | sf(6) | prob(2)]
  int n_params;  // number of tensors zest_net_pack expects
  bool packed;

  // ---- fp32 copy: one device blob, nn.Linear layout ([out,in] row-major) ----
  float* f32;
  int64_t f32_floats;
  // offsets (in floats) into f32
  int64_t w_pts[16], b_pts[16];  // pts_linears
  int64_t w_gate, b_gate;        // pts_bias
  int64_t w_feat, b_feat;        // feature_linear
  int64_t w_small, b_small;      // alpha_linear (+ w_linear | sf_linear, prob_linear) stacked rows
  int64_t w_views, b_views;      // views_linears[0]
  int64_t w_rgb, b_rgb;          // rgb_linear
  // where each user-visible parameter tensor lives in the blob (for pack / grad scatter)
  int64_t param_off[32];
  int64_t param_numel[32];

  // ---- bf16 tensor-core image (built by mlp_tc.cu) ----
  void* tc_blob;       // device: weight stages in UMMA smem-image order
  int64_t tc_bytes;
  float* tc_bias;      // device: [11][256] fp32 biases of the 256-wide ops, read by the tensor-core kernel's epilogue
  void* tc_plan_host;  // host: layer plan (opaque to everything but mlp_tc.cu)
  void* tc_desc_dev;   // device: pack descriptors (uploaded once)
  int* tc_counters;    // device: ring of tile-scheduler counters (one per launch in flight)
  std::atomic<unsigned> tc_counter_next;   // host threads launching the same net concurrently must get distinct counters
  bool tc_dirty;       // f32 changed since the bf16 image was built: rebuilt lazily by the next tensor-core launch
};

namespace zest {
int in_layer(const zest_net* n, int layer);  // input width of pts_linears[layer]
int tc_pack(zest_net* net, cudaStream_t st);  // mlp_tc.cu
void tc_free(zest_net* net);
}  // namespace zest
