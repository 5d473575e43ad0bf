// bf16 tensor-core radiance MLP for sm_100a: tcgen05.mma with fp32 accumulators AND the layer
// activations in TMEM.
//
// One persistent CTA per SM walks 128-sample tiles (M = 128 = one UMMA tile; one ray when S = 128)
// through the whole network of networks.py:150-221 without the activations ever leaving the SM:
//
//   prologue   PE(ndc[,t]) (networks.py:48-65), gathered feats, PE(dir)  -> bf16 A operands in smem
//   GATE       g = pts_bias(feat)                 -> bf16 pairs parked in TMEM (reused by 8 layers)
//   L0..L7     h = relu((W h + b) * g), skip [pe | h] into L5
//   FEAT       feature_linear(h7)  (+ the N = 16 heads [alpha | w | sf | prob] riding in the same op)
//   VIEWS/RGB  relu(views([feat | dirpe])) -> rgb_linear (N = 16 MMA) -> raw[M, out_ch] fp32
//
// TMEM map (512 columns x 128 lanes, lane = sample row):
//   [0,128) [128,256)  fp32 accumulators of the two N = 128 halves ("parts") of a 256-wide layer
//   [256,384)          the per-sample gate as bf16 pairs (dead after L7: the N = 16 head accumulators
//                      alias its first 32 columns)
//   [384,512)          the layer activation h as bf16 pairs = the A operand of the next layer's
//                      TS-form tcgen05.mma (A from TMEM, B = weights from shared memory)
// The activation is updated IN PLACE.  A layer is issued as four ring stages in the order
//   (part 0, K 0..127) (part 1, K 0..127) (part 0, K 128..255 + bias -> commit acc_full[0]) (part 1, ... -> acc_full[1])
// so that when acc_full[0] fires every MMA that reads the low K half of the old activation has
// retired and part 0's epilogue may overwrite exactly those columns (its outputs ARE the low K half
// of the next layer's input); the same holds for part 1 and the high half.  The next layer's low-K
// stages start as soon as part 0's epilogue has stored, while part 1 is still in its epilogue.
// Shared memory therefore only holds the small per-tile inputs (PE, dirPE, feats: SS-form MMAs)
// and a deep weight ring, and the MMA operand traffic out of shared memory is halved.
//
// Warp roles (384 or 448 threads): warps 0-7 epilogue (TMEM lane quarter = warp % 4, column half = warp / 4),
// warps 10-11 (or 10-13 when V > 6) loaders (two / one sample rows per thread: they gather the NEXT tile's features from the encoding
// volume / source views, encode PE and stage everything as bf16 operands while the current tile is in
// the tensor pipe), warp 8 = weight producer (cp.async.bulk / UBLKCP, weights pre-packed on the
// host side of the ABI in the exact smem image, consumption order), warp 9 = MMA issuer + TMEM
// owner.  The weight ring has 4 slots of 36 KB and every op consumes whole ring revolutions (the
// plan pads with empty stages), so slot numbers, descriptors and barrier addresses in the issue
// loop are compile-time constants: one 256x256 layer = 4 stages of 8 UMMAs (N = 128, K = 16).
// Biases ride in the MMA: the last weight stage of every accumulator group carries one extra K = 16
// block whose first two columns hold the bias split into bf16 hi + lo, multiplied against a constant
// "ones" A chunk, so the epilogue is multiply-by-gate / relu / pack only (packed f32x2 / bf16x2 ops).
//
// Synchronisation is mbarrier-only after start-up: ring full/empty, acc_full[part] (tcgen05.commit),
// acc_free[part] and a_ready[part] (epilogue -> MMA).  a_ready[0] also means "K columns 0..127 of
// the new activation are in TMEM", a_ready[1] the same for 128..255, so the next layer's first K half
// runs while part 1 is still in its epilogue.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "gather_core.cuh"
#include "net.cuh"
#include "tc_ptx.cuh"

namespace zest {

constexpr int kTile = 128;
constexpr int kStageBytes = 36864;       // 32 KB of weights (N = 128 x K = 128) + 4 KB bias block
constexpr int kStages = 4;
constexpr int kChunkBytes = kTile * 16;  // one 8-column k-chunk of a 128-row smem operand tile
constexpr int kMaxPlan = 64;
constexpr int kEpiWarps = 8;
// loader warps stage the next tile's inputs (gather + PE): 2 (two rows per thread, 168-register budget for the
// epilogue) or 4 (one row per thread, 128 registers: needed when many source views make the gather the long pole)
constexpr int kMaxViews = 14;             // 8 + 4 V <= 64 feature columns
#define ACC_COL_OF(part) ((uint32_t)(part) * 128u)
constexpr uint32_t GATE_COL = 256, HEAD_COL = 256, HEAD2_COL = 272, ACT_COL = 384;

// Biases of the 256-wide ops: folded into the MMA as one extra K = 16 step per accumulator group (default), or
// (-DZEST_TC_BIAS_IN_EPILOGUE) added by the epilogue in fp32 (packed FADD2 against a bias table in shared memory).
// Same-box A/B on cfg2: the epilogue variant saves 6 % of the issued UMMAs but is 2.4 % slower (the epilogue sits on
// the critical path between layers; the tensor pipe has the slack).
#ifdef ZEST_TC_BIAS_IN_EPILOGUE
constexpr bool kBiasInMma = false;
#else
constexpr bool kBiasInMma = true;
#endif
constexpr int kBiasOps = 11;   // GATE, L0..L7, FEAT, VIEWS: 256 fp32 each in the shared-memory bias table

struct TcStage {   // one ring slot's worth of weights (bytes = 0: an empty stage that only keeps the ring aligned)
  uint32_t src_off;  // byte offset in the packed blob
  uint32_t bytes;
};

struct PackDesc {  // how to build one weight block image from the fp32 blob
  int64_t src_off; int src_ld; int row0; int rows_valid; int col0; int cols_valid; int N; int K; int64_t dst_off;
  int64_t bias_src;  // >= 0: append a K = 16 block [n][0] = bf16 hi, [n][1] = bf16 lo of bias[row0 + n]
};

struct TcPlanHost {
  std::vector<TcStage> stages;
  int n_stages = 0, n_packs = 0;
  int P, Ppad, F, Fpad, C, overlap, gate_fp32, cluster2;
  size_t smem_bytes;
};

struct TcParams {
  TcStage plan[kMaxPlan]; int n_stages;  // by value: lives in the constant bank -> uniform loads in the producer loop
  const uint8_t* blob;
  const float* bias;   // [kBiasOps][256] fp32 (epilogue-side biases)
  // inputs (fused mode) or x (x mode)
  const float* ndc; int ndc_ld; int has_t; float t;
  const float* feats; int ldf;
  const float* dirs; int S;
  // fused gather (vol != nullptr): the loader warps sample the encoding volume + source views themselves
  const float* pts; const float* vol; int D, Hv, Wv; const float* img; int V, H, W; const float* cams;
  float* feats_out; int ldfo;   // optional fp32 copy of the gathered features (the reference's input_feat)
  const float* x; int ldx;
  int P, Ppad, F, Fpad;
  int kind, out_ch, overlap;
  int64_t M; int64_t n_tiles;
  int* tile_counter;   // dynamic tile scheduler: next unclaimed tile (zeroed before the launch); nullptr = static round-robin
  float* raw;
  unsigned long long* tl;  // debug timeline buffer (ZEST_TC_TIMELINE builds only)
};

// ---- debug timeline (compiled in with -DZEST_TC_TIMELINE): (tag << 48 | clock) records of CTA 0, tile #3 ----
#ifdef ZEST_TC_TIMELINE
#define TL(seg, tag) do { if (tl_on) { p.tl[(seg) * 256 + tl_n[0]] = ((unsigned long long)(tag) << 48) | (clock64() & 0xFFFFFFFFFFFFull); tl_n[0]++; } } while (0)
#else
#define TL(seg, tag) do { } while (0)
#endif

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void wait_bar(uint32_t bar, uint32_t parity, int tag) {
  if (ptx::mbar_try_wait(bar, parity)) return;
  uint32_t spins = 0;
  while (!ptx::mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 21)) {  // a protocol bug must not hang the GPU (each try_wait suspends for a while)
      printf("zest mlp_tc: barrier timeout tag=%d block=%d thread=%d parity=%u\n", tag, blockIdx.x, threadIdx.x, parity);
      __trap();
    }
  }
}

// PE via the double-angle recurrence from one accurate sincosf at 2^0 (error ~2^k ulp, two orders of
// magnitude below the bf16 rounding applied right after): out[(1+2k)*C + c] = sin(2^k v_c), next = cos.
template <int C, int NF>
__device__ __forceinline__ void pe_row(const float (&v)[4], float* out /* C*(2NF+1) */) {
#pragma unroll
  for (int c = 0; c < C; ++c) {
    out[c] = v[c];
    float s, co;
    sincosf(v[c], &s, &co);
#pragma unroll
    for (int k = 0; k < NF; ++k) {
      out[(1 + 2 * k) * C + c] = s;
      out[(2 + 2 * k) * C + c] = co;
      const float s2 = 2.f * s * co;
      co = 1.f - 2.f * s * s;
      s = s2;
    }
  }
}

// write NPAD fp32 values of one row as bf16 k-chunks of a K-major, no-swizzle smem operand tile
template <int NPAD>
__device__ __forceinline__ void store_row_chunks(uint32_t buf, int chunk0, int row, const float* v) {
#pragma unroll
  for (int c = 0; c < NPAD / 8; ++c)
    ptx::st_smem_v4(buf + (chunk0 + c) * kChunkBytes + row * 16, ptx::pack_bf16(v[8 * c], v[8 * c + 1]),
                    ptx::pack_bf16(v[8 * c + 2], v[8 * c + 3]), ptx::pack_bf16(v[8 * c + 4], v[8 * c + 5]),
                    ptx::pack_bf16(v[8 * c + 6], v[8 * c + 7]));
}

struct Bars {
  uint32_t full, empty, acc_full, acc_free, a_ready;  // barrier arrays (8 bytes per barrier)
};

// ---- epilogue of one accumulator part (bias already inside the accumulator) ----------------------
// MODE 0: bf16(acc) * gate, relu -> bf16 pairs  (L0..L7; product and relu in one HFMA2.BF16.RELU)
// MODE 4: acc * gate in fp32, relu -> bf16 pairs (L0..L7, ZEST_TC_GATE_FP32=1: one rounding less per activation)
// MODE 1: acc -> bf16 pairs                      (FEAT)
// MODE 2: acc, relu -> bf16 pairs                (VIEWS)
// MODE 3: acc -> bf16 pairs                      (GATE; stored to the gate columns)
// A thread owns two 32-column runs of the part's 128 columns: sub-half 0 = [hsel*32, +32) and sub-half 1 =
// [64 + hsel*32, +32), so that the eight warps together finish columns 0..63 first: those are K columns 0..63 of the
// next layer's input and get their own a_ready barrier (the next layer's first four UMMAs start half an epilogue earlier).
template <int MODE>
__device__ __forceinline__ uint32_t epi_pair(float v0, float v1, float b0, float b1, uint32_t g01) {
  if (!kBiasInMma) ptx::add_f32x2(v0, v1, b0, b1);
  if (MODE == 4) ptx::mul_f32x2(v0, v1, ptx::bf16_lo(g01), ptx::bf16_hi(g01));
  uint32_t pk = ptx::pack_bf16(v0, v1);
  if (MODE == 0) pk = ptx::mul_relu_bf16x2(pk, g01);
  if (MODE == 4 || MODE == 2) pk = ptx::relu_bf16x2(pk);
  return pk;
}

// 32 accumulator columns -> 16 packed pairs.  bias_s: shared-memory address of this run's 32 fp32 biases (the same
// address in every lane: broadcast reads).
template <int MODE>
__device__ __forceinline__ void epi_math(const uint32_t (&acc)[32], const uint32_t (&g)[16], uint32_t (&packed)[16], uint32_t bias_s) {
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    uint4 bq = make_uint4(0u, 0u, 0u, 0u);
    if (!kBiasInMma) bq = ptx::ld_smem_v4(bias_s + j * 4);
    packed[j / 2] = epi_pair<MODE>(__uint_as_float(acc[j]), __uint_as_float(acc[j + 1]), __uint_as_float(bq.x), __uint_as_float(bq.y), g[j / 2]);
    packed[j / 2 + 1] = epi_pair<MODE>(__uint_as_float(acc[j + 2]), __uint_as_float(acc[j + 3]), __uint_as_float(bq.z), __uint_as_float(bq.w), g[j / 2 + 1]);
  }
}

// out_base: TMEM column base of the destination (ACT_COL or GATE_COL).  bar_free: arrive once the accumulator is in
// registers.  bar_rdy0 / bar_rdy1: arrive after sub-half 0 / 1 has landed (pass the same barrier twice -> arrive once, at the end).
template <int MODE>
__device__ __forceinline__ void epilogue_part(uint32_t tmem_lane, int part, int hsel, uint32_t out_base, uint32_t bias_op,
                                              uint32_t bar_free, uint32_t bar_rdy0, uint32_t bar_rdy1, int lane, long long* t_ld = nullptr) {
  const uint32_t bias_s = bias_op + (uint32_t)(part * 128 + hsel * 32) * 4u;   // this thread's first run of 32 columns
  const uint32_t acc_t = tmem_lane + ACC_COL_OF(part) + hsel * 32;
  const uint32_t gate_t = tmem_lane + GATE_COL + part * 64 + hsel * 16;
  const uint32_t out_t = tmem_lane + out_base + part * 64 + hsel * 16;
  uint32_t acc[2][32], g[2][16], pk[16];
  ptx::tmem_ld32(acc_t, acc[0]);
  if (MODE == 0 || MODE == 4) ptx::tmem_ld16(gate_t, g[0]);
  ptx::tmem_ld32(acc_t + 64, acc[1]);
  if (MODE == 0 || MODE == 4) ptx::tmem_ld16(gate_t + 32, g[1]);
  ptx::tc_wait_ld();
  if (t_ld) *t_ld = clock64();
  // the accumulator half is drained: the MMA warp may overwrite it
  ptx::tc_fence_before();
  __syncwarp();
  if (lane == 0) ptx::mbar_arrive(bar_free);
  epi_math<MODE>(acc[0], g[0], pk, bias_s);
  ptx::tmem_st16(out_t, pk);
  if (bar_rdy0 != bar_rdy1) {
    ptx::tc_wait_st();
    ptx::tc_fence_before();
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(bar_rdy0);
  }
  epi_math<MODE>(acc[1], g[1], pk, bias_s + 64 * 4);
  ptx::tmem_st16(out_t + 32, pk);
  ptx::tc_wait_st();
  ptx::tc_fence_before();
  __syncwarp();
  if (lane == 0) ptx::mbar_arrive(bar_rdy1);
}

__device__ __forceinline__ void arrive_idle(uint32_t bar_free, uint32_t bar_ready, int lane) {
  __syncwarp();
  if (lane == 0) { ptx::mbar_arrive(bar_free); ptx::mbar_arrive(bar_ready); }
}

// two-part hidden op (L0..L7, FEAT): part p's outputs are K half p of the next layer's input, stored in place.
// a_ready: [0] = K columns 0..63, [2] = 64..127 (part 0's two sub-halves), [1] = 128..255 (part 1).
template <int MODE>
__device__ __forceinline__ void epilogue_two_part(uint32_t tmem_lane, int hsel, uint32_t bias_op, const Bars& b, uint32_t (&nfull)[2], int lane, int tag) {
  wait_bar(b.acc_full, nfull[0]++ & 1, tag);
  ptx::tc_fence_after();
  epilogue_part<MODE>(tmem_lane, 0, hsel, ACT_COL, bias_op, b.acc_free, b.a_ready, b.a_ready + 16, lane);
  wait_bar(b.acc_full + 8, nfull[1]++ & 1, tag + 1);
  ptx::tc_fence_after();
  epilogue_part<MODE>(tmem_lane, 1, hsel, ACT_COL, bias_op, b.acc_free + 8, b.a_ready + 8, b.a_ready + 8, lane);
}

// ---- MMA issue helpers (warp-convergent; every operand warp-uniform; one elected lane issues) ------
constexpr uint32_t kDescHi = (128u >> 4) | (1u << 14);           // SBO = 128 B, descriptor version 1
constexpr uint64_t kHi = (uint64_t)kDescHi << 32;
constexpr uint32_t kALbo = ((uint32_t)kChunkBytes >> 4) << 16;   // LBO(A in smem) = 128 rows * 16 B

struct MmaCtx {
  Bars b;
  uint32_t phase, par, n_issued;
  uint32_t ones_lo, ring_lo;
  uint32_t pair;     // 1: this CTA shares its weight ring with the other CTA of a 2-CTA cluster (slot releases go to both)
  __device__ __forceinline__ void next_op() { par ^= 1u; }
  __device__ __forceinline__ void release_slot(uint32_t bar) {   // elected lane: all MMAs reading the slot have been issued
    if (pair) ptx::mma_commit_multicast(bar, (uint16_t)3); else ptx::mma_commit(bar);
  }
  __device__ __forceinline__ void wait(uint32_t need) {
    if (need & 1u) wait_bar(b.a_ready, par, 200);
    if (need & 2u) wait_bar(b.a_ready + 8, par, 201);
    if (need & 4u) wait_bar(b.acc_free, par, 202);
    if (need & 8u) wait_bar(b.acc_free + 8, par, 203);
    if (need & 16u) wait_bar(b.a_ready + 16, par, 204);
  }
  __device__ __forceinline__ void end_revolution() { phase ^= 1u; }
};

// One ring stage at compile-time SLOT: n_k16 K = 16 steps of N weight rows.
//   TS: A = TMEM columns starting at `a` (8 columns per step);  !TS: A = smem descriptor low word `a`.
//   bias: + one step against the ones chunk (the stage image carries the bias block after the weights)
//   commit_part >= 0: the accumulator group ends here -> commit acc_full[commit_part]
template <int N, bool TS, int SLOT>
__device__ __forceinline__ void mma_stage(MmaCtx& c, uint32_t a, int n_k16, uint32_t d_tmem, bool first, bool bias, int commit_part,
                                          uint32_t mid_need = 0u) {
  constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
  wait_bar(c.b.full + 8 * SLOT, c.phase, 220 + SLOT);
  ptx::tc_fence_after();
  // B descriptor low word; one K = 16 step advances B by 2 chunks of N rows x 16 B
  uint32_t b_lo = (c.ring_lo + SLOT * (kStageBytes >> 4)) | ((uint32_t)N << 16);
  uint32_t acc = first ? 0u : 1u;
  if (mid_need) {   // first half of the K steps now, the rest once `mid_need` has been observed
    const int h = n_k16 / 2;
    if (ptx::elect_one()) {
#pragma unroll 4
      for (int k = 0; k < h; ++k) {
        if (TS) ptx::mma_bf16_ts(d_tmem, a + (TS ? 8u : ((2u * kChunkBytes) >> 4)) * k, kHi | (b_lo + 2u * N * k), kIdesc, k ? 1u : acc);
        else ptx::mma_bf16_ss(d_tmem, kHi | (a + ((2u * kChunkBytes) >> 4) * k), kHi | (b_lo + 2u * N * k), kIdesc, k ? 1u : acc);
      }
    }
    __syncwarp();
    a += (TS ? 8u : ((2u * kChunkBytes) >> 4)) * h; b_lo += 2u * N * h; n_k16 -= h; acc = 1u;
    c.wait(mid_need);
  }
  if (ptx::elect_one()) {
#pragma unroll 8
    for (int k = 0; k < n_k16; ++k) {
      if (TS) ptx::mma_bf16_ts(d_tmem, a, kHi | b_lo, kIdesc, acc);
      else ptx::mma_bf16_ss(d_tmem, kHi | a, kHi | b_lo, kIdesc, acc);
      acc = 1u;
      a += TS ? 8u : ((2u * kChunkBytes) >> 4);
      b_lo += 2u * N;
    }
    if (bias) ptx::mma_bf16_ss(d_tmem, kHi | c.ones_lo, kHi | b_lo, kIdesc, 1u);
    c.release_slot(c.b.empty + 8 * SLOT);
    if (commit_part >= 0) ptx::mma_commit(c.b.acc_full + 8 * commit_part);
  }
  __syncwarp();
  ++c.n_issued;
}

// an empty stage: hand the slot straight back to the producer
template <int SLOT>
__device__ __forceinline__ void mma_skip(MmaCtx& c) {
  wait_bar(c.b.full + 8 * SLOT, c.phase, 230 + SLOT);
  if (ptx::elect_one()) {
    ptx::mbar_arrive(c.b.empty + 8 * SLOT);
    if (c.pair) ptx::mbar_arrive_remote(c.b.empty + 8 * SLOT, ptx::cluster_ctarank() ^ 1u);
  }
  __syncwarp();
  ++c.n_issued;
}

template <int C, bool GATE32, int kLoadWarps, int CL>  // C = 3 (static: xyz) or 4 (dynamic: xyz + t); GATE32: fp32 gate multiply;
                                                       // CL = 2: CTA pairs (clusters) share every weight stage via TMA multicast
__global__ void __launch_bounds__(32 * (kEpiWarps + 2 + kLoadWarps), 1) mlp_tc_kernel(const __grid_constant__ TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t s_bars[2 * kStages + 14];
  __shared__ int s_tile[4];
  __shared__ float s_cams[kMaxViews * 24];
  __shared__ uint32_t s_tmem;

  constexpr int kThreads = 32 * (kEpiWarps + 2 + kLoadWarps), kRowsPerLoader = kTile / (32 * kLoadWarps);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // smem: S = [PE (Ppad/8 chunks) | dirPE x 2 (4 + 4) | ones (2) | feats (Fpad/8)] then the weight ring.
  // dirPE is double buffered by tile parity: it is read at the very end of a tile (VIEWS), after the
  // next tile's inputs have been staged; PE (last read by L5) and feats (GATE) are not.
  const uint32_t s_base = ptx::smem_u32(smem_raw);
  const int dir_chunk = p.Ppad / 8, ones_chunk = dir_chunk + 8, feat_chunk = ones_chunk + 2;
  const uint32_t bias_s0 = s_base + (feat_chunk + p.Fpad / 8) * kChunkBytes;   // [kBiasOps][256] fp32
  const uint32_t ring = bias_s0 + (kBiasInMma ? 0 : kBiasOps * 1024);
  const uint32_t bars = ptx::smem_u32(s_bars);
  Bars b;
  b.full = bars; b.empty = bars + 8 * kStages; b.acc_full = bars + 16 * kStages;
  b.acc_free = b.acc_full + 16; b.a_ready = b.acc_free + 16;
  // loader handshake: inputs of tile #it staged (4 loader warps) / feats operand free (GATE retired) / PE operand free (L5 retired)
  const uint32_t in_ready = b.a_ready + 24, feats_free = in_ready + 8, pe_free = in_ready + 16;
  const uint32_t tile_bar = in_ready + 24;   // 4 barriers: tile id #it of this CTA published in s_tile[it & 3]

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) { ptx::mbar_init(b.full + 8 * s, 1); ptx::mbar_init(b.empty + 8 * s, CL); }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(b.acc_full + 8 * i, 1);
      ptx::mbar_init(b.acc_free + 8 * i, kEpiWarps);
      ptx::mbar_init(b.a_ready + 8 * i, kEpiWarps);
    }
    ptx::mbar_init(b.a_ready + 16, kEpiWarps);
    ptx::mbar_init(in_ready, kLoadWarps); ptx::mbar_init(feats_free, 1); ptx::mbar_init(pe_free, 1);
    for (int i = 0; i < 4; ++i) ptx::mbar_init(tile_bar + 8 * i, 1);
    ptx::fence_mbar_init();
  }
  if (!kBiasInMma) for (int i = tid; i < kBiasOps * 256; i += kThreads) ptx::st_smem_u32(bias_s0 + i * 4, __float_as_uint(__ldg(p.bias + i)));
  if (tid < kTile) {  // the constant "ones" A chunk pair: columns 0, 1 = 1.0 (bias hi, lo), the rest 0
    ptx::st_smem_v4(s_base + ones_chunk * kChunkBytes + tid * 16, 0x3F803F80u, 0u, 0u, 0u);
    ptx::st_smem_v4(s_base + (ones_chunk + 1) * kChunkBytes + tid * 16, 0u, 0u, 0u, 0u);
    ptx::fence_proxy_async_smem();
  }
  if (p.vol) {
    for (int i = tid; i < p.V * 24; i += kThreads) s_cams[i] = __ldg(p.cams + i);
    for (int i = tid; i < (p.Fpad / 8) * kTile; i += kThreads) ptx::st_smem_v4(s_base + feat_chunk * kChunkBytes + i * 16, 0u, 0u, 0u, 0u);
    ptx::fence_proxy_async_smem();
  }
  if (warp == kEpiWarps + 1) { ptx::tmem_alloc(ptx::smem_u32(&s_tmem), 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = s_tmem;
  // CL = 2: both CTAs of a pair walk the same number of tiles (the even CTA's count; a surplus tile has no valid row)
  // because every weight stage is loaded half by each and multicast to both
  const int64_t first_cta = (CL == 2) ? (blockIdx.x & ~1u) : blockIdx.x;
  const int64_t my_tiles = (p.n_tiles > first_cta) ? (p.n_tiles - first_cta + gridDim.x - 1) / gridDim.x : 0;
  const uint32_t cta_rank = (CL == 2) ? ptx::cluster_ctarank() : 0u;
  // Tile schedule.  Dynamic (default): the first loader warp claims the CTA's next tile from a global counter and
  // publishes it in s_tile[it & 3] behind tile_bar[it & 3]; every other role picks it up there.  A CTA that starts late
  // or shares its SM with another stream's kernel (a collective waiting for a peer, a copy kernel) simply claims fewer
  // tiles - with the static round-robin one delayed CTA delayed the whole launch by its full tile list.  The leader is
  // never more than two tiles ahead of the slowest role (feats_free), so four slots cannot wrap.
  const bool dyn = (CL == 1) && p.tile_counter != nullptr;
  auto tile_static = [&](int64_t it) -> int64_t { return it < my_tiles ? (int64_t)blockIdx.x + it * (int64_t)gridDim.x : -1; };
  auto tile_lead = [&](int64_t it) -> int64_t {
    if (!dyn) return tile_static(it);
    int t = 0;
    if (lane == 0) {
      t = atomicAdd(p.tile_counter, 1);
      s_tile[it & 3] = t;
      ptx::mbar_arrive(tile_bar + 8 * (uint32_t)(it & 3));   // release: orders the store above
    }
    t = __shfl_sync(0xffffffffu, t, 0);
    return t < p.n_tiles ? (int64_t)t : -1;
  };
  auto tile_follow = [&](int64_t it) -> int64_t {
    if (!dyn) return tile_static(it);
    wait_bar(tile_bar + 8 * (uint32_t)(it & 3), (uint32_t)((it >> 2) & 1), 500);
    const int t = ((volatile int*)s_tile)[it & 3];
    return t < p.n_tiles ? (int64_t)t : -1;
  };
  if (CL == 2) ptx::cluster_sync();   // the peer's barriers are initialised before anything is multicast into this CTA

  if (warp == kEpiWarps) {
    // ===================== weight producer =====================
    // The whole warp runs the loop convergently (addresses stay in uniform registers); one elected
    // lane arms the barrier and issues the bulk copy.  Stage s always lands in slot s % 4.
    uint32_t phase = 0;
    for (int64_t it = 0;; ++it) {
      if (tile_follow(it) < 0) break;
      for (int s0 = 0; s0 < p.n_stages; s0 += kStages) {
#pragma unroll
        for (int slot = 0; slot < kStages; ++slot) {
          const uint32_t bytes = p.plan[s0 + slot].bytes, src_off = p.plan[s0 + slot].src_off;
          wait_bar(b.empty + 8 * slot, phase ^ 1, 100 + slot);
          if (ptx::elect_one()) {
            if (bytes) {
              ptx::mbar_arrive_expect_tx(b.full + 8 * slot, bytes);
              if (CL == 2)   // this CTA fetches its half of the stage and multicasts it into both rings
                ptx::bulk_g2s_multicast(ring + slot * kStageBytes + cta_rank * (bytes / 2), p.blob + src_off + cta_rank * (bytes / 2), bytes / 2,
                                        b.full + 8 * slot, (uint16_t)3);
              else
                ptx::bulk_g2s(ring + slot * kStageBytes, p.blob + src_off, bytes, b.full + 8 * slot);
            } else {
              ptx::mbar_arrive(b.full + 8 * slot);
            }
          }
          __syncwarp();
        }
        phase ^= 1;
      }
    }
  } else if (warp == kEpiWarps + 1) {
    // ===================== MMA issuer =====================
    // a_ready[0,1] / acc_free[0,1] complete exactly once per op: completion #(12 it + j) is the
    // epilogue of op j-1 (j = 0: the tile prologue, which also stands for the previous tile's RGB
    // epilogue).  mbarrier parity waits are only sound if the MMA warp observes EVERY completion, in
    // order, and before the next one can happen; every op below therefore waits on all four barriers
    // at least once before its last commit.  The schedule is straight-line code that must walk the
    // ring stages in exactly the order tc_pack() laid them out (checked per tile against n_stages).
    int tl_n[2] = {0, 0}; (void)tl_n;
    MmaCtx c{b, 0u, 1u, 0u, (((s_base + ones_chunk * kChunkBytes) >> 4) & 0x3FFF) | kALbo, (ring >> 4) & 0x3FFF, (uint32_t)(CL == 2)};
    const uint32_t pe_lo = ((s_base >> 4) & 0x3FFF) | kALbo;
    const uint32_t dir_lo[2] = {pe_lo + dir_chunk * (kChunkBytes >> 4), pe_lo + (dir_chunk + 4) * (kChunkBytes >> 4)};
    const uint32_t feat_lo = pe_lo + feat_chunk * (kChunkBytes >> 4);
    const uint32_t acc0 = tmem + ACC_COL_OF(0), acc1 = tmem + ACC_COL_OF(1), act = tmem + ACT_COL;
    const bool ov = p.overlap != 0;
    const int nk_f = p.Fpad / 16, nk_p = p.Ppad / 16;
    for (int64_t it = 0;; ++it) {
      if (!dyn && it >= my_tiles) break;
      c.n_issued = 0;
      // ---- revolution 0: GATE (feats, SS) | L0 (PE, SS) ----
      wait_bar(in_ready, (uint32_t)(it & 1), 210);   // the loader warps have staged this tile's operands
      // dynamic schedule: the loaders also arrive on in_ready for the end-of-work sentinel, published before that arrival
      if (dyn && ((volatile int*)s_tile)[it & 3] >= p.n_tiles) break;
      c.next_op(); c.wait(31u);
      mma_stage<128, false, 0>(c, feat_lo, nk_f, acc0, true, kBiasInMma, 0);
      mma_stage<128, false, 1>(c, feat_lo, nk_f, acc1, true, kBiasInMma, 1);
      if (ptx::elect_one()) ptx::mma_commit(feats_free);   // feats operand may be overwritten once GATE has retired
      __syncwarp();
      c.next_op();
      if (!ov) c.wait(31u);
      c.wait(4u | 1u | 16u);   // every barrier of part X must be observed before the commit that lets X complete again
      mma_stage<128, false, 2>(c, pe_lo, nk_p, acc0, true, kBiasInMma, 0);
      c.wait(8u | 2u);
      mma_stage<128, false, 3>(c, pe_lo, nk_p, acc1, true, kBiasInMma, 1);
      c.end_revolution();
      // ---- L1..L7: (p0, K lo) (p1, K lo) (p0, K hi) (p1, K hi); L5 = [pe | h4] starts with the PE blocks (smem) ----
      for (int l = 1; l < 8; ++l) {
        c.next_op();
        if (!ov) c.wait(31u);
        if (l == 5) {
          c.wait(4u);
          mma_stage<128, false, 0>(c, pe_lo, nk_p, acc0, true, false, -1);
          c.wait(8u);
          mma_stage<128, false, 1>(c, pe_lo, nk_p, acc1, true, false, -1);
          if (ptx::elect_one()) ptx::mma_commit(pe_free);   // last reader of the PE operand
          __syncwarp();
          c.wait(1u);
          mma_stage<128, true, 2>(c, act, 8, acc0, false, false, -1, 16u);
          mma_stage<128, true, 3>(c, act, 8, acc1, false, false, -1);
          c.end_revolution();
          c.wait(2u);
          mma_stage<128, true, 0>(c, act + 64, 8, acc0, false, kBiasInMma, 0);
          mma_stage<128, true, 1>(c, act + 64, 8, acc1, false, kBiasInMma, 1);
          mma_skip<2>(c);
          mma_skip<3>(c);
          c.end_revolution();
        } else {
#ifdef ZEST_TC_TIMELINE
          const bool tl_on = p.tl && blockIdx.x == 0 && it == 3 && lane == 0;
#endif
          c.wait(4u | 1u);
          TL(9, 100 * l + 50);
          mma_stage<128, true, 0>(c, act, 8, acc0, true, false, -1, 16u);
          TL(9, 100 * l + 51);
          c.wait(8u);
          TL(9, 100 * l + 52);
          mma_stage<128, true, 1>(c, act, 8, acc1, true, false, -1);
          TL(9, 100 * l + 53);
          c.wait(2u);
          TL(9, 100 * l + 54);
          mma_stage<128, true, 2>(c, act + 64, 8, acc0, false, kBiasInMma, 0);
          TL(9, 100 * l + 55);
          mma_stage<128, true, 3>(c, act + 64, 8, acc1, false, kBiasInMma, 1);
          TL(9, 100 * l + 56);
          c.end_revolution();
        }
      }
      // ---- FEAT (h7 -> feature); the N = 16 heads (same input; the gate columns are dead) ride behind the low-K stages ----
      c.next_op();
      if (!ov) c.wait(31u);
      c.wait(4u | 1u);
      mma_stage<128, true, 0>(c, act, 8, acc0, true, false, -1, 16u);
      c.wait(8u);
      mma_stage<128, true, 1>(c, act, 8, acc1, true, false, -1);
      c.wait(2u);
      mma_stage<16, true, 2>(c, act, 16, tmem + HEAD_COL, true, true, -1);
      mma_stage<128, true, 3>(c, act + 64, 8, acc0, false, kBiasInMma, 0);
      c.end_revolution();
      mma_stage<128, true, 0>(c, act + 64, 8, acc1, false, kBiasInMma, 1);
      // ---- VIEWS: [feature (TMEM) | dirPE (smem)] -> acc0 (N = 128, single part); RGB (N = 16) from v ----
      c.next_op();
      if (!ov) c.wait(31u);
      c.wait(4u | 1u);
      mma_stage<128, true, 1>(c, act, 8, acc0, true, false, -1, 16u);
      c.wait(2u | 8u);
      mma_stage<128, true, 2>(c, act + 64, 8, acc0, false, false, -1);
      mma_stage<128, false, 3>(c, dir_lo[it & 1], 2, acc0, false, kBiasInMma, 0);
      c.end_revolution();
      c.next_op(); c.wait(31u);
      mma_stage<16, true, 0>(c, act, 8, tmem + HEAD2_COL, true, true, 0);
      mma_skip<1>(c);
      mma_skip<2>(c);
      mma_skip<3>(c);
      c.end_revolution();
      if (c.n_issued != (uint32_t)p.n_stages) {
        if (lane == 0) printf("zest mlp_tc: MMA schedule walked %u stages, plan has %d\n", c.n_issued, p.n_stages);
        __trap();
      }
    }
  } else if (warp >= kEpiWarps + 2) {
    // ===================== loader warps: stage the operands of the tiles, one tile ahead of the tensor pipe =====
    // Two sample rows per thread.  feats (read by GATE only) and PE (last read by L5) are single buffers guarded by
    // feats_free / pe_free (tcgen05.commit behind their last readers); dirPE (read by VIEWS at the very end of a
    // tile) is double buffered by tile parity.
    const int row0 = (warp - (kEpiWarps + 2)) * 32 + lane;   // this thread stages rows row0 and row0 + 64
    const bool leader = warp == kEpiWarps + 2;
    for (int64_t it = 0;; ++it) {
      const int64_t tile = leader ? tile_lead(it) : tile_follow(it);
      if (tile < 0 && !dyn) break;
      if (it > 0) wait_bar(feats_free, (uint32_t)((it - 1) & 1), 400);
      if (tile < 0) {   // dynamic schedule, no tile left: wake the MMA warp (it reads the sentinel behind in_ready) and leave
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(in_ready);
        break;
      }
#pragma unroll 1
      for (int h = 0; h < kRowsPerLoader; ++h) {
        const int row = row0 + 32 * kLoadWarps * h;
        const int64_t m = tile * kTile + row;
        const bool valid = m < p.M;
        // ---- gathered features: fused trilinear volume sample + per-view bilinear RGB + mask (gather_core.cuh),
        //      or pre-gathered feats / the feat block of x.  The feats operand is free once GATE(it-1) has retired. ----
        if (p.vol) {
          // written straight into the bf16 operand (16 B for the 8 volume channels, 8 B per view) and, optionally,
          // into the fp32 input_feat copy; the pad columns [F, Fpad) were zeroed once at start-up
          const uint32_t frow = s_base + feat_chunk * kChunkBytes + row * 16;
          float* o = p.feats_out ? p.feats_out + m * p.ldfo : nullptr;
          const bool vec = (p.ldfo & 3) == 0;
          float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          float px = 0.f, py = 0.f, pz = 0.f;
          if (valid) {
            const float* n = p.ndc + m * p.ndc_ld;
            trilinear8(p.vol, p.D, p.Hv, p.Wv, __ldg(n), __ldg(n + 1), __ldg(n + 2), acc);
            const float* q = p.pts + m * 3;
            px = __ldg(q); py = __ldg(q + 1); pz = __ldg(q + 2);
            if (o) {
              if (vec) {
                reinterpret_cast<float4*>(o)[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
                reinterpret_cast<float4*>(o)[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
              } else {
#pragma unroll
                for (int k = 0; k < 8; ++k) o[k] = acc[k];
              }
            }
          }
          ptx::st_smem_v4(frow, ptx::pack_bf16(acc[0], acc[1]), ptx::pack_bf16(acc[2], acc[3]), ptx::pack_bf16(acc[4], acc[5]),
                          ptx::pack_bf16(acc[6], acc[7]));
#pragma unroll 1
          for (int v = 0; v < p.V; ++v) {
            float4 c4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid) {
              c4 = view_sample(reinterpret_cast<const float4*>(p.img) + (int64_t)v * p.H * p.W, p.H, p.W, s_cams + v * 24, px, py, pz);
              if (o) {
                if (vec) reinterpret_cast<float4*>(o + 8)[v] = c4;
                else { o[8 + 4 * v] = c4.x; o[9 + 4 * v] = c4.y; o[10 + 4 * v] = c4.z; o[11 + 4 * v] = c4.w; }
              }
            }
            ptx::st_smem_v2(frow + (uint32_t)(1 + (v >> 1)) * kChunkBytes + (uint32_t)(v & 1) * 8u, ptx::pack_bf16(c4.x, c4.y),
                            ptx::pack_bf16(c4.z, c4.w));
          }
        } else {
          float f[64];
#pragma unroll
          for (int i = 0; i < 64; ++i) f[i] = 0.f;
          if (valid) {
            const float* src = p.x ? (p.x + m * p.ldx + p.P) : (p.feats + m * p.ldf);
#pragma unroll
            for (int j = 0; j < 64; ++j) if (j < p.F) f[j] = __ldg(src + j);
          }
          if (p.Fpad <= 32) store_row_chunks<32>(s_base, feat_chunk, row, f);
          else if (p.Fpad <= 48) store_row_chunks<48>(s_base, feat_chunk, row, f);
          else store_row_chunks<64>(s_base, feat_chunk, row, f);
        }
      }
#pragma unroll 1
      for (int h = 0; h < kRowsPerLoader; ++h) {
        const int row = row0 + 32 * kLoadWarps * h;
        const int64_t m = tile * kTile + row;
        const bool valid = m < p.M;
        // ---- direction PE (this tile parity's buffer: its last reader, VIEWS two tiles ago, retired long ago) ----
        {
          float d[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) d[i] = 0.f;
          if (valid) {
            if (p.x) {
#pragma unroll
              for (int j = 0; j < 27; ++j) d[j] = __ldg(p.x + m * p.ldx + p.P + p.F + j);
            } else {
              const float* dp = p.dirs + (m / p.S) * 3;
              float v[4] = {__ldg(dp), __ldg(dp + 1), __ldg(dp + 2), 0.f};
              pe_row<3, 4>(v, d);
            }
          }
          store_row_chunks<32>(s_base, dir_chunk + 4 * (int)(it & 1), row, d);
        }
      }
      if (it > 0) wait_bar(pe_free, (uint32_t)((it - 1) & 1), 401);
#pragma unroll 1
      for (int h = 0; h < kRowsPerLoader; ++h) {
        const int row = row0 + 32 * kLoadWarps * h;
        const int64_t m = tile * kTile + row;
        const bool valid = m < p.M;
        // ---- point PE ----
        {
          float pe[C * 21 + 12];
#pragma unroll
          for (int i = 0; i < C * 21 + 12; ++i) pe[i] = 0.f;
          if (valid) {
            if (p.x) {
#pragma unroll
              for (int j = 0; j < C * 21; ++j) pe[j] = __ldg(p.x + m * p.ldx + j);  // already encoded (P = 21 C)
            } else {
              float v[4] = {__ldg(p.ndc + m * p.ndc_ld), __ldg(p.ndc + m * p.ndc_ld + 1), __ldg(p.ndc + m * p.ndc_ld + 2), p.t};
              pe_row<C, 10>(v, pe);
            }
          }
          if (C == 3) store_row_chunks<64>(s_base, 0, row, pe);
          else store_row_chunks<96>(s_base, 0, row, pe);
        }
      }
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(in_ready);
    }
  } else {
    // ===================== epilogue warps =====================
    const int q = warp & 3, hsel = warp >> 2;
    const int row = q * 32 + lane;
    uint32_t nfull[2] = {0, 0};
    int tl_n[2] = {0, 0}; (void)tl_n;
    const uint32_t tmem_lane = tmem + ((uint32_t)(q * 32) << 16);
    for (int64_t it = 0;; ++it) {
      const int64_t tile = tile_follow(it);
      if (tile < 0) break;
      const int64_t m = tile * kTile + row;
      const bool valid = m < p.M;
      // tile start: stands for "the previous tile's RGB epilogue is done" on all four epilogue -> MMA barriers
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        ptx::mbar_arrive(b.acc_free); ptx::mbar_arrive(b.acc_free + 8);
        ptx::mbar_arrive(b.a_ready); ptx::mbar_arrive(b.a_ready + 8); ptx::mbar_arrive(b.a_ready + 16);
      }
      // ---- GATE: bf16 pairs -> gate columns ----
      for (int part = 0; part < 2; ++part) {
        wait_bar(b.acc_full + 8 * part, nfull[part]++ & 1, 300 + part);
        ptx::tc_fence_after();
        if (part == 0) epilogue_part<3>(tmem_lane, 0, hsel, GATE_COL, bias_s0, b.acc_free, b.a_ready, b.a_ready + 16, lane);
        else epilogue_part<3>(tmem_lane, 1, hsel, GATE_COL, bias_s0, b.acc_free + 8, b.a_ready + 8, b.a_ready + 8, lane);
      }
      // ---- L0..L7 ----
      for (int l = 0; l < 8; ++l) {
#ifdef ZEST_TC_TIMELINE
        const bool tl_on = p.tl && blockIdx.x == 0 && it == 3 && lane == 0;
        for (int part = 0; part < 2; ++part) {
          long long t_ld = 0;
          wait_bar(b.acc_full + 8 * part, nfull[part]++ & 1, 310 + part);
          ptx::tc_fence_after();
          TL(warp, 100 * l + 10 * part + 0);
          if (part == 0) epilogue_part<GATE32 ? 4 : 0>(tmem_lane, 0, hsel, ACT_COL, bias_s0 + (1 + l) * 1024, b.acc_free, b.a_ready, b.a_ready + 16, lane, &t_ld);
          else epilogue_part<GATE32 ? 4 : 0>(tmem_lane, 1, hsel, ACT_COL, bias_s0 + (1 + l) * 1024, b.acc_free + 8, b.a_ready + 8, b.a_ready + 8, lane, &t_ld);
          if (tl_on) { p.tl[warp * 256 + tl_n[0]] = ((unsigned long long)(100 * l + 10 * part + 1) << 48) | (t_ld & 0xFFFFFFFFFFFFull); tl_n[0]++; }
          TL(warp, 100 * l + 10 * part + 3);
        }
#else
        epilogue_two_part<GATE32 ? 4 : 0>(tmem_lane, hsel, bias_s0 + (1 + l) * 1024, b, nfull, lane, 310);
#endif
      }
      // ---- FEAT + heads: the head accumulators (sigma + blend / scene flow / probs) are complete with
      //      FEAT part 0's commit; they are read into registers here, long before the next tile's
      //      GATE epilogue reuses the columns ----
      float head[12];
      wait_bar(b.acc_full, nfull[0]++ & 1, 320);
      ptx::tc_fence_after();
      if (hsel == 0) {
        uint32_t r[16];
        ptx::tmem_ld16(tmem_lane + HEAD_COL, r);
        ptx::tc_wait_ld();
        head[3] = __uint_as_float(r[0]);
        if (p.kind == 1) {
          head[4] = 1.f / (1.f + __expf(-__uint_as_float(r[1])));
        } else if (p.kind == 2) {
#pragma unroll
          for (int k = 0; k < 6; ++k) head[4 + k] = tanhf(__uint_as_float(r[1 + k]));
#pragma unroll
          for (int k = 0; k < 2; ++k) head[10 + k] = 1.f / (1.f + __expf(-__uint_as_float(r[7 + k])));
        }
      }
      epilogue_part<1>(tmem_lane, 0, hsel, ACT_COL, bias_s0 + 9 * 1024, b.acc_free, b.a_ready, b.a_ready + 16, lane);
      wait_bar(b.acc_full + 8, nfull[1]++ & 1, 321);
      ptx::tc_fence_after();
      epilogue_part<1>(tmem_lane, 1, hsel, ACT_COL, bias_s0 + 9 * 1024, b.acc_free + 8, b.a_ready + 8, b.a_ready + 8, lane);
      // ---- VIEWS: relu -> v = activation columns [0, 128) ----
      wait_bar(b.acc_full, nfull[0]++ & 1, 340);
      ptx::tc_fence_after();
      epilogue_part<2>(tmem_lane, 0, hsel, ACT_COL, bias_s0 + 10 * 1024, b.acc_free, b.a_ready, b.a_ready + 16, lane);
      arrive_idle(b.acc_free + 8, b.a_ready + 8, lane);
      // ---- RGB (N = 16) -> raw ----
      wait_bar(b.acc_full, nfull[0]++ & 1, 350);
      ptx::tc_fence_after();
      if (hsel == 0) {
        uint32_t r[16];
        ptx::tmem_ld16(tmem_lane + HEAD2_COL, r);
        ptx::tc_wait_ld();
        head[0] = __uint_as_float(r[0]);
        head[1] = __uint_as_float(r[1]);
        head[2] = __uint_as_float(r[2]);
        if (valid) {
          float* o = p.raw + m * p.out_ch;
          if (p.out_ch == 12) {
            reinterpret_cast<float4*>(o)[0] = make_float4(head[0], head[1], head[2], head[3]);
            reinterpret_cast<float4*>(o)[1] = make_float4(head[4], head[5], head[6], head[7]);
            reinterpret_cast<float4*>(o)[2] = make_float4(head[8], head[9], head[10], head[11]);
          } else {
            for (int k = 0; k < p.out_ch; ++k) o[k] = head[k];
          }
        }
        ptx::tc_fence_before();
      }
      // no arrival here: the next tile's prologue arrival doubles as "RGB epilogue done"
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (CL == 2) ptx::cluster_sync();   // no CTA leaves while its peer may still multicast into it / arrive on its barriers
  if (warp == kEpiWarps + 1) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem, 512); }
}

// ---- weight image builder: fp32 [out,in] -> bf16 [K/8][N][8] block images, zero padded -----------
__global__ void tc_pack_kernel(const float* __restrict__ f32, const PackDesc* __restrict__ descs, uint8_t* blob) {
  const PackDesc d = descs[blockIdx.x];
  __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(blob + d.dst_off);
  for (int i = threadIdx.x; i < d.N * d.K; i += blockDim.x) {
    const int n = i / d.K, k = i - n * d.K;
    float v = 0.f;
    if (n < d.rows_valid && k < d.cols_valid) v = f32[d.src_off + (int64_t)(d.row0 + n) * d.src_ld + d.col0 + k];
    dst[(int64_t)(k / 8) * (d.N * 8) + n * 8 + (k % 8)] = __float2bfloat16_rn(v);
  }
  if (d.bias_src >= 0) {  // trailing K = 16 block: [n][0] = hi, [n][1] = lo, zeros elsewhere
    __nv_bfloat16* bd = dst + (int64_t)d.N * d.K;
    for (int i = threadIdx.x; i < d.N * 16; i += blockDim.x) {
      const int n = (i >> 3) % d.N, c = (i & 7) + 8 * (i / (d.N * 8));
      float v = 0.f;
      if (c < 2 && n < d.rows_valid) {
        const float bv = f32[d.bias_src + d.row0 + n];
        const __nv_bfloat16 hi = __float2bfloat16_rn(bv);
        v = c == 0 ? __bfloat162float(hi) : bv - __bfloat162float(hi);
      }
      bd[i] = __float2bfloat16_rn(v);
    }
  }
}

// ---- self test: D[128, N] = A[128, K] * B[N, K]^T through the same layouts / descriptors ---------
// variant 0: A from shared memory (SS);  variant 1: A as bf16 pairs in TMEM (TS), written with
// tcgen05.st exactly like the epilogue writes the activation.
__global__ void __launch_bounds__(128, 1) tc_selftest_kernel(const uint16_t* __restrict__ A, const uint16_t* __restrict__ B,
                                                             float* __restrict__ D, int N, int K, int variant) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t a_s = ptx::smem_u32(smem_raw), b_s = a_s + 128 * K * 2;
  uint16_t* as = reinterpret_cast<uint16_t*>(smem_raw);
  uint16_t* bs = as + 128 * K;
  for (int i = tid; i < 128 * K; i += 128) { const int r = i / K, k = i % K; as[(k / 8) * (128 * 8) + r * 8 + (k % 8)] = A[i]; }
  for (int i = tid; i < N * K; i += 128) { const int r = i / K, k = i % K; bs[(k / 8) * (N * 8) + r * 8 + (k % 8)] = B[i]; }
  if (tid == 0) { ptx::mbar_init(ptx::smem_u32(&bar), 1); ptx::fence_mbar_init(); }
  if (warp == 0) { ptx::tmem_alloc(ptx::smem_u32(&s_tmem), 512); ptx::tmem_relinquish(); }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = s_tmem;
  const uint32_t tmem_lane = tmem + ((uint32_t)(warp * 32) << 16);
  if (variant == 1) {  // row = thread: K/2 packed pairs into columns [ACT_COL, ACT_COL + K/2)
    for (int c = 0; c < K / 2; c += 16) {
      uint32_t r[16];
      for (int j = 0; j < 16; ++j) r[j] = (uint32_t)A[tid * K + 2 * (c + j)] | ((uint32_t)A[tid * K + 2 * (c + j) + 1] << 16);
      ptx::tmem_st16(tmem_lane + ACT_COL + c, r);
    }
    ptx::tc_wait_st();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
  }
  if (tid == 0) {
    const uint32_t idesc = ptx::idesc_bf16(N);
    for (int k = 0; k < K / 16; ++k) {
      const uint64_t bd = ptx::smem_desc(b_s + k * 2 * N * 16, N * 16, 128);
      if (variant == 0) ptx::mma_bf16_ss(tmem, ptx::smem_desc(a_s + k * 2 * 128 * 16, 128 * 16, 128), bd, idesc, k > 0);
      else ptx::mma_bf16_ts(tmem, tmem + ACT_COL + 8 * k, bd, idesc, k > 0);
    }
    ptx::mma_commit(ptx::smem_u32(&bar));
  }
  wait_bar(ptx::smem_u32(&bar), 0, 900);
  ptx::tc_fence_after();
  for (int c = 0; c < N; c += 16) {
    uint32_t r[16];
    ptx::tmem_ld16(tmem_lane + c, r);
    ptx::tc_wait_ld();
    for (int j = 0; j < 16; ++j) D[(warp * 32 + lane) * N + c + j] = __uint_as_float(r[j]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem, 512); }
}

// ---- tensor-pipe rate probe: `reps` x 16 back-to-back UMMAs (M = 128, N, K = 16), optionally with the
// other three warps hammering TMEM loads, as the epilogue warps do.  out[0] = cycles from first issue to commit.
template <int N>
__global__ void __launch_bounds__(128, 1) tc_rate_kernel(int reps, int ts, int ld_traffic, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t s_tmem;
  __shared__ volatile int s_stop;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (128 + 256) * 256 * 2 / 16; i += 128) reinterpret_cast<uint4*>(smem_raw)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) { ptx::mbar_init(ptx::smem_u32(&bar), 1); ptx::fence_mbar_init(); s_stop = 0; }
  if (warp == 0) { ptx::tmem_alloc(ptx::smem_u32(&s_tmem), 512); ptx::tmem_relinquish(); }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = s_tmem;
  const uint32_t a_s = ptx::smem_u32(smem_raw), b_s = a_s + 128 * 256 * 2;
  if (warp == 0) {
    // warp-convergent issue with warp-uniform operands, like the MLP kernel's MMA warp
    constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t a_lo = ((a_s >> 4) & 0x3FFF) | kALbo, b_lo = ((b_s >> 4) & 0x3FFF) | ((uint32_t)N << 16);
    const uint32_t act = tmem + ACT_COL;
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      if (ptx::elect_one()) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          if (ts) ptx::mma_bf16_ts(tmem, act + 8 * k, kHi | (b_lo + 2u * N * k), kIdesc, 1u);
          else ptx::mma_bf16_ss(tmem, kHi | (a_lo + ((2u * kChunkBytes) >> 4) * k), kHi | (b_lo + 2u * N * k), kIdesc, 1u);
        }
      }
      __syncwarp();
    }
    if (ptx::elect_one()) ptx::mma_commit(ptx::smem_u32(&bar));
    __syncwarp();
    const long long t1 = clock64();
    while (!ptx::mbar_try_wait(ptx::smem_u32(&bar), 0)) { }
    const long long t2 = clock64();
    if (tid == 0) { out[0] = t2 - t0; out[1] = t1 - t0; s_stop = 1; }
  } else if (ld_traffic) {
    // TMEM read traffic on the gate columns (not touched by the MMAs) from the other three lane quarters
    uint32_t r[32];
    long long n = 0;
    while (!s_stop) {
      ptx::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + GATE_COL, r);
      ptx::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + GATE_COL + 32, r);
      ptx::tc_wait_ld();
      ++n;
    }
    if ((tid & 31) == 0) out[2 + warp] = n * 2 + (r[0] & 1);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem, 512); }
}

// ------------------------------------------------------------------------------------------------
static inline int up(int v, int m) { return (v + m - 1) / m * m; }
constexpr int kTileCounters = 64;

void tc_free(zest_net* net) {
  if (net->tc_blob) cudaFree(net->tc_blob);
  if (net->tc_bias) cudaFree(net->tc_bias);
  if (net->tc_desc_dev) cudaFree(net->tc_desc_dev);
  if (net->tc_counters) cudaFree(net->tc_counters);
  net->tc_desc_dev = nullptr; net->tc_counters = nullptr;
  if (net->tc_plan_host) delete (TcPlanHost*)net->tc_plan_host;
  net->tc_blob = nullptr; net->tc_bias = nullptr; net->tc_plan_host = nullptr;
}

static bool tc_supported(const zest_net* n) {
  return n->width == 256 && n->depth == 8 && n->skip == 4 && (n->in_pts == 63 || n->in_pts == 84) &&
         n->in_feat <= 64 && n->in_views == 27;
}

int tc_pack(zest_net* net, cudaStream_t st) {
  if (!tc_supported(net)) return ZEST_OK;  // fp32 path only; zest_mlp_fwd_tc reports the reason
  const bool overlap = !(getenv("ZEST_TC_OVERLAP") && atoi(getenv("ZEST_TC_OVERLAP")) == 0);
  const bool first = net->tc_plan_host == nullptr;
  TcPlanHost* ph = first ? new TcPlanHost() : (TcPlanHost*)net->tc_plan_host;
  net->tc_plan_host = ph;
  const int W = 256, P = net->in_pts, F = net->in_feat, Cv = net->in_views;
  ph->P = P; ph->Ppad = up(P, 32); ph->F = F; ph->Fpad = up(F, 16); ph->C = (P == 84) ? 4 : 3;
  ph->overlap = overlap ? 1 : 0;
  ph->cluster2 = (getenv("ZEST_TC_CLUSTER") && atoi(getenv("ZEST_TC_CLUSTER")) == 2) ? 1 : 0;
  ph->gate_fp32 = (getenv("ZEST_TC_GATE_FP32") && atoi(getenv("ZEST_TC_GATE_FP32")) != 0) ? 1 : 0;
  const int Ppad = ph->Ppad, Fpad = ph->Fpad;
  // S = [PE | dirPE x 2 (4 + 4 chunks) | ones (2) | feats] + ring
  ph->smem_bytes = (size_t)(Ppad / 8 + 10 + Fpad / 8) * kChunkBytes + (kBiasInMma ? 0 : (size_t)kBiasOps * 1024) + (size_t)kStages * kStageBytes;

  if (first) {   // the stage plan and the pack descriptors depend on the architecture only: built and uploaded once
  std::vector<TcStage> stages;
  std::vector<PackDesc> packs;
  int64_t blob_off = 0;
  // one ring stage = rows [row0, row0 + N) x columns [col0, col0 + kpad) of a weight matrix (+ bias block)
  auto stage = [&](int N, int rows_valid, int64_t w_off, int w_ld, int row0, int col0, int kpad, int cols_valid, int64_t b_off) {
    TcStage s{(uint32_t)blob_off, (uint32_t)(N * (kpad + (b_off >= 0 ? 16 : 0)) * 2)};
    packs.push_back(PackDesc{w_off, w_ld, row0, rows_valid, col0, cols_valid, N, kpad, blob_off, b_off});
    blob_off += s.bytes;
    stages.push_back(s);
  };
  auto skip = [&]() { stages.push_back(TcStage{0u, 0u}); };
  // revolution 0: GATE part 0, 1 (feats) | L0 part 0, 1 (PE)
  for (int part = 0; part < 2; ++part) stage(128, 128, net->w_gate, F, part * 128, 0, Fpad, F, kBiasInMma ? net->b_gate : -1);
  for (int part = 0; part < 2; ++part) stage(128, 128, net->w_pts[0], P, part * 128, 0, Ppad, P, kBiasInMma ? net->b_pts[0] : -1);
  // L1..L7: (p0, K lo) (p1, K lo) (p0, K hi + bias) (p1, K hi + bias); L5 = [pe | h4] starts with the PE blocks
  for (int l = 1; l < 8; ++l) {
    const int ld = (l == 5) ? P + W : W, c0 = (l == 5) ? P : 0;
    if (l == 5) for (int part = 0; part < 2; ++part) stage(128, 128, net->w_pts[l], ld, part * 128, 0, Ppad, P, -1);
    for (int part = 0; part < 2; ++part) stage(128, 128, net->w_pts[l], ld, part * 128, c0, 128, 128, -1);
    for (int part = 0; part < 2; ++part) stage(128, 128, net->w_pts[l], ld, part * 128, c0 + 128, 128, 128, kBiasInMma ? net->b_pts[l] : -1);
    if (l == 5) { skip(); skip(); }
  }
  // FEAT, with the stacked small heads (N = 16) behind the low-K stages
  for (int part = 0; part < 2; ++part) stage(128, 128, net->w_feat, W, part * 128, 0, 128, 128, -1);
  stage(16, net->n_small, net->w_small, W, 0, 0, W, W, net->b_small);
  for (int part = 0; part < 2; ++part) stage(128, 128, net->w_feat, W, part * 128, 128, 128, 128, kBiasInMma ? net->b_feat : -1);
  // VIEWS: [feature | dirPE] (N = 128); RGB (N = 16) from v
  stage(128, 128, net->w_views, W + Cv, 0, 0, 128, 128, -1);
  stage(128, 128, net->w_views, W + Cv, 0, 128, 128, 128, -1);
  stage(128, 128, net->w_views, W + Cv, 0, W, 32, Cv, kBiasInMma ? net->b_views : -1);
  stage(16, 3, net->w_rgb, W / 2, 0, 0, W / 2, W / 2, net->b_rgb);
  skip(); skip(); skip();

  ZEST_CHECK_ARG((int)stages.size() <= kMaxPlan && stages.size() % kStages == 0, "tc_pack: bad plan (%d stages)", (int)stages.size());
  for (auto& s : stages) ZEST_CHECK_ARG(s.bytes <= (uint32_t)kStageBytes && (s.bytes & 15u) == 0, "tc_pack: stage of %u bytes", s.bytes);
  ph->stages = stages; ph->n_stages = (int)stages.size();

    ZEST_CUDA(cudaMalloc(&net->tc_blob, (size_t)blob_off));
    ZEST_CUDA(cudaMalloc(&net->tc_bias, (size_t)kBiasOps * 256 * sizeof(float)));
    ZEST_CUDA(cudaMalloc(&net->tc_desc_dev, packs.size() * sizeof(PackDesc)));
    ZEST_CUDA(cudaMalloc(&net->tc_counters, kTileCounters * sizeof(int)));
    ZEST_CUDA(cudaMemcpy(net->tc_desc_dev, packs.data(), packs.size() * sizeof(PackDesc), cudaMemcpyHostToDevice));
    net->tc_bytes = blob_off;
    ph->n_packs = (int)packs.size();
  }
  {   // epilogue-side bias table: GATE, L0..L7, FEAT (256 each), VIEWS (128)
    ZEST_CUDA(cudaMemsetAsync(net->tc_bias, 0, (size_t)kBiasOps * 256 * sizeof(float), st));
    auto put = [&](int op, int64_t off, int n) {
      return cudaMemcpyAsync(net->tc_bias + op * 256, net->f32 + off, (size_t)n * sizeof(float), cudaMemcpyDeviceToDevice, st);
    };
    ZEST_CUDA(put(0, net->b_gate, W));
    for (int l = 0; l < 8; ++l) ZEST_CUDA(put(1 + l, net->b_pts[l], W));
    ZEST_CUDA(put(9, net->b_feat, W));
    ZEST_CUDA(put(10, net->b_views, W / 2));
  }
  // refresh of the weight image: stream-ordered, no allocation, no host synchronisation (capturable in a CUDA graph)
  tc_pack_kernel<<<(unsigned)ph->n_packs, 256, 0, st>>>(net->f32, (const PackDesc*)net->tc_desc_dev, (uint8_t*)net->tc_blob);
  ZEST_LAUNCH_CHECK();
  net->tc_dirty = false;
  return ZEST_OK;
}

static unsigned long long* g_timeline = nullptr;

static int tc_launch(const zest_net* net, TcParams& p, cudaStream_t st) {
  p.tl = g_timeline;
  if (!net->packed) { set_error("zest_mlp_fwd_tc: net not packed"); return ZEST_E_STATE; }
  if (tc_supported(net) && (net->tc_dirty || !net->tc_plan_host)) {
    // the bf16 weight image is rebuilt lazily, on the launch stream, the first time an inference launch follows a
    // zest_net_pack(): fine-tuning (which never reads it) pays nothing per optimiser step
    ZEST_TRY(tc_pack(const_cast<zest_net*>(net), st));
  }
  if (!tc_supported(net) || !net->tc_plan_host) {
    set_error("zest_mlp_fwd_tc: the tensor-core path supports width=256 depth=8 skip=4 in_pts in {63,84} in_feat<=64 in_views=27 "
              "(got width=%d depth=%d skip=%d in_pts=%d in_feat=%d in_views=%d); use the fp32 path",
              net->width, net->depth, net->skip, net->in_pts, net->in_feat, net->in_views);
    return ZEST_E_ARG;
  }
  const TcPlanHost* ph = (const TcPlanHost*)net->tc_plan_host;
  memcpy(p.plan, ph->stages.data(), ph->stages.size() * sizeof(TcStage)); p.n_stages = ph->n_stages; p.blob = (const uint8_t*)net->tc_blob; p.bias = net->tc_bias;
  p.P = ph->P; p.Ppad = ph->Ppad; p.F = ph->F; p.Fpad = ph->Fpad;
  p.kind = net->kind; p.out_ch = net->out_ch; p.overlap = ph->overlap;
  p.n_tiles = (p.M + kTile - 1) / kTile;
  if (p.M == 0) return ZEST_OK;
  ZEST_CHECK_ARG(p.n_tiles < (1ll << 30), "zest_mlp_fwd_tc: too many rows for one launch");
  // dynamic tile scheduler: one zeroed counter per launch out of a small ring (launches of one net that are in flight
  // at the same time on different streams must not share a counter); ZEST_TC_STATIC=1 keeps the static round-robin
  static const bool force_static = getenv("ZEST_TC_STATIC") && atoi(getenv("ZEST_TC_STATIC")) != 0;
  p.tile_counter = nullptr;
  if (!force_static && !ph->cluster2 && net->tc_counters) {
    zest_net* mut = const_cast<zest_net*>(net);
    p.tile_counter = net->tc_counters + (mut->tc_counter_next.fetch_add(1, std::memory_order_relaxed) % kTileCounters);
    ZEST_CUDA(cudaMemsetAsync(p.tile_counter, 0, sizeof(int), st));
  }
  int grid = (int)(p.n_tiles < num_sms() ? p.n_tiles : num_sms());
  const bool wide = p.vol && p.V > 6;   // many source views: the in-kernel gather needs four loader warps
  // ZEST_TC_CLUSTER=2 (experiment, default off): CTA pairs share every weight stage through TMA multicast, halving the
  // L2 -> SM weight traffic.  Measured 7 % SLOWER on cfg2 (same-box A/B): the pair runs in lock step on a ring that is
  // one layer deep, so any skew between the two CTAs stalls both; kept for the common variant only.
  const bool pair = ph->cluster2 && p.n_tiles >= 2 && !wide && !ph->gate_fp32;
  if (pair) grid = (grid + 1) & ~1;   // whole clusters; a CTA without tiles of its own still mirrors its peer's stage walk
#define ZEST_TC_GO(CC, G32, LW)                                                                                              \
  do {                                                                                                                       \
    ZEST_CUDA(cudaFuncSetAttribute(mlp_tc_kernel<CC, G32, LW, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ph->smem_bytes)); \
    mlp_tc_kernel<CC, G32, LW, 1><<<grid, 32 * (kEpiWarps + 2 + LW), ph->smem_bytes, st>>>(p);                               \
  } while (0)
#define ZEST_TC_GO_PAIR(CC)                                                                                                  \
  do {                                                                                                                       \
    ZEST_CUDA(cudaFuncSetAttribute(mlp_tc_kernel<CC, false, 2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ph->smem_bytes)); \
    cudaLaunchConfig_t cfg = {};                                                                                             \
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(32 * (kEpiWarps + 2 + 2)); cfg.dynamicSmemBytes = ph->smem_bytes; cfg.stream = st; \
    cudaLaunchAttribute at[1];                                                                                               \
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1; \
    cfg.attrs = at; cfg.numAttrs = 1;                                                                                        \
    ZEST_CUDA(cudaLaunchKernelEx(&cfg, mlp_tc_kernel<CC, false, 2, 2>, p));                                                  \
  } while (0)
  if (pair) {
    if (ph->C == 3) ZEST_TC_GO_PAIR(3); else ZEST_TC_GO_PAIR(4);
  } else if (ph->C == 3) {
    if (ph->gate_fp32) { if (wide) ZEST_TC_GO(3, true, 4); else ZEST_TC_GO(3, true, 2); }
    else { if (wide) ZEST_TC_GO(3, false, 4); else ZEST_TC_GO(3, false, 2); }
  } else {
    if (ph->gate_fp32) { if (wide) ZEST_TC_GO(4, true, 4); else ZEST_TC_GO(4, true, 2); }
    else { if (wide) ZEST_TC_GO(4, false, 4); else ZEST_TC_GO(4, false, 2); }
  }
#undef ZEST_TC_GO
#undef ZEST_TC_GO_PAIR
  ZEST_LAUNCH_CHECK();
  return ZEST_OK;
}

}  // namespace zest

using namespace zest;

extern "C" int zest_mlp_fwd_tc(const zest_net* net, const float* ndc, int ndc_ld, int has_t, float t, const float* feats,
                               int ldf, const float* dirs, int S, int64_t M, float* raw, void* stream) {
  ZEST_CHECK_ARG(net && ndc && feats && dirs && raw && M >= 0 && S > 0 && ndc_ld >= 3, "zest_mlp_fwd_tc: bad arguments");
  ZEST_CHECK_ARG((has_t != 0) == (net->in_pts == 84), "zest_mlp_fwd_tc: has_t does not match the net's input width");
  ZEST_CHECK_ARG(ldf >= net->in_feat, "zest_mlp_fwd_tc: ldf too small");
  TcParams p{};
  p.ndc = ndc; p.ndc_ld = ndc_ld; p.has_t = has_t; p.t = t; p.feats = feats; p.ldf = ldf; p.dirs = dirs; p.S = S;
  p.x = nullptr; p.ldx = 0; p.M = M; p.raw = raw;
  return tc_launch(net, p, (cudaStream_t)stream);
}

extern "C" int zest_gather_mlp_fwd_tc(const zest_net* net, const float* rays_pts, const float* rays_ndc, int ndc_ld, int has_t,
                                      float t, const float* vol_cl, int D, int Hv, int Wv, const float* img_cl, int V, int H,
                                      int W, const float* cams, const float* dirs, int S, int64_t M, float* feats_out, int ldfo,
                                      float* raw, void* stream) {
  ZEST_CHECK_ARG(net && rays_pts && rays_ndc && vol_cl && img_cl && cams && dirs && raw && M >= 0 && S > 0 && ndc_ld >= 3,
                 "zest_gather_mlp_fwd_tc: bad arguments");
  ZEST_CHECK_ARG((has_t != 0) == (net->in_pts == 84), "zest_gather_mlp_fwd_tc: has_t does not match the net's input width");
  ZEST_CHECK_ARG(V > 0 && V <= kMaxViews && net->in_feat == 8 + 4 * V, "zest_gather_mlp_fwd_tc: net expects %d feature columns, 8 + 4 x %d views given",
                 net->in_feat, V);
  ZEST_CHECK_ARG(D > 0 && Hv > 0 && Wv > 0 && H > 0 && W > 0, "zest_gather_mlp_fwd_tc: bad volume / image sizes");
  ZEST_CHECK_ARG(!feats_out || ldfo >= net->in_feat, "zest_gather_mlp_fwd_tc: ldfo too small");
  TcParams p{};
  p.ndc = rays_ndc; p.ndc_ld = ndc_ld; p.has_t = has_t; p.t = t; p.dirs = dirs; p.S = S;
  p.pts = rays_pts; p.vol = vol_cl; p.D = D; p.Hv = Hv; p.Wv = Wv; p.img = img_cl; p.V = V; p.H = H; p.W = W; p.cams = cams;
  p.feats_out = feats_out; p.ldfo = ldfo;
  p.M = M; p.raw = raw;
  return tc_launch(net, p, (cudaStream_t)stream);
}

extern "C" int zest_mlp_fwd_tc_x(const zest_net* net, const float* x, int ldx, int64_t M, float* raw, void* stream) {
  ZEST_CHECK_ARG(net && x && raw && M >= 0, "zest_mlp_fwd_tc_x: bad arguments");
  ZEST_CHECK_ARG(ldx >= net->in_pts + net->in_feat + net->in_views, "zest_mlp_fwd_tc_x: ldx too small");
  TcParams p{};
  p.x = x; p.ldx = ldx; p.M = M; p.raw = raw; p.S = 1;
  return tc_launch(net, p, (cudaStream_t)stream);
}

// debug / characterisation: out[0] = cycles for reps x 16 UMMAs (M = 128, N, K = 16), out[1] = issue cycles
extern "C" int zest_tc_rate_probe(int N, int reps, int ts, int ld_traffic, long long* out, void* stream) {
  ZEST_CHECK_ARG(out && N >= 16 && N <= 256 && (N % 16) == 0 && reps > 0, "zest_tc_rate_probe: bad arguments");
  const size_t smem = (size_t)(128 + 256) * 256 * 2;
#define ZEST_RATE(NN)                                                                                              \
  case NN:                                                                                                         \
    ZEST_CUDA(cudaFuncSetAttribute(tc_rate_kernel<NN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    tc_rate_kernel<NN><<<1, 128, smem, (cudaStream_t)stream>>>(reps, ts, ld_traffic, out);                       \
    break;
  switch (N) {
    ZEST_RATE(16) ZEST_RATE(64) ZEST_RATE(128) ZEST_RATE(256)
    default: set_error("zest_tc_rate_probe: N must be 16, 64, 128 or 256"); return ZEST_E_ARG;
  }
#undef ZEST_RATE
  ZEST_LAUNCH_CHECK();
  return ZEST_OK;
}

// debug: device buffer of 1024 u64 that ZEST_TC_TIMELINE builds fill with (tag << 48 | clock) records
extern "C" int zest_tc_set_timeline(unsigned long long* buf) { g_timeline = buf; return ZEST_OK; }

extern "C" int zest_tc_selftest(const uint16_t* A, const uint16_t* B, float* D, int N, int K, int variant, void* stream) {
  ZEST_CHECK_ARG(A && B && D && N >= 16 && N <= 256 && (N % 16) == 0 && K >= 16 && K <= 256 && (K % 16) == 0,
                 "zest_tc_selftest: N, K must be multiples of 16 in [16, 256]");
  ZEST_CHECK_ARG(variant == 0 || (variant == 1 && (K % 32) == 0), "zest_tc_selftest: variant 1 (A in TMEM) needs K % 32 == 0");
  const size_t smem = (size_t)(128 + N) * K * 2;
  ZEST_CUDA(cudaFuncSetAttribute(tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tc_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, B, D, N, K, variant);
  ZEST_LAUNCH_CHECK();
  return ZEST_OK;
}
