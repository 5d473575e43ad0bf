// bf16 tensor-core radiance MLP for sm_100a: tcgen05.mma with fp32 accumulators in TMEM.
//
// One persistent CTA per SM walks 128-sample tiles (M = 128 = one UMMA tile; one ray when S = 128)
// through the whole network of networks.py:150-221 without the activations ever leaving the SM:
//
//   prologue   PE(ndc[,t]) (networks.py:48-65), gathered feats, PE(dir)  -> bf16 A-operands in smem
//   GATE       g = pts_bias(feat)                 -> bf16 pairs parked in TMEM (reused by 8 layers)
//   L0..L7     h = relu((W h + b) * g), skip [pe | h] into L5      A ping-pongs between two smem tiles
//   FEAT/SMALL feature_linear, [alpha | w | sf | prob] heads (N = 16 MMA)
//   VIEWS/RGB  relu(views([feat | dirpe])) -> rgb_linear (N = 16 MMA) -> raw[M, out_ch] fp32
//
// Warp roles (320 threads): warps 0-7 epilogue (TMEM lane quarter = warp % 4, column half = warp / 4),
// warp 8 = weight producer (cp.async.bulk / UBLKCP into a 3-stage 20 KB ring, weights pre-packed on
// the host side of the ABI in the exact smem image, consumption order), warp 9 = MMA issuer (one
// lane walks a table of pre-built descriptor records in shared memory) and TMEM owner.
// Biases ride in the MMA: the last weight stage of every accumulator group carries one extra K = 16
// block whose first two columns hold the bias split into bf16 hi + lo, multiplied against a constant
// "ones" A chunk, so the epilogue is multiply-by-gate / relu / pack only (packed f32x2 / bf16x2 ops).  Operand layout: K-major, no swizzle: [K/8][rows][8 x bf16], so an
// epilogue thread (= one row) writes 16-byte chunks that are contiguous across the warp
// (conflict-free) and the UMMA descriptor is LBO = rows*16 B (K direction), SBO = 128 B.
//
// Synchronisation is mbarrier-only after start-up: ring full/empty, acc_full[part] (tcgen05.commit),
// acc_free[part] and a_ready[part] (epilogue -> MMA).  Each accumulator half ("part") is signalled
// separately so the next layer's first K half can start while the second half is still in its
// epilogue (ZEST_TC_OVERLAP=1, the default); ZEST_TC_OVERLAP=0 serialises layer by layer.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "net.cuh"
#include "tc_ptx.cuh"

namespace zest {

constexpr int kTile = 128;
constexpr int kStageBytes = 20480;       // 16 KB of weights (N = 128 x K = 64) + 4 KB bias block
constexpr int kStages = 3;
constexpr int kABytes = 65536;           // one 128 x 256 bf16 activation tile
constexpr int kChunkBytes = kTile * 16;  // one 8-column k-chunk of a 128-row tile
constexpr int kMaxPlan = 96;
constexpr int kEpiWarps = 8;
constexpr int kThreads = 32 * (kEpiWarps + 2);
// ops per tile: GATE, L0..L7, FEAT, SMALL, VIEWS, RGB (13)
#define ACC_COL_OF(part) ((uint32_t)(part) * 128u)
constexpr uint32_t GATE_COL = 256, HEAD_COL = 384, HEAD2_COL = 400;

enum : uint8_t { ST_FIRST = 1, ST_LAST = 2, ST_PART1 = 4, ST_OPSTART = 8, ST_OPEND = 16, ST_BIAS = 32 };
enum : uint8_t { NEED_A0 = 1, NEED_A1 = 2, NEED_ACC = 4 };

struct TcStage {   // 16 bytes, one ring slot's worth of weights and the MMAs that consume it
  uint32_t src_off;  // byte offset in the packed blob
  uint32_t bytes;
  uint16_t d_col;    // TMEM column of the accumulator
  uint8_t a_buf;     // 0 = A0, 1 = A1, 2 = S
  uint8_t a_chunk;   // first k-chunk (8 columns) of the A operand in that buffer
  uint8_t n_k16;     // K = 16 steps
  uint8_t n_div8;    // N / 8
  uint8_t flags;
  uint8_t need;      // barrier completions the MMA warp must observe before this stage (de-duplicated per op):
                     // bit0 a_ready[0], bit1 a_ready[1], bit2 acc_free[0], bit3 acc_free[1]
};

struct PackDesc {  // how to build one stage image from the fp32 blob
  int64_t src_off; int src_ld; int row0; int rows_valid; int col0; int cols_valid; int N; int K; int64_t dst_off;
  int64_t bias_src;  // >= 0: append a K = 16 block [n][0] = bf16 hi, [n][1] = bf16 lo of bias[row0 + n]
};

struct TcPlanHost {
  std::vector<TcStage> stages;
  int n_stages = 0;
  int P, Ppad, F, Fpad, C, nf_pts, s_chunks, overlap;
  size_t smem_bytes;
};

struct TcParams {
  TcStage plan[kMaxPlan]; int n_stages;  // by value: lives in the constant bank -> uniform loads in the issue loop
  const uint8_t* blob;
  // inputs (fused mode) or x (x mode)
  const float* ndc; int ndc_ld; int has_t; float t;
  const float* feats; int ldf;
  const float* dirs; int S;
  const float* x; int ldx;
  int P, Ppad, F, Fpad, Cv, nf_pts, nf_dir;
  int kind, out_ch, overlap;
  int64_t M; int64_t n_tiles;
  float* raw;
};

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void wait_bar(uint32_t bar, uint32_t parity, int tag) {
  if (ptx::mbar_try_wait(bar, parity)) return;
  uint32_t spins = 0;
  while (!ptx::mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 24)) {  // a protocol bug must not hang the GPU (each try_wait suspends for a while)
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

// write `n` fp32 values (zero padded to a multiple of 8) of one row as bf16 k-chunks
template <int NPAD>
__device__ __forceinline__ void store_row_chunks(uint32_t buf, int chunk0, int row, const float* v) {
#pragma unroll
  for (int c = 0; c < NPAD / 8; ++c)
    ptx::st_smem_v4(buf + (chunk0 + c) * kChunkBytes + row * 16, ptx::pack_bf16(v[8 * c], v[8 * c + 1]),
                    ptx::pack_bf16(v[8 * c + 2], v[8 * c + 3]), ptx::pack_bf16(v[8 * c + 4], v[8 * c + 5]),
                    ptx::pack_bf16(v[8 * c + 6], v[8 * c + 7]));
}

struct Smem {
  uint32_t a[2];      // activation tiles A0, A1
  uint32_t s;         // PE | dirPE | ones tile
  uint32_t ring;      // weight ring
  uint32_t full, empty, acc_full, acc_free, a_ready;  // barrier arrays
};

// ---- epilogue of one accumulator part of a hidden-type op (bias already inside the accumulator) ----
// MODE 0: acc * gate, relu -> bf16 A tile      (L0..L7)
// MODE 1: acc -> bf16 A tile                    (FEAT)
// MODE 2: acc, relu -> bf16 A tile              (VIEWS)
// MODE 3: acc -> bf16 pairs into the TMEM gate  (GATE)
template <int MODE>
__device__ __forceinline__ void epilogue_part(uint32_t tmem, int part, int q, int hsel, int row, uint32_t out_buf,
                                              uint32_t bar_free, uint32_t bar_ready, int lane) {
  const int col0 = part * 128 + hsel * 64;  // first output column of this thread
  const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
  uint32_t acc[2][32];
  ptx::tmem_ld32(tmem + lane_addr + ACC_COL_OF(part) + hsel * 64, acc[0]);
  ptx::tmem_ld32(tmem + lane_addr + ACC_COL_OF(part) + hsel * 64 + 32, acc[1]);
  uint32_t g[2][16];
  if (MODE == 0) {
    ptx::tmem_ld16(tmem + lane_addr + GATE_COL + col0 / 2, g[0]);
    ptx::tmem_ld16(tmem + lane_addr + GATE_COL + col0 / 2 + 16, g[1]);
  }
  ptx::tc_wait_ld();
  // the accumulator half is drained: the MMA warp may overwrite it
  ptx::tc_fence_before();
  __syncwarp();
  if (lane == 0) ptx::mbar_arrive(bar_free);
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    uint32_t packed[16];
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      float v0 = __uint_as_float(acc[h][j]), v1 = __uint_as_float(acc[h][j + 1]);
      if (MODE == 0) {
        const uint32_t g01 = g[h][j / 2];
        ptx::mul_f32x2(v0, v1, ptx::bf16_lo(g01), ptx::bf16_hi(g01));
      }
      uint32_t pk = ptx::pack_bf16(v0, v1);
      if (MODE == 0 || MODE == 2) pk = ptx::relu_bf16x2(pk);
      packed[j / 2] = pk;
    }
    if (MODE == 3) {
      ptx::tmem_st16(tmem + lane_addr + GATE_COL + col0 / 2 + h * 16, packed);
    } else {
      const int chunk0 = (col0 + h * 32) / 8;
#pragma unroll
      for (int c = 0; c < 4; ++c)
        ptx::st_smem_v4(out_buf + (chunk0 + c) * kChunkBytes + row * 16, packed[4 * c], packed[4 * c + 1],
                        packed[4 * c + 2], packed[4 * c + 3]);
    }
  }
  if (MODE == 3) { ptx::tc_wait_st(); ptx::tc_fence_before(); }
  else ptx::fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0) ptx::mbar_arrive(bar_ready);
}

__device__ __forceinline__ void arrive_idle(uint32_t bar_free, uint32_t bar_ready, int lane) {
  __syncwarp();
  if (lane == 0) { ptx::mbar_arrive(bar_free); ptx::mbar_arrive(bar_ready); }
}

struct MmaCtx {
  Smem sm;
  uint32_t slot, phase, par, n_issued;
  uint32_t ones_lo, ring_lo;
  __device__ __forceinline__ void next_op() { par ^= 1u; }
  __device__ __forceinline__ void wait(uint32_t need) {
    if (need & 1u) wait_bar(sm.a_ready, par, 200);
    if (need & 2u) wait_bar(sm.a_ready + 8, par, 201);
    if (need & 4u) wait_bar(sm.acc_free, par, 202);
    if (need & 8u) wait_bar(sm.acc_free + 8, par, 203);
  }
};

// one ring stage: n_k16 K = 16 steps of N weight rows against the A operand starting at a_lo;
// `last` = the accumulator group's last stage: + bias step against the ones chunk, + acc_full commit
template <int N>
__device__ __forceinline__ void mma_stage(MmaCtx& c, uint32_t a_lo, int n_k16, uint32_t d_tmem, bool first, bool last, int part) {
  constexpr uint32_t kDescHi = (128u >> 4) | (1u << 14);           // SBO = 128 B, descriptor version 1
  constexpr uint64_t kHi = (uint64_t)kDescHi << 32;
  constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
  wait_bar(c.sm.full + 8 * c.slot, c.phase, 220);
  ptx::tc_fence_after();
  // descriptor low words; one K = 16 step advances A by 2 chunks (4096 B) and B by 2 * N * 16 B
  uint32_t b_lo = (c.ring_lo + c.slot * (kStageBytes >> 4)) | ((uint32_t)N << 16);
  if (ptx::elect_one()) {
    uint32_t acc = first ? 0u : 1u;
    for (int k = 0; k < n_k16; ++k) {
      ptx::mma_bf16_ss(d_tmem, kHi | a_lo, kHi | b_lo, kIdesc, acc);
      acc = 1u;
      a_lo += (2u * kChunkBytes) >> 4;
      b_lo += 2u * N;
    }
    if (last) ptx::mma_bf16_ss(d_tmem, kHi | c.ones_lo, kHi | b_lo, kIdesc, 1u);
    ptx::mma_commit(c.sm.empty + 8 * c.slot);
    if (last) ptx::mma_commit(c.sm.acc_full + 8 * part);
  }
  __syncwarp();
  ++c.n_issued;
  if (++c.slot == kStages) { c.slot = 0; c.phase ^= 1u; }
}

// one K segment of an accumulator group, cut into ring stages exactly like tc_pack()'s add_group
template <int N>
__device__ __forceinline__ void mma_seg(MmaCtx& c, int part, uint32_t d_tmem, uint32_t a_lo, int kpad, bool first, bool last,
                                        uint32_t need_k0, uint32_t need_k128) {
  constexpr int kstep = (N == 16) ? 256 : 64;
  for (int k0 = 0; k0 < kpad; k0 += kstep) {
    const int kk = (kpad - k0 < kstep) ? kpad - k0 : kstep;
    if (k0 == 0 && need_k0) c.wait(need_k0);
    if (k0 == 128 && need_k128) c.wait(need_k128);
    mma_stage<N>(c, a_lo + (uint32_t)(k0 / 8) * (kChunkBytes >> 4), kk / 16, d_tmem, first && k0 == 0, last && k0 + kk >= kpad, part);
  }
}

template <int C>  // C = 3 (static: xyz) or 4 (dynamic: xyz + t)
__global__ void __launch_bounds__(kThreads, 1) mlp_tc_kernel(const __grid_constant__ TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t s_bars[2 * kStages + 6];
  __shared__ uint32_t s_tmem;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t base = ptx::smem_u32(smem_raw);
  Smem sm;
  sm.a[0] = base; sm.a[1] = base + kABytes; sm.s = base + 2 * kABytes;
  const int ones_chunk = p.Ppad / 8 + 4;     // S = [PE (Ppad/8 chunks) | dirPE (4) | ones (2)]
  const int s_chunks = ones_chunk + 2;
  sm.ring = sm.s + s_chunks * kChunkBytes;
  const uint32_t bars = ptx::smem_u32(s_bars);
  sm.full = bars; sm.empty = bars + 8 * kStages; sm.acc_full = bars + 16 * kStages;
  sm.acc_free = sm.acc_full + 16; sm.a_ready = sm.acc_free + 16;
  constexpr uint32_t kALo = ((uint32_t)kChunkBytes >> 4) << 16;    // LBO(A) = 128 rows * 16 B

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) { ptx::mbar_init(sm.full + 8 * s, 1); ptx::mbar_init(sm.empty + 8 * s, 1); }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(sm.acc_full + 8 * b, 1);
      ptx::mbar_init(sm.acc_free + 8 * b, kEpiWarps);
      ptx::mbar_init(sm.a_ready + 8 * b, kEpiWarps);
    }
    ptx::fence_mbar_init();
  }
  if (tid < kTile) {  // the constant "ones" A chunk pair: columns 0, 1 = 1.0 (bias hi, lo), the rest 0
    ptx::st_smem_v4(sm.s + ones_chunk * kChunkBytes + tid * 16, 0x3F803F80u, 0u, 0u, 0u);
    ptx::st_smem_v4(sm.s + (ones_chunk + 1) * kChunkBytes + tid * 16, 0u, 0u, 0u, 0u);
    ptx::fence_proxy_async_smem();
  }
  if (warp == kEpiWarps + 1) { ptx::tmem_alloc(ptx::smem_u32(&s_tmem), 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = s_tmem;
  const int64_t my_tiles = (p.n_tiles > blockIdx.x) ? (p.n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (warp == kEpiWarps) {
    // ===================== weight producer =====================
    // The whole warp runs the loop convergently (addresses stay in uniform registers); one elected
    // lane arms the barrier and issues the bulk copy.
    uint32_t slot = 0, phase = 0;
    for (int64_t it = 0; it < my_tiles; ++it) {
      for (int s = 0; s < p.n_stages; ++s) {
        const uint32_t bytes = p.plan[s].bytes, src_off = p.plan[s].src_off;
        wait_bar(sm.empty + 8 * slot, phase ^ 1, 100 + slot);
        if (ptx::elect_one()) {
          ptx::mbar_arrive_expect_tx(sm.full + 8 * slot, bytes);
          ptx::bulk_g2s(sm.ring + slot * kStageBytes, p.blob + src_off, bytes, sm.full + 8 * slot);
        }
        __syncwarp();
        if (++slot == kStages) { slot = 0; phase ^= 1; }
      }
    }
  } else if (warp == kEpiWarps + 1) {
    // ===================== MMA issuer =====================
    // a_ready[0,1] / acc_free[0,1] complete exactly once per op: completion #(13 it + j) is the
    // epilogue of op j-1 (j = 0: the tile prologue, which also stands for the previous tile's RGB
    // epilogue).  mbarrier parity waits are only sound if the MMA warp observes EVERY completion, in
    // order, and before the next one can happen; every op below therefore waits on all four barriers
    // at least once before its last commit.  The schedule is straight-line code (it must walk the
    // ring stages in exactly the order tc_pack() laid them out; checked per tile against n_stages):
    // the warp stays convergent, every operand is warp-uniform, one elected lane issues.
    MmaCtx c{sm, 0u, 0u, 1u, 0u, (((sm.s + ones_chunk * kChunkBytes) >> 4) & 0x3FFF) | kALo, (sm.ring >> 4) & 0x3FFF};
    const uint32_t s_lo = ((sm.s >> 4) & 0x3FFF) | kALo;
    const uint32_t a_lo[2] = {((sm.a[0] >> 4) & 0x3FFF) | kALo, ((sm.a[1] >> 4) & 0x3FFF) | kALo};
    const uint32_t dir_lo = s_lo + (p.Ppad / 8) * (kChunkBytes >> 4);
    const uint32_t acc_t[2] = {tmem + ACC_COL_OF(0), tmem + ACC_COL_OF(1)};
    const bool ov = p.overlap != 0;
    for (int64_t it = 0; it < my_tiles; ++it) {
      c.n_issued = 0;
      // GATE: feats staged in A1 chunks [0, Fpad/8)
      c.next_op(); c.wait(15u);
      for (int part = 0; part < 2; ++part) mma_seg<128>(c, part, acc_t[part], a_lo[1], p.Fpad, true, true, 0u, 0u);
      // L0: PE in S
      c.next_op(); c.wait(15u);
      for (int part = 0; part < 2; ++part) mma_seg<128>(c, part, acc_t[part], s_lo, p.Ppad, true, true, 0u, 0u);
      // L1..L7 (layer l reads A[(l-1)&1]; L5 = [pe | h4]) and FEAT (l = 8: h7 in A1 -> feature)
      for (int l = 1; l <= 8; ++l) {
        c.next_op();
        if (!ov) c.wait(15u);
        const uint32_t in_lo = a_lo[(l - 1) & 1];
        for (int part = 0; part < 2; ++part) {
          c.wait(4u << part);
          if (l == 5) mma_seg<128>(c, part, acc_t[part], s_lo, p.Ppad, true, false, 0u, 0u);
          mma_seg<128>(c, part, acc_t[part], in_lo, 256, l != 5, true, part == 0 ? 1u : 0u, part == 0 ? 2u : 0u);
        }
      }
      // SMALL heads (N = 16) from h7 (A1)
      c.next_op(); c.wait(15u);
      mma_seg<16>(c, 0, tmem + HEAD_COL, a_lo[1], 256, true, true, 0u, 0u);
      // VIEWS: [feature (A0) | dirPE (S)] -> ACC0 (N = 128, single part)
      c.next_op(); c.wait(15u);
      mma_seg<128>(c, 0, acc_t[0], a_lo[0], 256, true, false, 0u, 0u);
      mma_seg<128>(c, 0, acc_t[0], dir_lo, 32, false, true, 0u, 0u);
      // RGB (N = 16) from v (A1[:, 0:128])
      c.next_op(); c.wait(15u);
      mma_seg<16>(c, 0, tmem + HEAD2_COL, a_lo[1], 128, true, true, 0u, 0u);
      if (c.n_issued != (uint32_t)p.n_stages) {
        if (lane == 0) printf("zest mlp_tc: MMA schedule walked %u stages, plan has %d\n", c.n_issued, p.n_stages);
        __trap();
      }
    }
  } else {
    // ===================== prologue + epilogue warps =====================
    const int q = warp & 3, hsel = warp >> 2;
    const int row = q * 32 + lane;
    uint32_t nfull[2] = {0, 0};
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    for (int64_t it = 0; it < my_tiles; ++it) {
      const int64_t tile = blockIdx.x + it * gridDim.x;
      const int64_t m = tile * kTile + row;
      const bool valid = m < p.M;
      // ---- prologue: warps 0-3 encode the point, warps 4-7 stage feats and the direction ----
      if (hsel == 0) {
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
        if (C == 3) store_row_chunks<64>(sm.s, 0, row, pe);
        else store_row_chunks<96>(sm.s, 0, row, pe);
      } else {
        float f[64];
#pragma unroll
        for (int i = 0; i < 64; ++i) f[i] = 0.f;
        const float* src = p.x ? (p.x + m * p.ldx + p.P) : (p.feats + m * p.ldf);
        if (valid) {
#pragma unroll
          for (int j = 0; j < 64; ++j) if (j < p.F) f[j] = __ldg(src + j);
        }
        if (p.Fpad <= 32) store_row_chunks<32>(sm.a[1], 0, row, f);
        else if (p.Fpad <= 48) store_row_chunks<48>(sm.a[1], 0, row, f);
        else store_row_chunks<64>(sm.a[1], 0, row, f);
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
        store_row_chunks<32>(sm.s, p.Ppad / 8, row, d);
      }
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        ptx::mbar_arrive(sm.acc_free); ptx::mbar_arrive(sm.acc_free + 8);
        ptx::mbar_arrive(sm.a_ready); ptx::mbar_arrive(sm.a_ready + 8);
      }
      // ---- GATE ----
      for (int part = 0; part < 2; ++part) {
        wait_bar(sm.acc_full + 8 * part, nfull[part]++ & 1, 300 + part);
        ptx::tc_fence_after();
        epilogue_part<3>(tmem, part, q, hsel, row, 0, sm.acc_free + 8 * part, sm.a_ready + 8 * part, lane);
      }
      // ---- L0..L7: output ping-pongs A0, A1, ... (L0 reads S, writes A0) ----
      for (int l = 0; l < 8; ++l) {
        const uint32_t out = sm.a[l & 1];
        for (int part = 0; part < 2; ++part) {
          wait_bar(sm.acc_full + 8 * part, nfull[part]++ & 1, 310 + part);
          ptx::tc_fence_after();
          epilogue_part<0>(tmem, part, q, hsel, row, out, sm.acc_free + 8 * part, sm.a_ready + 8 * part, lane);
        }
      }
      // ---- FEAT: h7 (A1) -> feature (A0) ----
      for (int part = 0; part < 2; ++part) {
        wait_bar(sm.acc_full + 8 * part, nfull[part]++ & 1, 320 + part);
        ptx::tc_fence_after();
        epilogue_part<1>(tmem, part, q, hsel, row, sm.a[0], sm.acc_free + 8 * part, sm.a_ready + 8 * part, lane);
      }
      // ---- SMALL heads (N = 16): sigma + blend / scene flow / probs, kept in registers ----
      float head[12];
      wait_bar(sm.acc_full, nfull[0]++ & 1, 330);
      ptx::tc_fence_after();
      if (hsel == 0) {
        uint32_t r[16];
        ptx::tmem_ld16(tmem + lane_addr + HEAD_COL, r);
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
        ptx::tc_fence_before();
      }
      arrive_idle(sm.acc_free, sm.a_ready, lane);
      arrive_idle(sm.acc_free + 8, sm.a_ready + 8, lane);
      // ---- VIEWS: [feature (A0) | dirPE (S)] -> relu -> A1[:, 0:128] ----
      wait_bar(sm.acc_full, nfull[0]++ & 1, 340);
      ptx::tc_fence_after();
      epilogue_part<2>(tmem, 0, q, hsel, row, sm.a[1], sm.acc_free, sm.a_ready, lane);
      arrive_idle(sm.acc_free + 8, sm.a_ready + 8, lane);
      // ---- RGB (N = 16) -> raw ----
      wait_bar(sm.acc_full, nfull[0]++ & 1, 350);
      ptx::tc_fence_after();
      if (hsel == 0) {
        uint32_t r[16];
        ptx::tmem_ld16(tmem + lane_addr + HEAD2_COL, r);
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
  if (warp == kEpiWarps + 1) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem, 512); }
}

// ---- weight image builder: fp32 [out,in] -> bf16 [K/8][N][8] stage images, zero padded ----------
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
        const float b = f32[d.bias_src + d.row0 + n];
        const __nv_bfloat16 hi = __float2bfloat16_rn(b);
        v = c == 0 ? __bfloat162float(hi) : b - __bfloat162float(hi);
      }
      bd[i] = __float2bfloat16_rn(v);
    }
  }
}

// ---- self test: D[128, N] = A[128, K] * B[N, K]^T through the same layouts / descriptors ---------
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
  if (warp == 0) { ptx::tmem_alloc(ptx::smem_u32(&s_tmem), 256); ptx::tmem_relinquish(); }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = s_tmem;
  if (tid == 0) {
    const uint32_t idesc = ptx::idesc_bf16(N);
    for (int k = 0; k < K / 16; ++k) {
      uint64_t ad, bd;
      if (variant == 0) {
        ad = ptx::smem_desc(a_s + k * 2 * 128 * 16, 128 * 16, 128);
        bd = ptx::smem_desc(b_s + k * 2 * N * 16, N * 16, 128);
      } else {  // LBO / SBO swapped (diagnostic)
        ad = ptx::smem_desc(a_s + k * 2 * 128 * 16, 128, 128 * 16);
        bd = ptx::smem_desc(b_s + k * 2 * N * 16, 128, N * 16);
      }
      ptx::mma_bf16_ss(tmem, ad, bd, idesc, k > 0);
    }
    ptx::mma_commit(ptx::smem_u32(&bar));
  }
  wait_bar(ptx::smem_u32(&bar), 0, 900);
  ptx::tc_fence_after();
  for (int c = 0; c < N; c += 16) {
    uint32_t r[16];
    ptx::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c, r);
    ptx::tc_wait_ld();
    for (int j = 0; j < 16; ++j) D[(warp * 32 + lane) * N + c + j] = __uint_as_float(r[j]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem, 256); }
}

// ------------------------------------------------------------------------------------------------
static inline int up(int v, int m) { return (v + m - 1) / m * m; }

void tc_free(zest_net* net) {
  if (net->tc_blob) cudaFree(net->tc_blob);
  if (net->tc_plan_host) {
    TcPlanHost* ph = (TcPlanHost*)net->tc_plan_host;
    delete ph;
  }
  net->tc_blob = nullptr; net->tc_plan_host = nullptr;
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
  ph->P = P; ph->Ppad = up(P, 32); ph->F = F; ph->Fpad = up(F, 16); ph->C = (P == 84) ? 4 : 3; ph->nf_pts = 10;
  ph->overlap = overlap ? 1 : 0;
  ph->s_chunks = ph->Ppad / 8 + 6;   // PE | dirPE (4 chunks) | ones (2 chunks)
  ph->smem_bytes = 2 * kABytes + (size_t)ph->s_chunks * kChunkBytes + (size_t)kStages * kStageBytes;
  const int Ppad = ph->Ppad, Fpad = ph->Fpad, dir_chunk = Ppad / 8;

  std::vector<TcStage> stages;
  std::vector<PackDesc> packs;
  int64_t blob_off = 0;
  struct Seg { int a_buf, a_chunk, kpad, col0, cols_valid; };
  // one accumulator group: rows [row0, row0 + N) of a weight matrix against a list of K segments;
  // the group's last stage also carries the bias block (consumed against the "ones" A chunk)
  auto add_group = [&](int part, int N, int rows_valid, uint32_t d_col, int64_t w_off, int w_ld, int row0, int64_t b_off,
                       std::vector<Seg> segs, bool op_start, uint8_t need_first) {
    bool firststage = true;
    for (size_t si = 0; si < segs.size(); ++si) {
      const Seg& sg = segs[si];
      const int kstep = (N == 16) ? 256 : 64;  // 16 KB of B per stage at N = 128, 8 KB at N = 16
      for (int k0 = 0; k0 < sg.kpad; k0 += kstep) {
        const int kk = (sg.kpad - k0 < kstep) ? sg.kpad - k0 : kstep;
        TcStage s{};
        s.src_off = (uint32_t)blob_off; s.bytes = (uint32_t)(N * kk * 2); s.d_col = (uint16_t)d_col;
        s.a_buf = (uint8_t)sg.a_buf; s.a_chunk = (uint8_t)(sg.a_chunk + k0 / 8); s.n_k16 = (uint8_t)(kk / 16);
        s.n_div8 = (uint8_t)(N / 8);
        s.flags = (uint8_t)((firststage ? ST_FIRST : 0) | (part ? ST_PART1 : 0) | ((firststage && op_start) ? ST_OPSTART : 0));
        s.need = 0;
        if (firststage) s.need = need_first;
        if (overlap && sg.a_buf != 2) {  // A tile written by the previous op's epilogue, part by part
          const int c0 = sg.a_chunk + k0 / 8, c1 = c0 + kk / 8;
          if (c0 < 16) s.need |= NEED_A0;
          if (c1 > 16) s.need |= NEED_A1;
        }
        const int valid = sg.cols_valid - k0;
        packs.push_back(PackDesc{w_off, w_ld, row0, rows_valid, sg.col0 + k0, valid < 0 ? 0 : (valid > kk ? kk : valid), N, kk, blob_off, -1});
        blob_off += s.bytes;
        stages.push_back(s);
        firststage = false;
      }
    }
    stages.back().flags |= ST_LAST | ST_BIAS;
    stages.back().bytes += (uint32_t)(N * 16 * 2);
    packs.back().bias_src = b_off;
    blob_off += N * 16 * 2;
  };
  const uint8_t need_all = NEED_A0 | NEED_A1 | NEED_ACC;
  const uint8_t need_first = overlap ? (uint8_t)NEED_ACC : need_all;
  auto two_part = [&](int64_t w_off, int w_ld, int64_t b_off, std::vector<Seg> segs, bool conservative) {
    for (int part = 0; part < 2; ++part)
      add_group(part, 128, 128, ACC_COL_OF(part), w_off, w_ld, part * 128, b_off, segs, part == 0,
                conservative ? need_all : need_first);
  };
  // GATE: feats staged in A1 chunks [0, Fpad/8)
  two_part(net->w_gate, F, net->b_gate, {Seg{1, 0, Fpad, 0, F}}, true);
  // L0: PE in S
  two_part(net->w_pts[0], P, net->b_pts[0], {Seg{2, 0, Ppad, 0, P}}, true);
  // L1..L7; layer l reads A[(l-1)&1]; L5 = [pe | h4]
  for (int l = 1; l < 8; ++l) {
    const int in_buf = (l - 1) & 1;
    if (l == 5) two_part(net->w_pts[l], P + W, net->b_pts[l], {Seg{2, 0, Ppad, 0, P}, Seg{in_buf, 0, W, P, W}}, false);
    else two_part(net->w_pts[l], W, net->b_pts[l], {Seg{in_buf, 0, W, 0, W}}, false);
  }
  // FEAT: h7 is in A1 (L7 writes A[7&1]); feature -> A0
  two_part(net->w_feat, W, net->b_feat, {Seg{1, 0, W, 0, W}}, false);
  // SMALL heads (N = 16) from h7 (A1)
  add_group(0, 16, net->n_small, HEAD_COL, net->w_small, W, 0, net->b_small, {Seg{1, 0, W, 0, W}}, true, need_all);
  // VIEWS: [feature (A0) | dirPE (S)] -> ACC0 (N = 128, single part)
  add_group(0, 128, 128, ACC_COL_OF(0), net->w_views, W + Cv, 0, net->b_views, {Seg{0, 0, W, 0, W}, Seg{2, dir_chunk, 32, W, Cv}}, true, need_all);
  // RGB (N = 16) from v (A1[:, 0:128])
  add_group(0, 16, 3, HEAD2_COL, net->w_rgb, W / 2, 0, net->b_rgb, {Seg{1, 0, W / 2, 0, W / 2}}, true, need_all);

  for (size_t i = 0; i < stages.size(); ++i)
    if (i + 1 == stages.size() || (stages[i + 1].flags & ST_OPSTART)) stages[i].flags |= ST_OPEND;
  // fold the observation rule (see the MMA warp) into de-duplicated per-stage wait bits
  {
    uint32_t waited = 0;
    for (auto& s : stages) {
      if (s.flags & ST_OPSTART) waited = 0;
      const uint32_t part = (s.flags & ST_PART1) ? 1u : 0u;
      uint32_t need = (s.need & (NEED_A0 | NEED_A1)) | ((s.need & NEED_ACC) ? (4u << part) : 0u);
      if (s.flags & ST_LAST) need |= (1u << part) | (4u << part);
      if (s.flags & ST_OPEND) need |= 15u;
      need &= ~waited;
      waited |= need;
      s.need = (uint8_t)need;
    }
  }
  ZEST_CHECK_ARG((int)stages.size() <= kMaxPlan, "tc_pack: plan too long (%d stages)", (int)stages.size());
  for (auto& s : stages) ZEST_CHECK_ARG(s.bytes <= (uint32_t)kStageBytes && (s.bytes & 15u) == 0, "tc_pack: stage of %u bytes", s.bytes);
  ph->stages = stages; ph->n_stages = (int)stages.size();

  if (first) {
    ZEST_CUDA(cudaMalloc(&net->tc_blob, (size_t)blob_off));
    net->tc_bytes = blob_off;
  }
  // device-side scratch for the descriptors (freed after the pack kernel is enqueued + synced)
  PackDesc* d_packs = nullptr;
  ZEST_CUDA(cudaMalloc(&d_packs, packs.size() * sizeof(PackDesc)));
  ZEST_CUDA(cudaMemcpyAsync(d_packs, packs.data(), packs.size() * sizeof(PackDesc), cudaMemcpyHostToDevice, st));
  tc_pack_kernel<<<(unsigned)packs.size(), 256, 0, st>>>(net->f32, d_packs, (uint8_t*)net->tc_blob);
  ZEST_LAUNCH_CHECK();
  ZEST_CUDA(cudaStreamSynchronize(st));  // host vector / scratch are released below
  cudaFree(d_packs);
  return ZEST_OK;
}

static int tc_launch(const zest_net* net, TcParams& p, cudaStream_t st) {
  if (!net->packed) { set_error("zest_mlp_fwd_tc: net not packed"); return ZEST_E_STATE; }
  if (!tc_supported(net) || !net->tc_plan_host) {
    set_error("zest_mlp_fwd_tc: the tensor-core path supports width=256 depth=8 skip=4 in_pts in {63,84} in_feat<=64 in_views=27 "
              "(got width=%d depth=%d skip=%d in_pts=%d in_feat=%d in_views=%d); use the fp32 path",
              net->width, net->depth, net->skip, net->in_pts, net->in_feat, net->in_views);
    return ZEST_E_ARG;
  }
  const TcPlanHost* ph = (const TcPlanHost*)net->tc_plan_host;
  memcpy(p.plan, ph->stages.data(), ph->stages.size() * sizeof(TcStage)); p.n_stages = ph->n_stages; p.blob = (const uint8_t*)net->tc_blob;
  p.P = ph->P; p.Ppad = ph->Ppad; p.F = ph->F; p.Fpad = ph->Fpad; p.Cv = net->in_views; p.nf_pts = 10; p.nf_dir = 4;
  p.kind = net->kind; p.out_ch = net->out_ch; p.overlap = ph->overlap;
  p.n_tiles = (p.M + kTile - 1) / kTile;
  if (p.M == 0) return ZEST_OK;
  const int grid = (int)(p.n_tiles < num_sms() ? p.n_tiles : num_sms());
  if (ph->C == 3) {
    ZEST_CUDA(cudaFuncSetAttribute(mlp_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ph->smem_bytes));
    mlp_tc_kernel<3><<<grid, kThreads, ph->smem_bytes, st>>>(p);
  } else {
    ZEST_CUDA(cudaFuncSetAttribute(mlp_tc_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ph->smem_bytes));
    mlp_tc_kernel<4><<<grid, kThreads, ph->smem_bytes, st>>>(p);
  }
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

extern "C" int zest_mlp_fwd_tc_x(const zest_net* net, const float* x, int ldx, int64_t M, float* raw, void* stream) {
  ZEST_CHECK_ARG(net && x && raw && M >= 0, "zest_mlp_fwd_tc_x: bad arguments");
  ZEST_CHECK_ARG(ldx >= net->in_pts + net->in_feat + net->in_views, "zest_mlp_fwd_tc_x: ldx too small");
  TcParams p{};
  p.x = x; p.ldx = ldx; p.M = M; p.raw = raw; p.S = 1;
  return tc_launch(net, p, (cudaStream_t)stream);
}

extern "C" int zest_tc_selftest(const uint16_t* A, const uint16_t* B, float* D, int N, int K, int variant, void* stream) {
  ZEST_CHECK_ARG(A && B && D && N >= 16 && N <= 256 && (N % 16) == 0 && K >= 16 && K <= 256 && (K % 16) == 0,
                 "zest_tc_selftest: N, K must be multiples of 16 in [16, 256]");
  const size_t smem = (size_t)(128 + N) * K * 2;
  ZEST_CUDA(cudaFuncSetAttribute(tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tc_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, B, D, N, K, variant);
  ZEST_LAUNCH_CHECK();
  return ZEST_OK;
}
