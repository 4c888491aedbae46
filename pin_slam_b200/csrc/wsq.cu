// K1b, warp-specialised: gather + decoder + d/dq of the split pipeline for weighted_first maps, as ONE persistent CTA
// per SM whose warps have different jobs and only meet through mbarriers (no block-wide barrier after the prologue):
//
//   L  loader warps   TMA bulk copies (cp.async.bulk, completion counted in bytes on an mbarrier) of what search_kernel
//                     wrote for a 32-query block -- neighbour ids, IDW weights, position part of the decoder input and,
//                     when d/dq is wanted, the forward-mode seeds  omega_kj = d w_k / d q_j  and d(position part)/dq --
//                     into a ring of "meta" blocks in shared memory; no thread touches the data
//   G  gather teams   (2 x 4 warps, 120 registers) F/4 lanes per feature row (a 128-byte row = 8 lanes x LDG.128): ONE pass
//                     over the K rows gives the IDW-interpolated feature  xbar = sum_k w_k f_k  AND its three directional
//                     derivatives  T_j = sum_k omega_kj (f_k - f_0)  -> rows of an A tile (canonical K-major, hi / lo TF32)
//   M  MMA warps      (one per epilogue group, the elected lane issues) layer 0 as tcgen05.mma with the A tile from shared
//                     memory, layer 1 with the A operand IN TENSOR MEMORY, chunk by chunk behind the layer-0 epilogue;
//                     layer 0 of the next tile right behind layer 1 of this one (separate accumulator columns)
//   E  epilogue groups (2 x 4 warps, one TMEM lane quadrant per warp): thread = tile row.  tcgen05.ld of the accumulator,
//                     ReLU gate, hi / lo split, tcgen05.st of the next layer's A operand back to TMEM, mbarrier arrive;
//                     last layer: output head in registers, results to global memory.  Never issues an MMA, never meets
//                     another warp at a barrier.
//
// d sdf / d q is computed in FORWARD mode: a tile row is either a value row (decoder input x) or one of the three
// tangent rows (dx / dq_j) of the same query; tangent rows go through the same weights, without bias, and are gated
// by the value row's ReLU pattern.  Rows 4i .. 4i+3 of a tile belong to query i, so the gate is a quad shuffle.  This
// needs no transposed weight copies (52 KB instead of 108 KB of decoder state), no backward MMAs, no second pass over
// the feature rows (the C1 phase of decode_umma_kernel: 6.4 M sectors per 200 k queries) and no input-gradient tile,
// which is what makes room for a ring of A tiles: the gathers of later tiles run under the MMA chain of tile t.
// Without d/dq (mesher, dense RGB-D queries) a tile is 128 value rows.
//
// Replaces model/neural_points.py:598-731 (gathers, IDW, weighted_first), model/decoder.py:61-85,112 and the autograd
// call of utils/tools.py:247-260 for batches of >= PINB200_SPLIT_MIN_QUERIES_WF queries (inference mode).
#include "query_dev.cuh"
#include "umma_common.cuh"

namespace pinb {

constexpr int WS_EG = 2;       // epilogue groups of 4 warps
constexpr int WS_GT = 2;       // gather teams of 4 warps: team t fills the A tiles of the CTA's tiles i = t (mod WS_GT)
constexpr int WS_GW = 4;       // warps per gather team
constexpr int WS_LW = 2;       // loader warps; the other two warps of their 4-warp group issue the MMAs (one per epilogue group)
constexpr int WS_THREADS = (4 * WS_EG + WS_GT * WS_GW + WS_LW + WS_EG) * 32;
static_assert((WS_LW + WS_EG) % 4 == 0, "setmaxnreg works on groups of 4 warps");
// register budget per thread (setmaxnreg, one value per 4-warp group).  The pool is what the CTA was launched with
// (640 threads x 96 registers): 8*32*80 + 8*32*120 + 4*32*80 = 61440
constexpr int WS_REG_E = 80, WS_REG_G = 120, WS_REG_L = 80;
static_assert((4 * WS_EG * WS_REG_E + WS_GT * WS_GW * WS_REG_G + (WS_LW + WS_EG) * WS_REG_L) * 32 <= WS_THREADS * 96, "setmaxnreg pool");
constexpr int WS_A0 = 2;       // A-tile ring slots, one per gather team / epilogue group pair: layer 0 of a tile is issued early,
                               // so its slot is free again while the rest of the chain runs (3 slots measured the same)
constexpr int WS_TCOLS = 256;  // TMEM columns per epilogue group: [0,64) D0, [64,128) A1 hi, [128,192) A1 lo, [192,256) D1

struct WsMeta {  // float offsets inside a meta block of 32 queries, [field][k][lane]
  static constexpr int li = 0;                   // [8][32] neighbour id | REMAP, -1 invalid
  static constexpr int w = li + WT * 8;          // [8][32] IDW weight
  static constexpr int xn = w + WT * 8;          // [3][32] sum_k w_k n_k
  static constexpr int floats_ng = xn + WT * 3;
  static constexpr int om = floats_ng;           // [3][8][32] d w_k / d q_j
  static constexpr int P = om + WT * 24;         // [3 j][3 i][32] d (sum_k w_k n_k)_i / d q_j
  static constexpr int floats_g = P + WT * 9;
};

struct WsLayout {  // byte offsets from the dynamic shared memory base
  int w0_hi, w0_lo, w1_hi, w1_lo, b0, b1, wout, bout;
  int a0, a0_half, a0_stride;  // ring of A tiles: slot s = [a0 + s*stride: hi | + half: lo]
  int meta, meta_stride, n_meta;
  int bars, tmem, total;
};
// mbarrier indices
constexpr int WS_MB_MAX = 16;
constexpr int WSB_A0_FULL = 0, WSB_A0_EMPTY = WSB_A0_FULL + WS_A0, WSB_META_FULL = WSB_A0_EMPTY + WS_A0,
              WSB_META_EMPTY = WSB_META_FULL + WS_MB_MAX, WSB_MMA0 = WSB_META_EMPTY + WS_MB_MAX, WSB_MMA1 = WSB_MMA0 + WS_EG,
              WSB_A1_READY = WSB_MMA1 + WS_EG, WSB_D1_FREE = WSB_A1_READY + 4 * WS_EG, WSB_COUNT = WSB_D1_FREE + WS_EG;

// Optional cycle accounting (pinb200_set_option("ws_profile", 1)): per warp, clock64 deltas of up to 8 phases,
// summed over the tiles of the launch; read back with pinb200_debug_read("ws_profile", ...).
constexpr int WS_PROF_SLOTS = 8;
__device__ unsigned long long g_ws_prof[148 * (WS_THREADS / 32) * WS_PROF_SLOTS];
static int g_ws_profile = 0;

template <bool PROF>
struct WsClock {  // PROF = false: no code at all (64-bit counters in the production kernel spilled to local memory)
  unsigned int acc[WS_PROF_SLOTS];
  unsigned int last;
  __device__ __forceinline__ void start() {
    if (PROF) {
#pragma unroll
      for (int i = 0; i < WS_PROF_SLOTS; ++i) acc[i] = 0u;
      last = (unsigned int)clock();
    }
  }
  __device__ __forceinline__ void lap(int slot) {
    if (PROF) {
      const unsigned int now = (unsigned int)clock();
      acc[slot] += now - last;
      last = now;
    }
  }
  __device__ __forceinline__ void flush(int warp) {
    if (PROF) {
      if ((threadIdx.x & 31) == 0 && blockIdx.x < 148) {
#pragma unroll
        for (int i = 0; i < WS_PROF_SLOTS; ++i) g_ws_prof[(blockIdx.x * (WS_THREADS / 32) + warp) * WS_PROF_SLOTS + i] = acc[i];
      }
    }
  }
};

// mbarrier wait, executed by every lane of a converged warp (ONE warp instruction per attempt).  The suspend-time hint
// parks the warp in hardware until the phase completes: a software poll loop (round-2 first version: one lane
// spinning on try_wait, the rest at __syncwarp) made up 65 % of all executed instructions, stole issue slots and
// instruction-cache bandwidth from the working warps (profiles/r02_wsq_v1: icc hit rate 59 %, no_instruction 3.5 stalls
// per issue).  Bounded: a mis-programmed pipeline must trap, not hang the GPU.
__device__ __forceinline__ void ws_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
#pragma unroll 1
  for (int it = 0; it < 2048 && !done; ++it) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity), "r"(1000000)  // <= 1 ms per attempt
        : "memory");
  }
  if (!done) __trap();
}
// ring waits of the producer roles (not latency critical): back off between attempts instead of re-polling
__device__ __forceinline__ void ws_wait_relaxed(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
#pragma unroll 1
  for (int it = 0; it < (1 << 22) && !done; ++it) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity), "r"(1000000)
        : "memory");
    if (!done) __nanosleep(128);
  }
  if (!done) __trap();
}
__device__ __forceinline__ void ws_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void ws_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void ws_group_bar(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void ws_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void ws_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void ws_mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void ws_tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(
          taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}

// TMEM load split into issue and wait, so that shared-memory loads can be put under its latency; the wait takes the
// destination registers as in/out operands to keep their consumers behind it
__device__ __forceinline__ void ws_tmem_ld16_issue(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32"
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void ws_tmem_ld_wait(uint32_t (&v)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]),
                 "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])::"memory");
}
__device__ __forceinline__ void ws_lds16(const float* src, float (&b)[16]) {
#pragma unroll
  for (int e4 = 0; e4 < 4; ++e4) {
    const float4 t = *reinterpret_cast<const float4*>(src + 4 * e4);
    b[4 * e4] = t.x;
    b[4 * e4 + 1] = t.y;
    b[4 * e4 + 2] = t.z;
    b[4 * e4 + 3] = t.w;
  }
}

static_assert(WsMeta::P - WsMeta::om == Seeds::P - Seeds::om && WsMeta::floats_g - WsMeta::om == Seeds::floats,
              "the seed block of the search launch is copied verbatim into the meta block");

__device__ __forceinline__ bool ws_elect() {  // one lane of the (converged) warp; keeps the code warp-uniform for the compiler
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}
// TMA bulk copy global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void ws_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(um_smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void ws_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

template <int FT, bool GRAD, bool PROF>
__global__ void __launch_bounds__(WS_THREADS, 1) wsq_decode_kernel(const __grid_constant__ QueryParams p, const WsLayout lay) {
  constexpr int H = 64;
  using DM = UmmaDims<FT>;
  constexpr int K0 = DM::K0, F = FT, D = FT + 3;
  using M = RowMap<FT>;
  constexpr int QT = GRAD ? 32 : 128;   // queries per tile
  constexpr int BPT = QT / WT;          // meta blocks per tile
  constexpr int MB = GRAD ? 8 : 16;     // meta ring slots (blocks of 32 queries; a value-only tile has 4)
  constexpr int MSTRIDE = GRAD ? WsMeta::floats_g : WsMeta::floats_ng;
  constexpr int NPASS = WT / M::RPP;    // gather passes per meta block (F = 32: 8 passes of 4 queries)
  extern __shared__ __align__(1024) unsigned char ws_smem[];
  unsigned char* sm = ws_smem;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const pinb200_map_view& m = p.map;
  const int K = p.opts.nn_k, L = p.dec.n_hidden, OC = p.dec.out_dim;
  const float slope = p.dec.leaky_relu ? 0.01f : 0.f;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + lay.bars);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(sm + lay.tmem);
  float* meta = reinterpret_cast<float*>(sm + lay.meta);
  WsClock<PROF> clk;

  // ---- prologue (the only block-wide barrier): TMEM, mbarriers, decoder weights
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(um_smem_u32(s_tmem)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    auto init = [&](int i, int count) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(um_smem_u32(bars + i)), "r"(count));
    };
    for (int s = 0; s < WS_A0; ++s) {
      init(WSB_A0_FULL + s, WS_GW);
      init(WSB_A0_EMPTY + s, 1);
    }
    for (int s = 0; s < WS_MB_MAX; ++s) {
      init(WSB_META_FULL + s, 1);
      init(WSB_META_EMPTY + s, WS_GW);
    }
    for (int g = 0; g < WS_EG; ++g) {
      init(WSB_MMA0 + g, 1);
      init(WSB_MMA1 + g, 1);
      for (int c = 0; c < 4; ++c) init(WSB_A1_READY + 4 * g + c, 128);
      init(WSB_D1_FREE + g, 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  um_stage_weight(p.dec.w[0], D, H, D, H, K0, false, sm + lay.w0_hi, sm + lay.w0_lo);
  if (L > 1) um_stage_weight(p.dec.w[1], H, H, H, H, H, false, sm + lay.w1_hi, sm + lay.w1_lo);
  __syncthreads();
  // the layer-0 bias rides on the MMA: input column D is 1 for value rows (0 for tangent rows), weight column D = b0
  static_assert(D < K0, "a spare (padding) input column carries the bias");
  if (tid < H) {
    const float bv = p.dec.b[0] ? __ldg(p.dec.b[0] + tid) : 0.f;
    const float bh = __uint_as_float(__float_as_uint(bv) & TF32_MASK);
    const int off = (tid >> 3) * ((K0 >> 2) * UM_W_LBO) + (D >> 2) * UM_W_LBO + (tid & 7) * 16 + (D & 3) * 4;
    *reinterpret_cast<float*>(sm + lay.w0_hi + off) = bh;
    *reinterpret_cast<float*>(sm + lay.w0_lo + off) = bv - bh;
  }
  for (int e = tid; e < H; e += WS_THREADS) {
    reinterpret_cast<float*>(sm + lay.b0)[e] = 0.f;  // layer-0 bias: see the weight staging above
    reinterpret_cast<float*>(sm + lay.b1)[e] = (L > 1 && p.dec.b[1]) ? __ldg(p.dec.b[1] + e) : 0.f;
  }
  for (int e = tid; e < 4 * H; e += WS_THREADS) reinterpret_cast<float*>(sm + lay.wout)[e] = e < OC * H ? __ldg(p.dec.w_out + e) : 0.f;
  if (tid < 4) reinterpret_cast<float*>(sm + lay.bout)[tid] = (p.dec.b_out && tid < OC) ? __ldg(p.dec.b_out + tid) : 0.f;
  um_publish_and_sync();
  const uint32_t tmem_base = *s_tmem;
  clk.start();

  const long long n_tiles = (p.n + QT - 1) / QT;
  const int n_blocks = p.n_tiles;  // 32-query stash blocks

  if (warp < 4 * WS_EG) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(WS_REG_E));
    // =====================================================================================================
    // E: epilogue group g owns the tiles i = g, g + 2, ... of this CTA.  Its threads never issue an MMA and never
    // meet at a barrier: they wait for the group's MMA warp through mbarriers (bar0 / bar1: layer 0 / 1 complete)
    // and tell it through two more (a1_ready: A1 written + D0 read, d1_free: last accumulator read).
    // profile slots: 0 wait layer-0 MMAs, 1 layer-0 epilogue, 3 wait layer-1 MMAs, 4 last-layer epilogue, 5 outputs
    // =====================================================================================================
    const int g = warp >> 2, qd = warp & 3;
    const int r = qd * WT + lane;  // tile row == TMEM lane
    const uint32_t tb = tmem_base + g * WS_TCOLS;
    const uint32_t tl = tb + ((uint32_t)(qd * 32) << 16);
    const uint32_t bar0 = um_smem_u32(bars + WSB_MMA0 + g), bar1 = um_smem_u32(bars + WSB_MMA1 + g);
    uint32_t ph0 = 0, ph1 = 0;
    const int t = GRAD ? (lane & 3) : 0;  // row type: 0 value, 1..3 tangent d/dq_{t-1}
    const float bsel = t == 0 ? 1.f : 0.f;
    const int src = lane & ~3;  // the value row of this row's query
    const float* s_b0 = reinterpret_cast<const float*>(sm + lay.b0);
    const float* s_b1 = reinterpret_cast<const float*>(sm + lay.b1);
    const float* s_wout = reinterpret_cast<const float*>(sm + lay.wout);
    const float* s_bout = reinterpret_cast<const float*>(sm + lay.bout);
    const uint32_t tl1 = L > 1 ? tl + 192 : tl;  // accumulator of the last hidden layer
    const uint32_t a1_ready = um_smem_u32(bars + WSB_A1_READY + 4 * g), d1_free = um_smem_u32(bars + WSB_D1_FREE + g);

    // bias (value rows) + ReLU gate of 16 accumulator columns.  The gate of a tangent row is the sign pattern of its
    // query's value row: the 16 sign bits travel in ONE quad-leader shuffle per chunk.
    auto gated16 = [&](const uint32_t (&v)[16], const float (&bb)[16], float (&z)[16]) {
      uint32_t mk = 0u;
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        z[e] = fmaf(bsel, bb[e], __uint_as_float(v[e]));
        mk |= z[e] > 0.f ? (1u << e) : 0u;
      }
      if (GRAD) mk = __shfl_sync(FULL, mk, src);
#pragma unroll
      for (int e = 0; e < 16; ++e) z[e] = ((mk >> e) & 1u) ? z[e] : slope * z[e];
    };

    auto gate16 = [&](const uint32_t (&v)[16], float (&z)[16]) {  // gated16 without a bias
      uint32_t mk = 0u;
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        z[e] = __uint_as_float(v[e]);
        mk |= z[e] > 0.f ? (1u << e) : 0u;
      }
      if (GRAD) mk = __shfl_sync(FULL, mk, src);
#pragma unroll
      for (int e = 0; e < 16; ++e) z[e] = ((mk >> e) & 1u) ? z[e] : slope * z[e];
    };
    // where this row's results go: element (query qi, channel ch) at out_base[qi * out_qstride + ch * out_chstride]
    float* out_base;
    float* out_std = nullptr;
    int out_qstride, out_chstride, out_nch = p.is_color ? OC : 1;
    if (t == 0) {
      out_base = p.is_color ? p.out.color : p.out.sdf;
      out_qstride = p.is_color ? OC : 1;
      out_chstride = 1;
      if (!p.is_color) out_std = p.out.sdf_std;
    } else {
      out_base = p.is_color ? p.out.color_grad : p.out.grad;
      if (out_base) out_base += t - 1;
      out_qstride = p.is_color ? 3 * OC : 3;
      out_chstride = 3;
    }
    for (long long T = blockIdx.x + (long long)g * gridDim.x; T < n_tiles; T += (long long)WS_EG * gridDim.x) {
      ws_wait(bar0, ph0);
      ph0 ^= 1u;
      ws_fence_after();
      clk.lap(0);
      if (L > 1) {
        // ---- layer-0 epilogue: gated activations, hi / lo split -> A1 in TMEM
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t v[16];
          float z[16];
          um_tmem_ld16(tl + 16 * c, v);
          gate16(v, z);  // the bias came through the MMA
          uint32_t lo[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            v[e] = __float_as_uint(z[e]) & TF32_MASK;
            lo[e] = __float_as_uint(z[e] - __uint_as_float(v[e]));
          }
          ws_tmem_st16(tl + 64 + 16 * c, v);
          ws_tmem_st16(tl + 128 + 16 * c, lo);
          // this row's 16 A1 columns are written (and its D0 columns read): the group's MMA warp starts the two
          // k-steps of layer 1 that need them while the later chunks are still in the epilogue
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          ws_fence_before();
          ws_arrive(a1_ready + 8 * c);
        }
        clk.lap(1);
        ws_wait(bar1, ph1);
        ph1 ^= 1u;
        ws_fence_after();
        clk.lap(3);
      }
      // ---- last hidden layer: gated activations, output head(s) in registers
      float o[4] = {0.f, 0.f, 0.f, 0.f};
      {
        const float* bl = L > 1 ? s_b1 : s_b0;
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {  // 32 columns per step: two TMEM loads and their bias rows in flight together
          uint32_t va[16], vb[16];
          float za[16], zb[16];
          ws_tmem_ld16_issue(tl1 + 32 * c, va);
          ws_tmem_ld16_issue(tl1 + 32 * c + 16, vb);
          ws_lds16(bl + 32 * c, za);
          ws_lds16(bl + 32 * c + 16, zb);
          ws_tmem_ld_wait(va);
          ws_tmem_ld_wait(vb);
          gated16(va, za, za);
          gated16(vb, zb, zb);
#pragma unroll
          for (int ch = 0; ch < 4; ++ch)
            if (ch < OC) {
              float wa[16], wb[16];
              ws_lds16(s_wout + ch * H + 32 * c, wa);
              ws_lds16(s_wout + ch * H + 32 * c + 16, wb);
              float oa = 0.f, ob = 0.f;  // two independent chains
#pragma unroll
              for (int e = 0; e < 16; ++e) {
                oa = fmaf(za[e], wa[e], oa);
                ob = fmaf(zb[e], wb[e], ob);
              }
              o[ch] += oa + ob;
            }
        }
      }
      // this row has read its accumulator columns of the last layer
      ws_fence_before();
      ws_arrive(d1_free);
      clk.lap(4);
      // ---- outputs: value rows write the prediction, tangent rows one component of its gradient (destinations
      // resolved once per thread, see out_base)
      const long long qi = T * QT + (GRAD ? (r >> 2) : r);
      if (p.dec.sigmoid_out) {
#pragma unroll
        for (int ch = 0; ch < 4; ++ch)
          if (ch < OC) {
            const float oo = fmaf(bsel, s_bout[ch], o[ch]);
            const float val = 1.f / (1.f + expf(-oo));
            const float dv = val * (1.f - val);
            const float dvq = GRAD ? __shfl_sync(FULL, dv, src) : dv;
            o[ch] = t == 0 ? val : dvq * oo;
          }
      } else {
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) o[ch] = fmaf(bsel, s_bout[ch], o[ch]) * p.dec.out_scale;
      }
      if (qi < p.n) {
        if (out_base) {
          float* dst = out_base + qi * out_qstride;
#pragma unroll
          for (int ch = 0; ch < 4; ++ch)
            if (ch < out_nch) dst[ch * out_chstride] = o[ch];
        }
        if (out_std) out_std[qi] = 0.f;
      }
      clk.lap(5);
    }
  } else if (warp < 4 * WS_EG + WS_GT * WS_GW) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(WS_REG_G));
    // =====================================================================================================
    // G: gather teams -- the 4 warps of a team work on the same tile, each on its share of the queries: per round a
    // lane has GU passes x K feature-row loads (16 x LDG.128 for F = 32) in flight; the two teams (and the ring of A
    // tiles) overlap one team's load latency with the other team's reduction.  (A register double buffer over the
    // passes of ONE warp was tried: it spills inside the loop at 120 registers and lost 5 %.)
    // profile slots: 0 wait free A tile, 1 wait meta block, 2 feature-row loads issued, 3 reduce + A-tile stores,
    //                4 position rows + fence + arrive
    // =====================================================================================================
    const int team = (warp - 4 * WS_EG) / WS_GW, gw = (warp - 4 * WS_EG) % WS_GW;
    const int sub = lane / M::LPR, c4 = lane % M::LPR;
    const float4* __restrict__ f4 = reinterpret_cast<const float4*>(p.feat) + c4;
    constexpr int GU = NPASS / WS_GW >= 2 ? 2 : 1;  // passes in flight per gather warp
    int i = team;
    for (long long T = blockIdx.x + (long long)team * gridDim.x; T < n_tiles; T += (long long)WS_GT * gridDim.x, i += WS_GT) {
      const int slot = i % WS_A0;
      unsigned char* a_hi = sm + lay.a0 + slot * lay.a0_stride;
      unsigned char* a_lo = a_hi + lay.a0_half;
      bool slot_free = false;
#pragma unroll 1
      for (int b = 0; b < BPT; ++b) {
        const int blk = i * BPT + b, ms = blk % MB;
        ws_wait_relaxed(um_smem_u32(bars + WSB_META_FULL + ms), (uint32_t)(blk / MB) & 1u);
        clk.lap(1);
        const float* mt = meta + ms * MSTRIDE;
        const int* m_li = reinterpret_cast<const int*>(mt + WsMeta::li);
        const float* m_w = mt + WsMeta::w;
#pragma unroll 1
        for (int p0 = gw * GU; p0 < NPASS; p0 += WS_GW * GU) {
          float4 fv[GU][KREG];
#pragma unroll
          for (int u = 0; u < GU; ++u) {
            const int ql = (p0 + u) * M::RPP + sub;
#pragma unroll
            for (int k = 0; k < KREG; ++k) {
              fv[u][k] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (k < K) {
                const int lif = m_li[k * WT + ql];
                if (lif >= 0) fv[u][k] = __ldg(f4 + (size_t)(lif & ~REMAP) * M::LPR);
              }
            }
          }
          clk.lap(2);
          if (!slot_free) {  // the loads are in flight while the previous tile in this slot is consumed by layer 0
            ws_wait_relaxed(um_smem_u32(bars + WSB_A0_EMPTY + slot), ((uint32_t)(i / WS_A0) & 1u) ^ 1u);
            slot_free = true;
            clk.lap(0);
          }
#pragma unroll
          for (int u = 0; u < GU; ++u) {
            const int ql = (p0 + u) * M::RPP + sub;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            float4 tq[3];
#pragma unroll
            for (int j = 0; j < 3; ++j) tq[j] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int k = 0; k < KREG; ++k)
              if (k < K) {
                const float w = m_w[k * WT + ql];
                acc.x = fmaf(w, fv[u][k].x, acc.x);
                acc.y = fmaf(w, fv[u][k].y, acc.y);
                acc.z = fmaf(w, fv[u][k].z, acc.z);
                acc.w = fmaf(w, fv[u][k].w, acc.w);
                if (GRAD && k >= 1) {
                  const float4 d = make_float4(fv[u][k].x - fv[u][0].x, fv[u][k].y - fv[u][0].y, fv[u][k].z - fv[u][0].z,
                                               fv[u][k].w - fv[u][0].w);
#pragma unroll
                  for (int j = 0; j < 3; ++j) {
                    const float om = mt[WsMeta::om + (j * KREG + k) * WT + ql];  // 0 for invalid neighbours
                    tq[j].x = fmaf(om, d.x, tq[j].x);
                    tq[j].y = fmaf(om, d.y, tq[j].y);
                    tq[j].z = fmaf(om, d.z, tq[j].z);
                    tq[j].w = fmaf(om, d.w, tq[j].w);
                  }
                }
              }
            const int row = GRAD ? 4 * ql : b * WT + ql;
            float4 hi, lo;
            um_split4(acc, hi, lo);
            int off = um_a_off(row, c4, K0);
            *reinterpret_cast<float4*>(a_hi + off) = hi;
            *reinterpret_cast<float4*>(a_lo + off) = lo;
            if (GRAD) {
#pragma unroll
              for (int j = 0; j < 3; ++j) {
                um_split4(tq[j], hi, lo);
                off = um_a_off(row + 1 + j, c4, K0);
                *reinterpret_cast<float4*>(a_hi + off) = hi;
                *reinterpret_cast<float4*>(a_lo + off) = lo;
              }
            }
          }
          clk.lap(3);
        }
        if (!slot_free) {  // warps without a pass in this block (F = 8) still write position rows
          ws_wait_relaxed(um_smem_u32(bars + WSB_A0_EMPTY + slot), ((uint32_t)(i / WS_A0) & 1u) ^ 1u);
          slot_free = true;
        }
        // position part (columns F .. F+2) and zero padding of the rows: thread per row
        if (GRAD || (b % WS_GW) == gw) {
          const int row = GRAD ? gw * WT + lane : b * WT + lane;
          const int ql = GRAD ? row >> 2 : lane, tt = GRAD ? row & 3 : 0;
          const float* src3 = tt == 0 ? mt + WsMeta::xn : mt + WsMeta::P + (tt - 1) * 3 * WT;
          float4 hi, lo;
          um_split4(make_float4(src3[ql], src3[WT + ql], src3[2 * WT + ql], tt == 0 ? 1.f : 0.f), hi, lo);  // column D: bias input
          int off = um_a_off(row, F / 4, K0);
          *reinterpret_cast<float4*>(a_hi + off) = hi;
          *reinterpret_cast<float4*>(a_lo + off) = lo;
#pragma unroll
          for (int cz = F / 4 + 1; cz < K0 / 4; ++cz) {
            off = um_a_off(row, cz, K0);
            *reinterpret_cast<float4*>(a_hi + off) = make_float4(0.f, 0.f, 0.f, 0.f);
            *reinterpret_cast<float4*>(a_lo + off) = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // A-tile writes -> visible to the tensor core
        __syncwarp();
        if (lane == 0) ws_arrive(um_smem_u32(bars + WSB_META_EMPTY + ms));
        clk.lap(4);
      }
      if (lane == 0) ws_arrive(um_smem_u32(bars + WSB_A0_FULL + slot));
    }
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(WS_REG_L));  // one call site for the whole 4-warp group
    if (warp >= 4 * WS_EG + WS_GT * WS_GW + WS_LW) {
    // =====================================================================================================
    // M: one MMA warp per epilogue group (lane 0 issues).  Software pipeline over the group's tiles: layer 0 of tile
    // t+1 goes right behind layer 1 of tile t (separate accumulator columns), so it runs under the last-layer
    // epilogue of tile t.  With the issue loop on its own warp the epilogue warps are never blocked by a full
    // tensor-pipe queue (the group's first warp was: 3.8k of 11.5k cycles per tile, profiles/r02_k1_wsq_*).
    // profile slots: 0 wait A1 / D0, 1 wait D1 free, 2 layer-1 issue, 3 wait A tile, 4 layer-0 issue
    // =====================================================================================================
    const int g = warp - (4 * WS_EG + WS_GT * WS_GW + WS_LW);
    const uint32_t tb = tmem_base + g * WS_TCOLS;
    const uint32_t bar0 = um_smem_u32(bars + WSB_MMA0 + g), bar1 = um_smem_u32(bars + WSB_MMA1 + g);
    const uint32_t a1_ready = um_smem_u32(bars + WSB_A1_READY + 4 * g), d1_free = um_smem_u32(bars + WSB_D1_FREE + g);
    constexpr uint32_t A_SBO0 = (K0 / 4) * UM_A_LBO, W_SBO0 = (K0 / 4) * UM_W_LBO, W_SBO1 = (H / 4) * UM_W_LBO;
    // a k-step (8 columns = two 16-byte chunks) advances the start-address field of a descriptor by 2 * LBO / 16
    constexpr uint64_t A_STEP = (2 * UM_A_LBO) >> 4, W_STEP = (2 * UM_W_LBO) >> 4;
    const uint64_t w0h_d = um_desc(um_smem_u32(sm + lay.w0_hi), UM_W_LBO, W_SBO0), w0l_d = um_desc(um_smem_u32(sm + lay.w0_lo), UM_W_LBO, W_SBO0);
    const uint64_t w1h_d = um_desc(um_smem_u32(sm + lay.w1_hi), UM_W_LBO, W_SBO1), w1l_d = um_desc(um_smem_u32(sm + lay.w1_lo), UM_W_LBO, W_SBO1);
    const uint32_t idesc = um_idesc(H);
    const uint32_t td1 = L > 1 ? tb + 192 : tb;  // accumulator of the last hidden layer
    auto issue_l0 = [&](int ii) {  // layer 0 of the CTA's ii-th tile
      const int slot = ii % WS_A0;
      ws_wait(um_smem_u32(bars + WSB_A0_FULL + slot), (uint32_t)(ii / WS_A0) & 1u);
      clk.lap(3);
      if (ws_elect()) {
        ws_fence_after();
        const uint32_t a_hi = um_smem_u32(sm + lay.a0 + slot * lay.a0_stride);
        const uint64_t ah = um_desc(a_hi, UM_A_LBO, A_SBO0), al = um_desc(a_hi + lay.a0_half, UM_A_LBO, A_SBO0);
#pragma unroll
        for (int s = 0; s < K0 / 8; ++s) {
          um_mma(tb, al + s * A_STEP, w0h_d + s * W_STEP, idesc, s > 0);  // small terms first
          um_mma(tb, ah + s * A_STEP, w0l_d + s * W_STEP, idesc, 1);
          um_mma(tb, ah + s * A_STEP, w0h_d + s * W_STEP, idesc, 1);
        }
        ws_commit(bar0);
        ws_commit(um_smem_u32(bars + WSB_A0_EMPTY + slot));  // the A tile may be refilled once these MMAs have read it
      }
      __syncwarp();
      clk.lap(4);
    };
    int i = g;
    uint32_t n = 0;  // tiles of this group so far
    long long T = blockIdx.x + (long long)g * gridDim.x;
    if (T < n_tiles) issue_l0(i);
    for (; T < n_tiles; T += (long long)WS_EG * gridDim.x, i += WS_EG, ++n) {
      const bool has_next = T + (long long)WS_EG * gridDim.x < n_tiles;
      if (L > 1) {
        if (n > 0) ws_wait(d1_free, (n - 1u) & 1u);  // every row has read the previous tile's last accumulator
        clk.lap(1);
#pragma unroll
        for (int c = 0; c < 4; ++c) {  // layer 1 follows the layer-0 epilogue chunk by chunk (16 A1 columns = 2 k-steps)
          ws_wait(a1_ready + 8 * c, n & 1u);
          clk.lap(0);
          if (ws_elect()) {
            ws_fence_after();
#pragma unroll
            for (int s = 2 * c; s < 2 * c + 2; ++s) {
              ws_mma_ts(td1, tb + 128 + 8 * s, w1h_d + s * W_STEP, idesc, s > 0);
              ws_mma_ts(td1, tb + 64 + 8 * s, w1l_d + s * W_STEP, idesc, 1);
              ws_mma_ts(td1, tb + 64 + 8 * s, w1h_d + s * W_STEP, idesc, 1);
            }
            if (c == 3) ws_commit(bar1);
          }
          __syncwarp();
          clk.lap(2);
        }
      } else {
        ws_wait(d1_free, n & 1u);  // single hidden layer: the epilogue reads the layer-0 accumulator itself
        clk.lap(1);
      }
      if (has_next) issue_l0(i + WS_EG);
    }
    } else {
    // =====================================================================================================
    // L: loader warps -- TMA bulk copies of the search launch's results into the meta ring: neighbour ids + IDW weights
    // (2 KB), position part (384 B) and, with d/dq, the forward-mode seeds (4.1 KB) of a 32-query block, all counted in
    // bytes on the block's "full" mbarrier.  No thread touches the data.
    // profile slots: 0 wait free meta block, 1 copies issued
    // =====================================================================================================
    const int lw = warp - 4 * WS_EG - WS_GT * WS_GW;
    // the only role that reads what the search launch wrote: with programmatic dependent launch this grid may have
    // started before that one finished
    asm volatile("griddepcontrol.wait;" ::: "memory");
    int i = 0;
    for (long long T = blockIdx.x; T < n_tiles; T += gridDim.x, ++i) {
#pragma unroll 1
      for (int b = 0; b < BPT; ++b) {
        const int blk = i * BPT + b;
        if (blk % WS_LW != lw) continue;
        const int ms = blk % MB;
        ws_wait_relaxed(um_smem_u32(bars + WSB_META_EMPTY + ms), ((uint32_t)(blk / MB) & 1u) ^ 1u);
        clk.lap(0);
        float* mt = meta + ms * MSTRIDE;
        const uint32_t full = um_smem_u32(bars + WSB_META_FULL + ms);
        const long long st = T * BPT + b;  // stash block
        if (st < n_blocks) {
          if (ws_elect()) {
            const float* sb = p.stash + (size_t)st * Stash::floats;
            constexpr uint32_t B_LW = 2 * WT * 8 * 4, B_XN = 3 * WT * 4, B_SD = Seeds::floats * 4;
            ws_arrive_expect_tx(full, B_LW + B_XN + (GRAD ? B_SD : 0u));
            ws_bulk_g2s(mt + WsMeta::li, sb + Stash::li, B_LW, full);  // li | w are adjacent in both layouts
            ws_bulk_g2s(mt + WsMeta::xn, sb + Stash::pos, B_XN, full);
            if (GRAD) ws_bulk_g2s(mt + WsMeta::om, p.seeds + (size_t)st * Seeds::floats, B_SD, full);
          }
        } else {  // tail of the last value-only tile: a block without neighbours
          for (int e = lane; e < MSTRIDE; e += 32) mt[e] = e < WsMeta::w ? __int_as_float(-1) : 0.f;
          __syncwarp();
          if (lane == 0) ws_arrive(full);
        }
        __syncwarp();
        clk.lap(1);
      }
    }
  }
  }
  clk.flush(warp);

  // ---- teardown
  ws_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
template <int FT>
static WsLayout plan_ws_layout(const pinb200_decoder_view& d, bool grad) {
  using DM = UmmaDims<FT>;
  WsLayout l{};
  int o = 0;
  auto take = [&](int bytes) {
    const int at = o;
    o += (bytes + 127) & ~127;
    return at;
  };
  const int w0 = 64 * DM::K0 * 4, w1 = 64 * 64 * 4;
  l.w0_hi = take(w0);
  l.w0_lo = take(w0);
  if (d.n_hidden > 1) {
    l.w1_hi = take(w1);
    l.w1_lo = take(w1);
  }
  l.b0 = take(64 * 4);
  l.b1 = take(64 * 4);
  l.wout = take(4 * 64 * 4);
  l.bout = take(16);
  l.a0_half = ((UM_ROWS / 8) * (DM::K0 / 4) * UM_A_LBO + 127) & ~127;
  l.a0_stride = 2 * l.a0_half;
  l.a0 = take(WS_A0 * l.a0_stride);
  l.meta_stride = (grad ? WsMeta::floats_g : WsMeta::floats_ng) * 4;
  l.n_meta = grad ? 8 : 16;
  l.meta = take(l.n_meta * l.meta_stride);
  l.bars = take(WSB_COUNT * 8);
  l.tmem = take(4);
  l.total = o;
  return l;
}

template <int FT, bool GRAD, bool PROF>
static int launch_wsq(QueryParams& p, cudaStream_t stream) {
  const WsLayout lay = plan_ws_layout<FT>(p.dec, GRAD);
  const size_t smem_bytes = (size_t)lay.total;
  if (smem_bytes > 227 * 1024) {
    set_error("wsq_decode kernel needs %zu B shared memory (> 227 KB)", smem_bytes);
    return PINB200_ERR_UNSUPPORTED;
  }
  auto kern = wsq_decode_kernel<FT, GRAD, PROF>;
  static std::mutex mu;
  static std::vector<int> done;
  int dev = 0;
  cudaGetDevice(&dev);
  {
    std::lock_guard<std::mutex> lk(mu);
    if (std::find(done.begin(), done.end(), dev) == done.end()) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      if (e != cudaSuccess) {
        set_error("cudaFuncSetAttribute(wsq_decode): %s", cudaGetErrorString(e));
        return PINB200_ERR_CUDA;
      }
      done.push_back(dev);
    }
  }
  p.qpt = WT;
  p.n_tiles = (int)((p.n + WT - 1) / WT);
  constexpr int QT = GRAD ? 32 : 128;
  const long long n_tiles = (p.n + QT - 1) / QT;
  const int grid = (int)std::min<long long>(n_tiles, (long long)sm_count());
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(WS_THREADS);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = p.pdl ? 1 : 0;
  const QueryParams pk = p;
  const cudaError_t le = cudaLaunchKernelEx(&cfg, kern, pk, lay);
  if (le != cudaSuccess) {
    set_error("wsq_decode_kernel launch: %s", cudaGetErrorString(le));
    return PINB200_ERR_CUDA;
  }
  return check_launch("wsq_decode_kernel");
}

int dispatch_wsq(QueryParams& p, cudaStream_t stream) {
  const bool grad = p.opts.need_grad != 0;
  switch (p.dec.in_dim - 3) {
    case 8: return grad ? launch_wsq<8, true, false>(p, stream) : launch_wsq<8, false, false>(p, stream);
    case 16: return grad ? launch_wsq<16, true, false>(p, stream) : launch_wsq<16, false, false>(p, stream);
    case 32:
      if (g_ws_profile) return grad ? launch_wsq<32, true, true>(p, stream) : launch_wsq<32, false, true>(p, stream);
      return grad ? launch_wsq<32, true, false>(p, stream) : launch_wsq<32, false, false>(p, stream);
    default: break;
  }
  set_error("wsq_decode: feature_dim %d unsupported", p.dec.in_dim - 3);
  return PINB200_ERR_UNSUPPORTED;
}

void wsq_set_profile(int on) { g_ws_profile = on; }

// copies the cycle counters of the last profiled launch: [148 CTAs][20 warps][8 slots] uint64
int wsq_read_profile(unsigned long long* host_out, int64_t count) {
  const int64_t have = (int64_t)(sizeof(g_ws_prof) / sizeof(unsigned long long));
  if (count < have) {
    set_error("ws_profile: buffer of %lld < %lld counters", (long long)count, (long long)have);
    return PINB200_ERR_BAD_ARG;
  }
  const cudaError_t e = cudaMemcpyFromSymbol(host_out, g_ws_prof, sizeof(g_ws_prof));
  if (e != cudaSuccess) {
    set_error("ws_profile: %s", cudaGetErrorString(e));
    return PINB200_ERR_CUDA;
  }
  return PINB200_OK;
}

}  // namespace pinb
