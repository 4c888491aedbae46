// K1b, warp-specialised: gather + decoder + d/dq of the split pipeline for weighted_first maps, as ONE persistent CTA
// per SM whose warps have different jobs and only meet through mbarriers (no block-wide barrier after the prologue):
//
//   L  loader warps   thread per query: stash block (written by search_kernel) -> "meta" block in shared memory:
//                     neighbour ids, IDW weights, position part of the decoder input and, when d/dq is wanted, the
//                     derivatives of the weights  omega_kj = d w_k / d q_j  and of the position part
//   G  gather warps   F/4 lanes per feature row (a 128-byte row = 8 lanes x LDG.128): ONE pass over the K rows gives the
//                     IDW-interpolated feature  xbar = sum_k w_k f_k  AND its three directional derivatives
//                     T_j = sum_k omega_kj (f_k - f_0)  -> rows of the next A tile (canonical K-major, hi / lo TF32)
//   E  epilogue groups (2 x 4 warps, one TMEM lane quadrant per warp): thread = tile row.  Issue layer 0 as
//                     tcgen05.mma (A from shared memory), read the accumulator with tcgen05.ld, bias / ReLU, write the
//                     activations back to TENSOR MEMORY with tcgen05.st, issue layer 1 with the A operand IN TMEM,
//                     output head in registers, results to global memory.
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
// call of utils/tools.py:247-260 for batches of >= PINB200_SPLIT_MIN_QUERIES queries.
#include "query_dev.cuh"
#include "umma_common.cuh"

namespace pinb {

constexpr int WS_EG = 2;       // epilogue groups of 4 warps
constexpr int WS_GT = 2;       // gather teams of 4 warps: team t fills the A tiles of the CTA's tiles i = t (mod WS_GT)
constexpr int WS_GW = 4;       // warps per gather team
constexpr int WS_LW = 4;       // loader warps
constexpr int WS_THREADS = (4 * WS_EG + WS_GT * WS_GW + WS_LW) * 32;
// register budget per thread (setmaxnreg, one value per 4-warp group; 8*32*88 + 8*32*128 + 4*32*80 = 65536)
constexpr int WS_REG_E = 88, WS_REG_G = 128, WS_REG_L = 80;
constexpr int WS_A0 = 3;       // A-tile ring slots
constexpr int WS_TCOLS = 192;  // TMEM columns per epilogue group: [0,64) D, [64,128) A1 hi, [128,192) A1 lo

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
constexpr int WSB_A0_FULL = 0, WSB_A0_EMPTY = WSB_A0_FULL + WS_A0, WSB_META_FULL = WSB_A0_EMPTY + WS_A0,
              WSB_META_EMPTY = WSB_META_FULL + 8, WSB_MMA = WSB_META_EMPTY + 8, WSB_COUNT = WSB_MMA + WS_EG;

// Optional cycle accounting (pinb200_set_option("ws_profile", 1)): per warp, clock64 deltas of up to 8 phases,
// summed over the tiles of the launch; read back with pinb200_debug_read("ws_profile", ...).
constexpr int WS_PROF_SLOTS = 8;
__device__ unsigned long long g_ws_prof[148 * (WS_THREADS / 32) * WS_PROF_SLOTS];
static int g_ws_profile = 0;

struct WsClock {
  unsigned long long acc[WS_PROF_SLOTS];
  long long last;
  bool on;
  __device__ __forceinline__ void start(bool enable) {
    on = enable;
#pragma unroll
    for (int i = 0; i < WS_PROF_SLOTS; ++i) acc[i] = 0ull;
    last = on ? clock64() : 0;
  }
  __device__ __forceinline__ void lap(int slot) {
    if (on) {
      const long long now = clock64();
      acc[slot] += (unsigned long long)(now - last);
      last = now;
    }
  }
  __device__ __forceinline__ void flush(int warp) {
    if (on && (threadIdx.x & 31) == 0 && blockIdx.x < 148) {
#pragma unroll
      for (int i = 0; i < WS_PROF_SLOTS; ++i) g_ws_prof[(blockIdx.x * (WS_THREADS / 32) + warp) * WS_PROF_SLOTS + i] = acc[i];
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

// Meta block column of one query, from the search results (registers).  GRAD adds the forward-mode seeds:
//   w_k = u_k / sum u,  u_k = 1 / (d_k^2 + eps)  =>  d w_k / d q = w_k (c_k d_k - sum_m w_m c_m d_m),  c_k = -2 u_k,
//   d_k = q - p_k (the point the distance was measured to);  x_n = sum_k w_k n_k  =>
//   d x_n / d q_j = sum_k omega_kj (n_k - n_0) + sum_k w_k R_k e_j   (sum_k omega_kj = 0: the shift by the nearest
//   neighbour keeps the sum cancellation-free when neighbours coincide; R_k = I before loop closure)
template <bool GRAD>
__device__ __forceinline__ void ws_write_meta(const pinb200_map_view& m, int K, const int (&lif)[KREG], const float (&w)[KREG],
                                              const float (&dx)[KREG], const float (&dy)[KREG], const float (&dz)[KREG], float qx,
                                              float qy, float qz, float usum, int nn, float px, float py, float pz, float* mt,
                                              int lane) {
  int* m_li = reinterpret_cast<int*>(mt + WsMeta::li);
#pragma unroll
  for (int k = 0; k < KREG; ++k) {
    m_li[k * WT + lane] = k < K ? lif[k] : -1;
    mt[WsMeta::w + k * WT + lane] = k < K ? w[k] : 0.f;
  }
  mt[WsMeta::xn + 0 * WT + lane] = px;
  mt[WsMeta::xn + 1 * WT + lane] = py;
  mt[WsMeta::xn + 2 * WT + lane] = pz;
  if (GRAD) {
    float c[KREG], S0 = 0.f, S1 = 0.f, S2 = 0.f;
#pragma unroll
    for (int k = 0; k < KREG; ++k) {
      const bool v = k < K && lif[k] >= 0;
      c[k] = (v && nn > 0) ? -2.f * (w[k] * usum) : 0.f;
      const float wc = v ? w[k] * c[k] : 0.f;
      S0 = fmaf(wc, dx[k], S0);
      S1 = fmaf(wc, dy[k], S1);
      S2 = fmaf(wc, dz[k], S2);
    }
    float P[3][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
    float n0x = 0.f, n0y = 0.f, n0z = 0.f;
#pragma unroll
    for (int k = 0; k < KREG; ++k) {
      const bool v = k < K && lif[k] >= 0;
      float o0 = 0.f, o1 = 0.f, o2 = 0.f;
      if (v) {
        o0 = w[k] * (c[k] * dx[k] - S0);
        o1 = w[k] * (c[k] * dy[k] - S1);
        o2 = w[k] * (c[k] * dz[k] - S2);
        float nx, ny, nz;
        float4 quat;
        neighbour_vec(m, lif[k], dx[k], dy[k], dz[k], qx, qy, qz, nx, ny, nz, quat);
        if (k == 0) {
          n0x = nx;
          n0y = ny;
          n0z = nz;
        }
        const float ex = nx - n0x, ey = ny - n0y, ez = nz - n0z;
        P[0][0] = fmaf(o0, ex, P[0][0]);
        P[0][1] = fmaf(o0, ey, P[0][1]);
        P[0][2] = fmaf(o0, ez, P[0][2]);
        P[1][0] = fmaf(o1, ex, P[1][0]);
        P[1][1] = fmaf(o1, ey, P[1][1]);
        P[1][2] = fmaf(o1, ez, P[1][2]);
        P[2][0] = fmaf(o2, ex, P[2][0]);
        P[2][1] = fmaf(o2, ey, P[2][1]);
        P[2][2] = fmaf(o2, ez, P[2][2]);
        if (m.after_pgo) {  // column j of the point's (passive) rotation
          float r0, r1, r2;
          quat_rotate_passive(quat.x, quat.y, quat.z, quat.w, 1.f, 0.f, 0.f, r0, r1, r2);
          P[0][0] = fmaf(w[k], r0, P[0][0]);
          P[0][1] = fmaf(w[k], r1, P[0][1]);
          P[0][2] = fmaf(w[k], r2, P[0][2]);
          quat_rotate_passive(quat.x, quat.y, quat.z, quat.w, 0.f, 1.f, 0.f, r0, r1, r2);
          P[1][0] = fmaf(w[k], r0, P[1][0]);
          P[1][1] = fmaf(w[k], r1, P[1][1]);
          P[1][2] = fmaf(w[k], r2, P[1][2]);
          quat_rotate_passive(quat.x, quat.y, quat.z, quat.w, 0.f, 0.f, 1.f, r0, r1, r2);
          P[2][0] = fmaf(w[k], r0, P[2][0]);
          P[2][1] = fmaf(w[k], r1, P[2][1]);
          P[2][2] = fmaf(w[k], r2, P[2][2]);
        } else {
          P[0][0] += w[k];
          P[1][1] += w[k];
          P[2][2] += w[k];
        }
      }
      mt[WsMeta::om + (0 * KREG + k) * WT + lane] = o0;
      mt[WsMeta::om + (1 * KREG + k) * WT + lane] = o1;
      mt[WsMeta::om + (2 * KREG + k) * WT + lane] = o2;
    }
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
      for (int i = 0; i < 3; ++i) mt[WsMeta::P + (j * 3 + i) * WT + lane] = P[j][i];
  }
}

template <int FT, bool GRAD>
__global__ void __launch_bounds__(WS_THREADS, 1) wsq_decode_kernel(const __grid_constant__ QueryParams p, const WsLayout lay,
                                                                   const int profile) {
  constexpr int H = 64;
  using DM = UmmaDims<FT>;
  constexpr int K0 = DM::K0, F = FT, D = FT + 3;
  using M = RowMap<FT>;
  constexpr int QT = GRAD ? 32 : 128;   // queries per tile
  constexpr int BPT = QT / WT;          // meta blocks per tile
  constexpr int MB = GRAD ? 4 : 8;      // meta ring slots
  constexpr int MSTRIDE = GRAD ? WsMeta::floats_g : WsMeta::floats_ng;
  constexpr int NPASS = WT / M::RPP;    // gather passes per meta block (F = 32: 8 passes of 4 queries)
  constexpr int GU = NPASS / WS_GW >= 2 ? 2 : 1;  // passes in flight per gather warp
  extern __shared__ __align__(1024) unsigned char ws_smem[];
  unsigned char* sm = ws_smem;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const pinb200_map_view& m = p.map;
  const int K = p.opts.nn_k, L = p.dec.n_hidden, OC = p.dec.out_dim;
  const float slope = p.dec.leaky_relu ? 0.01f : 0.f;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + lay.bars);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(sm + lay.tmem);
  float* meta = reinterpret_cast<float*>(sm + lay.meta);
  WsClock clk;

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
    for (int s = 0; s < 8; ++s) {
      init(WSB_META_FULL + s, 1);
      init(WSB_META_EMPTY + s, WS_GW);
    }
    for (int g = 0; g < WS_EG; ++g) init(WSB_MMA + g, 1);
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  um_stage_weight(p.dec.w[0], D, H, D, H, K0, false, sm + lay.w0_hi, sm + lay.w0_lo);
  if (L > 1) um_stage_weight(p.dec.w[1], H, H, H, H, H, false, sm + lay.w1_hi, sm + lay.w1_lo);
  for (int e = tid; e < H; e += WS_THREADS) {
    reinterpret_cast<float*>(sm + lay.b0)[e] = p.dec.b[0] ? __ldg(p.dec.b[0] + e) : 0.f;
    reinterpret_cast<float*>(sm + lay.b1)[e] = (L > 1 && p.dec.b[1]) ? __ldg(p.dec.b[1] + e) : 0.f;
  }
  for (int e = tid; e < 4 * H; e += WS_THREADS) reinterpret_cast<float*>(sm + lay.wout)[e] = e < OC * H ? __ldg(p.dec.w_out + e) : 0.f;
  if (tid < 4) reinterpret_cast<float*>(sm + lay.bout)[tid] = (p.dec.b_out && tid < OC) ? __ldg(p.dec.b_out + tid) : 0.f;
  um_publish_and_sync();
  const uint32_t tmem_base = *s_tmem;
  clk.start(profile != 0);

  const long long n_tiles = (p.n + QT - 1) / QT;
  const int n_blocks = p.n_tiles;  // 32-query stash blocks

  if (warp < 4 * WS_EG) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(WS_REG_E));
    // =====================================================================================================
    // E: epilogue group g owns the tiles i = g, g + 2, ... of this CTA
    // profile slots: 0 group barrier, 1 wait A tile (issuing thread) + layer-0 MMAs, 2 layer-0 epilogue,
    //                3 layer-1 MMAs, 4 last-layer epilogue + outputs
    // =====================================================================================================
    const int g = warp >> 2, qd = warp & 3;
    const int r = qd * WT + lane;  // tile row == TMEM lane
    const uint32_t tb = tmem_base + g * WS_TCOLS;
    const uint32_t tl = tb + ((uint32_t)(qd * 32) << 16);
    const uint32_t mma_bar = um_smem_u32(bars + WSB_MMA + g);
    uint32_t mph = 0;
    const int t = GRAD ? (lane & 3) : 0;  // row type: 0 value, 1..3 tangent d/dq_{t-1}
    const float bsel = t == 0 ? 1.f : 0.f;
    const int src = lane & ~3;  // the value row of this row's query
    const float* s_b0 = reinterpret_cast<const float*>(sm + lay.b0);
    const float* s_b1 = reinterpret_cast<const float*>(sm + lay.b1);
    const float* s_wout = reinterpret_cast<const float*>(sm + lay.wout);
    const float* s_bout = reinterpret_cast<const float*>(sm + lay.bout);
    const uint32_t w0_hi = um_smem_u32(sm + lay.w0_hi), w0_lo = um_smem_u32(sm + lay.w0_lo);
    const uint32_t w1_hi = um_smem_u32(sm + lay.w1_hi), w1_lo = um_smem_u32(sm + lay.w1_lo);
    constexpr uint32_t A_SBO0 = (K0 / 4) * UM_A_LBO, W_SBO0 = (K0 / 4) * UM_W_LBO, W_SBO1 = (H / 4) * UM_W_LBO;
    const uint32_t idesc = um_idesc(H);
    int i = g;
    for (long long T = blockIdx.x + (long long)g * gridDim.x; T < n_tiles; T += (long long)WS_EG * gridDim.x, i += WS_EG) {
      const int slot = i % WS_A0;
      // every row of the group has read the previous tile's accumulator before the next MMA overwrites it
      ws_fence_before();
      ws_group_bar(1 + g);
      clk.lap(0);
      if (qd == 0) ws_wait(um_smem_u32(bars + WSB_A0_FULL + slot), (uint32_t)(i / WS_A0) & 1u);
      if (r == 0) {
        ws_fence_after();
        const uint32_t a_hi = um_smem_u32(sm + lay.a0 + slot * lay.a0_stride), a_lo = a_hi + lay.a0_half;
#pragma unroll 1
        for (int s = 0; s < K0 / 8; ++s) {
          const uint64_t ah = um_desc(a_hi + s * 2 * UM_A_LBO, UM_A_LBO, A_SBO0);
          const uint64_t al = um_desc(a_lo + s * 2 * UM_A_LBO, UM_A_LBO, A_SBO0);
          const uint64_t bh = um_desc(w0_hi + s * 2 * UM_W_LBO, UM_W_LBO, W_SBO0);
          const uint64_t bl = um_desc(w0_lo + s * 2 * UM_W_LBO, UM_W_LBO, W_SBO0);
          um_mma(tb, al, bh, idesc, s > 0);  // small terms first
          um_mma(tb, ah, bl, idesc, 1);
          um_mma(tb, ah, bh, idesc, 1);
        }
        ws_commit(mma_bar);
        ws_commit(um_smem_u32(bars + WSB_A0_EMPTY + slot));  // the A tile may be refilled once these MMAs have read it
      }
      ws_wait(mma_bar, mph);
      mph ^= 1u;
      ws_fence_after();
      clk.lap(1);
      if (L > 1) {
        // ---- layer-0 epilogue: bias (value rows), ReLU gate of the query's value row, hi / lo split -> A1 in TMEM
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t v[16], hi[16], lo[16];
          um_tmem_ld16(tl + 16 * c, v);
#pragma unroll
          for (int e4 = 0; e4 < 4; ++e4) {
            const float4 b = *reinterpret_cast<const float4*>(s_b0 + 16 * c + 4 * e4);
            const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float z = fmaf(bsel, bb[e], __uint_as_float(v[4 * e4 + e]));
              const float zv = GRAD ? __shfl_sync(FULL, z, src) : z;
              const float o = zv > 0.f ? z : slope * z;
              const uint32_t h = __float_as_uint(o) & TF32_MASK;
              hi[4 * e4 + e] = h;
              lo[4 * e4 + e] = __float_as_uint(o - __uint_as_float(h));
            }
          }
          ws_tmem_st16(tl + 64 + 16 * c, hi);
          ws_tmem_st16(tl + 128 + 16 * c, lo);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        ws_fence_before();
        ws_group_bar(1 + g);
        clk.lap(2);
        if (r == 0) {
          ws_fence_after();
#pragma unroll 1
          for (int s = 0; s < H / 8; ++s) {
            const uint64_t bh = um_desc(w1_hi + s * 2 * UM_W_LBO, UM_W_LBO, W_SBO1);
            const uint64_t bl = um_desc(w1_lo + s * 2 * UM_W_LBO, UM_W_LBO, W_SBO1);
            ws_mma_ts(tb, tb + 128 + 8 * s, bh, idesc, s > 0);
            ws_mma_ts(tb, tb + 64 + 8 * s, bl, idesc, 1);
            ws_mma_ts(tb, tb + 64 + 8 * s, bh, idesc, 1);
          }
          ws_commit(mma_bar);
        }
        ws_wait(mma_bar, mph);
        mph ^= 1u;
        ws_fence_after();
        clk.lap(3);
      }
      // ---- last hidden layer: bias, gate, output head(s) in registers
      float o[4] = {0.f, 0.f, 0.f, 0.f};
      {
        const float* bl = L > 1 ? s_b1 : s_b0;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t v[16];
          um_tmem_ld16(tl + 16 * c, v);
#pragma unroll
          for (int e4 = 0; e4 < 4; ++e4) {
            const float4 b = *reinterpret_cast<const float4*>(bl + 16 * c + 4 * e4);
            const float bb[4] = {b.x, b.y, b.z, b.w};
            float hh[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float z = fmaf(bsel, bb[e], __uint_as_float(v[4 * e4 + e]));
              const float zv = GRAD ? __shfl_sync(FULL, z, src) : z;
              hh[e] = zv > 0.f ? z : slope * z;
            }
#pragma unroll
            for (int ch = 0; ch < 4; ++ch)
              if (ch < OC) {
                const float4 wo = *reinterpret_cast<const float4*>(s_wout + ch * H + 16 * c + 4 * e4);
                o[ch] = fmaf(hh[3], wo.w, fmaf(hh[2], wo.z, fmaf(hh[1], wo.y, fmaf(hh[0], wo.x, o[ch]))));
              }
          }
        }
      }
      // ---- outputs: value rows write the prediction, tangent rows one component of its gradient
      const int ql = GRAD ? (r >> 2) : r;
      const long long qi = T * QT + ql;
      const bool live = qi < p.n;
#pragma unroll
      for (int ch = 0; ch < 4; ++ch)
        if (ch < OC) {
          const float oo = fmaf(bsel, s_bout[ch], o[ch]);
          float res;
          if (p.dec.sigmoid_out) {
            const float val = 1.f / (1.f + expf(-oo));
            const float dv = val * (1.f - val);
            const float dvq = GRAD ? __shfl_sync(FULL, dv, src) : dv;
            res = t == 0 ? val : dvq * oo;
          } else {
            res = oo * p.dec.out_scale;
          }
          if (live) {
            if (t == 0) {
              if (!p.is_color) {
                if (ch == 0) {
                  if (p.out.sdf) p.out.sdf[qi] = res;
                  if (p.out.sdf_std) p.out.sdf_std[qi] = 0.f;
                }
              } else if (p.out.color) {
                p.out.color[qi * OC + ch] = res;
              }
            } else {
              if (!p.is_color) {
                if (ch == 0 && p.out.grad) p.out.grad[3 * qi + (t - 1)] = res;
              } else if (p.out.color_grad) {
                p.out.color_grad[(qi * OC + ch) * 3 + (t - 1)] = res;
              }
            }
          }
        }
      clk.lap(4);
    }
  } else if (warp < 4 * WS_EG + WS_GT * WS_GW) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(WS_REG_G));
    // =====================================================================================================
    // G: gather teams -- the 4 warps of a team work on the same tile, each on its share of the queries
    // profile slots: 0 wait free A tile, 1 wait meta block, 2 feature-row loads issued, 3 reduce + A-tile stores,
    //                4 position rows + fence + arrive
    // =====================================================================================================
    const int team = (warp - 4 * WS_EG) / WS_GW, gw = (warp - 4 * WS_EG) % WS_GW;
    const int sub = lane / M::LPR, c4 = lane % M::LPR;
    const float4* __restrict__ f4 = reinterpret_cast<const float4*>(p.feat) + c4;
    int i = team;
    for (long long T = blockIdx.x + (long long)team * gridDim.x; T < n_tiles; T += (long long)WS_GT * gridDim.x, i += WS_GT) {
      const int slot = i % WS_A0;
      ws_wait(um_smem_u32(bars + WSB_A0_EMPTY + slot), ((uint32_t)(i / WS_A0) & 1u) ^ 1u);
      clk.lap(0);
      unsigned char* a_hi = sm + lay.a0 + slot * lay.a0_stride;
      unsigned char* a_lo = a_hi + lay.a0_half;
#pragma unroll 1
      for (int b = 0; b < BPT; ++b) {
        const int blk = i * BPT + b, ms = blk % MB;
        ws_wait(um_smem_u32(bars + WSB_META_FULL + ms), (uint32_t)(blk / MB) & 1u);
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
        // position part (columns F .. F+2) and zero padding of the rows: thread per row
        if (GRAD || (b % WS_GW) == gw) {
          const int row = GRAD ? gw * WT + lane : b * WT + lane;
          const int ql = GRAD ? row >> 2 : lane, tt = GRAD ? row & 3 : 0;
          const float* src3 = tt == 0 ? mt + WsMeta::xn : mt + WsMeta::P + (tt - 1) * 3 * WT;
          float4 hi, lo;
          um_split4(make_float4(src3[ql], src3[WT + ql], src3[2 * WT + ql], 0.f), hi, lo);
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
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(WS_REG_L));
    // =====================================================================================================
    // L: loader warps -- stash block of 32 queries -> meta block (thread per query)
    // profile slots: 0 wait free meta block, 1 stash loads, 2 seeds + meta stores
    // =====================================================================================================
    const int lw = warp - 4 * WS_EG - WS_GT * WS_GW;
    int i = 0;
    for (long long T = blockIdx.x; T < n_tiles; T += gridDim.x, ++i) {
#pragma unroll 1
      for (int b = 0; b < BPT; ++b) {
        const int blk = i * BPT + b;
        if (blk % WS_LW != lw) continue;
        const int ms = blk % MB;
        ws_wait(um_smem_u32(bars + WSB_META_EMPTY + ms), ((uint32_t)(blk / MB) & 1u) ^ 1u);
        clk.lap(0);
        float* mt = meta + ms * MSTRIDE;
        const long long st = T * BPT + b;  // stash block
        int lif[KREG], nn = 0;
        float w[KREG], dx[KREG], dy[KREG], dz[KREG], qx = 0.f, qy = 0.f, qz = 0.f, usum = 0.f, px = 0.f, py = 0.f, pz = 0.f;
#pragma unroll
        for (int k = 0; k < KREG; ++k) {
          lif[k] = -1;
          w[k] = dx[k] = dy[k] = dz[k] = 0.f;
        }
        if (st < n_blocks) {
          const float* __restrict__ sb = p.stash + (size_t)st * Stash::floats + lane;
#pragma unroll
          for (int k = 0; k < KREG; ++k)
            if (k < K) {
              lif[k] = __float_as_int(__ldg(sb + Stash::li + k * WT));
              w[k] = __ldg(sb + Stash::w + k * WT);
              if (GRAD) {
                dx[k] = __ldg(sb + Stash::dx + k * WT);
                dy[k] = __ldg(sb + Stash::dy + k * WT);
                dz[k] = __ldg(sb + Stash::dz + k * WT);
              }
            }
          px = __ldg(sb + Stash::pos);
          py = __ldg(sb + Stash::pos + WT);
          pz = __ldg(sb + Stash::pos + 2 * WT);
          if (GRAD) {
            qx = __ldg(sb + Stash::q);
            qy = __ldg(sb + Stash::q + WT);
            qz = __ldg(sb + Stash::q + 2 * WT);
            usum = __ldg(sb + Stash::usum);
            nn = __float_as_int(__ldg(sb + Stash::nn));
          }
        }
        clk.lap(1);
        ws_write_meta<GRAD>(m, K, lif, w, dx, dy, dz, qx, qy, qz, usum, nn, px, py, pz, mt, lane);
        __syncwarp();
        if (lane == 0) ws_arrive(um_smem_u32(bars + WSB_META_FULL + ms));
        clk.lap(2);
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
  l.n_meta = grad ? 4 : 8;
  l.meta = take(l.n_meta * l.meta_stride);
  l.bars = take(WSB_COUNT * 8);
  l.tmem = take(4);
  l.total = o;
  return l;
}

template <int FT, bool GRAD>
static int launch_wsq(QueryParams& p, cudaStream_t stream) {
  const WsLayout lay = plan_ws_layout<FT>(p.dec, GRAD);
  const size_t smem_bytes = (size_t)lay.total;
  if (smem_bytes > 227 * 1024) {
    set_error("wsq_decode kernel needs %zu B shared memory (> 227 KB)", smem_bytes);
    return PINB200_ERR_UNSUPPORTED;
  }
  auto kern = wsq_decode_kernel<FT, GRAD>;
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
  kern<<<grid, WS_THREADS, smem_bytes, stream>>>(p, lay, g_ws_profile);
  return check_launch("wsq_decode_kernel");
}

int dispatch_wsq(QueryParams& p, cudaStream_t stream) {
  const bool grad = p.opts.need_grad != 0;
  switch (p.dec.in_dim - 3) {
    case 8: return grad ? launch_wsq<8, true>(p, stream) : launch_wsq<8, false>(p, stream);
    case 16: return grad ? launch_wsq<16, true>(p, stream) : launch_wsq<16, false>(p, stream);
    case 32: return grad ? launch_wsq<32, true>(p, stream) : launch_wsq<32, false>(p, stream);
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
