// K1: fused voxel-hash kNN + IDW interpolation + decoder MLP + analytic d/dq.
//
// Every WARP owns an independent tile of 32 queries (no block-wide barriers after the one-off weight
// staging: 12 de-synchronised warps per SM overlap each other's memory and compute phases):
//   weighted_first=1 : 32 decoder rows, one per query (features IDW-averaged first)
//   weighted_first=0 : the 32 queries are decoded in row tiles of floor(32/K) queries x K rows
// Phase A (warp per query): hash the query's cell, probe the C neighbour cells of
//   the voxel hash table (one probe per lane, 32 at a time), age/distance filter,
//   in-register warp top-K, IDW weights, coalesced gather of the neighbour feature
//   rows (a 32-float row == one 128 B warp load), write the decoder input into the
//   transposed shared-memory tile.
// Phase B (thread per row): the tiny MLP forward and, if requested, the backward
//   pass w.r.t. the decoder input, weights broadcast from shared memory.
// Phase C (warp per query / thread per query): chain rule through the IDW weights
//   and the neighbour vectors -> d sdf / d query, sdf std, outputs.
//
// Replaces model/neural_points.py:530-746,950-1009, model/decoder.py:61-85,112,
// utils/tools.py:247-260, utils/tracker.py:313-328 of the reference.
#include <algorithm>
#include <mutex>
#include <vector>

#include "mlp.cuh"
#include "mlp_mma.cuh"

namespace pinb {

struct QueryLayout {  // float offsets into dynamic smem
  MmaDecSmem dec;
  int delta, warp0, warp_stride, n_warps;  // CTA-shared part, then n_warps per-warp blocks
  int act, knn_idx, knn_gidx, knn_d2, knn_w, knn_a, q, out, dv, nn, mask, total;  // offsets inside a warp block
};

struct QueryParams {
  pinb200_map_view map;
  pinb200_decoder_view dec;
  pinb200_query_opts opts;
  pinb200_query_out out;
  const float* query_xyz;
  const int32_t* query_ts;
  const float* feat;  // feature table decoded by `dec` (geo or colour)
  long long n;
  int use_saved_knn;  // 1: take kNN from out.knn_idx / out.knn_dist2 (decode-only launch, e.g. colour head)
  int is_color;       // outputs go to out.color / out.color_grad instead of sdf / grad
  int n_tiles;
  int qpt;  // queries per tile
  QueryLayout lay;
};

// ---------------------------------------------------------------------------
// warp-per-query feature movement.  The feature dimension FT is a template
// parameter so that the (neighbour, column) of every load is a compile-time
// function of the unrolled loop indices; each warp works on GQ queries at once and
// issues all of their row loads before consuming any (GQ * K independent 128-byte
// loads in flight per warp).
// ---------------------------------------------------------------------------
#ifndef PINB_K1_GQ
#define PINB_K1_GQ 8
#endif
constexpr int GQ_MAX = 8;  // queries gathered concurrently by one warp (4 for the widest rows: register budget)
template <int FT>
struct GqOf {
  static constexpr int value = FT >= 64 ? 4 : PINB_K1_GQ;
};
constexpr int WT = 32;  // queries (= threads) per warp tile
constexpr int WPB = 12; // max warps per CTA (one CTA per SM); fewer if shared memory does not fit

// FT >= 32: a neighbour row is FT/32 coalesced warp loads; FT < 32: 32/FT neighbour rows per warp load.
template <int FT>
struct FeatMap {
  static constexpr int NJ = FT >= 32 ? FT / 32 : 1;     // warp loads per neighbour row
  static constexpr int PER = FT >= 32 ? 1 : 32 / FT;    // neighbour rows per warp load
  static constexpr int R = FT >= 32 ? KREG * NJ : (KREG + PER - 1) / PER;  // warp loads per query (K = KREG)
};

// weighted_first: act[j][ql] = sum_k w_k f_k[j]  for the queries ql0 + g (g < GQ)
template <int FT, int LDX, int GQ>
__device__ __forceinline__ void gather_weighted_group(const float* __restrict__ feat, int K, const int* s_idx,
                                                      const float* s_w, int lane, float* s_act, int ql0, int qpt) {
  using M = FeatMap<FT>;
  float v[GQ][M::R], w[GQ][M::R];
#pragma unroll
  for (int g = 0; g < GQ; ++g) {
    const int ql = ql0 + g;
#pragma unroll
    for (int r = 0; r < M::R; ++r) {
      const int k = FT >= 32 ? r / M::NJ : r * M::PER + lane / (FT >= 32 ? 1 : FT);
      const int col = FT >= 32 ? 32 * (r % M::NJ) + lane : lane % (FT >= 32 ? 32 : FT);
      v[g][r] = 0.f;
      w[g][r] = 0.f;
      if (k < K && ql < qpt) {
        const int lk = s_idx[ql * K + k];
        w[g][r] = s_w[ql * K + k];
        v[g][r] = __ldg(feat + (size_t)(lk < 0 ? 0 : lk) * FT + col);
      }
    }
  }
#pragma unroll
  for (int g = 0; g < GQ; ++g) {
    const int ql = ql0 + g;
    if (FT >= 32) {
      float acc[M::NJ];
#pragma unroll
      for (int jj = 0; jj < M::NJ; ++jj) acc[jj] = 0.f;
#pragma unroll
      for (int r = 0; r < M::R; ++r) acc[r % M::NJ] = fmaf(w[g][r], v[g][r], acc[r % M::NJ]);
      if (ql < qpt)
#pragma unroll
        for (int jj = 0; jj < M::NJ; ++jj) s_act[ql * LDX + 32 * jj + lane] = acc[jj];
    } else {
      float a = 0.f;
#pragma unroll
      for (int r = 0; r < M::R; ++r) a = fmaf(w[g][r], v[g][r], a);
#pragma unroll
      for (int off = FT; off < 32; off <<= 1) a += __shfl_xor_sync(FULL, a, off);
      if (lane < FT && ql < qpt) s_act[ql * LDX + lane] = a;
    }
  }
}

// decode-every-neighbour: act[j][ql*K + k] = f_k[j] (0 if invalid)
template <int FT, int LDX, int GQ>
__device__ __forceinline__ void gather_rows_group(const float* __restrict__ feat, int K, const int* s_idx, int lane,
                                                  float* s_act, int ql0, int qpt, int sq0) {
  using M = FeatMap<FT>;
  float v[GQ][M::R];
#pragma unroll
  for (int g = 0; g < GQ; ++g) {
    const int ql = ql0 + g;
#pragma unroll
    for (int r = 0; r < M::R; ++r) {
      const int k = FT >= 32 ? r / M::NJ : r * M::PER + lane / (FT >= 32 ? 1 : FT);
      const int col = FT >= 32 ? 32 * (r % M::NJ) + lane : lane % (FT >= 32 ? 32 : FT);
      v[g][r] = 0.f;
      if (k < K && ql < qpt) {
        const int lk = s_idx[(sq0 + ql) * K + k];
        const float x = __ldg(feat + (size_t)(lk < 0 ? 0 : lk) * FT + col);
        v[g][r] = lk < 0 ? 0.f : x;
      }
    }
  }
#pragma unroll
  for (int g = 0; g < GQ; ++g) {
    const int ql = ql0 + g;
#pragma unroll
    for (int r = 0; r < M::R; ++r) {
      const int k = FT >= 32 ? r / M::NJ : r * M::PER + lane / (FT >= 32 ? 1 : FT);
      const int col = FT >= 32 ? 32 * (r % M::NJ) + lane : lane % (FT >= 32 ? 32 : FT);
      if (k < K && ql < qpt) s_act[(ql * K + k) * LDX + col] = v[g][r];
    }
  }
}

// a_k = <g_xbar[0..F), f_k> for the queries ql0 + 4*g  ->  s_a[ql*K + k]
template <int FT, int LDX, int GQ>
__device__ __forceinline__ void feature_dots_group(const float* __restrict__ feat, int K, const int* s_idx, int lane,
                                                   const float* s_act, float* s_a, int ql0, int qpt) {
  using M = FeatMap<FT>;
  float v[GQ][M::R];
#pragma unroll
  for (int g = 0; g < GQ; ++g) {
    const int ql = ql0 + g;
#pragma unroll
    for (int r = 0; r < M::R; ++r) {
      const int k = FT >= 32 ? r / M::NJ : r * M::PER + lane / (FT >= 32 ? 1 : FT);
      const int col = FT >= 32 ? 32 * (r % M::NJ) + lane : lane % (FT >= 32 ? 32 : FT);
      v[g][r] = 0.f;
      if (k < K && ql < qpt) {
        const int lk = s_idx[ql * K + k];
        const float x = __ldg(feat + (size_t)(lk < 0 ? 0 : lk) * FT + col);
        v[g][r] = lk < 0 ? 0.f : x;
      }
    }
  }
#pragma unroll
  for (int g = 0; g < GQ; ++g) {
    const int ql = ql0 + g;
    const int qs = ql < qpt ? ql : 0;
    if (FT >= 32) {
      float gx[M::NJ];
#pragma unroll
      for (int jj = 0; jj < M::NJ; ++jj) gx[jj] = s_act[qs * LDX + 32 * jj + lane];
      float part[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        part[k] = 0.f;
#pragma unroll
        for (int jj = 0; jj < M::NJ; ++jj) part[k] = fmaf(gx[jj], v[g][k * M::NJ + jj], part[k]);
      }
      const float tot = warp_reduce8(part, lane);
      const int k = warp_reduce8_owner(lane);
      if ((lane & 3) == 0 && k < K && ql < qpt) s_a[ql * K + k] = tot;
    } else {
      const float gxj = s_act[qs * LDX + (lane % (FT >= 32 ? 32 : FT))];
#pragma unroll
      for (int r = 0; r < M::R; ++r) {
        float a = gxj * v[g][r];
#pragma unroll
        for (int off = 1; off < (FT >= 32 ? 1 : FT); off <<= 1) a += __shfl_xor_sync(FULL, a, off);
        const int k = r * M::PER + lane / (FT >= 32 ? 1 : FT);
        if ((lane % (FT >= 32 ? 32 : FT)) == 0 && k < K && ql < qpt) s_a[ql * K + k] = a;
      }
    }
  }
}

// ---------------------------------------------------------------------------
// the fused kernel
// ---------------------------------------------------------------------------
template <int H, int FT>
__global__ void __launch_bounds__(WPB * 32, 1) query_kernel(const __grid_constant__ QueryParams p) {
  static_assert(H == 64, "the tensor-core decoder is written for hidden_dim 64");
  constexpr int KP0 = (FT + 3 + 7) / 8 * 8;        // decoder input width padded to the MMA k-step
  constexpr int KT0 = KP0 / 8;                     // k-steps of layer 0 == n-tiles of the input gradient
  constexpr int LDX = (KP0 > H ? KP0 : H) + 4;     // row-major tile leading dimension (== 4 or 12 mod 32)
  constexpr int GQ = GqOf<FT>::value;
  extern __shared__ __align__(16) float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const pinb200_map_view& m = p.map;
  const int K = p.opts.nn_k, L = p.dec.n_hidden;
  constexpr int F = FT, D = FT + 3;
  const int OC = p.dec.out_dim;
  const bool wf = p.opts.weighted_first != 0;
  const bool need_grad = p.opts.need_grad != 0;
  const bool leaky = p.dec.leaky_relu != 0;
  const float* __restrict__ feat = p.feat;

  uint32_t* s_delta = reinterpret_cast<uint32_t*>(smem + p.lay.delta);
  float* wsm = smem + p.lay.warp0 + warp * p.lay.warp_stride;  // this warp's private tile state
  float* s_act = wsm + p.lay.act;
  int* s_idx = reinterpret_cast<int*>(wsm + p.lay.knn_idx);
  int* s_gidx = reinterpret_cast<int*>(wsm + p.lay.knn_gidx);
  float* s_d2 = wsm + p.lay.knn_d2;
  float* s_w = wsm + p.lay.knn_w;
  float* s_a = wsm + p.lay.knn_a;
  float* s_q = wsm + p.lay.q;
  float* s_out = wsm + p.lay.out;
  float* s_dv = wsm + p.lay.dv;
  int* s_nn = reinterpret_cast<int*>(wsm + p.lay.nn);
  uint64_t* s_mask = reinterpret_cast<uint64_t*>(wsm + p.lay.mask);

  stage_mma_decoder(p.dec, p.lay.dec, smem);
  if (!p.use_saved_knn) fill_probe_deltas(m, s_delta);
  __syncthreads();

  const int QPT = wf ? WT : WT / K;                   // queries per 32-row decoder tile
  const int WQ = p.qpt;                               // queries per warp tile (<= WT; small launches use fewer so
                                                      // that every resident warp gets work: latency, not issue, bound)
  const int n_rt = wf ? 1 : (WQ + QPT - 1) / QPT;     // row tiles per warp tile
  const int tid = lane;                               // row / query owned by this thread inside the warp tile
  const int nwarp = blockDim.x >> 5;
  for (int st = blockIdx.x * nwarp + warp; st < p.n_tiles; st += gridDim.x * nwarp) {
    const long long q0s = (long long)st * WQ;

    // ============ phase A1: thread per query -- search, IDW weights, side effects ============
    {
      const long long qi = q0s + tid;
      const bool live = tid < WQ && qi < p.n;
      KnnRegs Lk;
      knn_regs_init(Lk);
      int cnt = 0;
      float qx = 0.f, qy = 0.f, qz = 0.f;
      if (live) {
        qx = __ldg(p.query_xyz + 3 * qi + 0);
        qy = __ldg(p.query_xyz + 3 * qi + 1);
        qz = __ldg(p.query_xyz + 3 * qi + 2);
        if (p.opts.transform) {  // q = T p in fp32 (utils/tools.py:534-553)
          const double* T = p.opts.transform;
          const float x = fmaf(qz, (float)T[2], fmaf(qy, (float)T[1], qx * (float)T[0])) + (float)T[3];
          const float y = fmaf(qz, (float)T[6], fmaf(qy, (float)T[5], qx * (float)T[4])) + (float)T[7];
          const float z = fmaf(qz, (float)T[10], fmaf(qy, (float)T[9], qx * (float)T[8])) + (float)T[11];
          qx = x;
          qy = y;
          qz = z;
        }
        if (p.use_saved_knn) {
#pragma unroll
          for (int k = 0; k < KREG; ++k)
            if (k < K) {
              Lk.idx[k] = __ldg(p.out.knn_idx + qi * K + k);
              Lk.gidx[k] = __ldg(p.out.knn_gidx + qi * K + k);
              Lk.d2[k] = __ldg(p.out.knn_dist2 + qi * K + k);
            }
          cnt = __ldg(p.out.nn_count + qi);
        } else {
          cnt = knn_search_thread(m, s_delta, qx, qy, qz, Lk, reinterpret_cast<int*>(s_act));
        }
      }
#ifdef PINB_K1_SMEM_SELECT
      __syncwarp();  // the selection scratch lives in the (idle) activation tile: every lane is done reading it
#endif
      // normalised inverse-distance weights, summed in neighbour order (model/neural_points.py:665-683)
      float u[KREG], w[KREG], usum = 0.f;
#pragma unroll
      for (int k = 0; k < KREG; ++k) {
        const bool v = k < K && Lk.idx[k] >= 0;
        u[k] = k < K ? (cnt == 0 ? IDW_EPS : (v ? __fdiv_rn(1.0f, Lk.d2[k] + IDW_EPS) : 0.f)) : 0.f;
        usum += u[k];
      }
#pragma unroll
      for (int k = 0; k < KREG; ++k) w[k] = (k < K && Lk.idx[k] >= 0) ? __fdiv_rn(u[k], usum) : 0.f;
#pragma unroll
      for (int k = 0; k < KREG; ++k)
        if (k < K) {
          s_idx[tid * K + k] = Lk.idx[k];
          s_gidx[tid * K + k] = Lk.gidx[k];
          s_d2[tid * K + k] = Lk.d2[k];
          s_w[tid * K + k] = w[k];
        }
      s_nn[tid] = cnt;
      s_q[3 * tid + 0] = qx;
      s_q[3 * tid + 1] = qy;
      s_q[3 * tid + 2] = qz;
      // neighbour vectors n_k = q - p_k (rotated into the point frame after PGO), certainty (:631-651)
      float sx = 0.f, sy = 0.f, sz = 0.f, qc = 0.f;
      {
        float px[KREG], py[KREG], pz[KREG], ce[KREG];
#pragma unroll
        for (int k = 0; k < KREG; ++k) {
          px[k] = py[k] = pz[k] = ce[k] = 0.f;
          if (k < K && Lk.idx[k] >= 0) {
            const float* pp = m.nb_points + 3 * (size_t)Lk.idx[k];
            px[k] = __ldg(pp);
            py[k] = __ldg(pp + 1);
            pz[k] = __ldg(pp + 2);
            if (!p.is_color) ce[k] = m.certainty[Lk.idx[k]];
          }
        }
#pragma unroll
        for (int k = 0; k < KREG; ++k) {
          if (k < K && Lk.idx[k] >= 0) {
            float nx = __fsub_rn(qx, px[k]), ny = __fsub_rn(qy, py[k]), nz = __fsub_rn(qz, pz[k]);
            if (m.after_pgo) {
              const float* qq = m.nb_orient + 4 * (size_t)Lk.idx[k];
              quat_rotate_passive(__ldg(qq), __ldg(qq + 1), __ldg(qq + 2), __ldg(qq + 3), nx, ny, nz, nx, ny, nz);
            }
            sx = fmaf(w[k], nx, sx);
            sy = fmaf(w[k], ny, sy);
            sz = fmaf(w[k], nz, sz);
            qc = fmaf(w[k], ce[k], qc);
          }
        }
      }
      if (wf) {  // the position part of the IDW-averaged decoder input (this thread's tile row) + zero padding
        s_act[tid * LDX + F + 0] = sx;
        s_act[tid * LDX + F + 1] = sy;
        s_act[tid * LDX + F + 2] = sz;
#pragma unroll
        for (int d = D; d < KP0; ++d) s_act[tid * LDX + d] = 0.f;
      }
      if (live && !p.is_color) {
        if (p.opts.training_mode && (p.opts.training_rows <= 0 || qi < p.opts.training_rows)) {
          // certainty scatter_add / ts amax (:685-710); invalid entries add 0 / max with 0 in the reference
          const int ts = (m.ts_update && p.query_ts) ? __ldg(p.query_ts + qi) : 0;
#pragma unroll
          for (int k = 0; k < KREG; ++k)
            if (k < K && Lk.idx[k] >= 0) {
              atomicAdd(m.certainty + Lk.idx[k], w[k]);
              if (m.ts_update && p.query_ts) atomicMax(m.ts_update + Lk.idx[k], ts);
            }
        }
        if (p.out.certainty) p.out.certainty[qi] = qc;
        if (!p.use_saved_knn) {
          if (p.out.nn_count) p.out.nn_count[qi] = cnt;
#pragma unroll
          for (int k = 0; k < KREG; ++k)
            if (k < K) {
              if (p.out.knn_idx) p.out.knn_idx[qi * K + k] = Lk.idx[k];
              if (p.out.knn_gidx) p.out.knn_gidx[qi * K + k] = Lk.gidx[k];
              if (p.out.knn_dist2) p.out.knn_dist2[qi * K + k] = Lk.d2[k];
              if (p.out.knn_weight) p.out.knn_weight[qi * K + k] = w[k];
            }
        }
        if (p.out.xyz) {
          p.out.xyz[3 * qi + 0] = qx;
          p.out.xyz[3 * qi + 1] = qy;
          p.out.xyz[3 * qi + 2] = qz;
        }
      }
    }
    __syncwarp();

    for (int rt = 0; rt < n_rt; ++rt) {
      const int sq0 = rt * QPT;                              // first query of this row tile inside the super tile
      const int qpt = min(QPT, WQ - sq0);                  // queries in this row tile
      const int used_rows = wf ? WT : qpt * K;

      // ============ phase A2: warp per query -- coalesced feature gathers into the tile ============
      if (wf) {
        for (int ql0 = 0; ql0 < qpt; ql0 += GQ) gather_weighted_group<FT, LDX, GQ>(feat, K, s_idx, s_w, lane, s_act, ql0, qpt);
      } else {
        for (int ql0 = 0; ql0 < qpt; ql0 += GQ) gather_rows_group<FT, LDX, GQ>(feat, K, s_idx, lane, s_act, ql0, qpt, sq0);
        // neighbour vectors of the (query, k) rows: thread per row
        if (tid < used_rows) {
          const int ql = tid / K, k = tid - ql * K, sq = sq0 + ql;
          const int lk = s_idx[sq * K + k];
          float nx = 0.f, ny = 0.f, nz = 0.f;
          if (lk >= 0) {
            const float* pp = m.nb_points + 3 * (size_t)lk;
            nx = __fsub_rn(s_q[3 * sq + 0], __ldg(pp));
            ny = __fsub_rn(s_q[3 * sq + 1], __ldg(pp + 1));
            nz = __fsub_rn(s_q[3 * sq + 2], __ldg(pp + 2));
            if (m.after_pgo) {
              const float* qq = m.nb_orient + 4 * (size_t)lk;
              quat_rotate_passive(__ldg(qq), __ldg(qq + 1), __ldg(qq + 2), __ldg(qq + 3), nx, ny, nz, nx, ny, nz);
            }
          }
          s_act[tid * LDX + F + 0] = nx;
          s_act[tid * LDX + F + 1] = ny;
          s_act[tid * LDX + F + 2] = nz;
#pragma unroll
          for (int d = D; d < KP0; ++d) s_act[tid * LDX + d] = 0.f;
        } else {
          for (int d = 0; d < KP0; ++d) s_act[tid * LDX + d] = 0.f;  // unused rows stay finite
        }
      }
      __syncwarp();

      // ============ phase B: decoder on the tensor cores (warp-level 3xTF32 MMA) ============
      float acc[2][8][4];
      warp_gemm_3xtf32<KT0, 8, false, LDX>(acc, s_act, smem + p.lay.dec.whi[0], smem + p.lay.dec.wlo[0], p.lay.dec.ldw[0], lane);
      s_mask[lane] = bias_act_frags<8>(acc, smem + p.lay.dec.b[0], leaky, lane);
      for (int l = 1; l < L; ++l) {
        __syncwarp();
        store_frags<8, LDX>(s_act, acc, lane);
        __syncwarp();
        warp_gemm_3xtf32<8, 8, false, LDX>(acc, s_act, smem + p.lay.dec.whi[l], smem + p.lay.dec.wlo[l], p.lay.dec.ldw[l], lane);
        s_mask[l * WT + lane] = bias_act_frags<8>(acc, smem + p.lay.dec.b[l], leaky, lane);
      }
      // output head(s): dot with w_out over the 64 hidden units; 4 lanes share a row
      for (int c = 0; c < OC; ++c) {
        const float* wo = smem + p.lay.dec.wout + c * H;
        float part[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int nt = 0; nt < 8; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) part[mt][e >> 1] = fmaf(acc[mt][nt][e], wo[frag_col(nt, e, lane)], part[mt][e >> 1]);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            float v = part[mt][hh];
            v += __shfl_xor_sync(FULL, v, 1);
            v += __shfl_xor_sync(FULL, v, 2);
            if ((lane & 3) == 0) {
              const int row = mt * 16 + (lane >> 2) + 8 * hh;
              const float o = v + smem[p.lay.dec.bout + c];
              float val, dv;
              if (p.dec.sigmoid_out) {
                val = 1.f / (1.f + expf(-o));
                dv = val * (1.f - val);
              } else {
                val = o * p.dec.out_scale;
                dv = p.dec.out_scale;
              }
              s_out[row * OC + c] = val;
              s_dv[row * 4 + c] = dv;
            }
          }
      }
      __syncwarp();

      // ======== per output channel: backward to the decoder input, then the IDW chain rule ========
      const int n_pass = need_grad ? OC : 1;
      for (int c = 0; c < n_pass; ++c) {
        if (need_grad) {
          const float* wo = smem + p.lay.dec.wout + c * H;
#pragma unroll
          for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 8; ++nt)
#pragma unroll
              for (int e = 0; e < 4; ++e) acc[mt][nt][e] = wo[frag_col(nt, e, lane)];
          mask_frags<8>(acc, s_mask[(L - 1) * WT + lane], leaky);
          for (int l = L - 1; l >= 1; --l) {
            __syncwarp();
            store_frags<8, LDX>(s_act, acc, lane);
            __syncwarp();
            warp_gemm_3xtf32<8, 8, true, LDX>(acc, s_act, smem + p.lay.dec.whi[l], smem + p.lay.dec.wlo[l], p.lay.dec.ldw[l], lane);
            mask_frags<8>(acc, s_mask[(l - 1) * WT + lane], leaky);
          }
          __syncwarp();
          store_frags<8, LDX>(s_act, acc, lane);
          __syncwarp();
          float gxa[2][KT0][4];
          warp_gemm_3xtf32<8, KT0, true, LDX>(gxa, s_act, smem + p.lay.dec.whi[0], smem + p.lay.dec.wlo[0], p.lay.dec.ldw[0], lane);
#pragma unroll
          for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < KT0; ++nt)
#pragma unroll
              for (int e = 0; e < 4; ++e) gxa[mt][nt][e] *= s_dv[frag_row(mt, e, lane) * 4 + c];
          __syncwarp();
          store_frags<KT0, LDX>(s_act, gxa, lane);
        }
        __syncwarp();

        if (wf) {
          // ---- C1: a_k = <g_xbar, f_k> (warp per query, coalesced re-read of the K feature rows)
          if (need_grad) {
            for (int ql0 = 0; ql0 < qpt; ql0 += GQ) feature_dots_group<FT, LDX, GQ>(feat, K, s_idx, lane, s_act, s_a, ql0, qpt);
            __syncwarp();
          }
          // ---- C2: thread per query -- chain rule through the IDW weights, outputs
          const long long qi = q0s + tid;
          if (tid < WQ && qi < p.n) {
            const float val = s_out[tid * OC + c];
            float gq0 = 0.f, gq1 = 0.f, gq2 = 0.f;
            if (need_grad) {
              const int nn = s_nn[tid];
              const float qx = s_q[3 * tid], qy = s_q[3 * tid + 1], qz = s_q[3 * tid + 2];
              const float gn0 = s_act[tid * LDX + F + 0], gn1 = s_act[tid * LDX + F + 1], gn2 = s_act[tid * LDX + F + 2];
              float ak[KREG], wk[KREG], ck[KREG], dx[KREG], dy[KREG], dz[KREG];
              float abar = 0.f;
              // issue every neighbour load first (local position, global position, quaternion), then compute
              float lpx[KREG], lpy[KREG], lpz[KREG], gpx[KREG], gpy[KREG], gpz[KREG];
              float4 qt[KREG];
              int lks[KREG];
#pragma unroll
              for (int k = 0; k < KREG; ++k) {
                lks[k] = k < K ? s_idx[tid * K + k] : -1;
                const int lk0 = lks[k] < 0 ? 0 : lks[k];
                const int gk0 = lks[k] < 0 ? 0 : s_gidx[tid * K + k];
                const float* pp = m.nb_points + 3 * (size_t)lk0;
                const float* pg = m.points + 3 * (size_t)gk0;
                lpx[k] = __ldg(pp);
                lpy[k] = __ldg(pp + 1);
                lpz[k] = __ldg(pp + 2);
                gpx[k] = __ldg(pg);
                gpy[k] = __ldg(pg + 1);
                gpz[k] = __ldg(pg + 2);
                qt[k] = m.after_pgo ? __ldg(reinterpret_cast<const float4*>(m.nb_orient) + lk0) : make_float4(1.f, 0.f, 0.f, 0.f);
              }
#pragma unroll
              for (int k = 0; k < KREG; ++k) {
                ak[k] = wk[k] = ck[k] = dx[k] = dy[k] = dz[k] = 0.f;
                if (lks[k] >= 0) {
                  const float ux = qx - lpx[k], uy = qy - lpy[k], uz = qz - lpz[k];
                  dx[k] = qx - gpx[k];  // what dist2 was measured to
                  dy[k] = qy - gpy[k];
                  dz[k] = qz - gpz[k];
                  float nx = ux, ny = uy, nz = uz, r0 = gn0, r1 = gn1, r2 = gn2;
                  if (m.after_pgo) {
                    quat_rotate_passive(qt[k].x, qt[k].y, qt[k].z, qt[k].w, ux, uy, uz, nx, ny, nz);
                    quat_rotate_active(qt[k].x, qt[k].y, qt[k].z, qt[k].w, gn0, gn1, gn2, r0, r1, r2);
                  }
                  wk[k] = s_w[tid * K + k];
                  ak[k] = s_a[tid * K + k] + gn0 * nx + gn1 * ny + gn2 * nz;
                  abar = fmaf(wk[k], ak[k], abar);
                  ck[k] = nn > 0 ? -2.f * __fdiv_rn(1.0f, s_d2[tid * K + k] + IDW_EPS) : 0.f;
                  gq0 = fmaf(wk[k], r0, gq0);
                  gq1 = fmaf(wk[k], r1, gq1);
                  gq2 = fmaf(wk[k], r2, gq2);
                }
              }
              // d w_k / d q = w_k (c_k - sum_j w_j c_j),  c_k = -2 u_k (q - p_k).
              // sum_k w_k (a_k - abar) c_k is invariant to a common shift of the a_k; shifting by the nearest
              // neighbour's a_0 first keeps (a_k - abar) exact when neighbours coincide (cancellation-free)
              {
                const float a0 = ak[0];
                abar = 0.f;
#pragma unroll
                for (int k = 0; k < KREG; ++k) {
                  ak[k] = wk[k] != 0.f ? ak[k] - a0 : 0.f;
                  abar = fmaf(wk[k], ak[k], abar);
                }
              }
#pragma unroll
              for (int k = 0; k < KREG; ++k) {
                const float coef = wk[k] * (ak[k] - abar) * ck[k];
                gq0 = fmaf(coef, dx[k], gq0);
                gq1 = fmaf(coef, dy[k], gq1);
                gq2 = fmaf(coef, dz[k], gq2);
              }
            }
            if (!p.is_color) {
              if (p.out.sdf) p.out.sdf[qi] = val;
              if (p.out.sdf_std) p.out.sdf_std[qi] = 0.f;
              if (need_grad && p.out.grad) {
                p.out.grad[3 * qi + 0] = gq0;
                p.out.grad[3 * qi + 1] = gq1;
                p.out.grad[3 * qi + 2] = gq2;
              }
            } else {
              if (p.out.color) {
                if (need_grad) {
                  p.out.color[qi * OC + c] = val;
                } else {
                  for (int cc = 0; cc < OC; ++cc) p.out.color[qi * OC + cc] = s_out[tid * OC + cc];
                }
              }
              if (need_grad && p.out.color_grad) {
                p.out.color_grad[(qi * OC + c) * 3 + 0] = gq0;
                p.out.color_grad[(qi * OC + c) * 3 + 1] = gq1;
                p.out.color_grad[(qi * OC + c) * 3 + 2] = gq2;
              }
            }
          }
        } else {
          // decode-every-neighbour: thread per query combines its K rows (tracker.py:317-323)
          if (tid < qpt && q0s + sq0 + tid < p.n) {
            const int ql = tid, sq = sq0 + ql;
            const long long qi = q0s + sq;
            const int nn = s_nn[sq];
            const float qx = s_q[3 * sq], qy = s_q[3 * sq + 1], qz = s_q[3 * sq + 2];
            const int n_ch = (need_grad || !p.is_color) ? 1 : OC;
            for (int ch = 0; ch < n_ch; ++ch) {
              const int cc = need_grad ? c : ch;
              float mean = 0.f, wsum = 0.f, msh = 0.f;
              const float s0 = s_out[(ql * K) * OC + cc];  // nearest neighbour's value: shift for exact differences
              for (int k = 0; k < K; ++k) {
                const float wk = s_w[sq * K + k];
                mean = fmaf(wk, s_out[(ql * K + k) * OC + cc], mean);
                msh = fmaf(wk, s_out[(ql * K + k) * OC + cc] - s0, msh);
                wsum += wk;
              }
              float var = 0.f, g0 = 0.f, g1 = 0.f, g2 = 0.f;
              for (int k = 0; k < K; ++k) {
                const int lk = s_idx[sq * K + k];
                if (lk < 0) continue;
                const float wk = s_w[sq * K + k];
                const float dm = (s_out[(ql * K + k) * OC + cc] - s0) - msh;  // == s_k - mean, cancellation-free
                var = fmaf(wk * dm, dm, var);
                if (need_grad) {
                  const int row = ql * K + k;
                  float r0 = s_act[row * LDX + F + 0], r1 = s_act[row * LDX + F + 1], r2 = s_act[row * LDX + F + 2];
                  const float* pg = m.points + 3 * (size_t)s_gidx[sq * K + k];  // the point dist2 was measured to
                  const float dx = qx - __ldg(pg), dy = qy - __ldg(pg + 1), dz = qz - __ldg(pg + 2);
                  if (m.after_pgo) {
                    const float* qq = m.nb_orient + 4 * (size_t)lk;
                    quat_rotate_active(__ldg(qq), __ldg(qq + 1), __ldg(qq + 2), __ldg(qq + 3), r0, r1, r2, r0, r1, r2);
                  }
                  const float uk = nn > 0 ? __fdiv_rn(1.0f, s_d2[sq * K + k] + IDW_EPS) : 0.f;
                  const float coef = wk * dm * (-2.f * uk);
                  g0 += fmaf(coef, dx, wk * r0);
                  g1 += fmaf(coef, dy, wk * r1);
                  g2 += fmaf(coef, dz, wk * r2);
                }
              }
              if (!p.is_color) {
                if (p.out.sdf) p.out.sdf[qi] = mean;
                if (p.out.sdf_std) p.out.sdf_std[qi] = sqrtf(var);
                if (need_grad && p.out.grad) {
                  p.out.grad[3 * qi + 0] = g0;
                  p.out.grad[3 * qi + 1] = g1;
                  p.out.grad[3 * qi + 2] = g2;
                }
              } else {
                if (p.out.color) p.out.color[qi * OC + cc] = mean;
                if (need_grad && p.out.color_grad) {
                  p.out.color_grad[(qi * OC + cc) * 3 + 0] = g0;
                  p.out.color_grad[(qi * OC + cc) * 3 + 1] = g1;
                  p.out.color_grad[(qi * OC + cc) * 3 + 2] = g2;
                }
              }
            }
          }
        }
        __syncwarp();
      }
    }
  }
}

// ---------------------------------------------------------------------------
// search-only kernels (warp per query, no decoder)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(TILE) knn_kernel(const __grid_constant__ pinb200_map_view m,
                                                   const float* __restrict__ query_xyz, long long n, int K,
                                                   int32_t* knn_idx, int32_t* knn_gidx, float* knn_d2, float* knn_w,
                                                   int32_t* nn_count) {
  extern __shared__ __align__(16) float smem[];
  uint32_t* s_delta = reinterpret_cast<uint32_t*>(smem);
  fill_probe_deltas(m, s_delta);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long wpb = blockDim.x >> 5;
  for (long long qi = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); qi < n; qi += (long long)gridDim.x * wpb) {
    const float qx = __ldg(query_xyz + 3 * qi), qy = __ldg(query_xyz + 3 * qi + 1), qz = __ldg(query_xyz + 3 * qi + 2);
    const Knn kn = knn_search_warp(m, s_delta, qx, qy, qz, K, lane);
    float u, inv_s;
    const float w = idw_weight(kn.d2, kn.idx >= 0, kn.count, K, lane, u, inv_s);
    if (lane < K) {
      if (knn_idx) knn_idx[qi * K + lane] = kn.idx;
      if (knn_gidx) knn_gidx[qi * K + lane] = kn.gidx;
      if (knn_d2) knn_d2[qi * K + lane] = kn.d2;
      if (knn_w) knn_w[qi * K + lane] = w;
    }
    if (lane == 0 && nn_count) nn_count[qi] = kn.count;
  }
}

// all-probe radius search: dist2 [N,C], global ids [N,C] (model/neural_points.py:950-1009)
__global__ void __launch_bounds__(TILE) radius_kernel(const __grid_constant__ pinb200_map_view m,
                                                      const float* __restrict__ query_xyz, long long n, float* dist2,
                                                      int32_t* idx, float* max_cert) {
  extern __shared__ __align__(16) float smem[];
  uint32_t* s_delta = reinterpret_cast<uint32_t*>(smem);
  fill_probe_deltas(m, s_delta);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long wpb = blockDim.x >> 5;
  const int C = m.n_probe;
  for (long long qi = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); qi < n; qi += (long long)gridDim.x * wpb) {
    const float qx = __ldg(query_xyz + 3 * qi), qy = __ldg(query_xyz + 3 * qi + 1), qz = __ldg(query_xyz + 3 * qi + 2);
    const uint32_t r0 = base_slot(m, qx, qy, qz);
    const float td_cur = m.time_filter ? __ldg(m.travel_dist + m.cur_ts) : 0.f;
    float best = 0.f;
    for (int c = lane; c < C; c += 32) {
      float d2;
      int li, gi;
      probe_cell(m, r0, s_delta[c], qx, qy, qz, td_cur, d2, li, gi);
      if (dist2) dist2[qi * C + c] = d2;
      if (idx) idx[qi * C + c] = gi;
      if (max_cert && gi >= 0) best = fmaxf(best, m.certainty[gi]);  // query_certainty (:1025-1028)
    }
    if (max_cert) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) best = fmaxf(best, __shfl_xor_sync(FULL, best, o));
      if (lane == 0) max_cert[qi] = best;
    }
  }
}

// query_feature's materialised feature vectors from saved kNN (compat path)
__global__ void __launch_bounds__(256) gather_kernel(const __grid_constant__ pinb200_map_view m,
                                                      const float* __restrict__ feat,
                                                      const float* __restrict__ query_xyz,
                                                      const int32_t* __restrict__ knn_idx,
                                                      const float* __restrict__ knn_w, long long n, int K, int wf,
                                                      float* out) {
  const int F = m.feature_dim, D = F + 3;
  const long long total = wf ? n * D : n * K * D;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int d = (int)(e % D);
    const long long r = e / D;
    const long long qi = wf ? r : r / K;
    float acc = 0.f;
    const int k_lo = wf ? 0 : (int)(r % K), k_hi = wf ? K : k_lo + 1;
    for (int k = k_lo; k < k_hi; ++k) {
      const int lk = knn_idx[qi * K + k];
      if (lk < 0) continue;
      float v;
      if (d < F) {
        v = __ldg(feat + (size_t)lk * F + d);
      } else {
        float nx = query_xyz[3 * qi] - m.nb_points[3 * (size_t)lk];
        float ny = query_xyz[3 * qi + 1] - m.nb_points[3 * (size_t)lk + 1];
        float nz = query_xyz[3 * qi + 2] - m.nb_points[3 * (size_t)lk + 2];
        if (m.after_pgo) {
          const float* qq = m.nb_orient + 4 * (size_t)lk;
          quat_rotate_passive(qq[0], qq[1], qq[2], qq[3], nx, ny, nz, nx, ny, nz);
        }
        v = d == F ? nx : (d == F + 1 ? ny : nz);
      }
      acc = wf ? fmaf(knn_w[qi * K + k], v, acc) : v;
    }
    out[e] = acc;
  }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
static int validate_map(const pinb200_map_view* m, bool need_feat) {
  if (!m || !m->slot_table || !m->points || !m->probe_dx || m->n_probe <= 0) {
    set_error("map view: null table/points/probe offsets");
    return PINB200_ERR_BAD_ARG;
  }
  if (m->buffer_size <= 0 || m->buffer_size >= (1LL << 31)) {
    set_error("map view: buffer_size %lld out of range (0, 2^31)", (long long)m->buffer_size);
    return PINB200_ERR_BAD_ARG;
  }
  if (m->time_filter && (!m->ts_create || !m->travel_dist)) {
    set_error("map view: time_filter needs ts_create and travel_dist");
    return PINB200_ERR_BAD_ARG;
  }
  if (need_feat) {
    if (!m->nb_points || !m->geo_feat || !m->certainty) {
      set_error("map view: null neighbour arrays");
      return PINB200_ERR_BAD_ARG;
    }
    const int F = m->feature_dim;
    const bool okF = F == 4 || F == 8 || F == 16 || F == 32 || F == 64;
    if (!okF) {
      set_error("feature_dim %d unsupported (4, 8, 16, 32, 64)", F);
      return PINB200_ERR_UNSUPPORTED;
    }
    if (m->after_pgo && !m->nb_orient) {
      set_error("map view: after_pgo needs nb_orient");
      return PINB200_ERR_BAD_ARG;
    }
  }
  return PINB200_OK;
}

static QueryLayout plan_layout(const QueryParams& p, int KP0) {
  QueryLayout l{};
  const int H = p.dec.hidden_dim, K = p.opts.nn_k;
  int o = 0;
  l.delta = o;
  o += align4(p.map.n_probe);
  l.dec = plan_mma_decoder_smem(p.dec, KP0, o);
  l.warp0 = align4(l.dec.end);
  // per-warp block
  int w = 0;
  const int ldx = (KP0 > H ? KP0 : H) + 4;
  l.act = w;
  w += align4(WT * ldx);
  l.knn_idx = w;
  w += WT * K;
  l.knn_gidx = w;
  w += WT * K;
  l.knn_d2 = w;
  w += WT * K;
  l.knn_w = w;
  w += WT * K;
  l.knn_a = w;
  w += WT * K;
  l.q = w;
  w += WT * 3;
  l.out = w;
  w += align4(WT * p.dec.out_dim);
  l.dv = w;
  w += WT * 4;
  l.nn = w;
  w += WT;
  w = (w + 1) & ~1;  // 8-byte align the 64-bit masks
  l.mask = w;
  w += 2 * WT * p.dec.n_hidden;
  l.warp_stride = align4(w);
  int nw = (227 * 1024 / 4 - l.warp0) / l.warp_stride;
  l.n_warps = nw > WPB ? WPB : nw;
  l.total = l.warp0 + l.n_warps * l.warp_stride;
  return l;
}

template <int H, int FT>
static int launch_query(QueryParams& p, cudaStream_t stream) {
  constexpr int KP0 = (FT + 3 + 7) / 8 * 8;
  p.lay = plan_layout(p, KP0);
  const size_t smem_bytes = (size_t)p.lay.total * sizeof(float);
  if (p.lay.n_warps < 1 || smem_bytes > 227 * 1024) {
    set_error("query kernel needs %zu B shared memory (> 227 KB)", smem_bytes);
    return PINB200_ERR_UNSUPPORTED;
  }
  // small batches (tracker: a few thousand points, mapper: ~26k rows): spread the warp tiles over all SMs
  // instead of packing 10-12 warps into a few CTAs
  int nw = p.lay.n_warps;
  {
    // queries per warp tile: a full 32 when there is enough work to occupy every resident warp, fewer (whole
    // decoder row tiles in decode-every-neighbour mode) for the tracker / mapper sized launches
    const long long slots = (long long)sm_count() * nw;
    const long long per_warp = (p.n + slots - 1) / slots;
    int wq;
    if (p.opts.weighted_first) {
      wq = (int)std::min<long long>(WT, std::max<long long>(GQ_MAX, (per_warp + GQ_MAX - 1) / GQ_MAX * GQ_MAX));
    } else {
      const int qpt_rows = WT / p.opts.nn_k;
      wq = (int)std::min<long long>(WT, (per_warp + qpt_rows - 1) / qpt_rows * qpt_rows);
    }
    p.qpt = wq;
    p.n_tiles = (int)((p.n + wq - 1) / wq);
    const long long per_sm = (p.n_tiles + sm_count() - 1) / sm_count();
    if (per_sm < nw) nw = (int)std::max<long long>(1, per_sm);
  }
  auto kern = query_kernel<H, FT>;
  // the attribute / occupancy calls cost a few microseconds each: remember the answer per (device, warps, smem)
  struct Cached {
    int dev, nw, occ;
    size_t smem;
  };
  static std::mutex mu;
  static std::vector<Cached> cache;
  int occ = 0, dev = 0;
  cudaGetDevice(&dev);
  {
    std::lock_guard<std::mutex> lk(mu);
    for (const Cached& c : cache)
      if (c.dev == dev && c.nw == nw && c.smem == smem_bytes) occ = c.occ;
    if (occ == 0) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      if (e != cudaSuccess) {
        set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        return PINB200_ERR_CUDA;
      }
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, nw * 32, smem_bytes);
      if (occ < 1) occ = 1;
      cache.push_back({dev, nw, occ, smem_bytes});
    }
  }
  const long long ctas_needed = (p.n_tiles + nw - 1) / nw;
  const int grid = (int)std::min<long long>(ctas_needed, (long long)sm_count() * occ);
  kern<<<grid, nw * 32, smem_bytes, stream>>>(p);
  return check_launch("query_kernel");
}

static int dispatch_query(QueryParams& p, cudaStream_t stream) {
  const int D = p.dec.in_dim;
  if (p.dec.hidden_dim != 64) {
    set_error("decoder hidden_dim %d unsupported (64)", p.dec.hidden_dim);
    return PINB200_ERR_UNSUPPORTED;
  }
  switch (D - 3) {
    case 4: return launch_query<64, 4>(p, stream);
    case 8: return launch_query<64, 8>(p, stream);
    case 16: return launch_query<64, 16>(p, stream);
    case 32: return launch_query<64, 32>(p, stream);
    case 64: return launch_query<64, 64>(p, stream);
    default: break;
  }
  set_error("feature_dim %d unsupported (4, 8, 16, 32, 64)", D - 3);
  return PINB200_ERR_UNSUPPORTED;
}

static int validate_decoder(const pinb200_decoder_view* d, int F) {
  if (!d || !d->w_out || d->n_hidden < 1 || d->n_hidden > PINB200_MAX_HIDDEN_LAYERS) {
    set_error("decoder view: bad n_hidden / null w_out");
    return PINB200_ERR_BAD_ARG;
  }
  for (int l = 0; l < d->n_hidden; ++l)
    if (!d->w[l]) {
      set_error("decoder view: null weight of layer %d", l);
      return PINB200_ERR_BAD_ARG;
    }
  if (d->in_dim != F + 3) {
    set_error("decoder in_dim %d != feature_dim+3 = %d (positional encoding is not supported)", d->in_dim, F + 3);
    return PINB200_ERR_UNSUPPORTED;
  }
  if (d->out_dim < 1 || d->out_dim > 4) {
    set_error("decoder out_dim %d unsupported (1..4)", d->out_dim);
    return PINB200_ERR_UNSUPPORTED;
  }
  return PINB200_OK;
}

}  // namespace pinb

using namespace pinb;

extern "C" int pinb200_query_sdf(const pinb200_map_view* map, const pinb200_decoder_view* sdf_dec,
                                 const pinb200_decoder_view* color_dec, const float* query_xyz,
                                 const int32_t* query_ts, int64_t n, const pinb200_query_opts* opts,
                                 const pinb200_query_out* out, void* stream) {
  int rc = validate_map(map, true);
  if (rc) return rc;
  if (!opts || !out || (!query_xyz && n > 0)) {
    set_error("query_sdf: null opts/out/query");
    return PINB200_ERR_BAD_ARG;
  }
  rc = validate_decoder(sdf_dec, map->feature_dim);
  if (rc) return rc;
  const int K = opts->nn_k;
  if (K < 1 || K > KREG || K > map->n_probe) {
    set_error("nn_k %d out of range (1..%d, <= n_probe %d)", K, KREG, map->n_probe);
    return PINB200_ERR_BAD_ARG;
  }
  if (n <= 0) return PINB200_OK;
  if (color_dec) {
    rc = validate_decoder(color_dec, map->feature_dim);
    if (rc) return rc;
    if (!map->color_feat || !out->knn_idx || !out->knn_gidx || !out->knn_dist2 || !out->nn_count) {
      set_error("colour head needs map->color_feat and out->knn_idx/knn_gidx/knn_dist2/nn_count as scratch");
      return PINB200_ERR_BAD_ARG;
    }
  }
  QueryParams p{};
  p.map = *map;
  p.dec = *sdf_dec;
  p.opts = *opts;
  p.out = *out;
  p.query_xyz = query_xyz;
  p.query_ts = query_ts;
  p.feat = map->geo_feat;
  p.n = n;
  p.qpt = WT;  // queries per warp tile
  p.n_tiles = (int)((n + WT - 1) / WT);
  rc = dispatch_query(p, (cudaStream_t)stream);
  if (rc) return rc;
  if (color_dec) {  // second launch: decode the colour features with the kNN the first launch saved
    QueryParams c = p;
    c.dec = *color_dec;
    c.feat = map->color_feat;
    c.use_saved_knn = 1;
    c.is_color = 1;
    c.opts.training_mode = 0;
    c.opts.need_grad = (opts->need_grad && out->color_grad) ? 1 : 0;
    rc = dispatch_query(c, (cudaStream_t)stream);
  }
  return rc;
}

extern "C" int pinb200_knn_search(const pinb200_map_view* map, const float* query_xyz, int64_t n, int32_t nn_k,
                                  int32_t* knn_idx, int32_t* knn_gidx, float* knn_dist2, float* knn_weight,
                                  int32_t* nn_count, void* stream) {
  int rc = validate_map(map, false);
  if (rc) return rc;
  if (nn_k < 1 || nn_k > PINB200_MAX_K || nn_k > map->n_probe) {
    set_error("nn_k %d out of range", nn_k);
    return PINB200_ERR_BAD_ARG;
  }
  if (n <= 0) return PINB200_OK;
  const int grid = (int)std::min<long long>((n + 3) / 4, (long long)sm_count() * 8);
  knn_kernel<<<grid, TILE, align4(map->n_probe) * sizeof(float), (cudaStream_t)stream>>>(
      *map, query_xyz, n, nn_k, knn_idx, knn_gidx, knn_dist2, knn_weight, nn_count);
  return check_launch("knn_kernel");
}

extern "C" int pinb200_radius_search(const pinb200_map_view* map, const float* query_xyz, int64_t n, float* dist2,
                                     int32_t* idx, void* stream) {
  int rc = validate_map(map, false);
  if (rc) return rc;
  if (n <= 0) return PINB200_OK;
  const int grid = (int)std::min<long long>((n + 3) / 4, (long long)sm_count() * 8);
  radius_kernel<<<grid, TILE, align4(map->n_probe) * sizeof(float), (cudaStream_t)stream>>>(*map, query_xyz, n, dist2,
                                                                                            idx, nullptr);
  return check_launch("radius_kernel");
}

extern "C" int pinb200_query_certainty(const pinb200_map_view* map, const float* query_xyz, int64_t n,
                                       float* out_certainty, void* stream) {
  int rc = validate_map(map, false);
  if (rc) return rc;
  if (map->global2local || map->time_filter || !map->certainty) {
    set_error("query_certainty works on the global arrays: global2local must be NULL, time_filter 0");
    return PINB200_ERR_BAD_ARG;
  }
  if (n <= 0) return PINB200_OK;
  const int grid = (int)std::min<long long>((n + 3) / 4, (long long)sm_count() * 8);
  radius_kernel<<<grid, TILE, align4(map->n_probe) * sizeof(float), (cudaStream_t)stream>>>(*map, query_xyz, n, nullptr,
                                                                                            nullptr, out_certainty);
  return check_launch("radius_kernel(certainty)");
}

extern "C" int pinb200_gather_features(const pinb200_map_view* map, const float* feat, const float* query_xyz,
                                       const int32_t* knn_idx, const float* knn_weight, int64_t n, int32_t nn_k,
                                       int32_t weighted_first, float* out, void* stream) {
  if (!map || !feat || !knn_idx || !knn_weight || !out || !map->nb_points) {
    set_error("gather_features: null argument");
    return PINB200_ERR_BAD_ARG;
  }
  if (n <= 0) return PINB200_OK;
  const long long total = (weighted_first ? n : n * nn_k) * (map->feature_dim + 3);
  const int grid = (int)std::min<long long>((total + 255) / 256, (long long)sm_count() * 16);
  gather_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*map, feat, query_xyz, knn_idx, knn_weight, n, nn_k,
                                                        weighted_first, out);
  return check_launch("gather_kernel");
}
