// Top-8 candidate selection of the thread-per-query search (K1 phase A1), kept in its own header so that the
// selection logic can also be compiled for the host and checked against std::stable_sort
// (tests/native/select_equiv.cu).
//
// Round 1 kept the 8 best candidates in an unsorted register list and replaced the worst one on every accepted
// probe: with 32 divergent lanes that ~160-instruction update ran on almost every one of the 33 probes and was 30 %
// of all K1 instructions (profiles/r01_k1_v52_phase_breakdown.csv).  This version is branch-free: the probes arrive
// in batches of 8, a batch is sorted with the optimal 19-comparator network, merged into the running sorted top-8
// with one half-cleaner (8 compare-selects) and re-sorted with a 12-comparator bitonic merger.  A comparator is
// 5 ALU instructions on a (distance, payload) pair, so 33 probes cost 19 + 3*39 + 8 = 144 comparators for every
// lane in lock step instead of ~5000 warp instructions.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define PINB_HD __host__ __device__ __forceinline__
#else
#define PINB_HD inline
#endif

namespace pinb {

constexpr float SEL_INVALID_D2 = 9e3f;  // == INVALID_D2 (model/neural_points.py:583)
constexpr int KREG = 8;

// Sorted (ascending distance) list of the best candidates seen so far.  `p` is an opaque payload (K1: the probe
// index, -1 for an invalid entry).  Ties: a later candidate with a distance equal to a kept one never displaces
// it; the order among exactly equal distances is unspecified, as it is for torch.sort.
struct KnnTop {
  float d[KREG];
  int p[KREG];
};

PINB_HD void knn_top_init(KnnTop& T) {
#pragma unroll
  for (int i = 0; i < KREG; ++i) {
    T.d[i] = SEL_INVALID_D2;
    T.p[i] = -1;
  }
}

PINB_HD void knn_cswap(float& da, int& pa, float& db, int& pb) {
  const bool s = db < da;
  const float lo = s ? db : da, hi = s ? da : db;
  const int pl = s ? pb : pa, ph = s ? pa : pb;
  da = lo;
  pa = pl;
  db = hi;
  pb = ph;
}

#define PINB_CS(a, b) knn_cswap(d[a], p[a], d[b], p[b])
// optimal 19-comparator sorting network for 8 keys (ascending)
PINB_HD void knn_sort8(float (&d)[KREG], int (&p)[KREG]) {
  PINB_CS(0, 1); PINB_CS(2, 3); PINB_CS(4, 5); PINB_CS(6, 7);
  PINB_CS(0, 2); PINB_CS(1, 3); PINB_CS(4, 6); PINB_CS(5, 7);
  PINB_CS(1, 2); PINB_CS(5, 6); PINB_CS(0, 4); PINB_CS(3, 7);
  PINB_CS(1, 5); PINB_CS(2, 6);
  PINB_CS(1, 4); PINB_CS(3, 6);
  PINB_CS(2, 4); PINB_CS(3, 5);
  PINB_CS(3, 4);
}
// bitonic merger: sorts any bitonic sequence of 8 keys (12 comparators)
PINB_HD void knn_bitonic8(float (&d)[KREG], int (&p)[KREG]) {
  PINB_CS(0, 4); PINB_CS(1, 5); PINB_CS(2, 6); PINB_CS(3, 7);
  PINB_CS(0, 2); PINB_CS(1, 3); PINB_CS(4, 6); PINB_CS(5, 7);
  PINB_CS(0, 1); PINB_CS(2, 3); PINB_CS(4, 5); PINB_CS(6, 7);
}
#undef PINB_CS

// T <- the 8 smallest of (T, batch), sorted.  The batch is sorted in place first.
PINB_HD void knn_top_merge8(KnnTop& T, float (&d)[KREG], int (&p)[KREG]) {
  knn_sort8(d, p);
  // half-cleaner of two ascending runs: min(T[i], batch[7-i]) is a bitonic sequence holding the 8 smallest
#pragma unroll
  for (int i = 0; i < KREG; ++i) {
    const bool s = d[KREG - 1 - i] < T.d[i];
    T.d[i] = s ? d[KREG - 1 - i] : T.d[i];
    T.p[i] = s ? p[KREG - 1 - i] : T.p[i];
  }
  knn_bitonic8(T.d, T.p);
}

// first batch: the running list is empty
PINB_HD void knn_top_first8(KnnTop& T, float (&d)[KREG], int (&p)[KREG]) {
  knn_sort8(d, p);
#pragma unroll
  for (int i = 0; i < KREG; ++i) {
    T.d[i] = d[i];
    T.p[i] = p[i];
  }
}

// single candidate (the 33rd probe of the default neighbourhood): 8 comparators
PINB_HD void knn_top_insert1(KnnTop& T, float d, int p) {
#pragma unroll
  for (int i = 0; i < KREG; ++i) knn_cswap(T.d[i], T.p[i], d, p);
}

}  // namespace pinb
