// Top-8 candidate selection of the thread-per-query search (K1 phase A1), kept in its own header so that the
// selection logic can also be compiled for the host and checked against std::partial_sort
// (tests/native/select_equiv.cu).
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

struct KnnRegs {
  float d2[KREG];
  int idx[KREG];
  int gidx[KREG];
};

PINB_HD void knn_regs_init(KnnRegs& L) {
#pragma unroll
  for (int i = 0; i < KREG; ++i) {
    L.d2[i] = SEL_INVALID_D2;
    L.idx[i] = -1;
    L.gidx[i] = -1;
  }
}

// Candidate selection keeps the KREG smallest distances UNSORTED (replace the current worst, then find the new
// worst: ~30 predicated instructions per accepted candidate instead of a sorted insertion) and sorts once at the
// end with a 19-comparator network.  Ties: a candidate equal to the current worst is rejected (earlier probe
// wins, like the reference's stable behaviour on duplicates); the final order among exactly equal distances is
// unspecified, as it is for torch.sort.
struct KnnSel {
  float worst;
  int wpos;
};

PINB_HD void knn_sel_replace(KnnRegs& L, KnnSel& S, float d2, int li, int gi) {
#pragma unroll
  for (int i = 0; i < KREG; ++i)
    if (i == S.wpos) {
      L.d2[i] = d2;
      L.idx[i] = li;
      L.gidx[i] = gi;
    }
  S.worst = L.d2[0];
  S.wpos = 0;
#pragma unroll
  for (int i = 1; i < KREG; ++i)
    if (L.d2[i] > S.worst) {  // first maximum: with several empty (9e3) slots the lowest index is refilled first
      S.worst = L.d2[i];
      S.wpos = i;
    }
}

PINB_HD void knn_cswap(KnnRegs& L, int a, int b) {
  if (L.d2[b] < L.d2[a]) {
    const float td = L.d2[a];
    L.d2[a] = L.d2[b];
    L.d2[b] = td;
    const int ti = L.idx[a];
    L.idx[a] = L.idx[b];
    L.idx[b] = ti;
    const int tg = L.gidx[a];
    L.gidx[a] = L.gidx[b];
    L.gidx[b] = tg;
  }
}

// optimal 19-comparator sorting network for 8 keys (ascending)
PINB_HD void knn_sort8(KnnRegs& L) {
  knn_cswap(L, 0, 1); knn_cswap(L, 2, 3); knn_cswap(L, 4, 5); knn_cswap(L, 6, 7);
  knn_cswap(L, 0, 2); knn_cswap(L, 1, 3); knn_cswap(L, 4, 6); knn_cswap(L, 5, 7);
  knn_cswap(L, 1, 2); knn_cswap(L, 5, 6); knn_cswap(L, 0, 4); knn_cswap(L, 3, 7);
  knn_cswap(L, 1, 5); knn_cswap(L, 2, 6);
  knn_cswap(L, 1, 4); knn_cswap(L, 3, 6);
  knn_cswap(L, 2, 4); knn_cswap(L, 3, 5);
  knn_cswap(L, 3, 4);
}


// ---- variant with the ids in a scratch (PINB_K1_SMEM_SELECT): only the 8 distances live in registers; the two ids
// of an accepted candidate are stored at the replaced slot of a transposed [slot][lane] scratch (conflict-free
// st.shared when the scratch is in shared memory).  Decisions and the final order are identical to
// knn_sel_replace + knn_sort8 (same comparisons on the same distances, same sorting network).
struct KnnKeys {
  float d2[KREG];
  float worst;
  int wpos;
};

PINB_HD void knn_keys_init(KnnKeys& S, int* sc_l, int* sc_g, int lane) {
#pragma unroll
  for (int i = 0; i < KREG; ++i) {
    S.d2[i] = SEL_INVALID_D2;
    sc_l[i * 32 + lane] = -1;
    sc_g[i * 32 + lane] = -1;
  }
  S.worst = SEL_INVALID_D2;
  S.wpos = 0;
}

PINB_HD void knn_keys_accept(KnnKeys& S, float d2, int li, int gi, int* sc_l, int* sc_g, int lane) {
  sc_l[S.wpos * 32 + lane] = li;
  sc_g[S.wpos * 32 + lane] = gi;
#pragma unroll
  for (int i = 0; i < KREG; ++i) S.d2[i] = (i == S.wpos) ? d2 : S.d2[i];
  S.worst = S.d2[0];
  S.wpos = 0;
#pragma unroll
  for (int i = 1; i < KREG; ++i)
    if (S.d2[i] > S.worst) {
      S.worst = S.d2[i];
      S.wpos = i;
    }
}

PINB_HD void knn_keys_cswap(float (&d)[KREG], int (&p)[KREG], int a, int b) {
  if (d[b] < d[a]) {
    const float td = d[a];
    d[a] = d[b];
    d[b] = td;
    const int tp = p[a];
    p[a] = p[b];
    p[b] = tp;
  }
}

// sort (distance, slot) pairs with the same 19-comparator network, then fetch the ids of each slot once
PINB_HD void knn_keys_finish(const KnnKeys& S, const int* sc_l, const int* sc_g, int lane, KnnRegs& L) {
  float d[KREG];
  int p[KREG];
#pragma unroll
  for (int i = 0; i < KREG; ++i) {
    d[i] = S.d2[i];
    p[i] = i;
  }
  knn_keys_cswap(d, p, 0, 1); knn_keys_cswap(d, p, 2, 3); knn_keys_cswap(d, p, 4, 5); knn_keys_cswap(d, p, 6, 7);
  knn_keys_cswap(d, p, 0, 2); knn_keys_cswap(d, p, 1, 3); knn_keys_cswap(d, p, 4, 6); knn_keys_cswap(d, p, 5, 7);
  knn_keys_cswap(d, p, 1, 2); knn_keys_cswap(d, p, 5, 6); knn_keys_cswap(d, p, 0, 4); knn_keys_cswap(d, p, 3, 7);
  knn_keys_cswap(d, p, 1, 5); knn_keys_cswap(d, p, 2, 6);
  knn_keys_cswap(d, p, 1, 4); knn_keys_cswap(d, p, 3, 6);
  knn_keys_cswap(d, p, 2, 4); knn_keys_cswap(d, p, 3, 5);
  knn_keys_cswap(d, p, 3, 4);
#pragma unroll
  for (int i = 0; i < KREG; ++i) {
    L.d2[i] = d[i];
    L.idx[i] = sc_l[p[i] * 32 + lane];
    L.gidx[i] = sc_g[p[i] * 32 + lane];
  }
}

}  // namespace pinb
