// Host-side check of the top-8 candidate selection used by the thread-per-query search of K1
// (pin_slam_b200/csrc/knn_select.cuh).  Compiled by nvcc as plain host code, runs without a GPU:
//   1. the register-only variant (knn_sel_replace + knn_sort8) returns the 8 smallest distances in ascending
//      order with their ids (checked against std::stable_sort; ids compared when the distances are distinct),
//   2. the scratch variant (PINB_K1_SMEM_SELECT: knn_keys_*) makes bit-identical decisions for every lane.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../../pin_slam_b200/csrc/knn_select.cuh"

using namespace pinb;

struct Cand {
  float d2;
  int li, gi;
};

int main() {
  std::mt19937 rng(1234);
  std::vector<int> sc_l(KREG * 32), sc_g(KREG * 32);
  long checked = 0;
  for (int trial = 0; trial < 200000; ++trial) {
    const int n = (int)(rng() % 40);
    const bool ties = (trial % 3) == 0;
    std::vector<Cand> c(n);
    std::vector<int> used;
    for (int i = 0; i < n; ++i) {
      float d;
      if (ties) {
        d = 0.25f * (float)(rng() % 12);
      } else {  // distinct, exactly representable distances
        int k;
        do k = (int)(rng() % (1 << 20)); while (std::find(used.begin(), used.end(), k) != used.end());
        used.push_back(k);
        d = (float)k * (1.0f / 1024.0f);
      }
      c[i] = {d, (int)(rng() % 100000), (int)(rng() % 10000000)};
    }
    const int lane = (int)(rng() % 32);
    // register-only variant
    KnnRegs A;
    knn_regs_init(A);
    KnnSel S;
    S.worst = SEL_INVALID_D2;
    S.wpos = 0;
    for (const Cand& x : c)
      if (x.d2 < S.worst) knn_sel_replace(A, S, x.d2, x.li, x.gi);
    knn_sort8(A);
    // scratch variant (other lanes' scratch entries are poisoned to catch addressing mistakes)
    std::fill(sc_l.begin(), sc_l.end(), -777);
    std::fill(sc_g.begin(), sc_g.end(), -777);
    KnnKeys Kk;
    knn_keys_init(Kk, sc_l.data(), sc_g.data(), lane);
    for (const Cand& x : c)
      if (x.d2 < Kk.worst) knn_keys_accept(Kk, x.d2, x.li, x.gi, sc_l.data(), sc_g.data(), lane);
    KnnRegs B;
    knn_keys_finish(Kk, sc_l.data(), sc_g.data(), lane, B);
    for (int i = 0; i < KREG; ++i)
      if (A.d2[i] != B.d2[i] || A.idx[i] != B.idx[i] || A.gidx[i] != B.gidx[i]) {
        std::printf("variant mismatch: trial %d slot %d: (%g,%d,%d) vs (%g,%d,%d)\n", trial, i, A.d2[i], A.idx[i],
                    A.gidx[i], B.d2[i], B.idx[i], B.gidx[i]);
        return 1;
      }
    for (int e = 0; e < KREG * 32; ++e)
      if ((e % 32) != lane && (sc_l[e] != -777 || sc_g[e] != -777)) {
        std::printf("scratch of another lane touched: trial %d entry %d\n", trial, e);
        return 1;
      }
    // against a plain sort
    std::vector<Cand> s = c;
    std::stable_sort(s.begin(), s.end(), [](const Cand& a, const Cand& b) { return a.d2 < b.d2; });
    for (int i = 0; i < KREG; ++i) {
      const float want = i < n ? s[i].d2 : SEL_INVALID_D2;
      if (A.d2[i] != want) {
        std::printf("distance mismatch: trial %d slot %d: %g vs %g\n", trial, i, A.d2[i], want);
        return 1;
      }
      if (i >= n && (A.idx[i] != -1 || A.gidx[i] != -1)) {
        std::printf("empty slot carries an id: trial %d slot %d\n", trial, i);
        return 1;
      }
      if (!ties && i < n && (A.idx[i] != s[i].li || A.gidx[i] != s[i].gi)) {
        std::printf("id mismatch: trial %d slot %d\n", trial, i);
        return 1;
      }
    }
    ++checked;
  }
  std::printf("ok %ld\n", checked);
  return 0;
}
