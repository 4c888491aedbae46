// Host-side check of the top-8 candidate selection used by the thread-per-query search of K1
// (pin_slam_b200/csrc/knn_select.cuh).  Compiled by nvcc as plain host code, runs without a GPU:
// candidate streams of 0..40 probes are fed in batches of 8 exactly as K1 does it (first batch -> knn_top_first8,
// full batches -> knn_top_merge8, a single trailing probe -> knn_top_insert1, a longer ragged tail -> a batch padded
// with invalid entries) and the result must be the 8 smallest distances in ascending order with their payloads
// (checked against std::stable_sort; payloads compared when the distances are distinct).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../../pin_slam_b200/csrc/knn_select.cuh"

using namespace pinb;

struct Cand {
  float d2;
  int pay;
};

int main() {
  std::mt19937 rng(1234);
  long checked = 0;
  for (int trial = 0; trial < 200000; ++trial) {
    const int n = (int)(rng() % 41);
    const bool ties = (trial % 3) == 0;
    std::vector<Cand> c(n);
    std::vector<int> used;
    for (int i = 0; i < n; ++i) {
      float d;
      const bool invalid = (rng() % 4) == 0;  // an empty / rejected probe
      if (invalid) {
        d = SEL_INVALID_D2;
      } else if (ties) {
        d = 0.25f * (float)(rng() % 12);
      } else {  // distinct, exactly representable distances
        int k;
        do k = (int)(rng() % (1 << 20)); while (std::find(used.begin(), used.end(), k) != used.end());
        used.push_back(k);
        d = (float)k * (1.0f / 1024.0f);
      }
      c[i] = {d, invalid ? -1 : i};
    }
    KnnTop T;
    knn_top_init(T);
    for (int c0 = 0; c0 < n; c0 += KREG) {
      if (n - c0 == 1 && c0 > 0) {
        knn_top_insert1(T, c[c0].d2, c[c0].pay);
        break;
      }
      float d[KREG];
      int p[KREG];
      for (int j = 0; j < KREG; ++j) {
        d[j] = c0 + j < n ? c[c0 + j].d2 : SEL_INVALID_D2;
        p[j] = c0 + j < n ? c[c0 + j].pay : -1;
      }
      if (c0 == 0)
        knn_top_first8(T, d, p);
      else
        knn_top_merge8(T, d, p);
    }
    std::vector<Cand> s = c;
    std::stable_sort(s.begin(), s.end(), [](const Cand& a, const Cand& b) { return a.d2 < b.d2; });
    for (int i = 0; i < KREG; ++i) {
      const float want = i < n ? s[i].d2 : SEL_INVALID_D2;
      if (T.d[i] != want) {
        std::printf("distance mismatch: trial %d slot %d: %g vs %g\n", trial, i, T.d[i], want);
        return 1;
      }
      if (want == SEL_INVALID_D2) {
        if (T.p[i] != -1) {
          std::printf("invalid slot carries payload %d: trial %d slot %d\n", T.p[i], trial, i);
          return 1;
        }
      } else if (!ties && T.p[i] != s[i].pay) {
        std::printf("payload mismatch: trial %d slot %d: %d vs %d\n", trial, i, T.p[i], s[i].pay);
        return 1;
      } else if (ties) {  // the payload must belong to SOME candidate with that distance
        if (T.p[i] < 0 || T.p[i] >= n || c[T.p[i]].d2 != want) {
          std::printf("tie payload %d does not carry distance %g: trial %d slot %d\n", T.p[i], want, trial, i);
          return 1;
        }
      }
    }
    if (ties) {  // no payload may appear twice
      for (int i = 0; i < KREG; ++i)
        for (int j = i + 1; j < KREG; ++j)
          if (T.p[i] >= 0 && T.p[i] == T.p[j]) {
            std::printf("payload %d selected twice: trial %d\n", T.p[i], trial);
            return 1;
          }
    }
    ++checked;
  }
  std::printf("ok %ld\n", checked);
  return 0;
}
