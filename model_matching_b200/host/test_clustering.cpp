// test_clustering -- reads poses (n, then n x (16 col-major floats + lcp)) and parameters from
// stdin, prints the indices greedy_clustering keeps; used by tests/test_clustering.py.
#include <cstdio>
#include <vector>

#include "pose_clustering.hpp"

int main() {
  int n, max_count;
  float frac, best, min_d, min_a, sym[3];
  if (scanf("%d %f %f %d %f %f %f %f %f", &n, &frac, &best, &max_count, &min_d, &min_a, &sym[0], &sym[1], &sym[2]) != 9) return 1;
  std::vector<PoseCandidate*> all, out;
  for (int i = 0; i < n; ++i) {
    Eigen::Matrix4f T;
    float lcp;
    for (int k = 0; k < 16; ++k) if (scanf("%f", &T.data()[k]) != 1) return 1;
    if (scanf("%f", &lcp) != 1) return 1;
    all.push_back(new PoseCandidate(T, lcp, (float)i));
  }
  clustering::greedy_clustering(all, frac, best, max_count, min_d, min_a, Eigen::Vector3f(sym[0], sym[1], sym[2]), out);
  for (auto* p : out) printf("%d\n", p->base_index);
  if (n >= 2) {
    float r, t;
    clustering::get_pose_diff(all[0]->transform, all[1]->transform, Eigen::Vector3f(sym[0], sym[1], sym[2]), r, t);
    printf("diff %.6f %.6f\n", r, t);
  }
  return 0;
}
