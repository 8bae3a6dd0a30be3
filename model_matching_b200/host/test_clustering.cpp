// test_clustering -- reads poses (n, then n x (16 col-major floats + lcp)) and parameters from
// stdin, prints the indices greedy_clustering keeps; used by tests/test_clustering.py.
// `test_clustering icp`: reads ns nt, ns source points (x y z), nt model points (x y z nx ny nz),
// runs clustering::point_to_plane_icp (needs the GPU) and prints the 16 column-major entries of
// the offset and the moved source; used by tests/test_icp_gpu.py.
#include <cstdio>
#include <cstring>
#include <memory>
#include <vector>

#include "pose_clustering.hpp"

// `icp stale`: offset_transform enters as a non-identity matrix, so a run that does not converge
// must come back as identity (src/pose_clustering.cpp:136-139)
static int icp_main(bool stale) {
  int ns, nt;
  if (scanf("%d %d", &ns, &nt) != 2) return 1;
  auto seg = std::make_shared<PCLPointCloud>(), model = std::make_shared<PCLPointCloud>();
  seg->points.resize(ns);
  model->points.resize(nt);
  for (auto& p : seg->points) { if (scanf("%f %f %f", &p.x, &p.y, &p.z) != 3) return 1; p.nx = 0; p.ny = 0; p.nz = 1; }
  for (auto& p : model->points) if (scanf("%f %f %f %f %f %f", &p.x, &p.y, &p.z, &p.nx, &p.ny, &p.nz) != 6) return 1;
  Eigen::Matrix4f off = Eigen::Matrix4f::Identity();
  if (stale) for (int k = 0; k < 16; ++k) off.data()[k] = 7.0f + (float)k;
  clustering::point_to_plane_icp(seg, model, off);
  for (int k = 0; k < 16; ++k) printf("%.9g ", off.data()[k]);
  printf("\n");
  for (auto& p : seg->points) printf("%.9g %.9g %.9g\n", p.x, p.y, p.z);
  return 0;
}

int main(int argc, char** argv) {
  if (argc > 1 && !strcmp(argv[1], "icp")) return icp_main(argc > 2 && !strcmp(argv[2], "stale"));
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
