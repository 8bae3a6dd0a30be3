// pose_clustering.cpp -- restatement of the reference's src/pose_clustering.cpp:5-122.
#include "pose_clustering.hpp"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <iostream>
#include <stdexcept>
#include <string>

#include "../../include/stocs_b200.h"

namespace clustering {

namespace {
// inverse of a 3x3 by cofactors (Eigen Matrix3f::inverse() uses the same closed form)
void inverse3(const float a[3][3], float inv[3][3]) {
  const float c00 = a[1][1] * a[2][2] - a[1][2] * a[2][1], c01 = a[1][2] * a[2][0] - a[1][0] * a[2][2],
              c02 = a[1][0] * a[2][1] - a[1][1] * a[2][0];
  const float det = a[0][0] * c00 + a[0][1] * c01 + a[0][2] * c02;
  const float id = 1.0f / det;
  inv[0][0] = c00 * id; inv[0][1] = (a[0][2] * a[2][1] - a[0][1] * a[2][2]) * id; inv[0][2] = (a[0][1] * a[1][2] - a[0][2] * a[1][1]) * id;
  inv[1][0] = c01 * id; inv[1][1] = (a[0][0] * a[2][2] - a[0][2] * a[2][0]) * id; inv[1][2] = (a[0][2] * a[1][0] - a[0][0] * a[1][2]) * id;
  inv[2][0] = c02 * id; inv[2][1] = (a[0][1] * a[2][0] - a[0][0] * a[2][1]) * id; inv[2][2] = (a[0][0] * a[1][1] - a[0][1] * a[1][0]) * id;
}
// Eigen::Quaternionf(Matrix3f) (QuaternionBase::operator=(rotation matrix), Shoemake's method)
void quat_from_matrix(const float m[3][3], float& w, float& x, float& y, float& z) {
  float t = m[0][0] + m[1][1] + m[2][2];
  if (t > 0.0f) {
    t = std::sqrt(t + 1.0f);
    w = 0.5f * t;
    t = 0.5f / t;
    x = (m[2][1] - m[1][2]) * t; y = (m[0][2] - m[2][0]) * t; z = (m[1][0] - m[0][1]) * t;
  } else {
    int i = 0;
    if (m[1][1] > m[0][0]) i = 1;
    if (m[2][2] > m[i][i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(m[i][i] - m[j][j] - m[k][k] + 1.0f);
    float q[3];
    q[i] = 0.5f * t;
    t = 0.5f / t;
    w = (m[k][j] - m[j][k]) * t;
    q[j] = (m[j][i] + m[i][j]) * t;
    q[k] = (m[k][i] + m[i][k]) * t;
    x = q[0]; y = q[1]; z = q[2];
  }
}
}  // namespace

// src/pose_clustering.cpp:5-72
void get_pose_diff(const Eigen::Matrix4f& test_pose, const Eigen::Matrix4f& base_pose, const Eigen::Vector3f& sym_info,
                   float& mean_rotation_error, float& translation_error) {
  float tr[3][3], br[3][3], ti[3][3], rd[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) { tr[i][j] = test_pose(i, j); br[i][j] = base_pose(i, j); }
  inverse3(tr, ti);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) rd[i][j] = ti[i][0] * br[0][j] + (ti[i][1] * br[1][j] + ti[i][2] * br[2][j]);
  float qw, qx, qy, qz;
  quat_from_matrix(rd, qw, qx, qy, qz);
  // quaternion_to_euler (:5-26)
  float e[3];
  const double sinr = +2.0 * (qw * qx + qy * qz), cosr = +1.0 - 2.0 * (qx * qx + qy * qy);
  e[0] = (float)std::atan2(sinr, cosr);
  const double sinp = +2.0 * (qw * qy - qz * qx);
  if (std::fabs(sinp) >= 1) e[1] = (float)std::copysign(M_PI / 2, sinp);
  else e[1] = (float)std::asin(sinp);
  const double siny = +2.0 * (qw * qz + qx * qy), cosy = +1.0 - 2.0 * (qy * qy + qz * qz);
  e[2] = (float)std::atan2(siny, cosy);
  for (int d = 0; d < 3; ++d) e[d] = (float)(e[d] * 180.0 / M_PI);
  for (int d = 0; d < 3; ++d) {
    e[d] = std::fabs(e[d]);
    if (sym_info(d) == 90) {
      e[d] = std::fabs(e[d] - 90);
      e[d] = std::min(e[d], 90 - e[d]);
    } else if (sym_info(d) == 180) {
      e[d] = std::min(e[d], 180 - e[d]);
    } else if (sym_info(d) == 360) {
      e[d] = 0;
    }
  }
  mean_rotation_error = std::max(std::max(e[0], e[1]), e[2]);
  translation_error = (float)std::sqrt(std::pow(base_pose(0, 3) - test_pose(0, 3), 2) + std::pow(base_pose(1, 3) - test_pose(1, 3), 2) +
                                       std::pow(base_pose(2, 3) - test_pose(2, 3), 2));
}

// src/pose_clustering.cpp:79-122
void greedy_clustering(std::vector<PoseCandidate*>& hypotheses_set, float acceptable_fraction, float best_score,
                       int maximum_pose_count, float min_distance, float min_angle, Eigen::Vector3f sym_info,
                       std::vector<PoseCandidate*>& clustered_hypotheses_set) {
  clustered_hypotheses_set.clear();
  std::vector<PoseCandidate*> pruned;
  for (auto pose_it : hypotheses_set)
    if (pose_it->lcp > acceptable_fraction * best_score) pruned.push_back(pose_it);
  // std::sort in the reference (order among equal scores unspecified); stable here
  std::stable_sort(pruned.begin(), pruned.end(), [](PoseCandidate* a, PoseCandidate* b) { return a->lcp > b->lcp; });
  for (auto candidate_it : pruned) {
    bool inValid = false;
    for (auto cluster_it : clustered_hypotheses_set) {
      float mean_rotation_error, translation_error;
      get_pose_diff(candidate_it->transform, cluster_it->transform, sym_info, mean_rotation_error, translation_error);
      if (mean_rotation_error < min_angle && translation_error < min_distance) { inValid = true; break; }
    }
    if (inValid == false) clustered_hypotheses_set.push_back(candidate_it);
    if ((int)clustered_hypotheses_set.size() > maximum_pose_count) break;
  }
}

// reference src/pose_clustering.cpp:123-141.  PCL's ICP loop runs on the device
// (stocs_b200_icp_point_to_plane); the context is created on first use and lives for the process.
void point_to_plane_icp(PCLPointCloud::Ptr segment_cloud, PCLPointCloud::Ptr model_cloud, Eigen::Matrix4f& offset_transform) {
  static stocs_b200_ctx* ctx = nullptr;
  if (!ctx) {
    const char* dev = getenv("STOCS_DEVICE");
    if (stocs_b200_create(&ctx, dev ? atoi(dev) : 0) != 0) {
      ctx = nullptr;  // create leaves *out NULL on failure; its message is kept per thread
      throw std::runtime_error(std::string("point_to_plane_icp: ") + stocs_b200_last_error(nullptr));
    }
  }
  const int ns = (int)segment_cloud->points.size(), nt = (int)model_cloud->points.size();
  // PCL's align() on an empty source or target does not converge; the reference then resets the
  // caller's matrix (src/pose_clustering.cpp:136-139)
  if (ns == 0 || nt == 0) { offset_transform.setIdentity(); return; }
  std::vector<float> sp((size_t)ns * 3), tp((size_t)nt * 3), tn((size_t)nt * 3), moved((size_t)ns * 3);
  for (int i = 0; i < ns; ++i) {
    const CloudPoint& p = segment_cloud->points[i];
    sp[3 * i] = p.x; sp[3 * i + 1] = p.y; sp[3 * i + 2] = p.z;
  }
  for (int i = 0; i < nt; ++i) {
    const CloudPoint& p = model_cloud->points[i];
    tp[3 * i] = p.x; tp[3 * i + 1] = p.y; tp[3 * i + 2] = p.z;
    tn[3 * i] = p.nx; tn[3 * i + 1] = p.ny; tn[3 * i + 2] = p.nz;
  }
  float T[16];
  int32_t converged = 0;
  if (stocs_b200_icp_point_to_plane(ctx, sp.data(), ns, tp.data(), tn.data(), nt, /*setMaximumIterations*/ 5,
                                    /*setMaxCorrespondenceDistance*/ 0.035f, T, moved.data(), nullptr, nullptr,
                                    &converged) != 0)
    throw std::runtime_error(std::string("point_to_plane_icp: ") + stocs_b200_last_error(ctx));
  for (int i = 0; i < ns; ++i) {  // icp->align(*segment_cloud) leaves the moved source in place
    CloudPoint& p = segment_cloud->points[i];
    p.x = moved[3 * i]; p.y = moved[3 * i + 1]; p.z = moved[3 * i + 2];
    const float nx = p.nx, ny = p.ny, nz = p.nz;
    p.nx = T[0] * nx + T[4] * ny + T[8] * nz;
    p.ny = T[1] * nx + T[5] * ny + T[9] * nz;
    p.nz = T[2] * nx + T[6] * ny + T[10] * nz;
  }
  // src/pose_clustering.cpp:136-139: final transformation when converged, identity otherwise
  if (converged) {
    for (int c = 0; c < 4; ++c)
      for (int r = 0; r < 4; ++r) offset_transform(r, c) = T[c * 4 + r];
  } else {
    offset_transform.setIdentity();
  }
}

}  // namespace clustering
