// stocs_scene_math.h -- leaf arithmetic of the scene-cloud construction (reference
// src/rgbd.cpp:190-279), shared by the CUDA kernels (csrc/scene_cloud.cu) and the CPU oracle so
// that both produce bit-identical clouds.  Same rules as stocs_math.h: only IEEE
// add/sub/mul/div/sqrt in a fixed order, no FMA contraction.
#pragma once
#include "stocs_math.h"

namespace stocsm {

// Smallest-eigenvalue eigenvector of a symmetric 3x3 matrix by cyclic Jacobi sweeps (binary64).
STOCS_HD void smallest_eigenvector3(double a[3][3], double v[3]) {
  double V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int sweep = 0; sweep < 24; ++sweep) {
    const double off = fabs(a[0][1]) + fabs(a[0][2]) + fabs(a[1][2]);
    if (off < 1e-30) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        if (fabs(a[p][q]) < 1e-300) continue;
        const double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; ++k) { const double akp = a[k][p], akq = a[k][q]; a[k][p] = c * akp - s * akq; a[k][q] = s * akp + c * akq; }
        for (int k = 0; k < 3; ++k) { const double apk = a[p][k], aqk = a[q][k]; a[p][k] = c * apk - s * aqk; a[q][k] = s * apk + c * aqk; }
        for (int k = 0; k < 3; ++k) { const double vkp = V[k][p], vkq = V[k][q]; V[k][p] = c * vkp - s * vkq; V[k][q] = s * vkp + c * vkq; }
      }
  }
  int m = 0;
  if (a[1][1] < a[m][m]) m = 1;
  if (a[2][2] < a[m][m]) m = 2;
  for (int k = 0; k < 3; ++k) v[k] = V[k][m];
}

// Depth-image normal at one pixel: stand-in for cv::rgbd::RgbdNormals(rows, cols, CV_32F, K, 5,
// RGBD_NORMALS_METHOD_LINEMOD) (reference src/rgbd.cpp:202-206; opencv_contrib's source is not in
// the reference tree).  Least-squares plane through the back-projected points of a 9x9 window
// (stride 2) that lie on the same surface as the centre pixel (depth within 2 % + 5 mm), oriented
// towards the camera.  Zero vector = invalid, which the caller treats as the reference treats an
// all-zero normal (src/rgbd.cpp:266).  xyz: organised H*W*3 cloud of the back-projection.
STOCS_HD void depth_normal_at(const float* xyz, int W, int H, int row, int col, float n_out[3]) {
  n_out[0] = n_out[1] = n_out[2] = 0.f;
  const float* c = xyz + 3 * ((size_t)row * W + col);
  if (!(c[2] > 0)) return;
  const float tol = 0.02f * c[2] + 0.005f;
  double m[3] = {0, 0, 0};
  float pts[25][3];
  int n = 0;
  for (int di = -4; di <= 4; di += 2)
    for (int dj = -4; dj <= 4; dj += 2) {
      const int i = row + di, j = col + dj;
      if (i < 0 || i >= H || j < 0 || j >= W) continue;
      const float* p = xyz + 3 * ((size_t)i * W + j);
      if (!(p[2] > 0) || fabsf(p[2] - c[2]) > tol) continue;
      pts[n][0] = p[0]; pts[n][1] = p[1]; pts[n][2] = p[2];
      m[0] += p[0]; m[1] += p[1]; m[2] += p[2];
      ++n;
    }
  if (n < 8) return;
  for (int k = 0; k < 3; ++k) m[k] /= n;
  double cov[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  for (int t = 0; t < n; ++t) {
    const double d[3] = {pts[t][0] - m[0], pts[t][1] - m[1], pts[t][2] - m[2]};
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) cov[a][b] += d[a] * d[b];
  }
  double v[3];
  smallest_eigenvector3(cov, v);
  if (v[0] * c[0] + v[1] * c[1] + v[2] * c[2] > 0) { v[0] = -v[0]; v[1] = -v[1]; v[2] = -v[2]; }
  n_out[0] = (float)v[0]; n_out[1] = (float)v[1]; n_out[2] = (float)v[2];
}

// pcl::VoxelGrid leaf coordinates of a point (floor(p * inverse_leaf_size))
STOCS_HD void voxel_coords(float x, float y, float z, float inv_leaf, long long ijk[3]) {
  ijk[0] = (long long)floorf(x * inv_leaf);
  ijk[1] = (long long)floorf(y * inv_leaf);
  ijk[2] = (long long)floorf(z * inv_leaf);
}

// Re-projection of a voxel centroid to (row, col) (reference src/rgbd.cpp:245-252):
// point2D = K * p with Eigen's 3-term order a + (b + c); int truncation of u/z, v/z.
STOCS_HD void reproject(float x, float y, float z, float fx, float cx, float fy, float cy, int* row, int* col) {
  const float u = fx * x + (0.0f * y + cx * z);
  const float v = 0.0f * x + (fy * y + cy * z);
  const float w = 0.0f * x + (0.0f * y + 1.0f * z);
  *col = (int)(u / w);
  *row = (int)(v / w);
}

}  // namespace stocsm
