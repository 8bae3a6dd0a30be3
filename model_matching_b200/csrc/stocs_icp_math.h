// stocs_icp_math.h -- leaf arithmetic of the point-to-plane ICP refinement (reference
// src/pose_clustering.cpp:123-141), shared by the CUDA kernels (csrc/icp.cu) and the CPU oracle.
//
// The reference delegates the whole loop to pcl::IterativeClosestPointWithNormals (PCL is an
// external dependency, not in the reference tree, version not pinned by its CMakeLists).  What is
// restated here is PCL's published algorithm for that class:
//   * correspondences: nearest target point of every (current) source point, kept when the
//     squared distance is <= max_correspondence_distance^2;
//   * TransformationEstimationPointToPlaneLLS: the small-angle linearisation.  Per pair
//     (s, d, n): a = (s x n), row = [a, n], rhs = n.(d - s); binary64 normal equations
//     ATA x = ATb, x = (alpha, beta, gamma, tx, ty, tz); R = Rz(gamma) Ry(beta) Rx(alpha);
//   * the source is moved by each step's 4x4 float matrix, final = step * final.
// Same rules as stocs_math.h: IEEE operations in a fixed order, no FMA contraction.
#pragma once
#include "stocs_math.h"

namespace stocsm {

constexpr int kIcpTerms = 29;  // 21 upper-triangle ATA + 6 ATb + pair count + sum of squared distances
constexpr int kIcpBlock = 256; // source points per reduction block (fixes the summation order)

// Contribution of one correspondence.  s: current source point, d: target point, n: target normal.
STOCS_HD void icp_pair_terms(V3 s, V3 d, V3 n, float d2, double* t) {
  const double sx = s.x, sy = s.y, sz = s.z, dx = d.x, dy = d.y, dz = d.z, nx = n.x, ny = n.y, nz = n.z;
  double r[6];
  r[0] = nz * sy - ny * sz;
  r[1] = nx * sz - nz * sx;
  r[2] = ny * sx - nx * sy;
  r[3] = nx; r[4] = ny; r[5] = nz;
  const double e = ((nx * dx + ny * dy) + nz * dz) - ((nx * sx + ny * sy) + nz * sz);
  int k = 0;
  for (int i = 0; i < 6; ++i)
    for (int j = i; j < 6; ++j) t[k++] = r[i] * r[j];
  for (int i = 0; i < 6; ++i) t[k++] = r[i] * e;
  t[k++] = 1.0;
  t[k++] = (double)d2;
}

// Solve the 6x6 symmetric system given as upper triangle (row-major, 21 values) and rhs; Gaussian
// elimination with partial pivoting in binary64.  False when a pivot vanishes.
STOCS_HD bool icp_solve6(const double* upper21, const double* rhs6, double* x) {
  double A[6][7];
  int k = 0;
  for (int i = 0; i < 6; ++i)
    for (int j = i; j < 6; ++j) { A[i][j] = upper21[k]; A[j][i] = upper21[k]; ++k; }
  for (int i = 0; i < 6; ++i) A[i][6] = rhs6[i];
  for (int c = 0; c < 6; ++c) {
    int p = c;
    double best = A[c][c] < 0 ? -A[c][c] : A[c][c];
    for (int r = c + 1; r < 6; ++r) {
      const double v = A[r][c] < 0 ? -A[r][c] : A[r][c];
      if (v > best) { best = v; p = r; }
    }
    if (!(best > 0.0)) return false;
    if (p != c)
      for (int j = 0; j < 7; ++j) { const double tmp = A[c][j]; A[c][j] = A[p][j]; A[p][j] = tmp; }
    for (int r = c + 1; r < 6; ++r) {
      const double f = A[r][c] / A[c][c];
      for (int j = c; j < 7; ++j) A[r][j] = A[r][j] - f * A[c][j];
    }
  }
  for (int i = 5; i >= 0; --i) {
    double s = A[i][6];
    for (int j = i + 1; j < 6; ++j) s = s - A[i][j] * x[j];
    x[i] = s / A[i][i];
  }
  for (int i = 0; i < 6; ++i)
    if (!(x[i] == x[i]) || x[i] > 1e300 || x[i] < -1e300) return false;
  return true;
}

// (alpha, beta, gamma, tx, ty, tz) -> column-major 4x4 float, R = Rz(gamma) Ry(beta) Rx(alpha)
STOCS_HD void icp_construct(const double* x, float* m) {
  double sa, ca, sb, cb, sg, cg;
  sincos_d(x[0], &sa, &ca);
  sincos_d(x[1], &sb, &cb);
  sincos_d(x[2], &sg, &cg);
  m[0] = (float)(cg * cb);
  m[4] = (float)(-sg * ca + (cg * sb) * sa);
  m[8] = (float)(sg * sa + (cg * sb) * ca);
  m[1] = (float)(sg * cb);
  m[5] = (float)(cg * ca + (sg * sb) * sa);
  m[9] = (float)(-cg * sa + (sg * sb) * ca);
  m[2] = (float)(-sb);
  m[6] = (float)(cb * sa);
  m[10] = (float)(cb * ca);
  m[12] = (float)x[3]; m[13] = (float)x[4]; m[14] = (float)x[5];
  m[3] = 0.f; m[7] = 0.f; m[11] = 0.f; m[15] = 1.f;
}

// C = A * B, column-major 4x4 float, terms added in k order
STOCS_HD void mat4_mul(const float* A, const float* B, float* C) {
  float out[16];
  for (int c = 0; c < 4; ++c)
    for (int r = 0; r < 4; ++r)
      out[c * 4 + r] = ((A[r] * B[c * 4] + A[4 + r] * B[c * 4 + 1]) + A[8 + r] * B[c * 4 + 2]) + A[12 + r] * B[c * 4 + 3];
  for (int i = 0; i < 16; ++i) C[i] = out[i];
}

// squared distance in binary32, the order the scoring path uses
STOCS_HD float icp_sqdist(V3 a, V3 b) {
  const float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
  return (dx * dx + dy * dy) + dz * dz;
}

}  // namespace stocsm
