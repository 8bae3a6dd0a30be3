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

// Depth-image normal at one pixel: cv::rgbd::RgbdNormals(rows, cols, CV_32F, K, 5,
// RGBD_NORMALS_METHOD_LINEMOD) applied to the raw CV_16U depth image (reference src/rgbd.cpp:202-206).
// opencv_contrib is NOT in the reference tree (nor in this image), so this restates its published
// algorithm -- modules/rgbd/src/normal.cpp, LINEMOD<float>::computeImpl<unsigned short, long>, after
// Hinterstoisser et al., "Gradient Response Maps", PAMI 2012, eq. (8) -- and is UNVERIFIED against the
// library ("parity unpinned"):
//   * pixels with row in [5, rows-6) and col in [5, cols-6) only; every other pixel keeps NaN;
//   * least-squares depth gradient over the 11x11 window: a neighbour takes part when its RAW depth
//     differs from the centre's by at most 50 units ("difference_threshold", integer arithmetic):
//       A0 += i*i, A1 += i*j, A3 += j*j, b0 += i*delta, b1 += j*delta   (i = column, j = row offset)
//       det = A0*A3 - A1*A1,  dx = A3*b0 - A1*b1,  dy = -A1*b0 + A0*b1   (gradient * det, no division)
//   * tangents (X1 - X)*det = K^-1 [d*det + (x+1)*dx, y*dx, dx],  (X2 - X)*det = K^-1 [x*dy, d*det + (y+1)*dy, dy]
//     with K^-1 written out in binary32 (K converted to float first), products and sums left to right;
//   * normal = cross product, scaled by 1.f / sqrt(n0*n0 + n1*n1 + n2*n2) and negated when n2 > 0
//     (towards the camera).  A window with no usable neighbour gives 0 * inf = NaN, as in OpenCV.
// The caller rejects NaN and all-zero normals (src/rgbd.cpp:262-266).
STOCS_HD void linemod_normal_at(const uint16_t* depth, int W, int H, int row, int col, float fx, float cx, float fy,
                                float cy, float n_out[3]) {
  const float qnan = bitsf(0x7fc00000u);
  n_out[0] = n_out[1] = n_out[2] = qnan;
  const int r = 5;
  if (row < r || row >= H - r - 1 || col < r || col >= W - r - 1) return;
  const long long d = (long long)depth[(size_t)row * W + col];
  long long A0 = 0, A1 = 0, A3 = 0, b0 = 0, b1 = 0;
  for (int j = -r; j <= r; ++j) {
    const uint16_t* line = depth + (size_t)(row + j) * W + col;
    for (int i = -r; i <= r; ++i) {
      const long long delta = (long long)line[i] - d;
      if ((delta < 0 ? -delta : delta) > 50) continue;
      A0 += (long long)(i * i); A1 += (long long)(i * j); A3 += (long long)(j * j);
      b0 += (long long)i * delta; b1 += (long long)j * delta;
    }
  }
  const long long det = A0 * A3 - A1 * A1;
  const long long dx = A3 * b0 - A1 * b1;
  const long long dy = -A1 * b0 + A0 * b1;
  // K^-1 "by hand, just for higher accuracy" (skew K(0,1) = 0)
  const float k00 = 1.0f / fx;
  const float k01 = -0.0f / (fx * fy);
  const float k02 = (0.0f * cy - cx * fy) / (fx * fy);
  const float k11 = 1.0f / fy;
  const float k12 = -cy / fy;
  const float a1 = (float)(d * det + (long long)(col + 1) * dx), b1f = (float)((long long)row * dx), c1 = (float)dx;
  const float a2 = (float)((long long)col * dy), b2f = (float)(d * det + (long long)(row + 1) * dy), c2 = (float)dy;
  const float u0 = (k00 * a1 + k01 * b1f) + k02 * c1, u1 = k11 * b1f + k12 * c1, u2 = c1;
  const float v0 = (k00 * a2 + k01 * b2f) + k02 * c2, v1 = k11 * b2f + k12 * c2, v2 = c2;
  const float n0 = u1 * v2 - u2 * v1, n1 = u2 * v0 - u0 * v2, n2 = u0 * v1 - u1 * v0;
  const float inv = 1.0f / sqrtf((n0 * n0 + n1 * n1) + n2 * n2);
  if (n2 > 0) { n_out[0] = (-n0) * inv; n_out[1] = (-n1) * inv; n_out[2] = (-n2) * inv; }
  else { n_out[0] = n0 * inv; n_out[1] = n1 * inv; n_out[2] = n2 * inv; }
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
