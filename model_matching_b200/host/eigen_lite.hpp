// eigen_lite.hpp -- the few Eigen types that appear in the public interface of
// stocs::stocs_estimator (reference include/stocs.hpp:8-10, include/point3d.hpp:13-14,141-156),
// for builds without Eigen (it is not installed in this image).  Define STOCS_USE_EIGEN to
// compile the shim against the real library instead; the memory layouts are identical
// (column-major fixed-size float matrices).
#pragma once
#ifdef STOCS_USE_EIGEN
#include <Eigen/Core>
#include <Eigen/Dense>
#else
#include <cmath>
#include <cstring>
#include <ostream>
#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW
namespace Eigen {

template <typename Scalar, int Rows, int Cols>
struct Matrix {
  Scalar m[Rows * Cols];  // column-major
  Matrix() { for (int i = 0; i < Rows * Cols; ++i) m[i] = Scalar(0); }
  Matrix(Scalar x, Scalar y, Scalar z) { static_assert(Rows * Cols == 3, "3-vector only"); m[0] = x; m[1] = y; m[2] = z; }
  static Matrix Zero() { return Matrix(); }
  void setIdentity() { *this = Identity(); }
  static Matrix Identity() { Matrix r; for (int i = 0; i < (Rows < Cols ? Rows : Cols); ++i) r(i, i) = Scalar(1); return r; }
  Scalar& operator()(int r, int c) { return m[c * Rows + r]; }
  const Scalar& operator()(int r, int c) const { return m[c * Rows + r]; }
  Scalar& operator()(int i) { return m[i]; }
  const Scalar& operator()(int i) const { return m[i]; }
  Scalar& operator[](int i) { return m[i]; }
  const Scalar& operator[](int i) const { return m[i]; }
  Scalar& coeffRef(int i) { return m[i]; }
  Scalar coeff(int i) const { return m[i]; }
  Scalar* data() { return m; }
  const Scalar* data() const { return m; }
  Scalar x() const { return m[0]; }
  Scalar y() const { return m[1]; }
  Scalar z() const { return m[2]; }
  Matrix operator-(const Matrix& o) const { Matrix r; for (int i = 0; i < Rows * Cols; ++i) r.m[i] = m[i] - o.m[i]; return r; }
  Matrix operator+(const Matrix& o) const { Matrix r; for (int i = 0; i < Rows * Cols; ++i) r.m[i] = m[i] + o.m[i]; return r; }
  Matrix& operator-=(const Matrix& o) { for (int i = 0; i < Rows * Cols; ++i) m[i] -= o.m[i]; return *this; }
  Matrix& operator+=(const Matrix& o) { for (int i = 0; i < Rows * Cols; ++i) m[i] += o.m[i]; return *this; }
  // 3-vector helpers in Eigen's evaluation order (redux a + (b + c)), see csrc/stocs_math.h
  Scalar squaredNorm() const { static_assert(Rows * Cols == 3, "3-vector only"); return m[0] * m[0] + (m[1] * m[1] + m[2] * m[2]); }
  Scalar norm() const { return std::sqrt(squaredNorm()); }
  Scalar dot(const Matrix& o) const { static_assert(Rows * Cols == 3, "3-vector only"); return m[0] * o.m[0] + (m[1] * o.m[1] + m[2] * o.m[2]); }
  Matrix normalized() const { Matrix r = *this; Scalar z = squaredNorm(); if (z > Scalar(0)) { Scalar n = std::sqrt(z); for (int i = 0; i < 3; ++i) r.m[i] = m[i] / n; } return r; }
  void normalize() { *this = normalized(); }
};
using Matrix4f = Matrix<float, 4, 4>;
using Matrix3f = Matrix<float, 3, 3>;
using Vector3f = Matrix<float, 3, 1>;
template <typename T> using Ref = T;  // const Eigen::Ref<const MatrixType>& == const MatrixType&

template <typename S, int R, int C>
std::ostream& operator<<(std::ostream& os, const Matrix<S, R, C>& a) {
  for (int r = 0; r < R; ++r) { for (int c = 0; c < C; ++c) os << (c ? " " : "") << a(r, c); if (r + 1 < R) os << "\n"; }
  return os;
}
}  // namespace Eigen
#endif
