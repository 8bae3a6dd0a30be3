// point3d.hpp -- point / hypothesis records of the estimator interface.
// Same public surface as the reference's include/point3d.hpp:11-156 (Point3D accessors and
// mutators, Quadrilateral with lexicographic <, PoseCandidate{transform, lcp, base_index}).
#ifndef STOCS_B200_POINT3D_HPP_
#define STOCS_B200_POINT3D_HPP_
#include <array>
#include <utility>
#include <vector>

#include "eigen_lite.hpp"

class Point3D {
 public:
  using Scalar = float;
  using VectorType = Eigen::Matrix<Scalar, 3, 1>;

  Point3D() {}
  Point3D(Scalar x, Scalar y, Scalar z) : pos_(x, y, z) {}
  explicit Point3D(const VectorType& p) : pos_(p) {}

  VectorType& pos() { return pos_; }
  const VectorType& pos() const { return pos_; }
  const VectorType& rgb() const { return rgb_; }
  const VectorType& normal() const { return normal_; }
  const float& probability() const { return current_probability_; }
  const float& class_probability() const { return class_probability_; }
  float edge_probability() const { return edge_probability_; }
  const std::pair<int, int>& pixel() const { return pixel_; }  // (row, col)

  void set_rgb(const VectorType& rgb) { rgb_ = rgb; }
  void set_normal(const VectorType& n) { normal_ = n.normalized(); }
  // addition: store a normal that is already the result of set_normal (GPU-built scene cloud)
  void set_unit_normal(const VectorType& n) { normal_ = n; }
  void set_pixel(const std::pair<int, int>& p) { pixel_ = p; }
  void set_probability(float class_probability, float edge_probability) {
    class_probability_ = class_probability;
    edge_probability_ = edge_probability;
    current_probability_ = class_probability;
  }
  void update_class_probability(float decay_fraction) { class_probability_ = decay_fraction * class_probability_; }
  void update_probability(float p) { current_probability_ = p; }
  void reset_probability() { current_probability_ = class_probability_; }
  void normalize() { pos_.normalize(); }
  bool hasColor() const { return rgb_.squaredNorm() > Scalar(0.001); }

  Scalar& x() { return pos_.coeffRef(0); }
  Scalar& y() { return pos_.coeffRef(1); }
  Scalar& z() { return pos_.coeffRef(2); }
  Scalar x() const { return pos_.coeff(0); }
  Scalar y() const { return pos_.coeff(1); }
  Scalar z() const { return pos_.coeff(2); }

 private:
  VectorType pos_{0.0f, 0.0f, 0.0f};
  VectorType normal_{0.0f, 0.0f, 0.0f};
  VectorType rgb_{-1.0f, -1.0f, -1.0f};
  std::pair<int, int> pixel_{0, 0};
  float class_probability_ = 0;
  float edge_probability_ = 0;
  float current_probability_ = 0;
};

// Four model indices congruent to a scene base.
struct Quadrilateral {
  std::array<int, 4> vertices;
  Quadrilateral(int v0, int v1, int v2, int v3) { vertices = {v0, v1, v2, v3}; }
  bool operator<(const Quadrilateral& rhs) const { return vertices < rhs.vertices; }
  bool operator==(const Quadrilateral& rhs) const { return vertices == rhs.vertices; }
  int operator[](int idx) const { return vertices[idx]; }
  int& operator[](int idx) { return vertices[idx]; }
};

class PoseCandidate {
 public:
  EIGEN_MAKE_ALIGNED_OPERATOR_NEW
  Eigen::Matrix4f transform;
  float lcp;
  int base_index;
  PoseCandidate(Eigen::Matrix4f transform, float lcp, float base_index) {
    this->transform = transform;
    this->lcp = lcp;
    this->base_index = (int)base_index;
  }
  ~PoseCandidate() {}
};
#endif
