// rgbd.cpp -- host layer around the GPU path: file formats and scene-cloud construction.
#include "rgbd.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <unordered_map>

#include "../../include/stocs_b200.h"
#include "../csrc/stocs_math.h"
#include "image_io.hpp"

namespace rgbd {

std::string remap_path(const std::string& path) {
  const char* e = std::getenv("STOCS_PATH_REMAP");
  if (!e) return path;
  std::stringstream rules(e);
  for (std::string rule; std::getline(rules, rule, ';');) {
    const size_t eq = rule.find('=');
    if (eq == std::string::npos || eq == 0) continue;
    const std::string from = rule.substr(0, eq), to = rule.substr(eq + 1);
    if (path.compare(0, from.size(), from) == 0) return to + path.substr(from.size());
  }
  return path;
}

// ---- PLY -------------------------------------------------------------------------------------
bool load_ply_file(const std::string& location_in, PCLPointCloud& cloud) {
  const std::string location = remap_path(location_in);
  std::ifstream f(location);
  if (!f) return false;
  std::string line;
  std::vector<std::string> props;
  size_t nvert = 0;
  bool in_vertex = false, ascii = false;
  while (std::getline(f, line)) {
    std::istringstream is(line);
    std::string tok;
    is >> tok;
    if (tok == "format") { std::string fmt; is >> fmt; ascii = (fmt == "ascii"); }
    else if (tok == "element") { std::string name; is >> name; in_vertex = (name == "vertex"); if (in_vertex) is >> nvert; }
    else if (tok == "property" && in_vertex) { std::string type, name; is >> type >> name; props.push_back(name); }
    else if (tok == "end_header") break;
  }
  if (!ascii) { std::cerr << "load_ply_file: only ASCII PLY is supported: " << location << std::endl; return false; }
  auto col = [&](const char* n) { for (size_t i = 0; i < props.size(); ++i) if (props[i] == n) return (int)i; return -1; };
  const int cx = col("x"), cy = col("y"), cz = col("z"), cnx = col("nx"), cny = col("ny"), cnz = col("nz");
  int cr = col("red"), cg = col("green"), cb = col("blue");
  if (cr < 0) { cr = col("r"); cg = col("g"); cb = col("b"); }
  cloud.points.clear();
  cloud.points.reserve(nvert);
  std::vector<double> v(props.size());
  for (size_t i = 0; i < nvert; ++i) {
    if (!std::getline(f, line)) break;
    std::istringstream is(line);
    for (size_t k = 0; k < props.size(); ++k) { std::string t; is >> t; v[k] = (t == "nan" || t == "-nan") ? NAN : atof(t.c_str()); }
    CloudPoint p{};
    p.x = (float)v[cx]; p.y = (float)v[cy]; p.z = (float)v[cz];
    p.nx = cnx >= 0 ? (float)v[cnx] : NAN; p.ny = cny >= 0 ? (float)v[cny] : NAN; p.nz = cnz >= 0 ? (float)v[cnz] : NAN;
    p.r = cr >= 0 ? (float)v[cr] : 0; p.g = cg >= 0 ? (float)v[cg] : 0; p.b = cb >= 0 ? (float)v[cb] : 0;
    cloud.points.push_back(p);
  }
  return true;
}

// reference src/rgbd.cpp:12-33
void load_ply_model(PCLPointCloud::Ptr cloud, std::vector<Point3D>& point3d, float scale) {
  for (auto v : cloud->points) {
    if (std::isfinite(v.nx) && std::isfinite(v.ny) && std::isfinite(v.nz)) {
      point3d.emplace_back(v.x * scale, v.y * scale, v.z * scale);
      point3d.back().set_normal(Point3D::VectorType(v.nx, v.ny, v.nz));
      point3d.back().set_rgb(Point3D::VectorType(v.r, v.g, v.b));
    }
  }
}

// reference src/rgbd.cpp:35-56; ASCII PointXYZRGBNormal layout as pcl::io::savePLYFile writes it.
// Floats are printed with 9 significant digits so that a write/read round trip is exact.
void save_as_ply(std::string location, std::vector<Point3D>& point3d, float scale) {
  location = remap_path(location);
  FILE* f = fopen(location.c_str(), "w");
  if (!f) { std::cerr << "save_as_ply: cannot write " << location << std::endl; return; }
  fprintf(f, "ply\nformat ascii 1.0\ncomment PCL generated\nelement vertex %zu\n", point3d.size());
  fprintf(f, "property float x\nproperty float y\nproperty float z\nproperty uchar red\nproperty uchar green\n"
             "property uchar blue\nproperty float nx\nproperty float ny\nproperty float nz\nproperty float curvature\n");
  fprintf(f, "element camera 1\nproperty float view_px\nproperty float view_py\nproperty float view_pz\n"
             "property float x_axisx\nproperty float x_axisy\nproperty float x_axisz\nproperty float y_axisx\n"
             "property float y_axisy\nproperty float y_axisz\nproperty float z_axisx\nproperty float z_axisy\n"
             "property float z_axisz\nproperty float focal\nproperty float scalex\nproperty float scaley\n"
             "property float centerx\nproperty float centery\nproperty int viewportx\nproperty int viewporty\n"
             "property float k1\nproperty float k2\nend_header\n");
  for (auto& v : point3d) {
    auto c8 = [](float c) { int i = (int)c; return i < 0 ? 0 : (i > 255 ? 255 : i); };
    fprintf(f, "%.9g %.9g %.9g %d %d %d %.9g %.9g %.9g 0\n", v.x() * scale, v.y() * scale, v.z() * scale, c8(v.rgb()[0]),
            c8(v.rgb()[1]), c8(v.rgb()[2]), v.normal()[0], v.normal()[1], v.normal()[2]);
  }
  fprintf(f, "0 0 0 1 0 0 0 1 0 0 0 1 0 0 0 0 0 %zu 1 0 0\n", point3d.size());
  fclose(f);
}

// reference src/rgbd.cpp:58-70
void transform_pointset(std::vector<Point3D>& input, std::vector<Point3D>& output,
                        Eigen::Matrix<Point3D::Scalar, 4, 4>& transform) {
  for (size_t i = 0; i < input.size(); ++i) {
    stocsm::V3 q = stocsm::xform_point(transform.data(), stocsm::v3(input[i].x(), input[i].y(), input[i].z()));
    output.push_back(Point3D(q.x, q.y, q.z));
  }
}

// reference src/rgbd.cpp:99-121 (the arithmetic is the shared, bit-pinned stocs_math.h version)
void ppf_compute(Point3D p1, Point3D p2, float tr, float rot, std::vector<int>& ppf_) {
  stocsm::Ppf4 f = stocsm::ppf_compute(stocsm::v3(p1.x(), p1.y(), p1.z()),
                                       stocsm::v3(p1.normal()[0], p1.normal()[1], p1.normal()[2]),
                                       stocsm::v3(p2.x(), p2.y(), p2.z()),
                                       stocsm::v3(p2.normal()[0], p2.normal()[1], p2.normal()[2]), (int)tr, (int)rot);
  for (int k = 0; k < 4; ++k) ppf_.push_back(f.f[k]);
}

// ---- compact PPF table file ("ppf_map"), replaces the Boost archive of src/rgbd.cpp:156-177 ----
static const char kMagic[8] = {'S', 'T', 'O', 'C', 'S', 'P', 'F', '1'};
void save_ppf_map(std::string location, PPFMapType& m) {
  location = remap_path(location);
  std::ofstream f(location, std::ios::binary);
  if (f.fail()) return;
  int64_t n = (int64_t)(m.pairs2.size() / 2);
  f.write(kMagic, 8);
  int32_t hdr[4] = {m.tr_discretization, m.rot_discretization, m.num_model_points, 0};
  f.write((const char*)hdr, sizeof(hdr));
  f.write((const char*)&m.expanded_keys, 8);
  f.write((const char*)&n, 8);
  f.write((const char*)m.keys4.data(), (std::streamsize)(n * 16));
  f.write((const char*)m.pairs2.data(), (std::streamsize)(n * 8));
}
void load_ppf_map(std::string location, PPFMapType& m) {
  location = remap_path(location);
  std::ifstream f(location, std::ios::binary);
  m = PPFMapType();
  char magic[8];
  if (!f.read(magic, 8) || memcmp(magic, kMagic, 8) != 0) {
    std::cerr << "load_ppf_map: " << location << " is not a STOCSPF1 table (run model_preprocess)" << std::endl;
    return;
  }
  int32_t hdr[4];
  int64_t n = 0;
  f.read((char*)hdr, sizeof(hdr));
  f.read((char*)&m.expanded_keys, 8);
  f.read((char*)&n, 8);
  m.tr_discretization = hdr[0]; m.rot_discretization = hdr[1]; m.num_model_points = hdr[2];
  m.keys4.resize((size_t)n * 4); m.pairs2.resize((size_t)n * 2);
  f.read((char*)m.keys4.data(), (std::streamsize)(n * 16));
  f.read((char*)m.pairs2.data(), (std::streamsize)(n * 8));
}

// ---- PCL operator restatements used by the OFFLINE model preprocessing (row 8f-2) ---------------
// pcl::VoxelGrid: centroid of every occupied leaf, output in increasing leaf index
// (x fastest), all fields averaged.
void voxel_grid_filter(PCLPointCloud& cloud, float leaf) {
  const float inv = 1.0f / leaf;
  float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (auto& p : cloud.points) {
    if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) continue;
    mn[0] = std::min(mn[0], p.x); mn[1] = std::min(mn[1], p.y); mn[2] = std::min(mn[2], p.z);
    mx[0] = std::max(mx[0], p.x); mx[1] = std::max(mx[1], p.y); mx[2] = std::max(mx[2], p.z);
  }
  if (!(mn[0] <= mx[0])) { cloud.points.clear(); return; }
  long long minb[3], divb[3];
  for (int k = 0; k < 3; ++k) { minb[k] = (long long)std::floor(mn[k] * inv); divb[k] = (long long)std::floor(mx[k] * inv) - minb[k] + 1; }
  std::vector<std::pair<long long, uint32_t>> idx;
  idx.reserve(cloud.points.size());
  for (uint32_t i = 0; i < cloud.points.size(); ++i) {
    const auto& p = cloud.points[i];
    if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) continue;
    long long ix = (long long)std::floor(p.x * inv) - minb[0], iy = (long long)std::floor(p.y * inv) - minb[1],
              iz = (long long)std::floor(p.z * inv) - minb[2];
    idx.emplace_back(ix + iy * divb[0] + iz * divb[0] * divb[1], i);
  }
  std::stable_sort(idx.begin(), idx.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
  std::vector<CloudPoint> out;
  size_t i = 0;
  while (i < idx.size()) {
    size_t j = i;
    float s[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    while (j < idx.size() && idx[j].first == idx[i].first) {
      const auto& p = cloud.points[idx[j].second];
      s[0] += p.x; s[1] += p.y; s[2] += p.z; s[3] += p.nx; s[4] += p.ny; s[5] += p.nz; s[6] += p.r; s[7] += p.g; s[8] += p.b;
      ++j;
    }
    const float c = (float)(j - i);
    out.push_back(CloudPoint{s[0] / c, s[1] / c, s[2] / c, s[3] / c, s[4] / c, s[5] / c, s[6] / c, s[7] / c, s[8] / c});
    i = j;
  }
  cloud.points.swap(out);
}

namespace {
struct HashGrid {
  float inv;
  std::unordered_map<uint64_t, std::vector<uint32_t>> cells;
  static uint64_t key(long long x, long long y, long long z) {
    return ((uint64_t)(x & 0x1fffff) << 42) | ((uint64_t)(y & 0x1fffff) << 21) | (uint64_t)(z & 0x1fffff);
  }
  HashGrid(const std::vector<CloudPoint>& pts, float cell) : inv(1.0f / cell) {
    for (uint32_t i = 0; i < pts.size(); ++i)
      cells[key((long long)std::floor(pts[i].x * inv), (long long)std::floor(pts[i].y * inv),
                (long long)std::floor(pts[i].z * inv))].push_back(i);
  }
  template <class Fn> void for_each_near(const CloudPoint& p, Fn fn) const {
    long long cx = (long long)std::floor(p.x * inv), cy = (long long)std::floor(p.y * inv), cz = (long long)std::floor(p.z * inv);
    for (long long z = cz - 1; z <= cz + 1; ++z)
      for (long long y = cy - 1; y <= cy + 1; ++y)
        for (long long x = cx - 1; x <= cx + 1; ++x) {
          auto it = cells.find(key(x, y, z));
          if (it == cells.end()) continue;
          for (uint32_t j : it->second) fn(j);
        }
  }
};

// smallest-eigenvalue eigenvector of a symmetric 3x3 matrix (cyclic Jacobi)
void smallest_eigenvector(double a[3][3], double v[3]) {
  double V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int sweep = 0; sweep < 32; ++sweep) {
    double off = std::fabs(a[0][1]) + std::fabs(a[0][2]) + std::fabs(a[1][2]);
    if (off < 1e-30) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        if (std::fabs(a[p][q]) < 1e-300) continue;
        double theta = (a[q][q] - a[p][p]) / (2 * a[p][q]);
        double t = (theta >= 0 ? 1 : -1) / (std::fabs(theta) + std::sqrt(theta * theta + 1));
        double c = 1 / std::sqrt(t * t + 1), s = t * c;
        for (int k = 0; k < 3; ++k) { double akp = a[k][p], akq = a[k][q]; a[k][p] = c * akp - s * akq; a[k][q] = s * akp + c * akq; }
        for (int k = 0; k < 3; ++k) { double apk = a[p][k], aqk = a[q][k]; a[p][k] = c * apk - s * aqk; a[q][k] = s * apk + c * aqk; }
        for (int k = 0; k < 3; ++k) { double vkp = V[k][p], vkq = V[k][q]; V[k][p] = c * vkp - s * vkq; V[k][q] = s * vkp + c * vkq; }
      }
  }
  int m = 0;
  if (a[1][1] < a[m][m]) m = 1;
  if (a[2][2] < a[m][m]) m = 2;
  for (int k = 0; k < 3; ++k) v[k] = V[k][m];
}
}  // namespace

// pcl::NormalEstimation with a radius search and the default viewpoint (0,0,0): plane fit by PCA
// of the neighbours, normal flipped towards the origin (reference src/rgbd.cpp:72-83).
void compute_normal_pcl(PCLPointCloud::Ptr cloud, float radius) {
  auto& pts = cloud->points;
  HashGrid grid(pts, radius);
  const float r2 = radius * radius;
  for (auto& p : pts) {
    double c[3] = {0, 0, 0};
    std::vector<uint32_t> nb;
    grid.for_each_near(p, [&](uint32_t j) {
      const auto& q = pts[j];
      const float dx = p.x - q.x, dy = p.y - q.y, dz = p.z - q.z;
      if (dx * dx + dy * dy + dz * dz <= r2) nb.push_back(j);
    });
    if (nb.size() < 3) { p.nx = p.ny = p.nz = NAN; continue; }
    for (uint32_t j : nb) { c[0] += pts[j].x; c[1] += pts[j].y; c[2] += pts[j].z; }
    for (int k = 0; k < 3; ++k) c[k] /= (double)nb.size();
    double cov[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for (uint32_t j : nb) {
      const double d[3] = {pts[j].x - c[0], pts[j].y - c[1], pts[j].z - c[2]};
      for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) cov[a][b] += d[a] * d[b];
    }
    double n[3];
    smallest_eigenvector(cov, n);
    if (n[0] * (0 - p.x) + n[1] * (0 - p.y) + n[2] * (0 - p.z) < 0) { n[0] = -n[0]; n[1] = -n[1]; n[2] = -n[2]; }
    p.nx = (float)n[0]; p.ny = (float)n[1]; p.nz = (float)n[2];
  }
}

// reference src/rgbd.cpp:179-281.  The whole body -- back-projection, VoxelGrid,
// RadiusOutlierRemoval, re-projection, class threshold, depth normals -- runs on the GPU
// (stocs_b200_build_scene_cloud, csrc/scene_cloud.cu); the host only decodes the PNGs.
void load_rgbd_data_sampled(std::string rgb_location, std::string depth_location, std::string class_probability_map_location,
                            const std::vector<uint8_t>& edge_map, int edge_w, int edge_h, std::vector<float> K,
                            float depth_scale, float voxel_size, float class_probability_threshold,
                            std::vector<Point3D>& point3d, stocs_b200_ctx* ctx) {
  std::vector<uint8_t> bgr;
  std::vector<uint16_t> depth, prob;
  int W = 0, H = 0, w2 = 0, h2 = 0;
  rgb_location = remap_path(rgb_location);
  depth_location = remap_path(depth_location);
  class_probability_map_location = remap_path(class_probability_map_location);
  if (!imgio::load_bgr8(rgb_location, bgr, W, H) || !imgio::load_gray16(depth_location, depth, w2, h2) || w2 != W || h2 != H) {
    std::cerr << "load_rgbd_data_sampled: cannot read " << rgb_location << " / " << depth_location << std::endl;
    return;
  }
  if (!imgio::load_gray16(class_probability_map_location, prob, w2, h2) || w2 != W || h2 != H) {
    std::cerr << "load_rgbd_data_sampled: cannot read " << class_probability_map_location << std::endl;
    return;
  }
  const bool have_edge = !edge_map.empty() && edge_w == W && edge_h == H;
  const int64_t cap = (int64_t)W * H;
  std::vector<float> pos((size_t)cap * 3), nrm((size_t)cap * 3), rgb((size_t)cap * 3), cls((size_t)cap), ep((size_t)cap);
  std::vector<int32_t> pix((size_t)cap * 2);
  int64_t n = 0;
  int rc = stocs_b200_build_scene_cloud(ctx, depth.data(), bgr.data(), prob.data(), have_edge ? edge_map.data() : nullptr, W, H,
                                        K[0], K[1], K[2], K[3], depth_scale, voxel_size, class_probability_threshold, pos.data(),
                                        nrm.data(), rgb.data(), pix.data(), cls.data(), ep.data(), cap, &n);
  if (rc != 0) { std::cerr << "stocs_b200_build_scene_cloud: " << stocs_b200_last_error(ctx) << std::endl; return; }
  point3d.reserve(point3d.size() + (size_t)n);
  for (int64_t i = 0; i < n; ++i) {
    point3d.emplace_back(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
    point3d.back().set_unit_normal(Point3D::VectorType(nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]));  // normalised on the device
    point3d.back().set_rgb(Point3D::VectorType(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]));
    point3d.back().set_pixel(std::make_pair((int)pix[2 * i], (int)pix[2 * i + 1]));
    point3d.back().set_probability(cls[i], ep[i]);
  }
}

}  // namespace rgbd
