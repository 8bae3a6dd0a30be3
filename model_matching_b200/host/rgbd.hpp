// rgbd.hpp -- namespace rgbd of the drop-in host layer: same free functions as the reference's
// include/rgbd.hpp:30-102, without PCL / OpenCV / Boost.
//
// PPFMapType: the reference's type is std::map<vector<int>, vector<pair<int,int>>> holding every
// ordered model pair under up to 128 keys (25-114 M entries).  Here it is the COMPACT table
// (each ordered pair once, under its own bin) that libstocs_b200 consumes; size() still reports
// the number of keys of the reference's expanded map, which is what the reference prints.
#ifndef STOCS_B200_RGBD_HPP_
#define STOCS_B200_RGBD_HPP_
// Standard headers that the reference's header chain (point3d.hpp, PCL, OpenCV, Boost) exposes and
// that its callers rely on without including them (src/stocs_match_one_object.cpp uses struct stat,
// std::ofstream, std::cout, std::random_shuffle, system): kept so that those callers compile unchanged.
#include <sys/stat.h>

#include <algorithm>
#include <array>
#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "point3d.hpp"

using micro = std::chrono::microseconds;
struct stocs_b200_ctx;  // libstocs_b200 context (include/stocs_b200.h)

struct PPFMapType {
  int tr_discretization = 0, rot_discretization = 0;
  int num_model_points = 0;
  int64_t expanded_keys = 0;            // == reference ppf_map.size()
  std::vector<int32_t> keys4;           // own-bin key of every stored ordered pair
  std::vector<int32_t> pairs2;          // (id1, id2), sorted by (key, id1, id2)
  size_t size() const { return (size_t)expanded_keys; }
  bool empty() const { return pairs2.empty(); }
};

// organised RGB-D cloud stand-in for pcl::PointCloud<PointXYZRGBNormal>
struct CloudPoint { float x, y, z, nx, ny, nz; float r, g, b; };
struct PCLPointCloud {
  std::vector<CloudPoint> points;
  using Ptr = std::shared_ptr<PCLPointCloud>;
};

namespace rgbd {

// a1 + scene-cloud construction (reference src/rgbd.cpp:179-281): runs on the GPU through
// stocs_b200_build_scene_cloud; the host decodes the PNG files.
void load_rgbd_data_sampled(std::string rgb_location, std::string depth_location,
                            std::string class_probability_map_location,
                            const std::vector<uint8_t>& edge_probability_map, int edge_w, int edge_h,
                            std::vector<float> camera_intrinsics, float depth_scale, float voxel_size,
                            float class_probability_threshold, std::vector<Point3D>& point3d,
                            stocs_b200_ctx* ctx);

// STOCS_PATH_REMAP="from=to[;from2=to2]": every path the shim opens has a leading `from` replaced by
// `to`.  Lets a caller with a compiled-in repository path (the reference's CLIs,
// src/stocs_match_one_object.cpp:4) run against a tree that lives elsewhere, without editing it.
std::string remap_path(const std::string& path);

bool load_ply_file(const std::string& location, PCLPointCloud& cloud);
void load_ply_model(PCLPointCloud::Ptr cloud, std::vector<Point3D>& point3d, float scale);
void save_as_ply(std::string location, std::vector<Point3D>& point3d, float scale);
void transform_pointset(std::vector<Point3D>& input, std::vector<Point3D>& output,
                        Eigen::Matrix<Point3D::Scalar, 4, 4>& transform);
void compute_normal_pcl(PCLPointCloud::Ptr cloud, float radius);
void voxel_grid_filter(PCLPointCloud& cloud, float leaf);
void ppf_compute(Point3D point_1, Point3D point_2, float tr_discretization, float rot_discretization,
                 std::vector<int>& ppf_);
void save_ppf_map(std::string location, PPFMapType& ppf_map);
void load_ppf_map(std::string ppf_map_location, PPFMapType& ppf_map);

}  // namespace rgbd
#endif
