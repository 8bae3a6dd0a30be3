// stocs.hpp -- drop-in replacement of the reference's include/stocs.hpp: the same class
// stocs::stocs_estimator (constructor arguments, 15 public methods, bool/NULL error convention,
// include/stocs.hpp:16-149) and free function stocs::pre_process_model (:182-191), implemented on
// libstocs_b200.so (include/stocs_b200.h).  Callers written against the reference -- e.g.
// run_stocs_estimation in src/stocs_match_one_object.cpp:51-185 -- compile unchanged.
//
// Differences a caller can observe (all documented in DESIGN.md):
//  * PPFMapType is the compact own-bin table, not a std::map (rgbd.hpp).  A non-empty
//    ppf_map_preloaded IS the table used online (uploaded with stocs_b200_upload_ppf_table); one
//    that does not belong to the model or the discretisations is a fatal error, never ignored.
//  * STOCS_DEVICES="0,1,..": compute_best_transform() shards the hypotheses over those GPUs
//    (stocs_b200_group_*: replicas of the scene index, one NCCL all-gather of the best records).
//  * Sampling is keyed by (seed, base number) instead of the wall clock; the seed comes from the
//    environment variable STOCS_SEED when set, else from the clock as in the reference.
//  * sample_class_base draws bases 128 at a time in one launch and find_congruent_sets_on_model
//    batches over the bases handed out so far; the values returned per call are unchanged.
//  * get_rigid_transform_from_congruent_pair queues the (base, quad) pair; transforms are fitted,
//    scored and reduced in one batched GPU pass at compute_best_transform() (or at the first
//    accessor that needs them).  The lists all_transforms / all_pose end up identical.
//  * sample_instance_base keeps its cached segmentation masks in device memory instead of writing
//    and re-reading dbg/seg_mask_<n>.png (src/stocs.cpp:625, src/rgbd.cpp:330); lossless either way.
#ifndef STOCS_B200_STOCS_HPP_
#define STOCS_B200_STOCS_HPP_
#include <cstdint>
#include <string>
#include <vector>

#include "image_io.hpp"
#include "rgbd.hpp"

struct stocs_b200_group;  // libstocs_b200 multi-GPU group (include/stocs_b200.h)

using Scalar = typename Point3D::Scalar;
using MatrixType = Eigen::Matrix<Scalar, 4, 4>;
using VectorType = typename Point3D::VectorType;

static constexpr Scalar kLargeNumber = 1e9;

namespace stocs {

class stocs_estimator {
 public:
  stocs_estimator(std::string model_location, PPFMapType& ppf_map_preloaded, std::string rgb_location,
                  std::string depth_location, std::string class_probability_map_location,
                  std::string edge_probability_map_location, std::string debug_location,
                  std::vector<float> camera_intrinsics, int image_width, int image_height, float read_depth_scale,
                  float write_depth_scale, float voxel_size, float distance_threshold, int ppf_tr_discretization,
                  int ppf_rot_discretization, float edge_threshold, float class_threshold);
  ~stocs_estimator();
  stocs_estimator(const stocs_estimator&) = delete;
  stocs_estimator& operator=(const stocs_estimator&) = delete;

  void load_object_info(std::string model_location, PPFMapType& ppf_map_preloaded);
  void load_scene_info(std::string rgb_location, std::string depth_location,
                       std::string class_probability_map_location, std::string edge_probability_map_location,
                       std::vector<float> camera_intrinsics, float read_depth_scale, float write_depth_scale,
                       float voxel_size, std::string dst_scene_location);

  bool sample_class_base(std::vector<int>& base_indices, float& invariant1, float& invariant2);
  bool sample_instance_base(std::vector<int>& base_indices, float& invariant1, float& invariant2,
                            std::vector<Point3D>& segment, float dispersion, int base_num);
  bool find_congruent_sets_on_model(std::vector<int>& base_indices, float invariant1, float invariant2,
                                    std::vector<Quadrilateral>* quadrilaterals);
  bool get_rigid_transform_from_congruent_pair(std::vector<int>& base_indices, Quadrilateral& congruent_quad,
                                               int base_index);
  Scalar compute_alignment_score_for_rigid_transform(const Eigen::Ref<const MatrixType>& mat);
  void compute_best_transform();
  void kdtree_initialize();
  void centroid_shift();

  VectorType get_scene_centroid() { return centroid_scene_; }
  std::vector<PoseCandidate*> get_pose_candidates() { flush_pending(); return all_pose; }
  Scalar get_best_score() { return best_lcp; }
  PoseCandidate* get_best_pose() {
    if (best_index == -1) return NULL;
    return all_pose[best_index];
  }
  void visualize_best_pose();

 protected:
  void flush_pending();  // fit every queued (base, quad) pair on the GPU
  void fail(const char* where);

  std::vector<Point3D> point3d_scene;
  std::vector<uint8_t> edge_probability_map;  // image_height x image_width, 8-bit
  int edge_w = 0, edge_h = 0;
  VectorType centroid_scene_;

  std::vector<Point3D> point3d_model;
  PPFMapType ppf_map;
  VectorType centroid_model_;

  std::vector<MatrixType> all_transforms;
  std::vector<PoseCandidate*> all_pose;

  std::string debug_location;
  float distance_threshold;
  int ppf_tr_discretization;
  int ppf_rot_discretization;
  float edge_threshold;
  float class_threshold;

  Scalar best_lcp;
  int best_index;
  int image_width, image_height;

  // GPU state
  ::stocs_b200_ctx* ctx_ = nullptr;        // device 0 of group_ when there is one
  ::stocs_b200_group* group_ = nullptr;    // STOCS_DEVICES with more than one entry
  uint64_t seed_ = 0;
  uint32_t next_base_no_ = 0;
  // class-mode prefetch: bases sampled kPrefetch at a time, congruent sets batched over them
  static constexpr int kPrefetch = 128;
  uint32_t cache_first_ = 0;
  std::vector<int32_t> cache_ids_;
  std::vector<float> cache_inv_;
  std::vector<uint8_t> cache_valid_;
  size_t handed_out_ = 0;
  bool cong_ready_ = false;
  std::vector<int> cong_slot_;
  std::vector<int64_t> cong_off_;
  std::vector<int32_t> cong_quads_;
  bool edge_uploaded_ = false;
  bool class_prob_dirty_ = false;
  std::vector<int32_t> pending_bases_, pending_quads_;
  std::vector<int> pending_base_index_;
};

// host half of pre_process_model (src/stocs.cpp:41-60): load, normals, flip, voxel grid, point list
bool load_and_sample_model(std::string src_model_location, float normal_radius, float read_depth_scale, float voxel_size,
                           std::vector<Point3D>& point3d_sampled);
void pre_process_model(std::string src_model_location, float normal_radius, float read_depth_scale,
                       float write_depth_scale, float voxel_size, float ppf_tr_discretization,
                       float ppf_rot_discretization, std::string dst_model_location,
                       std::string dst_ppf_map_location);

}  // namespace stocs
#endif
