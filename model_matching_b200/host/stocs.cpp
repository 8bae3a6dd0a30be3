// stocs.cpp -- stocs::stocs_estimator on libstocs_b200 (see stocs.hpp).  Host code only marshals:
// every online computation of the reference's src/stocs.cpp runs in a CUDA kernel behind the C ABI.
#include "stocs.hpp"

#include <sys/stat.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>

#include "../../include/stocs_b200.h"

namespace stocs {

void stocs_estimator::fail(const char* where) {
  std::cerr << "libstocs_b200: " << where << ": " << stocs_b200_last_error(ctx_);
  if (group_ && *stocs_b200_group_last_error(group_)) std::cerr << " / " << stocs_b200_group_last_error(group_);
  std::cerr << std::endl;
  std::exit(2);  // no CPU fallback: a GPU failure is fatal, loudly
}

stocs_estimator::stocs_estimator(std::string model_location, PPFMapType& ppf_map_preloaded, std::string rgb_location,
                                 std::string depth_location, std::string class_probability_map_location,
                                 std::string edge_probability_map_location, std::string debug_location,
                                 std::vector<float> camera_intrinsics, int image_width, int image_height,
                                 float read_depth_scale, float write_depth_scale, float voxel_size,
                                 float distance_threshold, int ppf_tr_discretization, int ppf_rot_discretization,
                                 float edge_threshold, float class_threshold) {
  this->debug_location = debug_location;
  this->distance_threshold = distance_threshold;
  this->ppf_tr_discretization = ppf_tr_discretization;
  this->ppf_rot_discretization = ppf_rot_discretization;
  this->edge_threshold = edge_threshold;
  this->class_threshold = class_threshold;
  this->image_width = image_width;
  this->image_height = image_height;
  all_transforms.clear();
  all_pose.clear();
  best_lcp = 0;
  best_index = -1;

  // STOCS_DEVICES="0,1,2,3": several GPUs (hypothesis sharding at compute_best_transform);
  // STOCS_DEVICE=k: one GPU; default device 0
  std::vector<int> devices;
  if (const char* list = std::getenv("STOCS_DEVICES")) {
    std::string tok;
    for (const char* p = list;; ++p) {
      if (*p == ',' || *p == 0) { if (!tok.empty()) devices.push_back(std::atoi(tok.c_str())); tok.clear(); if (!*p) break; }
      else tok.push_back(*p);
    }
  }
  int rc;
  if (devices.size() > 1) {
    rc = stocs_b200_group_create(&group_, devices.data(), (int)devices.size());
    if (rc == 0) ctx_ = stocs_b200_group_ctx(group_, 0);
  } else {
    const char* dev = std::getenv("STOCS_DEVICE");
    rc = stocs_b200_create(&ctx_, devices.size() == 1 ? devices[0] : (dev ? std::atoi(dev) : 0));
  }
  if (rc != 0) {
    std::cerr << "libstocs_b200: cannot create a GPU context (" << rc << "): " << stocs_b200_last_error(nullptr)
              << "\nThis build has no CPU path." << std::endl;
    std::exit(2);
  }
  if (group_) {
    if (stocs_b200_group_set_params(group_, distance_threshold, ppf_tr_discretization, ppf_rot_discretization) != 0) fail("set_params");
  } else if (stocs_b200_set_params(ctx_, distance_threshold, ppf_tr_discretization, ppf_rot_discretization) != 0) {
    fail("set_params");
  }
  const char* seed = std::getenv("STOCS_SEED");
  seed_ = seed ? std::strtoull(seed, nullptr, 10)
               : (uint64_t)std::chrono::system_clock::now().time_since_epoch().count();  // src/stocs.cpp:135

  load_object_info(model_location, ppf_map_preloaded);
  load_scene_info(rgb_location, depth_location, class_probability_map_location, edge_probability_map_location,
                  camera_intrinsics, read_depth_scale, write_depth_scale, voxel_size, debug_location + "/sampled_scene.ply");
  // Move the centroid of the point sets to 0 and index the scene (src/stocs.cpp:943-980): both
  // happen on the GPU when the point sets are uploaded.
  centroid_shift();
  kdtree_initialize();
}

stocs_estimator::~stocs_estimator() {
  // PoseCandidate* are left to the caller, as in the reference
  if (group_) stocs_b200_group_destroy(group_);
  else if (ctx_) stocs_b200_destroy(ctx_);
}

void stocs_estimator::load_object_info(std::string model_location, PPFMapType& ppf_map_preloaded) {
  point3d_model.clear();
  PCLPointCloud::Ptr cloud(new PCLPointCloud);
  if (!rgbd::load_ply_file(model_location, *cloud)) std::cerr << "cannot read " << model_location << std::endl;
  rgbd::load_ply_model(cloud, point3d_model, 1.0f);
  ppf_map = ppf_map_preloaded;
  std::cout << "|M| = " << point3d_model.size() << ",  |map(M)| = " << ppf_map.size() << std::endl;
}

void stocs_estimator::load_scene_info(std::string rgb_location, std::string depth_location,
                                      std::string class_probability_map_location,
                                      std::string edge_probability_map_location, std::vector<float> camera_intrinsics,
                                      float read_depth_scale, float write_depth_scale, float voxel_size,
                                      std::string dst_scene_location) {
  point3d_scene.clear();
  struct stat buffer;
  edge_probability_map.clear();
  if (stat(edge_probability_map_location.c_str(), &buffer) == 0)
    imgio::load_gray8(edge_probability_map_location, edge_probability_map, edge_w, edge_h);
  if (edge_probability_map.empty()) {
    edge_w = image_width; edge_h = image_height;
    edge_probability_map.assign((size_t)image_width * image_height, 0);
  }
  rgbd::load_rgbd_data_sampled(rgb_location, depth_location, class_probability_map_location, edge_probability_map, edge_w,
                               edge_h, camera_intrinsics, read_depth_scale, voxel_size, class_threshold, point3d_scene, ctx_);
  rgbd::save_as_ply(dst_scene_location, point3d_scene, write_depth_scale);
}

// src/stocs.cpp:943-964.  The sequential fp32 centroids are computed by the upload kernels; the
// host copies are shifted with the values read back so that accessors see centred clouds.
void stocs_estimator::centroid_shift() {
  const size_t S = point3d_scene.size(), M = point3d_model.size();
  if (S == 0 || M == 0) { std::cerr << "empty scene or model" << std::endl; return; }
  std::vector<float> mp(M * 3), mn(M * 3), sp(S * 3), sn(S * 3), sc(S);
  std::vector<int32_t> pix(S * 2);
  for (size_t i = 0; i < M; ++i)
    for (int k = 0; k < 3; ++k) { mp[3 * i + k] = point3d_model[i].pos()[k]; mn[3 * i + k] = point3d_model[i].normal()[k]; }
  for (size_t i = 0; i < S; ++i) {
    for (int k = 0; k < 3; ++k) { sp[3 * i + k] = point3d_scene[i].pos()[k]; sn[3 * i + k] = point3d_scene[i].normal()[k]; }
    sc[i] = point3d_scene[i].class_probability();
    pix[2 * i] = point3d_scene[i].pixel().first; pix[2 * i + 1] = point3d_scene[i].pixel().second;
  }
  if (const char* dump = std::getenv("STOCS_DUMP_INPUTS")) {
    // raw little-endian dump of what is uploaded (fixture generation for the parity tests):
    // int64 S, int64 M, then scene pos/nrm (S*3 f32 each), cls (S f32), pixel (S*2 i32), model pos/nrm
    FILE* f = fopen(dump, "wb");
    if (f) {
      int64_t hdr[2] = {(int64_t)S, (int64_t)M};
      fwrite(hdr, 8, 2, f);
      fwrite(sp.data(), 4, sp.size(), f); fwrite(sn.data(), 4, sn.size(), f); fwrite(sc.data(), 4, sc.size(), f);
      fwrite(pix.data(), 4, pix.size(), f); fwrite(mp.data(), 4, mp.size(), f); fwrite(mn.data(), 4, mn.size(), f);
      fclose(f);
    }
  }
  if (group_) {  // every device holds a replica of the model tables and the scene index
    if (stocs_b200_group_upload_model(group_, mp.data(), mn.data(), (int)M) != 0) fail("upload_model");
    if (stocs_b200_group_upload_scene(group_, sp.data(), sn.data(), sc.data(), pix.data(), (int)S) != 0) fail("upload_scene");
  } else {
    if (stocs_b200_upload_model(ctx_, mp.data(), mn.data(), (int)M) != 0) fail("upload_model");
    if (stocs_b200_upload_scene(ctx_, sp.data(), sn.data(), sc.data(), pix.data(), (int)S) != 0) fail("upload_scene");
  }
  // The preloaded map is THE table (reference: ppf_map = ppf_map_preloaded, src/stocs.cpp:94).  An
  // empty one (no file) leaves the table upload_model derived from the points; one that does not
  // fit the model or the discretisations stops the run.
  if (!ppf_map.empty()) {
    if (stocs_b200_upload_ppf_table(ctx_, ppf_map.keys4.data(), ppf_map.pairs2.data(), (int64_t)(ppf_map.pairs2.size() / 2),
                                    ppf_map.tr_discretization, ppf_map.rot_discretization, ppf_map.num_model_points) != 0)
      fail("upload_ppf_table (re-run model_preprocess for this model and these discretisations)");
  }
  float cs[3], cm[3];
  if (stocs_b200_get_centroids(ctx_, cs, cm) != 0) fail("get_centroids");
  if (stocs_b200_get_centred(ctx_, sp.data(), mp.data()) != 0) fail("get_centred");
  centroid_scene_ = VectorType(cs[0], cs[1], cs[2]);
  centroid_model_ = VectorType(cm[0], cm[1], cm[2]);
  for (size_t i = 0; i < S; ++i) point3d_scene[i].pos() = VectorType(sp[3 * i], sp[3 * i + 1], sp[3 * i + 2]);
  for (size_t i = 0; i < M; ++i) point3d_model[i].pos() = VectorType(mp[3 * i], mp[3 * i + 1], mp[3 * i + 2]);
}

void stocs_estimator::kdtree_initialize() {
  std::cout << "|S|: " << point3d_scene.size() << std::endl;  // the index was built by upload_scene
}

// src/stocs.cpp:363-519.  Class-mode bases are independent and keyed by (seed, base number), so
// the shim samples them kPrefetch at a time in ONE launch and hands them out call by call; the
// values are identical to sampling them one by one.
bool stocs_estimator::sample_class_base(std::vector<int>& base_indices, float& invariant1, float& invariant2) {
  const uint32_t no = next_base_no_++;
  if (no < cache_first_ || no >= cache_first_ + (uint32_t)cache_valid_.size()) {
    cache_first_ = no;
    cache_ids_.assign((size_t)kPrefetch * 4, -1);
    cache_inv_.assign((size_t)kPrefetch * 2, 0.f);
    cache_valid_.assign((size_t)kPrefetch, 0);
    if (stocs_b200_sample_bases(ctx_, seed_, no, kPrefetch, cache_ids_.data(), cache_inv_.data(), cache_valid_.data()) != 0)
      fail("sample_bases");
    cong_ready_ = false;
  }
  const size_t k = no - cache_first_;
  if (!cache_valid_[k]) {
    std::cout << "FAILED SAMPLING:: Zero probability returned!!!" << std::endl;
    return false;
  }
  for (int j = 0; j < 4; ++j) base_indices[j] = cache_ids_[4 * k + j];
  invariant1 = cache_inv_[2 * k];
  invariant2 = cache_inv_[2 * k + 1];
  handed_out_ = std::max(handed_out_, k + 1);
  return true;
}

// src/stocs.cpp:559-751.  One kernel launch per base (bases are sequentially coupled here).
bool stocs_estimator::sample_instance_base(std::vector<int>& base_indices, float& invariant1, float& invariant2,
                                           std::vector<Point3D>& segment, float dispersion, int base_num) {
  if (!edge_uploaded_) {
    if (stocs_b200_upload_edge_map(ctx_, edge_probability_map.data(), edge_w, edge_h) != 0) fail("upload_edge_map");
    edge_uploaded_ = true;
  }
  int32_t ids[4];
  float inv[2];
  uint8_t valid = 0;
  std::vector<uint32_t> seg_bits((point3d_scene.size() + 31) / 32, 0u);
  if (stocs_b200_sample_instance_base(ctx_, seed_, base_num, dispersion, ids, inv, &valid, nullptr, seg_bits.data()) != 0)
    fail("sample_instance_base");
  for (size_t i = 0; i < point3d_scene.size(); ++i)
    if ((seg_bits[i >> 5] >> (i & 31)) & 1u) segment.push_back(point3d_scene[i]);
  class_prob_dirty_ = true;  // the prior was decayed on the device (point3d.hpp:54-56 semantics)
  if (!valid) return false;
  for (int k = 0; k < 4; ++k) base_indices[k] = ids[k];
  invariant1 = inv[0];
  invariant2 = inv[1];
  return true;
}

// src/stocs.cpp:753-869.  When the queried base is one of the bases this estimator handed out, the
// congruent sets of ALL of them are computed in one batched call (the reference driver asks for
// every base in turn, src/stocs_match_one_object.cpp:111-118) and served from the cache.
bool stocs_estimator::find_congruent_sets_on_model(std::vector<int>& base_indices, float invariant1, float invariant2,
                                                   std::vector<Quadrilateral>* quadrilaterals) {
  quadrilaterals->clear();
  int32_t ids[4] = {base_indices[0], base_indices[1], base_indices[2], base_indices[3]};
  float inv[2] = {invariant1, invariant2};
  // cached?
  for (size_t k = 0; k < handed_out_ && k < cache_valid_.size(); ++k) {
    if (!cache_valid_[k]) continue;
    if (std::memcmp(&cache_ids_[4 * k], ids, 16) != 0 || cache_inv_[2 * k] != inv[0] || cache_inv_[2 * k + 1] != inv[1]) continue;
    if (!cong_ready_) {
      std::vector<int32_t> bids;
      std::vector<float> binv;
      cong_slot_.assign(cache_valid_.size(), -1);
      for (size_t j = 0; j < handed_out_; ++j)
        if (cache_valid_[j]) {
          cong_slot_[j] = (int)(bids.size() / 4);
          bids.insert(bids.end(), &cache_ids_[4 * j], &cache_ids_[4 * j] + 4);
          binv.push_back(cache_inv_[2 * j]); binv.push_back(cache_inv_[2 * j + 1]);
        }
      const int nb = (int)(bids.size() / 4);
      cong_off_.assign((size_t)nb + 1, 0);
      cong_quads_.assign(4 * 65536, 0);
      int rc = stocs_b200_find_congruent(ctx_, nb, bids.data(), binv.data(), cong_quads_.data(), (int64_t)cong_quads_.size() / 4, cong_off_.data());
      if (rc == STOCS_E_CAPACITY) {
        cong_quads_.assign((size_t)cong_off_[nb] * 4, 0);
        rc = stocs_b200_find_congruent(ctx_, nb, bids.data(), binv.data(), cong_quads_.data(), cong_off_[nb], cong_off_.data());
      }
      if (rc != 0) fail("find_congruent");
      cong_ready_ = true;
    }
    const int slot = cong_slot_[k];
    for (int64_t i = cong_off_[slot]; i < cong_off_[slot + 1]; ++i)
      quadrilaterals->emplace_back(cong_quads_[4 * i], cong_quads_[4 * i + 1], cong_quads_[4 * i + 2], cong_quads_[4 * i + 3]);
    return quadrilaterals->size() != 0;
  }
  // a base the estimator did not sample (e.g. instance mode, or supplied by the caller): single call
  int64_t off[2] = {0, 0};
  std::vector<int32_t> quads(4 * 4096);
  int rc = stocs_b200_find_congruent(ctx_, 1, ids, inv, quads.data(), (int64_t)quads.size() / 4, off);
  if (rc == STOCS_E_CAPACITY) {
    quads.resize((size_t)off[1] * 4);
    rc = stocs_b200_find_congruent(ctx_, 1, ids, inv, quads.data(), off[1], off);
  }
  if (rc != 0) fail("find_congruent");
  for (int64_t i = 0; i < off[1]; ++i)
    quadrilaterals->emplace_back(quads[4 * i], quads[4 * i + 1], quads[4 * i + 2], quads[4 * i + 3]);
  return quadrilaterals->size() != 0;
}

bool stocs_estimator::get_rigid_transform_from_congruent_pair(std::vector<int>& base_indices, Quadrilateral& q,
                                                              int base_index) {
  for (int k = 0; k < 4; ++k) { pending_bases_.push_back(base_indices[k]); pending_quads_.push_back(q[k]); }
  pending_base_index_.push_back(base_index);
  return true;  // src/stocs.cpp:940 always returns true
}

void stocs_estimator::flush_pending() {
  const int64_t n = (int64_t)pending_base_index_.size();
  if (n == 0) return;
  std::vector<float> Tc((size_t)n * 16), Tw((size_t)n * 16);
  std::vector<uint8_t> ok((size_t)n);
  if (stocs_b200_fit_transforms(ctx_, n, pending_bases_.data(), pending_quads_.data(), Tc.data(), Tw.data(), ok.data()) != 0)
    fail("fit_transforms");
  for (int64_t i = 0; i < n; ++i) {
    if (!ok[i]) continue;  // "if(ok && rms >= 0)" (src/stocs.cpp:922)
    MatrixType t, w;
    std::memcpy(t.data(), &Tc[16 * i], 64);
    std::memcpy(w.data(), &Tw[16 * i], 64);
    all_transforms.push_back(t);
    all_pose.push_back(new PoseCandidate(w, 0, (float)pending_base_index_[i]));
  }
  pending_bases_.clear(); pending_quads_.clear(); pending_base_index_.clear();
}

Scalar stocs_estimator::compute_alignment_score_for_rigid_transform(const Eigen::Ref<const MatrixType>& mat) {
  float lcp = 0;
  if (stocs_b200_score_lcp(ctx_, mat.data(), 1, &lcp, nullptr) != 0) fail("score_lcp");
  return lcp;
}

// src/stocs.cpp:982-1004
void stocs_estimator::compute_best_transform() {
  flush_pending();
  std::cout << "Transforms to verify: " << all_transforms.size() << std::endl;
  const int64_t H = (int64_t)all_transforms.size();
  best_lcp = 0;
  best_index = -1;
  if (H > 0) {
    std::vector<float> T((size_t)H * 16), lcp((size_t)H);
    for (int64_t i = 0; i < H; ++i) std::memcpy(&T[16 * i], all_transforms[i].data(), 64);
    int64_t bi = -1;
    float bl = 0;
    if (group_) {
      // instance sampling decays the class prior on device 0 only; LCP uses the decayed values
      // (src/stocs.cpp:577,1033), so the other replicas are brought up to date first
      if (class_prob_dirty_) {
        std::vector<float> cls(point3d_scene.size());
        if (stocs_b200_get_class_probability(ctx_, cls.data()) != 0) fail("get_class_probability");
        for (int d = 1; d < stocs_b200_group_size(group_); ++d)
          if (stocs_b200_set_class_probability(stocs_b200_group_ctx(group_, d), cls.data()) != 0) fail("set_class_probability");
      }
      stocs_b200_record top;
      if (stocs_b200_group_score_best(group_, T.data(), H, 1, &top, lcp.data(), nullptr) != 0) fail("group_score_best");
      bi = top.index;
      bl = top.lcp;
    } else {
      if (stocs_b200_score_lcp(ctx_, T.data(), H, lcp.data(), nullptr) != 0) fail("score_lcp");
      int64_t ti[1]; float tl[1];
      if (stocs_b200_reduce_best(ctx_, nullptr, H, 1, &bi, &bl, ti, tl) != 0) fail("reduce_best");
    }
    for (int64_t i = 0; i < H; ++i) all_pose[i]->lcp = lcp[i];
    best_lcp = bl;
    best_index = (int)bi;
  }
  std::cout << "best index: " << best_index << ", maximum score: " << best_lcp << std::endl;
}

// include/stocs.hpp:136-149
void stocs_estimator::visualize_best_pose() {
  if (best_index == -1) return;
  std::vector<Point3D> point3d_model_pose;
  rgbd::transform_pointset(point3d_model, point3d_model_pose, all_transforms[best_index]);
  rgbd::save_as_ply(debug_location + "/best_pose.ply", point3d_model_pose, 1);
  rgbd::save_as_ply(debug_location + "/scene.ply", point3d_scene, 1);
}

bool load_and_sample_model(std::string src_model_location, float normal_radius, float read_depth_scale, float voxel_size,
                           std::vector<Point3D>& point3d_sampled) {
  PCLPointCloud::Ptr cloud(new PCLPointCloud);
  if (!rgbd::load_ply_file(src_model_location, *cloud)) { std::cerr << "cannot read " << src_model_location << std::endl; return false; }
  rgbd::compute_normal_pcl(cloud, normal_radius);
  for (auto& p : cloud->points) { p.nx = -p.nx; p.ny = -p.ny; p.nz = -p.nz; }  // normals face outside
  rgbd::voxel_grid_filter(*cloud, voxel_size);
  rgbd::load_ply_model(cloud, point3d_sampled, read_depth_scale);
  return true;
}

// src/stocs.cpp:28-84.  Normal estimation and voxel-grid down-sampling are host restatements of
// the PCL operators; the O(|M|^2) pair loop runs on the GPU (ppf_table.cu).
void pre_process_model(std::string src_model_location, float normal_radius, float read_depth_scale,
                       float write_depth_scale, float voxel_size, float ppf_tr_discretization,
                       float ppf_rot_discretization, std::string dst_model_location, std::string dst_ppf_map_location) {
  std::vector<Point3D> point3d_sampled;
  if (!load_and_sample_model(src_model_location, normal_radius, read_depth_scale, voxel_size, point3d_sampled)) return;
  std::cout << "After sampling |M|= " << point3d_sampled.size() << std::endl;

  stocs_b200_ctx* ctx = nullptr;
  const char* dev = std::getenv("STOCS_DEVICE");
  if (stocs_b200_create(&ctx, dev ? std::atoi(dev) : 0) != 0) {
    std::cerr << "libstocs_b200: cannot create a GPU context: " << stocs_b200_last_error(nullptr) << std::endl;
    std::exit(2);
  }
  stocs_b200_set_params(ctx, 0.005f, (int)ppf_tr_discretization, (int)ppf_rot_discretization);
  const size_t M = point3d_sampled.size();
  std::vector<float> mp(M * 3), mn(M * 3);
  float max_distance = 0;
  for (size_t i = 0; i < M; ++i)
    for (int k = 0; k < 3; ++k) { mp[3 * i + k] = point3d_sampled[i].pos()[k]; mn[3 * i + k] = point3d_sampled[i].normal()[k]; }
  if (stocs_b200_upload_model(ctx, mp.data(), mn.data(), (int)M) != 0) {
    std::cerr << "libstocs_b200: upload_model: " << stocs_b200_last_error(ctx) << std::endl;
    std::exit(2);
  }
  for (size_t i = 0; i < M; ++i)
    for (size_t j = 0; j < M; ++j) {
      if (i == j) continue;
      float d = (point3d_sampled[i].pos() - point3d_sampled[j].pos()).norm();
      if (d > max_distance) max_distance = d;
    }
  std::cout << "max distance is: " << max_distance << std::endl;
  PPFMapType map;
  map.tr_discretization = (int)ppf_tr_discretization;
  map.rot_discretization = (int)ppf_rot_discretization;
  map.num_model_points = (int)M;
  int64_t n = 0;
  stocs_b200_ppf_export(ctx, nullptr, nullptr, 0, &n);
  map.keys4.resize((size_t)n * 4);
  map.pairs2.resize((size_t)n * 2);
  if (stocs_b200_ppf_export(ctx, map.keys4.data(), map.pairs2.data(), n, &n) != 0) {
    std::cerr << "libstocs_b200: ppf_export: " << stocs_b200_last_error(ctx) << std::endl;
    std::exit(2);
  }
  stocs_b200_ppf_num_expanded_keys(ctx, &map.expanded_keys);
  stocs_b200_destroy(ctx);
  rgbd::save_ppf_map(dst_ppf_map_location, map);
  rgbd::save_as_ply(dst_model_location, point3d_sampled, write_depth_scale);
}

}  // namespace stocs
