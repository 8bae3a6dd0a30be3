// stocs_single <scene_path> <object_name> -- online pose estimation CLI, same argv, same driver
// loop and same outputs as the reference's src/stocs_match_one_object.cpp:51-215:
//   <scene>/best_pose_candidate_<object>.txt (12 floats, rows 0-2 of the un-centred pose),
//   <scene>/dbg/{sampled_scene.ply, best_pose.ply, scene.ply}.
// The reference's compile-time globals (:4-24) keep their values; each can be overridden from the
// environment (STOCS_REPO_PATH, STOCS_CAM_INTRINSICS="fx,cx,fy,cy", STOCS_DEPTH_SCALE,
// STOCS_VOXEL_SIZE, STOCS_NUM_BASES, STOCS_MAX_SETS, STOCS_SEED) instead of editing the source.
#include <sys/stat.h>

#include <algorithm>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>

#include "stocs.hpp"

#ifndef STOCS_DEFAULT_REPO_PATH
#define STOCS_DEFAULT_REPO_PATH "."
#endif
std::string repo_path = STOCS_DEFAULT_REPO_PATH;

// rgbd parameters
float voxel_size = 0.005;          // In m
float distance_threshold = 0.005;  // for Congruent Set Matching and LCP computation
int ppf_tr_discretization = 5;     // In mm
int ppf_rot_discretization = 5;    // degrees
float edge_threshold = 0;          // Not used
float class_threshold = 0.10;      // Cut-off probability
float sample_dispersion = 0.9;

// stocs parameters
int number_of_bases = 100;
int maximum_congruent_sets = 200;

// camera parameters
std::vector<float> cam_intrinsics = {1066.778, 312.986, 1067.487, 241.310};  // YCB
float depth_scale = 1 / 10000.0f;

int image_width = 640;
int image_height = 480;

class BaseGraph {
 public:
  std::vector<int> baseIds_;
  float invariant1_;
  float invariant2_;
  std::vector<Quadrilateral> congruent_quads;
  BaseGraph(std::vector<int> base_ids, float invariant1, float invariant2) {
    baseIds_.assign(base_ids.begin(), base_ids.begin() + 4);
    invariant1_ = invariant1;
    invariant2_ = invariant2;
  }
};

static void env_overrides() {
  if (const char* e = std::getenv("STOCS_REPO_PATH")) repo_path = e;
  if (const char* e = std::getenv("STOCS_DEPTH_SCALE")) depth_scale = (float)atof(e);
  if (const char* e = std::getenv("STOCS_VOXEL_SIZE")) voxel_size = (float)atof(e);
  if (const char* e = std::getenv("STOCS_NUM_BASES")) number_of_bases = atoi(e);
  if (const char* e = std::getenv("STOCS_MAX_SETS")) maximum_congruent_sets = atoi(e);
  if (const char* e = std::getenv("STOCS_CAM_INTRINSICS")) {
    std::stringstream ss(e);
    std::string t;
    std::vector<float> k;
    while (std::getline(ss, t, ',')) k.push_back((float)atof(t.c_str()));
    if (k.size() == 4) cam_intrinsics = k;
  }
}

void run_stocs_estimation(std::string scene_path, std::string object_name, PPFMapType& ppf_map_preloaded) {
  std::string rgb_path = scene_path + "/rgb.png";
  std::string depth_path = scene_path + "/depth.png";
  std::string class_probability_path = scene_path + "/probability_maps/" + object_name + ".png";
  std::string edge_probability_path = scene_path + "/probability_maps/edge.png";
  std::string model_path = repo_path + "/models/" + object_name + "/model_search.ply";
  std::string output_pose_file = scene_path + "/best_pose_candidate_" + object_name + ".txt";

  std::vector<BaseGraph*> base_set;

  stocs::stocs_estimator stocs_ptr(model_path, ppf_map_preloaded, rgb_path, depth_path, class_probability_path,
                                   edge_probability_path, scene_path + "/dbg", cam_intrinsics, image_width, image_height,
                                   depth_scale, 1.0f, voxel_size, distance_threshold, ppf_tr_discretization,
                                   ppf_rot_discretization, edge_threshold, class_threshold);

  // Step 1: Sample n bases on scene
  auto start = std::chrono::high_resolution_clock::now();
  struct stat buffer;
  const bool instance_mode = stat(edge_probability_path.c_str(), &buffer) == 0;  // src/stocs_match_one_object.cpp:91
  for (int i = 0; i < number_of_bases; i++) {
    bool valid_base_found = false;
    std::vector<int> base_indices(4, -1);
    float invariant1, invariant2;
    std::vector<Point3D> segment;
    if (instance_mode)
      valid_base_found = stocs_ptr.sample_instance_base(base_indices, invariant1, invariant2, segment, sample_dispersion, i + 1);
    else
      valid_base_found = stocs_ptr.sample_class_base(base_indices, invariant1, invariant2);
    if (valid_base_found) base_set.push_back(new BaseGraph(base_indices, invariant1, invariant2));
  }
  auto finish = std::chrono::high_resolution_clock::now();
  std::cout << "Sampled " << base_set.size() << " bases in " << std::chrono::duration_cast<micro>(finish - start).count()
            << " microseconds\n";
  auto total_time = std::chrono::duration_cast<micro>(finish - start).count();

  // Step 2: congruent sets on the model for each sampled base
  start = std::chrono::high_resolution_clock::now();
  for (auto base_iterator : base_set)
    stocs_ptr.find_congruent_sets_on_model(base_iterator->baseIds_, base_iterator->invariant1_, base_iterator->invariant2_,
                                           &base_iterator->congruent_quads);

  // Step 3: at most k congruent pairs per base -> rigid transformations
  int total_congruent_set_found = 0;
  int base_number = 0;
  for (auto base_iterator : base_set) {
    int congruent_set_size = (int)base_iterator->congruent_quads.size();
    if (congruent_set_size < maximum_congruent_sets) {
      for (int i = 0; i < congruent_set_size; i++)
        stocs_ptr.get_rigid_transform_from_congruent_pair(base_iterator->baseIds_, base_iterator->congruent_quads[i], base_number);
    } else {
      // The reference shuffles an index vector that starts with congruent_set_size zeros
      // (src/stocs_match_one_object.cpp:134-139, quirk 5) with the unseeded rand(); here the
      // subset is the deterministic even spread floor(k * size / max), see DESIGN.md.
      for (int i = 0; i < maximum_congruent_sets; i++) {
        int pick = (int)(((long long)i * congruent_set_size) / maximum_congruent_sets);
        stocs_ptr.get_rigid_transform_from_congruent_pair(base_iterator->baseIds_, base_iterator->congruent_quads[pick], base_number);
      }
    }
    total_congruent_set_found += congruent_set_size;
    base_number++;
  }
  finish = std::chrono::high_resolution_clock::now();
  std::cout << "found " << total_congruent_set_found << " congruent sets in "
            << std::chrono::duration_cast<micro>(finish - start).count() << " microseconds\n";
  total_time += std::chrono::duration_cast<micro>(finish - start).count();

  // Verify all transforms to get the best pose
  start = std::chrono::high_resolution_clock::now();
  stocs_ptr.compute_best_transform();
  finish = std::chrono::high_resolution_clock::now();
  std::cout << "evaluated transforms in " << std::chrono::duration_cast<micro>(finish - start).count() << " microseconds\n";
  total_time += std::chrono::duration_cast<micro>(finish - start).count();
  std::cout << "total " << total_time << " microseconds\n";

  stocs_ptr.visualize_best_pose();
  PoseCandidate* best_pose = stocs_ptr.get_best_pose();
  if (best_pose != NULL) {
    std::ofstream out_file_ptr;
    out_file_ptr.open(output_pose_file, std::ofstream::out);
    out_file_ptr << best_pose->transform(0, 0) << " " << best_pose->transform(0, 1) << " " << best_pose->transform(0, 2) << " "
                 << best_pose->transform(0, 3) << " " << best_pose->transform(1, 0) << " " << best_pose->transform(1, 1) << " "
                 << best_pose->transform(1, 2) << " " << best_pose->transform(1, 3) << " " << best_pose->transform(2, 0) << " "
                 << best_pose->transform(2, 1) << " " << best_pose->transform(2, 2) << " " << best_pose->transform(2, 3)
                 << std::endl;
    out_file_ptr.close();
  } else {
    std::cout << "no pose found" << std::endl;
  }
}

int main(int argc, char** argv) {
  if (argc < 3) {
    std::cout << "Enter scene path and object name as arguments!" << std::endl;
    exit(-1);
  }
  env_overrides();
  std::string scene_path = argv[1];
  std::string object_name = argv[2];

  std::cout << "############# LOADING OBJECT MAPS ################" << std::endl;
  PPFMapType model_map;
  std::string model_map_path = repo_path + "/models/" + object_name + "/ppf_map";
  rgbd::load_ppf_map(model_map_path, model_map);
  std::cout << "############# LOADING OBJECT COMPLETE ################" << std::endl;

  if (system(("rm -rf " + scene_path + "/dbg").c_str()) != 0) return 1;
  if (system(("mkdir " + scene_path + "/dbg").c_str()) != 0) return 1;

  std::cout << "############# RUNNING STOCS for Scene: " << scene_path << ", Object: " << object_name << " ##############"
            << std::endl;
  run_stocs_estimation(scene_path, object_name, model_map);
  return 0;
}
