// model_preprocess <object_name> -- offline model preparation CLI, same argv and outputs as the
// reference's src/model_preprocess.cpp:14-39: <repo>/models/<object>/{model_search.ply, ppf_map}.
// Compile-time globals keep the reference's values (:3-12); environment overrides:
// STOCS_REPO_PATH, STOCS_MODEL_VOXEL_SIZE, STOCS_NORMAL_RADIUS, STOCS_MODEL_SCALE.
#include <cstdlib>
#include <iostream>

#include "stocs.hpp"

#ifndef STOCS_DEFAULT_REPO_PATH
#define STOCS_DEFAULT_REPO_PATH "."
#endif
std::string repo_path = STOCS_DEFAULT_REPO_PATH;

// All values in m
float voxel_size = 0.01;
float normal_radius = 0.005;
float model_scale = 1.0;

// All values in mm
int ppf_tr_discretization = 5;
int ppf_rot_discretization = 5;

int main(int argc, char** argv) {
  if (argc < 2) {
    std::cout << "Enter name of the object model!!" << std::endl;
    exit(-1);
  }
  if (const char* e = std::getenv("STOCS_REPO_PATH")) repo_path = e;
  if (const char* e = std::getenv("STOCS_MODEL_VOXEL_SIZE")) voxel_size = (float)atof(e);
  if (const char* e = std::getenv("STOCS_NORMAL_RADIUS")) normal_radius = (float)atof(e);
  if (const char* e = std::getenv("STOCS_MODEL_SCALE")) model_scale = (float)atof(e);
  std::string object_name = argv[1];
  std::string model_path = repo_path + "/models/" + object_name;

  if (system(("rm -rf " + model_path + "/model_search.ply").c_str()) != 0) return 1;
  if (system(("rm -rf " + model_path + "/ppf_map").c_str()) != 0) return 1;

  stocs::pre_process_model(model_path + "/textured_vertices.ply", normal_radius, model_scale, 1.0f, voxel_size,
                           ppf_tr_discretization, ppf_rot_discretization, model_path + "/model_search.ply",
                           model_path + "/ppf_map");
  return 0;
}
