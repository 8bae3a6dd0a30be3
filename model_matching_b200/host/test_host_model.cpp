// test_host_model <src.ply> <normal_radius> <read_scale> <voxel_size> <out.ply>
// The host-side half of stocs::pre_process_model (reference src/stocs.cpp:41-60): PLY load, normal
// estimation, normal flip, voxel grid, model point list -- WITHOUT the GPU part, so that
// tests/test_model_prep.py can check these restatements of the PCL operators on a CPU-only box.
#include <cstdlib>
#include <iostream>

#include "stocs.hpp"

int main(int argc, char** argv) {
  if (argc < 6) { std::cerr << "usage: test_host_model src.ply normal_radius read_scale voxel_size out.ply" << std::endl; return 2; }
  std::vector<Point3D> pts;
  if (!stocs::load_and_sample_model(argv[1], (float)atof(argv[2]), (float)atof(argv[3]), (float)atof(argv[4]), pts)) return 1;
  rgbd::save_as_ply(argv[5], pts, 1.0f);
  std::cout << pts.size() << std::endl;
  return 0;
}
