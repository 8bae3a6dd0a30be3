#!/bin/bash
# Runs the two CLIs on the reference's three example scenes ON THE GPU BOX and keeps, per scene, the
# point sets the estimator uploads (STOCS_DUMP_INPUTS) -- the INPUTS of tests/golden/golden_<scene>.npz
# (make_golden.py then computes every golden OUTPUT with the CPU oracle).
#   gpurun -- 'bash tests/golden/dump_inputs.sh gpurun_out/inputs'
set -e
OUT=${1:-gpurun_out/inputs}; mkdir -p "$OUT"; OUT=$(realpath "$OUT")
ROOT=$(cd "$(dirname "$0")/../.." && pwd)
run() {  # scene object env...
  local scene=$1 obj=$2; shift 2
  local T; T=$(mktemp -d)
  mkdir -p "$T/models/$obj" "$T/examples"
  cp -r "$ROOT/tests/golden/examples/$scene" "$T/examples/"
  cp "$ROOT/tests/golden/models/$obj/textured_vertices.ply" "$T/models/$obj/"
  env STOCS_REPO_PATH="$T" STOCS_SEED=7 "$@" "$ROOT/model_matching_b200/host/model_preprocess" "$obj" > "$OUT/${scene}_preprocess.log"
  env STOCS_REPO_PATH="$T" STOCS_SEED=7 STOCS_DUMP_INPUTS="$OUT/${scene}_inputs.bin" "$@" \
      "$ROOT/model_matching_b200/host/stocs_single" "$T/examples/$scene" "$obj" > "$OUT/${scene}_single.log"
  cp "$T/examples/$scene/best_pose_candidate_$obj.txt" "$OUT/${scene}_pose.txt" 2>/dev/null || true
  rm -rf "$T"
}
run ycb 024_bowl
run linemod obj_06 STOCS_CAM_INTRINSICS=572.4114,325.2611,573.57043,242.04899 STOCS_DEPTH_SCALE=0.001 STOCS_MODEL_VOXEL_SIZE=10 STOCS_NORMAL_RADIUS=5 STOCS_MODEL_SCALE=0.001
run packed dove STOCS_CAM_INTRINSICS=615.957763671875,308.1098937988281,615.9578247070312,246.33352661132812 STOCS_DEPTH_SCALE=0.000125 STOCS_MODEL_VOXEL_SIZE=0.005
ls -la "$OUT"
