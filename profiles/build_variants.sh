#!/bin/bash
# Build score-kernel variants for profiles/ab_probe.py:  profiles/build_variants.sh name:"-Dflags" ...
# Each variant is libstocs_b200.so with score.cu recompiled under the extra flags (the other objects
# are reused from csrc/.obj).  Output: gpurun_variants/lib_<name>.so (git-ignored, travels with gpurun).
set -e
cd "$(dirname "$0")/../model_matching_b200/csrc"
make -s -j8 >/dev/null
mkdir -p ../../gpurun_variants
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
ARCH="-gencode arch=compute_100a,code=sm_100a"
FLAGS="-O3 -std=c++17 $ARCH -lineinfo -fmad=false -Xcompiler -fPIC,-ffp-contract=off -Xptxas -v"
for spec in "$@"; do
  name="${spec%%:*}"; defs="${spec#*:}"; [ "$defs" = "$spec" ] && defs=""
  ( $NVCC $FLAGS $defs -c -o /tmp/score_$name.o score.cu 2>&1 | grep -A2 "score_lcp_kernelILb0" | grep -E "registers|spill" | tr '\n' ' ' | sed "s/^/$name: /"; echo
    objs=$(ls .obj/*.o | grep -v score.o)
    $NVCC $ARCH -shared -o ../../gpurun_variants/lib_$name.so $objs /tmp/score_$name.o -lcudart -ldl ) &
done
wait
