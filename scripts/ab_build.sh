#!/bin/bash
# Development A/B builds: compile the fp16-engine kernels with extra nvcc flags into ab/lib_<name>.so (the other objects are
# taken from the regular in-tree build).  usage: scripts/ab_build.sh <name> "<extra nvcc flags>" ; then
# SOCM_B200_LIB=$PWD/ab/lib_<name>.so python scripts/ab_time.py
set -e
cd "$(dirname "$0")/.."
NAME=$1; EXTRA=$2
OBJ=ab/obj_$NAME; mkdir -p $OBJ
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -diag-suppress 1886"
FILES=${FILES:-"loss_h rollout_h wgrad_h"}
for f in $FILES; do
  nvcc $FLAGS $EXTRA -c soc_matching_b200/csrc/$f.cu -o $OBJ/$f.o &
done
wait
PAT=$(for f in $FILES; do echo -n "/$f.o\|"; done); PAT=${PAT%\\|}
OTHERS=$(ls soc_matching_b200/build/*.o | grep -v "$PAT")
nvcc -shared -o ab/lib_$NAME.so $OBJ/*.o $OTHERS -gencode arch=compute_100a,code=sm_100a
echo ab/lib_$NAME.so
