#!/bin/bash
# Build a variant of the library beside the production one (A/B timing and trace builds on one GPU box):
#   tools/build_variant.sh <name> <translation unit without .cu> [nvcc defines...]
# compiles csrc/<tu>.cu with the extra defines into build/variant_<name>/ and links it with the production objects of
# the other translation units into multimodal_neuroimage_b200/libmmn_<name>.so; load it with MMN_LIB=<that path>.
set -e
name=$1; tu=$2; shift 2
pkg=$(dirname "$0")/../multimodal_neuroimage_b200
mkdir -p $pkg/build/variant_$name
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" -c $pkg/csrc/$tu.cu -o $pkg/build/variant_$name/$tu.o
objs=""
for o in $pkg/build/*.o; do
  if [ "$(basename $o)" = "$tu.o" ]; then objs="$objs $pkg/build/variant_$name/$tu.o"; else objs="$objs $o"; fi
done
nvcc -shared -o $pkg/libmmn_$name.so $objs
echo built $pkg/libmmn_$name.so
