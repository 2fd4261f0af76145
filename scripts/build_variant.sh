#!/bin/bash
# Builds an experimental variant of the library: scripts/build_variant.sh NAME -DMACRO=1 ...  -> build/exp/libtnf_NAME.so
# (select it at run time with TNF_B200_LIB=build/exp/libtnf_NAME.so)
set -e
name=$1; shift
out=build/exp/$name; mkdir -p $out
for f in thermo_nerf_b200/csrc/*.cu; do
  /usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC "$@" -I include -c $f -o $out/$(basename $f .cu).o &
done
wait
/usr/local/cuda/bin/nvcc -shared -gencode arch=compute_100a,code=sm_100a -o build/exp/libtnf_$name.so $out/*.o
echo built build/exp/libtnf_$name.so
