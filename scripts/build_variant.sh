#!/bin/bash
# Build a side-by-side variant of the library for A/B runs on the GPU box:
#   scripts/build_variant.sh <name> "<extra nvcc flags>" <file.cu> [<file.cu> ...]
# recompiles the named translation units with the extra flags and links them with the stock objects into
# multioutputihgp_b200/lib/ab/libmoihgp_<name>.so (git-ignored, travels with gpurun); select it with MOIHGP_B200_LIB=<path>.
set -e
cd "$(dirname "$0")/../multioutputihgp_b200/csrc"
name=$1; flags=$2; shift 2
mkdir -p ../lib/ab/obj_$name
make -j8 > /dev/null
objs=""
for o in ../lib/obj/*.o; do
  b=$(basename $o .o); keep=1
  for f in "$@"; do [ "$b" = "$(basename $f .cu)" ] && keep=0; done
  [ $keep = 1 ] && objs="$objs $o"
done
for f in "$@"; do
  b=$(basename $f .cu); extra=""
  [ "$b" = "setup" ] && extra="-fmad=false"
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -diag-suppress 128 $extra $flags -c $f -o ../lib/ab/obj_$name/$b.o
  objs="$objs ../lib/ab/obj_$name/$b.o"
done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../lib/ab/libmoihgp_$name.so $objs -lcudart
echo "built multioutputihgp_b200/lib/ab/libmoihgp_$name.so"
