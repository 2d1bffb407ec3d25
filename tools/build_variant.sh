#!/bin/bash
# usage: tools/build_variant.sh <name> <file.cu> [nvcc -D flags...]  -- rebuilds ONE source with extra flags and links it
# with the objects of the regular build into 3d-vlm-gd_b200/lib/variants/lib3dgd_<name>.so (for GD3_LIB=... experiments)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
P=$ROOT/3d-vlm-gd_b200
name=$1; src=$2; shift 2
mkdir -p $P/lib/variants $P/build/variants
obj=$P/build/variants/${name}.o
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC \
  --expt-relaxed-constexpr -I $ROOT/include -I $P/csrc "$@" -Xptxas -v -c $P/csrc/$src -o $obj 2> $P/build/variants/${name}.ptxas.txt
others=$(ls $P/build/*.o | grep -v "/${src%.cu}.o")
/usr/local/cuda/bin/nvcc -shared -o $P/lib/variants/lib3dgd_${name}.so $obj $others -lcudart
echo "built $name"
