#!/bin/bash
# A/B builds of the library with different compile-time knobs of the cell-tile kernel:
#   tools/build_variants.sh name "-DLJ_CT_UNROLL_MX=8" [name2 "flags2" ...]
# -> build_variants/liblj_b200_<name>.so (git-ignored; select with LJ_B200_LIB=...)
set -e
cd "$(dirname "$0")/../lj_gpu_b200/csrc"
mkdir -p ../../build_variants build
ARCH="-gencode arch=compute_100a,code=sm_100a"
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  /usr/local/cuda/bin/nvcc -O3 -std=c++17 $ARCH -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
    -cudart static $flags -c lj_force_celltile.cu -o build/lj_force_celltile_$name.o
  objs=$(ls build/*.o | grep -v "lj_force_celltile")
  /usr/local/cuda/bin/nvcc $ARCH -shared -cudart static -o ../../build_variants/liblj_b200_$name.so $objs build/lj_force_celltile_$name.o
  rm -f build/lj_force_celltile_$name.o
  echo "built build_variants/liblj_b200_$name.so ($flags)"
done
