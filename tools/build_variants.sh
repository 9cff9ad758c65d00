#!/bin/bash
# A/B builds of the library with different compile-time knobs of the cell-tile kernel:
#   tools/build_variants.sh name "-DLJ_CT_UNROLL_MX=8 -DLJ_TILE_ROWS_WIDE=88" [name2 "flags2" ...]
# -> build_variants/liblj_b200_<name>.so (git-ignored; select with LJ_B200_LIB=...)
set -e
cd "$(dirname "$0")/../lj_gpu_b200/csrc"
mkdir -p ../../build_variants build
ARCH="-gencode arch=compute_100a,code=sm_100a"
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  vobjs=""
  for f in lj_force_celltile lj_nlist; do   # the two files with compile-time knobs (LJ_CT_*, LJ_TILE_ROWS_*, LJ_TE_*)
    /usr/local/cuda/bin/nvcc -O3 -std=c++17 $ARCH -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
      -cudart static $flags -c $f.cu -o build/v_${f}_$name.o
    vobjs="$vobjs build/v_${f}_$name.o"
  done
  objs=$(ls build/lj_*.o | grep -v "lj_force_celltile\|lj_nlist")
  /usr/local/cuda/bin/nvcc $ARCH -shared -cudart static -o ../../build_variants/liblj_b200_$name.so $objs $vobjs
  rm -f $vobjs
  echo "built build_variants/liblj_b200_$name.so ($flags)"
done
