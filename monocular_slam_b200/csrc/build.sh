#!/bin/sh
# Builds liborbx.so (the C-ABI product library) for sm_100a.  Also usable by hand: sh monocular_slam_b200/csrc/build.sh
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-nvcc}
OUT=../liborbx.so
FLAGS="$ORBX_EXTRA_FLAGS -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC,-Wall,-Wextra,-fvisibility=hidden -Xptxas -v"
OBJS=""
for f in hamming fmat triangulate bow jpeg orb_pyramid orb_fast orb_select orb_describe orbx_api; do
    $NVCC $FLAGS -c $f.cu -o $f.o
    OBJS="$OBJS $f.o"
done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o $OUT $OBJS -cudart static -lpthread
echo "built $OUT"
