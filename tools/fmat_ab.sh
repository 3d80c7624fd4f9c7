#!/bin/sh
# A/B of k_fm_ransac's residency on one box: 2 CTAs of 256 threads per SM (128 registers, the default) against 3 (80 registers, so that
# the 320 pairs of a step are resident at once instead of 296 + 24).  Rebuilds fmat.o and relinks; ends on the default build.
cd "$(dirname "$0")/../monocular_slam_b200/csrc"
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC,-Wall,-Wextra,-fvisibility=hidden"
OBJS="hamming.o fmat.o triangulate.o bow.o jpeg.o orb_pyramid.o orb_fast.o orb_select.o orb_describe.o orbx_api.o"
for V in 3 2 3 2; do
    nvcc $FLAGS -DFM_CTAS_PER_SM=$V -c fmat.cu -o fmat.o && nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../liborbx.so $OBJS -cudart static -lpthread || exit 1
    echo "== $V CTAs per SM"
    python ../../tools/fmat_probe.py 320 1000 0.7 2>&1 | grep "pairs/s"
    python ../../tools/fmat_probe.py 640 1000 0.7 2>&1 | grep "conf 0.85"
done
