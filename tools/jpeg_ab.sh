#!/bin/sh
# A/B of two builds of jpeg.cu on one box: csrc/jpeg_old.o and csrc/jpeg_new.o (prepared by hand) linked in turn; frames/s of
# tools/jpeg_probe.py (camera-like frames) and of bench.py's ingest leg (stress frames).  JPEG_AB_ORDER="old new" runs each once,
# JPEG_AB_NOBENCH=1 leaves out the bench leg.
cd "$(dirname "$0")/../monocular_slam_b200/csrc"
OBJS="hamming.o fmat.o triangulate.o bow.o orb_pyramid.o orb_fast.o orb_select.o orb_describe.o orbx_api.o"
for V in ${JPEG_AB_ORDER:-old new old new}; do
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../liborbx.so $OBJS jpeg_$V.o -cudart static -lpthread || exit 1
    echo "== $V"
    python ../../tools/jpeg_probe.py 2>&1 | grep "frames/s" | sed -E 's/^(.{34}).* = ([0-9]+ frames\/s).*/\1 \2/'
    [ -n "$JPEG_AB_NOBENCH" ] || python ../../bench.py --no-hamming --no-fundamental --no-cpu --no-sustained --no-natural --no-cfg3 --no-single --no-triangulation --no-loop --no-bow 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read())['ingest']['jpeg_gpu']; print('stress rows: %.0f (decoder alone %.0f)  none: %.0f (decoder alone %.0f)' % (d['rows']['fps'], d['rows']['decode_only_fps'], d['none']['fps'], d['none']['decode_only_fps']))"
done
