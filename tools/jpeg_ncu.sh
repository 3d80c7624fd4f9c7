# ncu --set full of the JPEG kernels on the shape of tools/jpeg_probe.py (64 x 1080p; k_jpeg_sync: a restart marker per block row, the
# first configuration; k_jpeg_huff: a marker every 16 blocks, the second), after the probe itself exited 0:  sh tools/jpeg_ncu.sh
set -e
python tools/jpeg_probe.py > gpurun_out/jpeg_probe.txt 2>&1
for K in sync huff idct unstuff ycc; do
  ncu --set full --import-source on --clock-control none -k regex:k_jpeg_$K --launch-skip 3 --launch-count 1 -f -o gpurun_out/jpeg_$K \
      python tools/jpeg_probe.py > gpurun_out/jpeg_ncu_$K.log 2>&1
done
