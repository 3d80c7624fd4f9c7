# A/B of k_fast variants on one box: sh tools/fast_variants.sh   (expects /tmp-free repo copy; rebuilds liborbx.so per variant)
cp monocular_slam_b200/csrc/orb_fast.cu /tmp/orb_fast_new.cu
run() {
  python bench.py --steps 30 --warmup 5 --no-cpu --no-hamming --no-fundamental --no-loop --no-single --no-cfg3 --no-sustained > gpurun_out/fastv_$1.json 2> gpurun_out/fastv_$1.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/fastv_$1.json').read())
print("$1: dense fast %.4f step %.4f | natural fast %.4f step %.4f" % (d['stages_ms_per_step']['fast'], d['ms_per_step'], d['natural_images']['stages_ms_per_step']['fast'], d['natural_images']['ms_per_step']))
PY
}
for V in "$@"; do
  case $V in
    old) cp tools/orb_fast_r1.cu.txt monocular_slam_b200/csrc/orb_fast.cu; FLAGS="-DFT_OLD_SIGNATURE" ;;
    *) cp /tmp/orb_fast_new.cu monocular_slam_b200/csrc/orb_fast.cu; FLAGS="$V" ;;
  esac
  ORBX_EXTRA_FLAGS="$FLAGS" sh monocular_slam_b200/csrc/build.sh > /tmp/build.log 2>&1 || { tail -20 /tmp/build.log; exit 1; }
  run "$(echo $V | tr -c 'A-Za-z0-9\n' '_')"
done
cp /tmp/orb_fast_new.cu monocular_slam_b200/csrc/orb_fast.cu
