# ncu --set full of the three bag-of-words kernels on the bench shape (run under gpurun after tools/bow_probe.py exited 0):
# sh tools/bow_ncu.sh  ->  gpurun_out/bow_{descend,build,score}.ncu-rep + a launch list
set -e
python tools/bow_probe.py > gpurun_out/bow_probe.txt 2>&1
for K in descend build score; do
  ncu --set full --import-source on --clock-control none -k regex:k_bow_$K --launch-skip 4 --launch-count 1 -f -o gpurun_out/bow_$K \
      python tools/bow_probe.py > gpurun_out/bow_ncu_$K.log 2>&1
done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_bow_ --launch-count 40 --csv --log-file gpurun_out/bow_launches.csv \
    python tools/bow_probe.py > /dev/null 2>&1
