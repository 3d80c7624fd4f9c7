# e2e and plain-copy floor with pinned vs write-combined host frames at N ranks: sh tools/wc_probe.sh N
N=${1:-4}
for F in "" "--wc-frames"; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --no-hamming --no-fundamental --no-natural --no-cfg3 --no-single --no-sustained --no-loop --no-ingest $F > gpurun_out/wc.json 2> gpurun_out/wc.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/wc.json').read())
e=d['e2e']
print("N=$N '$F': e2e %.0f fps (%.3f ms) blocking %.0f floor h2d %.3f ms (%.1f GB/s/GPU) d2h %.3f duplex %.3f over h2d %.3f over duplex %.3f" % (e['value'], e['ms_per_step'], e['blocking_value'], e['h2d_floor_ms'], e['h2d_floor_gbs_per_gpu'], e['d2h_floor_ms'], e['duplex_floor_ms'], e['over_h2d_floor'], e['over_duplex_floor']))
PY
done
