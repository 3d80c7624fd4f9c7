for flags in "--no-cfg3 --no-single --no-natural --no-sustained" "--no-single --no-natural --no-sustained" "--no-cfg3 --no-natural --no-sustained" "--no-cfg3 --no-single --no-sustained" "--no-cfg3 --no-single --no-natural"; do
  python bench.py --steps 10 --warmup 3 --no-cpu --no-hamming --no-loop --no-triangulation $flags > gpurun_out/legs.json 2> gpurun_out/legs.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/legs.json').read())
f=d['fundamental']
print("$flags", "| e2e %.0f seq %.0f steady %.0f" % (d['e2e']['value'], f['sequence_pipeline']['value'], f['steady_state_pipeline']['value']))
PY
done
