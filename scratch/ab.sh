cd $GRAFT_REPO_ROOT
for ch in 16 25 34 50; do
  for sc in 0.125 1.0; do
  CALITAS_CHUNK=$ch timeout 300 python bench.py --guides 100 --scale $sc --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ab.json 2> gpurun_out/ab.err
  python - <<PY
import json
d=json.load(open("gpurun_out/ab.json")); print("chunk $ch scale $sc:", round(d["value"],1), round(d["ms_per_step"],2), {k: round(v,1) for k,v in d["breakdown_ms"].items()}, d["counts"]["hits"])
PY
  done
done
