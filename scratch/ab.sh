cd $GRAFT_REPO_ROOT
for v in prev new; do
  lib=""; [ $v = prev ] && lib="--lib scratch/lib_prev.so"
  for sc in 0.125 1.0; do
  timeout 300 python bench.py --guides 100 --scale $sc --steps 4 --warmup 3 --no-cpu-baseline $lib > gpurun_out/ab_${v}_$sc.json 2> gpurun_out/ab_$v.err
  python - <<PY
import json
d=json.load(open("gpurun_out/ab_${v}_$sc.json")); print("$v", $sc, round(d["value"],1), d["ms_per_step"], d["breakdown_ms"], d["counts"]["hits"])
PY
  done
done
