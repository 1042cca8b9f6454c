cd $GRAFT_REPO_ROOT
for v in base bytes; do
  lib=""; [ $v != base ] && lib="--lib scratch/lib_$v.so"
  timeout 300 python bench.py --guides 16 --scale 0.5 --steps 3 --warmup 2 --no-cpu-baseline $lib > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err; tail -2 gpurun_out/ab_$v.err
  python - <<PY
import json
d=json.load(open("gpurun_out/ab_$v.json")); print("$v", round(d["value"],1), d["breakdown_ms"], d["counts"]["hits"], d["counts"]["candidates"])
PY
done
