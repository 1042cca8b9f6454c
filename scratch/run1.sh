cd $GRAFT_REPO_ROOT
for g in 16 100; do
timeout 300 python bench.py --guides $g --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/p_$g.json 2> gpurun_out/p_$g.err; tail -3 gpurun_out/p_$g.err
python - <<PY
import json
d=json.load(open("gpurun_out/p_$g.json")); print($g, round(d["value"],1), round(d["e2e"]["value"],1), round(d["ms_per_step"],2), {k: round(v,2) for k,v in d["breakdown_ms"].items()}, d["counts"]["hits"], d["gpu_launches"])
PY
done
