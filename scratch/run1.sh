cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for sc in 1.0 0.125 0.25 0.5; do
timeout 300 python bench.py --guides 100 --scale $sc --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/p_s.json 2> gpurun_out/p_s.err
python - <<PY
import json
d=json.load(open("gpurun_out/p_s.json")); print("scale $sc", round(d["value"],1), round(d["ms_per_step"],2), {k: round(v,2) for k,v in d["breakdown_ms"].items()}, d["counts"]["hits"], d["gpu_launches"])
PY
done
CALITAS_TRACE=1 timeout 300 python bench.py --guides 100 --scale 0.125 --steps 1 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep "calitas trace" | tail -12
