cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for lib in scratch/lib_mid2.so calitas_b200/libcalitas_b200.so; do
  timeout 300 python bench.py --lib $lib --guides 100 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ab.json 2> gpurun_out/ab.err
  python - <<PY
import json
d=json.load(open("gpurun_out/ab.json")); print("$lib default:", round(d["value"],1), round(d["ms_per_step"],2), {k: round(v,1) for k,v in d["breakdown_ms"].items()}, d["counts"]["hits"], d["gpu_launches"])
PY
  timeout 300 python bench.py --lib $lib --guides 100 --steps 2 --warmup 2 --no-cpu-baseline --max-guide-diffs 6 --max-gaps 2 --pam ngg --aux-pams nag > gpurun_out/ab.json 2> gpurun_out/ab.err
  python - <<PY
import json
d=json.load(open("gpurun_out/ab.json")); print("$lib config4:", round(d["value"],1), round(d["ms_per_step"],2), {k: round(v,1) for k,v in d["breakdown_ms"].items()}, d["counts"]["hits"], d["gpu_launches"])
PY
done
