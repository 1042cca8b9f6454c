cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
CALITAS_DEDUP_TWO_SORTS=1 timeout 600 python -m pytest tests/test_parity_random.py tests/test_sharding.py -m gpu -x -q 2>&1 | tail -2
for lib in scratch/lib_mid.so calitas_b200/libcalitas_b200.so; do
  for sc in 0.125 1.0; do
  timeout 300 python bench.py --lib $lib --guides 100 --scale $sc --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ab.json 2> gpurun_out/ab.err
  python - <<PY
import json
d=json.load(open("gpurun_out/ab.json")); print("$lib scale $sc:", round(d["value"],1), round(d["ms_per_step"],2), {k: round(v,1) for k,v in d["breakdown_ms"].items()}, d["counts"]["hits"], d["gpu_launches"])
PY
  done
done
CALITAS_TRACE=1 timeout 300 python bench.py --guides 100 --scale 0.125 --steps 1 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep "calitas trace" | tail -8
