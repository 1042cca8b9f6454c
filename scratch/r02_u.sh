cd $GRAFT_REPO_ROOT
R=r02u
timeout 600 python -m pytest tests -x -q -m gpu -k "parity_random or golden or errors" > gpurun_out/${R}_pytest.log 2>&1; echo pytest rc=$?; tail -2 gpurun_out/${R}_pytest.log
run() { tag=$1; shift; timeout 300 python bench.py --steps 4 --warmup 2 --no-cpu-baseline "$@" > gpurun_out/${R}_$tag.json 2> gpurun_out/${R}_$tag.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${R}_$tag.json")); print("$tag", round(d["value"],1), round(d["ms_per_step"],2), d["breakdown_ms"], d.get("parity_check"))
except Exception as ex: print("$tag ERR", ex)
PY
}
run c4 --workload config4
run f
