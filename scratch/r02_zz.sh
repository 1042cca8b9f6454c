cd $GRAFT_REPO_ROOT
R=r02zz
timeout 100 python bench.py --workload config5 --guides 100 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/${R}_config5_100guides.json 2> gpurun_out/${R}_config5_100guides.err; echo c5_100 rc=$?
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${R}_config5_100guides.json")); print(round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms", round(d["ms_per_step"],2), d["breakdown_ms"], d["counts"], d.get("parity_check"), d["gpu_launches"])
except Exception as ex:
    print("ERR",ex); print(open("gpurun_out/${R}_config5_100guides.err").read()[-800:])
PY
