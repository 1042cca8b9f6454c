cd $GRAFT_REPO_ROOT
R=r02x
timeout 600 python bench.py --workload config5 --guides 100 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/${R}_config5_100guides.json 2> gpurun_out/${R}_config5_100guides.err; echo c5_100 rc=$?
python - <<PY
import json
for t in ("config5_100guides",):
    try:
        d=json.load(open("gpurun_out/${R}_%s.json"%t)); print(t, round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms", round(d["ms_per_step"],2), d["breakdown_ms"], d.get("parity_check"))
    except Exception as ex:
        print(t,"ERR",ex); print(open("gpurun_out/${R}_%s.err"%t).read()[-800:])
PY
