cd $GRAFT_REPO_ROOT
R=r02m
timeout 900 python -m pytest tests -x -q -m gpu -k "best or align_to_reference or a2r or pairwise or golden or frozen" > gpurun_out/${R}_pytest.log 2>&1; echo pytest rc=$?; tail -2 gpurun_out/${R}_pytest.log
timeout 300 python bench.py --workload config2 --steps 5 --warmup 3 > gpurun_out/${R}_c2.json 2> gpurun_out/${R}_c2.err; echo c2 rc=$?; tail -2 gpurun_out/${R}_c2.err
CALITAS_NO_ALL_COLUMNS=1 timeout 300 python bench.py --workload config2 --steps 5 --warmup 3 --no-cpu-baseline --no-parity-check > gpurun_out/${R}_c2_scan.json 2> gpurun_out/${R}_c2_scan.err; echo c2scan rc=$?
python - <<PY
import json
for t in ("c2","c2_scan"):
    try:
        d=json.load(open("gpurun_out/${R}_%s.json"%t)); print(t, round(d["value"],1), round(d["e2e"]["value"],1), round(d["ms_per_step"],2), d.get("breakdown_ms"), d.get("counts"), d.get("parity_check"))
    except Exception as ex: print(t,"ERR",ex)
PY
