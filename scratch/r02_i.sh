cd $GRAFT_REPO_ROOT
R=r02i
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/${R}_pytest.log 2>&1; echo pytest rc=$?; tail -5 gpurun_out/${R}_pytest.log
CALITAS_TOOL_TIMING=1 timeout 900 python bench.py --workload config5 --steps 3 --warmup 1 > gpurun_out/${R}_c5.json 2> gpurun_out/${R}_c5.err; echo c5 rc=$?; tail -12 gpurun_out/${R}_c5.err
timeout 600 python bench.py --workload config5 --guides 100 --steps 2 --warmup 1 --no-parity-check > gpurun_out/${R}_c5_100.json 2> gpurun_out/${R}_c5_100.err; echo c5_100 rc=$?; tail -3 gpurun_out/${R}_c5_100.err
timeout 300 python bench.py --one-process --gpus 1 --steps 2 --warmup 1 > gpurun_out/${R}_onep1.json 2> gpurun_out/${R}_onep1.err; echo onep rc=$?; tail -3 gpurun_out/${R}_onep1.err
python - <<PY
import json
for t in ("c5","c5_100","onep1"):
    try:
        d=json.load(open("gpurun_out/${R}_%s.json"%t)); print(t, round(d["value"],1), round(d["e2e"]["value"],1), round(d["ms_per_step"],2), d.get("breakdown_ms"), d.get("counts"), d.get("setup_s"), d.get("tool_e2e"), d.get("parity_check"))
    except Exception as ex: print(t,"ERR",ex)
PY
