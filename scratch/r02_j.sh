cd $GRAFT_REPO_ROOT
R=r02j
timeout 1500 python -m pytest tests -x -q -m gpu -k "not fullsize" > gpurun_out/${R}_pytest.log 2>&1; echo pytest rc=$?; tail -5 gpurun_out/${R}_pytest.log
timeout 900 python bench.py --workload config5 --steps 5 --warmup 2 > gpurun_out/${R}_c5.json 2> gpurun_out/${R}_c5.err; echo c5 rc=$?; tail -3 gpurun_out/${R}_c5.err
timeout 600 python bench.py --workload config5 --guides 100 --steps 2 --warmup 1 --no-parity-check > gpurun_out/${R}_c5_100.json 2> gpurun_out/${R}_c5_100.err; echo c5_100 rc=$?; tail -3 gpurun_out/${R}_c5_100.err
timeout 300 python bench.py --workload config2 --steps 3 --warmup 2 > gpurun_out/${R}_c2.json 2> gpurun_out/${R}_c2.err; echo c2 rc=$?; tail -3 gpurun_out/${R}_c2.err
CALITAS_NO_GROUP_WARP=1 timeout 300 python bench.py --workload config2 --steps 3 --warmup 2 --no-cpu-baseline --no-parity-check > gpurun_out/${R}_c2_old.json 2> gpurun_out/${R}_c2_old.err; echo c2old rc=$?
python - <<PY
import json
for t in ("c5","c5_100","c2","c2_old"):
    try:
        d=json.load(open("gpurun_out/${R}_%s.json"%t)); print(t, round(d["value"],1), round(d["e2e"]["value"],1), round(d["ms_per_step"],2), d.get("breakdown_ms"), d.get("counts"), d.get("setup_s"), d.get("tool_e2e"), d.get("parity_check"))
    except Exception as ex: print(t,"ERR",ex)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_align_group_warp -c 1 -o gpurun_out/${R}_groupwarp python bench.py --workload config2 --tasks 200000 --scale 0.1 --steps 1 --warmup 0 --no-cpu-baseline --no-parity-check > gpurun_out/${R}_ncu4.log 2>&1; echo rc=$?
