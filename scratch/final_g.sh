cd $GRAFT_REPO_ROOT
R=r01g
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/${R}_bench_default.json 2> gpurun_out/${R}_bench_default.err; echo rc=$?; tail -2 gpurun_out/${R}_bench_default.err
timeout 300 python bench.py --impl reference > gpurun_out/${R}_bench_reference.json 2> gpurun_out/${R}_ref.err; echo rc=$?
python - <<PY
import json
d=json.load(open("gpurun_out/r01g_bench_default.json")); print(round(d["value"],1), round(d["ms_per_step"],2), round(d["e2e"]["value"],1), d["gpu_launches"], round(d["roofline_int"]["frac"],3), {k: round(v,1) for k,v in d["breakdown_ms"].items()}, d["clocks"])
PY
