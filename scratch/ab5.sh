cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_parity_random.py tests/test_golden_sequential_guide_aligner.py tests/test_cli.py tests/test_errors_and_limits.py -m gpu -x -q 2>&1 | tail -2
for lib in scratch/lib_mid3.so calitas_b200/libcalitas_b200.so; do
timeout 300 python scratch/bench_a2r.py --lib $lib --tasks 1000000 --scale 0.2 --reps 2 > gpurun_out/ab_a2r.json 2> gpurun_out/a2r.err
python - <<PY
import json
d=json.load(open("gpurun_out/ab_a2r.json")); b=d["best"]; a=d["all_d5_p1_O10"]
print("$lib best:", round(b["tasks_per_s"]/1e6,2), "M/s dev", round(b["dev_ms"],1), "align", round(b["align_ms"],1), "hits", b["hits"], "| all:", round(a["tasks_per_s"]/1e6,1), "M/s")
PY
done
cp gpurun_out/ab_a2r.json gpurun_out/r01g_a2r_scale02.json
