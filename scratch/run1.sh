cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for g in 100 16; do
timeout 300 python bench.py --guides $g --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/p_$g.json 2> gpurun_out/p_$g.err; tail -3 gpurun_out/p_$g.err
python - <<PY
import json
d=json.load(open("gpurun_out/p_$g.json")); print($g, round(d["value"],1), round(d["e2e"]["value"],1), round(d["ms_per_step"],2), {k: round(v,2) for k,v in d["breakdown_ms"].items()}, d["counts"]["hits"])
PY
done
timeout 300 python bench.py --guides 100 --scale 0.125 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/p_s.json 2> gpurun_out/p_s.err
python - <<PY
import json
d=json.load(open("gpurun_out/p_s.json")); print("1/8", round(d["value"],1), round(d["ms_per_step"],2), {k: round(v,2) for k,v in d["breakdown_ms"].items()}, d["counts"]["hits"])
PY
timeout 300 python bench.py --guides 100 --steps 2 --warmup 2 --no-cpu-baseline --max-guide-diffs 6 --max-gaps 2 --pam ngg --aux-pams nag > gpurun_out/c4.json 2> gpurun_out/c4.err
python - <<PY
import json
d=json.load(open("gpurun_out/c4.json")); print("config4", round(d["value"],1), round(d["ms_per_step"],2), {k: round(v,2) for k,v in d["breakdown_ms"].items()}, d["counts"]["hits"])
PY
