cd $GRAFT_REPO_ROOT
for v in "ngg:nag" "nrg:" ":"; do
pam=${v%%:*}; aux=${v##*:}
timeout 300 python bench.py --guides 100 --steps 2 --warmup 2 --no-cpu-baseline --max-guide-diffs 6 --max-gaps 2 --pam "$pam" --aux-pams "$aux" > gpurun_out/c4_${pam}_${aux}.json 2> gpurun_out/c4.err; echo rc=$?; tail -3 gpurun_out/c4.err
python - <<PY
import json
d=json.load(open("gpurun_out/c4_${pam}_${aux}.json")); print("$v", round(d["value"],1), round(d["e2e"]["value"],1), d["ms_per_step"], d["breakdown_ms"], d["counts"]["hits"], d["counts"]["candidates"])
PY
done
