cd $GRAFT_REPO_ROOT
R=r02h
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/${R}_pytest.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/${R}_pytest.log
B="python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-parity-check"
run() { tag=$1; shift; env "$@" CALITAS_TRACE=1 timeout 300 $B $EXTRA > gpurun_out/${R}_$tag.json 2> gpurun_out/${R}_$tag.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${R}_$tag.json")); print("$tag", round(d["value"],1), round(d["ms_per_step"],2), d["breakdown_ms"])
except Exception as ex: print("$tag ERR", ex)
PY
}
EXTRA="--workload config4"
run c4 X=1
EXTRA=""
run f X=1
EXTRA="--scale 0.125"
run e X=1
C4="--workload config4 --guides 16 --scale 0.25 --steps 1 --warmup 1 --no-cpu-baseline --no-parity-check"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${R}_c4_launches.csv python bench.py $C4 > gpurun_out/${R}_ncu1.log 2>&1; echo rc=$?
