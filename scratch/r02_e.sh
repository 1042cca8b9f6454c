cd $GRAFT_REPO_ROOT
R=r02e
B="python bench.py --steps 4 --warmup 2 --no-cpu-baseline --no-parity-check"
run() { tag=$1; shift; env "$@" CALITAS_TRACE=1 timeout 300 $B $EXTRA > gpurun_out/${R}_$tag.json 2> gpurun_out/${R}_$tag.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${R}_$tag.json")); print("$tag", round(d["value"],1), round(d["ms_per_step"],2), d["breakdown_ms"])
except Exception as ex: print("$tag ERR", ex)
PY
}
EXTRA="--scale 0.125"
run e_two X=1
run e_one CALITAS_ONE_SCAN_STREAM=1
run e_p4 CALITAS_CHUNK_PLAN=24,24,24,16,8,4
EXTRA=""
run f_two X=1
run f_one CALITAS_ONE_SCAN_STREAM=1
EXTRA="--workload config4"
B="python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-parity-check"
run c4_two X=1
run c4_one CALITAS_ONE_SCAN_STREAM=1
