cd $GRAFT_REPO_ROOT
R=r02l
CALITAS_SCAN_INLINE_EMIT=1 timeout 900 python -m pytest tests -x -q -m gpu -k "parity_random or golden" > gpurun_out/${R}_pytest_inline.log 2>&1; echo pytest rc=$?; tail -2 gpurun_out/${R}_pytest_inline.log
run() { tag=$1; shift; env "$@" CALITAS_TRACE=1 timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-parity-check $EXTRA > gpurun_out/${R}_$tag.json 2> gpurun_out/${R}_$tag.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${R}_$tag.json")); print("$tag", round(d["value"],1), round(d["ms_per_step"],2), d["breakdown_ms"])
except Exception as ex: print("$tag ERR", ex)
PY
}
EXTRA="--workload config4"
run c4_replay CALITAS_SCAN_INLINE_EMIT=0
run c4_inline CALITAS_SCAN_INLINE_EMIT=1
run c4_replay2 CALITAS_SCAN_INLINE_EMIT=0
EXTRA=""
run f_replay CALITAS_SCAN_INLINE_EMIT=0
run f_inline CALITAS_SCAN_INLINE_EMIT=1
