cd $GRAFT_REPO_ROOT
R=r02k
run() { tag=$1; lib=$2; shift; shift; timeout 300 python bench.py --lib scratch/$lib --steps 4 --warmup 2 --no-cpu-baseline --no-parity-check "$@" > gpurun_out/${R}_$tag.json 2> gpurun_out/${R}_$tag.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${R}_$tag.json")); print("$tag", round(d["value"],1), round(d["ms_per_step"],2), d["breakdown_ms"])
except Exception as ex: print("$tag ERR", ex)
PY
}
for lib in lib_base.so lib_scan56.so lib_scan56_align64.so; do
run f_$lib $lib
run c4_$lib $lib --workload config4
run e_$lib $lib --scale 0.125
done
run f2_lib_base.so lib_base.so
