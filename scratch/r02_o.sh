cd $GRAFT_REPO_ROOT
R=r02o
python - <<'PY' > gpurun_out/r02o_micro.json 2> gpurun_out/r02o_micro.err
import json
from calitas_b200._capi import Engine
e = Engine(0)
names = ["alu_lop3","fma_imad","both_lop3_imad","fma_imad_hi","alu_lea_hi","vimnmx3","vimnmx3_s16x2","viaddmnmx_s16x2","shf","lop3_2r_imad_2r","vimnmx3x2_imad","lop3_imm"]
print(json.dumps({n: e.microbench_int(i) for i, n in enumerate(names)}))
PY
cat gpurun_out/r02o_micro.json; tail -3 gpurun_out/r02o_micro.err
timeout 600 python -m pytest tests -x -q -m gpu -k "parity_random or golden" > gpurun_out/${R}_pytest.log 2>&1; echo pytest rc=$?; tail -2 gpurun_out/${R}_pytest.log
run() { tag=$1; shift; timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline "$@" > gpurun_out/${R}_$tag.json 2> gpurun_out/${R}_$tag.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${R}_$tag.json")); print("$tag", round(d["value"],1), round(d["ms_per_step"],2), d["breakdown_ms"], d.get("parity_check"))
except Exception as ex: print("$tag ERR", ex)
PY
}
run c4 --workload config4
