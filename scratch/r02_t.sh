cd $GRAFT_REPO_ROOT
R=r02_2gpu
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 python -m pytest tests/test_cli.py tests/test_sharding.py -x -q -m gpu -k "sharded" > gpurun_out/${R}_pytest_multidevice.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/${R}_pytest_multidevice.log
timeout 300 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_default.json 2> gpurun_out/${R}_default.err; echo default rc=$?
timeout 300 python bench.py --one-process --gpus 2 --steps 5 --warmup 2 > gpurun_out/${R}_oneprocess.json 2> gpurun_out/${R}_oneprocess.err; echo onep rc=$?
python - <<PY
import json
for t in ("default","oneprocess"):
    try:
        d=json.load(open("gpurun_out/${R}_%s.json"%t)); print(t, round(d["value"],1), d["unit"], "e2e", round(d["e2e"]["value"],1), "ms", round(d["ms_per_step"],2), d.get("per_rank",{}).get("ms_per_step"), d.get("parity_check"), d["clocks"])
    except Exception as ex:
        print(t,"ERR",ex); print(open("gpurun_out/${R}_%s.err"%t).read()[-800:])
PY
