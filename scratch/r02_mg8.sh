# 8-GPU evidence of round 2: every BASELINE configuration that names 8 x B200, plus the one-process / one-table mode and the multi-device CLI test
cd $GRAFT_REPO_ROOT
N=${1:-8}
R=r02_${N}gpu
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
nvidia-smi -L | head -8
timeout 600 python -m pytest tests/test_cli.py tests/test_sharding.py -x -q -m gpu -k "sharded" > gpurun_out/${R}_pytest_multidevice.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/${R}_pytest_multidevice.log
CALITAS_TRACE=1 timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${R}_default.json 2> gpurun_out/${R}_default.err; echo default rc=$?
timeout 600 $TR bench.py --gpus $N --workload config4 --steps 5 --warmup 2 > gpurun_out/${R}_config4.json 2> gpurun_out/${R}_config4.err; echo c4 rc=$?
timeout 600 $TR bench.py --gpus $N --workload config4 --pam "" --aux-pams "" --steps 5 --warmup 2 > gpurun_out/${R}_config4_pamless.json 2> gpurun_out/${R}_config4_pamless.err; echo c4p rc=$?
timeout 600 $TR bench.py --gpus $N --workload config5 --steps 10 --warmup 3 > gpurun_out/${R}_config5.json 2> gpurun_out/${R}_config5.err; echo c5 rc=$?
timeout 600 $TR bench.py --gpus $N --workload config5 --guides 100 --steps 3 --warmup 2 > gpurun_out/${R}_config5_100guides.json 2> gpurun_out/${R}_config5_100guides.err; echo c5_100 rc=$?
timeout 600 $TR bench.py --gpus $N --workload config3 --steps 20 --warmup 5 > gpurun_out/${R}_config3.json 2> gpurun_out/${R}_config3.err; echo c3 rc=$?
timeout 600 $TR bench.py --gpus $N --workload config2 --steps 5 --warmup 2 > gpurun_out/${R}_config2.json 2> gpurun_out/${R}_config2.err; echo c2 rc=$?
timeout 600 python bench.py --one-process --gpus $N --steps 5 --warmup 2 > gpurun_out/${R}_oneprocess.json 2> gpurun_out/${R}_oneprocess.err; echo onep rc=$?
python - <<PY
import json
for t in ("default","config4","config4_pamless","config5","config5_100guides","config3","config2","oneprocess"):
    try:
        d=json.load(open("gpurun_out/${R}_%s.json"%t)); print(t, round(d["value"],1), d["unit"], "e2e", round(d["e2e"]["value"],1), "ms", round(d["ms_per_step"],2), d.get("per_rank",{}).get("ms_per_step"), d.get("parity_check"), d["clocks"])
    except Exception as ex:
        print(t,"ERR",ex); print(open("gpurun_out/${R}_%s.err"%t).read()[-800:])
PY
