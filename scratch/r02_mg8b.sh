# reduced 8-GPU refresh on the final tree: default, config2, config4 (ngg+nag); no CPU baseline (rank 0 would spend 12 s of 8-GPU box time on it)
cd $GRAFT_REPO_ROOT
N=${1:-8}
R=r02b_${N}gpu
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_default.json 2> gpurun_out/${R}_default.err; echo default rc=$?
timeout 300 $TR bench.py --gpus $N --workload config2 --steps 5 --warmup 2 --no-cpu-baseline > gpurun_out/${R}_config2.json 2> gpurun_out/${R}_config2.err; echo c2 rc=$?
timeout 300 $TR bench.py --gpus $N --workload config4 --steps 5 --warmup 2 --no-cpu-baseline > gpurun_out/${R}_config4.json 2> gpurun_out/${R}_config4.err; echo c4 rc=$?
python - <<PY
import json
for t in ("default","config2","config4"):
    try:
        d=json.load(open("gpurun_out/${R}_%s.json"%t)); print(t, round(d["value"],1), d["unit"], "e2e", round(d["e2e"]["value"],1), "ms", round(d["ms_per_step"],2), d.get("per_rank",{}).get("ms_per_step"), d.get("parity_check"), d["clocks"])
    except Exception as ex:
        print(t,"ERR",ex); print(open("gpurun_out/${R}_%s.err"%t).read()[-800:])
PY
