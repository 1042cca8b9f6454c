cd $GRAFT_REPO_ROOT
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r01f_bench_100guides_${N}gpu.json 2> gpurun_out/mg_$N.err
echo "rc=$?"; tail -3 gpurun_out/mg_$N.err
python - <<PY
import json
d=json.load(open("gpurun_out/r01f_bench_100guides_${N}gpu.json")); print(round(d["value"],1), round(d["ms_per_step"],2), d.get("per_rank"), d["clocks"], round(d["e2e"]["value"],1))
PY
