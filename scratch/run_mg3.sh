cd $GRAFT_REPO_ROOT
N=$1
nproc; lscpu | grep -i "numa\|socket\|model name" | head -8; nvidia-smi topo -m 2>/dev/null | head -14
for bind in 0 1; do
CALITAS_BENCH_VERBOSE=1 CALITAS_BENCH_BIND=$bind python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/mg_bind$bind.json 2> gpurun_out/mg_bind$bind.err
echo "rc=$?"; grep "\[bench\]" gpurun_out/mg_bind$bind.err | sort | head -8
python - <<PY
import json
d=json.load(open("gpurun_out/mg_bind$bind.json")); print("bind=$bind", round(d["value"],1), round(d["ms_per_step"],2), d.get("per_rank")["ms_per_step"], round(d["e2e"]["value"],1))
PY
done
