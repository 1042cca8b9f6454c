cd $GRAFT_REPO_ROOT
N=$1
env | grep -i nccl
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/mg_$N.json 2> gpurun_out/mg_$N.err
echo "rc=$?"; tail -3 gpurun_out/mg_$N.err; head -c 200 gpurun_out/mg_$N.json; echo; wc -l gpurun_out/mg_$N.json
