cd $GRAFT_REPO_ROOT
N=$1
nvidia-smi -L
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/mg_$N.json 2> gpurun_out/mg_$N.err
echo "rc=$?"; tail -20 gpurun_out/mg_$N.err; head -c 600 gpurun_out/mg_$N.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/mgref_$N.json 2> gpurun_out/mgref_$N.err
echo "rc=$?"; head -c 400 gpurun_out/mgref_$N.json
