cd $GRAFT_REPO_ROOT
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 500 python scratch/bench_a2r.py --tasks 1000000 --reps 2 > gpurun_out/a2r.json 2> gpurun_out/a2r.err; echo rc=$?; tail -3 gpurun_out/a2r.err; cat gpurun_out/a2r.json
