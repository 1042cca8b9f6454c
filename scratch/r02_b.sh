cd $GRAFT_REPO_ROOT
R=r02b
timeout 900 python -m pytest tests -x -q -m gpu -k "cheap or short_period or negative_overlap or defaults or variants_of_the_call" > gpurun_out/${R}_pytest.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/${R}_pytest.log
timeout 600 python bench.py --steps 3 --warmup 2 > gpurun_out/${R}_default.json 2> gpurun_out/${R}_default.err; echo default rc=$?; tail -3 gpurun_out/${R}_default.err
CALITAS_TRACE=1 timeout 400 python bench.py --workload config4 --steps 2 --warmup 1 --cpu-seconds 6 > gpurun_out/${R}_c4.json 2> gpurun_out/${R}_c4.err; echo c4 rc=$?
timeout 300 python bench.py --workload config2 --steps 3 --warmup 2 > gpurun_out/${R}_c2.json 2> gpurun_out/${R}_c2.err; echo c2 rc=$?; tail -3 gpurun_out/${R}_c2.err
timeout 300 python bench.py --workload config3 --steps 5 --warmup 3 --cpu-seconds 4 > gpurun_out/${R}_c3.json 2> gpurun_out/${R}_c3.err; echo c3 rc=$?; tail -3 gpurun_out/${R}_c3.err
timeout 300 python bench.py --workload config1 --steps 5 --warmup 3 --cpu-seconds 4 > gpurun_out/${R}_c1.json 2> gpurun_out/${R}_c1.err; echo c1 rc=$?; tail -3 gpurun_out/${R}_c1.err
CALITAS_TRACE=1 timeout 300 python bench.py --scale 0.125 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/${R}_eighth.json 2> gpurun_out/${R}_eighth.err; echo eighth rc=$?
