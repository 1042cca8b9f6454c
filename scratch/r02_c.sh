cd $GRAFT_REPO_ROOT
R=r02c
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/${R}_pytest.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/${R}_pytest.log
timeout 600 python bench.py --steps 3 --warmup 2 --cpu-seconds 4 > gpurun_out/${R}_default.json 2> gpurun_out/${R}_default.err; echo default rc=$?; tail -3 gpurun_out/${R}_default.err
CALITAS_TRACE=1 timeout 400 python bench.py --workload config4 --steps 2 --warmup 2 --cpu-seconds 4 > gpurun_out/${R}_c4.json 2> gpurun_out/${R}_c4.err; echo c4 rc=$?
CALITAS_TRACE=1 timeout 300 python bench.py --scale 0.125 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/${R}_eighth.json 2> gpurun_out/${R}_eighth.err; echo eighth rc=$?
C4="--workload config4 --guides 16 --scale 0.25 --steps 1 --warmup 1 --no-cpu-baseline --no-parity-check"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${R}_c4_launches.csv python bench.py $C4 > gpurun_out/${R}_ncu1.log 2>&1; echo rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_align_fast -s 1 -c 1 -o gpurun_out/${R}_align6 python bench.py $C4 > gpurun_out/${R}_ncu2.log 2>&1; echo rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_canon -s 1 -c 1 -o gpurun_out/${R}_canon6 python bench.py $C4 > gpurun_out/${R}_ncu3.log 2>&1; echo rc=$?
