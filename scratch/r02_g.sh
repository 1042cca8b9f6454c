cd $GRAFT_REPO_ROOT
R=r02g
C4="--workload config4 --guides 16 --scale 0.25 --steps 1 --warmup 1 --no-cpu-baseline --no-parity-check"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_canon -s 1 -c 1 -o gpurun_out/${R}_canon6 python bench.py $C4 > gpurun_out/${R}_ncu3.log 2>&1; echo rc=$?
timeout 900 python -m pytest tests -x -q -m gpu -k "vcf or shards or cli" > gpurun_out/${R}_pytest.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/${R}_pytest.log
