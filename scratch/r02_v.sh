cd $GRAFT_REPO_ROOT
R=r02_1gpu
Q="--scale 0.25 --guides 16 --steps 1 --warmup 1 --no-cpu-baseline --no-parity-check"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_canon$ -s 1 -c 1 -o gpurun_out/${R}_canon6 python bench.py --workload config4 $Q > gpurun_out/${R}_ncu4.log 2>&1; echo rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_canon_rank$ -s 1 -c 1 -o gpurun_out/${R}_canonrank6 python bench.py --workload config4 $Q > gpurun_out/${R}_ncu4b.log 2>&1; echo rc=$?
