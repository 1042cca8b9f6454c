cd $GRAFT_REPO_ROOT
R=r02_1gpu
timeout 400 ncu --set full --clock-control none --import-source on -k regex:^k_scan_explicit$ -s 2 -c 1 -o gpurun_out/${R}_scanexplicit python bench.py --workload config5 --guides 16 --scale 0.25 --records 750000 --steps 1 --warmup 1 --no-cpu-baseline --no-parity-check > gpurun_out/${R}_ncu7.log 2>&1; echo rc=$?
