cd $GRAFT_REPO_ROOT
R=r01f
timeout 600 python bench.py > gpurun_out/${R}_bench_default.json 2> gpurun_out/${R}_bench_default.err; echo rc=$?; tail -2 gpurun_out/${R}_bench_default.err
timeout 300 python bench.py --impl reference > gpurun_out/${R}_bench_reference.json 2> gpurun_out/${R}_ref.err; echo rc=$?
timeout 300 python bench.py --guides 16 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${R}_plain16.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${R}_launches.csv python bench.py --guides 16 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${R}_ncu1.log 2>&1
timeout 300 python bench.py --guides 16 --scale 0.25 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${R}_plain16q.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_scan_tiled -s 1 -c 1 -o gpurun_out/${R}_scan python bench.py --guides 16 --scale 0.25 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${R}_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_align5 -s 1 -c 1 -o gpurun_out/${R}_align python bench.py --guides 16 --scale 0.25 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${R}_ncu3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sweep -c 1 -o gpurun_out/${R}_sweep python bench.py --guides 16 --scale 0.25 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${R}_ncu4.log 2>&1
timeout 500 python scratch/bench_a2r.py --tasks 1000000 --reps 2 > gpurun_out/${R}_a2r.json 2> gpurun_out/a2r.err
tail -c 600 gpurun_out/${R}_bench_default.json; ls -la gpurun_out/${R}_*
