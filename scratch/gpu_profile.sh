cd $GRAFT_REPO_ROOT
timeout 600 python bench.py > gpurun_out/r01d_bench_default.json 2> gpurun_out/r01d_bench_default.err; echo rc=$?; tail -2 gpurun_out/r01d_bench_default.err
timeout 300 python bench.py --impl reference > gpurun_out/r01d_bench_reference.json 2> gpurun_out/r01d_ref.err; echo rc=$?
timeout 300 python bench.py --guides 16 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r01d_plain16.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r01d_launches.csv python bench.py --guides 16 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r01d_ncu1.log 2>&1
timeout 300 python bench.py --guides 16 --scale 0.25 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r01d_plain16q.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_scan_tiled -s 1 -c 1 -o gpurun_out/r01d_scan python bench.py --guides 16 --scale 0.25 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r01d_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_align5 -s 1 -c 1 -o gpurun_out/r01d_align python bench.py --guides 16 --scale 0.25 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r01d_ncu3.log 2>&1
timeout 300 python scratch/bench_a2r.py --tasks 200000 --scale 0.1 --reps 1 > gpurun_out/r01d_a2r_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_align_group|k_canon_warp" -c 2 -o gpurun_out/r01d_a2r python scratch/bench_a2r.py --tasks 200000 --scale 0.1 --reps 1 > gpurun_out/r01d_ncu4.log 2>&1
timeout 500 python scratch/bench_a2r.py --tasks 1000000 --reps 2 > gpurun_out/r01d_a2r.json 2> gpurun_out/a2r.err
tail -c 1500 gpurun_out/r01d_bench_default.json
