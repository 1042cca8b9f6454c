set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --guides 100 --steps 3 --warmup 3 > gpurun_out/r01b_bench100.json 2> gpurun_out/r01b_bench100.err
python bench.py --guides 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r01b_bench1.json 2> gpurun_out/r01b_bench1.err
python bench.py --guides 16 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r01b_plain16.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01b_launches.csv python bench.py --guides 16 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r01b_ncu1.log 2>&1
python bench.py --guides 16 --scale 0.25 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r01b_plain16q.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_scan_tiled -s 1 -c 1 -o gpurun_out/r01b_scan python bench.py --guides 16 --scale 0.25 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r01b_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_align -s 1 -c 1 -o gpurun_out/r01b_align python bench.py --guides 16 --scale 0.25 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r01b_ncu3.log 2>&1
tail -c 600 gpurun_out/r01b_bench100.json
