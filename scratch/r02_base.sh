# round-2 baseline data: d=6 (config 4) trace + ncu of the shipped k_align / k_scan_tiled at d=6, k_align_group, 1/8-genome trace
cd $GRAFT_REPO_ROOT
R=r02a
C4="--max-guide-diffs 6 --max-gaps 2 --pam ngg --aux-pams nag --no-cpu-baseline"
CALITAS_TRACE=1 timeout 400 python bench.py --guides 100 --steps 2 --warmup 1 $C4 > gpurun_out/${R}_c4.json 2> gpurun_out/${R}_c4.err; echo c4 rc=$?
CALITAS_TRACE=1 timeout 300 python bench.py --guides 100 --scale 0.125 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/${R}_eighth.json 2> gpurun_out/${R}_eighth.err; echo eighth rc=$?
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${R}_c4_launches.csv python bench.py --guides 16 --scale 0.25 --steps 1 --warmup 1 $C4 > gpurun_out/${R}_ncu1.log 2>&1; echo rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_align -s 1 -c 1 -o gpurun_out/${R}_align6 python bench.py --guides 16 --scale 0.25 --steps 1 --warmup 1 $C4 > gpurun_out/${R}_ncu2.log 2>&1; echo rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_scan_tiled -s 1 -c 1 -o gpurun_out/${R}_scan6 python bench.py --guides 16 --scale 0.25 --steps 1 --warmup 1 $C4 > gpurun_out/${R}_ncu3.log 2>&1; echo rc=$?
timeout 300 python scratch/bench_a2r.py --tasks 1000000 --reps 2 > gpurun_out/${R}_a2r.json 2> gpurun_out/${R}_a2r.err; echo a2r rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_align_group -c 1 -o gpurun_out/${R}_group python scratch/bench_a2r.py --tasks 200000 --reps 1 --scale 0.1 > gpurun_out/${R}_ncu4.log 2>&1; echo rc=$?
ls -la gpurun_out/${R}_*
