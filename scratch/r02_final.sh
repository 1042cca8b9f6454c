# round-2 evidence on one B200: tests, smoke, every bench workload (both arms for the default), ncu launch lists and full captures of the kernels DESIGN.md cites
cd $GRAFT_REPO_ROOT
R=r02_1gpu
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/${R}_pytest.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/${R}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${R}_smoke.log 2>&1; echo smoke rc=$?; tail -2 gpurun_out/${R}_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${R}_default.json 2> gpurun_out/${R}_default.err; echo default rc=$?; tail -2 gpurun_out/${R}_default.err
timeout 300 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/${R}_reference.json 2> gpurun_out/${R}_reference.err; echo ref rc=$?
timeout 300 python bench.py --workload config1 --steps 20 --warmup 5 > gpurun_out/${R}_config1.json 2> gpurun_out/${R}_config1.err; echo c1 rc=$?
timeout 300 python bench.py --workload config2 --steps 5 --warmup 3 > gpurun_out/${R}_config2.json 2> gpurun_out/${R}_config2.err; echo c2 rc=$?
timeout 300 python bench.py --workload config3 --steps 20 --warmup 5 --cpu-seconds 5 > gpurun_out/${R}_config3.json 2> gpurun_out/${R}_config3.err; echo c3 rc=$?
timeout 600 python bench.py --workload config4 --steps 5 --warmup 2 --cpu-seconds 8 > gpurun_out/${R}_config4_ngg_nag.json 2> gpurun_out/${R}_config4_ngg_nag.err; echo c4 rc=$?
timeout 600 python bench.py --workload config4 --pam nrg --aux-pams "" --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/${R}_config4_nrg.json 2> gpurun_out/${R}_config4_nrg.err; echo c4nrg rc=$?
timeout 600 python bench.py --workload config4 --pam "" --aux-pams "" --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/${R}_config4_pamless.json 2> gpurun_out/${R}_config4_pamless.err; echo c4p rc=$?
CALITAS_TOOL_TIMING=1 timeout 600 python bench.py --workload config5 --steps 10 --warmup 3 > gpurun_out/${R}_config5_1guide.json 2> gpurun_out/${R}_config5_1guide.err; echo c5 rc=$?; grep "calitas tool" gpurun_out/${R}_config5_1guide.err | tail -8
timeout 600 python bench.py --workload config5 --guides 100 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/${R}_config5_100guides.json 2> gpurun_out/${R}_config5_100guides.err; echo c5_100 rc=$?
python - <<PY
import json
for t in ("default","reference","config1","config2","config3","config4_ngg_nag","config4_nrg","config4_pamless","config5_1guide","config5_100guides"):
    try:
        d=json.load(open("gpurun_out/${R}_%s.json"%t)); print(t, round(d["value"],2), d["unit"], "e2e", round(d["e2e"]["value"],2), "ms", round(d["ms_per_step"],2), (d.get("roofline_int") or {}).get("frac"), d.get("parity_check"), (d.get("cpu_baseline") or {}).get("value"))
    except Exception as ex:
        print(t,"ERR",ex); print(open("gpurun_out/${R}_%s.err"%t).read()[-600:])
PY
NC="--no-cpu-baseline --no-parity-check"
D16="--guides 16 --steps 1 --warmup 1 $NC"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${R}_launches_default16.csv python bench.py $D16 > gpurun_out/${R}_ncu_a.log 2>&1; echo rc=$?
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${R}_launches_config4.csv python bench.py --workload config4 --scale 0.25 $D16 > gpurun_out/${R}_ncu_b.log 2>&1; echo rc=$?
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${R}_launches_config2.csv python bench.py --workload config2 --tasks 200000 --scale 0.1 --steps 1 --warmup 1 $NC > gpurun_out/${R}_ncu_c.log 2>&1; echo rc=$?
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/${R}_launches_config5.csv python bench.py --workload config5 --guides 16 --scale 0.25 --records 750000 --steps 1 --warmup 1 $NC > gpurun_out/${R}_ncu_d.log 2>&1; echo rc=$?
Q="--scale 0.25 $D16"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_scan_tiled -s 1 -c 1 -o gpurun_out/${R}_scan5 python bench.py $Q > gpurun_out/${R}_ncu1.log 2>&1; echo rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_align_fast5 -s 1 -c 1 -o gpurun_out/${R}_align5 python bench.py $Q > gpurun_out/${R}_ncu2.log 2>&1; echo rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_align_fast6 -s 1 -c 1 -o gpurun_out/${R}_align6 python bench.py --workload config4 $Q > gpurun_out/${R}_ncu3.log 2>&1; echo rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_canon$ -s 1 -c 1 -o gpurun_out/${R}_canon6 python bench.py --workload config4 $Q > gpurun_out/${R}_ncu4.log 2>&1; echo rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_scan_tiled -s 1 -c 1 -o gpurun_out/${R}_scan6 python bench.py --workload config4 $Q > gpurun_out/${R}_ncu5.log 2>&1; echo rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_align_group_warp -c 1 -o gpurun_out/${R}_groupwarp python bench.py --workload config2 --tasks 200000 --scale 0.1 --steps 1 --warmup 0 $NC > gpurun_out/${R}_ncu6.log 2>&1; echo rc=$?
ls -la gpurun_out/${R}_* | wc -l
