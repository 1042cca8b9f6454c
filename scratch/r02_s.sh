cd $GRAFT_REPO_ROOT
R=r02s
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/${R}_pytest.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/${R}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${R}_smoke.log 2>&1; echo smoke rc=$?; tail -2 gpurun_out/${R}_smoke.log
timeout 600 python bench.py > gpurun_out/${R}_default.json 2> gpurun_out/${R}_default.err; echo default rc=$?; cut -c1-400 gpurun_out/${R}_default.json
