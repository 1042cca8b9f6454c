cd $GRAFT_REPO_ROOT
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_plain.log 2>&1 && echo plain-ok &&
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file gpurun_out/memcheck.log python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_san.log 2>&1; echo rc=$?
tail -5 gpurun_out/smoke_san.log; tail -15 gpurun_out/memcheck.log
