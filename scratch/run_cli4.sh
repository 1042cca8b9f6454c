cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_cli.py tests/test_cabi_exports.py -m gpu -x -q 2>&1 | tail -2
timeout 300 python scratch/make_fasta.py --out /tmp/hg.fa --scale 1.0 --guides 100
cp /tmp/hg.guides.tsv /tmp/g100.tsv
for gf in g100; do
  s=$(date +%s%N)
  CALITAS_TOOL_TIMING=1 timeout 900 ./calitas_b200/calitas SearchReference --guides-file /tmp/$gf.tsv -r /tmp/hg.fa -o /tmp/out_$gf.tsv --stats 2> gpurun_out/cli4_$gf.err; echo rc=$?
  e=$(date +%s%N); echo "wall_ms $(( (e - s) / 1000000 ))"
  cat gpurun_out/cli4_$gf.err; ls -la /tmp/out_$gf.tsv; wc -l /tmp/out_$gf.tsv; rm -f /tmp/out_$gf.tsv
done
