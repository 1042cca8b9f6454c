cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python scratch/make_fasta.py --out /tmp/hg.fa --scale 1.0 --guides 100
head -1 /tmp/hg.guides.tsv > /tmp/g1.tsv; head -10 /tmp/hg.guides.tsv > /tmp/g10.tsv; cp /tmp/hg.guides.tsv /tmp/g100.tsv
for gf in g1 g10 g100; do
  s=$(date +%s%N)
  CALITAS_TOOL_TIMING=1 timeout 900 ./calitas_b200/calitas SearchReference --guides-file /tmp/$gf.tsv -r /tmp/hg.fa -o /tmp/out_$gf.tsv --stats 2> gpurun_out/cli3_$gf.err; echo rc=$?
  e=$(date +%s%N); echo "wall_ms $(( (e - s) / 1000000 ))"
  cat gpurun_out/cli3_$gf.err; ls -la /tmp/out_$gf.tsv; md5sum /tmp/out_$gf.tsv | cut -c1-32; rm -f /tmp/out_$gf.tsv
done
