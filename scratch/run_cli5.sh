cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_cli.py -m gpu -x -q -k "forty or streams or guide_batch" 2>&1 | tail -2
timeout 300 python scratch/make_fasta.py --out /tmp/hg.fa --scale 1.0 --guides 100
s=$(date +%s%N)
CALITAS_TOOL_TIMING=1 timeout 600 ./calitas_b200/calitas SearchReference --guides-file /tmp/hg.guides.tsv -r /tmp/hg.fa -o /tmp/out.tsv --stats 2> gpurun_out/cli5_g100.err; echo rc=$?
e=$(date +%s%N); echo "wall_ms $(( (e - s) / 1000000 ))"
cat gpurun_out/cli5_g100.err; wc -lc /tmp/out.tsv
