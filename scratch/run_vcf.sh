cd $GRAFT_REPO_ROOT
timeout 300 python scratch/make_fasta.py --out /tmp/hg.fa --scale 1.0 --guides 100
head -1 /tmp/hg.guides.tsv > /tmp/g1.tsv
timeout 900 python - <<'PY'
import sys, time
sys.path.insert(0, '.')
from calitas_b200 import synth
guides=[synth.BASELINE_GUIDE]+synth.random_guides(99)
g=synth.hg38_like_genome(1.0, guides=guides, sites_per_guide=200)
t=time.time(); arrays=[g.contig(c) for c in range(len(g.lengths))]
vcf=synth.synthetic_vcf(g, arrays, 1_000_000)
open('/tmp/v.vcf','w').write(vcf); print('vcf records', vcf.count('\n')-2, 'in', time.time()-t, 's')
PY
s=$(date +%s%N)
CALITAS_TOOL_TIMING=1 timeout 900 ./calitas_b200/calitas SearchReference --guides-file /tmp/g1.tsv -r /tmp/hg.fa -v /tmp/v.vcf -o /tmp/out_v.tsv --stats 2> gpurun_out/cli_vcf.err; echo rc=$?
e=$(date +%s%N); echo "wall_ms $(( (e - s) / 1000000 ))"
cat gpurun_out/cli_vcf.err; ls -la /tmp/out_v.tsv; grep -c "+variants" /tmp/out_v.tsv
