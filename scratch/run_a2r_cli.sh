cd $GRAFT_REPO_ROOT
timeout 300 python scratch/make_fasta.py --out /tmp/hg.fa --scale 1.0 --guides 100
timeout 300 python - <<'PY'
import sys, numpy as np
sys.path.insert(0, '.')
from calitas_b200 import synth
guides=[synth.BASELINE_GUIDE]+synth.random_guides(99)
g=synth.hg38_like_genome(1.0, guides=guides, sites_per_guide=200)
rng=np.random.default_rng(5); n=1_000_000
sites=[(c,pos,len(seq)) for c,lst in enumerate(g.planted) for (pos,seq) in lst]
near=rng.random(n)<0.5; si=rng.integers(0,len(sites),size=n); jit=rng.integers(-10,11,size=n)
p_contig=np.array(g.lengths,dtype=np.float64)/g.total(); cont=rng.choice(len(g.lengths),size=n,p=p_contig); u=rng.random(n); gi=rng.integers(0,len(guides),size=n)
with open('/tmp/tasks.tsv','w') as f:
    f.write('id\tquery\tchrom\tposition\n')
    for i in range(n):
        if near[i]:
            c,pos,ln=sites[si[i]]; p=pos+ln//2+int(jit[i])
        else:
            c=int(cont[i]); p=1+int(u[i]*(g.lengths[c]-1))
        f.write('t%d\t%s\t%s\t%d\n'%(i,guides[gi[i]],g.names[c],max(1,min(g.lengths[c],p))))
print('tasks written')
PY
for mode in "-w 60" "-w 60 -d 5 -p 1 -O 10"; do
  s=$(date +%s%N)
  CALITAS_TOOL_TIMING=1 timeout 900 ./calitas_b200/calitas AlignToReference -i /tmp/tasks.tsv -r /tmp/hg.fa -o /tmp/a2r_out.tsv $mode --stats 2> gpurun_out/cli_a2r.err; echo rc=$?
  e=$(date +%s%N); echo "mode [$mode] wall_ms $(( (e - s) / 1000000 ))"; cat gpurun_out/cli_a2r.err; wc -l /tmp/a2r_out.tsv
done
