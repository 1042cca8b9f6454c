import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo/oracle')
from conftest import read_fasta
import calitas_b200.testing as t
import pyoracle
ref = read_fasta('/root/repo/tests/golden/sga_test_ref.fa')
q = dict(ref)['chr1'][49:69]
for be in (pyoracle, t):
    rows = be.align_to_ref(ref, q, 'chr1', 65, best=False, max_guide_diffs=20, max_gaps=3, max_pam_diffs=0, max_total_diffs=23, max_overlap=0)
    print(be.__name__, [(r['startOffset'], r['endOffset'], r['strand'], r['score'], r['cigar']) for r in rows][:6])
    rows = be.align_to_ref(ref, q, 'chr1', 65, best=False, max_guide_diffs=20, max_gaps=3, max_pam_diffs=0, max_total_diffs=23, max_overlap=1000)
    print(be.__name__, len(rows), [(r['startOffset'], r['endOffset'], r['strand'], r['score'], r['cigar']) for r in rows if r['strand']=='+'][:50])
