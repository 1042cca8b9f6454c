"""Extracts the fixture genome of SequentialGuideAlignerTest.scala:12-44 into sga_test_ref.fa.
Runs only where /root/reference exists (the build container); the output is committed."""
import os, re
SRC = "/root/reference/calitas/src/test/scala/com/editasmedicine/aligner/SequentialGuideAlignerTest.scala"
if __name__ == "__main__" and len(__import__("sys").argv) == 1:
    src = open(SRC).read()
    chr1 = re.findall(r'\.add\("([ACGT]{100})"\)', src)
    assert len(chr1) == 24
    chr2 = ["GATACaaCTCGTACTGTCAGT", "GATACGTCTCGTACTGTCAtT"]
    with open(os.path.join(os.path.dirname(__file__), "sga_test_ref.fa"), "w") as f:
        f.write(">chr1\n")
        for l in chr1:
            f.write(l + "\n")
        f.write(">chr2\n" + "".join(chr2) + "\n")


def frozen_tables():
    """Frozen outputs of the ORACLE (not of the reference: no JVM here) on seeded synthetic inputs: they pin today's oracle + engine against
    drifting together unnoticed.  Regenerate only on a deliberate, documented change of semantics: python tests/golden/make_fixtures.py frozen"""
    import sys
    root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "oracle"))
    import pyoracle
    from calitas_b200 import synth
    here = os.path.dirname(os.path.abspath(__file__))
    g = synth.config1_genome(scale=0.02, n_sites=60)
    contigs = [(n, bytes(b)) for n, b in g.contigs()]
    cases = {
        "frozen_search_defaults.tsv": pyoracle.search_reference(contigs, synth.BASELINE_GUIDE, guide_id="g", assembly="SYN10M", raw=True),
        "frozen_search_d6_g2_ngg_nag.tsv": pyoracle.search_reference(contigs, "CTTGCCCCACAGGGCAGTAAngg", aux_pams=["nag"], guide_id="g", assembly="SYN10M", raw=True, d=6, g=2, p=1),
        "frozen_a2r_best_w60.tsv": pyoracle.align_to_reference(contigs, synth.a2r_tasks(g, [synth.BASELINE_GUIDE], 40), window_size=60, raw=True, assembly="SYN10M"),
    }
    for name, text in cases.items():
        open(os.path.join(here, name), "w").write(text)
        print(name, text.count("\n"), "lines")


if __name__ == "__main__" and len(__import__("sys").argv) > 1 and __import__("sys").argv[1] == "frozen":
    frozen_tables()
