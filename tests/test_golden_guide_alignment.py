"""GuideAlignmentTest.scala replayed against the oracle's GuideAlignment restatement (GuideAlignment.scala:10-50,99-163).
The product derives the same counters arithmetically from cigar ops; tests/test_host_render.py checks that path."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import pyoracle

CASES = [
    # (paddedQuery, paddedAlign, paddedTarget, start, end, strand, expected: gmm ggap gmmgap pmm pgap pmmgap mm gaps edits gstart gend)  -- GuideAlignmentTest.scala
    ("GCTGACTGCATGACTATAnrg", "|||||||||||||||||||||", "GCTGACTGCATGACTATAnrg", 1, 21, "+", (0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 18)),            # :11-28
    ("GCTGACT-GCATGACTATAnrg", "||.||||~|||.||~|||||||", "GCAGACTCGCACGA-TATAnrg", 1, 21, "+", (2, 2, 4, 0, 0, 0, 2, 2, 4, 1, 18)),         # :30-47
    ("GCTGACTGCATGACTATAnngrrn", "|||||||||||||||||||~||.|", "GCTGACTGCATGACTATAC-GATT", 1, 23, "+", (0, 0, 0, 1, 1, 2, 1, 1, 2, 1, 18)),   # :49-66
    ("GCTGAC---TGCATGACTATAnrg", "||||||~~~||||~~|||||||||", "GCTGACGGGTGCA--ACTATACGG", 1, 22, "-", (0, 5, 5, 0, 0, 0, 0, 5, 5, 4, 22)),   # :68-85
    ("---GCTGACTGCATGACTATAnrg--", "~~~|||||||||||||||||||||~~", "TGTGCTGACTGCATGACTATACGGCC", 1, 26, "+", (0, 3, 3, 0, 2, 2, 0, 5, 5, 4, 21)),  # :87-104
    ("GCTGACTGCATGACTATA--nrg", "||||||||||||||||||~~|||", "GCTGACTGCATGACTATATTCGG", 1, 23, "+", (0, 2, 2, 0, 0, 0, 0, 2, 2, 1, 18)),     # :106-123
]


def test_guide_alignment_counts():
    for pq, pa, pt, s, e, strand, exp in CASES:
        a = pyoracle.guide_alignment(pq, pa, pt, s, e, strand)
        got = (a["guideMismatches"], a["guideGapBases"], a["guideMmsPlusGaps"], a["pamMismatches"], a["pamGapBases"], a["pamMmsPlusGaps"],
               a["mismatches"], a["gapBases"], a["edits"], a["guideStartOffset"], a["guideEndOffset"])
        assert got == exp, (pq, got, exp)
