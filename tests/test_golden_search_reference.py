"""SearchReferenceTest.scala replayed (end-to-end SearchReference, variant windows, alleleCombos).  Each test cites the Scala lines."""
import os, sys
import pytest
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import pyoracle

GUIDE = "ACGTACATGCTCGATACGACGnngrrn"
PERFECT = "ACGTACATGCTCGATACGACGccgaat".upper()
MISMATCHED = "ACGcACAcGCcCGAcACGACGccgaat".upper()
FASTA = [("chr1", "N" * 5000 + "AATAT" * 1000 + "N" * 5000),                                                   # SearchReferenceTest.scala:17-33
         ("chr2", "N" * 3000 + PERFECT + "GT" * 500 + MISMATCHED + "CA" * 500 + "N" * 3000)]


def test_end_to_end(backend):  # :51-62
    hits = backend.search_reference(FASTA, GUIDE, guide_id="a", threads=1)
    assert len(hits) == 2
    assert all(h["chromosome"] == "chr2" for h in hits)
    assert hits[0]["coordinate_start"] == 3000 and hits[0]["total_mm_plus_gaps"] == 0
    assert hits[1]["coordinate_start"] == 4000 + len(PERFECT) and hits[1]["total_mm_plus_gaps"] == 4


def test_pamless_guide(backend):  # :64-69
    hits = backend.search_reference(FASTA, "".join(c for c in GUIDE if c.isupper()), guide_id="a", threads=1)
    assert len(hits) == 2


def test_adjacent_short_contigs(backend):  # :71-92
    ref = [("ref", "GTGCGTGACTTGAAGTCTCAGTATACCTTGCCACACGTTGCAGGTTGCCC"), ("alt", "GTGCGTGACTTGAAGTCTCAGTATgaaaTTGCCACACGTTGCAGGTTGCCC")]
    hits = backend.search_reference(ref, "GTGACTTGAAGTCTCAGTATA", guide_id="a", threads=1)
    assert len(hits) == 2
    assert (hits[0]["chromosome"], hits[0]["coordinate_start"], hits[0]["padded_alignment"]) == ("ref", 4, "|||||||||||||||||||||")
    assert (hits[1]["chromosome"], hits[1]["coordinate_start"], hits[1]["padded_alignment"]) == ("alt", 4, "||||||||||||||||||||.")


VCF_HEADER = "##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n"


def vcf(*records):
    return VCF_HEADER + "".join("\t".join([c, str(p), i, r, a, ".", "PASS", info]) + "\n" for (c, p, i, r, a, info) in records)


def test_flanks_for_ref_and_variant_windows(backend):  # :94-147
    query = "GCGTCACGGTCGAGCGATTGnrg"
    chr1 = ("ACACACACACACACACACACACACACACACACACACACAgcgtcacggtcgagcgattggggAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAA" +
            "ACACACACACACACACACACACACACACACACACACACAccccaatcgctcgaccgtgacgcAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAA" +
            "ACACACACACACACACACACACACACACACACACACACAcacggtcgagcgattggggAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAA" +
            "ACACACACACACACACACACACACACACACACACACACAaatcgctcgaccgtgacgcAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAA").upper()
    text = vcf(("chr1", 239, "insGAGGCGT", "A", "AGAGGCGT", "."), ("chr1", 339, "insTCGCCCC", "A", "ATCGCCCC", "."))
    hits = backend.search_reference([("chr1", chr1)], query, guide_id="test", vcf_text=text, g=0, d=0)
    assert len(hits) == 4
    h1, h2, h3, h4 = hits
    exp = [(39, "CACACACA", "AAAAAAAA", "CACACACACA", "GGGAAAAAAA"),
           (142, "TTTTTTTT", "TGTGTGTG", "TTTTTTTTTT", "GGGTGTGTGT"),
           (238, "ACACAGAG", "AAAAAAAA", "ACACACAGAG", "GGGAAAAAAA"),
           (338, "TTTTTTTT", "CGATGTGT", "TTTTTTTTTT", "GGGCGATGTG")]
    for h, (start, e5, e3, t5, t3) in zip(hits, exp):
        assert h["coordinate_start"] == start
        assert h["padded_extra_8_bases_5_prime"] == e5 and h["padded_extra_8_bases_3_prime"] == e3
        assert h["ten_bases_5_prime"] == t5 and h["ten_bases_3_prime"] == t3


# ---- pure functions of the variant path (host-side in both implementations; oracle checked here, product in test_host_variants.py) ----
def test_allele_combos_counts():  # :150-181
    assert pyoracle.allele_combos([2]) == [[0], [1]]
    assert pyoracle.allele_combos([3]) == [[0], [1], [2]]
    assert pyoracle.allele_combos([2, 2]) == [[0, 0], [0, 1], [1, 0], [1, 1]]
    assert pyoracle.allele_combos([3, 2]) == [[0, 0], [0, 1], [1, 0], [1, 1], [2, 0], [2, 1]]
    assert pyoracle.allele_combos([3, 2, 3]) == [[a, b, c] for a in range(3) for b in range(2) for c in range(3)]


REF50 = "CTAGACTGACTGACTAGCACTAGCCGCTTTATATATGCTATGGGACACCG"


def test_variant_window_snp():  # :183-196
    w = pyoracle.build_variant_window("chr1", REF50, vcf(("chr1", 20, "rs123", "C", "G", ".")), 15, [(0, True), (15, True), (20, True), (31, True)])
    assert w["bases"] == "ACTGACTGACTAGCAgTAGCCGCTTTATATA".upper() and w["cigar"] == "31M"
    assert w["offsets"] == [4, 19, 24, 35]


def test_variant_window_insertion():  # :198-215
    w = pyoracle.build_variant_window("chr1", REF50, vcf(("chr1", 20, "rs123", "C", "CGT", ".")), 15,
                                      [(0, True), (14, True), (15, True), (16, True), (17, True), (15, False), (16, False), (17, False)])
    assert w["bases"] == "ACTGACTGACTAGCAcgtTAGCCGCTTTATATA".upper() and w["cigar"] == "16M2I15M"
    assert w["offsets"] == [4, 18, 19, 19, 19, 19, 20, 20]


def test_variant_window_deletion():  # :217-230
    w = pyoracle.build_variant_window("chr1", REF50, vcf(("chr1", 20, "rs123", "CTA", "C", ".")), 15, [(0, True), (15, True), (16, True)])
    assert w["bases"] == "ACTGACTGACTAGCAcGCCGCTTTATATATG".upper() and w["cigar"] == "16M2D15M"
    assert w["offsets"] == [4, 19, 22]


def test_variant_window_multiple():  # :232-247
    ref = "CTAGACTGACTGACTAGCACTAGCCGCTTTATATATGCTAGGCGCTACTGAATGCTATAGCTCTGAGACTGGGACACCG"
    w = pyoracle.build_variant_window("chr1", ref, vcf(("chr1", 10, "snp", "C", "T", "."), ("chr1", 20, "ins", "C", "CG", "."), ("chr1", 30, "del", "TAT", "T", ".")), 15)
    assert w["bases"] == "CTAGACTGAtTGACTAGCAcgTAGCCGCTTtATATGCTAGGCGCTA".upper() and w["cigar"] == "20M1I10M2D15M"


def test_allele_combos_variants():  # :249-295
    assert pyoracle.variant_sets(vcf(("chr1", 20, "snp", "A", "C", ".")), 10) == [(("snp", "1"),)]
    assert sorted(pyoracle.variant_sets(vcf(("chr1", 20, "snp", "A", "C,G,T", ".")), 10)) == [(("snp", "1"),), (("snp", "2"),), (("snp", "3"),)]
    three = vcf(("chr1", 20, "a", "A", "C", "."), ("chr1", 25, "b", "C", "T", "."), ("chr1", 30, "c", "G", "A", "."))
    got = sorted(pyoracle.variant_sets(three, 10))
    exp = sorted([(("a", "1"),), (("b", "1"),), (("c", "1"),), (("a", "1"), ("b", "1")), (("a", "1"), ("c", "1")), (("b", "1"), ("c", "1")),
                  (("a", "1"), ("b", "1"), ("c", "1"))])
    assert got == exp
    assert len(pyoracle.variant_sets(three, 2)) == 1
    assert len(pyoracle.variant_sets(three, 3)) == 7
