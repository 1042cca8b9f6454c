"""Replays every expectation of the reference's SequentialGuideAlignerTest.scala (26 tests) against the oracle
(CPU, here) and the CUDA engine (-m gpu).  Each test cites the Scala lines it mirrors."""
import pytest

SGA = "SequentialGuideAlignerTest.scala"


def rc(s):
    comp = str.maketrans("ACGTUMRWSYKVHDBNacgtumrwsykvhdbn", "TGCAAKYWSRMBDHVNtgcaakywsrmbdhvn")
    return s.translate(comp)[::-1]


def _check(a, strand, so, eo, gso, geo, cigar, pg, pt):
    assert a["strand"] == strand
    assert (a["startOffset"], a["endOffset"], a["guideStartOffset"], a["guideEndOffset"]) == (so, eo, gso, geo)
    assert a["cigar"] == cigar and a["paddedGuide"] == pg and a["paddedTarget"] == pt


def test_perfect_pamless_f_strand(backend):  # :51-65
    alns = backend.align("AACCAACC", "TTTTAACCAACCGGGG", max_guide_diffs=0, max_pam_diffs=0, max_gaps=0, max_total_diffs=0)
    assert len(alns) == 1
    _check(alns[0], "+", 4, 12, 4, 12, "8=", "AACCAACC", "AACCAACC")


def test_perfect_pamless_r_strand(backend):  # :67-81
    alns = backend.align("GGTTGGTT", "TTAACCAACCGGGG", max_guide_diffs=0, max_pam_diffs=0, max_gaps=0, max_total_diffs=0)
    assert len(alns) == 1
    _check(alns[0], "-", 2, 10, 2, 10, "8=", "GGTTGGTT", "GGTTGGTT")


def test_r_strand_with_mismatch(backend):  # :83-97
    alns = backend.align("GGTTGGTT", "AGCCAACC", max_guide_diffs=1, max_pam_diffs=0, max_gaps=0, max_total_diffs=1)
    assert len(alns) == 1
    _check(alns[0], "-", 0, 8, 0, 8, "6=1X1=", "GGTTGGTT", "GGTTGGCT")


def test_pam_3prime_f(backend):  # :99-112
    alns = backend.align("AACCAACCAACCnrg", "CCAACCAACCAACCGAGGGGGG", max_guide_diffs=0, max_pam_diffs=0, max_gaps=1, max_total_diffs=1)
    assert len(alns) == 1
    _check(alns[0], "+", 2, 17, 2, 14, "15=", "AACCAACCAACCnrg", "AACCAACCAACCGAG")


def test_pam_3prime_r(backend):  # :114-127
    alns = backend.align("AACCAACCAACCnrg", "CCCTGGGTTGGTTGGTTGGGGGG", max_guide_diffs=0, max_pam_diffs=0, max_gaps=1, max_total_diffs=1)
    assert len(alns) == 1
    _check(alns[0], "-", 2, 17, 5, 17, "15=", "AACCAACCAACCnrg", "AACCAACCAACCCAG")


def test_pam_5prime_f(backend):  # :129-142
    alns = backend.align("tttvAACCAACCAACC", "CCTTTGAACCAACCAACCGAGG", max_guide_diffs=0, max_pam_diffs=0, max_gaps=1, max_total_diffs=1)
    assert len(alns) == 1
    _check(alns[0], "+", 2, 18, 6, 18, "16=", "tttvAACCAACCAACC", "TTTGAACCAACCAACC")


def test_pam_5prime_r(backend):  # :144-157
    query = "tttvAACCAACCAACC"
    target = "CC" + rc(query.replace("tttv", "TTTG")) + "GAGG"
    alns = backend.align(query, target, max_guide_diffs=0, max_pam_diffs=0, max_gaps=1, max_total_diffs=1)
    assert len(alns) == 1
    _check(alns[0], "-", 2, 18, 2, 14, "16=", "tttvAACCAACCAACC", "TTTGAACCAACCAACC")


def test_pam_5prime_mismatch_f(backend):  # :159-172
    alns = backend.align("tttvAACCAACCAACC", "CCTTTGAACCAACCAAGCGAGG", max_guide_diffs=1, max_pam_diffs=0, max_gaps=0, max_total_diffs=1)
    assert len(alns) == 1
    _check(alns[0], "+", 2, 18, 6, 18, "14=1X1=", "tttvAACCAACCAACC", "TTTGAACCAACCAAGC")


def test_pam_5prime_mismatch_r(backend):  # :174-187
    target = "CC" + rc("TTTGAACCAACCAAGC") + "GAGG"
    alns = backend.align("tttvAACCAACCAACC", target, max_guide_diffs=1, max_pam_diffs=0, max_gaps=0, max_total_diffs=1)
    assert len(alns) == 1
    _check(alns[0], "-", 2, 18, 2, 14, "14=1X1=", "tttvAACCAACCAACC", "TTTGAACCAACCAAGC")


def test_target_offset(backend):  # :189-220
    guide1, guide2 = "gggTTTTT", "TTTTTggg"
    target1 = "AGAGAGAGAGGGTTTTTGGGAGAGAGAGAGAGAG"
    target2 = "AGAGAGAGACCCAAAAACCCAGAGAGAGAGAGAG"
    kw = dict(max_guide_diffs=0, max_pam_diffs=0, max_gaps=0, max_total_diffs=0, target_offset=1000)
    r1 = backend.align(guide1, target1, **kw)[0]
    assert (r1["startOffset"], r1["endOffset"], r1["guideStartOffset"], r1["guideEndOffset"]) == (1009, 1017, 1012, 1017)
    r2 = backend.align(guide2, target1, **kw)[0]
    assert (r2["startOffset"], r2["endOffset"]) == (1012, 1020)
    r3 = backend.align(guide1, target2, **kw)[0]
    assert (r3["startOffset"], r3["endOffset"]) == (1012, 1020)
    r4 = backend.align(guide2, target2, **kw)[0]
    assert (r4["startOffset"], r4["endOffset"]) == (1009, 1017)


def test_rc_symmetry(backend):  # :222-233
    query = "AATTCcgg"
    for target in ["AATTCCGG", "AGTTCCGG", "AAATTCCGG", "AATTCCGAG", "AATTCCTG"]:
        f = backend.align_best(query, target)
        r = backend.align_best(rc(query), rc(target))
        for k in ("score", "guideMismatches", "guideGapBases", "pamMismatches", "pamGapBases"):
            assert r[k] == f[k], (target, k)


def test_penalize_n_in_reference(backend):  # :235-240
    result = backend.align_best("AACCGGTTnrg", "nnnnnnnnnnn")
    assert result["score"] == 8 * -60 + 3 * -130


def test_max_guide_diffs_with_indels(backend):  # :242-248
    results = backend.align("yttnAGGAAACTTCTGGCAGGACC", "GTTAGTTCCAGATCTTGAGGAAGCTATCCCAGGACCCTGTCGCCACAGCCA",
                            max_guide_diffs=5, max_gaps=1, max_pam_diffs=1, max_total_diffs=7, max_overlap=10)
    assert len(results) == 1
    assert results[0]["startOffset"] == 13


def test_pick_best_pam(backend):  # :250-256
    result = backend.align_best("AACCGGTTACGTnrg", "AACCGGTTACGTTTG", aux_pams=["ntg"])
    assert result["guide"] == "AACCGGTTACGTntg"
    assert result["pamMmsPlusGaps"] == 0


def test_prefer_longer_pam(backend):  # :258-263
    result = backend.align_best("AACCGGTTACGTnnn", "AACCGGTTACGTAAAAAAA", aux_pams=["nnnn", "nn"])
    assert result["guide"] == "AACCGGTTACGTnnnn"


def test_prefer_longer_pam_with_gap(backend):  # :265-271
    result = backend.align_best("AACCGGTTACGTacc", "AACCGGTTACGTACCCC", aux_pams=["cccc"])
    assert result["guide"] == "AACCGGTTACGTcccc"
    assert result["cigar"] == "12=1D4="


def _sub(ref, chrom, start1, end1):
    return dict(ref)[chrom][start1 - 1:end1]


def test_ref_perfect_f(backend, sga_ref):  # :274-285
    query = _sub(sga_ref, "chr1", 50, 69)
    r = backend.align_to_ref_best(sga_ref, query, "chr1", 65)
    assert (r["chrom"], r["startOffset"], r["endOffset"], r["strand"]) == ("chr1", 49, 69, "+")
    assert r["paddedGuide"] == r["paddedTarget"]
    assert set(r["paddedAlignment"]) == {"|"}
    assert r["score"] >= 0


def test_ref_u_same_as_t(backend, sga_ref):  # :287-296
    t_query = _sub(sga_ref, "chr1", 50, 69)
    u_query = t_query.replace("T", "U")
    assert u_query != t_query
    t = backend.align_to_ref_best(sga_ref, t_query, "chr1", 65)
    u = backend.align_to_ref_best(sga_ref, u_query, "chr1", 65)
    assert u["score"] == t["score"] and u["paddedAlignment"] == t["paddedAlignment"]


def test_ref_perfect_r(backend, sga_ref):  # :298-308
    query = rc(_sub(sga_ref, "chr1", 50, 69))
    r = backend.align_to_ref_best(sga_ref, query, "chr1", 65)
    assert (r["chrom"], r["startOffset"], r["endOffset"], r["strand"]) == ("chr1", 49, 69, "-")
    assert set(r["paddedAlignment"]) == {"|"}
    assert r["score"] >= 0


def test_ref_mismatch_f(backend, sga_ref):  # :310-321
    query = "GAGAATTGtTTGAACCCAGGnGG"
    r = backend.align_to_ref_best(sga_ref, query.upper(), "chr1", 515)
    assert (r["chrom"], r["startOffset"], r["endOffset"], r["strand"]) == ("chr1", 500, 523, "+")
    assert r["paddedAlignment"] == "||||||||.||||||||||||||"
    assert r["mismatches"] == 1


def test_ref_ambiguity_codes_in_pam(backend, sga_ref):  # :323-337
    r = backend.align_to_ref_best(sga_ref, "TCAGTGCCTGCGCCGCGCTCGCTCCCnrycwshdm", "chr1", 1820)
    assert (r["chrom"], r["startOffset"], r["endOffset"], r["guideStartOffset"], r["guideEndOffset"], r["strand"]) == ("chr1", 1800, 1835, 1800, 1826, "+")
    assert r["paddedAlignment"] == "||||||||||||||||||||||||||||||.||||"
    assert r["mismatches"] == 1


def test_ref_two_bulges_r(backend, sga_ref):  # :339-349
    query = "AGGCTGG-GGCGGTCGCtCGCNGG"
    r = backend.align_to_ref_best(sga_ref, "".join(c for c in query if c.isalpha()).upper(), "chr1", 1510)
    assert (r["chrom"], r["startOffset"], r["endOffset"], r["strand"]) == ("chr1", 1500, 1523, "-")
    assert r["paddedAlignment"] == "|||||||~|||||||||~||||||"


def test_ref_two_guide_mismatches_beat_one_pam_mismatch(backend, sga_ref):  # :351-359
    r = backend.align_to_ref_best(sga_ref, "GATACGTCTCGTACTGTnrg", "chr2", 22)
    assert (r["chrom"], r["startOffset"], r["endOffset"], r["gapBases"], r["mismatches"]) == ("chr2", 0, 20, 0, 2)


def test_prefer_mismatch_to_genome_bulge(backend):  # :361-368
    query = "GATACGTCTCGTACTGTnrg"
    target = query.replace("GATA", "GATT").replace("nrg", "AAG") + "TTTTT" + query.replace("TCTC", "TCTCC").replace("nrg", "AAG")
    r = backend.align_best(query, target)
    assert (r["startOffset"], r["mismatches"], r["gapBases"]) == (0, 1, 0)


def test_prefer_genome_bulge_to_guide_bulge(backend):  # :370-377
    query = "GATACGTCTCGTACTGTnrg"
    target = query.replace("TCTC", "TCTCC").replace("nrg", "AAG") + "NNNNN" + query.replace("TCTC", "TCT").replace("nrg", "AAG")
    r = backend.align_best(query, target)
    assert (r["startOffset"], r["mismatches"], r["gapBases"]) == (0, 0, 1)


def test_max_total_diffs_is_separate(backend):  # :379-389
    query = "GATACGTCTCGTACTGTnrg"
    target1 = "GAaACGTtTCGTACTGTaac".upper()
    r1 = backend.align(query, target1, max_guide_diffs=2, max_gaps=0, max_pam_diffs=1, max_total_diffs=3)
    assert len(r1) == 1
    r2 = backend.align(query, target1, max_guide_diffs=2, max_gaps=0, max_pam_diffs=1, max_total_diffs=2)
    assert len(r2) == 0
