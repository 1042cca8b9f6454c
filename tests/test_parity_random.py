"""Engine-vs-oracle parity on seeded synthetic inputs (SURVEY.md 8d shapes, reduced sizes): every column of every row must be identical.
Runs against the host simulation here (-m "not gpu") and against the CUDA library on the B200 (-m gpu)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import pyoracle
from calitas_b200 import synth

ENGINES = ["hostsim", pytest.param("gpu", marks=pytest.mark.gpu)]


@pytest.fixture(scope="module", params=ENGINES)
def eng(request):
    import backends
    return backends.get(request.param)


def rc(s):
    return synth.revcomp_bytes(s.encode()).decode()


IUPAC = "ACGTRYSWKMBDHVN"


def random_case(rng):
    lp = int(rng.integers(6, 25))
    alphabet = IUPAC if rng.random() < 0.25 else "ACGT"
    proto = "".join(rng.choice(list(alphabet), size=lp))
    kind = rng.integers(0, 4)
    pams = ["".join(rng.choice(list("acgtnryv"), size=int(rng.integers(1, 6)))) for _ in range(int(rng.integers(1, 4)))]
    if kind == 0:
        guide, aux = proto, []
    elif kind == 1:
        guide, aux = proto + pams[0], pams[1:]
    elif kind == 2:
        guide, aux = pams[0] + proto, pams[1:]
    else:
        guide, aux = proto + pams[0], []
    # target: random background with 1-3 mutated copies of the site, sometimes N runs / lower case / IUPAC
    tlen = int(rng.integers(lp, 160))
    t = list(rng.choice(list("ACGT"), size=tlen))
    concrete = "".join(c if c in "ACGT" else "ACGT"[int(rng.integers(0, 4))] for c in proto)
    for _ in range(int(rng.integers(0, 4))):
        site = synth.mutate_protospacer(rng, concrete.encode(), int(rng.integers(0, 4))).decode()
        site = site + "".join(rng.choice(list("ACGT"), size=int(rng.integers(0, 3)))) + "".join(rng.choice(list("ACGT"), size=3))
        if kind == 2:
            site = site[::-1]
        if rng.random() < 0.5:
            site = rc(site)
        if len(site) <= tlen:
            p = int(rng.integers(0, tlen - len(site) + 1))
            t[p:p + len(site)] = list(site)
    t = t[:tlen]
    if rng.random() < 0.3:
        p = int(rng.integers(0, tlen)); n = int(rng.integers(1, 8))
        for q in range(p, min(tlen, p + n)):
            t[q] = "N"
    if rng.random() < 0.2:
        t[int(rng.integers(0, tlen))] = str(rng.choice(list("RYMKnacgt")))
    target = "".join(t)
    d = int(rng.integers(0, 5)); p_ = int(rng.integers(0, 3)); g = int(rng.integers(0, 4))
    D = None if rng.random() < 0.5 else int(rng.integers(0, d + p_ + g + 1))
    return guide, aux, target, dict(max_guide_diffs=d, max_pam_diffs=p_, max_gaps=g, max_total_diffs=(d + g + p_ if D is None else D),
                                    max_overlap=int(rng.choice([0, 5, 10, 100])), target_offset=int(rng.integers(0, 1000)))


@pytest.mark.parametrize("paired", [False, True])
def test_align_random_cases(eng, monkeypatch, paired):
    if paired:
        monkeypatch.setenv("CALITAS_ALIGN_PAIR", "1")            # the opt-in two-candidates-per-thread aligner (align_pair) on the same cases
    rng = np.random.default_rng(7)
    n = 400 if eng.name == "hostsim" else 150
    for k in range(n):
        guide, aux, target, kw = random_case(rng)
        exp = pyoracle.align(guide, target, aux_pams=aux, **kw)
        got = eng.align(guide, target, aux_pams=aux, **kw)
        assert got == exp, (k, guide, aux, target, kw)


def test_align_best_random_cases(eng):
    rng = np.random.default_rng(8)
    n = 150 if eng.name == "hostsim" else 60
    for k in range(n):
        guide, aux, target, kw = random_case(rng)
        if len(target) < 4:
            continue
        try:
            exp = pyoracle.align_best(guide, target, aux_pams=aux, max_gaps=kw["max_gaps"])
        except pyoracle.OracleError:
            with pytest.raises(Exception):
                eng.align_best(guide, target, aux_pams=aux, max_gaps=kw["max_gaps"])
            continue
        got = eng.align_best(guide, target, aux_pams=aux, max_gaps=kw["max_gaps"])
        assert got == exp, (k, guide, aux, target)


def test_non_default_costs(eng):
    rng = np.random.default_rng(9)
    # the last two sets break  target_gap + query_gap <= mismatch  (band_align_h's condition): they must take the three-matrix kernels
    for costs in [(-100, -130, -110, -200), (-120, -121, -122, -260), (90, 95, 100, 300), (-400, -100, -100, -260), (-250, -120, -120, -260)]:      # (-250,-120,-120): k_edits = 2 d, so d <= 3 stays on the banded three-matrix kernel
        for k in range(40):
            guide, aux, target, kw = random_case(rng)
            exp = pyoracle.align(guide, target, aux_pams=aux, costs=costs, **kw)
            got = eng.align(guide, target, aux_pams=aux, costs=costs, **kw)
            assert got == exp, (costs, k, guide, aux, target, kw)


@pytest.fixture(scope="module")
def small_genome():
    g = synth.config1_genome(scale=0.03, n_sites=60)
    return g, [(n, bytes(b)) for n, b in g.contigs()]


def _lines(text):
    return [l for l in text.split("\n") if l]


def test_search_reference_defaults(eng, small_genome):
    g, contigs = small_genome
    exp = _lines(pyoracle.search_reference(contigs, synth.BASELINE_GUIDE, guide_id="g1", raw=True, assembly="SYN10M"))
    got = _lines(eng.search_reference(contigs, synth.BASELINE_GUIDE, guide_id="g1", raw=True, assembly="SYN10M"))
    assert len(exp) > 40
    assert got == exp


@pytest.mark.parametrize("guide,aux,kw", [
    ("CTTGCCCCACAGGGCAGTAAngg", ["nag"], dict(d=6, g=2, p=1)),
    ("CTTGCCCCACAGGGCAGTAA", [], dict(d=6, g=2, p=1)),
    ("tttvCTTGCCCCACAGGGCAGTAA", [], dict(d=4, g=1, p=1, O=0)),
    ("CTTGCCCCACAGGGCAGTAAnrg", [], dict(d=3, g=3, p=2, D=4, window_size=300, O=100)),
    ("CTTGCCCCACAGGGCAGTAAnrg", [], dict(window_size=3000)),            # 32 windows per scan tile
    ("CTTGCCCCACAGGGCAGTAAnrg", [], dict(window_size=20000, d=4)),      # 4 windows per scan tile
    ("CTTGCCCCACAGGGCAGTAAnrg", [], dict(window_size=120000)),          # one window per tile, close to the engine's window limit
    ("CTTGCCCCACAGGGCAGTAAnrg", [], dict(window_size=40)),              # step 10: every base is scanned four times
])
def test_search_reference_variants_of_the_call(eng, small_genome, guide, aux, kw):
    g, contigs = small_genome
    exp = _lines(pyoracle.search_reference(contigs, guide, aux_pams=aux, raw=True, **kw))
    got = _lines(eng.search_reference(contigs, guide, aux_pams=aux, raw=True, **kw))
    assert got == exp


@pytest.mark.parametrize("guide,aux,kw", [
    ("CTTGCCCCACAGGGCAGTAAngg", ["nag"], dict(d=6, g=2, p=1)),
    ("tttvCTTGCCCCACAGGGCAGTAA", [], dict(d=5, g=1, p=1)),
    ("CTTGCCCCACAGGGCAGTAAnrg", [], dict(d=4)),
])
def test_search_reference_with_the_paired_16_bit_aligner(eng, small_genome, monkeypatch, guide, aux, kw):
    """CALITAS_ALIGN_PAIR=1: align_pair (two candidates per thread, 16-bit packed scores) instead of align_fast.  Not the default (slower on the B200),
    kept exact: same rows as the oracle, odd and even candidate counts, both strands, 5' and 3' PAMs, band widths 4 / 5 / 6."""
    monkeypatch.setenv("CALITAS_ALIGN_PAIR", "1")
    g, contigs = small_genome
    exp = _lines(pyoracle.search_reference(contigs, guide, aux_pams=aux, raw=True, **kw))
    got = _lines(eng.search_reference(contigs, guide, aux_pams=aux, raw=True, **kw))
    assert len(exp) > 10
    assert got == exp


@pytest.mark.parametrize("kw", [dict(), dict(O=0), dict(O=1), dict(d=6, g=2), dict(window_size=200, O=25)])
def test_search_reference_repeats_on_both_strands(eng, kw):
    """Dense, overlapping hits on both strands: tandem copies of the target (guide + PAM), some edited, and of its reverse complement, packed
    closer than an alignment is long.  Stresses the per-window overlap filter, the two-strand removeOverlaps sweep (chains interleave by start,
    equal starts on both strands occur), sort ties and traceback choices in repeats."""
    rng = np.random.default_rng(11)
    site = "CTTGCCCCACAGGGCAGTAA"
    comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
    def edited(s):
        s = list(s)
        for _ in range(int(rng.integers(0, 3))):
            i = int(rng.integers(len(s)))
            r = rng.random()
            if r < 0.5:
                s[i] = "ACGT"[int(rng.integers(4))]
            elif r < 0.75:
                del s[i]
            else:
                s.insert(i, "ACGT"[int(rng.integers(4))])
        return "".join(s)
    contigs = []
    for c in range(3):
        parts = ["".join("ACGT"[i] for i in rng.integers(0, 4, 40))]
        for _ in range(120):
            unit = edited(site) + ["AGG", "TGG", "CAG", "GGG", "TGA"][int(rng.integers(5))]
            if rng.random() < 0.5:
                unit = "".join(comp[b] for b in reversed(unit))
            parts.append(unit)
            parts.append("".join("ACGT"[i] for i in rng.integers(0, 4, int(rng.integers(0, 9)))))      # 0-8 bases between copies
            if rng.random() < 0.05:
                parts.append("N" * int(rng.integers(1, 40)))
        contigs.append(("rep%d" % c, "".join(parts).encode()))
    exp = _lines(pyoracle.search_reference(contigs, synth.BASELINE_GUIDE, raw=True, **kw))
    got = _lines(eng.search_reference(contigs, synth.BASELINE_GUIDE, raw=True, **kw))
    assert len(exp) > (3 if kw.get("O") == 0 else 150)          # -O 0: every later hit of a group "overlaps" (>= 0), the sweep keeps a handful
    assert got == exp


def test_search_single_chrom(eng, small_genome):
    g, contigs = small_genome
    exp = _lines(pyoracle.search_reference(contigs, synth.BASELINE_GUIDE, chrom="chr2", raw=True))
    got = _lines(eng.search_reference(contigs, synth.BASELINE_GUIDE, chrom="chr2", raw=True))
    assert got == exp and len(exp) > 5


def test_search_reference_with_vcf(eng, small_genome):
    g, contigs = small_genome
    arrays = [np.frombuffer(b, dtype=np.uint8) for _, b in contigs]
    vcf = synth.synthetic_vcf(g, arrays, 600)
    exp = _lines(pyoracle.search_reference(contigs, synth.BASELINE_GUIDE, vcf_text=vcf, raw=True))
    got = _lines(eng.search_reference(contigs, synth.BASELINE_GUIDE, vcf_text=vcf, raw=True))
    assert got == exp
    assert any("+variants" in l for l in exp)


def test_align_to_reference(eng, small_genome):
    g, contigs = small_genome
    guides = [synth.BASELINE_GUIDE] + synth.random_guides(3)
    tasks = synth.a2r_tasks(g, guides, 120)
    for kw in (dict(window_size=60), dict(window_size=60, d=5, p=1, O=10), dict()):
        exp = _lines(pyoracle.align_to_reference(contigs, tasks, raw=True, **kw))
        got = _lines(eng.align_to_reference(contigs, tasks, raw=True, **kw))
        assert got == exp, kw


@pytest.mark.parametrize("n_shards", [1, 3])
def test_search_reference_batch_over_shards(eng, small_genome, n_shards):
    """A batch of guides over 1 or 3 engines (contig-range shards driven from host threads) = the oracle's per-guide tables, concatenated."""
    g, contigs = small_genome
    guides = [synth.BASELINE_GUIDE, ("CTTGCCCCACAGGGCAGTAAngg", ["nag"]), "tttvCTTGCCCCACAGGGCAGTAA", "GGGGCCACTAGGGACAGGAT"]
    ids = ["a", "b", "c", "d"]
    arrays = [np.frombuffer(b, dtype=np.uint8) for _, b in contigs]
    for vcf in (None, synth.synthetic_vcf(g, arrays, 300)):
        exp = []
        for gd, gid in zip(guides, ids):
            seq, aux = (gd, []) if isinstance(gd, str) else gd
            lines = _lines(pyoracle.search_reference(contigs, seq, guide_id=gid, aux_pams=aux, vcf_text=vcf, raw=True))
            exp += lines if not exp else lines[1:]
        got = _lines(eng.t.search_reference_batch(contigs, guides, ids, n_shards=n_shards, vcf_text=vcf))
        assert got == exp, (n_shards, vcf is not None)


def _tandem_repeat_contigs(seed, n_units):
    rng = np.random.default_rng(seed)
    comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
    parts = []
    for _ in range(n_units):
        s = list("CTTGCCCCACAGGGCAGTAA")
        for _ in range(int(rng.integers(0, 3))):
            s[int(rng.integers(len(s)))] = "ACGT"[int(rng.integers(4))]
        unit = "".join(s) + ["AGG", "TGG", "CAG", "GGG"][int(rng.integers(4))]
        if rng.random() < 0.5:
            unit = "".join(comp[b] for b in reversed(unit))
        parts.append(unit)
        parts.append("".join("ACGT"[i] for i in rng.integers(0, 4, int(rng.integers(0, 8)))))
    return [("rep0", "".join(parts).encode())]


@pytest.mark.parametrize("O", [10, 25, 0])
def test_shard_cuts_inside_dense_repeats(eng, O):
    """Consecutive windows overlap by 30 bases, so around a shard cut both neighbours report hits whose starts interleave whenever removeOverlaps
    does not collapse them (-O 25 exceeds every alignment's length: nothing is removed).  The host must merge the shard lists by the sort key
    at the cuts, not just concatenate them."""
    contigs = _tandem_repeat_contigs(503, 160)
    guides = [synth.BASELINE_GUIDE, ("CTTGCCCCACAGGGCAGTAAngg", ["nag"])]
    exp = []
    for gd, gid in zip(guides, "ab"):
        seq, aux = (gd, []) if isinstance(gd, str) else gd
        lines = _lines(pyoracle.search_reference(contigs, seq, guide_id=gid, aux_pams=aux, raw=True, O=O))
        exp += lines if not exp else lines[1:]
    assert len(exp) > {0: 3, 10: 300, 25: 1000}[O]        # -O 0: the sweep has unbounded reach (every later hit of a group "overlaps"), the tool de-duplicates on the host then
    for n_shards in (2, 3, 5):
        got = _lines(eng.t.search_reference_batch(contigs, guides, ["a", "b"], n_shards=n_shards, O=O))
        assert got == exp, n_shards


def test_align_to_reference_batches_of_10000_are_sorted_separately(eng, small_genome):
    """AlignToReference.scala:110,141: output is ReferenceHit.sort-ed per batch of 10 000 input rows, not globally."""
    g, contigs = small_genome
    guides = [synth.BASELINE_GUIDE] + synth.random_guides(2)
    tasks = synth.a2r_tasks(g, guides, 21000, near_fraction=0.9)
    kw = dict(window_size=60, d=5, p=1, O=10)
    exp = _lines(pyoracle.align_to_reference(contigs, tasks, raw=True, threads=8, **kw))
    got = _lines(eng.align_to_reference(contigs, tasks, raw=True, **kw))
    assert len(exp) > 3000 and got == exp
    starts = [int(l.split("\t")[4]) for l in exp[1:]]
    assert starts != sorted(starts)                       # three separately sorted batches


def test_search_reference_many_small_contigs(eng):
    """Hundreds of contigs from 1 base to a few windows long, with N runs, lower case and IUPAC codes: tile and window bookkeeping at contig
    edges (Range(0, len-1, step) yields nothing for a 1-base contig; windows shorter than the guide are dropped; upper-case N is trimmed, n is not)."""
    rng = np.random.default_rng(2026)
    site = "CTTGCCCCACAGGGCAGTAA"
    contigs = []
    for i in range(260):
        n = int(rng.choice([1, 2, 5, 19, 22, 23, 24, 40, 969, 970, 971, 1000, 1001, 1940, 2500, int(rng.integers(1, 3000))]))
        b = list(rng.choice(list("ACGT"), size=n))
        if n > 60 and rng.random() < 0.7:
            p = int(rng.integers(0, n - 30))
            s = synth.mutate_protospacer(rng, site.encode(), int(rng.integers(0, 4))).decode() + "TGG"
            if rng.random() < 0.5:
                s = rc(s)
            b[p:p + len(s)] = list(s)[:max(0, n - p)]
        for _ in range(int(rng.integers(0, 3))):
            if n > 4:
                p = int(rng.integers(0, n)); k = int(rng.integers(1, 40)); ch = str(rng.choice(["N", "N", "n", "R", "a"]))
                for q in range(p, min(n, p + k)):
                    b[q] = ch if ch != "a" else b[q].lower()
        if rng.random() < 0.1:
            b[:min(n, 15)] = ["N"] * min(n, 15)
        contigs.append(("ctg%d" % i, "".join(b[:n]).encode()))
    for guide, kw in ((synth.BASELINE_GUIDE, {}), ("CTTGCCCCRCAGGGCAGTAAngg", dict(d=4, O=0)), ("ccnCTTGCCCCACAGGGCAGTAA", dict(window_size=200, g=1))):
        exp = _lines(pyoracle.search_reference(contigs, guide, raw=True, **kw))
        got = _lines(eng.search_reference(contigs, guide, raw=True, **kw))
        assert got == exp, guide
        assert len(exp) > 20


def test_search_reference_cheap_guide_gap_costs(eng):
    """Cost sets where one gap kind is far cheaper than the rest: an accepted alignment can then hold more than lp gap bases, so the DP rectangle
    (GuideSpec.span) and the banded-kernel decision must follow the uncapped cost budget, not k_edits capped at lp.  With (-30,-30,-3,-60), d=3 the
    fgbio-optimal alignment of guide[:10] + 15 A + mutated guide[:10] + guide[10:] holds 15 deletions, is filtered by diffs <= d, and nothing may be
    reported in its place (the reference takes the optimal path, then filters)."""
    rng = np.random.default_rng(31)
    guide = "CTTGCCCCACAGGGCAGTAAngg"
    proto = guide[:20]
    second = list(proto[:10])
    for i in (1, 4, 7):
        second[i] = "ACGT"[("ACGT".index(second[i]) + 1) % 4]
    site = proto[:10] + "A" * 15 + "".join(second) + proto[10:] + "AGG"
    b = list(rng.choice(list("ACGT"), size=6000))
    b[3000:3000 + len(site)] = list(site)
    for k in range(12):                                   # plus ordinary sites, some with gaps
        p = 200 + 450 * k
        s = synth.mutate_protospacer(rng, proto.encode(), int(rng.integers(0, 4))).decode() + "TGG"
        if k % 2:
            s = rc(s)
        if not (3000 - 60 < p < 3000 + 120):
            b[p:p + len(s)] = list(s)
    contigs = [("c", "".join(b).encode())]
    for costs, d in (((-30, -30, -3, -60), 3), ((-30, -3, -30, -60), 3), ((-100, -100, -10, -200), 2), ((-120, -122, -121, -260), 5)):
        exp = _lines(pyoracle.search_reference(contigs, guide, raw=True, costs=costs, d=d, window_size=1000))
        got = _lines(eng.search_reference(contigs, guide, raw=True, costs=costs, d=d, window_size=1000))
        assert got == exp, (costs, d)
        target = "".join(b[2950:3150])
        kw = dict(max_guide_diffs=d, max_pam_diffs=1, max_gaps=3, max_total_diffs=d + 4, max_overlap=10)
        assert eng.align(guide, target, costs=costs, **kw) == pyoracle.align(guide, target, costs=costs, **kw), (costs, d)


@pytest.mark.parametrize("period,O", [(7, 5), (10, 10), (11, 15), (13, 10)])
def test_shard_cuts_inside_short_period_repeats(eng, period, O):
    """removeOverlaps chains across a shard cut: a perfect-ish tandem repeat with a period SHORTER than an alignment, several kb long (longer than the
    two halo windows), so that hits overlap their neighbours by >= -O and the greedy chain (A dropped for B, B for C, ...) runs through the cut.
    Units are mutated now and then so that scores vary along the chain."""
    rng = np.random.default_rng(1000 + period)
    unit = "CTTGCCCCACAGGGCAGTAA"[:period]
    n_units = 6500 // period
    parts = []
    for _ in range(n_units):
        u = list(unit)
        if rng.random() < 0.08:
            u[int(rng.integers(period))] = "ACGT"[int(rng.integers(4))]
        parts.append("".join(u))
    head = "".join("ACGT"[i] for i in rng.integers(0, 4, 700))
    contigs = [("rep0", (head + "".join(parts) + head[::-1]).encode())]
    proto = (unit * 4)[:20]                                # the repeat itself is the best guide for dense, chained hits
    guides = [proto + "nrg", proto]
    exp = []
    for gd, gid in zip(guides, "ab"):
        lines = _lines(pyoracle.search_reference(contigs, gd, guide_id=gid, raw=True, O=O, p=3))
        exp += lines if not exp else lines[1:]
    assert len(exp) > 50
    for n_shards in (2, 3, 5):
        got = _lines(eng.t.search_reference_batch(contigs, guides, ["a", "b"], n_shards=n_shards, O=O, p=3))
        assert got == exp, (period, O, n_shards)


def test_align_best_mode_negative_overlap_cannot_happen_but_wide_costs_clamp(eng):
    """Explicit mode with wide thresholds takes the warp-per-group canonicaliser on the GPU (k_canon_warp); max_overlap < 0 must behave like
    GuideAlignment.overlap's clamp at 0 (nothing overlaps by less than 0, so every alignment after the first is dropped ... except disjoint ones,
    whose clamped overlap 0 > -1 drops them too)."""
    rng = np.random.default_rng(77)
    guide = "CTTGCCCCACAGGGCAGTAAngg"
    for k in range(6):
        t = list(rng.choice(list("ACGT"), size=150))
        for p in (10, 80):
            s = synth.mutate_protospacer(rng, guide[:20].encode(), int(rng.integers(0, 3))).decode() + "TGG"
            t[p:p + len(s)] = list(s)
        target = "".join(t)
        for O in (-1, 0, 3):
            kw = dict(max_guide_diffs=8, max_pam_diffs=1, max_gaps=3, max_total_diffs=12, max_overlap=O)     # d = 8 -> k_edits > 6: the wide kernels
            assert eng.align(guide, target, **kw) == pyoracle.align(guide, target, **kw), (k, O)


@pytest.mark.parametrize("kw", [dict(), dict(d=6, g=2), dict(O=25), dict(O=1, max_variants=3)])
def test_search_reference_with_dense_vcf(eng, small_genome, kw):
    """SearchReference -v with clustered variants (allele combinations, suffix re-chunking), multi-allelic records and indels: the variant windows' hits
    are merged with the reference hits, grouped per variant set and de-duplicated on the device; rows must equal the oracle's, ties across variant groups included."""
    g, contigs = small_genome
    arrays = [np.frombuffer(b, dtype=np.uint8) for _, b in contigs]
    vcf = synth.synthetic_vcf_fast(g, arrays, [(0, l) for l in g.lengths], 2500, cluster_fraction=0.3)
    # variants right on planted sites, so that hits really differ between haplotypes
    extra = []
    for c, lst in enumerate(g.planted):
        for (pos, seq) in lst[:12]:
            p = pos + 5
            ref = chr(arrays[c][p])
            if ref in "ACGT":
                extra.append("%s\t%d\tsite%d_%d\t%s\t%s\t.\tPASS\tAF=0.25" % (g.names[c], p + 1, c, p, ref, "ACGT"[("ACGT".index(ref) + 1) % 4]))
    body = [l for l in vcf.split("\n") if l and not l.startswith("#")] + extra
    order = {n: i for i, n in enumerate(g.names)}
    seen = set()
    body = [l for l in sorted(body, key=lambda l: (order[l.split("\t")[0]], int(l.split("\t")[1]))) if not ((l.split("\t")[0], l.split("\t")[1]) in seen or seen.add((l.split("\t")[0], l.split("\t")[1])))]
    vcf = "##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n" + "\n".join(body) + "\n"
    for guide, aux in ((synth.BASELINE_GUIDE, []), ("CTTGCCCCACAGGGCAGTAAngg", ["nag"])):
        exp = _lines(pyoracle.search_reference(contigs, guide, aux_pams=aux, vcf_text=vcf, raw=True, **kw))
        got = _lines(eng.search_reference(contigs, guide, aux_pams=aux, vcf_text=vcf, raw=True, **kw))
        assert sum("+variants" in l for l in exp) > 5
        assert got == exp, (guide, kw)
