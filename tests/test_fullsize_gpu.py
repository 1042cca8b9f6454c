"""BASELINE-size checks on the B200 (3.1-Gbp synthetic hg38-sized genome, SURVEY.md 8d config 3/4 shapes).  The oracle cannot run a whole genome
in test time, so parity at this size is established through
  * the oracle on 2-Mbp slices aligned to the window grid: every engine hit in the slice interior must equal the oracle's row, all columns;
  * size-independent properties: every planted site is reported, the table is in ReferenceHit.sort order, dedup output is a subset of the raw
    output, and the union of 2 contig-range shards is byte-identical to the single-engine table."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))

pytestmark = pytest.mark.gpu

GUIDES_N = 8
STEP = 970            # window 1000, overlap 23 + 5 + 3 - 1 (SearchReference.scala:529-530)


@pytest.fixture(scope="module")
def world():
    from calitas_b200 import synth
    from calitas_b200._capi import Engine, Limits
    guides = [synth.BASELINE_GUIDE] + synth.random_guides(GUIDES_N - 1)
    genome = synth.hg38_like_genome(1.0, guides=guides, sites_per_guide=250)
    arrays = [genome.contig(c) for c in range(len(genome.lengths))]
    e = Engine(0)
    ref = e.load_reference(list(zip(genome.names, arrays)))
    lim = Limits(5, 1, 3, -1, 10)
    hs = e.search(ref, guides, lim, window_size=1000, dedup=True)
    rec = hs.records()
    hs.free()
    yield dict(genome=genome, arrays=arrays, guides=guides, engine=e, ref=ref, lim=lim, rec=rec)
    ref.free()
    e.close()


def test_hit_table_is_in_reference_sort_order(world):
    r = world["rec"]
    assert r.size > GUIDES_N * 200000                      # ~107 hits/Mbp/guide on random sequence
    key = np.stack([r["guide_idx"].astype(np.int64), r["contig_idx"].astype(np.int64), r["guide_start_offset"].astype(np.int64),
                    (r["strand"] == ord("-")).astype(np.int64), -r["score"].astype(np.int64)])
    order = np.lexsort(key[::-1])
    assert np.array_equal(key[:, order], key)              # ReferenceHit.scala:276-287, per guide
    assert set(np.unique(r["strand"]).tolist()) == {ord("+"), ord("-")}
    assert (r["score"] >= 590 - 130 * 3 - 121 * 3).all()
    assert (r["edits"] <= 9).all() and (r["end_offset"] > r["start_offset"]).all()


def test_every_planted_site_is_reported(world):
    g, r = world["genome"], world["rec"]
    protos = [x[:20] for x in world["guides"]]
    missing = 0
    total = 0
    by_contig = {}
    for c in range(len(g.lengths)):
        m = r[r["contig_idx"] == c]
        by_contig[c] = m[np.argsort(m["start_offset"], kind="stable")]
    for c, lst in enumerate(g.planted):
        m = by_contig[c]
        starts = m["start_offset"]
        for pos, seq in lst:
            total += 1
            lo = np.searchsorted(starts, pos - 64)
            hi = np.searchsorted(starts, pos + len(seq))
            cand = m[lo:hi]
            ok = ((cand["start_offset"] < pos + len(seq)) & (cand["end_offset"] > pos)).any()
            missing += 0 if ok else 1
    assert total == GUIDES_N * 250 or total > GUIDES_N * 240
    assert missing == 0


def _oracle_rows(text, offset):
    import pyoracle
    rows = pyoracle.hits_table(text)
    return [(r["coordinate_start"] + offset, r["coordinate_end"] + offset, r["strand"], r["score"], r["cigar"], r["padded_guide"], r["padded_alignment"],
             r["padded_target"], r["total_mm_plus_gaps"], r["pam_used"]) for r in rows]


def _engine_rows(world, m, guides):
    from calitas_b200 import testing
    g, arrays = world["genome"], world["arrays"]
    text = world["engine"].render_alignments(m, guides, list(zip(g.names, arrays)), upper_case=True)
    rows = testing._table(text, testing._INT_GA)
    return [(x["guideStartOffset"], x["guideEndOffset"], x["strand"], x["score"], x["cigar"], x["paddedGuide"], x["paddedAlignment"], x["paddedTarget"],
             x["edits"], "".join(ch for ch in x["guide"] if ch.islower())) for x in rows]


@pytest.mark.parametrize("contig,k0", [(0, 30000), (7, 101), (23, 40011)])
def test_slices_match_the_oracle(world, contig, k0):
    import pyoracle
    g, arrays, guides, r = world["genome"], world["arrays"], world["guides"], world["rec"]
    start, length = k0 * STEP, 2_000_000
    assert start + length < g.lengths[contig]
    sl = bytes(arrays[contig][start:start + length])
    for gi in range(GUIDES_N):
        exp = [t for t in _oracle_rows(pyoracle.search_reference([(g.names[contig], sl)], guides[gi], raw=True, threads=8), start)
               if t[0] >= start + 1500 and t[1] <= start + length - 1500]
        m = r[(r["guide_idx"] == gi) & (r["contig_idx"] == contig) & (r["guide_start_offset"] >= start + 1500) & (r["guide_end_offset"] <= start + length - 1500)]
        got = _engine_rows(world, m, guides)
        assert len(exp) > 50
        assert got == exp, (contig, gi)


# BASELINE configs[3] shapes at full size: 6 guide diffs / 2 gaps with two PAMs, PAM-less, and a 5' PAM (DP runs on the reverse complement)
CONFIG4_GUIDES = [("CTTGCCCCACAGGGCAGTAAngg", ["nag"]), "CTTGCCCCACAGGGCAGTAA", "tttvCTTGCCCCACAGGGCAGTAA", ("GGGGCCACTAGGGACAGGATngg", ["nag"])]


@pytest.fixture(scope="module")
def world6(world):
    from calitas_b200._capi import Limits
    lim = Limits(6, 1, 2, -1, 10)
    hs = world["engine"].search(world["ref"], CONFIG4_GUIDES, lim, window_size=1000, dedup=True)
    rec = hs.records()
    hs.free()
    return dict(rec=rec, lim=lim)


@pytest.mark.parametrize("contig,k0", [(2, 52001), (21, 977)])
def test_config4_shapes_match_the_oracle_on_slices(world, world6, contig, k0):
    import pyoracle
    g, arrays, r = world["genome"], world["arrays"], world6["rec"]
    assert r.size > 4 * 1_000_000                          # ~1 100 hits/Mbp/guide at 6 diffs
    start, length = k0 * (1000 - (23 + 6 + 2 - 1)), 1_000_000
    assert start + length < g.lengths[contig]
    sl = bytes(arrays[contig][start:start + length])
    for gi, gd in enumerate(CONFIG4_GUIDES):
        seq, aux = (gd, []) if isinstance(gd, str) else gd
        step_g = 1000 - (len(seq) + 6 + 2 - 1)            # the window grid depends on the raw guide length (SearchReference.scala:528-530)
        s0 = (start // step_g + 1) * step_g               # slice aligned to this guide's grid
        sl = bytes(arrays[contig][s0:s0 + length])
        exp = [t for t in _oracle_rows(pyoracle.search_reference([(g.names[contig], sl)], seq, aux_pams=aux, raw=True, threads=8, d=6, g=2), s0)
               if t[0] >= s0 + 1500 and t[1] <= s0 + length - 1500]
        m = r[(r["guide_idx"] == gi) & (r["contig_idx"] == contig) & (r["guide_start_offset"] >= s0 + 1500) & (r["guide_end_offset"] <= s0 + length - 1500)]
        got = _engine_rows(world, m, [x if isinstance(x, str) else x for x in CONFIG4_GUIDES])
        assert len(exp) > 100, (gi, len(exp))
        assert got == exp, (contig, gi)


def test_config5_vcf_slice_matches_the_oracle(world):
    """BASELINE configs[4]: SearchReference -v.  A synthetic PrepareVcf-shaped VCF over the first 2 Mbp of chrY, searched with -c chrY on the full genome
    (reference windows + variant windows merged and de-duplicated on the device); every row that ends inside the slice must equal the oracle's row
    (all 34 columns; the oracle sees the slice as a contig of the same name, so coordinates, flanks and variant descriptions need no shifting)."""
    import pyoracle
    from calitas_b200 import synth, testing
    g, arrays = world["genome"], world["arrays"]
    c, length = 23, 2_000_000
    sub = synth.Genome([g.names[c]], [length], g.seed)
    vcf = synth.synthetic_vcf(sub, [arrays[c][:length]], 6000)
    assert vcf.count("\n") > 5000
    fac = testing.Facade()
    contigs = list(zip(g.names, arrays))
    for guide, aux, kw in ((synth.BASELINE_GUIDE, [], {}), ("CTTGCCCCACAGGGCAGTAAngg", ["nag"], dict(d=6, g=2))):
        exp = [l for l in pyoracle.search_reference([(g.names[c], bytes(arrays[c][:length]))], guide, aux_pams=aux, vcf_text=vcf, raw=True, threads=8, **kw).split("\n") if l]
        got = [l for l in fac.search_reference(contigs, guide, aux_pams=aux, vcf_text=vcf, chrom=g.names[c], raw=True, **kw).split("\n") if l]
        inside = lambda l: int(l.split("\t")[5]) <= length - 1500            # coordinate_end
        e_rows, g_rows = [l for l in exp[1:] if inside(l)], [l for l in got[1:] if inside(l)]
        assert exp[0] == got[0] and len(e_rows) > 150 and any("+variants" in l for l in e_rows)
        assert g_rows == e_rows, guide


def test_dedup_output_is_a_subset_of_the_raw_output(world):
    e, ref, guides, lim = world["engine"], world["ref"], world["guides"], world["lim"]
    raw = e.search(ref, guides[:2], lim, window_size=1000, dedup=False).records()
    ded = world["rec"][world["rec"]["guide_idx"] < 2]
    assert raw.size > ded.size

    def keyset(a):
        v = a.copy()
        v["task_idx"] = 0                                  # the same locus found from two overlapping windows differs only in the window id
        b, sz = v.tobytes(), v.dtype.itemsize
        return {b[i:i + sz] for i in range(0, len(b), sz)}
    ks_raw, ks_ded = keyset(raw), keyset(ded)
    assert ks_ded <= ks_raw


def test_two_shards_union_equals_whole_at_full_size(world):
    from calitas_b200 import multi
    import ctypes as C
    e, g, arrays, guides, lim = world["engine"], world["genome"], world["arrays"], world["guides"], world["lim"]
    n = len(g.lengths)
    L = (C.c_int64 * n)(*g.lengths)
    parts = []
    for s in range(2):
        ob, oe, hb, he = [(C.c_int64 * n)() for _ in range(4)]
        e.lib.check(e.lib.L.calitas_shard_plan(n, L, s, 2, C.c_int64(4000), ob, oe, hb, he))
        ref = e.load_reference_ranges(g.names, g.lengths, [(hb[c], he[c]) for c in range(n)], [(ob[c], oe[c]) for c in range(n)],
                                      [np.ascontiguousarray(arrays[c][hb[c]:he[c]]) if he[c] > hb[c] else None for c in range(n)])
        hs = e.search(ref, guides, lim, window_size=1000, dedup=True)
        parts.append(hs.records())
        hs.free()
        ref.free()
    merged = multi.merge_shard_records(parts)
    whole = world["rec"]
    a, b = merged.copy(), whole.copy()
    a["task_idx"] = 0
    b["task_idx"] = 0                                       # window ids are shard-local bookkeeping
    assert a.size == b.size
    assert a.tobytes() == b.tobytes()
