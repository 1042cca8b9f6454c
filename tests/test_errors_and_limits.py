"""Error behaviour at the boundary (SURVEY.md 8b "Errors"): the reference's require()/validate() failures come back as CALITAS_EINVAL with the
reference's message, engine limits as CALITAS_ELIMIT, and nothing is ever silently truncated (a candidate pool that is too small is grown and
the scan repeated)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import pyoracle
from calitas_b200 import synth
from calitas_b200._capi import DEFAULT_COSTS, CalitasError, Engine, Limits

ENGINES = ["hostsim", pytest.param("gpu", marks=pytest.mark.gpu)]


@pytest.fixture(scope="module", params=ENGINES)
def eng(request):
    import backends
    return backends.get(request.param)


def code_of(fn):
    with pytest.raises(CalitasError) as ei:
        fn()
    return ei.value.code, ei.value.message


def test_guide_validation_messages(eng):
    t = "ACGTACGTACGTACGTACGTACGTACGTAGG"
    kw = dict(max_guide_diffs=2, max_gaps=1, max_pam_diffs=1, max_total_diffs=4)
    assert code_of(lambda: eng.align("acgtacgtacgtacgtacgt", t, **kw)) == (1, "requirement failed: Guide sequence cannot be all lower case.")     # SequentialGuideAligner.scala:84-87
    c, m = code_of(lambda: eng.align("ACGTacgtACGT", t, **kw))
    assert c == 1 and "Invalid Guide sequence" in m
    c, m = code_of(lambda: eng.align("ACGTACGTACGTACGTACGT", t, aux_pams=["ngg"], **kw))
    assert c == 1 and "Cannot provide auxiliary PAMs" in m
    c, m = code_of(lambda: eng.align("ACGTACGTACGTACGTACGTngg", t, aux_pams=["NAG"], **kw))
    assert c == 1 and "All PAMs must be lower case" in m


def test_engine_limits_are_reported_not_truncated(eng):
    t = "ACGT" * 30
    kw = dict(max_guide_diffs=2, max_gaps=1, max_pam_diffs=1, max_total_diffs=4)
    c, m = code_of(lambda: eng.align("A" * 33 + "ngg", t, **kw))
    assert c == 3 and "32" in m                                     # CALITAS_ELIMIT: protospacer > 32 nt
    c, m = code_of(lambda: eng.align("ACGTACGTACGTACGTACGT" + "n" * 17, t, **kw))
    assert c == 3
    c, m = code_of(lambda: eng.align("ACGTACGTACGTACGTACGTngg", t, aux_pams=["nag"] * 8, **kw))
    assert c == 3
    # exactly at the limits works and matches the oracle
    g32 = "ACGTTGCAACGTTGCAACGTTGCAACGTTGCA" + "nggnagnggnagngga"
    tgt = "TT" + "ACGTTGCAACGTTGCAACGTTGCAACGTTGCA" + "AGGTAGCGGAAGTGGA" + "CC"
    assert eng.align(g32, tgt, **kw) == pyoracle.align(g32, tgt, **kw)


def test_unknown_chromosome_and_window_size(eng):
    g = synth.config1_genome(scale=0.005, n_sites=5)
    contigs = [(n, bytes(b)) for n, b in g.contigs()]
    with pytest.raises(CalitasError) as ei:
        eng.search_reference(contigs, synth.BASELINE_GUIDE, chrom="chrNope")
    assert ei.value.code == 1 and "Unknown chromosome" in ei.value.message
    with pytest.raises(CalitasError) as ei:
        eng.search_reference(contigs, synth.BASELINE_GUIDE, window_size=30)          # step <= 0 (SearchReference.scala:529-530; Range throws)
    assert ei.value.code == 1


def test_candidate_pool_growth_never_truncates(eng):
    """First call with a deliberately tiny candidate pool hint: the engine must grow it and still return the full, exact table."""
    g = synth.config1_genome(scale=0.02, n_sites=120)
    contigs = [(n, b) for n, b in g.contigs()]
    guides = [synth.BASELINE_GUIDE] + synth.random_guides(19)
    e = Engine(0, lib=eng.t.lib)
    ref = e.load_reference(contigs)
    first = e.search(ref, guides, Limits(6, 1, 2, -1, 10)).records()             # d = 6: ~10x the candidates of d = 5; 20 guides -> 2 pipelined chunks
    again = e.search(ref, guides, Limits(6, 1, 2, -1, 10)).records()
    assert first.size > 500 and first.tobytes() == again.tobytes()
    exp = 0
    for gd in guides[:3]:
        exp += len(pyoracle.search_reference([(n, bytes(b)) for n, b in contigs], gd, d=6, g=2, p=1)) 
    assert int((first["guide_idx"] < 3).sum()) == exp
    ref.free()
    e.close()


def test_empty_and_tiny_inputs(eng):
    kw = dict(max_guide_diffs=5, max_gaps=3, max_pam_diffs=1, max_total_diffs=9)
    assert eng.align(synth.BASELINE_GUIDE, "ACGT", **kw) == pyoracle.align(synth.BASELINE_GUIDE, "ACGT", **kw) == []
    assert eng.align(synth.BASELINE_GUIDE, "N" * 50, **kw) == []
    contigs = [("tiny", b"A"), ("short", b"ACGTACGTAC"), ("n", b"N" * 2000)]
    assert [l for l in eng.search_reference(contigs, synth.BASELINE_GUIDE, raw=True).split("\n") if l] == \
           [l for l in pyoracle.search_reference(contigs, synth.BASELINE_GUIDE, raw=True).split("\n") if l]


def test_per_shard_dedup_with_nonpositive_max_overlap_is_refused(eng):
    """-O <= 0 makes every later hit of a group "overlap": the reference's sweep (SearchReference.scala:662-672) then reaches across the whole
    contig and a shard cannot reproduce it.  calitas_search says so instead of returning a table that only looks right; a whole reference, or
    dedup = 0 on a shard, is fine."""
    g = synth.config1_genome(scale=0.01, n_sites=20)
    contigs = [(n, b) for n, b in g.contigs()]
    e = Engine(0, lib=eng.t.lib)
    shard = e.load_reference(contigs, shard=(0, 2, 4000))
    c, m = code_of(lambda: e.search(shard, [synth.BASELINE_GUIDE], Limits(5, 1, 3, -1, 0), dedup=True))
    assert c == 1 and "max_overlap <= 0" in m
    assert len(e.search(shard, [synth.BASELINE_GUIDE], Limits(5, 1, 3, -1, 0), dedup=False)) >= 0
    assert len(e.search(shard, [synth.BASELINE_GUIDE], Limits(5, 1, 3, -1, 1), dedup=True)) >= 0
    shard.free()
    whole = e.load_reference(contigs)
    assert len(e.search(whole, [synth.BASELINE_GUIDE], Limits(5, 1, 3, -1, 0), dedup=True)) >= 1
    whole.free()
    e.close()


def test_align_targets_batch_equals_single_calls(eng):
    """calitas_align_targets gathers the tasks' bases on several host threads into one staging buffer and packs them on the device: a batch of ragged
    targets (1 base to a few hundred, two guides, non-zero offsets) must give, task by task, the records of the single-task calls."""
    rng = np.random.default_rng(23)
    e = eng.t.engine(DEFAULT_COSTS)
    guides = ["CTTGCCCCACAGGGCAGTAAnrg", "tttvGACCCCCTCCACCCCGCCTC"]
    lim = Limits(4, 1, 2, -1, 10)
    site = "CTTGCCCCACAGGGCAGTAATGG"
    tasks = []
    for i in range(3000):
        n = int(rng.integers(1, 300))
        t = "".join(rng.choice(list("ACGT"), size=n))
        if n > 40 and rng.random() < 0.5:
            p = int(rng.integers(0, n - len(site)))
            s = list(site); s[int(rng.integers(0, 20))] = "A"
            t = t[:p] + "".join(s) + t[p + len(site):]
        tasks.append((int(rng.integers(0, 2)), t, int(rng.integers(0, 5000))))
    batch = e.align_targets(guides, tasks, lim).records()
    assert batch.size > 500
    by_task = {}
    for r in batch:
        by_task.setdefault(int(r["task_idx"]), []).append(r)
    for i in list(rng.choice(len(tasks), size=60, replace=False)) + [0, len(tasks) - 1]:
        single = e.align_targets(guides, [tasks[int(i)]], lim).records()
        got = by_task.get(int(i), [])
        assert len(got) == single.size, i
        for a, b in zip(got, single):
            for f in batch.dtype.names:
                if f != "task_idx":
                    assert np.array_equal(a[f], b[f]), (i, f)
