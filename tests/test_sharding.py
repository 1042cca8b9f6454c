"""Contig-range sharding: the union of the per-shard hit sets must equal the single-engine hit set, bit for bit.
Shards are independent engines (one per GPU in production); here they run one after another on the same device / host simulation."""
import os
import sys

import numpy as np
import pytest

from calitas_b200 import synth
from calitas_b200._capi import Engine, Limits, Library

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ENGINES = ["hostsim", pytest.param("gpu", marks=pytest.mark.gpu)]


@pytest.fixture(scope="module", params=ENGINES)
def lib(request):
    if request.param == "hostsim":
        import backends
        return backends.get("hostsim").t.lib
    return Library()


def hit_tuple(h):
    return (h.guide_idx, h.pam_idx, h.contig_idx, h.start_offset, h.end_offset, h.guide_start_offset, h.guide_end_offset, h.score, chr(h.strand), h.n_ops,
            h.gap_bases, h.edits, tuple(h.ops))


@pytest.mark.parametrize("dedup", [True, False])
def test_shards_union_equals_whole(lib, dedup):
    g = synth.config1_genome(scale=0.02, n_sites=80)
    contigs = [(n, b) for n, b in g.contigs()]
    guides = [synth.BASELINE_GUIDE, ("CTTGCCCCACAGGGCAGTAAngg", ["nag"]), "GGGGCCACTAGGGACAGGAT"]
    lim = Limits(5, 1, 3, -1, 10)
    e = Engine(0, lib=lib)
    ref = e.load_reference(contigs)
    whole = [hit_tuple(h) for h in e.search(ref, guides, lim, dedup=dedup).hits()]
    assert len(whole) > 60
    ref.free()
    for n_shards in (2, 3, 7):
        parts = []
        for s in range(n_shards):
            r = e.load_reference(contigs, shard=(s, n_shards, 4000))
            parts.append([hit_tuple(h) for h in e.search(r, guides, lim, dedup=dedup).hits()])
            r.free()
        union = [t for p in parts for t in p]
        # per-shard lists are guide-major; the global order is guide-major too, so compare per guide in shard order
        merged = [t for gi in range(len(guides)) for p in parts for t in p if t[0] == gi]
        assert sorted(union) == sorted(whole), n_shards
        assert merged == whole, n_shards
    e.close()


@pytest.mark.parametrize("max_overlap", [10, 25, 1])
def test_merge_shard_records_at_cuts_in_dense_repeats(lib, max_overlap):
    """calitas_b200.multi.merge_shard_records (the one-process-per-GPU gather): merged per-shard records == the single-engine records, also where
    the window overlap makes hits of neighbouring shards interleave by start (large -O: removeOverlaps collapses nothing)."""
    from calitas_b200 import multi
    import test_parity_random as T
    contigs = [(n, np.frombuffer(b, dtype=np.uint8)) for n, b in T._tandem_repeat_contigs(507, 200)]
    guides = [synth.BASELINE_GUIDE, ("CTTGCCCCACAGGGCAGTAAngg", ["nag"])]
    lim = Limits(5, 1, 3, -1, max_overlap)
    e = Engine(0, lib=lib)
    ref = e.load_reference(contigs)
    whole = e.search(ref, guides, lim, dedup=True).records()
    ref.free()
    assert whole.size > 300
    for n_shards in (2, 3, 5):
        parts = []
        for s in range(n_shards):
            r = e.load_reference(contigs, shard=(s, n_shards, 4000))
            parts.append(e.search(r, guides, lim, dedup=True).records())
            r.free()
        merged = multi.merge_shard_records(parts)
        drop = "task_idx"                                                  # window ids are engine-local bookkeeping
        keep = [n for n in whole.dtype.names if n != drop]
        assert merged.size == whole.size and all((merged[n] == whole[n]).all() for n in keep), n_shards
    e.close()


def test_shard_plan_covers_genome(lib):
    import ctypes as C
    lengths = [1000, 5, 123456, 1, 777]
    n = len(lengths)
    L = (C.c_int64 * n)(*lengths)
    for shards in (1, 2, 3, 8):
        owned = [0] * n
        for s in range(shards):
            ob, oe, hb, he = [(C.c_int64 * n)() for _ in range(4)]
            lib.check(lib.L.calitas_shard_plan(n, L, s, shards, C.c_int64(100), ob, oe, hb, he))
            for c in range(n):
                assert 0 <= hb[c] <= ob[c] <= oe[c] <= he[c] <= lengths[c]
                owned[c] += oe[c] - ob[c]
        assert owned == lengths


@pytest.mark.parametrize("max_overlap", [10, 25])
def test_search_sharded_returns_one_merged_table(lib, max_overlap):
    """calitas_search_sharded: N engines on host threads, one table in one address space == the single-engine table (dense repeats across the cuts,
    where neighbouring shards' hits interleave by start)."""
    import test_parity_random as T
    contigs = [(n, np.frombuffer(b, dtype=np.uint8)) for n, b in T._tandem_repeat_contigs(511, 220)] + [(n, b) for n, b in synth.config1_genome(scale=0.01, n_sites=30).contigs()]
    guides = [synth.BASELINE_GUIDE, ("CTTGCCCCACAGGGCAGTAAngg", ["nag"]), "GGGGCCACTAGGGACAGGAT"]
    lim = Limits(5, 1, 3, -1, max_overlap)
    e = Engine(0, lib=lib)
    ref = e.load_reference(contigs)
    whole = e.search(ref, guides, lim, dedup=True).records()
    ref.free()
    assert whole.size > 200
    for n_shards in (2, 3, 5):
        engines = [e] + [Engine(0, lib=lib) for _ in range(n_shards - 1)]
        refs = [en.load_reference(contigs, shard=(s, n_shards, 4000)) for s, en in enumerate(engines)]
        hs = Engine.search_sharded(engines, refs, guides, lim)
        got = hs.records()
        hs.free()
        for r in refs:
            r.free()
        for en in engines[1:]:
            en.close()
        a, b = got.copy(), whole.copy()
        a["task_idx"] = 0
        b["task_idx"] = 0
        assert a.size == b.size and a.tobytes() == b.tobytes(), n_shards
    e.close()


def test_halo_sentinel_fails_the_call_and_the_tool_falls_back(lib, monkeypatch):
    """k_sweep's halo sentinel: when no restart of the removeOverlaps loop lies within reach of a shard's first owned hit, calitas_search fails with
    CALITAS_ELIMIT instead of guessing, and calitas_tool_search_reference_batch gathers raw hits and de-duplicates on the host.  No real input has
    triggered it, so the test shortens the reach (CALITAS_TEST_HALO_BASES) until ordinary dense repeats do."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import pyoracle
    import test_parity_random as T
    from calitas_b200 import _capi
    from calitas_b200.testing import Facade
    contigs = T._tandem_repeat_contigs(503, 160)
    lim = Limits(5, 1, 3, -1, 10)
    e = Engine(0, lib=lib)
    r = e.load_reference(contigs, shard=(1, 3, 4000))
    assert len(e.search(r, [synth.BASELINE_GUIDE], lim, dedup=True)) > 10          # the real reach: fine
    monkeypatch.setenv("CALITAS_TEST_HALO_BASES", "-100000")                          # nothing counts as within reach: every cut with halo hits is flagged
    with pytest.raises(_capi.CalitasError) as ei:
        e.search(r, [synth.BASELINE_GUIDE], lim, dedup=True)
    assert ei.value.code == 3 and "halo" in ei.value.message                          # CALITAS_ELIMIT
    r.free()
    e.close()
    exp = [l for l in pyoracle.search_reference(contigs, synth.BASELINE_GUIDE, guide_id="g0", raw=True).split("\n") if l]
    got = [l for l in Facade(lib).search_reference_batch(contigs, [synth.BASELINE_GUIDE], ["g0"], n_shards=3).split("\n") if l]
    assert got == exp and len(exp) > 100
