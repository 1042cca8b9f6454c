"""Frozen tables (tests/golden/frozen_*.tsv, written by tests/golden/make_fixtures.py frozen): the oracle's output on three seeded inputs as of
round 1.  They are NOT reference output (no JVM here); they stop the oracle and the engine from drifting together unnoticed -- every other
parity test compares the two with each other.  The oracle, the host simulation and (on a B200) the CUDA engine must all reproduce them."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pyoracle
from calitas_b200 import synth

GOLDEN = os.path.join(ROOT, "tests", "golden")
IMPLS = ["oracle", "hostsim", pytest.param("gpu", marks=pytest.mark.gpu)]


@pytest.fixture(scope="module")
def inputs():
    g = synth.config1_genome(scale=0.02, n_sites=60)
    return g, [(n, bytes(b)) for n, b in g.contigs()]


@pytest.fixture(scope="module", params=IMPLS)
def impl(request):
    import backends
    return backends.get(request.param)


def frozen(name):
    return open(os.path.join(GOLDEN, name)).read()


def test_search_defaults(impl, inputs):
    g, contigs = inputs
    assert impl.search_reference(contigs, synth.BASELINE_GUIDE, guide_id="g", assembly="SYN10M", raw=True) == frozen("frozen_search_defaults.tsv")


def test_search_d6_two_pams(impl, inputs):
    g, contigs = inputs
    got = impl.search_reference(contigs, "CTTGCCCCACAGGGCAGTAAngg", aux_pams=["nag"], guide_id="g", assembly="SYN10M", raw=True, d=6, g=2, p=1)
    assert got == frozen("frozen_search_d6_g2_ngg_nag.tsv")


def test_align_to_reference_best(impl, inputs):
    g, contigs = inputs
    got = impl.align_to_reference(contigs, synth.a2r_tasks(g, [synth.BASELINE_GUIDE], 40), window_size=60, raw=True, assembly="SYN10M")
    assert got == frozen("frozen_a2r_best_w60.tsv")
