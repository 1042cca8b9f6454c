"""A slice of the randomised hunt of scratch/fuzz_hostsim.py in the CPU suite: random guides (IUPAC, 5'/3'/no PAM, auxiliary PAMs), contigs from one
base to a few windows with planted sites, N runs, lower case, IUPAC and junk bytes, random limits, window sizes and cost sets; the engine sources
compiled for the host must give the oracle's table (or fail where the oracle fails) for every seed."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scratch"))


def test_forty_random_searches_match_the_oracle():
    import fuzz_hostsim
    assert fuzz_hostsim.hunt(40, 31000, verbose=False) == []
