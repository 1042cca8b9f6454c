"""Uniform facade over the two implementations the parity tests compare:
   "oracle" = oracle/pyoracle.py (CPU restatement, test infrastructure);
   "gpu"    = calitas_b200 (CUDA engine through the C ABI)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


class OracleBackend:
    name = "oracle"

    def __init__(self):
        import pyoracle
        pyoracle.build()
        self.o = pyoracle

    def align(self, guide, target, **kw):
        return self.o.align(guide, target, **kw)

    def align_best(self, guide, target, **kw):
        return self.o.align_best(guide, target, **kw)

    def align_to_ref_best(self, contigs, guide, chrom, pos, window_size=None, max_gaps=3):
        return self.o.align_to_ref(contigs, guide, chrom, pos, window_size=window_size, best=True, max_gaps=max_gaps)

    def align_to_ref(self, contigs, guide, chrom, pos, window_size=None, **kw):
        return self.o.align_to_ref(contigs, guide, chrom, pos, window_size=window_size, best=False, **kw)

    def search_reference(self, contigs, guide, **kw):
        return self.o.search_reference(contigs, guide, **kw)

    def align_to_reference(self, contigs, tasks, **kw):
        return self.o.align_to_reference(contigs, tasks, **kw)


class GpuBackend:
    """name == "gpu": the product library on a CUDA device.  name == "hostsim": the same engine sources compiled for the host
    by tests/hostsim (test-only; checks the engine's logic where no GPU exists)."""

    def __init__(self, name="gpu"):
        import calitas_b200.testing as t
        from calitas_b200._capi import Library
        self.name = name
        if name == "hostsim":
            import subprocess
            d = os.path.join(ROOT, "tests", "hostsim")
            subprocess.check_call(["make", "-C", d, "-s"])
            self.t = t.Facade(Library(os.path.join(d, "_build", "libcalitas_hostsim.so")))
        else:
            self.t = t.Facade()

    def align(self, guide, target, **kw):
        return self.t.align(guide, target, **kw)

    def align_best(self, guide, target, **kw):
        return self.t.align_best(guide, target, **kw)

    def align_to_ref_best(self, contigs, guide, chrom, pos, window_size=None, max_gaps=3):
        return self.t.align_to_ref(contigs, guide, chrom, pos, window_size=window_size, best=True, max_gaps=max_gaps)

    def align_to_ref(self, contigs, guide, chrom, pos, window_size=None, **kw):
        return self.t.align_to_ref(contigs, guide, chrom, pos, window_size=window_size, best=False, **kw)

    def search_reference(self, contigs, guide, **kw):
        return self.t.search_reference(contigs, guide, **kw)

    def align_to_reference(self, contigs, tasks, **kw):
        return self.t.align_to_reference(contigs, tasks, **kw)


_cache = {}


def get(name):
    if name not in _cache:
        _cache[name] = OracleBackend() if name == "oracle" else GpuBackend(name)
    return _cache[name]
