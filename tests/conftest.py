import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def read_fasta(path):
    contigs, name, chunks = [], None, []
    for line in open(path):
        line = line.rstrip("\n")
        if line.startswith(">"):
            if name is not None:
                contigs.append((name, "".join(chunks)))
            name, chunks = line[1:].split()[0], []
        else:
            chunks.append(line)
    if name is not None:
        contigs.append((name, "".join(chunks)))
    return contigs


@pytest.fixture(scope="session")
def sga_ref():
    """Fixture genome of SequentialGuideAlignerTest.scala:12-44."""
    return read_fasta(os.path.join(ROOT, "tests", "golden", "sga_test_ref.fa"))


@pytest.fixture(scope="session", params=["oracle", "hostsim", pytest.param("gpu", marks=pytest.mark.gpu)])
def backend(request):
    """The same reference-test bodies run against the CPU oracle (here) and against the CUDA engine through its C ABI (-m gpu)."""
    import backends
    return backends.get(request.param)
