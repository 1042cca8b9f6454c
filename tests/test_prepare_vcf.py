"""`calitas PrepareVcf` (PrepareVcf.scala:43-91) against the oracle's restatement and the reference's own test (PrepareVcfTest.scala:10-43), and the
prepared file fed to SearchReference -v.  PrepareVcf is host code (no alignment): the hostsim and the product CLI run the same sources."""
import gzip
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pyoracle

HEADER = """##fileformat=VCFv4.2
##FILTER=<ID=PASS,Description="Passes all filters.">
##FILTER=<ID=LowQual,Description="Low quality">
##INFO=<ID=AF,Number=A,Type=Float,Description="ALT allele frequency">
##INFO=<ID=DP,Number=1,Type=Integer,Description="Depth">
##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">
##contig=<ID=chr1,length=10000000>
##contig=<ID=1,length=10000000>
##reference=file:///old.fa
#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tsample1\tsample2
"""


@pytest.fixture(scope="module")
def calitas():
    d = os.path.join(ROOT, "tests", "hostsim")
    subprocess.check_call(["make", "-C", d, "-s", "all"])
    return os.path.join(d, "_build", "calitas_hostsim")


def run(calitas, *args):
    return subprocess.run([calitas] + [str(a) for a in args], capture_output=True, text=True, timeout=120)


def parse(text):
    head = [l for l in text.split("\n") if l.startswith("#")]
    recs = []
    for l in text.split("\n"):
        if l and not l.startswith("#"):
            f = l.split("\t")
            assert len(f) == 8 and f[6] == "PASS" and f[7].startswith("AF=")
            recs.append((f[0], f[1], f[2], f[3], f[4].split(","), f[5], f[7][3:].split(",")))
    return head, recs


def test_reference_test_case(calitas, tmp_path):
    """PrepareVcfTest.scala:10-43: ten chr1 SNPs with AF 0.5 and two genotyped samples -> no samples, ten records, .vcf.gz output."""
    body = "".join("chr1\t%d\t.\tA\tC\t.\tPASS\tAF=0.5\tGT\t0/1\t./.\n" % (1000 * (i + 1)) for i in range(10))
    vin, vout = tmp_path / "in.vcf", tmp_path / "prepared.vcf.gz"
    vin.write_text(HEADER + body)
    p = run(calitas, "PrepareVcf", "-i", vin, "-o", vout)
    assert p.returncode == 0, p.stderr
    raw = open(vout, "rb").read()
    assert raw[:4] == b"\x1f\x8b\x08\x04" and raw[12:14] == b"BC" and raw[-28:-26] == b"\x1f\x8b"      # BGZF members + the empty EOF block
    head, recs = parse(gzip.decompress(raw).decode())
    assert head[-1] == "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO"                                  # header.samples.size shouldBe 0
    assert len(recs) == 10 and all(r[4] == ["C"] and r[6] == ["0.500"] for r in recs)


def test_filters_and_rewrites_match_oracle(calitas, tmp_path):
    rng = np.random.default_rng(7)
    rows = []
    for i in range(400):
        chrom = ["1", "chr1", "X", "MT", "22", "GL000207.1"][int(rng.integers(6))]
        n_alt = int(rng.integers(1, 4))
        ref = "".join(rng.choice(list("ACGT"), int(rng.integers(1, 4))))
        alts = ["".join(rng.choice(list("ACGT"), int(rng.integers(1, 5)))) for _ in range(n_alt)]
        kind = int(rng.integers(12))
        if kind == 0:
            alts[0] = "<DEL>"
        elif kind == 1:
            alts[-1] = "*"
        elif kind == 2:
            alts[0] = "G]17:198982]"
        afs = ["%.6g" % v for v in rng.choice([0.5, 0.0625, 0.01, 0.0100001, 0.009, 0.0004996, 1.0, 0.25, 0.1235, 0.02, 3e-5], n_alt)]
        if kind == 3:
            afs[0] = "."
        flt = ["PASS", "PASS", "PASS", "LowQual", ".", "PASS;LowQual"][int(rng.integers(6))]
        qual = [".", "50", "37.455", "12.5", "99.995"][int(rng.integers(5))]
        info = ";".join(["DP=%d" % rng.integers(100), "AF=" + ",".join(afs), "DB"][: int(rng.integers(2, 4))])
        if not info.count("AF="):
            info += ";AF=" + ",".join(afs)
        rows.append("%s\t%d\t%s\t%s\t%s\t%s\t%s\t%s\tGT\t0/1\t1/1" % (chrom, 100 + 10 * i, "rs%d" % i if i % 3 else ".", ref, ",".join(alts), qual, flt, info))
    text = HEADER + "\n".join(rows) + "\n"
    half = len(rows) // 2
    a, b = tmp_path / "a.vcf", tmp_path / "b.vcf.gz"
    a.write_text(HEADER + "\n".join(rows[:half]) + "\n")
    with gzip.open(b, "wt") as f:
        f.write(HEADER.replace("sample2", "other") + "\n".join(rows[half:]) + "\n")
    for min_af, chr_flag in ((None, None), ("0.05", "false"), ("0.0004", "true")):
        out = tmp_path / "out.vcf"
        args = ["PrepareVcf", "-i", a, b, "-o", out] + (["-f", min_af] if min_af else []) + (["-c", chr_flag] if chr_flag else [])
        p = run(calitas, *args)
        assert p.returncode == 0, p.stderr
        head, recs = parse(open(out).read())
        exp = pyoracle.prepare_vcf([text], min_af=float(min_af) if min_af else 0.01, add_chr_prefix=chr_flag != "false")
        assert len(exp) > 20 and recs == exp
        assert head[0] == "##fileformat=VCFv4.2" and "##reference=file:///old.fa" in head and head[-1].count("\t") == 7


def test_dict_overrides_contigs(calitas, tmp_path):
    vin, out, dct = tmp_path / "in.vcf", tmp_path / "out.vcf", tmp_path / "ref.dict"
    vin.write_text(HEADER + "1\t100\t.\tA\tC,G\t.\tPASS\tAF=0.5,0.001\n")
    dct.write_text("@HD\tVN:1.5\n@SQ\tSN:chr1\tLN:248956422\tAS:hg38\tM5:abc\n@SQ\tSN:chr2\tLN:242193529\tAS:hg38\n")
    p = run(calitas, "PrepareVcf", "-i", vin, "-o", out, "-d", dct)
    assert p.returncode == 0, p.stderr
    head, recs = parse(open(out).read())
    assert [h for h in head if h.startswith("##contig")] == ["##contig=<ID=chr1,length=248956422,assembly=hg38>", "##contig=<ID=chr2,length=242193529,assembly=hg38>"]
    assert [h for h in head if h.startswith("##reference")] == ["##reference=hg38"]
    assert recs == [("chr1", "100", ".", "A", ["C"], ".", ["0.500"])]


def test_errors(calitas, tmp_path):
    vin = tmp_path / "in.vcf"
    vin.write_text(HEADER + "1\t100\t.\tA\tC\t.\tPASS\tDP=3\n")
    p = run(calitas, "PrepareVcf", "-i", vin, "-o", tmp_path / "o.vcf")
    assert p.returncode == 2 and "AF" in p.stderr                       # the reference throws NoSuchElementException on a PASS record without AF
    p = run(calitas, "PrepareVcf", "-o", tmp_path / "o.vcf")
    assert p.returncode == 2 and "input" in p.stderr
    p = run(calitas, "PrepareVcf", "-i", tmp_path / "missing.vcf", "-o", tmp_path / "o.vcf")
    assert p.returncode == 2 and "non-existent" in p.stderr


def test_number_formats():
    """htsjdk formatVCFDouble / formatQualValue cases, incl. the half-up ties java.util.Formatter rounds away from printf's half-even."""
    f = pyoracle.format_vcf_double
    assert [f(x) for x in (0.5, 0.0625, 1.0, 12.345, 0.005, 0.00049996, 0.0, 1e-21, 0.9996)] == ["0.500", "0.063", "1.00", "12.35", "5.000e-03", "5.000e-04", "0.00", "0.00", "1.000"]
    assert [pyoracle.format_qual(x) for x in (50.0, 37.455, 12.5, 99.995)] == ["50", "37.46", "12.50", "100"]
