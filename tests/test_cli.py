"""The `calitas` command line (SearchReference.scala:452-470, AlignToReference.scala:35-50 flags) end to end: FASTA + .fai + .dict (+ VCF, task table)
on disk -> hit table on disk, compared line by line with the oracle.  Here the CLI is linked against the test-only host simulation; the -m gpu
twin runs the product binary calitas_b200/calitas on the B200."""
import hashlib
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pyoracle
from calitas_b200 import synth

BINARIES = ["hostsim", pytest.param("gpu", marks=pytest.mark.gpu)]


@pytest.fixture(scope="module", params=BINARIES)
def calitas(request):
    if request.param == "hostsim":
        d = os.path.join(ROOT, "tests", "hostsim")
        subprocess.check_call(["make", "-C", d, "-s", "all"])
        return os.path.join(d, "_build", "calitas_hostsim")
    path = os.path.join(ROOT, "calitas_b200", "calitas")
    assert os.path.exists(path), "run __graft_entry__.build() first"
    return path


@pytest.fixture(scope="module")
def ref_dir(tmp_path_factory):
    d = tmp_path_factory.mktemp("ref")
    g = synth.config1_genome(scale=0.02, n_sites=60)
    contigs = [(n, bytes(b)) for n, b in g.contigs()]
    # soft-masked stretch: SearchReference upper-cases windows, AlignToReference does not (SequentialGuideAligner.scala:374)
    name0, b0 = contigs[0]
    contigs[0] = (name0, b0[:30000] + b0[30000:30400].lower() + b0[30400:])
    fa = d / "ref.fa"
    off, fai = 0, []
    with open(fa, "wb") as f:
        for n, b in contigs:
            hdr = (">%s some description\n" % n).encode()
            f.write(hdr)
            off += len(hdr)
            fai.append("%s\t%d\t%d\t60\t61" % (n, len(b), off))
            for i in range(0, len(b), 60):
                f.write(b[i:i + 60] + b"\n")
            off += len(b) + (len(b) + 59) // 60
    open(str(fa) + ".fai", "w").write("\n".join(fai) + "\n")
    open(d / "ref.dict", "w").write("@HD\tVN:1.5\n" + "".join("@SQ\tSN:%s\tLN:%d\tAS:SYN10M\n" % (n, len(b)) for n, b in contigs))
    return d, g, contigs


def run(calitas, *args):
    p = subprocess.run([calitas] + [str(a) for a in args], capture_output=True, text=True, timeout=600)
    return p


def lines(text):
    return [l for l in text.split("\n") if l]


def test_search_reference_cli(calitas, ref_dir):
    d, g, contigs = ref_dir
    out = d / "hits.tsv"
    p = run(calitas, "SearchReference", "-i", synth.BASELINE_GUIDE, "-I", "guide1", "-r", d / "ref.fa", "-o", out, "--time-stamp", "", "--aligner-version", "oracle", "--stats")
    assert p.returncode == 0, p.stderr
    exp = lines(pyoracle.search_reference(contigs, synth.BASELINE_GUIDE, guide_id="guide1", assembly="SYN10M", raw=True))
    assert len(exp) > 40 and lines(open(out).read()) == exp
    assert "hits" in p.stderr


def test_search_reference_cli_flags_and_vcf(calitas, ref_dir):
    d, g, contigs = ref_dir
    arrays = [np.frombuffer(b.upper(), dtype=np.uint8) for _, b in contigs]
    vcf = synth.synthetic_vcf(g, arrays, 600)
    vpath = d / "vars.vcf"
    open(vpath, "w").write(vcf)
    vid = "vars.vcf:" + hashlib.md5(vcf.encode()).hexdigest()
    out = d / "hits2.tsv"
    p = run(calitas, "SearchReference", "--guide=" + synth.BASELINE_GUIDE, "--guide-id", "g2", "-r", d / "ref.fa", "-v", vpath, "-o", out, "-O", "5", "-w", "500", "-V", "3", "-t", "4",
            "--time-stamp", "", "--aligner-version", "oracle")
    assert p.returncode == 0, p.stderr
    exp = lines(pyoracle.search_reference(contigs, synth.BASELINE_GUIDE, guide_id="g2", vcf_text=vcf, vcf_name=vid, assembly="SYN10M", raw=True, O=5, window_size=500, max_variants=3))
    got = lines(open(out).read())
    assert got == exp and any("+variants" in l and vid in l for l in got)
    # the same VCF gzip-compressed (two members, as bgzip writes): parsed alike, the id carries the MD5 of the compressed file
    import gzip
    half = vcf.index("\n", len(vcf) // 2) + 1
    gz = gzip.compress(vcf[:half].encode()) + gzip.compress(vcf[half:].encode())
    open(d / "vars.vcf.gz", "wb").write(gz)
    p = run(calitas, "SearchReference", "-i", synth.BASELINE_GUIDE, "-I", "g2", "-r", d / "ref.fa", "-v", d / "vars.vcf.gz", "-o", out, "-O", "5", "-w", "500", "-V", "3",
            "--time-stamp", "", "--aligner-version", "oracle")
    assert p.returncode == 0, p.stderr
    vid_gz = "vars.vcf.gz:" + hashlib.md5(gz).hexdigest()
    assert lines(open(out).read()) == [l.replace(vid, vid_gz) for l in exp]
    p = run(calitas, "SearchReference", "-i", "CTTGCCCCACAGGGCAGTAAngg", "-I", "g3", "-x", "nag", "nga", "-r", d / "ref.fa", "-o", out, "-d", "4", "-g", "2", "-p", "1", "-D", "5", "-c", "chr2",
            "-m", "-110", "-M=-250", "-b", "-125", "-B", "-119", "--time-stamp", "", "--aligner-version", "oracle")
    assert p.returncode == 0, p.stderr
    exp = lines(pyoracle.search_reference(contigs, "CTTGCCCCACAGGGCAGTAAngg", guide_id="g3", aux_pams=["nag", "nga"], chrom="chr2", assembly="SYN10M", raw=True, d=4, g=2, p=1, D=5,
                                          costs=(-110, -125, -119, -250)))
    assert lines(open(out).read()) == exp and len(exp) > 3


def test_prepare_vcf_then_search_reference(calitas, ref_dir):
    """The reference's two-step use: PrepareVcf (PrepareVcf.scala:43-91) on a raw call set -- genotypes, extra INFO keys, failing filters, rare and
    symbolic alleles -- then SearchReference -v on its output; expected = the oracle searching the oracle-prepared records."""
    d, g, contigs = ref_dir
    arrays = [np.frombuffer(b.upper(), dtype=np.uint8) for _, b in contigs]
    clean = [l for l in synth.synthetic_vcf(g, arrays, 300, seed=77).split("\n") if l and not l.startswith("#")]
    rng = np.random.default_rng(5)
    raw_rows = []
    for i, l in enumerate(clean):
        f = l.split("\t")
        k = int(rng.integers(10))
        if k == 0:
            f[6] = "LowQual"
        elif k == 1:
            f[4] += ",<DEL>"; f[7] += ",0.2"
        elif k == 2:
            f[7] = "AF=" + ",".join("0.004" for _ in f[4].split(","))
        elif k == 3 and "," not in f[4]:
            f[4] += "," + ("A" if f[4][0] != "A" else "C") + f[4][1:] + "T"; f[7] += ",0.003"     # second allele below the threshold: dropped, the first kept
        f[7] = "DP=%d;%s;DB" % (i, f[7])
        raw_rows.append("\t".join(f + ["GT", "0/1"]))
    head = "##fileformat=VCFv4.2\n##INFO=<ID=AF,Number=A,Type=Float,Description=\"af\">\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\ts1\n"
    raw_text = head + "\n".join(raw_rows) + "\n"
    open(d / "raw.vcf", "w").write(raw_text)
    prepared = d / "prepared.vcf.gz"
    p = run(calitas, "PrepareVcf", "-i", d / "raw.vcf", "-o", prepared, "-d", d / "ref.dict", "-c", "false")
    assert p.returncode == 0, p.stderr
    recs = pyoracle.prepare_vcf([raw_text], add_chr_prefix=False)
    assert 100 < len(recs) < len(clean)
    exp_vcf = "##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n" + "".join(
        "%s\t%s\t%s\t%s\t%s\t%s\tPASS\tAF=%s\n" % (c, pos, vid, ref, ",".join(alts), q, ",".join(afs)) for c, pos, vid, ref, alts, q, afs in recs)
    out = d / "hits_prepared.tsv"
    p = run(calitas, "SearchReference", "-i", synth.BASELINE_GUIDE, "-I", "g5", "-r", d / "ref.fa", "-v", prepared, "-o", out, "--time-stamp", "", "--aligner-version", "oracle")
    assert p.returncode == 0, p.stderr
    vid = "prepared.vcf.gz:" + hashlib.md5(open(prepared, "rb").read()).hexdigest()
    exp = lines(pyoracle.search_reference(contigs, synth.BASELINE_GUIDE, guide_id="g5", vcf_text=exp_vcf, vcf_name=vid, assembly="SYN10M", raw=True))
    got = lines(open(out).read())
    assert got == exp and any("+variants" in l for l in got)


def test_search_reference_cli_streams_many_blocks_in_order(calitas, ref_dir):
    """The table leaves the tool block by block while later blocks are rendered: with 3 rows per block and d=6 the run has dozens of blocks per
    guide; the file must still be the oracle's table, row for row, and the same through stdout."""
    d, g, contigs = ref_dir
    gf = d / "guides_stream.tsv"
    open(gf, "w").write("a\tCTTGCCCCACAGGGCAGTAA\nb\t%s\n" % synth.BASELINE_GUIDE)
    exp = None
    for gid, guide in (("a", "CTTGCCCCACAGGGCAGTAA"), ("b", synth.BASELINE_GUIDE)):
        t = lines(pyoracle.search_reference(contigs, guide, guide_id=gid, assembly="SYN10M", raw=True, d=6, g=2))
        exp = t if exp is None else exp + t[1:]
    assert len(exp) > 150
    env = dict(os.environ, CALITAS_ROW_BLOCK="3")
    out = d / "hits_stream.tsv"
    args = [calitas, "SearchReference", "--guides-file", str(gf), "-r", str(d / "ref.fa"), "-d", "6", "-g", "2", "--time-stamp", "", "--aligner-version", "oracle"]
    p = subprocess.run(args + ["-o", str(out)], capture_output=True, text=True, timeout=600, env=env)
    assert p.returncode == 0, p.stderr
    assert lines(open(out).read()) == exp
    p = subprocess.run(args, capture_output=True, text=True, timeout=600, env=env)
    assert p.returncode == 0 and lines(p.stdout) == exp
    # an unwritable output is an error, not a silent truncation
    p = subprocess.run(args + ["-o", str(d / "no_such_dir" / "x.tsv")], capture_output=True, text=True, timeout=600, env=env)
    assert p.returncode == 2 and "Cannot write" in p.stderr


def test_search_reference_cli_forty_guides_two_device_batches(calitas, ref_dir):
    """The streaming command line searches 32 guides per device call and renders one batch while a helper thread searches the next: 40 guides span
    two batches (guide indices restart in the second); the table must be the concatenation of the per-guide oracle tables."""
    d, g, contigs = ref_dir
    guides = [synth.BASELINE_GUIDE] + synth.random_guides(39)
    gf = d / "guides40.tsv"
    open(gf, "w").write("".join("id%d\t%s\n" % (i, s) for i, s in enumerate(guides)))
    out = d / "hits40.tsv"
    p = run(calitas, "SearchReference", "--guides-file", gf, "-r", d / "ref.fa", "-o", out, "--time-stamp", "", "--aligner-version", "oracle")
    assert p.returncode == 0, p.stderr
    exp = None
    for i, s in enumerate(guides):
        t = lines(pyoracle.search_reference(contigs, s, guide_id="id%d" % i, assembly="SYN10M", raw=True))
        exp = t if exp is None else exp + t[1:]
    got = lines(open(out).read())
    assert got == exp and len({l.split("\t")[0] for l in got[1:]}) >= 30


def test_search_reference_cli_guide_batch_and_stdout(calitas, ref_dir):
    d, g, contigs = ref_dir
    gf = d / "guides.tsv"
    open(gf, "w").write("# id\tguide\taux\na\t%s\nb\tCTTGCCCCACAGGGCAGTAAngg\tnag\nc\tGGGGCCACTAGGGACAGGAT\n" % synth.BASELINE_GUIDE)
    p = run(calitas, "SearchReference", "--guides-file", gf, "-r", d / "ref.fa", "--time-stamp", "", "--aligner-version", "oracle")
    assert p.returncode == 0, p.stderr
    exp = []
    for gid, seq, aux in (("a", synth.BASELINE_GUIDE, []), ("b", "CTTGCCCCACAGGGCAGTAAngg", ["nag"]), ("c", "GGGGCCACTAGGGACAGGAT", [])):
        l = lines(pyoracle.search_reference(contigs, seq, guide_id=gid, aux_pams=aux, assembly="SYN10M", raw=True))
        exp += l if not exp else l[1:]
    assert lines(p.stdout) == exp


def test_search_reference_cli_sharded_over_devices(calitas, ref_dir):
    """--devices a,b,c: one engine per device, contig-range shards driven from host threads; output identical to the single-device table."""
    d, g, contigs = ref_dir
    if calitas.endswith("hostsim"):
        devices = "0,0,0"                                  # the host simulation ignores device ids; three shards all the same
    else:
        import torch
        n = min(3, torch.cuda.device_count())
        if n < 2:
            pytest.skip("needs at least 2 GPUs")
        devices = ",".join(str(i) for i in range(n))
    gf = d / "guides2.tsv"
    open(gf, "w").write("a\t%s\nb\tCTTGCCCCACAGGGCAGTAAngg\tnag\n" % synth.BASELINE_GUIDE)
    outs = []
    for dev in ("0", devices):
        p = run(calitas, "SearchReference", "--guides-file", gf, "-r", d / "ref.fa", "--devices", dev, "--time-stamp", "", "--aligner-version", "oracle")
        assert p.returncode == 0, p.stderr
        outs.append(p.stdout)
    assert outs[0] == outs[1] and len(lines(outs[0])) > 80


def test_align_to_reference_cli(calitas, ref_dir):
    d, g, contigs = ref_dir
    guides = [synth.BASELINE_GUIDE] + synth.random_guides(2)
    tasks = synth.a2r_tasks(g, guides, 60)
    tasks[5] = (tasks[5][0], tasks[5][1], "chr1", 30200)                  # inside the soft-masked stretch
    tp = d / "tasks.tsv"
    open(tp, "w").write("id\tquery\tchrom\tposition\n" + "".join("%s\t%s\t%s\t%d\n" % t for t in tasks))
    for flags, kw in ((["-w", "60"], dict(window_size=60)), (["-w", "60", "-d", "5", "-p", "1", "-O", "10"], dict(window_size=60, d=5, p=1, O=10)), ([], dict())):
        out = d / "a2r.tsv"
        p = run(calitas, "AlignToReference", "-i", tp, "-r", d / "ref.fa", "-o", out, "--time-stamp", "", "--aligner-version", "oracle", *flags)
        assert p.returncode == 0, p.stderr
        exp = lines(pyoracle.align_to_reference(contigs, tasks, assembly="SYN10M", raw=True, **kw))
        assert lines(open(out).read()) == exp, flags


def test_cli_errors_mirror_the_reference(calitas, ref_dir, tmp_path):
    d, g, contigs = ref_dir
    p = run(calitas, "AlignToReference", "-i", d / "nope.tsv", "-r", d / "ref.fa")
    assert p.returncode != 0 and "non-existent" in p.stderr
    tp = tmp_path / "t.tsv"
    open(tp, "w").write("query\tchrom\tposition\nCTTGCCCCACAGGGCAGTAAnrg\tchr1\t5000\n")
    p = run(calitas, "AlignToReference", "-i", tp, "-r", d / "ref.fa", "-d", "3")
    assert p.returncode != 0 and "Must specify all or none of" in p.stderr                        # AlignToReference.scala:88-92
    p = run(calitas, "SearchReference", "-i", "cttgccccacagggcagtaa", "-I", "x", "-r", d / "ref.fa")
    assert p.returncode != 0 and "cannot be all lower case" in p.stderr                           # SequentialGuideAligner.scala:84-87
    p = run(calitas, "SearchReference", "-i", synth.BASELINE_GUIDE, "-I", "x", "-r", d / "ref.fa", "-c", "chrZ")
    assert p.returncode != 0 and "Unknown chromosome" in p.stderr
    bare = tmp_path / "bare.fa"
    open(bare, "w").write(">c\nACGTACGTACGTACGTACGTACGTACGTACGT\n")
    p = run(calitas, "SearchReference", "-i", synth.BASELINE_GUIDE, "-I", "x", "-r", bare)
    assert p.returncode != 0 and "sequence dictionary" in p.stderr                                # SearchReference.scala:478-484
    p = run(calitas, "SearchReference", "-I", "x", "-r", d / "ref.fa")
    assert p.returncode != 0 and "required" in p.stderr


def test_pairwise_align_sequences_cli(calitas, tmp_path):
    """PairwiseAlignSequences.scala:44-83: `query target` lines -> alignBest -> 11 columns."""
    rng = np.random.default_rng(11)
    pairs = [("AACCGGTTAACCGGTTAACC", "TTAACCGGGTTAACCGGTTAACCTT"), ("AACCGGTTAACCGGTTAACC", "ttaaccgttaaccggttaacctt"), ("CTTGCCCCACAGGGCAGTAAnrg", "GGCTTGCCCCACAGGGCAGTAACGGTT"),
             ("tttvCTTGCCCCACAGGGCAGTAA", "AATTTACTTGCCCCACTGGGCAGTAAGG")]
    for _ in range(40):
        q = "".join(rng.choice(list("ACGT"), size=int(rng.integers(8, 24))))
        t = list(rng.choice(list("ACGT"), size=int(rng.integers(30, 80))))
        p = int(rng.integers(0, len(t) - len(q) + 1))
        t[p:p + len(q)] = list(synth.mutate_protospacer(rng, q.encode(), int(rng.integers(0, 4))).decode())[:len(q)]
        pairs.append((q + ("ngg" if rng.random() < 0.5 else ""), "".join(t)))
    inp = tmp_path / "pairs.txt"
    open(inp, "w").write("\n".join("%s  %s" % p for p in pairs) + "\n\n")
    out = tmp_path / "pairs.tsv"
    p = run(calitas, "PairwiseAlignSequences", "-i", inp, "-o", out, "-t", "2")
    assert p.returncode == 0, p.stderr
    exp = ["query\ttarget\tscore\tquery_start\ttarget_start\tcigar\tmismatches\tgap_bases\tpadded_query\talignment\tpadded_target"]
    for q, t in pairs:
        a = pyoracle.align_best(q, t.upper())
        exp.append("\t".join(str(x) for x in (q, t.upper(), a["score"], 1, a["startOffset"], a["cigar"], a["mismatches"], a["gapBases"], a["paddedGuide"], a["paddedAlignment"], a["paddedTarget"])))
    assert lines(open(out).read()) == exp
    open(inp, "w").write("AAA CCC GGG\n")
    p = run(calitas, "PairwiseAlignSequences", "-i", inp)
    assert p.returncode != 0 and "Line found with 3 fields" in p.stderr
