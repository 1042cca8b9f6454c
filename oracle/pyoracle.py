"""ctypes binding for the CPU oracle (oracle/calitas_oracle.cpp).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never from calitas_b200/.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libcalitas_oracle.so")

DEFAULT_COSTS = (-120, -122, -121, -260)  # mismatch, genomeGap, guideGap, pamMismatch (SequentialGuideAligner.scala:17-21)


def build(force=False):
    src = os.path.join(_HERE, "calitas_oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        L.oracle_last_error.restype = C.c_char_p
        L.oracle_free.argtypes = [C.c_void_p]
        for name in ("oracle_align", "oracle_align_best", "oracle_align_to_ref", "oracle_fg_align", "oracle_guide_alignment",
                     "oracle_allele_combos", "oracle_variant_sets", "oracle_build_variant_window", "oracle_search_reference",
                     "oracle_align_to_reference"):
            getattr(L, name).restype = C.c_void_p
        L.oracle_search_reference_count.restype = C.c_int64
        _lib = L
    return _lib


class OracleError(Exception):
    pass


def _take(ptr):
    if not ptr:
        raise OracleError(lib().oracle_last_error().decode())
    s = C.string_at(ptr).decode("latin-1")
    lib().oracle_free(ptr)
    return s


def _b(s):
    return s.encode("latin-1") if isinstance(s, str) else bytes(s)


def _strarr(xs):
    arr = (C.c_char_p * max(1, len(xs)))()
    for i, x in enumerate(xs):
        arr[i] = _b(x)
    return arr


def _costs(costs):
    return (C.c_int * 4)(*costs)


def _table(text):
    lines = [l for l in text.split("\n") if l != ""]
    if not lines:
        return []
    hdr = lines[0].split("\t")
    return [dict(zip(hdr, l.split("\t"))) for l in lines[1:]]


_INT_GA = ("startOffset", "endOffset", "guideStartOffset", "guideEndOffset", "score", "mismatches", "gapBases", "edits", "guideMismatches",
           "guideGapBases", "guideMmsPlusGaps", "pamMismatches", "pamGapBases", "pamMmsPlusGaps")


def _ga(text):
    rows = _table(text)
    for r in rows:
        for k in _INT_GA:
            r[k] = int(r[k])
    return rows


def align(guide, target, aux_pams=(), target_name="n/a", target_offset=0, *, max_guide_diffs, max_gaps, max_pam_diffs, max_total_diffs,
          max_overlap=0, costs=DEFAULT_COSTS):
    t = _b(target)
    return _ga(_take(lib().oracle_align(_b(guide), _strarr(aux_pams), len(aux_pams), t, len(t), _b(target_name), target_offset,
                                        max_guide_diffs, max_gaps, max_pam_diffs, max_total_diffs, max_overlap, _costs(costs))))


def align_best(guide, target, aux_pams=(), max_gaps=3, costs=DEFAULT_COSTS):
    t = _b(target)
    return _ga(_take(lib().oracle_align_best(_b(guide), _strarr(aux_pams), len(aux_pams), t, len(t), max_gaps, _costs(costs))))[0]


def _contigs(contigs):
    names = _strarr([c[0] for c in contigs])
    bufs = [_b(c[1]) for c in contigs]
    lens = (C.c_int64 * max(1, len(contigs)))(*[len(b) for b in bufs])
    ptrs = (C.c_char_p * max(1, len(contigs)))(*bufs)
    return names, lens, ptrs, bufs


def align_to_ref(contigs, guide, chrom, pos, window_size=None, best=True, max_guide_diffs=0, max_gaps=3, max_pam_diffs=0, max_total_diffs=0,
                 max_overlap=0, costs=DEFAULT_COSTS):
    names, lens, ptrs, _keep = _contigs(contigs)
    rows = _ga(_take(lib().oracle_align_to_ref(len(contigs), names, lens, ptrs, _b(guide), _b(chrom), pos, -1 if window_size is None else window_size,
                                               1 if best else 0, max_guide_diffs, max_gaps, max_pam_diffs, max_total_diffs, max_overlap, _costs(costs))))
    return rows[0] if best else rows


def fg_align(query, target, min_score, costs=DEFAULT_COSTS):
    t = _b(target)
    out = []
    for l in _take(lib().oracle_fg_align(_b(query), t, len(t), min_score, _costs(costs))).split("\n"):
        if l:
            a, b, c, d = l.split("\t")
            out.append((int(a), int(b), int(c), d))
    return out


def guide_alignment(padded_guide, padded_align, padded_target, start, end, strand):
    return _ga(_take(lib().oracle_guide_alignment(_b(padded_guide), _b(padded_align), _b(padded_target), start, end, C.c_char(_b(strand)))))[0]


def allele_combos(counts):
    arr = (C.c_int * len(counts))(*counts)
    return [[int(x) for x in l.split(",")] for l in _take(lib().oracle_allele_combos(arr, len(counts))).split("\n") if l]


def variant_sets(vcf_text, max_variants):
    return [tuple(tuple(x.split(":")) for x in l.split(",")) for l in _take(lib().oracle_variant_sets(_b(vcf_text), max_variants)).split("\n") if l]


def build_variant_window(chrom, ref_bases, vcf_text, padding, queries=()):
    offs = (C.c_int * max(1, len(queries)))(*[q[0] for q in queries])
    prec = (C.c_int * max(1, len(queries)))(*[1 if q[1] else 0 for q in queries])
    rb = _b(ref_bases)
    lines = _take(lib().oracle_build_variant_window(_b(chrom), rb, C.c_int64(len(rb)), _b(vcf_text), padding, offs, prec, len(queries))).split("\n")
    bases, cigar, start = lines[0].split("\t")
    return {"bases": bases, "cigar": cigar, "start": int(start), "offsets": [int(x) for x in lines[1:1 + len(queries)]]}


def search_params(max_variants=16, window_size=1000, d=5, p=1, g=3, D=None, O=10, costs=DEFAULT_COSTS, threads=1, stage=0):
    mm, genome_gap, guide_gap, pam_mm = costs
    return (C.c_int * 13)(max_variants, window_size, d, p, g, -1 if D is None else D, O, mm, pam_mm, genome_gap, guide_gap, threads, stage)


_INT_HIT = ("coordinate_start", "coordinate_end", "score", "guide_mm", "guide_gaps", "guide_mm_plus_gaps", "pam_mm", "total_mm_plus_gaps",
            "unpadded_guide_sequence_length", "unpadded_target_sequence_length")


def hits_table(text):
    rows = _table(text)
    for r in rows:
        for k in _INT_HIT:
            r[k] = int(r[k])
    return rows


def search_reference(contigs, guide, guide_id="g", aux_pams=(), chrom=None, vcf_text=None, vcf_name="variants.vcf", assembly=None, raw=False, **kw):
    """SearchReference.execute on an in-memory genome: contigs = [(name, bases)], returns the hit table rows."""
    names, lens, ptrs, _keep = _contigs(contigs)
    nwin = C.c_int64(0)
    nhits = C.c_int64(0)
    text = _take(lib().oracle_search_reference(len(contigs), names, lens, ptrs, _b(assembly) if assembly else None, _b(guide), _b(guide_id),
                                               _strarr(aux_pams), len(aux_pams), _b(chrom) if chrom else None,
                                               _b(vcf_text) if vcf_text is not None else None, _b(vcf_name), search_params(**kw),
                                               C.byref(nwin), C.byref(nhits)))
    return text if raw else hits_table(text)


def search_reference_count(contigs, guide, aux_pams=(), **kw):
    names, lens, ptrs, _keep = _contigs(contigs)
    nwin = C.c_int64(0)
    n = lib().oracle_search_reference_count(len(contigs), names, lens, ptrs, _b(guide), _strarr(aux_pams), len(aux_pams), search_params(**kw), C.byref(nwin))
    if n < 0:
        raise OracleError(lib().oracle_last_error().decode())
    return n, nwin.value


def align_to_reference(contigs, tasks, window_size=None, d=None, p=None, g=3, D=None, O=None, costs=DEFAULT_COSTS, threads=1, assembly=None, raw=False):
    """AlignToReference.execute: tasks = [(id, query, chrom, position)]."""
    names, lens, ptrs, _keep = _contigs(contigs)
    mm, genome_gap, guide_gap, pam_mm = costs
    opt = lambda v: -1 if v is None else v
    ip = (C.c_int * 11)(opt(window_size), opt(d), opt(p), g, opt(D), opt(O), mm, pam_mm, genome_gap, guide_gap, threads)
    pos = (C.c_int * max(1, len(tasks)))(*[t[3] for t in tasks])
    text = _take(lib().oracle_align_to_reference(len(contigs), names, lens, ptrs, _b(assembly) if assembly else None, len(tasks),
                                                 _strarr([t[0] for t in tasks]), _strarr([t[1] for t in tasks]), _strarr([t[2] for t in tasks]), pos, ip))
    return text if raw else hits_table(text)


# ---------------------------------------------------------------------------------------------------------------------------------------
# PrepareVcf (PrepareVcf.scala:43-91) — pure-Python restatement; the records it keeps and rewrites, as (CHROM, POS, ID, REF, [ALT], QUAL, [AF]).
# Parity unpinned beyond the reference's one test (PrepareVcfTest.scala:10-43: samples removed, 10 of 10 records kept): the number
# formatting is htsjdk's VCFEncoder.formatVCFDouble / formatQualValue restated from its published source, not run here.
# ---------------------------------------------------------------------------------------------------------------------------------------
_CHROMS_TO_FIX = {str(i) for i in range(1, 23)} | {"X", "Y"}          # PrepareVcf.scala:13-15


def _java_round(value, places):
    """java.util.Formatter rounds HALF_UP on the shortest decimal form of the double."""
    from decimal import Decimal, ROUND_HALF_UP
    return Decimal(repr(float(value))).quantize(Decimal(1).scaleb(-places), rounding=ROUND_HALF_UP)


def format_vcf_double(d):
    """htsjdk VCFEncoder.formatVCFDouble."""
    from decimal import Decimal, ROUND_HALF_UP
    if d < 1:
        if d < 0.01:
            if abs(d) >= 1e-20:
                dec = Decimal(repr(float(d)))
                exp = dec.adjusted()
                mant = (dec.scaleb(-exp)).quantize(Decimal("0.001"), rounding=ROUND_HALF_UP)
                if abs(mant) >= 10:
                    mant = (mant / 10).quantize(Decimal("0.001"))
                    exp += 1
                return "%se%s%02d" % (mant, "-" if exp < 0 else "+", abs(exp))
            return "0.00"
        return str(_java_round(d, 3))
    return str(_java_round(d, 2))


def format_qual(q):
    """htsjdk VCFEncoder.formatQualValue."""
    s = str(_java_round(q, 2))
    return s[:-3] if s.endswith(".00") else s


def prepare_vcf(texts, min_af=0.01, add_chr_prefix=True):
    """texts: the input VCFs' contents in order.  Returns (header_has_samples_removed_chrom_line, records)."""
    import numpy as np
    records = []
    for text in texts:
        for line in text.split("\n"):
            line = line.rstrip("\r")
            if not line or line.startswith("#"):
                continue
            f = line.split("\t")
            if f[6] != "PASS":                                                          # :73
                continue
            af = None
            for kv in f[7].split(";"):
                if kv.startswith("AF="):
                    af = [float("nan") if x == "." else float(np.float32(x)) for x in kv[3:].split(",")]   # ArrayAttr[Float]
            if af is None:
                raise KeyError("AF")                                                    # :74 apply() on a missing key
            if not any(x >= min_af for x in af):                                        # :74
                continue
            alts = f[4].split(",")
            simple = lambda a: len(a) > 0 and all(c in "ACGTNacgtn" for c in a)         # fgbio SimpleAllele
            if not (simple(f[3]) and all(simple(a) for a in alts)):                     # :75
                continue
            kept = [(a, x) for a, x in zip(alts, af) if x >= min_af]                    # :77
            chrom = "chr" + f[0] if add_chr_prefix and f[0] in _CHROMS_TO_FIX else f[0]  # :79,91
            qual = f[5] if f[5] == "." else format_qual(float(f[5]))
            records.append((chrom, f[1], f[2], f[3], [a for a, _ in kept], qual, [format_vcf_double(x) for _, x in kept]))
    return records
