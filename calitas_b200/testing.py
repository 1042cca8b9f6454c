"""Dict-row facade over the tool-level C ABI (include/calitas_b200_tools.h), shaped like the CPU oracle binding so the parity
tests can run the same reference-test bodies against both.  Uses the product library by default."""
import ctypes as C

from . import _capi
from ._capi import Engine, Limits, SearchOptions, A2ROptions, A2RTask, make_guide, _b

_INT_GA = ("startOffset", "endOffset", "guideStartOffset", "guideEndOffset", "score", "mismatches", "gapBases", "edits", "guideMismatches",
           "guideGapBases", "guideMmsPlusGaps", "pamMismatches", "pamGapBases", "pamMmsPlusGaps")
_INT_HIT = ("coordinate_start", "coordinate_end", "score", "guide_mm", "guide_gaps", "guide_mm_plus_gaps", "pam_mm", "total_mm_plus_gaps",
            "unpadded_guide_sequence_length", "unpadded_target_sequence_length")


def _table(text, int_cols):
    lines = [l for l in text.split("\n") if l != ""]
    if not lines:
        return []
    hdr = lines[0].split("\t")
    rows = [dict(zip(hdr, l.split("\t"))) for l in lines[1:]]
    for r in rows:
        for k in int_cols:
            r[k] = int(r[k])
    return rows


class Facade:
    def __init__(self, lib=None, device=0):
        self.lib = lib or _capi.default_library()
        self.device = device
        self._engines = {}

    def engine(self, costs):
        costs = tuple(costs)
        if costs not in self._engines:
            self._engines[costs] = Engine(self.device, costs, self.lib)
        return self._engines[costs]

    def _text(self, ptr):
        return self.lib.take_text(ptr)

    def align(self, guide, target, aux_pams=(), target_name="n/a", target_offset=0, *, max_guide_diffs, max_gaps, max_pam_diffs, max_total_diffs,
              max_overlap=0, costs=_capi.DEFAULT_COSTS):
        e = self.engine(costs)
        g, keep = make_guide(guide, aux_pams)
        t = _b(target)
        lim = Limits(max_guide_diffs, max_pam_diffs, max_gaps, max_total_diffs, max_overlap)
        out = C.c_void_p()
        self.lib.check(self.lib.L.calitas_tool_align(e.ptr, C.byref(g), t, len(t), _b(target_name), target_offset, C.byref(lim), C.byref(out)))
        return _table(self._text(out), _INT_GA)

    def align_best(self, guide, target, aux_pams=(), max_gaps=3, costs=_capi.DEFAULT_COSTS):
        e = self.engine(costs)
        g, keep = make_guide(guide, aux_pams)
        t = _b(target)
        out = C.c_void_p()
        self.lib.check(self.lib.L.calitas_tool_align_best(e.ptr, C.byref(g), t, len(t), max_gaps, C.byref(out)))
        return _table(self._text(out), _INT_GA)[0]

    def _ref(self, e, contigs):
        key = (id(e), tuple((c[0], len(c[1]), c[1].ctypes.data if hasattr(c[1], "ctypes") else hash(c[1])) for c in contigs))
        cache = self.__dict__.setdefault("_refs", {})
        if key not in cache:
            if len(cache) > 8:
                for r in cache.values():
                    r[0].free()
                cache.clear()
            cache[key] = (e.load_reference(contigs), Engine.genome_view(contigs))
        return cache[key]

    def align_to_ref(self, contigs, guide, chrom, pos, window_size=None, best=True, max_guide_diffs=0, max_gaps=3, max_pam_diffs=0, max_total_diffs=0,
                     max_overlap=0, costs=_capi.DEFAULT_COSTS):
        e = self.engine(costs)
        ref, (view, keep) = self._ref(e, contigs)
        g, gkeep = make_guide(guide)
        lim = Limits(max_guide_diffs, max_pam_diffs, max_gaps, max_total_diffs, max_overlap)
        out = C.c_void_p()
        self.lib.check(self.lib.L.calitas_tool_align_to_ref(e.ptr, ref.ptr, C.byref(view), C.byref(g), _b(chrom), pos, -1 if window_size is None else window_size,
                                                            1 if best else 0, C.byref(lim), C.byref(out)))
        rows = _table(self._text(out), _INT_GA)
        return rows[0] if best else rows

    def search_reference(self, contigs, guide, guide_id="g", aux_pams=(), chrom=None, vcf_text=None, vcf_name="variants.vcf", assembly=None, raw=False,
                         max_variants=16, window_size=1000, d=5, p=1, g=3, D=None, O=10, costs=_capi.DEFAULT_COSTS, threads=1, stage=0):
        e = self.engine(costs)
        ref, _ = self._ref(e, contigs)
        view, keep = Engine.genome_view(contigs, assembly)
        gd, gkeep = make_guide(guide, aux_pams)
        opt = SearchOptions(_b(guide_id), max_variants, window_size, Limits(d, p, g, -1 if D is None else D, O), _b(chrom), _b(vcf_text), _b(vcf_name), b"", b"oracle")
        out = C.c_void_p()
        n = C.c_int64(0)
        self.lib.check(self.lib.L.calitas_tool_search_reference(e.ptr, ref.ptr, C.byref(view), C.byref(gd), C.byref(opt), C.byref(out), C.byref(n)))
        text = self._text(out)
        return text if raw else _table(text, _INT_HIT)

    def search_reference_batch(self, contigs, guides, guide_ids=None, n_shards=1, chrom=None, vcf_text=None, vcf_name="variants.vcf", assembly=None,
                               max_variants=16, window_size=1000, d=5, p=1, g=3, D=None, O=10, costs=_capi.DEFAULT_COSTS):
        """calitas_tool_search_reference_batch: guides = [str | (str, [aux pams])]; n_shards engines, each over its contig-range shard (here all on
        this facade's device; in production one per GPU).  Returns the TSV text."""
        from ._capi import make_guides
        engines = [self.engine(costs)] + [Engine(self.device, costs, self.lib) for _ in range(n_shards - 1)]
        refs = [e.load_reference(contigs, shard=None if n_shards == 1 else (s, n_shards, 4 * window_size)) for s, e in enumerate(engines)]
        try:
            view, keep = Engine.genome_view(contigs, assembly)
            arr, gkeep = make_guides(guides)
            ids = [_b(x) for x in (guide_ids or ["g%d" % i for i in range(len(guides))])]
            id_arr = (C.c_char_p * len(ids))(*ids)
            opt = SearchOptions(None, max_variants, window_size, Limits(d, p, g, -1 if D is None else D, O), _b(chrom), _b(vcf_text), _b(vcf_name), b"", b"oracle")
            e_arr = (C.c_void_p * n_shards)(*[e.ptr for e in engines])
            r_arr = (C.c_void_p * n_shards)(*[r.ptr for r in refs])
            out = C.c_void_p()
            n = C.c_int64(0)
            self.lib.check(self.lib.L.calitas_tool_search_reference_batch(n_shards, e_arr, r_arr, C.byref(view), len(guides), arr, id_arr, C.byref(opt), C.byref(out), C.byref(n)))
            return self._text(out)
        finally:
            for r in refs:
                r.free()
            for e in engines[1:]:
                e.close()

    def align_to_reference(self, contigs, tasks, window_size=None, d=None, p=None, g=3, D=None, O=None, costs=_capi.DEFAULT_COSTS, threads=1, assembly=None, raw=False):
        e = self.engine(costs)
        ref, _ = self._ref(e, contigs)
        view, keep = Engine.genome_view(contigs, assembly)
        opt_ = lambda v: -1 if v is None else v
        opt = A2ROptions(opt_(window_size), opt_(d), opt_(p), g, opt_(D), opt_(O), b"", b"oracle")
        arr = (A2RTask * max(1, len(tasks)))()
        keepalive = []
        for i, (tid, q, c, pos) in enumerate(tasks):
            vals = (_b(tid), _b(q), _b(c))
            keepalive.append(vals)
            arr[i] = A2RTask(vals[0], vals[1], vals[2], pos)
        out = C.c_void_p()
        n = C.c_int64(0)
        self.lib.check(self.lib.L.calitas_tool_align_to_reference(e.ptr, ref.ptr, C.byref(view), C.c_int64(len(tasks)), arr, C.byref(opt), C.byref(out), C.byref(n)))
        text = self._text(out)
        return text if raw else _table(text, _INT_HIT)

    def variant_windows(self, contigs, vcf_text, padding, max_variants=16, chrom=None):
        view, keep = Engine.genome_view(contigs)
        out = C.c_void_p()
        self.lib.check(self.lib.L.calitas_tool_variant_windows(C.byref(view), _b(vcf_text), _b(chrom), padding, max_variants, C.byref(out)))
        return [l.split("\t") for l in self._text(out).split("\n") if l]


_default = None


def _facade():
    global _default
    if _default is None:
        _default = Facade()
    return _default


def align(*a, **k):
    return _facade().align(*a, **k)


def align_best(*a, **k):
    return _facade().align_best(*a, **k)


def align_to_ref(*a, **k):
    return _facade().align_to_ref(*a, **k)


def search_reference(*a, **k):
    return _facade().search_reference(*a, **k)


def align_to_reference(*a, **k):
    return _facade().align_to_reference(*a, **k)
