"""ctypes binding of include/calitas_b200.h and include/calitas_b200_tools.h."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
PRODUCT_LIB = os.path.join(_HERE, "libcalitas_b200.so")

DEFAULT_COSTS = (-120, -122, -121, -260)  # mismatch, genome gap, guide gap, PAM mismatch (SequentialGuideAligner.scala:17-21)


class CalitasError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"calitas_b200 error {code}: {msg}")
        self.code = code
        self.message = msg


class Costs(C.Structure):
    _fields_ = [("mismatch_net_cost", C.c_int32), ("genome_gap_net_cost", C.c_int32), ("guide_gap_net_cost", C.c_int32), ("pam_mismatch_net_cost", C.c_int32)]


class Limits(C.Structure):
    _fields_ = [("max_guide_diffs", C.c_int32), ("max_pam_mismatches", C.c_int32), ("max_gaps_between_guide_and_pam", C.c_int32),
                ("max_total_diffs", C.c_int32), ("max_overlap", C.c_int32)]


class Guide(C.Structure):
    _fields_ = [("sequence", C.c_char_p), ("aux_pams", C.POINTER(C.c_char_p)), ("n_aux_pams", C.c_int32)]


class Hit(C.Structure):
    """calitas_hit: 32 bytes, 48 alignment columns (include/calitas_b200.h)"""
    _fields_ = [("start_offset", C.c_int32), ("task_idx", C.c_int32), ("score", C.c_int32), ("where", C.c_uint32), ("shape", C.c_uint32), ("ops", C.c_uint32 * 3)]


class HitWide(C.Structure):
    """calitas_hit_wide: the same header, 176 alignment columns"""
    _fields_ = [("start_offset", C.c_int32), ("task_idx", C.c_int32), ("score", C.c_int32), ("where", C.c_uint32), ("shape", C.c_uint32), ("ops", C.c_uint32 * 11)]


HIT_WORDS, HIT_WIDE_WORDS, HIT_HEADER_WORDS = 8, 16, 5        # include/calitas_b200.h: calitas_hit = 32 bytes, calitas_hit_wide = 64 bytes, both with a 5-word header
MAX_OP_WORDS = HIT_WIDE_WORDS - HIT_HEADER_WORDS


def hit_dtype():
    """numpy structured dtype of a DECODED hit: every logical field of a calitas_hit / calitas_hit_wide spelled out (see decode_hits)."""
    import numpy as np
    return np.dtype([("guide_idx", "<i4"), ("pam_idx", "<i4"), ("contig_idx", "<i4"), ("task_idx", "<i4"), ("start_offset", "<i4"), ("end_offset", "<i4"),
                     ("guide_start_offset", "<i4"), ("guide_end_offset", "<i4"), ("score", "<i4"), ("strand", "u1"), ("n_ops", "u1"), ("gap_bases", "u1"),
                     ("edits", "u1"), ("ops", "<u4", (MAX_OP_WORDS,))])


def decode_hits(raw_words):
    """raw_words: uint32 array of shape (n, 8) or (n, 16), the packed records of a hit set -> structured array of hit_dtype() (the accessors of
    include/calitas_b200.h, vectorised)."""
    import numpy as np
    w = np.asarray(raw_words, dtype=np.uint32)
    n, rw = w.shape
    out = np.zeros(n, dtype=hit_dtype())
    where, shape = w[:, 3], w[:, 4]
    out["start_offset"] = w[:, 0].view(np.int32)
    out["task_idx"] = w[:, 1].view(np.int32)
    out["score"] = w[:, 2].view(np.int32)
    out["guide_idx"] = where & 0x1FFF
    out["contig_idx"] = ((where >> 13) & 0x3FFFF).astype(np.int32) - 1
    out["strand"] = np.where(where >> 31, ord("-"), ord("+"))
    out["n_ops"] = shape & 0xFF
    out["end_offset"] = out["start_offset"] + ((shape >> 8) & 0xFF).astype(np.int32)
    out["guide_start_offset"] = out["start_offset"] + ((shape >> 16) & 0x3F).astype(np.int32)
    out["guide_end_offset"] = out["end_offset"] - ((shape >> 22) & 0x3F).astype(np.int32)
    out["pam_idx"] = (shape >> 28).astype(np.int32) - 1
    ops = w[:, HIT_HEADER_WORDS:]
    out["ops"][:, :rw - HIT_HEADER_WORDS] = ops
    bits = np.unpackbits(ops.view(np.uint8), axis=1) if n else np.zeros((0, 0), dtype=np.uint8)
    if n:
        hi = (ops & np.uint32(0xAAAAAAAA))
        anyb = ((ops | (ops >> 1)) & np.uint32(0x55555555))
        out["gap_bases"] = np.unpackbits(hi.view(np.uint8), axis=1).sum(axis=1)
        out["edits"] = np.unpackbits(anyb.view(np.uint8), axis=1).sum(axis=1)
    del bits
    return out


def encode_hits(records):
    """Decoded hits (hit_dtype()) -> packed calitas_hit_wide records, uint32 array of shape (n, 16) (what calitas_render_alignments takes)."""
    import numpy as np
    r = np.asarray(records, dtype=hit_dtype()).reshape(-1)
    w = np.zeros((r.size, HIT_WIDE_WORDS), dtype=np.uint32)
    w[:, 0] = r["start_offset"].view(np.uint32)
    w[:, 1] = r["task_idx"].view(np.uint32)
    w[:, 2] = r["score"].view(np.uint32)
    w[:, 3] = r["guide_idx"].astype(np.uint32) | ((r["contig_idx"] + 1).astype(np.uint32) << 13) | ((r["strand"] == ord("-")).astype(np.uint32) << 31)
    w[:, 4] = (r["n_ops"].astype(np.uint32) | ((r["end_offset"] - r["start_offset"]).astype(np.uint32) << 8) | ((r["guide_start_offset"] - r["start_offset"]).astype(np.uint32) << 16)
               | ((r["end_offset"] - r["guide_end_offset"]).astype(np.uint32) << 22) | ((r["pam_idx"] + 1).astype(np.uint32) << 28))
    w[:, HIT_HEADER_WORDS:] = r["ops"]
    return w


class RegionTask(C.Structure):
    _fields_ = [("guide_idx", C.c_int32), ("contig_idx", C.c_int32), ("start", C.c_int64), ("length", C.c_int32)]


class TargetTask(C.Structure):
    _fields_ = [("guide_idx", C.c_int32), ("bases", C.c_char_p), ("length", C.c_int32), ("target_offset", C.c_int32)]


class GenomeView(C.Structure):
    _fields_ = [("n_contigs", C.c_int32), ("names", C.POINTER(C.c_char_p)), ("lengths", C.POINTER(C.c_int64)), ("bases", C.POINTER(C.c_char_p)), ("assembly", C.c_char_p)]


class SearchOptions(C.Structure):
    _fields_ = [("guide_id", C.c_char_p), ("max_variants", C.c_int32), ("window_size", C.c_int32), ("limits", Limits), ("chrom", C.c_char_p),
                ("vcf_text", C.c_char_p), ("vcf_id", C.c_char_p), ("time_stamp", C.c_char_p), ("aligner_version", C.c_char_p)]


class A2RTask(C.Structure):
    _fields_ = [("id", C.c_char_p), ("query", C.c_char_p), ("chrom", C.c_char_p), ("position", C.c_int32)]


class A2ROptions(C.Structure):
    _fields_ = [("window_size", C.c_int32), ("max_guide_diffs", C.c_int32), ("max_pam_mismatches", C.c_int32), ("max_gaps_between_guide_and_pam", C.c_int32),
                ("max_total_diffs", C.c_int32), ("max_overlap", C.c_int32), ("time_stamp", C.c_char_p), ("aligner_version", C.c_char_p)]


EXPORTS = [
    # include/calitas_b200.h
    "calitas_engine_create", "calitas_engine_destroy", "calitas_engine_get_costs", "calitas_last_error", "calitas_reference_load", "calitas_reference_free", "calitas_reference_own_range",
    "calitas_shard_plan", "calitas_search", "calitas_search_sharded", "calitas_search_variants", "calitas_variant_set_load", "calitas_variant_set_free", "calitas_hitset_variant_info", "calitas_align_regions", "calitas_align_targets", "calitas_hitset_count", "calitas_hitset_data", "calitas_hitset_stride",
    "calitas_hitset_free", "calitas_hitset_stats", "calitas_render_alignments", "calitas_free_text", "calitas_microbench_int",
    # include/calitas_b200_tools.h
    "calitas_tool_align", "calitas_tool_align_best", "calitas_tool_align_to_ref", "calitas_tool_search_reference", "calitas_tool_search_reference_batch", "calitas_tool_search_reference_batch_fd", "calitas_tool_align_to_reference",
    "calitas_tool_variant_windows", "calitas_tool_pairwise_align", "calitas_tool_variant_plan_create", "calitas_tool_variant_plan_free", "calitas_tool_variant_plan_counts",
    "calitas_tool_variant_plan_search",
]


def _b(s):
    if s is None:
        return None
    return s.encode("latin-1") if isinstance(s, str) else bytes(s)


class Library:
    """One loaded shared library implementing the C ABI (the product .so; tests may load the host simulation explicitly)."""

    def __init__(self, path=PRODUCT_LIB):
        if not os.path.exists(path):
            raise CalitasError(-1, f"{path} is not built: run `python -c 'import __graft_entry__ as g; g.build()'` (no CPU fallback exists)")
        self.path = path
        L = self.L = C.CDLL(path)
        for name in EXPORTS:
            getattr(L, name)  # every declared symbol must be exported
        L.calitas_last_error.restype = C.c_char_p
        L.calitas_hitset_count.restype = C.c_int64
        L.calitas_hitset_count.argtypes = [C.c_void_p]
        L.calitas_hitset_data.restype = C.c_void_p
        L.calitas_hitset_data.argtypes = [C.c_void_p]
        L.calitas_hitset_stride.restype = C.c_int32
        L.calitas_hitset_variant_info.restype = C.c_void_p
        L.calitas_hitset_variant_info.argtypes = [C.c_void_p]
        L.calitas_tool_variant_plan_free.argtypes = [C.c_void_p]
        L.calitas_hitset_stride.argtypes = [C.c_void_p]
        L.calitas_hitset_free.argtypes = [C.c_void_p]
        L.calitas_hitset_stats.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
        L.calitas_engine_destroy.argtypes = [C.c_void_p]
        L.calitas_reference_free.argtypes = [C.c_void_p, C.c_void_p]
        L.calitas_free_text.argtypes = [C.c_void_p]

    def check(self, rc):
        if rc != 0:
            raise CalitasError(rc, (self.L.calitas_last_error() or b"").decode("latin-1"))

    def take_text(self, ptr):
        s = C.string_at(ptr).decode("latin-1")
        self.L.calitas_free_text(ptr)
        return s


_default = None


def default_library():
    global _default
    if _default is None:
        _default = Library()
    return _default


def make_guide(sequence, aux_pams=()):
    """Returns (Guide struct, keepalive)."""
    aux = [_b(a) for a in aux_pams]
    arr = (C.c_char_p * max(1, len(aux)))(*aux) if aux else None
    seq = _b(sequence)
    g = Guide(seq, arr if aux else None, len(aux))
    return g, (seq, aux, arr)


def make_guides(guides):
    """guides: list of str or (str, [aux pams])."""
    arr = (Guide * max(1, len(guides)))()
    keep = []
    for i, g in enumerate(guides):
        seq, aux = (g, ()) if isinstance(g, str) else (g[0], g[1])
        arr[i], k = make_guide(seq, aux)
        keep.append(k)
    return arr, keep


class HitSet:
    def __init__(self, lib, ptr):
        self.lib, self.ptr = lib, ptr

    def __len__(self):
        return self.lib.L.calitas_hitset_count(self.ptr)

    def stride(self):
        return self.lib.L.calitas_hitset_stride(self.ptr)

    def raw_words(self):
        """the packed records as a uint32 array of shape (n, stride / 4) (copy)"""
        import numpy as np
        n, stride = len(self), self.stride()
        if n == 0:
            return np.zeros((0, stride // 4), dtype=np.uint32)
        buf = C.cast(self.lib.L.calitas_hitset_data(self.ptr), C.POINTER(C.c_uint8 * (n * stride))).contents
        return np.frombuffer(buf, dtype=np.uint32).reshape(n, stride // 4).copy()

    def records(self):
        """hit records decoded into a numpy structured array (copy; see hit_dtype() / decode_hits)"""
        return decode_hits(self.raw_words())

    def hits(self):
        """decoded hits as a list of numpy records (attribute access: h.score, h.ops, ...)"""
        import numpy as np
        return list(self.records().view(np.recarray))

    def stats(self):
        ms = (C.c_double * 8)()
        cnt = (C.c_int64 * 8)()
        self.lib.check(self.lib.L.calitas_hitset_stats(self.ptr, ms, cnt))
        return {"ms_total": ms[0], "ms_scan": ms[1], "ms_align": ms[2], "ms_other": ms[3], "ms_d2h": ms[4], "windows": cnt[0], "candidates": cnt[1],
                "alignments": cnt[2], "launches": cnt[3], "h2d_bytes": cnt[4], "d2h_bytes": cnt[5], "scan_launches": cnt[6], "bases_scanned": cnt[7]}

    def free(self):
        if self.ptr:
            self.lib.L.calitas_hitset_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Reference:
    def __init__(self, engine, ptr, contigs, view_keep):
        self.engine, self.ptr, self.contigs, self._keep = engine, ptr, contigs, view_keep

    def free(self):
        if self.ptr:
            self.engine.lib.L.calitas_reference_free(self.engine.ptr, self.ptr)
            self.ptr = None


class Engine:
    """One engine per GPU (include/calitas_b200.h).  Mirrors `new SequentialGuideAligner(costs)`."""

    def __init__(self, device=0, costs=DEFAULT_COSTS, lib=None):
        self.lib = lib or default_library()
        self.costs = tuple(costs)
        c = Costs(*costs)
        p = C.c_void_p()
        self.lib.check(self.lib.L.calitas_engine_create(device, C.byref(c), C.byref(p)))
        self.ptr = p

    def close(self):
        if self.ptr:
            self.lib.L.calitas_engine_destroy(self.ptr)
            self.ptr = None

    def microbench_int(self, kind=0):
        v = C.c_double(0)
        self.lib.check(self.lib.L.calitas_microbench_int(self.ptr, kind, C.byref(v)))
        return v.value

    # ---- reference -------------------------------------------------------------------------------------------------------------
    @staticmethod
    def genome_view(contigs, assembly=None):
        """contigs: [(name, bases as bytes/str/numpy uint8 array)] -> (GenomeView, keepalive)."""
        n = len(contigs)
        names = (C.c_char_p * n)(*[_b(c[0]) for c in contigs])
        bufs, addrs, ptrs, lens = [], [], (C.c_char_p * n)(), (C.c_int64 * n)()
        for i, (_, b) in enumerate(contigs):
            if hasattr(b, "ctypes"):  # numpy array
                bufs.append(b)
                addrs.append(b.ctypes.data)
                lens[i] = b.size
            else:
                bb = _b(b)
                bufs.append(bb)
                addrs.append(C.cast(C.c_char_p(bb), C.c_void_p).value)
                lens[i] = len(bb)
        C.memmove(ptrs, (C.c_void_p * n)(*addrs), n * C.sizeof(C.c_void_p))
        asm = _b(assembly)
        view = GenomeView(n, names, lens, ptrs, asm)
        view._addrs = addrs
        return view, (names, bufs, ptrs, lens, asm, addrs)

    def load_reference(self, contigs, keep_raw=False, shard=None):
        """contigs: [(name, bases)] full contigs; shard=(index, count, halo) loads only that contig-range shard."""
        view, keep = self.genome_view(contigs)
        n = len(contigs)
        p = C.c_void_p()
        if shard is None:
            self.lib.check(self.lib.L.calitas_reference_load(self.ptr, n, view.names, view.lengths, C.cast(view.bases, C.POINTER(C.c_void_p)), None, None, None, None,
                                                             1 if keep_raw else 0, C.byref(p)))
        else:
            idx, cnt, halo = shard
            ob, oe, hb, he = [(C.c_int64 * n)() for _ in range(4)]
            self.lib.check(self.lib.L.calitas_shard_plan(n, view.lengths, idx, cnt, C.c_int64(halo), ob, oe, hb, he))
            ptrs = (C.c_void_p * n)()
            for i in range(n):
                ptrs[i] = view._addrs[i] + hb[i]
            self.lib.check(self.lib.L.calitas_reference_load(self.ptr, n, view.names, view.lengths, ptrs, hb, he, ob, oe, 1 if keep_raw else 0, C.byref(p)))
        return Reference(self, p, contigs, (view, keep))

    def load_reference_ranges(self, names, lengths, have, own, arrays, keep_raw=False):
        """Shard loader for callers that hold only their slice: arrays[i] = numpy uint8 bases [have[i][0], have[i][1]) of contig i."""
        n = len(names)
        cn = (C.c_char_p * n)(*[_b(x) for x in names])
        cl = (C.c_int64 * n)(*lengths)
        hb = (C.c_int64 * n)(*[h[0] for h in have])
        he = (C.c_int64 * n)(*[h[1] for h in have])
        ob = (C.c_int64 * n)(*[o[0] for o in own])
        oe = (C.c_int64 * n)(*[o[1] for o in own])
        ptrs = (C.c_void_p * n)(*[a.ctypes.data if a is not None and a.size else None for a in arrays])
        p = C.c_void_p()
        self.lib.check(self.lib.L.calitas_reference_load(self.ptr, n, cn, cl, ptrs, hb, he, ob, oe, 1 if keep_raw else 0, C.byref(p)))
        return Reference(self, p, None, (cn, cl, arrays))

    def render_alignments(self, records, guides, contigs, upper_case=True):
        """calitas_render_alignments over a numpy array of hit records (hit_dtype()); contigs = [(name, numpy uint8 array or bytes)] full contigs.
        Returns the GuideAlignment table text."""
        import numpy as np
        rec = np.ascontiguousarray(encode_hits(records))
        arr, keep = make_guides(guides)
        view, vkeep = self.genome_view(contigs)
        out = C.c_void_p()
        self.lib.check(self.lib.L.calitas_render_alignments(C.c_void_p(rec.ctypes.data), C.c_int64(rec.shape[0]), HIT_WIDE_WORDS * 4, len(guides), arr, len(contigs), view.names,
                                                            C.cast(view.bases, C.POINTER(C.c_void_p)), None, 1 if upper_case else 0, C.byref(out)))
        return self.lib.take_text(out)

    # ---- device entry points -----------------------------------------------------------------------------------------------------
    def search(self, ref, guides, limits, window_size=1000, chrom=None, dedup=True):
        arr, keep = make_guides(guides)
        lim = limits if isinstance(limits, Limits) else Limits(*limits)
        p = C.c_void_p()
        self.lib.check(self.lib.L.calitas_search(self.ptr, ref.ptr, len(guides), arr, C.byref(lim), window_size, _b(chrom), 1 if dedup else 0, C.byref(p)))
        return HitSet(self.lib, p)

    @staticmethod
    def search_sharded(engines, refs, guides, limits, window_size=1000, chrom=None):
        """calitas_search_sharded: engines[s] holds shard s in refs[s]; returns ONE merged hit set (owned by engines[0])."""
        lib = engines[0].lib
        arr, keep = make_guides(guides)
        lim = limits if isinstance(limits, Limits) else Limits(*limits)
        n = len(engines)
        e_arr = (C.c_void_p * n)(*[e.ptr for e in engines])
        r_arr = (C.c_void_p * n)(*[r.ptr for r in refs])
        p = C.c_void_p()
        lib.check(lib.L.calitas_search_sharded(n, e_arr, r_arr, len(guides), arr, C.byref(lim), window_size, _b(chrom), C.byref(p)))
        return HitSet(lib, p)

    def align_targets(self, guides, tasks, limits, best=False):
        """tasks: [(guide_idx, bases, target_offset)]"""
        arr, keep = make_guides(guides)
        lim = limits if isinstance(limits, Limits) else Limits(*limits)
        t = (TargetTask * max(1, len(tasks)))()
        bufs = []
        for i, (gi, bases, off) in enumerate(tasks):
            b = _b(bases)
            bufs.append(b)
            t[i] = TargetTask(gi, b, len(b), off)
        p = C.c_void_p()
        self.lib.check(self.lib.L.calitas_align_targets(self.ptr, len(guides), arr, C.c_int64(len(tasks)), t, C.byref(lim), 1 if best else 0, C.byref(p)))
        return HitSet(self.lib, p)

    def align_regions(self, ref, guides, tasks, limits, best=False):
        """tasks: [(guide_idx, contig_idx, start0, length)] or a prepared (RegionTask * n) array"""
        arr, keep = make_guides(guides)
        lim = limits if isinstance(limits, Limits) else Limits(*limits)
        if isinstance(tasks, C.Array):
            t, n = tasks, len(tasks)
        else:
            n = len(tasks)
            t = (RegionTask * max(1, n))()
            for i, (gi, ci, s, l) in enumerate(tasks):
                t[i] = RegionTask(gi, ci, s, l)
        p = C.c_void_p()
        self.lib.check(self.lib.L.calitas_align_regions(self.ptr, ref.ptr, len(guides), arr, C.c_int64(n), t, C.byref(lim), 1 if best else 0, C.byref(p)))
        return HitSet(self.lib, p)
