"""bench.py --workload config5: SearchReference -v (BASELINE configs[4]) — reference windows plus the variant windows of a synthetic PrepareVcf-shaped VCF,
merged, de-duplicated and sorted on the device.

One step = one calitas_search_variants call (through calitas_tool_variant_plan_search, include/calitas_b200_tools.h) on this rank's contig-range shard:
guide strings, the variant windows' bases and descriptors go host->device, the scan / align / canonicalise kernels run over the reference windows and the
variant windows, both hit lists are merged and swept per (guide, contig, strand, variant set), records and annotations come device->host.
The plan (VCF parsing, variant windows, variant-set numbering: host code mirroring SearchReference.scala:217-399) is built once per rank before the
timed region and reported under setup_s; `tool_e2e` times the whole operator (calitas_tool_search_reference: plan + search + rendered TSV) once."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np


def run_vcf(args, job):
    import bench as B
    from calitas_b200 import synth
    from calitas_b200._capi import Engine, GenomeView, Library, Limits, SearchOptions, HitSet, make_guides, _b
    w = args.w
    rank, world = job.rank, job.world
    guides = B.guide_list(w)
    genome = B.make_genome(w, args.scale, guides)
    n = len(genome.lengths)
    engine = Engine(job.local_rank, lib=Library(os.path.abspath(args.lib)) if args.lib else None)
    lib = engine.lib
    L = (C.c_int64 * n)(*genome.lengths)
    ob, oe, hb, he = [(C.c_int64 * n)() for _ in range(4)]
    halo = 4 * 1000
    lib.check(lib.L.calitas_shard_plan(n, L, rank, world, C.c_int64(halo), ob, oe, hb, he))
    t0 = time.perf_counter()
    arrays = [genome.range(c, hb[c], he[c]) if he[c] > hb[c] else None for c in range(n)]
    t_gen = time.perf_counter() - t0
    t0 = time.perf_counter()
    ref = engine.load_reference_ranges(genome.names, genome.lengths, [(hb[c], he[c]) for c in range(n)], [(ob[c], oe[c]) for c in range(n)], arrays)
    t_load = time.perf_counter() - t0
    own_bp = sum(oe[c] - ob[c] for c in range(n))
    # ---- this rank's share of the VCF: records whose POS lies in the window starts it owns --------------------------------------------------------
    t0 = time.perf_counter()
    vcf = synth.synthetic_vcf_fast(genome, [a[ob[c] - hb[c]:oe[c] - hb[c]] if a is not None else None for c, a in enumerate(arrays)], [(ob[c], oe[c]) for c in range(n)], w["records"])
    t_vcf = time.perf_counter() - t0
    n_records = vcf.count("\n") - 2
    # genome view over the shard: bases[c] must point at base 0 of contig c; only the loaded range is ever dereferenced (hits and variants lie inside it)
    names = (C.c_char_p * n)(*[_b(x) for x in genome.names])
    ptrs = (C.c_void_p * n)(*[(a.ctypes.data - hb[c]) if a is not None and a.size else None for c, a in enumerate(arrays)])
    view = GenomeView(n, names, L, C.cast(ptrs, C.POINTER(C.c_char_p)), None)
    lim = Limits(w["d"], w["p"], w["g"], -1, 10)
    garr, gkeep = make_guides(guides)
    vtext = _b(vcf)
    opt = SearchOptions(b"g", 16, 1000, lim, None, vtext, b"synthetic.vcf:0", b"", b"bench")
    refs = (C.c_void_p * 1)(ref.ptr)
    plan = C.c_void_p()
    t0 = time.perf_counter()
    lib.check(lib.L.calitas_tool_variant_plan_create(C.byref(view), C.byref(opt), len(guides), garr, 1, refs, C.byref(plan)))
    t_plan = time.perf_counter() - t0
    cnt = [C.c_int64(0) for _ in range(3)]
    lib.check(lib.L.calitas_tool_variant_plan_counts(plan, 0, C.byref(cnt[0]), C.byref(cnt[1]), C.byref(cnt[2])))

    def step():
        t0 = time.perf_counter()
        p = C.c_void_p()
        lib.check(lib.L.calitas_tool_variant_plan_search(plan, 0, engine.ptr, ref.ptr, len(guides), garr, C.byref(lim), 1000, None, C.byref(p)))
        hs = HitSet(lib, p)
        wall = time.perf_counter() - t0
        st = hs.stats(); st["hits"] = len(hs); st["wall_ms"] = wall * 1e3
        info = lib.L.calitas_hitset_variant_info(hs.ptr)
        if info and len(hs):
            arr = np.frombuffer(C.cast(info, C.POINTER(C.c_int32 * (6 * len(hs)))).contents, dtype=np.int32).reshape(-1, 6)
            st["variant_hits"] = int((arr[:, 5] > 0).sum())
        else:
            st["variant_hits"] = 0
        hs.free()
        return st

    stats, t_local, clocks = job.timed(step, args.warmup, args.steps)
    dev_ms = sum(s["ms_total"] for s in stats) / args.steps
    wall_ms = 1e3 * t_local / args.steps
    dev_ms_max, wall_ms_max = job.reduce([dev_ms, wall_ms], "max")
    total_hits, total_var, total_rec, total_win = job.reduce([float(stats[-1]["hits"]), float(stats[-1]["variant_hits"]), float(n_records), float(cnt[1].value)], "sum")
    per_rank = job.gather([dev_ms, stats[-1]["ms_scan"]])
    # the whole operator once (rank-local): VCF text in, TSV out
    t0 = time.perf_counter()
    out, nh = C.c_void_p(), C.c_int64(0)
    g0, k0 = make_guides(guides[:1])
    lib.check(lib.L.calitas_tool_search_reference(engine.ptr, ref.ptr, C.byref(view), g0, C.byref(opt), C.byref(out), C.byref(nh)))
    t_tool = time.perf_counter() - t0
    tsv = lib.take_text(out)
    if rank == 0:
        G = len(guides)
        bpg = genome.total() * G
        value, e2e = bpg / (dev_ms_max * 1e-3) / 1e9, bpg / (wall_ms_max * 1e-3) / 1e9
        st = stats[-1]
        peaks = B.peaks_file()
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        launches = max(1, st["scan_launches"])
        alg_bytes = st["bases_scanned"] / launches * 0.5
        scan_launch_ms = st["ms_scan"] / launches
        out_d = {"metric": B.METRIC, "value": value, "unit": "Gbp*guides/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms_max,
                 "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int32", "data": "synthetic", "config": B.workload_config(args, genome, guides),
                 "e2e": {"value": e2e, "unit": "Gbp*guides/s", "ms_per_step": wall_ms_max, "h2d_bytes_per_step": st["h2d_bytes"], "d2h_bytes_per_step": st["d2h_bytes"],
                         "api": "calitas_search_variants via calitas_tool_variant_plan_search (C ABI): guides, variant-window bases and descriptors in host memory -> merged, "
                                "de-duplicated, sorted hit records + variant annotations in pinned host memory"},
                 "gpu_launches": int(sum(s["launches"] for s in stats)), "clocks": clocks,
                 "roofline": {"kernel": "k_scan_tiled", "bound": "hbm", "achieved": alg_bytes / (scan_launch_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                              "frac": alg_bytes / (scan_launch_ms * 1e-3) / 1e9 / hbm_peak, "traffic": alg_bytes * B.NCU_SCAN["dram_over_algorithmic"], "avg_launch_ms": scan_launch_ms,
                              "launches_per_step": launches, "share_of_step": st["ms_scan"] / st["ms_total"],
                              "note": "scan kernels (tiled reference scan + explicit variant-window scan) are integer-ALU-bound; see the default workload's roofline_int"},
                 "breakdown_ms": {"scan": st["ms_scan"], "align": st["ms_align"], "sort_canon_dedup_merge": st["ms_other"], "d2h": st["ms_d2h"], "wall": st["wall_ms"]},
                 "counts": {"hits": total_hits, "hits_with_variants": total_var, "vcf_records": total_rec, "variant_windows": total_win, "genome_bp": genome.total(), "shard_bp_rank0": own_bp},
                 "setup_s": {"generate": t_gen, "load_and_pack": t_load, "vcf_text": t_vcf, "variant_plan_host": t_plan},
                 "tool_e2e": {"seconds": t_tool, "rows": int(nh.value), "what": "calitas_tool_search_reference on rank 0's shard, 1 guide: VCF text -> plan -> device search + merge -> rendered TSV (%d bytes)" % len(tsv)}}
        if per_rank is not None:
            out_d["per_rank"] = {"ms_per_step": [round(r[0], 3) for r in per_rank], "scan_ms": [round(r[1], 3) for r in per_rank]}
        if not args.no_parity_check:
            # the rendered rows of rank 0's first 2 Mbp against the oracle (same VCF records, the slice as a contig of the same name)
            po = B.oracle()
            length = min(2_000_000, oe[0] - ob[0])
            sl = bytes(arrays[0][ob[0] - hb[0]:ob[0] - hb[0] + length]) if ob[0] == 0 else None
            if sl is not None:
                sub_vcf = "\n".join(l for l in vcf.split("\n") if l.startswith("#") or (l.startswith(genome.names[0] + "\t") and int(l.split("\t", 2)[1]) < length - 100)) + "\n"
                g1 = guides[0]
                exp = [l for l in po.search_reference([(genome.names[0], sl)], B.guide_text(g1), guide_id="g", aux_pams=B.guide_aux(g1), vcf_text=sub_vcf, vcf_name="synthetic.vcf:0",
                                                      raw=True, threads=os.cpu_count() or 1, **B.limits_kw(w)).split("\n") if l]
                inside = lambda l: l.split("\t")[3] == genome.names[0] and int(l.split("\t")[5]) <= length - 1500
                fix = lambda l: "\t".join(f for i, f in enumerate(l.split("\t")) if i not in (30, 33))         # aligner_version, time_stamp: run-dependent columns
                e_rows = [fix(l) for l in exp[1:] if inside(l)]
                g_rows = [fix(l) for l in tsv.split("\n")[1:] if l and inside(l)]
                out_d["parity_check"] = {"checked_rows": len(e_rows), "rows_with_variants": sum("+variants" in l for l in e_rows), "equal": g_rows == e_rows,
                                         "region": "%s:0-%d" % (genome.names[0], length - 1500), "what": "all columns of the TSV except aligner_version and time_stamp"}
        print(json.dumps(out_d))
    lib.L.calitas_tool_variant_plan_free(plan)
    ref.free()
    engine.close()
