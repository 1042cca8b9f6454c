"""AlignToReference batch (BASELINE configs[1]): N (guide, locus) pairs, --window-size 60, one B200; times calitas_align_regions."""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from calitas_b200 import synth
from calitas_b200._capi import Engine, Library, Limits, RegionTask

ap = argparse.ArgumentParser()
ap.add_argument("--tasks", type=int, default=1_000_000)
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--guides", type=int, default=100)
ap.add_argument("--lib", default=None)
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()
guides = [synth.BASELINE_GUIDE] + synth.random_guides(args.guides - 1)
genome = synth.hg38_like_genome(args.scale, guides=guides, sites_per_guide=200)
e = Engine(0, lib=Library(os.path.abspath(args.lib)) if args.lib else None)
arrays = [genome.contig(c) for c in range(len(genome.lengths))]
ref = e.load_reference(list(zip(genome.names, arrays)))
rng = np.random.default_rng(20260105)
sites = [(c, pos, len(seq)) for c, lst in enumerate(genome.planted) for (pos, seq) in lst]
n = args.tasks
tasks = (RegionTask * n)()
p_contig = np.array(genome.lengths, dtype=np.float64) / genome.total()
near = rng.random(n) < 0.5
site_idx = rng.integers(0, len(sites), size=n)
jit = rng.integers(-10, 11, size=n)
cont = rng.choice(len(genome.lengths), size=n, p=p_contig)
u = rng.random(n)
gidx = rng.integers(0, len(guides), size=n)
pad = 30
for i in range(n):
    if near[i]:
        c, pos, ln = sites[site_idx[i]]
        p = pos + ln // 2 + int(jit[i])
    else:
        c = int(cont[i]); p = 1 + int(u[i] * (genome.lengths[c] - 1))
    p = max(1, min(genome.lengths[c], p))
    rs, re_ = max(p - pad, 1), min(p + pad, genome.lengths[c])
    tasks[i] = RegionTask(int(gidx[i]), c, rs - 1, re_ - rs + 1)
out = {}
for mode, best, lim in (("all_d5_p1_O10", False, Limits(5, 1, 3, -1, 10)), ("best", True, Limits(0, 0, 3, -1, 0))):
    times = []
    for rep in range(args.reps):
        t0 = time.perf_counter()
        hs = e.align_regions(ref, guides, tasks, lim, best=best)
        dt = time.perf_counter() - t0
        st = hs.stats(); nh = len(hs); hs.free()
        times.append(dt)
    out[mode] = {"tasks": n, "wall_s_best": min(times), "tasks_per_s": n / min(times), "hits": nh, "dev_ms": st["ms_total"], "scan_ms": st["ms_scan"], "align_ms": st["ms_align"],
                 "other_ms": st["ms_other"], "d2h_ms": st["ms_d2h"], "candidates": st["candidates"], "launches": st["launches"]}
print(json.dumps(out))
