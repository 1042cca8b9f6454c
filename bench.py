#!/usr/bin/env python
"""bench.py — headline benchmark: SearchReference throughput in Gbp*guides/s on a synthetic hg38-sized genome.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference ...                      (CPU arm: the oracle restatement of the reference on the host cores)

One step = one calitas_search call (include/calitas_b200.h) of G guides against this rank's contig-range shard of the 3.1-Gbp genome:
guides go host->device, every kernel of the path runs (scan, sort, align, canonicalise, removeOverlaps + sort), final hit records
come device->host.  Shards are independent: no collective on the data path.
  value  = genome bp x guides / device time of the step (CUDA events on the engine's stream, max over ranks)
  e2e    = same, wall clock around the C-ABI call with host buffers (guide strings in, hit records out), max over ranks
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# SASS instruction mix of one Myers column update of one guide in k_scan_tiled's branch-free inner loop (counted from cuobjdump -sass of the
# shipped library, see DESIGN.md "k_scan_tiled"; per 8 columns x 2 guides: 112 LOP3 + 32 LEA.HI + 8 VIMNMX3 | 56 IMAD.IADD | 16 LDS + 8 LDS.U8):
# ALU pipe 7 LOP3 + 2 LEA.HI + 0.5 VIMNMX3; FMA pipe 3.5 IMAD.IADD; LSU 1 LDS + 0.5 LDS.U8.
ALU_OPS_PER_COLUMN = 9.5
ISSUE_SLOTS_PER_COLUMN = 14.5
NCU_DRAM_OVER_ALGORITHMIC = 406.4 / 386.2   # k_scan_tiled, profiles/r01f_summary.txt
REF_OPS_PER_BP_GUIDE = 240   # 2 strands x 20 rows x 6 int32 ops of the reference's recurrence


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--guides", type=int, default=100, help="guides per step (north_star: 100 guides, defaults d=5 p=1 g=3)")
    ap.add_argument("--scale", type=float, default=1.0, help="genome scale; 1.0 = 3.1 Gbp with hg38 contig lengths")
    ap.add_argument("--max-guide-diffs", type=int, default=5)
    ap.add_argument("--max-pam-mismatches", type=int, default=1)
    ap.add_argument("--max-gaps", type=int, default=3)
    ap.add_argument("--pam", default="nrg", help="PAM appended to every guide ('' = PAM-less run of BASELINE configs[3])")
    ap.add_argument("--aux-pams", default="", help="comma-separated auxiliary PAMs (BASELINE configs[3]: --pam ngg --aux-pams nag)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--lib", default=None, help="alternative build of libcalitas_b200.so (kernel A/B experiments); default = the product library")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the cpu_baseline sample")
    return ap.parse_args()


def guide_list(args):
    """The BASELINE guide's protospacer + random 20-mers (seed 20260103), all with --pam / --aux-pams."""
    from calitas_b200 import synth
    aux = [a for a in args.aux_pams.split(",") if a]
    seqs = [synth.BASELINE_GUIDE[:20] + args.pam] + synth.random_guides(max(0, args.guides - 1), pam=args.pam)
    return [(s, aux) for s in seqs] if aux else seqs


def guide_text(g):
    return g if isinstance(g, str) else g[0]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_indices):
        self.gpu = ",".join(str(i) for i in gpu_indices)       # rank 0 samples every GPU of the job with ONE nvidia-smi process
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None

    def start(self):
        if not self.gpu:
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", self.gpu, "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            sm.sort()
            out.update({"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)})
        return out


def cpu_sample(genome, guides, threads, target_seconds, limits_kw, sample_bp=8_000_000):
    """Times the oracle (C++ restatement of the reference algorithm, oracle/) on a bounded sample of the same workload."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle
    pyoracle.build()
    # sample: the first contig's bases after the telomere N block
    length = min(genome.lengths[0], sample_bp + 20000)
    bases = bytes(genome.range(0, 0, length))
    contigs = [(genome.names[0], bases)]
    done, t_total, n_hits = 0, 0.0, 0
    t0 = time.perf_counter()
    while done < len(guides) and (done == 0 or t_total < target_seconds):
        g = guides[done]
        n, _ = pyoracle.search_reference_count(contigs, guide_text(g), aux_pams=() if isinstance(g, str) else g[1], threads=threads, **limits_kw)
        n_hits += n
        done += 1
        t_total = time.perf_counter() - t0
    value = len(bases) * done / t_total / 1e9
    return {"value": value, "unit": "Gbp*guides/s", "cores": threads, "kind": "port",
            "sample": "%d guide(s) x first %.1f Mbp of %s, same windows/limits, %d threads, %.1f s; C++ restatement of the reference algorithm (the JVM reference cannot run here)"
                      % (done, len(bases) / 1e6, genome.names[0], threads, t_total), "hits": n_hits}


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    from calitas_b200 import synth
    guides = guide_list(args)
    genome = synth.hg38_like_genome(args.scale, guides=[guide_text(g) for g in guides], sites_per_guide=200)
    threads = os.cpu_count() or 1
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle
    pyoracle.build()
    sample_bp = 4_000_000
    length = min(genome.lengths[0], sample_bp + 20000)
    contigs = [(genome.names[0], bytes(genome.range(0, 0, length)))]
    times = []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        g = guides[it % len(guides)]
        pyoracle.search_reference_count(contigs, guide_text(g), aux_pams=() if isinstance(g, str) else g[1], threads=threads,
                                        d=args.max_guide_diffs, p=args.max_pam_mismatches, g=args.max_gaps)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    value = length / (ms * 1e-3) / 1e9
    sample = "each step: 1 guide x first %.1f Mbp of %s (bounded sample of the workload), oracle C++ restatement, %d host threads" % (length / 1e6, genome.names[0], threads)
    print(json.dumps({
        "impl": "reference", "metric": "Gbp*guides/s SearchReference (hg38-size synthetic)", "value": value, "unit": "Gbp*guides/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int32",
        "data": "synthetic", "config": workload_config(args, genome),
        "cpu_baseline": {"value": value, "unit": "Gbp*guides/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Gbp*guides/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def bind_to_gpu_numa_node(gpu_index):
    """One process per GPU: keep the rank's host threads (and therefore its pinned buffers and its launch/read-back round trips) on the CPUs
    that are local to its GPU.  Returns the CPU list used, or None when the driver's affinity mask does not intersect the allowed CPUs."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        allowed = sorted(os.sched_getaffinity(0))
        n = max(allowed) + 1
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, (n + 63) // 64)
        local = [c for c in allowed if (mask[c // 64] >> (c % 64)) & 1]
        if local and len(local) < len(allowed):
            os.sched_setaffinity(0, local)
            return local
    except Exception:
        pass
    return None


def workload_config(args, genome):
    pams = ",".join([args.pam or "(none)"] + [a for a in args.aux_pams.split(",") if a])
    return {"workload": "SearchReference, %d guides (CTTGCCCCACAGGGCAGTAA + random 20-mers, PAMs %s) vs %.2f Gbp synthetic hg38-sized genome (24 contigs), "
                        "d=%d p=%d g=%d O=10 w=1000, contig-range sharded" % (args.guides, pams, genome.total() / 1e9, args.max_guide_diffs, args.max_pam_mismatches, args.max_gaps),
            "guides_per_step": args.guides, "genome_bp": genome.total(), "window_size": 1000, "max_guide_diffs": args.max_guide_diffs,
            "max_pam_mismatches": args.max_pam_mismatches, "max_gaps_between_guide_and_pam": args.max_gaps, "pams": pams, "max_overlap": 10,
            "dedup": "removeOverlaps+sort on device",
            "l2": "inputs larger than L2 (packed reference shard per scan launch >> 126 MB)", "parallelism": "contig-range shards, 1 rank per GPU, no collective"}


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"          # the version banner would land on stdout next to the one JSON line
    import numpy as np
    import torch
    import torch.distributed as dist
    from calitas_b200 import synth
    from calitas_b200._capi import Engine, Library, Limits

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU path")
    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 and os.environ.get("CALITAS_BENCH_BIND", "1") != "0" else None
    if os.environ.get("CALITAS_BENCH_VERBOSE"):
        print("[bench] rank %d cpus %s" % (rank, sorted(os.sched_getaffinity(0))), file=sys.stderr)
    if world > 1:
        # NCCL prints its version banner on stdout when the first communicator comes up; the contract is ONE JSON line on stdout,
        # so stdout is pointed at stderr while the process group initialises.
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            try:
                import ctypes
                ctypes.CDLL(None).fflush(None)        # NCCL printf()s into the C stdio buffer: drain it while fd 1 still points at stderr
            except Exception:
                pass
            os.dup2(saved, 1)
            os.close(saved)
    guides = guide_list(args)
    genome = synth.hg38_like_genome(args.scale, guides=[guide_text(g) for g in guides], sites_per_guide=200)
    n = len(genome.lengths)
    engine = Engine(local_rank, lib=Library(os.path.abspath(args.lib)) if args.lib else None)

    # ---- this rank's contig-range shard: generate only the bases it holds ------------------------------------------------------
    import ctypes as C
    L = (C.c_int64 * n)(*genome.lengths)
    ob, oe, hb, he = [(C.c_int64 * n)() for _ in range(4)]
    halo = 4 * 1000
    engine.lib.check(engine.lib.L.calitas_shard_plan(n, L, rank, world, C.c_int64(halo), ob, oe, hb, he))
    t0 = time.perf_counter()
    arrays = [genome.range(c, hb[c], he[c]) if he[c] > hb[c] else None for c in range(n)]
    t_gen = time.perf_counter() - t0
    t0 = time.perf_counter()
    ref = engine.load_reference_ranges(genome.names, genome.lengths, [(hb[c], he[c]) for c in range(n)], [(ob[c], oe[c]) for c in range(n)], arrays)
    t_load = time.perf_counter() - t0
    own_bp = sum(oe[c] - ob[c] for c in range(n))
    del arrays
    lim = Limits(args.max_guide_diffs, args.max_pam_mismatches, args.max_gaps, -1, 10)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        t0 = time.perf_counter()
        hs = engine.search(ref, guides, lim, window_size=1000, dedup=True)
        wall = time.perf_counter() - t0
        st = hs.stats()
        st["hits"] = len(hs)
        st["wall_ms"] = wall * 1e3
        hs.free()
        return st

    # rank 0 samples the job's GPUs: CUDA_VISIBLE_DEVICES entries (indices or UUIDs, both accepted by nvidia-smi -i) when the launcher set it
    visible = [x.strip() for x in os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",") if x.strip()]
    sampler = ClockSampler((visible[:world] if len(visible) >= world else range(world)) if rank == 0 else [])
    sampler.start()                      # nvidia-smi needs ~1 s to start: launched before the warm-up so that it is sampling during the timed steps
    for _ in range(args.warmup):
        step()
    barrier()
    t_begin = time.perf_counter()
    stats = [step() for _ in range(args.steps)]
    torch.cuda.synchronize()
    t_local = time.perf_counter() - t_begin
    barrier()
    clocks = sampler.stop()

    dev_ms = sum(s["ms_total"] for s in stats) / args.steps        # CUDA events, whole step on the device incl. D2H of hits
    wall_ms = 1e3 * t_local / args.steps
    scan_ms = sum(s["ms_scan"] for s in stats) / args.steps
    red = torch.tensor([dev_ms, wall_ms, scan_ms], dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(stats[-1]["hits"]), float(own_bp), float(stats[-1]["candidates"])], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    per_rank = None
    if world > 1:
        mine = torch.tensor([dev_ms, scan_ms, float(stats[-1]["candidates"])], dtype=torch.float64, device="cuda")
        allr = torch.zeros(world * 3, dtype=torch.float64, device="cuda")
        dist.all_gather_into_tensor(allr, mine)
        per_rank = allr.view(world, 3).tolist()
    dev_ms_max, wall_ms_max, scan_ms_max = [float(x) for x in red.tolist()]
    total_hits, total_bp, total_cand = [float(x) for x in tot.tolist()]

    if rank == 0:
        G = len(guides)
        bpg = genome.total() * G
        value = bpg / (dev_ms_max * 1e-3) / 1e9
        e2e = bpg / (wall_ms_max * 1e-3) / 1e9
        st = stats[-1]
        # dominant kernel: k_scan_tiled (rank 0's launches)
        launches = max(1, st["scan_launches"])
        scan_launch_ms = st["ms_scan"] / launches
        alg_bytes = st["bases_scanned"] / launches * 0.5                     # 4-bit packed reference, read once per launch
        hbm_achieved = alg_bytes / (scan_launch_ms * 1e-3) / 1e9
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        int_peaks = {k: engine.microbench_int(i) for i, k in enumerate(("alu_lop3", "fma_imad", "both_lop3_imad", "fma_imad_hi", "alu_lea_hi"))}
        own0 = sum(oe[c] - ob[c] for c in range(n))
        columns = own0 * G * 2                                                  # rank 0's shard: one Myers column per base, strand and guide
        int_achieved = columns * ALU_OPS_PER_COLUMN / (st["ms_scan"] * 1e-3) / 1e12
        out = {
            "metric": "Gbp*guides/s SearchReference (hg38-size synthetic)", "value": value, "unit": "Gbp*guides/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms_max, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": workload_config(args, genome),
            "e2e": {"value": e2e, "unit": "Gbp*guides/s", "ms_per_step": wall_ms_max, "h2d_bytes_per_step": st["h2d_bytes"], "d2h_bytes_per_step": st["d2h_bytes"],
                    "api": "calitas_search (C ABI): guide strings + limits in host memory -> deduplicated, sorted hit records in pinned host memory",
                    "reference": "the packed reference shard stays resident in HBM between calls, like the reference's in-memory FASTA; uploading and packing it "
                                 "from host bytes (calitas_reference_load) took setup_s.load_and_pack on rank 0",
                    "first_call_value_incl_reference_upload": bpg / (wall_ms_max * 1e-3 + t_load) / 1e9},
            "gpu_launches": int(sum(s["launches"] for s in stats)),
            "clocks": clocks,
            "roofline": {"kernel": "k_scan_tiled", "bound": "hbm", "achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_achieved / hbm_peak,
                         "traffic": alg_bytes * NCU_DRAM_OVER_ALGORITHMIC, "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture "
                         "(profiles/r01f_summary.txt: 388.2 + 18.1 = 406.4 MB for 386.2 MB algorithmic), scaled to this launch's algorithmic bytes", "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s",
                         "note": "streaming read of the 4-bit packed shard once per launch; the kernel is integer-ALU-bound (see roofline_int), so the HBM fraction is small by design",
                         "avg_launch_ms": scan_launch_ms, "launches_per_step": launches, "share_of_step": st["ms_scan"] / st["ms_total"]},
            "roofline_int": {"kernel": "k_scan_tiled", "bound": "int_alu_pipe", "achieved": int_achieved, "peak": int_peaks["alu_lop3"], "unit": "Tiop/s (ALU-pipe thread instructions)",
                             "frac": int_achieved / int_peaks["alu_lop3"] if int_peaks["alu_lop3"] else None,
                             "alu_ops_per_column": ALU_OPS_PER_COLUMN, "issue_slots_per_column": ISSUE_SLOTS_PER_COLUMN,
                             "issue_frac": columns * ISSUE_SLOTS_PER_COLUMN / (st["ms_scan"] * 1e-3) / 1e12 / int_peaks["both_lop3_imad"] if int_peaks["both_lop3_imad"] else None,
                             "peak_source": "measured on this GPU by calitas_microbench_int (LOP3 chains; no integer peak in MEASURED_PEAKS.json)",
                             "measured_peaks_tiops": int_peaks,
                             "reference_equivalent_tiops": bpg * REF_OPS_PER_BP_GUIDE / (dev_ms_max * 1e-3) / 1e12,
                             "gcups_equivalent": value * 40},
            "breakdown_ms": {"scan": st["ms_scan"], "align": st["ms_align"], "sort_canon_dedup": st["ms_other"], "d2h": st["ms_d2h"], "wall": st["wall_ms"]},
            "counts": {"hits": total_hits, "candidates": total_cand, "windows_rank0": st["windows"], "genome_bp": genome.total(), "shard_bp_sum": total_bp},
            "setup_s": {"generate": t_gen, "load_and_pack": t_load},
        }
        out["config"]["cpu_binding_rank0"] = ("cpus %d-%d (GPU-local NUMA node)" % (numa[0], numa[-1])) if numa else "none"
        if per_rank is not None:
            out["per_rank"] = {"ms_per_step": [round(r[0], 3) for r in per_rank], "scan_ms": [round(r[1], 3) for r in per_rank], "candidates": [int(r[2]) for r in per_rank]}
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_sample(genome, guides, os.cpu_count() or 1, args.cpu_seconds, dict(d=args.max_guide_diffs, p=args.max_pam_mismatches, g=args.max_gaps))
        print(json.dumps(out))
    ref.free()
    engine.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
