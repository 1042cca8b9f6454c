"""Turns the ncu outputs brought back in gpurun_out/ into the small text summaries committed under profiles/.
usage: python profiles/summarize.py <round-tag> <launches.csv> [name=report.ncu-rep ...]"""
import collections
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size", "sm__inst_executed.sum.per_cycle_elapsed",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.avg.per_second", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__average_warps_issue_stalled", "local_load", "local_store", "l1tex__t_sectors_pipe_lsu_mem_local",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct"]


def launches(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        a = agg.setdefault(r[ki].split("(")[0][:70], [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    out.write("# kernel launches (ncu --metrics gpu__time_duration.sum --clock-control none): cold-cache, serialised -> compare SHARES\n")
    out.write("%-72s %6s %14s %7s\n" % ("kernel", "count", "total_us", "share"))
    for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        out.write("%-72s %6d %14.1f %6.1f%%\n" % (n, c, t / 1e3, 100 * t / tot))


def report(name, path, out):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    out.write("\n# %s: ncu --set full --clock-control none (%s)\n" % (name, path.split("/")[-1]))
    for vals in rows[2:]:
        out.write("kernel: %s\n" % vals[hdr.index("Kernel Name")])
        for i, h in enumerate(hdr):
            if any(h == w or h.startswith(w) for w in WANT):
                out.write("  %-80s %14s %s\n" % (h, vals[i], units[i]))


if __name__ == "__main__":
    tag = sys.argv[1]
    with open("profiles/%s_summary.txt" % tag, "w") as out:
        launches(sys.argv[2], out)
        for spec in sys.argv[3:]:
            name, path = spec.split("=")
            report(name, path, out)
    print(open("profiles/%s_summary.txt" % tag).read())
