"""Per-source-line instruction counts of one kernel: joins the SASS page of an ncu report (ncu -i X.ncu-rep --page source --csv) with the
line table of the cubin (nvdisasm -g) by instruction order.  usage: python profiles/hotlines.py <report.ncu-rep> <lib.so> <mangled-substring> [top]"""
import csv, re, subprocess, sys, tempfile, os, glob

rep, lib, func = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = next(r for r in rows if r and r[0] == "Address")
ix, isrc, iex, ismp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
ins = [(r[isrc], int(r[iex] or 0), int(r[ismp] or 0)) for r in rows if len(r) == len(hdr) and r[0] != "Address"]
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, capture_output=True)
cubin = max(glob.glob(d + "/*.cubin"), key=os.path.getsize)
sass = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
lines, cur, on = [], None, False
stack = ""
for l in sass:
    if l.startswith(".text."):
        on = func in l
        continue
    if not on:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)), m.group(3)); continue
    if re.match(r"\s+/\*[0-9a-f]{4}\*/", l):
        lines.append(cur)
assert len(lines) == len(ins), (len(lines), len(ins))
agg, tot, opc = {}, 0, {}
for (src, ex, smp), loc in zip(ins, lines):
    k = (loc[0], loc[1]) if loc else ("?", 0)
    a = agg.setdefault(k, [0, 0]); a[0] += ex; a[1] += smp; tot += ex
    op = src.split()[0] if not src.startswith("@") else src.split()[1]
    op = op.split(".")[0]; opc[op] = opc.get(op, 0) + ex
print("total warp instructions", tot)
src_cache = {}
for k, (ex, smp) in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    text = ""
    for root in ("calitas_b200/csrc/", ""):
        p = root + k[0]
        if os.path.exists(p):
            src_cache.setdefault(p, open(p).read().splitlines()); text = src_cache[p][k[1] - 1].strip()[:120]; break
    print("%5.1f%% %7d smp  %s:%d  %s" % (100.0 * ex / tot, smp, k[0], k[1], text))
print("opcodes:", ", ".join("%s %.1f%%" % (o, 100.0 * c / tot) for o, c in sorted(opc.items(), key=lambda x: -x[1])[:14]))
