"""Copies the bench lines of a round from gpurun_out/ into profiles/ under stable names and writes profiles/<round>_summary.md.
usage: python profiles/collect.py r02 name=gpurun_out/file.json ..."""
import json
import os
import shutil
import sys

tag = sys.argv[1]
rows = []
for spec in sys.argv[2:]:
    name, path = spec.split("=")
    if not os.path.exists(path):
        print("missing", path)
        continue
    try:
        d = json.load(open(path))
    except Exception as ex:
        print("bad", path, ex)
        continue
    dst = "profiles/%s_%s.json" % (tag, name)
    shutil.copyfile(path, dst)
    rf, ri, pc, cb = d.get("roofline") or {}, d.get("roofline_int") or {}, d.get("parity_check") or {}, d.get("cpu_baseline") or {}
    rows.append((name, d["n_gpus"], d["value"], d["unit"], d["e2e"]["value"], d["ms_per_step"], d["e2e"].get("d2h_bytes_per_step", 0) / 1e6, rf.get("kernel", ""),
                 ri.get("frac"), rf.get("frac"), pc.get("equal"), pc.get("checked_rows"), cb.get("value"), cb.get("cores"), (d.get("clocks") or {}).get("sm_mhz"), (d.get("clocks") or {}).get("reasons")))
with open("profiles/%s_summary.md" % tag, "w") as out:
    out.write("# bench lines of %s (copied from gpurun_out/ by profiles/collect.py; one JSON line each in profiles/%s_<name>.json)\n\n" % (tag, tag))
    out.write("| line | GPUs | value | unit | e2e | ms/step | D2H MB/step | dominant kernel | ALU-pipe frac | HBM frac | parity_check | rows checked | CPU port | cores | SM MHz | throttle |\n|" + "---|" * 16 + "\n")
    for r in rows:
        f = lambda v, fmt="%.3g": "" if v is None else (fmt % v if isinstance(v, float) else str(v))
        out.write("| %s | %d | %.1f | %s | %.1f | %.2f | %.0f | %s | %s | %s | %s | %s | %s | %s | %s | %s |\n" % (r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7], f(r[8]), f(r[9], "%.2g"), f(r[10]), f(r[11]), f(r[12]), f(r[13]), f(r[14]), f(r[15])))
print(open("profiles/%s_summary.md" % tag).read())
