"""Writes the synthetic hg38-sized genome (or a scaled one) as FASTA + .fai + .dict; prints the planted guide list file."""
import argparse, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from calitas_b200 import synth
ap = argparse.ArgumentParser(); ap.add_argument("--out", required=True); ap.add_argument("--scale", type=float, default=1.0); ap.add_argument("--guides", type=int, default=100)
a = ap.parse_args()
guides = [synth.BASELINE_GUIDE] + synth.random_guides(a.guides - 1)
g = synth.hg38_like_genome(a.scale, guides=guides, sites_per_guide=200)
t0 = time.time(); off = 0; fai = []
with open(a.out, "wb") as f:
    for c, name in enumerate(g.names):
        b = g.contig(c); hdr = (">%s\n" % name).encode(); f.write(hdr); off += len(hdr)
        fai.append("%s\t%d\t%d\t60\t61" % (name, b.size, off))
        full = (b.size // 60) * 60
        m = np.empty((full // 60, 61), dtype=np.uint8); m[:, :60] = b[:full].reshape(-1, 60); m[:, 60] = 10; m.tofile(f)
        if b.size > full: f.write(bytes(b[full:]) + b"\n")
        off += b.size + (b.size + 59) // 60
open(a.out + ".fai", "w").write("\n".join(fai) + "\n")
open(os.path.splitext(a.out)[0] + ".dict", "w").write("@HD\tVN:1.5\n" + "".join("@SQ\tSN:%s\tLN:%d\tAS:SYNHG38\n" % (n, l) for n, l in zip(g.names, g.lengths)))
open(os.path.splitext(a.out)[0] + ".guides.tsv", "w").write("".join("g%d\t%s\n" % (i, s) for i, s in enumerate(guides)))
print("wrote %s (%.2f Gbp) in %.1f s" % (a.out, g.total() / 1e9, time.time() - t0))
