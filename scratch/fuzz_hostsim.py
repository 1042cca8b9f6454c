"""Randomised hostsim-vs-oracle hunt (not collected by pytest): random guides (IUPAC, 5'/3'/no PAM, auxiliary PAMs), contigs from 1 base to a few
windows with planted sites, N runs, lower case, IUPAC and junk bytes, random limits / window sizes / cost sets.  usage: python scratch/fuzz_hostsim.py [n_seeds] [first_seed]
Round 1: 600 seeds from 20000 clean; the tandem-repeat variant of this hunt found the shard-cut merge bug fixed in 7c79f4e."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "tests"), ROOT, os.path.join(ROOT, "oracle")]
import pyoracle, backends
from calitas_b200 import synth
import test_parity_random as T

IUP = "ACGTRYSWKMBDHVN"


def hunt(n_seeds, first, eng=None, verbose=True):
  """Returns the list of seeds whose tables differ between the oracle and `eng` (default: the host simulation)."""
  eng = eng or backends.get("hostsim")
  bad = []
  for seed in range(first, first + n_seeds):
      rng = np.random.default_rng(seed)
      lp = int(rng.integers(6, 33))
      proto = "".join(rng.choice(list("ACGT" if rng.random() < 0.7 else IUP), size=lp))
      concrete = "".join(c if c in "ACGT" else "ACGT"[int(rng.integers(4))] for c in proto)
      pam = "".join(rng.choice(list("acgtnry"), size=int(rng.integers(1, 7))))
      kind = seed % 3
      guide = proto if kind == 0 else (proto + pam if kind == 1 else pam + proto)
      aux = [] if kind == 0 or rng.random() < 0.5 else ["".join(rng.choice(list("acgtn"), size=int(rng.integers(1, 7)))) for _ in range(int(rng.integers(1, 4)))]
      contigs = []
      for i in range(int(rng.integers(1, 25))):
          L = max(1, int(rng.choice([1, 2, lp - 1, lp, lp + 3, 40, 969, 970, 971, 1000, 1001, 1940, int(rng.integers(1, 5000))])))
          b = list(rng.choice(list("ACGT"), size=L))
          for _ in range(int(rng.integers(0, 6))):
              if L > lp + 12:
                  p = int(rng.integers(0, L - lp - 10))
                  s = synth.mutate_protospacer(rng, concrete.encode(), int(rng.integers(0, 5))).decode()
                  s = (s + "".join(rng.choice(list("ACGT"), size=6))) if kind != 2 else ("".join(rng.choice(list("ACGT"), size=6)) + s)
                  if rng.random() < 0.5:
                      s = T.rc(s)
                  b[p:p + len(s)] = list(s)[:max(0, L - p)]
          for _ in range(int(rng.integers(0, 3))):
              if L > 4:
                  p = int(rng.integers(0, L)); k = int(rng.integers(1, 40)); ch = str(rng.choice(["N", "N", "n", "R", "a", "X"]))
                  for q in range(p, min(L, p + k)):
                      b[q] = ch if ch != "a" else b[q].lower()
          contigs.append(("c%d" % i, "".join(b[:L]).encode()))
      d = int(rng.integers(0, min(7, lp))); p_ = int(rng.integers(0, 3)); g_ = int(rng.integers(0, 4))
      w = int(rng.choice([1000, 1000, 300, lp + len(pam) + d + g_ + int(rng.integers(0, 40))]))
      costs = [(-120, -122, -121, -260), (-100, -130, -110, -200), (-250, -120, -120, -260), (-120, -121, -122, -260)][seed % 4]
      kw = dict(d=d, p=p_, g=g_, O=int(rng.choice([0, 1, 10, 50])), window_size=w, costs=costs)
      if rng.random() < 0.3:
          kw["D"] = int(rng.integers(0, d + p_ + g_ + 1))
      res = []
      for impl in (pyoracle, eng):
          try:
              res.append(impl.search_reference(contigs, guide, aux_pams=aux, raw=True, **kw))
          except Exception as ex:
              res.append(("ERR", str(ex)[:80]))
      if res[0] != res[1] and not (isinstance(res[0], tuple) and isinstance(res[1], tuple)):
          bad.append(seed)
          if verbose:
              print("MISMATCH seed", seed, guide, aux, kw)
  return bad


if __name__ == "__main__":
    n_seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    first = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
    t0 = time.time()
    bad = hunt(n_seeds, first)
    print("seeds", n_seeds, "mismatches", len(bad), "in %.0f s" % (time.time() - t0))
    sys.exit(1 if bad else 0)
