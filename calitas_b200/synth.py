"""Deterministic synthetic inputs for the parity tests and bench.py: genomes, planted off-target sites,
guide sets, AlignToReference task lists and PrepareVcf-shaped VCFs.  numpy only; block-wise seeded so that any rank can
generate any base range of the same genome independently (contig-range sharding needs no broadcast)."""
import numpy as np

BLOCK = 1 << 20
ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.arange(256, dtype=np.uint8)
for a, b in zip(b"ACGTUMRWSYKVHDBNacgtumrwsykvhdbn", b"TGCAAKYWSRMBDHVNtgcaakywsrmbdhvn"):
    _COMP[a] = b

HG38_LENGTHS = [248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636, 138394717, 133797422, 135086622, 133275309,
                114364328, 107043718, 101991189, 90338345, 83257441, 80373285, 58617616, 64444167, 46709983, 50818468, 156040895, 57227415]
HG38_NAMES = ["chr%d" % i for i in range(1, 23)] + ["chrX", "chrY"]

BASELINE_GUIDE = "CTTGCCCCACAGGGCAGTAAnrg"   # BASELINE.json configs[0]


def revcomp_bytes(b):
    return bytes(_COMP[np.frombuffer(b, dtype=np.uint8)][::-1])


class Genome:
    """A synthetic genome defined by (names, lengths, seed, N blocks, planted sites); bases are produced on demand per range."""

    def __init__(self, names, lengths, seed, n_blocks=None, planted=None):
        self.names, self.lengths, self.seed = list(names), [int(x) for x in lengths], int(seed)
        self.n_blocks = n_blocks if n_blocks is not None else [[] for _ in names]      # per contig: [(begin, end)]
        self.planted = planted if planted is not None else [[] for _ in names]         # per contig: [(pos, bytes)]

    def range(self, contig, begin, end):
        """bases [begin, end) of contig as a numpy uint8 array"""
        begin, end = max(0, int(begin)), min(self.lengths[contig], int(end))
        out = np.empty(max(0, end - begin), dtype=np.uint8)
        if end <= begin:
            return out
        for blk in range(begin // BLOCK, (end - 1) // BLOCK + 1):
            rng = np.random.default_rng([self.seed, contig, blk])
            codes = rng.integers(0, 4, size=BLOCK, dtype=np.uint8)
            b0 = blk * BLOCK
            lo, hi = max(begin, b0), min(end, b0 + BLOCK)
            out[lo - begin:hi - begin] = ACGT[codes[lo - b0:hi - b0]]
        pl = self.planted[contig]
        if pl:
            import bisect
            i = max(0, bisect.bisect_left(pl, (begin - 256, b"")))
            while i < len(pl) and pl[i][0] < end:
                pos, seq = pl[i]
                lo, hi = max(begin, pos), min(end, pos + len(seq))
                if lo < hi:
                    out[lo - begin:hi - begin] = np.frombuffer(seq, dtype=np.uint8)[lo - pos:hi - pos]
                i += 1
        for (nb, ne) in self.n_blocks[contig]:
            lo, hi = max(begin, nb), min(end, ne)
            if lo < hi:
                out[lo - begin:hi - begin] = ord("N")
        return out

    def contig(self, c):
        return self.range(c, 0, self.lengths[c])

    def contigs(self):
        return [(n, self.contig(i)) for i, n in enumerate(self.names)]

    def total(self):
        return sum(self.lengths)


def mutate_protospacer(rng, proto, n_edits):
    """0..n random edits (substitution / 1-bp insertion / 1-bp deletion; 60 % of indels inside a homopolymer run >= 2)."""
    s = bytearray(proto)
    for _ in range(n_edits):
        kind = rng.integers(0, 3)
        if kind == 0 or len(s) < 4:
            i = int(rng.integers(0, len(s)))
            s[i] = rng.choice([c for c in b"ACGT" if c != s[i]])
        else:
            runs = [i for i in range(1, len(s)) if s[i] == s[i - 1]]
            i = int(rng.choice(runs)) if (runs and rng.random() < 0.6) else int(rng.integers(1, len(s) - 1))
            if kind == 1:
                s.insert(i, s[i] if (runs and i in runs) else int(rng.choice(list(b"ACGT"))))
            else:
                del s[i]
    return bytes(s)


def plant_sites(rng, lengths, guides, n_sites, n_blocks, max_edits=5, pams=(b"AGG", b"TGG", b"CAG", b"GGG", b"TGA"), margin=200):
    """planted[contig] = [(pos, bytes)], non-overlapping, outside N blocks.  guides: list of 'PROTOSPACERpam' strings (3' PAM)."""
    planted = [[] for _ in lengths]
    taken = [set() for _ in lengths]                      # occupied 128-base bins
    blocked = []
    for c, blocks in enumerate(n_blocks):
        b = set()
        for (nb, ne) in blocks:
            if ne - nb <= (1 << 22):                      # big blocks are handled by interval test below
                b.update(range(max(0, nb - 64) >> 7, ((ne + 64) >> 7) + 1))
        blocked.append(b)
    big = [[(nb, ne) for (nb, ne) in blocks if ne - nb > (1 << 22)] for blocks in n_blocks]
    p_contig = np.array(lengths, dtype=np.float64) / float(sum(lengths))
    protos = ["".join(c for c in g if c.isupper()).encode() for g in guides]
    for k in range(n_sites):
        site = mutate_protospacer(rng, protos[k % len(protos)], int(rng.integers(0, max_edits + 1)))
        site += bytes(rng.integers(0, 4, size=int(rng.integers(0, 4))).astype(np.uint8).tolist()).translate(bytes.maketrans(bytes(range(4)), b"ACGT"))
        site += pams[int(rng.integers(0, len(pams)))]
        if rng.random() < 0.5:
            site = revcomp_bytes(site)
        for _try in range(100):
            c = int(rng.choice(len(lengths), p=p_contig))
            if lengths[c] < 2 * margin + 64:
                continue
            pos = int(rng.integers(margin, lengths[c] - margin - len(site)))
            bins = range((pos - 64) >> 7, ((pos + len(site) + 64) >> 7) + 1)
            if any(b in taken[c] or b in blocked[c] for b in bins):
                continue
            if any(pos < ne + 64 and nb - 64 < pos + len(site) for (nb, ne) in big[c]):
                continue
            taken[c].update(bins)
            planted[c].append((pos, site))
            break
    for lst in planted:
        lst.sort()
    return planted


def config1_genome(scale=1.0, n_sites=200, guides=(BASELINE_GUIDE,)):
    """BASELINE configs[0]: 10 Mbp, 4 contigs of 4/3/2/1 Mbp, seed 20260101, 10-kb N telomeres, one 50-kb N block in contig 1."""
    lengths = [int(4e6 * scale), int(3e6 * scale), int(2e6 * scale), int(1e6 * scale)]
    names = ["chr1", "chr2", "chr3", "chr4"]
    tel = max(100, int(10000 * scale))
    n_blocks = [[(0, tel), (l - tel, l)] for l in lengths]
    mid = lengths[0] // 2
    n_blocks[0].append((mid, mid + max(500, int(50000 * scale))))
    rng = np.random.default_rng(20260101)
    planted = plant_sites(rng, lengths, list(guides), n_sites, n_blocks)
    return Genome(names, lengths, 20260101, n_blocks, planted)


def hg38_like_genome(scale=1.0, guides=(BASELINE_GUIDE,), sites_per_guide=2000, seed=20260102):
    """BASELINE configs[2]: 24 contigs with hg38 primary-assembly lengths (x scale), N telomeres, a centromere-like block and 30 scattered N blocks per contig."""
    lengths = [max(20000, int(l * scale)) for l in HG38_LENGTHS]
    rng = np.random.default_rng(seed)
    n_blocks = []
    for l in lengths:
        tel = max(50, int(10000 * scale))
        blocks = [(0, tel), (l - tel, l)]
        cen = int(rng.integers(int(1e6 * scale) + 1, int(3e6 * scale) + 2))
        cs = int(rng.integers(l // 3, l // 2))
        blocks.append((cs, min(l - tel, cs + cen)))
        for _ in range(30):
            bl = int(rng.integers(max(1, int(1000 * scale)), max(2, int(50000 * scale))))
            bs = int(rng.integers(tel, max(tel + 1, l - tel - bl)))
            blocks.append((bs, bs + bl))
        n_blocks.append(blocks)
    planted = plant_sites(rng, lengths, list(guides), sites_per_guide * len(guides), n_blocks)
    return Genome(list(HG38_NAMES), lengths, seed, n_blocks, planted)


def random_guides(n, seed=20260103, length=20, pam="nrg"):
    """BASELINE configs[3]: n random 20-mers, homopolymers >= 5 rejected, with the given PAM suffix ('' for PAM-less)."""
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < n:
        s = "".join("ACGT"[i] for i in rng.integers(0, 4, size=length))
        if any(c * 5 in s for c in "ACGT"):
            continue
        out.append(s + pam)
    return out


def a2r_tasks(genome, guides, n, seed=20260105, near_fraction=0.5):
    """BASELINE configs[1]: rows (id, query, chrom, position): half at planted sites +/- U(-10, 10), half uniform random."""
    rng = np.random.default_rng(seed)
    sites = [(c, pos, len(seq)) for c, lst in enumerate(genome.planted) for (pos, seq) in lst]
    total = float(genome.total())
    tasks = []
    for i in range(n):
        q = guides[int(rng.integers(0, len(guides)))]
        if sites and rng.random() < near_fraction:
            c, pos, ln = sites[int(rng.integers(0, len(sites)))]
            p = pos + ln // 2 + int(rng.integers(-10, 11))
        else:
            c = int(rng.choice(len(genome.lengths), p=[l / total for l in genome.lengths]))
            p = int(rng.integers(1, genome.lengths[c] + 1))
        tasks.append(("t%d" % i, q, genome.names[c], max(1, min(genome.lengths[c], p))))
    return tasks


def synthetic_vcf(genome, contig_bases, n_records, seed=20260104, cluster_fraction=0.05):
    """BASELINE configs[4]: PrepareVcf-shaped records (PrepareVcf.scala:67-79): sorted in contig order, FILTER=PASS, INFO=AF only, no samples.
    85 % SNPs, 7.5 % insertions, 7.5 % deletions (1-10 bp), 2 % multi-allelic, 5 % of records in clusters of 2-6 within 30 bp."""
    rng = np.random.default_rng(seed)
    total = float(genome.total())
    lines = ["##fileformat=VCFv4.2", "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO"]
    per_contig = [max(1, int(round(n_records * l / total))) for l in genome.lengths]
    rid = 0
    for c, n in enumerate(per_contig):
        bases = contig_bases[c]
        L = len(bases)
        pos_set = set()
        while len(pos_set) < n:
            p = int(rng.integers(20, max(21, L - 40)))
            pos_set.add(p)
            if rng.random() < cluster_fraction:
                for _ in range(int(rng.integers(1, 6))):
                    pos_set.add(min(L - 40, p + int(rng.integers(1, 30))))
        for p in sorted(pos_set):
            ref1 = chr(bases[p - 1])
            if ref1 not in "ACGT":
                continue
            kind = rng.random()
            if kind < 0.85:
                alts = [x for x in "ACGT" if x != ref1]
                rng.shuffle(alts)
                k = 2 if rng.random() < 0.02 else 1
                ref, alt = ref1, alts[:k]
            elif kind < 0.925:
                ins = "".join("ACGT"[i] for i in rng.integers(0, 4, size=int(rng.integers(1, 11))))
                ref, alt = ref1, [ref1 + ins]
            else:
                dl = int(rng.integers(1, 11))
                seg = bytes(bases[p - 1:p + dl]).decode()
                if "N" in seg or len(seg) < dl + 1:
                    continue
                ref, alt = seg, [ref1]
            afs = ",".join("%.4g" % max(0.01, float(rng.random()) ** 3) for _ in alt)
            rid += 1
            lines.append("\t".join([genome.names[c], str(p), "rs%d" % rid, ref, ",".join(alt), ".", "PASS", "AF=" + afs]))
    return "\n".join(lines) + "\n"


def synthetic_vcf_fast(genome, arrays, ranges, n_records, seed=20260104, cluster_fraction=0.05):
    """The shape of synthetic_vcf (BASELINE configs[4]) for millions of records: positions and alleles are drawn with numpy, only the text is built in a
    Python loop.  arrays[c] holds bases [ranges[c][0], ranges[c][1]) of contig c (None / empty range: no records there); n_records counts the WHOLE genome,
    each contig range gets its share.  85 % SNPs, 7.5 % insertions, 7.5 % deletions (1-10 bp), ~2 % multi-allelic SNPs, ~5 % of records inside 30-bp clusters."""
    total = float(genome.total())
    lines = ["##fileformat=VCFv4.2", "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO"]
    rid = 0
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    code = np.full(256, 255, dtype=np.uint8)
    for i, b in enumerate(b"ACGT"):
        code[b] = i
    for c, (lo, hi) in enumerate(ranges):
        a = arrays[c]
        if a is None or hi - lo < 200:
            continue
        rng = np.random.default_rng([seed, c, int(lo)])
        n = max(1, int(round(n_records * (hi - lo) / total)))
        pos = rng.integers(lo + 20, hi - 40, size=int(n * (1 - cluster_fraction)) + 1)           # 1-based POS = 0-based index + 1 below
        k = int(n * cluster_fraction)
        if k:
            seeds = rng.choice(pos, size=max(1, k // 3))
            pos = np.concatenate([pos, np.repeat(seeds, 3)[:k] + rng.integers(1, 30, size=min(k, 3 * len(seeds)))])
        pos = np.unique(np.clip(pos, lo + 20, hi - 41))
        ref_code = code[a[pos - lo]]                                                              # base at 0-based index pos (POS = pos + 1)
        pos = pos[ref_code < 4]; ref_code = ref_code[ref_code < 4]
        m = pos.size
        kind = rng.random(m)
        alt1 = (ref_code + rng.integers(1, 4, size=m)) % 4
        alt2 = (ref_code + 1 + (alt1 - ref_code) % 4 % 3) % 4                                       # a second, different ALT for the multi-allelic SNPs
        multi = rng.random(m) < 0.02
        ins_len = rng.integers(1, 11, size=m)
        ins_bases = acgt[rng.integers(0, 4, size=(m, 10))]
        del_len = rng.integers(1, 11, size=m)
        af = np.maximum(0.01, rng.random((m, 2)) ** 3)
        name = genome.names[c]
        for j in range(m):
            p = int(pos[j]); r1 = "ACGT"[ref_code[j]]
            if kind[j] < 0.85:
                if multi[j] and alt2[j] != alt1[j] and alt2[j] != ref_code[j]:
                    ref, alt, afs = r1, "ACGT"[alt1[j]] + "," + "ACGT"[alt2[j]], "%.4g,%.4g" % (af[j, 0], af[j, 1])
                else:
                    ref, alt, afs = r1, "ACGT"[alt1[j]], "%.4g" % af[j, 0]
            elif kind[j] < 0.925:
                ref, alt, afs = r1, r1 + ins_bases[j, :ins_len[j]].tobytes().decode(), "%.4g" % af[j, 0]
            else:
                seg = a[p - lo:p - lo + del_len[j] + 1].tobytes()
                if b"N" in seg or len(seg) < del_len[j] + 1:
                    continue
                ref, alt, afs = seg.decode(), r1, "%.4g" % af[j, 0]
            rid += 1
            lines.append("%s\t%d\trs%d\t%s\t%s\t.\tPASS\tAF=%s" % (name, p + 1, rid, ref, alt, afs))
    return "\n".join(lines) + "\n"
