// cal_core.cuh — per-thread building blocks of the CALITAS B200 engine (device code; also compiled for the host by the
// test-only hostsim build, tests/hostsim/, so the logic can be checked without a GPU).
//
// Reference behaviour implemented here (files under calitas/src/main/scala/com/editasmedicine/aligner/):
//   scorer              SequentialGuideAligner.scala:139-153, 192-208
//   glocal DP+traceback fgbio 2.0.0 alignment.Aligner (Mode.Glocal) as called at SequentialGuideAligner.scala:261,278,295,299
//   PAM extension       SequentialGuideAligner.scala:433-492
//   record coordinates  SequentialGuideAligner.scala:263-310, 505-524; GuideAlignment.scala:21-31
//   per-window filter   SequentialGuideAligner.scala:315-322; GuideAlignment.scala:119-129
//   genome-wide sweep   SearchReference.scala:653-675; ReferenceHit.scala:135-144
#pragma once
#include <stdint.h>
#include "../../include/calitas_b200.h"

#if defined(__CUDACC__) && !defined(CAL_HOSTSIM)
#define CAL_HD __host__ __device__ __forceinline__
#define CAL_D __device__ __forceinline__
#else
#define CAL_HD inline
#define CAL_D inline
#endif

namespace cal {

// ---- target codes ------------------------------------------------------------------------------------------------
// One 4-bit code per reference base: the IUPAC base set (A=1, C=2, G=4, T/U=8, ambiguity codes = unions), with two
// specials: 15 = upper-case 'N' (never matches, SequentialGuideAligner.scala:144; trimmed from window ends,
// SearchReference.scala:58-59) and 0 = 'n' or any non-IUPAC byte (never matches, not trimmed).
enum { CODE_N = 15 };

CAL_HD uint32_t iupac_set(uint8_t b) {
  switch (b & 0xDF) {  // case-insensitive for letters
    case 'A': return 1;  case 'C': return 2;  case 'G': return 4;  case 'T': return 8;  case 'U': return 8;
    case 'M': return 3;  case 'R': return 5;  case 'W': return 9;  case 'S': return 6;  case 'Y': return 10; case 'K': return 12;
    case 'V': return 7;  case 'H': return 11; case 'D': return 13; case 'B': return 14; case 'N': return 15;
    default:  return 0;
  }
}
CAL_HD uint32_t target_code(uint8_t b) {
  if (b == 'N') return CODE_N;
  uint32_t s = iupac_set(b);          // 0 for any non-IUPAC byte
  return s == 15 ? 0u : s;            // 'n'
}
// complement of a base set: A<->T, C<->G = reverse the 4 bits; 0 and 15 map to themselves
CAL_HD uint32_t comp_code(uint32_t c) { return ((c & 1) << 3) | ((c & 2) << 1) | ((c & 4) >> 1) | ((c & 8) >> 3); }
// does a query base set pair with a target code?  (scorePairing: N never matches)
CAL_HD bool pairs(uint32_t qset, uint32_t tcode) { return tcode != CODE_N && (qset & tcode) != 0; }

// ---- hit records ---------------------------------------------------------------------------------------------------------------------
// HitX: one alignment with every field spelled out (host-side code and the generic device paths work on it).  The device pipeline and the
// C ABI move the packed form of include/calitas_b200.h instead: `rw` 32-bit words per record (8 = calitas_hit, 16 = calitas_hit_wide), five
// header words + 2-bit ops.  Unused op bits are zero, so gap and edit counts are population counts over the op words.
struct HitX {
  int32_t guide_idx, pam_idx, contig_idx, task_idx, start_offset, end_offset, guide_start_offset, guide_end_offset, score;
  uint8_t strand, n_ops, gap_bases, edits;
  uint32_t ops[CALITAS_MAX_OPS / 16];
};
CAL_HD int popc32(uint32_t v) {
#if defined(__CUDA_ARCH__)
  return __popc(v);
#else
  return __builtin_popcount(v);
#endif
}
CAL_HD uint32_t rec_make_where(int32_t guide_idx, int32_t contig_idx, bool neg) { return (uint32_t)guide_idx | ((uint32_t)(contig_idx + 1) << 13) | (neg ? 0x80000000u : 0u); }
CAL_HD uint32_t rec_make_shape(int n_ops, int span, int lead, int trail, int pam_idx) {
  return (uint32_t)n_ops | ((uint32_t)span << 8) | ((uint32_t)lead << 16) | ((uint32_t)trail << 22) | ((uint32_t)(pam_idx + 1) << 28);
}
CAL_HD int32_t rec_start(const uint32_t* r) { return (int32_t)r[0]; }
CAL_HD int32_t rec_task(const uint32_t* r) { return (int32_t)r[1]; }
CAL_HD int32_t rec_score(const uint32_t* r) { return (int32_t)r[2]; }
CAL_HD int32_t rec_guide(const uint32_t* r) { return (int32_t)(r[3] & 0x1FFFu); }
CAL_HD int32_t rec_contig(const uint32_t* r) { return (int32_t)((r[3] >> 13) & 0x3FFFFu) - 1; }
CAL_HD uint32_t rec_neg(const uint32_t* r) { return r[3] >> 31; }
CAL_HD int32_t rec_nops(const uint32_t* r) { return (int32_t)(r[4] & 0xFFu); }
CAL_HD int32_t rec_span(const uint32_t* r) { return (int32_t)((r[4] >> 8) & 0xFFu); }
CAL_HD int32_t rec_end(const uint32_t* r) { return (int32_t)r[0] + rec_span(r); }
CAL_HD int32_t rec_gstart(const uint32_t* r) { return (int32_t)r[0] + (int32_t)((r[4] >> 16) & 0x3Fu); }
CAL_HD int32_t rec_gend(const uint32_t* r) { return rec_end(r) - (int32_t)((r[4] >> 22) & 0x3Fu); }
CAL_HD int32_t rec_pam(const uint32_t* r) { return (int32_t)(r[4] >> 28) - 1; }
CAL_HD int32_t rec_gap_bases(const uint32_t* r, int rw) { int c = 0; for (int k = CALITAS_HIT_HEADER_WORDS; k < rw; ++k) c += popc32(r[k] & 0xAAAAAAAAu); return c; }
CAL_HD int32_t rec_edits(const uint32_t* r, int rw) { int c = 0; for (int k = CALITAS_HIT_HEADER_WORDS; k < rw; ++k) c += popc32((r[k] | (r[k] >> 1)) & 0x55555555u); return c; }
CAL_HD void pack_hit(const HitX& h, uint32_t* r, int rw) {
  r[0] = (uint32_t)h.start_offset; r[1] = (uint32_t)h.task_idx; r[2] = (uint32_t)h.score; r[3] = rec_make_where(h.guide_idx, h.contig_idx, h.strand == '-');
  r[4] = rec_make_shape(h.n_ops, h.end_offset - h.start_offset, h.guide_start_offset - h.start_offset, h.end_offset - h.guide_end_offset, h.pam_idx);
  for (int k = CALITAS_HIT_HEADER_WORDS; k < rw; ++k) { const int w = k - CALITAS_HIT_HEADER_WORDS; r[k] = w < CALITAS_MAX_OPS / 16 ? h.ops[w] : 0u; }
}
CAL_HD void unpack_hit(const uint32_t* r, int rw, HitX& h) {
  h.start_offset = rec_start(r); h.task_idx = rec_task(r); h.score = rec_score(r); h.guide_idx = rec_guide(r); h.contig_idx = rec_contig(r); h.strand = (uint8_t)(rec_neg(r) ? '-' : '+');
  h.n_ops = (uint8_t)rec_nops(r); h.end_offset = rec_end(r); h.guide_start_offset = rec_gstart(r); h.guide_end_offset = rec_gend(r); h.pam_idx = rec_pam(r);
  h.gap_bases = (uint8_t)rec_gap_bases(r, rw); h.edits = (uint8_t)rec_edits(r, rw);
  for (int w = 0; w < CALITAS_MAX_OPS / 16; ++w) h.ops[w] = w + CALITAS_HIT_HEADER_WORDS < rw ? r[w + CALITAS_HIT_HEADER_WORDS] : 0u;
}
// 32-byte records (rw = 8) hold 48 alignment columns, 64-byte records 176
CAL_HD int rec_words_for(int max_columns) { return max_columns <= 16 * (CALITAS_HIT_WORDS - CALITAS_HIT_HEADER_WORDS) ? CALITAS_HIT_WORDS : CALITAS_HIT_WIDE_WORDS; }

// ---- scores (SequentialGuideAligner.scala:192-208, 213) ---------------------------------------------------------
struct Scores {
  int32_t match, mismatch, pam_match, pam_mismatch, query_gap /* cigar D */, target_gap /* cigar I */, worst_guide_diff;
  int32_t abs_mm, abs_genome_gap, abs_guide_gap;
};
CAL_HD int32_t iabs(int32_t v) { return v < 0 ? -v : v; }
CAL_HD Scores make_scores(const calitas_costs& c) {
  Scores s;
  s.abs_mm = iabs(c.mismatch_net_cost); s.abs_genome_gap = iabs(c.genome_gap_net_cost); s.abs_guide_gap = iabs(c.guide_gap_net_cost);
  s.match = s.abs_mm / 2;
  s.mismatch = -(s.abs_mm - s.match);
  s.query_gap = -s.abs_guide_gap;
  s.target_gap = -s.abs_genome_gap + s.match;
  s.pam_match = iabs(c.pam_mismatch_net_cost) / 2;
  s.pam_mismatch = -(iabs(c.pam_mismatch_net_cost) - s.pam_match);
  int32_t w = -s.abs_mm; if (-s.abs_genome_gap < w) w = -s.abs_genome_gap; if (-s.abs_guide_gap < w) w = -s.abs_guide_gap;
  s.worst_guide_diff = w;
  return s;
}

// ---- per-guide device descriptor ------------------------------------------------------------------------------------
// The DP always runs with the PAM on the right (SequentialGuideAligner.scala:255-259): `q` is the protospacer for a
// 3' PAM and its reverse complement for a 5' PAM; `pam` likewise.  dir 0 scans the target as given, dir 1 scans its
// reverse complement; strand = (dir ^ five_prime) ? '-' : '+'.
struct GuideSpec {
  uint32_t peq[2][16];                       // Myers match masks per scan direction and target code, top-aligned (row i = bit 32-lp+i), low bits all 1
  uint8_t  q[CALITAS_MAX_PROTOSPACER];       // base set per DP-query row
  uint16_t qmask[CALITAS_MAX_PROTOSPACER];   // per row: bit c set iff the row's base pairs with target code c
  uint8_t  pam[CALITAS_MAX_PAMS][CALITAS_MAX_PAM_LEN];
  uint8_t  pam_len[CALITAS_MAX_PAMS];
  int32_t  lp, n_pams, five_prime;
  int32_t  d, p, g;                          // maxGuideDiffs, maxPamDiffs, maxGapsBetweenGuideAndPam
  int32_t  max_tot_filter;                   // d + g + p (SequentialGuideAligner.scala:249)
  int32_t  max_total_diffs, max_overlap;     // post filter (:317)
  int32_t  min_score;                        // :239-243
  int32_t  k_edits;                          // candidate threshold of the bit-parallel scan: unit edits <= k_edits is necessary for score >= min_score
  int32_t  span;                             // max target columns any co-optimal alignment of an accepted end column can cover
  int32_t  band_k;                           // max gap bases of either kind on such an alignment (== k_edits unless k_edits was capped at lp): decides the banded kernels
  int32_t  slots;                            // alignment slots per candidate end column = max(1, n_pams)
  int32_t  max_cols;                         // most alignment columns a hit of this guide can have (decides the record size of a call)
};

// ---- bit-parallel candidate scan (Myers/Hyyro, semi-global: free target start) ----------------------------------------
// State of one (window, direction, guide) scan.  Column update for target code `eq` = peq[dir][code].
struct MyersState { uint32_t pv, mv; int32_t score; };
CAL_HD void myers_init(MyersState& s, int lp) {
  s.pv = lp >= 32 ? 0xFFFFFFFFu : ~(0xFFFFFFFFu >> lp);   // D[i][0] = i : leading insertions at the window's left edge
  s.mv = 0; s.score = lp;
}
// 7 LOP3 + 2 LEA.HI on the ALU pipe, 3 IMAD.IADD on the FMA pipe (the compiler turns the add and both x+x shifts into IMADs).
// Measured alternatives that were slower on B200: IMAD.WIDE (x*2 gives shift and top bit at once) runs well below IMAD rate.
CAL_HD void myers_step(MyersState& s, uint32_t eq) {
  uint32_t xv = eq | s.mv;
  uint32_t xh = (((eq & s.pv) + s.pv) ^ s.pv) | eq;
  uint32_t ph = s.mv | ~(xh | s.pv);
  uint32_t mh = s.pv & xh;
  s.score += (int32_t)(ph >> 31) - (int32_t)(mh >> 31);
  ph <<= 1; mh <<= 1;
  s.pv = mh | ~(xv | ph);
  s.mv = ph & xv;
}

// ---- banded glocal DP with traceback for one candidate end column ------------------------------------------------------
enum { OP_EQ = 0, OP_X = 1, OP_I = 2, OP_D = 3 };
enum { TR_LEFT = 0, TR_UP = 1, TR_DIAG = 2, TR_DONE = 3 };
const int32_t NEG_SCORE = -(1 << 28);
const int MAX_SPAN = 2 * CALITAS_MAX_PROTOSPACER + 2;     // columns of the DP rectangle
const int MAX_GUIDE_OPS = CALITAS_MAX_PROTOSPACER + MAX_SPAN;

struct GuideAln {          // one fgbio `Alignment` of the protospacer, in DP orientation
  int32_t score;
  int32_t t_start;         // 1-based first target column (targetStart)
  int32_t t_end;           // 1-based last target column (targetEnd) == the candidate end column
  int32_t n_ops;
  int32_t diffs;           // non-'=' columns (SequentialGuideAligner.scala:442)
  int32_t terminal_gap;    // length of a trailing I or D run (:452)
  int32_t terminal_d;      // length of a trailing D run (columns after the last guide base)
  uint8_t ops[MAX_GUIDE_OPS];   // in DP orientation, first column first
};

// Target accessor: code of DP column p (1-based) of the scanned target, already complemented for dir 1.
// `Fetch` is a functor: uint32_t operator()(int p).
template <class Fetch>
CAL_HD bool band_align(const GuideSpec& g, const Scores& sc, Fetch fetch, int32_t j, GuideAln& out, uint8_t* trace /* (lp+1)*(MAX_SPAN+1) */) {
  const int n = g.lp;
  const int jlo = j - g.span > 0 ? j - g.span : 0;
  const int width = j - jlo;                         // local columns 0..width
  const int W = MAX_SPAN + 1;
  int32_t pd[MAX_SPAN + 1], pl[MAX_SPAN + 1], pu[MAX_SPAN + 1];   // previous row
  int32_t cd[MAX_SPAN + 1], cl[MAX_SPAN + 1], cu[MAX_SPAN + 1];   // current row
  uint8_t tc[MAX_SPAN + 1];
  for (int c = 1; c <= width; ++c) { uint32_t t = fetch(jlo + c); tc[c] = (uint8_t)t; }
  for (int c = 0; c <= width; ++c) { pd[c] = 0; pl[c] = 0; pu[c] = 0; }      // row 0: free leading target
  const int32_t gI = sc.target_gap, gD = sc.query_gap;
  for (int i = 1; i <= n; ++i) {
    const uint32_t qs = g.q[i - 1];
    cd[0] = NEG_SCORE; cl[0] = NEG_SCORE; cu[0] = pu[0] + gI;                // column 0: leading insertions only
    trace[i * W] = (uint8_t)((i == 1 ? TR_DIAG : TR_UP) << 2);
    for (int c = 1; c <= width; ++c) {
      uint32_t tr;
      {   // Diagonal cell: predecessor Diagonal > Left > Up on ties
        int32_t add = pairs(qs, tc[c]) ? sc.match : sc.mismatch;
        int32_t d = pd[c - 1], l = pl[c - 1], u = pu[c - 1];
        if (d >= l && d >= u) { cd[c] = d + add; tr = TR_DIAG; } else if (l >= u) { cd[c] = l + add; tr = TR_LEFT; } else { cd[c] = u + add; tr = TR_UP; }
      }
      {   // Up cell (cigar I): from Diagonal or Up
        int32_t d = pd[c] + gI, u = pu[c] + gI;
        if (d >= u) { cu[c] = d; tr |= TR_DIAG << 2; } else { cu[c] = u; tr |= TR_UP << 2; }
      }
      {   // Left cell (cigar D): from Diagonal, Left or Up
        int32_t d = cd[c - 1] + gD, l = cl[c - 1] + gD, u = cu[c - 1] + gD;
        if (d >= l && d >= u) { cl[c] = d; tr |= TR_DIAG << 4; } else if (l >= u) { cl[c] = l; tr |= TR_LEFT << 4; } else { cl[c] = u; tr |= TR_UP << 4; }
      }
      trace[i * W + c] = (uint8_t)tr;
    }
    for (int c = 0; c <= width; ++c) { pd[c] = cd[c]; pl[c] = cl[c]; pu[c] = cu[c]; }
  }
  int32_t best = pd[width]; int dir = TR_DIAG;
  if (pl[width] > best) { best = pl[width]; dir = TR_LEFT; }
  if (pu[width] > best) { best = pu[width]; dir = TR_UP; }
  if (best < g.min_score) return false;
  // traceback
  int ci = n, cc = width, cdir = dir, nrev = 0;
  uint8_t rev[MAX_GUIDE_OPS];
  for (;;) {
    int next;
    if (ci == 0) next = TR_DONE;
    else { uint32_t t = trace[ci * W + cc]; next = cdir == TR_DIAG ? (t & 3) : (cdir == TR_UP ? ((t >> 2) & 3) : ((t >> 4) & 3)); }
    if (next == TR_DONE) break;
    if (cdir == TR_DIAG) { rev[nrev++] = pairs(g.q[ci - 1], tc[cc]) ? OP_EQ : OP_X; --ci; --cc; }
    else if (cdir == TR_LEFT) { rev[nrev++] = OP_D; --cc; }
    else { rev[nrev++] = OP_I; --ci; }
    cdir = next;
  }
  out.score = best; out.t_start = jlo + cc + 1; out.t_end = j; out.n_ops = nrev; out.diffs = 0;
  for (int k = 0; k < nrev; ++k) { uint8_t o = rev[nrev - 1 - k]; out.ops[k] = o; if (o != OP_EQ) ++out.diffs; }
  out.terminal_gap = 0; out.terminal_d = 0;
  if (nrev > 0 && out.ops[nrev - 1] >= OP_I) {
    uint8_t o = out.ops[nrev - 1]; int k = nrev; while (k > 0 && out.ops[k - 1] == o) { --k; ++out.terminal_gap; }
    if (o == OP_D) out.terminal_d = out.terminal_gap;
  }
  return true;
}

// Register-resident variant for the common case k_edits <= KB: only the 2*KB+1 diagonals around the end cell's diagonal can hold an
// alignment with <= k_edits edits, so one DP row is 2*KB+1 cells kept in registers (compile-time indices).  Values outside the band count as
// unreachable; every co-optimal path of an accepted end cell lies inside it, hence scores, tie-breaks and traceback are identical to
// band_align's ("exactness of the band").
//
// Tagged scores.  band_align picks a cell's predecessor matrix with `if (d >= l && d >= u) DIAG else if (l >= u) LEFT else UP` (and
// `d >= u` for the Up matrix): the best score, ties to Diagonal, then Left, then Up.  Here every value is stored as 4 * score + tag with
// tag 3 in the Diagonal matrix, 2 in Left, 1 in Up, so ONE integer max (3-input VIMNMX3) returns the winning score and, in its two low
// bits, the matrix it came from under exactly that tie order: a > b implies 4a + s > 4b + t for any tags, a == b leaves the tags to
// decide.  The same addend applies to all predecessors of a cell, so it is added after the max.  The 2-bit winners go into one trace word
// per matrix and row with a funnel shift each (cell t of a row ends up at bits 32 - 2B + 2t), the match flags into a fourth word.  About
// 18 instructions per cell instead of 35 (compare / select chains, field packing); cells, scores and traces are those of band_align.
enum { TG_DONE = 0, TG_UP = 1, TG_LEFT = 2, TG_DIAG = 3 };
CAL_HD int32_t max3_s32(int32_t a, int32_t b, int32_t c) {
#if defined(__CUDA_ARCH__)
  return __vimax3_s32(a, b, c);
#else
  const int32_t m = a > b ? a : b; return m > c ? m : c;
#endif
}
CAL_HD uint32_t shift_in_low_bits(uint32_t acc, uint32_t v, int nbits) {     // (acc >> nbits) | (v << (32 - nbits)): v's low bits enter at the top
#if defined(__CUDA_ARCH__)
  return __funnelshift_r(acc, v, nbits);
#else
  return (uint32_t)(((((uint64_t)v) << 32) | acc) >> nbits);
#endif
}
template <int KB, class Fetch>
CAL_HD bool band_align_k(const GuideSpec& g, const Scores& sc, Fetch fetch, int32_t j, GuideAln& out) {
  constexpr int B = 2 * KB + 1;
  static_assert(2 * B <= 32, "one trace word per matrix and row");
  const int n = g.lp;
  const int base = j - n - KB;                  // cell (i, t) is target column c = i + base + t
  const int32_t NEG4 = 4 * NEG_SCORE;
  int32_t d[B], l[B], u[B];                     // previous row, tagged
  uint32_t trd[CALITAS_MAX_PROTOSPACER + 1], tru[CALITAS_MAX_PROTOSPACER + 1], trl[CALITAS_MAX_PROTOSPACER + 1], trm[CALITAS_MAX_PROTOSPACER + 1];
#pragma unroll
  for (int t = 0; t < B; ++t) { const int c = base + t; const int32_t v = (c >= 0 && c <= j) ? 0 : NEG4; d[t] = v + TG_DIAG; l[t] = v + TG_LEFT; u[t] = v + TG_UP; }
  uint64_t win = 0;                             // target codes of the current row's band, one nibble per diagonal
#pragma unroll
  for (int t = 0; t < B; ++t) { const int c = 1 + base + t; const uint64_t code = (c >= 1 && c <= j) ? fetch(c) : 0u; win |= code << (4 * t); }
  const int32_t gI4 = 4 * sc.target_gap, gD4 = 4 * sc.query_gap, mis4 = 4 * sc.mismatch, dmatch4 = 4 * (sc.match - sc.mismatch);
  for (int i = 1; i <= n; ++i) {
    const uint32_t qm = g.qmask[i - 1];
    uint32_t wd = 0, wu = 0, wl = 0, wm = 0;
    int32_t left_d = NEG4 + TG_DIAG, left_l = NEG4 + TG_LEFT, left_u = NEG4 + TG_UP;
#pragma unroll
    for (int t = 0; t < B; ++t) {
      const uint32_t code = (uint32_t)(win >> (4 * t)) & 15u;
      const uint32_t mt = (qm >> code) & 1u;
      const int32_t add4 = mis4 + (int32_t)mt * dmatch4;
      const int32_t md = max3_s32(d[t], l[t], u[t]);                                  // Diagonal: predecessor (i-1, c-1) is the same diagonal
      const int32_t mu = t + 1 < B ? (d[t + 1] > u[t + 1] ? d[t + 1] : u[t + 1]) : NEG4 + TG_DIAG;   // Up: (i-1, c) is diagonal t+1
      const int32_t ml = max3_s32(left_d, left_l, left_u);                            // Left: (i, c-1) is diagonal t-1
      const int32_t nd = (md | 3) + add4, nu = ((mu & ~3) | TG_UP) + gI4, nl = ((ml & ~3) | TG_LEFT) + gD4;
      wd = shift_in_low_bits(wd, (uint32_t)md, 2); wu = shift_in_low_bits(wu, (uint32_t)mu, 2); wl = shift_in_low_bits(wl, (uint32_t)ml, 2); wm = shift_in_low_bits(wm, mt, 1);
      d[t] = nd; u[t] = nu; l[t] = nl; left_d = nd; left_l = nl; left_u = nu;
    }
    trd[i] = wd; tru[i] = wu; trl[i] = wl; trm[i] = wm;
    const int cn = (i + 1) + base + (B - 1);                   // slide the band one column to the right
    const uint64_t code = (cn >= 1 && cn <= j) ? fetch(cn) : 0u;
    win = (win >> 4) | (code << (4 * (B - 1)));
  }
  const int32_t mbest = max3_s32(d[KB], l[KB], u[KB]);
  const int32_t best = mbest >> 2;
  if (best < g.min_score) return false;
  int ci = n, ct = KB, cdir = mbest & 3, nrev = 0;
  uint8_t rev[MAX_GUIDE_OPS];
  for (;;) {
    if (ci == 0) break;
    const int sh = 32 - 2 * B + 2 * ct;
    const int next = (int)(((cdir == TG_DIAG ? trd[ci] : (cdir == TG_UP ? tru[ci] : trl[ci])) >> sh) & 3u);
    if (cdir == TG_DIAG) { rev[nrev++] = ((trm[ci] >> (32 - B + ct)) & 1u) ? OP_EQ : OP_X; --ci; }
    else if (cdir == TG_LEFT) { rev[nrev++] = OP_D; --ct; }
    else { rev[nrev++] = OP_I; --ci; ++ct; }
    if (ct < 0 || ct >= B) return false;                       // cannot happen for an accepted end cell; keeps indexing safe
    cdir = next;
  }
  out.score = best; out.t_start = ci + base + ct + 1; out.t_end = j; out.n_ops = nrev; out.diffs = 0;
  for (int k = 0; k < nrev; ++k) { uint8_t o = rev[nrev - 1 - k]; out.ops[k] = o; if (o != OP_EQ) ++out.diffs; }
  out.terminal_gap = 0; out.terminal_d = 0;
  if (nrev > 0 && out.ops[nrev - 1] >= OP_I) {
    uint8_t o = out.ops[nrev - 1]; int k = nrev; while (k > 0 && out.ops[k - 1] == o) { --k; ++out.terminal_gap; }
    if (o == OP_D) out.terminal_d = out.terminal_gap;
  }
  return true;
}

// Whole-range variant for a GROUP of candidate end columns of one (window, strand), used where nearly every column is a candidate
// (alignBest / alignToRefBest, SequentialGuideAligner.scala:333-345,402-418: d = protospacer length): columns jlo+1 .. last candidate are
// filled ONCE — fgbio fills its matrices once per window too — column-major, with the previous column of all three matrices in registers,
// and each candidate column is traced back as soon as it is complete.  Cells, tie-breaks and tracebacks are those of band_align; every
// band_align rectangle [j - span, j] is contained in the range filled here, and both equal the whole-window DP on accepted end cells.
// `col_at(k)` = k-th candidate column (ascending), `emit(k, aln)` receives each accepted alignment.  Requires last - jlo <= GW.
template <int GW, class Fetch, class ColAt, class Emit>
CAL_HD void band_align_group(const GuideSpec& g, const Scores& sc, Fetch fetch, int n_cols, ColAt col_at, Emit emit, uint8_t* trace /* (MAX_PROTOSPACER+1)*(GW+1) */) {
  constexpr int R = CALITAS_MAX_PROTOSPACER, TW = GW + 1;
  const int n = g.lp;
  const int first = col_at(0), last = col_at(n_cols - 1);
  const int jlo = first - g.span > 0 ? first - g.span : 0;
  // tagged scores as in band_align_k: 4 * score + (3 Diagonal, 2 Left, 1 Up); one integer max picks score and predecessor matrix in the reference's tie order
  const int32_t gI4 = 4 * sc.target_gap, gD4 = 4 * sc.query_gap, mis4 = 4 * sc.mismatch, dmatch4 = 4 * (sc.match - sc.mismatch), NEG4 = 4 * NEG_SCORE;
  int32_t D[R + 1], L[R + 1], U[R + 1];          // previous column; compile-time indices only -> registers
  uint32_t qm[R / 2];
#pragma unroll
  for (int i = 0; i < R / 2; ++i) qm[i] = (uint32_t)g.qmask[2 * i] | ((uint32_t)g.qmask[2 * i + 1] << 16);
#pragma unroll
  for (int i = 0; i <= R; ++i) { D[i] = (i == 0 ? 0 : NEG4) + TG_DIAG; L[i] = (i == 0 ? 0 : NEG4) + TG_LEFT; U[i] = i * gI4 + TG_UP; }      // local column 0: leading insertions only
  for (int i = 1; i <= n; ++i) trace[i * TW] = (uint8_t)((i == 1 ? TG_DIAG : TG_UP) << 2);
  int k = 0, next_col = first;
  for (int c = 1; c <= last - jlo; ++c) {
    const uint32_t code = fetch(jlo + c);
    int32_t dgD = TG_DIAG, dgL = TG_LEFT, dgU = TG_UP;   // row i-1 of the previous column (row 0 is all zero: free leading target)
    int32_t upD = TG_DIAG, upU = TG_UP;                  // row i-1 of this column
    int32_t lastD = TG_DIAG, lastL = TG_LEFT, lastU = TG_UP;
#pragma unroll
    for (int i = 1; i <= R; ++i) {
      if (i <= n) {
        const int32_t tD = D[i], tL = L[i], tU = U[i];
        const uint32_t mt = (((qm[(i - 1) >> 1] >> (((i - 1) & 1) * 16)) & 0xFFFFu) >> code) & 1u;
        const int32_t add4 = mis4 + (int32_t)mt * dmatch4;
        const int32_t md = max3_s32(dgD, dgL, dgU), mu = upD > upU ? upD : upU, ml = max3_s32(tD, tL, tU);
        const int32_t nd = (md | 3) + add4, nu = ((mu & ~3) | TG_UP) + gI4, nl = ((ml & ~3) | TG_LEFT) + gD4;
        const uint32_t cell = ((uint32_t)md & 3u) + ((uint32_t)mu & 3u) * 4u + ((uint32_t)ml & 3u) * 16u + mt * 64u;
        D[i] = nd; L[i] = nl; U[i] = nu; dgD = tD; dgL = tL; dgU = tU; upD = nd; upU = nu;
        trace[i * TW + c] = (uint8_t)cell;
        if (i == n) { lastD = nd; lastL = nl; lastU = nu; }
      }
    }
    if (jlo + c != next_col) continue;
    const int j = next_col, kk = k;
    ++k; next_col = k < n_cols ? col_at(k) : -1;
    const int32_t mbest = max3_s32(lastD, lastL, lastU);
    const int32_t best = mbest >> 2;
    if (best < g.min_score) continue;
    GuideAln out;
    int ci = n, cc = c, cdir = mbest & 3, nrev = 0;
    uint8_t rev[MAX_GUIDE_OPS];
    for (;;) {
      if (ci == 0 || nrev >= MAX_GUIDE_OPS) break;
      const uint32_t cell = trace[ci * TW + cc];
      const int next = (int)(cdir == TG_DIAG ? (cell & 3) : (cdir == TG_UP ? ((cell >> 2) & 3) : ((cell >> 4) & 3)));
      if (cdir == TG_DIAG) { rev[nrev++] = (cell >> 6) ? OP_EQ : OP_X; --ci; --cc; }
      else if (cdir == TG_LEFT) { rev[nrev++] = OP_D; --cc; }
      else { rev[nrev++] = OP_I; --ci; }
      cdir = next;
    }
    out.score = best; out.t_start = jlo + cc + 1; out.t_end = j; out.n_ops = nrev; out.diffs = 0;
    for (int q = 0; q < nrev; ++q) { uint8_t o = rev[nrev - 1 - q]; out.ops[q] = o; if (o != OP_EQ) ++out.diffs; }
    out.terminal_gap = 0; out.terminal_d = 0;
    if (nrev > 0 && out.ops[nrev - 1] >= OP_I) {
      uint8_t o = out.ops[nrev - 1]; int q = nrev; while (q > 0 && out.ops[q - 1] == o) { --q; ++out.terminal_gap; }
      if (o == OP_D) out.terminal_d = out.terminal_gap;
    }
    emit(kk, out);
  }
}

// ---- 2-bit op packing ------------------------------------------------------------------------------------------------------
CAL_HD void ops_set(uint32_t* words, int idx, uint32_t op) { words[idx >> 4] |= op << ((idx & 15) * 2); }
CAL_HD uint32_t ops_get(const uint32_t* words, int idx) { return (words[idx >> 4] >> ((idx & 15) * 2)) & 3u; }

// ---- PAM extension + record (SequentialGuideAligner.scala:433-492, 505-524, 263-310) ---------------------------------------
// Window geometry in contig (or task) coordinates: the scanned target is bases [w_begin, w_end) (dir 1: reverse complemented).
struct WindowGeom { int32_t w_begin, w_end; };

template <class Fetch>
CAL_HD bool extend_pam(const GuideSpec& g, const Scores& sc, Fetch fetch, int32_t m /* scanned target length */, const GuideAln& a, int pam_idx,
                       int32_t& best_score, int32_t& best_offset, uint32_t& best_xmask) {
  const int pam_len = g.pam_len[pam_idx];
  int max_extra = g.g - a.terminal_gap; int alt = g.max_tot_filter - a.diffs; if (alt < max_extra) max_extra = alt;
  bool have = false;
  for (int offset = 0; offset <= max_extra; ++offset) {
    int t_off = a.t_end + offset;                                  // 0-based offset of the first PAM base in the scanned target
    int limit = g.p; int l2 = g.max_tot_filter - a.diffs - offset; if (l2 < limit) limit = l2;
    if (t_off + pam_len > m || limit < 0) continue;
    int32_t score = 0; int nx = 0; uint32_t xmask = 0;
    for (int i = 0; i < pam_len; ++i) {
      int32_t add = pairs(g.pam[pam_idx][i], fetch(t_off + i + 1)) ? sc.pam_match : sc.pam_mismatch;
      score += add;
      if (!(add > 0)) { ++nx; xmask |= 1u << i; }                  // op '=' iff addend > 0 (:468)
    }
    if (nx > limit) continue;
    int32_t total = a.score + score + offset * sc.query_gap;
    if (!have || total > best_score) { have = true; best_score = total; best_offset = offset; best_xmask = xmask; }   // maxBy keeps the first maximum
  }
  return have;
}

// Builds the hit for guide alignment `a` extended with PAM `pam_idx` (or PAM-less when pam_idx < 0).
// The record is assembled in thread-local storage and stored once: `out` usually lives in global memory, and setting ~30 two-bit ops there
// one read-modify-write at a time is what made this the second-largest part of k_align.
CAL_HD void make_hit(const GuideSpec& g, const GuideAln& a, int pam_idx, int32_t score, int32_t offset, uint32_t xmask, int dir,
                     const WindowGeom& w, int32_t guide_idx, int32_t contig_idx, int32_t task_idx, HitX& out) {
  HitX h;
  const int pam_len = pam_idx >= 0 ? g.pam_len[pam_idx] : 0;
  const int n_ops = a.n_ops + (pam_idx >= 0 ? offset + pam_len : 0);
  h.guide_idx = guide_idx; h.pam_idx = pam_idx; h.contig_idx = contig_idx; h.task_idx = task_idx; h.score = score;
  for (int k = 0; k < CALITAS_MAX_OPS / 16; ++k) h.ops[k] = 0;
  int gaps = 0, edits = 0;
  // ops in guide orientation: DP orientation for a 3' PAM, reversed for a 5' PAM (Cigar.reverse, :267,284)
  for (int k = 0; k < n_ops; ++k) {
    uint32_t op;
    if (k < a.n_ops) op = a.ops[k];
    else if (k < a.n_ops + offset) op = OP_D;
    else op = ((xmask >> (k - a.n_ops - offset)) & 1u) ? OP_X : OP_EQ;
    if (op >= OP_I) ++gaps;
    if (op != OP_EQ) ++edits;
    ops_set(h.ops, g.five_prime ? n_ops - 1 - k : k, op);
  }
  h.n_ops = (uint8_t)n_ops; h.gap_bases = (uint8_t)gaps; h.edits = (uint8_t)edits;
  // coordinates in the scanned target (toGuideAlignment with strand '+', GuideAlignment.scala:21-31): leftDelta is always 0 because
  // a glocal alignment starts on a guide base; rightDelta = trailing D run + guide-PAM gap + PAM.
  const int32_t s0 = a.t_start - 1;
  const int32_t e0 = a.t_end + (pam_idx >= 0 ? offset + pam_len : 0);
  const int32_t gs0 = s0;
  const int32_t ge0 = e0 - (a.terminal_d + (pam_idx >= 0 ? offset + pam_len : 0));
  if (dir == 0) {
    h.start_offset = w.w_begin + s0; h.end_offset = w.w_begin + e0; h.guide_start_offset = w.w_begin + gs0; h.guide_end_offset = w.w_begin + ge0;
  } else {   // flip about the window (:271-274, 305-308)
    h.start_offset = w.w_end - e0; h.end_offset = w.w_end - s0; h.guide_start_offset = w.w_end - ge0; h.guide_end_offset = w.w_end - gs0;
  }
  h.strand = (uint8_t)((dir ^ g.five_prime) ? '-' : '+');
  out = h;
}

// ---- per-window canonicalisation (SequentialGuideAligner.scala:315-322) ---------------------------------------------------
// The alignments of one (guide, window, strand) in emission order (end column, then PAM index) are stably sorted by (score desc, gapBases asc)
// and taken greedily: one is kept iff edits <= maxTotalDiffs and it overlaps no alignment kept before it by more than maxOverlap.
// The align kernels leave a 16-byte canon key per alignment slot, so this step never touches the records.
struct CKey { int32_t score, start, end; uint32_t w; };      // w = gap bases | edits << 8 | state << 16 (0: empty slot, 1: owned window, 2: halo window)
CAL_HD CKey ckey_make(int32_t score, int32_t start, int32_t end, int gaps, int edits, int state) { return CKey{ score, start, end, (uint32_t)gaps | ((uint32_t)edits << 8) | ((uint32_t)state << 16) }; }
CAL_HD int ck_gaps(const CKey& k) { return (int)(k.w & 255u); }
CAL_HD int ck_edits(const CKey& k) { return (int)((k.w >> 8) & 255u); }
CAL_HD int ck_state(const CKey& k) { return (int)(k.w >> 16); }
CAL_HD bool ck_before(const CKey& a, int ia, const CKey& b, int ib) {      // a sorts before b in the stable (score desc, gapBases asc, arrival) order
  if (a.score != b.score) return a.score > b.score;
  if (ck_gaps(a) != ck_gaps(b)) return ck_gaps(a) < ck_gaps(b);
  return ia < ib;
}
CAL_HD int32_t ck_overlap(const CKey& a, const CKey& b) {                  // GuideAlignment.overlap (GuideAlignment.scala:119-123), clamped at 0
  const int32_t lo = a.start > b.start ? a.start : b.start, hi = a.end < b.end ? a.end : b.end;
  const int32_t ov = hi - lo; return ov < 0 ? 0 : ov;
}
// Position of slot k in the kept list of its group keys[0..n), or -1 when it is dropped; n <= 32.  Every slot's thread can call this on its own:
// the usual group is a few adjacent end columns of one site, where the best acceptable alignment is kept and everything else overlaps it, which
// takes one pass over the keys; only a slot that clears the best one replays the greedy selection (state in two bit masks).
CAL_HD int canon_slot_rank(const CKey* keys, int n, int k, int32_t max_total_diffs, int32_t max_overlap) {
  const CKey me = keys[k];
  if (ck_state(me) == 0 || ck_edits(me) > max_total_diffs) return -1;
  int b = -1; CKey best = me;
  for (int i = 0; i < n; ++i) {
    const CKey c = keys[i];
    if (ck_state(c) == 0 || ck_edits(c) > max_total_diffs) continue;
    if (b < 0 || ck_before(c, i, best, b)) { b = i; best = c; }
  }
  if (b == k) return 0;                                    // nothing acceptable sorts before it
  if (ck_overlap(me, best) > max_overlap) return -1;       // the first kept alignment already rules it out
  uint32_t decided = 0, kept = 0; int n_kept = 0;
  for (int i = 0; i < n; ++i) { const CKey c = keys[i]; if (ck_state(c) == 0 || ck_edits(c) > max_total_diffs) decided |= 1u << i; }
  for (;;) {
    int p = -1; CKey pk = me;
    for (int i = 0; i < n; ++i) { if ((decided >> i) & 1u) continue; const CKey c = keys[i]; if (p < 0 || ck_before(c, i, pk, p)) { p = i; pk = c; } }
    if (p < 0) return -1;                                  // unreachable: slot k itself is undecided until it is picked
    bool keep = true;
    for (int i = 0; i < n && keep; ++i) if (((kept >> i) & 1u) && ck_overlap(keys[i], pk) > max_overlap) keep = false;
    if (p == k) return keep ? n_kept : -1;
    decided |= 1u << p; if (keep) { kept |= 1u << p; ++n_kept; }
  }
}
// Any group size, one thread for the whole group: rank[i] = position in the kept list or -1 (rank doubles as state while running).
CAL_HD int canon_group(const CKey* keys, int32_t* rank, int n, int32_t max_total_diffs, int32_t max_overlap) {
  for (int i = 0; i < n; ++i) rank[i] = (ck_state(keys[i]) == 0 || ck_edits(keys[i]) > max_total_diffs) ? -1 : -2;      // -2 = not yet visited
  int kept = 0;
  for (;;) {
    int b = -1;
    for (int i = 0; i < n; ++i) if (rank[i] == -2 && (b < 0 || ck_before(keys[i], i, keys[b], b))) b = i;
    if (b < 0) break;
    bool keep = true;
    for (int i = 0; i < n && keep; ++i) if (rank[i] >= 0 && ck_overlap(keys[i], keys[b]) > max_overlap) keep = false;
    rank[b] = keep ? kept++ : -1;
  }
  return kept;
}

// ---- genome-wide sweep (SearchReference.scala:653-675) --------------------------------------------------------------------------
// ReferenceHit.end (ReferenceHit.scala:135-138): coordinate_start + cigar.lengthOnTarget - 1, with coordinate_start the guide-only start.
CAL_HD int32_t rec_sweep_end(const uint32_t* r) { return rec_gstart(r) + rec_span(r) - 1; }

}  // namespace cal
