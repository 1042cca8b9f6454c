// cal_host.cpp — see cal_host.h
#include "cal_host.h"
#include <cctype>
#include <cstring>

namespace cal {

static bool is_lower(char c) { return c >= 'a' && c <= 'z'; }
static bool is_upper(char c) { return c >= 'A' && c <= 'Z'; }
static bool is_alpha(char c) { return is_lower(c) || is_upper(c); }

std::string GuideDef::with_pam(int pam_idx) const {
  std::string pam = pam_idx >= 0 ? pams[(size_t)pam_idx] : std::string();
  return five_prime ? pam + protospacer : protospacer + pam;
}

// Guide.apply(sequence, auxPams): SequentialGuideAligner.scala:81-107 (case runs :110-121)
GuideDef parse_guide(const std::string& sequence, const std::vector<std::string>& aux) {
  size_t a = 0, b = sequence.size();
  while (a < b && (unsigned char)sequence[a] <= ' ') ++a;
  while (b > a && (unsigned char)sequence[b - 1] <= ' ') --b;
  const std::string seq = sequence.substr(a, b - a);
  std::vector<std::string> runs;
  for (size_t i = 0; i < seq.size();) {
    size_t j = i; const bool low = is_lower(seq[i]);
    while (j < seq.size() && is_lower(seq[j]) == low) ++j;
    runs.push_back(seq.substr(i, j - i)); i = j;
  }
  if (runs.empty() || runs.size() > 2) throw InvalidArgument("requirement failed: Invalid Guide sequence " + sequence + ".");
  if (runs.size() == 1 && !is_upper(runs[0][0])) throw InvalidArgument("requirement failed: Guide sequence cannot be all lower case.");
  if (!aux.empty() && runs.size() != 2) throw InvalidArgument("requirement failed: Cannot provide auxiliary PAMs without providing a PAM in the guide sequence.");
  for (auto& p : aux) for (char c : p) if (is_upper(c)) throw InvalidArgument("requirement failed: All PAMs must be lower case.");
  GuideDef g; g.raw = sequence;
  if (runs.size() == 1) g.protospacer = runs[0];
  else if (is_upper(runs[0][0])) { g.protospacer = runs[0]; g.pams.push_back(runs[1]); g.three_prime = true; }
  else { g.protospacer = runs[1]; g.pams.push_back(runs[0]); g.five_prime = true; }
  for (auto& p : aux) g.pams.push_back(p);
  for (auto& c : g.protospacer) c = (char)std::toupper((unsigned char)c);
  for (auto& p : g.pams) for (auto& c : p) c = (char)std::tolower((unsigned char)c);
  return g;
}
GuideDef parse_guide(const calitas_guide& g) {
  if (!g.sequence) throw InvalidArgument("guide sequence is NULL");
  std::vector<std::string> aux;
  for (int i = 0; i < g.n_aux_pams; ++i) aux.push_back(g.aux_pams[i] ? g.aux_pams[i] : "");
  return parse_guide(g.sequence, aux);
}

char complement_base(char b) {
  switch (b) {
    case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A'; case 'U': return 'A';
    case 'M': return 'K'; case 'K': return 'M'; case 'R': return 'Y'; case 'Y': return 'R'; case 'W': return 'W'; case 'S': return 'S';
    case 'V': return 'B'; case 'B': return 'V'; case 'H': return 'D'; case 'D': return 'H'; case 'N': return 'N';
    case 'a': return 't'; case 'c': return 'g'; case 'g': return 'c'; case 't': return 'a'; case 'u': return 'a';
    case 'm': return 'k'; case 'k': return 'm'; case 'r': return 'y'; case 'y': return 'r'; case 'w': return 'w'; case 's': return 's';
    case 'v': return 'b'; case 'b': return 'v'; case 'h': return 'd'; case 'd': return 'h'; case 'n': return 'n';
    default: return b;
  }
}
std::string revcomp(const std::string& s) {
  std::string r(s.size(), ' ');
  for (size_t i = 0; i < s.size(); ++i) r[s.size() - 1 - i] = complement_base(s[i]);
  return r;
}

GuideSpec make_guide_spec(const GuideDef& g, const Scores& sc, const calitas_limits& lim, bool best) {
  GuideSpec s; std::memset(&s, 0, sizeof s);
  const int lp = g.protospacer_length();
  if (lp < 1) throw InvalidArgument("empty protospacer");
  if (lp > CALITAS_MAX_PROTOSPACER) throw LimitExceeded("protospacer longer than " + std::to_string(CALITAS_MAX_PROTOSPACER) + " nt is not supported: " + g.raw);
  if ((int)g.pams.size() > CALITAS_MAX_PAMS) throw LimitExceeded("more than " + std::to_string(CALITAS_MAX_PAMS) + " PAMs");
  s.lp = lp; s.five_prime = g.five_prime ? 1 : 0;
  // The DP keeps the PAM on the right: a 5' PAM guide is aligned as its reverse complement (SequentialGuideAligner.scala:260-262).
  const std::string q = g.five_prime ? revcomp(g.protospacer) : g.protospacer;
  for (int i = 0; i < lp; ++i) {
    s.q[i] = (uint8_t)iupac_set((uint8_t)q[(size_t)i]);
    uint16_t m = 0; for (uint32_t code = 0; code < 16; ++code) if (pairs(s.q[i], code)) m |= (uint16_t)(1u << code);
    s.qmask[i] = m;
  }
  const bool no_pams = g.pams.empty() || (g.pams.size() == 1 && g.pams[0].empty());   // :446
  s.n_pams = no_pams ? 0 : (int)g.pams.size();
  for (int k = 0; k < s.n_pams; ++k) {
    const std::string pam = g.five_prime ? revcomp(g.pams[(size_t)k]) : g.pams[(size_t)k];
    if (pam.size() > CALITAS_MAX_PAM_LEN) throw LimitExceeded("PAM longer than " + std::to_string(CALITAS_MAX_PAM_LEN) + " nt");
    s.pam_len[k] = (uint8_t)pam.size();
    for (size_t i = 0; i < pam.size(); ++i) s.pam[k][i] = (uint8_t)iupac_set((uint8_t)pam[i]);
  }
  s.slots = s.n_pams > 0 ? s.n_pams : 1;
  const uint32_t low = lp >= 32 ? 0u : (0xFFFFFFFFu >> lp);
  for (uint32_t code = 0; code < 16; ++code) {
    uint32_t m = low;
    for (int i = 0; i < lp; ++i) if (pairs(s.q[i], code)) m |= 1u << (32 - lp + i);
    s.peq[0][code] = m;
  }
  for (uint32_t code = 0; code < 16; ++code) s.peq[1][code] = s.peq[0][comp_code(code)];
  if (best) { s.d = lp; s.p = g.pam_length(); s.g = lim.max_gaps_between_guide_and_pam; s.max_total_diffs = s.d + s.g + s.p; s.max_overlap = 0; }
  else {
    s.d = lim.max_guide_diffs; s.p = lim.max_pam_mismatches; s.g = lim.max_gaps_between_guide_and_pam;
    s.max_total_diffs = lim.max_total_diffs >= 0 ? lim.max_total_diffs : s.d + s.g + s.p; s.max_overlap = lim.max_overlap;
  }
  if (s.d < 0 || s.p < 0 || s.g < 0) throw InvalidArgument("limits must be non-negative");
  s.max_tot_filter = s.d + s.g + s.p;
  s.min_score = sc.match * lp + sc.worst_guide_diff * s.d;
  // An end column can reach min_score only through an alignment whose net cost is <= |worst| * d; each edit costs at least
  // min_cost, so it has at most k_edits edits, hence unit edit distance <= k_edits (lossless prefilter, exact for the default costs).
  int min_cost = sc.abs_mm; if (sc.abs_genome_gap < min_cost) min_cost = sc.abs_genome_gap; if (sc.abs_guide_gap < min_cost) min_cost = sc.abs_guide_gap;
  if (min_cost <= 0 || sc.abs_guide_gap <= 0) throw LimitExceeded("net costs must be non-zero");
  const long long budget = (long long)(-sc.worst_guide_diff) * s.d;      // net cost an accepted alignment can spend
  const long long k = budget / min_cost;
  s.k_edits = k > lp ? lp : (int)k;
  // Deletions (genome bases opposite a guide gap) any co-optimal alignment of an accepted end column can hold: each costs |guideGap|, so never
  // more than budget / |guideGap| (NOT the capped k_edits: with a cheap guide gap an accepted alignment can hold more than lp of them), and never
  // so many that inserting the whole guide (score lp * target_gap) would beat it.  Insertions: at most lp rows, each costing |genomeGap|.
  const long long del_cost = budget / sc.abs_guide_gap, del_max = (long long)lp * sc.abs_genome_gap / sc.abs_guide_gap;
  const long long dels = del_cost < del_max ? del_cost : del_max;
  long long ins = budget / sc.abs_genome_gap; if (ins > lp) ins = lp;
  const long long span = lp + dels;
  if (span > MAX_SPAN) throw LimitExceeded("costs/limits need a DP band wider than this build supports");
  s.span = (int)span;
  s.band_k = (int)(dels > ins ? dels : ins);   // a co-optimal path of an accepted end cell never leaves the 2 * band_k + 1 diagonals around the end cell's
  if (lp + dels + s.g + g.pam_length() > CALITAS_MAX_OPS) throw LimitExceeded("alignment longer than CALITAS_MAX_OPS columns");
  if (dels + s.g + g.pam_length() > 63) throw LimitExceeded("trailing deletions + guide-PAM gap + PAM longer than 63 bases do not fit the hit record");
  s.max_cols = (int)(lp + dels + (s.n_pams > 0 ? s.g + g.pam_length() : 0));
  return s;
}

// Allocation-free core of the rendering: every ReferenceHit row goes through it (tens of millions per run), so it works on fixed arrays.
// `guide` = guide + PAM text in guide orientation (GuideDef::with_pam); `target_fwd` = the hit's bases as stored, forward orientation.
void render_hit_fix(const HitX& h, const char* guide, int guide_len, const char* target_fwd, int target_len, bool upper_case, RenderedFix& r) {
  static const char kOps[4] = { '=', 'X', 'I', 'D' };
  const bool neg = h.strand == '-';
  const int n = h.n_ops;
  if (n > CALITAS_MAX_OPS) throw std::runtime_error("hit has more alignment columns than CALITAS_MAX_OPS");
  r.n = n; r.cigar_len = 0; r.unpadded_len = 0;
  r.mismatches = r.gap_bases = r.guide_mm = r.guide_gaps = r.pam_mm = r.pam_gaps = 0;
  int qi = 0, ti = 0, run = 0; char run_op = 0;
  char* pg = r.padded_guide; char* pa = r.padded_alignment; char* pt = r.padded_target;
  auto flush_run = [&] { if (run) { char tmp[12]; int k = 0, v = run; do { tmp[k++] = (char)('0' + v % 10); v /= 10; } while (v); while (k) r.cigar[r.cigar_len++] = tmp[--k]; r.cigar[r.cigar_len++] = run_op; } };
  for (int k = 0; k < n; ++k) {
    const uint32_t op = ops_get(h.ops, k);
    const bool has_q = op != OP_D, has_t = op != OP_I;
    if ((has_q && qi >= guide_len) || (has_t && ti >= target_len)) throw std::runtime_error("hit ops do not fit the guide/target");
    pg[k] = has_q ? guide[qi++] : '-';
    if (has_t) {
      char c = neg ? complement_base(target_fwd[target_len - 1 - ti]) : target_fwd[ti];
      ++ti;
      if (upper_case) c = (char)std::toupper((unsigned char)c);
      pt[k] = c;
    } else pt[k] = '-';
    pa[k] = op == OP_EQ ? '|' : (op == OP_X ? '.' : '~');
    if (kOps[op] != run_op) { flush_run(); run_op = kOps[op]; run = 0; }
    ++run;
  }
  flush_run();
  if (qi != guide_len || ti != target_len) throw std::runtime_error("hit ops do not cover the guide/target");
  // counters of GuideAlignment.scala:99-108,139-163, computed per column
  int first_upper = -1, last_upper = -1;
  for (int i = 0; i < n; ++i) if (is_upper(pg[i])) { if (first_upper < 0) first_upper = i; last_upper = i; }
  for (int i = 0; i < n; ++i) {
    const char a = pa[i], q = pg[i];
    if (a == '.') { ++r.mismatches; if (is_lower(q)) ++r.pam_mm; else ++r.guide_mm; }
    else if (a == '~') {
      ++r.gap_bases;
      bool in_guide, in_pam;
      if (q != '-') { in_guide = !is_lower(q); in_pam = is_lower(q); }
      else {
        int lo = i; while (lo > 0 && pg[lo] == '-') --lo;
        int hi = i; while (hi < n - 1 && pg[hi] == '-') ++hi;
        const char prev = pg[lo], next = pg[hi];
        in_guide = (is_alpha(prev) && !is_lower(prev)) || (is_alpha(next) && !is_lower(next));
        in_pam = (prev == '-' || is_lower(prev)) && (next == '-' || is_lower(next));
      }
      if (in_guide) ++r.guide_gaps;
      if (in_pam) ++r.pam_gaps;
    }
  }
  r.edits = r.mismatches + r.gap_bases;
  r.guide_mm_plus_gaps = r.guide_mm + r.guide_gaps; r.pam_mm_plus_gaps = r.pam_mm + r.pam_gaps;
  if (first_upper >= 0) for (int i = first_upper; i <= last_upper; ++i) if (is_alpha(pt[i])) r.unpadded[r.unpadded_len++] = pt[i];
}

Rendered render_hit(const HitX& h, const GuideDef& g, const std::string& target_fwd, bool upper_case) {
  Rendered r; RenderedFix f;
  r.guide = g.with_pam(h.pam_idx);
  render_hit_fix(h, r.guide.data(), (int)r.guide.size(), target_fwd.data(), (int)target_fwd.size(), upper_case, f);
  r.padded_guide.assign(f.padded_guide, (size_t)f.n); r.padded_alignment.assign(f.padded_alignment, (size_t)f.n); r.padded_target.assign(f.padded_target, (size_t)f.n);
  r.cigar.assign(f.cigar, (size_t)f.cigar_len); r.unpadded_target_without_pam.assign(f.unpadded, (size_t)f.unpadded_len);
  r.mismatches = f.mismatches; r.gap_bases = f.gap_bases; r.edits = f.edits; r.guide_mm = f.guide_mm; r.guide_gaps = f.guide_gaps; r.guide_mm_plus_gaps = f.guide_mm_plus_gaps;
  r.pam_mm = f.pam_mm; r.pam_gaps = f.pam_gaps; r.pam_mm_plus_gaps = f.pam_mm_plus_gaps;
  return r;
}

std::string alignment_header() {
  return "guide\tchrom\tstartOffset\tendOffset\tguideStartOffset\tguideEndOffset\tstrand\tscore\tcigar\tpaddedGuide\tpaddedAlignment\tpaddedTarget\t"
         "mismatches\tgapBases\tedits\tguideMismatches\tguideGapBases\tguideMmsPlusGaps\tpamMismatches\tpamGapBases\tpamMmsPlusGaps\tunpaddedTargetWithoutPam\n";
}
std::string alignment_row(const HitX& h, const Rendered& r, const std::string& chrom) {
  std::string s = r.guide; auto add = [&](const std::string& v) { s += '\t'; s += v; }; auto I = [](int v) { return std::to_string(v); };
  add(chrom); add(I(h.start_offset)); add(I(h.end_offset)); add(I(h.guide_start_offset)); add(I(h.guide_end_offset)); add(std::string(1, (char)h.strand));
  add(I(h.score)); add(r.cigar); add(r.padded_guide); add(r.padded_alignment); add(r.padded_target); add(I(r.mismatches)); add(I(r.gap_bases)); add(I(r.edits));
  add(I(r.guide_mm)); add(I(r.guide_gaps)); add(I(r.guide_mm_plus_gaps)); add(I(r.pam_mm)); add(I(r.pam_gaps)); add(I(r.pam_mm_plus_gaps)); add(r.unpadded_target_without_pam);
  s += '\n'; return s;
}

}  // namespace cal
