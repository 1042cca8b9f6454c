// cal_tools.cpp — host-side mirror of the reference's operators (see include/calitas_b200_tools.h).
// Alignments are computed only by the device engine (calitas_search / calitas_align_regions / calitas_align_targets).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <unistd.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "../../include/calitas_b200_tools.h"
#include "cal_host.h"

using namespace cal;

namespace {

typedef std::string Str;
struct ToolError { int code; Str msg; };
[[noreturn]] void bad(const Str& m) { throw ToolError{ CALITAS_EINVAL, m }; }
void ck(int rc) { if (rc != CALITAS_OK) throw ToolError{ rc, calitas_last_error() }; }

extern "C" int calitas_tools_set_error(int code, const char* msg);   // defined in cal_engine.cu

template <class F> int guarded(F f) {
  try { return f(); }
  catch (const ToolError& e) { return calitas_tools_set_error(e.code, e.msg.c_str()); }
  catch (const InvalidArgument& e) { return calitas_tools_set_error(CALITAS_EINVAL, e.what()); }
  catch (const LimitExceeded& e) { return calitas_tools_set_error(CALITAS_ELIMIT, e.what()); }
  catch (const std::exception& e) { return calitas_tools_set_error(CALITAS_ESTATE, e.what()); }
}
char* dup_text(const Str& s) { char* p = (char*)std::malloc(s.size() + 1); std::memcpy(p, s.data(), s.size() + 1); return p; }
Str to_upper(Str s) { for (auto& c : s) c = (char)std::toupper((unsigned char)c); return s; }

struct HitSet {   // RAII over calitas_hitset
  calitas_hitset* h = nullptr; ~HitSet() { calitas_hitset_free(h); }
  int64_t n() const { return calitas_hitset_count(h); }
  // record i of the set (packed, calitas_hitset_stride() bytes apart) and its expanded form
  const calitas_hit* at(int64_t i) const { return reinterpret_cast<const calitas_hit*>(reinterpret_cast<const char*>(calitas_hitset_data(h)) + (size_t)i * (size_t)calitas_hitset_stride(h)); }
  HitX x(int64_t i) const { HitX v; unpack_hit(reinterpret_cast<const uint32_t*>(at(i)), calitas_hitset_stride(h) / 4, v); return v; }
  std::vector<HitX> all() const { std::vector<HitX> v((size_t)n()); for (int64_t i = 0; i < n(); ++i) v[(size_t)i] = x(i); return v; }
};

// Row rendering is the host-side bottleneck once the search takes milliseconds: rows are independent, so they are
// rendered on all host threads.  fn(begin, end) must only touch its own index range.
// CALITAS_TOOL_TIMING=1: phase timings of the tool layer on stderr
struct PhaseTimer {
  bool on = std::getenv("CALITAS_TOOL_TIMING") != nullptr; std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
  void lap(const char* what) { if (!on) return; auto n = std::chrono::steady_clock::now(); std::fprintf(stderr, "[calitas tool] %-28s %8.3f s\n", what, std::chrono::duration<double>(n - t).count()); t = n; }
};

template <class F> void parallel_for(int64_t n, int64_t grain, F fn) {
  const int64_t want = (n + grain - 1) / grain;
  const int nt = (int)std::max<int64_t>(1, std::min<int64_t>(want, std::min<unsigned>(32u, std::max(1u, std::thread::hardware_concurrency()))));
  if (nt <= 1) { fn((int64_t)0, n); return; }
  std::vector<std::thread> th; std::vector<Str> err((size_t)nt); std::vector<int> code((size_t)nt, 0);
  for (int t = 0; t < nt; ++t) th.emplace_back([&, t]() {
    try { fn(n * t / nt, n * (t + 1) / nt); }
    catch (const ToolError& e) { code[(size_t)t] = e.code; err[(size_t)t] = e.msg; }
    catch (const std::exception& e) { code[(size_t)t] = CALITAS_ESTATE; err[(size_t)t] = e.what(); } });
  for (auto& t : th) t.join();
  for (int t = 0; t < nt; ++t) if (code[(size_t)t]) throw ToolError{ code[(size_t)t], err[(size_t)t] };
}

int contig_index(const calitas_genome_view& g, const Str& name) { for (int i = 0; i < g.n_contigs; ++i) if (name == g.names[i]) return i; return -1; }

// ---- GuideAlignment ordering (GuideAlignment.scala:125-129): stable sort by score desc, gap bases asc ----
void sort_alignments(std::vector<HitX>& v) {
  std::stable_sort(v.begin(), v.end(), [](const HitX& a, const HitX& b) { return a.score != b.score ? a.score > b.score : a.gap_bases < b.gap_bases; });
}

// ---- ReferenceHit rows (ReferenceHit.scala:99-132, 210-254) ----------------------------------------------------------------------
struct VariantAllele { Str id; int pos; Str ref, alt; float af; const void* rec = nullptr; int alt_idx = 0; /* the VCF record and which of its ALTs */ };
Str variant_display(const VariantAllele& v) {   // SearchReference.scala:106-109
  char buf[64]; std::snprintf(buf, sizeof buf, "%.3f", (double)v.af);
  return (v.id.empty() ? Str(".") : v.id) + ":" + std::to_string(v.pos - 1) + ":" + v.ref + ">" + v.alt + ":" + buf;
}
Str format_af(double d) {   // fgbio Metric: DecimalFormat("0.0####")
  char buf[64]; std::snprintf(buf, sizeof buf, "%.5f", d); Str s(buf);
  while (s.size() > 3 && s.back() == '0' && s[s.size() - 2] != '.') s.pop_back();
  return s;
}

struct Row {   // what removeOverlaps / sort need, plus the rendered line
  int contig; int start; char strand; int score; int sweep_end; Str key /* chromosome:strand:variant_description */; Str line;
};

struct RowContext {
  const calitas_genome_view* genome; Str guide_id, aligner_id, arguments, time_stamp, version, vcf_id; bool has_vcf = false;
};

Str fetch_bases(const calitas_genome_view& g, int contig, int start1, int end1, bool rc) {   // ReferenceHit.scala:261-266 (1-based inclusive)
  const int64_t len = g.lengths[contig];
  const int as = std::max(1, start1), ae = (int)std::min<int64_t>(len, end1);
  Str s((size_t)(as - start1), 'N');
  if (ae >= as) s.append((const char*)g.bases[contig] + as - 1, (size_t)(ae - as + 1));
  s.append((size_t)std::max(0, end1 - std::max(ae, as - 1)), 'N');
  return to_upper(rc ? revcomp(s) : s);
}

struct Flanks { bool has[4] = { false, false, false, false }; Str v[4]; };   // left10, right10, left8, right8 in guide orientation

// ---- the row writer: one pass, no per-field strings (a 100-guide run renders 3 x 10^7 rows) -----------------------------------------
inline void put_int(Str& out, int v) {
  char tmp[12]; int k = 0; unsigned u = v < 0 ? 0u - (unsigned)v : (unsigned)v;
  do { tmp[k++] = (char)('0' + u % 10); u /= 10; } while (u);
  if (v < 0) tmp[k++] = '-';
  char rev[12]; for (int i = 0; i < k; ++i) rev[i] = tmp[k - 1 - i];
  out.append(rev, (size_t)k);
}
// fetch_bases straight into the row: bases [start1, end1] (1-based inclusive), N beyond the contig ends, reverse-complemented for '-', upper case
inline void put_flank(Str& out, const calitas_genome_view& g, int contig, int start1, int end1, bool rc) {
  const int64_t len = g.lengths[contig]; const char* b = (const char*)g.bases[contig];
  char buf[32]; const int n = end1 - start1 + 1;
  if (n <= 0 || n > 32) { out += fetch_bases(g, contig, start1, end1, rc); return; }
  for (int i = 0; i < n; ++i) { const int64_t p = (int64_t)start1 + i; buf[i] = (p >= 1 && p <= len) ? b[p - 1] : 'N'; }
  if (rc) { for (int i = 0; i < n / 2; ++i) { const char t = buf[i]; buf[i] = buf[n - 1 - i]; buf[n - 1 - i] = t; } for (int i = 0; i < n; ++i) buf[i] = complement_base(buf[i]); }
  for (int i = 0; i < n; ++i) buf[i] = (char)std::toupper((unsigned char)buf[i]);
  out.append(buf, (size_t)n);
}
// Per-guide constants of the 34 columns (ReferenceHit.scala:99-132): columns 1-2 and 28, 30-34.
struct RowConst {
  Str head;        // guide_id \t unpadded_guide_sequence \t
  Str proto_len;   // unpadded_guide_sequence_length
  Str tail;        // \t aligner \t aligner_version \t aligner_search_pam \t aligner_other_parameters \t time_stamp \n
  Str build_ref, build_var;
  std::vector<Str> guide_text;   // guide + PAM per pam index; [n_pams] = PAM-less
};
RowConst make_row_const(const RowContext& cx, const GuideDef& gd) {
  RowConst rc; rc.head = cx.guide_id + "\t" + gd.protospacer + "\t"; rc.proto_len = std::to_string(gd.protospacer.size());
  Str pams; for (size_t i = 0; i < gd.pams.size(); ++i) { if (i) pams += ','; pams += gd.pams[i]; }                       // ReferenceHit.scala:207
  rc.tail = "\t" + cx.aligner_id + "\t" + cx.version + "\t" + pams + "\t" + cx.arguments + "\t" + cx.time_stamp + "\n";
  const calitas_genome_view& g = *cx.genome;
  rc.build_ref = (g.assembly && g.assembly[0]) ? g.assembly : "unknown"; rc.build_var = rc.build_ref + "+variants";
  for (size_t i = 0; i < gd.pams.size(); ++i) rc.guide_text.push_back(gd.with_pam((int)i));
  rc.guide_text.push_back(gd.with_pam(-1));
  return rc;
}
const Str& guide_text_of(const RowConst& rc, int pam_idx) { return pam_idx >= 0 ? rc.guide_text[(size_t)pam_idx] : rc.guide_text.back(); }

// Appends one ReferenceHit line (ReferenceHit.scala:210-254).  `var` = variant columns (id, description, vcf, allele frequency), or NULL.
void write_row(Str& out, const RowContext& cx, const RowConst& rc, const GuideDef& gd, const HitX& h, const RenderedFix& r, int contig, int so, int eo, int gso, int geo,
               const Str* var /* [4] */, const Flanks& fl) {
  const calitas_genome_view& g = *cx.genome;
  const bool neg = h.strand == '-';
  out += rc.head; out += var ? rc.build_var : rc.build_ref; out += '\t'; out += g.names[contig]; out += '\t'; put_int(out, gso); out += '\t'; put_int(out, geo); out += '\t';
  out += (char)h.strand; out += '\t'; out.append(r.unpadded, (size_t)r.unpadded_len); out += '\t';
  // ten bases 5' / 3' of the protospacer span, eight of the full span, in guide orientation (ReferenceHit.scala:226-249)
  if (fl.has[0]) out += fl.v[0]; else if (neg) put_flank(out, g, contig, geo + 1, geo + 10, true); else put_flank(out, g, contig, gso + 1 - 10, gso, false);
  out += '\t';
  if (fl.has[1]) out += fl.v[1]; else if (neg) put_flank(out, g, contig, gso + 1 - 10, gso, true); else put_flank(out, g, contig, geo + 1, geo + 10, false);
  out += '\t';
  if (h.pam_idx >= 0) out += gd.pams[(size_t)h.pam_idx];
  out += '\t';
  if (var) { out += var[0]; out += '\t'; out += var[1]; out += '\t'; out += var[2]; out += '\t'; out += var[3]; out += '\t'; } else out.append("\t\t\t\t", 4);
  put_int(out, h.score); out += '\t'; put_int(out, r.guide_mm); out += '\t'; put_int(out, r.guide_gaps); out += '\t'; put_int(out, r.guide_mm_plus_gaps); out += '\t';
  put_int(out, r.pam_mm); out += '\t'; put_int(out, r.edits); out += '\t';
  out.append(r.padded_guide, (size_t)r.n); out += '\t'; out.append(r.padded_alignment, (size_t)r.n); out += '\t'; out.append(r.padded_target, (size_t)r.n); out += '\t';
  if (fl.has[2]) out += fl.v[2]; else if (neg) put_flank(out, g, contig, eo + 1, eo + 8, true); else put_flank(out, g, contig, so + 1 - 8, so, false);
  out += '\t';
  if (fl.has[3]) out += fl.v[3]; else if (neg) put_flank(out, g, contig, so + 1 - 8, so, true); else put_flank(out, g, contig, eo + 1, eo + 8, false);
  out += '\t';
  out.append(r.cigar, (size_t)r.cigar_len); out += '\t'; out += rc.proto_len; out += '\t'; put_int(out, r.unpadded_len); out += rc.tail;
}

// ReferenceHit.sort order (ReferenceHit.scala:276-287) on hit records: contig, coordinate_start, strand ("+" < "-"), score descending.
// A hit of one engine's result set: the record, its annotation when the set came from calitas_search_variants (else NULL), the engine.
struct HitRef {
  const calitas_hit* h; const calitas_variant_hit_info* v; int engine;
  int gstart() const { return v ? v->guide_start_offset : calitas_hit_guide_start_offset(h); }       // reference coordinate (variant-window hits keep window offsets in the record)
};
bool hit_sorts_before(const HitRef& a, const HitRef& b) {
  const int ca = calitas_hit_contig_idx(a.h), cb = calitas_hit_contig_idx(b.h); if (ca != cb) return ca < cb;
  const int sa = a.gstart(), sb = b.gstart(); if (sa != sb) return sa < sb;
  const char ta = calitas_hit_strand(a.h), tb = calitas_hit_strand(b.h); if (ta != tb) return ta < tb;
  return a.h->score > b.h->score;
}
// order[0, seg) and order[seg, end) are each sorted (one guide's hits of the shards so far, and of the next shard).  Shards are ascending base
// ranges, but consecutive windows overlap by guide length + d + g - 1 bases, so the last window of one shard and the first of the next can
// report hits whose starts interleave (whenever removeOverlaps does not collapse them, e.g. a large -O).  Only that stretch is merged: stable,
// the earlier shard first on equal keys, which is the single engine's arrival order.
void merge_at_cut(std::vector<HitRef>& order, size_t seg) {
  if (seg == 0 || seg >= order.size() || !hit_sorts_before(order[seg], order[seg - 1])) return;
  auto first = std::upper_bound(order.begin(), order.begin() + (long)seg, order[seg], hit_sorts_before);                 // prefix elements <= the segment's first stay put
  auto last = std::lower_bound(order.begin() + (long)seg, order.end(), order[seg - 1], hit_sorts_before);                // segment elements >= the prefix's last stay put
  std::inplace_merge(first, order.begin() + (long)seg, last, hit_sorts_before);
}

// A row with what removeOverlaps / sort on the host need (VCF runs only: variant-window hits are merged with reference hits there).
Row make_row(const RowContext& cx, const RowConst& rc, const HitX& h, const RenderedFix& r, const GuideDef& gd, int contig, int so, int eo, int gso, int geo,
             const std::vector<VariantAllele>& variants, const Flanks& fl) {
  const calitas_genome_view& g = *cx.genome;
  std::vector<const VariantAllele*> vs;
  for (auto& v : variants) if (v.pos - 1 >= so && v.pos - 1 <= eo) vs.push_back(&v);       // ReferenceHit.scala:211
  Str var[4]; float min_af = 0;
  for (size_t i = 0; i < vs.size(); ++i) { if (i) { var[0] += ';'; var[1] += ';'; } var[0] += vs[i]->id; var[1] += variant_display(*vs[i]); if (i == 0 || vs[i]->af < min_af) min_af = vs[i]->af; }
  if (!vs.empty()) { if (cx.has_vcf) var[2] = cx.vcf_id; var[3] = format_af((double)min_af); }
  Row row; row.contig = contig; row.start = gso; row.strand = (char)h.strand; row.score = h.score;
  row.sweep_end = gso + (h.end_offset - h.start_offset) - 1;           // ReferenceHit.scala:135-138 (cigar.lengthOnTarget is window-relative span)
  row.key = Str("{") + g.names[contig] + ":" + (char)h.strand + ":" + var[1];
  write_row(row.line, cx, rc, gd, h, r, contig, so, eo, gso, geo, vs.empty() ? nullptr : var, fl);
  return row;
}

// Writes all of `data` to a file descriptor (sequential: works for files, pipes and terminals alike).
void write_all(int fd, const char* data, size_t n) {
  size_t off = 0;
  while (off < n) { const ssize_t put = ::write(fd, data + off, n - off); if (put <= 0) throw ToolError{ CALITAS_ESTATE, "short write to the output" }; off += (size_t)put; }
}
// Renders n_blocks text blocks on all host threads and hands them to `consume` strictly in order on the calling thread, while later blocks are
// still being rendered: the table goes to disk as it is produced instead of being assembled in memory first (18 GB for 100 guides on hg38).
// Blocks are claimed in index order and a worker never runs more than `lookahead` blocks ahead of the consumer, which bounds the memory held.
template <class Render, class Consume> void ordered_pipeline(int64_t n_blocks, int64_t lookahead, Render render, Consume consume) {
  if (n_blocks <= 0) return;
  const int nt = (int)std::max<int64_t>(1, std::min<int64_t>(n_blocks, std::min<unsigned>(32u, std::max(1u, std::thread::hardware_concurrency()))));
  std::vector<Str> blocks((size_t)n_blocks); std::vector<char> ready((size_t)n_blocks, 0);
  std::mutex mu; std::condition_variable cv; int64_t next = 0, consumed = 0; bool failed = false; Str err; int err_code = CALITAS_ESTATE;
  auto worker = [&]() {
    for (;;) {
      int64_t k;
      { std::unique_lock<std::mutex> lk(mu);
        k = next++;
        if (k >= n_blocks) return;
        cv.wait(lk, [&] { return failed || k < consumed + lookahead; });
        if (failed) return; }
      try { render(k, blocks[(size_t)k]); }
      catch (const ToolError& e) { std::lock_guard<std::mutex> lk(mu); failed = true; err = e.msg; err_code = e.code; }
      catch (const std::exception& e) { std::lock_guard<std::mutex> lk(mu); failed = true; err = e.what(); }
      { std::lock_guard<std::mutex> lk(mu); ready[(size_t)k] = 1; }
      cv.notify_all();
    } };
  std::vector<std::thread> th; for (int t = 0; t < nt; ++t) th.emplace_back(worker);
  try {
    for (int64_t k = 0; k < n_blocks; ++k) {
      { std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&] { return failed || ready[(size_t)k]; }); if (failed) break; }
      consume(blocks[(size_t)k]); Str().swap(blocks[(size_t)k]);
      { std::lock_guard<std::mutex> lk(mu); consumed = k + 1; }
      cv.notify_all();
    }
  } catch (const ToolError& e) { std::lock_guard<std::mutex> lk(mu); failed = true; err = e.msg; err_code = e.code; }
  catch (const std::exception& e) { std::lock_guard<std::mutex> lk(mu); failed = true; err = e.what(); }       // the workers must be joined before anything propagates
  { std::lock_guard<std::mutex> lk(mu); if (failed) next = n_blocks; }
  cv.notify_all();
  for (auto& t : th) t.join();
  if (failed) throw ToolError{ err_code, err };
}

Str hit_header() {
  return "guide_id\tunpadded_guide_sequence\tgenome_build\tchromosome\tcoordinate_start\tcoordinate_end\tstrand\tunpadded_target_sequence\tten_bases_5_prime\t"
         "ten_bases_3_prime\tpam_used\tvariant_id\tvariant_description\tvariant_vcf\tallele_frequency\tscore\tguide_mm\tguide_gaps\tguide_mm_plus_gaps\tpam_mm\t"
         "total_mm_plus_gaps\tpadded_guide\tpadded_alignment\tpadded_target\tpadded_extra_8_bases_5_prime\tpadded_extra_8_bases_3_prime\tcigar\t"
         "unpadded_guide_sequence_length\tunpadded_target_sequence_length\taligner\taligner_version\taligner_search_pam\taligner_other_parameters\ttime_stamp\n";
}

void sort_rows(std::vector<Row>& rows) {   // ReferenceHit.sort, ReferenceHit.scala:284
  std::stable_sort(rows.begin(), rows.end(), [](const Row& a, const Row& b) {
    if (a.contig != b.contig) return a.contig < b.contig;
    if (a.start != b.start) return a.start < b.start;
    if (a.strand != b.strand) return a.strand < b.strand;
    return a.score > b.score;
  });
}
int row_overlap(const Row& a, const Row& b) { return std::max(0, std::min(a.sweep_end, b.sweep_end) - std::max(a.start, b.start)); }

// removeOverlaps on the host — used only when variant-window hits must be merged with reference hits (SearchReference.scala:653-675);
// the plain reference search does this on the device.
std::vector<Row> remove_overlaps_host(std::vector<Row>& hits, int max_overlap) {
  std::map<Str, std::vector<Row>> groups;
  for (auto& h : hits) groups[h.key].push_back(std::move(h));
  std::vector<Row> keep;
  for (auto& kv : groups) {
    std::vector<Row>& v = kv.second; sort_rows(v);
    for (size_t i = 0; i < v.size();) {
      const size_t cur = i++;
      while (i < v.size() && row_overlap(v[i], v[cur]) >= max_overlap && v[i].score <= v[cur].score) ++i;
      if (i >= v.size() || row_overlap(v[i], v[cur]) < max_overlap) keep.push_back(v[cur]);
    }
  }
  return keep;
}

Str contig_slice(const calitas_genome_view& g, int contig, int start, int end) { return Str((const char*)g.bases[contig] + start, (size_t)(end - start)); }

// ---- variant windows (SearchReference.scala:101-400) -----------------------------------------------------------------------------------
struct VcfRecord { Str chrom; int pos; Str id, ref; std::vector<Str> alts; std::vector<float> afs; bool has_af = false; int end() const { return pos + (int)ref.size() - 1; } };


// One VCF data line [b, e) (no newline) -> record; false for blank lines and header lines.  Fields are cut with pointer arithmetic: a 3-million-record file
// is parsed in a fraction of a second per host thread.
bool parse_vcf_line(const char* b, const char* e, VcfRecord& v) {
  if (e > b && e[-1] == '\r') --e;
  if (e == b || *b == '#') return false;
  const char* f[9]; int nf = 0; f[nf++] = b;
  for (const char* p = b; p < e && nf < 9; ++p) if (*p == '\t') f[nf++] = p + 1;
  if (nf < 5) bad("malformed VCF record: " + Str(b, e));
  auto end_of = [&](int k) { return k + 1 < nf ? f[k + 1] - 1 : e; };
  v.chrom.assign(f[0], end_of(0)); v.pos = std::atoi(Str(f[1], end_of(1)).c_str());
  v.id.assign(f[2], end_of(2)); if (v.id == ".") v.id.clear();
  v.ref.assign(f[3], end_of(3));
  for (const char* p = f[4], *q = end_of(4);;) { const char* c = std::find(p, q, ','); if (!(c - p == 1 && *p == '.')) v.alts.emplace_back(p, c); if (c == q) break; p = c + 1; }
  if (nf >= 8) {
    for (const char* p = f[7], *q = end_of(7);;) {
      const char* c = std::find(p, q, ';');
      if (c - p >= 3 && p[0] == 'A' && p[1] == 'F' && p[2] == '=') {
        v.has_af = true;
        for (const char* x = p + 3;;) { const char* y = std::find(x, c, ','); v.afs.push_back((y - x == 1 && *x == '.') ? 0.0f : std::strtof(Str(x, y).c_str(), nullptr)); if (y == c) break; x = y + 1; }
      }
      if (c == q) break;
      p = c + 1;
    }
  }
  return true;
}

std::vector<VcfRecord> parse_vcf(const char* text) {
  const size_t n = std::strlen(text);
  const int nt = (int)std::max<size_t>(1, std::min<size_t>(n / (4u << 20) + 1, std::min<unsigned>(32u, std::max(1u, std::thread::hardware_concurrency()))));
  std::vector<size_t> cut((size_t)nt + 1, n); cut[0] = 0;
  for (int t = 1; t < nt; ++t) { size_t p = n * (size_t)t / (size_t)nt; while (p < n && text[p] != '\n') ++p; cut[(size_t)t] = p < n ? p + 1 : n; }
  std::vector<std::vector<VcfRecord>> part((size_t)nt);
  parallel_for(nt, 1, [&](int64_t tb, int64_t te) {
    for (int64_t t = tb; t < te; ++t) {
      std::vector<VcfRecord>& out = part[(size_t)t];
      for (size_t p = cut[(size_t)t]; p < cut[(size_t)t + 1];) {
        const char* b = text + p; const char* e = (const char*)std::memchr(b, '\n', cut[(size_t)t + 1] - p); if (!e) e = text + cut[(size_t)t + 1];
        VcfRecord v; if (parse_vcf_line(b, e, v)) out.push_back(std::move(v));
        p = (size_t)(e - text) + 1;
      }
    } });
  if (nt == 1) return std::move(part[0]);
  size_t total = 0; for (auto& v : part) total += v.size();
  std::vector<VcfRecord> out; out.reserve(total);
  for (auto& v : part) for (auto& r : v) out.push_back(std::move(r));
  return out;
}

struct VariantWindow {
  int contig; int start /* 1-based */; std::vector<VariantAllele> alleles; Str cigar_ops /* M I D per unit */; Str bases;
  Str cigar_text() const { Str o; for (size_t i = 0; i < cigar_ops.size();) { size_t j = i; while (j < cigar_ops.size() && cigar_ops[j] == cigar_ops[i]) ++j; o += std::to_string(j - i); o += cigar_ops[i]; i = j; } return o; }
  // SearchReference.scala:133-156
  int ref_offset_at(int offset, bool preceding) const {
    if (offset == (int)bases.size()) { int lot = 0; for (char c : cigar_ops) lot += (c != 'I'); return start - 1 + lot; }
    int ref_off = start - 1, base_off = 0;
    for (size_t i = 0; i < cigar_ops.size();) {
      size_t j = i; while (j < cigar_ops.size() && cigar_ops[j] == cigar_ops[i]) ++j;
      const char op = cigar_ops[i]; const int len = (int)(j - i), on_query = op == 'D' ? 0 : len, on_ref = op == 'I' ? 0 : len;
      if (offset < base_off + on_query) {
        if (op == 'I') return preceding ? ref_off - 1 : ref_off;
        return ref_off + (offset - base_off);
      }
      ref_off += on_ref; base_off += on_query; i = j;
    }
    bad("window offset outside the variant window");
  }
};

// all allele index vectors in the order of SearchReference.scala:377-399: first variant varies slowest
void allele_vectors(const std::vector<int>& counts, std::vector<std::vector<int>>& out) {
  size_t total = 1; for (int c : counts) total *= (size_t)c;
  out.assign(total, std::vector<int>(counts.size(), 0));
  for (size_t r = 0; r < total; ++r) { size_t rem = r; for (size_t i = counts.size(); i-- > 0;) { out[r][i] = (int)(rem % (size_t)counts[i]); rem /= (size_t)counts[i]; } }
}

// Contig bases as the variant iterator sees them: upper-cased (SearchReference.scala:225); only the window is copied.
struct ContigRef { const char* p; int64_t len; };
VariantWindow build_variant_window(const std::vector<const VcfRecord*>& vars, const std::vector<int>& alleles, int contig, const ContigRef& ref, int padding) {
  const int ws = std::max(1, vars.front()->pos - padding), we = (int)std::min<int64_t>(ref.len, (int64_t)vars.back()->end() + padding);
  VariantWindow w; w.contig = contig; w.start = ws;
  if (ws <= we) { w.bases.assign(ref.p + ws - 1, (size_t)(we - ws + 1)); for (auto& c : w.bases) c = (char)std::toupper((unsigned char)c); }
  for (size_t i = 0; i < vars.size(); ++i) {
    const VcfRecord* v = vars[i]; const int a = alleles[i];
    const float af = (v->has_af && a - 1 < (int)v->afs.size()) ? v->afs[(size_t)a - 1] : 0.0f;     // SearchReference.scala:199
    w.alleles.push_back(VariantAllele{ v->id, v->pos, v->ref, v->alts[(size_t)a - 1], af, v, a - 1 });
  }
  for (size_t k = w.alleles.size(); k-- > 0;) {   // right to left (:270-279)
    const VariantAllele& a = w.alleles[k]; const size_t at = (size_t)(a.pos - ws);
    if (at > w.bases.size()) bad("variant lies outside its contig: " + a.id);
    if (a.ref.size() == a.alt.size()) w.bases.replace(at, a.alt.size(), a.alt);
    else w.bases = w.bases.substr(0, at) + a.alt + (at + a.ref.size() < w.bases.size() ? w.bases.substr(at + a.ref.size()) : Str());
  }
  int ref_pos = ws, base_off = 0;
  for (auto& a : w.alleles) {   // :282-319
    const int pre = a.pos - ref_pos;
    if (pre > 0) { w.cigar_ops.append((size_t)pre, 'M'); ref_pos += pre; base_off += pre; }
    if (a.ref.size() == a.alt.size()) w.cigar_ops.append(a.ref.size(), 'M');
    else if (a.ref.size() == 1) { w.cigar_ops += 'M'; w.cigar_ops.append(a.alt.size() - 1, 'I'); }
    else if (a.alt.size() == 1) { w.cigar_ops += 'M'; w.cigar_ops.append(a.ref.size() - 1, 'D'); }
    else { w.cigar_ops.append(a.ref.size(), 'D'); w.cigar_ops.append(a.alt.size(), 'I'); }
    ref_pos += (int)a.ref.size(); base_off += (int)a.alt.size();
  }
  const int tail = (int)w.bases.size() - base_off;
  if (tail < 0) bad("variant window: alleles extend past the window");
  w.cigar_ops.append((size_t)tail, 'M');
  return w;
}

// Every window of one cluster of nearby variants, in the reference's order: reChunk (:343-347) x alleleCombos (:351-369, 377-399).
void cluster_windows(const std::vector<const VcfRecord*>& cluster, int ref_idx, const ContigRef& ref, int padding, int max_variants, std::vector<VariantWindow>& out) {
  for (size_t t = 0; t < cluster.size(); ++t) {
    std::vector<const VcfRecord*> sub;
    for (size_t u = t; u < cluster.size() && cluster[u]->pos - cluster[t]->end() <= padding; ++u) sub.push_back(cluster[u]);
    if ((int)sub.size() > max_variants) {                                                                  // :352-356
      for (size_t a = 0; a < sub[0]->alts.size(); ++a) out.push_back(build_variant_window({ sub[0] }, { (int)a + 1 }, ref_idx, ref, padding));
      continue;
    }
    std::vector<int> counts; for (auto* v : sub) counts.push_back(1 + (int)v->alts.size());
    std::vector<std::vector<int>> combos; allele_vectors(counts, combos);
    for (auto& combo : combos) {
      std::vector<const VcfRecord*> sv; std::vector<int> sa;
      for (size_t k = 0; k < sub.size(); ++k) if (combo[k] != 0) { sv.push_back(sub[k]); sa.push_back(combo[k]); }
      if (sv.empty()) continue;
      bool valid = true;                                                                                   // VariantSet.isValid :182-193
      for (size_t k = 0; k + 1 < sv.size() && valid; ++k) if (sv[k]->pos <= sv[k + 1]->end() && sv[k + 1]->pos <= sv[k]->end()) valid = false;
      if (valid) out.push_back(build_variant_window(sv, sa, ref_idx, ref, padding));
    }
  }
}

std::vector<VariantWindow> variant_windows(const calitas_genome_view& g, const std::vector<VcfRecord>& all, int chrom_idx, int padding, int max_variants) {
  std::vector<const VcfRecord*> vs;
  for (auto& v : all) if (chrom_idx < 0 || v.chrom == g.names[chrom_idx]) vs.push_back(&v);
  // pass 1 (sequential, cheap): clusters of variants within `padding` of each other, each on its contig — nextChunk :326-337, contig walk :251
  struct Cluster { int ref_idx; size_t begin, end; };
  std::vector<Cluster> clusters;
  int ref_idx = chrom_idx >= 0 ? chrom_idx : 0;
  for (size_t i = 0; i < vs.size();) {
    const size_t first = i; const VcfRecord* last = vs[i++];
    while (i < vs.size() && vs[i]->chrom == last->chrom && vs[i]->pos <= last->end() + padding) last = vs[i++];
    while (vs[first]->chrom != g.names[ref_idx]) { if (++ref_idx >= g.n_contigs) bad("VCF contig " + vs[first]->chrom + " not found in FASTA order"); }
    clusters.push_back(Cluster{ ref_idx, first, i });
  }
  // pass 2: clusters are independent -> built on all host threads, concatenated in order
  std::vector<std::vector<VariantWindow>> per((size_t)clusters.size());
  parallel_for((int64_t)clusters.size(), 512, [&](int64_t b, int64_t e) {
    for (int64_t c = b; c < e; ++c) {
      const Cluster& cl = clusters[(size_t)c];
      std::vector<const VcfRecord*> cluster(vs.begin() + (long)cl.begin, vs.begin() + (long)cl.end);
      const ContigRef ref{ (const char*)g.bases[cl.ref_idx], g.lengths[cl.ref_idx] };
      cluster_windows(cluster, cl.ref_idx, ref, padding, max_variants, per[(size_t)c]);
    } });
  size_t total = 0; for (auto& v : per) total += v.size();
  std::vector<VariantWindow> out; out.reserve(total);
  for (auto& v : per) for (auto& w : v) out.push_back(std::move(w));
  return out;
}

Str core_parameters_search(const calitas_search_options& o, const calitas_costs& c) {   // SearchReference.scala:496-508
  const calitas_limits& l = o.limits;
  const int maxtot = l.max_total_diffs >= 0 ? l.max_total_diffs : l.max_guide_diffs + l.max_gaps_between_guide_and_pam + l.max_pam_mismatches;
  std::vector<Str> kv = { "max-variants=" + std::to_string(o.max_variants), "window-size=" + std::to_string(o.window_size), "max-guide-diffs=" + std::to_string(l.max_guide_diffs),
    "max-pam-mismatches=" + std::to_string(l.max_pam_mismatches), "max-gaps-between-guide-and-pam=" + std::to_string(l.max_gaps_between_guide_and_pam),
    "max-total-diffs=" + std::to_string(maxtot), "max-overlap=" + std::to_string(l.max_overlap), "guide-mismatch-net-cost=" + std::to_string(c.mismatch_net_cost),
    "pam-mismatch-net-cost=" + std::to_string(c.pam_mismatch_net_cost), "genome-gap-net-cost=" + std::to_string(c.genome_gap_net_cost), "guide-gap-net-cost=" + std::to_string(c.guide_gap_net_cost) };
  std::sort(kv.begin(), kv.end()); Str s; for (size_t i = 0; i < kv.size(); ++i) { if (i) s += ';'; s += kv[i]; } return s;
}
Str core_parameters_a2r(const calitas_a2r_options& o, const calitas_costs& c) {         // AlignToReference.scala:77-86
  auto opt = [](int v) { return v >= 0 ? "Some(" + std::to_string(v) + ")" : Str("None"); };
  std::vector<Str> kv = { "max-guide-diffs=" + opt(o.max_guide_diffs), "max-pam-mismatches=" + opt(o.max_pam_mismatches), "max-gaps-between-guide-and-pam=" + std::to_string(o.max_gaps_between_guide_and_pam),
    "max-overlap=" + opt(o.max_overlap), "guide-mismatch-net-cost=" + std::to_string(c.mismatch_net_cost), "pam-mismatch-net-cost=" + std::to_string(c.pam_mismatch_net_cost),
    "genome-gap-net-cost=" + std::to_string(c.genome_gap_net_cost), "guide-gap-net-cost=" + std::to_string(c.guide_gap_net_cost) };
  std::sort(kv.begin(), kv.end()); Str s; for (size_t i = 0; i < kv.size(); ++i) { if (i) s += ';'; s += kv[i]; } return s;
}

Str render_rows(const std::vector<HitX>& hits, const GuideDef& gd, const Str& chrom, const Str& target_fwd_base, int target_offset, const calitas_genome_view* genome) {
  Str text = alignment_header();
  for (auto& h : hits) {
    Str fwd;
    if (h.contig_idx >= 0) fwd = contig_slice(*genome, h.contig_idx, h.start_offset, h.end_offset);
    else fwd = target_fwd_base.substr((size_t)(h.start_offset - target_offset), (size_t)(h.end_offset - h.start_offset));
    Rendered r = render_hit(h, gd, fwd, false);
    text += alignment_row(h, r, chrom);
  }
  return text;
}

// ---- SearchReference -v on the device: what calitas_search_variants needs, built from the host-side variant windows -------------------------
// Variant sets are named by numbers that are equal iff the sets are equal (all the device needs to group hits): a single allele gets
// 1 + its index among all (record, ALT) pairs of the VCF, larger sets are interned.  The reference orders rows that tie on
// (contig, start, strand, score) across variant groups by a HashMap; the oracle orders them by the group key's text: fix_variant_ties does that on the
// few rows concerned instead of ranking millions of descriptions up front.
struct VcfDevicePlan {
  std::vector<VcfRecord> recs;
  std::vector<std::vector<VariantWindow>> by_class;            // host windows per padding class
  std::vector<int32_t> guide_class;
  std::vector<const VariantWindow*> flat;                       // class-major
  std::vector<int32_t> flat_class;
  std::vector<calitas_variant_allele> alleles; std::vector<uint32_t> set_id;
  std::vector<int32_t> first_allele, first_set;
  std::vector<std::vector<calitas_variant_window>> eng_windows; std::vector<std::vector<int32_t>> eng_flat;     // per engine: the windows it processes, and their index in `flat`
  std::vector<calitas_variant_set*> eng_set;                    // per engine: the windows on its device (loaded by the first search)
  std::mutex mu;
  VcfDevicePlan() {}
  VcfDevicePlan(const VcfDevicePlan&) = delete;
  ~VcfDevicePlan() { for (auto* v : eng_set) calitas_variant_set_free(v); }
  const calitas_variant_set* device_set(int s, calitas_engine* e, const calitas_reference* ref) {
    std::lock_guard<std::mutex> lk(mu);
    if (eng_set.size() < eng_windows.size()) eng_set.resize(eng_windows.size(), nullptr);
    if (!eng_set[(size_t)s]) ck(calitas_variant_set_load(e, ref, (int64_t)eng_windows[(size_t)s].size(), eng_windows[(size_t)s].data(), (int64_t)alleles.size(), alleles.data(),
                                                         (int64_t)set_id.size(), set_id.data(), &eng_set[(size_t)s]));
    return eng_set[(size_t)s];
  }
};
extern "C" int calitas_reference_own_range(const calitas_reference* r, int32_t contig, int64_t* own_begin, int64_t* own_end);

void build_vcf_device_plan(VcfDevicePlan& P, const calitas_genome_view& genome, const calitas_search_options& opt, const std::vector<GuideDef>& defs, int chrom_idx,
                           int n_engines, const calitas_reference* const* refs) {
  PhaseTimer pt;
  P.recs = parse_vcf(opt.vcf_text);
  pt.lap("  vcf: parse");
  std::map<int, int> class_of_padding;
  for (size_t g = 0; g < defs.size(); ++g) {
    const int padding = defs[g].length() - 1 + opt.limits.max_guide_diffs + opt.limits.max_gaps_between_guide_and_pam;      // SearchReference.scala:575
    auto it = class_of_padding.find(padding);
    if (it == class_of_padding.end()) { it = class_of_padding.emplace(padding, (int)P.by_class.size()).first; P.by_class.push_back(variant_windows(genome, P.recs, chrom_idx, padding, opt.max_variants)); }
    P.guide_class.push_back(it->second);
  }
  pt.lap("  vcf: variant windows");
  // allele numbers: index of the (record, ALT) pair in the VCF; the same pair in different windows is the same variant
  std::vector<uint32_t> alt_base(P.recs.size() + 1, 0);
  for (size_t r = 0; r < P.recs.size(); ++r) alt_base[r + 1] = alt_base[r] + (uint32_t)P.recs[r].alts.size();
  const uint32_t n_single = alt_base.back();
  std::map<std::vector<uint32_t>, uint32_t> set_no;
  auto number_of = [&](int, const VariantAllele& a) -> uint32_t { return alt_base[(size_t)((const VcfRecord*)a.rec - P.recs.data())] + (uint32_t)a.alt_idx; };
  for (size_t c = 0; c < P.by_class.size(); ++c) for (auto& w : P.by_class[c]) { P.flat.push_back(&w); P.flat_class.push_back((int32_t)c); }
  // offsets first (cheap), then the alleles and the single-allele sets on all host threads; only the sets of several alleles are interned one by one
  const size_t nw = P.flat.size();
  P.first_allele.resize(nw); P.first_set.resize(nw);
  { size_t na = 0, ns = 0; for (size_t i = 0; i < nw; ++i) { const size_t m = P.flat[i]->alleles.size(); P.first_allele[i] = (int32_t)na; P.first_set[i] = (int32_t)ns; na += m; ns += m * (m + 1) / 2;
      if (na >= (1ull << 31) || ns >= (1ull << 31)) bad("too many variant alleles / variant sets"); }
    P.alleles.resize(na); P.set_id.assign(ns, 0u); }
  std::mutex multi_mu; std::vector<std::pair<size_t, std::vector<uint32_t>>> multi;       // (position in set_id, allele numbers)
  parallel_for((int64_t)nw, 4096, [&](int64_t wb, int64_t we) {
    std::vector<std::pair<size_t, std::vector<uint32_t>>> mine; std::vector<uint32_t> num;
    for (int64_t i = wb; i < we; ++i) {
      const VariantWindow& w = *P.flat[(size_t)i]; const int m = (int)w.alleles.size();
      num.clear();
      for (int k = 0; k < m; ++k) { const VariantAllele& a = w.alleles[(size_t)k]; P.alleles[(size_t)P.first_allele[(size_t)i] + (size_t)k] = calitas_variant_allele{ a.pos, (int32_t)a.ref.size(), (int32_t)a.alt.size() }; num.push_back(number_of(w.contig, a)); }
      size_t at = (size_t)P.first_set[(size_t)i];
      for (int a = 0; a < m; ++a) for (int b = a + 1; b <= m; ++b, ++at) {
        if (b - a == 1) P.set_id[at] = 1u + num[(size_t)a];
        else mine.emplace_back(at, std::vector<uint32_t>(num.begin() + a, num.begin() + b));
      }
    }
    if (!mine.empty()) { std::lock_guard<std::mutex> lk(multi_mu); for (auto& x : mine) multi.push_back(std::move(x)); } });
  for (auto& x : multi) {
    auto it = set_no.find(x.second); if (it == set_no.end()) it = set_no.emplace(x.second, (uint32_t)set_no.size()).first;
    P.set_id[x.first] = 1u + n_single + it->second;
  }
  pt.lap("  vcf: alleles and variant sets");
  // per engine: the windows whose first base it owns, plus those within two reference windows of its cuts (halo: they take part in removeOverlaps there, too)
  P.eng_windows.resize((size_t)n_engines); P.eng_flat.resize((size_t)n_engines);
  const int64_t halo = 2 * (int64_t)opt.window_size;
  for (int s = 0; s < n_engines; ++s) {
    std::vector<int64_t> ob((size_t)genome.n_contigs), oe((size_t)genome.n_contigs);
    for (int c = 0; c < genome.n_contigs; ++c) ck(calitas_reference_own_range(refs[s], c, &ob[(size_t)c], &oe[(size_t)c]));
    for (size_t i = 0; i < P.flat.size(); ++i) {
      const VariantWindow& w = *P.flat[i]; const int64_t st = w.start - 1; const int c = w.contig;
      const bool owned = st >= ob[(size_t)c] && st < oe[(size_t)c];
      const bool near = oe[(size_t)c] > ob[(size_t)c] && st >= ob[(size_t)c] - halo && st < oe[(size_t)c] + halo;
      if (!owned && !near) continue;
      P.eng_windows[(size_t)s].push_back(calitas_variant_window{ (const uint8_t*)w.bases.data(), (int32_t)w.bases.size(), c, w.start, (int32_t)w.alleles.size(), P.first_allele[i], P.first_set[i],
                                                                  P.flat_class[i], owned ? 1 : 0 });
      P.eng_flat[(size_t)s].push_back((int32_t)i);
    }
  }
  pt.lap("  vcf: per-engine window lists");
}

// Flanks of a variant-window hit, taken from the window where it reaches far enough (SearchReference.scala:599-612); `h` holds window offsets.
Flanks window_flanks(const VariantWindow& w, const HitX& h) {
  const int wl = (int)w.bases.size();
  Flanks raw;   // window-orientation flanks (:599-602)
  if (h.guide_start_offset >= 10) { raw.has[0] = true; raw.v[0] = w.bases.substr((size_t)h.guide_start_offset - 10, 10); }
  if (wl - h.guide_end_offset >= 10) { raw.has[1] = true; raw.v[1] = w.bases.substr((size_t)h.guide_end_offset, 10); }
  if (h.start_offset >= 8) { raw.has[2] = true; raw.v[2] = w.bases.substr((size_t)h.start_offset - 8, 8); }
  if (wl - h.end_offset >= 8) { raw.has[3] = true; raw.v[3] = w.bases.substr((size_t)h.end_offset, 8); }
  Flanks fl = raw;
  if (h.strand == '-') {   // :604-612
    fl.has[0] = raw.has[1]; fl.v[0] = revcomp(raw.v[1]); fl.has[1] = raw.has[0]; fl.v[1] = revcomp(raw.v[0]);
    fl.has[2] = raw.has[3]; fl.v[2] = revcomp(raw.v[3]); fl.has[3] = raw.has[2]; fl.v[3] = revcomp(raw.v[2]);
  }
  return fl;
}

// variant_description of the variants of `w` a hit with reference span [so, eo] overlaps (ReferenceHit.scala:211, 231)
Str variant_description_of(const VariantWindow& w, int so, int eo) {
  Str d; bool first = true;
  for (auto& v : w.alleles) if (v.pos - 1 >= so && v.pos - 1 <= eo) { if (!first) d += ';'; d += variant_display(v); first = false; }
  return d;
}

}  // namespace

extern "C" {

int calitas_tool_align(calitas_engine* e, const calitas_guide* guide, const uint8_t* target, int32_t target_len, const char* target_name,
                       int32_t target_offset, const calitas_limits* limits, char** out_text) {
  return guarded([&]() -> int {
    if (!e || !guide || !limits || !out_text || (target_len && !target)) bad("bad arguments");
    *out_text = nullptr;
    GuideDef gd = parse_guide(*guide);
    calitas_target_task task{ 0, target, target_len, target_offset };
    HitSet hs; ck(calitas_align_targets(e, 1, guide, 1, &task, limits, 0, &hs.h));
    std::vector<HitX> hits = hs.all();
    *out_text = dup_text(render_rows(hits, gd, target_name ? target_name : "n/a", Str((const char*)target, (size_t)target_len), target_offset, nullptr));
    return CALITAS_OK;
  });
}

int calitas_tool_align_best(calitas_engine* e, const calitas_guide* guide, const uint8_t* target, int32_t target_len, int32_t max_gaps, char** out_text) {
  return guarded([&]() -> int {
    if (!e || !guide || !out_text || (target_len && !target)) bad("bad arguments");
    *out_text = nullptr;
    GuideDef gd = parse_guide(*guide);
    calitas_limits lim{ 0, 0, max_gaps, -1, 0 };
    calitas_target_task task{ 0, target, target_len, 0 };
    HitSet hs; ck(calitas_align_targets(e, 1, guide, 1, &task, &lim, 1, &hs.h));
    if (hs.n() == 0) throw ToolError{ CALITAS_ESTATE, "empty.maxBy" };                               // SequentialGuideAligner.scala:344
    int64_t best = 0; for (int64_t i = 1; i < hs.n(); ++i) if (hs.at(i)->score > hs.at(best)->score) best = i;   // maxBy keeps the first maximum
    *out_text = dup_text(render_rows({ hs.x(best) }, gd, "n/a", Str((const char*)target, (size_t)target_len), 0, nullptr));
    return CALITAS_OK;
  });
}

int calitas_tool_align_to_ref(calitas_engine* e, const calitas_reference* ref, const calitas_genome_view* genome, const calitas_guide* guide, const char* chrom,
                              int32_t pos, int32_t window_size, int32_t best, const calitas_limits* limits, char** out_text) {
  return guarded([&]() -> int {
    if (!e || !ref || !genome || !guide || !chrom || !limits || !out_text) bad("bad arguments");
    *out_text = nullptr;
    GuideDef gd = parse_guide(*guide);
    const int ci = contig_index(*genome, chrom);
    if (ci < 0) bad(Str("requirement failed: Unknown chromosome: ") + chrom);                        // SequentialGuideAligner.scala:370
    const int padding = window_size >= 0 ? window_size / 2 : gd.length() * 2;                        // :372
    const int64_t rs = std::max<int64_t>((int64_t)pos - padding, 1), re = std::min<int64_t>((int64_t)pos + padding, genome->lengths[ci]);   // :373
    calitas_region_task task{ 0, ci, rs - 1, (int32_t)std::max<int64_t>(0, re - rs + 1) };
    HitSet hs; ck(calitas_align_regions(e, ref, 1, guide, 1, &task, limits, best, &hs.h));
    std::vector<HitX> hits = hs.all();
    sort_alignments(hits);                                                                           // :386
    if (best) { if (hits.empty()) throw ToolError{ CALITAS_ESTATE, "head of empty list" }; hits.resize(1); }   // :417
    *out_text = dup_text(render_rows(hits, gd, chrom, Str(), 0, genome));
    return CALITAS_OK;
  });
}

int calitas_tool_search_reference(calitas_engine* e, const calitas_reference* ref, const calitas_genome_view* genome, const calitas_guide* guide,
                                  const calitas_search_options* opt, char** out_tsv, int64_t* n_hits) {
  const char* id = opt ? opt->guide_id : nullptr;
  return calitas_tool_search_reference_batch(1, &e, &ref, genome, 1, guide, &id, opt, out_tsv, n_hits);
}

// SearchReference.execute for a batch of guides over 1..N engines (one per GPU, each holding a contig-range shard of the same genome,
// calitas_shard_plan): the engines run concurrently on host threads, shards are independent, the host concatenates per guide in shard order.
// out_fd >= 0: the table is written to that descriptor (streamed block by block when there is no VCF) and *out_tsv stays NULL.
static thread_local bool g_force_host_dedup = false;      // set for the retry after an engine's halo sentinel fired (see search_reference_batch_impl's callers)
static int search_reference_batch_impl(int32_t n_engines, calitas_engine* const* engines, const calitas_reference* const* refs, const calitas_genome_view* genome,
                                       int32_t n_guides, const calitas_guide* guides, const char* const* guide_ids, const calitas_search_options* opt,
                                       int out_fd, char** out_tsv, int64_t* n_hits, int64_t* n_bytes) {
  return guarded([&]() -> int {
    if (n_engines <= 0 || !engines || !refs || !genome || n_guides <= 0 || !guides || !opt || (out_fd < 0 && !out_tsv)) bad("bad arguments");
    for (int s = 0; s < n_engines; ++s) if (!engines[s] || !refs[s]) bad("engine or reference is NULL");
    if (out_tsv) *out_tsv = nullptr;
    // removeOverlaps + sort run on the device per engine -- with a VCF too: the variant windows' hits join the reference hits' groups there
    // (calitas_search_variants) -- except with -O <= 0 on several engines: every later hit of a group then "overlaps" (>= 0), the reference's sweep
    // (SearchReference.scala:662-672) has unbounded reach, and no halo can make a shard's sweep see what it would need; that case gathers raw hits
    // and de-duplicates on the host.
    const bool host_dedup = n_engines > 1 && (opt->limits.max_overlap <= 0 || g_force_host_dedup);
    const bool device_vcf = opt->vcf_text != nullptr && !host_dedup;      // variant windows are aligned, merged and de-duplicated by calitas_search_variants
    const bool stream = out_fd >= 0 && !host_dedup;
    int64_t streamed_bytes = 0;
    std::vector<GuideDef> defs; for (int g = 0; g < n_guides; ++g) defs.push_back(parse_guide(guides[g]));
    calitas_costs costs; ck(calitas_engine_get_costs(engines[0], &costs));
    std::vector<RowContext> cxs((size_t)n_guides);
    for (int g = 0; g < n_guides; ++g) {
      RowContext& cx = cxs[(size_t)g]; cx.genome = genome; cx.aligner_id = "CALITAS:SearchReference";
      cx.guide_id = (guide_ids && guide_ids[g]) ? guide_ids[g] : (opt->guide_id ? opt->guide_id : "");
      cx.arguments = core_parameters_search(*opt, costs); cx.time_stamp = opt->time_stamp ? opt->time_stamp : ""; cx.version = opt->aligner_version ? opt->aligner_version : "calitas-b200";
      cx.has_vcf = opt->vcf_text != nullptr; cx.vcf_id = opt->vcf_id ? opt->vcf_id : "";
    }
    int chrom_idx = -1;
    if (opt->chrom && opt->chrom[0]) { chrom_idx = contig_index(*genome, opt->chrom); if (chrom_idx < 0) bad(Str("Unknown chromosome: ") + opt->chrom); }
    const bool with_vcf = opt->vcf_text != nullptr;
    std::vector<std::vector<Row>> rows((size_t)n_guides);
    std::vector<RowConst> rcs; for (int g = 0; g < n_guides; ++g) rcs.push_back(make_row_const(cxs[(size_t)g], defs[(size_t)g]));
    const int64_t ROW_BLOCK = std::getenv("CALITAS_ROW_BLOCK") ? std::max(1, std::atoi(std::getenv("CALITAS_ROW_BLOCK"))) : 2048;   // rows per rendered text block (the variable lets tests force many blocks)
    std::vector<std::vector<Str>> row_text((size_t)n_guides); int64_t n_final = 0;      // plain runs: rendered text blocks per guide
    PhaseTimer pt;
    // every engine's work runs on its own host thread; errors travel back as (code, message)
    auto run_all = [&](const std::function<void(int)>& job) {
      std::vector<int> rc((size_t)n_engines, CALITAS_OK); std::vector<Str> msg((size_t)n_engines);
      auto body = [&](int s) { try { job(s); } catch (const ToolError& te) { rc[(size_t)s] = te.code; msg[(size_t)s] = te.msg; } catch (const std::exception& ex) { rc[(size_t)s] = CALITAS_ESTATE; msg[(size_t)s] = ex.what(); } };
      if (n_engines == 1) body(0);
      else { std::vector<std::thread> th; for (int s = 0; s < n_engines; ++s) th.emplace_back(body, s); for (auto& t : th) t.join(); }
      for (int s = 0; s < n_engines; ++s) if (rc[(size_t)s] != CALITAS_OK) throw ToolError{ rc[(size_t)s], msg[(size_t)s] };
    };
    {  // reference windows: SearchReference.scala:527-564 (+ removeOverlaps/sort on the device when there is no VCF)
      // The streaming path searches GUIDE_BATCH guides per device call and renders a batch while the next one is searched on a helper thread:
      // the first call's buffer set-up (page-locking the result buffer costs ~0.5 s per GB) shrinks with the batch and hides behind rendering,
      // and the pinned result memory is two batches instead of the whole run.  Other paths search all guides in one call.
      const int GUIDE_BATCH = stream ? 32 : n_guides;
      VcfDevicePlan vplan;
      if (device_vcf) { build_vcf_device_plan(vplan, *genome, *opt, defs, chrom_idx, n_engines, refs); pt.lap("vcf: parse, windows, variant sets"); }
      auto search_batch = [&](int g0, int g1, std::vector<HitSet>& into) {
        run_all([&](int s) {
          if (device_vcf) ck(calitas_search_variants(engines[s], refs[s], g1 - g0, guides + g0, vplan.guide_class.data() + g0, &opt->limits, opt->window_size, opt->chrom,
                                                     vplan.device_set(s, engines[s], refs[s]), &into[(size_t)s].h));
          else ck(calitas_search(engines[s], refs[s], g1 - g0, guides + g0, &opt->limits, opt->window_size, opt->chrom, host_dedup ? 0 : 1, &into[(size_t)s].h)); });
      };
      std::vector<HitSet> hs((size_t)n_engines);
      search_batch(0, std::min(GUIDE_BATCH, n_guides), hs);
      pt.lap("reference search (device, first batch)");
      for (int g0 = 0; g0 < n_guides; g0 += GUIDE_BATCH) {
        const int g1 = std::min(n_guides, g0 + GUIDE_BATCH);
        std::vector<HitSet> next_hs((size_t)n_engines); std::thread prefetch; int pf_code = CALITAS_OK; Str pf_msg;
        if (g1 < n_guides) prefetch = std::thread([&]() {
          try { search_batch(g1, std::min(n_guides, g1 + GUIDE_BATCH), next_hs); }
          catch (const ToolError& te) { pf_code = te.code; pf_msg = te.msg; } catch (const std::exception& ex) { pf_code = CALITAS_ESTATE; pf_msg = ex.what(); } });
        struct Joiner { std::thread& t; ~Joiner() { if (t.joinable()) t.join(); } } joiner{ prefetch };      // an exception below must not leave the helper running
        const Flanks none;
        std::vector<int64_t> cursor((size_t)n_engines, 0);          // every hit set is guide-major; guide_idx counts from the batch's first guide
        const int rec_words = calitas_hitset_stride(hs[0].h) / 4;     // one record size per call: the same guides and limits went to every engine
        for (int g = g0; g < g1; ++g) {
          std::vector<HitRef> order;                                // this guide's hits: shard 0's, then shard 1's, ... = ReferenceHit.sort order
          for (int s = 0; s < n_engines; ++s) {
            int64_t& i = cursor[(size_t)s]; const HitSet& h_ = hs[(size_t)s]; const size_t seg = order.size();
            const calitas_variant_hit_info* vi = calitas_hitset_variant_info(h_.h);
            for (; i < h_.n() && calitas_hit_guide_idx(h_.at(i)) == g - g0; ++i) order.push_back(HitRef{ h_.at(i), vi ? vi + i : nullptr, s });
            if (!host_dedup) merge_at_cut(order, seg);
          }
          // the host window behind a variant-window hit
          auto window_of = [&](const HitRef& r) -> const VariantWindow& { return *vplan.flat[(size_t)vplan.eng_flat[(size_t)r.engine][(size_t)r.v->window_idx]]; };
          if (device_vcf) {   // rows tying on (contig, start, strand, score) across variant groups: the oracle's order is the group key's text (see VcfDevicePlan)
            for (size_t a = 0; a < order.size();) {
              size_t b = a + 1; while (b < order.size() && !hit_sorts_before(order[a], order[b]) && !hit_sorts_before(order[b], order[a])) ++b;
              bool mixed = false; for (size_t k = a + 1; k < b; ++k) if (order[k].v->set_rank != order[a].v->set_rank) mixed = true;
              if (mixed) {
                std::vector<std::pair<Str, HitRef>> run;
                for (size_t k = a; k < b; ++k) run.emplace_back(order[k].v->window_idx >= 0 && order[k].v->set_rank ? variant_description_of(window_of(order[k]), order[k].v->start_offset, order[k].v->end_offset) : Str(), order[k]);
                std::stable_sort(run.begin(), run.end(), [](const std::pair<Str, HitRef>& x, const std::pair<Str, HitRef>& y) { return x.first < y.first; });
                for (size_t k = a; k < b; ++k) order[k] = run[k - a].second;
              }
              a = b;
            }
          }
          const GuideDef& gd = defs[(size_t)g]; const RowContext& cx = cxs[(size_t)g]; const RowConst& rc = rcs[(size_t)g];
          // without a VCF the device has already de-duplicated and sorted: rows are final, rendered straight into text blocks of ROW_BLOCK rows
          std::vector<Row>& out = rows[(size_t)g]; if (host_dedup) out.resize(order.size());
          std::vector<Str>& blocks = row_text[(size_t)g]; if (!host_dedup && !stream) blocks.resize((order.size() + ROW_BLOCK - 1) / ROW_BLOCK);
          n_final += host_dedup ? 0 : (int64_t)order.size();
          auto render = [&](int64_t b, int64_t e_, Str* blk) {
            if (blk) blk->reserve((size_t)(e_ - b) * 640);
            RenderedFix r;
            for (int64_t k = b; k < e_; ++k) {
              HitX h; unpack_hit(reinterpret_cast<const uint32_t*>(order[(size_t)k].h), rec_words, h);
              const Str& gt = guide_text_of(rc, h.pam_idx);
              if (order[(size_t)k].v && order[(size_t)k].v->window_idx >= 0) {        // hit of a variant window (SearchReference.scala:596-622): bases and flanks come from the window
                const calitas_variant_hit_info& v = *order[(size_t)k].v; const VariantWindow& w = window_of(order[(size_t)k]);
                render_hit_fix(h, gt.data(), (int)gt.size(), w.bases.data() + h.start_offset, h.end_offset - h.start_offset, false, r);
                const Row row = make_row(cx, rc, h, r, gd, w.contig, v.start_offset, v.end_offset, v.guide_start_offset, v.guide_end_offset, w.alleles, window_flanks(w, h));
                if (blk) *blk += row.line; else out[(size_t)k] = row;
                continue;
              }
              render_hit_fix(h, gt.data(), (int)gt.size(), (const char*)genome->bases[h.contig_idx] + h.start_offset, h.end_offset - h.start_offset, true, r);   // windows are upper-cased, SearchReference.scala:67
              if (blk) write_row(*blk, cx, rc, gd, h, r, h.contig_idx, h.start_offset, h.end_offset, h.guide_start_offset, h.guide_end_offset, nullptr, none);
              else out[(size_t)k] = make_row(cx, rc, h, r, gd, h.contig_idx, h.start_offset, h.end_offset, h.guide_start_offset, h.guide_end_offset, {}, none);
            } };
          const int64_t n_rows = (int64_t)order.size();
          if (host_dedup) parallel_for(n_rows, ROW_BLOCK, [&](int64_t b, int64_t e_) { render(b, e_, nullptr); });
          else if (stream) {
            if (g == 0) { const Str header = hit_header(); write_all(out_fd, header.data(), header.size()); streamed_bytes += (int64_t)header.size(); }
            ordered_pipeline((n_rows + ROW_BLOCK - 1) / ROW_BLOCK, 256,
                             [&](int64_t k, Str& blk) { render(k * ROW_BLOCK, std::min(n_rows, (k + 1) * ROW_BLOCK), &blk); },
                             [&](const Str& blk) { write_all(out_fd, blk.data(), blk.size()); streamed_bytes += (int64_t)blk.size(); });
          }
          else parallel_for((int64_t)blocks.size(), 1, [&](int64_t bb, int64_t be) { for (int64_t k = bb; k < be; ++k) render(k * ROW_BLOCK, std::min(n_rows, (k + 1) * ROW_BLOCK), &blocks[(size_t)k]); });
        }
        for (int s = 0; s < n_engines; ++s) if (cursor[(size_t)s] != hs[(size_t)s].n()) throw ToolError{ CALITAS_ESTATE, "hit set is not guide-major" };
        if (prefetch.joinable()) prefetch.join();                  // the engines are idle again: only now may this batch's hit sets go back to their pools
        if (pf_code != CALITAS_OK) throw ToolError{ pf_code, pf_msg };
        for (int s = 0; s < n_engines; ++s) std::swap(hs[(size_t)s].h, next_hs[(size_t)s].h);
      }
      pt.lap("reference rows");
    }
    if (with_vcf && host_dedup) {  // SearchReference.scala:570-630 on the host (several engines with -O <= 0)
      std::vector<VcfRecord> recs = parse_vcf(opt->vcf_text);
      pt.lap("parse vcf");
      // variant windows depend on the guide only through the padding (Guide.length, :575): build once per distinct padding
      std::map<int, std::vector<VariantWindow>> by_padding;
      std::vector<const std::vector<VariantWindow>*> windows_of((size_t)n_guides);
      for (int g = 0; g < n_guides; ++g) {
        const int padding = defs[(size_t)g].length() - 1 + opt->limits.max_guide_diffs + opt->limits.max_gaps_between_guide_and_pam;
        auto it = by_padding.find(padding);
        if (it == by_padding.end()) it = by_padding.emplace(padding, variant_windows(*genome, recs, chrom_idx, padding, opt->max_variants)).first;
        windows_of[(size_t)g] = &it->second;
      }
      pt.lap("variant windows");
      struct TaskRef { int guide; int window; };
      std::vector<std::vector<calitas_target_task>> tasks((size_t)n_engines); std::vector<std::vector<TaskRef>> task_ref((size_t)n_engines);
      int64_t rr = 0;
      for (int g = 0; g < n_guides; ++g) { const auto& ws = *windows_of[(size_t)g];
        for (size_t w = 0; w < ws.size(); ++w, ++rr) { const size_t s = (size_t)(rr % n_engines);
          tasks[s].push_back(calitas_target_task{ g, (const uint8_t*)ws[w].bases.data(), (int32_t)ws[w].bases.size(), 0 }); task_ref[s].push_back(TaskRef{ g, (int)w }); } }
      std::vector<HitSet> hs((size_t)n_engines);
      run_all([&](int s) { if (!tasks[(size_t)s].empty()) ck(calitas_align_targets(engines[s], n_guides, guides, (int64_t)tasks[(size_t)s].size(), tasks[(size_t)s].data(), &opt->limits, 0, &hs[(size_t)s].h)); });
      pt.lap("variant windows (device)");
      // rows must arrive in the reference's window order per guide (it decides ties in removeOverlaps): walk tasks in global order
      std::vector<int64_t> cursor((size_t)n_engines, 0); std::vector<int64_t> next_task((size_t)n_engines, 0);
      rr = 0;
      for (int g = 0; g < n_guides; ++g) { const auto& ws = *windows_of[(size_t)g];
        for (size_t wi = 0; wi < ws.size(); ++wi, ++rr) {
          const size_t s = (size_t)(rr % n_engines); const int64_t t = next_task[s]++; int64_t& i = cursor[s]; const HitSet& h_ = hs[s];
          const VariantWindow& w = ws[wi]; const GuideDef& gd = defs[(size_t)g];
          for (; i < h_.n() && h_.at(i)->task_idx == t; ++i) {
            const HitX h = h_.x(i);
            const Str& gt = guide_text_of(rcs[(size_t)g], h.pam_idx);
            RenderedFix r; render_hit_fix(h, gt.data(), (int)gt.size(), w.bases.data() + h.start_offset, h.end_offset - h.start_offset, false, r);
            const Flanks fl = window_flanks(w, h);
            const int so = w.ref_offset_at(h.start_offset, true), eo = w.ref_offset_at(h.end_offset, false);             // :615-620
            const int gso = w.ref_offset_at(h.guide_start_offset, true), geo = w.ref_offset_at(h.guide_end_offset, false);
            rows[(size_t)g].push_back(make_row(cxs[(size_t)g], rcs[(size_t)g], h, r, gd, w.contig, so, eo, gso, geo, w.alleles, fl));
          }
        } }
      for (int s = 0; s < n_engines; ++s) if (cursor[(size_t)s] != hs[(size_t)s].n()) throw ToolError{ CALITAS_ESTATE, "variant hit set is not task-major" };
      pt.lap("variant rows");
    }
    if (host_dedup) {
      for (int g = 0; g < n_guides; ++g) {
        rows[(size_t)g] = remove_overlaps_host(rows[(size_t)g], opt->limits.max_overlap);                                 // :641
        sort_rows(rows[(size_t)g]);                                                                                      // :647
      }
      pt.lap("removeOverlaps + sort (host)");
    }
    if (stream) { if (n_hits) *n_hits = n_final; if (n_bytes) *n_bytes = streamed_bytes; return CALITAS_OK; }
    const Str header = hit_header(); int64_t total = n_final; size_t bytes = header.size();
    for (auto& v : rows) { for (auto& r : v) bytes += r.line.size(); total += (int64_t)v.size(); }
    std::vector<Str*> blocks; std::vector<size_t> at;                                                                    // text blocks of the plain run, in table order
    for (auto& v : row_text) for (auto& b : v) { blocks.push_back(&b); at.push_back(bytes); bytes += b.size(); }
    char* text = (char*)std::malloc(bytes + 1); if (!text) throw ToolError{ CALITAS_ESTATE, "out of memory for the hit table" };
    char* w = text; std::memcpy(w, header.data(), header.size()); w += header.size();
    for (auto& v : rows) for (auto& r : v) { std::memcpy(w, r.line.data(), r.line.size()); w += r.line.size(); }
    parallel_for((int64_t)blocks.size(), 16, [&](int64_t b, int64_t e_) { for (int64_t k = b; k < e_; ++k) { Str& s = *blocks[(size_t)k]; std::memcpy(text + at[(size_t)k], s.data(), s.size()); Str().swap(s); } });
    text[bytes] = 0;
    pt.lap("table assembly");
    if (n_hits) *n_hits = total;
    if (n_bytes) *n_bytes = (int64_t)bytes;
    if (out_fd >= 0) { write_all(out_fd, text, bytes); std::free(text); }
    else *out_tsv = text;
    return CALITAS_OK;
  });
}

int calitas_tool_search_reference_batch(int32_t n_engines, calitas_engine* const* engines, const calitas_reference* const* refs, const calitas_genome_view* genome,
                                        int32_t n_guides, const calitas_guide* guides, const char* const* guide_ids, const calitas_search_options* opt,
                                        char** out_tsv, int64_t* n_hits) {
  int rc = search_reference_batch_impl(n_engines, engines, refs, genome, n_guides, guides, guide_ids, opt, -1, out_tsv, n_hits, nullptr);
  if (rc == CALITAS_ELIMIT && n_engines > 1 && std::strstr(calitas_last_error(), "removeOverlaps: a chain")) {
    // an engine's halo sentinel: this input chains overlapping hits through a whole shard halo; gather raw hits and run removeOverlaps + sort on the host instead
    g_force_host_dedup = true;
    rc = search_reference_batch_impl(n_engines, engines, refs, genome, n_guides, guides, guide_ids, opt, -1, out_tsv, n_hits, nullptr);
    g_force_host_dedup = false;
  }
  return rc;
}
int calitas_tool_search_reference_batch_fd(int32_t n_engines, calitas_engine* const* engines, const calitas_reference* const* refs, const calitas_genome_view* genome,
                                           int32_t n_guides, const calitas_guide* guides, const char* const* guide_ids, const calitas_search_options* opt,
                                           int32_t out_fd, int64_t* n_hits, int64_t* n_bytes) {
  if (out_fd < 0) return calitas_tools_set_error(CALITAS_EINVAL, "bad output descriptor");
  return search_reference_batch_impl(n_engines, engines, refs, genome, n_guides, guides, guide_ids, opt, out_fd, nullptr, n_hits, n_bytes);
}

int calitas_tool_align_to_reference(calitas_engine* e, const calitas_reference* ref, const calitas_genome_view* genome, int64_t n_tasks,
                                    const calitas_a2r_task* tasks, const calitas_a2r_options* opt, char** out_tsv, int64_t* n_hits) {
  return guarded([&]() -> int {
    if (!e || !ref || !genome || !opt || !out_tsv || (n_tasks && !tasks)) bad("bad arguments");
    *out_tsv = nullptr;
    const int given = (opt->max_guide_diffs >= 0) + (opt->max_pam_mismatches >= 0) + (opt->max_overlap >= 0);
    if (given != 0 && given != 3) bad("Must specify all or none of: --max-guide-diffs, --max-pam-mismatches, --max-overlap");   // AlignToReference.scala:88-92
    const bool best = given == 0;
    calitas_costs costs; ck(calitas_engine_get_costs(e, &costs));
    RowContext cx; cx.genome = genome; cx.aligner_id = "CALITAS:AlignToReference"; cx.arguments = core_parameters_a2r(*opt, costs);
    cx.time_stamp = opt->time_stamp ? opt->time_stamp : ""; cx.version = opt->aligner_version ? opt->aligner_version : "calitas-b200";
    calitas_limits lim{ best ? 0 : opt->max_guide_diffs, best ? 0 : opt->max_pam_mismatches, opt->max_gaps_between_guide_and_pam,
                        best ? -1 : (opt->max_total_diffs >= 0 ? opt->max_total_diffs : opt->max_guide_diffs + opt->max_gaps_between_guide_and_pam + opt->max_pam_mismatches),
                        best ? 0 : opt->max_overlap };
    // One device call per span of tasks (bounded only by the engine's guide table): the reference's batches of 10 000 rows
    // (AlignToReference.scala:110) matter for the ORDER of the output — each batch is sorted on its own (:141) — not for how the work is cut.
    Str text = hit_header(); int64_t total = 0;
    const Flanks none;
    const int64_t MAX_DISTINCT = 16000;
    for (int64_t s0 = 0; s0 < n_tasks;) {
      std::map<Str, int> guide_of; std::vector<Str> queries; std::vector<GuideDef> defs; std::vector<calitas_region_task> rt;
      int64_t s1 = s0;
      for (; s1 < n_tasks; ++s1) {
        const calitas_a2r_task& t = tasks[s1];
        if (!t.query || !t.chrom) bad("task query/chrom is NULL");
        auto it = guide_of.find(t.query);
        if (it == guide_of.end()) {
          if ((int64_t)queries.size() >= MAX_DISTINCT && (s1 - s0) % 10000 == 0) break;                  // cut only on a batch boundary
          it = guide_of.emplace(t.query, (int)queries.size()).first; queries.push_back(t.query); defs.push_back(parse_guide(t.query, {}));   // :112
        }
        const GuideDef& gd = defs[(size_t)it->second];
        const int ci = contig_index(*genome, t.chrom);
        if (ci < 0) bad(Str("requirement failed: Unknown chromosome: ") + t.chrom);
        const int padding = opt->window_size >= 0 ? opt->window_size / 2 : gd.length() * 2;
        const int64_t rs = std::max<int64_t>((int64_t)t.position - padding, 1), re = std::min<int64_t>((int64_t)t.position + padding, genome->lengths[ci]);
        rt.push_back(calitas_region_task{ it->second, ci, rs - 1, (int32_t)std::max<int64_t>(0, re - rs + 1) });
      }
      std::vector<calitas_guide> cg; for (auto& q : queries) cg.push_back(calitas_guide{ q.c_str(), nullptr, 0 });
      HitSet hs; ck(calitas_align_regions(e, ref, (int32_t)cg.size(), cg.data(), (int64_t)rt.size(), rt.data(), &lim, best ? 1 : 0, &hs.h));
      // hits arrive grouped by task in retval order: per task apply `.sorted` (+ `.head` in best mode), render on all host threads
      const int64_t nt = s1 - s0;
      std::vector<int64_t> first((size_t)nt + 1, 0);
      { int64_t i = 0; for (int64_t t = 0; t < nt; ++t) { first[(size_t)t] = i; while (i < hs.n() && hs.at(i)->task_idx == t) ++i; } first[(size_t)nt] = i; if (i != hs.n()) throw ToolError{ CALITAS_ESTATE, "hit set is not task-major" }; }
      if (best) for (int64_t t = 0; t < nt; ++t) if (first[(size_t)t] == first[(size_t)t + 1])          // alignToRefBest(...).head on an empty result throws in the reference
        throw ToolError{ CALITAS_ESTATE, Str("head of empty list: no alignment for query ") + tasks[s0 + t].query };
      std::vector<std::vector<Row>> task_rows((size_t)nt);
      parallel_for(nt, 256, [&](int64_t tb, int64_t te) {
        for (int64_t t = tb; t < te; ++t) {
          std::vector<HitX> alns; for (int64_t i = first[(size_t)t]; i < first[(size_t)t + 1]; ++i) alns.push_back(hs.x(i));
          sort_alignments(alns);
          if (best) alns.resize(1);
          const GuideDef& gd = defs[(size_t)rt[(size_t)t].guide_idx];
          RowContext c2 = cx; c2.guide_id = tasks[s0 + t].id ? tasks[s0 + t].id : tasks[s0 + t].query;    // :100
          const RowConst rc = make_row_const(c2, gd);
          RenderedFix r;
          for (auto& h : alns) {
            const Str& gt = guide_text_of(rc, h.pam_idx);
            render_hit_fix(h, gt.data(), (int)gt.size(), (const char*)genome->bases[h.contig_idx] + h.start_offset, h.end_offset - h.start_offset, false, r);   // region is not upper-cased (SequentialGuideAligner.scala:374)
            task_rows[(size_t)t].push_back(make_row(c2, rc, h, r, gd, h.contig_idx, h.start_offset, h.end_offset, h.guide_start_offset, h.guide_end_offset, {}, none));
          }
        } });
      for (int64_t b0 = 0; b0 < nt; b0 += 10000) {                                                     // ReferenceHit.sort per batch of 10 000 input rows (:110,141)
        std::vector<Row> rows;
        for (int64_t t = b0; t < std::min<int64_t>(nt, b0 + 10000); ++t) for (auto& r : task_rows[(size_t)t]) rows.push_back(std::move(r));
        sort_rows(rows);
        for (auto& r : rows) text += r.line;
        total += (int64_t)rows.size();
      }
      s0 = s1;
    }
    if (n_hits) *n_hits = total;
    *out_tsv = dup_text(text);
    return CALITAS_OK;
  });
}

// PairwiseAlignSequences.execute (PairwiseAlignSequences.scala:44-83): alignBest of each (query, upper-cased target) pair, 11 columns.
// The reference calls alignBest(guide, target) without forwarding its own -g/-O flags, so the gap limit is always the default 3.
int calitas_tool_pairwise_align(calitas_engine* e, int64_t n_pairs, const char* const* queries, const char* const* targets, char** out_tsv) {
  return guarded([&]() -> int {
    if (!e || !out_tsv || n_pairs < 0 || (n_pairs && (!queries || !targets))) bad("bad arguments");
    *out_tsv = nullptr;
    Str text = "query\ttarget\tscore\tquery_start\ttarget_start\tcigar\tmismatches\tgap_bases\tpadded_query\talignment\tpadded_target\n";
    calitas_limits lim{ 0, 0, 3, -1, 0 };
    for (int64_t b0 = 0; b0 < n_pairs; b0 += 10000) {                                                  // :63
      const int64_t b1 = std::min<int64_t>(n_pairs, b0 + 10000);
      std::map<Str, int> guide_of; std::vector<Str> qs; std::vector<GuideDef> defs; std::vector<Str> ups; std::vector<calitas_target_task> tasks;
      for (int64_t i = b0; i < b1; ++i) {
        if (!queries[i] || !targets[i]) bad("query/target is NULL");
        auto it = guide_of.find(queries[i]);
        if (it == guide_of.end()) { it = guide_of.emplace(queries[i], (int)qs.size()).first; qs.push_back(queries[i]); defs.push_back(parse_guide(queries[i], {})); }
        ups.push_back(to_upper(targets[i]));                                                           // :53
      }
      for (int64_t i = b0; i < b1; ++i) { const Str& t = ups[(size_t)(i - b0)]; tasks.push_back(calitas_target_task{ guide_of[queries[i]], (const uint8_t*)t.data(), (int32_t)t.size(), 0 }); }
      std::vector<calitas_guide> cg; for (auto& q : qs) cg.push_back(calitas_guide{ q.c_str(), nullptr, 0 });
      HitSet hs; ck(calitas_align_targets(e, (int32_t)cg.size(), cg.data(), (int64_t)tasks.size(), tasks.data(), &lim, 1, &hs.h));
      int64_t i = 0;
      for (int64_t t = 0; t < (int64_t)tasks.size(); ++t) {
        int64_t best_i = -1;
        for (; i < hs.n() && hs.at(i)->task_idx == t; ++i) if (best_i < 0 || hs.at(i)->score > hs.at(best_i)->score) best_i = i;     // maxBy keeps the first maximum (:344)
        if (best_i < 0) throw ToolError{ CALITAS_ESTATE, Str("empty.maxBy: no alignment for ") + queries[b0 + t] };
        const HitX best_hit = hs.x(best_i); const HitX* best = &best_hit;
        const Str& tgt = ups[(size_t)t]; const GuideDef& gd = defs[(size_t)tasks[(size_t)t].guide_idx];
        Rendered r = render_hit(*best, gd, tgt.substr((size_t)best->start_offset, (size_t)(best->end_offset - best->start_offset)), false);
        text += Str(queries[b0 + t]) + "\t" + tgt + "\t" + std::to_string(best->score) + "\t1\t" + std::to_string(best->start_offset) + "\t" + r.cigar + "\t" +
                std::to_string(r.mismatches) + "\t" + std::to_string(r.gap_bases) + "\t" + r.padded_guide + "\t" + r.padded_alignment + "\t" + r.padded_target + "\n";
      }
    }
    *out_tsv = dup_text(text);
    return CALITAS_OK;
  });
}

// ---- the variant plan as an object: built once per (VCF, guide batch, engines), searched any number of times ---------------------------------------
struct calitas_variant_plan { VcfDevicePlan plan; std::vector<GuideDef> defs; };

int calitas_tool_variant_plan_create(const calitas_genome_view* genome, const calitas_search_options* opt, int32_t n_guides, const calitas_guide* guides,
                                     int32_t n_engines, const calitas_reference* const* refs, calitas_variant_plan** out) {
  return guarded([&]() -> int {
    if (!genome || !opt || !opt->vcf_text || n_guides <= 0 || !guides || n_engines <= 0 || !refs || !out) bad("bad arguments");
    *out = nullptr;
    std::unique_ptr<calitas_variant_plan> p(new calitas_variant_plan());
    for (int g = 0; g < n_guides; ++g) p->defs.push_back(parse_guide(guides[g]));
    int chrom_idx = -1; if (opt->chrom && opt->chrom[0]) { chrom_idx = contig_index(*genome, opt->chrom); if (chrom_idx < 0) bad(Str("Unknown chromosome: ") + opt->chrom); }
    build_vcf_device_plan(p->plan, *genome, *opt, p->defs, chrom_idx, n_engines, refs);
    *out = p.release();
    return CALITAS_OK;
  });
}
void calitas_tool_variant_plan_free(calitas_variant_plan* p) { delete p; }
int calitas_tool_variant_plan_counts(const calitas_variant_plan* p, int32_t engine, int64_t* n_records, int64_t* n_windows, int64_t* n_window_bases) {
  if (!p || engine < 0 || engine >= (int32_t)p->plan.eng_windows.size()) return calitas_tools_set_error(CALITAS_EINVAL, "bad arguments");
  if (n_records) *n_records = (int64_t)p->plan.recs.size();
  if (n_windows) *n_windows = (int64_t)p->plan.eng_windows[(size_t)engine].size();
  if (n_window_bases) { int64_t b = 0; for (auto& w : p->plan.eng_windows[(size_t)engine]) b += w.length; *n_window_bases = b; }
  return CALITAS_OK;
}
// calitas_search_variants for engine `engine` of the plan (the guides must be the ones the plan was built for)
int calitas_tool_variant_plan_search(calitas_variant_plan* p, int32_t engine, calitas_engine* e, const calitas_reference* ref, int32_t n_guides, const calitas_guide* guides,
                                     const calitas_limits* limits, int32_t window_size, const char* chrom, calitas_hitset** out) {
  if (!p || engine < 0 || engine >= (int32_t)p->plan.eng_windows.size() || n_guides != (int32_t)p->plan.guide_class.size()) return calitas_tools_set_error(CALITAS_EINVAL, "bad arguments");
  return guarded([&]() -> int {
    VcfDevicePlan& P = p->plan;
    ck(calitas_search_variants(e, ref, n_guides, guides, P.guide_class.data(), limits, window_size, chrom, P.device_set(engine, e, ref), out));   // the first call uploads the windows
    return CALITAS_OK;
  });
}

int calitas_tool_variant_windows(const calitas_genome_view* genome, const char* vcf_text, const char* chrom, int32_t padding, int32_t max_variants, char** out_text) {
  return guarded([&]() -> int {
    if (!genome || !vcf_text || !out_text) bad("bad arguments");
    *out_text = nullptr;
    int chrom_idx = -1; if (chrom && chrom[0]) { chrom_idx = contig_index(*genome, chrom); if (chrom_idx < 0) bad(Str("Unknown chromosome: ") + chrom); }
    std::vector<VcfRecord> recs = parse_vcf(vcf_text);
    Str text;
    for (auto& w : variant_windows(*genome, recs, chrom_idx, padding, max_variants)) {
      text += Str(genome->names[w.contig]) + "\t" + std::to_string(w.start) + "\t" + w.cigar_text() + "\t" + w.bases + "\t";
      for (size_t i = 0; i < w.alleles.size(); ++i) { if (i) text += ';'; text += variant_display(w.alleles[i]); }
      text += '\n';
    }
    *out_text = dup_text(text);
    return CALITAS_OK;
  });
}

}  // extern "C"
