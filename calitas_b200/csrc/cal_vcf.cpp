// cal_vcf.cpp — `calitas PrepareVcf`: the VCF clean-up step that feeds SearchReference -v (PrepareVcf.scala:43-91).  Host code only; no
// alignment happens here.  Restates what the reference does through fgbio's VcfSource / VcfWriter (htsjdk underneath):
//   * the header of the FIRST input is kept, without samples (PrepareVcf.scala:46-63); with -d the contig lines are replaced by the
//     dictionary's sequences and ##reference by the first sequence's assembly (:54-59)
//   * records must have FILTER == PASS (:73), an AF value >= min-af (:74; AF is a Float compared with the Double threshold, so AF=0.01
//     is 0.0099999998 and fails the default 0.01) and only plain-base alleles (:75, fgbio SimpleAllele: no '*', '.', '<SYM>', breakends)
//   * ALT alleles and their AF values below min-af are removed, genotypes and every INFO field but AF are dropped, chromosomes 1-22, X, Y
//     get a "chr" prefix unless -c false (:77-83,91)
// Not byte-pinned (no JVM here): header lines keep the input's order (htsjdk re-sorts them), and no .tbi index is written next to a .gz.
// Numbers follow htsjdk's VCFEncoder: AF through formatVCFDouble (%.3f, %.3e below 0.01, %.2f from 1), QUAL as %.2f without a trailing .00,
// both rounded half-up on the shortest decimal form of the value as java.util.Formatter does.
#include "cal_vcf.h"

#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <set>
#include <zlib.h>

namespace cal { namespace io {

namespace {

typedef std::string Str;

std::vector<Str> split_on(const Str& s, char sep) {
  std::vector<Str> out; size_t a = 0;
  for (;;) { const size_t b = s.find(sep, a); if (b == Str::npos) { out.push_back(s.substr(a)); return out; } out.push_back(s.substr(a, b - a)); a = b + 1; }
}

// Decimal digits and exponent of the shortest string that round-trips `v` (> 0): v = 0.d0 d1 d2 ... x 10^exp10
void shortest_digits(double v, Str& digits, int& exp10) {
  char buf[64]; auto r = std::to_chars(buf, buf + sizeof buf, v, std::chars_format::scientific);
  const Str s(buf, r.ptr);                                   // d.ddddde[+-]xx
  const size_t e = s.find('e'); Str mant = s.substr(0, e); const int ex = std::atoi(s.c_str() + e + 1);
  digits.clear(); for (char c : mant) if (c >= '0' && c <= '9') digits += c;
  exp10 = ex + 1;
}
// Rounds the digit string half-up to `keep` digits (keep may be <= 0); may carry into a new leading digit (exp10 grows).
void round_half_up(Str& digits, int& exp10, int keep) {
  if (keep < 0) { digits = "0"; return; }
  if ((int)digits.size() <= keep) { digits.append((size_t)(keep - (int)digits.size()), '0'); return; }
  const bool up = digits[(size_t)keep] >= '5';
  digits.resize((size_t)keep);
  if (up) {
    int i = keep - 1;
    while (i >= 0 && digits[(size_t)i] == '9') { digits[(size_t)i] = '0'; --i; }
    if (i >= 0) ++digits[(size_t)i]; else { digits.insert(digits.begin(), '1'); ++exp10; }   // the caller re-trims to its own width
  }
}
// java.util.Formatter "%.<prec>f" of a non-negative double
Str java_fixed(double v, int prec) {
  if (v == 0) { Str z = "0"; if (prec) { z += '.'; z.append((size_t)prec, '0'); } return z; }
  Str d; int e; shortest_digits(v, d, e);
  int keep = e + prec;                                      // digits before the cut
  if (keep < 0) { d = ""; keep = 0; e = -prec; }            // far below the last printed place: rounds to zero
  else { const size_t before = d.size(); (void)before; round_half_up(d, e, keep); }
  // d now holds (e + prec) digits (or one more leading digit after a carry, reflected in e)
  Str ip, fp;
  const int n_int = e > 0 ? e : 0;
  if ((int)d.size() < n_int + prec) d.insert(0, (size_t)(n_int + prec - (int)d.size()), '0');
  ip = d.substr(0, d.size() - (size_t)prec); fp = d.substr(d.size() - (size_t)prec);
  if (ip.empty()) ip = "0";
  return prec ? ip + "." + fp : ip;
}
// java.util.Formatter "%.3e" of a positive double: d.ddde[+-]xx
Str java_sci3(double v) {
  Str d; int e; shortest_digits(v, d, e);
  round_half_up(d, e, 4);
  if (d.size() > 4) d.resize(4);                            // carry produced "10000": keep "1000", exponent already bumped
  const int ex = e - 1;
  char tail[16]; std::snprintf(tail, sizeof tail, "e%c%02d", ex < 0 ? '-' : '+', ex < 0 ? -ex : ex);
  return d.substr(0, 1) + "." + d.substr(1) + tail;
}
// htsjdk VCFEncoder.formatVCFDouble
Str format_vcf_double(double d) {
  if (d < 1) {
    if (d < 0.01) { if (std::fabs(d) >= 1e-20) return d < 0 ? "-" + java_sci3(-d) : java_sci3(d); return "0.00"; }
    return java_fixed(d, 3);
  }
  return java_fixed(d, 2);
}
// htsjdk VCFEncoder.formatQualValue
Str format_qual(double q) {
  Str s = q < 0 ? "-" + java_fixed(-q, 2) : java_fixed(q, 2);
  if (s.size() > 3 && s.compare(s.size() - 3, 3, ".00") == 0) s.resize(s.size() - 3);
  return s;
}

bool simple_allele(const Str& a) {                          // fgbio Allele.apply -> SimpleAllele: bases only
  if (a.empty()) return false;
  for (char c : a) { switch (c) { case 'A': case 'C': case 'G': case 'T': case 'N': case 'a': case 'c': case 'g': case 't': case 'n': break; default: return false; } }
  return true;
}

const std::set<Str>& chroms_to_fix() {                      // PrepareVcf.scala:13-15
  static const std::set<Str> s = [] { std::set<Str> t; for (int i = 1; i <= 22; ++i) t.insert(std::to_string(i)); t.insert("X"); t.insert("Y"); return t; }();
  return s;
}

struct DictSeq { Str name, assembly; long long length; };
std::vector<DictSeq> read_dict(const Str& path) {           // SAMSequenceDictionaryExtractor on a .dict / SAM header
  std::vector<DictSeq> out;
  for (const Str& raw : split_on(gunzip_if_needed(read_file(path)), '\n')) {
    Str line = raw; if (!line.empty() && line.back() == '\r') line.pop_back();
    if (line.compare(0, 3, "@SQ") != 0) continue;
    DictSeq s; s.length = 0;
    for (const Str& f : split_on(line, '\t')) {
      if (f.compare(0, 3, "SN:") == 0) s.name = f.substr(3); else if (f.compare(0, 3, "LN:") == 0) s.length = std::atoll(f.c_str() + 3); else if (f.compare(0, 3, "AS:") == 0) s.assembly = f.substr(3);
    }
    out.push_back(s);
  }
  if (out.empty()) throw IoError{ "No sequences in sequence dictionary: " + path };
  return out;
}

void bgzf_block(const char* data, size_t n, Str& out) {
  z_stream zs; std::memset(&zs, 0, sizeof zs);
  if (deflateInit2(&zs, Z_DEFAULT_COMPRESSION, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) throw IoError{ "zlib: deflateInit2 failed" };
  std::vector<unsigned char> buf(deflateBound(&zs, (uLong)n) + 64);
  zs.next_in = (Bytef*)data; zs.avail_in = (uInt)n; zs.next_out = buf.data(); zs.avail_out = (uInt)buf.size();
  const int rc = deflate(&zs, Z_FINISH); const size_t clen = buf.size() - zs.avail_out; deflateEnd(&zs);
  if (rc != Z_STREAM_END) throw IoError{ "zlib: deflate failed" };
  const size_t total = 18 + clen + 8;
  if (total > 65536) throw IoError{ "BGZF block too large" };
  const unsigned char hdr[18] = { 31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 'B', 'C', 2, 0, (unsigned char)((total - 1) & 0xFF), (unsigned char)((total - 1) >> 8) };
  out.append((const char*)hdr, 18); out.append((const char*)buf.data(), clen);
  const uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), (const Bytef*)data, (uInt)n), isize = (uint32_t)n;
  unsigned char tail[8]; for (int i = 0; i < 4; ++i) { tail[i] = (unsigned char)(crc >> (8 * i)); tail[4 + i] = (unsigned char)(isize >> (8 * i)); }
  out.append((const char*)tail, 8);
}

}  // namespace

std::string bgzf_compress(const std::string& text) {
  Str out; const size_t BLOCK = 0xff00;
  for (size_t a = 0; a < text.size(); a += BLOCK) bgzf_block(text.data() + a, std::min(BLOCK, text.size() - a), out);
  bgzf_block("", 0, out);                                   // end-of-file marker block
  return out;
}

PrepareVcfStats prepare_vcf(const std::vector<std::string>& inputs, const std::string& output, double min_af, const std::string& dict_path, bool add_chr_prefix) {
  if (inputs.empty()) throw IoError{ "Argument 'input' is required" };
  PrepareVcfStats st;
  Str out;
  // ---- header: PrepareVcf.scala:46-63 -------------------------------------------------------------------------------------------------
  {
    const Str text = gunzip_if_needed(read_file(inputs[0]));
    std::vector<Str> meta; bool have_chrom_line = false;
    size_t a = 0;
    while (a < text.size()) {
      size_t b = text.find('\n', a); if (b == Str::npos) b = text.size();
      Str line = text.substr(a, b - a); a = b + 1; if (!line.empty() && line.back() == '\r') line.pop_back();
      if (line.compare(0, 2, "##") == 0) meta.push_back(line);
      else { have_chrom_line = line.compare(0, 6, "#CHROM") == 0; break; }
    }
    if (!have_chrom_line) throw IoError{ "VCF has no #CHROM header line: " + inputs[0] };
    if (!dict_path.empty()) {
      const std::vector<DictSeq> seqs = read_dict(dict_path);
      std::vector<Str> contig_lines;
      for (const DictSeq& s : seqs) { Str l = "##contig=<ID=" + s.name + ",length=" + std::to_string(s.length); if (!s.assembly.empty()) l += ",assembly=" + s.assembly; contig_lines.push_back(l + ">"); }
      std::vector<Str> kept; bool placed = false;
      for (const Str& l : meta) {
        if (l.compare(0, 9, "##contig=") == 0) { if (!placed) { kept.insert(kept.end(), contig_lines.begin(), contig_lines.end()); placed = true; } continue; }
        if (l.compare(0, 12, "##reference=") == 0) continue;
        kept.push_back(l);
      }
      if (!placed) kept.insert(kept.end(), contig_lines.begin(), contig_lines.end());
      kept.push_back("##reference=" + seqs[0].assembly);
      meta.swap(kept);
    }
    for (const Str& l : meta) { out += l; out += '\n'; }
    out += "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n";     // samples = IndexedSeq.empty (:62)
  }
  // ---- records: PrepareVcf.scala:68-88 ------------------------------------------------------------------------------------------------
  for (const Str& path : inputs) {
    const Str text = gunzip_if_needed(read_file(path));
    size_t a = 0;
    while (a < text.size()) {
      size_t b = text.find('\n', a); if (b == Str::npos) b = text.size();
      Str line = text.substr(a, b - a); a = b + 1; if (!line.empty() && line.back() == '\r') line.pop_back();
      if (line.empty() || line[0] == '#') continue;
      ++st.records_in;
      const std::vector<Str> f = split_on(line, '\t');
      if (f.size() < 8) throw IoError{ "malformed VCF record in " + path + ": " + line };
      if (f[6] != "PASS") continue;                                                                  // v.filters == Variant.PassingFilters (:73)
      std::vector<float> afs; bool has_af = false;
      for (const Str& kv : split_on(f[7], ';')) if (kv.compare(0, 3, "AF=") == 0) { has_af = true; for (const Str& x : split_on(kv.substr(3), ',')) afs.push_back(x == "." ? std::nanf("") : std::strtof(x.c_str(), nullptr)); }
      if (!has_af) throw IoError{ "key not found: AF (" + f[0] + ":" + f[1] + " in " + path + ")" };  // v[ArrayAttr[Float]]("AF") throws NoSuchElementException (:74)
      bool any = false; for (float x : afs) if ((double)x >= min_af) any = true;
      if (!any) continue;
      const std::vector<Str> alts = split_on(f[4], ',');
      bool simple = simple_allele(f[3]); for (const Str& x : alts) if (!simple_allele(x)) simple = false;
      if (!simple) continue;                                                                          // (:75)
      Str alt_out, af_out;
      for (size_t i = 0; i < alts.size() && i < afs.size(); ++i) {                                     // zip + filter (:77)
        if (!((double)afs[i] >= min_af)) continue;
        if (!alt_out.empty()) { alt_out += ','; af_out += ','; }
        alt_out += alts[i]; af_out += format_vcf_double((double)afs[i]);
      }
      if (alt_out.empty()) alt_out = ".";
      const Str chrom = (add_chr_prefix && chroms_to_fix().count(f[0])) ? "chr" + f[0] : f[0];         // fixChrom (:91)
      Str qual = f[5]; if (qual != ".") { char* e = nullptr; const double q = std::strtod(qual.c_str(), &e); if (e != qual.c_str() && !*e) qual = format_qual(q); }
      out += chrom; out += '\t'; out += f[1]; out += '\t'; out += f[2]; out += '\t'; out += f[3]; out += '\t'; out += alt_out; out += '\t'; out += qual;
      out += "\tPASS\tAF="; out += af_out; out += '\n';
      ++st.records_out;
    }
  }
  const bool gz = output.size() > 3 && output.compare(output.size() - 3, 3, ".gz") == 0;
  if (gz) { const Str z = bgzf_compress(out); write_file(output, z.data(), z.size()); } else write_file(output, out.data(), out.size());
  return st;
}

}}  // namespace cal::io
