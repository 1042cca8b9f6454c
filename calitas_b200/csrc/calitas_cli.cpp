// calitas_cli.cpp — `calitas SearchReference ...` / `AlignToReference ...` / `PairwiseAlignSequences ...` on the B200 engine, and `PrepareVcf`.
//
// Keeps the reference's command-line surface (SearchReference.scala:452-470, AlignToReference.scala:35-50: same flags, defaults and
// validation messages) and its 34-column hit table; underneath, the whole FASTA is read into memory, packed into HBM and searched
// by libcalitas_b200.so.  Extensions, all optional: --devices (comma-separated GPU ids; the genome is sharded by contig range),
// --guides-file (a batch of guides in one run: `id<TAB>guide[<TAB>auxPam,auxPam]` per line; the engine scans 16 guides per pass),
// --time-stamp / --aligner-version (fix the two run-dependent columns, for reproducible comparisons), --stats (timings to stderr).
// -t/--threads is accepted and ignored (the reference's CPU thread count).
#include <cctype>
#include <fcntl.h>
#include <unistd.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/calitas_b200_tools.h"
#include "cal_io.h"
#include "cal_vcf.h"

using namespace cal::io;
typedef std::string Str;

namespace {

struct UsageError { Str msg; };
double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

struct FlagDef { char short_name; const char* long_name; bool multi; bool boolean = false; };

// sopt-style parsing: -x v, --long v, --long=v, repeated flags or several values after a multi-valued flag.
std::map<Str, std::vector<Str>> parse_flags(int argc, char** argv, int first, const std::vector<FlagDef>& defs) {
  std::map<Str, std::vector<Str>> out;
  const FlagDef* cur = nullptr;
  for (int i = first; i < argc; ++i) {
    Str a = argv[i];
    const bool is_flag = a.size() >= 2 && a[0] == '-' && !(a[1] >= '0' && a[1] <= '9');
    if (is_flag) {
      Str name, value; bool has_value = false;
      if (a[1] == '-') { const size_t eq = a.find('='); name = a.substr(2, eq == Str::npos ? Str::npos : eq - 2); if (eq != Str::npos) { value = a.substr(eq + 1); has_value = true; } }
      else { name = a.substr(1, 1); if (a.size() > 2) { value = a.substr(a[2] == '=' ? 3 : 2); has_value = true; } }
      cur = nullptr;
      for (const FlagDef& d : defs) if (name == d.long_name || (name.size() == 1 && name[0] == d.short_name)) cur = &d;
      if (!cur) throw UsageError{ "No option found with name '" + name + "'" };
      std::vector<Str>& v = out[cur->long_name];
      if (has_value) v.push_back(value);
      else if (cur->boolean) {             // sopt booleans: bare flag = true, or an explicit true/false value
        const Str nx = i + 1 < argc ? argv[i + 1] : "";
        if (nx == "true" || nx == "false" || nx == "True" || nx == "False" || nx == "T" || nx == "F" || nx == "yes" || nx == "no") { v.push_back(nx); ++i; } else v.push_back("true");
      }
      else if (i + 1 < argc && !(std::strlen(argv[i + 1]) >= 2 && argv[i + 1][0] == '-' && !(argv[i + 1][1] >= '0' && argv[i + 1][1] <= '9'))) v.push_back(argv[++i]);
      else if (!cur->multi) throw UsageError{ Str("Option '") + cur->long_name + "' requires a value" };
    } else {
      if (!cur || !cur->multi) throw UsageError{ "Unexpected argument: " + a };
      out[cur->long_name].push_back(a);
    }
  }
  return out;
}
struct Flags {
  std::map<Str, std::vector<Str>> m;
  bool has(const char* k) const { auto it = m.find(k); return it != m.end() && !it->second.empty(); }
  Str str(const char* k, const Str& dflt = Str()) const { return has(k) ? m.at(k).back() : dflt; }
  Str required(const char* k) const { if (!has(k)) throw UsageError{ Str("Argument '") + k + "' is required" }; return m.at(k).back(); }
  int integer(const char* k, int dflt) const {
    if (!has(k)) return dflt;
    const Str& s = m.at(k).back(); char* e = nullptr; const long v = std::strtol(s.c_str(), &e, 10);
    if (e == s.c_str() || *e) throw UsageError{ Str("Value for '") + k + "' is not an integer: " + s };
    return (int)v;
  }
  std::vector<Str> list(const char* k) const { return has(k) ? m.at(k) : std::vector<Str>(); }
};

void ck(int rc) { if (rc != CALITAS_OK) throw UsageError{ calitas_last_error() }; }

struct Session {       // engines (one per device), the genome on the host and its shards on the devices
  Genome genome; std::vector<const char*> names; std::vector<int64_t> lengths; std::vector<const uint8_t*> bases; calitas_genome_view view;
  std::vector<calitas_engine*> engines; std::vector<calitas_reference*> refs;
  double load_s = 0, pack_s = 0, init_s = 0;
  ~Session() { for (size_t i = 0; i < engines.size(); ++i) { if (i < refs.size()) calitas_reference_free(engines[i], refs[i]); calitas_engine_destroy(engines[i]); } }
};

std::vector<int> parse_devices(const Str& s) {
  std::vector<int> out; size_t a = 0;
  while (a <= s.size()) { size_t b = s.find(',', a); if (b == Str::npos) b = s.size(); if (b > a) out.push_back(std::atoi(s.substr(a, b - a).c_str())); a = b + 1; }
  if (out.empty()) out.push_back(0);
  return out;
}

void open_session(Session& S, const Flags& f, bool need_fai, int halo) {
  const Str ref = f.required("ref");
  double t0 = now_s();
  S.genome = load_fasta(ref);
  if (!S.genome.has_dict) throw UsageError{ "Reference genome must have a sequence dictionary: " + ref };           // SearchReference.scala:478-484, AlignToReference.scala:61-62
  if (need_fai && !S.genome.has_fai) throw UsageError{ "Reference genome must have a fasta index: " + ref };        // AlignToReference.scala:60
  S.load_s = now_s() - t0;
  const int n = (int)S.genome.names.size();
  for (int c = 0; c < n; ++c) { S.names.push_back(S.genome.names[(size_t)c].c_str()); S.lengths.push_back((int64_t)S.genome.seqs[(size_t)c].size()); S.bases.push_back((const uint8_t*)S.genome.seqs[(size_t)c].data()); }
  S.view = calitas_genome_view{ n, S.names.data(), S.lengths.data(), S.bases.data(), S.genome.assembly.empty() ? nullptr : S.genome.assembly.c_str() };
  calitas_costs costs{ f.integer("guide-mismatch-net-cost", -120), f.integer("genome-gap-net-cost", -122), f.integer("guide-gap-net-cost", -121), f.integer("pam-mismatch-net-cost", -260) };
  const std::vector<int> devices = parse_devices(f.str("devices", "0"));
  t0 = now_s();
  for (size_t s = 0; s < devices.size(); ++s) {
    const double ti = now_s();
    calitas_engine* e = nullptr; ck(calitas_engine_create(devices[s], &costs, &e)); S.engines.push_back(e);
    S.init_s += now_s() - ti;
    calitas_reference* r = nullptr;
    if (devices.size() == 1) ck(calitas_reference_load(e, n, S.names.data(), S.lengths.data(), S.bases.data(), nullptr, nullptr, nullptr, nullptr, 0, &r));
    else {
      std::vector<int64_t> ob((size_t)n), oe((size_t)n), hb((size_t)n), he((size_t)n); std::vector<const uint8_t*> ptr((size_t)n);
      ck(calitas_shard_plan(n, S.lengths.data(), (int)s, (int)devices.size(), halo, ob.data(), oe.data(), hb.data(), he.data()));
      for (int c = 0; c < n; ++c) ptr[(size_t)c] = S.bases[(size_t)c] + hb[(size_t)c];
      ck(calitas_reference_load(e, n, S.names.data(), S.lengths.data(), ptr.data(), hb.data(), he.data(), ob.data(), oe.data(), 0, &r));
    }
    S.refs.push_back(r);
  }
  S.pack_s = now_s() - t0 - S.init_s;
}

const std::vector<FlagDef> kCommon = {
  { 'r', "ref", false }, { 'o', "output", false }, { 't', "threads", false }, { 'w', "window-size", false }, { 'd', "max-guide-diffs", false },
  { 'p', "max-pam-mismatches", false }, { 'g', "max-gaps-between-guide-and-pam", false }, { 'D', "max-total-diffs", false }, { 'O', "max-overlap", false },
  { 'm', "guide-mismatch-net-cost", false }, { 'M', "pam-mismatch-net-cost", false }, { 'b', "genome-gap-net-cost", false }, { 'B', "guide-gap-net-cost", false },
  { 0, "devices", false }, { 0, "time-stamp", false }, { 0, "aligner-version", false }, { 0, "stats", false, true } };

int search_reference(int argc, char** argv) {
  std::vector<FlagDef> defs = kCommon;
  for (const FlagDef& d : std::vector<FlagDef>{ { 'i', "guide", false }, { 'I', "guide-id", false }, { 'x', "auxiliary-pams", true }, { 'v', "variants", false },
                                                  { 'V', "max-variants", false }, { 'c', "chrom", false }, { 0, "guides-file", false } }) defs.push_back(d);
  Flags f{ parse_flags(argc, argv, 2, defs) };
  // guides: the reference's single -i/-I/-x, or a batch file
  std::vector<Str> ids, seqs; std::vector<std::vector<Str>> aux;
  if (f.has("guides-file")) {
    const Str text = read_file(f.str("guides-file")); size_t a = 0;
    while (a < text.size()) {
      size_t b = text.find('\n', a); if (b == Str::npos) b = text.size();
      Str line = text.substr(a, b - a); a = b + 1; if (!line.empty() && line.back() == '\r') line.pop_back();
      if (line.empty() || line[0] == '#') continue;
      std::vector<Str> c; size_t x = 0; for (;;) { size_t y = line.find('\t', x); if (y == Str::npos) { c.push_back(line.substr(x)); break; } c.push_back(line.substr(x, y - x)); x = y + 1; }
      if (c.size() < 2) throw UsageError{ "guides file lines are id<TAB>guide[<TAB>aux,pams]: " + line };
      ids.push_back(c[0]); seqs.push_back(c[1]); aux.emplace_back();
      if (c.size() > 2) { size_t u = 0; while (u <= c[2].size()) { size_t v = c[2].find(',', u); if (v == Str::npos) v = c[2].size(); if (v > u) aux.back().push_back(c[2].substr(u, v - u)); u = v + 1; } }
    }
    if (seqs.empty()) throw UsageError{ "no guides in " + f.str("guides-file") };
  } else { seqs.push_back(f.required("guide")); ids.push_back(f.required("guide-id")); aux.push_back(f.list("auxiliary-pams")); }
  calitas_search_options opt; std::memset(&opt, 0, sizeof opt);
  opt.max_variants = f.integer("max-variants", 16); opt.window_size = f.integer("window-size", 1000);
  opt.limits = calitas_limits{ f.integer("max-guide-diffs", 5), f.integer("max-pam-mismatches", 1), f.integer("max-gaps-between-guide-and-pam", 3), f.integer("max-total-diffs", -1), f.integer("max-overlap", 10) };
  const Str chrom = f.str("chrom"); opt.chrom = chrom.empty() ? nullptr : chrom.c_str();
  Str vcf_text, vcf_id; if (f.has("variants")) {                                  // .vcf or .vcf.gz / bgzip; the id carries the MD5 of the file as stored (ReferenceHit.scala:175-183)
    const Str raw = read_file(f.str("variants")); vcf_id = file_name_of(f.str("variants")) + ":" + md5_hex(raw); vcf_text = gunzip_if_needed(raw);
    opt.vcf_text = vcf_text.c_str(); opt.vcf_id = vcf_id.c_str();
  }
  const Str stamp = f.str("time-stamp", utc_time_stamp()), version = f.str("aligner-version", "calitas-b200-0.1");
  opt.time_stamp = stamp.c_str(); opt.aligner_version = version.c_str();
  Session S; open_session(S, f, false, 4 * opt.window_size);
  std::vector<calitas_guide> guides; std::vector<std::vector<const char*>> aux_ptr(seqs.size()); std::vector<const char*> id_ptr;
  for (size_t i = 0; i < seqs.size(); ++i) { for (auto& a : aux[i]) aux_ptr[i].push_back(a.c_str()); guides.push_back(calitas_guide{ seqs[i].c_str(), aux_ptr[i].empty() ? nullptr : aux_ptr[i].data(), (int32_t)aux_ptr[i].size() }); id_ptr.push_back(ids[i].c_str()); }
  int64_t n_hits = 0, n_bytes = 0;
  // the table leaves through a descriptor while it is rendered (stdout for "-")
  const Str out_path = f.str("output", "-");
  const bool to_stdout = out_path.empty() || out_path == "-" || out_path == "/dev/stdout";
  int fd = 1;
  if (to_stdout) std::fflush(stdout);
  else { fd = ::open(out_path.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0666); if (fd < 0) throw IoError{ "Cannot write to path: " + out_path }; }
  double t0 = now_s();
  const int rc = calitas_tool_search_reference_batch_fd((int32_t)S.engines.size(), S.engines.data(), (const calitas_reference* const*)S.refs.data(), &S.view, (int32_t)guides.size(), guides.data(),
                                                        id_ptr.data(), &opt, fd, &n_hits, &n_bytes);
  const Str err = rc ? calitas_last_error() : "";
  if (!to_stdout && ::close(fd) != 0 && !rc) throw IoError{ "Short write to " + out_path };
  if (rc) throw UsageError{ err };
  const double search_s = now_s() - t0;
  if (f.has("stats"))
    std::fprintf(stderr, "calitas-b200 SearchReference: %zu guide(s), %lld hits, %lld bytes; fasta read %.3f s + parse %.3f s, CUDA init %.3f s, upload+pack %.3f s on %zu GPU(s), search+render+write %.3f s\n",
                 guides.size(), (long long)n_hits, (long long)n_bytes, S.genome.read_s, S.genome.parse_s, S.init_s, S.pack_s, S.engines.size(), search_s);
  return 0;
}

int align_to_reference(int argc, char** argv) {
  std::vector<FlagDef> defs = kCommon; defs.push_back(FlagDef{ 'i', "input", false });
  Flags f{ parse_flags(argc, argv, 2, defs) };
  const std::vector<A2RRow> rows = load_a2r_tasks(f.required("input"));
  calitas_a2r_options opt; std::memset(&opt, 0, sizeof opt);
  opt.window_size = f.integer("window-size", -1); opt.max_guide_diffs = f.integer("max-guide-diffs", -1); opt.max_pam_mismatches = f.integer("max-pam-mismatches", -1);
  opt.max_gaps_between_guide_and_pam = f.integer("max-gaps-between-guide-and-pam", 3); opt.max_total_diffs = f.integer("max-total-diffs", -1); opt.max_overlap = f.integer("max-overlap", -1);
  const Str stamp = f.str("time-stamp", utc_time_stamp()), version = f.str("aligner-version", "calitas-b200-0.1");
  opt.time_stamp = stamp.c_str(); opt.aligner_version = version.c_str();
  Flags one = f; one.m["devices"] = { Str(std::to_string(parse_devices(f.str("devices", "0"))[0])) };     // AlignToReference batches are small: one GPU
  Session S; open_session(S, one, true, 0);
  std::vector<calitas_a2r_task> tasks; for (const A2RRow& r : rows) tasks.push_back(calitas_a2r_task{ r.id.c_str(), r.query.c_str(), r.chrom.c_str(), r.position });
  char* tsv = nullptr; int64_t n_hits = 0;
  double t0 = now_s();
  ck(calitas_tool_align_to_reference(S.engines[0], S.refs[0], &S.view, (int64_t)tasks.size(), tasks.data(), &opt, &tsv, &n_hits));
  const double run_s = now_s() - t0;
  write_file(f.str("output", "-"), tsv, std::strlen(tsv));
  calitas_free_text(tsv);
  if (f.has("stats")) std::fprintf(stderr, "calitas-b200 AlignToReference: %zu task(s), %lld rows; fasta %.3f s, upload+pack %.3f s, align+render %.3f s\n", tasks.size(), (long long)n_hits, S.load_s, S.pack_s, run_s);
  return 0;
}

int pairwise_align(int argc, char** argv) {           // PairwiseAlignSequences.scala:24-36
  std::vector<FlagDef> defs = { { 'i', "input", false }, { 'o', "output", false }, { 't', "threads", false }, { 'g', "max-gaps-between-guide-and-pam", false }, { 'O', "max-overlap", false },
                                { 'm', "guide-mismatch-net-cost", false }, { 'M', "pam-mismatch-net-cost", false }, { 'b', "genome-gap-net-cost", false }, { 'B', "guide-gap-net-cost", false }, { 0, "devices", false } };
  Flags f{ parse_flags(argc, argv, 2, defs) };
  const Str text = read_file(f.required("input"));
  std::vector<Str> qs, ts; size_t a = 0;
  while (a < text.size()) {
    size_t b = text.find('\n', a); if (b == Str::npos) b = text.size();
    const Str line = text.substr(a, b - a); a = b + 1;
    std::vector<Str> fields; size_t x = 0;
    while (x < line.size()) { while (x < line.size() && std::isspace((unsigned char)line[x])) ++x; size_t y = x; while (y < line.size() && !std::isspace((unsigned char)line[y])) ++y; if (y > x) fields.push_back(line.substr(x, y - x)); x = y; }
    if (fields.empty()) continue;
    if (fields.size() != 2) { Str joined; for (auto& s : fields) { if (!joined.empty()) joined += ' '; joined += s; } throw UsageError{ "requirement failed: Line found with " + std::to_string(fields.size()) + " fields: " + joined }; }
    qs.push_back(fields[0]); ts.push_back(fields[1]);
  }
  calitas_costs costs{ f.integer("guide-mismatch-net-cost", -120), f.integer("genome-gap-net-cost", -122), f.integer("guide-gap-net-cost", -121), f.integer("pam-mismatch-net-cost", -260) };
  calitas_engine* e = nullptr; ck(calitas_engine_create(parse_devices(f.str("devices", "0"))[0], &costs, &e));
  std::vector<const char*> qp, tp; for (size_t i = 0; i < qs.size(); ++i) { qp.push_back(qs[i].c_str()); tp.push_back(ts[i].c_str()); }
  char* tsv = nullptr; const int rc = calitas_tool_pairwise_align(e, (int64_t)qs.size(), qp.data(), tp.data(), &tsv);
  const Str err = rc ? calitas_last_error() : "";
  if (!rc) { write_file(f.str("output", "-"), tsv, std::strlen(tsv)); calitas_free_text(tsv); }
  calitas_engine_destroy(e);
  if (rc) throw UsageError{ err };
  return 0;
}

int prepare_vcf_tool(int argc, char** argv) {       // PrepareVcf.scala:32-38
  std::vector<FlagDef> defs = { { 'i', "input", true }, { 'o', "output", false }, { 'f', "min-af", false }, { 'd', "dict", false }, { 'c', "add-chr-prefix", false, true }, { 0, "stats", false, true } };
  Flags f{ parse_flags(argc, argv, 2, defs) };
  if (!f.has("input")) throw UsageError{ "Argument 'input' is required" };
  double min_af = 0.01;
  if (f.has("min-af")) { const Str s = f.str("min-af"); char* e = nullptr; min_af = std::strtod(s.c_str(), &e); if (e == s.c_str() || *e) throw UsageError{ "Value for 'min-af' is not a number: " + s }; }
  const Str c = f.str("add-chr-prefix", "true");
  const bool add_chr = !(c == "false" || c == "False" || c == "F" || c == "no");
  const PrepareVcfStats st = prepare_vcf(f.list("input"), f.required("output"), min_af, f.str("dict"), add_chr);
  if (f.has("stats")) std::fprintf(stderr, "calitas-b200 PrepareVcf: %lld of %lld records kept\n", st.records_out, st.records_in);
  return 0;
}

void usage() {
  std::fprintf(stderr,
    "calitas (B200 engine)\nUSAGE: calitas SearchReference -i GUIDEpam -I ID -r ref.fa [-x pam ...] [-v variants.vcf] [-V 16] [-o out.tsv] [-w 1000] [-d 5] [-p 1] [-g 3] [-D n] [-O 10]\n"
    "                               [-m -120] [-M -260] [-b -122] [-B -121] [-c chrom] [--devices 0,1,...] [--guides-file file]\n"
    "       calitas AlignToReference -i tasks.tsv -r ref.fa [-o out.tsv] [-w n] [-d n -p n -O n] [-g 3] [-D n] [-m -M -b -B]\n"
    "       calitas PairwiseAlignSequences -i pairs.txt [-o out.tsv] [-m -M -b -B]\n"
    "       calitas PrepareVcf -i in.vcf[.gz] [more.vcf ...] -o out.vcf[.gz] [-f 0.01] [-d ref.dict] [-c true|false]\n");
}

}  // namespace

int main(int argc, char** argv) {
  if (argc < 2) { usage(); return 1; }
  try {
    const Str tool = argv[1];
    if (tool == "SearchReference") return search_reference(argc, argv);
    if (tool == "AlignToReference") return align_to_reference(argc, argv);
    if (tool == "PairwiseAlignSequences") return pairwise_align(argc, argv);
    if (tool == "PrepareVcf") return prepare_vcf_tool(argc, argv);
    usage(); std::fprintf(stderr, "Unknown tool: %s\n", argv[1]); return 1;
  } catch (const UsageError& e) { std::fprintf(stderr, "calitas: %s\n", e.msg.c_str()); return 2; }
  catch (const IoError& e) { std::fprintf(stderr, "calitas: %s\n", e.msg.c_str()); return 2; }
  catch (const std::exception& e) { std::fprintf(stderr, "calitas: %s\n", e.what()); return 3; }
}
