// cal_io.cpp — see cal_io.h
#include "cal_io.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <thread>
#include <vector>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

namespace cal { namespace io {

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

std::string read_file(const std::string& path) {
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) throw IoError{ "Cannot read non-existent path: " + path };
  std::string out;
  if (std::fseek(f, 0, SEEK_END) == 0) {
    const long n = std::ftell(f);
    std::rewind(f);
    if (n > 0) { out.resize((size_t)n); const size_t got = std::fread(&out[0], 1, (size_t)n, f); out.resize(got); }
  } else {                                 // not seekable (pipe): read in chunks
    char buf[1 << 16]; size_t got;
    while ((got = std::fread(buf, 1, sizeof buf, f)) > 0) out.append(buf, got);
  }
  std::fclose(f);
  return out;
}

void write_file(const std::string& path, const char* data, size_t n) {
  if (path.empty() || path == "-" || path == "/dev/stdout") { std::fwrite(data, 1, n, stdout); std::fflush(stdout); return; }
  const int fd = ::open(path.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0666);
  if (fd < 0) throw IoError{ "Cannot write to path: " + path };
  struct stat st;
  if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode)) {                      // pipe, /dev/null, ...: sequential writes
    size_t off = 0; while (off < n) { const ssize_t put = ::write(fd, data + off, n - off); if (put <= 0) { ::close(fd); throw IoError{ "Short write to " + path }; } off += (size_t)put; }
    ::close(fd); return;
  }
  // a 100-guide hit table is ~18 GB: slices go out on several threads (pwrite), a small file on one
  const size_t SLICE = 64u << 20;
  const size_t n_slices = (n + SLICE - 1) / SLICE;
  const unsigned nt = (unsigned)std::max<size_t>(1, std::min<size_t>(std::min<size_t>(8, n_slices), std::max(1u, std::thread::hardware_concurrency())));
  std::atomic_size_t next(0); std::atomic_bool failed(false);
  auto work = [&]() {
    for (;;) {
      const size_t k = next.fetch_add(1); if (k >= n_slices || failed.load()) return;
      size_t off = k * SLICE; const size_t stop = std::min(n, off + SLICE);
      while (off < stop) { const ssize_t put = pwrite(fd, data + off, stop - off, (off_t)off); if (put <= 0) { failed.store(true); return; } off += (size_t)put; }
    } };
  std::vector<std::thread> th; for (unsigned t = 1; t < nt; ++t) th.emplace_back(work);
  work(); for (auto& t : th) t.join();
  if (::close(fd) != 0 || failed.load()) throw IoError{ "Short write to " + path };
}

std::string gunzip_if_needed(const std::string& raw) {
  if (raw.size() < 2 || (unsigned char)raw[0] != 0x1f || (unsigned char)raw[1] != 0x8b) return raw;
  std::string out; out.reserve(raw.size() * 4);
  z_stream zs; std::memset(&zs, 0, sizeof zs);
  if (inflateInit2(&zs, 16 + MAX_WBITS) != Z_OK) throw IoError{ "zlib: inflateInit2 failed" };
  zs.next_in = (Bytef*)raw.data(); zs.avail_in = (uInt)std::min<size_t>(raw.size(), 1u << 30);
  size_t consumed = 0; std::vector<char> buf(1 << 20);
  for (;;) {
    zs.next_out = (Bytef*)buf.data(); zs.avail_out = (uInt)buf.size();
    const uInt in_before = zs.avail_in;
    const int rc = inflate(&zs, Z_NO_FLUSH);
    consumed += in_before - zs.avail_in;
    out.append(buf.data(), buf.size() - zs.avail_out);
    if (rc == Z_STREAM_END) {                                   // end of one member: BGZF and `cat a.gz b.gz` continue with the next
      if (consumed >= raw.size()) break;
      if (inflateReset(&zs) != Z_OK) { inflateEnd(&zs); throw IoError{ "zlib: inflateReset failed" }; }
    } else if (rc != Z_OK && !(rc == Z_BUF_ERROR && zs.avail_out == 0)) { inflateEnd(&zs); throw IoError{ "corrupt gzip data" }; }
    if (zs.avail_in == 0 && consumed < raw.size()) { zs.next_in = (Bytef*)raw.data() + consumed; zs.avail_in = (uInt)std::min<size_t>(raw.size() - consumed, 1u << 30); }
    else if (zs.avail_in == 0 && rc != Z_STREAM_END && zs.avail_out != 0) { inflateEnd(&zs); throw IoError{ "truncated gzip data" }; }
  }
  inflateEnd(&zs);
  return out;
}

static bool exists(const std::string& path) { FILE* f = std::fopen(path.c_str(), "rb"); if (!f) return false; std::fclose(f); return true; }

static std::vector<std::string> split_tabs(const std::string& line) {
  std::vector<std::string> out; size_t a = 0;
  for (;;) { const size_t b = line.find('\t', a); if (b == std::string::npos) { out.push_back(line.substr(a)); return out; } out.push_back(line.substr(a, b - a)); a = b + 1; }
}

// One contig's sequence lines -> bases: copies line by line with memchr (the text is ~61/60 of the sequence, so this is a streaming copy).
static void strip_newlines(const char* p, const char* end, std::string& out) {
  out.clear(); out.reserve((size_t)(end - p));
  while (p < end) {
    const char* nl = (const char*)std::memchr(p, '\n', (size_t)(end - p));
    const char* stop = nl ? nl : end;
    const char* q = stop; if (q > p && q[-1] == '\r') --q;
    out.append(p, (size_t)(q - p));
    p = nl ? nl + 1 : end;
  }
}

// The FASTA text: a read-only mapping of a plain regular file (no 3-GB copy; the pages are touched by the stripping threads), else the
// file's contents read and, if gzip-compressed, inflated.
struct TextView {
  const char* data = nullptr; size_t size = 0; std::string owned; void* map = nullptr; size_t map_len = 0;
  ~TextView() { if (map) munmap(map, map_len); }
};
static void open_text(const std::string& path, TextView& v) {
  const int fd = ::open(path.c_str(), O_RDONLY);
  if (fd < 0) throw IoError{ "Cannot read non-existent path: " + path };
  struct stat st; unsigned char magic[2] = { 0, 0 };
  const bool regular = fstat(fd, &st) == 0 && S_ISREG(st.st_mode) && st.st_size > 0;
  if (regular && pread(fd, magic, 2, 0) == 2 && !(magic[0] == 0x1f && magic[1] == 0x8b)) {
    void* m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    if (m != MAP_FAILED) { madvise(m, (size_t)st.st_size, MADV_WILLNEED); v.map = m; v.map_len = (size_t)st.st_size; v.data = (const char*)m; v.size = v.map_len; ::close(fd); return; }
  }
  ::close(fd);
  v.owned = gunzip_if_needed(read_file(path)); v.data = v.owned.data(); v.size = v.owned.size();
}

Genome load_fasta(const std::string& path) {
  Genome g;
  double t0 = now_s();
  TextView text; open_text(path, text);
  g.read_s = now_s() - t0; t0 = now_s();
  struct Span { const char* b; const char* e; };
  std::vector<Span> spans;
  const char* p = text.data; const char* end = p + text.size;
  while (p < end) {
    if (*p != '>') { const char* nl = (const char*)std::memchr(p, '\n', (size_t)(end - p)); p = nl ? nl + 1 : end; continue; }   // text before the first header
    {
      const char* nl = (const char*)std::memchr(p, '\n', (size_t)(end - p)); const char* he = nl ? nl : end;
      const char* q = p + 1; while (q < he && *q != ' ' && *q != '\t' && *q != '\r') ++q;
      g.names.emplace_back(p + 1, (size_t)(q - p - 1));
      const char* sb = nl ? nl + 1 : end;
      const char* se = sb;
      for (;;) {                                                        // next header = '>' at the start of a line
        const char* gt = (const char*)std::memchr(se, '>', (size_t)(end - se));
        if (!gt) { se = end; break; }
        if (gt == sb || gt[-1] == '\n') { se = gt; break; }
        se = gt + 1;
      }
      spans.push_back(Span{ sb, se });
      p = se;
    }
  }
  if (g.names.empty()) throw IoError{ "No sequences in FASTA: " + path };
  g.seqs.resize(g.names.size());
  {  // contigs are independent: strip line ends on a few threads
    const unsigned nt = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    std::vector<size_t> order(spans.size()); for (size_t i = 0; i < order.size(); ++i) order[i] = i;                      // longest contig first: the tail of the schedule is short
    std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return spans[a].e - spans[a].b > spans[b].e - spans[b].b; });
    std::vector<std::thread> th; std::atomic_size_t* next = new std::atomic_size_t(0);
    auto work = [&]() { for (;;) { const size_t k = next->fetch_add(1); if (k >= spans.size()) return; const size_t i = order[k]; strip_newlines(spans[i].b, spans[i].e, g.seqs[i]); } };
    for (unsigned t = 1; t < nt; ++t) th.emplace_back(work);
    work(); for (auto& t : th) t.join(); delete next;
  }
  // .fai: name, length, offset, line bases, line width — only the first two are checked (the engine holds the whole FASTA in memory)
  if (exists(path + ".fai")) {
    g.has_fai = true;
    const std::string fai = read_file(path + ".fai"); size_t a = 0, row = 0;
    while (a < fai.size()) {
      size_t b = fai.find('\n', a); if (b == std::string::npos) b = fai.size();
      const std::string line = fai.substr(a, b - a); a = b + 1;
      if (line.empty()) continue;
      const std::vector<std::string> f = split_tabs(line);
      if (row >= g.names.size() || f.size() < 2 || f[0] != g.names[row] || std::atoll(f[1].c_str()) != (long long)g.seqs[row].size())
        throw IoError{ "FASTA index does not match the FASTA: " + path + ".fai" };
      ++row;
    }
    if (row != g.names.size()) throw IoError{ "FASTA index does not match the FASTA: " + path + ".fai" };
  }
  // sequence dictionary (SAMSequenceDictionaryExtractor: <fasta>.dict or the FASTA's extension replaced by .dict)
  std::vector<std::string> dicts = { path + ".dict" };
  { const size_t dot = path.rfind('.'); const size_t slash = path.rfind('/'); if (dot != std::string::npos && (slash == std::string::npos || dot > slash)) dicts.push_back(path.substr(0, dot) + ".dict"); }
  for (const std::string& d : dicts) {
    if (!exists(d)) continue;
    g.has_dict = true;
    const std::string txt = read_file(d); size_t a = 0;
    while (a < txt.size() && g.assembly.empty()) {
      size_t b = txt.find('\n', a); if (b == std::string::npos) b = txt.size();
      const std::string line = txt.substr(a, b - a); a = b + 1;
      if (line.compare(0, 3, "@SQ") != 0) continue;
      for (const std::string& f : split_tabs(line)) if (f.compare(0, 3, "AS:") == 0) { g.assembly = f.substr(3); break; }
    }
    break;
  }
  g.parse_s = now_s() - t0;
  return g;
}

std::vector<A2RRow> load_a2r_tasks(const std::string& path) {
  const std::string text = gunzip_if_needed(read_file(path));
  std::vector<A2RRow> rows; std::vector<std::string> hdr; size_t a = 0; int ci = -1, cq = -1, cc = -1, cp = -1;
  while (a < text.size()) {
    size_t b = text.find('\n', a); if (b == std::string::npos) b = text.size();
    std::string line = text.substr(a, b - a); a = b + 1;
    if (!line.empty() && line.back() == '\r') line.pop_back();
    if (line.empty()) continue;
    const std::vector<std::string> f = split_tabs(line);
    if (hdr.empty()) {
      hdr = f;
      for (size_t i = 0; i < f.size(); ++i) { if (f[i] == "id") ci = (int)i; else if (f[i] == "query") cq = (int)i; else if (f[i] == "chrom") cc = (int)i; else if (f[i] == "position") cp = (int)i; }
      if (cq < 0 || cc < 0 || cp < 0) throw IoError{ "Input must have a header with the columns query, chrom, position (and optionally id): " + path };
      continue;
    }
    if (f.size() != hdr.size()) throw IoError{ "Line has " + std::to_string(f.size()) + " fields, header has " + std::to_string(hdr.size()) + ": " + line };
    A2RRow r; r.query = f[(size_t)cq]; r.chrom = f[(size_t)cc]; r.id = ci >= 0 ? f[(size_t)ci] : r.query;
    char* endp = nullptr; const long v = std::strtol(f[(size_t)cp].c_str(), &endp, 10);
    if (endp == f[(size_t)cp].c_str() || *endp != 0) throw IoError{ "position is not an integer: " + f[(size_t)cp] };
    r.position = (int32_t)v; rows.push_back(r);
  }
  return rows;
}

// ---- MD5 (RFC 1321) -------------------------------------------------------------------------------------------------------------
namespace {
inline uint32_t rol(uint32_t x, int c) { return (x << c) | (x >> (32 - c)); }
void md5_block(uint32_t st[4], const uint8_t* p) {
  static const uint32_t K[64] = {
    0xd76aa478, 0xe8c7b756, 0x242070db, 0xc1bdceee, 0xf57c0faf, 0x4787c62a, 0xa8304613, 0xfd469501, 0x698098d8, 0x8b44f7af, 0xffff5bb1, 0x895cd7be, 0x6b901122, 0xfd987193, 0xa679438e, 0x49b40821,
    0xf61e2562, 0xc040b340, 0x265e5a51, 0xe9b6c7aa, 0xd62f105d, 0x02441453, 0xd8a1e681, 0xe7d3fbc8, 0x21e1cde6, 0xc33707d6, 0xf4d50d87, 0x455a14ed, 0xa9e3e905, 0xfcefa3f8, 0x676f02d9, 0x8d2a4c8a,
    0xfffa3942, 0x8771f681, 0x6d9d6122, 0xfde5380c, 0xa4beea44, 0x4bdecfa9, 0xf6bb4b60, 0xbebfbc70, 0x289b7ec6, 0xeaa127fa, 0xd4ef3085, 0x04881d05, 0xd9d4d039, 0xe6db99e5, 0x1fa27cf8, 0xc4ac5665,
    0xf4292244, 0x432aff97, 0xab9423a7, 0xfc93a039, 0x655b59c3, 0x8f0ccc92, 0xffeff47d, 0x85845dd1, 0x6fa87e4f, 0xfe2ce6e0, 0xa3014314, 0x4e0811a1, 0xf7537e82, 0xbd3af235, 0x2ad7d2bb, 0xeb86d391 };
  static const int S[64] = { 7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 5, 9, 14, 20, 5, 9, 14, 20, 5, 9, 14, 20, 5, 9, 14, 20,
                             4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21 };
  uint32_t m[16]; for (int i = 0; i < 16; ++i) m[i] = (uint32_t)p[4 * i] | ((uint32_t)p[4 * i + 1] << 8) | ((uint32_t)p[4 * i + 2] << 16) | ((uint32_t)p[4 * i + 3] << 24);
  uint32_t a = st[0], b = st[1], c = st[2], d = st[3];
  for (int i = 0; i < 64; ++i) {
    uint32_t f; int g;
    if (i < 16) { f = (b & c) | (~b & d); g = i; } else if (i < 32) { f = (d & b) | (~d & c); g = (5 * i + 1) & 15; }
    else if (i < 48) { f = b ^ c ^ d; g = (3 * i + 5) & 15; } else { f = c ^ (b | ~d); g = (7 * i) & 15; }
    const uint32_t t = d; d = c; c = b; b = b + rol(a + f + K[i] + m[g], S[i]); a = t;
  }
  st[0] += a; st[1] += b; st[2] += c; st[3] += d;
}
}  // namespace

std::string md5_hex(const std::string& data) {
  uint32_t st[4] = { 0x67452301, 0xefcdab89, 0x98badcfe, 0x10325476 };
  const uint8_t* p = (const uint8_t*)data.data(); size_t n = data.size(), i = 0;
  for (; i + 64 <= n; i += 64) md5_block(st, p + i);
  uint8_t tail[128]; size_t r = n - i; std::memcpy(tail, p + i, r); tail[r++] = 0x80;
  const size_t pad = r <= 56 ? 56 : 120; std::memset(tail + r, 0, pad - r);
  const uint64_t bits = (uint64_t)n * 8; for (int k = 0; k < 8; ++k) tail[pad + k] = (uint8_t)(bits >> (8 * k));
  md5_block(st, tail); if (pad == 120) md5_block(st, tail + 64);
  char hex[33]; for (int k = 0; k < 16; ++k) std::snprintf(hex + 2 * k, 3, "%02x", (unsigned)((st[k >> 2] >> (8 * (k & 3))) & 0xFF));
  return std::string(hex, 32);
}

std::string file_name_of(const std::string& path) { const size_t s = path.rfind('/'); return s == std::string::npos ? path : path.substr(s + 1); }

std::string utc_time_stamp() {
  std::time_t t = std::time(nullptr); std::tm tm; gmtime_r(&t, &tm);
  char buf[64]; std::strftime(buf, sizeof buf, "%a %b %d %H:%M:%S UTC %Y", &tm);
  return buf;
}

}}  // namespace cal::io
