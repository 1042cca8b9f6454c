// cal_io.h — file formats either side of the hot path: FASTA + .fai + .dict in, AlignToReference task table in,
// plain-text VCF in, hit table out.  Host code for the `calitas` command-line tool; nothing here computes alignments.
#pragma once
#include <stdint.h>
#include <string>
#include <vector>

namespace cal { namespace io {

struct IoError { std::string msg; };

std::string read_file(const std::string& path);                       // whole file; throws IoError
void write_file(const std::string& path, const char* data, size_t n); // "-" or "" = stdout
// gzip / bgzip (BGZF is a series of gzip members) -> text; anything else is returned unchanged.  fgbio's VcfSource reads .vcf and .vcf.gz alike.
std::string gunzip_if_needed(const std::string& raw);

// A reference genome as SearchReference / AlignToReference see it (htsjdk ReferenceSequenceFile + SAMSequenceDictionary):
// contig name = header up to the first white space; bases exactly as in the file (case kept, line ends removed).
struct Genome {
  std::vector<std::string> names, seqs;
  std::string assembly;            // first AS tag of the .dict (ReferenceHit.scala:207), empty if none
  bool has_dict = false, has_fai = false;
  double read_s = 0, parse_s = 0;  // timing of the load, seconds
};
// Reads <path>, <path>.fai (checked against the sequences when present) and the sequence dictionary (<path>.dict or <path minus extension>.dict).
Genome load_fasta(const std::string& path);

// AlignToReference input (AlignToReference.scala:97-102): tab-delimited with a header naming `query`, `chrom`, `position` and optionally `id`.
struct A2RRow { std::string id, query, chrom; int32_t position; };
std::vector<A2RRow> load_a2r_tasks(const std::string& path);

std::string md5_hex(const std::string& data);                         // ReferenceHit.scala:175-183
std::string file_name_of(const std::string& path);
std::string utc_time_stamp();                                         // "EEE MMM dd HH:mm:ss z yyyy" in UTC (ReferenceHit.scala:169-173)

}}  // namespace cal::io
