// cal_vcf.h — PrepareVcf (PrepareVcf.scala:43-91): VCF clean-up for SearchReference -v.  Host code for the `calitas` command-line tool.
#pragma once
#include <string>
#include <vector>
#include "cal_io.h"

namespace cal { namespace io {

struct PrepareVcfStats { long long records_in = 0, records_out = 0; };
// inputs: one or more .vcf / .vcf.gz (disjoint, merged in the order given; the first one's header is used); output: .vcf, or .vcf.gz (BGZF);
// dict_path: optional sequence dictionary that overrides the contig lines; add_chr_prefix: "chr" for chromosomes 1-22, X, Y.
PrepareVcfStats prepare_vcf(const std::vector<std::string>& inputs, const std::string& output, double min_af, const std::string& dict_path, bool add_chr_prefix);
std::string bgzf_compress(const std::string& text);       // BGZF: 64-KB gzip members with the BC extra field + the empty end-of-file block

}}  // namespace cal::io
