// cal_host.h — host-side mirror of the reference's operator interface around the device path:
// Guide parsing (SequentialGuideAligner.scala:32-122), limits/threshold derivation (:239-249), and the rendering of
// device hit records into GuideAlignment / ReferenceHit text (GuideAlignment.scala, ReferenceHit.scala:210-266).
#pragma once
#include <stdexcept>
#include <string>
#include <vector>
#include "cal_core.cuh"

namespace cal {

struct InvalidArgument : std::runtime_error { using std::runtime_error::runtime_error; };   // reference: require(...) -> IllegalArgumentException
struct LimitExceeded : std::runtime_error { using std::runtime_error::runtime_error; };

struct GuideDef {
  std::string raw;                  // as given
  std::string protospacer;          // upper case
  std::vector<std::string> pams;    // lower case, primary first
  bool five_prime = false, three_prime = false;
  int protospacer_length() const { return (int)protospacer.size(); }
  int pam_length() const { size_t m = 0; for (auto& p : pams) if (p.size() > m) m = p.size(); return (int)m; }
  int length() const { return protospacer_length() + pam_length(); }
  std::string with_pam(int pam_idx) const;   // guide + PAM text in guide orientation
};

GuideDef parse_guide(const calitas_guide& g);
GuideDef parse_guide(const std::string& sequence, const std::vector<std::string>& aux);

// best != 0: alignBest / alignToRefBest limits (SequentialGuideAligner.scala:336-343, 407-417)
GuideSpec make_guide_spec(const GuideDef& g, const Scores& sc, const calitas_limits& lim, bool best);

char complement_base(char b);
std::string revcomp(const std::string& s);

// One rendered alignment (GuideAlignment + its derived counters)
struct Rendered {
  std::string guide, padded_guide, padded_alignment, padded_target, cigar, unpadded_target_without_pam;
  int mismatches = 0, gap_bases = 0, edits = 0, guide_mm = 0, guide_gaps = 0, guide_mm_plus_gaps = 0, pam_mm = 0, pam_gaps = 0, pam_mm_plus_gaps = 0;
};
// The same without allocations (fixed arrays, explicit lengths): the ReferenceHit row writer renders tens of millions of hits through it.
struct RenderedFix {
  char padded_guide[CALITAS_MAX_OPS], padded_alignment[CALITAS_MAX_OPS], padded_target[CALITAS_MAX_OPS]; int n;
  char cigar[CALITAS_MAX_OPS * 4]; int cigar_len;
  char unpadded[CALITAS_MAX_OPS]; int unpadded_len;
  int mismatches, gap_bases, edits, guide_mm, guide_gaps, guide_mm_plus_gaps, pam_mm, pam_gaps, pam_mm_plus_gaps;
};
void render_hit_fix(const HitX& h, const char* guide, int guide_len, const char* target_fwd, int target_len, bool upper_case, RenderedFix& r);
// `target` = bases [start_offset, end_offset) of the hit in forward orientation (as stored); rendered in guide orientation.
Rendered render_hit(const HitX& h, const GuideDef& g, const std::string& target_fwd, bool upper_case);
std::string alignment_header();
std::string alignment_row(const HitX& h, const Rendered& r, const std::string& chrom);

}  // namespace cal
