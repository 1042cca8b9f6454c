/*
 * calitas_b200_tools.h — host-side mirror of the reference's operator interface, on top of the C ABI in calitas_b200.h.
 *
 * These entry points keep the names, argument meaning and error behaviour of the reference operators so that a JVM shim, the
 * `calitas` CLI in this repo and the parity tests all drive the GPU engine the way the reference drives its CPU aligner:
 *   calitas_tool_align              SequentialGuideAligner.align          (SequentialGuideAligner.scala:228-323)
 *   calitas_tool_align_best         SequentialGuideAligner.alignBest      (:333-345)
 *   calitas_tool_align_to_ref       alignToRef / alignToRefBest           (:359-418)
 *   calitas_tool_search_reference   SearchReference.execute               (SearchReference.scala:513-649)
 *   calitas_tool_align_to_reference AlignToReference.execute              (AlignToReference.scala:95-147)
 *   calitas_tool_pairwise_align     PairwiseAlignSequences.execute        (PairwiseAlignSequences.scala:44-83)
 * All computation of alignments happens on the device through calitas_search / calitas_align_regions / calitas_align_targets;
 * this layer parses, builds windows for variants, renders text and (for VCF runs) merges reference and variant hits.
 * Text results are malloc'd; release with calitas_free_text.  Alignment rows use the columns of calitas_render_alignments;
 * hit tables are the reference's 34-column TSV (ReferenceHit.scala:99-132) with a header line.
 */
#ifndef CALITAS_B200_TOOLS_H
#define CALITAS_B200_TOOLS_H
#include "calitas_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

/* Host view of the genome the engine's reference was loaded from (caller-owned, full contigs): needed to render target
 * strings and flanks (ReferenceHit.scala:213-216, 261-266).  assembly = first AS of the .dict or NULL ("unknown"). */
typedef struct calitas_genome_view {
  int32_t n_contigs; const char* const* names; const int64_t* lengths; const uint8_t* const* bases; const char* assembly;
} calitas_genome_view;

int calitas_tool_align(calitas_engine* e, const calitas_guide* guide, const uint8_t* target, int32_t target_len, const char* target_name,
                       int32_t target_offset, const calitas_limits* limits, char** out_text);
int calitas_tool_align_best(calitas_engine* e, const calitas_guide* guide, const uint8_t* target, int32_t target_len,
                            int32_t max_gaps_between_guide_and_pam, char** out_text);
/* window_size < 0: None (2 x guide length each side).  best != 0: alignToRefBest (one row). */
int calitas_tool_align_to_ref(calitas_engine* e, const calitas_reference* ref, const calitas_genome_view* genome, const calitas_guide* guide,
                              const char* chrom, int32_t pos, int32_t window_size, int32_t best, const calitas_limits* limits, char** out_text);

typedef struct calitas_search_options {      /* SearchReference.scala:452-470 */
  const char* guide_id;                      /* -I */
  int32_t max_variants;                      /* -V, default 16 */
  int32_t window_size;                       /* -w, default 1000 */
  calitas_limits limits;                     /* -d -p -g -D -O */
  const char* chrom;                         /* -c or NULL */
  const char* vcf_text;                      /* -v: contents of the (plain-text) VCF, or NULL */
  const char* vcf_id;                        /* "<file name>:<md5>" (ReferenceHit.scala:175-183) */
  const char* time_stamp;                    /* run-dependent column; NULL -> empty */
  const char* aligner_version;               /* run-dependent column; NULL -> "calitas-b200" */
} calitas_search_options;
int calitas_tool_search_reference(calitas_engine* e, const calitas_reference* ref, const calitas_genome_view* genome, const calitas_guide* guide,
                                  const calitas_search_options* opt, char** out_tsv, int64_t* n_hits);

/* The same for a batch of guides over n_engines engines (one per GPU; engine s holds shard s of calitas_shard_plan(.., s, n_engines, ..) in refs[s]).
 * The reference runs one guide per process; the engine batches guides per scan and the shards run concurrently on host threads.
 * Rows come guide by guide, each guide's block exactly as the single-guide call produces it; guide_ids[g] is that guide's -I. */
int calitas_tool_search_reference_batch(int32_t n_engines, calitas_engine* const* engines, const calitas_reference* const* refs, const calitas_genome_view* genome,
                                        int32_t n_guides, const calitas_guide* guides, const char* const* guide_ids, const calitas_search_options* opt,
                                        char** out_tsv, int64_t* n_hits);

/* The same, with the table written to an open file descriptor instead of returned: without a VCF the rows are rendered on all host threads and
 * leave block by block, in order, while later blocks are still being rendered (the 100-guide hg38 table is 18 GB; it never sits in memory whole).
 * Replaces Metric.writer over the keepers (SearchReference.scala:646-648).  *n_bytes = bytes written. */
int calitas_tool_search_reference_batch_fd(int32_t n_engines, calitas_engine* const* engines, const calitas_reference* const* refs, const calitas_genome_view* genome,
                                           int32_t n_guides, const calitas_guide* guides, const char* const* guide_ids, const calitas_search_options* opt,
                                           int32_t out_fd, int64_t* n_hits, int64_t* n_bytes);

typedef struct calitas_a2r_task { const char* id; const char* query; const char* chrom; int32_t position; } calitas_a2r_task;   /* AlignToReference.scala:97-102 */
typedef struct calitas_a2r_options {         /* AlignToReference.scala:35-50; -1 = not given */
  int32_t window_size, max_guide_diffs, max_pam_mismatches, max_gaps_between_guide_and_pam, max_total_diffs, max_overlap;
  const char* time_stamp; const char* aligner_version;
} calitas_a2r_options;
int calitas_tool_align_to_reference(calitas_engine* e, const calitas_reference* ref, const calitas_genome_view* genome, int64_t n_tasks,
                                    const calitas_a2r_task* tasks, const calitas_a2r_options* opt, char** out_tsv, int64_t* n_hits);

/* PairwiseAlignSequences.execute (PairwiseAlignSequences.scala:44-83): one alignBest per (query, target) pair; 11-column table with header. */
int calitas_tool_pairwise_align(calitas_engine* e, int64_t n_pairs, const char* const* queries, const char* const* targets, char** out_tsv);

/* SearchReference -v in two steps, for callers that search the same VCF repeatedly: the plan holds the parsed VCF, the variant windows of every padding
 * class of the guides (SearchReference.scala:217-399, 575) and, per engine, the windows it processes (owned, or halo around its shard's cuts);
 * calitas_tool_variant_plan_search runs calitas_search_variants for one engine of it (its first call uploads that
 * engine's windows: calitas_variant_set_load).  calitas_tool_search_reference* build such a plan internally. */
typedef struct calitas_variant_plan calitas_variant_plan;
int calitas_tool_variant_plan_create(const calitas_genome_view* genome, const calitas_search_options* opt, int32_t n_guides, const calitas_guide* guides,
                                     int32_t n_engines, const calitas_reference* const* refs, calitas_variant_plan** out);
void calitas_tool_variant_plan_free(calitas_variant_plan* p);
int calitas_tool_variant_plan_counts(const calitas_variant_plan* p, int32_t engine, int64_t* n_records, int64_t* n_windows, int64_t* n_window_bases);
int calitas_tool_variant_plan_search(calitas_variant_plan* p, int32_t engine, calitas_engine* e, const calitas_reference* ref, int32_t n_guides, const calitas_guide* guides,
                                     const calitas_limits* limits, int32_t window_size, const char* chrom, calitas_hitset** out);

/* Inspection hook for the variant path (SearchReference.scala:217-399): one line per variant window
 * "chrom \t start \t cigar \t bases \t id:pos:ref>alt;..." in iterator order. */
int calitas_tool_variant_windows(const calitas_genome_view* genome, const char* vcf_text, const char* chrom, int32_t padding, int32_t max_variants, char** out_text);

#ifdef __cplusplus
}
#endif
#endif
