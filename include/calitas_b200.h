/*
 * calitas_b200.h — C ABI of the B200-native CALITAS off-target search engine.
 *
 * This is the drop-in boundary for the reference's hot path.  The reference (Scala/JVM) has no FFI of its
 * own; the functions below are what a JNI/Panama shim would bind in place of the calls cited beside each one
 * (reference files under calitas/src/main/scala/com/editasmedicine/aligner/; see INTEGRATION.md for the stub).
 *
 * Conventions
 *   - every function returns 0 on success and a non-zero CALITAS_E* code on failure; the message is
 *     available from calitas_last_error() (thread-local) — the JVM side rethrows it as
 *     IllegalArgumentException (CALITAS_EINVAL, mirrors the reference's `require`) or IllegalStateException;
 *   - plain pointers and sizes only; the caller owns every input for the duration of the call, the engine
 *     copies what it keeps; result sets are owned by the engine until *_free;
 *   - coordinates are 0-based half-open, as in GuideAlignment.scala:60-63;
 *   - one engine may be used from one host thread at a time; engines are independent (one per GPU);
 *   - there is no CPU fallback: every entry point that computes fails with CALITAS_ECUDA without a device.
 */
#ifndef CALITAS_B200_H
#define CALITAS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CALITAS_OK      0
#define CALITAS_EINVAL  1   /* bad argument: mirrors the reference's require()/validate() failures */
#define CALITAS_ECUDA   2   /* CUDA runtime / device failure, or no device */
#define CALITAS_ELIMIT  3   /* input exceeds a documented engine limit (protospacer > 32 nt, PAM > 16 nt, > 8 PAMs, ...) */
#define CALITAS_ESTATE  4   /* wrong call order / unknown handle */

#define CALITAS_MAX_PROTOSPACER 32
#define CALITAS_MAX_PAM_LEN     16
#define CALITAS_MAX_PAMS        8
#define CALITAS_MAX_OPS         128   /* alignment columns per hit (48 in a calitas_hit, more in a calitas_hit_wide) */
#define CALITAS_MAX_GUIDES      8192  /* guides per call (13 bits of calitas_hit.where) */

typedef struct calitas_engine calitas_engine;
typedef struct calitas_reference calitas_reference;
typedef struct calitas_hitset calitas_hitset;

/* Net costs, SequentialGuideAligner.scala:17-21,170-175 (either sign accepted, :193-198). */
typedef struct calitas_costs {
  int32_t mismatch_net_cost;      /* -m, default -120 */
  int32_t genome_gap_net_cost;    /* -b, default -122 */
  int32_t guide_gap_net_cost;     /* -B, default -121 */
  int32_t pam_mismatch_net_cost;  /* -M, default -260 */
} calitas_costs;

/* Limits of one SequentialGuideAligner.align call (SequentialGuideAligner.scala:228-236). */
typedef struct calitas_limits {
  int32_t max_guide_diffs;                 /* -d */
  int32_t max_pam_mismatches;              /* -p */
  int32_t max_gaps_between_guide_and_pam;  /* -g */
  int32_t max_total_diffs;                 /* -D; < 0 means d + g + p (SearchReference.scala:493) */
  int32_t max_overlap;                     /* -O */
} calitas_limits;

/* One guide: "PROTOSPACERpam" / "pamPROTOSPACER" / "PROTOSPACER" plus auxiliary PAMs (Guide.apply, SequentialGuideAligner.scala:81-107). */
typedef struct calitas_guide {
  const char* sequence;
  const char* const* aux_pams;
  int32_t n_aux_pams;
} calitas_guide;

/* One alignment: the POD image of GuideAlignment (GuideAlignment.scala:72-88), 32 bytes.  Strings (padded guide / alignment / target, cigar
 * text, the derived counters of GuideAlignment.scala:99-163) are rendered on the host from `ops` and the reference bases by
 * calitas_render_*; ops are in guide orientation, 2 bits per alignment column, column k in bits 2(k mod 16) of ops[k / 16]:
 * 0 '=', 1 'X', 2 'I' (guide base opposite a genome gap), 3 'D' (genome base opposite a guide gap).
 * A result set holds records of ONE size, calitas_hitset_stride(): 32 bytes (calitas_hit, up to 48 alignment columns) when no guide of the
 * call can produce a longer alignment, else 64 bytes (calitas_hit_wide: the same 20-byte header, up to 176 columns).  Record i starts at
 * (const char*)calitas_hitset_data(h) + i * stride; the accessors below work on either form. */
#define CALITAS_HIT_WORDS       8    /* 32-bit words of a calitas_hit */
#define CALITAS_HIT_WIDE_WORDS 16    /* 32-bit words of a calitas_hit_wide */
#define CALITAS_HIT_HEADER_WORDS 5
typedef struct calitas_hit {
  int32_t  start_offset;      /* guide+PAM span [start_offset, end_offset) */
  int32_t  task_idx;          /* window index (search) or task index (align_regions / align_targets) */
  int32_t  score;
  uint32_t where;             /* bits 0-12 guide_idx (index into the guides array of the call); bits 13-30 contig_idx + 1 (0: none, calitas_align_targets); bit 31 strand ('-' = 1) */
  uint32_t shape;             /* bits 0-7 n_ops; 8-15 end_offset - start_offset; 16-21 guide_start_offset - start_offset; 22-27 end_offset - guide_end_offset; 28-31 pam_idx + 1 */
  uint32_t ops[CALITAS_HIT_WORDS - CALITAS_HIT_HEADER_WORDS];
} calitas_hit;
typedef struct calitas_hit_wide {
  int32_t start_offset, task_idx, score; uint32_t where, shape;
  uint32_t ops[CALITAS_HIT_WIDE_WORDS - CALITAS_HIT_HEADER_WORDS];
} calitas_hit_wide;
static inline int32_t calitas_hit_guide_idx(const calitas_hit* h) { return (int32_t)(h->where & 0x1FFFu); }
static inline int32_t calitas_hit_contig_idx(const calitas_hit* h) { return (int32_t)((h->where >> 13) & 0x3FFFFu) - 1; }
static inline char    calitas_hit_strand(const calitas_hit* h) { return (h->where >> 31) ? '-' : '+'; }
static inline int32_t calitas_hit_n_ops(const calitas_hit* h) { return (int32_t)(h->shape & 0xFFu); }
static inline int32_t calitas_hit_end_offset(const calitas_hit* h) { return h->start_offset + (int32_t)((h->shape >> 8) & 0xFFu); }
/* protospacer-only span: coordinate_start / coordinate_end of ReferenceHit.scala:223-224 */
static inline int32_t calitas_hit_guide_start_offset(const calitas_hit* h) { return h->start_offset + (int32_t)((h->shape >> 16) & 0x3Fu); }
static inline int32_t calitas_hit_guide_end_offset(const calitas_hit* h) { return calitas_hit_end_offset(h) - (int32_t)((h->shape >> 22) & 0x3Fu); }
/* index into that guide's PAM list (primary = 0); -1 for a PAM-less guide */
static inline int32_t calitas_hit_pam_idx(const calitas_hit* h) { return (int32_t)(h->shape >> 28) - 1; }
static inline uint32_t calitas_hit_op(const calitas_hit* h, int32_t k) { return (h->ops[k >> 4] >> ((k & 15) * 2)) & 3u; }
/* GuideAlignment.gapBases (columns with op I or D) and GuideAlignment.edits (columns with op != '=') */
static inline int32_t calitas_hit_gap_bases(const calitas_hit* h) { int32_t n = calitas_hit_n_ops(h), c = 0, k; for (k = 0; k < n; ++k) c += calitas_hit_op(h, k) >= 2u; return c; }
static inline int32_t calitas_hit_edits(const calitas_hit* h) { int32_t n = calitas_hit_n_ops(h), c = 0, k; for (k = 0; k < n; ++k) c += calitas_hit_op(h, k) != 0u; return c; }

/* ---- engine ------------------------------------------------------------------------------------------- */
/* Replaces `new SequentialGuideAligner(costs)` (SearchReference.scala:486-491, AlignToReference.scala:64-70). */
int calitas_engine_create(int32_t device_id, const calitas_costs* costs, calitas_engine** out);
void calitas_engine_destroy(calitas_engine* e);
int calitas_engine_get_costs(const calitas_engine* e, calitas_costs* out);
const char* calitas_last_error(void);

/* ---- reference loading and packing ------------------------------------------------------------------------
 * Replaces SearchReference.windowIterator's per-contig byte arrays (SearchReference.scala:39-49) and the indexed
 * FASTA behind alignToRef (SequentialGuideAligner.scala:369-374).  Contig c has total length lengths[c]; this engine
 * is given bases[c][0 .. have_end[c]-have_begin[c]) = contig bases [have_begin[c], have_end[c]) and owns the reference
 * windows whose nominal start lies in [own_begin[c], own_end[c]).  Pass NULL for the four range arrays to load and own
 * everything (single GPU).  The bytes are uploaded, then packed on the device to 4-bit IUPAC-set codes. */
int calitas_reference_load(calitas_engine* e, int32_t n_contigs, const char* const* names, const int64_t* lengths,
                           const uint8_t* const* bases, const int64_t* have_begin, const int64_t* have_end,
                           const int64_t* own_begin, const int64_t* own_end, int32_t keep_raw, calitas_reference** out);
void calitas_reference_free(calitas_engine* e, calitas_reference* r);
/* The window starts of contig `contig` this reference owns (the own_begin / own_end it was loaded with; the whole contig for an unsharded load). */
int calitas_reference_own_range(const calitas_reference* r, int32_t contig, int64_t* own_begin, int64_t* own_end);
/* Contiguous split of the genome into n_shards base ranges for contig-range sharding: fills, for `shard`,
 * own_begin/own_end per contig (window starts owned) and have_begin/have_end (bases needed, incl. a halo of `halo` bases). */
int calitas_shard_plan(int32_t n_contigs, const int64_t* lengths, int32_t shard, int32_t n_shards, int64_t halo,
                       int64_t* own_begin, int64_t* own_end, int64_t* have_begin, int64_t* have_end);

/* ---- SearchReference hot path ---------------------------------------------------------------------------------
 * Replaces the window loop of SearchReference.execute (SearchReference.scala:527-564): windowIterator + one
 * SequentialGuideAligner.align per window, for every guide of the batch, then (dedup != 0) removeOverlaps + ReferenceHit.sort
 * (SearchReference.scala:641-648, 653-675; ReferenceHit.scala:276-287).  chrom may be NULL (all contigs, -c absent).
 * On a sharded reference (own ranges that do not cover every contig) dedup != 0 requires max_overlap >= 1: with max_overlap <= 0 every later
 * hit of a group "overlaps", the reference's sweep reaches across the whole contig and a shard cannot reproduce it (CALITAS_EINVAL; search with
 * dedup = 0 and run removeOverlaps over the gathered hits, as calitas_tool_search_reference_batch does). */
int calitas_search(calitas_engine* e, const calitas_reference* ref, int32_t n_guides, const calitas_guide* guides,
                   const calitas_limits* limits, int32_t window_size, const char* chrom, int32_t dedup, calitas_hitset** out);

/* The same over n_engines engines (one per GPU; engine s holds shard s of calitas_shard_plan(.., s, n_engines, ..) in refs[s]), ending with ONE table in
 * one address space, as SearchReference.execute does (SearchReference.scala:641-648): the engines run concurrently on host threads, then each guide's
 * per-shard lists are concatenated in shard order and merged by the ReferenceHit.sort key where they meet (consecutive windows overlap, so hits of
 * neighbouring shards can interleave around a cut).  The result is owned by engines[0]; stats: ms[] = maximum over the engines, ms[6] = the host-side
 * merge, counts[] = sums.  Needs max_overlap >= 1 (see calitas_search). */
int calitas_search_sharded(int32_t n_engines, calitas_engine* const* engines, const calitas_reference* const* refs, int32_t n_guides,
                           const calitas_guide* guides, const calitas_limits* limits, int32_t window_size, const char* chrom, calitas_hitset** out);

/* ---- SearchReference -v: reference windows and variant windows, merged on the device -------------------------------------------------------
 * Replaces the variant loop of SearchReference.execute (SearchReference.scala:570-630) together with removeOverlaps + ReferenceHit.sort over the union
 * of both hit lists (:641-648, 653-675).  The host builds the variant windows (VariantWindow / VariantSet, :101-400); the engine aligns every guide g
 * against the windows w with windows[w].guide_class == guide_class[g] (windows depend on the guide only through its padding, :575), maps each hit back
 * to reference coordinates (VariantWindow.refOffsetAtBaseOffset, :133-156), names the variant set it overlaps (ReferenceHit.scala:211) and
 * de-duplicates per (guide, contig, strand, variant set).
 *   alleles[]   the chosen allele of each variant of each window, windows refer to runs of it (position order);
 *   set_rank[]  for a window with m alleles, set_rank[first_set + a*m - a*(a-1)/2 + (b-a-1)] names the variant set alleles[a..b) (0 <= a < b <= m): equal
 *               sets carry equal values > 0, and the values order the groups the way the caller wants rows that tie on (contig, start, strand, score)
 *               ordered (the reference's own order among such rows is a HashMap's);
 *   owned = 0   marks a halo window of a sharded run: its hits take part in removeOverlaps and are never reported.
 * The windows live on the device as a calitas_variant_set; calitas_search_variants builds the (guide, window) tasks there.
 * The result set holds records in ReferenceHit.sort order per guide, and one calitas_variant_hit_info per record (calitas_hitset_variant_info):
 * hits of variant windows keep window-relative offsets in the record (task_idx = window index) and carry the reference offsets here. */
typedef struct calitas_variant_allele { int32_t pos /* 1-based POS */, ref_len, alt_len; } calitas_variant_allele;
typedef struct calitas_variant_window {
  const uint8_t* bases; int32_t length;     /* the window's bases with its alleles applied (upper case) */
  int32_t contig_idx; int32_t ref_start;    /* 1-based reference position of the first base */
  int32_t n_alleles; int32_t first_allele; int32_t first_set; int32_t guide_class; int32_t owned;
} calitas_variant_window;
typedef struct calitas_variant_hit_info {
  int32_t window_idx;                        /* -1: hit of a reference window (offsets below = the record's) */
  int32_t start_offset, end_offset, guide_start_offset, guide_end_offset;      /* reference coordinates (SearchReference.scala:615-620) */
  int32_t set_rank;                          /* 0: the hit overlaps no variant of its window */
} calitas_variant_hit_info;
/* The variant windows of one VCF on the engine's device (uploaded and packed once, like the reference; sorted by guide_class). */
typedef struct calitas_variant_set calitas_variant_set;
int calitas_variant_set_load(calitas_engine* e, const calitas_reference* ref, int64_t n_windows, const calitas_variant_window* windows,
                             int64_t n_alleles, const calitas_variant_allele* alleles, int64_t n_sets, const uint32_t* set_rank, calitas_variant_set** out);
void calitas_variant_set_free(calitas_variant_set* variants);
int calitas_search_variants(calitas_engine* e, const calitas_reference* ref, int32_t n_guides, const calitas_guide* guides, const int32_t* guide_class,
                            const calitas_limits* limits, int32_t window_size, const char* chrom, const calitas_variant_set* variants, calitas_hitset** out);
const calitas_variant_hit_info* calitas_hitset_variant_info(const calitas_hitset* h);   /* NULL unless the set came from calitas_search_variants */

/* ---- AlignToReference / variant windows -----------------------------------------------------------------------
 * One SequentialGuideAligner.align per task (SequentialGuideAligner.scala:228), batched.  best != 0 applies the
 * alignBest/alignToRefBest limits per guide (SequentialGuideAligner.scala:336-343, 407-417: d = protospacer length,
 * p = PAM length, maxTotal = d + g + p, maxOverlap = 0) and `limits` supplies only max_gaps_between_guide_and_pam. */
typedef struct calitas_region_task { int32_t guide_idx; int32_t contig_idx; int64_t start; int32_t length; } calitas_region_task;        /* alignToRef: bases [start, start+length) */
typedef struct calitas_target_task { int32_t guide_idx; const uint8_t* bases; int32_t length; int32_t target_offset; } calitas_target_task; /* align(guide, target, targetOffset) */
int calitas_align_regions(calitas_engine* e, const calitas_reference* ref, int32_t n_guides, const calitas_guide* guides,
                          int64_t n_tasks, const calitas_region_task* tasks, const calitas_limits* limits, int32_t best, calitas_hitset** out);
int calitas_align_targets(calitas_engine* e, int32_t n_guides, const calitas_guide* guides,
                          int64_t n_tasks, const calitas_target_task* tasks, const calitas_limits* limits, int32_t best, calitas_hitset** out);

/* ---- result sets ---------------------------------------------------------------------------------------------
 * Hits are ordered as the reference would emit them: search with dedup: ReferenceHit.sort order per guide;
 * otherwise by (guide, window/task, '+' before '-', rank in the per-window retval of SequentialGuideAligner.scala:315-322). */
int64_t calitas_hitset_count(const calitas_hitset* h);
const calitas_hit* calitas_hitset_data(const calitas_hitset* h);   /* pinned host memory, valid until free; records are calitas_hitset_stride() bytes apart */
int32_t calitas_hitset_stride(const calitas_hitset* h);            /* 32 (calitas_hit) or 64 (calitas_hit_wide) */
void calitas_hitset_free(calitas_hitset* h);
/* Timings of the call that produced the set, milliseconds between CUDA events recorded on the stream each piece runs on:
 *   ms[0] whole call on the device, first launch to the last byte of the final D2H;
 *   ms[1] scan kernels (summed over guide chunks), ms[2] align kernels, ms[3] sorts + canonicalise + dedup, ms[4] D2H of hit records.
 *   calitas_search overlaps the scan of guide chunk c+1 with [2]-[4] of chunk c on separate streams, so [1]-[4] add up to more than [0];
 *   ms[5] = ms[0] - ms[1], the part of the call not hidden behind the scan kernels (0 for the align_* entry points).
 * counts[0] owned windows, [1] candidate end columns, [2] alignment slots before canonicalisation, [3] kernel launches,
 * [4] bytes copied host->device, [5] bytes copied device->host, [6] scan-kernel launches, [7] reference bases scanned (summed over scan launches). */
int calitas_hitset_stats(const calitas_hitset* h, double ms[8], int64_t counts[8]);

/* Integer-issue microbenchmark used as the roofline denominator of the scan kernel (no integer peak is published or in
 * MEASURED_PEAKS.json).  Result = executed thread-level integer instructions per second / 1e12.  kind 0: LOP3 chains (ALU pipe),
 * 1: IMAD chains (FMA pipe), 2: LOP3 and IMAD 1:1 (both pipes, scheduler issue limit), 3: IMAD.HI, 4: LEA.HI. */
int calitas_microbench_int(calitas_engine* e, int32_t kind, double* tera_ops_per_s);

/* ---- host-side rendering (ReferenceHit.Builder.build, ReferenceHit.scala:210-254; GuideAlignment.scala:10-50,99-163) ---- */
/* Renders hits as tab-separated GuideAlignment rows (same columns as oracle_alignment_header()).  `contig_bases[c]` must
 * point at base 0 of contig c (NULL allowed for contigs without hits); for align_targets pass the task bases through
 * `target_bases[task]` instead.  upper_case != 0 upper-cases target bases (SearchReference windows, SearchReference.scala:67).
 * *out_text is malloc'd; release with calitas_free_text. */
int calitas_render_alignments(const void* hits, int64_t n_hits, int32_t stride, int32_t n_guides, const calitas_guide* guides,
                              int32_t n_contigs, const char* const* names, const uint8_t* const* contig_bases,
                              const calitas_target_task* targets, int32_t upper_case, char** out_text);
void calitas_free_text(char* text);

#ifdef __cplusplus
}
#endif
#endif
