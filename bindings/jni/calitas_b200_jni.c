/* calitas_b200_jni.c — JNI shim between the Scala host (bindings/scala/.../Native.scala) and the C ABI of include/calitas_b200.h.
 *
 * Replaces, on the JVM side, `new SequentialGuideAligner(costs)` and the window / task loops of SearchReference.execute
 * (SearchReference.scala:486-491, 527-564, 584-594, 641-648) and AlignToReference.execute (AlignToReference.scala:64-70, 116-134).
 *
 * Build where a JDK exists:
 *   gcc -O2 -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -I../../include calitas_b200_jni.c \
 *       -L../../calitas_b200 -lcalitas_b200 -Wl,-rpath,'$ORIGIN' -o libcalitas_b200_jni.so
 * This image has no JDK: tests/test_bindings.py compiles the file against bindings/jni/stub/jni.h (declarations only) and checks the exported
 * Java_* symbols against Native.scala; it has never run inside a JVM.
 *
 * Conventions: handles travel as jlong; a failing call throws IllegalArgumentException (CALITAS_EINVAL = the reference's require())
 * or IllegalStateException with calitas_last_error() and returns 0 / NULL; hit records come back as a direct ByteBuffer over the
 * engine's pinned host memory (72-byte little-endian calitas_hit records), valid until hitsetFree(handle).
 */
#include <jni.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "calitas_b200.h"

#define JFN(name) Java_com_editasmedicine_aligner_b200_Native_##name
#define PTR(type, h) ((type*)(intptr_t)(h))

static void throw_for(JNIEnv* env, int rc, const char* fallback) {
  const char* cls = rc == CALITAS_EINVAL ? "java/lang/IllegalArgumentException" : "java/lang/IllegalStateException";
  const char* msg = calitas_last_error();
  jclass c = (*env)->FindClass(env, cls);
  if (c) (*env)->ThrowNew(env, c, (msg && msg[0]) ? msg : fallback);
}

/* String[] guides + String[][] auxPams -> calitas_guide[]; everything is released by free_guides after the call (the engine copies what it keeps). */
typedef struct guide_pack {
  jsize n; calitas_guide* g; jstring* seq_ref; const char** seq; jsize* n_aux; jstring** aux_ref; const char*** aux;
} guide_pack;

static void free_guides(JNIEnv* env, guide_pack* p) {
  if (!p->g) return;
  for (jsize i = 0; i < p->n; ++i) {
    if (p->seq && p->seq[i]) (*env)->ReleaseStringUTFChars(env, p->seq_ref[i], p->seq[i]);
    if (p->aux && p->aux[i]) { for (jsize k = 0; k < p->n_aux[i]; ++k) if (p->aux[i][k]) (*env)->ReleaseStringUTFChars(env, p->aux_ref[i][k], p->aux[i][k]); }
    if (p->aux) free((void*)p->aux[i]);
    if (p->aux_ref) free(p->aux_ref[i]);
  }
  free(p->g); free(p->seq_ref); free((void*)p->seq); free(p->n_aux); free(p->aux_ref); free((void*)p->aux);
  memset(p, 0, sizeof *p);
}

static int pack_guides(JNIEnv* env, jobjectArray guides, jobjectArray auxPams, guide_pack* p) {
  memset(p, 0, sizeof *p);
  if (!guides) return 0;
  p->n = (*env)->GetArrayLength(env, guides);
  const size_t n = (size_t)(p->n > 0 ? p->n : 1);
  p->g = (calitas_guide*)calloc(n, sizeof *p->g); p->seq_ref = (jstring*)calloc(n, sizeof *p->seq_ref); p->seq = (const char**)calloc(n, sizeof *p->seq);
  p->n_aux = (jsize*)calloc(n, sizeof *p->n_aux); p->aux_ref = (jstring**)calloc(n, sizeof *p->aux_ref); p->aux = (const char***)calloc(n, sizeof *p->aux);
  if (!p->g || !p->seq_ref || !p->seq || !p->n_aux || !p->aux_ref || !p->aux) return 0;
  for (jsize i = 0; i < p->n; ++i) {
    p->seq_ref[i] = (jstring)(*env)->GetObjectArrayElement(env, guides, i);
    if (!p->seq_ref[i]) return 0;
    p->seq[i] = (*env)->GetStringUTFChars(env, p->seq_ref[i], NULL);
    if (!p->seq[i]) return 0;
    jobjectArray aux = auxPams ? (jobjectArray)(*env)->GetObjectArrayElement(env, auxPams, i) : NULL;
    p->n_aux[i] = aux ? (*env)->GetArrayLength(env, aux) : 0;
    if (p->n_aux[i] > 0) {
      p->aux_ref[i] = (jstring*)calloc((size_t)p->n_aux[i], sizeof(jstring)); p->aux[i] = (const char**)calloc((size_t)p->n_aux[i], sizeof(char*));
      if (!p->aux_ref[i] || !p->aux[i]) return 0;
      for (jsize k = 0; k < p->n_aux[i]; ++k) {
        p->aux_ref[i][k] = (jstring)(*env)->GetObjectArrayElement(env, aux, k);
        if (!p->aux_ref[i][k]) return 0;
        p->aux[i][k] = (*env)->GetStringUTFChars(env, p->aux_ref[i][k], NULL);
        if (!p->aux[i][k]) return 0;
      }
    }
    p->g[i].sequence = p->seq[i]; p->g[i].aux_pams = p->aux[i]; p->g[i].n_aux_pams = (int32_t)p->n_aux[i];
  }
  return 1;
}

static int read_limits(JNIEnv* env, jintArray limits, calitas_limits* out) {      /* d, p, g, D (< 0 = d + g + p), O */
  jint v[5];
  if (!limits || (*env)->GetArrayLength(env, limits) != 5) return 0;
  (*env)->GetIntArrayRegion(env, limits, 0, 5, v);
  out->max_guide_diffs = v[0]; out->max_pam_mismatches = v[1]; out->max_gaps_between_guide_and_pam = v[2]; out->max_total_diffs = v[3]; out->max_overlap = v[4];
  return 1;
}

static jobject wrap_hits(JNIEnv* env, calitas_hitset* hs, jlongArray outHandle) {
  jlong h = (jlong)(intptr_t)hs;
  (*env)->SetLongArrayRegion(env, outHandle, 0, 1, &h);                                   /* the caller frees with hitsetFree(handle) */
  return (*env)->NewDirectByteBuffer(env, (void*)calitas_hitset_data(hs), (jlong)calitas_hitset_count(hs) * (jlong)sizeof(calitas_hit));
}

/* ---- engine: new SequentialGuideAligner(mismatchNetCost, genomeGapNetCost, guideGapNetCost, pamMismatchNetCost) ------------------------ */
JNIEXPORT jlong JNICALL JFN(engineCreate)(JNIEnv* env, jclass cls, jint device, jintArray costs) {
  (void)cls;
  jint v[4];
  if (!costs || (*env)->GetArrayLength(env, costs) != 4) { throw_for(env, CALITAS_EINVAL, "costs must hold 4 values"); return 0; }
  (*env)->GetIntArrayRegion(env, costs, 0, 4, v);
  calitas_costs cc = { v[0], v[1], v[2], v[3] };                                          /* mismatch, genomeGap, guideGap, pamMismatch */
  calitas_engine* e = NULL;
  const int rc = calitas_engine_create(device, &cc, &e);
  if (rc) { throw_for(env, rc, "calitas_engine_create failed"); return 0; }
  return (jlong)(intptr_t)e;
}

JNIEXPORT void JNICALL JFN(engineDestroy)(JNIEnv* env, jclass cls, jlong engine) { (void)env; (void)cls; calitas_engine_destroy(PTR(calitas_engine, engine)); }

/* ---- contig-range sharding: fills own/have begin/end (each long[nContigs]) for `shard` of `nShards` ---------------------------------- */
JNIEXPORT void JNICALL JFN(shardPlan)(JNIEnv* env, jclass cls, jlongArray lengths, jint shard, jint nShards, jlong halo,
                                      jlongArray ownBegin, jlongArray ownEnd, jlongArray haveBegin, jlongArray haveEnd) {
  (void)cls;
  const jsize n = (*env)->GetArrayLength(env, lengths);
  int64_t* buf = (int64_t*)calloc((size_t)(n > 0 ? n : 1) * 5, sizeof(int64_t));
  if (!buf) { throw_for(env, CALITAS_ESTATE, "out of memory"); return; }
  (*env)->GetLongArrayRegion(env, lengths, 0, n, (jlong*)buf);
  const int rc = calitas_shard_plan((int32_t)n, buf, shard, nShards, halo, buf + n, buf + 2 * n, buf + 3 * n, buf + 4 * n);
  if (rc) throw_for(env, rc, "calitas_shard_plan failed");
  else {
    (*env)->SetLongArrayRegion(env, ownBegin, 0, n, (const jlong*)(buf + n)); (*env)->SetLongArrayRegion(env, ownEnd, 0, n, (const jlong*)(buf + 2 * n));
    (*env)->SetLongArrayRegion(env, haveBegin, 0, n, (const jlong*)(buf + 3 * n)); (*env)->SetLongArrayRegion(env, haveEnd, 0, n, (const jlong*)(buf + 4 * n));
  }
  free(buf);
}

/* ---- reference: the contigs SearchReference.windowIterator reads (SearchReference.scala:39-49), as direct ByteBuffers --------------------
 * bases[c] holds contig bases [haveBegin[c], haveEnd[c]); pass null range arrays to load and own everything. */
JNIEXPORT jlong JNICALL JFN(referenceLoad)(JNIEnv* env, jclass cls, jlong engine, jobjectArray names, jlongArray lengths, jobjectArray bases,
                                           jlongArray haveBegin, jlongArray haveEnd, jlongArray ownBegin, jlongArray ownEnd) {
  (void)cls;
  const jsize n = (*env)->GetArrayLength(env, names);
  const size_t nn = (size_t)(n > 0 ? n : 1);
  jstring* name_ref = (jstring*)calloc(nn, sizeof(jstring)); const char** name = (const char**)calloc(nn, sizeof(char*));
  const uint8_t** ptr = (const uint8_t**)calloc(nn, sizeof(uint8_t*)); int64_t* num = (int64_t*)calloc(nn * 5, sizeof(int64_t));
  calitas_reference* ref = NULL; int rc = CALITAS_ESTATE; int ok = name_ref && name && ptr && num;
  if (ok) {
    (*env)->GetLongArrayRegion(env, lengths, 0, n, (jlong*)num);
    const int ranges = haveBegin && haveEnd && ownBegin && ownEnd;
    if (ranges) {
      (*env)->GetLongArrayRegion(env, haveBegin, 0, n, (jlong*)(num + n)); (*env)->GetLongArrayRegion(env, haveEnd, 0, n, (jlong*)(num + 2 * n));
      (*env)->GetLongArrayRegion(env, ownBegin, 0, n, (jlong*)(num + 3 * n)); (*env)->GetLongArrayRegion(env, ownEnd, 0, n, (jlong*)(num + 4 * n));
    }
    for (jsize c = 0; c < n && ok; ++c) {
      name_ref[c] = (jstring)(*env)->GetObjectArrayElement(env, names, c);
      name[c] = name_ref[c] ? (*env)->GetStringUTFChars(env, name_ref[c], NULL) : NULL;
      jobject b = (*env)->GetObjectArrayElement(env, bases, c);
      ptr[c] = b ? (const uint8_t*)(*env)->GetDirectBufferAddress(env, b) : NULL;
      if (!name[c]) ok = 0;
    }
    if (ok) rc = calitas_reference_load(PTR(calitas_engine, engine), (int32_t)n, name, num, ptr, ranges ? num + n : NULL, ranges ? num + 2 * n : NULL,
                                        ranges ? num + 3 * n : NULL, ranges ? num + 4 * n : NULL, 0, &ref);
    for (jsize c = 0; c < n; ++c) if (name[c]) (*env)->ReleaseStringUTFChars(env, name_ref[c], name[c]);
  }
  free(name_ref); free((void*)name); free((void*)ptr); free(num);
  if (!ok) { throw_for(env, CALITAS_EINVAL, "bad reference arguments"); return 0; }
  if (rc) { throw_for(env, rc, "calitas_reference_load failed"); return 0; }
  return (jlong)(intptr_t)ref;
}

JNIEXPORT void JNICALL JFN(referenceFree)(JNIEnv* env, jclass cls, jlong engine, jlong ref) { (void)env; (void)cls; calitas_reference_free(PTR(calitas_engine, engine), PTR(calitas_reference, ref)); }

/* ---- SearchReference: the window loop + removeOverlaps + ReferenceHit.sort for a batch of guides ---------------------------------------- */
JNIEXPORT jobject JNICALL JFN(search)(JNIEnv* env, jclass cls, jlong engine, jlong ref, jobjectArray guides, jobjectArray auxPams, jintArray limits,
                                      jint windowSize, jstring chrom, jboolean dedup, jlongArray outHandle) {
  (void)cls;
  calitas_limits lim; guide_pack gp; calitas_hitset* hs = NULL; jobject out = NULL;
  if (!read_limits(env, limits, &lim)) { throw_for(env, CALITAS_EINVAL, "limits must hold 5 values"); return NULL; }
  if (!pack_guides(env, guides, auxPams, &gp)) { free_guides(env, &gp); throw_for(env, CALITAS_EINVAL, "bad guide arguments"); return NULL; }
  const char* chrom_utf = chrom ? (*env)->GetStringUTFChars(env, chrom, NULL) : NULL;
  const int rc = calitas_search(PTR(calitas_engine, engine), PTR(const calitas_reference, ref), (int32_t)gp.n, gp.g, &lim, windowSize, chrom_utf, dedup ? 1 : 0, &hs);
  if (chrom_utf) (*env)->ReleaseStringUTFChars(env, chrom, chrom_utf);
  free_guides(env, &gp);
  if (rc) throw_for(env, rc, "calitas_search failed"); else out = wrap_hits(env, hs, outHandle);
  return out;
}

/* ---- aligner.align(query, target, chrom, targetOffset, ...) for many targets (variant windows, SearchReference.scala:584-594) ------------- */
JNIEXPORT jobject JNICALL JFN(alignTargets)(JNIEnv* env, jclass cls, jlong engine, jobjectArray guides, jobjectArray auxPams, jintArray guideIdx,
                                            jobjectArray targets, jintArray targetOffsets, jintArray limits, jboolean best, jlongArray outHandle) {
  (void)cls;
  calitas_limits lim; guide_pack gp; calitas_hitset* hs = NULL; jobject out = NULL;
  if (!read_limits(env, limits, &lim)) { throw_for(env, CALITAS_EINVAL, "limits must hold 5 values"); return NULL; }
  if (!pack_guides(env, guides, auxPams, &gp)) { free_guides(env, &gp); throw_for(env, CALITAS_EINVAL, "bad guide arguments"); return NULL; }
  const jsize n = (*env)->GetArrayLength(env, targets);
  const size_t nn = (size_t)(n > 0 ? n : 1);
  calitas_target_task* tasks = (calitas_target_task*)calloc(nn, sizeof *tasks); jbyteArray* arr = (jbyteArray*)calloc(nn, sizeof *arr);
  jint* gi = (jint*)calloc(nn, sizeof *gi); jint* off = (jint*)calloc(nn, sizeof *off);
  int rc = CALITAS_ESTATE;
  if (tasks && arr && gi && off) {
    (*env)->GetIntArrayRegion(env, guideIdx, 0, n, gi); (*env)->GetIntArrayRegion(env, targetOffsets, 0, n, off);
    for (jsize t = 0; t < n; ++t) {                                                     /* byte[] targets are copied by the JVM or pinned; released below */
      arr[t] = (jbyteArray)(*env)->GetObjectArrayElement(env, targets, t);
      tasks[t].guide_idx = gi[t]; tasks[t].target_offset = off[t];
      tasks[t].length = arr[t] ? (int32_t)(*env)->GetArrayLength(env, arr[t]) : 0;
      tasks[t].bases = arr[t] ? (const uint8_t*)(*env)->GetByteArrayElements(env, arr[t], NULL) : NULL;
    }
    rc = calitas_align_targets(PTR(calitas_engine, engine), (int32_t)gp.n, gp.g, (int64_t)n, tasks, &lim, best ? 1 : 0, &hs);
    for (jsize t = 0; t < n; ++t) if (tasks[t].bases) (*env)->ReleaseByteArrayElements(env, arr[t], (jbyte*)tasks[t].bases, JNI_ABORT);
  }
  free(tasks); free(arr); free(gi); free(off);
  free_guides(env, &gp);
  if (rc) throw_for(env, rc, "calitas_align_targets failed"); else out = wrap_hits(env, hs, outHandle);
  return out;
}

/* ---- aligner.alignToRef / alignToRefBest for many (guide, contig, start, length) regions (AlignToReference.scala:116-134) ---------------- */
JNIEXPORT jobject JNICALL JFN(alignRegions)(JNIEnv* env, jclass cls, jlong engine, jlong ref, jobjectArray guides, jobjectArray auxPams, jintArray guideIdx,
                                            jintArray contigIdx, jlongArray starts, jintArray lengths, jintArray limits, jboolean best, jlongArray outHandle) {
  (void)cls;
  calitas_limits lim; guide_pack gp; calitas_hitset* hs = NULL; jobject out = NULL;
  if (!read_limits(env, limits, &lim)) { throw_for(env, CALITAS_EINVAL, "limits must hold 5 values"); return NULL; }
  if (!pack_guides(env, guides, auxPams, &gp)) { free_guides(env, &gp); throw_for(env, CALITAS_EINVAL, "bad guide arguments"); return NULL; }
  const jsize n = (*env)->GetArrayLength(env, guideIdx);
  const size_t nn = (size_t)(n > 0 ? n : 1);
  calitas_region_task* tasks = (calitas_region_task*)calloc(nn, sizeof *tasks);
  jint* gi = (jint*)calloc(nn, sizeof *gi); jint* ci = (jint*)calloc(nn, sizeof *ci); jint* len = (jint*)calloc(nn, sizeof *len); jlong* st = (jlong*)calloc(nn, sizeof *st);
  int rc = CALITAS_ESTATE;
  if (tasks && gi && ci && len && st) {
    (*env)->GetIntArrayRegion(env, guideIdx, 0, n, gi); (*env)->GetIntArrayRegion(env, contigIdx, 0, n, ci);
    (*env)->GetIntArrayRegion(env, lengths, 0, n, len); (*env)->GetLongArrayRegion(env, starts, 0, n, st);
    for (jsize t = 0; t < n; ++t) { tasks[t].guide_idx = gi[t]; tasks[t].contig_idx = ci[t]; tasks[t].start = st[t]; tasks[t].length = len[t]; }
    rc = calitas_align_regions(PTR(calitas_engine, engine), PTR(const calitas_reference, ref), (int32_t)gp.n, gp.g, (int64_t)n, tasks, &lim, best ? 1 : 0, &hs);
  }
  free(tasks); free(gi); free(ci); free(len); free(st);
  free_guides(env, &gp);
  if (rc) throw_for(env, rc, "calitas_align_regions failed"); else out = wrap_hits(env, hs, outHandle);
  return out;
}

JNIEXPORT void JNICALL JFN(hitsetFree)(JNIEnv* env, jclass cls, jlong handle) { (void)env; (void)cls; calitas_hitset_free(PTR(calitas_hitset, handle)); }
