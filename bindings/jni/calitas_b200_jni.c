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
 * engine's pinned host memory (little-endian calitas_hit records, hitsetStride(handle) = 32 or 64 bytes each), valid until hitsetFree(handle).
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

/* A Java String copied into malloc'd memory; the local reference and the JVM's UTF buffer are released at once, so that loops over thousands of
 * elements never fill the local-reference table (JNI guarantees 16 slots) nor hold JVM buffers. */
static char* dup_jstring(JNIEnv* env, jstring s) {
  if (!s) return NULL;
  const char* u = (*env)->GetStringUTFChars(env, s, NULL);
  char* c = NULL;
  if (u) { size_t n = strlen(u); c = (char*)malloc(n + 1); if (c) memcpy(c, u, n + 1); (*env)->ReleaseStringUTFChars(env, s, u); }
  return c;
}
static char* dup_array_string(JNIEnv* env, jobjectArray arr, jsize i) {
  jstring s = (jstring)(*env)->GetObjectArrayElement(env, arr, i);
  char* c = dup_jstring(env, s);
  if (s) (*env)->DeleteLocalRef(env, s);
  return c;
}

/* String[] guides + String[][] auxPams -> calitas_guide[] over owned copies; free_guides releases them after the call (the engine copies what it keeps). */
typedef struct guide_pack { jsize n; calitas_guide* g; char** seq; jsize* n_aux; char*** aux; } guide_pack;

static void free_guides(guide_pack* p) {
  if (!p->g) return;
  for (jsize i = 0; i < p->n; ++i) {
    if (p->seq) free(p->seq[i]);
    if (p->aux && p->aux[i]) { for (jsize k = 0; k < p->n_aux[i]; ++k) free(p->aux[i][k]); free(p->aux[i]); }
  }
  free(p->g); free(p->seq); free(p->n_aux); free(p->aux);
  memset(p, 0, sizeof *p);
}

static int pack_guides(JNIEnv* env, jobjectArray guides, jobjectArray auxPams, guide_pack* p) {
  memset(p, 0, sizeof *p);
  if (!guides) return 0;
  p->n = (*env)->GetArrayLength(env, guides);
  const size_t n = (size_t)(p->n > 0 ? p->n : 1);
  p->g = (calitas_guide*)calloc(n, sizeof *p->g); p->seq = (char**)calloc(n, sizeof *p->seq);
  p->n_aux = (jsize*)calloc(n, sizeof *p->n_aux); p->aux = (char***)calloc(n, sizeof *p->aux);
  if (!p->g || !p->seq || !p->n_aux || !p->aux) return 0;
  for (jsize i = 0; i < p->n; ++i) {
    p->seq[i] = dup_array_string(env, guides, i);
    if (!p->seq[i]) return 0;
    jobjectArray aux = auxPams ? (jobjectArray)(*env)->GetObjectArrayElement(env, auxPams, i) : NULL;
    p->n_aux[i] = aux ? (*env)->GetArrayLength(env, aux) : 0;
    if (p->n_aux[i] > 0) {
      p->aux[i] = (char**)calloc((size_t)p->n_aux[i], sizeof(char*));
      if (!p->aux[i]) { (*env)->DeleteLocalRef(env, aux); return 0; }
      for (jsize k = 0; k < p->n_aux[i]; ++k) {
        p->aux[i][k] = dup_array_string(env, aux, k);
        if (!p->aux[i][k]) { (*env)->DeleteLocalRef(env, aux); return 0; }
      }
    }
    if (aux) (*env)->DeleteLocalRef(env, aux);
    p->g[i].sequence = p->seq[i]; p->g[i].aux_pams = (const char* const*)p->aux[i]; p->g[i].n_aux_pams = (int32_t)p->n_aux[i];
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
  return (*env)->NewDirectByteBuffer(env, (void*)calitas_hitset_data(hs), (jlong)calitas_hitset_count(hs) * (jlong)calitas_hitset_stride(hs));   /* 32- or 64-byte records: hitsetStride(handle) */
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
  char** name = (char**)calloc(nn, sizeof(char*));
  const uint8_t** ptr = (const uint8_t**)calloc(nn, sizeof(uint8_t*)); int64_t* num = (int64_t*)calloc(nn * 5, sizeof(int64_t));
  calitas_reference* ref = NULL; int rc = CALITAS_ESTATE; int ok = name && ptr && num;
  const char* why = "bad reference arguments";
  if (ok) {
    (*env)->GetLongArrayRegion(env, lengths, 0, n, (jlong*)num);
    const int ranges = haveBegin && haveEnd && ownBegin && ownEnd;
    if (ranges) {
      (*env)->GetLongArrayRegion(env, haveBegin, 0, n, (jlong*)(num + n)); (*env)->GetLongArrayRegion(env, haveEnd, 0, n, (jlong*)(num + 2 * n));
      (*env)->GetLongArrayRegion(env, ownBegin, 0, n, (jlong*)(num + 3 * n)); (*env)->GetLongArrayRegion(env, ownEnd, 0, n, (jlong*)(num + 4 * n));
    }
    for (jsize c = 0; c < n && ok; ++c) {
      name[c] = dup_array_string(env, names, c);
      if (!name[c]) { ok = 0; break; }
      const int64_t need = ranges ? num[2 * n + c] - num[n + c] : num[c];                 /* bases the engine will read from this buffer */
      jobject b = (*env)->GetObjectArrayElement(env, bases, c);
      if (b) {
        ptr[c] = (const uint8_t*)(*env)->GetDirectBufferAddress(env, b);
        const jlong cap = (*env)->GetDirectBufferCapacity(env, b);
        (*env)->DeleteLocalRef(env, b);
        if (!ptr[c] || cap < need) { ok = 0; why = "a contig's direct buffer is shorter than its base range (or not a direct buffer)"; }
      } else if (need > 0) { ok = 0; why = "a contig with bases to load has no buffer"; }
    }
    if (ok) rc = calitas_reference_load(PTR(calitas_engine, engine), (int32_t)n, (const char* const*)name, num, ptr, ranges ? num + n : NULL, ranges ? num + 2 * n : NULL,
                                        ranges ? num + 3 * n : NULL, ranges ? num + 4 * n : NULL, 0, &ref);
    for (jsize c = 0; c < n; ++c) free(name[c]);
  }
  free(name); free((void*)ptr); free(num);
  if (!ok) { jclass c = (*env)->FindClass(env, "java/lang/IllegalArgumentException"); if (c) (*env)->ThrowNew(env, c, why); return 0; }
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
  if (!pack_guides(env, guides, auxPams, &gp)) { free_guides(&gp); throw_for(env, CALITAS_EINVAL, "bad guide arguments"); return NULL; }
  char* chrom_c = dup_jstring(env, chrom);
  const int rc = calitas_search(PTR(calitas_engine, engine), PTR(const calitas_reference, ref), (int32_t)gp.n, gp.g, &lim, windowSize, chrom_c, dedup ? 1 : 0, &hs);
  free(chrom_c);
  free_guides(&gp);
  if (rc) throw_for(env, rc, "calitas_search failed"); else out = wrap_hits(env, hs, outHandle);
  return out;
}

/* ---- aligner.align(query, target, chrom, targetOffset, ...) for many targets (variant windows, SearchReference.scala:584-594) ------------- */
JNIEXPORT jobject JNICALL JFN(alignTargets)(JNIEnv* env, jclass cls, jlong engine, jobjectArray guides, jobjectArray auxPams, jintArray guideIdx,
                                            jobjectArray targets, jintArray targetOffsets, jintArray limits, jboolean best, jlongArray outHandle) {
  (void)cls;
  calitas_limits lim; guide_pack gp; calitas_hitset* hs = NULL; jobject out = NULL;
  if (!read_limits(env, limits, &lim)) { throw_for(env, CALITAS_EINVAL, "limits must hold 5 values"); return NULL; }
  if (!pack_guides(env, guides, auxPams, &gp)) { free_guides(&gp); throw_for(env, CALITAS_EINVAL, "bad guide arguments"); return NULL; }
  const jsize n = (*env)->GetArrayLength(env, targets);
  const size_t nn = (size_t)(n > 0 ? n : 1);
  calitas_target_task* tasks = (calitas_target_task*)calloc(nn, sizeof *tasks);
  jint* gi = (jint*)calloc(nn, sizeof *gi); jint* off = (jint*)calloc(nn, sizeof *off); size_t* at = (size_t*)calloc(nn + 1, sizeof *at);
  uint8_t* pool = NULL;
  int rc = CALITAS_ESTATE;
  if (tasks && gi && off && at) {
    (*env)->GetIntArrayRegion(env, guideIdx, 0, n, gi); (*env)->GetIntArrayRegion(env, targetOffsets, 0, n, off);
    /* two passes, one local reference alive at a time: sizes, then one copy of every target into a single pool (no per-array pinning) */
    for (jsize t = 0; t < n; ++t) {
      jbyteArray a = (jbyteArray)(*env)->GetObjectArrayElement(env, targets, t);
      const jsize len = a ? (*env)->GetArrayLength(env, a) : 0;
      if (a) (*env)->DeleteLocalRef(env, a);
      at[t + 1] = at[t] + (size_t)len;
    }
    pool = (uint8_t*)malloc(at[n] ? at[n] : 1);
    if (pool) {
      for (jsize t = 0; t < n; ++t) {
        const jsize len = (jsize)(at[t + 1] - at[t]);
        jbyteArray a = (jbyteArray)(*env)->GetObjectArrayElement(env, targets, t);
        if (a && len > 0) (*env)->GetByteArrayRegion(env, a, 0, len, (jbyte*)(pool + at[t]));
        if (a) (*env)->DeleteLocalRef(env, a);
        tasks[t].guide_idx = gi[t]; tasks[t].target_offset = off[t]; tasks[t].length = (int32_t)len; tasks[t].bases = pool + at[t];
      }
      rc = calitas_align_targets(PTR(calitas_engine, engine), (int32_t)gp.n, gp.g, (int64_t)n, tasks, &lim, best ? 1 : 0, &hs);
    }
  }
  free(tasks); free(gi); free(off); free(at); free(pool);
  free_guides(&gp);
  if (rc) throw_for(env, rc, "calitas_align_targets failed"); else out = wrap_hits(env, hs, outHandle);
  return out;
}

/* ---- aligner.alignToRef / alignToRefBest for many (guide, contig, start, length) regions (AlignToReference.scala:116-134) ---------------- */
JNIEXPORT jobject JNICALL JFN(alignRegions)(JNIEnv* env, jclass cls, jlong engine, jlong ref, jobjectArray guides, jobjectArray auxPams, jintArray guideIdx,
                                            jintArray contigIdx, jlongArray starts, jintArray lengths, jintArray limits, jboolean best, jlongArray outHandle) {
  (void)cls;
  calitas_limits lim; guide_pack gp; calitas_hitset* hs = NULL; jobject out = NULL;
  if (!read_limits(env, limits, &lim)) { throw_for(env, CALITAS_EINVAL, "limits must hold 5 values"); return NULL; }
  if (!pack_guides(env, guides, auxPams, &gp)) { free_guides(&gp); throw_for(env, CALITAS_EINVAL, "bad guide arguments"); return NULL; }
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
  free_guides(&gp);
  if (rc) throw_for(env, rc, "calitas_align_regions failed"); else out = wrap_hits(env, hs, outHandle);
  return out;
}

/* ---- the same over several engines (one per GPU), ONE merged table back (calitas_search_sharded) --------------------------------------------------- */
JNIEXPORT jobject JNICALL JFN(searchSharded)(JNIEnv* env, jclass cls, jlongArray engines, jlongArray refs, jobjectArray guides, jobjectArray auxPams, jintArray limits,
                                             jint windowSize, jstring chrom, jlongArray outHandle) {
  (void)cls;
  calitas_limits lim; guide_pack gp; calitas_hitset* hs = NULL; jobject out = NULL;
  if (!read_limits(env, limits, &lim)) { throw_for(env, CALITAS_EINVAL, "limits must hold 5 values"); return NULL; }
  const jsize ne = engines ? (*env)->GetArrayLength(env, engines) : 0;
  if (ne <= 0 || !refs || (*env)->GetArrayLength(env, refs) != ne) { throw_for(env, CALITAS_EINVAL, "engines and refs must have the same, positive length"); return NULL; }
  if (!pack_guides(env, guides, auxPams, &gp)) { free_guides(&gp); throw_for(env, CALITAS_EINVAL, "bad guide arguments"); return NULL; }
  jlong* eh = (jlong*)calloc((size_t)ne, sizeof *eh); jlong* rh = (jlong*)calloc((size_t)ne, sizeof *rh);
  calitas_engine** ep = (calitas_engine**)calloc((size_t)ne, sizeof *ep); const calitas_reference** rp = (const calitas_reference**)calloc((size_t)ne, sizeof *rp);
  int rc = CALITAS_ESTATE;
  if (eh && rh && ep && rp) {
    (*env)->GetLongArrayRegion(env, engines, 0, ne, eh); (*env)->GetLongArrayRegion(env, refs, 0, ne, rh);
    for (jsize s = 0; s < ne; ++s) { ep[s] = PTR(calitas_engine, eh[s]); rp[s] = PTR(const calitas_reference, rh[s]); }
    char* chrom_c = dup_jstring(env, chrom);
    rc = calitas_search_sharded((int32_t)ne, ep, rp, (int32_t)gp.n, gp.g, &lim, windowSize, chrom_c, &hs);
    free(chrom_c);
  }
  free(eh); free(rh); free(ep); free((void*)rp);
  free_guides(&gp);
  if (rc) throw_for(env, rc, "calitas_search_sharded failed"); else out = wrap_hits(env, hs, outHandle);
  return out;
}

JNIEXPORT jint JNICALL JFN(hitsetStride)(JNIEnv* env, jclass cls, jlong handle) { (void)env; (void)cls; return (jint)calitas_hitset_stride(PTR(calitas_hitset, handle)); }
JNIEXPORT void JNICALL JFN(hitsetFree)(JNIEnv* env, jclass cls, jlong handle) { (void)env; (void)cls; calitas_hitset_free(PTR(calitas_hitset, handle)); }
