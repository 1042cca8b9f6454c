package com.editasmedicine.aligner.b200

import java.nio.ByteBuffer

/** JNI entry points of libcalitas_b200_jni.so (bindings/jni/calitas_b200_jni.c), one per function of include/calitas_b200.h.
  *
  * Uncompiled here (no JDK / scalac in this image); tests/test_bindings.py checks that every @native method below has its
  * Java_com_editasmedicine_aligner_b200_Native_* export in the shim and vice versa.
  *
  * Handles are opaque pointers carried as Long.  A failing call throws IllegalArgumentException where the reference's `require`
  * would (CALITAS_EINVAL) and IllegalStateException otherwise.  Hit records come back as a direct ByteBuffer over the engine's
  * pinned host memory: little-endian records of hitsetStride(handle) = 32 or 64 bytes (HitDecoder), valid until hitsetFree(handle).
  */
object Native {
  System.loadLibrary("calitas_b200_jni")

  /** costs = (guideMismatchNetCost, genomeGapNetCost, guideGapNetCost, pamMismatchNetCost): `new SequentialGuideAligner(...)`. */
  @native def engineCreate(device: Int, costs: Array[Int]): Long
  @native def engineDestroy(engine: Long): Unit

  /** Contig-range sharding for `nShards` GPUs: fills the four arrays (length = lengths.length) for `shard`. */
  @native def shardPlan(lengths: Array[Long], shard: Int, nShards: Int, halo: Long,
                        ownBegin: Array[Long], ownEnd: Array[Long], haveBegin: Array[Long], haveEnd: Array[Long]): Unit

  /** bases(c) = direct ByteBuffer with contig bases [haveBegin(c), haveEnd(c)); null range arrays = load and own everything. */
  @native def referenceLoad(engine: Long, names: Array[String], lengths: Array[Long], bases: Array[ByteBuffer],
                            haveBegin: Array[Long], haveEnd: Array[Long], ownBegin: Array[Long], ownEnd: Array[Long]): Long
  @native def referenceFree(engine: Long, ref: Long): Unit

  /** limits = (maxGuideDiffs, maxPamMismatches, maxGapsBetweenGuideAndPam, maxTotalDiffs (< 0: d + g + p), maxOverlap). */
  @native def search(engine: Long, ref: Long, guides: Array[String], auxPams: Array[Array[String]], limits: Array[Int],
                     windowSize: Int, chrom: String, dedup: Boolean, outHandle: Array[Long]): ByteBuffer

  /** One SequentialGuideAligner.align(guide(guideIdx(t)), targets(t), targetOffset = targetOffsets(t)) per task t. */
  @native def alignTargets(engine: Long, guides: Array[String], auxPams: Array[Array[String]], guideIdx: Array[Int],
                           targets: Array[Array[Byte]], targetOffsets: Array[Int], limits: Array[Int], best: Boolean,
                           outHandle: Array[Long]): ByteBuffer

  /** One alignToRef / alignToRefBest per task: contig bases [starts(t), starts(t) + lengths(t)). */
  @native def alignRegions(engine: Long, ref: Long, guides: Array[String], auxPams: Array[Array[String]], guideIdx: Array[Int],
                           contigIdx: Array[Int], starts: Array[Long], lengths: Array[Int], limits: Array[Int], best: Boolean,
                           outHandle: Array[Long]): ByteBuffer

  /** The same over several engines (engine s holds shard s of shardPlan(.., s, engines.length, ..) in refs(s)): the engines run concurrently on native
    * threads and ONE table comes back, merged per guide at the shard cuts (calitas_search_sharded).  Needs maxOverlap >= 1. */
  @native def searchSharded(engines: Array[Long], refs: Array[Long], guides: Array[String], auxPams: Array[Array[String]], limits: Array[Int],
                            windowSize: Int, chrom: String, outHandle: Array[Long]): ByteBuffer

  /** Bytes per record of a result set: 32 (calitas_hit) or 64 (calitas_hit_wide). */
  @native def hitsetStride(handle: Long): Int
  @native def hitsetFree(handle: Long): Unit
}
