package com.editasmedicine.aligner.b200

import java.nio.ByteBuffer

import com.editasmedicine.aligner.GuideAlignment
import com.editasmedicine.aligner.SequentialGuideAligner.Guide

/** What SearchReference.execute calls instead of its window loop (SearchReference.scala:527-564) and of removeOverlaps + ReferenceHit.sort
  * (:641-648) when no VCF is given: one engine per GPU, the genome sharded by contig range, every guide of the run in one call.
  * The alignments come back de-duplicated and in ReferenceHit.sort order per guide, in ONE table whatever the number of engines (the shard
  * lists are merged at the cuts on the native side); the caller feeds them to `hitBuilder.build` as before.
  * Uncompiled here (no scalac in this image).
  *
  * @param devices     CUDA device ids, one engine each
  * @param costs       (guideMismatchNetCost, genomeGapNetCost, guideGapNetCost, pamMismatchNetCost)
  * @param contigNames / contigBases  the reference as SearchReference reads it: bases exactly as in the FASTA, in direct buffers
  */
final class B200Search(devices: Seq[Int], costs: Array[Int], contigNames: Array[String], contigBases: Array[ByteBuffer]) extends AutoCloseable {
  private val lengths = contigBases.map(_.capacity.toLong)
  private val engines = devices.map(d => Native.engineCreate(d, costs)).toArray
  private val refs    = engines.indices.map { s =>
    if (engines.length == 1) Native.referenceLoad(engines(s), contigNames, lengths, contigBases, null, null, null, null)
    else {
      val n = lengths.length
      val (ob, oe, hb, he) = (new Array[Long](n), new Array[Long](n), new Array[Long](n), new Array[Long](n))
      Native.shardPlan(lengths, s, engines.length, 4L * 1000, ob, oe, hb, he)          // halo: 4 x the default window size
      val slices = contigBases.indices.map { c => val b = contigBases(c).duplicate(); b.position(hb(c).toInt); b.limit(he(c).toInt); b.slice() }.toArray
      Native.referenceLoad(engines(s), contigNames, lengths, slices, hb, he, ob, oe)
    }
  }.toArray

  /** guides(i) = (guide with its primary PAM, auxiliary PAMs).  Returns, per guide, its alignments in ReferenceHit.sort order. */
  def search(guides: IndexedSeq[(String, Seq[String])], maxGuideDiffs: Int, maxPamMismatches: Int, maxGapsBetweenGuideAndPam: Int,
             maxTotalDiffs: Int, maxOverlap: Int, windowSize: Int, chrom: Option[String]): IndexedSeq[IndexedSeq[GuideAlignment]] = {
    val parsed  = guides.map { case (g, aux) => Guide(g, aux) }
    val pams    = guides.map { case (g, aux) => (g.filter(_.isLower) +: aux).filter(_.nonEmpty).toIndexedSeq }
    val limits  = Array(maxGuideDiffs, maxPamMismatches, maxGapsBetweenGuideAndPam, maxTotalDiffs, maxOverlap)
    val out     = Array.fill(guides.length)(IndexedSeq.newBuilder[GuideAlignment])
    require(maxOverlap >= 1 || engines.length == 1,
      "with --max-overlap <= 0 removeOverlaps reaches across the whole contig and cannot run per shard: use one engine (or search with dedup = false and run removeOverlaps over the gathered hits)")
    // The shards run concurrently inside the native call.  Their per-guide lists are NOT simply concatenated: consecutive windows overlap by
    // guide length + d + g - 1 bases, so the last window of one shard and the first of the next can report hits whose starts interleave; the native
    // side merges each guide's lists by the ReferenceHit.sort key where they meet (calitas_search_sharded, merge_guide_segments) and returns ONE table.
    val handle = new Array[Long](1)
    val buf    = Native.searchSharded(engines, refs, guides.map(_._1).toArray, guides.map(_._2.toArray).toArray, limits, windowSize, chrom.orNull, handle)
    val stride = Native.hitsetStride(handle(0))
    var i = 0
    while (i < HitDecoder.count(buf, stride)) {
      val r = HitDecoder.raw(buf, i, stride)
      val fetch = (start: Int, end: Int) => { val a = new Array[Byte](end - start); val b = contigBases(r.contigIdx).duplicate(); b.position(start); b.get(a); new String(a).toUpperCase.getBytes }
      out(r.guideIdx) += HitDecoder.decode(r, parsed(r.guideIdx), pams(r.guideIdx), contigNames(r.contigIdx), fetch)
      i += 1
    }
    Native.hitsetFree(handle(0))
    out.map(_.result()).toIndexedSeq
  }

  override def close(): Unit = engines.indices.foreach { s => Native.referenceFree(engines(s), refs(s)); Native.engineDestroy(engines(s)) }
}
