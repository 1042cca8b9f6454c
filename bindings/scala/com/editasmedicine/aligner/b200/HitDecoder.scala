package com.editasmedicine.aligner.b200

import java.nio.{ByteBuffer, ByteOrder}

import com.editasmedicine.aligner.GuideAlignment
import com.editasmedicine.aligner.SequentialGuideAligner.Guide
import com.fulcrumgenomics.alignment.Cigar
import com.fulcrumgenomics.util.Sequences

/** Decodes the 72-byte `calitas_hit` records (include/calitas_b200.h) the engine returns into the reference's GuideAlignment.
  *
  * Record layout (little endian): 0 guide_idx, 4 pam_idx, 8 contig_idx, 12 task_idx, 16 start_offset, 20 end_offset,
  * 24 guide_start_offset, 28 guide_end_offset, 32 score (all Int); 36 strand ('+'/'-'), 37 n_ops, 38 gap_bases, 39 edits (bytes);
  * 40.. ops, 2 bits per alignment column in guide orientation: 0 '=', 1 'X', 2 'I' (guide base opposite a genome gap),
  * 3 'D' (genome base opposite a guide gap).
  *
  * The padded strings are rebuilt exactly as fgbio's Alignment.paddedString(gapChar = '~') + SequentialGuideAligner.toGuideAlignment
  * (SequentialGuideAligner.scala:505-524) produce them: the C++ host code does the same in render_hit_fix (calitas_b200/csrc/cal_host.cpp),
  * which the parity tests compare with the oracle column by column.  Uncompiled here (no scalac in this image).
  */
object HitDecoder {
  val RecordBytes = 72
  private val OpChars = Array('=', 'X', 'I', 'D')

  final case class Raw(guideIdx: Int, pamIdx: Int, contigIdx: Int, taskIdx: Int, startOffset: Int, endOffset: Int,
                       guideStartOffset: Int, guideEndOffset: Int, score: Int, strand: Char, ops: Array[Int])

  def count(buf: ByteBuffer): Int = buf.capacity / RecordBytes

  def raw(buf: ByteBuffer, i: Int): Raw = {
    val b = buf.duplicate().order(ByteOrder.LITTLE_ENDIAN)
    val o = i * RecordBytes
    val nOps = b.get(o + 37) & 0xff
    val ops  = Array.tabulate(nOps)(k => (b.getInt(o + 40 + 4 * (k >> 4)) >>> ((k & 15) * 2)) & 3)
    Raw(b.getInt(o), b.getInt(o + 4), b.getInt(o + 8), b.getInt(o + 12), b.getInt(o + 16), b.getInt(o + 20),
        b.getInt(o + 24), b.getInt(o + 28), b.getInt(o + 32), b.get(o + 36).toChar, ops)
  }

  /** @param guide  the guide the hit belongs to (guides(raw.guideIdx) of the call)
    * @param pams   that guide's PAMs in call order: the primary PAM first, then the auxiliary PAMs (all lower case)
    * @param chrom  name of the contig / target
    * @param fetch  (start, end) => forward-strand bases [start, end) of the target, already upper-cased where the reference
    *               upper-cases its windows (SearchReference.scala:67)
    */
  def decode(r: Raw, guide: Guide, pams: IndexedSeq[String], chrom: String, fetch: (Int, Int) => Array[Byte]): GuideAlignment = {
    val pam       = if (r.pamIdx >= 0) pams(r.pamIdx) else ""
    val guideText = if (guide.pamIsFivePrime) pam + guide.guide else guide.guide + pam     // guide + PAM in guide orientation
    val fwd       = new String(fetch(r.startOffset, r.endOffset))
    val target    = if (r.strand == '-') Sequences.revcomp(fwd) else fwd
    val pg = new StringBuilder; val pa = new StringBuilder; val pt = new StringBuilder
    var qi = 0; var ti = 0
    r.ops.foreach { op =>
      if (op != 3) { pg.append(guideText.charAt(qi)); qi += 1 } else pg.append('-')
      if (op != 2) { pt.append(target.charAt(ti)); ti += 1 } else pt.append('-')
      pa.append(if (op == 0) '|' else if (op == 1) '.' else '~')
    }
    require(qi == guideText.length && ti == target.length, "hit ops do not cover the guide/target")
    val cigar = {                                                                           // run-length encoding of = X I D
      val sb = new StringBuilder; var k = 0
      while (k < r.ops.length) { var j = k; while (j < r.ops.length && r.ops(j) == r.ops(k)) j += 1; sb.append(j - k).append(OpChars(r.ops(k))); k = j }
      Cigar(sb.toString)
    }
    GuideAlignment(guide = guideText, chrom = chrom, startOffset = r.startOffset, endOffset = r.endOffset,
                   guideStartOffset = r.guideStartOffset, guideEndOffset = r.guideEndOffset, strand = r.strand, score = r.score,
                   cigar = cigar, paddedGuide = pg.toString, paddedAlignment = pa.toString, paddedTarget = pt.toString)
  }
}
