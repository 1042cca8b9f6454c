package com.editasmedicine.aligner.b200

import java.nio.{ByteBuffer, ByteOrder}

import com.editasmedicine.aligner.GuideAlignment
import com.editasmedicine.aligner.SequentialGuideAligner.Guide
import com.fulcrumgenomics.alignment.Cigar
import com.fulcrumgenomics.util.Sequences

/** Decodes the packed hit records (include/calitas_b200.h) the engine returns into the reference's GuideAlignment.
  *
  * A result set holds records of one size, `stride` = calitas_hitset_stride(): 32 bytes (calitas_hit, up to 48 alignment columns) or
  * 64 bytes (calitas_hit_wide, up to 176).  Both start with the same 20-byte header (little endian):
  *   0 start_offset, 4 task_idx, 8 score (Int);
  *   12 where: bits 0-12 guide_idx, bits 13-30 contig_idx + 1 (0 = none), bit 31 strand ('-' = 1);
  *   16 shape: bits 0-7 n_ops, 8-15 end_offset - start_offset, 16-21 guide_start_offset - start_offset,
  *             22-27 end_offset - guide_end_offset, 28-31 pam_idx + 1;
  *   20.. ops, 2 bits per alignment column in guide orientation: 0 '=', 1 'X', 2 'I' (guide base opposite a genome gap),
  *   3 'D' (genome base opposite a guide gap).
  *
  * The padded strings are rebuilt exactly as fgbio's Alignment.paddedString(gapChar = '~') + SequentialGuideAligner.toGuideAlignment
  * (SequentialGuideAligner.scala:505-524) produce them: the C++ host code does the same in render_hit_fix (calitas_b200/csrc/cal_host.cpp),
  * which the parity tests compare with the oracle column by column.  Uncompiled here (no scalac in this image).
  */
object HitDecoder {
  val HeaderBytes = 20
  private val OpChars = Array('=', 'X', 'I', 'D')

  final case class Raw(guideIdx: Int, pamIdx: Int, contigIdx: Int, taskIdx: Int, startOffset: Int, endOffset: Int,
                       guideStartOffset: Int, guideEndOffset: Int, score: Int, strand: Char, ops: Array[Int])

  def count(buf: ByteBuffer, stride: Int): Int = buf.capacity / stride

  def raw(buf: ByteBuffer, i: Int, stride: Int): Raw = {
    val b = buf.duplicate().order(ByteOrder.LITTLE_ENDIAN)
    val o = i * stride
    val start = b.getInt(o); val where = b.getInt(o + 12); val shape = b.getInt(o + 16)
    val nOps  = shape & 0xFF
    val end   = start + ((shape >>> 8) & 0xFF)
    val ops   = Array.tabulate(nOps)(k => (b.getInt(o + HeaderBytes + 4 * (k >> 4)) >>> ((k & 15) * 2)) & 3)
    Raw(guideIdx = where & 0x1FFF, pamIdx = (shape >>> 28) - 1, contigIdx = ((where >>> 13) & 0x3FFFF) - 1, taskIdx = b.getInt(o + 4),
        startOffset = start, endOffset = end, guideStartOffset = start + ((shape >>> 16) & 0x3F), guideEndOffset = end - ((shape >>> 22) & 0x3F),
        score = b.getInt(o + 8), strand = if ((where >>> 31) != 0) '-' else '+', ops = ops)
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
