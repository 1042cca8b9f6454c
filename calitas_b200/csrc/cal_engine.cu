// cal_engine.cu — kernels, engine orchestration and the C ABI of include/calitas_b200.h.
//
// Data path of one SearchReference batch (SearchReference.scala:527-564, 641-648) on one GPU:
//   k_scan_tiled   every owned reference window x both strands x every guide of the chunk: bit-parallel semi-global edit
//                  distance of the protospacer (one 32-bit word per scan, tile of 64 windows staged in shared memory);
//                  emits the candidate end columns (a lossless superset of fgbio's `score >= minScore` columns)
//   sort           candidate keys (guide, window, strand, end column) -> fgbio's emission order
//   k_align        per candidate: exact 3-matrix glocal DP on the band that can hold a co-optimal alignment, traceback with
//                  the oracle's tie-breaks, PAM extension, hit record
//   k_canon        per (guide, window, strand): the reference's sort + greedy overlap filter
//   dedup          removeOverlaps + ReferenceHit.sort: radix sorts + per-(guide, contig, strand) sweep
// AlignToReference / variant windows use k_scan_explicit (one thread per window and strand) and the same tail.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <thread>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "cal_dev.h"
#include "cal_core.cuh"
#include "cal_host.h"

namespace cal {

// ------------------------------------------------------------------------------------------------------------------------------------
// device-side descriptors
// ------------------------------------------------------------------------------------------------------------------------------------
struct ContigDev { int64_t len; int64_t nib_base; /* nibble index of contig base 0 */ int64_t win_base; /* global index of window 0 */
                   int64_t own_lo, own_hi; /* global ids of the windows this engine owns; windows outside are halo (processed, never reported) */ };
struct Tile { int32_t contig; int32_t nwin; int64_t first_k; };
struct ExplicitWindow { int64_t nib_start; int32_t len; int32_t target_offset; int32_t guide_idx; int32_t contig_idx; int32_t task_id /* task_idx of the hits */; int32_t owned /* 0: halo, hits take part in removeOverlaps and are never reported */; };

const int ALIGN_KB = 6;          // k_align keeps a 2*6+1-diagonal DP band in registers when the candidate threshold allows
const int SCAN_SMEM_LIMIT = 200 * 1024;   // dynamic shared memory a k_scan_tiled CTA may ask for
const int TILE_WINDOWS = 64;      // windows per scan tile at the default window size; halved until the tile fits shared memory for larger -w
const int HALO_WINDOWS = 2;      // windows processed beyond each interior shard cut so that removeOverlaps sees both sides of the cut
#ifndef CAL_SCAN_NG
#define CAL_SCAN_NG 2
#endif
const int SCAN_NG = CAL_SCAN_NG;     // guides a scan thread advances together (independent dependency chains sharing the base-code extraction)
const int KEY_COL_BITS_MAX = 17;
const uint32_t MAX_WINDOW_LEN = (1u << KEY_COL_BITS_MAX) - 1;
const int MAX_GUIDES_PER_CALL = CALITAS_MAX_GUIDES;     // 13 bits of the hit record

// Candidate key = guide | window | strand | end column, each field only as wide as the call needs (<= 14 + 32 + 1 + 17 bits), so that the
// LSD radix sort of the keys runs over `bits` bits: 5 eight-bit passes instead of 8 for 100 guides x hg38 at -w 1000.  Every pass is a
// dependent launch that has to find room on SMs held by the next chunk's scan kernel.
struct KeyLayout { int32_t col_bits, win_shift, guide_shift, bits; };
CAL_HD int bit_length(uint64_t v) { int b = 0; while (v) { ++b; v >>= 1; } return b; }
inline KeyLayout make_key_layout(uint32_t max_col, uint64_t n_windows, uint32_t n_guides) {
  KeyLayout L; L.col_bits = bit_length(max_col); if (L.col_bits < 1) L.col_bits = 1;
  L.win_shift = L.col_bits + 1;
  int wb = bit_length(n_windows > 0 ? n_windows - 1 : 0); if (wb < 1) wb = 1;
  L.guide_shift = L.win_shift + wb;
  L.bits = L.guide_shift + bit_length(n_guides > 0 ? n_guides - 1 : 0);
  return L;
}
CAL_HD uint64_t make_key(const KeyLayout& L, uint32_t guide, uint32_t window, uint32_t strandbit, uint32_t col) {
  return ((uint64_t)guide << L.guide_shift) | ((uint64_t)window << L.win_shift) | ((uint64_t)strandbit << L.col_bits) | col;
}
CAL_HD int32_t key_col(const KeyLayout& L, uint64_t key) { return (int32_t)(key & ((1ull << L.col_bits) - 1)); }
CAL_HD uint32_t key_strandbit(const KeyLayout& L, uint64_t key) { return (uint32_t)(key >> L.col_bits) & 1u; }
CAL_HD uint32_t key_window(const KeyLayout& L, uint64_t key) { return (uint32_t)((key >> L.win_shift) & ((1ull << (L.guide_shift - L.win_shift)) - 1)); }
CAL_HD int32_t key_guide(const KeyLayout& L, uint64_t key) { return (int32_t)(key >> L.guide_shift); }
CAL_HD uint64_t key_group(const KeyLayout& L, uint64_t key) { return key >> L.col_bits; }      // (guide, window, strand)
CAL_HD uint32_t nibble_at(const uint32_t* words, int64_t idx) { return (words[idx >> 3] >> ((uint32_t)(idx & 7) * 4)) & 15u; }

// ------------------------------------------------------------------------------------------------------------------------------------
// k_pack: raw bytes -> 4-bit target codes.  Streaming, HBM-bound: 1 B read + 0.5 B written per base.
// ------------------------------------------------------------------------------------------------------------------------------------
CAL_KERNEL __launch_bounds__(256) k_pack(const uint8_t* __restrict__ raw, uint32_t* __restrict__ nib, int64_t n_words) {
  CAL_SHARED_DYN(uint8_t, lut);
  CAL_PHASE(0) { for (int i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = (uint8_t)target_code((uint8_t)i); }
  __syncthreads();
  CAL_PHASE(1) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < n_words; w += stride) {
      const uint64_t v = reinterpret_cast<const uint64_t*>(raw)[w];
      uint32_t out = 0;
#pragma unroll
      for (int k = 0; k < 8; ++k) out |= (uint32_t)lut[(v >> (8 * k)) & 0xFF] << (4 * k);
      nib[w] = out;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------------------------
// k_scan_tiled: the hot kernel
// ------------------------------------------------------------------------------------------------------------------------------------
struct ScanArgs {
  const uint32_t* nib; const ContigDev* contigs; const Tile* tiles; const GuideSpec* specs;
  int32_t g_begin, g_end, window_size, step, min_len, scan_slots, tile_windows;
  uint64_t* cand; unsigned long long* cand_count; unsigned long long cand_cap; KeyLayout key;
};

struct Emitter {
  uint64_t* cand; unsigned long long* count; unsigned long long cap; uint64_t key_base;
  CAL_D void operator()(int32_t col) const {
    unsigned long long i = atomicAdd(count, 1ull);
    if (i < cap) cand[i] = key_base | (uint32_t)col;
  }
};

// One guide of a thread's scan: its match-mask table for the thread's direction (shared memory), threshold, candidate key base.
struct ScanGuide { const uint32_t* peq; int32_t lp, k_edits; uint64_t key_base; };
#ifdef CAL_HOSTSIM
inline int __vimin3_s32(int a, int b, int c) { int m = a < b ? a : b; return m < c ? m : c; }
#endif

// Scans window [tb + rs, tb + re) of the shared-memory byte tile for NG guides at once (independent Myers chains that share the base fetch
// and give the scheduler instruction-level parallelism).  The tile holds one byte per base, already scaled to the byte offset of that base's
// entry in a 16-word mask table (code * 4): a column costs one LDS.U8 and one add (FMA pipe) before the table lookups, nothing on the ALU pipe
// the kernel is bound by (the packed-nibble form needed a shift and a funnel shift per column and ran 7 % slower).  DIR 0: left to right;
// DIR 1: right to left, the tables then hold the complemented masks.  Column p = 1.. in scan order.  Blocks of 8 columns run branch-free:
// the running minimum of the distance decides, once per block, whether the (rare) per-column emission replay is needed.  Tables of a
// thread's guides are adjacent (32 words apart).
template <int DIR, int NG>
CAL_D void scan_window(const uint8_t* tb, int32_t rs, int32_t re, int32_t c_lo, int32_t c_hi, const ScanGuide* sg, uint64_t* cand, unsigned long long* count, unsigned long long cap) {
  // Columns [c_lo, c_hi) (0-based, scan order) of the window are this thread's; when c_lo > 0 the chains are warmed up over the lp + k_edits
  // columns before c_lo from the fresh state: an alignment with <= k_edits edits that ends at or after c_lo starts inside that stretch, and a
  // later start can only raise distances, so "distance <= k_edits" comes out exactly as in a scan from the window's first column.
  const uint8_t* t0 = reinterpret_cast<const uint8_t*>(sg[0].peq);
  MyersState st[NG];
#pragma unroll
  for (int j = 0; j < NG; ++j) myers_init(st[j], sg[j].lp);
  const int32_t n = c_hi;
  const uint8_t* p = DIR == 0 ? tb + rs : tb + (re - 1);
#define CAL_EQ(J, B) (*reinterpret_cast<const uint32_t*>(t0 + 128 * (J) + (B)))
#define CAL_STEP1(J, STATE, B, COL) { myers_step(STATE, CAL_EQ(J, B)); if (STATE.score <= sg[J].k_edits) { Emitter em{ cand, count, cap, sg[J].key_base }; em(COL); } }
  int32_t c = c_lo;
  if (c_lo > 0) {
    int32_t warm = 0;
#pragma unroll
    for (int j = 0; j < NG; ++j) { const int32_t w = sg[j].lp + sg[j].k_edits; warm = w > warm ? w : warm; }
    for (int32_t w = c_lo - warm > 0 ? c_lo - warm : 0; w < c_lo; ++w) { const uint32_t b = DIR == 0 ? p[w] : p[-w];
#pragma unroll
      for (int j = 0; j < NG; ++j) myers_step(st[j], CAL_EQ(j, b)); }
  }
  // (An in-loop test of every column instead of the block minimum + replay was measured at thresholds of 6 edits, where a third of the warp-blocks
  //  replay: 71.4 instead of 63.5 ms per 16-guide scan, and 449 instead of 395 ms per default step; the replay stays.)
  for (; c + 8 <= n; c += 8) {
    const uint8_t* q = DIR == 0 ? p + c : p - c;
    MyersState save[NG]; int32_t mn[NG], prev[NG];
#pragma unroll
    for (int j = 0; j < NG; ++j) { save[j] = st[j]; mn[j] = 0x7FFFFFFF; prev[j] = 0x7FFFFFFF; }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const uint32_t b = DIR == 0 ? q[k] : q[-k];
#pragma unroll
      for (int j = 0; j < NG; ++j) {
        myers_step(st[j], CAL_EQ(j, b));
        if (k & 1) mn[j] = __vimin3_s32(mn[j], prev[j], st[j].score); else prev[j] = st[j].score;      // one 3-input min per two columns
      }
    }
#pragma unroll
    for (int j = 0; j < NG; ++j) if (mn[j] <= sg[j].k_edits) { MyersState t = save[j]; for (int k = 0; k < 8; ++k) { const uint32_t b = DIR == 0 ? q[k] : q[-k]; CAL_STEP1(j, t, b, c + k + 1) } }
  }
  for (; c < n; ++c) { const uint32_t b = DIR == 0 ? p[c] : p[-c];
#pragma unroll
    for (int j = 0; j < NG; ++j) CAL_STEP1(j, st[j], b, c + 1) }
#undef CAL_STEP1
#undef CAL_EQ
}

// Block = TILE_WINDOWS windows x 2 directions x 4 (guide slot, window part) pairs = 512 threads; a thread owns one window part, one
// direction and every scan_slots-th pair of guides of the chunk.
CAL_KERNEL CAL_MAXNREG(56) k_scan_tiled(ScanArgs a) {
  CAL_SHARED_DYN(uint32_t, smem);
  const int ng = a.g_end - a.g_begin;
  uint32_t* s_peq = smem;                       // ng * 32 words
  int32_t* s_meta = (int32_t*)(smem + ng * 32); // ng * 4: lp, k_edits, five_prime, pad
  uint32_t* s_tile = smem + ng * 36 + ((((a.tile_windows - 1) * a.step + a.window_size + 16 + 15) & ~15) >> 2);      // packed words are staged behind the byte tile (which is word-aligned with them: + 2 words)
  const Tile tile = a.tiles[blockIdx.x];
  const ContigDev ctg = a.contigs[tile.contig];
  const int64_t tile_start = tile.first_k * (int64_t)a.step;
  int64_t tile_end = (tile.first_k + tile.nwin - 1) * (int64_t)a.step + a.window_size; if (tile_end > ctg.len) tile_end = ctg.len;
  const int64_t nib0 = ctg.nib_base + tile_start;
  const int64_t word0 = nib0 >> 3;
  const int32_t n_words = (int32_t)(((ctg.nib_base + tile_end + 7) >> 3) - word0);
#ifndef CAL_HOSTSIM
  // The tile's packed bases (~31 KB) come in with one TMA bulk copy (cp.async.bulk, completion on an mbarrier) issued by one thread, while
  // all threads stage the guides' mask tables: no per-thread LDG/STS loop, no address arithmetic on the ALU pipe this kernel is bound by.
  __shared__ __align__(8) unsigned long long s_mbar;
  {
    const uint32_t mbar = (uint32_t)__cvta_generic_to_shared(&s_mbar);
    const int64_t word0a = word0 & ~3ll;                                // 16-byte aligned source; the tile starts `lead` words in
    const int32_t lead = (int32_t)(word0 - word0a);
    const uint32_t bytes = (uint32_t)((((uint32_t)(n_words + lead)) * 4u + 15u) & ~15u);
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(mbar) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mbar), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   :: "r"((uint32_t)__cvta_generic_to_shared(s_tile)), "l"(a.nib + word0a), "r"(bytes), "r"(mbar) : "memory");
    }
    for (int i = threadIdx.x; i < ng * 32; i += blockDim.x) s_peq[i] = a.specs[a.g_begin + (i >> 5)].peq[(i >> 4) & 1][i & 15];
    for (int i = threadIdx.x; i < ng; i += blockDim.x) {
      const GuideSpec& sp = a.specs[a.g_begin + i];
      s_meta[4 * i] = sp.lp; s_meta[4 * i + 1] = sp.k_edits; s_meta[4 * i + 2] = sp.five_prime; s_meta[4 * i + 3] = 0;
    }
    asm volatile("{\n\t.reg .pred p;\n\tTILE_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0, 0x989680;\n\t@p bra TILE_DONE;\n\tbra TILE_WAIT;\n\tTILE_DONE:\n\t}" :: "r"(mbar) : "memory");
    s_tile += lead;
  }
  __syncthreads();
#else
  CAL_PHASE(0) {
    for (int i = threadIdx.x; i < n_words; i += blockDim.x) s_tile[i] = __ldg(a.nib + word0 + i);
    for (int i = threadIdx.x; i < ng * 32; i += blockDim.x) s_peq[i] = a.specs[a.g_begin + (i >> 5)].peq[(i >> 4) & 1][i & 15];
    for (int i = threadIdx.x; i < ng; i += blockDim.x) {
      const GuideSpec& s = a.specs[a.g_begin + i];
      s_meta[4 * i] = s.lp; s_meta[4 * i + 1] = s.k_edits; s_meta[4 * i + 2] = s.five_prime; s_meta[4 * i + 3] = 0;
    }
  }
  __syncthreads();
#endif
  // expand the packed words into the byte tile: byte = code * 4.  The byte tile is laid out word-aligned with the packed words (byte 8 i + q
  // holds nibble q of word i), so a word becomes two aligned 32-bit stores; the tile's first base sits r_off0 bytes in.
  const int32_t r_off0 = (int32_t)(nib0 & 7);
  uint8_t* s_bytes = reinterpret_cast<uint8_t*>(smem + ng * 36) + 0;
  CAL_PHASE(1) {
    uint32_t* out = smem + ng * 36;
    for (int i = threadIdx.x; i < n_words; i += blockDim.x) {
      const uint32_t w = s_tile[i];
      const uint32_t lo = (w & 0x0F0F0F0Fu) << 2, hi = ((w >> 4) & 0x0F0F0F0Fu) << 2;      // even / odd nibbles, one per byte, already scaled by 4
#ifndef CAL_HOSTSIM
      out[2 * i] = __byte_perm(lo, hi, 0x5140); out[2 * i + 1] = __byte_perm(lo, hi, 0x7362);
#else
      out[2 * i] = (lo & 0xFFu) | ((hi & 0xFFu) << 8) | ((lo & 0xFF00u) << 8) | ((hi & 0xFF00u) << 16);
      out[2 * i + 1] = ((lo >> 16) & 0xFFu) | (((hi >> 16) & 0xFFu) << 8) | (((lo >> 24) & 0xFFu) << 16) | (((hi >> 24) & 0xFFu) << 24);
#endif
    }
  }
  __syncthreads();
  CAL_PHASE(2) {
    // thread = (window, direction, guide slot, window part): with few guides the block keeps its 512 threads by cutting each window into parts
    const int tw = a.tile_windows, n_slots = a.scan_slots, n_parts = (int)(blockDim.x / (2 * tw)) / n_slots;
    // A warp takes 32 windows of the same parity: consecutive windows are `step` bytes apart in the byte tile, and with the default step (970 =
    // 2 x 485) windows 2m lie 1940 m bytes apart = 485 m words, 485 m mod 32 = 5 m mod 32 is a permutation, so the 32 LDS.U8 of a column
    // hit 32 different banks (consecutive windows in one warp collide pairwise: measured 7x the bank conflicts).
    const int x = threadIdx.x % tw;
    const int kk = tw == 64 ? (((x & 31) << 1) | (x >> 5)) : x;
    const int dir = (threadIdx.x / tw) & 1, sp = threadIdx.x / (2 * tw), slot = sp % n_slots, part = sp / n_slots;
    if (kk >= tile.nwin) return;
    const int64_t ws = (tile.first_k + kk) * (int64_t)a.step;
    int64_t we = ws + a.window_size; if (we > ctg.len) we = ctg.len;
    int32_t rs = r_off0 + (int32_t)(ws - tile_start), re = r_off0 + (int32_t)(we - tile_start);
    while (rs < re && s_bytes[rs] == (CODE_N << 2)) ++rs;            // SearchReference.scala:58-59
    while (rs < re && s_bytes[re - 1] == (CODE_N << 2)) --re;
    const int32_t m = re - rs;
    if (m <= 0 || m < a.min_len) return;                              // SearchReference.scala:536
    const int32_t c_lo = (int32_t)((int64_t)m * part / n_parts), c_hi = (int32_t)((int64_t)m * (part + 1) / n_parts);
    const uint32_t wid = (uint32_t)(ctg.win_base + tile.first_k + kk);
    for (int g = SCAN_NG * slot; g < ng; g += SCAN_NG * n_slots) {
      ScanGuide sg[SCAN_NG];
      const int cnt = ng - g < SCAN_NG ? ng - g : SCAN_NG;
      for (int j = 0; j < cnt; ++j) {
        sg[j].peq = s_peq + (g + j) * 32 + dir * 16; sg[j].lp = s_meta[4 * (g + j)]; sg[j].k_edits = s_meta[4 * (g + j) + 1];
        sg[j].key_base = make_key(a.key, (uint32_t)(a.g_begin + g + j), wid, (uint32_t)(dir ^ s_meta[4 * (g + j) + 2]), 0);
      }
#define CAL_SCAN_CALL(N) { if (dir == 0) scan_window<0, N>(s_bytes, rs, re, c_lo, c_hi, sg, a.cand, a.cand_count, a.cand_cap); else scan_window<1, N>(s_bytes, rs, re, c_lo, c_hi, sg, a.cand, a.cand_count, a.cand_cap); }
      if (cnt == SCAN_NG) CAL_SCAN_CALL(SCAN_NG)
      else if (cnt >= 2) { CAL_SCAN_CALL(2) if (cnt == 3) { sg[0] = sg[2]; CAL_SCAN_CALL(1) } }
      else CAL_SCAN_CALL(1)
#undef CAL_SCAN_CALL
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------------------------
// k_scan_explicit: one thread per (window, direction); windows are short (AlignToReference regions, variant windows)
// ------------------------------------------------------------------------------------------------------------------------------------
struct ScanExplicitArgs {
  const uint32_t* nib; const ExplicitWindow* windows; int64_t n_windows; const GuideSpec* specs;
  uint64_t* cand; unsigned long long* cand_count; unsigned long long cand_cap; KeyLayout key;
};
// One window in one direction, word by word: a packed word is loaded once and its codes are taken from the low (left to right) or the high end (right to
// left: DIR 1 scans the reverse complement with the complemented tables), so a column costs a mask/shift pair instead of an index computation and a load.
CAL_D void scan_words(const uint32_t* nib, int64_t nib_start, int32_t len, int dir, const uint32_t* peq, int lp, int k_edits, const Emitter& emit) {
  MyersState st; myers_init(st, lp);
  int32_t col = 1, left = len;
  if (dir == 0) {
    int64_t wi = nib_start >> 3; int k = (int)(nib_start & 7);
    uint32_t word = __ldg(nib + wi) >> (4 * k);
    while (left > 0) {
      int n = 8 - k; if (n > left) n = left;
      for (int q = 0; q < n; ++q) { const uint32_t c = word & 15u; word >>= 4; myers_step(st, peq[c]); if (st.score <= k_edits) emit(col); ++col; }
      left -= n; k = 0; ++wi; if (left > 0) word = __ldg(nib + wi);
    }
  } else {
    const int64_t last = nib_start + len - 1;
    int64_t wi = last >> 3; int k = (int)(last & 7);
    uint32_t word = __ldg(nib + wi) << (4 * (7 - k));
    while (left > 0) {
      int n = k + 1; if (n > left) n = left;
      for (int q = 0; q < n; ++q) { const uint32_t c = word >> 28; word <<= 4; myers_step(st, peq[c]); if (st.score <= k_edits) emit(col); ++col; }
      left -= n; k = 7; --wi; if (left > 0) word = __ldg(nib + wi);
    }
  }
}
// thread = (window, direction); a block is 128 consecutive windows of one direction.  Window lists are guide-major (variant tasks) or sorted by guide
// in practice, so the mask tables of the block's first window's guide are staged in shared memory; a window of another guide reads its tables from global memory.
CAL_KERNEL __launch_bounds__(128) k_scan_explicit(ScanExplicitArgs a) {
  CAL_SHARED_DYN(uint32_t, smem);                      // 32 words of mask tables (direction, code) + the guide they belong to
  const int64_t tb = (int64_t)blockIdx.x * blockDim.x, t = tb + threadIdx.x;
  CAL_PHASE(0) {
    const int64_t wb = tb >= a.n_windows ? tb - a.n_windows : tb;
    const int32_t g0 = a.windows[wb].guide_idx;
    if (threadIdx.x < 32) smem[threadIdx.x] = a.specs[g0].peq[threadIdx.x >> 4][threadIdx.x & 15];
    if (threadIdx.x == 32) smem[32] = (uint32_t)g0;
  }
  __syncthreads();
  CAL_PHASE(1) {
    if (t >= 2 * a.n_windows) return;
    const int dir = t >= a.n_windows ? 1 : 0;
    const int64_t w = dir ? t - a.n_windows : t;
    const ExplicitWindow ew = a.windows[w];
    if (ew.len <= 0) return;
    const GuideSpec& s = a.specs[ew.guide_idx];
    Emitter emit{ a.cand, a.cand_count, a.cand_cap, make_key(a.key, 0, (uint32_t)w, (uint32_t)(dir ^ s.five_prime), 0) };
    if (ew.guide_idx == (int32_t)smem[32]) scan_words(a.nib, ew.nib_start, ew.len, dir, smem + 16 * dir, s.lp, s.k_edits, emit);
    else scan_words(a.nib, ew.nib_start, ew.len, dir, s.peq[dir], s.lp, s.k_edits, emit);
  }
}

// Best mode (alignBest / alignToRefBest: d = protospacer length) makes EVERY column of both strands a candidate -- the semi-global edit distance of an
// lp-row pattern never exceeds lp -- so the candidate list is written directly, already in (window, strand, column) order: no scan, no sort.
CAL_KERNEL __launch_bounds__(256) k_window_columns(const ExplicitWindow* windows, int64_t n_windows, uint32_t* cnt) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < 2 * n_windows) { const int32_t len = windows[t >> 1].len; cnt[t] = len > 0 ? (uint32_t)len : 0u; }
}
CAL_KERNEL __launch_bounds__(256) k_all_columns(const ExplicitWindow* windows, int64_t n_windows, const uint32_t* off, KeyLayout key, uint64_t* out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 2 * n_windows) return;
  const int32_t len = windows[t >> 1].len;
  const uint64_t base = make_key(key, 0, (uint32_t)(t >> 1), (uint32_t)(t & 1), 0);
  uint64_t* o = out + off[t];
  for (int32_t c = 1; c <= len; ++c) o[c - 1] = base | (uint32_t)c;
}

// ------------------------------------------------------------------------------------------------------------------------------------
// k_align: candidate -> exact DP + traceback + PAM extension -> hit slots
// ------------------------------------------------------------------------------------------------------------------------------------
struct AlignArgs {
  const uint64_t* cand; int64_t n_cand; const GuideSpec* specs; Scores sc; int32_t slots;
  int32_t explicit_mode;
  const uint32_t* nib;
  const ContigDev* contigs; int32_t n_contigs; int32_t window_size, step;     // tiled
  const ExplicitWindow* windows;                                             // explicit (window ids index this array)
  uint32_t* recs; int32_t rw; CKey* ckeys; KeyLayout key;        // packed hit records, rw words each, and one canon key per alignment slot (cal_core.cuh)
  int64_t nib_last_word;                                        // last valid word of `nib` (align_fast clamps its nine loads to it)
  int32_t match4, mis4, up4, left4;                             // align_pair: 4 * match, 4 * mismatch, 4 * target_gap + TG_UP, 4 * query_gap + TG_LEFT
};
struct NibFetch {
  const uint32_t* nib; int64_t first; int32_t m; int dir;
  CAL_D uint32_t operator()(int32_t p) const {       // DP column p (1-based) of the scanned target
    const int64_t idx = first + (dir == 0 ? p - 1 : m - p);
    const uint32_t c = (__ldg(nib + (idx >> 3)) >> ((uint32_t)(idx & 7) * 4)) & 15u;
    return dir == 0 ? c : comp_code(c);
  }
};
// window id -> contig, trimmed [begin, end) in contig coordinates (tiled mode)
CAL_D void locate_window(const uint32_t* nib, const ContigDev* contigs, int32_t n_contigs, int32_t window_size, int32_t step, uint32_t wid,
                         int32_t& contig, int64_t& wb, int64_t& we) {
  int lo = 0, hi = n_contigs - 1;
  while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (contigs[mid].win_base <= (int64_t)wid) lo = mid; else hi = mid - 1; }
  contig = lo;
  const ContigDev c = contigs[lo];
  wb = ((int64_t)wid - c.win_base) * step; we = wb + window_size; if (we > c.len) we = c.len;
  while (wb < we && nibble_at(nib, c.nib_base + wb) == CODE_N) ++wb;
  while (wb < we && nibble_at(nib, c.nib_base + we - 1) == CODE_N) --we;
}
// Everything k_align needs to know about the (window, strand) of a candidate key.
struct CandCtx { int32_t gidx, contig_idx, m, dir, task; uint32_t wid; WindowGeom geom; int64_t first; uint8_t owned; };
CAL_D CandCtx decode_candidate(const AlignArgs& a, uint64_t key) {
  CandCtx x;
  const uint32_t strandbit = key_strandbit(a.key, key);
  x.wid = key_window(a.key, key); x.gidx = key_guide(a.key, key); x.owned = 1;
  if (a.explicit_mode) {
    const ExplicitWindow ew = a.windows[x.wid];
    x.gidx = ew.guide_idx; x.contig_idx = ew.contig_idx; x.geom.w_begin = ew.target_offset; x.geom.w_end = ew.target_offset + ew.len; x.first = ew.nib_start;
    x.task = ew.task_id; x.owned = ew.owned ? 1 : 2;
  } else {
    int64_t wb, we; locate_window(a.nib, a.contigs, a.n_contigs, a.window_size, a.step, x.wid, x.contig_idx, wb, we);
    x.geom.w_begin = (int32_t)wb; x.geom.w_end = (int32_t)we; x.first = a.contigs[x.contig_idx].nib_base + wb;
    x.owned = ((int64_t)x.wid >= a.contigs[x.contig_idx].own_lo && (int64_t)x.wid < a.contigs[x.contig_idx].own_hi) ? 1 : 2;
    x.task = (int32_t)x.wid;
  }
  x.dir = (int)(strandbit ^ (uint32_t)a.specs[x.gidx].five_prime);
  x.m = x.geom.w_end - x.geom.w_begin;
  return x;
}
// Guide alignment of candidate i -> PAM extension(s) -> hit slots (SequentialGuideAligner.scala:433-492, 505-524).
CAL_D void post_alignment(const AlignArgs& a, const CandCtx& x, const GuideSpec& g, const NibFetch& fetch, const GuideAln& aln, int64_t i) {
  if (aln.diffs > g.d) return;                                   // SequentialGuideAligner.scala:447,450
  const int64_t base = i * a.slots;
  HitX h;
  if (g.n_pams == 0) {
    make_hit(g, aln, -1, aln.score, 0, 0u, x.dir, x.geom, x.gidx, x.contig_idx, x.task, h); pack_hit(h, a.recs + base * a.rw, a.rw);
    a.ckeys[base] = ckey_make(h.score, h.start_offset, h.end_offset, h.gap_bases, h.edits, x.owned);
  } else {
    for (int pi = 0; pi < g.n_pams; ++pi) {
      int32_t score = 0, offset = 0; uint32_t xmask = 0;
      if (extend_pam(g, a.sc, fetch, x.m, aln, pi, score, offset, xmask)) {
        make_hit(g, aln, pi, score, offset, xmask, x.dir, x.geom, x.gidx, x.contig_idx, x.task, h); pack_hit(h, a.recs + (base + pi) * a.rw, a.rw);
        a.ckeys[base + pi] = ckey_make(h.score, h.start_offset, h.end_offset, h.gap_bases, h.edits, x.owned);
      }
    }
  }
}
template <int KB>                         // KB > 0: register-resident band of 2*KB+1 diagonals; KB == 0: full rectangle in local memory
CAL_D void align_body(const AlignArgs& a) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n_cand) return;
  const uint64_t key = a.cand[i];
  const int32_t col = key_col(a.key, key);
  const CandCtx x = decode_candidate(a, key);
  const GuideSpec& g = a.specs[x.gidx];
  const NibFetch fetch{ a.nib, x.first, x.m, x.dir };
  for (int s = 0; s < a.slots; ++s) a.ckeys[i * a.slots + s] = CKey{ 0, 0, 0, 0u };
  GuideAln aln;
  if (KB > 0) {                           // every guide of the launch has k_edits <= KB (defaults: d = 5 -> 5, d = 6 -> 6)
    if (!band_align_k<(KB > 0 ? KB : 1)>(g, a.sc, fetch, col, aln)) return;
  } else {                                // wide thresholds: full rectangle in local memory
    uint8_t trace[(CALITAS_MAX_PROTOSPACER + 1) * (MAX_SPAN + 1)];
    if (!band_align(g, a.sc, fetch, col, aln, trace)) return;
  }
  post_alignment(a, x, g, fetch, aln, i);
}

// ------------------------------------------------------------------------------------------------------------------------------------
// align_fast: the hot-path form of align_body<KB> for the common shapes (lp + KB + g + longest PAM <= 64 columns).  Same cells, scores,
// tie-breaks and traceback as band_align_k (tagged scores, see cal_core.cuh), same extension rule as extend_pam, same record as make_hit —
// but nothing goes through byte arrays in local memory:
//   * the 64 target codes the candidate can touch (band + PAM extension) are loaded once as nine 32-bit words and kept in registers; the reverse
//     strand costs one BREV per word (reversing the bits of a word reverses the order of its 4-bit codes AND complements each code: A=1 <-> T=8,
//     C=2 <-> G=4) instead of a complement per fetched base; from them four bit vectors say which window positions hold A, C, G, T, and a DP row
//     reads the match flags of its whole band from them with one shift (no per-cell code extraction, no per-cell match-trace bit);
//   * the traceback shifts each 2-bit op straight into a 128-bit register pair (first column ends up in the low bits = guide orientation for a
//     3' PAM; a 5' PAM reverses the fields at the end), counts and the terminal gap run are taken on the way;
//   * the extension appends the guide-PAM gap and the PAM ops with shifts, and the record leaves as two (four) 16-byte stores.
// Codes outside the window never matter: columns < 1 only feed cells whose predecessors are unreachable, columns > j are never on a path to
// the end cell, and the extension checks t_off + pam_len <= m before it looks at a base.
// ------------------------------------------------------------------------------------------------------------------------------------
CAL_D uint32_t brev32(uint32_t v) {
#if defined(__CUDA_ARCH__)
  return __brev(v);
#else
  v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1); v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2); v = ((v >> 4) & 0x0F0F0F0Fu) | ((v & 0x0F0F0F0Fu) << 4);
  return (v >> 24) | ((v >> 8) & 0xFF00u) | ((v << 8) & 0xFF0000u) | (v << 24);
#endif
}
CAL_D uint64_t rev_fields2(uint64_t v) {     // reverses the order of the 32 two-bit fields of v
  const uint32_t lo = brev32((uint32_t)v), hi = brev32((uint32_t)(v >> 32));
  const uint64_t r = ((uint64_t)lo << 32) | hi;                                        // all 64 bits reversed
  return ((r & 0x5555555555555555ull) << 1) | ((r >> 1) & 0x5555555555555555ull);      // put the two bits of each field back in order
}
struct Ops128 { uint64_t lo, hi; };
CAL_D Ops128 shl128(Ops128 v, int s) {        // 0 <= s < 128
  if (s == 0) return v;
  if (s >= 64) return Ops128{ 0ull, v.lo << (s - 64) };
  return Ops128{ v.lo << s, (v.hi << s) | (v.lo >> (64 - s)) };
}
CAL_D Ops128 shr128(Ops128 v, int s) {
  if (s == 0) return v;
  if (s >= 64) return Ops128{ v.hi >> (s - 64), 0ull };
  return Ops128{ (v.lo >> s) | (v.hi << (64 - s)), v.hi >> s };
}
CAL_D uint32_t spread_bits16(uint32_t v) {    // bit i of v -> bit 2 i
  v &= 0xFFFFu; v = (v | (v << 8)) & 0x00FF00FFu; v = (v | (v << 4)) & 0x0F0F0F0Fu; v = (v | (v << 2)) & 0x33333333u; v = (v | (v << 1)) & 0x55555555u;
  return v;
}
// target codes right of an alignment's last guide column j, for the PAM extension: code(k) = base at column j + 1 + k
struct RegCodes { uint64_t e0, e1, e2, e3; int first;
  CAL_D uint32_t operator()(int k) const { const int idx = first + k; const uint64_t e = idx < 32 ? (idx < 16 ? e0 : e1) : (idx < 48 ? e2 : e3); return (uint32_t)(e >> (4 * (idx & 15))) & 15u; } };
// Everything after the traceback of one guide alignment, shared by the register-resident aligners: the diffs filter (SequentialGuideAligner.scala:447,450),
// extend_pam for every PAM (:433-492), the ops of guide-PAM gap and PAM, Cigar.reverse for a 5' PAM, the coordinates of make_hit (:263-310, 505-524),
// the record as 16-byte stores and the canon key.  `ops`: n_g two-bit ops, first alignment column in the low bits; term / term_op: the trailing gap run.
template <class Codes>
CAL_D void emit_hit_slots(const AlignArgs& a, const CandCtx& x, const GuideSpec& g, Ops128 ops, int n_g, int term, int term_op, int32_t best, int t_start, int32_t j,
                          const Codes& after, int64_t slot0) {
  const Scores& sc = a.sc;
  const int diffs = popc32((uint32_t)((ops.lo | (ops.lo >> 1)) & 0x55555555u)) + popc32((uint32_t)(((ops.lo | (ops.lo >> 1)) >> 32) & 0x55555555u)) +
                    popc32((uint32_t)((ops.hi | (ops.hi >> 1)) & 0x55555555u)) + popc32((uint32_t)(((ops.hi | (ops.hi >> 1)) >> 32) & 0x55555555u));
  if (diffs > g.d) return;                               // SequentialGuideAligner.scala:447,450
  const int terminal_d = term_op == OP_D ? term : 0;
  const int n_pams = g.n_pams;
  for (int pi = 0; pi < (n_pams > 0 ? n_pams : 1); ++pi) {
    int32_t score = best, offset = 0; uint32_t xmask = 0; int pam_len = 0;
    if (n_pams > 0) {                                    // extend_pam (SequentialGuideAligner.scala:433-492)
      pam_len = g.pam_len[pi];
      int max_extra = g.g - term; const int alt = g.max_tot_filter - diffs; if (alt < max_extra) max_extra = alt;
      bool have = false;
      for (int off = 0; off <= max_extra; ++off) {
        int limit = g.p; const int l2 = g.max_tot_filter - diffs - off; if (l2 < limit) limit = l2;
        if (j + off + pam_len > x.m || limit < 0) continue;
        int32_t ps = 0; int nx = 0; uint32_t xm = 0;
        for (int q = 0; q < pam_len; ++q) {
          const bool pr = pairs(g.pam[pi][q], after(off + q));
          ps += pr ? sc.pam_match : sc.pam_mismatch;
          if (!((pr ? sc.pam_match : sc.pam_mismatch) > 0)) { ++nx; xm |= 1u << q; }                // op '=' iff addend > 0 (:468)
        }
        if (nx > limit) continue;
        const int32_t total = best + ps + off * sc.query_gap;
        if (!have || total > score) { have = true; score = total; offset = off; xmask = xm; }      // maxBy keeps the first maximum
      }
      if (!have) continue;
    }
    const int tail = n_pams > 0 ? offset + pam_len : 0;
    const int n_ops = n_g + tail;
    Ops128 all = ops;
    if (tail) {
      Ops128 ext = shl128(Ops128{ (uint64_t)spread_bits16(xmask), 0ull }, 2 * offset);       // the PAM columns (X = 1, '=' = 0) ...
      ext.lo |= offset ? ((1ull << (2 * offset)) - 1) : 0ull;                                 // ... behind offset x D (3): the guide-PAM gap (offset <= 26)
      ext = shl128(ext, 2 * n_g);
      all.lo |= ext.lo; all.hi |= ext.hi;
    }
    if (g.five_prime) {                                  // Cigar.reverse (:267,284): reverse the n_ops fields
      Ops128 r{ rev_fields2(all.hi), rev_fields2(all.lo) };
      all = shr128(r, 2 * (64 - n_ops));
    }
    const int32_t s0 = t_start - 1, e0c = j + tail, trail_dp = terminal_d + tail;
    int32_t start; int lead, trail;
    if (x.dir == 0) { start = x.geom.w_begin + s0; lead = 0; trail = trail_dp; }
    else { start = x.geom.w_end - e0c; lead = trail_dp; trail = 0; }                               // flip about the window (:271-274, 305-308)
    uint32_t* rec = a.recs + (slot0 + pi) * a.rw;
    const uint32_t where = rec_make_where(x.gidx, x.contig_idx, (x.dir ^ g.five_prime) != 0);
    const uint32_t shape = rec_make_shape(n_ops, e0c - s0, lead, trail, n_pams > 0 ? pi : -1);
#ifndef CAL_HOSTSIM
    uint4* r4 = reinterpret_cast<uint4*>(rec);
    r4[0] = make_uint4((uint32_t)start, (uint32_t)(x.task), (uint32_t)score, where);
    r4[1] = make_uint4(shape, (uint32_t)all.lo, (uint32_t)(all.lo >> 32), (uint32_t)all.hi);
    if (a.rw > CALITAS_HIT_WORDS) { r4[2] = make_uint4((uint32_t)(all.hi >> 32), 0u, 0u, 0u); r4[3] = make_uint4(0u, 0u, 0u, 0u); }
#else
    rec[0] = (uint32_t)start; rec[1] = (uint32_t)(x.task); rec[2] = (uint32_t)score; rec[3] = where; rec[4] = shape;
    rec[5] = (uint32_t)all.lo; rec[6] = (uint32_t)(all.lo >> 32); rec[7] = (uint32_t)all.hi;
    if (a.rw > CALITAS_HIT_WORDS) { rec[8] = (uint32_t)(all.hi >> 32); for (int k = 9; k < a.rw; ++k) rec[k] = 0u; }
#endif
    const int gaps = popc32((uint32_t)all.lo & 0xAAAAAAAAu) + popc32((uint32_t)(all.lo >> 32) & 0xAAAAAAAAu) + popc32((uint32_t)all.hi & 0xAAAAAAAAu) + popc32((uint32_t)(all.hi >> 32) & 0xAAAAAAAAu);
    const int edits = diffs + offset + popc32(xmask);    // non-'=' columns: the guide part, the guide-PAM gap, the PAM mismatches
    a.ckeys[slot0 + pi] = ckey_make(score, start, start + (e0c - s0), gaps, edits, x.owned);
  }
}
// The 64 target codes from DP column base + 1 on, in scan order, as eight words (reverse strand: reversed and complemented by BREV).
CAL_D void load_code_window(const AlignArgs& a, const CandCtx& x, int base, uint32_t A[8]) {
  const int64_t q_lo = x.dir == 0 ? x.first + base : x.first + x.m - (base + 1) - 63;       // lowest nibble index of the stretch
  const int64_t w0 = q_lo >> 3; const int sh = (int)(q_lo & 7) * 4;
  uint32_t W[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) { int64_t w = w0 + k; w = w < 0 ? 0 : (w > a.nib_last_word ? a.nib_last_word : w); W[k] = __ldg(a.nib + w); }
#pragma unroll
  for (int k = 0; k < 8; ++k) { const uint32_t v = shift_in_low_bits(W[k], W[k + 1], sh); if (x.dir == 0) A[k] = v; else A[7 - k] = brev32(v); }
}
template <int KB>
CAL_D void align_fast(const AlignArgs& a) {
  constexpr int B = 2 * KB + 1;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n_cand) return;
  const uint64_t key = a.cand[i];
  const int32_t j = key_col(a.key, key);
  const CandCtx x = decode_candidate(a, key);
  const GuideSpec& g = a.specs[x.gidx];
  for (int s = 0; s < a.slots; ++s) a.ckeys[i * a.slots + s] = CKey{ 0, 0, 0, 0u };
  const int n = g.lp;
  const int base = j - n - KB;                       // cell (i, t) is target column c = i + base + t; the register window starts at column base + 1
  // ---- the 64 codes from column base + 1 on, in scan order -----------------------------------------------------------------------------
  uint32_t A[8];
  load_code_window(a, x, base, A);
  // ---- which positions of the window pair with which base: four 64-bit vectors (bit k = the code at window position k holds base b), N excluded.
  //      A DP row then gets the match flags of its whole band with one shift: position of diagonal t in row r is r - 1 + t.
  uint64_t T[4] = { 0ull, 0ull, 0ull, 0ull };
#pragma unroll
  for (int k = 0; k < 6; ++k) {                        // 48 positions cover lp + 2 KB + 1 <= 45
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      uint32_t v = (A[k] >> b) & 0x11111111u;          // bit b of each of the word's eight codes, at stride 4 ...
      v = (v | (v >> 3)) & 0x03030303u; v = (v | (v >> 6)) & 0x000F000Fu; v = (v | (v >> 12)) & 0xFFu;     // ... compressed to eight adjacent bits
      T[b] |= (uint64_t)v << (8 * k);
    }
  }
  { const uint64_t is_n = T[0] & T[1] & T[2] & T[3];   // code 15: upper-case N never matches (SequentialGuideAligner.scala:144)
#pragma unroll
    for (int b = 0; b < 4; ++b) T[b] &= ~is_n; }
  // ---- DP fill: band_align_k's recurrence -------------------------------------------------------------------------------------------------
  const Scores& sc = a.sc;
  const int32_t NEG4 = 4 * NEG_SCORE;
  int32_t d[B], l[B], u[B];
  uint32_t trd[CALITAS_MAX_PROTOSPACER + 1], tru[CALITAS_MAX_PROTOSPACER + 1], trl[CALITAS_MAX_PROTOSPACER + 1], trm[CALITAS_MAX_PROTOSPACER + 1];
#pragma unroll
  for (int t = 0; t < B; ++t) { const int c = base + t; const int32_t v = (c >= 0 && c <= j) ? 0 : NEG4; d[t] = v + TG_DIAG; l[t] = v + TG_LEFT; u[t] = v + TG_UP; }
  const int32_t gI4 = 4 * sc.target_gap, gD4 = 4 * sc.query_gap, mis4 = 4 * sc.mismatch, match4 = 4 * sc.match;
  for (int r = 1; r <= n; ++r) {
    const uint32_t qs = g.q[r - 1];                     // the row's base set
    const uint64_t tq = ((qs & 1u) ? T[0] : 0ull) | ((qs & 2u) ? T[1] : 0ull) | ((qs & 4u) ? T[2] : 0ull) | ((qs & 8u) ? T[3] : 0ull);
    const uint32_t m = (uint32_t)(tq >> (r - 1)) & ((1u << B) - 1u);     // bit t: the row's base pairs with the target base of diagonal t
    uint32_t wd = 0, wu = 0, wl = 0;
    int32_t left_d = NEG4 + TG_DIAG, left_l = NEG4 + TG_LEFT, left_u = NEG4 + TG_UP;
#pragma unroll
    for (int t = 0; t < B; ++t) {
      const int32_t add4 = (m & (1u << t)) ? match4 : mis4;
      const int32_t md = max3_s32(d[t], l[t], u[t]);
      const int32_t mu = t + 1 < B ? (d[t + 1] > u[t + 1] ? d[t + 1] : u[t + 1]) : NEG4 + TG_DIAG;
      const int32_t ml = max3_s32(left_d, left_l, left_u);
      const int32_t nd = (md | 3) + add4, nu = ((mu & ~3) | TG_UP) + gI4, nl = ((ml & ~3) | TG_LEFT) + gD4;
      wd = shift_in_low_bits(wd, (uint32_t)md, 2); wu = shift_in_low_bits(wu, (uint32_t)mu, 2); wl = shift_in_low_bits(wl, (uint32_t)ml, 2);
      d[t] = nd; u[t] = nu; l[t] = nl; left_d = nd; left_l = nl; left_u = nu;
    }
    trd[r] = wd; tru[r] = wu; trl[r] = wl; trm[r] = m;
  }
  const int32_t mbest = max3_s32(d[KB], l[KB], u[KB]);
  const int32_t best = mbest >> 2;
  if (best < g.min_score) return;
  // ---- traceback: ops shifted into a register pair, first alignment column in the low bits ------------------------------------------------
  Ops128 ops{ 0ull, 0ull };
  int ci = n, ct = KB, cdir = mbest & 3, n_g = 0, term = 0, term_op = 0;
  while (ci > 0) {
    const int shb = 32 - 2 * B + 2 * ct;
    const int next = (int)(((cdir == TG_DIAG ? trd[ci] : (cdir == TG_UP ? tru[ci] : trl[ci])) >> shb) & 3u);
    uint32_t op;
    if (cdir == TG_DIAG) { op = ((trm[ci] >> ct) & 1u) ? OP_EQ : OP_X; --ci; }
    else if (cdir == TG_LEFT) { op = OP_D; --ct; }
    else { op = OP_I; --ci; ++ct; }
    if (ct < 0 || ct >= B) return;                       // cannot happen for an accepted end cell; keeps indexing safe
    if (n_g == term && op >= OP_I && (term == 0 || (int)op == term_op)) { ++term; term_op = (int)op; }     // trailing run of one gap kind (:452)
    ops.hi = (ops.hi << 2) | (ops.lo >> 62); ops.lo = (ops.lo << 2) | op;
    ++n_g;
    cdir = next;
  }
  // window position k is column base + 1 + k: the base right of the alignment (column j + 1) is position n + KB.  The window is loaded again (from
  // L1/L2) rather than kept in eight registers across the DP loop.
  load_code_window(a, x, base, A);
  const RegCodes after{ (uint64_t)A[0] | ((uint64_t)A[1] << 32), (uint64_t)A[2] | ((uint64_t)A[3] << 32), (uint64_t)A[4] | ((uint64_t)A[5] << 32), (uint64_t)A[6] | ((uint64_t)A[7] << 32), n + KB };
  emit_hit_slots(a, x, g, ops, n_g, term, term_op, best, base + ct + 1, j, after, i * a.slots);
}
CAL_KERNEL __launch_bounds__(128) k_align_fast6(AlignArgs a) { align_fast<6>(a); }
CAL_KERNEL __launch_bounds__(128) k_align_fast5(AlignArgs a) { align_fast<5>(a); }
CAL_KERNEL __launch_bounds__(128) k_align_fast4(AlignArgs a) { align_fast<4>(a); }

// ------------------------------------------------------------------------------------------------------------------------------------
// align_pair: align_fast for TWO candidates per thread (candidates 2p and 2p + 1 of the sorted list), their DP cells side by side in the two 16-bit
// halves of one register.  The packed forms of the cell's instructions (VIMNMX3.S16x2, VIMNMX.S16x2, VIADDMNMX.S16x2) issue at the rate of the 32-bit
// ones (measured: 17.3 against 18.5 T thread-instructions/s), the tag and mask operations are plain 32-bit logic, so one instruction stream fills two
// band matrices.  Same cells, tie order and traces as align_fast; what changes is how the numbers are held:
//   * a value is 4 * score + tag in 16 bits.  Unreachable cells start at PAIR_FLOOR and every sum is clamped there by the VIADDMNMX that forms it, so
//     nothing wraps.  A clamped value is too large by construction but never above PAIR_FLOOR + 3 + r * match4 in row r, and every cell on a path to an
//     accepted end cell holds at least 4 * min_score - (n - r) * match4: the launch takes this kernel only when
//     PAIR_FLOOR + 3 + n * match4 < 4 * min_score for every guide (pair_fits; defaults: -27 197 < 1 872), so on those cells every maximum is decided
//     between exact values, exactly as in 32 bits;
//   * the row's match flags come from bit-REVERSED position vectors, cell 0 in the top bit of each half: one PRMT turns the two top bits into two
//     half-word masks that select 4 * match or 4 * mismatch per half, and the flags move up one bit per cell;
//   * the 2-bit winners are not extracted cell by cell: acc = 4 * acc + value and ref = 4 * ref + (value with its tag forced) are two IMADs on the
//     otherwise idle FMA pipe, whatever spills out of a half spills out of both alike, and ref - acc (Diagonal: tag forced to 3) or acc - ref (Up, Left:
//     tag cleared) is the clean sum of 4^k * (3 - tag) or 4^k * tag per half; eight cells per half-word, so cells 8 .. B-1 use a second accumulator.
// MEASURED, AND NOT THE DEFAULT (CALITAS_ALIGN_PAIR=1 selects it; parity green on hostsim and on the GPU): the fill drops from 2 x 207 to 310 instructions
// per row of two candidates (ALU-pipe instructions 2 x 150 -> 175), but on 2.0 x 10^7 candidates (config 4, a quarter genome) the kernel takes 7.22 ms
// against k_align_fast6's 6.56 ms (ncu, profiles/r02r_summary.txt): traceback, PAM extension and record emission are per candidate and already were
// 32 % of align_fast's instructions, so the whole kernel executes only 15 % fewer (4.78 G against 5.65 G warp instructions), and at 96 registers it runs
// 18 warps per SM instead of 24 with two dependent local-memory walks per thread (long-scoreboard stall 3.2 per issue against 1.6; issue slots 58 %
// against 74 %).  With both walks interleaved it went from 8.5 to 7.2 ms; config 4 as a whole: 420 against 442 Gbp*guides/s.
// ------------------------------------------------------------------------------------------------------------------------------------
const int32_t PAIR_FLOOR = -32000;                      // a multiple of 4; leaves room for one addend below it (pair_fits bounds the addends)
CAL_D uint32_t p2_pack(int32_t lo, int32_t hi) { return ((uint32_t)lo & 0xFFFFu) | ((uint32_t)hi << 16); }
CAL_D int32_t p2_half(uint32_t v, int lane) { return (int32_t)(int16_t)(uint16_t)(lane ? (v >> 16) : (v & 0xFFFFu)); }
#if defined(__CUDA_ARCH__)
CAL_D uint32_t p2_max3(uint32_t a, uint32_t b, uint32_t c) { return __vimax3_s16x2(a, b, c); }
CAL_D uint32_t p2_max(uint32_t a, uint32_t b) { return __vmaxs2(a, b); }
CAL_D uint32_t p2_addmax(uint32_t a, uint32_t b, uint32_t c) { return __viaddmax_s16x2(a, b, c); }           // per half: max(a + b, c)
CAL_D uint32_t p2_signmask(uint32_t v) { uint32_t r; asm("prmt.b32 %0, %1, 0, 0xbb99;" : "=r"(r) : "r"(v)); return r; }   // per half: 0xFFFF when its top bit is set (selector bit 3: replicate the byte's sign; __byte_perm drops that bit)
CAL_D uint32_t p2_shift_in(uint32_t acc, uint32_t v) { uint32_t r; asm("mad.lo.u32 %0, %1, 4, %2;" : "=r"(r) : "r"(acc), "r"(v)); return r; }   // 4 * acc + v as ONE IMAD (FMA pipe)
#else
CAL_D uint32_t p2_max3(uint32_t a, uint32_t b, uint32_t c) {
  auto m3 = [](int32_t x, int32_t y, int32_t z) { const int32_t m = x > y ? x : y; return m > z ? m : z; };
  return p2_pack(m3(p2_half(a, 0), p2_half(b, 0), p2_half(c, 0)), m3(p2_half(a, 1), p2_half(b, 1), p2_half(c, 1)));
}
CAL_D uint32_t p2_max(uint32_t a, uint32_t b) { return p2_max3(a, b, b); }
CAL_D uint32_t p2_addmax(uint32_t a, uint32_t b, uint32_t c) {
  auto am = [](int32_t x, int32_t y, int32_t z) { const int32_t v = (int32_t)(int16_t)(uint16_t)(x + y); return v > z ? v : z; };
  return p2_pack(am(p2_half(a, 0), p2_half(b, 0), p2_half(c, 0)), am(p2_half(a, 1), p2_half(b, 1), p2_half(c, 1)));
}
CAL_D uint32_t p2_signmask(uint32_t v) { return ((v & 0x8000u) ? 0xFFFFu : 0u) | ((v & 0x80000000u) ? 0xFFFF0000u : 0u); }
CAL_D uint32_t p2_shift_in(uint32_t acc, uint32_t v) { return acc * 4u + v; }
#endif
// bit-reversed position vectors of one candidate's code window: bit 63 - k of R[b] = the code at window position k holds base b (N excluded)
CAL_D void pair_vectors(const AlignArgs& a, const CandCtx& x, int base, uint64_t R[4]) {
  uint32_t A[8];
  load_code_window(a, x, base, A);
  uint64_t T[4] = { 0ull, 0ull, 0ull, 0ull };
#pragma unroll
  for (int k = 0; k < 6; ++k) {
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      uint32_t v = (A[k] >> b) & 0x11111111u;
      v = (v | (v >> 3)) & 0x03030303u; v = (v | (v >> 6)) & 0x000F000Fu; v = (v | (v >> 12)) & 0xFFu;
      T[b] |= (uint64_t)v << (8 * k);
    }
  }
  const uint64_t is_n = T[0] & T[1] & T[2] & T[3];
#pragma unroll
  for (int b = 0; b < 4; ++b) { const uint64_t t = T[b] & ~is_n; R[b] = ((uint64_t)brev32((uint32_t)t) << 32) | (uint64_t)brev32((uint32_t)(t >> 32)); }
}
template <int KB>
CAL_D void align_pair(const AlignArgs& a) {
  constexpr int B = 2 * KB + 1;
  constexpr int B1 = B < 8 ? B : 8;                    // cells in the first trace accumulator
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t i0 = 2 * p;
  if (i0 >= a.n_cand) return;
  const bool two = i0 + 1 < a.n_cand;                  // an odd list ends with a thread whose second half repeats the first and reports nothing
  const uint64_t key0 = a.cand[i0], key1 = a.cand[two ? i0 + 1 : i0];
  const int32_t j0 = key_col(a.key, key0), j1 = key_col(a.key, key1);
  const CandCtx x0 = decode_candidate(a, key0), x1 = decode_candidate(a, key1);
  const GuideSpec& g0 = a.specs[x0.gidx]; const GuideSpec& g1 = a.specs[x1.gidx];
  for (int s = 0; s < (two ? 2 : 1) * a.slots; ++s) a.ckeys[i0 * a.slots + s] = CKey{ 0, 0, 0, 0u };
  const int n = g0.lp;                                 // the same for every guide of the launch (pair_fits)
  const int base0 = j0 - n - KB, base1 = j1 - n - KB;
  uint64_t R0[4], R1[4];
  pair_vectors(a, x0, base0, R0); pair_vectors(a, x1, base1, R1);
  // ---- DP fill ------------------------------------------------------------------------------------------------------------------------------
  const uint32_t FLOORP = p2_pack(PAIR_FLOOR, PAIR_FLOOR);
  const uint32_t MISP = p2_pack(a.mis4, a.mis4), XP = p2_pack(a.match4 ^ a.mis4, a.match4 ^ a.mis4), UPP = p2_pack(a.up4, a.up4), LEFTP = p2_pack(a.left4, a.left4);
  const uint32_t T3 = 0x00030003u, TC = 0xFFFCFFFCu;
  uint32_t d[B], l[B], u[B];
  uint32_t trd[CALITAS_MAX_PROTOSPACER + 1], tru[CALITAS_MAX_PROTOSPACER + 1], trl[CALITAS_MAX_PROTOSPACER + 1], trm[CALITAS_MAX_PROTOSPACER + 1];
  uint32_t trd2[CALITAS_MAX_PROTOSPACER + 1], tru2[CALITAS_MAX_PROTOSPACER + 1], trl2[CALITAS_MAX_PROTOSPACER + 1];
#pragma unroll
  for (int t = 0; t < B; ++t) {
    const int c0 = base0 + t, c1 = base1 + t;
    const int32_t v0 = (c0 >= 0 && c0 <= j0) ? 0 : PAIR_FLOOR, v1 = (c1 >= 0 && c1 <= j1) ? 0 : PAIR_FLOOR;
    d[t] = p2_pack(v0 + TG_DIAG, v1 + TG_DIAG); l[t] = p2_pack(v0 + TG_LEFT, v1 + TG_LEFT); u[t] = p2_pack(v0 + TG_UP, v1 + TG_UP);
  }
  for (int r = 1; r <= n; ++r) {
    const uint32_t qs0 = g0.q[r - 1], qs1 = g1.q[r - 1];
    const uint64_t tq0 = ((qs0 & 1u) ? R0[0] : 0ull) | ((qs0 & 2u) ? R0[1] : 0ull) | ((qs0 & 4u) ? R0[2] : 0ull) | ((qs0 & 8u) ? R0[3] : 0ull);
    const uint64_t tq1 = ((qs1 & 1u) ? R1[0] : 0ull) | ((qs1 & 2u) ? R1[1] : 0ull) | ((qs1 & 4u) ? R1[2] : 0ull) | ((qs1 & 8u) ? R1[3] : 0ull);
    // cell t of the row sits at window position r - 1 + t = bit 63 - (r - 1 + t) of the reversed vector: cell 0 goes to the top bit of its half
    uint32_t mm = (uint32_t)(((tq0 << (r - 1)) >> 48) & 0xFFFFu) | ((uint32_t)((tq1 << (r - 1)) >> 48) << 16);
    trm[r] = mm;
    uint32_t ad = 0, rd = 0, au = 0, ru = 0, al = 0, rl = 0;
    uint32_t left_d = p2_pack(PAIR_FLOOR + TG_DIAG, PAIR_FLOOR + TG_DIAG), left_l = p2_pack(PAIR_FLOOR + TG_LEFT, PAIR_FLOOR + TG_LEFT), left_u = p2_pack(PAIR_FLOOR + TG_UP, PAIR_FLOOR + TG_UP);
#pragma unroll
    for (int t = 0; t < B; ++t) {
      const uint32_t add4 = (p2_signmask(mm) & XP) ^ MISP;              // 4 * match where the half's flag is set, else 4 * mismatch
      mm += mm;                                                         // the next cell's flags move to the top (what crosses into the upper half stays below the bits still to be read: B <= 16)
      const uint32_t md = p2_max3(d[t], l[t], u[t]);
      const uint32_t mu = t + 1 < B ? p2_max(d[t + 1], u[t + 1]) : p2_pack(PAIR_FLOOR + TG_DIAG, PAIR_FLOOR + TG_DIAG);
      const uint32_t ml = p2_max3(left_d, left_l, left_u);
      const uint32_t fd = md | T3, fu = mu & TC, fl = ml & TC;          // the value with its tag forced to 3 / cleared
      const uint32_t nd = p2_addmax(fd, add4, FLOORP), nu = p2_addmax(fu, UPP, FLOORP), nl = p2_addmax(fl, LEFTP, FLOORP);
      if (t == B1) { trd[r] = rd - ad; tru[r] = au - ru; trl[r] = al - rl; ad = rd = au = ru = al = rl = 0; }
      ad = p2_shift_in(ad, md); rd = p2_shift_in(rd, fd); au = p2_shift_in(au, mu); ru = p2_shift_in(ru, fu); al = p2_shift_in(al, ml); rl = p2_shift_in(rl, fl);
      d[t] = nd; u[t] = nu; l[t] = nl; left_d = nd; left_l = nl; left_u = nu;
    }
    if (B > B1) { trd2[r] = rd - ad; tru2[r] = au - ru; trl2[r] = al - rl; }
    else { trd[r] = rd - ad; tru[r] = au - ru; trl[r] = al - rl; }
  }
  const uint32_t mbest2 = p2_max3(d[KB], l[KB], u[KB]);
  // ---- traceback: the two candidates' walks run side by side (two independent chains of dependent local-memory loads in flight) -----------------
  struct Walk { Ops128 ops; int ci, ct, cdir, n_g, term, term_op; };
  Walk w0{ Ops128{ 0ull, 0ull }, 0, KB, 0, 0, 0, 0 }, w1 = w0;
  const int32_t mb0 = p2_half(mbest2, 0), mb1 = p2_half(mbest2, 1);
  if ((mb0 >> 2) >= g0.min_score) { w0.ci = n; w0.cdir = mb0 & 3; }
  if (two && (mb1 >> 2) >= g1.min_score) { w1.ci = n; w1.cdir = mb1 & 3; }
  const bool live0 = w0.ci > 0, live1 = w1.ci > 0;
  bool bad0 = false, bad1 = false;
  auto walk_step = [&](Walk& w, const int lsh, bool& bad) {
    const bool second = w.ct >= B1;
    const int shb = lsh + 2 * (second ? B - 1 - w.ct : B1 - 1 - w.ct);
    const uint32_t tw = w.cdir == TG_DIAG ? (second ? trd2[w.ci] : trd[w.ci]) : (w.cdir == TG_UP ? (second ? tru2[w.ci] : tru[w.ci]) : (second ? trl2[w.ci] : trl[w.ci]));
    const int f = (int)((tw >> shb) & 3u);
    const int next = w.cdir == TG_DIAG ? 3 - f : f;
    uint32_t op;
    if (w.cdir == TG_DIAG) { op = ((trm[w.ci] >> (lsh + 15 - w.ct)) & 1u) ? OP_EQ : OP_X; --w.ci; }
    else if (w.cdir == TG_LEFT) { op = OP_D; --w.ct; }
    else { op = OP_I; --w.ci; ++w.ct; }
    if (w.ct < 0 || w.ct >= B) { bad = true; w.ci = 0; w.ct = 0; return; }          // cannot happen for an accepted end cell; keeps indexing safe
    if (w.n_g == w.term && op >= OP_I && (w.term == 0 || (int)op == w.term_op)) { ++w.term; w.term_op = (int)op; }
    w.ops.hi = (w.ops.hi << 2) | (w.ops.lo >> 62); w.ops.lo = (w.ops.lo << 2) | op;
    ++w.n_g;
    w.cdir = next;
  };
  while (w0.ci > 0 || w1.ci > 0) {
    if (w0.ci > 0) walk_step(w0, 0, bad0);
    if (w1.ci > 0) walk_step(w1, 16, bad1);
  }
  // ---- emission, as in align_fast ----------------------------------------------------------------------------------------------------------
#pragma unroll 1
  for (int lane = 0; lane < 2; ++lane) {
    if (lane ? (!live1 || bad1) : (!live0 || bad0)) continue;
    const CandCtx& x = lane ? x1 : x0; const GuideSpec& g = lane ? g1 : g0; const Walk& w = lane ? w1 : w0;
    const int32_t j = lane ? j1 : j0; const int base = lane ? base1 : base0;
    uint32_t A[8];
    load_code_window(a, x, base, A);
    const RegCodes after{ (uint64_t)A[0] | ((uint64_t)A[1] << 32), (uint64_t)A[2] | ((uint64_t)A[3] << 32), (uint64_t)A[4] | ((uint64_t)A[5] << 32), (uint64_t)A[6] | ((uint64_t)A[7] << 32), n + KB };
    emit_hit_slots(a, x, g, w.ops, w.n_g, w.term, w.term_op, (lane ? mb1 : mb0) >> 2, base + w.ct + 1, j, after, (i0 + lane) * a.slots);
  }
}
CAL_KERNEL __launch_bounds__(128, 5) k_align_pair6(AlignArgs a) { align_pair<6>(a); }
CAL_KERNEL __launch_bounds__(128, 5) k_align_pair5(AlignArgs a) { align_pair<5>(a); }
CAL_KERNEL __launch_bounds__(128, 5) k_align_pair4(AlignArgs a) { align_pair<4>(a); }

// Wide thresholds on short explicit windows (alignBest / alignToRefBest: every end column is a candidate): one thread per (window, strand)
// group of consecutive sorted candidates fills the DP once and traces every candidate column (band_align_group).
const int GROUP_W = 160;
struct GroupColAt { const uint64_t* cand; int64_t i0; KeyLayout key; CAL_D int operator()(int k) const { return (int)key_col(key, cand[i0 + k]); } };
struct GroupEmit { const AlignArgs* a; const CandCtx* x; const GuideSpec* g; const NibFetch* fetch; int64_t i0;
                   CAL_D void operator()(int k, const GuideAln& aln) const { post_alignment(*a, *x, *g, *fetch, aln, i0 + k); } };
CAL_KERNEL __launch_bounds__(128) k_mark_groups(const uint64_t* cand, int64_t n, int32_t col_bits, uint32_t* flag) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flag[i] = (i == 0 || (cand[i - 1] >> col_bits) != (cand[i] >> col_bits)) ? 1u : 0u;
}
CAL_KERNEL __launch_bounds__(128) k_group_starts(const uint32_t* flag, const uint32_t* pos, int64_t n, uint32_t* gstart) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && flag[i]) gstart[pos[i]] = (uint32_t)i;
}
CAL_KERNEL __launch_bounds__(128) k_align_group(AlignArgs a, const uint32_t* gstart, int64_t n_groups, const uint32_t* only /* NULL: every group; else the groups flagged != 0 */) {
  const int64_t gi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gi >= n_groups || (only && !only[gi])) return;
  const int64_t i0 = gstart[gi], i1 = gi + 1 < n_groups ? (int64_t)gstart[gi + 1] : a.n_cand;
  const CandCtx x = decode_candidate(a, a.cand[i0]);
  const GuideSpec& g = a.specs[x.gidx];
  const NibFetch fetch{ a.nib, x.first, x.m, x.dir };
  for (int64_t s = i0 * a.slots; s < i1 * a.slots; ++s) a.ckeys[s] = CKey{ 0, 0, 0, 0u };
  uint8_t trace[(CALITAS_MAX_PROTOSPACER + 1) * (GROUP_W + 1)];
  const GroupColAt col_at{ a.cand, i0, a.key };
  const int first = col_at(0), last = col_at((int)(i1 - i0) - 1);
  const int jlo = first - g.span > 0 ? first - g.span : 0;
  if (last - jlo <= GROUP_W) {
    const GroupEmit emit{ &a, &x, &g, &fetch, i0 };
    band_align_group<GROUP_W>(g, a.sc, fetch, (int)(i1 - i0), col_at, emit, trace);
  } else {                                // long window: candidate by candidate
    for (int64_t i = i0; i < i1; ++i) { GuideAln aln; if (band_align(g, a.sc, fetch, col_at((int)(i - i0)), aln, trace)) post_alignment(a, x, g, fetch, aln, i); }
  }
}

#ifndef CAL_HOSTSIM
// The same work with a WARP per (window, strand) group: lane r - 1 owns DP row r (a protospacer has at most 32 rows) and the matrices are filled along
// anti-diagonals, cell (r, c) at step r + c - 1.  A cell needs max3(D, L, U) of (r-1, c-1) and max(D, U) of (r-1, c) from the lane above -- two
// shuffles -- and max3 of its own previous cell; scores stay tagged as in band_align_k, so the predecessor choices fall out of the maxima.  One trace
// byte per cell goes to shared memory (33 x 161 bytes per warp), the last row's end scores too; then the lanes trace 32 candidate columns at a time,
// packing ops into registers and finishing with emit_hit_slots.  Nothing lives in local memory: against the thread-per-group kernel above (168
// registers, 7.3 KB of stack, 18 % of the warps resident, latency-bound on its byte trace) this one is issue-bound.
// Groups that do not fit (window wider than GROUP_W columns, alignments longer than 64 columns) are flagged for k_align_group.
const int GROUP_WARPS = 4;
struct SmemCodes { const uint8_t* codes; int first; CAL_D uint32_t operator()(int k) const { return codes[first + k]; } };
CAL_KERNEL __launch_bounds__(32 * GROUP_WARPS) k_align_group_warp(AlignArgs a, const uint32_t* gstart, int64_t n_groups, uint32_t* slow) {
  constexpr int TW = GROUP_W + 1, EXT = 64;
  __shared__ uint8_t s_trace[GROUP_WARPS][(CALITAS_MAX_PROTOSPACER + 1) * TW + 15];
  __shared__ uint8_t s_codes[GROUP_WARPS][TW + EXT + 3];
  __shared__ int32_t s_end[GROUP_WARPS][TW + 3];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  uint8_t* trace = s_trace[wib]; uint8_t* codes = s_codes[wib]; int32_t* ends = s_end[wib];
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const Scores& sc = a.sc;
  const int32_t gI4 = 4 * sc.target_gap, gD4 = 4 * sc.query_gap, mis4 = 4 * sc.mismatch, dmatch4 = 4 * (sc.match - sc.mismatch), NEG4 = 4 * NEG_SCORE;
  for (int64_t gi = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); gi < n_groups; gi += n_warps) {
    const int64_t i0 = gstart[gi], i1 = gi + 1 < n_groups ? (int64_t)gstart[gi + 1] : a.n_cand;
    const int n_cols = (int)(i1 - i0);
    const CandCtx x = decode_candidate(a, a.cand[i0]);
    const GuideSpec& g = a.specs[x.gidx];
    const int n = g.lp;
    const int first = key_col(a.key, a.cand[i0]), last = key_col(a.key, a.cand[i1 - 1]);
    const int jlo = first - g.span > 0 ? first - g.span : 0;
    const int W = last - jlo;
    int pam_max = 0; for (int k = 0; k < g.n_pams; ++k) pam_max = pam_max > g.pam_len[k] ? pam_max : g.pam_len[k];
    const bool fits = W <= GROUP_W && g.max_cols <= 64 && g.g + pam_max <= EXT;
    if (lane == 0) slow[gi] = fits ? 0u : 1u;
    if (!fits) continue;
    for (int64_t sl = i0 * a.slots + lane; sl < i1 * a.slots; sl += 32) a.ckeys[sl] = CKey{ 0, 0, 0, 0u };
    const NibFetch fetch{ a.nib, x.first, x.m, x.dir };
    for (int c = 1 + lane; c <= W + EXT; c += 32) codes[c] = (uint8_t)(jlo + c <= x.m ? fetch(jlo + c) : 0u);
    if (lane < n) trace[(lane + 1) * TW] = (uint8_t)((lane == 0 ? TG_DIAG : TG_UP) << 2);     // local column 0: leading insertions only
    __syncwarp();
    // ---- fill along anti-diagonals ---------------------------------------------------------------------------------------------------------
    const uint32_t qm = lane < n ? g.qmask[lane] : 0u;
    int32_t m3_p1, m3_p2, mup_p1;                       // of this row's cell at the previous step (p1) and the one before (p2): max3(D, L, U) and max(D, U), tagged
    { const int32_t d0 = NEG4 + TG_DIAG, l0 = NEG4 + TG_LEFT, u0 = (lane + 1) * gI4 + TG_UP; m3_p1 = max3_s32(d0, l0, u0); m3_p2 = m3_p1; mup_p1 = d0 > u0 ? d0 : u0; }
    const int steps = W + n - 1;
    for (int t = 1; t <= steps; ++t) {
      int32_t diag_in = __shfl_up_sync(0xFFFFFFFFu, m3_p2, 1), up_in = __shfl_up_sync(0xFFFFFFFFu, mup_p1, 1);
      if (lane == 0) { diag_in = TG_DIAG; up_in = TG_DIAG; }                                 // row 0 is all zero (free leading target): Diagonal wins its ties
      const int c = t - lane;
      int32_t n3 = m3_p1, nup = mup_p1;
      if (lane < n && c >= 1 && c <= W) {
        const uint32_t mt = (qm >> codes[c]) & 1u;
        const int32_t add4 = mis4 + (int32_t)mt * dmatch4;
        const int32_t md = diag_in, mu = up_in, ml = m3_p1;
        const int32_t nd = (md | 3) + add4, nu = ((mu & ~3) | TG_UP) + gI4, nl = ((ml & ~3) | TG_LEFT) + gD4;
        trace[(lane + 1) * TW + c] = (uint8_t)(((uint32_t)md & 3u) | (((uint32_t)mu & 3u) << 2) | (((uint32_t)ml & 3u) << 4) | (mt << 6));
        n3 = max3_s32(nd, nl, nu); nup = nd > nu ? nd : nu;
        if (lane == n - 1) ends[c] = n3;
      }
      m3_p2 = m3_p1; m3_p1 = n3; mup_p1 = nup;
    }
    __syncwarp();
    // ---- every candidate column: traceback + extension + records, 32 at a time --------------------------------------------------------------------
    for (int k = lane; k < n_cols; k += 32) {
      const int32_t j = key_col(a.key, a.cand[i0 + k]);
      const int c = j - jlo;
      const int32_t mbest = ends[c], best = mbest >> 2;
      if (best < g.min_score) continue;
      Ops128 ops{ 0ull, 0ull };
      int ci = n, cc = c, cdir = mbest & 3, n_g = 0, term = 0, term_op = 0;
      while (ci > 0 && n_g < 64) {
        const uint32_t cell = trace[ci * TW + cc];
        const int next = (int)(cdir == TG_DIAG ? (cell & 3u) : (cdir == TG_UP ? ((cell >> 2) & 3u) : ((cell >> 4) & 3u)));
        uint32_t op;
        if (cdir == TG_DIAG) { op = (cell >> 6) ? OP_EQ : OP_X; --ci; --cc; }
        else if (cdir == TG_LEFT) { op = OP_D; --cc; }
        else { op = OP_I; --ci; }
        if (n_g == term && op >= OP_I && (term == 0 || (int)op == term_op)) { ++term; term_op = (int)op; }
        ops.hi = (ops.hi << 2) | (ops.lo >> 62); ops.lo = (ops.lo << 2) | op;
        ++n_g;
        cdir = next;
      }
      const SmemCodes after{ codes, c + 1 };
      emit_hit_slots(a, x, g, ops, n_g, term, term_op, best, jlo + cc + 1, j, after, (i0 + k) * a.slots);
    }
    __syncwarp();
  }
}
#endif

CAL_KERNEL __launch_bounds__(128) k_align(AlignArgs a) { align_body<ALIGN_KB>(a); }
CAL_KERNEL __launch_bounds__(128) k_align5(AlignArgs a) { align_body<5>(a); }
CAL_KERNEL __launch_bounds__(128) k_align4(AlignArgs a) { align_body<4>(a); }
CAL_KERNEL __launch_bounds__(128) k_align_wide(AlignArgs a) { align_body<0>(a); }

// ------------------------------------------------------------------------------------------------------------------------------------
// k_canon: per (guide, window, strand) group — SequentialGuideAligner.scala:315-322
// ------------------------------------------------------------------------------------------------------------------------------------
struct CanonArgs {
  const uint64_t* cand; int64_t n_cand; const GuideSpec* specs; int32_t slots; int32_t explicit_mode; const ExplicitWindow* windows;
  const CKey* ckeys; int32_t* rank; uint32_t* gbase; uint32_t* flag; uint8_t* slot_owned; int32_t drop_halo; KeyLayout key;
};
// k_canon, two passes, one thread per candidate in both.
// Pass 1 (k_canon): which alignments are kept.  Only alignments that overlap by more than max_overlap interact, and two alignments whose end columns
// differ by at least max_cols - max_overlap cannot (an alignment has at most max_cols columns), so a thread looks no further than its CLUSTER: the run of
// neighbouring candidates of its group whose end columns follow each other closer than that -- one site, a few adjacent end columns -- and decides its
// own slots there (canon_slot_rank; clusters of more than 32 slots, i.e. tandem repeats, are decided by their first candidate's thread with canon_group).
// Pass 2 (k_canon_rank): the position of every kept alignment in its GROUP's kept list (retval order: score desc, gapBases asc, arrival), by counting
// the kept alignments of the group that sort before it.  Results per slot: flag (kept and reported), rank, gbase (first slot of the group), slot_owned;
// the r-th kept alignment of a group later goes to output position pos[gbase] + r (pos = exclusive scan of flag).
CAL_KERNEL __launch_bounds__(128) k_canon(CanonArgs a) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n_cand) return;
  const uint64_t ck = a.cand[i]; const uint64_t grp = key_group(a.key, ck);
  const int32_t gidx = a.explicit_mode ? a.windows[key_window(a.key, ck)].guide_idx : key_guide(a.key, ck);
  const int32_t max_total = a.specs[gidx].max_total_diffs, max_overlap = a.specs[gidx].max_overlap;
  const int32_t reach = max_overlap < 0 ? 0x7FFFFFFF : a.specs[gidx].max_cols - max_overlap;   // end columns this far apart (or further) never overlap by more than max_overlap (a negative limit: even disjoint ones "overlap" by 0 > limit)
  int64_t i0 = i; int32_t col = key_col(a.key, ck);
  while (i0 > 0) { const uint64_t p = a.cand[i0 - 1]; if (key_group(a.key, p) != grp || col - key_col(a.key, p) >= reach) break; col = key_col(a.key, p); --i0; }
  int64_t i1 = i + 1; col = key_col(a.key, ck);
  while (i1 < a.n_cand) { const uint64_t p = a.cand[i1]; if (key_group(a.key, p) != grp || key_col(a.key, p) - col >= reach) break; col = key_col(a.key, p); ++i1; }
  const int64_t base = i0 * a.slots; const int64_t n = (i1 - i0) * a.slots;
  const CKey* keys = a.ckeys + base;
  // the usual group is one cluster: then the cluster's kept list is the group's and pass 2 has nothing to do for these slots (gbase != ~0 tells it)
  const bool whole_group = (i0 == 0 || key_group(a.key, a.cand[i0 - 1]) != grp) && (i1 == a.n_cand || key_group(a.key, a.cand[i1]) != grp);
  if (n > 32) {
    if (i != i0) return;
    for (int64_t k = 0; k < n; ++k) a.gbase[base + k] = 0xFFFFFFFFu;
    canon_group(keys, a.rank + base, (int)n, max_total, max_overlap);
    for (int64_t k = 0; k < n; ++k) {
      const bool halo = ck_state(keys[k]) == 2;
      a.slot_owned[base + k] = halo ? 0 : 1; a.flag[base + k] = (a.rank[base + k] >= 0 && !(halo && a.drop_halo)) ? 1u : 0u;
    }
    return;
  }
  for (int q = 0; q < a.slots; ++q) {
    const int k = (int)(i - i0) * a.slots + q; const int64_t s = base + k;
    const int st = ck_state(keys[k]); const bool halo = st == 2;
    const int r = st == 0 ? -1 : canon_slot_rank(keys, (int)n, k, max_total, max_overlap);
    a.slot_owned[s] = halo ? 0 : 1; a.flag[s] = (r >= 0 && !(halo && a.drop_halo)) ? 1u : 0u;
    a.rank[s] = r; a.gbase[s] = whole_group ? (uint32_t)base : 0xFFFFFFFFu;
  }
}
CAL_KERNEL __launch_bounds__(128) k_canon_rank(CanonArgs a) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n_cand) return;
  if (a.gbase[i * a.slots] != 0xFFFFFFFFu) return;            // pass 1 ranked this candidate's slots: its cluster was its whole group
  const uint64_t grp = key_group(a.key, a.cand[i]);
  int64_t i0 = i; while (i0 > 0 && key_group(a.key, a.cand[i0 - 1]) == grp) --i0;
  int64_t i1 = i + 1; while (i1 < a.n_cand && key_group(a.key, a.cand[i1]) == grp) ++i1;
  const int64_t base = i0 * a.slots; const int64_t n = (i1 - i0) * a.slots;
  for (int q = 0; q < a.slots; ++q) {
    const int64_t k = (i - i0) * a.slots + q, s = base + k;
    a.gbase[s] = (uint32_t)base;
    if (!a.flag[s]) { a.rank[s] = -1; continue; }
    const CKey me = a.ckeys[s];
    int r = 0;
    for (int64_t y = 0; y < n; ++y) if (y != k && a.flag[base + y] && ck_before(a.ckeys[base + y], (int)y, me, (int)k)) ++r;
    a.rank[s] = r;
  }
}
#ifndef CAL_HOSTSIM
// Warp-per-group variant for large groups (best mode: every end column of a window is a candidate, 60-250 alignments per group).  Same
// result as canon_group: the best undecided alignment in (score desc, gapBases asc, arrival) order is decided next; when it is kept, every
// undecided alignment overlapping it by more than max_overlap is dropped at once (it would be dropped on its turn anyway, and deciding it
// early cannot change a later decision because dropped alignments never block anything).  Persistent grid, warps stride over the groups.
CAL_KERNEL __launch_bounds__(128) k_canon_warp(CanonArgs a, const uint32_t* gstart, const uint32_t* gpos, const uint32_t* gflag) {
  const int64_t n_groups = (int64_t)gpos[a.n_cand - 1] + gflag[a.n_cand - 1];
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t gi = warp; gi < n_groups; gi += n_warps) {
    const int64_t i0 = gstart[gi], i1 = gi + 1 < n_groups ? (int64_t)gstart[gi + 1] : a.n_cand;
    const int64_t base = i0 * a.slots; const int n = (int)((i1 - i0) * a.slots);
    const int32_t gidx = a.explicit_mode ? a.windows[key_window(a.key, a.cand[i0])].guide_idx : key_guide(a.key, a.cand[i0]);
    const int32_t max_total = a.specs[gidx].max_total_diffs, max_overlap = a.specs[gidx].max_overlap;
    const CKey* keys = a.ckeys + base;
    bool halo = false;
    for (int k = lane; k < n; k += 32) { const int st = ck_state(keys[k]); a.rank[base + k] = (st == 0 || ck_edits(keys[k]) > max_total) ? -1 : -2; halo |= st == 2; }
    halo = __any_sync(0xFFFFFFFFu, halo);
    __syncwarp();
    int kept = 0;
    for (;;) {
      unsigned long long best = ~0ull;
      for (int k = lane; k < n; k += 32) {
        if (a.rank[base + k] != -2) continue;
        const unsigned long long key = ((unsigned long long)(uint32_t)(0x7FFFFFFF - keys[k].score) << 32) | ((unsigned long long)ck_gaps(keys[k]) << 20) | (unsigned long long)k;
        if (key < best) best = key;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { const unsigned long long other = __shfl_xor_sync(0xFFFFFFFFu, best, o); if (other < best) best = other; }
      if (best == ~0ull) break;
      const int b = (int)(best & 0xFFFFFu);
      const CKey hb = keys[b];
      __syncwarp();
      if (lane == 0) a.rank[base + b] = kept;                 // nothing kept so far overlaps it: those would have dropped it when they were kept
      ++kept;
      for (int k = lane; k < n; k += 32) {
        if (k == b || a.rank[base + k] != -2) continue;
        if (ck_overlap(keys[k], hb) > max_overlap) a.rank[base + k] = -1;
      }
      __syncwarp();
    }
    for (int k = lane; k < n; k += 32) { a.gbase[base + k] = (uint32_t)base; a.slot_owned[base + k] = halo ? 0 : 1; a.flag[base + k] = (a.rank[base + k] >= 0 && !(halo && a.drop_halo)) ? 1u : 0u; }
    __syncwarp();
  }
}
#endif
// record copy: rw words as 16-byte pieces (records are 32 or 64 bytes, 32-byte aligned)
CAL_D void copy_rec(uint32_t* dst, const uint32_t* src, int rw) {
#ifndef CAL_HOSTSIM
  const uint4* s4 = reinterpret_cast<const uint4*>(src); uint4* d4 = reinterpret_cast<uint4*>(dst);
  for (int k = 0; k < rw / 4; ++k) d4[k] = s4[k];
#else
  for (int k = 0; k < rw; ++k) dst[k] = src[k];
#endif
}
// flagged slot s -> output position pos[gbase[s]] + rank[s]: groups in candidate order, the kept alignments of a group in retval order
CAL_KERNEL __launch_bounds__(256) k_gather_flagged(const uint32_t* recs, int32_t rw, const int32_t* rank, const uint32_t* gbase, const uint32_t* flag, const uint32_t* pos, int64_t n, uint32_t* out,
                                                   const uint8_t* slot_owned, uint8_t* out_owned) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n || !flag[s]) return;
  const int64_t at = (int64_t)pos[gbase[s]] + rank[s];
  copy_rec(out + at * rw, recs + s * rw, rw);
  if (out_owned) out_owned[at] = slot_owned[s];
}

// ------------------------------------------------------------------------------------------------------------------------------------
// dedup: removeOverlaps (SearchReference.scala:653-675) + ReferenceHit.sort (ReferenceHit.scala:276-287)
// ------------------------------------------------------------------------------------------------------------------------------------
// Sort key = guide (relative to the chunk) | contig | coordinate_start | strand | score_hi - score, each field only as wide as the call needs
// (49 bits for a 16-guide chunk on hg38 with the default costs).  ONE stable radix sort over it yields the order
// (guide, contig, start, strand, -score, arrival): that is ReferenceHit.sort's order already, and within it the hits of one
// (guide, contig, strand) group -- what removeOverlaps sweeps -- appear by (start, -score, arrival), interleaved with the other strand's.
// So the sweep runs the two strands' chains side by side and no second sort follows.  When the fields do not fit 64 bits (not reachable with
// real genomes and costs) the score goes into keyA instead and two stable sorts give the same order.  Arrival order is the array index.
// A value outside its field would mis-sort silently, so it raises *overflow, which the host reads with the sweep's counters.
struct DedupLayout { int32_t start_bits, contig_shift, guide_shift, bits /* without the score */, g0, score_hi, score_bits, merged; };
// One thread per alignment slot: the kept ones (flag) get their sort key at pos[s] (their arrival rank: group order, then retval order), with the
// slot of the record as the sort value; nothing is copied until the keepers are known.
CAL_KERNEL __launch_bounds__(256) k_dedup_keys(const uint32_t* recs, int32_t rw, const int32_t* rank, const uint32_t* gbase, const uint32_t* flag, const uint32_t* pos, int64_t n_slots, DedupLayout L,
                                               uint64_t* key, uint64_t* keyA, uint32_t* idx, uint32_t* overflow) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_slots || !flag[s]) return;
  const int64_t i = (int64_t)pos[gbase[s]] + rank[s]; const uint32_t src = (uint32_t)s;
  const uint32_t* h = recs + (int64_t)src * rw;
  const int64_t start = rec_gstart(h), sc = (int64_t)L.score_hi - rec_score(h), g = (int64_t)rec_guide(h) - L.g0, c = rec_contig(h);
  if (start < 0 || (start >> L.start_bits) != 0 || sc < 0 || (sc >> L.score_bits) != 0 || g < 0 || c < 0 ||
      (uint64_t)c >= (1ull << (L.guide_shift - L.contig_shift)) || (L.bits < 64 && ((uint64_t)g >> (L.bits - L.guide_shift)) != 0)) *overflow = 1u;
  const uint64_t k = ((uint64_t)g << L.guide_shift) | ((uint64_t)c << L.contig_shift) | ((uint64_t)start << 1) | (uint64_t)rec_neg(h);
  if (L.merged) key[i] = (k << L.score_bits) | (uint64_t)sc;
  else { key[i] = k; keyA[i] = (uint64_t)sc; }
  idx[i] = src;
}
CAL_KERNEL __launch_bounds__(256) k_gather_u64(const uint64_t* in, const uint32_t* idx, int64_t n, uint64_t* out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[idx[i]];
}
CAL_KERNEL __launch_bounds__(256) k_gather_u32(const uint32_t* in, const uint32_t* idx, int64_t n, uint32_t* out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[idx[i]];
}
CAL_KERNEL __launch_bounds__(256) k_iota_u32(uint32_t* out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (uint32_t)i;
}
CAL_KERNEL __launch_bounds__(256) k_sweep_prepare(const uint32_t* recs, int32_t rw, const uint8_t* owned, const uint32_t* idx, int64_t n, int32_t* s_start, int32_t* s_end, int32_t* s_score, uint8_t* s_owned) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t* h = recs + (int64_t)idx[i] * rw;
  s_start[i] = rec_gstart(h); s_end[i] = rec_sweep_end(h); s_score[i] = rec_score(h); s_owned[i] = owned[idx[i]];
}
CAL_HD int32_t soa_overlap(const int32_t* s_start, const int32_t* s_end, int64_t a, int64_t b) {
  const int32_t hi = s_end[a] < s_end[b] ? s_end[a] : s_end[b], lo = s_start[a] > s_start[b] ? s_start[a] : s_start[b];
  const int32_t o = hi - lo; return o > 0 ? o : 0;
}
// One thread per sweep segment of a (guide, contig) run of the sorted list.  The two strands are separate removeOverlaps groups
// (SearchReference.scala:656) whose hits interleave here by start, so the thread keeps one "current hit" per strand.  Per strand the
// reference's loop (:662-672) is: take cur; skip the following hits while they overlap cur by >= max_overlap and score no more; the hit that
// ends the skipping decides cur (kept iff it overlaps cur by < max_overlap, or there is none) and becomes the next cur.
// With segmented != 0 (max_overlap >= 1) a segment also starts where a hit begins more than CALITAS_MAX_OPS bases after its predecessor in the
// list: no earlier hit of either strand can then overlap it at all, so both chains restart there unconditionally and segments are
// independent.  With max_overlap <= 0 every later hit "overlaps" (>= 0) and a (guide, contig) run is swept by one thread.
// Halo sentinel (sharded references): a run that begins with halo hits begins at an interior shard cut, where this engine does not know the state of the
// reference's loop (its current hit may be one this shard never saw).  A hit R is a RESTART -- it ends any skipping and becomes the current hit whatever came
// before, in the reference's loop as in this one -- when it overlaps every earlier hit of its strand by less than max_overlap.  For the hits this shard sees
// that is checked directly; the hits it does not see lie in windows more than HALO_WINDOWS before its first owned window, i.e. they end before
// (first owned hit's start) - halo_bases, so they cannot touch a hit that starts at or after that.  The first owned hit of a run is therefore decided exactly
// iff such a restart exists at or before it; if not, the halo was too short for this input and *unsafe is raised: the call fails with CALITAS_ELIMIT
// instead of returning a guess ("Multi-GPU").
CAL_KERNEL __launch_bounds__(128) k_sweep(const uint64_t* key, const int32_t* s_start, const int32_t* s_end, const int32_t* s_score, const uint8_t* s_owned, int64_t n, int32_t max_overlap,
                                          int32_t segmented, int32_t strand_shift /* bit of the strand; the (guide, contig) run is key >> (strand_shift + 1 + start_bits) */, int32_t start_bits,
                                          uint32_t* keep, uint32_t* unsafe, int32_t halo_bases) {
  const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i0 >= n) return;
  const int gshift = strand_shift + 1 + start_bits;
  const uint64_t grp = key[i0] >> gshift;
  const bool run_start = i0 == 0 || (key[i0 - 1] >> gshift) != grp;
  if (!run_start && !(segmented && s_start[i0] - s_start[i0 - 1] > CALITAS_MAX_OPS)) return;
  int64_t cur[2] = { -1, -1 };
  const bool at_cut = run_start && !s_owned[i0];            // the list continues to the left, out of this shard's sight
  bool unknown[2] = { at_cut, at_cut }; int32_t max_end[2] = { -0x7FFFFFFF, -0x7FFFFFFF }, last_restart[2] = { -0x7FFFFFFF, -0x7FFFFFFF };
  for (int64_t i = i0; i < n; ++i) {
    if (i > i0 && ((key[i] >> gshift) != grp || (segmented && s_start[i] - s_start[i - 1] > CALITAS_MAX_OPS))) break;
    const int st = (int)((key[i] >> strand_shift) & 1ull);
    if (unknown[st]) {
      if (max_end[st] == -0x7FFFFFFF || max_end[st] - s_start[i] < max_overlap) last_restart[st] = s_start[i];      // overlaps every earlier visible hit by < max_overlap
      if (s_owned[i]) {                                      // the first owned hit of this strand: was there a restart out of the unseen hits' reach?
        if (last_restart[st] == -0x7FFFFFFF || last_restart[st] < s_start[i] - halo_bases) *unsafe = 1u;
        unknown[st] = false;
      }
      if (s_end[i] > max_end[st]) max_end[st] = s_end[i];
    }
    const int64_t c = cur[st];
    if (c < 0) cur[st] = i;
    else if (soa_overlap(s_start, s_end, i, c) >= max_overlap && s_score[i] <= s_score[c]) keep[i] = 0;
    else { keep[c] = (soa_overlap(s_start, s_end, i, c) < max_overlap && s_owned[c]) ? 1u : 0u; cur[st] = i; }      // halo hits take part, are never reported
  }
  for (int st = 0; st < 2; ++st) if (cur[st] >= 0) keep[cur[st]] = s_owned[cur[st]] ? 1u : 0u;
}
// out[pos[i]] = hits[idx[i]] for the keepers: the sorted order is the final order
CAL_KERNEL __launch_bounds__(256) k_gather_keepers(const uint32_t* recs, int32_t rw, const uint32_t* idx, const uint32_t* keep, const uint32_t* pos, int64_t n, uint32_t* out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && keep[i]) copy_rec(out + (int64_t)pos[i] * rw, recs + (int64_t)idx[i] * rw, rw);
}

// ------------------------------------------------------------------------------------------------------------------------------------
// SearchReference -v on the device (SearchReference.scala:570-630, 641-648, 653-675): hits of the reference windows and of the variant windows are
// merged, de-duplicated per (guide, contig, strand, variant set) and sorted here; the host only builds the windows and renders the keepers.
// ------------------------------------------------------------------------------------------------------------------------------------
struct VarWindowDev { int64_t nib_start; int32_t length, contig_idx, ref_start /* 1-based */, n_alleles, first_allele, first_set, owned, pad_; };
struct VarInfo { int32_t window_idx, start_offset, end_offset, guide_start_offset, guide_end_offset, set_rank; };      // == calitas_variant_hit_info
// VariantWindow.refOffsetAtBaseOffset (SearchReference.scala:133-156) from the allele list: the window's cigar is, per allele in position order, an
// M run up to the allele, then M x len (same length) | M + I x (alt - 1) | M + D x (ref - 1) | D x ref + I x alt (:282-319), then an M tail.
CAL_D int32_t var_ref_offset(const VarWindowDev& w, const calitas_variant_allele* al, int32_t offset, bool preceding) {
  int32_t ref_off = w.ref_start - 1, base_off = 0, ref_pos = w.ref_start;
  // one cigar element; returns true when `offset` falls into it
#define CAL_VSTEP(OP, LEN) { const int32_t len_ = (LEN); const int32_t onq_ = (OP) == 2 ? 0 : len_, onr_ = (OP) == 1 ? 0 : len_; \
    if (offset < base_off + onq_ && offset < w.length) return (OP) == 1 ? (preceding ? ref_off - 1 : ref_off) : ref_off + (offset - base_off); \
    ref_off += onr_; base_off += onq_; }
  for (int k = 0; k < w.n_alleles; ++k) {
    const calitas_variant_allele a = al[w.first_allele + k];
    const int32_t pre = a.pos - ref_pos;
    if (pre > 0) { CAL_VSTEP(0, pre) ref_pos += pre; }
    if (a.ref_len == a.alt_len) { CAL_VSTEP(0, a.ref_len) }
    else if (a.ref_len == 1) { CAL_VSTEP(0, 1) CAL_VSTEP(1, a.alt_len - 1) }
    else if (a.alt_len == 1) { CAL_VSTEP(0, 1) CAL_VSTEP(2, a.ref_len - 1) }
    else { CAL_VSTEP(2, a.ref_len) CAL_VSTEP(1, a.alt_len) }
    ref_pos += a.ref_len;
  }
  CAL_VSTEP(0, w.length - base_off)
#undef CAL_VSTEP
  return ref_off;                           // offset == window length: start - 1 + cigar.lengthOnTarget (:134-136)
}
struct VarKeyLayout { int32_t set_bits, contig_bits, score_hi, score_bits, start_bits; };
// What calitas_search_variants adds to a search: the variant windows whose hits are merged with the reference windows' hits.
struct VariantPlan { const calitas_variant_set* vs; const int32_t* guide_class; };
// tasks of a span of guides: guide g (class c) is aligned against the windows [class_begin[c], class_end[c]) of the set; task t of the span belongs to
// the guide whose offset range holds t
struct VarTaskGuide { int64_t first_task; int32_t guide, w_begin; };
CAL_KERNEL __launch_bounds__(256) k_variant_tasks(const VarWindowDev* windows, const VarTaskGuide* guides, int32_t n_guides, int64_t n_tasks, ExplicitWindow* out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_tasks) return;
  int lo = 0, hi = n_guides - 1;
  while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (guides[mid].first_task <= t) lo = mid; else hi = mid - 1; }
  const VarTaskGuide g = guides[lo];
  const int32_t w = g.w_begin + (int32_t)(t - g.first_task);
  const VarWindowDev v = windows[w];
  out[t] = ExplicitWindow{ v.nib_start, v.length, 0, g.guide, v.contig_idx, w, v.owned };
}
// One thread per merged hit (reference hits first, then variant-window hits, each in arrival order): reference coordinates, the variant set the hit
// overlaps (ReferenceHit.scala:211), and the two sort keys of the sweep order: major = guide | contig | set, minor = start | strand | score_hi - score.
CAL_KERNEL __launch_bounds__(256) k_variant_keys(const uint32_t* recs, int32_t rw, int64_t n_ref, int64_t n, const VarWindowDev* windows, const calitas_variant_allele* alleles,
                                                 const uint32_t* set_rank, VarKeyLayout L, VarInfo* info, uint64_t* major, uint64_t* minor, uint32_t* idx, uint32_t* overflow) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t* h = recs + i * rw;
  VarInfo v;
  if (i < n_ref) { v.window_idx = -1; v.start_offset = rec_start(h); v.end_offset = rec_end(h); v.guide_start_offset = rec_gstart(h); v.guide_end_offset = rec_gend(h); v.set_rank = 0; }
  else {
    v.window_idx = rec_task(h);
    const VarWindowDev w = windows[v.window_idx];
    v.start_offset = var_ref_offset(w, alleles, rec_start(h), true); v.end_offset = var_ref_offset(w, alleles, rec_end(h), false);                 // SearchReference.scala:615-620
    v.guide_start_offset = var_ref_offset(w, alleles, rec_gstart(h), true); v.guide_end_offset = var_ref_offset(w, alleles, rec_gend(h), false);
    int a = 0; while (a < w.n_alleles && alleles[w.first_allele + a].pos - 1 < v.start_offset) ++a;                                               // ReferenceHit.scala:211: pos - 1 in [start, end]
    int b = a; while (b < w.n_alleles && alleles[w.first_allele + b].pos - 1 <= v.end_offset) ++b;
    v.set_rank = b > a ? (int32_t)set_rank[w.first_set + a * w.n_alleles - a * (a - 1) / 2 + (b - a - 1)] : 0;
  }
  info[i] = v;
  const int64_t sc = (int64_t)L.score_hi - rec_score(h), start = v.guide_start_offset, c = rec_contig(h);
  if (start < 0 || (start >> L.start_bits) != 0 || sc < 0 || (sc >> L.score_bits) != 0 || c < 0 || (c >> L.contig_bits) != 0 || ((uint64_t)(uint32_t)v.set_rank >> L.set_bits) != 0) *overflow = 1u;
  major[i] = (((uint64_t)rec_guide(h) << L.contig_bits | (uint64_t)c) << L.set_bits) | (uint64_t)(uint32_t)v.set_rank;
  minor[i] = ((((uint64_t)start << 1) | (uint64_t)rec_neg(h)) << L.score_bits) | (uint64_t)sc;
  idx[i] = (uint32_t)i;
}
// sorted position i -> what k_sweep reads: key = major << 1 | strand (group = key >> 1), start, sweep end (ReferenceHit.scala:135-138: the hit's own span), score, owned
CAL_KERNEL __launch_bounds__(256) k_variant_sweep_prepare(const uint32_t* recs, int32_t rw, const VarInfo* info, const uint8_t* owned, const uint64_t* major_sorted, const uint32_t* idx, int64_t n,
                                                          uint64_t* key, int32_t* s_start, int32_t* s_end, int32_t* s_score, uint8_t* s_owned) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t src = idx[i]; const uint32_t* h = recs + (int64_t)src * rw;
  key[i] = (major_sorted[i] << 1) | (uint64_t)rec_neg(h);
  s_start[i] = info[src].guide_start_offset; s_end[i] = info[src].guide_start_offset + rec_span(h) - 1; s_score[i] = rec_score(h); s_owned[i] = owned[src];
}
// keepers, in sweep order: their final sort key (ReferenceHit.sort: guide, contig, coordinate_start, strand, -score) and their index into the merged list
CAL_KERNEL __launch_bounds__(256) k_variant_final_keys(const uint32_t* recs, int32_t rw, const VarInfo* info, const uint32_t* idx, const uint32_t* keep, const uint32_t* pos, int64_t n, DedupLayout L,
                                                       uint64_t* key, uint32_t* out_idx, uint32_t* overflow) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !keep[i]) return;
  const uint32_t src = idx[i]; const uint32_t* h = recs + (int64_t)src * rw;
  const int64_t start = info[src].guide_start_offset, sc = (int64_t)L.score_hi - rec_score(h), g = rec_guide(h), c = rec_contig(h);
  if (start < 0 || (start >> L.start_bits) != 0 || sc < 0 || (sc >> L.score_bits) != 0 || c < 0 || (uint64_t)c >= (1ull << (L.guide_shift - L.contig_shift))) *overflow = 1u;
  const uint64_t k = ((uint64_t)g << L.guide_shift) | ((uint64_t)c << L.contig_shift) | ((uint64_t)start << 1) | (uint64_t)rec_neg(h);
  key[pos[i]] = (k << L.score_bits) | (uint64_t)sc;
  out_idx[pos[i]] = src;
}
CAL_KERNEL __launch_bounds__(256) k_variant_gather(const uint32_t* recs, int32_t rw, const VarInfo* info, const uint32_t* idx, int64_t n, uint32_t* out, VarInfo* out_info) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  copy_rec(out + i * rw, recs + (int64_t)idx[i] * rw, rw); out_info[i] = info[idx[i]];
}

// Counters the host waits for are stored straight into mapped pinned host memory by these one-thread kernels (see dev::alloc_host_mapped).
CAL_KERNEL k_publish_u64(const unsigned long long* src, unsigned long long* dst_host) {
  *(volatile unsigned long long*)dst_host = *src;
#ifndef CAL_HOSTSIM
  __threadfence_system();
#endif
}
// dst_host[0] = *a + *b (last exclusive-scan position + last flag = number of flagged items); dst_host[1] = *c when c is given
CAL_KERNEL k_publish_sum(const uint32_t* a, const uint32_t* b, const unsigned long long* c, unsigned long long* dst_host) {
  ((volatile unsigned long long*)dst_host)[0] = (unsigned long long)*a + (unsigned long long)*b;
  if (c) ((volatile unsigned long long*)dst_host)[1] = *c;
#ifndef CAL_HOSTSIM
  __threadfence_system();
#endif
}

// ------------------------------------------------------------------------------------------------------------------------------------
// k_int_peak: integer-issue microbenchmark for the roofline denominator (no integer peak in MEASURED_PEAKS.json).
// Eight chains per thread; every statement reads two other chains, so ptxas cannot fold consecutive operations of a chain into one.
//   kind 0  LOP3 only            -> ALU-pipe issue rate (the pipe k_scan_tiled is bound by)
//   kind 1  IMAD only            -> FMA-pipe issue rate
//   kind 2  LOP3 and IMAD, 1:1   -> both pipes together (scheduler issue limit)
//   kind 3  IMAD.HI only, kind 4 LEA.HI only: rates of the two candidate instructions for the score update
//   kind 5-11: the instructions of the DP cell (VIMNMX3, its packed 16-bit forms, SHF) and ALU + FMA mixes with two register operands per instruction
// ------------------------------------------------------------------------------------------------------------------------------------
#ifndef CAL_HOSTSIM
CAL_D uint32_t pk_lop3(uint32_t a, uint32_t b, uint32_t c) { uint32_t d; asm volatile("lop3.b32 %0, %1, %2, %3, 0xe8;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
CAL_D uint32_t pk_imad(uint32_t a, uint32_t b, uint32_t c) { uint32_t d; asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
CAL_D uint32_t pk_imadhi(uint32_t a, uint32_t b, uint32_t c) { uint32_t d; asm volatile("mad.hi.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
CAL_D uint32_t pk_leahi(uint32_t a, uint32_t b, uint32_t c) { return b + (a >> 31) + (c & 0u); }
CAL_D uint32_t pk_max3(uint32_t a, uint32_t b, uint32_t c) { return (uint32_t)__vimax3_s32((int)a, (int)b, (int)c); }
CAL_D uint32_t pk_max3x2(uint32_t a, uint32_t b, uint32_t c) { return __vimax3_s16x2(a, b, c); }
CAL_D uint32_t pk_addmaxx2(uint32_t a, uint32_t b, uint32_t c) { return __viaddmax_s16x2(a, b, c); }
CAL_D uint32_t pk_shf(uint32_t a, uint32_t b, uint32_t c) { return __funnelshift_r(a, b, 2) + (c & 0u); }
CAL_D uint32_t pk_lop2i(uint32_t a, uint32_t b, uint32_t c) { uint32_t d; asm volatile("lop3.b32 %0, %1, %2, 0x00030003, 0xf8;" : "=r"(d) : "r"(a), "r"(b)); return d + (c & 0u); }     // two registers + an immediate
CAL_D uint32_t pk_imad1(uint32_t a, uint32_t b, uint32_t c) { uint32_t d; asm volatile("mad.lo.u32 %0, %1, 5, %2;" : "=r"(d) : "r"(a), "r"(b)); return d + (c & 0u); }                // register x immediate + register
#else
inline uint32_t pk_lop3(uint32_t a, uint32_t b, uint32_t c) { return (a & b) | (a & c) | (b & c); }
inline uint32_t pk_imad(uint32_t a, uint32_t b, uint32_t c) { return a * b + c; }
inline uint32_t pk_imadhi(uint32_t a, uint32_t b, uint32_t c) { return (uint32_t)(((uint64_t)a * b) >> 32) + c; }
inline uint32_t pk_leahi(uint32_t a, uint32_t b, uint32_t c) { return b + (a >> 31) + (c & 0u); }
inline uint32_t pk_max3(uint32_t a, uint32_t b, uint32_t c) { return a > b ? (a > c ? a : c) : (b > c ? b : c); }
inline uint32_t pk_max3x2(uint32_t a, uint32_t b, uint32_t c) { return pk_max3(a, b, c); }
inline uint32_t pk_addmaxx2(uint32_t a, uint32_t b, uint32_t c) { return a + b > c ? a + b : c; }
inline uint32_t pk_shf(uint32_t a, uint32_t b, uint32_t c) { return (a >> 2) | (b << 30) | (c & 0u); }
inline uint32_t pk_lop2i(uint32_t a, uint32_t b, uint32_t c) { return (a | (b & 0x00030003u)) + (c & 0u); }
inline uint32_t pk_imad1(uint32_t a, uint32_t b, uint32_t c) { return a * 5u + b + (c & 0u); }
#endif
#define CAL_PEAK_ROUND(OP_EVEN, OP_ODD) _Pragma("unroll") for (int k = 0; k < 8; ++k) a[k] = (k & 1) ? OP_ODD(a[k], a[(k + 3) & 7], a[(k + 5) & 7]) : OP_EVEN(a[k], a[(k + 3) & 7], a[(k + 5) & 7]);
CAL_KERNEL __launch_bounds__(256) k_int_peak(uint32_t* out, int iters, int kind, uint32_t seed) {
  uint32_t a[8], c = seed * 2654435761u + threadIdx.x;
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] = c + k * 0x9E3779B9u;
  if (kind == 0)      for (int it = 0; it < iters; ++it) { CAL_PEAK_ROUND(pk_lop3, pk_lop3) CAL_PEAK_ROUND(pk_lop3, pk_lop3) }
  else if (kind == 1) for (int it = 0; it < iters; ++it) { CAL_PEAK_ROUND(pk_imad, pk_imad) CAL_PEAK_ROUND(pk_imad, pk_imad) }
  else if (kind == 2) for (int it = 0; it < iters; ++it) { CAL_PEAK_ROUND(pk_lop3, pk_imad) CAL_PEAK_ROUND(pk_imad, pk_lop3) }
  else if (kind == 3) for (int it = 0; it < iters; ++it) { CAL_PEAK_ROUND(pk_imadhi, pk_imadhi) CAL_PEAK_ROUND(pk_imadhi, pk_imadhi) }
  else if (kind == 4) for (int it = 0; it < iters; ++it) { CAL_PEAK_ROUND(pk_leahi, pk_leahi) CAL_PEAK_ROUND(pk_leahi, pk_leahi) }
  else if (kind == 5) for (int it = 0; it < iters; ++it) { CAL_PEAK_ROUND(pk_max3, pk_max3) CAL_PEAK_ROUND(pk_max3, pk_max3) }                  // VIMNMX3
  else if (kind == 6) for (int it = 0; it < iters; ++it) { CAL_PEAK_ROUND(pk_max3x2, pk_max3x2) CAL_PEAK_ROUND(pk_max3x2, pk_max3x2) }          // VIMNMX3.S16x2
  else if (kind == 7) for (int it = 0; it < iters; ++it) { CAL_PEAK_ROUND(pk_addmaxx2, pk_addmaxx2) CAL_PEAK_ROUND(pk_addmaxx2, pk_addmaxx2) }  // VIADDMNMX.S16x2
  else if (kind == 8) for (int it = 0; it < iters; ++it) { CAL_PEAK_ROUND(pk_shf, pk_shf) CAL_PEAK_ROUND(pk_shf, pk_shf) }                      // SHF (funnel shift)
  else if (kind == 9) for (int it = 0; it < iters; ++it) { CAL_PEAK_ROUND(pk_lop2i, pk_imad1) CAL_PEAK_ROUND(pk_imad1, pk_lop2i) }              // LOP3 and IMAD with two register operands each, 1:1
  else if (kind == 10) for (int it = 0; it < iters; ++it) { CAL_PEAK_ROUND(pk_max3x2, pk_imad1) CAL_PEAK_ROUND(pk_imad1, pk_max3x2) }           // VIMNMX3.S16x2 and IMAD, 1:1
  else                for (int it = 0; it < iters; ++it) { CAL_PEAK_ROUND(pk_lop2i, pk_lop2i) CAL_PEAK_ROUND(pk_lop2i, pk_lop2i) }              // kind 11: LOP3 with an immediate
  uint32_t r = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) r ^= a[k];
  if (r == 0x12345678u) out[blockIdx.x * blockDim.x + threadIdx.x] = r;               // keeps the chains alive
}

// ------------------------------------------------------------------------------------------------------------------------------------
// host-side engine
// ------------------------------------------------------------------------------------------------------------------------------------
struct DBuf {
  void* p = nullptr; size_t cap = 0;
  void ensure(size_t bytes) { if (bytes > cap) { dev::free_(p); p = nullptr; cap = 0; size_t want = bytes + bytes / 4 + 256; p = dev::alloc(want); cap = want; } }
  void ensure_keep(size_t bytes, size_t used, dev::Stream s) {
    if (bytes <= cap) return;
    size_t want = bytes + bytes / 2 + 256; void* q = dev::alloc(want);
    if (used) { dev::d2d(q, p, used, s); dev::stream_sync(s); }
    dev::free_(p); p = q; cap = want;
  }
  void release() { dev::free_(p); p = nullptr; cap = 0; }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct PinnedBuf { void* p = nullptr; size_t cap = 0; };

}  // namespace cal

using namespace cal;

struct calitas_reference {
  calitas_engine* owner = nullptr;
  std::vector<std::string> names; std::vector<int64_t> len, have_b, have_e, own_b, own_e, nib_off;
  uint32_t* d_nib = nullptr; uint8_t* d_raw = nullptr; int64_t total_padded = 0;
  struct TileSet { std::vector<Tile> tiles; std::vector<ContigDev> contigs; Tile* d_tiles = nullptr; ContigDev* d_contigs = nullptr; int64_t n_windows = 0, total_windows = 0; int tile_windows = TILE_WINDOWS; };
  std::map<std::pair<int, int>, TileSet> tilesets;   // (window_size, step)
};

struct calitas_hitset {
  calitas_engine* owner = nullptr; PinnedBuf buf; int64_t n = 0; int32_t stride = CALITAS_HIT_WORDS * 4; PinnedBuf info; /* calitas_variant_hit_info per hit, search_variants only */ double ms[8] = { 0, 0, 0, 0, 0, 0, 0, 0 }; int64_t counts[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
};

// Variant windows of one VCF, resident on the device of `owner` (calitas_variant_set_load): packed bases, descriptors, alleles, variant-set names.
struct calitas_variant_set {
  calitas_engine* owner = nullptr; int64_t n_windows = 0, n_alleles = 0, n_sets = 0, nib_words = 0; uint32_t max_len = 1, max_rank = 0; int32_t n_contigs = 0;
  DBuf nib, windows, alleles, sets;
  std::vector<int64_t> class_begin, class_end;          // windows are sorted by guide_class
};

enum { CNT_GROUPS = 3, CNT_KEPT = 4, CNT_KEEPERS = 5 /* + 6: overflow flag as published */, CNT_DEDUP_OVERFLOW = 7 /* low word: key overflow; high word: halo sentinel */ };
enum { CE_SCAN_B = 0, CE_SCAN_E, CE_COUNT, CE_SORTED, CE_ALIGN_B, CE_ALIGN_E, CE_TAIL_B, CE_TAIL_E, CE_COPY_B, CE_COPY_E, CE_N };
struct ChunkEvents { dev::Event ev[CE_N]; };

struct calitas_engine {
  int device = 0; Scores sc; calitas_costs costs;
  // stream: sort/align/canonicalise/dedup (greatest priority); scan_stream: k_scan_tiled of calitas_search (least priority, so that the tail of
  // guide chunk c slips in between the blocks of chunk c+1's scan); copy_stream: D2H of finished hit segments
  dev::Stream stream, scan_stream, scan_stream2, copy_stream, pub_stream;   // scans alternate between the two scan streams: the next chunk's blocks fill the SMs the previous scan's last wave leaves idle
  dev::Event ev[8];
  std::vector<ChunkEvents> chunk_ev;
  size_t out_hits_hint = 1u << 16;
  DBuf cand_b, cand_c;
  DBuf specs, cand, cand_sorted, hits, valid, rank, perm, flag, pos, out, tmp, key1, keyA, idx, idx2, key_b, sstart, send, sscore, windows, nib_tmp, slot_owned, sowned, out_owned, var_info, var_info2, out2, vk1, vk2, vk3, vi1, vi2, vi3;
  // eight 64-bit device counters and their pinned host mirror: slots 0..2 = candidate counts of the three scan buffers (run_explicit uses 0),
  // slot CNT_DEDUP_OVERFLOW = k_dedup_keys' "a field does not fit its sort-key width" flag
  unsigned long long* h_count = nullptr;       // pinned + mapped: the device stores into it through h_count_dev (no copy engine involved)
  unsigned long long* h_count_dev = nullptr;   // device alias of h_count
  unsigned long long* d_count = nullptr;
  std::vector<PinnedBuf> pinned_pool;
  size_t cand_cap_hint = 1u << 20;
  int64_t launches = 0;
};

namespace {

thread_local std::string g_last_error;
int set_error(int code, const std::string& msg) { g_last_error = msg; return code; }

template <class F> int guarded(F f) {
  try { g_last_error.clear(); return f(); }
  catch (const InvalidArgument& e) { return set_error(CALITAS_EINVAL, e.what()); }
  catch (const LimitExceeded& e) { return set_error(CALITAS_ELIMIT, e.what()); }
  catch (const dev::Error& e) { return set_error(CALITAS_ECUDA, e.msg); }
  catch (const std::exception& e) { return set_error(CALITAS_ESTATE, e.what()); }
}

inline unsigned blocks_for(int64_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }

PinnedBuf take_pinned(calitas_engine* e, size_t bytes) {
  for (size_t i = 0; i < e->pinned_pool.size(); ++i) if (e->pinned_pool[i].cap >= bytes) { PinnedBuf b = e->pinned_pool[i]; e->pinned_pool.erase(e->pinned_pool.begin() + (long)i); return b; }
  PinnedBuf b; b.cap = bytes + bytes / 4 + 4096; b.p = dev::alloc_host(b.cap); return b;
}

// Dynamic shared memory of one k_scan_tiled CTA: mask tables + meta, the byte tile (word-aligned with the packed words), the staged packed
// words incl. the lead-in and rounding of the 16-byte-granular bulk copy.
size_t scan_smem_bytes(int ng, int tile_windows, int window_size, int step) {
  const int64_t tile_bases = (int64_t)(tile_windows - 1) * step + window_size;
  return (size_t)ng * 36 * 4 + (size_t)((tile_bases + 16 + 15) & ~15ll) + (size_t)((tile_bases + 7) / 8 + 2 + 8) * 4;
}

// Builds (or returns the cached) tiling of the owned reference windows for (window_size, step): SearchReference.scala:52-53.
calitas_reference::TileSet& tileset_for(calitas_engine* e, calitas_reference* r, int window_size, int step, const char* chrom) {
  auto key = std::make_pair(window_size, step);
  auto it = r->tilesets.find(key);
  if (it == r->tilesets.end()) {
    calitas_reference::TileSet ts;
    // windows per tile: as many (<= 64, power of two) as keep the byte tile + staged packed words + 16 guides' tables within shared memory
    while (ts.tile_windows > 1 && scan_smem_bytes(16, ts.tile_windows, window_size, step) > (size_t)SCAN_SMEM_LIMIT) ts.tile_windows /= 2;
    int64_t win_base = 0;
    for (size_t c = 0; c < r->len.size(); ++c) {
      const int64_t len = r->len[c];
      const int64_t n_win = len - 1 > 0 ? (len - 1 + step - 1) / step : 0;       // Range(0, len-1, step).size
      int64_t k_lo = (r->own_b[c] + step - 1) / step;
      int64_t k_hi = n_win;                                                      // exclusive
      if (r->own_e[c] < len) { int64_t lim = (r->own_e[c] + step - 1) / step; if (lim < k_hi) k_hi = lim; }
      if (k_lo > k_hi) k_lo = k_hi;
      ts.contigs.push_back(ContigDev{ len, r->nib_off[c] - r->have_b[c], win_base, win_base + k_lo, win_base + k_hi });
      // halo windows on each interior cut (removeOverlaps needs both sides of a cut; their hits are never reported)
      int64_t p_lo = k_lo, p_hi = k_hi;
      if (k_lo < k_hi) { if (k_lo > 0) p_lo = std::max<int64_t>(0, k_lo - HALO_WINDOWS); if (k_hi < n_win) p_hi = std::min<int64_t>(n_win, k_hi + HALO_WINDOWS); }
      if (p_lo < p_hi) {
        int64_t last_end = std::min(len, (p_hi - 1) * step + window_size);
        if (p_lo * step < r->have_b[c] || last_end > r->have_e[c])
          throw InvalidArgument("reference shard of contig " + r->names[c] + " does not hold the bases of its owned windows plus " + std::to_string(HALO_WINDOWS) + " halo windows per cut");
      }
      for (int64_t k = p_lo; k < p_hi; k += ts.tile_windows) ts.tiles.push_back(Tile{ (int32_t)c, (int32_t)std::min<int64_t>(ts.tile_windows, p_hi - k), k });
      ts.n_windows += std::max<int64_t>(0, k_hi - k_lo);
      win_base += n_win;
    }
    if (win_base >= (1ll << 32)) throw LimitExceeded("more than 2^32 reference windows");
    ts.total_windows = win_base;
    ts.d_tiles = (Tile*)dev::alloc(ts.tiles.size() * sizeof(Tile));
    ts.d_contigs = (ContigDev*)dev::alloc(ts.contigs.size() * sizeof(ContigDev));
    dev::h2d(ts.d_tiles, ts.tiles.data(), ts.tiles.size() * sizeof(Tile), e->stream);
    dev::h2d(ts.d_contigs, ts.contigs.data(), ts.contigs.size() * sizeof(ContigDev), e->stream);
    dev::stream_sync(e->stream);
    it = r->tilesets.emplace(key, std::move(ts)).first;
  }
  (void)chrom;
  return it->second;
}

struct Pipeline {   // tail shared by the tiled and explicit paths: sort -> align -> canon -> (removeOverlaps + sort | compaction) -> e->out
  calitas_engine* e; const uint64_t* cand; dev::Event ev_sorted, ev_align_b, ev_align_e;   // candidate keys; events recorded after the sort / around k_align
  const GuideSpec* d_specs; int slots; bool explicit_mode; int banded;   // banded: 0 = wide thresholds, else the largest k_edits of the launch (<= ALIGN_KB)
  const uint32_t* nib; const ContigDev* d_contigs; int n_contigs; int window_size; int step; const ExplicitWindow* d_windows; bool drop_halo;
  KeyLayout key; int rw;                                                   // rw: 32-bit words per hit record of this call
  int fast; int64_t nib_words;                                             // fast: 0 = no, 1 = every guide of the launch fits align_fast's 64-column window, 2 = and align_pair's 16-bit scores (align_level); nib_words: size of `nib`
  const DedupLayout* dedup; int32_t max_overlap;                           // dedup != nullptr: removeOverlaps + ReferenceHit.sort over the kept alignments
  DBuf* out_owned;                                                         // plain compaction only: when given, one byte per output record (1 = owned, 0 = halo) goes here
  bool presorted = false;                                                  // the candidate keys are already in (guide, window, strand, column) order
  int32_t halo_bases = 0;                                                  // k_sweep's halo sentinel: HALO_WINDOWS * step - window overlap - CALITAS_MAX_OPS
  bool any_wide_group = true;                                              // best mode: some group may not fit k_align_group_warp's tile (window > GROUP_W columns or alignments > 64 columns)
};

// align_fast<KB> keeps 64 target codes per candidate in registers: band + guide-PAM gap + longest PAM must fit, and the extension indexes 32 of them
bool fits_align_fast(const GuideSpec& sp, int banded) {
  const int kb = banded > 5 ? 6 : (banded == 5 ? 5 : 4);
  int pam = 0; for (int k = 0; k < sp.n_pams; ++k) pam = std::max<int>(pam, sp.pam_len[k]);
  return !std::getenv("CALITAS_NO_ALIGN_FAST") && sp.lp + kb + sp.g + pam <= 64;
}

// align_pair<KB> holds 4 * score + tag in 16 bits: see its header for the argument behind the second condition
bool pair_fits(const GuideSpec& sp, const Scores& sc) {
  const int32_t match4 = 4 * sc.match, big = 700;
  const int32_t adds[4] = { match4, 4 * sc.mismatch, 4 * sc.target_gap + TG_UP, 4 * sc.query_gap + TG_LEFT };
  for (int k = 0; k < 4; ++k) if (adds[k] > big || adds[k] < -big) return false;                    // PAIR_FLOOR + an addend stays inside 16 bits
  if (match4 <= 0 || (int64_t)match4 * sp.lp + 3 > 32767) return false;
  return (int64_t)PAIR_FLOOR + 3 + (int64_t)sp.lp * match4 < 4 * (int64_t)sp.min_score;
}
// 0: the general kernels; 1: align_fast; 2: align_pair (all guides of the launch of one protospacer length)
int align_level(const std::vector<GuideSpec>& specs, size_t g0, size_t g1, int banded, const Scores& sc) {
  if (banded <= 0 || g1 <= g0) return 0;
  int level = std::getenv("CALITAS_ALIGN_PAIR") ? 2 : 1;        // align_pair is exact but measured slower than align_fast (see its header): opt-in
  for (size_t g = g0; g < g1; ++g) {
    if (!fits_align_fast(specs[g], banded)) return 0;
    if (specs[g].lp != specs[g0].lp || !pair_fits(specs[g], sc)) level = 1;
  }
  return level;
}

// e->out must hold (out_n + more) records; the first out_n are kept
void ensure_out(calitas_engine* e, int rw, int64_t out_n, int64_t more, size_t projected_total) {
  const size_t need = (size_t)(out_n + more) * rw * 4;
  if (need > e->out.cap) e->out.ensure_keep(std::max(need, projected_total * rw * 4), (size_t)out_n * rw * 4, e->stream);
}

// Runs sort/align/canon on the candidate keys, then either removeOverlaps + sort (P.dedup) or a plain compaction, appending the surviving records to
// e->out at out_n (arrival order without dedup: group order, then retval order).  Returns their number; n_alignments = alignment slots examined.
int64_t run_tail(const Pipeline& P, int64_t n_cand, int64_t out_n, size_t projected_total, int64_t& n_alignments) {
  calitas_engine* e = P.e; dev::Stream s = e->stream; const int rw = P.rw;
  n_alignments = 0;
  if (n_cand == 0) { dev::event_record(P.ev_sorted, s); dev::event_record(P.ev_align_b, s); dev::event_record(P.ev_align_e, s); return 0; }
  if (n_cand * (int64_t)P.slots >= (1ll << 32)) throw LimitExceeded("too many candidate alignments in one batch");
  // 1. sort candidate keys -> (guide, window, strand, end column)
  size_t tb = 0;
  if (!P.presorted) {
    e->cand_sorted.ensure((size_t)n_cand * 8);
    tb = dev::sort_keys_u64_tmp((size_t)n_cand, 0, P.key.bits); e->tmp.ensure(tb);
    dev::sort_keys_u64(e->tmp.p, tb, P.cand, e->cand_sorted.as<uint64_t>(), (size_t)n_cand, 0, P.key.bits, s); ++e->launches;
  }
  dev::event_record(P.ev_sorted, s);                      // the candidate buffer may be refilled by the next scan from here on
  // 2. align
  const int64_t n_slots = n_cand * P.slots;
  e->hits.ensure((size_t)n_slots * rw * 4); e->valid.ensure((size_t)n_slots * sizeof(CKey));
  AlignArgs aa; std::memset(&aa, 0, sizeof aa);
  aa.cand = P.presorted ? P.cand : e->cand_sorted.as<uint64_t>(); aa.n_cand = n_cand; aa.specs = P.d_specs; aa.sc = e->sc; aa.slots = P.slots; aa.explicit_mode = P.explicit_mode ? 1 : 0;
  aa.nib = P.nib; aa.contigs = P.d_contigs; aa.n_contigs = P.n_contigs; aa.window_size = P.window_size; aa.step = P.step; aa.windows = P.d_windows;
  aa.recs = e->hits.as<uint32_t>(); aa.rw = rw; aa.ckeys = e->valid.as<CKey>(); aa.key = P.key; aa.nib_last_word = P.nib_words - 1;
  dev::event_record(P.ev_align_b, s);
  aa.match4 = 4 * e->sc.match; aa.mis4 = 4 * e->sc.mismatch; aa.up4 = 4 * e->sc.target_gap + TG_UP; aa.left4 = 4 * e->sc.query_gap + TG_LEFT;
  if (std::getenv("CALITAS_TRACE")) std::fprintf(stderr, "[calitas trace] align: %lld candidates, banded %d, level %d (2 = align_pair, 1 = align_fast)\n", (long long)n_cand, P.banded, P.fast);
  if (P.banded > 0 && P.fast == 2) {
    const unsigned blocks = blocks_for((n_cand + 1) / 2, 128);
    if (P.banded > 5) { CAL_LAUNCH(k_align_pair6, blocks, 128, 0, s, 1, aa); dev::launch_check("k_align_pair6"); }
    else if (P.banded == 5) { CAL_LAUNCH(k_align_pair5, blocks, 128, 0, s, 1, aa); dev::launch_check("k_align_pair5"); }
    else { CAL_LAUNCH(k_align_pair4, blocks, 128, 0, s, 1, aa); dev::launch_check("k_align_pair4"); }
  }
  else if (P.banded > 0 && P.fast) {
    if (P.banded > 5) { CAL_LAUNCH(k_align_fast6, blocks_for(n_cand, 128), 128, 0, s, 1, aa); dev::launch_check("k_align_fast6"); }
    else if (P.banded == 5) { CAL_LAUNCH(k_align_fast5, blocks_for(n_cand, 128), 128, 0, s, 1, aa); dev::launch_check("k_align_fast5"); }
    else { CAL_LAUNCH(k_align_fast4, blocks_for(n_cand, 128), 128, 0, s, 1, aa); dev::launch_check("k_align_fast4"); }
  }
  else if (P.banded > 5) { CAL_LAUNCH(k_align, blocks_for(n_cand, 128), 128, 0, s, 1, aa); dev::launch_check("k_align"); }
  else if (P.banded == 5) { CAL_LAUNCH(k_align5, blocks_for(n_cand, 128), 128, 0, s, 1, aa); dev::launch_check("k_align5"); }
  else if (P.banded > 0) { CAL_LAUNCH(k_align4, blocks_for(n_cand, 128), 128, 0, s, 1, aa); dev::launch_check("k_align4"); }
  else if (P.explicit_mode) {             // short windows, nearly every column a candidate: one DP fill per (window, strand)
    e->key1.ensure((size_t)n_cand * 4); e->keyA.ensure((size_t)n_cand * 4); e->idx.ensure((size_t)n_cand * 4);   // group flag / group index / group start (u32 views)
    uint32_t* gflag = e->key1.as<uint32_t>(); uint32_t* gpos = e->keyA.as<uint32_t>();
    CAL_LAUNCH(k_mark_groups, blocks_for(n_cand, 128), 128, 0, s, 1, aa.cand, n_cand, P.key.col_bits, gflag); dev::launch_check("k_mark_groups"); ++e->launches;
    size_t tb2 = dev::exclusive_sum_u32_tmp((size_t)n_cand); e->tmp.ensure(tb2);
    dev::exclusive_sum_u32(e->tmp.p, tb2, gflag, gpos, (size_t)n_cand, s); ++e->launches;
    CAL_LAUNCH(k_publish_sum, 1, 1, 0, s, 1, gpos + (n_cand - 1), gflag + (n_cand - 1), (const unsigned long long*)nullptr, e->h_count_dev + CNT_GROUPS); dev::launch_check("k_publish_sum"); dev::stream_sync(s);
    const int64_t n_groups = (int64_t)((volatile unsigned long long*)e->h_count)[CNT_GROUPS];
    CAL_LAUNCH(k_group_starts, blocks_for(n_cand, 128), 128, 0, s, 1, gflag, gpos, n_cand, e->idx.as<uint32_t>()); dev::launch_check("k_group_starts"); ++e->launches;
    dev::event_record(P.ev_align_b, s);
#ifndef CAL_HOSTSIM
    if (!std::getenv("CALITAS_NO_GROUP_WARP")) {          // warp per group; what does not fit its shared-memory tile is flagged and goes through the thread-per-group kernel
      e->idx2.ensure((size_t)n_groups * 4);
      CAL_LAUNCH(k_align_group_warp, (unsigned)dev::sm_count(e->device) * 8, 32 * GROUP_WARPS, 0, s, 1, aa, e->idx.as<uint32_t>(), n_groups, e->idx2.as<uint32_t>()); dev::launch_check("k_align_group_warp"); ++e->launches;
      if (P.any_wide_group) { CAL_LAUNCH(k_align_group, blocks_for(n_groups, 128), 128, 0, s, 1, aa, e->idx.as<uint32_t>(), n_groups, (const uint32_t*)e->idx2.as<uint32_t>()); dev::launch_check("k_align_group"); }
      else --e->launches;
    } else
#endif
    { CAL_LAUNCH(k_align_group, blocks_for(n_groups, 128), 128, 0, s, 1, aa, e->idx.as<uint32_t>(), n_groups, (const uint32_t*)nullptr); dev::launch_check("k_align_group"); }
  }
  else { CAL_LAUNCH(k_align_wide, blocks_for(n_cand, 128), 128, 0, s, 1, aa); dev::launch_check("k_align_wide"); }
  ++e->launches;
  dev::event_record(P.ev_align_e, s);
  // 3. canonicalise per (guide, window, strand)
  e->rank.ensure((size_t)n_slots * 4); e->perm.ensure((size_t)n_slots * 4); e->flag.ensure((size_t)n_slots * 4); e->pos.ensure((size_t)n_slots * 4); e->slot_owned.ensure((size_t)n_slots);
  CanonArgs ca; std::memset(&ca, 0, sizeof ca);
  ca.cand = aa.cand; ca.n_cand = n_cand; ca.specs = P.d_specs; ca.slots = P.slots; ca.explicit_mode = aa.explicit_mode; ca.windows = P.d_windows;
  ca.ckeys = aa.ckeys; ca.rank = e->rank.as<int32_t>(); ca.gbase = e->perm.as<uint32_t>(); ca.flag = e->flag.as<uint32_t>(); ca.slot_owned = e->slot_owned.as<uint8_t>(); ca.drop_halo = P.drop_halo ? 1 : 0; ca.key = P.key;
#ifndef CAL_HOSTSIM
  if (P.explicit_mode && !P.banded) {     // large groups: a warp per group (group starts, indices and flags were built for k_align_group)
    CAL_LAUNCH(k_canon_warp, (unsigned)dev::sm_count(e->device) * 16, 128, 0, s, 1, ca, e->idx.as<uint32_t>(), e->keyA.as<uint32_t>(), e->key1.as<uint32_t>()); dev::launch_check("k_canon_warp"); ++e->launches;
  } else
#endif
  {
    CAL_LAUNCH(k_canon, blocks_for(n_cand, 128), 128, 0, s, 1, ca); dev::launch_check("k_canon"); ++e->launches;
    CAL_LAUNCH(k_canon_rank, blocks_for(n_cand, 128), 128, 0, s, 1, ca); dev::launch_check("k_canon_rank"); ++e->launches;
  }
  tb = dev::exclusive_sum_u32_tmp((size_t)n_slots); e->tmp.ensure(tb);
  dev::exclusive_sum_u32(e->tmp.p, tb, e->flag.as<uint32_t>(), e->pos.as<uint32_t>(), (size_t)n_slots, s); ++e->launches;
  CAL_LAUNCH(k_publish_sum, 1, 1, 0, s, 1, e->pos.as<uint32_t>() + (n_slots - 1), e->flag.as<uint32_t>() + (n_slots - 1), (const unsigned long long*)nullptr, e->h_count_dev + CNT_KEPT); dev::launch_check("k_publish_sum");
  dev::stream_sync(s);
  const int64_t n = (int64_t)((volatile unsigned long long*)e->h_count)[CNT_KEPT];      // kept alignments
  n_alignments = n_slots;
  if (n == 0) return 0;
  if (!P.dedup) {                          // plain compaction straight into the output
    ensure_out(e, rw, out_n, n, projected_total);
    uint8_t* oo = nullptr;
    if (P.out_owned) { P.out_owned->ensure_keep((size_t)(out_n + n), (size_t)out_n, s); oo = P.out_owned->as<uint8_t>() + out_n; }
    CAL_LAUNCH(k_gather_flagged, blocks_for(n_slots, 256), 256, 0, s, 1, aa.recs, rw, e->rank.as<int32_t>(), e->perm.as<uint32_t>(), e->flag.as<uint32_t>(), e->pos.as<uint32_t>(), n_slots, e->out.as<uint32_t>() + out_n * rw,
               (const uint8_t*)e->slot_owned.as<uint8_t>(), oo); dev::launch_check("k_gather_flagged"); ++e->launches;
    return n;
  }
  // 4. removeOverlaps (SearchReference.scala:653-675) + ReferenceHit.sort: keys of the kept slots, ONE stable radix sort (two when the fields do not fit
  //    64 bits), the two-strand sweep, and only then a copy: the keepers go from their alignment slots straight to the output.
  const DedupLayout& L = *P.dedup;
  if (n >= (1ll << 32)) throw LimitExceeded("too many hits in one batch");
  e->key1.ensure((size_t)n * 8); e->keyA.ensure((size_t)n * 8); e->idx.ensure((size_t)n * 4); e->idx2.ensure((size_t)n * 4);
  tb = dev::sort_pairs_u64_tmp((size_t)n, 0, 64); e->tmp.ensure(tb);
  const uint64_t* skey; const uint32_t* sidx; int strand_shift;
  uint32_t* d_overflow = (uint32_t*)(e->d_count + CNT_DEDUP_OVERFLOW);
  if (L.merged) {
    CAL_LAUNCH(k_dedup_keys, blocks_for(n_slots, 256), 256, 0, s, 1, aa.recs, rw, e->rank.as<int32_t>(), e->perm.as<uint32_t>(), e->flag.as<uint32_t>(), e->pos.as<uint32_t>(), n_slots, L, e->keyA.as<uint64_t>(), (uint64_t*)nullptr, e->idx.as<uint32_t>(), d_overflow); dev::launch_check("k_dedup_keys"); ++e->launches;
    dev::sort_pairs_u64(e->tmp.p, tb, e->keyA.as<uint64_t>(), e->key1.as<uint64_t>(), e->idx.as<uint32_t>(), e->idx2.as<uint32_t>(), (size_t)n, 0, L.bits + L.score_bits, s); ++e->launches;
    skey = e->key1.as<uint64_t>(); sidx = e->idx2.as<uint32_t>(); strand_shift = L.score_bits;
  } else {      // stable by -score (arrival order = input order), then stable by (guide, contig, start, strand)
    e->key_b.ensure((size_t)n * 8); e->sstart.ensure((size_t)n * 4);
    uint32_t* ord = e->sstart.as<uint32_t>();    // positions 0..n-1, carried through the first sort to permute the second sort's keys
    CAL_LAUNCH(k_dedup_keys, blocks_for(n_slots, 256), 256, 0, s, 1, aa.recs, rw, e->rank.as<int32_t>(), e->perm.as<uint32_t>(), e->flag.as<uint32_t>(), e->pos.as<uint32_t>(), n_slots, L, e->key1.as<uint64_t>(), e->keyA.as<uint64_t>(), e->idx.as<uint32_t>(), d_overflow); dev::launch_check("k_dedup_keys"); ++e->launches;
    CAL_LAUNCH(k_iota_u32, blocks_for(n, 256), 256, 0, s, 1, ord, n); dev::launch_check("k_iota_u32"); ++e->launches;
    dev::sort_pairs_u64(e->tmp.p, tb, e->keyA.as<uint64_t>(), e->key_b.as<uint64_t>(), ord, e->idx2.as<uint32_t>(), (size_t)n, 0, L.score_bits, s); ++e->launches;
    CAL_LAUNCH(k_gather_u64, blocks_for(n, 256), 256, 0, s, 1, e->key1.as<uint64_t>(), e->idx2.as<uint32_t>(), n, e->keyA.as<uint64_t>()); dev::launch_check("k_gather_u64"); ++e->launches;
    CAL_LAUNCH(k_gather_u32, blocks_for(n, 256), 256, 0, s, 1, e->idx.as<uint32_t>(), e->idx2.as<uint32_t>(), n, ord); dev::launch_check("k_gather_u32"); ++e->launches;
    dev::sort_pairs_u64(e->tmp.p, tb, e->keyA.as<uint64_t>(), e->key1.as<uint64_t>(), ord, e->idx.as<uint32_t>(), (size_t)n, 0, L.bits, s); ++e->launches;
    skey = e->key1.as<uint64_t>(); sidx = e->idx.as<uint32_t>(); strand_shift = 0;
  }
  // skey[i], sidx[i] (= alignment slot) sorted by (guide, contig, start, strand, -score, arrival)
  e->sstart.ensure((size_t)n * 4); e->send.ensure((size_t)n * 4); e->sscore.ensure((size_t)n * 4); e->sowned.ensure((size_t)n);
  e->key_b.ensure((size_t)n * 4);
  uint32_t* keep = e->rank.as<uint32_t>(); uint32_t* kpos = e->key_b.as<uint32_t>();   // rank (one word per slot) is free again after k_canon
  CAL_LAUNCH(k_sweep_prepare, blocks_for(n, 256), 256, 0, s, 1, aa.recs, rw, e->slot_owned.as<uint8_t>(), sidx, n, e->sstart.as<int32_t>(), e->send.as<int32_t>(), e->sscore.as<int32_t>(), e->sowned.as<uint8_t>()); dev::launch_check("k_sweep_prepare"); ++e->launches;
  CAL_LAUNCH(k_sweep, blocks_for(n, 128), 128, 0, s, 1, skey, e->sstart.as<int32_t>(), e->send.as<int32_t>(), e->sscore.as<int32_t>(), e->sowned.as<uint8_t>(), n, P.max_overlap, P.max_overlap >= 1 ? 1 : 0, strand_shift, L.start_bits, keep, d_overflow + 1, P.halo_bases); dev::launch_check("k_sweep"); ++e->launches;
  tb = dev::exclusive_sum_u32_tmp((size_t)n); e->tmp.ensure(tb);
  dev::exclusive_sum_u32(e->tmp.p, tb, keep, kpos, (size_t)n, s); ++e->launches;
  CAL_LAUNCH(k_publish_sum, 1, 1, 0, s, 1, kpos + (n - 1), keep + (n - 1), (const unsigned long long*)(e->d_count + CNT_DEDUP_OVERFLOW), e->h_count_dev + CNT_KEEPERS); dev::launch_check("k_publish_sum");
  dev::stream_sync(s);
  { const unsigned long long fl = ((volatile unsigned long long*)e->h_count)[CNT_KEEPERS + 1];
    if (fl & 0xFFFFFFFFull) throw std::runtime_error("internal error: a hit field exceeds its sort-key width in removeOverlaps");
    if (fl >> 32) throw LimitExceeded("removeOverlaps: a chain of overlapping hits runs through the whole halo of a shard cut; this input cannot be de-duplicated per shard (search with dedup = 0 and run removeOverlaps over the gathered hits, or use fewer shards)"); }
  const int64_t nk = (int64_t)((volatile unsigned long long*)e->h_count)[CNT_KEEPERS];
  if (nk == 0) return 0;
  ensure_out(e, rw, out_n, nk, projected_total);
  CAL_LAUNCH(k_gather_keepers, blocks_for(n, 256), 256, 0, s, 1, aa.recs, rw, sidx, keep, kpos, n, e->out.as<uint32_t>() + out_n * rw); dev::launch_check("k_gather_keepers"); ++e->launches;
  return nk;
}

calitas_hitset* finish_hitset(calitas_engine* e, int64_t n_out, int rw, const double ms[8], const int64_t counts[8]) {
  std::unique_ptr<calitas_hitset> hs(new calitas_hitset());
  hs->owner = e; hs->n = n_out; hs->stride = rw * 4;
  hs->buf = take_pinned(e, (size_t)std::max<int64_t>(1, n_out) * rw * 4);
  dev::event_record(e->ev[6], e->stream);
  dev::d2h(hs->buf.p, e->out.p, (size_t)n_out * rw * 4, e->stream);
  dev::event_record(e->ev[1], e->stream);
  dev::stream_sync(e->stream);
  for (int i = 0; i < 8; ++i) { hs->ms[i] = ms[i]; hs->counts[i] = counts[i]; }
  hs->ms[0] = dev::event_ms(e->ev[0], e->ev[1]);
  hs->ms[4] = dev::event_ms(e->ev[6], e->ev[1]);
  hs->ms[3] = hs->ms[0] - hs->ms[1] - hs->ms[2] - hs->ms[4];
  hs->counts[5] += (int64_t)((size_t)n_out * rw * 4);
  return hs.release();
}

std::vector<GuideSpec> build_specs(calitas_engine* e, int32_t n_guides, const calitas_guide* guides, const calitas_limits* limits, bool best, std::vector<GuideDef>* defs_out) {
  if (n_guides <= 0 || !guides) throw InvalidArgument("no guides given");
  if (n_guides > MAX_GUIDES_PER_CALL) throw LimitExceeded("more than " + std::to_string(MAX_GUIDES_PER_CALL) + " guides in one call");
  if (!limits) throw InvalidArgument("limits is NULL");
  std::vector<GuideSpec> specs;
  for (int i = 0; i < n_guides; ++i) { GuideDef d = parse_guide(guides[i]); specs.push_back(make_guide_spec(d, e->sc, *limits, best)); if (defs_out) defs_out->push_back(d); }
  return specs;
}

// explicit-window path shared by align_regions / align_targets
// Shape of a launch over a set of guides: alignment slots per candidate, which align kernel, record size.
struct LaunchShape { int slots, banded, rw; int fast; bool all_columns; /* every guide's threshold admits every column (best mode) */ };
LaunchShape launch_shape(const std::vector<GuideSpec>& specs, size_t g0, size_t g1, int rw_all, const Scores& sc) {
  LaunchShape L{ 1, 1, rw_all, 0, g1 > g0 };
  for (size_t g = g0; g < g1; ++g) { L.slots = std::max(L.slots, specs[g].slots); L.banded = std::max(L.banded, std::max(specs[g].k_edits, specs[g].band_k)); }
  if (L.banded > ALIGN_KB) L.banded = 0;
  L.fast = align_level(specs, g0, g1, L.banded, sc);
  for (size_t g = g0; g < g1; ++g) L.all_columns = L.all_columns && specs[g].k_edits >= specs[g].lp;
  if (L.banded > 0 || std::getenv("CALITAS_NO_ALL_COLUMNS")) L.all_columns = false;
  return L;
}
int rec_words_of(const std::vector<GuideSpec>& specs) { int m = 1; for (auto& sp : specs) m = std::max(m, sp.max_cols); return rec_words_for(m); }

// Explicit windows already on the device (e->windows[0..n_windows), guide table in e->specs): scan + tail in batches, the kept alignments are appended
// to e->out at n_out in (window, strand, retval) order.  drop_halo / out_owned as in Pipeline.
void explicit_core(calitas_engine* e, const uint32_t* d_nib, int64_t nib_words, int64_t n_windows, uint32_t max_len, const LaunchShape& L, bool drop_halo, DBuf* out_owned,
                   int64_t& n_out, double ms[8], int64_t counts[8]) {
  dev::Stream s = e->stream;
  // batches bound the candidate buffer: in best mode every column of every window is a candidate
  const int64_t batch = std::max<int64_t>(1, std::min<int64_t>(n_windows, L.banded ? (1 << 25) : (1 << 18)));
  const KeyLayout key = make_key_layout(max_len, (uint64_t)batch, 1);       // window ids are relative to the batch; the guide comes from the window
  for (int64_t w0 = 0; w0 < n_windows; w0 += batch) {
    const int64_t nw = std::min(batch, n_windows - w0);
    unsigned long long n_cand = 0;
    if (L.all_columns) {
      e->flag.ensure((size_t)(2 * nw) * 4); e->pos.ensure((size_t)(2 * nw) * 4);
      dev::event_record(e->ev[4], s);
      CAL_LAUNCH(k_window_columns, blocks_for(2 * nw, 256), 256, 0, s, 1, e->windows.as<ExplicitWindow>() + w0, nw, e->flag.as<uint32_t>()); dev::launch_check("k_window_columns"); ++e->launches;
      size_t tb0 = dev::exclusive_sum_u32_tmp((size_t)(2 * nw)); e->tmp.ensure(tb0);
      dev::exclusive_sum_u32(e->tmp.p, tb0, e->flag.as<uint32_t>(), e->pos.as<uint32_t>(), (size_t)(2 * nw), s); ++e->launches;
      CAL_LAUNCH(k_publish_sum, 1, 1, 0, s, 1, e->pos.as<uint32_t>() + (2 * nw - 1), e->flag.as<uint32_t>() + (2 * nw - 1), (const unsigned long long*)nullptr, e->h_count_dev); dev::launch_check("k_publish_sum");
      dev::stream_sync(s);
      n_cand = *(volatile unsigned long long*)e->h_count;
      if (n_cand >= (1ull << 32)) throw LimitExceeded("too many candidate columns in one batch");
      e->cand.ensure((size_t)std::max<unsigned long long>(n_cand, 1) * 8);
      CAL_LAUNCH(k_all_columns, blocks_for(2 * nw, 256), 256, 0, s, 1, e->windows.as<ExplicitWindow>() + w0, nw, e->pos.as<uint32_t>(), key, e->cand.as<uint64_t>()); dev::launch_check("k_all_columns"); ++e->launches;
      dev::event_record(e->ev[5], s);
      dev::event_sync(e->ev[5]);                            // its time is read right below
    } else
    for (;;) {
      e->cand.ensure(e->cand_cap_hint * 8);
      dev::zero(e->d_count, 8, s);
      ScanExplicitArgs sa{ d_nib, e->windows.as<ExplicitWindow>() + w0, nw, e->specs.as<GuideSpec>(), e->cand.as<uint64_t>(), e->d_count, (unsigned long long)e->cand_cap_hint, key };
      dev::event_record(e->ev[4], s);
      CAL_LAUNCH(k_scan_explicit, blocks_for(2 * nw, 128), 128, 256, s, 2, sa); dev::launch_check("k_scan_explicit"); ++e->launches;
      dev::event_record(e->ev[5], s);
      CAL_LAUNCH(k_publish_u64, 1, 1, 0, s, 1, e->d_count, e->h_count_dev); dev::launch_check("k_publish_u64"); dev::stream_sync(s);
      n_cand = *(volatile unsigned long long*)e->h_count;
      if (n_cand <= e->cand_cap_hint) break;
      e->cand_cap_hint = (size_t)(n_cand + n_cand / 4);
    }
    ms[1] += dev::event_ms(e->ev[4], e->ev[5]);
    counts[1] += (int64_t)n_cand;
    // window ids inside the batch are relative to w0: rebase through the pointer passed to the tail
    Pipeline P{ e, e->cand.as<uint64_t>(), e->ev[7], e->ev[2], e->ev[3], e->specs.as<GuideSpec>(), L.slots, true, L.banded, d_nib, nullptr, 0, 0, 0, e->windows.as<ExplicitWindow>() + w0, drop_halo, key, L.rw, L.fast, nib_words, nullptr, 0, out_owned };
    P.presorted = L.all_columns;
    P.any_wide_group = max_len > (uint32_t)GROUP_W || L.rw > CALITAS_HIT_WORDS;          // record size > 32 bytes <=> some guide can produce more than 48 columns (the warp kernel takes 64)
    int64_t n_aln = 0;
    const int64_t n_kept = run_tail(P, (int64_t)n_cand, n_out, 0, n_aln);
    if (n_cand) ms[2] += dev::event_ms(e->ev[2], e->ev[3]);
    counts[2] += n_aln;
    n_out += n_kept;
  }
}

// explicit-window path shared by align_regions / align_targets
calitas_hitset* run_explicit(calitas_engine* e, const uint32_t* d_nib, int64_t nib_words, const std::vector<ExplicitWindow>& windows, const std::vector<GuideSpec>& specs) {
  dev::Stream s = e->stream; dev::set_device(e->device);
  e->launches = 0;
  dev::event_record(e->ev[0], s);
  const LaunchShape L = launch_shape(specs, 0, specs.size(), rec_words_of(specs), e->sc);
  e->specs.ensure(specs.size() * sizeof(GuideSpec)); dev::h2d(e->specs.p, specs.data(), specs.size() * sizeof(GuideSpec), s);
  e->windows.ensure(windows.size() * sizeof(ExplicitWindow)); dev::h2d(e->windows.p, windows.data(), windows.size() * sizeof(ExplicitWindow), s);
  double ms[8] = { 0, 0, 0, 0, 0, 0, 0, 0 }; int64_t counts[8] = { (int64_t)windows.size(), 0, 0, 0, 0, 0, 0, 0 };
  counts[4] = (int64_t)(specs.size() * sizeof(GuideSpec) + windows.size() * sizeof(ExplicitWindow));
  int64_t n_out = 0;
  uint32_t max_len = 1; for (auto& w : windows) if (w.len > 0 && (uint32_t)w.len > max_len) max_len = (uint32_t)w.len;
  explicit_core(e, d_nib, nib_words, (int64_t)windows.size(), max_len, L, true, nullptr, n_out, ms, counts);
  counts[3] = e->launches;
  return finish_hitset(e, n_out, L.rw, ms, counts);
}

// Score bounds of the hits a set of guides can produce (for the width of the score field in the sort keys): a hit's score = guide alignment
// (min_score ... lp rows each worth at most a match or an inserted base) + PAM bases + offset * queryGap.
void score_bounds(const std::vector<GuideSpec>& specs, size_t g0, size_t g1, const Scores& sc, int32_t& score_hi, int32_t& score_bits) {
  int64_t hi = 0, lo = 0; bool first = true;
  for (size_t g = g0; g < g1; ++g) {
    const GuideSpec& sp = specs[g];
    int pam_len = 0; for (int k = 0; k < sp.n_pams; ++k) pam_len = std::max<int>(pam_len, sp.pam_len[k]);
    const int64_t per_row = std::max<int64_t>(std::max<int64_t>(sc.match, sc.target_gap), 0), per_pam = std::max<int64_t>(iabs(sc.pam_match), iabs(sc.pam_mismatch));
    const int64_t h = per_row * sp.lp + per_pam * pam_len + (int64_t)std::max(0, sp.g) * std::max<int64_t>(sc.query_gap, 0);
    const int64_t l = (int64_t)sp.min_score - per_pam * pam_len - (int64_t)std::max(0, sp.g) * iabs(sc.query_gap);
    if (first || h > hi) hi = h;
    if (first || l < lo) lo = l;
    first = false;
  }
  if (hi - lo >= (1ll << 31) || hi > 0x7FFFFFFFll || hi < -0x7FFFFFFFll) { score_hi = 0x7FFFFFFF; score_bits = 33; }   // 0x7FFFFFFF - score fits 33 bits for any int32 score
  else { score_hi = (int32_t)hi; score_bits = std::max(1, bit_length((uint64_t)(hi - lo))); }
}

// Variant windows of a SearchReference -v run (SearchReference.scala:570-630): aligned like any explicit target, then merged with the n_ref reference
// hits already in e->out (arrival order; e->out_owned says which are reported), de-duplicated per (guide, contig, strand, variant set) and put in
// ReferenceHit.sort order (:641-648, 653-675).  Result: records in e->out2, annotations in e->var_info2; returns their number.
int64_t merge_variant_hits(calitas_engine* e, const VariantPlan& vp, const std::vector<GuideSpec>& specs, const calitas_limits* limits, const calitas_reference* ref, int rw, int64_t n_ref,
                           int32_t halo_bases, double ms[8], int64_t counts[8]) {
  dev::Stream s = e->stream;
  const calitas_variant_set& vs = *vp.vs;
  if (vs.owner != e) throw InvalidArgument("the variant set belongs to another engine");
  if (vs.n_contigs != (int32_t)ref->len.size()) throw InvalidArgument("the variant set was loaded for another reference");
  int64_t n = n_ref;
  if (vs.n_windows > 0) {
    // ---- tasks: every guide against the windows of its class, built on the device for spans of guides that keep a task list under 2^25 entries ------
    const LaunchShape L = launch_shape(specs, 0, specs.size(), rw, e->sc);
    const int64_t TASK_SPAN = (int64_t)1 << 25;
    std::vector<VarTaskGuide> span; int64_t span_tasks = 0;
    auto flush = [&]() {
      if (span_tasks == 0) { span.clear(); return; }
      e->var_info2.ensure(span.size() * sizeof(VarTaskGuide)); dev::h2d(e->var_info2.p, span.data(), span.size() * sizeof(VarTaskGuide), s);
      e->windows.ensure((size_t)span_tasks * sizeof(ExplicitWindow));
      CAL_LAUNCH(k_variant_tasks, blocks_for(span_tasks, 256), 256, 0, s, 1, vs.windows.as<VarWindowDev>(), e->var_info2.as<VarTaskGuide>(), (int32_t)span.size(), span_tasks, e->windows.as<ExplicitWindow>());
      dev::launch_check("k_variant_tasks"); ++e->launches;
      dev::stream_sync(s);                                 // `span` is reused
      counts[4] += (int64_t)(span.size() * sizeof(VarTaskGuide));
      explicit_core(e, vs.nib.as<uint32_t>(), vs.nib_words, span_tasks, vs.max_len, L, false, &e->out_owned, n, ms, counts);
      span.clear(); span_tasks = 0;
    };
    for (size_t g = 0; g < specs.size(); ++g) {
      const int32_t c = vp.guide_class[g];
      if (c < 0 || c >= (int32_t)vs.class_begin.size()) continue;                    // no windows of that class
      const int64_t nw = vs.class_end[(size_t)c] - vs.class_begin[(size_t)c];
      if (nw == 0) continue;
      if (span_tasks + nw > TASK_SPAN && span_tasks > 0) flush();
      span.push_back(VarTaskGuide{ span_tasks, (int32_t)g, (int32_t)vs.class_begin[(size_t)c] });
      span_tasks += nw;
    }
    flush();
  }
  if (n == 0) return 0;
  if (n >= (1ll << 32)) throw LimitExceeded("too many hits in one batch");
  // ---- keys: sweep order = (guide, contig, variant set | start, strand, -score, arrival) via two stable sorts, minor key first ---------------------------
  VarKeyLayout V; DedupLayout F;
  {
    const uint32_t max_rank = vs.max_rank;
    int64_t max_clen = 1; for (int64_t l : ref->len) max_clen = std::max(max_clen, l);
    V.set_bits = std::max(1, bit_length(max_rank)); V.contig_bits = std::max(1, bit_length((uint64_t)(ref->len.size() - 1))); V.start_bits = std::min(31, bit_length((uint64_t)max_clen));
    score_bounds(specs, 0, specs.size(), e->sc, V.score_hi, V.score_bits);
    const int guide_bits = std::max(1, bit_length((uint64_t)(specs.size() - 1)));
    if (guide_bits + V.contig_bits + V.set_bits > 64 || V.start_bits + 1 + V.score_bits > 64) throw LimitExceeded("variant sort keys do not fit 64 bits");
    F.start_bits = V.start_bits; F.contig_shift = F.start_bits + 1; F.guide_shift = F.contig_shift + V.contig_bits; F.g0 = 0; F.bits = F.guide_shift + guide_bits;
    F.score_hi = V.score_hi; F.score_bits = V.score_bits; F.merged = 1;
    if (F.bits + F.score_bits > 64) throw LimitExceeded("final sort key does not fit 64 bits");
  }
  const size_t nn = (size_t)n;
  e->var_info.ensure(nn * sizeof(VarInfo)); e->vk1.ensure(nn * 8); e->vk2.ensure(nn * 8); e->vk3.ensure(nn * 8); e->vi1.ensure(nn * 4); e->vi2.ensure(nn * 4); e->vi3.ensure(nn * 4);
  e->out_owned.ensure_keep(nn, nn, s);
  uint32_t* d_overflow = (uint32_t*)(e->d_count + CNT_DEDUP_OVERFLOW);
  const uint32_t* recs = e->out.as<uint32_t>();
  uint64_t* major = e->vk1.as<uint64_t>(); uint64_t* minor = e->vk2.as<uint64_t>(); uint64_t* ktmp = e->vk3.as<uint64_t>();
  uint32_t* i1 = e->vi1.as<uint32_t>(); uint32_t* i2 = e->vi2.as<uint32_t>(); uint32_t* i3 = e->vi3.as<uint32_t>();
  CAL_LAUNCH(k_variant_keys, blocks_for(n, 256), 256, 0, s, 1, recs, rw, n_ref, n, vs.windows.as<VarWindowDev>(), vs.alleles.as<calitas_variant_allele>(), vs.sets.as<uint32_t>(), V,
             e->var_info.as<VarInfo>(), major, minor, i1, d_overflow); dev::launch_check("k_variant_keys"); ++e->launches;
  size_t tb = dev::sort_pairs_u64_tmp(nn, 0, 64); e->tmp.ensure(tb);
  dev::sort_pairs_u64(e->tmp.p, tb, minor, ktmp, i1, i2, nn, 0, V.start_bits + 1 + V.score_bits, s); ++e->launches;          // i2: merged-list index, in minor order
  CAL_LAUNCH(k_gather_u64, blocks_for(n, 256), 256, 0, s, 1, major, i2, n, minor); dev::launch_check("k_gather_u64"); ++e->launches;                 // minor := major keys in minor order
  const int guide_bits = F.bits - F.guide_shift;
  dev::sort_pairs_u64(e->tmp.p, tb, minor, ktmp, i2, i3, nn, 0, guide_bits + V.contig_bits + V.set_bits, s); ++e->launches;  // ktmp: sorted major keys, i3: merged-list index
  // ---- sweep ------------------------------------------------------------------------------------------------------------------------------------
  e->sstart.ensure(nn * 4); e->send.ensure(nn * 4); e->sscore.ensure(nn * 4); e->sowned.ensure(nn); e->flag.ensure(nn * 4); e->pos.ensure(nn * 4);
  CAL_LAUNCH(k_variant_sweep_prepare, blocks_for(n, 256), 256, 0, s, 1, recs, rw, e->var_info.as<VarInfo>(), e->out_owned.as<uint8_t>(), ktmp, i3, n, major,
             e->sstart.as<int32_t>(), e->send.as<int32_t>(), e->sscore.as<int32_t>(), e->sowned.as<uint8_t>()); dev::launch_check("k_variant_sweep_prepare"); ++e->launches;
  CAL_LAUNCH(k_sweep, blocks_for(n, 128), 128, 0, s, 1, major, e->sstart.as<int32_t>(), e->send.as<int32_t>(), e->sscore.as<int32_t>(), e->sowned.as<uint8_t>(), n, limits->max_overlap,
             limits->max_overlap >= 1 ? 1 : 0, 0, 0, e->flag.as<uint32_t>(), d_overflow + 1, halo_bases); dev::launch_check("k_sweep"); ++e->launches;
  tb = dev::exclusive_sum_u32_tmp(nn); e->tmp.ensure(tb);
  dev::exclusive_sum_u32(e->tmp.p, tb, e->flag.as<uint32_t>(), e->pos.as<uint32_t>(), nn, s); ++e->launches;
  CAL_LAUNCH(k_publish_sum, 1, 1, 0, s, 1, e->pos.as<uint32_t>() + (n - 1), e->flag.as<uint32_t>() + (n - 1), (const unsigned long long*)nullptr, e->h_count_dev + CNT_KEEPERS); dev::launch_check("k_publish_sum");
  dev::stream_sync(s);
  const int64_t nk = (int64_t)((volatile unsigned long long*)e->h_count)[CNT_KEEPERS];
  if (nk == 0) return 0;
  // ---- final order: ReferenceHit.sort over the keepers, stable on the sweep order (variant set rank, then arrival) -----------------------------------
  CAL_LAUNCH(k_variant_final_keys, blocks_for(n, 256), 256, 0, s, 1, recs, rw, e->var_info.as<VarInfo>(), i3, e->flag.as<uint32_t>(), e->pos.as<uint32_t>(), n, F, minor, i1, d_overflow); dev::launch_check("k_variant_final_keys"); ++e->launches;
  tb = dev::sort_pairs_u64_tmp((size_t)nk, 0, 64); e->tmp.ensure(tb);
  dev::sort_pairs_u64(e->tmp.p, tb, minor, ktmp, i1, i2, (size_t)nk, 0, F.bits + F.score_bits, s); ++e->launches;
  e->out2.ensure((size_t)nk * rw * 4); e->var_info2.ensure((size_t)nk * sizeof(VarInfo));
  CAL_LAUNCH(k_variant_gather, blocks_for(nk, 256), 256, 0, s, 1, recs, rw, e->var_info.as<VarInfo>(), i2, nk, e->out2.as<uint32_t>(), e->var_info2.as<VarInfo>()); dev::launch_check("k_variant_gather"); ++e->launches;
  CAL_LAUNCH(k_publish_u64, 1, 1, 0, s, 1, e->d_count + CNT_DEDUP_OVERFLOW, e->h_count_dev + CNT_KEEPERS + 1); dev::launch_check("k_publish_u64");
  dev::stream_sync(s);
  { const unsigned long long fl = ((volatile unsigned long long*)e->h_count)[CNT_KEEPERS + 1];
    if (fl & 0xFFFFFFFFull) throw std::runtime_error("internal error: a hit field exceeds its sort-key width in the variant merge");
    if (fl >> 32) throw LimitExceeded("removeOverlaps: a chain of overlapping hits runs through the whole halo of a shard cut; this input cannot be de-duplicated per shard (use fewer shards)"); }
  return nk;
}

}  // namespace

// ====================================================================================================================================
// C ABI
// ====================================================================================================================================
extern "C" {

const char* calitas_last_error(void) { return g_last_error.c_str(); }
int calitas_tools_set_error(int code, const char* msg) { return set_error(code, msg ? msg : ""); }   // used by cal_tools.cpp
int calitas_engine_get_costs(const calitas_engine* e, calitas_costs* out) {
  if (!e || !out) return set_error(CALITAS_EINVAL, "bad arguments");
  *out = e->costs; return CALITAS_OK;
}

int calitas_engine_create(int32_t device_id, const calitas_costs* costs, calitas_engine** out) {
  return guarded([&]() -> int {
    if (!out) throw InvalidArgument("out is NULL");
    *out = nullptr;
    calitas_costs c = costs ? *costs : calitas_costs{ -120, -122, -121, -260 };
    dev::init(device_id);
    std::unique_ptr<calitas_engine> e(new calitas_engine());
    e->device = device_id; e->costs = c; e->sc = make_scores(c);
    e->stream = dev::stream_create_prio(1); e->scan_stream = dev::stream_create_prio(0); e->scan_stream2 = dev::stream_create_prio(0); e->pub_stream = dev::stream_create_prio(1); e->copy_stream = dev::stream_create_prio(1);
#ifndef CAL_HOSTSIM
    // once, to the most any launch asks for (SCAN_SMEM_LIMIT): the attribute is per device and function, and engines of several host threads
    // share it — setting it per launch let one thread lower it under another's launch
    dev::check(cudaFuncSetAttribute(k_scan_tiled, cudaFuncAttributeMaxDynamicSharedMemorySize, SCAN_SMEM_LIMIT), "cudaFuncSetAttribute");
#endif
    for (auto& ev : e->ev) ev = dev::event_create();
    { void* alias = nullptr; e->h_count = (unsigned long long*)dev::alloc_host_mapped(128, &alias); e->h_count_dev = (unsigned long long*)alias; }
    e->d_count = (unsigned long long*)dev::alloc(64);
    *out = e.release();
    return CALITAS_OK;
  });
}

void calitas_engine_destroy(calitas_engine* e) {
  if (!e) return;
  try {
    dev::set_device(e->device);
    dev::stream_sync(e->scan_stream); dev::stream_sync(e->scan_stream2); dev::stream_sync(e->pub_stream); dev::stream_sync(e->stream); dev::stream_sync(e->copy_stream);
    for (DBuf* b : { &e->cand_b, &e->cand_c, &e->specs, &e->cand, &e->cand_sorted, &e->hits, &e->valid, &e->rank, &e->perm, &e->flag, &e->pos, &e->out, &e->tmp, &e->key1, &e->keyA,
                     &e->idx, &e->idx2, &e->key_b, &e->sstart, &e->send, &e->sscore, &e->windows, &e->nib_tmp, &e->slot_owned, &e->sowned, &e->out_owned, &e->var_info, &e->var_info2, &e->out2, &e->vk1, &e->vk2, &e->vk3, &e->vi1, &e->vi2, &e->vi3 }) b->release();
    for (auto& p : e->pinned_pool) dev::free_host(p.p);
    dev::free_host(e->h_count); dev::free_(e->d_count);
    for (auto& ev : e->ev) dev::event_destroy(ev);
    for (auto& ce : e->chunk_ev) for (auto& ev : ce.ev) dev::event_destroy(ev);
    dev::stream_destroy(e->stream); dev::stream_destroy(e->scan_stream); dev::stream_destroy(e->scan_stream2); dev::stream_destroy(e->pub_stream); dev::stream_destroy(e->copy_stream);
  } catch (...) {}
  delete e;
}

int calitas_shard_plan(int32_t n_contigs, const int64_t* lengths, int32_t shard, int32_t n_shards, int64_t halo,
                       int64_t* own_begin, int64_t* own_end, int64_t* have_begin, int64_t* have_end) {
  return guarded([&]() -> int {
    if (n_contigs <= 0 || !lengths || n_shards <= 0 || shard < 0 || shard >= n_shards || halo < 0) throw InvalidArgument("bad shard plan arguments");
    int64_t tot = 0; for (int c = 0; c < n_contigs; ++c) tot += lengths[c];
    const int64_t lo = (int64_t)((__int128)tot * shard / n_shards), hi = (int64_t)((__int128)tot * (shard + 1) / n_shards);
    int64_t off = 0;
    for (int c = 0; c < n_contigs; ++c) {
      const int64_t len = lengths[c];
      int64_t b = std::min(std::max<int64_t>(lo - off, 0), len), en = std::min(std::max<int64_t>(hi - off, 0), len);
      own_begin[c] = b; own_end[c] = en; have_begin[c] = en > b ? std::max<int64_t>(0, b - halo) : b; have_end[c] = en > b ? std::min(len, en + halo) : b;
      off += len;
    }
    return CALITAS_OK;
  });
}

int calitas_reference_load(calitas_engine* e, int32_t n_contigs, const char* const* names, const int64_t* lengths, const uint8_t* const* bases,
                           const int64_t* have_begin, const int64_t* have_end, const int64_t* own_begin, const int64_t* own_end, int32_t keep_raw,
                           calitas_reference** out) {
  return guarded([&]() -> int {
    if (!e || !out || n_contigs <= 0 || !names || !lengths || !bases) throw InvalidArgument("bad reference arguments");
    if (n_contigs >= (1 << 18)) throw LimitExceeded("more than 262143 contigs");
    *out = nullptr; dev::set_device(e->device);
    std::unique_ptr<calitas_reference> r(new calitas_reference());
    r->owner = e; int64_t off = 64;
    for (int c = 0; c < n_contigs; ++c) {
      const int64_t len = lengths[c];
      if (len < 0 || len >= (1ll << 31)) throw LimitExceeded("contig longer than 2^31-1 bases");
      const int64_t hb = have_begin ? have_begin[c] : 0, he = have_end ? have_end[c] : len, ob = own_begin ? own_begin[c] : 0, oe = own_end ? own_end[c] : len;
      if (hb < 0 || he < hb || he > len || ob < 0 || oe < ob || oe > len) throw InvalidArgument("bad base range for contig");
      r->names.push_back(names[c] ? names[c] : ""); r->len.push_back(len); r->have_b.push_back(hb); r->have_e.push_back(he); r->own_b.push_back(ob); r->own_e.push_back(oe);
      r->nib_off.push_back(off);
      off += ((he - hb) + 63) / 64 * 64 + 64;
    }
    r->total_padded = off;
    r->d_raw = (uint8_t*)dev::alloc((size_t)off);
    r->d_nib = (uint32_t*)dev::alloc((size_t)off / 2);
    dev::zero(r->d_raw, (size_t)off, e->stream);
    {  // caller memory is pageable: stage through two pinned buffers so that the host copy of chunk k+1 overlaps the H2D of chunk k
      const size_t STAGE = 32u << 20;
      void* stage[2] = { dev::alloc_host(STAGE), dev::alloc_host(STAGE) }; int turn = 0; bool used[2] = { false, false };
      try {
        for (int c = 0; c < n_contigs; ++c) {
          const int64_t n = r->have_e[(size_t)c] - r->have_b[(size_t)c];
          if (n > 0 && !bases[c]) throw InvalidArgument("bases pointer is NULL");
          for (int64_t off2 = 0; off2 < n; off2 += (int64_t)STAGE) {
            const size_t len = (size_t)std::min<int64_t>((int64_t)STAGE, n - off2);
            if (used[turn]) dev::event_sync(e->ev[4 + turn]);
            std::memcpy(stage[turn], bases[c] + off2, len);
            dev::h2d(r->d_raw + r->nib_off[(size_t)c] + off2, stage[turn], len, e->stream);
            dev::event_record(e->ev[4 + turn], e->stream); used[turn] = true; turn ^= 1;
          }
        }
        dev::stream_sync(e->stream);
      } catch (...) { try { dev::stream_sync(e->stream); } catch (...) {} dev::free_host(stage[0]); dev::free_host(stage[1]); dev::free_(r->d_raw); dev::free_(r->d_nib); throw; }
      dev::free_host(stage[0]); dev::free_host(stage[1]);
    }
    const int64_t n_words = off / 8;
    const unsigned grid = (unsigned)std::min<int64_t>((n_words + 255) / 256, (int64_t)dev::sm_count(e->device) * 16);
    CAL_LAUNCH(k_pack, grid, 256, 256, e->stream, 2, r->d_raw, r->d_nib, n_words); dev::launch_check("k_pack");
    dev::stream_sync(e->stream);
    if (!keep_raw) { dev::free_(r->d_raw); r->d_raw = nullptr; }
    *out = r.release();
    return CALITAS_OK;
  });
}

void calitas_reference_free(calitas_engine* e, calitas_reference* r) {
  if (!r) return;
  try {
    if (e) dev::set_device(e->device);
    dev::free_(r->d_nib); dev::free_(r->d_raw);
    for (auto& kv : r->tilesets) { dev::free_(kv.second.d_tiles); dev::free_(kv.second.d_contigs); }
  } catch (...) {}
  delete r;
}

// One guide chunk of a search: consecutive guides sharing the raw guide length, hence the window tiling (SearchReference.scala:528-530).
struct SearchChunk { int g0, g1, raw_len, step, slots, scan_slots, banded; int fast; calitas_reference::TileSet* ts; size_t t_begin, n_tiles, smem; int64_t bases; KeyLayout key; DedupLayout dedup; };

static int search_impl(calitas_engine* e, const calitas_reference* ref_c, int32_t n_guides, const calitas_guide* guides, const calitas_limits* limits,
                       int32_t window_size, const char* chrom, int32_t dedup, const VariantPlan* vp, calitas_hitset** out) {
  return guarded([&]() -> int {
    if (!e || !ref_c || !out) throw InvalidArgument("bad search arguments");
    *out = nullptr;
    calitas_reference* ref = const_cast<calitas_reference*>(ref_c);
    dev::set_device(e->device); dev::Stream s = e->stream, ss0 = e->scan_stream, cs = e->copy_stream;
    const bool two_scan_streams = !std::getenv("CALITAS_ONE_SCAN_STREAM");
    dev::Stream scan_streams[2] = { e->scan_stream, two_scan_streams ? e->scan_stream2 : e->scan_stream };
    std::vector<GuideDef> defs; std::vector<GuideSpec> specs = build_specs(e, n_guides, guides, limits, false, &defs);
    if (window_size <= 0 || (uint32_t)window_size > MAX_WINDOW_LEN) throw InvalidArgument("window size out of range");
    if (dedup && limits->max_overlap <= 0) {      // every later hit of a group then "overlaps" (>= 0): the sweep has unbounded reach, no halo makes a shard exact
      for (size_t c = 0; c < ref->len.size(); ++c) if (ref->own_b[c] != 0 || ref->own_e[c] != ref->len[c])
        throw InvalidArgument("max_overlap <= 0 cannot be de-duplicated per shard: search with dedup = 0 and run removeOverlaps over the gathered hits");
    }
    int chrom_idx = -1;
    if (chrom && chrom[0]) { for (size_t c = 0; c < ref->names.size(); ++c) if (ref->names[c] == chrom) chrom_idx = (int)c; if (chrom_idx < 0) throw InvalidArgument(std::string("Unknown chromosome: ") + chrom); }
    // ---- plan: guide chunks and their window tilings ---------------------------------------------------------------------------
    // measured on B200: chunks of 25-50 guides, or one tail per group of 4 chunks, shorten nothing at 1/8 genome scale (a tail that shares the SMs with
    // a scan kernel slows down in proportion to its work) and expose a longer last tail at full scale (-2 to -5 %)
    std::vector<int> plan;                                      // guides per chunk, in order; the last entry repeats.  CALITAS_CHUNK_PLAN=24,24,16,8,4 overrides (A/B runs)
    if (const char* pl = std::getenv("CALITAS_CHUNK_PLAN")) { for (const char* q = pl; *q;) { const int v = std::atoi(q); if (v > 0) plan.push_back(std::min(v, 64)); while (*q && *q != ',') ++q; if (*q) ++q; } }
    if (plan.empty()) plan.push_back(16);
    std::vector<SearchChunk> chunks;
    for (int g0 = 0; g0 < n_guides;) {
      SearchChunk ch; ch.g0 = g0; ch.raw_len = (int)defs[(size_t)g0].raw.size();
      const int G_CHUNK = plan[std::min(chunks.size(), plan.size() - 1)];
      int g1 = g0 + 1; while (g1 < n_guides && g1 - g0 < G_CHUNK && (int)defs[(size_t)g1].raw.size() == ch.raw_len) ++g1;
      ch.g1 = g1;
      const int overlap = ch.raw_len + limits->max_guide_diffs + limits->max_gaps_between_guide_and_pam - 1;
      ch.step = window_size - overlap;
      if (ch.step <= 0) throw InvalidArgument("window size must exceed guide length + max-guide-diffs + max-gaps-between-guide-and-pam - 1");
      ch.ts = &tileset_for(e, ref, window_size, ch.step, chrom);
      // tiles of one contig are contiguous: restrict the launch for -c
      ch.t_begin = 0; size_t t_end = ch.ts->tiles.size();
      if (chrom_idx >= 0) { ch.t_begin = t_end = 0; bool in = false; for (size_t t = 0; t < ch.ts->tiles.size(); ++t) { if (ch.ts->tiles[t].contig == chrom_idx) { if (!in) { ch.t_begin = t; in = true; } t_end = t + 1; } } }
      ch.n_tiles = t_end - ch.t_begin;
      ch.slots = 1; ch.banded = 1; for (int g = g0; g < g1; ++g) { ch.slots = std::max(ch.slots, specs[(size_t)g].slots); ch.banded = std::max(ch.banded, std::max(specs[(size_t)g].k_edits, specs[(size_t)g].band_k)); }
      if (ch.banded > ALIGN_KB) ch.banded = 0;
      ch.fast = align_level(specs, (size_t)g0, (size_t)g1, ch.banded, e->sc);
      const int ng = g1 - g0;
      ch.smem = scan_smem_bytes(ng, ch.ts->tile_windows, window_size, ch.step);
      if (ch.smem > (size_t)SCAN_SMEM_LIMIT) throw LimitExceeded("window size too large for the shared-memory tile");
      { const int want = (ng + SCAN_NG - 1) / SCAN_NG; ch.scan_slots = want >= 3 ? 4 : (want == 2 ? 2 : 1); }     // guide slots per window (1, 2 or 4); the rest of the block's 4 slots split the window into parts
      ch.key = make_key_layout((uint32_t)window_size, (uint64_t)ch.ts->total_windows, (uint32_t)n_guides);
      {   // removeOverlaps sort keys: field widths from the reference and the chunk; score bounds from the guides' thresholds and costs
        DedupLayout& D = ch.dedup; const Scores& sc = e->sc;
        int64_t max_len = 1; for (int64_t l : ref->len) max_len = std::max(max_len, l);
        D.start_bits = std::min(31, bit_length((uint64_t)max_len)); D.contig_shift = D.start_bits + 1;        // bit 0 = strand, then the start
        D.guide_shift = D.contig_shift + bit_length((uint64_t)(ref->len.size() - 1)); D.g0 = g0; D.bits = D.guide_shift + bit_length((uint64_t)(g1 - g0 - 1));
        score_bounds(specs, (size_t)g0, (size_t)g1, sc, D.score_hi, D.score_bits);
        D.merged = (D.bits + D.score_bits <= 64 && !std::getenv("CALITAS_DEDUP_TWO_SORTS")) ? 1 : 0;      // the variable forces the wide-key path in tests
      }
      ch.bases = 0; for (size_t t = ch.t_begin; t < t_end; ++t) ch.bases += (int64_t)(ch.ts->tiles[t].nwin - 1) * ch.step + window_size;
      chunks.push_back(ch);
      g0 = g1;
    }
    const size_t n_chunks = chunks.size();
    while (e->chunk_ev.size() < n_chunks) { ChunkEvents ce; for (auto& ev : ce.ev) ev = dev::event_create(); e->chunk_ev.push_back(ce); }
    e->launches = 0;
    dev::zero(e->d_count + CNT_DEDUP_OVERFLOW, 8, s);                         // field-overflow flag of k_dedup_keys
    dev::event_record(e->ev[0], ss0);
    e->specs.ensure(specs.size() * sizeof(GuideSpec)); dev::h2d(e->specs.p, specs.data(), specs.size() * sizeof(GuideSpec), ss0);
    dev::event_record(e->ev[2], ss0); dev::stream_wait(scan_streams[1], e->ev[2]);      // the guide table is uploaded on the first scan stream
    double ms[8] = { 0, 0, 0, 0, 0, 0, 0, 0 }; int64_t counts[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
    counts[4] = (int64_t)(specs.size() * sizeof(GuideSpec));
    { const calitas_reference::TileSet& ts = *chunks[0].ts; for (size_t c = 0; c < ts.contigs.size(); ++c) if (chrom_idx < 0 || (int)c == chrom_idx) counts[0] += ts.contigs[c].own_hi - ts.contigs[c].own_lo; }
    // ---- pipeline: scan(c+1) runs on the scan stream while sort/align/canon/dedup of chunk c run on the main stream and finished hit
    //      segments go to pinned host memory on the copy stream -------------------------------------------------------------------
    // Three candidate slots: two scans are always queued ahead of the chunk whose tail the host is driving, so the scan stream never
    // waits for the host (the tail's small dependent launches and counter read-backs are slow while a scan kernel holds the SMs).
    const int N_SLOTS = 3;
    DBuf* cand_slot[N_SLOTS] = { &e->cand, &e->cand_b, &e->cand_c };
    auto launch_scan = [&](size_t c) {
      const SearchChunk& ch = chunks[c]; const int slot = (int)(c % N_SLOTS); ChunkEvents& ce = e->chunk_ev[c];
      dev::Stream ss = scan_streams[c & 1];
      cand_slot[slot]->ensure(e->cand_cap_hint * 8);
      if (c >= (size_t)N_SLOTS) dev::stream_wait(ss, e->chunk_ev[c - N_SLOTS].ev[CE_SORTED]);       // the previous user of this candidate slot has been sorted away
      dev::zero(e->d_count + slot, 8, ss);
      dev::event_record(ce.ev[CE_SCAN_B], ss);
      if (ch.n_tiles) {
        ScanArgs sa{ ref->d_nib, ch.ts->d_contigs, ch.ts->d_tiles + ch.t_begin, e->specs.as<GuideSpec>(), ch.g0, ch.g1, window_size, ch.step, ch.raw_len, ch.scan_slots, ch.ts->tile_windows,
                     cand_slot[slot]->as<uint64_t>(), e->d_count + slot, (unsigned long long)e->cand_cap_hint, ch.key };
        // always 2 * tile_windows * 4 threads (512 at the default window size): 4 / scan_slots window parts per guide slot
        CAL_LAUNCH(k_scan_tiled, (unsigned)ch.n_tiles, 2 * ch.ts->tile_windows * 4, ch.smem, ss, 3, sa);
        dev::launch_check("k_scan_tiled"); ++e->launches;
        counts[6] += 1; counts[7] += ch.bases;
      }
      dev::event_record(ce.ev[CE_SCAN_E], ss);
      // the count is published from a stream of its own with the greatest priority: on a scan stream the one-thread kernel would queue behind every
      // pending block of the next chunk's scan (same priority, other stream) and the tail would start a whole scan late
      dev::stream_wait(e->pub_stream, ce.ev[CE_SCAN_E]);
      CAL_LAUNCH(k_publish_u64, 1, 1, 0, e->pub_stream, 1, e->d_count + slot, e->h_count_dev + slot); dev::launch_check("k_publish_u64");
      dev::event_record(ce.ev[CE_COUNT], e->pub_stream);
    };
    int max_cols = 1; for (auto& sp : specs) max_cols = std::max(max_cols, sp.max_cols);
    const int rw = rec_words_for(max_cols); const size_t rec_bytes = (size_t)rw * 4;
    PinnedBuf pin = take_pinned(e, std::max<size_t>(e->out_hits_hint, 1024) * rec_bytes);
    PinnedBuf var_info_pin;
    int64_t n_out = 0; size_t copies = 0;
    try {
      for (size_t c = 0; c < (size_t)(N_SLOTS - 1) && c < n_chunks; ++c) launch_scan(c);
      for (size_t c = 0; c < n_chunks; ++c) {
        const SearchChunk& ch = chunks[c]; const int slot = (int)(c % N_SLOTS); ChunkEvents& ce = e->chunk_ev[c];
        if (c + N_SLOTS - 1 < n_chunks) launch_scan(c + N_SLOTS - 1);
        unsigned long long n_cand = 0;
        for (;;) {
          dev::event_sync(ce.ev[CE_COUNT]);
          n_cand = ((volatile unsigned long long*)e->h_count)[slot];
          if (n_cand <= e->cand_cap_hint) break;
          // pool too small: grow and re-run this chunk's scan (and the one queued behind it), never truncate
          dev::stream_sync(scan_streams[0]); dev::stream_sync(scan_streams[1]); dev::stream_sync(e->pub_stream);
          e->cand_cap_hint = (size_t)(n_cand + n_cand / 4);
          for (size_t r = c; r < c + N_SLOTS && r < n_chunks; ++r) { counts[6] -= chunks[r].n_tiles ? 1 : 0; counts[7] -= chunks[r].bases; launch_scan(r); }
        }
        counts[1] += (int64_t)n_cand;
        dev::event_record(ce.ev[CE_TAIL_B], s);
        Pipeline P{ e, cand_slot[slot]->as<uint64_t>(), ce.ev[CE_SORTED], ce.ev[CE_ALIGN_B], ce.ev[CE_ALIGN_E], e->specs.as<GuideSpec>(), ch.slots, false, ch.banded,
                    ref->d_nib, ch.ts->d_contigs, (int)ch.ts->contigs.size(), window_size, ch.step, nullptr, dedup == 0 && !vp, ch.key, rw, ch.fast, ref->total_padded / 8, (dedup && !vp) ? &ch.dedup : nullptr, limits->max_overlap, vp ? &e->out_owned : nullptr };
        P.halo_bases = HALO_WINDOWS * ch.step - (window_size - ch.step) - CALITAS_MAX_OPS;
        if (const char* hb = std::getenv("CALITAS_TEST_HALO_BASES")) P.halo_bases = std::atoi(hb);      // tests: a reach so short that the sentinel fires on ordinary input
        int64_t n_aln = 0;
        // room in e->out: what this chunk adds, and (when it has to grow) the rest of the call projected from the hits per guide so far
        const size_t projected = ch.g0 > 0 ? (size_t)((double)n_out * (double)n_guides / (double)ch.g0 * 1.15) + 4096 : e->out_hits_hint;
        const int64_t n_new = run_tail(P, (int64_t)n_cand, n_out, projected, n_aln);
        counts[2] += n_aln;
        dev::event_record(ce.ev[CE_TAIL_E], s);
        // the finished segment goes to the host while later chunks compute
        dev::stream_wait(cs, ce.ev[CE_TAIL_E]);
        dev::event_record(ce.ev[CE_COPY_B], cs);
        if (vp) n_out += n_new;                                    // with variant windows the hits stay on the device until the merge below
        else if (n_new) {
          if ((size_t)(n_out + n_new) * rec_bytes > pin.cap) {
            dev::stream_sync(cs);
            // first call on this engine: size the result buffer once from the hits per guide seen so far (page-locking gigabytes is slow,
            // so growing it chunk by chunk cost seconds); later calls start from the previous call's total
            const size_t projected = (size_t)((double)(n_out + n_new) * (double)n_guides / (double)std::max(1, ch.g1) * 1.15) + 4096;
            PinnedBuf bigger = take_pinned(e, std::max((size_t)(n_out + n_new) * 3 / 2, projected) * rec_bytes);
            std::memcpy(bigger.p, pin.p, (size_t)n_out * rec_bytes);
            e->pinned_pool.push_back(pin); pin = bigger;
          }
          dev::d2h((char*)pin.p + (size_t)n_out * rec_bytes, e->out.as<char>() + (size_t)n_out * rec_bytes, (size_t)n_new * rec_bytes, cs);
          n_out += n_new;
        }
        dev::event_record(ce.ev[CE_COPY_E], cs);
        ++copies;
      }
      PinnedBuf info_pin;
      if (vp) {        // variant windows: align, merge with the reference hits, removeOverlaps + sort, then one copy of records and annotations
        dev::stream_sync(cs); dev::stream_sync(s);
        int32_t halo_bases = 0x7FFFFFFF; for (auto& ch : chunks) halo_bases = std::min(halo_bases, HALO_WINDOWS * ch.step - (window_size - ch.step) - CALITAS_MAX_OPS);
        n_out = merge_variant_hits(e, *vp, specs, limits, ref, rw, n_out, halo_bases, ms, counts);
        if ((size_t)n_out * rec_bytes > pin.cap) { e->pinned_pool.push_back(pin); pin = take_pinned(e, (size_t)n_out * rec_bytes); }
        info_pin = take_pinned(e, std::max<size_t>(1, (size_t)n_out) * sizeof(calitas_variant_hit_info));
        dev::event_record(e->ev[6], s);
        dev::d2h(pin.p, e->out2.p, (size_t)n_out * rec_bytes, s);
        dev::d2h(info_pin.p, e->var_info2.p, (size_t)n_out * sizeof(calitas_variant_hit_info), s);
        dev::event_record(e->ev[7], s);
        dev::stream_wait(cs, e->ev[7]);
        counts[5] += (int64_t)((size_t)n_out * sizeof(calitas_variant_hit_info));
      }
      var_info_pin = info_pin;
      dev::event_record(e->ev[1], cs);
      dev::stream_sync(cs); dev::stream_sync(s); dev::stream_sync(scan_streams[0]); dev::stream_sync(scan_streams[1]); dev::stream_sync(e->pub_stream);
      if (vp) ms[4] += dev::event_ms(e->ev[6], e->ev[7]);
    } catch (...) {
      try { dev::stream_sync(scan_streams[0]); dev::stream_sync(scan_streams[1]); dev::stream_sync(e->pub_stream); dev::stream_sync(s); dev::stream_sync(cs); } catch (...) {}
      e->pinned_pool.push_back(pin);
      throw;
    }
    double scan_until = 0;                                      // scans of consecutive chunks overlap at their ends (two scan streams): ms[1] is the union of the intervals
    for (size_t c = 0; c < n_chunks; ++c) {
      const ChunkEvents& ce = e->chunk_ev[c];
      const double sb = std::max(scan_until, dev::event_ms(e->ev[0], ce.ev[CE_SCAN_B])), se = dev::event_ms(e->ev[0], ce.ev[CE_SCAN_E]);
      if (se > sb) { ms[1] += se - sb; scan_until = se; }
      const double al = dev::event_ms(ce.ev[CE_ALIGN_B], ce.ev[CE_ALIGN_E]);
      ms[2] += al;
      ms[3] += dev::event_ms(ce.ev[CE_TAIL_B], ce.ev[CE_TAIL_E]) - al;
      ms[4] += dev::event_ms(ce.ev[CE_COPY_B], ce.ev[CE_COPY_E]);
    }
    ms[0] = dev::event_ms(e->ev[0], e->ev[1]);
    if (std::getenv("CALITAS_TRACE")) {       // per-chunk timeline (ms since the start of the call) on stderr
      for (size_t c = 0; c < n_chunks; ++c) {
        const ChunkEvents& ce = e->chunk_ev[c];
        std::fprintf(stderr, "[calitas trace] chunk %zu guides %d-%d: scan %.2f-%.2f  tail %.2f-%.2f (sorted %.2f, align %.2f-%.2f)  copy %.2f-%.2f\n", c, chunks[c].g0, chunks[c].g1,
                     dev::event_ms(e->ev[0], ce.ev[CE_SCAN_B]), dev::event_ms(e->ev[0], ce.ev[CE_SCAN_E]), dev::event_ms(e->ev[0], ce.ev[CE_TAIL_B]), dev::event_ms(e->ev[0], ce.ev[CE_TAIL_E]),
                     dev::event_ms(e->ev[0], ce.ev[CE_SORTED]), dev::event_ms(e->ev[0], ce.ev[CE_ALIGN_B]), dev::event_ms(e->ev[0], ce.ev[CE_ALIGN_E]),
                     dev::event_ms(e->ev[0], ce.ev[CE_COPY_B]), dev::event_ms(e->ev[0], ce.ev[CE_COPY_E]));
      }
      std::fprintf(stderr, "[calitas trace] total %.2f ms\n", ms[0]);
    }
    ms[5] = ms[0] - ms[1];                                   // time of the call not hidden behind the scan kernels
    counts[3] = e->launches;
    counts[5] += (int64_t)((size_t)n_out * rec_bytes);
    e->out_hits_hint = std::max<size_t>(e->out_hits_hint, (size_t)n_out + (size_t)n_out / 8);
    std::unique_ptr<calitas_hitset> hs(new calitas_hitset());
    hs->owner = e; hs->n = n_out; hs->buf = pin; hs->stride = (int32_t)rec_bytes; hs->info = var_info_pin;
    for (int i = 0; i < 8; ++i) { hs->ms[i] = ms[i]; hs->counts[i] = counts[i]; }
    *out = hs.release();
    return CALITAS_OK;
  });
}

int calitas_search(calitas_engine* e, const calitas_reference* ref, int32_t n_guides, const calitas_guide* guides, const calitas_limits* limits,
                   int32_t window_size, const char* chrom, int32_t dedup, calitas_hitset** out) {
  return search_impl(e, ref, n_guides, guides, limits, window_size, chrom, dedup, nullptr, out);
}

int calitas_variant_set_load(calitas_engine* e, const calitas_reference* ref, int64_t n_windows, const calitas_variant_window* windows, int64_t n_alleles, const calitas_variant_allele* alleles,
                             int64_t n_sets, const uint32_t* set_rank, calitas_variant_set** out) {
  return guarded([&]() -> int {
    if (!e || !ref || !out || n_windows < 0 || (n_windows && !windows) || n_alleles < 0 || (n_alleles && !alleles) || n_sets < 0 || (n_sets && !set_rank)) throw InvalidArgument("bad variant arguments");
    *out = nullptr;
    if (n_windows >= (1ll << 31)) throw LimitExceeded("too many variant windows");
    dev::set_device(e->device); dev::Stream s = e->stream;
    std::unique_ptr<calitas_variant_set> vs(new calitas_variant_set());
    vs->owner = e; vs->n_windows = n_windows; vs->n_alleles = n_alleles; vs->n_sets = n_sets; vs->n_contigs = (int32_t)ref->len.size();
    // window bases: concatenated at multiples of 8 with a gap, uploaded raw and packed on the device like the reference
    std::vector<VarWindowDev> wd((size_t)n_windows);
    int64_t total = 64; int32_t prev_class = 0;
    for (int64_t w = 0; w < n_windows; ++w) {
      const calitas_variant_window& v = windows[w];
      if (v.length < 0 || (uint32_t)v.length > MAX_WINDOW_LEN || (v.length && !v.bases) || v.n_alleles < 0 || v.first_allele < 0 || v.first_allele + (int64_t)v.n_alleles > n_alleles ||
          v.first_set < 0 || v.first_set + (int64_t)v.n_alleles * (v.n_alleles + 1) / 2 > n_sets || v.contig_idx < 0 || v.contig_idx >= vs->n_contigs || v.guide_class < 0 || v.guide_class >= (1 << 20))
        throw InvalidArgument("bad variant window");
      if (v.guide_class < prev_class) throw InvalidArgument("variant windows must be sorted by guide_class");
      prev_class = v.guide_class;
      if ((size_t)v.guide_class >= vs->class_begin.size()) { vs->class_begin.resize((size_t)v.guide_class + 1, w); vs->class_end.resize((size_t)v.guide_class + 1, w); }
      vs->class_end[(size_t)v.guide_class] = w + 1;
      vs->max_len = std::max<uint32_t>(vs->max_len, (uint32_t)v.length);
      wd[(size_t)w] = VarWindowDev{ total, v.length, v.contig_idx, v.ref_start, v.n_alleles, v.first_allele, v.first_set, v.owned ? 1 : 0, 0 };
      total += (v.length + 7) / 8 * 8 + 8;
    }
    for (size_t c = 0; c < vs->class_begin.size(); ++c) if (vs->class_end[c] < vs->class_begin[c]) vs->class_end[c] = vs->class_begin[c];
    total += 64;
    for (int64_t k = 0; k < n_sets; ++k) vs->max_rank = std::max(vs->max_rank, set_rank[k]);
    {
      // the bases are copied into one staging buffer on all host threads (millions of small windows)
      std::vector<uint8_t> raw((size_t)total, 0);
      const int nt = (int)std::max<int64_t>(1, std::min<int64_t>(n_windows / 65536 + 1, std::min<unsigned>(16u, std::max(1u, std::thread::hardware_concurrency()))));
      std::vector<std::thread> th;
      for (int t = 0; t < nt; ++t) th.emplace_back([&, t]() { for (int64_t w = n_windows * t / nt; w < n_windows * (t + 1) / nt; ++w) if (windows[w].length) std::memcpy(raw.data() + wd[(size_t)w].nib_start, windows[w].bases, (size_t)windows[w].length); });
      for (auto& t : th) t.join();
      DBuf d_raw; struct Guard { DBuf& b; ~Guard() { try { b.release(); } catch (...) {} } } guard{ d_raw };
      d_raw.ensure((size_t)total); vs->nib.ensure((size_t)total / 2 + 8);
      dev::h2d(d_raw.p, raw.data(), (size_t)total, s);
      const int64_t n_words = total / 8;
      const unsigned grid = (unsigned)std::min<int64_t>((n_words + 255) / 256, (int64_t)dev::sm_count(e->device) * 16);
      CAL_LAUNCH(k_pack, grid, 256, 256, s, 2, d_raw.as<uint8_t>(), vs->nib.as<uint32_t>(), n_words); dev::launch_check("k_pack");
      dev::stream_sync(s);
      vs->nib_words = n_words;
    }
    vs->windows.ensure(std::max<size_t>(1, wd.size()) * sizeof(VarWindowDev)); dev::h2d(vs->windows.p, wd.data(), wd.size() * sizeof(VarWindowDev), s);
    vs->alleles.ensure(std::max<size_t>(1, (size_t)n_alleles) * sizeof(calitas_variant_allele)); dev::h2d(vs->alleles.p, alleles, (size_t)n_alleles * sizeof(calitas_variant_allele), s);
    vs->sets.ensure(std::max<size_t>(1, (size_t)n_sets) * 4); dev::h2d(vs->sets.p, set_rank, (size_t)n_sets * 4, s);
    dev::stream_sync(s);
    *out = vs.release();
    return CALITAS_OK;
  });
}
void calitas_variant_set_free(calitas_variant_set* vs) {
  if (!vs) return;
  try { if (vs->owner) dev::set_device(vs->owner->device); vs->nib.release(); vs->windows.release(); vs->alleles.release(); vs->sets.release(); } catch (...) {}
  delete vs;
}

int calitas_search_variants(calitas_engine* e, const calitas_reference* ref, int32_t n_guides, const calitas_guide* guides, const int32_t* guide_class, const calitas_limits* limits,
                            int32_t window_size, const char* chrom, const calitas_variant_set* variants, calitas_hitset** out) {
  if (!variants || !guide_class) return set_error(CALITAS_EINVAL, "bad variant arguments");
  VariantPlan vp{ variants, guide_class };
  return search_impl(e, ref, n_guides, guides, limits, window_size, chrom, 1, &vp, out);
}

// One table in one address space from N engines (one per GPU, each holding shard s of calitas_shard_plan): the engines search concurrently on host
// threads; per guide the shards' lists (each in ReferenceHit.sort order, ascending base ranges) are concatenated and merged where they meet -- consecutive
// windows overlap by guide length + d + g - 1 bases, so the last window of one shard and the first of the next can report hits whose starts interleave --
// stably, the earlier shard first on equal keys (the single engine's arrival order).
}  // extern "C"
namespace {
template <int RW> struct RecN { uint32_t w[RW]; };
template <int RW> bool rec_sorts_before(const RecN<RW>& a, const RecN<RW>& b) {      // ReferenceHit.sort (ReferenceHit.scala:276-287): contig, coordinate_start, strand, -score
  const int ca = rec_contig(a.w), cb = rec_contig(b.w); if (ca != cb) return ca < cb;
  const int sa = rec_gstart(a.w), sb = rec_gstart(b.w); if (sa != sb) return sa < sb;
  if (rec_neg(a.w) != rec_neg(b.w)) return rec_neg(a.w) < rec_neg(b.w);
  return rec_score(a.w) > rec_score(b.w);
}
template <int RW> void merge_guide_segments(RecN<RW>* out, const std::vector<std::pair<const RecN<RW>*, int64_t>>& segs) {
  int64_t n = 0;
  for (auto& sg : segs) {
    if (sg.second == 0) continue;
    std::memcpy(out + n, sg.first, (size_t)sg.second * sizeof(RecN<RW>));
    const int64_t seg = n; n += sg.second;
    if (seg == 0 || !rec_sorts_before<RW>(out[seg], out[seg - 1])) continue;
    RecN<RW>* first = std::upper_bound(out, out + seg, out[seg], rec_sorts_before<RW>);               // prefix elements <= the segment's first stay put
    RecN<RW>* last = std::lower_bound(out + seg, out + n, out[seg - 1], rec_sorts_before<RW>);          // segment elements >= the prefix's last stay put
    std::inplace_merge(first, out + seg, last, rec_sorts_before<RW>);
  }
}
}  // namespace
extern "C" {

int calitas_search_sharded(int32_t n_engines, calitas_engine* const* engines, const calitas_reference* const* refs, int32_t n_guides, const calitas_guide* guides,
                           const calitas_limits* limits, int32_t window_size, const char* chrom, calitas_hitset** out) {
  if (n_engines <= 0 || !engines || !refs || !out) return set_error(CALITAS_EINVAL, "bad arguments");
  if (n_engines == 1) return search_impl(engines[0], refs[0], n_guides, guides, limits, window_size, chrom, 1, nullptr, out);
  *out = nullptr;
  std::vector<calitas_hitset*> hs((size_t)n_engines, nullptr); std::vector<int> rc((size_t)n_engines, CALITAS_OK); std::vector<std::string> msg((size_t)n_engines);
  {
    std::vector<std::thread> th;
    for (int s = 0; s < n_engines; ++s) th.emplace_back([&, s]() {
      rc[(size_t)s] = search_impl(engines[s], refs[s], n_guides, guides, limits, window_size, chrom, 1, nullptr, &hs[(size_t)s]);
      if (rc[(size_t)s] != CALITAS_OK) msg[(size_t)s] = g_last_error; });
    for (auto& t : th) t.join();
  }
  auto free_all = [&] { for (auto* h : hs) calitas_hitset_free(h); };
  for (int s = 0; s < n_engines; ++s) if (rc[(size_t)s] != CALITAS_OK) { free_all(); return set_error(rc[(size_t)s], msg[(size_t)s]); }
  return guarded([&]() -> int {
    const auto t0 = std::chrono::steady_clock::now();
    const int stride = hs[0]->stride; const int rw = stride / 4;
    for (auto* h : hs) if (h->stride != stride) { free_all(); throw std::runtime_error("engines returned different record sizes"); }
    // per engine and guide: where the guide's records begin (every list is guide-major)
    std::vector<std::vector<int64_t>> begin((size_t)n_engines, std::vector<int64_t>((size_t)n_guides + 1, 0));
    int64_t total = 0;
    for (int s = 0; s < n_engines; ++s) {
      const char* base = (const char*)hs[(size_t)s]->buf.p; const int64_t n = hs[(size_t)s]->n; int64_t i = 0;
      for (int g = 0; g < n_guides; ++g) {
        begin[(size_t)s][(size_t)g] = i;
        int64_t lo = i, hi = n;                         // first record with guide_idx > g
        while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (rec_guide((const uint32_t*)(base + (size_t)mid * stride)) <= g) lo = mid + 1; else hi = mid; }
        i = lo;
      }
      begin[(size_t)s][(size_t)n_guides] = i; total += n;
      if (i != n) { free_all(); throw std::runtime_error("hit set is not guide-major"); }
    }
    std::vector<int64_t> at((size_t)n_guides + 1, 0);
    for (int g = 0; g < n_guides; ++g) { int64_t c = 0; for (int s = 0; s < n_engines; ++s) c += begin[(size_t)s][(size_t)g + 1] - begin[(size_t)s][(size_t)g]; at[(size_t)g + 1] = at[(size_t)g] + c; }
    calitas_engine* e0 = engines[0]; dev::set_device(e0->device);
    PinnedBuf pin = take_pinned(e0, (size_t)std::max<int64_t>(1, total) * stride);
    // guides are independent: merged on host threads
    const int nt = (int)std::max<int64_t>(1, std::min<int64_t>(n_guides, std::min<unsigned>(32u, std::max(1u, std::thread::hardware_concurrency()))));
    std::atomic<int> next(0); std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t) th.emplace_back([&]() {
      for (;;) {
        const int g = next.fetch_add(1); if (g >= n_guides) return;
        char* dst = (char*)pin.p + (size_t)at[(size_t)g] * stride;
        if (rw == CALITAS_HIT_WORDS) {
          std::vector<std::pair<const RecN<CALITAS_HIT_WORDS>*, int64_t>> segs;
          for (int s = 0; s < n_engines; ++s) segs.emplace_back((const RecN<CALITAS_HIT_WORDS>*)((const char*)hs[(size_t)s]->buf.p + (size_t)begin[(size_t)s][(size_t)g] * stride), begin[(size_t)s][(size_t)g + 1] - begin[(size_t)s][(size_t)g]);
          merge_guide_segments<CALITAS_HIT_WORDS>((RecN<CALITAS_HIT_WORDS>*)dst, segs);
        } else {
          std::vector<std::pair<const RecN<CALITAS_HIT_WIDE_WORDS>*, int64_t>> segs;
          for (int s = 0; s < n_engines; ++s) segs.emplace_back((const RecN<CALITAS_HIT_WIDE_WORDS>*)((const char*)hs[(size_t)s]->buf.p + (size_t)begin[(size_t)s][(size_t)g] * stride), begin[(size_t)s][(size_t)g + 1] - begin[(size_t)s][(size_t)g]);
          merge_guide_segments<CALITAS_HIT_WIDE_WORDS>((RecN<CALITAS_HIT_WIDE_WORDS>*)dst, segs);
        }
      } });
    for (auto& t : th) t.join();
    std::unique_ptr<calitas_hitset> m(new calitas_hitset());
    m->owner = e0; m->n = total; m->buf = pin; m->stride = stride;
    for (int s = 0; s < n_engines; ++s) { for (int k = 0; k < 8; ++k) { m->ms[k] = std::max(m->ms[k], hs[(size_t)s]->ms[k]); m->counts[k] += hs[(size_t)s]->counts[k]; } }
    m->ms[6] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();      // host-side merge into one table
    free_all();
    *out = m.release();
    return CALITAS_OK;
  });
}

int calitas_align_regions(calitas_engine* e, const calitas_reference* ref, int32_t n_guides, const calitas_guide* guides, int64_t n_tasks,
                          const calitas_region_task* tasks, const calitas_limits* limits, int32_t best, calitas_hitset** out) {
  return guarded([&]() -> int {
    if (!e || !ref || !out || n_tasks < 0 || (n_tasks && !tasks)) throw InvalidArgument("bad align_regions arguments");
    *out = nullptr;
    if (n_tasks >= (1ll << 31)) throw LimitExceeded("too many tasks");
    std::vector<GuideSpec> specs = build_specs(e, n_guides, guides, limits, best != 0, nullptr);
    std::vector<ExplicitWindow> windows((size_t)n_tasks);
    for (int64_t i = 0; i < n_tasks; ++i) {
      const calitas_region_task& t = tasks[i];
      if (t.guide_idx < 0 || t.guide_idx >= n_guides) throw InvalidArgument("task guide_idx out of range");
      if (t.contig_idx < 0 || t.contig_idx >= (int)ref->len.size()) throw InvalidArgument("Unknown chromosome index");
      const size_t c = (size_t)t.contig_idx;
      if (t.length < 0 || (uint32_t)t.length > MAX_WINDOW_LEN) throw InvalidArgument("region length out of range");
      if (t.start < ref->have_b[c] || t.start + t.length > ref->have_e[c]) throw InvalidArgument("region outside the loaded bases of contig " + ref->names[c]);
      windows[(size_t)i] = ExplicitWindow{ ref->nib_off[c] - ref->have_b[c] + t.start, t.length, (int32_t)t.start, t.guide_idx, t.contig_idx, (int32_t)i, 1 };
    }
    *out = run_explicit(e, ref->d_nib, ref->total_padded / 8, windows, specs);
    return CALITAS_OK;
  });
}

int calitas_align_targets(calitas_engine* e, int32_t n_guides, const calitas_guide* guides, int64_t n_tasks, const calitas_target_task* tasks,
                          const calitas_limits* limits, int32_t best, calitas_hitset** out) {
  return guarded([&]() -> int {
    if (!e || !out || n_tasks < 0 || (n_tasks && !tasks)) throw InvalidArgument("bad align_targets arguments");
    *out = nullptr;
    if (n_tasks >= (1ll << 31)) throw LimitExceeded("too many tasks");
    dev::set_device(e->device);
    std::vector<GuideSpec> specs = build_specs(e, n_guides, guides, limits, best != 0, nullptr);
    std::vector<ExplicitWindow> windows((size_t)n_tasks);
    int64_t total = 8;
    for (int64_t i = 0; i < n_tasks; ++i) {
      const calitas_target_task& t = tasks[i];
      if (t.guide_idx < 0 || t.guide_idx >= n_guides) throw InvalidArgument("task guide_idx out of range");
      if (t.length < 0 || (uint32_t)t.length > MAX_WINDOW_LEN || (t.length && !t.bases)) throw InvalidArgument("target length out of range");
      windows[(size_t)i] = ExplicitWindow{ total, t.length, t.target_offset, t.guide_idx, -1, (int32_t)i, 1 };
      total += (t.length + 7) / 8 * 8 + 8;
    }
    // the target bytes are gathered into one staging buffer on all host threads, uploaded and packed by k_pack, like the reference and the variant windows
    total = (total + 7) / 8 * 8 + 64;
    {
      std::vector<uint8_t> raw((size_t)total, 0);
      const int nt = (int)std::max<int64_t>(1, std::min<int64_t>(n_tasks / 65536 + 1, std::min<unsigned>(16u, std::max(1u, std::thread::hardware_concurrency()))));
      std::vector<std::thread> th;
      for (int t = 0; t < nt; ++t) th.emplace_back([&, t]() { for (int64_t i = n_tasks * t / nt; i < n_tasks * (t + 1) / nt; ++i) if (tasks[i].length) std::memcpy(raw.data() + windows[(size_t)i].nib_start, tasks[i].bases, (size_t)tasks[i].length); });
      for (auto& t : th) t.join();
      DBuf d_raw; struct Guard { DBuf& b; ~Guard() { try { b.release(); } catch (...) {} } } guard{ d_raw };
      d_raw.ensure((size_t)total); e->nib_tmp.ensure((size_t)total / 2 + 8);
      dev::h2d(d_raw.p, raw.data(), (size_t)total, e->stream);
      const int64_t n_words = total / 8;
      const unsigned grid = (unsigned)std::min<int64_t>((n_words + 255) / 256, (int64_t)dev::sm_count(e->device) * 16);
      CAL_LAUNCH(k_pack, grid, 256, 256, e->stream, 2, d_raw.as<uint8_t>(), e->nib_tmp.as<uint32_t>(), n_words); dev::launch_check("k_pack");
      dev::stream_sync(e->stream);
    }
    *out = run_explicit(e, e->nib_tmp.as<uint32_t>(), total / 8, windows, specs);
    return CALITAS_OK;
  });
}

int calitas_microbench_int(calitas_engine* e, int32_t kind, double* tera_ops_per_s) {
  return guarded([&]() -> int {
    if (!e || !tera_ops_per_s || kind < 0 || kind > 11) throw InvalidArgument("bad microbench arguments");
    dev::set_device(e->device); dev::Stream s = e->stream;
    const int iters = 4096; const unsigned grid = (unsigned)dev::sm_count(e->device) * 16, block = 256;
    e->tmp.ensure((size_t)grid * block * 4);
    double best = 0;
    for (int rep = 0; rep < 4; ++rep) {
      dev::event_record(e->ev[4], s);
      CAL_LAUNCH(k_int_peak, grid, block, 0, s, 1, e->tmp.as<uint32_t>(), iters, (int)kind, 12345u + (uint32_t)rep); dev::launch_check("k_int_peak");
      dev::event_record(e->ev[5], s); dev::stream_sync(s);
      const double ms = dev::event_ms(e->ev[4], e->ev[5]);
      const double ops = (double)grid * block * (double)iters * 16.0;     // 8 chains x 2 integer instructions per iteration, per thread
      if (rep > 0 && ms > 0) best = std::max(best, ops / (ms * 1e-3) / 1e12);
    }
    *tera_ops_per_s = best;
    return CALITAS_OK;
  });
}

int64_t calitas_hitset_count(const calitas_hitset* h) { return h ? h->n : 0; }
const calitas_hit* calitas_hitset_data(const calitas_hitset* h) { return h ? (const calitas_hit*)h->buf.p : nullptr; }
int32_t calitas_hitset_stride(const calitas_hitset* h) { return h ? h->stride : CALITAS_HIT_WORDS * 4; }
int calitas_reference_own_range(const calitas_reference* r, int32_t contig, int64_t* own_begin, int64_t* own_end) {
  if (!r || contig < 0 || contig >= (int32_t)r->len.size() || !own_begin || !own_end) return set_error(CALITAS_EINVAL, "bad arguments");
  *own_begin = r->own_b[(size_t)contig]; *own_end = r->own_e[(size_t)contig]; return CALITAS_OK;
}
void calitas_hitset_free(calitas_hitset* h) { if (!h) return; if (h->owner) { h->owner->pinned_pool.push_back(h->buf); if (h->info.p) h->owner->pinned_pool.push_back(h->info); } delete h; }
const calitas_variant_hit_info* calitas_hitset_variant_info(const calitas_hitset* h) { return h ? (const calitas_variant_hit_info*)h->info.p : nullptr; }
int calitas_hitset_stats(const calitas_hitset* h, double ms[8], int64_t counts[8]) {
  if (!h) return set_error(CALITAS_EINVAL, "hitset is NULL");
  for (int i = 0; i < 8; ++i) { if (ms) ms[i] = h->ms[i]; if (counts) counts[i] = h->counts[i]; }
  return CALITAS_OK;
}

int calitas_render_alignments(const void* hits, int64_t n_hits, int32_t stride, int32_t n_guides, const calitas_guide* guides, int32_t n_contigs, const char* const* names,
                              const uint8_t* const* contig_bases, const calitas_target_task* targets, int32_t upper_case, char** out_text) {
  return guarded([&]() -> int {
    if (!out_text || (n_hits && !hits) || !guides || (stride != CALITAS_HIT_WORDS * 4 && stride != CALITAS_HIT_WIDE_WORDS * 4)) throw InvalidArgument("bad render arguments");
    *out_text = nullptr;
    std::vector<GuideDef> defs; for (int i = 0; i < n_guides; ++i) defs.push_back(parse_guide(guides[i]));
    std::string text = alignment_header();
    for (int64_t i = 0; i < n_hits; ++i) {
      HitX h; unpack_hit(reinterpret_cast<const uint32_t*>((const char*)hits + (size_t)i * stride), stride / 4, h);
      if (h.guide_idx < 0 || h.guide_idx >= n_guides) throw InvalidArgument("hit guide_idx out of range");
      std::string fwd, chrom = "n/a";
      const int32_t len = h.end_offset - h.start_offset;
      if (h.contig_idx >= 0) {
        if (h.contig_idx >= n_contigs || !contig_bases || !contig_bases[h.contig_idx]) throw InvalidArgument("contig bases missing for a hit");
        fwd.assign((const char*)contig_bases[h.contig_idx] + h.start_offset, (size_t)len);
        if (names && names[h.contig_idx]) chrom = names[h.contig_idx];
      } else {
        if (!targets) throw InvalidArgument("targets missing for align_targets hits");
        const calitas_target_task& t = targets[h.task_idx];
        fwd.assign((const char*)t.bases + (h.start_offset - t.target_offset), (size_t)len);
      }
      Rendered r = render_hit(h, defs[(size_t)h.guide_idx], fwd, upper_case != 0);
      text += alignment_row(h, r, chrom);
    }
    char* p = (char*)std::malloc(text.size() + 1); std::memcpy(p, text.data(), text.size() + 1); *out_text = p;
    return CALITAS_OK;
  });
}
void calitas_free_text(char* text) { std::free(text); }

}  // extern "C"

#ifdef CAL_HOSTSIM
namespace cal { namespace sim { thread_local Dim3 threadIdx_, blockIdx_, blockDim_, gridDim_; thread_local int phase_ = 0; thread_local unsigned char smem_[256 * 1024]; } }
#endif
