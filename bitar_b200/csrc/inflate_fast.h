// inflate_fast.h -- raw DEFLATE (RFC 1951) decoder, one LANE per stream, for the sm_100a inflate kernel.
//
// Why a lane per stream (measured on B200, DESIGN.md "Inflate"): a DEFLATE stream decodes serially, so
// a 1 GiB buffer cut into 59 460-byte segments offers ~18 k independent dependency chains and nothing
// else.  A warp per stream spends 32 issue lanes on one chain (issue-bound near 33 GB/s); here every lane
// of a warp owns a stream, all ~18 k chains are in flight at once, and throughput is set by how SHORT the
// per-symbol chain is.  The step below is therefore written for the fewest instructions per symbol:
//   * u16 decode tables (root + second level) per lane in shared memory, one LDS per symbol,
//   * up to four literals, then one match, per step: lanes of a warp that walk the same kind of column
//     stay converged,
//   * output bytes go to a short per-lane ring in shared memory and leave as aligned 16-byte vector
//     stores the moment a vector completes; matches inside the ring never touch HBM, farther ones read
//     the lane's own earlier output (L1/L2),
//   * capacity is checked once per 16 output bytes, not per literal,
//   * block headers, stored blocks and error paths are cold, out-of-line code.
//
// BITAR_HD: the same source is compiled for the CPU with one lane (tools/model/core_host.cc,
// tests/test_core_host.py), so header parsing, table construction and every error path run without a GPU.
//
// Replaces the inflate half of the codec behind rte_compressdev (/root/reference/src/device.cc:240-318,
// decompress xform at src/config.cc:93-105); per-op status as consumed at src/device.cc:512-520.
#pragma once
#include <stdint.h>

#include "../../include/bitar_cuda.h"
#include "checksum.h"
#include "deflate_common.h"
#include "inflate_core.h"  // status codes, ld_in8 / ld_in32, brev32

namespace bitar {
namespace fl {

using inf::kStatusDataError;
using inf::kStatusOk;
using inf::kStatusOutOfSpace;
using inf::kStatusTruncated;

// ---- u16 table entries ----------------------------------------------------------------------------------
// litlen: [3:0] code length, [7:4] kind/extra, [15:8] value
//   literal      : [7:4] = 0,            value = byte
//   length       : [7] = 1, [6:4] = number of extra bits (0..5), value = base length - 3
//   end of block : [7:4] = 0xE
//   link         : [7:4] = 0xF, [3:0] = index bits of the second-level table (0 = none: canonical slow
//                  path / invalid code), value = (second-level offset - root size) / 4
// dist  : [3:0] code length, [8:4] distance symbol 0..29; symbol 31 = link ([15:9] (offset - root size) / 4)
constexpr uint32_t kBadEntry = 0x00F0u;    // "no such code"
constexpr uint32_t kNoEntry = 0x10F0u;     // step(): nothing pending after the literal run
constexpr uint32_t kBadDist = 31u << 4;

BITAR_HD uint16_t ll_entry(int sym, int nbits) {
  if (sym < 256) return (uint16_t)((sym << 8) | nbits);
  if (sym == 256) return (uint16_t)(0xE0 | nbits);
  if (sym < 286) {
    const int s = sym - 257;
    return (uint16_t)(((dfl::len_base(s) - 3) << 8) | 0x80 | (dfl::len_extra_bits(s) << 4) | nbits);
  }
  return (uint16_t)kBadEntry;  // 286 / 287 never appear in valid data
}
BITAR_HD uint16_t d_entry(int sym, int nbits) { return sym < 30 ? (uint16_t)((sym << 4) | nbits) : (uint16_t)kBadDist; }
BITAR_HD uint16_t ll_link(int off, int sub_bits) { return (uint16_t)(((off >> 2) << 8) | 0xF0 | sub_bits); }
BITAR_HD uint16_t d_link(int off, int sub_bits) { return (uint16_t)(((off >> 2) << 9) | kBadDist | sub_bits); }

// distance symbol -> base | extra bits << 16 (shared by the lanes of a CTA)
BITAR_HD uint32_t dist_info(int sym) {
  return sym < 30 ? ((uint32_t)dfl::dist_base(sym) | ((uint32_t)dfl::dist_extra_bits(sym) << 16)) : 0u;
}

// per-lane global scratch: header parsing and the canonical slow path (cold)
struct LaneScratch {
  uint8_t lens[320];
  uint16_t ll_sorted[288];
  uint16_t d_sorted[32];
  uint16_t ll_count[16], ll_first[16], ll_offs[16];
  uint16_t d_count[16], d_first[16], d_offs[16];
};

enum TableKind { kLitLen = 0, kDist = 1, kCodeLen = 2 };

// Root table of 2^tbits entries followed by second-level tables (up to `capacity` entries in total) for
// codes longer than tbits; codes whose second level does not fit keep a "link with 0 bits" root entry and
// are decoded canonically from count/first/offs/sorted.  Validity rules are zlib's inflate_table:
// over-subscribed sets are errors, incomplete sets are errors except a single 1-bit code, an all-zero
// alphabet is accepted (any use of it is then an error).
BITAR_HD_NOINLINE uint32_t build_table(const uint8_t* lens, int n, int kind, uint16_t* table, int tbits, int capacity,
                                       uint16_t* count, uint16_t* first, uint16_t* offs, uint16_t* sorted) {
  for (int b = 0; b < 16; ++b) count[b] = 0;
  for (int i = 0; i < n; ++i) count[lens[i]]++;
  int left = 1, maxl = 0;
  for (int b = 1; b <= 15; ++b) {
    left = (left << 1) - (int)count[b];
    if (count[b]) maxl = b;
    if (left < 0) return kStatusDataError;
  }
  const int used = n - (int)count[0];
  if (left > 0 && used > 0 && (kind == kCodeLen || maxl != 1)) return kStatusDataError;
  uint32_t code0 = 0, o = 0;
  uint16_t at[16];
  first[0] = offs[0] = at[0] = 0;
  for (int b = 1; b <= 15; ++b) {
    first[b] = (uint16_t)code0;
    offs[b] = at[b] = (uint16_t)o;
    code0 = (code0 + count[b]) << 1;
    o += count[b];
  }
  for (int i = 0; i < n; ++i)
    if (lens[i]) sorted[at[lens[i]]++] = (uint16_t)i;
  const uint32_t fill = kind == kLitLen ? kBadEntry : kind == kDist ? kBadDist : 0u;
  uint32_t* t32 = reinterpret_cast<uint32_t*>(table);
  for (int j = 0; j < capacity / 2; ++j) t32[j] = fill | (fill << 16);
  int idx = 0;
  for (; idx < used; ++idx) {   // canonical order: by length, then by symbol
    const int sym = sorted[idx], l = lens[sym];
    if (l > tbits) break;
    const uint32_t code = (uint32_t)first[l] + (uint32_t)(idx - (int)offs[l]);
    const uint32_t r = inf::brev32(code) >> (32 - l);
    const uint16_t e = kind == kLitLen ? ll_entry(sym, l) : kind == kDist ? d_entry(sym, l) : (uint16_t)((sym << 4) | l);
    for (uint32_t k = r; k < (1u << tbits); k += (1u << l)) table[k] = e;
  }
  int next_free = 1 << tbits;
  while (idx < used) {          // longer codes: those sharing a tbits-bit prefix are contiguous
    const int l = lens[sorted[idx]];
    const uint32_t prefix = ((uint32_t)first[l] + (uint32_t)(idx - (int)offs[l])) >> (l - tbits);
    int j = idx, lmax = l;
    while (j < used) {
      const int l2 = lens[sorted[j]];
      const uint32_t c2 = (uint32_t)first[l2] + (uint32_t)(j - (int)offs[l2]);
      if ((c2 >> (l2 - tbits)) != prefix) break;
      lmax = l2;
      ++j;
    }
    int sub_bits = lmax - tbits;
    if (sub_bits < 2) sub_bits = 2;   // offsets are stored / 4
    const int size = 1 << sub_bits;
    if (next_free + size <= capacity) {
      const int rel = next_free - (1 << tbits);   // links hold the offset past the root, / 4
      table[inf::brev32(prefix) >> (32 - tbits)] = kind == kLitLen ? ll_link(rel, sub_bits) : d_link(rel, sub_bits);
      for (int k = idx; k < j; ++k) {
        const int sym = sorted[k], lk = lens[sym], rest = lk - tbits;
        const uint32_t ck = (uint32_t)first[lk] + (uint32_t)(k - (int)offs[lk]);
        const uint32_t r = inf::brev32(ck & ((1u << rest) - 1u)) >> (32 - rest);
        const uint16_t e = kind == kLitLen ? ll_entry(sym, lk) : d_entry(sym, lk);
        for (int t = (int)r; t < size; t += (1 << rest)) table[next_free + t] = e;
      }
      next_free += size;
    }
    idx = j;
  }
  return kStatusOk;
}

// canonical bit-by-bit decode: entry with the full code length, or the kind's "bad" entry
BITAR_HD_NOINLINE uint32_t canonical_decode(uint32_t bits, int kind, const uint16_t* count, const uint16_t* first,
                                            const uint16_t* offs, const uint16_t* sorted) {
  uint32_t code = 0;
  for (int l = 1; l <= 15; ++l) {
    code = (code << 1) | ((bits >> (l - 1)) & 1u);
    const uint32_t rel = code - (uint32_t)first[l];
    if (code >= first[l] && rel < count[l]) {
      const int sym = sorted[offs[l] + rel];
      return kind == kLitLen ? ll_entry(sym, l) : d_entry(sym, l);
    }
  }
  return kind == kLitLen ? kBadEntry : kBadDist;
}

// ---- shared-memory accessors: 32-bit shared addresses + LDS/STS on the device, pointers on the host ----
#if defined(__CUDA_ARCH__)
typedef uint32_t sptr;
BITAR_HD sptr sp_of(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
BITAR_HD uint32_t s_ld16(sptr a) { uint32_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
BITAR_HD uint32_t s_ld32(sptr a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
BITAR_HD uint32_t s_ld8(sptr a) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
BITAR_HD void s_st8(sptr a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v)); }
BITAR_HD void s_ld128(sptr a, uint32_t& w0, uint32_t& w1, uint32_t& w2, uint32_t& w3) {
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3) : "r"(a));
}
BITAR_HD uint32_t fsr(uint32_t lo, uint32_t hi, uint32_t n) { return __funnelshift_r(lo, hi, n); }
BITAR_HD uint32_t fsl_hi(uint32_t v, uint32_t n) { return __funnelshift_l(v, 0u, n); }   // v >> (32 - n), 0 for n = 0
#else
typedef const uint8_t* sptr;
BITAR_HD sptr sp_of(const void* p) { return static_cast<const uint8_t*>(p); }
BITAR_HD uint32_t s_ld16(sptr a) { return *reinterpret_cast<const uint16_t*>(a); }
BITAR_HD uint32_t s_ld32(sptr a) { return *reinterpret_cast<const uint32_t*>(a); }
BITAR_HD uint32_t s_ld8(sptr a) { return *a; }
BITAR_HD void s_st8(sptr a, uint32_t v) { *const_cast<uint8_t*>(a) = (uint8_t)v; }
BITAR_HD void s_ld128(sptr a, uint32_t& w0, uint32_t& w1, uint32_t& w2, uint32_t& w3) {
  const uint32_t* p = reinterpret_cast<const uint32_t*>(a);
  w0 = p[0]; w1 = p[1]; w2 = p[2]; w3 = p[3];
}
BITAR_HD uint32_t fsr(uint32_t lo, uint32_t hi, uint32_t n) { return n ? (lo >> n) | (hi << (32u - n)) : lo; }
BITAR_HD uint32_t fsl_hi(uint32_t v, uint32_t n) { return n ? v >> (32u - n) : 0u; }
#endif

// ---- the parallel-inflate index (deflate_common.h) as found at the end of a chunk ------------------------
struct IndexInfo {
  const uint8_t* words;     // first index word (unaligned)
  uint32_t entries;         // hdr_bit / sub_bit words
  uint32_t end_bit;         // bit offset just past the final end-of-block symbol
  uint32_t total_out;       // uncompressed bytes
  uint32_t stream_bytes;    // bytes of the DEFLATE stream proper
};
BITAR_HD uint32_t load_u32le(const uint8_t* p) {
  return (uint32_t)inf::ld_in8(p) | ((uint32_t)inf::ld_in8(p + 1) << 8) | ((uint32_t)inf::ld_in8(p + 2) << 16) |
         ((uint32_t)inf::ld_in8(p + 3) << 24);
}
// True when the chunk ends in a structurally consistent index: magic, sizes that add up to src_len exactly.
BITAR_HD bool parse_index(const uint8_t* src, uint32_t src_len, IndexInfo* ix) {
  if (src_len < 16u || load_u32le(src + src_len - 4u) != dfl::kIndexMagic) return false;
  ix->total_out = load_u32le(src + src_len - 8u);
  ix->end_bit = load_u32le(src + src_len - 12u);
  if (ix->total_out <= dfl::kSub || ix->total_out > BITAR_MAX_SEG_SIZE) return false;
  ix->entries = dfl::idx_entries(ix->total_out);
  const uint32_t nbytes = 4u * (ix->entries + 3u);
  ix->stream_bytes = (ix->end_bit + 7u) >> 3;
  if (ix->end_bit < 10u || (uint64_t)ix->stream_bytes + nbytes != (uint64_t)src_len) return false;
  ix->words = src + ix->stream_bytes;
  return true;
}
BITAR_HD uint32_t index_word(const IndexInfo& ix, uint32_t k) { return load_u32le(ix.words + 4u * k); }

// Bit offsets of block b / of its sub-range s as the index gives them, checked against each other and against
// the stream: nothing derived from a damaged index may point outside the DEFLATE stream.
struct BlockBits {
  uint32_t hdr, end;   // first bit of the block header; first bit after the block
};
BITAR_HD bool index_block_bits(const IndexInfo& ix, uint32_t b, uint32_t n_blocks, BlockBits* bb) {
  bb->hdr = index_word(ix, b * 33u);
  bb->end = b + 1u < n_blocks ? index_word(ix, (b + 1u) * 33u) : ix.end_bit;
  return bb->hdr < bb->end && bb->end <= ix.end_bit;
}
BITAR_HD bool index_sub_bits(const IndexInfo& ix, uint32_t b, uint32_t s, uint32_t n_subs, const BlockBits& bb, uint32_t* sbit,
                             uint32_t* ebit) {
  *sbit = index_word(ix, b * 33u + 1u + s);
  *ebit = s + 1u < n_subs ? index_word(ix, b * 33u + 2u + s) : bb.end;
  return *sbit >= bb.hdr && *sbit <= *ebit && *ebit <= bb.end;
}

// CTA-shared constants: distance symbol info and (when checksums are on) four CRC-32 slicing tables
struct CtaTables {
  uint32_t dinfo[32];
  uint32_t crc[4][256];
  uint32_t x2n[32];       // x^(2^k) mod P, for combining the partial CRCs of sub-ranges (checksum.h)
};

// LT / DT: total u16 entries of the litlen / distance tables (root + second level)
template <int LBITS, int LT, int DBITS, int DT, int RING>
struct LaneLayout {
  static_assert(LT >= (1 << LBITS) && DT >= (1 << DBITS) && LT % 8 == 0 && DT % 8 == 0, "table sizes");
  static_assert(RING >= 256 && (RING & (RING - 1)) == 0, "ring: power of two >= 256");
  static constexpr int kBytes = 2 * LT + 2 * DT + RING;
  // 16-byte aligned lanes; stride = 16 (mod 128) so that the 8 lanes of a quarter warp cover all 32
  // banks with their 16-byte ring vectors
  static constexpr int kStride = ((kBytes + 127) / 128) * 128 + 16;
};

// SUB = false: a whole stream per lane (any valid raw DEFLATE: header parsing and table construction are
//               part of the lane's state machine, tables are private to the lane).
// SUB = true : one 2 KiB sub-range of a block per lane, located through the parallel-inflate index
//               (deflate_common.h): the decode tables belong to the warp (built once per block by
//               inflate_indexed_kernel.cuh), the lane starts at a given bit offset, must produce exactly
//               the sub-range and must end exactly where the index says the next sub-range starts.
// CK = false compiles the checksum folding out (the kernels pick it when no checksum is configured).
template <int LBITS, int LT, int DBITS, int DT, int RING, bool SUB = false, bool CK = true>
struct FastLane {
  static_assert(DBITS >= 7, "the distance table also hosts the 7-bit code-length code");
  static_assert(RING >= 128 && (RING & (RING - 1)) == 0, "ring: power of two >= 128");
  static constexpr uint32_t RM = RING - 1;
  static constexpr uint32_t kPiece = RING / 2;   // bytes copied between two flushes (<= RING - 16 - 19)
  static constexpr uint32_t LMASK = (1u << LBITS) - 1u, DMASK = (1u << DBITS) - 1u;
  enum : uint32_t { kHeader = 0, kDecode = 1, kStored = 2, kFinish = 3, kDone = 4, kSubEnd = 5 };

  // shared memory of this lane
  uint16_t* lt;
  uint16_t* dt;
  sptr lt_s, dt_s, ring_s, dinfo_s, crc_s;
  LaneScratch* sc;
  // input: 64-bit bit buffer (lo, hi), refilled one aligned 32-bit word at a time, one word prefetched
  const uint8_t* in;
  const uint32_t* words;
  uint32_t in_len, nwords, wpos, next, skip, start_off;
  uint32_t lo, hi, cnt;
  // output: "virtual" positions v = offset + (dst & 15), so that v % 16 == address % 16
  uint8_t* vbase;
  uint32_t vstart, vpos, vflushed, vcap;
  uint32_t state, status, last, blocks;
  uint32_t stored_rem, stored_at;
  uint32_t sub_end_bit, sub_eob;     // SUB: where the sub-range must end (stream bit offset); block ends here
  // checksums of the produced bytes (folded in as the bytes leave the ring)
  uint32_t ck_type, crc, ad_a;
  uint64_t ad_b;

  BITAR_HD void bind(uint8_t* smem_lane, const CtaTables* cta, LaneScratch* scratch, uint32_t checksum_type) {
    bind_parts(reinterpret_cast<uint16_t*>(smem_lane), reinterpret_cast<uint16_t*>(smem_lane + 2 * LT),
               smem_lane + 2 * LT + 2 * DT, cta, scratch, checksum_type);
  }
  // tables and ring given separately (SUB: the tables are the warp's, the ring is the lane's)
  BITAR_HD void bind_parts(uint16_t* lt_, uint16_t* dt_, uint8_t* ring_, const CtaTables* cta, LaneScratch* scratch,
                           uint32_t checksum_type) {
    lt = lt_;
    dt = dt_;
    lt_s = sp_of(lt_);
    dt_s = sp_of(dt_);
    ring_s = sp_of(ring_);
    dinfo_s = sp_of(cta->dinfo);
    crc_s = sp_of(cta->crc);
    sc = scratch;
    ck_type = checksum_type;
    state = kDone;
    status = kStatusOk;
    in = nullptr;
    words = nullptr;
    vbase = nullptr;
    in_len = nwords = wpos = next = skip = start_off = lo = hi = cnt = 0;
    vstart = vpos = vflushed = vcap = last = blocks = stored_rem = stored_at = 0;
    sub_end_bit = sub_eob = 0;
    crc = ad_a = 0;
    ad_b = 0;
  }

  // SUB: decode `len` bytes to dst from the symbol at stream bit `start_bit`; the sub-range must end at
  // `end_bit`, after an end-of-block symbol when `eob` (it is the last sub-range of its block).
  // `first`: the sub-range starts the chunk (its CRC register starts at ~0; all partial sums are combined by the
  // kernel with cks::crc_contrib / adler_b_contrib).
  BITAR_HD void start_sub(const uint8_t* src, uint32_t stream_len, uint32_t start_bit, uint32_t end_bit, bool eob,
                          uint8_t* dst, uint32_t len, bool first = false) {
    in = src;
    in_len = stream_len;
    bits_init(start_bit >> 3);
    drop(start_bit & 7u);
    const uint32_t mis = (uint32_t)((uintptr_t)dst & 15u);
    vbase = dst - mis;
    vstart = vpos = vflushed = mis;
    vcap = mis + len;
    state = len ? (uint32_t)kDecode : (uint32_t)kSubEnd;
    status = kStatusOk;
    last = blocks = stored_rem = 0;
    sub_end_bit = end_bit;
    sub_eob = eob ? 1u : 0u;
    crc = first ? 0xFFFFFFFFu : 0u;   // partial sums: combined by the kernel (checksum.h)
    ad_a = 0;
    ad_b = 0;
  }

  BITAR_HD void start(const uint8_t* src, uint32_t len, uint8_t* dst, uint32_t cap) {
    in = src;
    in_len = len;
    bits_init(0);
    const uint32_t mis = (uint32_t)((uintptr_t)dst & 15u);
    vbase = dst - mis;
    vstart = vpos = vflushed = mis;
    vcap = mis + cap;
    state = kHeader;
    status = kStatusOk;
    last = blocks = stored_rem = 0;
    crc = 0xFFFFFFFFu;
    ad_a = 1;
    ad_b = 0;
  }
  BITAR_HD uint32_t produced() const { return vpos - vstart; }
  BITAR_HD uint64_t checksum() const {
    const uint32_t n = produced();
    const uint32_t c = (ck_type & BITAR_CHECKSUM_CRC32) ? (n ? crc ^ 0xFFFFFFFFu : 0u) : 0u;
    const uint32_t a = (ck_type & BITAR_CHECKSUM_ADLER32)
                           ? (((uint32_t)(ad_b % cks::kAdlerMod) << 16) | (ad_a % cks::kAdlerMod)) : 0u;
    return cks::pack(c, a);
  }

  // ---- bit reader ----
  BITAR_HD void bits_init(uint32_t off) {
    const uint8_t* a = in + off;
    start_off = off;
    const uint32_t mis = (uint32_t)((uintptr_t)a & 3u);
    words = reinterpret_cast<const uint32_t*>(a - mis);
    const uint32_t bytes = off < in_len ? in_len - off : 0u;
    nwords = bytes ? (mis + bytes + 3u) >> 2 : 0u;
    const uint32_t w0 = nwords ? inf::ld_in32(words) : 0u;
    lo = w0 >> (8u * mis);
    hi = 0;
    cnt = 32u - 8u * mis;
    skip = 8u * mis;
    next = nwords > 1 ? inf::ld_in32(words + 1) : 0u;
    wpos = 2;
  }
  BITAR_HD void refill() {   // afterwards cnt is in [32, 63]
    if (cnt < 32u) {
      lo |= next << cnt;
      hi = fsl_hi(next, cnt);
      cnt += 32u;
      next = wpos < nwords ? inf::ld_in32(words + wpos) : 0u;
      wpos++;
    }
  }
  BITAR_HD void drop(uint32_t n) {   // n < 32
    lo = fsr(lo, hi, n);
    hi >>= n;
    cnt -= n;
  }
  BITAR_HD uint32_t take(uint32_t n) {   // n <= 16
    const uint32_t v = lo & ((1u << n) - 1u);
    drop(n);
    return v;
  }
  BITAR_HD int64_t consumed_bits() const { return 32ll * ((int64_t)wpos - 1) - (int64_t)skip - (int64_t)cnt; }
  BITAR_HD bool overrun() const { return consumed_bits() > 8ll * ((int64_t)in_len - (int64_t)start_off); }
  BITAR_HD uint32_t consumed_bytes() const {
    const int64_t used = (int64_t)start_off + ((consumed_bits() + 7) >> 3);
    return used > (int64_t)in_len ? in_len : (uint32_t)used;
  }

  BITAR_HD void fail(uint32_t st) {
    if (status == kStatusOk) status = st;
    state = kFinish;
  }

  // ---- checksum folding ----
  BITAR_HD void ck_byte(uint32_t b) {
    if (ck_type & BITAR_CHECKSUM_CRC32) crc = s_ld32(crc_s + (((crc ^ b) & 0xFFu) << 2)) ^ (crc >> 8);
    if (ck_type & BITAR_CHECKSUM_ADLER32) {
      ad_a += b;
      ad_b += ad_a;
    }
  }
  BITAR_HD void ck_word(uint32_t w) {
    if (ck_type & BITAR_CHECKSUM_CRC32) {   // slicing-by-4
      const uint32_t x = crc ^ w;
      crc = s_ld32(crc_s + 3072u + ((x & 0xFFu) << 2)) ^ s_ld32(crc_s + 2048u + (((x >> 8) & 0xFFu) << 2)) ^
            s_ld32(crc_s + 1024u + (((x >> 16) & 0xFFu) << 2)) ^ s_ld32(crc_s + ((x >> 24) << 2));
    }
    if (ck_type & BITAR_CHECKSUM_ADLER32) {
      const uint32_t b0 = w & 0xFFu, b1 = (w >> 8) & 0xFFu, b2 = (w >> 16) & 0xFFu, b3 = w >> 24;
      ad_b += 4ull * ad_a + (uint64_t)(4u * b0 + 3u * b1 + 2u * b2 + b3);
      ad_a += b0 + b1 + b2 + b3;
      if (ad_a >= 0x80000000u) ad_a %= cks::kAdlerMod;
    }
  }

  // ---- output ----
  BITAR_HD void emit(uint32_t byte) {
    s_st8(ring_s + (vpos & RM), byte);
    vpos++;
  }

  // Store every complete 16-byte vector below vpos (and the unaligned head of the chunk).
  BITAR_HD void flush() {
    if (vpos > vcap) {   // the capacity check of the literal path lives here: at most 16 + 4 bytes late
      vpos = vcap;
      fail(SUB ? kStatusDataError : overrun() ? kStatusTruncated : kStatusOutOfSpace);
    }
    while (vpos - vflushed >= 16u || ((vflushed & 15u) && vpos >= ((vflushed + 15u) & ~15u))) {
      if (vflushed & 15u) {
        const uint32_t a = (vflushed + 15u) & ~15u;
        for (uint32_t v = vflushed; v < a; ++v) {
          const uint32_t b = s_ld8(ring_s + (v & RM));
          vbase[v] = (uint8_t)b;
          if (CK && ck_type) ck_byte(b);
        }
        vflushed = a;
        continue;
      }
      uint32_t w0, w1, w2, w3;
      s_ld128(ring_s + (vflushed & RM), w0, w1, w2, w3);
#if defined(__CUDA_ARCH__)
      *reinterpret_cast<uint4*>(vbase + vflushed) = make_uint4(w0, w1, w2, w3);
#else
      uint32_t* o32 = reinterpret_cast<uint32_t*>(vbase + vflushed);
      o32[0] = w0; o32[1] = w1; o32[2] = w2; o32[3] = w3;
#endif
      if (CK && ck_type) {
        ck_word(w0); ck_word(w1); ck_word(w2); ck_word(w3);
      }
      vflushed += 16u;
    }
  }

  BITAR_HD void finish() {
    flush();
    for (uint32_t v = vflushed; v < vpos; ++v) {
      const uint32_t b = s_ld8(ring_s + (v & RM));
      vbase[v] = (uint8_t)b;
      if (CK && ck_type) ck_byte(b);
    }
    vflushed = vpos;
    state = kDone;
  }

  // LZ77 copy (1 <= dist <= produced(), vpos + len <= vcap), in pieces that fit the ring
  BITAR_HD void copy(uint32_t len, uint32_t dist) {
    for (;;) {
      const uint32_t piece = len < kPiece ? len : kPiece;
      const uint32_t src = vpos - dist;
      uint32_t j = 0;
      if (dist < (uint32_t)RING) {           // the source is still in the ring
        if (dist >= 4u) {
          for (; j + 4u <= piece; j += 4u) {
            const uint32_t b0 = s_ld8(ring_s + ((src + j) & RM)), b1 = s_ld8(ring_s + ((src + j + 1u) & RM));
            const uint32_t b2 = s_ld8(ring_s + ((src + j + 2u) & RM)), b3 = s_ld8(ring_s + ((src + j + 3u) & RM));
            s_st8(ring_s + ((vpos + j) & RM), b0);
            s_st8(ring_s + ((vpos + j + 1u) & RM), b1);
            s_st8(ring_s + ((vpos + j + 2u) & RM), b2);
            s_st8(ring_s + ((vpos + j + 3u) & RM), b3);
          }
        }
        for (; j < piece; ++j) s_st8(ring_s + ((vpos + j) & RM), s_ld8(ring_s + ((src + j) & RM)));
      } else {                               // flushed long ago: dist >= RING, so src + piece <= vflushed
        const uint8_t* g = vbase + src;
        for (; j + 4u <= piece; j += 4u) {
          const uint32_t b0 = g[j], b1 = g[j + 1u], b2 = g[j + 2u], b3 = g[j + 3u];
          s_st8(ring_s + ((vpos + j) & RM), b0);
          s_st8(ring_s + ((vpos + j + 1u) & RM), b1);
          s_st8(ring_s + ((vpos + j + 2u) & RM), b2);
          s_st8(ring_s + ((vpos + j + 3u) & RM), b3);
        }
        for (; j < piece; ++j) s_st8(ring_s + ((vpos + j) & RM), g[j]);
      }
      vpos += piece;
      len -= piece;
      if (len == 0) return;
      flush();
    }
  }

  // ---- cold paths ----
  BITAR_HD uint32_t ll_resolve(uint32_t e) {
    const uint32_t sb = e & 15u;
    if (sb) {
      e = s_ld16(lt_s + 2u * ((1u << LBITS) + ((e >> 8) << 2) + ((lo >> LBITS) & ((1u << sb) - 1u))));
      if ((e & 0xF0u) != 0xF0u) return e;
    }
    return canonical_decode(lo, kLitLen, sc->ll_count, sc->ll_first, sc->ll_offs, sc->ll_sorted);
  }
  BITAR_HD uint32_t d_resolve(uint32_t d) {
    const uint32_t sb = d & 15u;
    if (sb) {
      d = s_ld16(dt_s + 2u * ((1u << DBITS) + ((d >> 9) << 2) + ((lo >> DBITS) & ((1u << sb) - 1u))));
      if ((d & kBadDist) != kBadDist) return d;
    }
    return canonical_decode(lo, kDist, sc->d_count, sc->d_first, sc->d_offs, sc->d_sorted);
  }

  BITAR_HD void header() {
    if (last) {
      state = kFinish;
      return;
    }
    refill();
    last = take(1);
    const uint32_t type = take(2);
    blocks++;
    if (overrun()) return fail(kStatusTruncated);
    if (type == 0) {
      drop(cnt & 7u);   // to the byte boundary (cnt and the bit position share parity mod 8)
      refill();
      const uint32_t len = take(16);
      refill();
      const uint32_t nlen = take(16);
      if (overrun()) return fail(kStatusTruncated);
      if ((len ^ 0xFFFFu) != nlen) return fail(kStatusDataError);
      const uint32_t at = start_off + (uint32_t)(consumed_bits() >> 3);
      if ((uint64_t)at + len > in_len) return fail(kStatusTruncated);
      if (vpos + len > vcap) return fail(kStatusOutOfSpace);
      stored_rem = len;
      stored_at = at;
      state = kStored;
      return;
    }
    if (type == 3) return fail(kStatusDataError);
    int nlen, ndist;
    if (type == 1) {
      for (int i = 0; i < 288; ++i) sc->lens[i] = (uint8_t)dfl::fixed_ll_len(i);
      for (int i = 0; i < 32; ++i) sc->lens[288 + i] = 5;
      nlen = 288;
      ndist = 32;
    } else {
      nlen = (int)take(5) + 257;
      ndist = (int)take(5) + 1;
      const int ncode = (int)take(4) + 4;
      if (nlen > 286 || ndist > 30) return fail(kStatusDataError);
      for (int i = 0; i < 19; ++i) sc->lens[i] = 0;
      for (int i = 0; i < ncode; ++i) {
        refill();
        sc->lens[dfl::cl_order(i)] = (uint8_t)take(3);
      }
      if (overrun()) return fail(kStatusTruncated);
      uint32_t st = build_table(sc->lens, 19, kCodeLen, dt, 7, 128, sc->d_count, sc->d_first, sc->d_offs, sc->d_sorted);
      if (st != kStatusOk) return fail(st);
      int idx = 0, prev = 0;
      const int total = nlen + ndist;
      while (idx < total) {
        refill();
        const uint32_t e = dt[lo & 127u];
        if ((e & 15u) == 0) return fail(kStatusDataError);
        drop(e & 15u);
        const int sym = (int)(e >> 4);
        int rep, val;
        if (sym < 16) { rep = 1; val = sym; prev = sym; }
        else if (sym == 16) {
          if (idx == 0) return fail(kStatusDataError);
          rep = 3 + (int)take(2); val = prev;
        } else if (sym == 17) { rep = 3 + (int)take(3); val = 0; prev = 0; }
        else { rep = 11 + (int)take(7); val = 0; prev = 0; }
        if (idx + rep > total) return fail(kStatusDataError);
        for (int k = 0; k < rep; ++k) sc->lens[idx + k] = (uint8_t)val;
        idx += rep;
      }
      if (overrun()) return fail(kStatusTruncated);
      if (sc->lens[256] == 0) return fail(kStatusDataError);
    }
    uint32_t st = build_table(sc->lens + nlen, ndist, kDist, dt, DBITS, DT, sc->d_count, sc->d_first, sc->d_offs, sc->d_sorted);
    if (st != kStatusOk) return fail(st);
    st = build_table(sc->lens, nlen, kLitLen, lt, LBITS, LT, sc->ll_count, sc->ll_first, sc->ll_offs, sc->ll_sorted);
    if (st != kStatusOk) return fail(st);
#if defined(__CUDA_ARCH__)
    asm volatile("" ::: "memory");   // table stores (generic pointers) stay ahead of the LDS look-ups
#endif
    state = kDecode;
  }

  BITAR_HD void stored() {   // the payload of a stored block, kPiece bytes per call
    const uint32_t m = stored_rem < kPiece ? stored_rem : kPiece;
    for (uint32_t j = 0; j < m; ++j) s_st8(ring_s + ((vpos + j) & RM), inf::ld_in8(in + stored_at + j));
    vpos += m;
    stored_at += m;
    stored_rem -= m;
    flush();
    if (stored_rem == 0 && state == kStored) {
      bits_init(stored_at);
      state = kHeader;
    }
  }

  // SUB: the sub-range is complete -- it must have ended exactly where the index says
  BITAR_HD void sub_end() {
    state = kFinish;
    if (vpos != vcap) return fail(kStatusDataError);
    if (sub_eob) {
      refill();
      const uint32_t e = ll_lookup();
      if ((e & 0xF0u) != 0xE0u) return fail(kStatusDataError);
      drop(e & 15u);
    }
    if (8ll * (int64_t)start_off + consumed_bits() != (int64_t)sub_end_bit) fail(kStatusDataError);
  }

  // SUB: fewer than 5 bytes left -- one symbol at a time, so that the lane stops exactly at the end
  BITAR_HD void tail_step() {
    if (vpos >= vcap) {
      state = kSubEnd;
      return;
    }
    refill();
    const uint32_t e = ll_lookup();
    if ((e & 0xF0u) == 0) {
      drop(e & 15u);
      emit(e >> 8);
    } else if ((e & 0x80u) && (e & 0x70u) < 0x60u) {
      match(e);
    } else {
      fail(kStatusDataError);   // end of block (or no such code) inside a sub-range
    }
  }

  BITAR_HD void cold() {
    if (SUB) {
      if (state == kSubEnd) sub_end();
      else if (state == kFinish) finish();
      return;
    }
    if (state == kHeader) header();
    else if (state == kStored) stored();
    else if (state == kFinish) finish();
  }

  BITAR_HD void special(uint32_t e) {   // end of block, or a code that does not exist
    if (SUB) return fail(kStatusDataError);
    if ((e & 0xF0u) == 0xE0u) {
      drop(e & 15u);
      if (overrun()) return fail(kStatusTruncated);
      state = last ? kFinish : kHeader;
      return;
    }
    fail(overrun() ? kStatusTruncated : kStatusDataError);
  }

  BITAR_HD uint32_t ll_lookup() {
    uint32_t e = s_ld16(lt_s + ((lo & LMASK) << 1));
    if ((e & 0xF0u) == 0xF0u) e = ll_resolve(e);
    return e;
  }

  // the match whose length code is e (looked up, not yet dropped)
  BITAR_HD void match(uint32_t e) {
    drop(e & 15u);
    refill();
    const uint32_t len = (e >> 8) + 3u + take((e >> 4) & 7u);
    uint32_t d = s_ld16(dt_s + ((lo & DMASK) << 1));
    if ((d & kBadDist) == kBadDist) d = d_resolve(d);
    if ((d & kBadDist) == kBadDist) return fail(overrun() ? kStatusTruncated : kStatusDataError);
    drop(d & 15u);
    refill();
    const uint32_t di = s_ld32(dinfo_s + ((d >> 4) << 2));
    const uint32_t dist = (di & 0xFFFFu) + take(di >> 16);
    if (overrun()) return fail(kStatusTruncated);
    if (dist > produced()) return fail(kStatusDataError);
    if (vpos + len > vcap) return fail(SUB ? kStatusDataError : kStatusOutOfSpace);
    copy(len, dist);
  }

  // ---- one step: up to four literals, then at most one match (or a cold-path action) ----
  BITAR_HD void step() {
    if (state != kDecode) {
      cold();
      return;
    }
    if (SUB && vcap - vpos < 5u) {              // (vpos <= vcap while decoding)
      tail_step();
      return;
    }
    refill();                                   // cnt >= 32
    uint32_t e = ll_lookup();
    if ((e & 0xF0u) == 0) {
      drop(e & 15u);                            // cnt >= 17
      emit(e >> 8);
      e = ll_lookup();
      if ((e & 0xF0u) == 0) {
        drop(e & 15u);                          // cnt >= 2
        emit(e >> 8);
        refill();                               // cnt >= 32
        e = ll_lookup();
        if ((e & 0xF0u) == 0) {
          drop(e & 15u);                        // cnt >= 17
          emit(e >> 8);
          e = ll_lookup();
          if ((e & 0xF0u) == 0) {
            drop(e & 15u);                      // cnt >= 2
            emit(e >> 8);
            e = kNoEntry;
          }
        }
      }
    }
    if ((e & 0x80u) && (e & 0x70u) < 0x60u) {   // length code: the match
      match(e);
      if (state != kDecode) return;
    } else if (e != kNoEntry) {
      special(e);
    }
    if (vpos - vflushed >= 16u) flush();
  }
};

}  // namespace fl
}  // namespace bitar
