// inflate_tok.h -- phase A of the two-phase inflate of indexed chunks: Huffman-decode one 2 KiB sub-range per LANE into
// a TOKEN MAP (no LZ77 copies), and the map format that phase B (inflate_tok_kernel.cuh) resolves in stream order.
//
// Why two phases (DESIGN.md "Inflate"): a match of a sub-range may copy from anywhere in the preceding 32 KiB of its
// block, i.e. from bytes that the lane of ANOTHER sub-range is still producing.  Phase A therefore only does what is
// independent per sub-range -- the serial bit-level work, 32 chains per block -- and leaves behind, per sub-range,
// two things in its SLOT of a scratch area:
//     tokens   16 bits per token, compact, in stream order: the byte of a literal, or D - 1 of a match   (<= 2048)
//     starts   one bit per OUTPUT byte of the sub-range: set where a token (literal or match) starts
// A set bit followed by a set bit is a literal; a set bit followed by a clear bit is the head of a match whose length
// is the distance to the next set bit (matches are at least 3 bytes; a token never straddles a sub-range, so the byte
// after a sub-range always counts as a start).  Phase B is then BYTE-parallel: lane i of a group takes output byte i of
// a 32-byte piece; the number of start bits up to its own says which token of the sub-range its byte belongs to, the
// bit after its own whether that token is a literal: no prefix sums over token lengths, no loop over matches.
//
// BITAR_HD: the same source is compiled for the CPU (tools/model/core_host.cc, tests/test_core_host.py).
//
// Replaces the inflate half of the codec behind rte_compressdev (/root/reference/src/device.cc:240-318, decompress xform
// at src/config.cc:93-105).
#pragma once
#include <stdint.h>
#include <string.h>

#include "inflate_fast.h"   // tables, entries, bit-level helpers, the index

namespace bitar {
namespace tk {

using fl::kStatusDataError;
using fl::kStatusOk;
using fl::kStatusTruncated;
using fl::sptr;

// slot of one sub-range (16-byte aligned parts; vectors are stored whole and phase B reads up to a step past the last
// token without looking at it, hence the slack)
constexpr uint32_t kSlotToks = 0u;                                 // 2048 x u16 + 64 spare
constexpr uint32_t kSlotBits = 4096u + 128u;                       // 64 words of start bits
constexpr uint32_t kSlotBytes = kSlotBits + 256u;                  // 4480: a multiple of 16
constexpr uint32_t kSubsPerTask = 32u;                             // a task = one 64 KiB block
constexpr size_t kTaskBytes = (size_t)kSubsPerTask * kSlotBytes;
static_assert(kSlotBytes % 16u == 0 && kSlotBits % 16u == 0, "slot parts are vector aligned");
constexpr uint32_t kLaneRingBytes = 32u;                           // per lane in shared memory: 16 tokens

BITAR_HD void s_st16(sptr a, uint32_t v) {
#if defined(__CUDA_ARCH__)
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((unsigned short)v));
#else
  *reinterpret_cast<uint16_t*>(const_cast<uint8_t*>(a)) = (uint16_t)v;
#endif
}

// One sub-range of a Huffman-coded block: symbols in, token map out.  The decode tables belong to the group (built once
// per block by the kernel); the lane starts at an indexed bit offset, must produce exactly `olen` bytes worth of
// tokens and must end exactly where the index says the next sub-range starts.
// The lane stages 16 tokens in shared memory; they leave as aligned 16-byte vectors.  A step appends at most 4
// literals and one match.
template <int LBITS, int LT, int DBITS, int DT>
struct TokLane {
  static constexpr uint32_t LMASK = (1u << LBITS) - 1u, DMASK = (1u << DBITS) - 1u;
  enum : uint32_t { kDecode = 1, kFinish = 3, kDone = 4, kSubEnd = 5 };

  sptr lt_s, dt_s, ring_s, dinfo_s;
  const fl::LaneScratch* sc;
  // input: 64-bit bit buffer (lo, hi), refilled one aligned 32-bit word at a time, one word prefetched
  const uint8_t* in;
  const uint32_t* words;
  uint32_t in_len, nwords, wpos, next, skip, start_off;
  uint32_t lo, hi, cnt;
  // output
  uint8_t* slot;                 // 16-byte aligned
  uint32_t tpos, tflushed;       // tokens appended / stored
  uint32_t sbits;                // start bits of the 32 output bytes around opos
  uint32_t opos, olen, before;   // bytes produced, bytes to produce, bytes of the block before this sub-range
  uint32_t state, status;
  uint32_t sub_end_bit, sub_eob;

  BITAR_HD void bind(const uint16_t* lt_, const uint16_t* dt_, uint8_t* ring_, const uint32_t* dinfo, const fl::LaneScratch* scratch) {
    lt_s = fl::sp_of(lt_);
    dt_s = fl::sp_of(dt_);
    ring_s = fl::sp_of(ring_);
    dinfo_s = fl::sp_of(dinfo);
    sc = scratch;
    state = kDone;
    status = kStatusOk;
    in = nullptr;
    words = nullptr;
    slot = nullptr;
    in_len = nwords = wpos = next = skip = start_off = lo = hi = cnt = 0;
    tpos = tflushed = sbits = opos = olen = before = sub_end_bit = sub_eob = 0;
  }

  // Decode `len` bytes worth of tokens from the symbol at stream bit `start_bit`; the sub-range must end at `end_bit`,
  // after an end-of-block symbol when `eob`.  `block_before` = bytes of the block that precede the sub-range (a
  // distance may reach that far back and no farther: phase B never reads outside the block).
  BITAR_HD void start_sub(const uint8_t* src, uint32_t stream_len, uint32_t start_bit, uint32_t end_bit, bool eob, uint8_t* slot_,
                          uint32_t len, uint32_t block_before) {
    in = src;
    in_len = stream_len;
    bits_init(start_bit >> 3);
    drop(start_bit & 7u);
    slot = slot_;
    tpos = tflushed = sbits = opos = 0;
    olen = len;
    before = block_before;
    state = len ? (uint32_t)kDecode : (uint32_t)kSubEnd;
    status = kStatusOk;
    sub_end_bit = end_bit;
    sub_eob = eob ? 1u : 0u;
  }
  BITAR_HD uint32_t tokens() const { return tpos; }

  // ---- bit reader (as fl::FastLane) ----
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
      hi = fl::fsl_hi(next, cnt);
      cnt += 32u;
      next = wpos < nwords ? inf::ld_in32(words + wpos) : 0u;
      wpos++;
    }
  }
  BITAR_HD void drop(uint32_t n) {   // n < 32
    lo = fl::fsr(lo, hi, n);
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

  BITAR_HD void fail(uint32_t st) {
    if (status == kStatusOk) status = st;
    state = kFinish;
  }

  // ---- output ----
  BITAR_HD void st_bits(uint32_t w, uint32_t v) { reinterpret_cast<uint32_t*>(slot + kSlotBits)[w] = v; }
  // a token of n bytes starts at opos: v = the byte of a literal, the distance - 1 of a match
  BITAR_HD void token(uint32_t v, uint32_t n) {
    s_st16(ring_s + ((tpos & 15u) << 1), v);
    tpos++;
    sbits |= 1u << (opos & 31u);
    const uint32_t np = opos + n;
    if ((np ^ opos) >> 5) {                     // the word of start bits is complete (a long match skips whole words)
      uint32_t w = opos >> 5;
      st_bits(w, sbits);
      sbits = 0;
#if defined(__CUDACC__)
#pragma unroll 1
#endif
      for (++w; w < (np >> 5); ++w) st_bits(w, 0u);
    }
    opos = np;
  }
  BITAR_HD void literal(uint32_t byte) { token(byte, 1u); }
  BITAR_HD void flush_vec(sptr from, uint8_t* to) {
    uint32_t w0, w1, w2, w3;
    fl::s_ld128(from, w0, w1, w2, w3);
#if defined(__CUDA_ARCH__)
    *reinterpret_cast<uint4*>(to) = make_uint4(w0, w1, w2, w3);
#else
    uint32_t* o32 = reinterpret_cast<uint32_t*>(to);
    o32[0] = w0; o32[1] = w1; o32[2] = w2; o32[3] = w3;
#endif
  }
  BITAR_HD void flush() {   // a complete vector of 8 tokens (a step adds at most 5)
    if (tpos - tflushed >= 8u) {
      flush_vec(ring_s + ((tflushed & 15u) << 1), slot + kSlotToks + 2u * tflushed);
      tflushed += 8u;
    }
  }
  BITAR_HD void finish() {
    flush();
    if (tpos > tflushed) flush_vec(ring_s + ((tflushed & 15u) << 1), slot + kSlotToks + 2u * tflushed);   // (a whole vector: the slot has the room)
    if (opos & 31u) st_bits(opos >> 5, sbits);
    state = kDone;
  }

  // ---- cold paths ----
  BITAR_HD uint32_t ll_resolve(uint32_t e) {
    const uint32_t sb = e & 15u;
    if (sb) {
      e = fl::s_ld16(lt_s + 2u * ((1u << LBITS) + ((e >> 8) << 2) + ((lo >> LBITS) & ((1u << sb) - 1u))));
      if ((e & 0xF0u) != 0xF0u) return e;
    }
    return fl::canonical_decode(lo, fl::kLitLen, sc->ll_count, sc->ll_first, sc->ll_offs, sc->ll_sorted);
  }
  BITAR_HD uint32_t d_resolve(uint32_t d) {
    const uint32_t sb = d & 15u;
    if (sb) {
      d = fl::s_ld16(dt_s + 2u * ((1u << DBITS) + ((d >> 9) << 2) + ((lo >> DBITS) & ((1u << sb) - 1u))));
      if ((d & fl::kBadDist) != fl::kBadDist) return d;
    }
    return fl::canonical_decode(lo, fl::kDist, sc->d_count, sc->d_first, sc->d_offs, sc->d_sorted);
  }
  BITAR_HD uint32_t ll_lookup() {
    uint32_t e = fl::s_ld16(lt_s + ((lo & LMASK) << 1));
    if ((e & 0xF0u) == 0xF0u) e = ll_resolve(e);
    return e;
  }

  // the sub-range is complete -- it must have ended exactly where the index says
  BITAR_HD void sub_end() {
    state = kFinish;
    if (opos != olen) return fail(kStatusDataError);
    if (sub_eob) {
      refill();
      const uint32_t e = ll_lookup();
      if ((e & 0xF0u) != 0xE0u) return fail(kStatusDataError);
      drop(e & 15u);
    }
    if (8ll * (int64_t)start_off + consumed_bits() != (int64_t)sub_end_bit) fail(kStatusDataError);
  }

  // the match whose length code is e (looked up, not yet dropped)
  BITAR_HD void match(uint32_t e) {
    drop(e & 15u);
    refill();
    const uint32_t len = (e >> 8) + 3u + take((e >> 4) & 7u);
    uint32_t d = fl::s_ld16(dt_s + ((lo & DMASK) << 1));
    if ((d & fl::kBadDist) == fl::kBadDist) d = d_resolve(d);
    if ((d & fl::kBadDist) == fl::kBadDist) return fail(kStatusDataError);
    drop(d & 15u);
    refill();
    const uint32_t di = fl::s_ld32(dinfo_s + ((d >> 4) << 2));
    const uint32_t dist = (di & 0xFFFFu) + take(di >> 16);
    // (no test for running past the input here: the reader yields zeros there, and a lane that used them cannot end on
    // the bit offset the index demands -- sub_end() reports it; an indexed chunk is never truncated, its trailer says so)
    if (dist > before + opos || opos + len > olen) return fail(kStatusDataError);   // outside the block / the sub-range
    token(dist - 1u, len);
  }

  // fewer than 5 bytes left -- one symbol at a time, so that the lane stops exactly at the end
  BITAR_HD void tail_step() {
    if (opos >= olen) {
      state = kSubEnd;
      return;
    }
    refill();
    const uint32_t e = ll_lookup();
    if ((e & 0xF0u) == 0) {
      drop(e & 15u);
      literal(e >> 8);
    } else if ((e & 0x80u) && (e & 0x70u) < 0x60u) {
      match(e);
    } else {
      fail(kStatusDataError);   // end of block (or no such code) inside a sub-range
    }
    flush();
  }

  // ---- one step: up to four literals, then at most one match ----
  BITAR_HD void step() {
    if (state != kDecode) {
      if (state == kSubEnd) sub_end();
      else if (state == kFinish) finish();
      return;
    }
    if (olen - opos < 5u) {
      tail_step();
      return;
    }
    refill();                                   // cnt >= 32
    uint32_t e = ll_lookup();
    if ((e & 0xF0u) == 0) {
      drop(e & 15u);                            // cnt >= 17
      literal(e >> 8);
      e = ll_lookup();
      if ((e & 0xF0u) == 0) {
        drop(e & 15u);                          // cnt >= 2
        literal(e >> 8);
        refill();                               // cnt >= 32
        e = ll_lookup();
        if ((e & 0xF0u) == 0) {
          drop(e & 15u);                        // cnt >= 17
          literal(e >> 8);
          e = ll_lookup();
          if ((e & 0xF0u) == 0) {
            drop(e & 15u);                      // cnt >= 2
            literal(e >> 8);
            e = fl::kNoEntry;
          }
        }
      }
    }
    if ((e & 0x80u) && (e & 0x70u) < 0x60u) {   // length code: the match
      match(e);
    } else if (e != fl::kNoEntry) {
      fail(kStatusDataError);                   // end of block or no such code inside a sub-range
    }
    flush();
  }
};

// Phase B stated serially (host tests, and the definition the kernel's byte-parallel resolver is checked against):
// the `len` output bytes of one sub-range from its slot; base[pos ..] receives them, everything below pos is final.
// Returns false on a map that phase A cannot have produced.
inline bool resolve_sub_serial(const uint8_t* slot, uint8_t* base, uint32_t pos, uint32_t len) {
  const uint32_t* bits = reinterpret_cast<const uint32_t*>(slot + kSlotBits);
  uint32_t ti = 0, dist = 0;
  for (uint32_t i = 0; i < len; ++i) {
    const bool start = (bits[i >> 5] >> (i & 31u)) & 1u;
    const bool next_start = i + 1u == len || ((bits[(i + 1u) >> 5] >> ((i + 1u) & 31u)) & 1u);
    uint16_t t16 = 0;
    if (start) {
      if (ti >= 2048u) return false;
      memcpy(&t16, slot + kSlotToks + 2u * ti++, 2);
    }
    if (start && next_start) {
      if (t16 > 0xFFu) return false;
      base[pos + i] = (uint8_t)t16;
      continue;
    }
    if (start) dist = (uint32_t)t16 + 1u;
    if (dist == 0 || dist > pos + i) return false;
    base[pos + i] = base[pos + i - dist];
  }
  return true;
}

}  // namespace tk
}  // namespace bitar
