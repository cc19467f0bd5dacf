// inflate_tok.h -- phase A of the two-phase inflate of indexed chunks: Huffman-decode one 2 KiB sub-range per LANE into
// TOKEN UNITS (no LZ77 copies), and the unit format that phase B (inflate_resolve_kernel.cuh) resolves in stream order.
//
// Why two phases (DESIGN.md "Inflate"): a match of a sub-range may copy from anywhere in the preceding 32 KiB of its
// block, i.e. from bytes that the lane of ANOTHER sub-range is still producing.  Phase A therefore only does what is
// independent per sub-range -- the serial bit-level work, 32 chains per block -- and leaves behind, per sub-range, a
// sequence of 16-bit units:
//     0x00bb            literal byte bb
//     0x8000 | (L - 3)  match head, length L = 3..258, ALWAYS followed by
//     D - 1             its distance D = 1..32768 (bit 15 clear)
//     0x4000            nothing (padding)
// A head never sits at a unit index == 7 (mod 8): phase B takes the units of a sub-range in groups of 8 (one per lane
// of a group) and finds a head's distance in the next lane without looking into the next group.  The units of a
// sub-range go to its fixed slot of a scratch area as aligned 16-byte vectors; the unit count (padded to a multiple of
// 8) to a side array.
//
// BITAR_HD: the same source is compiled for the CPU (tools/model/core_host.cc, tests/test_core_host.py).
//
// Replaces the inflate half of the codec behind rte_compressdev (/root/reference/src/device.cc:240-318, decompress xform
// at src/config.cc:93-105).
#pragma once
#include <stdint.h>

#include "inflate_fast.h"   // tables, entries, bit-level helpers, the index

namespace bitar {
namespace tk {

using fl::kStatusDataError;
using fl::kStatusOk;
using fl::kStatusTruncated;
using fl::sptr;

constexpr uint32_t kUnitNop = 0x4000u;
constexpr uint32_t kUnitHead = 0x8000u;
// units of one sub-range, worst case: 2048 literals, one padding unit per 7 units, rounded up to whole vectors
constexpr uint32_t kSlotUnits = 2352u;
constexpr uint32_t kSlotBytes = kSlotUnits * 2u;                   // 4704: a multiple of 16
constexpr uint32_t kSubsPerTask = 32u;                             // a task = one 64 KiB block
constexpr size_t kTaskBytes = (size_t)kSubsPerTask * kSlotBytes;   // 150 528
static_assert(kSlotUnits % 8u == 0 && kSlotUnits >= (dfl::kSub * 8u + 6u) / 7u + 8u, "slot holds the worst case");

BITAR_HD void s_st16(sptr a, uint32_t v) {
#if defined(__CUDA_ARCH__)
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((unsigned short)v));
#else
  *reinterpret_cast<uint16_t*>(const_cast<uint8_t*>(a)) = (uint16_t)v;
#endif
}

// One sub-range of a Huffman-coded block: symbols in, units out.  The decode tables belong to the group (built once
// per block by the kernel); the lane starts at an indexed bit offset, must produce exactly `olen` bytes worth of
// tokens and must end exactly where the index says the next sub-range starts.
// URING: units of the lane's staging ring in shared memory (a step appends at most 7: 4 literals, a pad, a match).
template <int LBITS, int LT, int DBITS, int DT, int URING = 16>
struct TokLane {
  static_assert(URING >= 16 && (URING & (URING - 1)) == 0, "ring: power of two, two vectors at least");
  static constexpr uint32_t UM = URING - 1;
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
  uint32_t upos, uflushed;       // units appended / stored
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
    upos = uflushed = opos = olen = before = sub_end_bit = sub_eob = 0;
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
    upos = uflushed = opos = 0;
    olen = len;
    before = block_before;
    state = len ? (uint32_t)kDecode : (uint32_t)kSubEnd;
    status = kStatusOk;
    sub_end_bit = end_bit;
    sub_eob = eob ? 1u : 0u;
  }
  BITAR_HD uint32_t units() const { return upos; }   // after kDone: a multiple of 8

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
  BITAR_HD void unit(uint32_t u) {
    s_st16(ring_s + ((upos & UM) << 1), u);
    upos++;
  }
  BITAR_HD void flush() {   // every complete vector of 8 units
    while (upos - uflushed >= 8u) {
      uint32_t w0, w1, w2, w3;
      fl::s_ld128(ring_s + ((uflushed & UM) << 1), w0, w1, w2, w3);
#if defined(__CUDA_ARCH__)
      *reinterpret_cast<uint4*>(slot + 2u * uflushed) = make_uint4(w0, w1, w2, w3);
#else
      uint32_t* o32 = reinterpret_cast<uint32_t*>(slot + 2u * uflushed);
      o32[0] = w0; o32[1] = w1; o32[2] = w2; o32[3] = w3;
#endif
      uflushed += 8u;
    }
  }
  BITAR_HD void finish() {
    while (upos & 7u) unit(kUnitNop);
    flush();
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
    if ((d & fl::kBadDist) == fl::kBadDist) return fail(overrun() ? kStatusTruncated : kStatusDataError);
    drop(d & 15u);
    refill();
    const uint32_t di = fl::s_ld32(dinfo_s + ((d >> 4) << 2));
    const uint32_t dist = (di & 0xFFFFu) + take(di >> 16);
    if (overrun()) return fail(kStatusTruncated);
    if (dist > before + opos || opos + len > olen) return fail(kStatusDataError);   // outside the block / the sub-range
    if ((upos & 7u) == 7u) unit(kUnitNop);
    unit(kUnitHead | (len - 3u));
    unit(dist - 1u);
    opos += len;
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
      unit(e >> 8);
      opos++;
    } else if ((e & 0x80u) && (e & 0x70u) < 0x60u) {
      match(e);
    } else {
      fail(kStatusDataError);   // end of block (or no such code) inside a sub-range
    }
    if (upos - uflushed >= 8u) flush();
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
      unit(e >> 8);
      opos++;
      e = ll_lookup();
      if ((e & 0xF0u) == 0) {
        drop(e & 15u);                          // cnt >= 2
        unit(e >> 8);
        opos++;
        refill();                               // cnt >= 32
        e = ll_lookup();
        if ((e & 0xF0u) == 0) {
          drop(e & 15u);                        // cnt >= 17
          unit(e >> 8);
          opos++;
          e = ll_lookup();
          if ((e & 0xF0u) == 0) {
            drop(e & 15u);                      // cnt >= 2
            unit(e >> 8);
            opos++;
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
    if (upos - uflushed >= 8u) flush();
  }
};

// ---- phase B: one LANE resolves one block ------------------------------------------------------------------------
// The units of the block's sub-ranges in stream order: literal bytes and LZ77 copies through a short ring in shared
// memory that leaves as aligned 16-byte vector stores; matches farther back than the ring read the lane's own earlier
// output (L1/L2).  The copy chain of a block is serial, so the parallelism of this phase is across blocks: 32 per
// warp, every lane a state machine of its own (step() = one token).
template <int RING>
struct ResolveLane {
  static_assert(RING >= 128 && (RING & (RING - 1)) == 0, "ring: power of two >= 128");
  static constexpr uint32_t RM = RING - 1;
  static constexpr uint32_t kPiece = RING / 2;   // bytes copied between two flushes (<= RING - 16 - 19)
  enum : uint32_t { kIdle = 0, kRun = 1, kBad = 2 };

  sptr ring_s;
  const uint16_t* up;        // next unit of the current sub-range
  uint32_t urem;             // units left in it
  const uint8_t* slots;      // the block's unit slots (kSlotBytes each)
  const uint16_t* cnts;      // units per slot
  uint32_t s, ns;            // next sub-range, sub-ranges of the block
  // output: "virtual" positions v = offset + (dst & 15), so that v % 16 == address % 16
  uint8_t* vbase;
  uint32_t vstart, vpos, vflushed, vcap;
  uint32_t state;

  BITAR_HD void bind(uint8_t* ring_) {
    ring_s = fl::sp_of(ring_);
    up = nullptr;
    slots = nullptr;
    cnts = nullptr;
    vbase = nullptr;
    urem = s = ns = vstart = vpos = vflushed = vcap = 0;
    state = kIdle;
  }
  BITAR_HD void start_block(uint8_t* dst, uint32_t len, const uint8_t* slots_, const uint16_t* cnts_, uint32_t n_subs) {
    const uint32_t mis = (uint32_t)((uintptr_t)dst & 15u);
    vbase = dst - mis;
    vstart = vpos = vflushed = mis;
    vcap = mis + len;
    slots = slots_;
    cnts = cnts_;
    s = 0;
    ns = n_subs;
    urem = 0;
    state = kRun;
  }
  BITAR_HD static uint32_t ld_unit(const uint16_t* p) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)__ldg(p);
#else
    return *p;
#endif
  }
  BITAR_HD void emit(uint32_t byte) {
    fl::s_st8(ring_s + (vpos & RM), byte);
    vpos++;
  }
  // Store every complete 16-byte vector below vpos (and the unaligned head of the block).
  BITAR_HD void flush() {
    while (vpos - vflushed >= 16u || ((vflushed & 15u) && vpos >= ((vflushed + 15u) & ~15u))) {
      if (vflushed & 15u) {
        const uint32_t a = (vflushed + 15u) & ~15u;
        for (uint32_t v = vflushed; v < a; ++v) vbase[v] = (uint8_t)fl::s_ld8(ring_s + (v & RM));
        vflushed = a;
        continue;
      }
      uint32_t w0, w1, w2, w3;
      fl::s_ld128(ring_s + (vflushed & RM), w0, w1, w2, w3);
#if defined(__CUDA_ARCH__)
      *reinterpret_cast<uint4*>(vbase + vflushed) = make_uint4(w0, w1, w2, w3);
#else
      uint32_t* o32 = reinterpret_cast<uint32_t*>(vbase + vflushed);
      o32[0] = w0; o32[1] = w1; o32[2] = w2; o32[3] = w3;
#endif
      vflushed += 16u;
    }
  }
  BITAR_HD void finish() {
    flush();
    for (uint32_t v = vflushed; v < vpos; ++v) vbase[v] = (uint8_t)fl::s_ld8(ring_s + (v & RM));
    vflushed = vpos;
  }
  // LZ77 copy (1 <= dist <= vpos - vstart, vpos + len <= vcap), in pieces that fit the ring
  BITAR_HD void copy(uint32_t len, uint32_t dist) {
    for (;;) {
      const uint32_t piece = len < kPiece ? len : kPiece;
      const uint32_t src = vpos - dist;
      uint32_t j = 0;
      if (dist < (uint32_t)RING) {           // the source is still in the ring
        if (dist >= 4u) {
          for (; j + 4u <= piece; j += 4u) {
            const uint32_t b0 = fl::s_ld8(ring_s + ((src + j) & RM)), b1 = fl::s_ld8(ring_s + ((src + j + 1u) & RM));
            const uint32_t b2 = fl::s_ld8(ring_s + ((src + j + 2u) & RM)), b3 = fl::s_ld8(ring_s + ((src + j + 3u) & RM));
            fl::s_st8(ring_s + ((vpos + j) & RM), b0);
            fl::s_st8(ring_s + ((vpos + j + 1u) & RM), b1);
            fl::s_st8(ring_s + ((vpos + j + 2u) & RM), b2);
            fl::s_st8(ring_s + ((vpos + j + 3u) & RM), b3);
          }
        }
        for (; j < piece; ++j) fl::s_st8(ring_s + ((vpos + j) & RM), fl::s_ld8(ring_s + ((src + j) & RM)));
      } else {                               // flushed long ago: dist >= RING, so src + piece <= vflushed
        const uint8_t* g = vbase + src;
        for (; j + 4u <= piece; j += 4u) {
          const uint32_t b0 = g[j], b1 = g[j + 1u], b2 = g[j + 2u], b3 = g[j + 3u];
          fl::s_st8(ring_s + ((vpos + j) & RM), b0);
          fl::s_st8(ring_s + ((vpos + j + 1u) & RM), b1);
          fl::s_st8(ring_s + ((vpos + j + 2u) & RM), b2);
          fl::s_st8(ring_s + ((vpos + j + 3u) & RM), b3);
        }
        for (; j < piece; ++j) fl::s_st8(ring_s + ((vpos + j) & RM), g[j]);
      }
      vpos += piece;
      len -= piece;
      if (len == 0) return;
      flush();
    }
  }
  // the next sub-range's units, or the end of the block
  BITAR_HD void next_slot() {
    if (s < ns) {
      up = reinterpret_cast<const uint16_t*>(slots + (size_t)s * kSlotBytes);
      urem = cnts[s];
      if (urem > kSlotUnits) state = kBad;
      ++s;
      return;
    }
    if (vpos != vcap) {
      state = kBad;
      return;
    }
    finish();
    state = kIdle;
  }
  // one token
  BITAR_HD void step() {
    if (state != kRun) return;
    if (urem == 0) {
      next_slot();
      return;
    }
    const uint32_t u = ld_unit(up);
    ++up;
    --urem;
    if (u < 0x100u) {
      if (vpos >= vcap) {
        state = kBad;
        return;
      }
      emit(u);
    } else if (u & kUnitHead) {
      const uint32_t len = (u & 0xFFu) + 3u;
      if (urem == 0) {
        state = kBad;
        return;
      }
      const uint32_t dist = ld_unit(up) + 1u;
      ++up;
      --urem;
      if (dist > vpos - vstart || vpos + len > vcap) {   // (phase A checked both against the block)
        state = kBad;
        return;
      }
      copy(len, dist);
    }
    if (vpos - vflushed >= 16u) flush();
  }
};

// Phase B stated serially (host tests, and the definition the kernel is checked against): resolve the units of one
// sub-range into out[pos ..]; `base` = first byte of the block.  Returns the new position, or 0xFFFFFFFF on a unit
// sequence that phase A cannot have produced.
BITAR_HD uint32_t resolve_units_serial(const uint16_t* units, uint32_t n_units, uint8_t* base, uint32_t pos, uint32_t limit) {
  for (uint32_t i = 0; i < n_units; ++i) {
    const uint32_t u = units[i];
    if (u == kUnitNop) continue;
    if (u < 0x100u) {
      if (pos >= limit) return 0xFFFFFFFFu;
      base[pos++] = (uint8_t)u;
    } else if (u & kUnitHead) {
      const uint32_t len = (u & 0xFFu) + 3u;
      if (i + 1 >= n_units || (i & 7u) == 7u) return 0xFFFFFFFFu;
      const uint32_t dist = (uint32_t)units[++i] + 1u;
      if (dist > pos || pos + len > limit || dist > 32768u) return 0xFFFFFFFFu;
      for (uint32_t k = 0; k < len; ++k, ++pos) base[pos] = base[pos - dist];
    } else {
      return 0xFFFFFFFFu;
    }
  }
  return pos;
}

}  // namespace tk
}  // namespace bitar
