// inflate_spec.h -- lane-parallel inflate of DEFLATE streams that carry NO parallel-inflate index (the reference's own
// output: zlib / hardware streams, anything RFC 1951), by SPECULATION on the bit offsets the index would have held.
//
// A Huffman-coded block is cut into ranges of B bits; lane r of a warp starts decoding at bit  first + r * B  WITHOUT
// knowing whether a symbol starts there.  Prefix codes resynchronise: after a few dozen symbols the lane's chain of symbol
// boundaries joins the true chain.  Every lane therefore
//   * records the bit position (and its output / token counts) at the start of its first kRec steps,
//   * decodes to its range's end (the next lane's guess), writing a token map exactly as tk::TokLane does,
//   * then WALKS on, one symbol at a time, until its own position equals a recorded position of the next lane: from
//     there on the next lane's chain is this lane's continuation.  Same bit position + same decoder state (a
//     literal/length symbol is due) = same future: a false match is impossible.
// Lane 0 starts at a known symbol, so validity propagates: lane r is good from record j(r) on when lanes 0 .. r-1 all
// found their successor.  What a good lane decoded BEFORE its join point is garbage and is skipped (token and start-bit
// offsets of the record); what it decoded after the next lane's join point is the next lane's business.  A lane that
// runs out of slot space, finds no join point, meets the end of the block or a bad code simply ENDS the round there;
// the next round (or block) starts at its last position, which is a true symbol boundary.
//
// Phase B (inflate_spec_kernel.cuh) then resolves the ranges of a round in order, byte-parallel, as for indexed chunks:
// the ranges are cut at token boundaries by construction, only their output lengths vary (at most 2048 bytes each).
//
// The decoder never reports an error of its own: whatever it cannot finish cleanly (stored blocks, damaged streams,
// truncated input, output that does not fit, a distance before the start of the output) is DECLINED and decoded by the
// whole-stream kernel (inflate_kernel.cuh), which owns the status words.
//
// BITAR_HD: compiled for the CPU too (tools/model/core_host.cc: host_inflate_spec runs the 32 lanes of a round one
// after the other; tests/test_core_host.py).
//
// Replaces the inflate half of the codec behind rte_compressdev for buffers the reference compressed itself
// (/root/reference/src/memory.cc:432-505 takes any buffer; decompress xform at src/config.cc:93-105).
#pragma once
#include <stdint.h>
#include <string.h>

#include "inflate_tok.h"

namespace bitar {
namespace sp {

constexpr uint32_t kRec = 32u;                                  // recorded step starts per lane
constexpr uint32_t kSlotToks = tk::kSlotToks;                   // 2048 + 64 tokens, as tk::
constexpr uint32_t kSlotBits = tk::kSlotBits;                   // start bits: 72 words (a shifted read looks two words ahead)
constexpr uint32_t kBitsWords = 72u;
constexpr uint32_t kSlotRec = kSlotBits + 4u * kBitsWords;      // kRec x {bit position, output bytes | tokens << 16}
constexpr uint32_t kSlotBytes = kSlotRec + 8u * kRec;           // 4768
static_assert(kSlotBytes % 16u == 0, "slots are vector aligned");
constexpr uint32_t kOutStop = 2048u - 262u;                     // no fast step starts at or past this many output bytes (a step adds <= 4 + 258)
constexpr uint32_t kOutStopWalk = 2048u - 258u;                 // no symbol of the walk either (adds <= 258): a lane never passes 2048
constexpr uint32_t kTokStop = 2048u - 5u;
constexpr uint32_t kMinRangeBits = 1024u;                       // a range shorter than the record window (kRec steps) joins badly
constexpr uint32_t kSeek = 48u;                                 // symbols a lane decodes past its successor's last record, looking for the end of the block
enum : uint32_t { kEndNone = 0, kEndSync = 1, kEndStop = 2, kEndEob = 3, kEndBad = 4, kEndFull = 5 };   // Stop: no join point in the successor's records; Full: the slot has no room for another step

// Bits per range for a round that starts at bit `first`: aim at `target` output bytes per lane, judged by the stream's
// overall ratio (what is left of the output capacity over what is left of the input), then cut what is left of the input
// into a whole number of rounds of 32 equal ranges, so that the last round of a block is as full as the first.
BITAR_HD uint32_t range_bits(uint32_t first, uint32_t in_len, uint32_t produced, uint32_t cap, uint32_t target) {
  const uint32_t in_bits = 8u * in_len, rem_bits = in_bits > first ? in_bits - first : 0u;
  const uint32_t rem_out = cap > produced ? cap - produced : 1u;
  uint64_t b = (uint64_t)target * rem_bits / rem_out;
  if (b < kMinRangeBits) b = kMinRangeBits;
  if (b > 16384u) b = 16384u;
  const uint32_t rounds = (uint32_t)((rem_bits + 32u * b - 1u) / (32u * b));
  const uint32_t even = rem_bits / (32u * (rounds ? rounds : 1u)) + 1u;
  return even < kMinRangeBits ? kMinRangeBits : even;   // (a short rest keeps fewer lanes busy rather than all of them on crumbs)
}

template <int LBITS, int LT, int DBITS, int DT>
struct SpecLane : tk::TokLane<LBITS, LT, DBITS, DT> {
  using Base = tk::TokLane<LBITS, LT, DBITS, DT>;
  using Base::cnt;
  using Base::dinfo_s;
  using Base::dt_s;
  using Base::lo;
  using Base::opos;
  using Base::skip;
  using Base::slot;
  using Base::start_off;
  using Base::state;
  using Base::tpos;
  using Base::wpos;
  enum : uint32_t { kFast = Base::kDecode, kWalk = 6u, kDone = Base::kDone };

  uint32_t goal;                 // the next lane's guess: the walk starts at the first step that begins at or past it
  uint32_t nrec;                 // step starts recorded
  uint32_t end_kind, end_bit, end_opos, end_tpos, sync_j;
  const uint32_t* nx_rec;        // the next lane's records
  uint32_t nx_n, nx_j, nx_pos, seek;

  BITAR_HD uint32_t pos() const { return 8u * start_off + 32u * (wpos - 1u) - skip - cnt; }
  BITAR_HD uint32_t* rec() const { return reinterpret_cast<uint32_t*>(slot + kSlotRec); }

  // decode from bit `start` (a guess, or a known symbol for lane 0) into slot_
  BITAR_HD void start_spec(const uint8_t* src, uint32_t stream_len, uint32_t start, uint32_t goal_, uint8_t* slot_) {
    Base::start_sub(src, stream_len, start, 0u, false, slot_, 0x40000000u, 0u);
    state = kFast;
    goal = goal_;
    nrec = 0;
    end_kind = kEndNone;
    end_bit = end_opos = end_tpos = sync_j = 0;
    nx_rec = nullptr;
    nx_n = nx_j = nx_pos = seek = 0;
  }
  BITAR_HD void idle() {
    state = kDone;
    nrec = 0;
    end_kind = kEndBad;
    end_bit = end_opos = end_tpos = sync_j = 0;
  }
  // after every lane has run its first kRec steps: the successor's records
  // (may_seek: past the successor's last record the lane goes on for up to kSeek symbols -- a successor that ran into
  // the end of the block a few symbols after its start has left few records, and the end is near for this lane too)
  BITAR_HD void set_next(const uint8_t* next_slot, uint32_t n, bool may_seek) {
    nx_rec = reinterpret_cast<const uint32_t*>(next_slot + kSlotRec);
    nx_n = n;
    nx_j = 0;
    nx_pos = n ? ld_rec(0) : 0u;
    seek = may_seek ? kSeek : 0u;
  }
  BITAR_HD uint32_t ld_rec(uint32_t j) const {
#if defined(__CUDA_ARCH__)
    return __ldcg(nx_rec + 2u * j);
#else
    return nx_rec[2u * j];
#endif
  }

  // (the staged tokens and the open word of start bits are stored by the NEXT step -- one copy of that code instead of one
  // per way a lane can end)
  BITAR_HD void end(uint32_t kind, uint32_t p) {
    end_kind = kind;
    end_bit = p;
    end_opos = opos;
    end_tpos = tpos;
    state = Base::kFinish;
  }

  // the match whose length code is e (looked up, not yet dropped); distances are checked by phase B
  BITAR_HD void spec_match(uint32_t e) {
    Base::drop(e & 15u);
    Base::refill();
    const uint32_t len = (e >> 8) + 3u + Base::take((e >> 4) & 7u);
    uint32_t d = fl::s_ld16(dt_s + ((lo & Base::DMASK) << 1));
    if ((d & fl::kBadDist) == fl::kBadDist) d = Base::d_resolve(d);
    if ((d & fl::kBadDist) == fl::kBadDist) return end(kEndBad, pos());
    Base::drop(d & 15u);
    Base::refill();
    const uint32_t di = fl::s_ld32(dinfo_s + ((d >> 4) << 2));
    const uint32_t dist = (di & 0xFFFFu) + Base::take(di >> 16);
    Base::token(dist - 1u, len);
  }
  // neither a literal nor a length: the end of the block, or no such code
  BITAR_HD void special(uint32_t e) {
    if ((e & 0xF0u) == 0xE0u) {
      Base::drop(e & 15u);
      return end(kEndEob, pos());
    }
    end(kEndBad, pos());
  }

  // One step: up to four literals, then at most one match (a lane inside its range); one symbol (a lane that walks).
  // `walk` false: the first kRec steps of a round, during which nobody reads the records yet -- a lane whose range is
  // already complete waits.  One symbol loop serves both modes: the code of a step is what every warp of the kernel
  // keeps fetching, and it should be small.
  BITAR_HD void step(bool walk) {
    if (state == kDone) return;
    if (state == Base::kFinish) return Base::finish();            // state = kDone
    const uint32_t p = pos();
    uint32_t nsym = 4u;
    if (state == kFast) {
      if (p >= goal) {
        state = kWalk;
      } else if (opos >= kOutStop || tpos >= kTokStop) {
        return end(kEndFull, p);
      } else if (nrec < kRec) {
        uint32_t* r = rec();
        r[2u * nrec] = p;
        r[2u * nrec + 1u] = opos | (tpos << 16);
        ++nrec;
      }
    }
    if (state == kWalk) {
      if (!walk) return;
      while (nx_j < nx_n && nx_pos < p) {
        ++nx_j;
        if (nx_j < nx_n) nx_pos = ld_rec(nx_j);
      }
      if (nx_j >= nx_n) {                                         // past everything the successor recorded (or no successor)
        if (seek == 0) return end(kEndStop, p);
        --seek;
      } else if (nx_pos == p) {
        sync_j = nx_j;
        return end(kEndSync, p);
      }
      if (opos >= kOutStopWalk || tpos >= 2048u) return end(kEndFull, p);
      nsym = 1u;
    }
    // (four symbols per trip, two per refill -- a code is at most 15 bits; with nsym = 4 the trip is the whole step, with
    // nsym = 1 it ends after the first symbol: one copy of the code for both modes, no loop overhead in the common one)
    for (;;) {
      Base::refill();                                             // cnt >= 32
      uint32_t e = Base::ll_lookup();
      if ((e & 0xF0u) == 0) {
        Base::drop(e & 15u);                                      // cnt >= 17
        Base::literal(e >> 8);
        if (--nsym == 0) break;
        e = Base::ll_lookup();
        if ((e & 0xF0u) == 0) {
          Base::drop(e & 15u);                                    // cnt >= 2
          Base::literal(e >> 8);
          if (--nsym == 0) break;
          Base::refill();                                         // cnt >= 32
          e = Base::ll_lookup();
          if ((e & 0xF0u) == 0) {
            Base::drop(e & 15u);                                  // cnt >= 17
            Base::literal(e >> 8);
            if (--nsym == 0) break;
            e = Base::ll_lookup();
            if ((e & 0xF0u) == 0) {
              Base::drop(e & 15u);                                // cnt >= 2
              Base::literal(e >> 8);
              if (--nsym == 0) break;
              continue;
            }
          }
        }
      }
      if ((e & 0x80u) && (e & 0x70u) < 0x60u) spec_match(e);
      else special(e);
      break;
    }
    Base::flush();
  }
};

// what a lane that turned out good contributes: its record at the join point tells where its true part starts
struct RangeOut {
  uint32_t len;      // output bytes
  uint32_t tskip;    // tokens of the slot before the join point
  uint32_t bskip;    // output bytes (= start bits) of the slot before the join point
};
template <class Lane>
BITAR_HD RangeOut range_out(const Lane& l, uint32_t j) {
  const uint32_t* r = reinterpret_cast<const uint32_t*>(l.slot + kSlotRec);
#if defined(__CUDA_ARCH__)
  const uint32_t ot = __ldcg(r + 2u * j + 1u);
#else
  const uint32_t ot = r[2u * j + 1u];
#endif
  RangeOut o;
  o.bskip = ot & 0xFFFFu;
  o.tskip = ot >> 16;
  o.len = l.end_opos - o.bskip;
  return o;
}

// Phase B stated serially (host model; the kernel's byte-parallel resolver is checked against it on the GPU): the `len`
// output bytes of one range from its slot, skipping what the lane decoded before its join point.  base[pos ..] receives
// them, everything below pos is final.  False: a distance reaches below the start of the output.
inline bool resolve_range_serial(const uint8_t* slot, uint8_t* base, uint32_t pos, uint32_t len, uint32_t tskip, uint32_t bskip) {
  const uint32_t* bits = reinterpret_cast<const uint32_t*>(slot + kSlotBits);
  uint32_t ti = tskip, dist = 0;
  for (uint32_t i = 0; i < len; ++i) {
    const uint32_t a = bskip + i;
    const bool start = (bits[a >> 5] >> (a & 31u)) & 1u;
    const bool next_start = i + 1u == len || ((bits[(a + 1u) >> 5] >> ((a + 1u) & 31u)) & 1u);
    uint16_t t16 = 0;
    if (start) memcpy(&t16, slot + kSlotToks + 2u * ti++, 2);
    if (start && next_start) {
      base[pos + i] = (uint8_t)t16;
      continue;
    }
    if (start) dist = (uint32_t)t16 + 1u;
    if (dist == 0 || dist > pos + i) return false;
    base[pos + i] = base[pos + i - dist];
  }
  return true;
}

}  // namespace sp
}  // namespace bitar
