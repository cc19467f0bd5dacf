// inflate_core.h -- raw DEFLATE (RFC 1951) decoder core for the sm_100a inflate kernel.
//
// One *group* of G lanes (G = 1,2,4,8,16,32; a sub-warp) decodes one chunk.  Every lane of a group
// runs the Huffman decode redundantly on register-resident bit-buffer state (no shuffles on the
// critical path, table look-ups are shared-memory broadcasts), and the G lanes cooperate on the
// byte moves: back-reference copies, stored blocks and 16-byte vector flushes to HBM.
//
// Recently produced output lives in a per-group shared-memory ring; back-references that fit the
// ring never touch HBM, farther ones read the already flushed output (L1/L2 hits).  The ring is
// flushed to global memory in aligned 16-byte vectors.
//
// The file is BITAR_HD code: compiled by nvcc for the kernel and by g++ with G = 1 for the CPU unit
// tests (tests/test_core_host.py through tools/model/core_host.cc) -- same source, so table
// construction, header parsing and all error paths are exercised without a GPU.
//
// Replaces the inflate half of the codec the reference reaches through rte_compressdev
// (/root/reference/src/device.cc:240-318, decompress xform at src/config.cc:93-105); per-op status
// mirrors rte_comp_op_status as consumed at src/device.cc:512-520.
#pragma once
#include <stdint.h>

#include "deflate_common.h"

namespace bitar {
namespace inf {

// per-chunk status words (shared with include/bitar_cuda.h: BITAR_OP_*)
enum : uint32_t {
  kStatusOk = 0,
  kStatusOutOfSpace = 1,   // output capacity exceeded (RTE_COMP_OP_STATUS_OUT_OF_SPACE_TERMINATED)
  kStatusDataError = 2,    // invalid DEFLATE stream (RTE_COMP_OP_STATUS_ERROR)
  kStatusTruncated = 3,    // input ended before the final block completed
  kStatusNotRun = 0xFFFFFFFFu
};

// decode-table entry:  [31:16] value | [10:9] kind | [8:5] extra bits | [4:0] code length
enum : uint32_t { kKindLiteral = 0, kKindBase = 1, kKindEob = 2, kKindSlow = 3 };
BITAR_HD uint32_t make_entry(uint32_t value, uint32_t kind, uint32_t extra, uint32_t nbits) {
  return (value << 16) | (kind << 9) | (extra << 5) | nbits;
}
BITAR_HD uint32_t e_nbits(uint32_t e) { return e & 31u; }
BITAR_HD uint32_t e_extra(uint32_t e) { return (e >> 5) & 15u; }
BITAR_HD uint32_t e_kind(uint32_t e) { return (e >> 9) & 3u; }
BITAR_HD uint32_t e_value(uint32_t e) { return e >> 16; }

struct alignas(16) Vec16 {
  uint32_t x, y, z, w;
};

template <int LBITS, int DBITS, int RING>
struct alignas(16) GroupSmem {
  uint8_t ring[RING];          // recent output, indexed by (virtual position & (RING-1))
  uint32_t lt[1 << LBITS];     // litlen primary table
  uint32_t dt[1 << DBITS];     // distance primary table (also hosts the code-length code table)
  uint16_t ll_sorted[288];     // symbols sorted by (code length, symbol): canonical order
  uint16_t d_sorted[32];
  uint16_t ll_count[16], ll_first[16], ll_offs[16];
  uint16_t d_count[16], d_first[16], d_offs[16];
  uint8_t lens[320];           // code lengths of the block being set up
};

// ---- group abstraction ---------------------------------------------------------------------------
template <int G>
struct Group {
  int lane;        // 0..G-1
  unsigned mask;   // lanes of this group inside the warp
  BITAR_HD void sync() const {
#if defined(__CUDA_ARCH__)
    __syncwarp(mask);
#endif
  }
};

BITAR_HD uint32_t ld_in32(const uint32_t* p) {
#if defined(__CUDA_ARCH__)
  return __ldg(p);
#else
  return *p;
#endif
}
BITAR_HD uint8_t ld_in8(const uint8_t* p) {
#if defined(__CUDA_ARCH__)
  return __ldg(p);
#else
  return *p;
#endif
}
BITAR_HD uint32_t brev32(uint32_t v) {
#if defined(__CUDA_ARCH__)
  return __brev(v);
#else
  v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
  v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
  v = ((v >> 4) & 0x0F0F0F0Fu) | ((v & 0x0F0F0F0Fu) << 4);
  v = ((v >> 8) & 0x00FF00FFu) | ((v & 0x00FF00FFu) << 8);
  return (v >> 16) | (v << 16);
#endif
}

// ---- bit reader: 64-bit buffer, refilled one aligned 32-bit word at a time, one word prefetched ---
struct BitReader {
  const uint32_t* words;
  uint32_t nwords, wpos, next, start_off;
  uint64_t buf;
  int cnt;
  int64_t loaded;   // bits ever placed in buf (excludes the alignment skip)
  int64_t avail;    // bits really present in the input from the start offset

  BITAR_HD void init(const uint8_t* base, uint32_t len, uint32_t off) {
    const uint8_t* a = base + off;
    start_off = off;
    uint32_t mis = (uint32_t)((uintptr_t)a & 3u);
    words = (const uint32_t*)(a - mis);
    uint32_t bytes = len - off;
    avail = (int64_t)bytes * 8;
    nwords = bytes ? (mis + bytes + 3u) >> 2 : 0u;
    uint32_t first = nwords ? ld_in32(words) : 0u;
    buf = (uint64_t)(first >> (8u * mis));
    cnt = 32 - 8 * (int)mis;
    loaded = cnt;
    next = nwords > 1 ? ld_in32(words + 1) : 0u;
    wpos = 2;
  }
  BITAR_HD void refill() {  // afterwards cnt is in [32, 63]
    if (cnt < 32) {
      buf |= (uint64_t)next << cnt;
      cnt += 32;
      loaded += 32;
      next = wpos < nwords ? ld_in32(words + wpos) : 0u;
      wpos++;
    }
  }
  BITAR_HD uint32_t peek(int n) const { return (uint32_t)buf & ((1u << n) - 1u); }
  BITAR_HD void drop(int n) {
    buf >>= n;
    cnt -= n;
  }
  BITAR_HD uint32_t take(int n) {
    uint32_t v = peek(n);
    drop(n);
    return v;
  }
  BITAR_HD int64_t consumed() const { return loaded - cnt; }
  BITAR_HD bool overrun() const { return consumed() > avail; }
};

// ---- decode table construction --------------------------------------------------------------------
enum TableKind { kTableLitLen = 0, kTableDist = 1, kTableCodeLen = 2 };

BITAR_HD uint32_t entry_for(int kind, int sym, int nbits) {
  if (kind == kTableLitLen) {
    if (sym < 256) return make_entry((uint32_t)sym, kKindLiteral, 0, (uint32_t)nbits);
    if (sym == 256) return make_entry(0, kKindEob, 0, (uint32_t)nbits);
    if (sym < 286)
      return make_entry((uint32_t)dfl::len_base(sym - 257), kKindBase,
                        (uint32_t)dfl::len_extra_bits(sym - 257), (uint32_t)nbits);
    return make_entry(0, kKindSlow, 0, 0);  // 286/287 never valid: slow path reports the error
  }
  if (kind == kTableDist) {
    if (sym < 30)
      return make_entry((uint32_t)dfl::dist_base(sym), kKindBase, (uint32_t)dfl::dist_extra_bits(sym),
                        (uint32_t)nbits);
    return make_entry(0, kKindSlow, 0, 0);
  }
  return make_entry((uint32_t)sym, kKindLiteral, 0, (uint32_t)nbits);
}

// Builds count/first/offs/sorted and the primary table for one alphabet.  Returns a status word.
// Validity rules follow zlib's inflate_table: over-subscribed sets are errors; incomplete sets are
// errors except a litlen/dist alphabet holding a single 1-bit code; an all-zero alphabet is accepted
// (any use of it is then an error).
template <int G>
BITAR_HD_NOINLINE uint32_t build_table(const uint8_t* lens, int n, int kind, uint32_t* table, int tbits,
                                       uint16_t* count, uint16_t* first, uint16_t* offs,
                                       uint16_t* sorted, const Group<G>& g) {
  g.sync();
  uint32_t status = kStatusOk;
  if (g.lane == 0) {
    for (int b = 0; b < 16; ++b) count[b] = 0;
    for (int i = 0; i < n; ++i) count[lens[i]]++;
    int left = 1, maxl = 0;
    for (int b = 1; b <= 15; ++b) {
      left = (left << 1) - (int)count[b];
      if (count[b]) maxl = b;
      if (left < 0) break;
    }
    int used = n - (int)count[0];
    if (left < 0) status = kStatusDataError;
    else if (left > 0 && used > 0 && (kind == kTableCodeLen || maxl != 1)) status = kStatusDataError;
    uint32_t f = 0, o = 0;
    first[0] = 0;
    offs[0] = 0;
    for (int b = 1; b <= 15; ++b) {
      first[b] = (uint16_t)f;
      offs[b] = (uint16_t)o;
      f = (f + count[b]) << 1;
      o += count[b];
    }
    if (status == kStatusOk) {
      uint16_t at[16];
      for (int b = 0; b < 16; ++b) at[b] = offs[b];
      for (int i = 0; i < n; ++i)
        if (lens[i]) sorted[at[lens[i]]++] = (uint16_t)i;
    }
    count[0] = (uint16_t)status;  // broadcast slot (count[0] is never used by the decoder)
  }
  g.sync();
  status = count[0];
  if (status != kStatusOk) return status;
  const uint32_t slow = make_entry(0, kKindSlow, 0, 0);
  for (int j = g.lane; j < (1 << tbits); j += G) table[j] = slow;
  g.sync();
  int used = (int)offs[15] + (int)count[15];
  for (int idx = g.lane; idx < used; idx += G) {
    int sym = sorted[idx];
    int l = lens[sym];
    if (l > tbits) continue;
    uint32_t code = (uint32_t)first[l] + (uint32_t)(idx - (int)offs[l]);
    uint32_t r = brev32(code) >> (32 - l);
    uint32_t e = entry_for(kind, sym, l);
    for (uint32_t k = r; k < (1u << tbits); k += (1u << l)) table[k] = e;
  }
  g.sync();
  return kStatusOk;
}

// Canonical bit-by-bit decode for codes longer than the primary table (and for invalid prefixes).
// Returns an entry with the full code length, or kind==kKindSlow/nbits==0 on an invalid code.
BITAR_HD_NOINLINE uint32_t slow_decode(uint64_t buf, int kind, const uint16_t* count,
                                       const uint16_t* first, const uint16_t* offs,
                                       const uint16_t* sorted) {
  uint32_t code = 0;
  for (int l = 1; l <= 15; ++l) {
    code = (code << 1) | (uint32_t)((buf >> (l - 1)) & 1u);
    uint32_t rel = code - (uint32_t)first[l];
    if (code >= first[l] && rel < count[l]) return entry_for(kind, sorted[offs[l] + rel], l);
  }
  return make_entry(0, kKindSlow, 0, 0);
}

// ---- output window ---------------------------------------------------------------------------------
// Positions are "virtual": v = (output offset) + (dst address & 15), so that v % 16 == address % 16
// and ring slots line up with 16-byte vectors of the destination.
template <int G, int RING>
struct Output {
  uint8_t* vbase;      // dst - mis (16-byte aligned)
  uint8_t* ring;
  uint32_t vstart;     // mis
  uint32_t vpos;       // next virtual position to write
  uint32_t vflushed;   // everything below is in global memory
  uint32_t vcap;       // mis + capacity

  static constexpr uint32_t RM = RING - 1;
  static constexpr uint32_t kFlushAt = 256;
  static_assert(RING >= 1024 && (RING & (RING - 1)) == 0, "ring must be a power of two >= 1024");

  BITAR_HD void init(uint8_t* dst, uint32_t cap, uint8_t* ring_) {
    uint32_t mis = (uint32_t)((uintptr_t)dst & 15u);
    vbase = dst - mis;
    ring = ring_;
    vstart = vpos = vflushed = mis;
    vcap = mis + cap;
  }
  BITAR_HD uint32_t produced() const { return vpos - vstart; }

  BITAR_HD void store16(uint32_t v) {
#if defined(__CUDA_ARCH__)
    *reinterpret_cast<uint4*>(vbase + v) = *reinterpret_cast<const uint4*>(ring + (v & RM));
#else
    *reinterpret_cast<Vec16*>(vbase + v) = *reinterpret_cast<const Vec16*>(ring + (v & RM));
#endif
  }

  // Flush whole 16-byte vectors below vpos (bytes below vstart are never touched).
  BITAR_HD void flush(const Group<G>& g, bool final) {
    g.sync();
    uint32_t lo = vflushed, hi = final ? vpos : (vpos & ~15u);
    if (hi <= lo) return;
    uint32_t a = (lo + 15u) & ~15u;   // first vector boundary at/after lo
    uint32_t b = hi & ~15u;           // last vector boundary at/below hi
    if (a > b) {                      // lo and hi inside one vector
      for (uint32_t v = lo + (uint32_t)g.lane; v < hi; v += G) vbase[v] = ring[v & RM];
    } else {
      for (uint32_t v = lo + (uint32_t)g.lane; v < a; v += G) vbase[v] = ring[v & RM];
      for (uint32_t v = a + 16u * (uint32_t)g.lane; v < b; v += 16u * G) store16(v);
      for (uint32_t v = b + (uint32_t)g.lane; v < hi; v += G) vbase[v] = ring[v & RM];
    }
    vflushed = hi;
    g.sync();
  }
  BITAR_HD void maybe_flush(const Group<G>& g) {
    if (vpos - vflushed >= kFlushAt) flush(g, false);
  }

  BITAR_HD void literal(uint32_t byte, const Group<G>& g) {
    if (g.lane == 0) ring[vpos & RM] = (uint8_t)byte;
    vpos++;
  }

  // LZ77 copy of len bytes from distance dist (1 <= dist <= produced(), vpos + len <= vcap).
  BITAR_HD void copy(uint32_t len, uint32_t dist, const Group<G>& g) {
    g.sync();
    uint32_t src = vpos - dist;
    if (dist + len <= (uint32_t)RING) {
      if (dist >= (uint32_t)G) {
        const bool overlap = dist < len;
        for (uint32_t base = 0; base < len; base += G) {
          uint32_t i = base + (uint32_t)g.lane;
          if (i < len) ring[(vpos + i) & RM] = ring[(src + i) & RM];
          if (overlap) g.sync();
        }
      } else {  // period shorter than the group: read only the original pattern
        for (uint32_t i = (uint32_t)g.lane; i < len; i += G) ring[(vpos + i) & RM] = ring[(src + i % dist) & RM];
      }
    } else {  // far reference: dist > RING - len >= len, the source is disjoint from the destination
      if (src + len > vflushed) flush(g, false);
      for (uint32_t i = (uint32_t)g.lane; i < len; i += G) {
#if defined(__CUDA_ARCH__)
        uint8_t b = *reinterpret_cast<volatile const uint8_t*>(vbase + src + i);
#else
        uint8_t b = vbase[src + i];
#endif
        ring[(vpos + i) & RM] = b;
      }
    }
    vpos += len;
  }

  // stored block payload: n bytes from the (read-only) input
  BITAR_HD void copy_in(const uint8_t* in, uint32_t n, const Group<G>& g) {
    g.sync();
    uint32_t done = 0;
    while (done < n) {
      uint32_t part = n - done < kFlushAt ? n - done : kFlushAt;
      for (uint32_t i = (uint32_t)g.lane; i < part; i += G) ring[(vpos + i) & RM] = ld_in8(in + done + i);
      vpos += part;
      done += part;
      flush(g, false);
    }
  }
};

struct ChunkResult {
  uint32_t produced;
  uint32_t status;
  uint32_t consumed;   // input bytes used
  uint32_t blocks;
};

// ---- the decoder -------------------------------------------------------------------------------------
template <int G, int LBITS, int DBITS, int RING>
BITAR_HD ChunkResult inflate_chunk(const uint8_t* in, uint32_t in_len, uint8_t* out, uint32_t cap,
                                            GroupSmem<LBITS, DBITS, RING>* sm, const Group<G>& g) {
  static_assert(DBITS >= 7, "distance table also hosts the 7-bit code-length code");
  ChunkResult res{0, kStatusOk, 0, 0};
  BitReader br;
  br.init(in, in_len, 0);
  Output<G, RING> o;
  o.init(out, cap, sm->ring);
  uint32_t status = kStatusOk;
  uint32_t last = 0;
  constexpr uint32_t LMASK = (1u << LBITS) - 1u, DMASK = (1u << DBITS) - 1u;

  while (!last && status == kStatusOk) {
    br.refill();
    last = br.take(1);
    uint32_t type = br.take(2);
    res.blocks++;
    if (type == 0) {
      br.drop(br.cnt & 7);  // to the next byte boundary (cnt and position share parity mod 8)
      br.refill();
      uint32_t len = br.take(16);
      uint32_t nlen = br.take(16);
      if (br.overrun()) { status = kStatusTruncated; break; }
      if ((len ^ 0xFFFFu) != nlen) { status = kStatusDataError; break; }
      uint32_t at = br.start_off + (uint32_t)(br.consumed() >> 3);
      if ((uint64_t)at + len > in_len) { status = kStatusTruncated; break; }
      if (o.vpos + len > o.vcap) { status = kStatusOutOfSpace; break; }
      o.copy_in(in + at, len, g);
      br.init(in, in_len, at + len);
      continue;
    }
    if (type == 3) { status = kStatusDataError; break; }

    int nlen, ndist;
    if (type == 1) {
      for (int i = g.lane; i < 288; i += G) sm->lens[i] = (uint8_t)dfl::fixed_ll_len(i);
      for (int i = g.lane; i < 32; i += G) sm->lens[288 + i] = 5;
      nlen = 288;
      ndist = 32;
    } else {
      nlen = (int)br.take(5) + 257;
      ndist = (int)br.take(5) + 1;
      int ncode = (int)br.take(4) + 4;
      if (nlen > 286 || ndist > 30) { status = kStatusDataError; break; }
      g.sync();
      if (g.lane == 0)
        for (int i = 0; i < 19; ++i) sm->lens[i] = 0;
      g.sync();
      for (int i = 0; i < ncode; ++i) {
        br.refill();
        uint32_t v = br.take(3);
        if (g.lane == 0) sm->lens[dfl::cl_order(i)] = (uint8_t)v;
      }
      if (br.overrun()) { status = kStatusTruncated; break; }
      status = build_table<G>(sm->lens, 19, kTableCodeLen, sm->dt, 7, sm->d_count, sm->d_first,
                              sm->d_offs, sm->d_sorted, g);
      if (status != kStatusOk) break;
      // code lengths of both alphabets, decoded redundantly by every lane; lane 0 stores them
      int idx = 0, prev = 0;
      const int total = nlen + ndist;
      while (idx < total) {
        br.refill();
        uint32_t e = sm->dt[br.peek(7)];
        if (e_nbits(e) == 0) { status = kStatusDataError; break; }
        br.drop((int)e_nbits(e));
        int sym = (int)e_value(e);
        int rep, val;
        if (sym < 16) { rep = 1; val = sym; prev = sym; }
        else if (sym == 16) {
          if (idx == 0) { status = kStatusDataError; break; }
          rep = 3 + (int)br.take(2); val = prev;
        } else if (sym == 17) { rep = 3 + (int)br.take(3); val = 0; prev = 0; }
        else { rep = 11 + (int)br.take(7); val = 0; prev = 0; }
        if (idx + rep > total) { status = kStatusDataError; break; }
        if (g.lane == 0)
          for (int k = 0; k < rep; ++k) sm->lens[idx + k] = (uint8_t)val;
        idx += rep;
      }
      if (status != kStatusOk) break;
      if (br.overrun()) { status = kStatusTruncated; break; }
      g.sync();
      if (sm->lens[256] == 0) { status = kStatusDataError; break; }
    }
    // distance lengths are moved out of the way first: lens[nlen..] -> built before lt overwrites nothing
    status = build_table<G>(sm->lens + nlen, ndist, kTableDist, sm->dt, DBITS, sm->d_count, sm->d_first,
                            sm->d_offs, sm->d_sorted, g);
    if (status != kStatusOk) break;
    status = build_table<G>(sm->lens, nlen, kTableLitLen, sm->lt, LBITS, sm->ll_count, sm->ll_first,
                            sm->ll_offs, sm->ll_sorted, g);
    if (status != kStatusOk) break;

    // ---- symbol loop ----
    for (;;) {
      br.refill();
      uint32_t e = sm->lt[(uint32_t)br.buf & LMASK];
      if (e_kind(e) == kKindSlow) {
        e = slow_decode(br.buf, kTableLitLen, sm->ll_count, sm->ll_first, sm->ll_offs, sm->ll_sorted);
        if (e_nbits(e) == 0) { status = kStatusDataError; break; }
      }
      br.drop((int)e_nbits(e));
      uint32_t kind = e_kind(e);
      if (kind == kKindLiteral) {
        if (o.vpos >= o.vcap) { status = br.overrun() ? kStatusTruncated : kStatusOutOfSpace; break; }
        o.literal(e_value(e), g);
        o.maybe_flush(g);
        continue;
      }
      if (kind == kKindEob) break;
      uint32_t len = e_value(e) + br.take((int)e_extra(e));
      br.refill();
      uint32_t d = sm->dt[(uint32_t)br.buf & DMASK];
      if (e_kind(d) == kKindSlow) {
        d = slow_decode(br.buf, kTableDist, sm->d_count, sm->d_first, sm->d_offs, sm->d_sorted);
        if (e_nbits(d) == 0) { status = kStatusDataError; break; }
      }
      br.drop((int)e_nbits(d));
      uint32_t dist = e_value(d) + br.take((int)e_extra(d));
      if (br.overrun()) { status = kStatusTruncated; break; }
      if (dist > o.produced()) { status = kStatusDataError; break; }
      if (o.vpos + len > o.vcap) { status = kStatusOutOfSpace; break; }
      o.copy(len, dist, g);
      o.maybe_flush(g);
    }
    if (status == kStatusOk && br.overrun()) status = kStatusTruncated;
  }
  o.flush(g, true);
  res.produced = o.produced();
  res.status = status;
  int64_t used = (int64_t)br.start_off + ((br.consumed() + 7) >> 3);
  res.consumed = used > (int64_t)in_len ? in_len : (uint32_t)used;
  return res;
}

}  // namespace inf
}  // namespace bitar
