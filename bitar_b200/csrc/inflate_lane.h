// inflate_lane.h -- raw DEFLATE (RFC 1951) decoder, one LANE per chunk, written as a bounded-step
// state machine so that the 32 lanes of a warp (32 independent streams) stay converged.
//
// Why this shape (measured on B200): a DEFLATE stream decodes serially, so throughput is
// (streams in flight) / (latency of one symbol step).  A warp-per-stream decoder leaves 31 of 32 issue
// lanes idle and is issue-bound near 26 GB/s; sub-warp groups diverge and serialise.  Here every lane
// owns a stream and all lanes execute the same instruction sequence:
//     step():  [decode one litlen symbol (+ its distance)]  [copy <= kCopyStep bytes of a pending match]
//              [move <= kCopyStep stored bytes]  [flush one 16-byte vector]
// Long matches, stored blocks and flushes are spread over several steps instead of looping, block
// headers are parsed with all lanes of a wave entering together.
//
// Per-lane shared memory is small (u16 decode tables + a short output ring), so > 100 streams are
// resident per SM; code lengths, canonical-code side arrays (slow path for codes longer than the
// primary table) live in a per-lane global scratch that stays in L1/L2.
//
// BITAR_HD: the same source runs on the CPU with one lane (tests/test_core_host.py).
//
// Replaces the inflate half of the codec behind rte_compressdev (/root/reference/src/device.cc:240-318,
// xform at src/config.cc:93-105).
#pragma once
#include <stdint.h>

#include "deflate_common.h"
#include "inflate_core.h"  // status codes, BitReader, ld_in8/ld_in32, brev32

namespace bitar {
namespace infl {

using inf::BitReader;
using inf::kStatusDataError;
using inf::kStatusOk;
using inf::kStatusOutOfSpace;
using inf::kStatusTruncated;

// ---- table entries (u16) ----------------------------------------------------------------------------
// litlen: [3:0] code length (0 = not in the primary table: slow path / invalid)
//         [6:4] literal: 0 | length symbol: extra-bit count 0..5 | 6 = end of block
//         [7]   0 literal, 1 length / end of block
//         [15:8] literal byte, or (length base - 3)
// dist  : [3:0] code length (0 = slow path / invalid), [8:4] distance symbol 0..29
BITAR_HD uint16_t ll_entry(int sym, int nbits) {
  if (sym < 256) return (uint16_t)((sym << 8) | nbits);
  if (sym == 256) return (uint16_t)(0x80 | (6 << 4) | nbits);
  if (sym < 286) {
    int s = sym - 257;
    return (uint16_t)(((dfl::len_base(s) - 3) << 8) | 0x80 | (dfl::len_extra_bits(s) << 4) | nbits);
  }
  return 0;  // 286/287: invalid
}
BITAR_HD uint16_t d_entry(int sym, int nbits) { return sym < 30 ? (uint16_t)((sym << 4) | nbits) : (uint16_t)0; }
// second-level links (codes longer than the root index):
//   litlen: [3:0] sub-table index bits (2..6), [7:4] = 0xF, [15:8] sub-table offset / 4
//   dist  : [3:0] sub-table index bits,        [8:4] = 31,  [15:9] sub-table offset / 4
BITAR_HD uint16_t ll_link(int off, int sub_bits) { return (uint16_t)(((off >> 2) << 8) | 0xF0 | sub_bits); }
BITAR_HD uint16_t d_link(int off, int sub_bits) { return (uint16_t)(((off >> 2) << 9) | (31 << 4) | sub_bits); }

enum LaneState : uint32_t { kHeader = 0, kDecode = 1, kStored = 2, kFinish = 3, kDone = 4 };

constexpr int kCopyStep = 8;   // bytes of a match / stored block moved per step

// per-lane global scratch (header parsing + fallback slow path)
struct LaneScratch {
  uint8_t lens[320];
  uint16_t ll_sorted[288];
  uint16_t d_sorted[32];
  uint16_t ll_count[16], ll_first[16], ll_offs[16];
  uint16_t d_count[16], d_first[16], d_offs[16];
};

// LT / DT = total u16 entries of the litlen / distance table (root + sub-tables)
template <int LBITS, int LT, int DBITS, int DT, int RING>
struct LaneSmem {
  static_assert(LT >= (1 << LBITS) && DT >= (1 << DBITS) && LT % 4 == 0 && DT % 4 == 0, "table sizes");
  static constexpr int kBytes = 2 * LT + 2 * DT + RING;
  static constexpr int kStride = kBytes + 4;   // +4: lane i starts at bank i for equal offsets
};

// Builds the decode table (u16 entries: root of 2^tbits, then sub-tables for longer codes up to
// `capacity` entries) and the canonical side arrays for one alphabet.
// kind: 0 litlen, 1 dist, 2 code-length code (entry = sym << 4 | nbits, like dist).
// Codes whose sub-table does not fit `capacity` keep a 0 root entry and take the slow path.
BITAR_HD_NOINLINE uint32_t build_table_lane(const uint8_t* lens, int n, int kind, uint16_t* table, int tbits,
                                            int capacity, uint16_t* count, uint16_t* first, uint16_t* offs,
                                            uint16_t* sorted) {
  for (int b = 0; b < 16; ++b) count[b] = 0;
  for (int i = 0; i < n; ++i) count[lens[i]]++;
  int left = 1, maxl = 0;
  for (int b = 1; b <= 15; ++b) {
    left = (left << 1) - (int)count[b];
    if (count[b]) maxl = b;
    if (left < 0) return kStatusDataError;               // over-subscribed
  }
  int used = n - (int)count[0];
  if (left > 0 && used > 0 && (kind == 2 || maxl != 1)) return kStatusDataError;   // incomplete
  uint32_t f = 0, o = 0;
  uint16_t at[16];
  first[0] = offs[0] = at[0] = 0;
  for (int b = 1; b <= 15; ++b) {
    first[b] = (uint16_t)f;
    offs[b] = at[b] = (uint16_t)o;
    f = (f + count[b]) << 1;
    o += count[b];
  }
  for (int i = 0; i < n; ++i)
    if (lens[i]) sorted[at[lens[i]]++] = (uint16_t)i;
  uint32_t* t32 = reinterpret_cast<uint32_t*>(table);
  for (int j = 0; j < capacity / 2; ++j) t32[j] = 0;
  int idx = 0;
  for (; idx < used; ++idx) {
    int sym = sorted[idx];
    int l = lens[sym];
    if (l > tbits) break;                                 // sorted by length: the rest is longer too
    uint32_t code = (uint32_t)first[l] + (uint32_t)(idx - (int)offs[l]);
    uint32_t r = inf::brev32(code) >> (32 - l);
    const uint16_t e = kind == 0 ? ll_entry(sym, l) : kind == 1 ? d_entry(sym, l) : (uint16_t)((sym << 4) | l);
    for (uint32_t k = r; k < (1u << tbits); k += (1u << l)) table[k] = e;
  }
  // codes longer than the root: canonical order keeps codes with the same tbits-bit prefix contiguous
  int next_free = 1 << tbits;
  while (idx < used) {
    int l = lens[sorted[idx]];
    uint32_t code = (uint32_t)first[l] + (uint32_t)(idx - (int)offs[l]);
    const uint32_t prefix = code >> (l - tbits);
    int j = idx, lmax = l;
    while (j < used) {
      int l2 = lens[sorted[j]];
      uint32_t c2 = (uint32_t)first[l2] + (uint32_t)(j - (int)offs[l2]);
      if ((c2 >> (l2 - tbits)) != prefix) break;
      lmax = l2;
      ++j;
    }
    int sub_bits = lmax - tbits;
    if (sub_bits < 2) sub_bits = 2;
    const int size = 1 << sub_bits;
    if (next_free + size <= capacity) {
      table[inf::brev32(prefix) >> (32 - tbits)] = kind == 0 ? ll_link(next_free, sub_bits) : d_link(next_free, sub_bits);
      for (int k = idx; k < j; ++k) {
        int sym = sorted[k];
        int lk = lens[sym];
        uint32_t ck = (uint32_t)first[lk] + (uint32_t)(k - (int)offs[lk]);
        int rest = lk - tbits;                            // bits after the prefix
        uint32_t low = ck & ((1u << rest) - 1u);
        uint32_t r = inf::brev32(low) >> (32 - rest);
        const uint16_t e = kind == 0 ? ll_entry(sym, lk) : d_entry(sym, lk);
        for (int t = (int)r; t < size; t += (1 << rest)) table[next_free + t] = e;
      }
      next_free += size;
    }
    idx = j;
  }
  return kStatusOk;
}

// canonical bit-by-bit decode (codes longer than the primary table, or invalid prefixes): returns the
// entry with the full code length, or 0.
BITAR_HD_NOINLINE uint16_t slow_decode_lane(uint64_t buf, int kind, const uint16_t* count, const uint16_t* first,
                                            const uint16_t* offs, const uint16_t* sorted) {
  uint32_t code = 0;
  for (int l = 1; l <= 15; ++l) {
    code = (code << 1) | (uint32_t)((buf >> (l - 1)) & 1u);
    uint32_t rel = code - (uint32_t)first[l];
    if (code >= first[l] && rel < count[l]) {
      int sym = sorted[offs[l] + rel];
      return kind == 0 ? ll_entry(sym, l) : d_entry(sym, l);
    }
  }
  return 0;
}

// Shared-memory accessors.  On the device they take 32-bit shared addresses and emit LDS/STS directly
// (a generic pointer member would compile to LD/ST with 64-bit address math); on the host they are
// plain pointer accesses.
#if defined(__CUDACC__) && defined(BITAR_LANE_DEBUG)
__device__ unsigned int g_dbg[16];   // [0] count, [1] tag, [2..] details of the first violation, [10..11] smem window
#endif
#if defined(__CUDA_ARCH__)
typedef uint32_t saddr_t;
BITAR_HD saddr_t to_saddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
#if defined(BITAR_LANE_DEBUG)
// debug build: accesses outside the CTA's shared window / the chunk's buffers are recorded and skipped
__device__ __forceinline__ bool dbg_bad(uint32_t tag, unsigned long long a, unsigned long long lo, unsigned long long hi) {
  if (a >= lo && a < hi) return false;
  if (atomicAdd(&g_dbg[0], 1u) == 0) {
    g_dbg[1] = tag; g_dbg[2] = (unsigned)a; g_dbg[3] = (unsigned)(a >> 32); g_dbg[4] = (unsigned)lo; g_dbg[5] = (unsigned)(lo >> 32);
    g_dbg[6] = (unsigned)hi; g_dbg[7] = (unsigned)(hi >> 32); g_dbg[8] = blockIdx.x; g_dbg[9] = threadIdx.x;
  }
  return true;
}
#define BITAR_SCHK(a, n, tag) if (dbg_bad(tag, (a), g_dbg[10], g_dbg[11] - (n) + 1)) return 0
#define BITAR_SCHKV(a, n, tag) if (dbg_bad(tag, (a), g_dbg[10], g_dbg[11] - (n) + 1)) return
#else
#define BITAR_SCHK(a, n, tag)
#define BITAR_SCHKV(a, n, tag)
#endif
BITAR_HD uint32_t s_ld16(saddr_t a) { BITAR_SCHK(a, 2, 1); uint32_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
BITAR_HD uint32_t s_ld8(saddr_t a) { BITAR_SCHK(a, 1, 2); uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
BITAR_HD uint32_t s_ld32(saddr_t a) { BITAR_SCHK(a, 4, 3); uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
BITAR_HD void s_st8(saddr_t a, uint32_t v) { BITAR_SCHKV(a, 1, 4); asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
#else
typedef const uint8_t* saddr_t;
BITAR_HD saddr_t to_saddr(const void* p) { return static_cast<const uint8_t*>(p); }
BITAR_HD uint32_t s_ld16(saddr_t a) { return *reinterpret_cast<const uint16_t*>(a); }
BITAR_HD uint32_t s_ld8(saddr_t a) { return *a; }
BITAR_HD uint32_t s_ld32(saddr_t a) { return *reinterpret_cast<const uint32_t*>(a); }
BITAR_HD void s_st8(saddr_t a, uint32_t v) { *const_cast<uint8_t*>(a) = (uint8_t)v; }
#endif

template <int LBITS, int LT, int DBITS, int DT, int RING>
struct Lane {
  static_assert(RING >= 256 && (RING & (RING - 1)) == 0, "ring: power of two >= 256");
  static_assert(DBITS >= 7, "the distance table also hosts the 7-bit code-length code");
  static constexpr uint32_t RM = RING - 1;
  static constexpr uint32_t kNear = RING - 64;   // matches at most this far are served from the ring

  // shared memory of this lane
  uint16_t* lt;            // generic pointers: table construction (cold)
  uint16_t* dt;
  saddr_t lt_s, dt_s, ring_s;   // shared addresses: the hot path
  LaneScratch* sc;
  // stream state
  BitReader br;
  const uint8_t* in;
  uint32_t in_len;
  uint8_t* vbase;          // dst - (dst & 15)
  uint32_t vstart, vpos, vflushed, vcap;
  uint32_t state, status, last, blocks;
  uint32_t copy_rem, copy_dist;      // pending match
  uint32_t far_n, far_sh;            // far match: bytes to write next step, bit shift of the source
  uint64_t far_w0, far_w1;           // raw aligned words loaded last step (consumed one step later)
  uint32_t stored_rem, stored_at;    // pending stored payload (input offset)

  BITAR_HD void bind(uint8_t* smem_lane, LaneScratch* scratch) {
    lt = reinterpret_cast<uint16_t*>(smem_lane);
    dt = reinterpret_cast<uint16_t*>(smem_lane + 2 * LT);
    lt_s = to_saddr(smem_lane);
    dt_s = lt_s + 2 * LT;
    ring_s = dt_s + 2 * DT;
    sc = scratch;
    state = kDone;                  // a lane that never start()s must be inert in every step
    status = kStatusOk;
    in = nullptr;
    vbase = nullptr;
    in_len = vstart = vpos = vflushed = vcap = last = blocks = 0;
    copy_rem = copy_dist = far_n = far_sh = stored_rem = stored_at = 0;
    far_w0 = far_w1 = 0;
  }

  BITAR_HD void start(const uint8_t* src, uint32_t len, uint8_t* dst, uint32_t cap) {
    in = src;
    in_len = len;
    br.init(src, len, 0);
    uint32_t mis = (uint32_t)((uintptr_t)dst & 15u);
    vbase = dst - mis;
    vstart = vpos = vflushed = mis;
    vcap = mis + cap;
    state = kHeader;
    status = kStatusOk;
    last = 0;
    blocks = 0;
    copy_rem = stored_rem = far_n = 0;
  }
  BITAR_HD uint32_t produced() const { return vpos - vstart; }
  BITAR_HD void fail(uint32_t st) {
    status = st;
    state = kFinish;
    copy_rem = stored_rem = far_n = 0;
  }

  // ---- block header (serial per lane; lanes of a wave enter together) ----
  BITAR_HD void header() {
    if (last) {
      state = kFinish;
      return;
    }
    br.refill();
    last = br.take(1);
    uint32_t type = br.take(2);
    blocks++;
    if (br.overrun()) return fail(kStatusTruncated);
    if (type == 0) {
      br.drop(br.cnt & 7);
      br.refill();
      uint32_t len = br.take(16), nlen = br.take(16);
      if (br.overrun()) return fail(kStatusTruncated);
      if ((len ^ 0xFFFFu) != nlen) return fail(kStatusDataError);
      uint32_t at = br.start_off + (uint32_t)(br.consumed() >> 3);
      if ((uint64_t)at + len > in_len) return fail(kStatusTruncated);
      if (vpos + len > vcap) return fail(kStatusOutOfSpace);
      stored_rem = len;
      stored_at = at;
      state = kStored;
      if (len == 0) {
        br.init(in, in_len, at);
        state = kHeader;
      }
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
      nlen = (int)br.take(5) + 257;
      ndist = (int)br.take(5) + 1;
      int ncode = (int)br.take(4) + 4;
      if (nlen > 286 || ndist > 30) return fail(kStatusDataError);
      for (int i = 0; i < 19; ++i) sc->lens[i] = 0;
      for (int i = 0; i < ncode; ++i) {
        br.refill();
        sc->lens[dfl::cl_order(i)] = (uint8_t)br.take(3);
      }
      if (br.overrun()) return fail(kStatusTruncated);
      uint32_t st = build_table_lane(sc->lens, 19, 2, dt, 7, 128, sc->d_count, sc->d_first, sc->d_offs, sc->d_sorted);
      if (st != kStatusOk) return fail(st);
      int idx = 0, prev = 0;
      const int total = nlen + ndist;
      while (idx < total) {
        br.refill();
        uint32_t e = dt[br.peek(7)];
        if ((e & 15u) == 0) return fail(kStatusDataError);
        br.drop((int)(e & 15u));
        int sym = (int)(e >> 4), rep, val;
        if (sym < 16) { rep = 1; val = sym; prev = sym; }
        else if (sym == 16) {
          if (idx == 0) return fail(kStatusDataError);
          rep = 3 + (int)br.take(2); val = prev;
        } else if (sym == 17) { rep = 3 + (int)br.take(3); val = 0; prev = 0; }
        else { rep = 11 + (int)br.take(7); val = 0; prev = 0; }
        if (idx + rep > total) return fail(kStatusDataError);
        for (int k = 0; k < rep; ++k) sc->lens[idx + k] = (uint8_t)val;
        idx += rep;
      }
      if (br.overrun()) return fail(kStatusTruncated);
      if (sc->lens[256] == 0) return fail(kStatusDataError);
    }
    uint32_t st = build_table_lane(sc->lens + nlen, ndist, 1, dt, DBITS, DT, sc->d_count, sc->d_first, sc->d_offs, sc->d_sorted);
    if (st != kStatusOk) return fail(st);
    st = build_table_lane(sc->lens, nlen, 0, lt, LBITS, LT, sc->ll_count, sc->ll_first, sc->ll_offs, sc->ll_sorted);
    if (st != kStatusOk) return fail(st);
    state = kDecode;
  }

  // ---- one litlen symbol (and its distance) ----
  BITAR_HD void decode_one() {
    br.refill();
    uint32_t e = s_ld16(lt_s + 2u * ((uint32_t)br.buf & ((1u << LBITS) - 1u)));
    if ((e & 0xF0u) == 0xF0u)                              // second level: code longer than LBITS
      e = s_ld16(lt_s + 2u * (((e >> 8) << 2) + (((uint32_t)(br.buf >> LBITS)) & ((1u << (e & 15u)) - 1u))));
    if ((e & 15u) == 0) {
      e = slow_decode_lane(br.buf, 0, sc->ll_count, sc->ll_first, sc->ll_offs, sc->ll_sorted);
      if (e == 0) return fail(br.overrun() ? kStatusTruncated : kStatusDataError);
    }
    br.drop((int)(e & 15u));
    if (!(e & 0x80u)) {                                   // literal
      if (vpos >= vcap) return fail(br.overrun() ? kStatusTruncated : kStatusOutOfSpace);
      s_st8(ring_s + (vpos & RM), e >> 8);
      vpos++;
      return;
    }
    const uint32_t x = (e >> 4) & 7u;
    if (x == 6u) {                                        // end of block
      if (br.overrun()) return fail(kStatusTruncated);
      state = kHeader;
      return;
    }
    const uint32_t len = (e >> 8) + 3u + br.take((int)x);
    br.refill();
    uint32_t d = s_ld16(dt_s + 2u * ((uint32_t)br.buf & ((1u << DBITS) - 1u)));
    if (((d >> 4) & 31u) == 31u)
      d = s_ld16(dt_s + 2u * (((d >> 9) << 2) + (((uint32_t)(br.buf >> DBITS)) & ((1u << (d & 15u)) - 1u))));
    if ((d & 15u) == 0) {
      d = slow_decode_lane(br.buf, 1, sc->d_count, sc->d_first, sc->d_offs, sc->d_sorted);
      if (d == 0) return fail(br.overrun() ? kStatusTruncated : kStatusDataError);
    }
    br.drop((int)(d & 15u));
    const int ds = (int)(d >> 4);
    const uint32_t dist = (uint32_t)dfl::dist_base(ds) + br.take(dfl::dist_extra_bits(ds));
    if (br.overrun()) return fail(kStatusTruncated);
    if (dist > produced()) return fail(kStatusDataError);
    if (vpos + len > vcap) return fail(kStatusOutOfSpace);
    copy_rem = len;
    copy_dist = dist;
  }

  // ---- up to kCopyStep bytes of the pending match; trip = uniform bound supplied by the caller ----
  BITAR_HD void copy_some(uint32_t trip) {
    const uint32_t m = copy_rem < (uint32_t)kCopyStep ? copy_rem : (uint32_t)kCopyStep;
    const uint32_t src = vpos - copy_dist;
    if (copy_dist <= kNear) {
      if (copy_dist >= (uint32_t)kCopyStep) {             // no overlap inside the step: loads first
        uint32_t b[kCopyStep];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (uint32_t j = 0; j < (uint32_t)kCopyStep; ++j)
          if (j < trip && j < m) b[j] = s_ld8(ring_s + ((src + j) & RM));
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (uint32_t j = 0; j < (uint32_t)kCopyStep; ++j)
          if (j < trip && j < m) s_st8(ring_s + ((vpos + j) & RM), b[j]);
      } else {
        for (uint32_t j = 0; j < trip; ++j)
          if (j < m) s_st8(ring_s + ((vpos + j) & RM), s_ld8(ring_s + ((src + j) & RM)));
      }
    } else {
      // far: the source was flushed to global memory long ago.  Issue one unaligned 8-byte read now and
      // write the bytes to the ring at the start of the NEXT step: the load latency (L2) is then covered
      // by a whole step of other lanes' work instead of stalling the warp.
      const uint8_t* a = vbase + src;
      const uint64_t* a8 = reinterpret_cast<const uint64_t*>(reinterpret_cast<uintptr_t>(a) & ~(uintptr_t)7);
      const uint32_t sh = ((uint32_t)reinterpret_cast<uintptr_t>(a) & 7u) * 8u;
#if defined(__CUDA_ARCH__) && defined(BITAR_LANE_DEBUG)
      if (dbg_bad(10, (unsigned long long)a8, (unsigned long long)vbase, (unsigned long long)(vbase + vcap) - 15)) { far_n = m; far_w0 = far_w1 = 0; far_sh = 0; return; }
#endif
      far_w0 = a8[0];                                    // NOT consumed in this step
      far_w1 = a8[1];
      far_sh = sh;
      far_n = m;
      return;
    }
    vpos += m;
    copy_rem -= m;
  }

  BITAR_HD void far_commit() {
    const uint64_t v = far_sh ? (far_w0 >> far_sh) | (far_w1 << (64u - far_sh)) : far_w0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (uint32_t j = 0; j < (uint32_t)kCopyStep; ++j)
      if (j < far_n) s_st8(ring_s + ((vpos + j) & RM), (uint32_t)(v >> (8u * j)) & 0xFFu);
    vpos += far_n;
    copy_rem -= far_n;
    far_n = 0;
  }

  BITAR_HD void stored_some() {
    const uint32_t m = stored_rem < (uint32_t)kCopyStep ? stored_rem : (uint32_t)kCopyStep;
    for (uint32_t j = 0; j < m; ++j) s_st8(ring_s + ((vpos + j) & RM), inf::ld_in8(in + stored_at + j));
    vpos += m;
    stored_at += m;
    stored_rem -= m;
    if (stored_rem == 0) {
      br.init(in, in_len, stored_at);
      state = kHeader;
    }
  }

  // ---- move one 16-byte vector from the ring to global memory when one is complete ----
  BITAR_HD void flush_some() {
    if (vflushed & 15u) {                                 // chunk head: dst is not 16-byte aligned
      const uint32_t a = (vflushed + 15u) & ~15u;
      if (vpos < a) return;
      for (uint32_t v = vflushed; v < a; ++v) vbase[v] = (uint8_t)s_ld8(ring_s + (v & RM));
      vflushed = a;
      return;
    }
    if (vpos - vflushed < 16u) return;
    const saddr_t r = ring_s + (vflushed & RM);
    const uint32_t w0 = s_ld32(r), w1 = s_ld32(r + 4), w2 = s_ld32(r + 8), w3 = s_ld32(r + 12);
#if defined(__CUDA_ARCH__)
#if defined(BITAR_LANE_DEBUG)
    if (dbg_bad(11, (unsigned long long)(vbase + vflushed), (unsigned long long)vbase, (unsigned long long)(vbase + vcap))) { vflushed += 16u; return; }
#endif
    *reinterpret_cast<uint4*>(vbase + vflushed) = make_uint4(w0, w1, w2, w3);
#else
    uint32_t* o32 = reinterpret_cast<uint32_t*>(vbase + vflushed);
    o32[0] = w0; o32[1] = w1; o32[2] = w2; o32[3] = w3;
#endif
    vflushed += 16u;
  }

  BITAR_HD void finish() {
    for (;;) {
      const uint32_t before = vflushed;
      flush_some();
      if (vflushed == before) break;
    }
    for (uint32_t v = vflushed; v < vpos; ++v) vbase[v] = (uint8_t)s_ld8(ring_s + (v & RM));
    vflushed = vpos;
    state = kDone;
  }

  // One bounded step.  `trip` must be >= min(copy_rem, kCopyStep) of this lane (the kernel passes the
  // warp-wide maximum so the copy loop has a uniform trip count).
  BITAR_HD void step_pre() {
    if (state == kHeader && copy_rem == 0) header();
    if (state == kDecode && copy_rem == 0) decode_one();
  }
  BITAR_HD uint32_t want_copy() const {
    return state == kDone ? 0u : (copy_rem < (uint32_t)kCopyStep ? copy_rem : (uint32_t)kCopyStep);
  }
  BITAR_HD void step_post(uint32_t trip) {
    if (state == kDone) return;
    if (far_n) far_commit();
    else if (copy_rem) copy_some(trip);
    if (state == kStored) stored_some();
    flush_some();
    if (state == kFinish) finish();
  }

  BITAR_HD uint32_t consumed_bytes() const {
    int64_t used = (int64_t)br.start_off + ((br.consumed() + 7) >> 3);
    return used > (int64_t)in_len ? in_len : (uint32_t)used;
  }
};

}  // namespace infl
}  // namespace bitar
