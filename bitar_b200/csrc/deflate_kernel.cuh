// deflate_kernel.cuh -- K1 + K2 + K3 + K5: one CTA compresses one chunk into one raw DEFLATE stream.
//
// Pipeline inside a CTA (16 warps; 4 in the small-chunk instance), per <=64 KiB sub-block:
//   load    : the sub-block is pulled into shared memory with a 1-D TMA bulk copy (cp.async.bulk +
//             mbarrier), so match extension and literal look-ups never touch HBM again.
//   far     : for every position the last earlier position BEFORE ITS 2 KiB SUB-RANGE that holds the same 4 bytes
//             (a block-wide table of up to 8192 u16 entries keyed by a 13-bit hash; the CTA walks the sub-ranges in
//             order: look up, barrier, insert, barrier), written as one u16 per position to an L2-resident scratch.
//             These FAR candidates give the match finder the 32 KiB window of the reference's level-1 zlib.
//   match   : a token never straddles a sub-range boundary (deflate_common.h: that is what lets the inflate kernels
//             Huffman-decode the 32 sub-ranges of a block in parallel), so the parse of each sub-range is
//             independent: every WARP owns one sub-range at a time, with its own 2-way table of 512 buckets in
//             shared memory (the two most recent positions of a hash), and walks it in windows of 32 positions:
//             4-byte hash, the two most recent earlier positions with that hash (inside the window through
//             __match_any_sync, else the table), both matches extended, the far candidate extended where the near
//             match is shorter than 8 bytes, greedy parse of the window (a walk over its match lanes, one shuffle
//             each) with the carry in a register.  No CTA-wide barrier, no atomics on the tables, deterministic.
//   count   : literal/length and distance frequencies with shared-memory atomics.
//   plan    : rank sort of the used symbols, two-queue Huffman merge (the one serial step: thread 0 for the
//             literal/length tree, thread 32 for the distance tree, CRC-32 / Adler-32 of the block on the
//             other warps meanwhile), then in parallel: node depths, zlib's over-long-tree repair, length
//             assignment, canonical codes, body sizes, code-length RLE (one run per thread) and the 19-symbol
//             code-length code (one warp).  Same results as dfl::build_dynamic_plan (deflate_common.h).
//   encode  : cheapest of stored / fixed / dynamic; every thread encodes 8 consecutive tokens,
//             a block-wide prefix sum of code lengths (warp shuffles) gives each thread its bit
//             offset, codes are packed into a shared-memory stage and leave the SM as aligned
//             16-byte vector stores.  The bit offsets of the sub-range starts are collected on the way
//             and appended as the parallel-inflate index.
//
// The token stream between match and encode is compact: one u32 per token (literal byte, or length and
// distance), each sub-range's tokens in order at its own offset of a per-CTA global scratch area (L2
// resident, written and read once).
//
// Output is bit-identical to tools/model/deflate_model.h (tests pin this) and always a valid
// RFC 1951 stream that zlib inflates to the input.
//
// Replaces: the compress ops assembled at /root/reference/src/memory.cc:350-430 and executed behind
// src/device.cc:464-535 with the xform of src/config.cc:83-91 (DEFLATE, level 1, fixed | dynamic).
//
// This header is a template over four macros and may be included once per configuration (capi.cu includes it twice):
//   BITAR_DK_NS        namespace of the instance                     default dk    small chunks: dks
//   BITAR_DK_WARPS     warps per CTA                                         16                  4
//   BITAR_DK_BLOCK_MAX bytes per block (= largest chunk when < 64 KiB)       65536               16384
//   BITAR_DK_MIN_CTAS  resident CTAs per SM the kernel is compiled for       2                   6
// The output does not depend on the configuration (the far pass and the plan work per block, the match phase per 2 KiB
// sub-range; the far table's size follows the block length, not the instance);
// the small instance exists because a 4 KiB chunk keeps 2 of a CTA's warps busy: it trades warps per CTA for
// CTAs per SM.  It is only valid for chunks of at most BITAR_DK_BLOCK_MAX bytes.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/bitar_cuda.h"
#include "checksum.h"
#include "deflate_common.h"

#ifndef BITAR_DK_RANK_SORT
#define BITAR_DK_RANK_SORT 1       // 0: the bitonic network in the 16-warp instance (A/B)
#endif
#ifndef BITAR_DK_NS
#define BITAR_DK_NS dk
#define BITAR_DK_WARPS 16
#define BITAR_DK_BLOCK_MAX 65536
#define BITAR_DK_MIN_CTAS 2
#endif

namespace bitar {
namespace BITAR_DK_NS {

constexpr int kWarps = BITAR_DK_WARPS;
constexpr int kThreads = kWarps * 32;
constexpr int kBlockMax = BITAR_DK_BLOCK_MAX;    // sub-block size (positions fit 16 bits)
constexpr uint32_t kMaxSeg = kBlockMax < 65536 ? (uint32_t)kBlockMax : (uint32_t)BITAR_MAX_SEG_SIZE;   // largest chunk
constexpr int kNearBits = 9;                     // per-warp 2-way table: 512 buckets for a 2048-position sub-range
constexpr int kNearSlots = (1 << kNearBits) + 32;  // + one always-empty dummy slot per lane
constexpr int kFarTabMax = 1 << (kBlockMax >= 65536 ? 13 : 11);   // dfl::far_hash_bits(kBlockMax) entries at most
constexpr int kFarNeed = 8;                      // the far candidate is tried where the near match is shorter (model: far_need)
constexpr int kFarMin = 4;                       // shortest far match (model: far_min)
constexpr uint32_t kNoFar = 0xFFFFu;
constexpr int kTokPerThread = 8;                 // encode: consecutive tokens per thread
constexpr int kTile = kThreads * kTokPerThread;  // encode: tokens per tile (4096 with 16 warps, at most 48 bits each)
constexpr int kStageWords = kTile * 3 / 2 + 160; // bit stage: a tile's 6 bytes per token + the block header (<= 141 words) + the partial unit
constexpr int kSeqPerThread = (328 + kThreads - 1) / kThreads;   // plan: code lengths (<= 316) per thread
constexpr int kHdrPerThread = (19 + 316 + kThreads - 1) / kThreads;   // header: code-length items per thread
constexpr uint32_t kTokNone = 0x100u;            // compact token stream: padding (emits nothing)
constexpr uint32_t kTokEob = 0x101u;             //                       end of block
// other compact tokens: < 0x100 literal byte; bit 31 set: a match with its distance symbol already worked out by the match
// phase (which needs it for the frequency counts anyway):
//   bits 0..7 length - 3, 8..22 distance - 1, 23..27 distance symbol
// (the encode phase looks the length code up by length - 3 in a per-block table, the distance code by symbol)
constexpr uint32_t kNoCand = 0xFFFFu;
constexpr int kWinUnroll = 1;                    // windows per trip of the match loop (1 / 2 / 3 / 4 in one gpurun call: 56.6 / 56.1 / 56.4 / 56.6 GB/s)

struct PlanPar {                      // scratch of the parallel half of the plan
  uint32_t ll_bl[16], d_bl[16];      // leaves per code length
  uint32_t ll_at[16], d_at[16];      // canonical next-code counters
  uint32_t ll_over, d_over;          // nodes below the length limit
  uint32_t hlit_max, hdist_max;      // highest used symbol
  unsigned long long dyn_bits, fix_bits;
  uint32_t d_node_freq[64];
  uint16_t d_parent[64];
  uint8_t seq[328];                  // litlen lengths then distance lengths, as the code-length RLE scans them
};
struct EncodeArea {
  uint32_t stage[kStageWords];       // output bit stage, stage[0] is virtual byte `sbase`
  uint32_t sort_keys[512];
  dfl::PlanScratch scratch;
  PlanPar pp;
};
struct __align__(16) Smem {
  union {                            // the match phase and the plan/encode phases never overlap in time
    struct {
      union {
        uint32_t head[kWarps][kNearSlots];            // per warp: hash -> the two most recent positions, 16 bits each (kNoCand = empty)
        uint32_t far_tab[kFarTabMax];                 // far pass (before the match phase): hash -> last position + 1 before the sub-range
      };
      // pre-shifted token fields (tok_pack): length - 3 -> (length symbol index << 26) | (length - 3);
      // distance - 1 -> distance symbol << 23 through zlib's two-level map d < 256 ? lut[d] : lut[256 + (d >> 7)]
      uint32_t len_sym_lut[256];
      uint32_t dist_sym_lut[512];
    } m;                             // (the tables are longer than the bit stage: the look-up tables lie behind it)
    EncodeArea enc;
  } u;
  uint32_t keep[4];                  // the partially filled 16-byte unit of the stage while the tables use its space
  uint32_t ll_freq[288];
  uint32_t d_freq[32];
  uint32_t ll_enc[288];              // code | (length << 16) under the chosen block type
  uint32_t d_enc[32];
  uint32_t len_enc[256];             // ll_enc of the length symbol of every length - 3
  uint32_t d_sorted[32];
  uint32_t warp_sums[kThreads / 32];
  uint32_t crc_tab[256];
  uint32_t x2n[32];
  uint32_t cks_crc, cks_a, cks_b;    // checksum accumulators
  uint32_t ll_m, d_m;
  uint32_t next_idx;                 // the chunk this CTA compresses next (fetched early, its input is prefetched)
  uint32_t sub_cnt[32];              // tokens of each sub-range of the block (compact, at tokens + sub * 2048)
  uint32_t sub_voff[34];             // encode: first slot of each sub-range (counts rounded up to kTokPerThread)
  uint32_t index[((kMaxSeg + 65535u) >> dfl::kIdxBlockLog2) * 33 + 4];   // parallel-inflate index of the chunk (deflate_common.h)
  uint32_t any_coded;                // some Huffman-coded block spans more than one sub-range
  uint32_t block_type;
  uint32_t tile_bits;
  unsigned long long mbar;           // TMA completion barrier
  dfl::BlockPlan plan;
  // input block, shifted so that raw + (src & 15) is the first byte.  Last member: a launch whose chunks are all
  // shorter than kBlockMax allocates only what they need (smem_bytes), which raises the CTAs per SM.
  __align__(16) uint8_t raw[kBlockMax + 48];
};
inline size_t smem_bytes(uint32_t max_len) {
  const uint32_t len = max_len < (uint32_t)kBlockMax ? (max_len + 15u) & ~15u : (uint32_t)kBlockMax;
  return sizeof(Smem) - (size_t)kBlockMax + len;
}

// ---- small helpers -----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D TMA bulk copy global -> shared (16-byte aligned addresses, size multiple of 16)
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                            unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// L2 prefetch of a global range (16-byte aligned address, size a multiple of 16): the next chunk's first block is
// requested while the current chunk is being compressed, so that its TMA load -- over PCIe when the input lives
// in pinned host memory -- does not start from cold
__device__ __forceinline__ void tma_prefetch_l2(const void* src_gmem, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_gmem), "r"(bytes) : "memory");
}

// Shared-memory reads of the input block go through explicit 32-bit shared addresses (`ds` = shared
// address of byte 0 of the block): a generic pointer would compile to LD.E with 64-bit address math.
__device__ __forceinline__ uint32_t lds_u32(uint32_t saddr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr));
  return v;
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t saddr) {
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(saddr));
  return v;
}
// unaligned little-endian 32-bit read (two aligned loads + funnel shift)
__device__ __forceinline__ uint32_t ld32u(uint32_t saddr) {
  uint32_t a = saddr & ~3u;
  return __funnelshift_r(lds_u32(a), lds_u32(a + 4), (saddr & 3u) * 8u);
}

// Lengths of the common prefixes of the bytes at p with the bytes at three candidates (all in the block held in shared
// memory): the two near candidates c1 / c2, at most max1 / max2 bytes, and the far candidate c3, at most max3 bytes, which
// only counts where both near matches stay below kFarNeed bytes (len3 = 0 otherwise).  max = 0 switches a candidate off
// (the caller then passes p itself as the candidate).  All streams are read as aligned words, one new word per stream
// and step; the words of p are loaded once for the three.  The first 8 bytes go without a branch (most matches of
// columnar data end there); the loop then runs while any candidate still matches.
__device__ __forceinline__ int prefix8(uint32_t x0, uint32_t x1) {
  // equal leading bytes of the first differing word: clz(brev(z)) >> 3 is 0..3, or 4 when z == 0 (both words equal)
  const uint32_t z = x0 ? x0 : x1;
  return (x0 ? 0 : 4) + (__clz((int)__brev(z)) >> 3);
}
__device__ __forceinline__ void match_len3(uint32_t ds, int p, int c1, int c2, int c3, int max1, int max2, int max3, int& len1, int& len2,
                                           int& len3) {
  const uint32_t ap = (ds + (uint32_t)p) & ~3u, a1 = (ds + (uint32_t)c1) & ~3u, a2 = (ds + (uint32_t)c2) & ~3u, a3 = (ds + (uint32_t)c3) & ~3u;
  const uint32_t sp = ((ds + (uint32_t)p) & 3u) * 8u, s1 = ((ds + (uint32_t)c1) & 3u) * 8u, s2 = ((ds + (uint32_t)c2) & 3u) * 8u,
                 s3 = ((ds + (uint32_t)c3) & 3u) * 8u;
  const uint32_t p0 = lds_u32(ap), p1 = lds_u32(ap + 4u);
  const uint32_t q0 = lds_u32(a1), q1 = lds_u32(a1 + 4u), r0 = lds_u32(a2), r1 = lds_u32(a2 + 4u), t0 = lds_u32(a3), t1 = lds_u32(a3 + 4u);
  uint32_t wp = lds_u32(ap + 8u), w1 = lds_u32(a1 + 8u), w2 = lds_u32(a2 + 8u), w3 = lds_u32(a3 + 8u);
  const uint32_t pa = __funnelshift_r(p0, p1, sp), pb = __funnelshift_r(p1, wp, sp);
  const uint32_t x0 = pa ^ __funnelshift_r(q0, q1, s1), x1 = pb ^ __funnelshift_r(q1, w1, s1);
  const uint32_t y0 = pa ^ __funnelshift_r(r0, r1, s2), y1 = pb ^ __funnelshift_r(r1, w2, s2);
  const uint32_t z0 = pa ^ __funnelshift_r(t0, t1, s3), z1 = pb ^ __funnelshift_r(t1, w3, s3);
  int l1 = prefix8(x0, x1), l2 = prefix8(y0, y1), l3 = prefix8(z0, z1);
  // the far candidate counts where the near matches are short: min(l, max) < kFarNeed for both (kFarNeed == 8: a near
  // match that reaches 8 bytes here either stops there or grows in the loop)
  static_assert(kFarNeed == 8, "the far rule is decided on the first 8 bytes");
  const bool far_on = min(l1, max1) < kFarNeed && min(l2, max2) < kFarNeed;
  bool g1 = l1 == 8 && max1 > 8, g2 = l2 == 8 && max2 > 8, g3 = far_on && l3 == 8 && max3 > 8;
  int l = 8;                                          // bytes compared so far by a candidate that is still going
  while (g1 || g2 || g3) {                            // 8 bytes per round
    const uint32_t np0 = lds_u32(ap + (uint32_t)l + 4u), np1 = lds_u32(ap + (uint32_t)l + 8u);
    const uint32_t m10 = lds_u32(a1 + (uint32_t)l + 4u), m11 = lds_u32(a1 + (uint32_t)l + 8u);
    const uint32_t m20 = lds_u32(a2 + (uint32_t)l + 4u), m21 = lds_u32(a2 + (uint32_t)l + 8u);
    const uint32_t m30 = lds_u32(a3 + (uint32_t)l + 4u), m31 = lds_u32(a3 + (uint32_t)l + 8u);
    const uint32_t xp0 = __funnelshift_r(wp, np0, sp), xp1 = __funnelshift_r(np0, np1, sp);
    if (g1) {
      const uint32_t x0 = xp0 ^ __funnelshift_r(w1, m10, s1), x1 = xp1 ^ __funnelshift_r(m10, m11, s1);
      l1 = l + prefix8(x0, x1);
      g1 = (x0 | x1) == 0 && l + 8 < max1;
    }
    if (g2) {
      const uint32_t y0 = xp0 ^ __funnelshift_r(w2, m20, s2), y1 = xp1 ^ __funnelshift_r(m20, m21, s2);
      l2 = l + prefix8(y0, y1);
      g2 = (y0 | y1) == 0 && l + 8 < max2;
    }
    if (g3) {
      const uint32_t z0 = xp0 ^ __funnelshift_r(w3, m30, s3), z1 = xp1 ^ __funnelshift_r(m30, m31, s3);
      l3 = l + prefix8(z0, z1);
      g3 = (z0 | z1) == 0 && l + 8 < max3;
    }
    wp = np1;
    w1 = m11;
    w2 = m21;
    w3 = m31;
    l += 8;
  }
  len1 = min(l1, max1);
  len2 = min(l2, max2);
  len3 = far_on ? min(l3, max3) : 0;
}

// ---- output stream: bit stage in shared memory, flushed as aligned 16-byte vectors ------------------
// Byte positions are "virtual" (offset + (dst & 15)) so that multiples of 16 are 16-byte aligned
// addresses.  All fields are CTA-uniform values held redundantly in registers.
struct OutStream {
  uint8_t* vbase;      // dst - mis
  uint32_t vstart;     // mis
  uint32_t vcap;       // mis + dst_cap
  uint32_t sbase;      // virtual byte position of stage[0] (multiple of 16)
  uint32_t vflushed;   // bytes below are in global memory
  uint64_t bit;        // next bit to write, in virtual bits (8 * virtual byte + bit)

  __device__ __forceinline__ void init(uint8_t* dst, uint32_t cap) {
    uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 15u);
    vbase = dst - mis;
    vstart = mis;
    vcap = mis + cap;
    sbase = 0;
    vflushed = mis;
    bit = 8ull * mis;
  }
  __device__ __forceinline__ uint32_t produced_bytes() const { return (uint32_t)((bit + 7) >> 3) - vstart; }
};

// OR `nbits` (<= 32) bits of v into the stage at virtual bit position `at`
__device__ __forceinline__ void stage_or(Smem& sm, const OutStream& o, uint64_t at, uint32_t v, int nbits) {
  if (nbits == 0) return;
  uint32_t rel = (uint32_t)(at - 8ull * o.sbase);
  uint32_t w = rel >> 5, sh = rel & 31u;
  atomicOr(&sm.u.enc.stage[w], v << sh);
  if (sh + (uint32_t)nbits > 32u) atomicOr(&sm.u.enc.stage[w + 1], v >> (32u - sh));
}

// Flush the stage up to virtual byte `upto` (all threads).  When `slide`, whole 16-byte units below
// `upto` are retired and the partially filled unit moves to the front of the stage.
__device__ void stream_flush(Smem& sm, OutStream& o, uint32_t upto, bool slide) {
  __syncthreads();
  const uint8_t* sbytes = reinterpret_cast<const uint8_t*>(sm.u.enc.stage);
  // when the stage slides: the 4 words of the partial unit move to its front (read now, written after the barrier)
  const uint32_t new_base = upto & ~15u;
  const uint32_t shift_words = slide ? (new_base - o.sbase) >> 2 : 0u;
  uint32_t keep = 0;
  if (shift_words && threadIdx.x < 4) keep = sm.u.enc.stage[shift_words + threadIdx.x];
  uint32_t lo = o.vflushed, hi = upto;
  if (hi > lo) {
    uint32_t a = (lo + 15u) & ~15u, b = hi & ~15u;
    if (a > b) {
      for (uint32_t v = lo + threadIdx.x; v < hi; v += kThreads) o.vbase[v] = sbytes[v - o.sbase];
    } else {
      for (uint32_t v = lo + threadIdx.x; v < a; v += kThreads) o.vbase[v] = sbytes[v - o.sbase];
      for (uint32_t v = a + 16u * threadIdx.x; v < b; v += 16u * kThreads)
        *reinterpret_cast<uint4*>(o.vbase + v) = *reinterpret_cast<const uint4*>(sbytes + (v - o.sbase));
      for (uint32_t v = b + threadIdx.x; v < hi; v += kThreads) o.vbase[v] = sbytes[v - o.sbase];
    }
    o.vflushed = hi;
  }
  if (shift_words == 0) return;
  __syncthreads();   // every read of the stage is done
  // clear everything that was used; words 0..3 (first iteration of threads 0..3) receive the partial unit instead
  const uint32_t used_words = min((uint32_t)kStageWords, shift_words + 8u);
  for (uint32_t i = threadIdx.x; i < used_words; i += kThreads) sm.u.enc.stage[i] = i < 4u ? keep : 0u;
  o.sbase = new_base;
  __syncthreads();
}

// serial bit writer used by thread 0 for block headers
struct HeaderWriter {
  Smem& sm;
  const OutStream& o;
  uint64_t at;
  __device__ __forceinline__ void put(uint32_t v, int n) {
    stage_or(sm, o, at, v, n);
    at += (uint64_t)n;
  }
};

// CTA-wide bitonic sort of sm.u.enc.sort_keys[0..512)
__device__ void sort512(Smem& sm) {
  for (uint32_t k = 2; k <= 512; k <<= 1) {
    for (uint32_t j = k >> 1; j > 0; j >>= 1) {
      __syncthreads();
      for (uint32_t i = threadIdx.x; i < 512; i += kThreads) {
        uint32_t ixj = i ^ j;
        if (ixj > i) {
          uint32_t a = sm.u.enc.sort_keys[i], b = sm.u.enc.sort_keys[ixj];
          bool up = (i & k) == 0;
          if ((a > b) == up) {
            sm.u.enc.sort_keys[i] = b;
            sm.u.enc.sort_keys[ixj] = a;
          }
        }
      }
    }
  }
  __syncthreads();
}

// The same order by rank: a key's place is the number of smaller keys (keys are distinct: the symbol is part of the
// key).  Every thread scans all 288 keys for each of its own with broadcast 16-byte reads; no barrier per step,
// which is what the bitonic network costs a small CTA.  sm.ll_m (the number of used symbols) must be complete.
__device__ void sort_rank(Smem& sm) {
  uint32_t* tmp = sm.u.enc.scratch.sorted;
  const uint4* k4 = reinterpret_cast<const uint4*>(sm.u.enc.sort_keys);
  __syncthreads();
  for (int i = threadIdx.x; i < 288; i += kThreads) {
    const uint32_t k = sm.u.enc.sort_keys[i];
    if (k == 0xFFFFFFFFu) continue;
    uint32_t r = 0;
#pragma unroll 8
    for (int j = 0; j < 288 / 4; ++j) {
      const uint4 v = k4[j];
      r += (uint32_t)(v.x < k) + (uint32_t)(v.y < k) + (uint32_t)(v.z < k) + (uint32_t)(v.w < k);
    }
    tmp[r] = k;
  }
  __syncthreads();
  const int m = (int)sm.ll_m;
  for (int i = threadIdx.x; i < 288; i += kThreads) sm.u.enc.sort_keys[i] = i < m ? tmp[i] : 0xFFFFFFFFu;
  __syncthreads();
}

// ---- far pass ---------------------------------------------------------------------------------------
// far[p] = the last position c before the 2 KiB sub-range of p whose 4 bytes hash like those at p (kNoFar when the table's
// entry for the hash of p is empty or lies farther back than max_dist; a candidate that holds other bytes is weeded out
// by the match phase).  The whole CTA takes the
// sub-ranges in order: look up all positions of the sub-range, barrier, insert them (the largest position of a hash
// wins, whatever the order: atomicMax), barrier.  Semantics == the far pass of tools/model/deflate_model.h.
__device__ __forceinline__ void far_pass(Smem& sm, uint32_t ds, int n, int max_dist, uint16_t* __restrict__ far, int tid) {
  constexpr int kPer = ((int)dfl::kSub + kThreads - 1) / kThreads;   // positions per thread and sub-range
  constexpr bool kExact = kPer * kThreads == (int)dfl::kSub;           // (else the last pass of the CTA is partly idle)
  const int fb = dfl::far_hash_bits((uint32_t)n);
  uint32_t* tab = sm.u.m.far_tab;
  for (int i = tid; i < (1 << fb); i += kThreads) tab[i] = 0u;
  __syncthreads();
  for (int s0 = 0; s0 < n; s0 += (int)dfl::kSub) {
    uint32_t h[kPer];
    if (kPer == 4 && kExact && s0 + (int)dfl::kSub + 4 <= n) {
      // a full sub-range with a thread per 4 CONSECUTIVE positions: their four 4-byte windows come out of three aligned
      // words, their far candidates leave as one 8-byte store
      const int p0 = s0 + 4 * tid;
      const uint32_t a = (ds + (uint32_t)p0) & ~3u, sh = (ds + (uint32_t)p0) & 3u;
      const uint32_t w0 = lds_u32(a), w1 = lds_u32(a + 4u), w2 = lds_u32(a + 8u);
      uint32_t f[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t o = sh + (uint32_t)k;            // byte offset of the window in (w0, w1, w2): 0 .. 6
        const uint32_t w = __funnelshift_r(o < 4u ? w0 : w1, o < 4u ? w1 : w2, (o & 3u) * 8u);
        h[k] = dfl::hash_far(w, fb);
        f[k] = kNoFar;
        if (s0) {
          const uint32_t c1 = tab[h[k]];
          if (c1 && p0 + k - (int)(c1 - 1u) <= max_dist) f[k] = c1 - 1u;
        }
      }
      *reinterpret_cast<uint2*>(far + p0) = make_uint2(f[0] | (f[1] << 16), f[2] | (f[3] << 16));
      __syncthreads();
#pragma unroll
      for (int k = 0; k < 4; ++k) atomicMax(&tab[h[k]], (uint32_t)(p0 + k + 1));
      __syncthreads();
      continue;
    }
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
      const int p = s0 + tid + k * kThreads;
      h[k] = 0xFFFFFFFFu;                             // no position / no hash
      if (p + 4 <= n && (kExact || tid + k * kThreads < (int)dfl::kSub)) {
        const uint32_t w = ld32u(ds + (uint32_t)p);
        h[k] = dfl::hash_far(w, fb);
        uint32_t f = kNoFar;
        if (s0) {
          const uint32_t c1 = tab[h[k]];              // position + 1, 0 = empty
          // (whether the candidate's 4 bytes really equal those at p is left to the match phase: its extension of the
          // far candidate yields fewer than kFarMin bytes otherwise, which the selection rejects -- same tokens, one
          // unaligned read less per position here)
          if (c1 && p - (int)(c1 - 1u) <= max_dist) f = c1 - 1u;
        }
        far[p] = (uint16_t)f;
      }
    }
    if (s0 + (int)dfl::kSub >= n) break;              // the last sub-range's positions are nobody's far candidates
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kPer; ++k)
      if (h[k] != 0xFFFFFFFFu) atomicMax(&tab[h[k]], (uint32_t)(s0 + tid + k * kThreads + 1));
    __syncthreads();
  }
  __syncthreads();                                    // far[] (global) and the table's space are handed to the match phase
}

// ---- match phase -------------------------------------------------------------------------------------
// One warp, one 2 KiB sub-range [s0, s1) of the block: windows of 32 positions in order.
// Semantics == the near pass / selection / parse of tools/model/deflate_model.h: candidates = the two most recent earlier
// positions with the same 4-byte hash -- lower lanes of the window first, then the two ways of this sub-range's table;
// the longer match wins (ties: the nearer); a near match shorter than kFarNeed yields to a strictly longer far match of
// at least kFarMin bytes; greedy parse in position order with a one-position lazy step.
__device__ __forceinline__ void match_subrange(Smem& sm, uint32_t ds, int n, int s0, int s1, int warp, int lane, int max_dist,
                                               const uint16_t* far, uint32_t* __restrict__ tokens) {
  constexpr unsigned kFull = 0xFFFFFFFFu;
  volatile uint32_t* head = sm.u.m.head[warp];          // low half: the most recent position, high half: the one before; [512 ..]: per-lane dummy slots, always empty
  for (int i = lane; i < kNearSlots; i += 32) head[i] = 0xFFFFFFFFu;
  __syncwarp();
  const unsigned lt_mask = (1u << lane) - 1u;
  const uint32_t dummy = (1u << kNearBits) + (uint32_t)lane;
  const int sub_end = min(n, s0 + (int)dfl::kSub);
  int carry = s0;                                     // next token start
  // this sub-range's tokens go to tokens[s0 ..], compact and in order
  uint32_t* __restrict__ tokp = tokens + s0;
  uint32_t cnt = 0;
  // far candidates of a window, read one window ahead (sub-range 0 has none: the pass wrote kNoFar; entries of positions
  // without four bytes are never used: `valid` below; the scratch has a window of slack past the block)
  const uint16_t* farp = far + s0 + lane;
  uint32_t far_next = __ldcg(farp);
#pragma unroll kWinUnroll
  for (int base = s0; base < s1; base += 32) {
    const int p = base + lane;
    const bool valid = p + 4 <= n;
    const uint32_t fc = valid ? far_next : kNoFar;
    farp += 32;
    far_next = __ldcg(farp);                          // the next window's, one iteration ahead
    const uint32_t w4 = ld32u(ds + p);                // (the block is padded: reading past its end is harmless)
    const uint32_t h = valid ? dfl::hash_near(w4, kNearBits) : dummy;
    const unsigned m = __match_any_sync(kFull, h);    // dummies are unique per lane
    const unsigned lower = m & lt_mask;
    const uint32_t t01 = head[h], t0 = t01 & 0xFFFFu, t1 = t01 >> 16;
    const int k1 = 31 - __clz((int)lower);            // nearest lower lane with this hash (-1: none)
    const unsigned lower2 = lower & ~(lower ? 1u << k1 : 0u);
    __syncwarp();
    if (valid && (m >> lane) == 1u) {                 // the window's highest position for this hash: the bucket after the window
      head[h] = (uint32_t)p | ((lower ? (uint32_t)(base + k1) : t0) << 16);
    }
    __syncwarp();
    const uint32_t cand1 = lower ? (uint32_t)(base + k1) : t0;
    const uint32_t cand2 = lower ? (lower2 ? (uint32_t)(base + 31 - __clz((int)lower2)) : t0) : t1;
    const int a = carry - base;                       // where the parse enters this window (>= 0)
    if (a >= 32) continue;                            // the whole window lies inside the previous match
    // every lane runs the (branch-free) first 8 bytes of the match extension; lanes without a usable candidate
    // compare their position with itself under a length limit of 0
    const bool live = valid && p >= carry;
    const int maxl = min(dfl::kMaxMatch, sub_end - p);
    const bool has1 = live && cand1 != kNoCand && p - (int)cand1 <= max_dist;
    const bool has2 = live && cand2 != kNoCand && p - (int)cand2 <= max_dist;
    const bool has3 = live && fc != kNoFar;           // the far candidate: tried where the near matches are short
    int len1, len2, len3;
    match_len3(ds, p, has1 ? (int)cand1 : p, has2 ? (int)cand2 : p, has3 ? (int)fc : p, has1 ? maxl : 0, has2 ? maxl : 0, has3 ? maxl : 0, len1,
               len2, len3);
    int len = len1, c = (int)cand1;
    if (len2 > len1) {
      len = len2;
      c = (int)cand2;
    }
    if (len3 >= kFarMin && len3 > len) {
      len = len3;
      c = (int)fc;
    }
    const bool hit = len >= dfl::kMinMatch;
    int adv = hit ? len : 1, dist = hit ? p - c : 0;
    {   // lazy step: a match yields to a strictly longer match that starts at the next position of the window
      const int next_adv = __shfl_down_sync(kFull, adv, 1);
      if (lane < 31 && adv > 1 && next_adv > adv) {
        adv = 1;
        dist = 0;
      }
    }
    // Greedy parse of the window from a: the token starts are the positions reached from a by t -> t + adv[t].  Every
    // lane collects the positions reached from ITS position, doubling the number of hops per round (five rounds cover
    // the 32 hops of a window of literals); lane a's set is the parse.
    unsigned reach;
    {
      const int nxt = lane + adv;
      unsigned r = (lt_mask + 1u) | (nxt < 32 ? 1u << nxt : 0u);
#pragma unroll
      for (int round = 0; round < 5; ++round) r |= __shfl_sync(kFull, r, 31 - __clz((int)r));
      reach = __shfl_sync(kFull, r, a);
      const int last = 31 - __clz((int)reach);      // the last token start of the window
      carry = base + last + __shfl_sync(kFull, adv, last);
    }
    const bool start = (reach & (lt_mask + 1u)) != 0u && p < n;   // lt_mask + 1 == this lane's bit
    const unsigned starts = __ballot_sync(kFull, start);
    {   // token + symbol counts, without a literal / match branch (the loads are harmless for the other kind)
      const bool is_match = adv > 1;
      const uint32_t byte = w4 & 0xFFu;
      const uint32_t d1 = is_match ? (uint32_t)dist - 1u : 0u;
      const uint32_t len_f = sm.u.m.len_sym_lut[is_match ? adv - 3 : 0];                  // symbol << 26 | length - 3
      const uint32_t dist_f = sm.u.m.dist_sym_lut[d1 < 256u ? d1 : 256u + (d1 >> 7)];     // symbol << 23
      if (start) {
        atomicAdd(&sm.ll_freq[is_match ? 257u + (len_f >> 26) : byte], 1u);
        if (is_match) atomicAdd(&sm.d_freq[dist_f >> 23], 1u);
        // (one 32-bit index from the scratch base: a single wide multiply-add forms the address)
        tokp[cnt + (uint32_t)__popc(starts & lt_mask)] = is_match ? (0x80000000u | (len_f & 0xFFu) | dist_f | (d1 << 8)) : byte;
      }
    }
    cnt += (uint32_t)__popc(starts);
  }
  if (lane == 0) sm.sub_cnt[s0 >> dfl::kSubLog2] = cnt;
}

// ---- plan, first half, in parallel --------------------------------------------------------------------
// The same code lengths, codes, hlit / hdist and body sizes as dfl::build_dynamic_plan() (the model pins
// this bit for bit), computed by the whole CTA: only the two-queue Huffman merge stays serial, and it keeps
// its queue heads in registers.

// Two-queue merge over leaves sorted ascending by (freq << 9 | symbol); same picks (ties prefer leaves) as
// dfl::huff_lengths_from_sorted.  Leaves 0..m-1, internal nodes m..2m-2, parent[] for every node but the root.
// Sentinels instead of bounds checks (this one thread is bound by its instruction count, ~30 per pick): the caller
// sets key[m] = 0xFFFFFFFF (as a frequency: 2^23 - 1, above every real sum) and node_freq[m .. 2m-1] = 0xFFFFFFFF
// (a node that does not exist yet).
__device__ void huff_merge(const uint32_t* __restrict__ key, int m, uint32_t* node_freq, uint16_t* parent) {
  constexpr uint32_t kInf = 0xFFFFFFFFu;
  uint32_t lf = key[0] >> 9, lf2 = key[1] >> 9, nf = kInf, nf2 = kInf;   // heads of the leaf / internal queues
  int leaf = 0, inode = m, next = m;
  // branch-free picks (a taken branch costs this single thread ~20 cycles, a select 2): both queues' next entries
  // are loaded unconditionally and chosen by the comparison
  for (int k = 0; k < m - 1; ++k) {
    uint32_t f = 0;
#pragma unroll
    for (int pick = 0; pick < 2; ++pick) {
      const bool tl = lf <= nf;                       // ties prefer leaves
      f += tl ? lf : nf;
      parent[tl ? leaf : inode] = (uint16_t)next;
      leaf += tl ? 1 : 0;
      inode += tl ? 0 : 1;
      const uint32_t new_lf2 = key[leaf + 1] >> 9;
      const uint32_t new_nf2 = node_freq[inode + 1];
      lf = tl ? lf2 : lf;
      nf = tl ? nf : nf2;
      lf2 = tl ? new_lf2 : lf2;
      nf2 = tl ? nf2 : new_nf2;
    }
    node_freq[next] = f;
    nf = inode == next ? f : nf;
    nf2 = inode + 1 == next ? f : nf2;
    ++next;
  }
}

// depth of every node (walk to the root), leaves counted per clamped length, nodes below the limit counted
__device__ __forceinline__ void huff_depths(const uint16_t* parent, int m, int max_bits, uint32_t* bl, uint32_t* over,
                                            int t, int nthr) {
  const int root = 2 * m - 2;
  for (int i = t; i < root; i += nthr) {
    int d = 0, j = i;
    while (j != root) {
      j = parent[j];
      ++d;
    }
    if (i < m) atomicAdd(&bl[min(d, max_bits)], 1u);
    if (d > max_bits) atomicAdd(over, 1u);
  }
}

// zlib gen_bitlen repair of an over-long tree (serial; rare)
__device__ __forceinline__ void huff_repair(uint32_t* bl, int overflow, int max_bits) {
  while (overflow > 0) {
    int bits = max_bits - 1;
    while (bl[bits] == 0) bits--;
    bl[bits]--;
    bl[bits + 1] += 2;
    bl[max_bits]--;
    overflow -= 2;
  }
}

// longest codes to the least frequent symbols
__device__ __forceinline__ void huff_assign(const uint32_t* key, int m, int max_bits, const uint32_t* bl, uint8_t* len,
                                            int t, int nthr) {
  for (int i = t; i < m; i += nthr) {
    int acc = 0;
    for (int bits = max_bits; bits >= 1; --bits) {
      const int c = (int)bl[bits];
      if (i < acc + c) {
        len[key[i] & 511u] = (uint8_t)bits;
        break;
      }
      acc += c;
    }
  }
}

// canonical codes (bit-reversed) by one warp: next_code[len]++ in symbol order, 32 symbols per round
__device__ __forceinline__ void huff_codes_warp(const uint8_t* len, int n, const uint32_t* bl, uint32_t* at, uint16_t* code,
                                                int lane) {
  constexpr unsigned kFull = 0xFFFFFFFFu;
  uint32_t c = 0, mine = 0;
  for (int b = 1; b <= dfl::kMaxBits; ++b) {
    c = (c + bl[b - 1]) << 1;
    if (lane == b) mine = c;
  }
  if (lane >= 1 && lane <= dfl::kMaxBits) at[lane] = mine;
  __syncwarp();
  const unsigned lt_mask = (1u << lane) - 1u;
  for (int base = 0; base < n; base += 32) {
    const int i = base + lane;
    const uint32_t l = i < n ? len[i] : 0u;
    const unsigned peers = __match_any_sync(kFull, l);
    if (i < n) code[i] = l ? (uint16_t)dfl::bitrev(at[l] + (uint32_t)__popc(peers & lt_mask), (int)l) : (uint16_t)0;
    __syncwarp();
    if (l && (peers & lt_mask) == 0) at[l] += (uint32_t)__popc(peers);
    __syncwarp();
  }
}

// CTA-wide exclusive prefix sum (one value per thread); *total receives the sum.  Two barriers.
__device__ __forceinline__ uint32_t block_excl_scan(Smem& sm, uint32_t v, uint32_t* total, int lane, int warp) {
  uint32_t incl = v;
#pragma unroll
  for (int o2 = 1; o2 < 32; o2 <<= 1) {
    const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o2);
    if (lane >= o2) incl += t;
  }
  __syncthreads();   // warp_sums free again
  if (lane == 31) sm.warp_sums[warp] = incl;
  __syncthreads();
  uint32_t off = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < kThreads / 32; ++w) {
    const uint32_t ws = sm.warp_sums[w];
    if (w < warp) off += ws;
    tot += ws;
  }
  *total = tot;
  return off + incl - v;
}

// Code-length RLE (dfl::cl_rle over the litlen lengths, then over the distance lengths) with one run per
// position: thread t owns positions [t * kSeqPerThread, +kSeqPerThread) of the concatenated length sequence (one
// position per thread with 16 warps) and emits the tokens of the runs that start there.  Same tokens, in the same
// order, as the serial function.
__device__ __forceinline__ void cl_rle_parallel(Smem& sm, int tid, int lane, int warp) {
  dfl::BlockPlan& pl = sm.plan;
  const int hlit = pl.hlit, total_syms = pl.hlit + pl.hdist;
  if (tid < dfl::kNumCl) sm.u.enc.scratch.cl_freq[tid] = 0;
  uint8_t* seq = sm.u.enc.pp.seq;
  for (int i = tid; i < total_syms; i += kThreads) seq[i] = i < hlit ? pl.ll_len[i] : pl.d_len[i - hlit];
  __syncthreads();
  int v[kSeqPerThread], run[kSeqPerThread];
  uint32_t ntok = 0;
#pragma unroll
  for (int k = 0; k < kSeqPerThread; ++k) {
    const int pos = tid * kSeqPerThread + k;
    v[k] = 0;
    run[k] = 0;
    if (pos < total_syms) {
      v[k] = seq[pos];
      const bool start = pos == 0 || pos == hlit || seq[pos - 1] != v[k];
      if (start) {
        const int end = pos < hlit ? hlit : total_syms;
        int r = 1;
        while (pos + r < end && ((pos + r) & 3) && seq[pos + r] == v[k]) ++r;   // to a word boundary
        if (((pos + r) & 3) == 0) {                                             // then four lengths per step
          const uint32_t pat = (uint32_t)v[k] * 0x01010101u;
          while (pos + r + 4 <= end && *reinterpret_cast<const uint32_t*>(seq + pos + r) == pat) r += 4;
        }
        while (pos + r < end && seq[pos + r] == v[k]) ++r;
        run[k] = r;
        if (v[k] == 0) {
          const int rem = r % 138;
          ntok += (uint32_t)(r / 138 + (rem >= 3 ? 1 : rem));
        } else {
          const int rem = (r - 1) % 6;
          ntok += (uint32_t)(1 + (r - 1) / 6 + (rem >= 3 ? 1 : rem));
        }
      }
    }
  }
  uint32_t total = 0;
  uint32_t o = block_excl_scan(sm, ntok, &total, lane, warp);
  uint32_t* cl_freq = sm.u.enc.scratch.cl_freq;
  const auto emit = [&](int sym, int extra) {
    pl.cl_tok[o++] = (uint16_t)(sym | (extra << 5));
    atomicAdd(&cl_freq[sym], 1u);
  };
#pragma unroll
  for (int k = 0; k < kSeqPerThread; ++k) {
    int r = run[k];
    if (r <= 0) continue;
    if (v[k] == 0) {
      while (r >= 11) {
        const int c = r > 138 ? 138 : r;
        emit(18, c - 11);
        r -= c;
      }
      if (r >= 3) {
        emit(17, r - 3);
        r = 0;
      }
      while (r-- > 0) emit(0, 0);
    } else {
      emit(v[k], 0);
      r--;
      while (r >= 3) {
        const int c = r > 6 ? 6 : r;
        emit(16, c - 3);
        r -= c;
      }
      while (r-- > 0) emit(v[k], 0);
    }
  }
  if (tid == 0) pl.n_cl_tok = (int)total;
  __syncthreads();
}

// The code-length code (19 symbols) and the header size by one warp: dfl::plan_cl_tree() with the symbol-parallel
// steps of the big trees (rank sort through shuffles, depths, length assignment, canonical codes); only the merge
// stays serial.  cl_freq comes from cl_rle_parallel.  Same results as the serial function (the model uses that one).
__device__ __forceinline__ void plan_cl_tree_warp(Smem& sm, int lane) {
  constexpr unsigned kFull = 0xFFFFFFFFu;
  dfl::BlockPlan& pl = sm.plan;
  PlanPar& pp = sm.u.enc.pp;                       // the big trees are done: their counters and the distance
  uint32_t* sorted = sm.u.enc.scratch.sorted;      // tree's node arrays are free
  const uint32_t f = lane < dfl::kNumCl ? sm.u.enc.scratch.cl_freq[lane] : 0u;
  uint32_t fs = f;
  {
    const unsigned used = __ballot_sync(kFull, f != 0);
    if (__popc(used) == 1) {   // complete the code: a second symbol gets a 1-bit code too
      const int other = (__ffs((int)used) - 1) == 0 ? 1 : 0;
      if (lane == other) fs = 1;
    }
  }
  const int cm = __popc(__ballot_sync(kFull, fs != 0));
  const uint32_t key = fs ? (fs << 9) | (uint32_t)lane : 0xFFFFFFFFu;
  uint32_t rank = 0;
#pragma unroll
  for (int j = 0; j < dfl::kNumCl; ++j) rank += (uint32_t)(__shfl_sync(kFull, key, j) < key);
  if (fs) sorted[rank] = key;
  if (lane == 0) sorted[cm] = 0xFFFFFFFFu;                     // sentinels of the merge
  pp.d_node_freq[lane] = pp.d_node_freq[32 + lane] = 0xFFFFFFFFu;
  if (lane < 16) pp.ll_bl[lane] = 0;
  if (lane < dfl::kNumCl) pl.cl_len[lane] = 0;
  if (lane == 0) pp.ll_over = 0;
  __syncwarp();
  if (lane == 0) huff_merge(sorted, cm, pp.d_node_freq, pp.d_parent);
  __syncwarp();
  huff_depths(pp.d_parent, cm, dfl::kMaxClBits, pp.ll_bl, &pp.ll_over, lane, 32);
  __syncwarp();
  if (lane == 0) huff_repair(pp.ll_bl, (int)pp.ll_over, dfl::kMaxClBits);
  __syncwarp();
  huff_assign(sorted, cm, dfl::kMaxClBits, pp.ll_bl, pl.cl_len, lane, 32);
  __syncwarp();
  huff_codes_warp(pl.cl_len, dfl::kNumCl, pp.ll_bl, pp.ll_at, pl.cl_code, lane);
  const uint32_t my_len = lane < dfl::kNumCl ? pl.cl_len[lane] : 0u;
  const unsigned nz = __ballot_sync(kFull, lane < dfl::kNumCl && pl.cl_len[dfl::cl_order(lane)] != 0);
  const int hclen = max(4, 32 - __clz((int)nz));
  uint32_t hb = lane < dfl::kNumCl ? f * (my_len + (uint32_t)dfl::cl_extra_bits(lane)) : 0u;
  hb = __reduce_add_sync(kFull, hb);
  if (lane == 0) {
    pl.hclen = hclen;
    pl.header_bits = 3 + 5 + 5 + 4 + 3 * (uint32_t)hclen + hb;
  }
  __syncwarp();
}

// ---- checksum of the block held in shared memory (threads tid0..tid0+nthr) ---------------------------
__device__ __forceinline__ void block_checksum(Smem& sm, const uint8_t* d, uint32_t n, uint32_t tail_after,
                                               bool first_block, int type, int t, int nthr) {
  uint32_t per = (n + nthr - 1) / nthr;
  uint32_t lo = min(n, per * (uint32_t)t), hi = min(n, lo + per);
  if (lo >= hi) return;
  uint32_t state = (first_block && lo == 0) ? 0xFFFFFFFFu : 0u, s1 = 0, s2 = 0;
  const bool want_crc = type & BITAR_CHECKSUM_CRC32;
  for (uint32_t i = lo; i < hi; ++i) {
    uint32_t b = d[i];
    if (want_crc) state = cks::crc_step(state, b, sm.crc_tab);
    s1 += b;
    s2 += (hi - i) * b;  // per <= 128 -> < 2^23
  }
  uint32_t tail = (n - hi) + tail_after;
  if (want_crc) atomicXor(&sm.cks_crc, cks::crc_contrib(state, tail, sm.x2n));
  if (type & BITAR_CHECKSUM_ADLER32) {
    atomicAdd(&sm.cks_a, s1 % cks::kAdlerMod);
    atomicAdd(&sm.cks_b, cks::adler_b_contrib(s1 % cks::kAdlerMod, s2 % cks::kAdlerMod, tail % cks::kAdlerMod));
  }
}

// bits of one compact token under sm.ll_enc / sm.d_enc; also returns the two code words
__device__ __forceinline__ int token_bits(const Smem& sm, uint32_t tok, uint32_t& lo_bits, int& lo_n, uint32_t& hi_bits,
                                          int& hi_n) {
  lo_n = hi_n = 0;
  lo_bits = hi_bits = 0;
  if (tok == kTokNone) return 0;
  if (tok < 0x200u) {   // literal byte, or end of block (0x101 -> symbol 256)
    const uint32_t e = sm.ll_enc[tok < 0x100u ? tok : (uint32_t)dfl::kEob];
    lo_bits = e & 0xFFFFu;
    lo_n = (int)(e >> 16);
    return lo_n;
  }
  // len_enc / d_enc entries carry the symbol's number of extra bits in bits 24..; the extra value is the low bits of
  // length - 3 / distance - 1 (RFC 1951's bases are aligned that way)
  const uint32_t e = sm.len_enc[tok & 0xFFu];
  const uint32_t cl = (e >> 16) & 0xFFu, leb = e >> 24;
  lo_bits = (e & 0xFFFFu) | (((tok & 0xFFu) & ((1u << leb) - 1u)) << cl);
  lo_n = (int)(cl + leb);  // <= 20
  const uint32_t f = sm.d_enc[(tok >> 23) & 31u];
  const uint32_t dl = (f >> 16) & 0xFFu, deb = f >> 24;
  hi_bits = (f & 0xFFFFu) | ((((tok >> 8) & 0x7FFFu) & ((1u << deb) - 1u)) << dl);
  hi_n = (int)(dl + deb);  // <= 28
  return lo_n + hi_n;
}

// ---- the kernel ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, BITAR_DK_MIN_CTAS)
    deflate_kernel(const bitar_chunk* __restrict__ ops, uint32_t n_ops, bitar_result* __restrict__ results,
                   unsigned int* __restrict__ counter, uint32_t* __restrict__ token_scratch, uint16_t* far_scratch, int huffman,
                   int checksum_type, int max_dist, int emit_index, unsigned long long* __restrict__ prof) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t* tokens = token_scratch + (size_t)blockIdx.x * kBlockMax;
  uint16_t* far = far_scratch + (size_t)blockIdx.x * kBlockMax;

  if (tid == 0) {
    mbar_init(&sm.mbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (checksum_type != BITAR_CHECKSUM_NONE) {
    for (int i = tid; i < 256; i += kThreads) sm.crc_tab[i] = cks::crc_table_entry((uint32_t)i);
    if (tid == 0) cks::crc_x2n_init(sm.x2n);
  }
  for (int i = tid; i < kStageWords; i += kThreads) sm.u.enc.stage[i] = 0;
  __syncthreads();
  uint32_t tma_parity = 0;
  // optional phase timers (cycles, thread 0): 0 load, 1 match, 2 sort, 3 Huffman merge (serial), 4 tables+header, 5 encode,
  // 6 finish, 7 depths / lengths / codes / sizes, 8 code-length RLE, 9 code-length code + block type (serial), 10 far pass
  long long t_prev = prof ? clock64() : 0;
#define BITAR_PHASE(k)                                          \
  if (prof && tid == 0) {                                       \
    long long t_now = clock64();                                \
    atomicAdd(&prof[k], (unsigned long long)(t_now - t_prev));  \
    t_prev = t_now;                                             \
  }

  uint32_t idx = blockIdx.x;
  while (idx < n_ops) {
    const bitar_chunk op = ops[idx];
    const uint8_t* src = static_cast<const uint8_t*>(op.src);
    const uint32_t total = op.src_len;
    OutStream o;
    o.init(static_cast<uint8_t*>(op.dst), op.dst_cap);
    uint32_t status = BITAR_OP_OK;
    if (tid == 0) {
      sm.cks_crc = 0;
      sm.cks_a = 0;
      sm.cks_b = 0;
      sm.any_coded = 0;
    }
    {
      const uint32_t idx_words = min(dfl::idx_entries(total), (uint32_t)(sizeof(sm.index) / 4));
      for (uint32_t i = tid; i < idx_words; i += kThreads) sm.index[i] = 0;
    }

    if (total > kMaxSeg) status = BITAR_OP_DATA_ERROR;   // (the C-ABI rejects such ops before the launch: capi.cu qp_submit)

    if (total == 0) {  // empty input: a fixed block holding only end-of-block (03 00), as zlib emits
      __syncthreads();
      if (op.dst_cap < 2) status = BITAR_OP_OUT_OF_SPACE;
      else {
        if (tid == 0) {
          HeaderWriter hw{sm, o, o.bit};
          hw.put(1, 1);
          hw.put(1, 2);
          hw.put(0, 7);
        }
        o.bit += 10;
      }
    }

    for (uint32_t off = 0; off < total && status == BITAR_OP_OK; off += kBlockMax) {
      const int n = (int)min((uint32_t)kBlockMax, total - off);
      const bool final_block = off + (uint32_t)n == total;
      // ---- load: TMA bulk copy of [src+off, +n) widened to 16-byte boundaries ----
      const uint8_t* g0 = src + off;
      const uint32_t gmis = (uint32_t)(reinterpret_cast<uintptr_t>(g0) & 15u);
      const uint32_t bytes = (gmis + (uint32_t)n + 15u) & ~15u;
      __syncthreads();  // previous block fully consumed
      if (tid == 0) {
        mbar_expect_tx(&sm.mbar, bytes);
        tma_load_1d(sm.raw, g0 - gmis, bytes, &sm.mbar);
        if (off == 0) {   // claim the next chunk now and ask L2 for its first block
          const uint32_t nx = gridDim.x + atomicAdd(counter, 1u);
          sm.next_idx = nx;
          if (nx < n_ops) {
            const uint8_t* ns = static_cast<const uint8_t*>(ops[nx].src);
            const uint32_t nl = min(ops[nx].src_len, (uint32_t)kBlockMax);
            const uint32_t nmis = (uint32_t)(reinterpret_cast<uintptr_t>(ns) & 15u);
            if (nl) tma_prefetch_l2(ns - nmis, (nmis + nl + 15u) & ~15u);
          }
        }
      }
      // the hash tables take the space of the bit stage for the duration of the match phase: park the
      // partially filled 16-byte unit at its front
      if (tid < 4) sm.keep[tid] = sm.u.enc.stage[tid];
      static_assert(sizeof(sm.u.m.head) >= sizeof(sm.u.enc.stage) || sizeof(sm.u.m.far_tab) >= sizeof(sm.u.enc.stage),
                    "the look-up tables must lie behind the bit stage");
      for (int i = tid; i < 256; i += kThreads) sm.u.m.len_sym_lut[i] = ((uint32_t)dfl::len_sym(i + 3) << 26) | (uint32_t)i;
      for (int i = tid; i < 512; i += kThreads)
        sm.u.m.dist_sym_lut[i] = (uint32_t)dfl::dist_sym(i < 256 ? i + 1 : ((i - 256) << 7) + 1) << 23;
      for (int i = tid; i < 288; i += kThreads) sm.ll_freq[i] = 0;
      if (tid < 32) sm.d_freq[tid] = 0;
      mbar_wait(&sm.mbar, tma_parity);
      tma_parity ^= 1u;
      __syncthreads();
      const uint8_t* d = sm.raw + gmis;
      BITAR_PHASE(0)

      // ---- match + parse + count: one 2 KiB sub-range per warp at a time, no CTA-wide barrier ----
      const uint32_t ds = smem_u32(d);
      far_pass(sm, ds, n, max_dist, far, tid);
      BITAR_PHASE(10)
      for (int s0 = warp * (int)dfl::kSub; s0 < n; s0 += kWarps * (int)dfl::kSub)
        match_subrange(sm, ds, n, s0, min(n, s0 + (int)dfl::kSub), warp, lane, max_dist, far, tokens);
      __syncthreads();
      for (int i = tid; i < kStageWords; i += kThreads) sm.u.enc.stage[i] = tid < 4 && i < 4 ? sm.keep[i] : 0u;
      BITAR_PHASE(1)
      // ---- plan: sort used symbols, Huffman lengths, header; checksums in parallel ----
      if (tid == 0) {
        sm.ll_freq[dfl::kEob] = 1;
        sm.ll_m = 0;
      }
      __syncthreads();
      {
        uint32_t used = 0;
        for (int i = tid; i < 512; i += kThreads) {
          uint32_t key = 0xFFFFFFFFu;
          if (i < dfl::kNumLitLen && sm.ll_freq[i]) key = (sm.ll_freq[i] << 9) | (uint32_t)i;
          sm.u.enc.sort_keys[i] = key;
          used += key != 0xFFFFFFFFu;
        }
        used = __reduce_add_sync(0xFFFFFFFFu, used);
        if (lane == 0 && used) atomicAdd(&sm.ll_m, used);
      }
      // sentinels of the Huffman merge: nodes that do not exist yet
      for (int i = tid; i < 2 * 288; i += kThreads) sm.u.enc.scratch.hs.node_freq[i] = 0xFFFFFFFFu;
      if (tid < 64) sm.u.enc.pp.d_node_freq[tid] = 0xFFFFFFFFu;
      if (kThreads >= 512 && !BITAR_DK_RANK_SORT) sort512(sm);
      else sort_rank(sm);
      BITAR_PHASE(2)
      PlanPar& pp = sm.u.enc.pp;
      dfl::HuffScratch& hs = sm.u.enc.scratch.hs;
      for (int i = tid; i < 288; i += kThreads) sm.plan.ll_len[i] = 0;
      if (tid < 32) sm.plan.d_len[tid] = 0;
      if (tid < 16) {
        pp.ll_bl[tid] = 0;
        pp.d_bl[tid] = 0;
      }
      if (tid == 0) {
        pp.ll_over = pp.d_over = pp.hlit_max = pp.hdist_max = 0;
        pp.dyn_bits = pp.fix_bits = 0;
        huff_merge(sm.u.enc.sort_keys, (int)sm.ll_m, hs.node_freq, hs.parent);   // the serial part
      } else if (tid == 32) {
        // distance tree: at least two used symbols (dummies of frequency 1, as zlib's build_tree)
        uint32_t df[32];
        int used = 0;
        for (int i = 0; i < dfl::kNumDist; ++i) {
          df[i] = sm.d_freq[i];
          used += df[i] != 0;
        }
        for (int i = 0; used < 2 && i < dfl::kNumDist; ++i)
          if (!df[i]) {
            df[i] = 1;
            used++;
          }
        const int dm = dfl::sort_used_small(df, dfl::kNumDist, sm.d_sorted);
        sm.d_sorted[dm < 31 ? dm : 31] = 0xFFFFFFFFu;   // sentinel (30 distance symbols at most)
        sm.d_m = (uint32_t)dm;
        huff_merge(sm.d_sorted, dm, pp.d_node_freq, pp.d_parent);
      } else if (warp >= 2 && checksum_type != BITAR_CHECKSUM_NONE) {
        block_checksum(sm, d, (uint32_t)n, total - (off + (uint32_t)n), off == 0, checksum_type, tid - 64,
                       kThreads - 64);
      }
      __syncthreads();
      BITAR_PHASE(3)
      {
        const int ll_m = (int)sm.ll_m, d_m = (int)sm.d_m;
        if (tid < kThreads - 64) huff_depths(hs.parent, ll_m, dfl::kMaxBits, pp.ll_bl, &pp.ll_over, tid, kThreads - 64);
        else huff_depths(pp.d_parent, d_m, dfl::kMaxBits, pp.d_bl, &pp.d_over, tid - (kThreads - 64), 64);
        __syncthreads();
        if (tid == 0) huff_repair(pp.ll_bl, (int)pp.ll_over, dfl::kMaxBits);
        if (tid == 32) huff_repair(pp.d_bl, (int)pp.d_over, dfl::kMaxBits);
        __syncthreads();
        if (tid < kThreads - 64) huff_assign(sm.u.enc.sort_keys, ll_m, dfl::kMaxBits, pp.ll_bl, sm.plan.ll_len, tid, kThreads - 64);
        else huff_assign(sm.d_sorted, d_m, dfl::kMaxBits, pp.d_bl, sm.plan.d_len, tid - (kThreads - 64), 64);
        __syncthreads();
        if (warp == 0) huff_codes_warp(sm.plan.ll_len, dfl::kNumLitLen, pp.ll_bl, pp.ll_at, sm.plan.ll_code, lane);
        else if (warp == 1) huff_codes_warp(sm.plan.d_len, dfl::kNumDist, pp.d_bl, pp.d_at, sm.plan.d_code, lane);
        else {
          // body sizes under the dynamic and the fixed code, highest used symbols: one symbol per thread
          unsigned long long dyn = 0, fix = 0;
          for (int i = tid - 64; i < dfl::kNumLitLen + dfl::kNumDist; i += kThreads - 64) {
            if (i < dfl::kNumLitLen) {
              const uint32_t f = sm.ll_freq[i];
              const int eb = i > 256 ? dfl::len_extra_bits(i - 257) : 0;
              dyn += (unsigned long long)f * (uint32_t)(sm.plan.ll_len[i] + eb);
              fix += (unsigned long long)f * (uint32_t)(dfl::fixed_ll_len(i) + eb);
              if (sm.plan.ll_len[i]) atomicMax(&pp.hlit_max, (uint32_t)i);
            } else {
              const int j = i - dfl::kNumLitLen;
              const uint32_t f = sm.d_freq[j];
              const int eb = dfl::dist_extra_bits(j);
              dyn += (unsigned long long)f * (uint32_t)(sm.plan.d_len[j] + eb);
              fix += (unsigned long long)f * (uint32_t)(5 + eb);
              if (sm.plan.d_len[j]) atomicMax(&pp.hdist_max, (uint32_t)j);
            }
          }
#pragma unroll
          for (int o2 = 16; o2 > 0; o2 >>= 1) {
            dyn += __shfl_xor_sync(0xFFFFFFFFu, dyn, o2);
            fix += __shfl_xor_sync(0xFFFFFFFFu, fix, o2);
          }
          if (lane == 0 && (dyn | fix)) {
            atomicAdd(&pp.dyn_bits, dyn);
            atomicAdd(&pp.fix_bits, fix);
          }
        }
        __syncthreads();
      }
      if (tid == 0) {
        sm.plan.hlit = max(257, (int)pp.hlit_max + 1);
        sm.plan.hdist = max(1, (int)pp.hdist_max + 1);
        sm.plan.dyn_body_bits = pp.dyn_bits;
        sm.plan.fixed_body_bits = pp.fix_bits;
      }
      __syncthreads();
      BITAR_PHASE(7)
      cl_rle_parallel(sm, tid, lane, warp);            // code-length RLE, one run per thread
      BITAR_PHASE(8)
      if (warp == 0) plan_cl_tree_warp(sm, lane);          // code-length code, header size
      if (tid == 0) {
        // choose the block type (same rule as the model)
        uint64_t dyn_bits = (uint64_t)sm.plan.header_bits + sm.plan.dyn_body_bits;
        uint64_t fix_bits = 3 + sm.plan.fixed_body_bits;
        int pieces = (n + 65534) / 65535;
        uint64_t cur = o.bit - 8ull * o.vstart;
        uint64_t stored_bits = ((cur + 3 + 7) / 8 * 8 - cur) + 32 + (uint64_t)n * 8 + (uint64_t)(pieces - 1) * 40;
        uint32_t type;
        uint64_t best;
        if (huffman == BITAR_HUFFMAN_FIXED) {
          type = fix_bits <= stored_bits ? dfl::kFixed : dfl::kStored;
          best = type == dfl::kFixed ? fix_bits : stored_bits;
        } else {
          type = dfl::kDynamic;
          best = dyn_bits;
          if (fix_bits <= best) {
            type = dfl::kFixed;
            best = fix_bits;
          }
          if (stored_bits < best) {
            type = dfl::kStored;
            best = stored_bits;
          }
        }
        sm.block_type = type;
        sm.tile_bits = (uint32_t)best;
      }
      __syncthreads();
      BITAR_PHASE(9)
      const uint32_t type = sm.block_type;
      {
        uint64_t end_bit = o.bit + sm.tile_bits;
        if (((end_bit + 7) >> 3) > (uint64_t)o.vcap) {
          status = BITAR_OP_OUT_OF_SPACE;
          break;
        }
      }
      const uint32_t iblk = (off >> dfl::kIdxBlockLog2) * 33u;   // this block's slots in the index
      if (tid == 0) {
        sm.index[iblk] = (uint32_t)(o.bit - 8ull * o.vstart);
        if (type != dfl::kStored && (uint32_t)n > dfl::kSub) sm.any_coded = 1;
      }

      if (type == dfl::kStored) {
        // ---- stored: header bits, byte align, then raw copy straight from shared memory ----
        int pieces = (n + 65534) / 65535;
        int done = 0;
        for (int k = 0; k < pieces; ++k) {
          int len = min(65535, n - done);
          bool last = final_block && k == pieces - 1;
          if (tid == 0) {
            HeaderWriter hw{sm, o, o.bit};
            hw.put(last ? 1u : 0u, 1);
            hw.put(0, 2);
          }
          o.bit = (o.bit + 3 + 7) & ~7ull;
          if (tid == 0) {
            HeaderWriter hw{sm, o, o.bit};
            hw.put((uint32_t)len, 16);
            hw.put((uint32_t)(~len) & 0xFFFFu, 16);
          }
          o.bit += 32;
          uint32_t vb = (uint32_t)(o.bit >> 3);
          stream_flush(sm, o, vb, false);
          // payload (byte stores are coalesced; this path only runs for incompressible blocks)
          for (int i = tid; i < len; i += kThreads) o.vbase[vb + i] = d[done + i];
          done += len;
          vb += (uint32_t)len;
          o.bit = 8ull * vb;
          // restart the stage at the new position
          __syncthreads();
          for (int i = tid; i < kStageWords; i += kThreads) sm.u.enc.stage[i] = 0;
          o.sbase = vb & ~15u;
          o.vflushed = vb;
          __syncthreads();
        }
        continue;
      }

      // ---- code tables for the encode loop ----
      for (int i = tid; i < 288; i += kThreads) {
        uint32_t e;
        if (type == dfl::kDynamic) e = (uint32_t)sm.plan.ll_code[i] | ((uint32_t)sm.plan.ll_len[i] << 16);
        else {
          int l = dfl::fixed_ll_len(i);
          uint32_t code = i < 144 ? 0x30u + i : i < 256 ? 0x190u + (i - 144) : i < 280 ? (uint32_t)(i - 256) : 0xC0u + (i - 280);
          e = dfl::bitrev(code, l) | ((uint32_t)l << 16);
        }
        if (i > 256 && i < 257 + 29) e |= (uint32_t)dfl::len_extra_bits(i - 257) << 24;
        sm.ll_enc[i] = e;
      }
      if (tid < 32) {
        uint32_t e;
        if (type == dfl::kDynamic) e = (uint32_t)sm.plan.d_code[tid] | ((uint32_t)sm.plan.d_len[tid] << 16);
        else e = dfl::bitrev((uint32_t)tid, 5) | (5u << 16);
        if (tid < dfl::kNumDist) e |= (uint32_t)dfl::dist_extra_bits(tid) << 24;
        sm.d_enc[tid] = e;
      }
      __syncthreads();
      for (int i = tid; i < 256; i += kThreads) sm.len_enc[i] = sm.ll_enc[257 + dfl::len_sym(i + 3)];
      // ---- header: fixed fields by thread 0, then one code-length-code length / RLE token per thread at the bit
      //      offset given by a CTA-wide prefix sum ----
      uint32_t hdr_bits = type == dfl::kDynamic ? sm.plan.header_bits : 3u;
      if (tid == 0) {
        HeaderWriter hw{sm, o, o.bit};
        hw.put(final_block ? 1u : 0u, 1);
        hw.put(type, 2);
        if (type == dfl::kDynamic) {
          hw.put((uint32_t)(sm.plan.hlit - 257), 5);
          hw.put((uint32_t)(sm.plan.hdist - 1), 5);
          hw.put((uint32_t)(sm.plan.hclen - 4), 4);
        }
      }
      if (type == dfl::kDynamic) {
        const dfl::BlockPlan& pl = sm.plan;
        // item i: the i-th code-length-code length (3 bits), then the RLE tokens; thread t owns items
        // [t * kHdrPerThread, +kHdrPerThread) (one item per thread with 16 warps)
        uint32_t bits[kHdrPerThread], nb[kHdrPerThread], mine = 0;
#pragma unroll
        for (int k = 0; k < kHdrPerThread; ++k) {
          const int i = tid * kHdrPerThread + k;
          bits[k] = nb[k] = 0;
          if (i < pl.hclen) {
            bits[k] = pl.cl_len[dfl::cl_order(i)];
            nb[k] = 3;
          } else if (i < pl.hclen + pl.n_cl_tok) {
            const int t = pl.cl_tok[i - pl.hclen], sym = t & 31, ev = t >> 5;
            bits[k] = (uint32_t)pl.cl_code[sym] | ((uint32_t)ev << pl.cl_len[sym]);
            nb[k] = (uint32_t)pl.cl_len[sym] + (uint32_t)dfl::cl_extra_bits(sym);
          }
          mine += nb[k];
        }
        uint32_t total = 0;
        uint32_t off2 = block_excl_scan(sm, mine, &total, lane, warp);
#pragma unroll
        for (int k = 0; k < kHdrPerThread; ++k) {
          stage_or(sm, o, o.bit + 17u + off2, bits[k], (int)nb[k]);
          off2 += nb[k];
        }
      }
      o.bit += hdr_bits;
      __syncthreads();
      BITAR_PHASE(4)

      // ---- encode: the compact tokens of all sub-ranges as one virtual sequence of slots; every sub-range
      //      is padded to a multiple of kTokPerThread so that a thread's slots never straddle two sub-ranges,
      //      and the end-of-block symbol gets a group of its own at the end ----
      const int n_sub = (int)dfl::idx_subs((uint32_t)n);
      if (warp == 0) {
        const uint32_t c = lane < n_sub ? (sm.sub_cnt[lane] + kTokPerThread - 1u) & ~(uint32_t)(kTokPerThread - 1) : 0u;
        uint32_t incl = c;
#pragma unroll
        for (int off2 = 1; off2 < 32; off2 <<= 1) {
          uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, off2);
          if (lane >= off2) incl += t;
        }
        sm.sub_voff[lane] = incl - c;
        if (lane == 31) sm.sub_voff[32] = sm.sub_voff[33] = incl;
      }
      __syncthreads();
      const int n_slots = (int)sm.sub_voff[32] + kTokPerThread;
      for (int tb = 0; tb < n_slots; tb += kTile) {
        const int g = tb + tid * kTokPerThread;
        uint32_t tk[kTokPerThread];
        int sr = 0;                          // largest sub-range with voff[sr] <= g (n_sub = the end-of-block group)
#pragma unroll
        for (int step = 16; step > 0; step >>= 1)
          if (sr + step <= n_sub && (int)sm.sub_voff[sr + step] <= g) sr += step;
        if (sr + 1 <= n_sub && (int)sm.sub_voff[sr + 1] <= g) sr += 1;
        const int l = g - (int)sm.sub_voff[sr];
        if (g >= n_slots) {
#pragma unroll
          for (int j = 0; j < kTokPerThread; ++j) tk[j] = kTokNone;
        } else if (sr >= n_sub) {
#pragma unroll
          for (int j = 0; j < kTokPerThread; ++j) tk[j] = j == 0 ? kTokEob : kTokNone;
        } else {
          const uint4 v = *reinterpret_cast<const uint4*>(tokens + sr * (int)dfl::kSub + l);
          const uint4 w = *reinterpret_cast<const uint4*>(tokens + sr * (int)dfl::kSub + l + 4);
          const int have = (int)sm.sub_cnt[sr] - l;
          tk[0] = v.x;
          tk[1] = have > 1 ? v.y : kTokNone;
          tk[2] = have > 2 ? v.z : kTokNone;
          tk[3] = have > 3 ? v.w : kTokNone;
          tk[4] = have > 4 ? w.x : kTokNone;
          tk[5] = have > 5 ? w.y : kTokNone;
          tk[6] = have > 6 ? w.z : kTokNone;
          tk[7] = have > 7 ? w.w : kTokNone;
        }
        uint32_t mybits = 0;
#pragma unroll
        for (int j = 0; j < kTokPerThread; ++j) {
          uint32_t lb, hb;
          int ln, hn;
          mybits += (uint32_t)token_bits(sm, tk[j], lb, ln, hb, hn);
        }
        // block-wide exclusive prefix sum of mybits
        uint32_t incl = mybits;
#pragma unroll
        for (int off2 = 1; off2 < 32; off2 <<= 1) {
          uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, off2);
          if (lane >= off2) incl += t;
        }
        if (lane == 31) sm.warp_sums[warp] = incl;
        __syncthreads();
        uint32_t warp_off = 0, tile_total = 0;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) {
          uint32_t ws = sm.warp_sums[w];
          if (w < warp) warp_off += ws;
          tile_total += ws;
        }
        uint64_t at = o.bit + warp_off + (incl - mybits);
        if (l == 0 && sr < n_sub && g < n_slots)   // first symbol of a sub-range
          sm.index[iblk + 1u + (uint32_t)sr] = (uint32_t)(at - 8ull * o.vstart);
        // emit: accumulate into a 64-bit window, flush whole words
        {
          uint32_t rel = (uint32_t)(at - 8ull * o.sbase);
          uint32_t w = rel >> 5;
          uint32_t fill = rel & 31u;   // bits already owned by the previous thread in word w
          uint64_t acc = 0;
          uint32_t accn = fill;
          bool first = true;
#pragma unroll
          for (int j = 0; j < kTokPerThread; ++j) {
            uint32_t lb, hb;
            int ln, hn;
            token_bits(sm, tk[j], lb, ln, hb, hn);
            acc |= (uint64_t)lb << accn;
            accn += (uint32_t)ln;
            if (accn >= 32) {
              if (first) { atomicOr(&sm.u.enc.stage[w], (uint32_t)acc); first = false; }
              else sm.u.enc.stage[w] = (uint32_t)acc;
              acc >>= 32; accn -= 32; ++w;
            }
            acc |= (uint64_t)hb << accn;
            accn += (uint32_t)hn;
            if (accn >= 32) {
              if (first) { atomicOr(&sm.u.enc.stage[w], (uint32_t)acc); first = false; }
              else sm.u.enc.stage[w] = (uint32_t)acc;
              acc >>= 32; accn -= 32; ++w;
            }
          }
          if (accn > (first ? fill : 0u) || (uint32_t)acc != 0u) atomicOr(&sm.u.enc.stage[w], (uint32_t)acc);
        }
        o.bit += tile_total;
        stream_flush(sm, o, (uint32_t)(o.bit >> 3) & ~15u, true);
      }
    }

    BITAR_PHASE(5)
    // ---- finish the chunk ----
    if (status == BITAR_OP_OK) {
      uint32_t end_byte = (uint32_t)((o.bit + 7) >> 3);
      if (end_byte > o.vcap) status = BITAR_OP_OUT_OF_SPACE;
      else {
        __syncthreads();
        // the parallel-inflate index goes after the byte-aligned end of the stream when it is useful and fits
        const uint32_t entries = dfl::idx_entries(total);
        if (emit_index && sm.any_coded && (uint64_t)end_byte + 4ull * (entries + 3u) <= (uint64_t)o.vcap) {
          const uint32_t end_bit = (uint32_t)(o.bit - 8ull * o.vstart);
          for (uint32_t t = tid; t < entries + 3u; t += kThreads) {
            const uint32_t w = t < entries ? sm.index[t] : t == entries ? end_bit : t == entries + 1u ? total : dfl::kIndexMagic;
            stage_or(sm, o, 8ull * end_byte + 32ull * t, w, 32);
          }
          end_byte += 4u * (entries + 3u);
          o.bit = 8ull * end_byte;
        }
        stream_flush(sm, o, end_byte, false);
      }
    }
    __syncthreads();
    if (tid == 0) {
      bitar_result r;
      r.status = status;
      r.produced = status == BITAR_OP_OK ? o.produced_bytes() : 0u;
      uint32_t crc = 0, adler = 0;
      if (checksum_type & BITAR_CHECKSUM_CRC32) crc = total ? (sm.cks_crc ^ 0xFFFFFFFFu) : 0u;
      if (checksum_type & BITAR_CHECKSUM_ADLER32) adler = cks::adler_finish(sm.cks_a, sm.cks_b, total);
      r.checksum = cks::pack(crc, adler);
      results[idx] = r;
    }
    // reset the stage for the next chunk and fetch its index
    __syncthreads();
    for (int i = tid; i < kStageWords; i += kThreads) sm.u.enc.stage[i] = 0;
    if (tid == 0 && (total == 0 || total > kMaxSeg)) sm.next_idx = gridDim.x + atomicAdd(counter, 1u);   // (no block was loaded)
    __syncthreads();
    idx = sm.next_idx;
    __syncthreads();
    BITAR_PHASE(6)
  }
#undef BITAR_PHASE
}

inline size_t deflate_scratch_bytes(int grid) { return (size_t)grid * kBlockMax * sizeof(uint32_t); }   // tokens
inline size_t deflate_far_bytes(int grid) { return ((size_t)grid * kBlockMax + 64u) * sizeof(uint16_t); }   // far candidates (+ a window of slack)

// resident CTAs per SM for chunks of at most max_len bytes (the shared-memory footprint follows the chunk size)
inline cudaError_t deflate_ctas_per_sm(int device, uint32_t max_len, int* ctas_out) {
  static bool configured[64] = {false};
  if (!configured[device & 63]) {
    cudaError_t e = cudaFuncSetAttribute(deflate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
    if (e != cudaSuccess) return e;
    cudaFuncSetAttribute(deflate_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    configured[device & 63] = true;
  }
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_out, deflate_kernel, kThreads, smem_bytes(max_len));
  if (e != cudaSuccess) return e;
  return *ctas_out < 1 ? cudaErrorLaunchOutOfResources : cudaSuccess;
}

// largest grid any launch on this device can use (sizes the token scratch)
inline cudaError_t deflate_grid(int device, int sm_count, int* grid_out) {
  int ctas = 0;
  cudaError_t e = deflate_ctas_per_sm(device, kBlockMax < 65536 ? 1u : (uint32_t)kBlockMax, &ctas);
  if (e != cudaSuccess) return e;
  *grid_out = sm_count * ctas;
  return cudaSuccess;
}

// launches a call of n ops is split into when `want` are asked for
inline uint32_t deflate_splits(uint32_t n, uint32_t want) {
  if (want > 8u) want = 8u;
  return (want < 1u || n < 4u * want) ? 1u : want;
}

inline cudaError_t deflate_launch(const bitar_chunk* ops, uint32_t n, bitar_result* res, unsigned int* counter,
                                  uint32_t* scratch, uint16_t* far, int device, int sm_count, int grid_override, uint32_t max_len,
                                  int huffman, int checksum_type, int max_dist, int emit_index, unsigned long long* prof,
                                  cudaStream_t stream, uint32_t splits = 1) {
  if (n == 0) return cudaSuccess;
  int ctas = 0;
  cudaError_t e = deflate_ctas_per_sm(device, max_len, &ctas);
  if (e != cudaSuccess) return e;
  // `splits` > 1: the call goes out as that many launches over consecutive parts of the op list (each with its own work
  // counter, at most 8).  A persistent grid holds every SM until its list is drained; between launches the kernels other
  // queue pairs have waiting -- an inflate whose copy-back should run beside this compress -- get their turn.
  splits = deflate_splits(n, splits);
  const uint32_t per = (n + splits - 1u) / splits;
  for (uint32_t k = 0, first = 0; first < n; ++k, first += per) {
    const uint32_t count = n - first < per ? n - first : per;
    int grid = grid_override > 0 ? grid_override : sm_count * ctas;
    if ((uint32_t)grid > count) grid = (int)count;
    deflate_kernel<<<grid, kThreads, smem_bytes(max_len), stream>>>(ops + first, count, res + first, counter + k, scratch, far, huffman,
                                                                    checksum_type, max_dist, emit_index, prof);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

}  // namespace BITAR_DK_NS
}  // namespace bitar

#undef BITAR_DK_NS
#undef BITAR_DK_WARPS
#undef BITAR_DK_BLOCK_MAX
#undef BITAR_DK_MIN_CTAS
