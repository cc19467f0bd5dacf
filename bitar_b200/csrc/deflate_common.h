// deflate_common.h -- RFC 1951 building blocks shared by the sm_100a kernels and by the host-side
// model/unit tests (tools/model).  Everything here is plain serial code marked BITAR_HD so that the
// exact source that runs inside the kernels can be exercised on the CPU.
//
// Replaces (by construction, not by translation) the codec that the reference reaches through
// rte_compressdev_enqueue_burst (/root/reference/src/device.cc:464-488) with the xform built at
// /root/reference/src/config.cc:83-105: DEFLATE, level 1, dynamic or fixed Huffman.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define BITAR_HD __host__ __device__ __forceinline__
#define BITAR_HD_NOINLINE __host__ __device__ __noinline__
#else
#define BITAR_HD inline
#define BITAR_HD_NOINLINE inline
#endif

namespace bitar {
namespace dfl {

constexpr int kNumLitLen = 286;   // 0..255 literals, 256 EOB, 257..285 lengths
constexpr int kNumDist = 30;
constexpr int kNumCl = 19;
constexpr int kMaxBits = 15;
constexpr int kMaxClBits = 7;
constexpr int kMinMatch = 3;
constexpr int kMaxMatch = 258;
constexpr int kMaxDist = 32768;
constexpr int kEob = 256;

// ---- symbol maps (RFC 1951 3.2.5), computed instead of tabulated -------------------------------
BITAR_HD int ilog2(uint32_t v) {
#if defined(__CUDA_ARCH__)
  return 31 - __clz((int)v);
#else
  return 31 - __builtin_clz(v);
#endif
}

// match length 3..258 -> length symbol index 0..28 (symbol = 257 + index)
BITAR_HD int len_sym(int len) {
  int l = len - 3;
  if (l < 8) return l;
  if (len == 258) return 28;
  int n = ilog2((uint32_t)l);
  return 4 * (n - 1) + ((l >> (n - 2)) & 3);
}
BITAR_HD int len_extra_bits(int sym) {  // sym index 0..28
  return (sym < 8 || sym == 28) ? 0 : ((sym >> 2) - 1);
}
BITAR_HD int len_extra_val(int len, int sym) {
  int eb = len_extra_bits(sym);
  return (len - 3) & ((1 << eb) - 1);
}
BITAR_HD int len_base(int sym) {  // inverse map: first length of symbol index
  if (sym < 8) return sym + 3;
  if (sym == 28) return 258;
  int eb = (sym >> 2) - 1;
  return 3 + ((4 + (sym & 3)) << eb);
}

// distance 1..32768 -> distance symbol 0..29
BITAR_HD int dist_sym(int dist) {
  int d = dist - 1;
  if (d < 4) return d;
  int n = ilog2((uint32_t)d);
  return 2 * n + ((d >> (n - 1)) & 1);
}
BITAR_HD int dist_extra_bits(int sym) { return sym < 4 ? 0 : ((sym >> 1) - 1); }
BITAR_HD int dist_extra_val(int dist, int sym) {
  return (dist - 1) & ((1 << dist_extra_bits(sym)) - 1);
}
BITAR_HD int dist_base(int sym) {
  if (sym < 4) return sym + 1;
  int eb = (sym >> 1) - 1;
  return 1 + ((2 + (sym & 1)) << eb);
}

// fixed Huffman code lengths (RFC 1951 3.2.6)
BITAR_HD int fixed_ll_len(int sym) { return sym < 144 ? 8 : sym < 256 ? 9 : sym < 280 ? 7 : 8; }

BITAR_HD uint32_t bitrev(uint32_t code, int len) {
#if defined(__CUDA_ARCH__)
  return __brev(code) >> (32 - len);
#else
  uint32_t r = 0;
  for (int i = 0; i < len; ++i) r |= ((code >> i) & 1u) << (len - 1 - i);
  return r;
#endif
}

// order in which code-length-code lengths are transmitted
BITAR_HD int cl_order(int i) {
  // 16,17,18,0,8,7,9,6,10,5,11,4,12,3,13,2,14,1,15
  return i < 3 ? 16 + i : ((i & 1) ? (i == 3 ? 0 : 8 - ((i - 3) >> 1)) : 8 + ((i - 4) >> 1));
}

// ---- Huffman construction ------------------------------------------------------------------------
// Input: the used symbols sorted ascending by (freq, symbol): sorted_key[i] = (freq << 9) | symbol,
// i in [0, m).  Output: len[] for those symbols (other entries untouched -- caller zero-fills).
// Scratch: node_freq[2*m], parent[2*m] (uint16), both caller-provided.
// The construction is the classic two-queue merge on sorted leaves (optimal prefix code), followed by
// zlib-style length limiting on the per-length counts and re-assignment of lengths in sorted order
// (longest codes to the least frequent symbols).
struct HuffScratch {
  uint32_t node_freq[2 * 288];
  uint16_t parent[2 * 288];
  uint8_t depth[2 * 288];
};

BITAR_HD_NOINLINE void huff_lengths_from_sorted(const uint32_t* sorted_key, int m, int max_bits,
                                                uint8_t* len, uint16_t* bl_count /*[16]*/,
                                                HuffScratch* s) {
  for (int b = 0; b <= kMaxBits; ++b) bl_count[b] = 0;
  if (m == 0) return;
  if (m == 1) {  // callers avoid this for deflate trees; keep it valid anyway
    len[sorted_key[0] & 511] = 1;
    bl_count[1] = 1;
    return;
  }
  // leaves 0..m-1, internal nodes m..2m-2
  for (int i = 0; i < m; ++i) s->node_freq[i] = sorted_key[i] >> 9;
  int leaf = 0, inode = m, next = m;  // inode = head of the internal-node queue, next = next slot
  for (int k = 0; k < m - 1; ++k) {
    uint32_t f = 0;
    for (int pick = 0; pick < 2; ++pick) {
      bool take_leaf;
      if (leaf >= m) take_leaf = false;
      else if (inode >= next) take_leaf = true;
      else take_leaf = s->node_freq[leaf] <= s->node_freq[inode];  // ties prefer leaves (shallower)
      int c = take_leaf ? leaf++ : inode++;
      f += s->node_freq[c];
      s->parent[c] = (uint16_t)next;
    }
    s->node_freq[next++] = f;
  }
  int root = next - 1;
  s->depth[root] = 0;
  for (int i = root - 1; i >= 0; --i) s->depth[i] = (uint8_t)(s->depth[s->parent[i]] + 1);
  // zlib gen_bitlen counts EVERY node below the length limit (leaves and internal nodes, whose clamped
  // parents push them to max_bits + 1 again); the repair loop below relies on that count
  int overflow = 0;
  for (int i = 0; i < m; ++i) {
    int d = s->depth[i];
    if (d > max_bits) {
      d = max_bits;
      overflow++;
    }
    bl_count[d]++;
  }
  for (int i = m; i < root; ++i)
    if (s->depth[i] > max_bits) overflow++;
  if (overflow > 0) {
    // zlib gen_bitlen: move one leaf down from the deepest non-full level, pairing an overflow leaf
    do {
      int bits = max_bits - 1;
      while (bl_count[bits] == 0) bits--;
      bl_count[bits]--;
      bl_count[bits + 1] = (uint16_t)(bl_count[bits + 1] + 2);
      bl_count[max_bits]--;
      overflow -= 2;
    } while (overflow > 0);
  }
  // longest codes to the least frequent symbols (sorted ascending by freq)
  int i = 0;
  for (int bits = max_bits; bits >= 1; --bits)
    for (int c = bl_count[bits]; c > 0; --c) len[sorted_key[i++] & 511] = (uint8_t)bits;
}

// Canonical codes (RFC 1951 3.2.2), returned bit-reversed so they can be emitted LSB-first.
BITAR_HD_NOINLINE void huff_codes(const uint8_t* len, int n, const uint16_t* bl_count, uint16_t* code) {
  uint16_t next_code[kMaxBits + 2];
  uint32_t c = 0;
  next_code[0] = 0;
  for (int b = 1; b <= kMaxBits; ++b) {
    c = (c + bl_count[b - 1]) << 1;
    next_code[b] = (uint16_t)c;
  }
  for (int i = 0; i < n; ++i) {
    int l = len[i];
    code[i] = l ? (uint16_t)bitrev(next_code[l]++, l) : 0;
  }
}

// Insertion sort of the used symbols of a small alphabet into sorted_key; returns m.
BITAR_HD_NOINLINE int sort_used_small(const uint32_t* freq, int n, uint32_t* sorted_key) {
  int m = 0;
  for (int i = 0; i < n; ++i) {
    if (!freq[i]) continue;
    uint32_t key = (freq[i] << 9) | (uint32_t)i;
    int j = m++;
    while (j > 0 && sorted_key[j - 1] > key) {
      sorted_key[j] = sorted_key[j - 1];
      --j;
    }
    sorted_key[j] = key;
  }
  return m;
}

// ---- dynamic block header --------------------------------------------------------------------------
// RLE of a code-length sequence into code-length-alphabet tokens: tok = sym | (extra_val << 5).
// Runs never cross the litlen / dist boundary (as in zlib's scan_tree, called once per tree).
BITAR_HD_NOINLINE int cl_rle(const uint8_t* len, int n, uint16_t* tok, int ntok, uint32_t* cl_freq) {
  int i = 0;
  while (i < n) {
    int v = len[i], run = 1;
    while (i + run < n && len[i + run] == v) run++;
    i += run;
    if (v == 0) {
      while (run >= 11) {
        int r = run > 138 ? 138 : run;
        tok[ntok++] = (uint16_t)(18 | ((r - 11) << 5));
        cl_freq[18]++;
        run -= r;
      }
      if (run >= 3) {
        tok[ntok++] = (uint16_t)(17 | ((run - 3) << 5));
        cl_freq[17]++;
        run = 0;
      }
      while (run-- > 0) {
        tok[ntok++] = 0;
        cl_freq[0]++;
      }
    } else {
      tok[ntok++] = (uint16_t)v;
      cl_freq[v]++;
      run--;
      while (run >= 3) {
        int r = run > 6 ? 6 : run;
        tok[ntok++] = (uint16_t)(16 | ((r - 3) << 5));
        cl_freq[16]++;
        run -= r;
      }
      while (run-- > 0) {
        tok[ntok++] = (uint16_t)v;
        cl_freq[v]++;
      }
    }
  }
  return ntok;
}

BITAR_HD int cl_extra_bits(int sym) { return sym == 16 ? 2 : sym == 17 ? 3 : sym == 18 ? 7 : 0; }

// Everything a CTA needs to emit one dynamic block: produced by build_dynamic_plan().
struct BlockPlan {
  uint8_t ll_len[288];
  uint8_t d_len[32];
  uint16_t ll_code[288];
  uint16_t d_code[32];
  uint8_t cl_len[kNumCl];
  uint16_t cl_code[kNumCl];
  uint16_t cl_tok[kNumLitLen + kNumDist];
  int n_cl_tok;
  int hlit, hdist, hclen;       // counts (not the minus-offset wire values)
  uint32_t header_bits;         // 3 (BFINAL/BTYPE) + tables
  uint64_t dyn_body_bits;       // sum over symbols incl. extra bits and EOB
  uint64_t fixed_body_bits;     // same symbols under the fixed code (+3 header bits not included)
};

struct PlanScratch {
  HuffScratch hs;
  uint32_t sorted[288];
  uint32_t cl_freq[kNumCl];
  uint16_t bl_count[kMaxBits + 1];
};

// Last part of the plan: the code-length code and the header size, from the RLE tokens (p->cl_tok,
// p->n_cl_tok) and their frequencies (s->cl_freq).  Serial (19 symbols).
BITAR_HD_NOINLINE void plan_cl_tree(BlockPlan* p, PlanScratch* s) {
  int cm = sort_used_small(s->cl_freq, kNumCl, s->sorted);
  for (int i = 0; i < kNumCl; ++i) p->cl_len[i] = 0;
  if (cm == 1) {  // complete the code: give a second symbol a 1-bit code too
    int only = (int)(s->sorted[0] & 511);
    int other = only == 0 ? 1 : 0;
    s->cl_freq[other] = 1;
    cm = sort_used_small(s->cl_freq, kNumCl, s->sorted);
    s->cl_freq[other] = 0;
  }
  huff_lengths_from_sorted(s->sorted, cm, kMaxClBits, p->cl_len, s->bl_count, &s->hs);
  huff_codes(p->cl_len, kNumCl, s->bl_count, p->cl_code);
  int hclen = kNumCl;
  while (hclen > 4 && p->cl_len[cl_order(hclen - 1)] == 0) hclen--;
  p->hclen = hclen;

  uint32_t hb = 3 + 5 + 5 + 4 + 3 * (uint32_t)hclen;
  for (int i = 0; i < kNumCl; ++i) hb += s->cl_freq[i] * (uint32_t)(p->cl_len[i] + cl_extra_bits(i));
  p->header_bits = hb;
}

// Second half of the plan: code-length RLE, the code-length code and the header size, from the
// litlen / distance code lengths already in p (ll_len, d_len, hlit, hdist).  Serial; the deflate kernel runs
// the RLE with one run per thread and shares plan_cl_tree().
BITAR_HD_NOINLINE void plan_header(BlockPlan* p, PlanScratch* s) {
  for (int i = 0; i < kNumCl; ++i) s->cl_freq[i] = 0;
  int nt = cl_rle(p->ll_len, p->hlit, p->cl_tok, 0, s->cl_freq);
  nt = cl_rle(p->d_len, p->hdist, p->cl_tok, nt, s->cl_freq);
  p->n_cl_tok = nt;
  plan_cl_tree(p, s);
}

// ll_sorted / d_sorted: used symbols sorted ascending by (freq<<9 | sym) -- the caller sorts (the
// kernel does it with a CTA-wide bitonic sort; the host model with std::sort).  ll_freq[256] must
// already include the end-of-block count.  As in zlib (build_tree), a tree with fewer than two used
// symbols gets dummy symbols of frequency 1 so that the emitted code is always complete.
// This is the sequential statement of the plan; the deflate kernel computes the same first half (code
// lengths, codes, hlit / hdist, body sizes) with all its threads and shares plan_header().
BITAR_HD_NOINLINE void build_dynamic_plan(const uint32_t* ll_freq, const uint32_t* d_freq,
                                          const uint32_t* ll_sorted, int ll_m,
                                          const uint32_t* d_sorted, int d_m, BlockPlan* p,
                                          PlanScratch* s) {
  for (int i = 0; i < 288; ++i) p->ll_len[i] = 0;
  for (int i = 0; i < 32; ++i) p->d_len[i] = 0;
  huff_lengths_from_sorted(ll_sorted, ll_m, kMaxBits, p->ll_len, s->bl_count, &s->hs);
  huff_codes(p->ll_len, kNumLitLen, s->bl_count, p->ll_code);
  huff_lengths_from_sorted(d_sorted, d_m, kMaxBits, p->d_len, s->bl_count, &s->hs);
  huff_codes(p->d_len, kNumDist, s->bl_count, p->d_code);

  int hlit = kNumLitLen;
  while (hlit > 257 && p->ll_len[hlit - 1] == 0) hlit--;
  int hdist = kNumDist;
  while (hdist > 1 && p->d_len[hdist - 1] == 0) hdist--;
  p->hlit = hlit;
  p->hdist = hdist;

  plan_header(p, s);

  uint64_t dyn = 0, fix = 0;
  for (int i = 0; i < kNumLitLen; ++i) {
    uint32_t f = ll_freq[i];
    if (!f) continue;
    int eb = i > 256 ? len_extra_bits(i - 257) : 0;
    dyn += (uint64_t)f * (uint32_t)(p->ll_len[i] + eb);
    fix += (uint64_t)f * (uint32_t)(fixed_ll_len(i) + eb);
  }
  for (int i = 0; i < kNumDist; ++i) {
    uint32_t f = d_freq[i];
    if (!f) continue;
    int eb = dist_extra_bits(i);
    dyn += (uint64_t)f * (uint32_t)(p->d_len[i] + eb);
    fix += (uint64_t)f * (uint32_t)(5 + eb);
  }
  p->dyn_body_bits = dyn;
  p->fixed_body_bits = fix;
}

// Token word written by the match finder, one per input position:
//   0                      position is covered by a previous match (emits nothing)
//   1                      literal (the byte is read from the input itself)
//   (dist << 9) | len      match, len 3..258, dist 1..32768
BITAR_HD uint32_t tok_match(int len, int dist) { return ((uint32_t)dist << 9) | (uint32_t)len; }
BITAR_HD int tok_len(uint32_t t) { return (int)(t & 511u); }
BITAR_HD int tok_dist(uint32_t t) { return (int)(t >> 9); }

// Hashes of the 4 input bytes at a position (little-endian word w): one multiply serves both tables of the match finder.
// near: bucket of the per-sub-range 2-way table; far: slot of the block-wide "last position before this sub-range" table,
// whose size follows the block length n (2^9 .. 2^13 entries) so that small chunks need little shared memory.
BITAR_HD uint32_t hash_mul(uint32_t w) { return w * 0x9E3779B1u; }
BITAR_HD uint32_t hash_near(uint32_t w, int near_bits) { return hash_mul(w) >> (32 - near_bits); }
BITAR_HD uint32_t hash_far(uint32_t w, int far_bits) { return hash_mul(w) >> (32 - far_bits); }
BITAR_HD int far_hash_bits(uint32_t n) {
  const int b = n > 1u ? ilog2(n - 1u) - 2 : 0;
  return b < 9 ? 9 : b > 13 ? 13 : b;
}

enum BlockType { kStored = 0, kFixed = 1, kDynamic = 2 };

// ---- parallel-inflate index ------------------------------------------------------------------------
// The compressor never lets a TOKEN straddle the boundary of a 2 KiB SUB-RANGE (a match starts and ends inside the
// sub-range of its position; its SOURCE may lie anywhere in the preceding 32 KiB of the same 64 KiB block), so the
// symbols of every sub-range can be Huffman-decoded on their own once the decoder knows (a) the block's Huffman code and
// (b) the bit offset of the sub-range's first symbol.  (b) is appended to the chunk AFTER the final block's end-of-block
// symbol and byte padding: a stock RFC 1951 decoder stops at the final block and never looks at it (zlib returns
// Z_STREAM_END with the trailer left in avail_in).  The sm_100a inflate path finds it from the end of the buffer, decodes
// the symbols of 32 sub-ranges per warp into tokens (inflate_tok_kernel.cuh) and then resolves each block's back-references
// in stream order (inflate_resolve_kernel.cuh).  All values are little-endian u32, bit offsets count from the first bit
// of the chunk's stream:
//     for each 64 KiB block b:  hdr_bit[b], then sub_bit[b][s] for each sub-range s (0 for stored blocks)
//     end_bit      bit offset just past the final end-of-block symbol
//     total_out    uncompressed bytes of the chunk
//     kIndexMagic
// The index is omitted when no Huffman-coded block is longer than one sub-range, when it does not fit the output slot,
// or when the device was opened with emit_index = 0 (the chunk is then bare RFC 1951 and inflates through the
// whole-stream kernel).
constexpr int kSubLog2 = 11;
constexpr uint32_t kSub = 1u << kSubLog2;
constexpr int kIdxBlockLog2 = 16;                  // == the compressor's sub-block size (kBlockMax)
constexpr uint32_t kIndexMagic = 0xB17A0B02u;      // "bitar", sub-range log2, version (2: match sources leave the sub-range)
BITAR_HD uint32_t idx_blocks(uint32_t total) { return (total + 65535u) >> kIdxBlockLog2; }
BITAR_HD uint32_t idx_subs(uint32_t block_len) { return (block_len + kSub - 1u) >> kSubLog2; }
BITAR_HD uint32_t idx_entry(uint32_t block, uint32_t sub) { return block * 33u + 1u + sub; }   // hdr_bit at block * 33
BITAR_HD uint32_t idx_entries(uint32_t total) { return (total >> kIdxBlockLog2) * 33u + ((total & 65535u) ? 1u + idx_subs(total & 65535u) : 0u); }
BITAR_HD uint32_t idx_bytes(uint32_t total) { return 4u * (idx_entries(total) + 3u); }

}  // namespace dfl
}  // namespace bitar
