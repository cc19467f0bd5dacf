// deflate_model.h -- sequential CPU model of the GPU deflate kernel (TEST / TUNING TOOL).
//
// The sm_100a kernel in bitar_b200/csrc/deflate_kernel.cuh is deterministic by construction
// (CTA-synchronous match-finding steps, atomicMax hash inserts, order-preserving parse), so its
// output is a pure function of (input bytes, parameters).  This file states that function in plain
// sequential C++, sharing the Huffman / header code with the kernel through deflate_common.h.  It is
// used to (a) tune the compression ratio against zlib level 1 without a GPU and (b) pin the kernel
// bit-for-bit in tests.  It is not part of the product and not the parity oracle (that is
// oracle/): streams produced here are themselves checked with zlib and oracle/rfc1951.c.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../bitar_b200/csrc/deflate_common.h"

namespace bitar_model {

using namespace bitar::dfl;

struct Params {
  int step = 32;        // positions per dictionary step: one warp window (exact nearest-previous semantics)
  int hash_bits = 10;   // one 1024-entry table per 2 KiB sub-range (the kernel keeps one per warp)
  int min_match = 4;    // bytes hashed (3 or 4); emitted matches are always >= 3
  int cand_mode = 1;    // 0 table only, 1 warp-near else table, 2 best of both
  int far3 = 4096;      // reject length-3 matches farther than this (0 = keep all)
  int huffman = 2;      // 1 fixed only, 2 dynamic (min of stored/fixed/dynamic)
  int block = 65536;    // sub-block size inside a chunk (window restarts at sub-block start)
  int near_mode = 0;    // 0: nearest lower lane of the window with the same HASH; 1..: same 4-byte WORD at a distance of set #near_mode
  int lazy = 1;         // 1: a match yields to a strictly longer match starting at the next position of its window
  int sub_log2 = kSubLog2;  // matches stay inside the 2^sub_log2-byte sub-range of their position and the
                        // parallel-inflate index is appended (deflate_common.h); 0 = off
};

struct BitWriter {
  std::vector<uint8_t> out;
  uint64_t acc = 0;
  int nbits = 0;
  void put(uint32_t v, int n) {
    acc |= (uint64_t)v << nbits;
    nbits += n;
    while (nbits >= 8) {
      out.push_back((uint8_t)acc);
      acc >>= 8;
      nbits -= 8;
    }
  }
  void align() {
    if (nbits) put(0, 8 - nbits);
  }
  uint64_t bits() const { return (uint64_t)out.size() * 8 + (uint64_t)nbits; }
};

inline uint32_t load32(const uint8_t* d, size_t n, size_t p) {  // zero padded past the end
  uint32_t w = 0;
  for (int i = 0; i < 4; ++i)
    if (p + i < n) w |= (uint32_t)d[p + i] << (8 * i);
  return w;
}

inline int match_len(const uint8_t* d, int n, int p, int c, int sub_log2 = 0) {
  int maxl = std::min(kMaxMatch, n - p), l = 0;
  if (sub_log2) maxl = std::min(maxl, (((p >> sub_log2) + 1) << sub_log2) - p);
  while (l < maxl && d[p + l] == d[c + l]) ++l;
  return l;
}

// Tokenise one sub-block d[0..n): tok[p] as in deflate_common.h.
inline void find_tokens(const uint8_t* d, int n, const Params& P, std::vector<uint32_t>& tok) {
  tok.assign((size_t)n, 0);
  std::vector<uint32_t> head((size_t)1 << P.hash_bits, 0);
  std::vector<uint32_t> h(P.step), old(P.step);
  std::vector<int> near(P.step), adv(P.step), mdist(P.step);
  std::vector<uint8_t> valid(P.step);
  int carry = 0;
  for (int base = 0; base < n; base += P.step) {
    if (P.sub_log2 && (base & ((1 << P.sub_log2) - 1)) == 0) std::fill(head.begin(), head.end(), 0u);   // per sub-range table
    for (int t = 0; t < P.step; ++t) {
      int p = base + t;
      valid[t] = p + 4 <= n;
      h[t] = valid[t] ? hash_word(load32(d, (size_t)n, (size_t)p), P.hash_bits, P.min_match) : 0;
      old[t] = valid[t] ? head[h[t]] : 0;
      near[t] = -1;
    }
    static const int kSets[4][16] = {{0}, {1, 2, 3, 4, 6, 8, 12, 16, 24, 0}, {1, 2, 3, 4, 5, 6, 7, 8, 10, 12, 16, 20, 24, 28, 0}, {1, 2, 4, 8, 16, 0}};
    for (int w0 = 0; w0 < P.step; w0 += 32)
      for (int l = 1; l < 32 && w0 + l < P.step; ++l) {
        if (!valid[w0 + l]) continue;
        if (P.near_mode == 0) {
          for (int k = l - 1; k >= 0; --k)
            if (valid[w0 + k] && h[w0 + k] == h[w0 + l]) {
              near[w0 + l] = base + w0 + k;
              break;
            }
        } else {
          for (const int* dd = kSets[P.near_mode]; *dd; ++dd)
            if (*dd <= l && load32(d, (size_t)n, (size_t)(base + w0 + l)) == load32(d, (size_t)n, (size_t)(base + w0 + l - *dd))) {
              near[w0 + l] = base + w0 + l - *dd;
              break;
            }
        }
      }
    for (int t = 0; t < P.step; ++t)
      if (valid[t]) head[h[t]] = std::max(head[h[t]], (uint32_t)(base + t + 1));
    for (int t = 0; t < P.step; ++t) {
      int p = base + t;
      adv[t] = 1;
      mdist[t] = 0;
      if (!valid[t]) continue;
      int best = 0, bdist = 0;
      auto consider = [&](int c) {
        if (c < 0 || p - c > kMaxDist) return;
        if (P.sub_log2 && (c >> P.sub_log2) != (p >> P.sub_log2)) return;   // other sub-range
        int l = match_len(d, n, p, c, P.sub_log2);
        if (l > best) {  // first considered wins ties
          best = l;
          bdist = p - c;
        }
      };
      int c_old = old[t] ? (int)old[t] - 1 : -1;
      if (P.cand_mode == 0) consider(c_old);
      else if (P.cand_mode == 1) consider(near[t] >= 0 ? near[t] : c_old);
      else {
        consider(near[t]);
        if (c_old != near[t]) consider(c_old);
      }
      if (best >= kMinMatch && !(best == 3 && P.far3 && bdist > P.far3)) {
        adv[t] = best;
        mdist[t] = bdist;
      }
    }
    if (P.lazy)   // decided on the matches as found (not on already reduced neighbours): one pass, in place is fine going up
      for (int t = 0; t + 1 < P.step; ++t)
        if (adv[t] > 1 && adv[t + 1] > adv[t] + (P.lazy - 1)) adv[t] = 1, mdist[t] = 0;
    int lim = std::min(n, base + P.step);
    int pos = carry;
    while (pos < lim) {
      int t = pos - base;
      tok[(size_t)pos] = adv[t] > 1 ? tok_match(adv[t], mdist[t]) : 1u;
      pos += adv[t];
    }
    carry = pos;
  }
}

struct BlockStats {
  uint64_t literals = 0, matches = 0, match_bytes = 0;
  uint64_t far128 = 0, far128_bytes = 0, far256 = 0, far256_bytes = 0, far512 = 0, far512_bytes = 0;   // matches by distance (inflate ring sizes)
  int type = 0;
  uint64_t bits = 0;
};

inline void put_fixed_sym(BitWriter& bw, int sym) {
  int l = fixed_ll_len(sym);
  uint32_t code = sym < 144 ? 0x30 + sym : sym < 256 ? 0x190 + (sym - 144) : sym < 280 ? sym - 256 : 0xC0 + (sym - 280);
  bw.put(bitrev(code, l), l);
}

// Encode one sub-block (tokens already found) as the cheapest of stored / fixed / dynamic.
inline void encode_block(const uint8_t* d, int n, const std::vector<uint32_t>& tok, bool final_block,
                         const Params& P, BitWriter& bw, BlockStats* st, std::vector<uint32_t>* index = nullptr) {
  uint32_t ll_freq[288] = {0}, d_freq[32] = {0};
  for (int p = 0; p < n; ++p) {
    uint32_t t = tok[(size_t)p];
    if (t == 0) continue;
    if (t == 1) {
      ll_freq[d[p]]++;
      if (st) st->literals++;
    } else {
      ll_freq[257 + len_sym(tok_len(t))]++;
      d_freq[dist_sym(tok_dist(t))]++;
      if (st) {
        st->matches++;
        st->match_bytes += (uint64_t)tok_len(t);
        if (tok_dist(t) > 128) st->far128++, st->far128_bytes += (uint64_t)tok_len(t);
        if (tok_dist(t) > 256) st->far256++, st->far256_bytes += (uint64_t)tok_len(t);
        if (tok_dist(t) > 512) st->far512++, st->far512_bytes += (uint64_t)tok_len(t);
      }
    }
  }
  ll_freq[kEob] = 1;
  // zlib build_tree: force at least two used symbols per tree
  uint32_t llf[288], df[32];
  memcpy(llf, ll_freq, sizeof llf);
  memcpy(df, d_freq, sizeof df);
  auto force2 = [](uint32_t* f, int n_) {
    int used = 0;
    for (int i = 0; i < n_; ++i) used += f[i] != 0;
    for (int i = 0; used < 2 && i < n_; ++i)
      if (!f[i]) {
        f[i] = 1;
        used++;
      }
  };
  force2(llf, kNumLitLen);
  force2(df, kNumDist);
  uint32_t ll_sorted[288], d_sorted[32];
  int ll_m = 0, d_m = 0;
  for (int i = 0; i < kNumLitLen; ++i)
    if (llf[i]) ll_sorted[ll_m++] = (llf[i] << 9) | (uint32_t)i;
  for (int i = 0; i < kNumDist; ++i)
    if (df[i]) d_sorted[d_m++] = (df[i] << 9) | (uint32_t)i;
  std::sort(ll_sorted, ll_sorted + ll_m);
  std::sort(d_sorted, d_sorted + d_m);
  static thread_local BlockPlan plan;
  static thread_local PlanScratch scratch;
  // body costs are computed on the real frequencies (dummy symbols are never emitted)
  build_dynamic_plan(ll_freq, d_freq, ll_sorted, ll_m, d_sorted, d_m, &plan, &scratch);

  uint64_t dyn_bits = plan.header_bits + plan.dyn_body_bits;
  uint64_t fix_bits = 3 + plan.fixed_body_bits;
  // stored: align to byte, then 4 bytes LEN/NLEN + data per <=65535-byte piece
  int pieces = n == 0 ? 1 : (n + 65534) / 65535;
  uint64_t cur = bw.bits();
  uint64_t stored_bits = ((cur + 3 + 7) / 8 * 8 - cur) + 32 + (uint64_t)n * 8 + (uint64_t)(pieces - 1) * 40;
  int type;
  if (P.huffman == 1) type = fix_bits <= stored_bits ? kFixed : kStored;
  else {
    type = kDynamic;
    uint64_t best = dyn_bits;
    if (fix_bits <= best) {
      type = kFixed;
      best = fix_bits;
    }
    if (stored_bits < best) type = kStored;
  }
  if (st) st->type = type;
  uint64_t before = bw.bits();
  if (index) {   // hdr_bit, then one slot per sub-range (stays 0 for stored blocks)
    index->push_back((uint32_t)before);
    index->resize(index->size() + idx_subs((uint32_t)n), 0u);
  }
  const size_t idx_base = index ? index->size() - idx_subs((uint32_t)n) : 0;
  if (type == kStored) {
    int off = 0;
    for (int k = 0; k < pieces; ++k) {
      int len = std::min(65535, n - off);
      bool last = final_block && k == pieces - 1;
      bw.put(last ? 1 : 0, 1);
      bw.put(0, 2);
      bw.align();
      bw.put((uint32_t)len, 16);
      bw.put((uint32_t)(~len) & 0xffff, 16);
      for (int i = 0; i < len; ++i) bw.put(d[off + i], 8);
      off += len;
    }
  } else {
    bw.put(final_block ? 1 : 0, 1);
    bw.put((uint32_t)type, 2);
    if (type == kDynamic) {
      bw.put((uint32_t)(plan.hlit - 257), 5);
      bw.put((uint32_t)(plan.hdist - 1), 5);
      bw.put((uint32_t)(plan.hclen - 4), 4);
      for (int i = 0; i < plan.hclen; ++i) bw.put(plan.cl_len[cl_order(i)], 3);
      for (int i = 0; i < plan.n_cl_tok; ++i) {
        int sym = plan.cl_tok[i] & 31, ev = plan.cl_tok[i] >> 5;
        bw.put(plan.cl_code[sym], plan.cl_len[sym]);
        if (sym >= 16) bw.put((uint32_t)ev, cl_extra_bits(sym));
      }
    }
    for (int p = 0; p < n; ++p) {
      uint32_t t = tok[(size_t)p];
      if (index && (p & (int)(kSub - 1)) == 0) (*index)[idx_base + ((size_t)p >> kSubLog2)] = (uint32_t)bw.bits();
      if (t == 0) continue;
      if (t == 1) {
        if (type == kDynamic) bw.put(plan.ll_code[d[p]], plan.ll_len[d[p]]);
        else put_fixed_sym(bw, d[p]);
      } else {
        int len = tok_len(t), dist = tok_dist(t), ls = len_sym(len), ds = dist_sym(dist);
        if (type == kDynamic) bw.put(plan.ll_code[257 + ls], plan.ll_len[257 + ls]);
        else put_fixed_sym(bw, 257 + ls);
        bw.put((uint32_t)len_extra_val(len, ls), len_extra_bits(ls));
        if (type == kDynamic) bw.put(plan.d_code[ds], plan.d_len[ds]);
        else bw.put(bitrev((uint32_t)ds, 5), 5);
        bw.put((uint32_t)dist_extra_val(dist, ds), dist_extra_bits(ds));
      }
    }
    if (type == kDynamic) bw.put(plan.ll_code[kEob], plan.ll_len[kEob]);
    else put_fixed_sym(bw, kEob);
  }
  if (st) st->bits = bw.bits() - before;
}

// One chunk = one complete raw DEFLATE stream (BFINAL on the last sub-block).
inline std::vector<uint8_t> deflate_chunk(const uint8_t* d, size_t n, const Params& P,
                                          std::vector<BlockStats>* stats = nullptr) {
  BitWriter bw;
  if (n == 0) {  // fixed block holding only end-of-block: 03 00
    bw.put(1, 1);
    bw.put(1, 2);
    put_fixed_sym(bw, kEob);
    bw.align();
    return bw.out;
  }
  std::vector<uint32_t> tok, index;
  const bool want_index = P.sub_log2 == kSubLog2 && P.block == (1 << kIdxBlockLog2);
  bool any_coded = false;
  for (size_t off = 0; off < n; off += (size_t)P.block) {
    int len = (int)std::min((size_t)P.block, n - off);
    find_tokens(d + off, len, P, tok);
    BlockStats st;
    encode_block(d + off, len, tok, off + (size_t)len == n, P, bw, &st, want_index ? &index : nullptr);
    any_coded |= st.type != kStored && len > (int)kSub;   // a block the index can split
    if (stats) stats->push_back(st);
  }
  const uint32_t end_bit = (uint32_t)bw.bits();
  bw.align();
  if (want_index && any_coded) {   // the parallel-inflate index (deflate_common.h)
    index.push_back(end_bit);
    index.push_back((uint32_t)n);
    index.push_back(kIndexMagic);
    for (uint32_t w : index) bw.put(w, 32);
  }
  return bw.out;
}

}  // namespace bitar_model
