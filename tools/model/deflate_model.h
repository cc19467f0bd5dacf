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
  // The match finder of the sm_100a kernel (deflate_kernel.cuh), stated sequentially.  Per 64 KiB block:
  //   far pass   : a table of 2^far_bits(n) entries holds, for every hash of 4 bytes, the last position BEFORE the current
  //                2 KiB sub-range; every position gets the entry of its hash as its FAR candidate when the 4 bytes
  //                really are equal (the kernel runs this pass sub-range by sub-range with the whole CTA).
  //   near pass  : per sub-range a 2-way table of 2^near_bits buckets (the two most recent positions of a hash before
  //                the current window of 32 positions) and the positions of the window itself; the two most recent
  //                positions with the hash of p are candidates, the longer match wins (ties: the nearer).
  //   selection  : when the near match is shorter than far_need bytes, a far match of at least far_min bytes that is
  //                strictly longer replaces it.  A token never straddles a sub-range boundary (that is what lets the
  //                inflate kernel Huffman-decode 32 sub-ranges of a block in parallel, deflate_common.h).
  int near_bits = 9;
  int far_need = 8;
  int far_min = 4;
  int pair = 0;         // EXPERIMENT: the near tables survive from an even sub-range into the odd one after it
  int select = 0;       // EXPERIMENT: candidate selection variants (0 = the kernel's)
  int lazy = 1;         // 1: a match yields to a strictly longer match starting at the next position of its window
  int max_dist = kMaxDist;   // the configured window (1 << window_size)
  int huffman = 2;      // 1 fixed only, 2 dynamic (min of stored/fixed/dynamic)
  int block = 65536;    // sub-block size inside a chunk (window restarts at sub-block start)
  int sub_log2 = kSubLog2;
};

static long g_win_total = 0, g_win_nocand = 0, g_win_skipped = 0, g_win_1cand = 0;
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

inline int match_len(const uint8_t* d, int p, int c, int maxl) {
  int l = 0;
  while (l < maxl && d[p + l] == d[c + l]) ++l;
  return l;
}

// Tokenise one sub-block d[0..n): tok[p] as in deflate_common.h.
inline void find_tokens(const uint8_t* d, int n, const Params& P, std::vector<uint32_t>& tok) {
  tok.assign((size_t)n, 0);
  const int SUB = 1 << P.sub_log2;
  constexpr int kNone = -1;
  // ---- far pass ----
  std::vector<int> far((size_t)n, kNone);
  {
    const int fb = far_hash_bits((uint32_t)n);
    std::vector<int> C((size_t)1 << fb, kNone);
    for (int s0 = 0; s0 < n; s0 += SUB) {
      const int s1 = std::min(n, s0 + SUB);
      if (s0)
        for (int p = s0; p < s1 && p + 4 <= n; ++p) {
          const uint32_t w = load32(d, (size_t)n, (size_t)p);
          const int c = C[hash_far(w, fb)];
          if (c != kNone && p - c <= P.max_dist && load32(d, (size_t)n, (size_t)c) == w) far[(size_t)p] = c;
        }
      for (int p = s0; p < s1 && p + 4 <= n; ++p) C[hash_far(load32(d, (size_t)n, (size_t)p), fb)] = p;
    }
  }
  // ---- near pass, selection and parse, windows of 32 positions ----
  std::vector<int> head0((size_t)1 << P.near_bits), head1((size_t)1 << P.near_bits);
  int carry = 0;
  for (int base = 0; base < n; base += 32) {
    if ((base & (SUB - 1)) == 0) {
      if (!(P.pair && ((base / SUB) % P.pair) != 0)) {
        std::fill(head0.begin(), head0.end(), kNone);
        std::fill(head1.begin(), head1.end(), kNone);
      }
      carry = base;
    }
    const int sub_end = std::min(n, (base & ~(SUB - 1)) + SUB);
    uint32_t h[32];
    int adv[32], mdist[32];
    bool valid[32];
    int c1[32], c2[32];
    for (int t = 0; t < 32; ++t) {
      const int p = base + t;
      valid[t] = p + 4 <= n;
      h[t] = valid[t] ? hash_near(load32(d, (size_t)n, (size_t)p), P.near_bits) : 0;
      c1[t] = c2[t] = kNone;
      if (!valid[t]) continue;
      int k1 = -1, k2 = -1;   // the two nearest lower positions of the window with this hash
      for (int k = t - 1; k >= 0; --k)
        if (valid[k] && h[k] == h[t]) {
          if (k1 < 0) k1 = k;
          else {
            k2 = k;
            break;
          }
        }
      if (k1 >= 0) {
        c1[t] = base + k1;
        c2[t] = k2 >= 0 ? base + k2 : head0[h[t]];
      } else {
        c1[t] = head0[h[t]];
        c2[t] = head1[h[t]];
      }
    }
    for (int t = 0; t < 32; ++t)   // the table after this window: the two most recent positions of every hash
      if (valid[t]) {
        head1[h[t]] = head0[h[t]];
        head0[h[t]] = base + t;
      }
    {   // EXPERIMENT: windows in which no live lane has a candidate with 3 equal bytes
      g_win_total++;
      if (carry - base >= 32) g_win_skipped++;
      else {
        bool any = false, two = false;
        for (int t = 0; t < 32; ++t) {
          const int p = base + t;
          if (!valid[t] || p < carry) continue;
          const uint32_t wp = load32(d, (size_t)n, (size_t)p) & 0xFFFFFFu;
          int cnt = 0;
          for (int c : {c1[t], c2[t], far[(size_t)p]})
            if (c != kNone && p - c <= P.max_dist && (load32(d, (size_t)n, (size_t)c) & 0xFFFFFFu) == wp) cnt++;
          any |= cnt > 0;
          two |= cnt > 1;
        }
        if (!any) g_win_nocand++;
        else if (!two) g_win_1cand++;
      }
    }
    for (int t = 0; t < 32; ++t) {
      const int p = base + t;
      adv[t] = 1;
      mdist[t] = 0;
      if (!valid[t]) continue;
      const int maxl = std::min(kMaxMatch, sub_end - p);
      int best = 0, bc = kNone;
      if (P.select == 0) {
      if (c1[t] != kNone && p - c1[t] <= P.max_dist) best = match_len(d, p, c1[t], maxl), bc = c1[t];
      if (c2[t] != kNone && p - c2[t] <= P.max_dist) {
        const int l = match_len(d, p, c2[t], maxl);
        if (l > best) best = l, bc = c2[t];
      }
      if (best < P.far_need && far[(size_t)p] != kNone) {
        const int l = match_len(d, p, far[(size_t)p], maxl);
        if (l >= P.far_min && l > best) best = l, bc = far[(size_t)p];
      }
      } else {
        const uint32_t wp = load32(d, (size_t)n, (size_t)p);
        auto ok4 = [&](int c) { return c != kNone && p - c <= P.max_dist && load32(d, (size_t)n, (size_t)c) == wp; };
        int cands[3] = {c1[t], c2[t], far[(size_t)p]};
        if (P.select == 1) {          // first candidate whose 4 bytes are equal
          for (int k = 0; k < 3; ++k) if (ok4(cands[k])) { best = match_len(d, p, cands[k], maxl); bc = cands[k]; break; }
        } else if (P.select == 2) {   // c1 else c2 (first verified), then far when short
          for (int k = 0; k < 2; ++k) if (ok4(cands[k])) { best = match_len(d, p, cands[k], maxl); bc = cands[k]; break; }
          if (best < P.far_need && far[(size_t)p] != kNone) {
            const int l = match_len(d, p, far[(size_t)p], maxl);
            if (l >= P.far_min && l > best) best = l, bc = far[(size_t)p];
          }
        } else if (P.select == 3) {   // longer of c1 / c2 judged on 8 bytes, winner extended; far when short
          int l1 = ok4(cands[0]) ? std::min(8, match_len(d, p, cands[0], maxl)) : 0;
          int l2 = ok4(cands[1]) ? std::min(8, match_len(d, p, cands[1], maxl)) : 0;
          int w = l2 > l1 ? 1 : 0;
          if (std::max(l1, l2) >= 4) { best = match_len(d, p, cands[w], maxl); bc = cands[w]; }
          if (best < P.far_need && far[(size_t)p] != kNone) {
            const int l = match_len(d, p, far[(size_t)p], maxl);
            if (l >= P.far_min && l > best) best = l, bc = far[(size_t)p];
          }
        } else if (P.select == 4) {   // c1 only, then far when short
          if (ok4(cands[0])) { best = match_len(d, p, cands[0], maxl); bc = cands[0]; }
          if (best < P.far_need && far[(size_t)p] != kNone) {
            const int l = match_len(d, p, far[(size_t)p], maxl);
            if (l >= P.far_min && l > best) best = l, bc = far[(size_t)p];
          }
        }
      }
      if (best >= kMinMatch) {
        adv[t] = best;
        mdist[t] = p - bc;
      }
    }
    if (P.lazy)   // decided on the matches as found (not on already reduced neighbours): one pass, in place is fine going up
      for (int t = 0; t + 1 < 32; ++t)
        if (adv[t] > 1 && adv[t + 1] > adv[t]) adv[t] = 1, mdist[t] = 0;
    const int lim = std::min(n, base + 32);
    int pos = carry;
    while (pos < lim) {
      const int t = pos - base;
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
