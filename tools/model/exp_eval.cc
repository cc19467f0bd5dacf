// exp_eval -- SCRATCH experiment driver for match-finder rules (round 2).  Not product, not a test.
// usage: exp_eval FILE SEG [key=val ...]
#include <zlib.h>

#include <cstdio>
#include <cstdlib>
#include <string>

#include "deflate_model.h"

using namespace bitar_model;

struct XP {
  int far_mode = 0;      // 0 off; 1 zlib-like block-wide chain (upper bound); 2 per-sub-range last-occurrence tables
  int chain = 4;         // far_mode 1: chain depth
  int fbits = 10;        // far_mode 2: table bits per sub-range
  int tagbits = 5;       // far_mode 2: tag bits (0 = none)
  int walk = 16;         // far_mode 2: sub-ranges walked back
  int margin = -1;       // diagonal rule margin (-1 = off)
  int far_min = 4;       // min length of a far match
  int need = 0;          // only search far when near best < need (0 = always)
  int nearblock = 0;
  int ins = 0;          // 1: near table holds token starts only (inserted after the window's parse)
  int bonus = 0;        // a far match must beat the near one by more than this
  int far8 = 0;         // far table keyed by 8 bytes
  int ways = 1;         // near table ways (2: also the second most recent position)
  int nice = 8;         // stop trying older candidates once a match this long is found
  int zone = 0;         // far_mode 3: far sources only from the first `zone` bytes of a sub-range, far matches only after them
  int cbits = 12;        // far_mode 3: cumulative table bits
  int both = 0;         // far_mode 3: also look into T_{s-1} exact table     // 1: near table is block-wide sequential (no cage, no reset) -- ablation
};
static XP X;
static std::vector<uint32_t> g_dump;
static long g_lit_lazy = 0, g_lit_nocand = 0, g_lit_short = 0, g_lits = 0, g_matches = 0;

static inline uint32_t hash32(uint32_t w, int bits) { return (w * 0x9E3779B1u) >> (32 - bits); }

static void find_tokens_x(const uint8_t* d, int n, const Params& P, std::vector<uint32_t>& tok) {
  tok.assign((size_t)n, 0);
  const int SUB = 1 << P.sub_log2;
  const int nsub = (n + SUB - 1) / SUB;
  // far structures
  std::vector<int> zhead, zprev;   // mode 1
  std::vector<std::vector<uint32_t>> T;   // mode 2: per sub-range, entry = (pos+1) | tag << 20
  if (X.far_mode == 1) {
    zhead.assign(1 << 15, -1);
    zprev.assign((size_t)n, -1);
  }
  if (X.far_mode == 2) {
    T.assign(nsub, std::vector<uint32_t>((size_t)1 << X.fbits, 0u));
    for (int p = 0; p + 4 <= n; ++p) {
      uint32_t w = load32(d, n, p);
      uint32_t hh = hash32(w, X.fbits + X.tagbits);
      uint32_t slot = hh >> X.tagbits, tag = hh & ((1u << X.tagbits) - 1u);
      T[p >> P.sub_log2][slot] = (uint32_t)(p + 1) | (tag << 20);   // last occurrence wins
    }
  }
  std::vector<int> C;   // mode 3: last occurrence before the current sub-range start
  std::vector<int> Cpend;
  if (X.far_mode == 3) C.assign((size_t)1 << X.cbits, -1);
  int c_upto = 0;       // positions < c_upto are in C
  std::vector<uint32_t> head((size_t)1 << P.hash_bits, 0);
  std::vector<uint32_t> head2((size_t)1 << P.hash_bits, 0);
  std::vector<uint32_t> h(P.step), old(P.step), old2(P.step);
  std::vector<int> near(P.step), adv(P.step), mdist(P.step);
  std::vector<uint8_t> valid(P.step);
  int carry = 0;
  int zins = 0;   // mode 1: positions < zins are inserted
  for (int base = 0; base < n; base += P.step) {
    if (!X.nearblock && (base & (SUB - 1)) == 0) std::fill(head.begin(), head.end(), 0u), std::fill(head2.begin(), head2.end(), 0u);
    if (X.far_mode == 3 && (base & (SUB - 1)) == 0)
      for (; c_upto < base; ++c_upto)
        if (c_upto + 4 <= n && (!X.zone || (c_upto & (SUB - 1)) + 4 <= X.zone)) C[hash32(load32(d, n, c_upto), X.cbits)] = c_upto;
    for (int t = 0; t < P.step; ++t) {
      int p = base + t;
      valid[t] = p + 4 <= n;
      h[t] = valid[t] ? hash_word(load32(d, (size_t)n, (size_t)p), P.hash_bits, P.min_match) : 0;
      old[t] = valid[t] ? head[h[t]] : 0;
      old2[t] = valid[t] ? head2[h[t]] : 0;
      near[t] = -1;
    }
    for (int l = 1; l < 32 && l < P.step; ++l) {
      if (!valid[l]) continue;
      for (int k = l - 1; k >= 0; --k)
        if (valid[k] && h[k] == h[l]) {
          near[l] = base + k;
          break;
        }
    }
    if (!X.ins)
      for (int t = 0; t < P.step; ++t)
        if (valid[t]) { uint32_t v = (uint32_t)(base + t + 1); if (v > head[h[t]]) { head2[h[t]] = head[h[t]]; head[h[t]] = v; } }
    for (int t = 0; t < P.step; ++t) {
      int p = base + t;
      adv[t] = 1;
      mdist[t] = 0;
      if (!valid[t]) continue;
      int best = 0, bdist = 0;
      const int s = p >> P.sub_log2, r = p & (SUB - 1);
      auto consider_near = [&](int c) {
        if (c < 0 || p - c > kMaxDist) return;
        if (!X.nearblock && (c >> P.sub_log2) != s) return;
        int l = match_len(d, n, p, c, X.nearblock ? 0 : P.sub_log2);
        if (l > best) best = l, bdist = p - c;
      };
      int c_old = old[t] ? (int)old[t] - 1 : -1;
      if (X.ways == 2) {
        int c_old2 = old2[t] ? (int)old2[t] - 1 : -1;
        // second nearest in the window
        int near2 = -1;
        if (near[t] >= 0) for (int k = near[t] - base - 1; k >= 0; --k) if (valid[k] && h[k] == h[t]) { near2 = base + k; break; }
        int cands[4] = {near[t], near[t] >= 0 ? (near2 >= 0 ? near2 : c_old) : c_old, -1, -1};
        if (near[t] < 0) cands[0] = c_old, cands[1] = c_old2;
        consider_near(cands[0]);
        if (best < X.nice) consider_near(cands[1]);
      } else
      if (X.both == 0) consider_near(near[t] >= 0 ? near[t] : c_old);
      else { consider_near(near[t]); if (c_old != near[t]) consider_near(c_old); }
      if (X.far_mode == 1) {
        // insert positions < p lazily (sequential semantics)
        for (; zins < p; ++zins)
          if (zins + 4 <= n) {
            uint32_t hh = hash32(load32(d, n, zins), 15);
            zprev[zins] = zhead[hh];
            zhead[hh] = zins;
          }
        if (!X.need || best < X.need) {
          int c = zhead[hash32(load32(d, n, p), 15)];
          for (int k = 0; k < X.chain && c >= 0 && p - c <= kMaxDist; ++k, c = zprev[c]) {
            int l = match_len(d, n, p, c, 0);
            if (l > best && l >= X.far_min) best = l, bdist = p - c;
          }
        }
      } else if (X.far_mode == 3 && (!X.need || best < X.need)) {
        uint32_t w = load32(d, n, p);
        int c = C[hash32(w, X.cbits)];
        if (c >= 0 && p - c <= kMaxDist && load32(d, n, c) == w && (!X.zone || r >= X.zone)) {
          int l = match_len(d, n, p, c, P.sub_log2);
          if (X.zone) l = std::min(l, X.zone - (c & (SUB - 1)));
          if (l >= X.far_min && l > best + (best >= 3 ? X.bonus : 0)) best = l, bdist = p - c;
        }
      } else if (X.far_mode == 2 && (!X.need || best < X.need)) {
        uint32_t w = load32(d, n, p);
        uint32_t hh = hash32(w, X.fbits + X.tagbits);
        uint32_t slot = hh >> X.tagbits, tag = hh & ((1u << X.tagbits) - 1u);
        for (int k = 1; k <= X.walk && s - k >= 0; ++k) {
          uint32_t e = T[s - k][slot];
          if (!e || (e >> 20) != tag) continue;
          int c = (int)(e & 0xFFFFF) - 1;
          if (p - c > kMaxDist) break;
          if (load32(d, n, c) != w) continue;
          int l = match_len(d, n, p, c, 0);
          if (X.margin >= 0) {
            int rp = c & (SUB - 1);
            // source must end at relative offset <= r - margin, and stay inside its sub-range
            int lim = std::min(r - X.margin - rp, SUB - rp);
            if (lim < l) l = lim;
          }
          if (l >= X.far_min && l > best) {
            best = l, bdist = p - c;
          }
          if (l >= X.far_min) break;   // first usable far candidate
        }
      }
      if (best >= kMinMatch && !(best == 3 && P.far3 && bdist > P.far3)) {
        adv[t] = best;
        mdist[t] = bdist;
      }
    }
    std::vector<int> why(P.step, 0);
    for (int t = 0; t < P.step; ++t) if (adv[t] == 1) why[t] = (near[t] < 0 && !old[t]) ? 1 : 2;
    if (P.lazy)
      for (int t = 0; t + 1 < P.step; ++t)
        if (adv[t] > 1 && adv[t + 1] > adv[t] + (P.lazy - 1)) adv[t] = 1, mdist[t] = 0, why[t] = 3;
    int lim = std::min(n, base + P.step);
    int pos = carry;
    while (pos < lim) {
      int t = pos - base;
      tok[(size_t)pos] = adv[t] > 1 ? tok_match(adv[t], mdist[t]) : 1u;
      if (adv[t] > 1) g_dump.push_back(((uint32_t)adv[t] << 16) | (uint32_t)mdist[t]);
      if (adv[t] > 1) g_matches++; else { g_lits++; if (why[t] == 1) g_lit_nocand++; else if (why[t] == 2) g_lit_short++; else g_lit_lazy++; }
      if (X.ins && valid[t]) head[h[t]] = std::max(head[h[t]], (uint32_t)(pos + 1));
      if (X.ins == 2 && adv[t] > 1 && t + 1 < P.step && valid[t + 1]) head[h[t + 1]] = std::max(head[h[t + 1]], (uint32_t)(pos + 2));
      pos += adv[t];
    }
    carry = pos;
  }
}

static std::vector<uint8_t> zdeflate(const uint8_t* d, size_t n, int level) {
  z_stream zs{};
  deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY);
  std::vector<uint8_t> out(n + n / 8 + 256);
  zs.next_in = (Bytef*)d;
  zs.avail_in = (uInt)n;
  zs.next_out = out.data();
  zs.avail_out = (uInt)out.size();
  deflate(&zs, Z_FINISH);
  out.resize(out.size() - zs.avail_out);
  deflateEnd(&zs);
  return out;
}
static bool zcheck(const std::vector<uint8_t>& c, const uint8_t* d, size_t n) {
  z_stream zs{};
  inflateInit2(&zs, -15);
  std::vector<uint8_t> out(n + 16);
  zs.next_in = (Bytef*)c.data();
  zs.avail_in = (uInt)c.size();
  zs.next_out = out.data();
  zs.avail_out = (uInt)out.size();
  int rc = inflate(&zs, Z_FINISH);
  size_t got = out.size() - zs.avail_out;
  inflateEnd(&zs);
  return rc == Z_STREAM_END && got == n && memcmp(out.data(), d, n) == 0;
}

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return perror("open"), 1;
  std::vector<uint8_t> data;
  uint8_t buf[1 << 16];
  size_t r;
  while ((r = fread(buf, 1, sizeof buf, f)) > 0) data.insert(data.end(), buf, buf + r);
  fclose(f);
  size_t seg = (size_t)atol(argv[2]);
  Params P;
  for (int i = 3; i < argc; ++i) {
    std::string kv = argv[i];
    auto eq = kv.find('=');
    std::string k = kv.substr(0, eq);
    int v = atoi(kv.substr(eq + 1).c_str());
    if (k == "hash_bits") P.hash_bits = v;
    else if (k == "lazy") P.lazy = v;
    else if (k == "far3") P.far3 = v;
    else if (k == "sub_log2") P.sub_log2 = v;
    else if (k == "far_mode") X.far_mode = v;
    else if (k == "chain") X.chain = v;
    else if (k == "fbits") X.fbits = v;
    else if (k == "tagbits") X.tagbits = v;
    else if (k == "walk") X.walk = v;
    else if (k == "margin") X.margin = v;
    else if (k == "far_min") X.far_min = v;
    else if (k == "need") X.need = v;
    else if (k == "nearblock") X.nearblock = v;
    else if (k == "cbits") X.cbits = v;
    else if (k == "both") X.both = v;
    else if (k == "zone") X.zone = v;
    else if (k == "ways") X.ways = v;
    else if (k == "nice") X.nice = v;
    else if (k == "ins") X.ins = v;
    else if (k == "bonus") X.bonus = v;
    else return fprintf(stderr, "unknown key %s\n", k.c_str()), 2;
  }
  uint64_t zsum = 0, msum = 0;
  for (size_t off = 0; off < data.size(); off += seg) {
    size_t n = std::min(seg, data.size() - off);
    auto z = zdeflate(data.data() + off, n, 1);
    BitWriter bw;
    std::vector<uint32_t> tok, index;
    for (size_t o2 = 0; o2 < n; o2 += 65536) {
      int len = (int)std::min((size_t)65536, n - o2);
      find_tokens_x(data.data() + off + o2, len, P, tok);
      encode_block(data.data() + off + o2, len, tok, o2 + len == n, P, bw, nullptr, &index);
    }
    bw.align();
    for (size_t k = 0; k < index.size() + 3; ++k) bw.put(0, 32);
    if (!zcheck(bw.out, data.data() + off, n)) return fprintf(stderr, "INVALID at chunk %zu\n", off / seg), 1;
    zsum += z.size();
    msum += bw.out.size();
  }
  if (getenv("DUMP")) { FILE* g = fopen(getenv("DUMP"), "wb"); fwrite(g_dump.data(), 4, g_dump.size(), g); fclose(g); }
  if (getenv("STATS")) printf("lits %ld (nocand %ld short/collision %ld lazy %ld) matches %ld\n", g_lits, g_lit_nocand, g_lit_short, g_lit_lazy, g_matches);
  printf("%-28s zlib1 %.3f model %.3f model/zlib %.4f\n", strrchr(argv[1], '/') + 1, (double)data.size() / zsum,
         (double)data.size() / msum, (double)msum / zsum);
  return 0;
}
