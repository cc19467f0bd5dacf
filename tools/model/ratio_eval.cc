// ratio_eval -- tune the GPU match finder's parameters on the CPU model and compare the compressed
// size with zlib level 1 (the reference's level, /root/reference/src/config.cc:87) chunk by chunk.
// Every model stream is verified by inflating it with zlib.   usage: ratio_eval FILE SEG [key=val..]
#include <zlib.h>

#include <cstdio>
#include <cstdlib>
#include <string>

#include "deflate_model.h"

static std::vector<uint8_t> zdeflate(const uint8_t* d, size_t n, int level, int strategy) {
  z_stream zs{};
  deflateInit2(&zs, level, Z_DEFLATED, -15, 8, strategy);
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
  return rc == Z_STREAM_END && got == n && memcmp(out.data(), d, n) == 0;   // the index trailer stays in avail_in
}

int main(int argc, char** argv) {
  if (argc < 3) return fprintf(stderr, "usage: %s FILE SEG [key=val ...]\n", argv[0]), 2;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return perror("open"), 1;
  std::vector<uint8_t> data;
  uint8_t buf[1 << 16];
  size_t r;
  while ((r = fread(buf, 1, sizeof buf, f)) > 0) data.insert(data.end(), buf, buf + r);
  fclose(f);
  size_t seg = (size_t)atol(argv[2]);
  bitar_model::Params P;
  for (int i = 3; i < argc; ++i) {
    std::string kv = argv[i];
    auto eq = kv.find('=');
    std::string k = kv.substr(0, eq);
    int v = atoi(kv.substr(eq + 1).c_str());
    if (k == "near_bits") P.near_bits = v;
    else if (k == "far_need") P.far_need = v;
    else if (k == "far_min") P.far_min = v;
    else if (k == "max_dist") P.max_dist = v;
    else if (k == "huffman") P.huffman = v;
    else if (k == "block") P.block = v;
    else if (k == "lazy") P.lazy = v;
    else if (k == "select") P.select = v;
    else if (k == "pair") P.pair = v;
    else if (k == "sub_log2") P.sub_log2 = v;
    else return fprintf(stderr, "unknown key %s\n", k.c_str()), 2;
  }
  uint64_t g_far = 0, g_deferred = 0, g_all = 0, g_dirty_bytes = 0, g_sub_dirty = 0, g_subs = 0;
  uint64_t zsum = 0, msum = 0, lits = 0, matches = 0, mbytes = 0, fr[6] = {0, 0, 0, 0, 0, 0};
  int types[3] = {0, 0, 0};
  for (size_t off = 0; off < data.size(); off += seg) {
    size_t n = std::min(seg, data.size() - off);
    auto z = zdeflate(data.data() + off, n, 1, P.huffman == 1 ? Z_FIXED : Z_DEFAULT_STRATEGY);
    std::vector<bitar_model::BlockStats> st;
    auto m = bitar_model::deflate_chunk(data.data() + off, n, P, &st);
    {   // how much a decoder that resolves sub-range-local matches itself would have to defer (experiment)
      std::vector<uint32_t> tok;
      for (size_t o2 = 0; o2 < n; o2 += 65536) {
        int bl = (int)std::min((size_t)65536, n - o2);
        bitar_model::find_tokens(data.data() + off + o2, bl, P, tok);
        std::vector<uint8_t> dirty((size_t)bl, 0);
        for (int s0 = 0; s0 < bl; s0 += 2048) {
          bool any = false;
          for (int p = s0; p < std::min(bl, s0 + 2048); ++p) {
            uint32_t t = tok[(size_t)p];
            if (t <= 1) continue;
            int len = bitar_model::tok_len(t), dist = bitar_model::tok_dist(t);
            bool far = p - dist < s0, d = far;
            for (int k = 0; k < len && !d; ++k) d = dirty[(size_t)(p - dist + k)];
            g_all++; g_far += far; g_deferred += d;
            if (d) { for (int k = 0; k < len; ++k) dirty[(size_t)(p + k)] = 1; g_dirty_bytes += len; any = true; }
          }
          g_subs++; g_sub_dirty += any;
        }
      }
    }
    if (!zcheck(m, data.data() + off, n)) {
      if (getenv("DUMP_BAD")) { FILE* g = fopen(getenv("DUMP_BAD"), "wb"); fwrite(m.data(), 1, m.size(), g); fclose(g); }
      return fprintf(stderr, "MODEL STREAM INVALID at chunk %zu\n", off / seg), 1;
    }
    zsum += z.size();
    msum += m.size();
    for (auto& s : st) {
      lits += s.literals;
      matches += s.matches;
      mbytes += s.match_bytes;
      fr[0] += s.far128; fr[1] += s.far128_bytes; fr[2] += s.far256; fr[3] += s.far256_bytes; fr[4] += s.far512; fr[5] += s.far512_bytes;
      types[s.type]++;
    }
  }
  printf("bytes=%zu zlib1=%llu (%.3f) model=%llu (%.3f) model/zlib=%.4f  lits=%llu matches=%llu avg_mlen=%.1f B/sym=%.2f blocks s/f/d=%d/%d/%d\n",
         data.size(), (unsigned long long)zsum, (double)data.size() / zsum, (unsigned long long)msum,
         (double)data.size() / msum, (double)msum / zsum, (unsigned long long)lits,
         (unsigned long long)matches, matches ? (double)mbytes / matches : 0.0,
         (double)data.size() / (double)(lits + matches), types[0], types[1], types[2]);
  printf("windows: %ld total, %.1f%% skipped (covered), %.1f%% without any 3-byte candidate, %.1f%% with at most one per lane\n", bitar_model::g_win_total, 100.0 * bitar_model::g_win_skipped / bitar_model::g_win_total, 100.0 * bitar_model::g_win_nocand / bitar_model::g_win_total, 100.0 * bitar_model::g_win_1cand / bitar_model::g_win_total);
  printf("far matches %.1f%%, deferred (far or reading deferred bytes) %.1f%% of matches, %.1f%% of bytes; sub-ranges with any: %.1f%%\n", 100.0 * g_far / g_all,
         100.0 * g_deferred / g_all, 100.0 * g_dirty_bytes / data.size(), 100.0 * g_sub_dirty / g_subs);
  printf("matches farther than 128 / 256 / 512 B: %.1f%% / %.1f%% / %.1f%% of matches, %.1f%% / %.1f%% / %.1f%% of all bytes\n", 100.0 * fr[0] / matches, 100.0 * fr[2] / matches,
         100.0 * fr[4] / matches, 100.0 * fr[1] / data.size(), 100.0 * fr[3] / data.size(), 100.0 * fr[5] / data.size());
  return 0;
}
