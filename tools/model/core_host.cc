// core_host.cc -- builds the kernels' host/device-shared code for the CPU (TEST TOOL).
//   * host_inflate_chunk : bitar_b200/csrc/inflate_core.h instantiated with G = 1
//   * model_deflate_chunk: sequential model of the deflate kernel (deflate_model.h)
// Loaded through ctypes by tests/; never part of the product library.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../bitar_b200/csrc/inflate_core.h"
#include "../../bitar_b200/csrc/inflate_fast.h"
#include "../../bitar_b200/csrc/inflate_tok.h"
#include "../../bitar_b200/csrc/inflate_spec.h"
#include "deflate_model.h"

#define API extern "C" __attribute__((visibility("default")))

API int host_inflate_chunk(const uint8_t* in, uint32_t in_len, uint8_t* out, uint32_t cap,
                           uint32_t* result4, int lbits) {
  using namespace bitar::inf;
  Group<1> g{0, 1u};
  ChunkResult r;
  if (lbits == 9) {
    static thread_local GroupSmem<9, 7, 1024> sm;
    r = inflate_chunk<1, 9, 7, 1024>(in, in_len, out, cap, &sm, g);
  } else {
    static thread_local GroupSmem<10, 8, 1024> sm;
    r = inflate_chunk<1, 10, 8, 1024>(in, in_len, out, cap, &sm, g);
  }
  result4[0] = r.produced;
  result4[1] = r.status;
  result4[2] = r.consumed;
  result4[3] = r.blocks;
  return 0;
}

// the lane-per-stream decoder of the production inflate kernel (inflate_fast.h), one lane on the CPU
template <int LB, int LT, int DB, int DT, int RG>
static void run_fast(const uint8_t* in, uint32_t in_len, uint8_t* out, uint32_t cap, uint32_t* result6, int checksum_type) {
  using namespace bitar::fl;
  using L = FastLane<LB, LT, DB, DT, RG>;
  alignas(16) static thread_local uint8_t smem[LaneLayout<LB, LT, DB, DT, RG>::kStride];
  static thread_local LaneScratch scratch;
  static CtaTables cta;
  static bool init = false;
  if (!init) {
    for (int i = 0; i < 32; ++i) cta.dinfo[i] = dist_info(i);
    for (uint32_t i = 0; i < 256; ++i) cta.crc[0][i] = bitar::cks::crc_table_entry(i);
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = cta.crc[0][i];
      for (int k = 1; k < 4; ++k) {
        c = (c >> 8) ^ cta.crc[0][c & 0xFFu];
        cta.crc[k][i] = c;
      }
    }
    init = true;
  }
  L lane;
  lane.bind(smem, &cta, &scratch, (uint32_t)checksum_type);
  lane.start(in, in_len, out, cap);
  uint64_t steps = 0;
  while (lane.state != L::kDone) {
    lane.step();
    if (++steps > (1ull << 32)) break;
  }
  result6[0] = lane.produced();
  result6[1] = lane.status;
  result6[2] = lane.consumed_bytes();
  result6[3] = lane.blocks;
  const uint64_t ck = lane.checksum();
  result6[4] = (uint32_t)ck;
  result6[5] = (uint32_t)(ck >> 32);
}

API int host_inflate_fast(const uint8_t* in, uint32_t in_len, uint8_t* out, uint32_t cap, uint32_t* result6, int lbits,
                          int checksum_type) {
  if (lbits == 9) run_fast<9, 576, 7, 128, 256>(in, in_len, out, cap, result6, checksum_type);
  else if (lbits == 8) run_fast<8, 264, 7, 128, 256>(in, in_len, out, cap, result6, checksum_type);   // starved second level: slow path
  else run_fast<10, 1152, 8, 288, 512>(in, in_len, out, cap, result6, checksum_type);
  return 0;
}

// The two-phase path of the production inflate kernels on the CPU: the index is parsed with the kernels' own
// parse_index(), block headers are parsed by a whole-stream lane, every 2 KiB sub-range is Huffman-decoded into its token
// map by a tk::TokLane that starts at its indexed bit offset (phase A: the kernel runs 32 of these per warp), and the
// maps of a block are resolved in order by tk::resolve_sub_serial (phase B, stated serially).
// result: [0] produced, [1] status, [2] 1 when the chunk carried an index, [3] sub-ranges decoded.
API int host_inflate_indexed(const uint8_t* in, uint32_t in_len, uint8_t* out, uint32_t cap, uint32_t* result4) {
  using namespace bitar::fl;
  using namespace bitar;
  constexpr int LB = 10, LT = 1344, DB = 8, DT = 512, RG = 128;
  using Gen = FastLane<LB, LT, DB, DT, RG, false>;
  using Tok = tk::TokLane<LB, LT, DB, DT>;
  alignas(16) static thread_local uint8_t smem[LaneLayout<LB, LT, DB, DT, 256>::kStride];
  alignas(16) static thread_local uint8_t uring[tk::kLaneRingBytes];
  alignas(16) static thread_local uint8_t slots[32][tk::kSlotBytes];
  static thread_local LaneScratch scratch;
  static CtaTables cta;
  for (int i = 0; i < 32; ++i) cta.dinfo[i] = dist_info(i);
  result4[0] = result4[1] = result4[2] = result4[3] = 0;
  IndexInfo ix;
  if (!parse_index(in, in_len, &ix)) return 0;
  result4[2] = 1;
  if (ix.total_out > cap) {
    result4[1] = kStatusOutOfSpace;
    return 0;
  }
  const uint32_t nb = dfl::idx_blocks(ix.total_out);
  uint32_t status = kStatusOk;
  for (uint32_t b = 0; b < nb && status == kStatusOk; ++b) {
    const uint32_t blen = ix.total_out - (b << 16) < 65536u ? ix.total_out - (b << 16) : 65536u;
    const uint32_t ns = dfl::idx_subs(blen);
    BlockBits bb;
    if (!index_block_bits(ix, b, nb, &bb)) { status = kStatusDataError; break; }
    const uint32_t hdr = bb.hdr;
    Gen g;
    g.bind(smem, &cta, &scratch, 0);
    g.start(in, ix.stream_bytes, out + (b << 16), blen);
    g.bits_init(hdr >> 3);
    g.drop(hdr & 7u);
    g.header();
    if (g.status != kStatusOk) { status = g.status; break; }
    if (g.state == Gen::kStored) {   // stored block(s): the whole-stream lane copies them
      uint64_t steps = 0;
      while (g.state != Gen::kDone && g.produced() < blen && ++steps < (1ull << 30)) g.step();
      while (g.state != Gen::kDone) { g.state = Gen::kFinish; g.step(); }
      if (g.status != kStatusOk || g.produced() != blen) status = g.status ? g.status : (uint32_t)kStatusDataError;
      continue;
    }
    if ((uint32_t)(8ll * g.start_off + g.consumed_bits()) != index_word(ix, b * 33u + 1u)) { status = kStatusDataError; break; }
    for (uint32_t s = 0; s < ns; ++s) {   // phase A
      Tok l;
      l.bind(g.lt, g.dt, uring, cta.dinfo, &scratch);
      const uint32_t len = blen - s * dfl::kSub < dfl::kSub ? blen - s * dfl::kSub : dfl::kSub;
      uint32_t sbit, ebit;
      if (!index_sub_bits(ix, b, s, ns, bb, &sbit, &ebit)) { status = kStatusDataError; break; }
      l.start_sub(in, ix.stream_bytes, sbit, ebit, s + 1 == ns, slots[s], len, s * dfl::kSub);
      uint64_t steps = 0;
      while (l.state != Tok::kDone && ++steps < (1ull << 30)) l.step();
      result4[3]++;
      if (l.status != kStatusOk) { status = l.status; break; }
      if (l.tokens() > dfl::kSub) { status = kStatusDataError; break; }
    }
    if (status == kStatusOk)   // phase B
      for (uint32_t s = 0; s < ns; ++s) {
        const uint32_t len = blen - s * dfl::kSub < dfl::kSub ? blen - s * dfl::kSub : dfl::kSub;
        if (!tk::resolve_sub_serial(slots[s], out + (b << 16), s * dfl::kSub, len)) { status = kStatusDataError; break; }
      }
  }
  result4[0] = status == kStatusOk ? ix.total_out : 0;
  result4[1] = status;
  return 0;
}

// The speculative path of streams WITHOUT an index (inflate_spec.h) on the CPU: block headers by a whole-stream lane,
// then rounds of 32 sp::SpecLane ranges run one after the other (the kernel runs them as the lanes of a warp), the same
// round bookkeeping as inflate_spec_kernel.cuh, ranges resolved by sp::resolve_range_serial.
// result: [0] produced, [1] 0 = decoded / 1 = declined (the whole-stream kernel's business), [2] blocks, [3] rounds,
//         [4] ranges that contributed, [5] rounds cut short by a lane without a join point, [6] walk symbols,
//         [7] rounds cut short by a full slot, [8] sum over the rounds of the longest lane's steps (the warp's time),
//         [9] steps of the lanes that contributed
API int host_inflate_spec(const uint8_t* in, uint32_t in_len, uint8_t* out, uint32_t cap, uint32_t* result12, uint32_t target) {
  uint32_t* result8 = result12;
  using namespace bitar::fl;
  using namespace bitar;
  constexpr int LB = 9, LT = 864, DB = 7, DT = 256;
  using Gen = FastLane<LB, LT, DB, DT, 256, false>;
  using Lane = sp::SpecLane<LB, LT, DB, DT>;
  alignas(16) static thread_local uint8_t smem[LaneLayout<LB, LT, DB, DT, 256>::kStride];
  alignas(16) static thread_local uint8_t uring[32][tk::kLaneRingBytes];
  alignas(16) static thread_local uint8_t slots[32][sp::kSlotBytes];
  static thread_local LaneScratch scratch;
  static CtaTables cta;
  for (int i = 0; i < 32; ++i) cta.dinfo[i] = dist_info(i);
  for (int i = 0; i < 12; ++i) result8[i] = 0;
  result8[1] = 1;
  if (in_len < 8 || cap == 0) return 0;
  uint32_t total = 0, bit = 0, tgt = target;
  for (;;) {
    Gen g;
    g.bind(smem, &cta, &scratch, 0);
    g.start(in, in_len, out, cap);
    g.bits_init(bit >> 3);
    g.drop(bit & 7u);
    g.header();
    result8[2]++;
    if (g.status != kStatusOk || g.state != Gen::kDecode) return 0;   // stored blocks, bad headers: declined
    const uint32_t last = g.last;
    uint32_t first = (uint32_t)(8ll * g.start_off + g.consumed_bits());
    for (;;) {   // rounds
      const uint32_t B = sp::range_bits(first, in_len, total, cap, tgt);
      static thread_local Lane lanes[32];
      for (int r = 0; r < 32; ++r) {
        lanes[r].bind(g.lt, g.dt, uring[r], cta.dinfo, &scratch);
        const uint64_t start = (uint64_t)first + (uint64_t)r * B;
        if (start < 8ull * in_len) lanes[r].start_spec(in, in_len, (uint32_t)start, (uint32_t)start + B, slots[r]);
        else lanes[r].idle();
      }
      result8[3]++;
      for (int r = 0; r < 32; ++r)
        for (uint32_t i = 0; i < sp::kRec; ++i) lanes[r].step(false);
      for (int r = 0; r < 32; ++r) lanes[r].set_next(slots[(r + 1) & 31], r < 31 ? lanes[r + 1].nrec : 0u, r < 31);
      uint64_t longest = 0;
      uint32_t lane_steps[32];
      for (int r = 0; r < 32; ++r) {
        uint64_t steps = 0;
        while (lanes[r].state != Lane::kDone && ++steps < (1ull << 24)) {
          if (lanes[r].state == Lane::kWalk) result8[6]++;
          lanes[r].step(true);
        }
        if (lanes[r].state != Lane::kDone) return 0;
        longest = steps > longest ? steps : longest;
        lane_steps[r] = (uint32_t)steps + sp::kRec;
      }
      result8[8] += (uint32_t)longest + sp::kRec;
      int m = 0;
      while (m < 31 && lanes[m].end_kind == sp::kEndSync) ++m;
      if (lanes[m].end_kind == sp::kEndSync) return 0;   // (lane 31 has no successor: it stops)
      uint32_t j = 0;
      for (int r = 0; r <= m; ++r) {
        const sp::RangeOut ro = sp::range_out(lanes[r], j);
        if ((uint64_t)total + ro.len > cap) return 0;
        if (!sp::resolve_range_serial(slots[r], out, total, ro.len, ro.tskip, ro.bskip)) return 0;
        total += ro.len;
        j = lanes[r].sync_j;
        result8[4]++;
        result8[9] += lane_steps[r];
      }
      const uint32_t ek = lanes[m].end_kind;
      if (ek == sp::kEndStop && m < 31) result8[5]++;
      if (ek == sp::kEndFull) result8[7]++;
      if (ek == sp::kEndBad) return 0;
      if (lanes[m].end_bit > 8u * in_len) return 0;     // the chain ran past the input
      if (ek != sp::kEndEob && lanes[0].end_bit == first) return 0;   // no progress
      if (ek == sp::kEndFull && tgt > 128u) tgt >>= 1;   // ranges too long for their slots: shorter ones from here on
      first = lanes[m].end_bit;
      if (lanes[m].end_kind == sp::kEndEob) break;
    }
    bit = first;
    if (last) break;
  }
  result8[0] = total;
  result8[1] = 0;
  return 0;
}

API long model_deflate_chunk(const uint8_t* src, uint32_t n, uint8_t* dst, uint32_t cap, int huffman,
                             int block) {
  bitar_model::Params P;
  P.huffman = huffman;
  if (block > 0) P.block = block;
  auto out = bitar_model::deflate_chunk(src, n, P);
  if (out.size() > cap) return -1;
  memcpy(dst, out.data(), out.size());
  return (long)out.size();
}

// Length-limited Huffman construction (deflate_common.h, shared with the deflate kernel) on random and
// adversarial frequency sets: returns the number of sets whose code is not complete (Kraft sum != 1) or
// exceeds the length limit.
API int model_huffman_fuzz(int trials, unsigned seed) {
  using namespace bitar::dfl;
  uint64_t x = seed * 0x9E3779B97F4A7C15ull + 1;
  auto rnd = [&]() {
    x ^= x << 13; x ^= x >> 7; x ^= x << 17;
    return (uint32_t)(x >> 16);
  };
  static thread_local HuffScratch hs;
  int bad = 0;
  for (int t = 0; t < trials; ++t) {
    const int n = (t & 1) ? 286 : (t & 2) ? 30 : 19, maxb = n == 19 ? 7 : 15;
    const int mode = (int)(rnd() % 5);
    uint32_t key[288];
    int m = 0;
    for (int i = 0; i < n; ++i) {
      const uint32_t r = rnd();
      uint32_t f = mode == 0 ? (r % 3 == 0 ? 0 : 1u << (r % 20)) : mode == 1 ? r % 1000 : mode == 2 ? (i < 30 ? 1u << (i % 22) : 0)
                   : mode == 3 ? (r % 5 == 0 ? r % 100000 : r % 3) : (i < 40 ? (uint32_t)(1.0 * (1u << 22) / (1 + i * i * i)) : r % 2);
      if (f > (1u << 22)) f = 1u << 22;
      if (f) key[m++] = (f << 9) | (uint32_t)i;
    }
    if (m < 2) continue;
    std::sort(key, key + m);
    uint8_t len[288] = {0};
    uint16_t blc[16];
    huff_lengths_from_sorted(key, m, maxb, len, blc, &hs);
    uint64_t kraft = 0;
    bool too_long = false;
    for (int i = 0; i < n; ++i)
      if (len[i]) {
        too_long |= len[i] > maxb;
        kraft += 1ull << (maxb - len[i]);
      }
    bad += too_long || kraft != (1ull << maxb);
  }
  return bad;
}

// symbol-map self check used by tests: returns 0 when the computed maps agree with RFC 1951 tables
API int model_selfcheck(void) {
  using namespace bitar::dfl;
  static const uint16_t lb[29] = {3,  4,  5,  6,  7,  8,  9,  10, 11,  13,  15,  17,  19,  23, 27,
                                  31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
  static const uint8_t le[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
  static const uint16_t db[30] = {1,   2,   3,   4,   5,   7,    9,    13,   17,   25,   33,   49,   65,    97,    129,
                                  193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
  static const uint8_t de[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
  static const uint8_t co[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
  for (int s = 0; s < 29; ++s)
    if (len_base(s) != lb[s] || len_extra_bits(s) != le[s]) return 1;
  for (int s = 0; s < 30; ++s)
    if (dist_base(s) != db[s] || dist_extra_bits(s) != de[s]) return 2;
  for (int len = 3; len <= 258; ++len) {
    int s = len_sym(len);
    if (s < 0 || s > 28 || len < lb[s] || (s < 28 && len >= lb[s + 1] && s != 27) ||
        lb[s] + len_extra_val(len, s) != len)
      return 3;
    if (s == 27 && len > 257) return 3;
  }
  for (int d = 1; d <= 32768; ++d) {
    int s = dist_sym(d);
    if (s < 0 || s > 29 || db[s] + dist_extra_val(d, s) != d || dist_extra_val(d, s) >= (1 << de[s])) return 4;
  }
  for (int i = 0; i < 19; ++i)
    if (cl_order(i) != co[i]) return 5;
  return 0;
}
