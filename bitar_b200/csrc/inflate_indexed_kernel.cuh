// inflate_indexed_kernel.cuh -- K4: sub-range parallel inflate of chunks that carry the parallel-inflate
// index (deflate_common.h), i.e. every chunk this library's deflate kernel produced.
//
//   plan kernel    one thread per op: finds the index at the end of the buffer (fl::parse_index), turns every
//                  64 KiB block of an indexed chunk into a TASK and sends everything else (zlib streams,
//                  stored-only chunks, tiny chunks) to the whole-stream kernel (inflate_kernel.cuh).
//   indexed kernel persistent CTAs, one task per WARP at a time, fetched from a global counter:
//                    1. all lanes parse the block header together (same bits, same registers),
//                    2. the warp builds the two decode tables cooperatively in its shared memory,
//                    3. lane s decodes sub-range s (2 KiB of output) from its indexed bit offset with
//                       fl::FastLane<SUB = true>::step() and checks that it ends exactly at the next offset.
//                  A block offers 32 independent symbol chains instead of one, so a 1 GiB buffer keeps
//                  ~90 k chains in flight (148 SMs x ~20 warps) where a stream-per-warp decoder has ~4 k and
//                  a stream-per-lane decoder ~19 k; that is what the throughput follows.
//
// Replaces: rte_compressdev decompress ops assembled at /root/reference/src/memory.cc:432-505 and executed
// behind src/device.cc:464-535 (dst segment i at out + i*S, src/memory.cc:482-493).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/bitar_cuda.h"
#include "inflate_fast.h"

namespace bitar {
namespace xk {

struct Task {
  uint32_t op;
  uint32_t block;
};

// device-side work counters of one inflate call (zeroed before the plan kernel)
struct Counters {
  unsigned int n_tasks, n_generic, task_next, generic_next;
  unsigned int n_small, small_next;   // blocks of at most kSmallSubs sub-ranges: four per warp (GROUP = 8)
};
constexpr uint32_t kSmallSubs = 8;

// per-op checksum accumulators (only touched when a checksum is configured): partial sums of the blocks /
// sub-ranges are folded in with atomics, the task that finishes last publishes the checksum
struct CkAcc {
  uint32_t crc, a, b, remaining;
};

__global__ void __launch_bounds__(128)
    inflate_plan_kernel(const bitar_chunk* __restrict__ ops, uint32_t n_ops, bitar_result* __restrict__ results,
                        Task* __restrict__ tasks, Task* __restrict__ small_tasks, uint32_t* __restrict__ generic,
                        Counters* __restrict__ pc, CkAcc* __restrict__ acc, int use_index, int allow_small) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_ops) return;
  const bitar_chunk op = ops[i];
  fl::IndexInfo ix;
  if (use_index && op.src && fl::parse_index(static_cast<const uint8_t*>(op.src), op.src_len, &ix)) {
    bitar_result r;
    r.checksum = 0;
    if (ix.total_out > op.dst_cap) {
      r.produced = 0;
      r.status = BITAR_OP_OUT_OF_SPACE;
      results[i] = r;
      return;
    }
    r.produced = ix.total_out;
    r.status = BITAR_OP_OK;
    results[i] = r;
    const uint32_t nb = dfl::idx_blocks(ix.total_out);
    if (acc) acc[i] = CkAcc{0u, 0u, 0u, nb};
    // a chunk of one block with few sub-ranges would leave most lanes of a warp idle: such blocks go to the
    // list that is decoded four to a warp
    if (allow_small && nb == 1u && dfl::idx_subs(ix.total_out) <= kSmallSubs) {
      small_tasks[atomicAdd(&pc->n_small, 1u)] = Task{i, 0u};
      return;
    }
    const uint32_t base = atomicAdd(&pc->n_tasks, nb);
    for (uint32_t b = 0; b < nb; ++b) tasks[base + b] = Task{i, b};
    return;
  }
  generic[atomicAdd(&pc->n_generic, 1u)] = i;
}

// shared memory of one GROUP of lanes (a whole warp, or 8 lanes for small blocks): the block's tables + the lanes' rings
template <int LT, int DT, int RING, int GROUP>
struct __align__(16) WarpSmem {
  static constexpr int kRingStride = RING + 16;   // 16-byte aligned, lanes spread over the banks
  uint16_t lt[LT];
  uint16_t dt[DT];
  fl::LaneScratch sc;        // code lengths + canonical side arrays of the block (shared by the group)
  uint32_t cnt[16], at[16];  // table construction scratch
  uint8_t ring[GROUP * kRingStride];
};

// Cooperative construction of one decode table (same layout and validity rules as fl::build_table) by a group of
// G lanes of a warp: gl = lane index inside the group, gmask = the group's lanes.
template <int G>
__device__ __forceinline__ uint32_t warp_build_table(const uint8_t* lens, int n, int kind, uint16_t* table, int tbits,
                                                     int capacity, uint16_t* count, uint16_t* first, uint16_t* offs,
                                                     uint16_t* sorted, uint32_t* cnt32, uint32_t* at32, int gl, unsigned gmask) {
  const int lane = (int)(threadIdx.x & 31u);
  for (int b = gl; b < 16; b += G) cnt32[b] = 0;
  __syncwarp(gmask);
  for (int i = gl; i < n; i += G) atomicAdd(&cnt32[lens[i]], 1u);
  __syncwarp(gmask);
  int left = 1, maxl = 0;
  bool over = false;
  uint32_t code0 = 0, o = 0;
  if (gl == 0) {
    count[0] = (uint16_t)cnt32[0];
    first[0] = offs[0] = 0;
    at32[0] = 0;
  }
  for (int b = 1; b <= 15; ++b) {   // every lane walks the 15 lengths (broadcast reads), the leader records them
    const uint32_t c = cnt32[b];
    left = (left << 1) - (int)c;
    if (c) maxl = b;
    if (left < 0) over = true;
    if (gl == 0) {
      count[b] = (uint16_t)c;
      first[b] = (uint16_t)code0;
      offs[b] = (uint16_t)o;
      at32[b] = o;
    }
    code0 = (code0 + c) << 1;
    o += c;
  }
  const int used = n - (int)cnt32[0];
  __syncwarp(gmask);
  if (over) return fl::kStatusDataError;
  if (left > 0 && used > 0 && (kind == fl::kCodeLen || maxl != 1)) return fl::kStatusDataError;
  // stable counting sort by (length, symbol): G symbols per round
  const unsigned lt_mask = (1u << lane) - 1u;
  for (int base = 0; base < n; base += G) {
    const int i = base + gl;
    const uint32_t l = i < n ? lens[i] : 0u;
    const unsigned peers = __match_any_sync(gmask, l);
    if (l) sorted[at32[l] + __popc(peers & lt_mask)] = (uint16_t)i;
    __syncwarp(gmask);
    if (l && (peers & lt_mask) == 0) at32[l] += (uint32_t)__popc(peers);
    __syncwarp(gmask);
  }
  const uint32_t fill = kind == fl::kLitLen ? fl::kBadEntry : kind == fl::kDist ? fl::kBadDist : 0u;
  uint32_t* t32 = reinterpret_cast<uint32_t*>(table);
  for (int j = gl; j < capacity / 2; j += G) t32[j] = fill | (fill << 16);
  __syncwarp(gmask);
  // root entries, one symbol per lane
  const int n_root = tbits < 15 ? (int)offs[tbits] + (int)count[tbits] : used;   // symbols with length <= tbits
  for (int idx = gl; idx < n_root; idx += G) {
    const int sym = sorted[idx], l = lens[sym];
    const uint32_t code = (uint32_t)first[l] + (uint32_t)(idx - (int)offs[l]);
    const uint32_t r = __brev(code) >> (32 - l);
    const uint16_t e = kind == fl::kLitLen ? fl::ll_entry(sym, l) : kind == fl::kDist ? fl::d_entry(sym, l) : (uint16_t)((sym << 4) | l);
    for (uint32_t k = r; k < (1u << tbits); k += (1u << l)) table[k] = e;
  }
  __syncwarp(gmask);
  // codes longer than the root: second-level tables, allocated in canonical order by the leader (few symbols)
  if (gl == 0 && n_root < used) {
    int idx = n_root, next_free = 1 << tbits;
    while (idx < used) {
      const int l = lens[sorted[idx]];
      const uint32_t prefix = ((uint32_t)first[l] + (uint32_t)(idx - (int)offs[l])) >> (l - tbits);
      int j = idx, lmax = l;
      while (j < used) {
        const int l2 = lens[sorted[j]];
        const uint32_t c2 = (uint32_t)first[l2] + (uint32_t)(j - (int)offs[l2]);
        if ((c2 >> (l2 - tbits)) != prefix) break;
        lmax = l2;
        ++j;
      }
      int sub_bits = lmax - tbits;
      if (sub_bits < 2) sub_bits = 2;
      const int size = 1 << sub_bits;
      if (next_free + size <= capacity) {
        const int rel = next_free - (1 << tbits);
        table[__brev(prefix) >> (32 - tbits)] = kind == fl::kLitLen ? fl::ll_link(rel, sub_bits) : fl::d_link(rel, sub_bits);
        for (int k = idx; k < j; ++k) {
          const int sym = sorted[k], lk = lens[sym], rest = lk - tbits;
          const uint32_t ck = (uint32_t)first[lk] + (uint32_t)(k - (int)offs[lk]);
          const uint32_t r = __brev(ck & ((1u << rest) - 1u)) >> (32 - rest);
          const uint16_t e = kind == fl::kLitLen ? fl::ll_entry(sym, lk) : fl::d_entry(sym, lk);
          for (int t = (int)r; t < size; t += (1 << rest)) table[next_free + t] = e;
        }
        next_free += size;
      }
      idx = j;
    }
  }
  __syncwarp(gmask);
  return fl::kStatusOk;
}

// GROUP = 32: a warp per block (up to 32 sub-ranges).  GROUP = 8: four blocks of at most 8 sub-ranges per warp
// (small segments), each with its own tables; the groups of a warp run the same code on their own tasks and only
// ever synchronise among their own lanes.
template <int LBITS, int LT, int DBITS, int DT, int RING, int WARPS, bool CK, int GROUP>
__global__ void __launch_bounds__(WARPS * 32, 1)
    inflate_indexed_kernel(const bitar_chunk* __restrict__ ops, bitar_result* __restrict__ results,
                           const Task* __restrict__ tasks, Counters* __restrict__ pc, CkAcc* __restrict__ acc,
                           int checksum_type) {
  using Lane = fl::FastLane<LBITS, LT, DBITS, DT, RING, true, CK>;
  using WS = WarpSmem<LT, DT, RING, GROUP>;
  constexpr int kGroupsPerWarp = 32 / GROUP;
  if (!CK) checksum_type = BITAR_CHECKSUM_NONE;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  fl::CtaTables* cta = reinterpret_cast<fl::CtaTables*>(smem_raw + (size_t)WARPS * kGroupsPerWarp * sizeof(WS));
  if (threadIdx.x < 32) cta->dinfo[threadIdx.x] = fl::dist_info((int)threadIdx.x);
  if (checksum_type & BITAR_CHECKSUM_CRC32) {
    for (unsigned i = threadIdx.x; i < 256; i += WARPS * 32) cta->crc[0][i] = cks::crc_table_entry(i);
    if (threadIdx.x == 0) cks::crc_x2n_init(cta->x2n);
    __syncthreads();
    for (unsigned i = threadIdx.x; i < 256; i += WARPS * 32) {
      uint32_t c = cta->crc[0][i];
      for (int k = 1; k < 4; ++k) {
        c = (c >> 8) ^ cta->crc[0][c & 0xFFu];
        cta->crc[k][i] = c;
      }
    }
  }
  __syncthreads();

  const int wlane = (int)(threadIdx.x & 31u);
  const int lane = wlane % GROUP;                     // lane inside the group
  const int gbase = wlane - lane;                     // first lane of the group inside the warp
  const unsigned kFull = GROUP == 32 ? 0xFFFFFFFFu : (((1u << GROUP) - 1u) << gbase);   // the group's lanes
  WS& ws = *reinterpret_cast<WS*>(smem_raw + (size_t)((threadIdx.x >> 5) * kGroupsPerWarp + gbase / GROUP) * sizeof(WS));
  Lane L;
  L.bind_parts(ws.lt, ws.dt, ws.ring + lane * WS::kRingStride, cta, &ws.sc, (uint32_t)checksum_type);
  const uint32_t n_tasks = GROUP == 32 ? pc->n_tasks : pc->n_small;
  unsigned int* next_task = GROUP == 32 ? &pc->task_next : &pc->small_next;

  for (;;) {
    uint32_t t = 0;
    if (lane == 0) t = atomicAdd(next_task, 1u);
    t = __shfl_sync(kFull, t, gbase);
    if (t >= n_tasks) break;
    const Task tk = tasks[t];
    const bitar_chunk op = ops[tk.op];
    const uint8_t* src = static_cast<const uint8_t*>(op.src);
    fl::IndexInfo ix;
    fl::parse_index(src, op.src_len, &ix);   // validated by the plan kernel
    const uint32_t nb = dfl::idx_blocks(ix.total_out), b = tk.block;
    const uint32_t blen = min(65536u, ix.total_out - (b << 16)), ns = dfl::idx_subs(blen);
    fl::BlockBits bb;
    uint32_t status = fl::index_block_bits(ix, b, nb, &bb) ? fl::kStatusOk : fl::kStatusDataError;
    const uint32_t hdr = status == fl::kStatusOk ? bb.hdr : 0u, block_end = bb.end;
    uint8_t* out = static_cast<uint8_t*>(op.dst) + ((size_t)b << 16);
    // this lane's share of the block's checksum: (crc register, sum of bytes, weighted sum) of `ck_len` bytes
    // that are followed by `ck_tail` more bytes of the chunk
    uint32_t ck_crc = 0, ck_s1 = 0, ck_s2 = 0, ck_len = 0, ck_tail = 0;

    // ---- block header: every lane reads the same bits ----
    L.in = src;
    L.in_len = ix.stream_bytes;
    L.status = fl::kStatusOk;
    L.bits_init(hdr >> 3);
    L.drop(hdr & 7u);
    L.refill();
    const uint32_t last = L.take(1);
    const uint32_t type = L.take(2);
    const uint32_t want_last = b + 1u == nb ? 1u : 0u;
    if ((type != 0u && last != want_last) || type == 3u) status = fl::kStatusDataError;
    if (status == fl::kStatusOk && type == 0u) {
      // stored block: one or two pieces (65535 + 1; only the last carries the block's BFINAL), copied by the whole warp
      uint32_t done = 0, at = ((hdr + 3u + 7u) >> 3), piece_last = last;
      for (;;) {
        if ((uint64_t)at + 4u > ix.stream_bytes) { status = fl::kStatusDataError; break; }
        const uint32_t len = (uint32_t)src[at] | ((uint32_t)src[at + 1] << 8);
        const uint32_t nlen = (uint32_t)src[at + 2] | ((uint32_t)src[at + 3] << 8);
        at += 4u;
        if ((len ^ 0xFFFFu) != nlen || done + len > blen || (uint64_t)at + len > ix.stream_bytes) { status = fl::kStatusDataError; break; }
        for (uint32_t i = (uint32_t)lane; i < len; i += (uint32_t)GROUP) out[done + i] = src[at + i];
        if (checksum_type != BITAR_CHECKSUM_NONE && ck_len == 0) {   // stored pieces: lane j sums the j-th slice of the block
          const uint32_t per = (blen + (uint32_t)GROUP - 1u) / (uint32_t)GROUP, lo = min(blen, per * (uint32_t)lane), hi = min(blen, lo + per);
          // (the payload of a two-piece block is contiguous in the output, which the warp reads back below)
          ck_len = hi - lo;
          ck_tail = lo;   // temporarily: the slice's offset inside the block
        }
        done += len;
        at += len;
        if (done == blen) {
          if (8u * at != block_end || piece_last != want_last) status = fl::kStatusDataError;
          break;
        }
        // the next piece: BTYPE 0 at the byte boundary
        if ((uint64_t)at + 1u > ix.stream_bytes || (src[at] & 6u) != 0u) { status = fl::kStatusDataError; break; }
        piece_last = src[at] & 1u;
        at += 1u;
      }
    } else if (status == fl::kStatusOk) {
      int nlen, ndist;
      if (type == 1u) {
        for (int i = lane; i < 288; i += GROUP) ws.sc.lens[i] = (uint8_t)dfl::fixed_ll_len(i);
        for (int i = lane; i < 32; i += GROUP) ws.sc.lens[288 + i] = 5;
        nlen = 288;
        ndist = 32;
      } else {
        nlen = (int)L.take(5) + 257;
        ndist = (int)L.take(5) + 1;
        const int ncode = (int)L.take(4) + 4;
        if (nlen > 286 || ndist > 30) status = fl::kStatusDataError;
        for (int i = lane; i < 19; i += GROUP) ws.sc.lens[i] = 0;
        __syncwarp(kFull);
        for (int i = 0; i < ncode; ++i) {
          L.refill();
          const uint32_t v = L.take(3);
          if (lane == 0) ws.sc.lens[dfl::cl_order(i)] = (uint8_t)v;
        }
        __syncwarp(kFull);
        if (status == fl::kStatusOk)
          status = warp_build_table<GROUP>(ws.sc.lens, 19, fl::kCodeLen, ws.dt, 7, 128, ws.sc.d_count, ws.sc.d_first, ws.sc.d_offs,
                                           ws.sc.d_sorted, ws.cnt, ws.at, lane, kFull);
        if (status == fl::kStatusOk) {
          int idx = 0, prev = 0;
          const int total = nlen + ndist;
          while (idx < total) {
            L.refill();
            const uint32_t e = ws.dt[L.lo & 127u];
            if ((e & 15u) == 0) { status = fl::kStatusDataError; break; }
            L.drop(e & 15u);
            const int sym = (int)(e >> 4);
            int rep, val;
            if (sym < 16) { rep = 1; val = sym; prev = sym; }
            else if (sym == 16) {
              if (idx == 0) { status = fl::kStatusDataError; break; }
              rep = 3 + (int)L.take(2); val = prev;
            } else if (sym == 17) { rep = 3 + (int)L.take(3); val = 0; prev = 0; }
            else { rep = 11 + (int)L.take(7); val = 0; prev = 0; }
            if (idx + rep > total) { status = fl::kStatusDataError; break; }
            for (int k = lane; k < rep; k += GROUP) ws.sc.lens[idx + k] = (uint8_t)val;
            idx += rep;
          }
          __syncwarp(kFull);
          if (status == fl::kStatusOk && (L.overrun() || ws.sc.lens[256] == 0)) status = fl::kStatusDataError;
        }
      }
      // the first symbol must sit where the index says sub-range 0 starts
      if (status == fl::kStatusOk &&
          (uint32_t)(8ll * (long long)L.start_off + L.consumed_bits()) != fl::index_word(ix, b * 33u + 1u))
        status = fl::kStatusDataError;
      __syncwarp(kFull);
      if (status == fl::kStatusOk)
        status = warp_build_table<GROUP>(ws.sc.lens + nlen, ndist, fl::kDist, ws.dt, DBITS, DT, ws.sc.d_count, ws.sc.d_first,
                                         ws.sc.d_offs, ws.sc.d_sorted, ws.cnt, ws.at, lane, kFull);
      if (status == fl::kStatusOk)
        status = warp_build_table<GROUP>(ws.sc.lens, nlen, fl::kLitLen, ws.lt, LBITS, LT, ws.sc.ll_count, ws.sc.ll_first,
                                         ws.sc.ll_offs, ws.sc.ll_sorted, ws.cnt, ws.at, lane, kFull);
      // ---- lane s decodes sub-range s ----
      if (status == fl::kStatusOk) {
        L.state = Lane::kDone;
        if ((uint32_t)lane < ns) {
          const uint32_t s = (uint32_t)lane;
          uint32_t sbit, ebit;
          const uint32_t len = min(dfl::kSub, blen - s * dfl::kSub);
          if (!fl::index_sub_bits(ix, b, s, ns, bb, &sbit, &ebit)) L.status = fl::kStatusDataError;
          else {
            L.start_sub(src, ix.stream_bytes, sbit, ebit, s + 1u == ns, out + (size_t)s * dfl::kSub, len, b == 0u && s == 0u);
            ck_len = len;
            ck_tail = ix.total_out - ((b << 16) + s * dfl::kSub + len);
          }
        }
        while (L.state != Lane::kDone) L.step();
        status = L.status;
        if (ck_len) {
          ck_crc = L.crc;
          ck_s1 = L.ad_a % cks::kAdlerMod;
          ck_s2 = (uint32_t)(L.ad_b % cks::kAdlerMod);
        }
      }
    }
    __syncwarp(kFull);
    if (status != fl::kStatusOk) atomicMax(&results[tk.op].status, status);
    if (checksum_type != BITAR_CHECKSUM_NONE) {
      if (type == 0u && ck_len) {   // stored block: sum this lane's slice of what the warp just wrote
        const uint32_t lo = ck_tail;
        const uint8_t* p = out + lo;
        uint32_t state = (b == 0u && lo == 0u) ? 0xFFFFFFFFu : 0u, s1 = 0;
        uint64_t s2 = 0;
        for (uint32_t i = 0; i < ck_len; ++i) {
          const uint32_t byte = *reinterpret_cast<const volatile uint8_t*>(p + i);
          if (checksum_type & BITAR_CHECKSUM_CRC32) state = cta->crc[0][(state ^ byte) & 0xFFu] ^ (state >> 8);
          s1 += byte;
          s2 += (uint64_t)(ck_len - i) * byte;
        }
        ck_crc = state;
        ck_s1 = s1 % cks::kAdlerMod;
        ck_s2 = (uint32_t)(s2 % cks::kAdlerMod);
        ck_tail = ix.total_out - ((b << 16) + lo + ck_len);
      }
      uint32_t c = 0, a = 0, bb = 0;
      if (ck_len) {
        if (checksum_type & BITAR_CHECKSUM_CRC32) c = cks::crc_contrib(ck_crc, ck_tail, cta->x2n);
        a = ck_s1;
        bb = cks::adler_b_contrib(ck_s1, ck_s2, ck_tail % cks::kAdlerMod);
      }
#pragma unroll
      for (int o2 = GROUP / 2; o2 > 0; o2 >>= 1) {
        c ^= __shfl_xor_sync(kFull, c, o2);
        a += __shfl_xor_sync(kFull, a, o2);
        bb += __shfl_xor_sync(kFull, bb, o2);
      }
      if (lane == 0) {
        CkAcc* k = acc + tk.op;
        atomicXor(&k->crc, c);
        atomicAdd(&k->a, a % cks::kAdlerMod);
        atomicAdd(&k->b, bb % cks::kAdlerMod);
        __threadfence();
        if (atomicSub(&k->remaining, 1u) == 1u) {   // the chunk's last block: publish
          __threadfence();
          const uint32_t crc = (checksum_type & BITAR_CHECKSUM_CRC32) ? (atomicXor(&k->crc, 0u) ^ 0xFFFFFFFFu) : 0u;
          const uint32_t adler = (checksum_type & BITAR_CHECKSUM_ADLER32)
                                     ? cks::adler_finish(atomicAdd(&k->a, 0u), atomicAdd(&k->b, 0u), ix.total_out) : 0u;
          results[tk.op].checksum = cks::pack(crc, adler);
        }
      }
    }
  }
}

template <int LBITS, int LT, int DBITS, int DT, int RING, int WARPS, int GROUP = 32>
struct IndexedConfig {
  static constexpr int kThreads = WARPS * 32;
  static constexpr size_t kSmem = (size_t)WARPS * (32 / GROUP) * sizeof(WarpSmem<LT, DT, RING, GROUP>) + sizeof(fl::CtaTables);
  static int ctas_per_sm(int device) {
    static int per_device[64] = {0};
    int& c = per_device[device & 63];
    if (c == 0) {
      auto kern = inflate_indexed_kernel<LBITS, LT, DBITS, DT, RING, WARPS, false, GROUP>;
      auto kern_ck = inflate_indexed_kernel<LBITS, LT, DBITS, DT, RING, WARPS, true, GROUP>;
      if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem) != cudaSuccess) return 0;
      if (cudaFuncSetAttribute(kern_ck, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem) != cudaSuccess) return 0;
      cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
      cudaFuncSetAttribute(kern_ck, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c, kern_ck, kThreads, kSmem) != cudaSuccess) return 0;
    }
    return c;
  }
  // n_tasks_max: upper bound of the task count (the real count lives on the device)
  static cudaError_t launch(const bitar_chunk* ops, bitar_result* res, const Task* tasks, Counters* pc, CkAcc* acc,
                            int checksum_type, uint32_t n_tasks_max, int device, int sm_count, cudaStream_t stream) {
    const int c = ctas_per_sm(device);
    if (c < 1) return cudaErrorLaunchOutOfResources;
    uint32_t grid = (uint32_t)(sm_count * c);
    const uint32_t per_cta = WARPS * (32 / GROUP);
    const uint32_t want = (n_tasks_max + per_cta - 1) / per_cta;
    if (want < grid) grid = want;
    if (grid == 0) return cudaSuccess;
    if (checksum_type == BITAR_CHECKSUM_NONE)
      inflate_indexed_kernel<LBITS, LT, DBITS, DT, RING, WARPS, false, GROUP><<<grid, kThreads, kSmem, stream>>>(ops, res, tasks, pc, acc, 0);
    else
      inflate_indexed_kernel<LBITS, LT, DBITS, DT, RING, WARPS, true, GROUP><<<grid, kThreads, kSmem, stream>>>(ops, res, tasks, pc, acc, checksum_type);
    return cudaGetLastError();
  }
};

}  // namespace xk
}  // namespace bitar
