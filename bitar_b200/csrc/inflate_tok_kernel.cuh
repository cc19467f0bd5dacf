// inflate_tok_kernel.cuh -- K4: two-phase inflate of chunks that carry the parallel-inflate index (deflate_common.h),
// i.e. every chunk this library's deflate kernel produced.
//
//   plan kernel     one thread per op: finds the index at the end of the buffer (fl::parse_index), turns every
//                   64 KiB block of an indexed chunk into a TASK and sends everything else (zlib streams, stored-only
//                   chunks, tiny chunks) to the whole-stream kernel (inflate_kernel.cuh).
//   inflate kernel  persistent CTAs, one task per WARP at a time, fetched from a global counter; two phases per block:
//   phase A (tok)     1. all lanes parse the block header together (same bits, same registers),
//                     2. the warp builds the two decode tables cooperatively in its shared memory,
//                     3. lane s Huffman-decodes sub-range s (2 KiB of output) from its indexed bit offset into token
//                        units (tk::TokLane, inflate_tok.h) in the warp's own scratch (L2 resident, reused per block)
//                        and checks that it ends exactly at the next offset.
//                   A block offers 32 independent symbol chains instead of one.  Stored blocks are copied here.
//   phase B (res)   the same warp takes the block's units in stream order, 32 at a time (one per lane: prefix sum of the
//                   token lengths, literal bytes stored at once, LZ77 copies one match after the other with the lanes
//                   as bytes) through a 4 KiB ring in shared memory -- the space of the tables, which are done with;
//                   sources farther back come from the block's own flushed output; the ring leaves as aligned
//                   16-byte vector stores, 512 bytes per flush.  The copy chain of a block is serial: this phase lives
//                   on few instructions per token and on the other warps of the SM being in phase A meanwhile.
//   checksum        (only when configured) one warp per indexed op over the finished output.
//
// Replaces: rte_compressdev decompress ops assembled at /root/reference/src/memory.cc:432-505 and executed
// behind src/device.cc:464-535 (dst segment i at out + i*S, src/memory.cc:482-493).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/bitar_cuda.h"
#include "inflate_kernel.cuh"   // ik::CksSmem, ik::group_checksum
#include "inflate_tok.h"

namespace bitar {
namespace xk {

struct Task {
  uint32_t op;
  uint32_t block;
};

// device-side work counters of one inflate call (zeroed before the plan kernel)
struct Counters {
  unsigned int n_tasks, n_generic, task_next, generic_next;
  unsigned int n_indexed, res_next, ck_next, pad;
};
constexpr uint32_t kSmallSubs = 8;

__global__ void __launch_bounds__(128)
    inflate_plan_kernel(const bitar_chunk* __restrict__ ops, uint32_t n_ops, bitar_result* __restrict__ results,
                        Task* __restrict__ tasks, uint32_t* __restrict__ indexed, uint32_t* __restrict__ generic,
                        Counters* __restrict__ pc, int use_index) {
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
    indexed[atomicAdd(&pc->n_indexed, 1u)] = i;
    const uint32_t base = atomicAdd(&pc->n_tasks, nb);
    for (uint32_t b = 0; b < nb; ++b) tasks[base + b] = Task{i, b};
    return;
  }
  generic[atomicAdd(&pc->n_generic, 1u)] = i;
}

// shared memory of one GROUP of lanes (a whole warp, or 8 lanes for small blocks): the block's tables + the lanes' rings
template <int LT, int DT, int URING, int GROUP>
struct __align__(16) WarpSmem {
  static constexpr int kRingStride = 2 * URING + 16;   // 16-byte aligned, lanes spread over the banks
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

// ---- phase B ------------------------------------------------------------------------------------------
// One group of G lanes resolves one block right after decoding it: the units of its sub-ranges in stream order, G at a
// time (one per lane): prefix sum of the token lengths, literal bytes stored at once, LZ77 copies one match after the
// other with the lanes as bytes.  Positions are "virtual" (offset in the block + (address of the block & 15)), so that
// multiples of 16 are 16-byte aligned addresses.
// CTA-shared constants: lane % d and the largest multiple of d that fits the group, for copies whose source overlaps them.
struct ResolveLut {
  uint8_t mod[32][32];   // [d][lane]
  uint8_t per32[32];     // [d] for a group of 32 lanes
  uint8_t per8[32];      //     ... of 8 lanes
};
__device__ __forceinline__ void resolve_lut_init(ResolveLut* lut) {
  for (unsigned i = threadIdx.x; i < 1024u; i += blockDim.x) lut->mod[i >> 5][i & 31u] = (uint8_t)((i >> 5) ? (i & 31u) % (i >> 5) : (i & 31u));   // row 0: the identity
  if (threadIdx.x < 32u) {
    const unsigned d = threadIdx.x ? threadIdx.x : 1u;
    lut->per32[threadIdx.x] = (uint8_t)(32u - 32u % d);
    lut->per8[threadIdx.x] = (uint8_t)(d <= 8u ? 8u - 8u % d : 8u);
  }
}
// shared-memory accesses by 32-bit shared address (device only)
__device__ __forceinline__ uint32_t r_ld8(uint32_t a) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ void r_st8(uint32_t a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v)); }
__device__ __forceinline__ uint4 r_ld128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
  return v;
}

template <int G, int RING>
struct ResolveGroup {
  static_assert(RING >= 2048 && (RING & (RING - 1)) == 0, "ring: power of two");
  static constexpr uint32_t RM = RING - 1;
  static constexpr uint32_t kPartMax = 4u * 258u + 8u;          // bytes 8 units can produce: the most written ahead of a match
  static constexpr uint32_t kFlush = 16u * G;                   // one vector per lane
  static constexpr uint32_t kNear = RING - kPartMax - 258u - 32u;   // matches at most this far back find their source in the ring
  uint32_t ring_s;          // shared address of the group's ring (followed by 16 match records of 8 bytes when G == 32)
  uint32_t lut_s;           // shared address of the CTA's ResolveLut
  uint8_t* vbase;           // block address - mis
  uint32_t flushed;         // virtual position below which everything is in global memory
  unsigned gmask;
  int gl;

  // store the complete 16-byte vectors below `upto` (all bytes below are final), and the unaligned head of the block
  __device__ __forceinline__ void flush(uint32_t upto) {
    if (flushed & 15u) {
      const uint32_t a = (flushed + 15u) & ~15u;
      if (upto < a) return;
      for (uint32_t v = flushed + (uint32_t)gl; v < a; v += G) vbase[v] = (uint8_t)r_ld8(ring_s + (v & RM));
      flushed = a;
    }
    const uint32_t end = upto & ~15u;
    for (uint32_t v = flushed + 16u * (uint32_t)gl; v < end; v += kFlush) *reinterpret_cast<uint4*>(vbase + v) = r_ld128(ring_s + (v & RM));
    if (end > flushed) flushed = end;
    __syncwarp(gmask);      // the stores are ordered before later reads of the group (far sources)
  }
  __device__ __forceinline__ void finish(uint32_t upto) {
    flush(upto);
    for (uint32_t v = flushed + (uint32_t)gl; v < upto; v += G) vbase[v] = (uint8_t)r_ld8(ring_s + (v & RM));
    flushed = upto;
    __syncwarp(gmask);
  }

  // out[dst .. dst + len) = out[dst - dist ..]; everything below dst is final.  All arguments are uniform in the group,
  // so the branches do not diverge.
  __device__ __forceinline__ void copy(uint32_t dst, uint32_t len, uint32_t dist) {
    const uint32_t lane_u = (uint32_t)gl;
    const uint32_t dp = dst + lane_u, sp = dp - dist;
    if (dist >= len || dist >= (uint32_t)G) {
      if (dist <= kNear) {
        if (lane_u < len) r_st8(ring_s + (dp & RM), r_ld8(ring_s + (sp & RM)));
        for (uint32_t k = G; k < len; k += G) {             // a match longer than a pass of the group
          __syncwarp(gmask);                                  // (this pass may read what the last one wrote: dist < 2 G)
          if (k + lane_u < len) r_st8(ring_s + ((dp + k) & RM), r_ld8(ring_s + ((sp + k) & RM)));
        }
      } else {
        // flushed long ago: dist > kNear, so the source ends below `flushed`
        for (uint32_t k = lane_u; k < len; k += G) r_st8(ring_s + ((dst + k) & RM), (uint32_t)__ldcg(vbase + (dst - dist) + k));
      }
    } else {
      // the source overlaps the copy and repeats inside one pass: every lane keeps its byte, a pass writes a whole
      // number of periods
      const uint32_t r = r_ld8(lut_s + dist * 32u + lane_u);
      const uint32_t per = r_ld8(lut_s + 1024u + (G == 32 ? 0u : 32u) + dist);
      const uint32_t byte = r_ld8(ring_s + ((dst - dist + r) & RM));
      if (lane_u < per)
        for (uint32_t k = lane_u; k < len; k += per) r_st8(ring_s + ((dst + k) & RM), byte);
    }
    __syncwarp(gmask);
  }

  // The units of one sub-range (n_units of them, a multiple of 8, at `units`) from virtual position `pos`; returns the
  // new position, or 0xFFFFFFFF when the units do not add up to `sub_limit` (they always do when phase A succeeded).
  __device__ __forceinline__ uint32_t resolve(const uint16_t* units, uint32_t n_units, uint32_t pos, uint32_t sub_limit) {
    const int wlane = (int)(threadIdx.x & 31u);
    const int gbase = wlane - gl;
    // two groups of units in flight ahead of the one being resolved (ld.cg: the scratch is rewritten for every block)
    uint32_t u0 = (uint32_t)gl < n_units ? __ldcg(units + gl) : tk::kUnitNop;
    uint32_t u1 = (uint32_t)(G + gl) < n_units ? __ldcg(units + G + gl) : tk::kUnitNop;
    for (uint32_t i = 0; i < n_units; i += G) {
      const uint32_t u = u0;
      u0 = u1;
      u1 = i + 2u * G + (uint32_t)gl < n_units ? __ldcg(units + i + 2u * G + (uint32_t)gl) : tk::kUnitNop;
      const bool is_head = (u & tk::kUnitHead) != 0u;
      const unsigned hm = __ballot_sync(gmask, is_head);
      const bool is_cont = gl > 0 && ((hm >> (wlane - 1)) & 1u);
      const bool is_lit = !is_head && !is_cont && u < 0x100u;
      const uint32_t olen = is_lit ? 1u : is_head ? (u & 0xFFu) + 3u : 0u;
      uint32_t incl = olen;
#pragma unroll
      for (int d = 1; d < G; d <<= 1) {
        const uint32_t v = __shfl_up_sync(gmask, incl, d, G);
        if (gl >= d) incl += v;
      }
      const uint32_t total = __shfl_sync(gmask, incl, G - 1, G);
      if (pos + total > sub_limit) return 0xFFFFFFFFu;
      const uint32_t my = pos + incl - olen;            // where this lane's token starts
      // groups of 8 units are resolved together when the G units produce more than fits ahead of a match in the ring
      const int parts = G > 8 && total > kPartMax ? G / 8 : 1;
      for (int part = 0; part < parts; ++part) {
        const bool mine = parts == 1 || (gl >> 3) == part;
        if (is_lit && mine) r_st8(ring_s + (my & RM), u);
        unsigned heads = (hm >> gbase) & (G == 32 ? 0xFFFFFFFFu : ((1u << G) - 1u));
        if (parts > 1) heads &= 0xFFu << (8 * part);
        if (G == 32) {
          // the heads leave their (start, length, distance) in shared memory in stream order: a match is then one
          // broadcast read away for the whole warp
          const uint32_t rec_s = ring_s + (uint32_t)RING;
          const int nh = __popc(heads);
          const uint32_t next_u = __shfl_down_sync(gmask, u, 1, G);          // a head's distance sits in the next lane
          if (is_head && mine) {
            const uint32_t rank = (uint32_t)__popc(heads & ((1u << gl) - 1u));
            asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(rec_s + 8u * rank), "r"(my | (olen << 20)), "r"(next_u + 1u) : "memory");
          }
          __syncwarp(gmask);
          for (int t = 0; t < nh; ++t) {
            uint32_t w0, dist;
            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(w0), "=r"(dist) : "r"(rec_s + 8u * (uint32_t)t) : "memory");
            copy(w0 & 0xFFFFFu, w0 >> 20, dist);
          }
        } else {
          __syncwarp(gmask);
          while (heads) {
            const int h = __ffs((int)heads) - 1;
            heads &= heads - 1u;
            const uint32_t len = __shfl_sync(gmask, olen, h, G);
            const uint32_t dist = __shfl_sync(gmask, u, h + 1, G) + 1u;
            const uint32_t dst = __shfl_sync(gmask, my, h, G);
            copy(dst, len, dist);
            if (dst + len - flushed >= 2u * kFlush) flush(dst + len);
          }
        }
        if (parts > 1) flush(pos + __shfl_sync(gmask, incl, 8 * part + 7, G));   // everything up to the end of this part is final
      }
      pos += total;
      if (pos - flushed >= kFlush) flush(pos);
    }
    return pos == sub_limit ? pos : 0xFFFFFFFFu;
  }
};

// The same for a whole warp per block (the 64 KiB blocks of ordinary segments), laid out for few instructions per match:
// the block goes through a LINEAR window in shared memory -- the previous sub-range and the current one side by side,
// indexed by position, no wrap-around -- so that a copy is `buf[d + lane] = buf[s + f(lane)]` with everything worked out
// by the match's head lane beforehand (all heads of a batch in parallel) and left in a record the warp reads by
// broadcast.  Matches whose source lies below the window (flushed output) are independent of the batch: they are copied
// first, their loads in flight together; the others follow one after the other.  A sub-range leaves the window as
// aligned 16-byte vectors when it is complete; then the window slides by one sub-range.
struct ResolveWarp {
  static constexpr uint32_t kWin = dfl::kSub;                 // bytes of a sub-range
  static constexpr uint32_t kBuf = 2u * kWin + 16u;           // previous + current sub-range + the block's misalignment
  static constexpr uint32_t kBytes = kBuf + 16u * 16u + 16u * 8u;   // + a record per possible match of a batch (near, far)
  uint32_t buf_s;           // shared address of the window (16-byte aligned)
  uint32_t lut_s;           // shared address of the CTA's ResolveLut
  uint8_t* vbase;           // block address - mis
  uint32_t flushed;         // virtual position below which everything is in global memory
  uint32_t cur0;            // virtual position of window index kWin (a multiple of 16)
  int gl;

  __device__ __forceinline__ void begin(uint8_t* out) {
    const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(out) & 15u);
    vbase = out - mis;
    flushed = mis;
    cur0 = 0;
  }
  // shared address of virtual position v
  __device__ __forceinline__ uint32_t at(uint32_t v) const { return buf_s + kWin + (v - cur0); }

  // store what is final: the unaligned head of the block, then the complete 16-byte vectors below `upto`
  __device__ __forceinline__ void flush(uint32_t upto) {
    if (flushed & 15u) {
      const uint32_t a = (flushed + 15u) & ~15u;
      if (upto < a) return;
      for (uint32_t v = flushed + (uint32_t)gl; v < a; v += 32u) vbase[v] = (uint8_t)r_ld8(at(v));
      flushed = a;
    }
    const uint32_t end = upto & ~15u;
    for (uint32_t v = flushed + 16u * (uint32_t)gl; v < end; v += 512u) *reinterpret_cast<uint4*>(vbase + v) = r_ld128(at(v));
    if (end > flushed) flushed = end;
  }
  __device__ __forceinline__ void finish(uint32_t upto) {
    flush(upto);
    for (uint32_t v = flushed + (uint32_t)gl; v < upto; v += 32u) vbase[v] = (uint8_t)r_ld8(at(v));
    flushed = upto;
    __syncwarp();
  }
  // the current sub-range becomes the previous one (no pass reads what another pass writes, except lane 0's first and
  // last vector, which program order takes care of)
  __device__ __forceinline__ void slide() {
    for (uint32_t i = 16u * (uint32_t)gl; i < kWin + 16u; i += 512u) {
      const uint4 v = r_ld128(buf_s + kWin + i);
      asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(buf_s + i), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    }
    cur0 += kWin;
    __syncwarp();
  }

  // the rare shapes, out of line: what is left of far matches longer than a pass of the warp ...
  __device__ __noinline__ void far_rest(uint32_t a0, uint32_t s0, uint32_t a1, uint32_t s1) {
    const uint32_t lane_u = (uint32_t)gl, l0 = a0 >> 20, l1 = a1 >> 20;
#pragma unroll 1
    for (uint32_t k = 32u + lane_u; k < l0; k += 32u) r_st8((a0 & 0xFFFFFu) + k, (uint32_t)__ldcg(vbase + s0 + k));
#pragma unroll 1
    for (uint32_t k = 32u + lane_u; k < l1; k += 32u) r_st8((a1 & 0xFFFFFu) + k, (uint32_t)__ldcg(vbase + s1 + k));
  }
  // ... and near matches longer than a pass: a pattern (every lane keeps the byte of its place in the period, a pass
  // writes a whole number of periods), or a plain copy pass by pass
  __device__ __noinline__ void near_long(uint32_t a, uint32_t len, uint32_t sa, uint32_t row, uint32_t per) {
    const uint32_t lane_u = (uint32_t)gl;
    if (row != lut_s) {
      const uint32_t byte = r_ld8(sa + r_ld8(row + lane_u));
      if (lane_u < per) {
#pragma unroll 1
        for (uint32_t k = lane_u; k < len; k += per) r_st8(a + k, byte);
      }
    } else {
#pragma unroll 1
      for (uint32_t k = lane_u; k < len + lane_u; k += 32u) {
        if (k < len) r_st8(a + k, r_ld8(sa + k));
        __syncwarp();            // (the next pass may read what this one wrote: dist < 64)
      }
    }
  }

  // The units of one sub-range (n_units of them, a multiple of 8, at `units`) from virtual position `pos`; returns the
  // new position, or 0xFFFFFFFF when the units do not add up to `sub_limit` (they always do when phase A succeeded).
  __device__ __forceinline__ uint32_t resolve(const uint16_t* units, uint32_t n_units, uint32_t pos, uint32_t sub_limit) {
    const uint32_t lane_u = (uint32_t)gl;
    const uint32_t near_s = buf_s + kBuf, far_s = near_s + 16u * 16u;
    const uint32_t bias = buf_s + kWin - cur0;                // shared address of virtual position 0 (mod 2^32)
    // two batches of units in flight ahead of the one being resolved (ld.cg: the scratch is rewritten for every block)
    uint32_t u0 = lane_u < n_units ? __ldcg(units + lane_u) : tk::kUnitNop;
    uint32_t u1 = 32u + lane_u < n_units ? __ldcg(units + 32u + lane_u) : tk::kUnitNop;
    for (uint32_t i = 0; i < n_units; i += 32u) {
      const uint32_t u = u0;
      u0 = u1;
      u1 = i + 64u + lane_u < n_units ? __ldcg(units + i + 64u + lane_u) : tk::kUnitNop;
      const bool is_head = (u & tk::kUnitHead) != 0u;
      const unsigned heads = __ballot_sync(0xFFFFFFFFu, is_head);
      const bool is_cont = ((heads << 1) >> lane_u) & 1u;     // the unit after a head: its distance (a head is never in lane 31)
      const bool is_lit = !is_head && !is_cont && u < 0x100u;
      const uint32_t olen = is_lit ? 1u : is_head ? (u & 0xFFu) + 3u : 0u;
      uint32_t incl = olen;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (gl >= d) incl += v;
      }
      const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
      if (pos + total > sub_limit) return 0xFFFFFFFFu;
      const uint32_t my = pos + incl - olen;                  // where this lane's token starts
      const uint32_t da = bias + my;
      if (is_lit) r_st8(da, u);
      const uint32_t dist = __shfl_down_sync(0xFFFFFFFFu, u, 1) + 1u;   // a head's distance sits in the next lane
      const bool is_far = is_head && dist > my - cur0 + kWin;           // the source starts below the window
      const unsigned fars = __ballot_sync(0xFFFFFFFFu, is_far);
      const unsigned nears = heads & ~fars;
      if (is_head) {
        const unsigned below = (1u << lane_u) - 1u;
        if (is_far) {
          // destination address | length, virtual position of the source
          asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(far_s + 8u * (uint32_t)__popc(fars & below)), "r"(da | (olen << 20)), "r"(my - dist) : "memory");
        } else {
          // destination address | length, source address, row of lane -> source byte (lane % dist for a source that
          // overlaps the copy: a repeating pattern; the identity else), bytes a pass may write (whole periods)
          const bool pattern = dist < olen && dist < 32u;
          const uint32_t row = lut_s + (pattern ? dist * 32u : 0u);
          const uint32_t per = pattern ? r_ld8(lut_s + 1024u + dist) : 32u;
          asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(near_s + 16u * (uint32_t)__popc(nears & below)), "r"(da | (olen << 20)),
                       "r"(da - dist), "r"(row), "r"(per) : "memory");
        }
      }
      __syncwarp();
      {   // the far matches: nothing in this batch depends on their sources, so their loads go out together
        const int nf = __popc(fars);
#pragma unroll 1
        for (int t = 0; t < nf; t += 2) {
          uint32_t a0, s0, a1 = 0, s1 = 0;
          asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(a0), "=r"(s0) : "r"(far_s + 8u * (uint32_t)t) : "memory");
          if (t + 1 < nf) asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(a1), "=r"(s1) : "r"(far_s + 8u * (uint32_t)(t + 1)) : "memory");
          const uint32_t l0 = a0 >> 20, l1 = a1 >> 20;
          uint32_t b0 = 0, b1 = 0;
          if (lane_u < l0) b0 = (uint32_t)__ldcg(vbase + s0 + lane_u);
          if (lane_u < l1) b1 = (uint32_t)__ldcg(vbase + s1 + lane_u);
          if (lane_u < l0) r_st8((a0 & 0xFFFFFu) + lane_u, b0);
          if (lane_u < l1) r_st8((a1 & 0xFFFFFu) + lane_u, b1);
          if (l0 > 32u || l1 > 32u) far_rest(a0, s0, a1, s1);
        }
        __syncwarp();
      }
      const int nn = __popc(nears);
#pragma unroll 1
      for (int t = 0; t < nn; ++t) {
        uint32_t w0, sa, row, per;
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(w0), "=r"(sa), "=r"(row), "=r"(per) : "r"(near_s + 16u * (uint32_t)t) : "memory");
        const uint32_t len = w0 >> 20, a = w0 & 0xFFFFFu;
        if (len <= 32u) {
          // one pass: lane k takes byte k of the match from byte (k mod dist) of its source (k itself when they do not overlap)
          const uint32_t byte = r_ld8(sa + r_ld8(row + lane_u));
          if (lane_u < len) r_st8(a + lane_u, byte);
        } else {
          near_long(a, len, sa, row, per);
        }
        __syncwarp();
      }
      pos += total;
    }
    if (pos != sub_limit) return 0xFFFFFFFFu;
    flush(pos);
    return pos;
  }
};

// ---- phase A ------------------------------------------------------------------------------------------
// GROUP = 32: a warp per block (up to 32 sub-ranges).  GROUP = 8: four blocks of at most 8 sub-ranges per warp
// (small segments), each with its own tables; the groups of a warp run the same code on their own tasks and only
// ever synchronise among their own lanes.  Every group owns `subs` slots of the unit scratch (tk::kSlotBytes each).
template <int LBITS, int LT, int DBITS, int DT, int WARPS, int GROUP, int MIN_CTAS>
__global__ void __launch_bounds__(WARPS * 32, MIN_CTAS)
    inflate_tok_kernel(const bitar_chunk* __restrict__ ops, bitar_result* __restrict__ results, const Task* __restrict__ tasks,
                       Counters* __restrict__ pc, uint8_t* scratch, uint32_t subs) {
  using Lane = tk::TokLane<LBITS, LT, DBITS, DT, 16>;
  using WS = WarpSmem<LT, DT, 16, GROUP>;
  constexpr int kGroupsPerWarp = 32 / GROUP;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint32_t* dinfo = reinterpret_cast<uint32_t*>(smem_raw + (size_t)WARPS * kGroupsPerWarp * sizeof(WS));
  ResolveLut* lut = reinterpret_cast<ResolveLut*>(dinfo + 32);
  if (threadIdx.x < 32) dinfo[threadIdx.x] = fl::dist_info((int)threadIdx.x);
  resolve_lut_init(lut);
  __syncthreads();

  const int wlane = (int)(threadIdx.x & 31u);
  const int lane = wlane % GROUP;                     // lane inside the group
  const int gbase = wlane - lane;                     // first lane of the group inside the warp
  const unsigned kFull = GROUP == 32 ? 0xFFFFFFFFu : (((1u << GROUP) - 1u) << gbase);   // the group's lanes
  WS& ws = *reinterpret_cast<WS*>(smem_raw + (size_t)((threadIdx.x >> 5) * kGroupsPerWarp + gbase / GROUP) * sizeof(WS));
  Lane L;
  L.bind(ws.lt, ws.dt, ws.ring + lane * WS::kRingStride, dinfo, &ws.sc);
  // phase B: the resolver's window / ring takes the space of the tables and unit rings once phase A is done with them
  constexpr int kRing = 2048;
  static_assert(sizeof(WS) >= (GROUP == 32 ? (size_t)ResolveWarp::kBytes : (size_t)kRing), "the resolver borrows the group's phase-A space");
  ResolveGroup<GROUP == 32 ? 8 : GROUP, kRing> R;    // (groups of 8 lanes; unused by the warp-per-block instance)
  R.gl = lane;
  R.gmask = kFull;
  R.ring_s = (uint32_t)__cvta_generic_to_shared(&ws);
  R.lut_s = (uint32_t)__cvta_generic_to_shared(lut);
  ResolveWarp W;
  W.gl = lane;
  W.buf_s = (uint32_t)__cvta_generic_to_shared(&ws);
  W.lut_s = (uint32_t)__cvta_generic_to_shared(lut);
  // the group's unit scratch: `subs` slots, rewritten for every block
  uint8_t* const slots = scratch + ((size_t)(blockIdx.x * WARPS + (threadIdx.x >> 5)) * kGroupsPerWarp + gbase / GROUP) * subs * tk::kSlotBytes;
  const uint32_t n_tasks = pc->n_tasks;

  for (;;) {
    uint32_t t = 0;
    if (lane == 0) t = atomicAdd(&pc->task_next, 1u);
    t = __shfl_sync(kFull, t, gbase);
    if (t >= n_tasks) break;
    const Task tk_ = tasks[t];
    const bitar_chunk op = ops[tk_.op];
    const uint8_t* src = static_cast<const uint8_t*>(op.src);
    fl::IndexInfo ix;
    fl::parse_index(src, op.src_len, &ix);   // validated by the plan kernel
    const uint32_t nb = dfl::idx_blocks(ix.total_out), b = tk_.block;
    const uint32_t blen = min(65536u, ix.total_out - (b << 16)), ns = dfl::idx_subs(blen);
    fl::BlockBits bb;
    uint32_t status = fl::index_block_bits(ix, b, nb, &bb) && ns <= subs ? fl::kStatusOk : fl::kStatusDataError;
    const uint32_t hdr = status == fl::kStatusOk ? bb.hdr : 0u, block_end = bb.end;
    uint8_t* out = static_cast<uint8_t*>(op.dst) + ((size_t)b << 16);
    uint32_t type = 3u;

    // ---- block header: every lane reads the same bits ----
    L.in = src;
    L.in_len = ix.stream_bytes;
    L.status = fl::kStatusOk;
    L.bits_init(hdr >> 3);
    L.drop(hdr & 7u);
    L.refill();
    if (status == fl::kStatusOk) {
      const uint32_t last = L.take(1);
      type = L.take(2);
      const uint32_t want_last = b + 1u == nb ? 1u : 0u;
      if ((type != 0u && last != want_last) || type == 3u) status = fl::kStatusDataError;
      if (status == fl::kStatusOk && type == 0u) {
        // stored block: one or two pieces (65535 + 1; only the last carries the block's BFINAL), copied by the whole group
        uint32_t done = 0, at = ((hdr + 3u + 7u) >> 3), piece_last = last;
        for (;;) {
          if ((uint64_t)at + 4u > ix.stream_bytes) { status = fl::kStatusDataError; break; }
          const uint32_t len = (uint32_t)src[at] | ((uint32_t)src[at + 1] << 8);
          const uint32_t nlen = (uint32_t)src[at + 2] | ((uint32_t)src[at + 3] << 8);
          at += 4u;
          if ((len ^ 0xFFFFu) != nlen || done + len > blen || (uint64_t)at + len > ix.stream_bytes) { status = fl::kStatusDataError; break; }
          for (uint32_t i = (uint32_t)lane; i < len; i += (uint32_t)GROUP) out[done + i] = src[at + i];
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
        // ---- lane s decodes sub-range s into units ----
        if (status == fl::kStatusOk) {
          L.state = Lane::kDone;
          L.upos = 0;
          if ((uint32_t)lane < ns) {
            const uint32_t s = (uint32_t)lane;
            uint32_t sbit, ebit;
            const uint32_t len = min(dfl::kSub, blen - s * dfl::kSub);
            if (!fl::index_sub_bits(ix, b, s, ns, bb, &sbit, &ebit)) L.status = fl::kStatusDataError;
            else
              L.start_sub(src, ix.stream_bytes, sbit, ebit, s + 1u == ns, slots + (size_t)s * tk::kSlotBytes, len, s * dfl::kSub);
          }
          while (L.state != Lane::kDone) L.step();
          status = L.status;
          // ---- phase B: the block's units in stream order ----
          const uint32_t my_units = (uint32_t)lane < ns ? L.units() : 0u;
          if (!__any_sync(kFull, status != fl::kStatusOk)) {   // (also orders the unit stores before the loads below)
            const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(out) & 15u);
            uint32_t pos = mis;
            if (GROUP == 32) {
              W.begin(out);
              for (uint32_t s = 0; s < ns; ++s) {
                const uint32_t nu = __shfl_sync(kFull, my_units, (int)s, GROUP);
                if (s) W.slide();
                pos = W.resolve(reinterpret_cast<const uint16_t*>(slots + (size_t)s * tk::kSlotBytes), nu, pos, mis + min(blen, (s + 1u) * dfl::kSub));
                if (pos == 0xFFFFFFFFu) break;
              }
              if (pos == 0xFFFFFFFFu) status = fl::kStatusDataError;
              else W.finish(pos);
            } else {
              R.vbase = out - mis;
              R.flushed = mis;
              for (uint32_t s = 0; s < ns; ++s) {
                const uint32_t nu = __shfl_sync(kFull, my_units, (int)s, GROUP);
                pos = R.resolve(reinterpret_cast<const uint16_t*>(slots + (size_t)s * tk::kSlotBytes), nu, pos, mis + min(blen, (s + 1u) * dfl::kSub));
                if (pos == 0xFFFFFFFFu) break;
              }
              if (pos == 0xFFFFFFFFu) status = fl::kStatusDataError;
              else R.finish(pos);
            }
          }
          __syncwarp(kFull);   // the ring's space goes back to phase A
        }
      }
    }
    if (status != fl::kStatusOk) atomicMax(&results[tk_.op].status, status);   // any lane's failure fails the op
    __syncwarp(kFull);
  }
}

// ---- checksum of the indexed ops (only launched when a checksum is configured): one warp per op ----
__global__ void __launch_bounds__(128)
    inflate_checksum_kernel(const bitar_chunk* __restrict__ ops, bitar_result* __restrict__ results, const uint32_t* __restrict__ indexed,
                            Counters* __restrict__ pc, int checksum_type) {
  __shared__ ik::CksSmem ck;
  for (unsigned i = threadIdx.x; i < 256; i += blockDim.x) ck.crc_tab[i] = cks::crc_table_entry(i);
  if (threadIdx.x == 0) cks::crc_x2n_init(ck.x2n);
  __syncthreads();
  inf::Group<32> g;
  g.lane = (int)(threadIdx.x & 31u);
  g.mask = 0xFFFFFFFFu;
  const uint32_t n = pc->n_indexed;
  for (;;) {
    uint32_t k = 0;
    if (g.lane == 0) k = atomicAdd(&pc->ck_next, 1u);
    k = __shfl_sync(0xFFFFFFFFu, k, 0);
    if (k >= n) break;
    const uint32_t i = indexed[k];
    if (results[i].status != BITAR_OP_OK) continue;
    const uint64_t sum = ik::group_checksum<32>(static_cast<const uint8_t*>(ops[i].dst), results[i].produced, checksum_type, &ck, g);
    if (g.lane == 0) results[i].checksum = sum;
  }
}

template <int LBITS, int LT, int DBITS, int DT, int WARPS, int GROUP = 32, int MIN_CTAS = 2>
struct TokConfig {
  static constexpr int kThreads = WARPS * 32;
  static constexpr size_t kSmem = (size_t)WARPS * (32 / GROUP) * sizeof(WarpSmem<LT, DT, 16, GROUP>) + 32 * sizeof(uint32_t) + sizeof(ResolveLut);
  static int ctas_per_sm(int device) {
    static int per_device[64] = {0};
    int& c = per_device[device & 63];
    if (c == 0) {
      auto kern = inflate_tok_kernel<LBITS, LT, DBITS, DT, WARPS, GROUP, MIN_CTAS>;
      if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem) != cudaSuccess) return 0;
      cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c, kern, kThreads, kSmem) != cudaSuccess) return 0;
    }
    return c;
  }
  // bytes of unit scratch the largest grid needs: `subs` slots per group
  static size_t scratch_bytes(int device, int sm_count, uint32_t subs) {
    return (size_t)sm_count * (size_t)ctas_per_sm(device) * WARPS * (32 / GROUP) * subs * tk::kSlotBytes;
  }
  // n_tasks_max: upper bound of the task count (the real count lives on the device)
  static cudaError_t launch(const bitar_chunk* ops, bitar_result* res, const Task* tasks, Counters* pc, uint8_t* scratch, uint32_t subs,
                            uint32_t n_tasks_max, int device, int sm_count, cudaStream_t stream) {
    const int c = ctas_per_sm(device);
    if (c < 1) return cudaErrorLaunchOutOfResources;
    uint32_t grid = (uint32_t)(sm_count * c);
    const uint32_t per_cta = WARPS * (32 / GROUP);
    const uint32_t want = (n_tasks_max + per_cta - 1) / per_cta;
    if (want < grid) grid = want;
    if (grid == 0) return cudaSuccess;
    inflate_tok_kernel<LBITS, LT, DBITS, DT, WARPS, GROUP, MIN_CTAS><<<grid, kThreads, kSmem, stream>>>(ops, res, tasks, pc, scratch, subs);
    return cudaGetLastError();
  }
};

}  // namespace xk
}  // namespace bitar
