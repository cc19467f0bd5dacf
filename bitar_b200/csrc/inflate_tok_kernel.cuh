// inflate_tok_kernel.cuh -- K4: two-phase inflate of chunks that carry the parallel-inflate index (deflate_common.h),
// i.e. every chunk this library's deflate kernel produced.
//
//   plan kernel     one thread per op: finds the index at the end of the buffer (fl::parse_index), turns every
//                   64 KiB block of an indexed chunk into a TASK and sends everything else (zlib streams, stored-only
//                   chunks, tiny chunks) to the whole-stream kernel (inflate_kernel.cuh).
//   inflate kernel  persistent CTAs, one task per WARP at a time, fetched from a global counter; two phases per block:
//   phase A (tok)     1. all lanes parse the block header together (same bits, same registers),
//                     2. the warp builds the two decode tables cooperatively in its shared memory,
//                     3. lane s Huffman-decodes sub-range s (2 KiB of output) from its indexed bit offset into a token
//                        map (tk::TokLane, inflate_tok.h: 16 bits per token and one start bit per output byte)
//                        in the warp's own scratch (L2 resident, reused per block) and checks that it ends exactly
//                        at the next offset.
//                   A block offers 32 independent symbol chains instead of one.  Stored blocks are copied here.
//   phase B (res)   the same warp resolves the block in stream order, BYTE-parallel: 32 output bytes per step, one per
//                   lane.  A population count on the start bits tells a lane which token of the sub-range its byte
//                   belongs to, the bit after its own whether that is a literal; tokens are read one step ahead.  A match byte
//                   copies from a 4 KiB ring of the latest output in shared memory -- the space of the tables, which are
//                   done with -- or, for sources farther back, from the block's own flushed output (L2); sources inside
//                   the step's own 32 bytes are followed by pointer jumping over shuffles (five rounds cover any
//                   chain).  The ring leaves as aligned 16-byte vector stores once per sub-range.  The cost of a step
//                   does not depend on how many tokens it holds.
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
  unsigned int n_indexed, spec_next, ck_next, n_declined;   // spec_next / n_declined: inflate_spec_kernel.cuh
  unsigned long long sum_generic;                           // compressed bytes of the ops without an index
  unsigned int pad[6];
};
static_assert(sizeof(Counters) == 16 * sizeof(unsigned int), "capi.cu: kCountersPerBatch");

// Longest first (the speculative kernel's list).  Work is handed out by a counter, in list order; a buffer of columns puts
// its costliest chunks (the least compressible column) wherever that column lies -- at the end in the benchmark's buffer,
// where they leave the last wave of warps half empty.  That list is therefore walked twice: the first pass takes the ops
// whose compressed size is at least the mean of the list (cost follows the token count, i.e. the compressed size), the
// second pass the rest.  (Not for the tasks of indexed chunks: measured, no gain at a warp per block -- and at four blocks
// per warp a group that skips a task falls out of step with the other three groups of its warp, which then execute one
// after the other: 4 KiB segments 44 -> 28 GB/s.)
__device__ __forceinline__ bool first_pass_op(uint32_t src_len, uint32_t n, unsigned long long sum) {
  return (unsigned long long)src_len * n >= sum;
}
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
  atomicAdd(&pc->sum_generic, (unsigned long long)op.src_len);
}

// shared memory of one GROUP of lanes (a whole warp, or 8 lanes for small blocks): the block's tables + the lanes' rings
template <int LT, int DT, int GROUP>
struct __align__(16) WarpSmem {
  static constexpr int kRingStride = (int)tk::kLaneRingBytes + 16;   // 16-byte aligned, lanes spread over the banks
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
// One group of G lanes resolves one block right after decoding it, sub-range by sub-range, G output bytes per step.
// Positions are "virtual" (offset in the block + (address of the block & 15)), so that multiples of 16 are 16-byte
// aligned addresses, in the ring as in global memory.
// shared-memory accesses by 32-bit shared address (device only)
__device__ __forceinline__ uint32_t r_ld8(uint32_t a) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ uint32_t r_ld32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void r_st8(uint32_t a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void r_st32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ uint4 r_ld128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void r_st128(uint32_t a, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// RING : bytes of the latest output kept in shared memory.  FLUSH: the ring is stored every FLUSH bytes of a sub-range.
// A match byte whose source lies at most kNear back reads the ring; a source farther back lies below the last flush.
template <int G, int RING, int FLUSH>
struct ResolveBytes {
  static_assert(G == 32 || G == 8, "a warp, or a quarter of one");
  static_assert((RING & (RING - 1)) == 0 && dfl::kSub % FLUSH == 0 && FLUSH % G == 0, "ring: power of two; whole steps between flushes");
  static constexpr uint32_t RM = RING - 1;
  static constexpr uint32_t kNear = RING - G;                   // (the step's own stores may already have replaced anything older)
  static_assert(kNear >= (uint32_t)FLUSH + G + 16u, "a source that is not in the ring must have been flushed");
  static_assert(RING >= FLUSH + 16 + G, "a flush reads what the ring still holds");
  static constexpr uint32_t kBitsBytes = 288u;                  // staged start bits: 64 words, then words that read as "all starts"
  static constexpr uint32_t kBytes = RING + kBitsBytes;
  static constexpr uint32_t kAll = G == 32 ? 0xFFFFFFFFu : ((1u << G) - 1u);
  uint32_t ring_s;          // shared address of the group's ring (16-byte aligned), followed by the staged start bits
  uint8_t* vbase;           // block address - mis
  uint32_t flushed;         // virtual position below which everything is in global memory
  unsigned gmask;           // the group's lanes in the warp
  int gl;                   // lane inside the group

  __device__ __forceinline__ void begin(uint8_t* out) {
    const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(out) & 15u);
    vbase = out - mis;
    flushed = mis;
  }
  // store what is final: the unaligned head of the block, then the complete 16-byte vectors below `upto`
  __device__ __forceinline__ void flush(uint32_t upto) {
    if (flushed & 15u) {
      const uint32_t a = (flushed + 15u) & ~15u;
      if (upto < a) return;
      for (uint32_t v = flushed + (uint32_t)gl; v < a; v += G) vbase[v] = (uint8_t)r_ld8(ring_s + (v & RM));
      flushed = a;
    }
    const uint32_t end = upto & ~15u;
    for (uint32_t v = flushed + 16u * (uint32_t)gl; v < end; v += 16u * G) *reinterpret_cast<uint4*>(vbase + v) = r_ld128(ring_s + (v & RM));
    if (end > flushed) flushed = end;
    __syncwarp(gmask);      // the stores are ordered before later reads of the group (far sources)
  }
  __device__ __forceinline__ void finish(uint32_t upto) {
    flush(upto);
    for (uint32_t v = flushed + (uint32_t)gl; v < upto; v += G) vbase[v] = (uint8_t)r_ld8(ring_s + (v & RM));
    flushed = upto;
    __syncwarp(gmask);
  }

  // start bits of step k (s) and which of them are literals (a start whose NEXT byte is a start too)
  __device__ __forceinline__ void masks(uint32_t k, uint32_t& s, uint32_t& lit) const {
    const uint32_t bm_s = ring_s + (uint32_t)RING;
    uint32_t nx;
    if (G == 32) {
      s = r_ld32(bm_s + 4u * k);
      nx = r_ld32(bm_s + 4u * k + 4u);
    } else {
      s = r_ld8(bm_s + k);
      nx = r_ld8(bm_s + k + 1u);
    }
    lit = s & ((s >> 1) | (nx << (G - 1)));
    if (G != 32) lit &= kAll;
  }

  // The `len` bytes of one sub-range, from its slot, to virtual position v0 on.
  __device__ __forceinline__ void resolve_sub(const uint8_t* slot, uint32_t v0, uint32_t len) {
    const uint32_t lane_u = (uint32_t)gl;
    const uint32_t bm_s = ring_s + (uint32_t)RING;
    // stage the start bits.  Everything from bit `len` on reads as a start: the byte after the sub-range is one (tokens
    // never straddle sub-ranges), and the lanes of the last step that lie past the end then look like literals -- they
    // fetch a token of the slot's slack and store it to ring positions nobody reads before they are written again.
    for (uint32_t i = lane_u; i < 16u; i += G) r_st128(bm_s + 16u * i, __ldcg(reinterpret_cast<const uint4*>(slot + tk::kSlotBits) + i));
    __syncwarp(gmask);
    if (lane_u == 0) {
      const uint32_t w = len >> 5, sh = len & 31u;
      const uint32_t old = sh ? r_ld32(bm_s + 4u * w) & ((1u << sh) - 1u) : 0u;
      r_st32(bm_s + 4u * w, old | (0xFFFFFFFFu << sh));
      r_st32(bm_s + 4u * w + 4u, 0xFFFFFFFFu);
      r_st32(bm_s + 4u * w + 8u, 0xFFFFFFFFu);
      r_st32(bm_s + 4u * w + 12u, 0xFFFFFFFFu);
    }
    __syncwarp(gmask);
    const uint16_t* toks = reinterpret_cast<const uint16_t*>(slot + tk::kSlotToks);
    const uint32_t le = kAll >> (G - 1 - gl);       // the lanes up to and including this one
    const uint32_t nsteps = (len + G - 1u) / G;
    // Software pipeline, two steps deep.  This lane's token for a step -- the byte of a literal, or the distance - 1 of
    // the match its byte belongs to -- is the one of the nearest start at or below its byte; its place follows from
    // the start bits alone, so it is fetched two steps ahead (x2), and a source that lies below the ring (a FAR match
    // byte: its address needs the token) one step ahead (fb1).
    uint32_t tbase = 0;                             // tokens of the sub-range that start before the step being fetched
    uint32_t s2, lit1, lit2;
    masks(0, s2, lit1);
    uint32_t x1 = (uint32_t)__ldcg(toks + ((uint32_t)__popc(s2 & le) - 1u));
    tbase += (uint32_t)__popc(s2);
    masks(1, s2, lit2);
    uint32_t x2 = (uint32_t)__ldcg(toks + (tbase + (uint32_t)__popc(s2 & le) - 1u));
    uint32_t pv = v0 + lane_u;
    uint32_t fb1 = 0;
    if (!((lit1 >> lane_u) & 1u) && x1 + 1u > kNear) fb1 = (uint32_t)__ldcg(vbase + (pv - (x1 + 1u)));
#pragma unroll 2
    for (uint32_t k = 0; k < nsteps; ++k) {
      const uint32_t lit = lit1, x = x1, fb = fb1;
      lit1 = lit2;
      x1 = x2;
      tbase += (uint32_t)__popc(s2);
      masks(k + 2u, s2, lit2);
      x2 = (uint32_t)__ldcg(toks + (tbase + (uint32_t)__popc(s2 & le) - 1u));
      fb1 = 0;
      if (!((lit1 >> lane_u) & 1u) && x1 + 1u > kNear) fb1 = (uint32_t)__ldcg(vbase + (pv + G - (x1 + 1u)));
      uint32_t b = x;
      if (lit != kAll) {                            // the step holds match bytes
        const uint32_t dist = x + 1u;
        const bool mb = !((lit >> lane_u) & 1u);
        const bool inside = mb && dist <= lane_u;   // the source is a byte of this very step
        const bool far = mb && dist > kNear;
        if (mb && !inside && !far) b = r_ld8(ring_s + ((pv - dist) & RM));
        if (far) b = fb;
        if (__any_sync(gmask, inside)) {            // follow the chain of sources to a byte that is known: G - 1 hops at most
          uint32_t ptr = inside ? lane_u - dist : lane_u;
#pragma unroll
          for (int r = 1; r < G; r <<= 1) ptr = __shfl_sync(gmask, ptr, (int)ptr, G);
          b = __shfl_sync(gmask, b, (int)ptr, G);
        }
      }
      r_st8(ring_s + (pv & RM), b);
      pv += G;
      __syncwarp(gmask);
      if (FLUSH < (int)dfl::kSub && ((k + 1u) * G) % FLUSH == 0u && (k + 1u) * G < len) flush(v0 + (k + 1u) * G);
    }
    flush(v0 + len);
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
  using Lane = tk::TokLane<LBITS, LT, DBITS, DT>;
  using WS = WarpSmem<LT, DT, GROUP>;
  using Res = ResolveBytes<GROUP, GROUP == 32 ? 4096 : 2048, GROUP == 32 ? 2048 : 1024>;
  constexpr int kGroupsPerWarp = 32 / GROUP;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint32_t* dinfo = reinterpret_cast<uint32_t*>(smem_raw + (size_t)WARPS * kGroupsPerWarp * sizeof(WS));
  if (threadIdx.x < 32) dinfo[threadIdx.x] = fl::dist_info((int)threadIdx.x);
  __syncthreads();

  const int wlane = (int)(threadIdx.x & 31u);
  const int lane = wlane % GROUP;                     // lane inside the group
  const int gbase = wlane - lane;                     // first lane of the group inside the warp
  const unsigned kFull = GROUP == 32 ? 0xFFFFFFFFu : (((1u << GROUP) - 1u) << gbase);   // the group's lanes
  WS& ws = *reinterpret_cast<WS*>(smem_raw + (size_t)((threadIdx.x >> 5) * kGroupsPerWarp + gbase / GROUP) * sizeof(WS));
  Lane L;
  L.bind(ws.lt, ws.dt, ws.ring + lane * WS::kRingStride, dinfo, &ws.sc);
  // phase B: the resolver's ring takes the space of the tables and lane rings once phase A is done with them
  static_assert(sizeof(WS) >= (size_t)Res::kBytes, "the resolver borrows the group's phase-A space");
  Res R;
  R.gl = lane;
  R.gmask = kFull;
  R.ring_s = (uint32_t)__cvta_generic_to_shared(&ws);
  // the group's token-map scratch: `subs` slots, rewritten for every block
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
        // ---- lane s decodes sub-range s into its token map ----
        if (status == fl::kStatusOk) {
          L.state = Lane::kDone;
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
          // ---- phase B: the block's bytes in stream order ----
          if (!__any_sync(kFull, status != fl::kStatusOk)) {   // (also orders the lanes' map stores before the loads below)
            const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(out) & 15u);
            R.begin(out);
            for (uint32_t s = 0; s < ns; ++s) {
              // the next sub-range's map may have left L2 since its lane wrote it (4 736 warps x ~85 KB are in flight):
              // ask for it now (a 128-byte line per lane covers the tokens, two more the start bits)
              if (GROUP == 32 && s + 1u < ns) {
                const uint8_t* nx = slots + (size_t)(s + 1u) * tk::kSlotBytes;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + 128u * (uint32_t)lane));
                if (lane < 2) asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + tk::kSlotBits + 128u * (uint32_t)lane));
              }
              R.resolve_sub(slots + (size_t)s * tk::kSlotBytes, mis + s * dfl::kSub, min(dfl::kSub, blen - s * dfl::kSub));
            }
            R.finish(mis + blen);
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
  static constexpr size_t kSmem = (size_t)WARPS * (32 / GROUP) * sizeof(WarpSmem<LT, DT, GROUP>) + 32 * sizeof(uint32_t);
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
