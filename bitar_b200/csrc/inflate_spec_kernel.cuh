// inflate_spec_kernel.cuh -- K4b: lane-parallel inflate of streams WITHOUT a parallel-inflate index (zlib / hardware
// output, i.e. what the reference itself compressed), by speculation on symbol boundaries (inflate_spec.h).
//
// One WARP per stream, taken from the list the plan kernel left for "everything else".  Per Huffman-coded block:
//   header    all lanes read the same bits; the warp builds the two decode tables cooperatively (xk::warp_build_table);
//   rounds    the rest of the input is cut into ranges of B bits (sp::range_bits: ~`target` output bytes each, judged by
//             the stream's ratio).  Lane r decodes from bit first + r * B -- a guess for r > 0 -- into its slot of the
//             warp's scratch: token map + start bits as tk::TokLane, plus a record of its first step starts.  Past its
//             range it walks on symbol by symbol until it stands on a recorded position of lane r + 1 (sp::SpecLane).
//             Lanes 0 .. m, m = the first lane that found no successor (slot full, no join point, end of block), are
//             good; a shuffle hands each its join record (where its true part starts), a warp scan places the outputs;
//   phase B   the warp resolves ranges 0 .. m in order, byte-parallel, 32 output bytes per step (SpecResolve: the
//             resolver of inflate_tok_kernel.cuh with a token / start-bit offset per range and a check that no source
//             lies before the output).  Its ring is separate from the tables here (they serve the next round).
//   The next round starts at lane m's last position -- a true symbol boundary --, the next block after its end-of-block.
// Anything irregular (stored blocks, bad codes, truncation, output that does not fit, a fixed-length code that never
// resynchronises AND makes no progress) DECLINES the stream: it goes to a second list that the whole-stream kernel
// decodes afterwards, so status words and partial results are that kernel's, exactly as before.
//
// Replaces: rte_compressdev decompress ops on buffers compressed elsewhere (/root/reference/src/memory.cc:432-505,
// src/device.cc:464-535).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/bitar_cuda.h"
#include "inflate_spec.h"
#include "inflate_tok_kernel.cuh"

namespace bitar {
namespace sk {

template <int LT, int DT, int RING>
struct __align__(16) SpecSmem {
  static constexpr int kRingStride = (int)tk::kLaneRingBytes + 16;   // 16-byte aligned, lanes spread over the banks
  uint16_t lt[LT];
  uint16_t dt[DT];
  fl::LaneScratch sc;
  uint32_t cnt[16], at[16];
  uint8_t lane_ring[32 * kRingStride];
  uint8_t ring[RING + 288];                                          // phase B: latest output, then the staged start bits
};

// xk::ResolveBytes for ranges that start anywhere in their slot
template <int RING, int FLUSH>
struct SpecResolve : xk::ResolveBytes<32, RING, FLUSH> {
  using Base = xk::ResolveBytes<32, RING, FLUSH>;
  using Base::flushed;
  using Base::gl;
  using Base::gmask;
  using Base::ring_s;
  using Base::vbase;
  static constexpr uint32_t RM = Base::RM, kNear = Base::kNear, kAll = 0xFFFFFFFFu, G = 32u;
  uint32_t lo_v;     // virtual position of output byte 0: no source lies below it
  uint32_t bad;      // this lane saw a distance that does

  // The `len` bytes (1 .. 2048) of one range to virtual position v0 on: tokens from `tskip`, start bits from `bskip`.
  __device__ __forceinline__ void resolve_range(const uint8_t* slot, uint32_t v0, uint32_t len, uint32_t tskip, uint32_t bskip) {
    const uint32_t lane_u = (uint32_t)gl;
    const uint32_t bm_s = ring_s + (uint32_t)RING;
    {   // stage the start bits, shifted down by bskip; everything from bit `len` on reads as a start
      const uint32_t* bw = reinterpret_cast<const uint32_t*>(slot + sp::kSlotBits) + (bskip >> 5);
      const uint32_t sh = bskip & 31u;
      for (uint32_t i = lane_u; i <= (len >> 5) + 1u; i += G) xk::r_st32(bm_s + 4u * i, __funnelshift_r(__ldcg(bw + i), __ldcg(bw + i + 1u), sh));
      __syncwarp(gmask);
      if (lane_u == 0) {
        const uint32_t w = len >> 5, s2 = len & 31u;
        const uint32_t old = s2 ? xk::r_ld32(bm_s + 4u * w) & ((1u << s2) - 1u) : 0u;
        xk::r_st32(bm_s + 4u * w, old | (0xFFFFFFFFu << s2));
        xk::r_st32(bm_s + 4u * w + 4u, 0xFFFFFFFFu);
        xk::r_st32(bm_s + 4u * w + 8u, 0xFFFFFFFFu);
        xk::r_st32(bm_s + 4u * w + 12u, 0xFFFFFFFFu);
      }
      __syncwarp(gmask);
    }
    const uint16_t* toks = reinterpret_cast<const uint16_t*>(slot + sp::kSlotToks) + tskip;
    const uint32_t le = kAll >> (G - 1u - lane_u);  // the lanes up to and including this one
    const uint32_t nsteps = (len + G - 1u) / G;
    // two steps deep, as xk::ResolveBytes::resolve_sub: token two steps ahead, a far source one step ahead
    uint32_t tbase = 0;
    uint32_t s2, lit1, lit2;
    Base::masks(0, s2, lit1);
    uint32_t x1 = (uint32_t)__ldcg(toks + ((uint32_t)__popc(s2 & le) - 1u));
    tbase += (uint32_t)__popc(s2);
    Base::masks(1, s2, lit2);
    uint32_t x2 = (uint32_t)__ldcg(toks + (tbase + (uint32_t)__popc(s2 & le) - 1u));
    uint32_t pv = v0 + lane_u;
    // A distance that reaches below the output (possible in its first 32 KiB only) is looked for where a FAR source is
    // fetched, and for NEAR sources (the ring: reading it is harmless) while the range lies within the ring's reach of
    // the start.  Predicated, no branches: `bad` only ever collects.
    const bool early = v0 - lo_v < (uint32_t)RING;
    uint32_t fb1 = 0;
    {
      const bool farp = !((lit1 >> lane_u) & 1u) && x1 + 1u > kNear, inb = x1 + 1u <= pv - lo_v;
      if (farp && inb) fb1 = (uint32_t)__ldcg(vbase + (pv - (x1 + 1u)));
      bad |= (uint32_t)(farp && !inb);
    }
#pragma unroll 2
    for (uint32_t k = 0; k < nsteps; ++k) {
      const uint32_t lit = lit1, x = x1, fb = fb1;
      lit1 = lit2;
      x1 = x2;
      tbase += (uint32_t)__popc(s2);
      Base::masks(k + 2u, s2, lit2);
      x2 = (uint32_t)__ldcg(toks + (tbase + (uint32_t)__popc(s2 & le) - 1u));
      fb1 = 0;
      {
        const bool farp = !((lit1 >> lane_u) & 1u) && x1 + 1u > kNear, inb = x1 + 1u <= pv + G - lo_v;
        if (farp && inb) fb1 = (uint32_t)__ldcg(vbase + (pv + G - (x1 + 1u)));
        bad |= (uint32_t)(farp && !inb);
      }
      uint32_t b = x;
      if (lit != kAll) {                            // the step holds match bytes
        const uint32_t dist = x + 1u;
        const bool mb = !((lit >> lane_u) & 1u);
        bad |= (uint32_t)(early && mb && dist > pv - lo_v);
        const bool inside = mb && dist <= lane_u;   // the source is a byte of this very step
        const bool far = mb && dist > kNear;
        if (mb && !inside && !far) b = xk::r_ld8(ring_s + ((pv - dist) & RM));
        if (far) b = fb;
        if (__any_sync(gmask, inside)) {            // follow the chain of sources to a byte that is known
          uint32_t ptr = inside ? lane_u - dist : lane_u;
#pragma unroll
          for (int r = 1; r < 32; r <<= 1) ptr = __shfl_sync(gmask, ptr, (int)ptr);
          b = __shfl_sync(gmask, b, (int)ptr);
        }
      }
      xk::r_st8(ring_s + (pv & RM), b);
      pv += G;
      __syncwarp(gmask);
      if (((k + 1u) * G) % FLUSH == 0u && (k + 1u) * G < len) Base::flush(v0 + (k + 1u) * G);
    }
    Base::flush(v0 + len);
  }
};

// What the header of a Huffman-coded block yields.  ok = 0: a stored block, a bad header, bad code lengths -- the stream is
// the whole-stream kernel's.
struct BlockHead {
  uint32_t ok, last, first;   // first: bit position of the block's first symbol
};

// Block header at bit `bit` and the block's two decode tables, by the whole warp: every lane reads the same bits, the
// tables are built cooperatively (xk::warp_build_table).  Out of line, as is the checksum below: this kernel's warps are
// spread over very different code (header, lane decode, walk, resolve), and what does not sit in the instruction cache
// stalls every one of them -- the first version, everything inlined, lost more issue slots to instruction fetch than to
// any data dependency (ncu: 3.8 warps per issue waiting on "no instruction").
template <int LBITS, int LT, int DBITS, int DT, class WS>
__device__ __noinline__ BlockHead block_head(const uint8_t* src, uint32_t src_len, uint32_t bit, WS* wsp, int lane) {
  constexpr unsigned kFull = 0xFFFFFFFFu;
  WS& ws = *wsp;
  tk::TokLane<LBITS, LT, DBITS, DT> L;
  L.in = src;
  L.in_len = src_len;
  L.bits_init(bit >> 3);
  L.drop(bit & 7u);
  L.refill();
  BlockHead h{0u, 0u, 0u};
  h.last = L.take(1);
  const uint32_t type = L.take(2);
  if (type == 0u || type == 3u) return h;          // stored blocks (and bad headers) are the whole-stream kernel's
  uint32_t status = fl::kStatusOk;
  int nlen, ndist;
  if (type == 1u) {
    for (int i = lane; i < 288; i += 32) ws.sc.lens[i] = (uint8_t)dfl::fixed_ll_len(i);
    ws.sc.lens[288 + lane] = 5;
    nlen = 288;
    ndist = 32;
    __syncwarp(kFull);
  } else {
    nlen = (int)L.take(5) + 257;
    ndist = (int)L.take(5) + 1;
    const int ncode = (int)L.take(4) + 4;
    if (nlen > 286 || ndist > 30) status = fl::kStatusDataError;
    if (lane < 19) ws.sc.lens[lane] = 0;
    __syncwarp(kFull);
    for (int i = 0; i < ncode; ++i) {
      L.refill();
      const uint32_t v = L.take(3);
      if (lane == 0) ws.sc.lens[dfl::cl_order(i)] = (uint8_t)v;
    }
    __syncwarp(kFull);
    if (status == fl::kStatusOk)
      status = xk::warp_build_table<32>(ws.sc.lens, 19, fl::kCodeLen, ws.dt, 7, 128, ws.sc.d_count, ws.sc.d_first, ws.sc.d_offs,
                                        ws.sc.d_sorted, ws.cnt, ws.at, lane, kFull);
    if (status == fl::kStatusOk) {
      int i2 = 0, prev = 0;
      const int tot = nlen + ndist;
      while (i2 < tot) {
        L.refill();
        const uint32_t e = ws.dt[L.lo & 127u];
        if ((e & 15u) == 0) { status = fl::kStatusDataError; break; }
        L.drop(e & 15u);
        const int sym = (int)(e >> 4);
        int rep, val;
        if (sym < 16) { rep = 1; val = sym; prev = sym; }
        else if (sym == 16) {
          if (i2 == 0) { status = fl::kStatusDataError; break; }
          rep = 3 + (int)L.take(2); val = prev;
        } else if (sym == 17) { rep = 3 + (int)L.take(3); val = 0; prev = 0; }
        else { rep = 11 + (int)L.take(7); val = 0; prev = 0; }
        if (i2 + rep > tot) { status = fl::kStatusDataError; break; }
        for (int k = lane; k < rep; k += 32) ws.sc.lens[i2 + k] = (uint8_t)val;
        i2 += rep;
      }
      __syncwarp(kFull);
      if (status == fl::kStatusOk && (L.overrun() || ws.sc.lens[256] == 0)) status = fl::kStatusDataError;
    }
  }
  h.first = (uint32_t)(8ll * (long long)L.start_off + L.consumed_bits());
  __syncwarp(kFull);
  // the two tables of the block, one after the other through the same code (kind 1 = distances first: its lengths lie behind
  // the literal/length ones, and the code-length table it replaces is done with)
#pragma unroll 1
  for (int k = 0; k < 2 && status == fl::kStatusOk; ++k) {
    const bool d = k == 0;
    status = xk::warp_build_table<32>(d ? ws.sc.lens + nlen : ws.sc.lens, d ? ndist : nlen, d ? fl::kDist : fl::kLitLen, d ? ws.dt : ws.lt,
                                      d ? DBITS : LBITS, d ? DT : LT, d ? ws.sc.d_count : ws.sc.ll_count, d ? ws.sc.d_first : ws.sc.ll_first,
                                      d ? ws.sc.d_offs : ws.sc.ll_offs, d ? ws.sc.d_sorted : ws.sc.ll_sorted, ws.cnt, ws.at, lane, kFull);
  }
  h.ok = status == fl::kStatusOk ? 1u : 0u;
  return h;
}

__device__ __noinline__ uint64_t stream_checksum(const uint8_t* dst, uint32_t n, int type, const ik::CksSmem* ck, int lane) {
  inf::Group<32> g;
  g.lane = lane;
  g.mask = 0xFFFFFFFFu;
  return ik::group_checksum<32>(dst, n, type, ck, g);
}

template <int LBITS, int LT, int DBITS, int DT, int WARPS, int MIN_CTAS, int RING, int FLUSH>
__global__ void __launch_bounds__(WARPS * 32, MIN_CTAS)
    inflate_spec_kernel(const bitar_chunk* __restrict__ ops, bitar_result* __restrict__ results, const uint32_t* __restrict__ list,
                        xk::Counters* __restrict__ pc, uint32_t* __restrict__ declined, uint8_t* scratch, int checksum_type, uint32_t target) {
  using Lane = sp::SpecLane<LBITS, LT, DBITS, DT>;
  constexpr int kRing = RING, kFlush = FLUSH;
  using WS = SpecSmem<LT, DT, kRing>;
  using Res = SpecResolve<kRing, kFlush>;
  static_assert(sizeof(WS) % 16 == 0, "warp areas stay vector aligned");
  constexpr unsigned kFull = 0xFFFFFFFFu;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint32_t* dinfo = reinterpret_cast<uint32_t*>(smem_raw + (size_t)WARPS * sizeof(WS));
  ik::CksSmem* ck = reinterpret_cast<ik::CksSmem*>(dinfo + 32);
  if (threadIdx.x < 32) dinfo[threadIdx.x] = fl::dist_info((int)threadIdx.x);
  if (checksum_type != BITAR_CHECKSUM_NONE) {
    for (unsigned i = threadIdx.x; i < 256; i += blockDim.x) ck->crc_tab[i] = cks::crc_table_entry(i);
    if (threadIdx.x == 0) cks::crc_x2n_init(ck->x2n);
  }
  __syncthreads();

  const int lane = (int)(threadIdx.x & 31u);
  WS& ws = *reinterpret_cast<WS*>(smem_raw + (size_t)(threadIdx.x >> 5) * sizeof(WS));
  Lane L;
  L.bind(ws.lt, ws.dt, ws.lane_ring + lane * WS::kRingStride, dinfo, &ws.sc);
  Res R;
  R.gl = lane;
  R.gmask = kFull;
  R.ring_s = (uint32_t)__cvta_generic_to_shared(ws.ring);
  uint8_t* const slots = scratch + (size_t)(blockIdx.x * WARPS + (threadIdx.x >> 5)) * 32u * sp::kSlotBytes;
  uint8_t* const my_slot = slots + (size_t)lane * sp::kSlotBytes;
  const uint32_t n_list = pc->n_generic;

  for (;;) {
    uint32_t t = 0;
    if (lane == 0) t = atomicAdd(&pc->spec_next, 1u);
    t = __shfl_sync(kFull, t, 0);
    if (t >= 2u * n_list) break;
    const bool second = t >= n_list;                // longest first: two passes over the list (xk::first_pass_op)
    const uint32_t idx = list[second ? t - n_list : t];
    const bitar_chunk op = ops[idx];
    if (xk::first_pass_op(op.src_len, n_list, pc->sum_generic) == second) continue;
    const uint8_t* src = static_cast<const uint8_t*>(op.src);
    uint8_t* dst = static_cast<uint8_t*>(op.dst);
    const uint32_t src_len = op.src_len, cap = op.dst_cap;
    // (tiny streams are not worth a round; the bit positions of a stream fit 32 bits)
    bool ok = src != nullptr && dst != nullptr && src_len >= 64u && src_len < (1u << 28) && cap > 0u;
    const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 15u);
    R.begin(dst);
    R.lo_v = mis;
    R.bad = 0u;
    uint32_t total = 0, bit = 0, last = 0, tgt = target;
    while (ok && !last) {
      // ---- block header and tables ----
      const BlockHead bh = block_head<LBITS, LT, DBITS, DT, WS>(src, src_len, bit, &ws, lane);
      if (!bh.ok) {
        ok = false;
        break;
      }
      last = bh.last;
      uint32_t first = bh.first;
      // ---- rounds of 32 ranges ----
      for (;;) {
        const uint32_t B = sp::range_bits(first, src_len, total, cap, tgt);
        const unsigned long long start = (unsigned long long)first + (unsigned long long)lane * B;
        if (start < 8ull * src_len) L.start_spec(src, src_len, (uint32_t)start, (uint32_t)start + B, my_slot);
        else L.idle();
        // One loop, one copy of the step: iterations 0 .. kRec - 1 are the lock-step ones during which the lanes record
        // (nobody walks yet); every lane passes iteration kRec, where the records are complete and the successors' are
        // handed out; from there on a lane leaves when it is done.
#pragma unroll 1
        for (uint32_t i = 0; i <= sp::kRec || L.state != Lane::kDone; ++i) {
          if (i == sp::kRec) {
            __syncwarp(kFull);                       // the records are complete (and visible) before anybody walks
            uint32_t n_next = __shfl_down_sync(kFull, L.nrec, 1);
            if (lane == 31) n_next = 0;
            L.set_next(slots + (size_t)((lane + 1) & 31) * sp::kSlotBytes, n_next, lane < 31);
          }
          L.step(i >= sp::kRec);
        }
        __syncwarp(kFull);                           // the maps are complete before phase B reads them
        const unsigned synced = __ballot_sync(kFull, L.end_kind == sp::kEndSync);   // (lane 31 never is: it has no successor)
        const int m = __ffs((int)~synced) - 1;       // lanes 0 .. m are good
        uint32_t j = __shfl_up_sync(kFull, L.sync_j, 1);
        if (lane == 0) j = 0;
        sp::RangeOut ro{0u, 0u, 0u};
        if (lane <= m) ro = sp::range_out(L, j);
        uint32_t incl = ro.len;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const uint32_t v = __shfl_up_sync(kFull, incl, d);
          if (lane >= d) incl += v;
        }
        const uint32_t round_total = __shfl_sync(kFull, incl, m);
        const uint32_t ek = __shfl_sync(kFull, L.end_kind, m), eb = __shfl_sync(kFull, L.end_bit, m);
        if ((unsigned long long)total + round_total > cap || ek == sp::kEndBad || eb > 8u * src_len || (ek != sp::kEndEob && eb == first)) {
          ok = false;
          break;
        }
        if (ek == sp::kEndFull && tgt > 128u) tgt >>= 1;   // ranges too long for their slots: shorter ones from here on
        for (int r = 0; r <= m; ++r) {
          const uint32_t len_r = __shfl_sync(kFull, ro.len, r), off_r = __shfl_sync(kFull, incl - ro.len, r);
          const uint32_t ts = __shfl_sync(kFull, ro.tskip, r), bs = __shfl_sync(kFull, ro.bskip, r);
          if (len_r) R.resolve_range(slots + (size_t)r * sp::kSlotBytes, mis + total + off_r, len_r, ts, bs);
        }
        total += round_total;
        first = eb;
        if (ek == sp::kEndEob) break;
      }
      bit = first;
    }
    if (ok) R.finish(mis + total);
    if (__any_sync(kFull, R.bad != 0u)) ok = false;
    if (ok) {
      uint64_t sum = 0;
      if (checksum_type != BITAR_CHECKSUM_NONE) sum = stream_checksum(dst, total, checksum_type, ck, lane);
      if (lane == 0) {
        bitar_result out;
        out.produced = total;
        out.status = BITAR_OP_OK;
        out.checksum = sum;
        results[idx] = out;
      }
    } else if (lane == 0) {
      declined[atomicAdd(&pc->n_declined, 1u)] = idx;
    }
    __syncwarp(kFull);
  }
}

template <int LBITS, int LT, int DBITS, int DT, int WARPS, int MIN_CTAS, int RING, int FLUSH>
struct SpecConfig {
  static constexpr int kThreads = WARPS * 32;
  static constexpr size_t kSmem = (size_t)WARPS * sizeof(SpecSmem<LT, DT, RING>) + 32 * sizeof(uint32_t) + sizeof(ik::CksSmem);
  static int ctas_per_sm(int device) {
    static int per_device[64] = {0};
    int& c = per_device[device & 63];
    if (c == 0) {
      auto kern = inflate_spec_kernel<LBITS, LT, DBITS, DT, WARPS, MIN_CTAS, RING, FLUSH>;
      if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem) != cudaSuccess) return 0;
      cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c, kern, kThreads, kSmem) != cudaSuccess) return 0;
    }
    return c;
  }
  // bytes of scratch the largest grid needs: 32 slots per warp
  static size_t scratch_bytes(int device, int sm_count) {
    return (size_t)sm_count * (size_t)ctas_per_sm(device) * WARPS * 32u * sp::kSlotBytes;
  }
  // n_max: upper bound of the listed ops (the real count lives on the device)
  static cudaError_t launch(const bitar_chunk* ops, bitar_result* res, const uint32_t* list, xk::Counters* pc, uint32_t* declined,
                            uint8_t* scratch, int checksum_type, uint32_t target, uint32_t n_max, int device, int sm_count,
                            cudaStream_t stream) {
    const int c = ctas_per_sm(device);
    if (c < 1) return cudaErrorLaunchOutOfResources;
    uint32_t grid = (uint32_t)(sm_count * c);
    const uint32_t want = (n_max + WARPS - 1) / WARPS;
    if (want < grid) grid = want;
    if (grid == 0) return cudaSuccess;
    inflate_spec_kernel<LBITS, LT, DBITS, DT, WARPS, MIN_CTAS, RING, FLUSH><<<grid, kThreads, kSmem, stream>>>(ops, res, list, pc, declined, scratch,
                                                                                                  checksum_type, target);
    return cudaGetLastError();
  }
};

}  // namespace sk
}  // namespace bitar
