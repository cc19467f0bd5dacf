// inflate_lane_kernel.cuh -- K4 (+K5): lane-per-chunk raw DEFLATE inflate for sm_100a.
//
// A warp decodes a WAVE of 32 consecutive chunks, one per lane, through the bounded-step state machine
// of inflate_lane.h; waves are handed out by a global atomic counter.  Consecutive chunks of a columnar
// buffer have similar symbol counts, so the lanes of a wave finish close together.
//   * decode tables (u16) and a short output ring per lane in shared memory; 1.5 - 3 KiB per stream,
//     so 64 - 128 streams are resident per SM,
//   * input words through the read-only path with one word of prefetch per lane,
//   * output leaves the SM as 16-byte vector stores, one per lane and step,
//   * optional CRC-32 / Adler-32 of the produced bytes once the wave is done (lanes converged).
//
// Replaces: rte_compressdev decompress ops assembled at /root/reference/src/memory.cc:432-505 and
// executed behind src/device.cc:464-535 (dst segment i at out + i*S, src/memory.cc:482-493).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/bitar_cuda.h"
#include "checksum.h"
#include "inflate_lane.h"

namespace bitar {
namespace ilk {

template <int LBITS, int LT, int DBITS, int DT, int RING, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
    inflate_lane_kernel(const bitar_chunk* __restrict__ ops, uint32_t n_ops, bitar_result* __restrict__ results,
                        unsigned int* __restrict__ counter, infl::LaneScratch* __restrict__ scratch,
                        int checksum_type) {
  using LaneT = infl::Lane<LBITS, LT, DBITS, DT, RING>;
  constexpr int kStride = infl::LaneSmem<LBITS, LT, DBITS, DT, RING>::kStride;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint32_t* crc_tab = reinterpret_cast<uint32_t*>(smem_raw + (size_t)WARPS * 32 * kStride);

  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
#if defined(BITAR_LANE_DEBUG)
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    infl::g_dbg[10] = (uint32_t)__cvta_generic_to_shared(smem_raw);
    infl::g_dbg[11] = infl::g_dbg[10] + (uint32_t)(WARPS * 32 * kStride + 1024);
  }
  __syncthreads();
#endif
  LaneT L;
  L.bind(smem_raw + (size_t)threadIdx.x * kStride, scratch + ((size_t)blockIdx.x * WARPS * 32 + threadIdx.x));

  if (checksum_type & BITAR_CHECKSUM_CRC32) {
    for (unsigned i = threadIdx.x; i < 256; i += WARPS * 32) crc_tab[i] = cks::crc_table_entry(i);
    __syncthreads();
  }

  const uint32_t static_waves = gridDim.x * WARPS;
  uint32_t wave = blockIdx.x * WARPS + warp;
  while ((uint64_t)wave * 32u < n_ops) {
    const uint32_t idx = wave * 32u + lane;
    const bool active = idx < n_ops;
    bitar_chunk op;
    if (active) {
      op = ops[idx];
      L.start(static_cast<const uint8_t*>(op.src), op.src_len, static_cast<uint8_t*>(op.dst), op.dst_cap);
    } else {
      L.state = infl::kDone;
    }
    while (__any_sync(0xFFFFFFFFu, L.state != infl::kDone)) {
      L.step_pre();
      const uint32_t trip = __reduce_max_sync(0xFFFFFFFFu, L.want_copy());
      L.step_post(trip);
    }
    uint64_t sum = 0;
    if (checksum_type != BITAR_CHECKSUM_NONE) {
      // all lanes walk their own output (L1/L2 resident) in lock step
      const uint32_t nb = (active && L.status == infl::kStatusOk) ? L.produced() : 0u;
      const uint8_t* p = active ? static_cast<const uint8_t*>(op.dst) : nullptr;
      uint32_t crc = 0xFFFFFFFFu, a = 1, b = 0;
      const bool want_crc = checksum_type & BITAR_CHECKSUM_CRC32, want_adler = checksum_type & BITAR_CHECKSUM_ADLER32;
      for (uint32_t i = 0; i < nb; ++i) {
        const uint32_t byte = *reinterpret_cast<const volatile uint8_t*>(p + i);
        if (want_crc) crc = crc_tab[(crc ^ byte) & 0xFFu] ^ (crc >> 8);
        if (want_adler) {
          a += byte;
          b += a;
          if ((i & 0xFFFu) == 0xFFFu) {   // 4096 * 255 * 4096 / 2 < 2^32
            a %= cks::kAdlerMod;
            b %= cks::kAdlerMod;
          }
        }
      }
      a %= cks::kAdlerMod;
      b %= cks::kAdlerMod;
      sum = cks::pack(want_crc ? (nb ? crc ^ 0xFFFFFFFFu : 0u) : 0u, want_adler ? ((b << 16) | a) : 0u);
    }
    if (active) {
      bitar_result r;
      r.produced = L.produced();
      r.status = L.status;
      r.checksum = sum;
      results[idx] = r;
    }
    uint32_t next = 0;
    if (lane == 0) next = static_waves + atomicAdd(counter, 1u);
    wave = __shfl_sync(0xFFFFFFFFu, next, 0);
  }
}

template <int LBITS, int LT, int DBITS, int DT, int RING, int WARPS>
struct LaneConfig {
  static constexpr int kThreads = WARPS * 32;
  static constexpr size_t kSmem = (size_t)WARPS * 32 * infl::LaneSmem<LBITS, LT, DBITS, DT, RING>::kStride + 1024;
  // resident CTAs per SM on `device` (0 on error)
  static int ctas_per_sm(int device) {
    static int per_device[64] = {0};
    int& c = per_device[device & 63];
    if (c == 0) {
      auto kern = inflate_lane_kernel<LBITS, LT, DBITS, DT, RING, WARPS>;
      if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem) != cudaSuccess) return 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c, kern, kThreads, kSmem) != cudaSuccess) return 0;
    }
    return c;
  }
  static size_t scratch_bytes(int device, int sm_count) {
    return (size_t)sm_count * ctas_per_sm(device) * kThreads * sizeof(infl::LaneScratch);
  }
  static cudaError_t launch(const bitar_chunk* ops, uint32_t n, bitar_result* res, unsigned int* counter,
                            void* scratch, int checksum_type, int device, int sm_count, cudaStream_t stream) {
    const int c = ctas_per_sm(device);
    if (c < 1) return cudaErrorLaunchOutOfResources;
    const uint32_t waves = (n + 31u) / 32u;
    uint32_t grid = (uint32_t)(sm_count * c);
    const uint32_t want = (waves + WARPS - 1) / WARPS;
    if (want < grid) grid = want;
    if (grid == 0) return cudaSuccess;
    inflate_lane_kernel<LBITS, LT, DBITS, DT, RING, WARPS><<<grid, kThreads, kSmem, stream>>>(
        ops, n, res, counter, static_cast<infl::LaneScratch*>(scratch), checksum_type);
    return cudaGetLastError();
  }
};

}  // namespace ilk
}  // namespace bitar
