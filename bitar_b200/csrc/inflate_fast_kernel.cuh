// inflate_fast_kernel.cuh -- K4 (+K5): lane-per-stream raw DEFLATE inflate for sm_100a.
//
// One persistent CTA per SM; every lane decodes one stream at a time through fl::FastLane::step()
// (inflate_fast.h) and fetches its next stream from a global atomic counter the moment it is done, so
// lanes never wait for the slowest stream of a warp.  The first assignment interleaves warps across
// CTAs: consecutive streams (same column, similar symbol mix -> converged lanes) share a warp, while
// every SM receives warps from all over the buffer (balanced column mix per SM).
//
// Shared memory per lane: u16 decode tables + output ring (LaneLayout); per CTA: distance-symbol info and
// the CRC-32 slicing tables.  Input words arrive through the read-only path with one word of prefetch,
// output leaves as aligned 16-byte vector stores.
//
// Replaces: rte_compressdev decompress ops assembled at /root/reference/src/memory.cc:432-505 and
// executed behind src/device.cc:464-535 (dst segment i at out + i*S, src/memory.cc:482-493).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/bitar_cuda.h"
#include "inflate_fast.h"

namespace bitar {
namespace fk {

template <int LBITS, int LT, int DBITS, int DT, int RING, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1)
    inflate_fast_kernel(const bitar_chunk* __restrict__ ops, uint32_t n_ops, bitar_result* __restrict__ results,
                        unsigned int* __restrict__ counter, fl::LaneScratch* __restrict__ scratch, int checksum_type) {
  using Lane = fl::FastLane<LBITS, LT, DBITS, DT, RING>;
  constexpr int kStride = fl::LaneLayout<LBITS, LT, DBITS, DT, RING>::kStride;
  constexpr int kThreads = WARPS * 32;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  fl::CtaTables* cta = reinterpret_cast<fl::CtaTables*>(smem_raw + (size_t)kThreads * kStride);

  if (threadIdx.x < 32) cta->dinfo[threadIdx.x] = fl::dist_info((int)threadIdx.x);
  if (checksum_type & BITAR_CHECKSUM_CRC32) {
    for (unsigned i = threadIdx.x; i < 256; i += kThreads) {
      uint32_t c = cks::crc_table_entry(i);
      cta->crc[0][i] = c;
    }
    __syncthreads();
    for (unsigned i = threadIdx.x; i < 256; i += kThreads) {
      uint32_t c = cta->crc[0][i];
      for (int k = 1; k < 4; ++k) {
        c = (c >> 8) ^ cta->crc[0][c & 0xFFu];
        cta->crc[k][i] = c;
      }
    }
  }
  __syncthreads();

  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  Lane L;
  L.bind(smem_raw + (size_t)threadIdx.x * kStride, cta, scratch + ((size_t)blockIdx.x * kThreads + threadIdx.x),
         (uint32_t)checksum_type);

  const uint32_t total_lanes = gridDim.x * kThreads;
  uint32_t idx = (warp * gridDim.x + blockIdx.x) * 32u + lane;   // first assignment: warps interleaved over CTAs
  uint32_t cur = 0xFFFFFFFFu;
  for (;;) {
    if (L.state == Lane::kDone) {
      if (cur != 0xFFFFFFFFu) {
        bitar_result r;
        r.produced = L.produced();
        r.status = L.status;
        r.checksum = (checksum_type != BITAR_CHECKSUM_NONE && L.status == fl::kStatusOk) ? L.checksum() : 0ull;
        results[cur] = r;
        idx = total_lanes + atomicAdd(counter, 1u);
      }
      if (idx >= n_ops) break;
      cur = idx;
      const bitar_chunk op = ops[cur];
      L.start(static_cast<const uint8_t*>(op.src), op.src_len, static_cast<uint8_t*>(op.dst), op.dst_cap);
    }
    L.step();
  }
}

template <int LBITS, int LT, int DBITS, int DT, int RING, int WARPS>
struct FastConfig {
  static constexpr int kThreads = WARPS * 32;
  static constexpr size_t kSmem = (size_t)kThreads * fl::LaneLayout<LBITS, LT, DBITS, DT, RING>::kStride + sizeof(fl::CtaTables);
  static int ctas_per_sm(int device) {
    static int per_device[64] = {0};
    int& c = per_device[device & 63];
    if (c == 0) {
      auto kern = inflate_fast_kernel<LBITS, LT, DBITS, DT, RING, WARPS>;
      if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem) != cudaSuccess) return 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c, kern, kThreads, kSmem) != cudaSuccess) return 0;
    }
    return c;
  }
  static size_t scratch_bytes(int device, int sm_count) {
    return (size_t)sm_count * ctas_per_sm(device) * kThreads * sizeof(fl::LaneScratch);
  }
  static cudaError_t launch(const bitar_chunk* ops, uint32_t n, bitar_result* res, unsigned int* counter, void* scratch,
                            int checksum_type, int device, int sm_count, cudaStream_t stream) {
    const int c = ctas_per_sm(device);
    if (c < 1) return cudaErrorLaunchOutOfResources;
    uint32_t grid = (uint32_t)(sm_count * c);
    const uint32_t want = (n + kThreads - 1) / kThreads;
    if (want < grid) grid = want;
    if (grid == 0) return cudaSuccess;
    inflate_fast_kernel<LBITS, LT, DBITS, DT, RING, WARPS><<<grid, kThreads, kSmem, stream>>>(
        ops, n, res, counter, static_cast<fl::LaneScratch*>(scratch), checksum_type);
    return cudaGetLastError();
  }
};

}  // namespace fk
}  // namespace bitar
