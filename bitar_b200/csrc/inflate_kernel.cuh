// inflate_kernel.cuh -- K4 (+K5): sub-warp-cooperative raw DEFLATE inflate for sm_100a.
//
// Grid: persistent CTAs (a multiple of the SM count); every group of G lanes pulls chunk indices
// from a global atomic counter, so long and short streams balance dynamically.  Per group the
// decode tables, the canonical-code side arrays and the output ring live in shared memory
// (inflate_core.h::GroupSmem); input words are read through the read-only path with one word of
// prefetch, output leaves the SM as aligned 16-byte vector stores.
//
// Replaces: rte_compressdev decompress ops assembled at /root/reference/src/memory.cc:432-505 and
// executed behind src/device.cc:464-535 (dst segment i at out + i*S, src/memory.cc:482-493).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/bitar_cuda.h"
#include "checksum.h"
#include "inflate_core.h"

namespace bitar {
namespace ik {

// CTA-shared checksum tables, placed after the per-group areas
struct CksSmem {
  uint32_t crc_tab[256];
  uint32_t x2n[32];
};

// Checksum of a finished chunk, computed by the G lanes of the group over the bytes they just wrote
// (still L1/L2 resident).  Lane j takes the j-th contiguous slice.
template <int G>
__device__ __forceinline__ uint64_t group_checksum(const uint8_t* p, uint32_t n, int type,
                                                   const CksSmem* ck, const inf::Group<G>& g) {
  g.sync();
  uint32_t per = (n + G - 1) / G;
  uint32_t lo = min(n, per * (uint32_t)g.lane), hi = min(n, lo + per);
  uint32_t state = lo == 0 ? 0xFFFFFFFFu : 0u, s1 = 0;
  uint64_t s2 = 0;  // weights reach the slice length (128 KiB for 1 MiB chunks at G = 8): keep 64 bit
  const bool want_crc = type & BITAR_CHECKSUM_CRC32, want_adler = type & BITAR_CHECKSUM_ADLER32;
  for (uint32_t i = lo; i < hi; ++i) {
    uint32_t b = *reinterpret_cast<const volatile uint8_t*>(p + i);
    if (want_crc) state = cks::crc_step(state, b, ck->crc_tab);
    s1 += b;
    s2 += (uint64_t)(hi - i) * b;
  }
  uint32_t tail = n - hi;
  uint32_t crc_part = want_crc ? cks::crc_contrib(state, tail, ck->x2n) : 0u;
  if (lo >= hi && lo != 0) crc_part = 0;  // empty slice contributes nothing
  uint32_t b_part = want_adler ? cks::adler_b_contrib(s1 % cks::kAdlerMod, (uint32_t)(s2 % cks::kAdlerMod), tail % cks::kAdlerMod) : 0u;
  uint32_t a_part = s1 % cks::kAdlerMod;
  const unsigned lane_in_warp = threadIdx.x & 31u;
  const unsigned base = lane_in_warp & ~(unsigned)(G - 1);
#pragma unroll
  for (int off = G / 2; off > 0; off >>= 1) {
    unsigned src = base + ((lane_in_warp - base) ^ (unsigned)off);
    crc_part ^= __shfl_sync(g.mask, crc_part, (int)src);
    a_part += __shfl_sync(g.mask, a_part, (int)src);
    b_part += __shfl_sync(g.mask, b_part, (int)src);
  }
  uint32_t crc = want_crc ? (crc_part ^ 0xFFFFFFFFu) : 0u;
  if (n == 0) crc = 0;
  uint32_t adler = want_adler ? cks::adler_finish(a_part, b_part, n) : 0u;
  return cks::pack(crc, adler);
}

template <int G, int LBITS, int DBITS, int RING, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
    inflate_kernel(const bitar_chunk* __restrict__ ops, uint32_t n_ops, bitar_result* __restrict__ results,
                   unsigned int* __restrict__ counter, int checksum_type, const uint32_t* __restrict__ list,
                   const unsigned int* __restrict__ n_list) {
  using GS = inf::GroupSmem<LBITS, DBITS, RING>;
  constexpr int kGroupsPerWarp = 32 / G;
  constexpr int kGroupsPerCta = WARPS * kGroupsPerWarp;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  GS* all = reinterpret_cast<GS*>(smem_raw);
  CksSmem* ck = reinterpret_cast<CksSmem*>(smem_raw + sizeof(GS) * kGroupsPerCta);

  const unsigned lane_in_warp = threadIdx.x & 31u;
  const unsigned warp = threadIdx.x >> 5;
  const unsigned group_in_warp = lane_in_warp / G;
  GS* sm = all + warp * kGroupsPerWarp + group_in_warp;
  inf::Group<G> g;
  g.lane = (int)(lane_in_warp % G);
  g.mask = G == 32 ? 0xFFFFFFFFu : (((1u << G) - 1u) << (group_in_warp * G));

  if (checksum_type != BITAR_CHECKSUM_NONE) {
    for (unsigned i = threadIdx.x; i < 256; i += blockDim.x) ck->crc_tab[i] = cks::crc_table_entry(i);
    if (threadIdx.x == 0) cks::crc_x2n_init(ck->x2n);
    __syncthreads();
  }

  // with a list (written by the plan kernel of inflate_indexed_kernel.cuh) only the listed ops are decoded
  if (list) n_ops = *n_list;
  const uint32_t total_groups = gridDim.x * kGroupsPerCta;
  uint32_t slot = blockIdx.x * kGroupsPerCta + warp * kGroupsPerWarp + group_in_warp;
  while (slot < n_ops) {
    const uint32_t idx = list ? list[slot] : slot;
    const bitar_chunk op = ops[idx];
    inf::ChunkResult r = inf::inflate_chunk<G, LBITS, DBITS, RING>(
        static_cast<const uint8_t*>(op.src), op.src_len, static_cast<uint8_t*>(op.dst), op.dst_cap, sm, g);
    uint64_t sum = 0;
    if (checksum_type != BITAR_CHECKSUM_NONE && r.status == inf::kStatusOk)
      sum = group_checksum<G>(static_cast<const uint8_t*>(op.dst), r.produced, checksum_type, ck, g);
    if (g.lane == 0) {
      bitar_result out;
      out.produced = r.produced;
      out.status = r.status;
      out.checksum = sum;
      results[idx] = out;
    }
    uint32_t next = 0;
    if (g.lane == 0) next = total_groups + atomicAdd(counter, 1u);
    slot = __shfl_sync(g.mask, next, (int)(group_in_warp * G));
  }
}

template <int G, int LBITS, int DBITS, int RING, int WARPS>
struct InflateConfig {
  static constexpr int kThreads = WARPS * 32;
  static constexpr size_t kSmem =
      sizeof(inf::GroupSmem<LBITS, DBITS, RING>) * (WARPS * (32 / G)) + sizeof(CksSmem);
  static cudaError_t launch(const bitar_chunk* ops, uint32_t n, bitar_result* res, unsigned int* counter,
                            int checksum_type, int device, int sm_count, cudaStream_t stream,
                            const uint32_t* list = nullptr, const unsigned int* n_list = nullptr) {
    auto kern = inflate_kernel<G, LBITS, DBITS, RING, WARPS>;
    static int per_device[64] = {0};   // resident CTAs per SM, resolved once per device
    int& ctas_per_sm = per_device[device & 63];
    if (ctas_per_sm == 0) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem);
      if (e != cudaSuccess) return e;
      // every kernel of the library asks for the same shared-memory carve-out: an SM has to drain before it can
      // change the split, which serialises kernels of different queue pairs (and of one stream) otherwise
      cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
      e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, kThreads, kSmem);
      if (e != cudaSuccess) return e;
      if (ctas_per_sm < 1) return cudaErrorLaunchOutOfResources;
    }
    constexpr uint32_t groups_per_cta = WARPS * (32 / G);
    uint32_t want = (n + groups_per_cta - 1) / groups_per_cta;
    uint32_t grid = (uint32_t)(sm_count * ctas_per_sm);
    if (want < grid) grid = want;
    if (grid == 0) return cudaSuccess;
    kern<<<grid, kThreads, kSmem, stream>>>(ops, n, res, counter, checksum_type, list, n_list);
    return cudaGetLastError();
  }
};

}  // namespace ik
}  // namespace bitar
