// capi.cu -- implementation of include/bitar_cuda.h: CUDA device probe, queue pairs (one stream
// each), the output-slot pool, memory allocation and the kernel launches.
//
// Mapping to the reference (/root/reference):
//   bitar_dev            <- CompressDevice state + DeviceMemory        (src/device.cc:114-154, src/memory.cc:120-235)
//   QueuePair            <- QueuePairMemory: op/mbuf pools, pending ops (src/memory.cc:237-348, 507-575)
//   bitar_qp_deflate/... <- AssembleFrom + EnqueueBurst/DequeueBurst    (src/memory.cc:350-505, src/device.cc:464-535)
//   slot pool            <- DeviceMemory::Take/Put                      (src/memory.cc:160-209)
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../../include/bitar_cuda.h"
#include "deflate_kernel.cuh"      // bitar::dk: 16 warps per CTA, 2 CTAs per SM, chunks up to 1 MiB
#define BITAR_DK_NS dks            // bitar::dks: 4 warps per CTA, 6 .. 8 CTAs per SM, chunks up to 16 KiB
#define BITAR_DK_WARPS 4
#define BITAR_DK_BLOCK_MAX 16384
#define BITAR_DK_MIN_CTAS 8
#include "deflate_kernel.cuh"
#include "inflate_kernel.cuh"
#include "inflate_tok_kernel.cuh"
#include "inflate_spec_kernel.cuh"

namespace {

thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};
std::atomic<int> g_inflate_variant{-1};
std::atomic<int> g_stage_strided{1};                        // tests / A-B runs: 0 = always gather with the kernel
std::atomic<size_t> g_stage_batch_bytes{0};                 // tests: least bytes per batch of a staged inflate call (0 = default)
std::atomic<unsigned long long*> g_deflate_prof{nullptr};  // device buffer of 8 phase counters (debug)

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
  return code;
}
#define CU_TRY(expr, code)                                                                   \
  do {                                                                                       \
    cudaError_t e_ = (expr);                                                                 \
    if (e_ != cudaSuccess) return fail(code, "%s: %s", #expr, cudaGetErrorString(e_));       \
  } while (0)

// Up to 16 streams per device (8 queue pairs, each with a copy-back stream for staged calls): more hardware queues
// than the driver's default 8, or streams sharing one serialise.  Takes effect when the library is loaded before the
// process initialises CUDA (the env var is read then); an existing setting is left alone.
struct ConnectionsDefault {
  ConnectionsDefault() { setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0); }
} g_connections_default;

constexpr uint32_t kCountersPerBatch = 16;             // device work counters of one batch (xk::Counters)
constexpr uint32_t kMaxStageBatches = 8;               // measured: ~8 batches of >= 16 MiB per call, whatever its size
constexpr uint32_t kStageLanes = 3;                     // + the queue pair's own stream: 8 queue pairs fit 32 hardware queues
constexpr size_t kStageBatchBytes = (size_t)16 << 20;   // least inflated bytes per batch of a staged call

struct QueuePair {
  cudaStream_t stream = nullptr;
  cudaEvent_t ev_start = nullptr, ev_k0 = nullptr, ev_k1 = nullptr, ev_stop = nullptr;
  bitar_chunk* h_ops = nullptr;    // pinned descriptor ring
  bitar_chunk* d_ops = nullptr;    // device descriptor ring
  bitar_result* h_res = nullptr;   // pinned results
  bitar_result* d_res = nullptr;
  uint32_t cap = 0;
  unsigned int* d_counter = nullptr;
  uint32_t* d_tokens = nullptr;    // deflate token scratch (grid * 64 Ki u32), allocated on first use
  uint16_t* d_far = nullptr;       // deflate far-candidate scratch (grid * 64 Ki u16)
  bitar::xk::Task* d_tasks = nullptr;   // indexed inflate: one task per 64 KiB block
  size_t tasks_cap = 0;
  uint8_t* d_units = nullptr;           //   token units between the two phases: `subs` slots per resident group of lanes
  size_t units_cap = 0;
  uint32_t* d_generic = nullptr;        // ops without an index: the speculative kernel's list
  uint32_t* d_declined = nullptr;       // what that kernel leaves to the whole-stream kernel
  uint32_t* d_indexed = nullptr;        // ops on the two-phase path
  size_t generic_cap = 0;
  // inflate with host (pinned / registered) buffers: stage through device memory
  bitar_chunk* h_orig = nullptr;        // the caller's pointers (pinned), n entries
  bitar_chunk* d_orig = nullptr;
  uint32_t orig_cap = 0;
  uint8_t* d_stage_in = nullptr;
  size_t stage_in_cap = 0;
  uint8_t* d_stage_out = nullptr;
  size_t stage_out_cap = 0;
  bool stage_src = false, stage_dst = false;
  // deflate of a host-resident buffer: the input goes to the device stage in pieces by the copy engine (its own stream),
  // the kernel of piece b starts when piece b has arrived
  bool dstage = false;
  const uint8_t* dstage_base = nullptr;
  size_t dstage_total = 0;
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_copy[kMaxStageBatches] = {};
  bool stage_dst_contig = false;        // destinations form one range of equal-capacity segments (Decompress())
  size_t stage_out_bytes = 0;           // sum of the call's destination capacities
  uint32_t nb = 1, per = 0;             // batches of the call, ops per batch
  // compressed inputs at a constant stride (pool slots handed out by take_n: Compress() -> Decompress()): a batch's
  // gather is one strided copy-engine transfer instead of a kernel reading host memory
  bool stage_src_strided = false;
  size_t src_stride = 0;

  uint32_t batch_pitch[kMaxStageBatches] = {};
  size_t batch_in_off[kMaxStageBatches] = {};
  // staged calls run in batches spread over kStageLanes extra streams: one batch is gather -> inflate -> copy-back
  // in order on its lane, the lanes overlap each other (PCIe both ways and the SMs busy at once)
  cudaStream_t lane[kStageLanes] = {};
  cudaEvent_t ev_lane[kStageLanes] = {};
  cudaEvent_t ev_fork = nullptr;
  bitar_result* user_out = nullptr;
  uint32_t pending_n = 0;
  std::atomic<int> busy{0};
  std::atomic<int> last_status{0};
  bool timed = false;
};

}  // namespace

struct bitar_dev {
  int id = 0;
  int sm_count = 0;
  bitar_cfg cfg{};
  std::vector<QueuePair*> qps;
  int deflate_grid = 0;
  int deflate_grid_small = 0;       // largest grid of the small-chunk instance (bitar::dks)
  int deflate_grid_override = 0;    // BITAR_DEBUG_DEFLATE_GRID (16-warp instance only)
  // slot pool
  std::mutex mu;
  std::vector<void*> slabs;
  std::vector<void*> free_slots;                 // LIFO
  std::unordered_set<const void*> occupied;
  // chained segments (max_sgl_segs = k > 1): the pool hands out GROUPS of k slots that lie back to back (a stream's
  // destination is one range); free_slots then holds group bases, and a group returns when all its slots came back
  uint32_t group = 1;
  std::vector<std::pair<uint8_t*, size_t>> slab_ranges;
  std::unordered_map<const void*, uint32_t> group_out;   // group base -> slots of it that are out
  uint32_t slot_stride = 0;
  uint32_t grow_warned = 0;
};

namespace {

constexpr uint32_t kRteMaxMemzone = 2560;  // RTE_MAX_MEMZONE, the reference's default pool size

uint32_t ref_compressed_seg_size(uint32_t seg) {  // src/config.cc:59-73 with its 16-bit arithmetic
  if (seg == 0 || seg > 65535u) return 0;
  uint32_t lower = seg << 1, num = 65536u;
  while ((num & lower) == 0) num >>= 1;
  return num > 32768u ? (uint32_t)((double)seg * 1.1) : num;
}
uint32_t stored_bound(uint32_t seg) { return seg + 5u * ((seg + 65534u) / 65535u); }

int alloc_kind(int kind, int device, size_t size, void** out) {
  *out = nullptr;
  if (size == 0) size = 1;
  if (kind == BITAR_MEM_DEVICE) {
    CU_TRY(cudaSetDevice(device), BITAR_E_INVALID);
    cudaError_t e = cudaMallocAsync(out, size, cudaStreamPerThread);
    if (e == cudaSuccess) e = cudaStreamSynchronize(cudaStreamPerThread);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return fail(BITAR_E_OUT_OF_MEMORY, "cudaMallocAsync(%zu) on device %d: %s", size, device, cudaGetErrorString(e));
    }
    return BITAR_OK;
  }
  if (kind == BITAR_MEM_PINNED) {
    if (device >= 0) CU_TRY(cudaSetDevice(device), BITAR_E_INVALID);
    cudaError_t e = cudaHostAlloc(out, size, cudaHostAllocPortable | cudaHostAllocMapped);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return fail(BITAR_E_OUT_OF_MEMORY, "cudaHostAlloc(%zu): %s", size, cudaGetErrorString(e));
    }
    return BITAR_OK;
  }
  return fail(BITAR_E_INVALID, "unknown memory kind %d", kind);
}

int free_kind(int kind, int device, void* p) {
  if (!p) return BITAR_OK;
  if (kind == BITAR_MEM_DEVICE) {
    CU_TRY(cudaSetDevice(device), BITAR_E_INVALID);
    CU_TRY(cudaFreeAsync(p, cudaStreamPerThread), BITAR_E_INVALID);
    CU_TRY(cudaStreamSynchronize(cudaStreamPerThread), BITAR_E_IO_ERROR);
    return BITAR_OK;
  }
  if (kind == BITAR_MEM_PINNED) {
    CU_TRY(cudaFreeHost(p), BITAR_E_INVALID);
    return BITAR_OK;
  }
  return fail(BITAR_E_INVALID, "unknown memory kind %d", kind);
}

// grow the slot pool by `count` slots carved from one slab (caller holds dev->mu)
int pool_grow(bitar_dev* dev, uint32_t count) {
  void* slab = nullptr;
  count = (count + dev->group - 1) / dev->group * dev->group;
  int rc = alloc_kind(dev->cfg.slot_mem_kind, dev->id, (size_t)count * dev->slot_stride, &slab);
  if (rc) return rc;
  dev->slabs.push_back(slab);
  dev->slab_ranges.emplace_back(static_cast<uint8_t*>(slab), (size_t)count * dev->slot_stride);
  if (dev->group > 1) {   // group bases, ascending addresses out first
    for (uint32_t i = count / dev->group; i-- > 0;)
      dev->free_slots.push_back(static_cast<uint8_t*>(slab) + (size_t)i * dev->group * dev->slot_stride);
    return BITAR_OK;
  }
  // push in reverse so that Take() hands out ascending addresses (contiguous runs for take_n)
  for (uint32_t i = count; i-- > 0;) dev->free_slots.push_back(static_cast<uint8_t*>(slab) + (size_t)i * dev->slot_stride);
  return BITAR_OK;
}

// chained segments: take n slots as ceil(n / group) groups (caller holds dev->mu); the last group may hand out fewer
int pool_take_groups(bitar_dev* dev, uint32_t n, void** slots) {
  const uint32_t k = dev->group, need = (n + k - 1) / k;
  if (dev->free_slots.size() < need) {
    if (dev->grow_warned++ % 32 == 0)
      fprintf(stderr, "[bitar] WARNING: allocating output slots in the critical path (performance will be impacted)\n");
    std::vector<void*> old;
    old.swap(dev->free_slots);
    int rc = pool_grow(dev, (need - (uint32_t)old.size()) * k);
    if (rc) {
      dev->free_slots.swap(old);
      return rc;
    }
    dev->free_slots.insert(dev->free_slots.begin(), old.begin(), old.end());
  }
  for (uint32_t g = 0, i = 0; g < need; ++g) {
    uint8_t* base = static_cast<uint8_t*>(dev->free_slots.back());
    dev->free_slots.pop_back();
    const uint32_t cnt = n - i < k ? n - i : k;
    dev->group_out[base] = cnt;
    for (uint32_t j = 0; j < cnt; ++j, ++i) {
      slots[i] = base + (size_t)j * dev->slot_stride;
      dev->occupied.insert(slots[i]);
    }
  }
  return BITAR_OK;
}
// chained segments: a slot comes back; its group is free again when all of its slots are (caller holds dev->mu)
int pool_put_grouped(bitar_dev* dev, const void* addr) {
  if (dev->occupied.erase(addr) == 0) return 0;
  const uint8_t* a = static_cast<const uint8_t*>(addr);
  for (const auto& r : dev->slab_ranges)
    if (a >= r.first && a < r.first + r.second) {
      const size_t gbytes = (size_t)dev->group * dev->slot_stride;
      const uint8_t* base = r.first + (size_t)(a - r.first) / gbytes * gbytes;
      auto it = dev->group_out.find(base);
      if (it != dev->group_out.end() && --it->second == 0) {
        dev->group_out.erase(it);
        dev->free_slots.push_back(const_cast<uint8_t*>(base));
      }
      break;
    }
  return 1;
}

int qp_reserve(bitar_dev* dev, QueuePair* q, uint32_t n) {
  if (n <= q->cap) return BITAR_OK;
  uint32_t cap = q->cap ? q->cap : 1024;
  while (cap < n) cap *= 2;
  CU_TRY(cudaStreamSynchronize(q->stream), BITAR_E_IO_ERROR);
  if (q->h_ops) cudaFreeHost(q->h_ops);
  if (q->h_res) cudaFreeHost(q->h_res);
  if (q->d_ops) cudaFree(q->d_ops);
  if (q->d_res) cudaFree(q->d_res);
  q->h_ops = nullptr; q->h_res = nullptr; q->d_ops = nullptr; q->d_res = nullptr;
  q->cap = 0;
  CU_TRY(cudaHostAlloc((void**)&q->h_ops, (size_t)cap * sizeof(bitar_chunk), cudaHostAllocPortable | cudaHostAllocMapped), BITAR_E_OUT_OF_MEMORY);
  CU_TRY(cudaHostAlloc((void**)&q->h_res, (size_t)cap * sizeof(bitar_result), cudaHostAllocPortable | cudaHostAllocMapped), BITAR_E_OUT_OF_MEMORY);
  CU_TRY(cudaMalloc((void**)&q->d_ops, (size_t)cap * sizeof(bitar_chunk)), BITAR_E_OUT_OF_MEMORY);
  CU_TRY(cudaMalloc((void**)&q->d_res, (size_t)cap * sizeof(bitar_result)), BITAR_E_OUT_OF_MEMORY);
  q->cap = cap;
  (void)dev;
  return BITAR_OK;
}

// Runs on a CUDA driver thread after the result download: publish results, mark the QP idle.
void CUDART_CB qp_finish(void* arg) {
  QueuePair* q = static_cast<QueuePair*>(arg);
  int bad = 0;
  for (uint32_t i = 0; i < q->pending_n; ++i)
    if (q->h_res[i].status != BITAR_OP_OK) bad = 1;
  if (q->user_out) memcpy(q->user_out, q->h_res, (size_t)q->pending_n * sizeof(bitar_result));
  q->last_status.store(bad ? BITAR_E_IO_ERROR : BITAR_OK, std::memory_order_release);
  q->busy.store(0, std::memory_order_release);
}

// ---- staging of host-resident inflate buffers ---------------------------------------------------------
// A DEFLATE decoder reads its input in dependent 4-byte steps and writes 16 bytes at a time: over PCIe that
// is latency-bound (measured 2 GB/s).  When the compressed buffers and / or the destination of an inflate
// call live in pinned / registered HOST memory, the call therefore gathers the inputs into device memory
// with wide coalesced reads, inflates there, and scatters the result back with wide coalesced writes
// (staged addresses keep the misalignment of the originals so that both sides move 16-byte vectors).
__global__ void __launch_bounds__(256) stage_copy_kernel(const bitar_chunk* __restrict__ from, const bitar_chunk* __restrict__ to,
                                                         const bitar_result* __restrict__ results, int out_direction) {
  const uint32_t i = blockIdx.x;
  const uint8_t* src;
  uint8_t* dst;
  uint32_t n;
  if (!out_direction) {   // gather: caller's compressed bytes -> stage
    src = static_cast<const uint8_t*>(from[i].src);
    dst = const_cast<uint8_t*>(static_cast<const uint8_t*>(to[i].src));
    n = from[i].src_len;
  } else {                // scatter: inflated bytes in the stage -> caller's destination
    src = static_cast<const uint8_t*>(from[i].dst);
    dst = static_cast<uint8_t*>(to[i].dst);
    n = min(results[i].produced, to[i].dst_cap);
  }
  if (src == dst) return;   // this op's buffer is device memory, used in place
  const uint32_t head = min(n, (16u - (uint32_t)(reinterpret_cast<uintptr_t>(src) & 15u)) & 15u);
  if (threadIdx.x < head) dst[threadIdx.x] = src[threadIdx.x];
  const uint32_t vecs = (n - head) >> 4;
  const uint4* s4 = reinterpret_cast<const uint4*>(src + head);
  uint4* d4 = reinterpret_cast<uint4*>(dst + head);
  for (uint32_t v = threadIdx.x; v < vecs; v += blockDim.x) d4[v] = s4[v];
  const uint32_t done = head + (vecs << 4);
  if (threadIdx.x < n - done) dst[done + threadIdx.x] = src[done + threadIdx.x];
}

// Resets a queue pair's work counters.  A kernel of our own rather than cudaMemsetAsync: the driver's memset kernel
// runs with the default shared-memory carve-out, and switching the carve-out back and forth drains the SMs.
__global__ void zero_counters_kernel(unsigned int* c) { c[threadIdx.x] = 0; }

// Descriptor upload / result download by the SMs (8-byte words between pinned host memory and device memory).  The
// copy engines serve their requests in order across ALL streams: a 400 KB descriptor upload issued after another
// queue pair's (or batch's) 128 MiB staging copy waits for it, and so does every kernel behind the upload
// (measured: a deflate call enqueued after a 1 GiB host-to-device copy on another stream started 19 ms late).
__global__ void __launch_bounds__(256) words_copy_kernel(uint2* __restrict__ dst, const uint2* __restrict__ src, uint32_t n) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = src[i];
}
static_assert(sizeof(bitar_chunk) % 8 == 0 && sizeof(bitar_result) % 8 == 0, "descriptors move as 8-byte words");
inline cudaError_t words_copy(void* dst, const void* src, size_t bytes, cudaStream_t st) {
  const uint32_t n = (uint32_t)(bytes / 8);
  if (n == 0) return cudaSuccess;
  uint32_t blocks = (n + 255u) / 256u;
  if (blocks > 64u) blocks = 64u;
  words_copy_kernel<<<blocks, 256, 0, st>>>(static_cast<uint2*>(dst), static_cast<const uint2*>(src), n);
  g_launches.fetch_add(1);
  return cudaGetLastError();
}

bool is_host_memory(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost;
}

// The same per op of a call, one driver query per 2 MiB of address space (host and device allocations never share
// such a region under unified addressing; a 1 GiB call costs ~500 queries instead of one per op).
struct HostProbe {
  uintptr_t region = ~(uintptr_t)0;
  bool host = false;
  bool operator()(const void* p) {
    const uintptr_t r = reinterpret_cast<uintptr_t>(p) >> 21;
    if (r != region) {
      region = r;
      host = is_host_memory(p);
    }
    return host;
  }
};

// Rewrites q->h_ops to staged device addresses when the call's buffers are host memory (caller's ops are
// kept in q->h_orig).  Called with the queue pair idle, before the descriptors are uploaded.
int inflate_prepare_staging(QueuePair* q, uint32_t n, bool allow_batches) {
  q->nb = 1;
  q->per = n;
  // Every op is classified: host-resident sources / destinations are staged, device-resident ones are used in place
  // (a call may mix them: the C++ facade stages pageable buffers itself and passes pool slots as they are).
  uint32_t host_src = 0, host_dst = 0, with_src = 0, with_dst = 0;
  {
    HostProbe ps, pd;
    for (uint32_t i = 0; i < n; ++i) {
      const bitar_chunk& c = q->h_ops[i];
      if (c.src != nullptr && c.src_len > 0) {
        ++with_src;
        host_src += ps(c.src) ? 1u : 0u;
      }
      if (c.dst != nullptr && c.dst_cap > 0) {
        ++with_dst;
        host_dst += pd(c.dst) ? 1u : 0u;
      }
    }
  }
  q->stage_src = host_src != 0;
  q->stage_dst = host_dst != 0;
  const bool all_src_host = host_src == with_src && with_src == n, all_dst_host = host_dst == with_dst && with_dst == n;
  if (!q->stage_src && !q->stage_dst) return BITAR_OK;
  if (q->orig_cap < n) {
    if (q->h_orig) cudaFreeHost(q->h_orig);
    if (q->d_orig) cudaFree(q->d_orig);
    q->h_orig = nullptr; q->d_orig = nullptr; q->orig_cap = 0;
    CU_TRY(cudaHostAlloc((void**)&q->h_orig, (size_t)q->cap * sizeof(bitar_chunk), cudaHostAllocPortable | cudaHostAllocMapped), BITAR_E_OUT_OF_MEMORY);
    CU_TRY(cudaMalloc((void**)&q->d_orig, (size_t)q->cap * sizeof(bitar_chunk)), BITAR_E_OUT_OF_MEMORY);
    q->orig_cap = q->cap;
  }
  memcpy(q->h_orig, q->h_ops, (size_t)n * sizeof(bitar_chunk));
  size_t need_in = 0, need_out = 0;
  uint32_t max_len = 0;
  for (uint32_t i = 0; i < n; ++i) {
    need_in += (((size_t)q->h_ops[i].src_len + 15u) & ~(size_t)15u) + 32u;
    need_out += (((size_t)q->h_ops[i].dst_cap + 15u) & ~(size_t)15u) + 32u;
    max_len = q->h_ops[i].src_len > max_len ? q->h_ops[i].src_len : max_len;
  }
  q->stage_out_bytes = need_out;
  // Staged calls with a host-resident destination run in batches (qp_submit): at most kMaxStageBatches, each of at
  // least kStageBatchBytes of output.
  if (q->stage_dst && all_dst_host && n > 1 && allow_batches) {
    static const size_t env_bytes = getenv("BITAR_STAGE_BATCH_MIB") ? (size_t)atoi(getenv("BITAR_STAGE_BATCH_MIB")) << 20 : kStageBatchBytes;
    const size_t tuned = g_stage_batch_bytes.load();
    const size_t batch_bytes = tuned ? tuned : env_bytes;
    const size_t want = need_out / (batch_bytes ? batch_bytes : 1);
    q->nb = (uint32_t)(want < 1 ? 1 : want > kMaxStageBatches ? kMaxStageBatches : want);
    if (q->nb > n) q->nb = n;
    q->per = (n + q->nb - 1) / q->nb;
    q->nb = (n + q->per - 1) / q->per;
  }
  // sources at a constant stride with room for the widest row?
  const size_t mis0 = reinterpret_cast<uintptr_t>(q->h_orig[0].src) & 15u;
  q->stage_src_strided = false;
  if (q->stage_src && all_src_host && n > 1 && g_stage_strided.load()) {
    const uint8_t* s0 = static_cast<const uint8_t*>(q->h_orig[0].src);
    const uint8_t* s1 = static_cast<const uint8_t*>(q->h_orig[1].src);
    const size_t stride = s1 > s0 ? (size_t)(s1 - s0) : 0;
    bool ok = stride > 0 && (stride & 15u) == 0 && ((mis0 + max_len + 15u) & ~(size_t)15u) <= stride;
    for (uint32_t i = 2; i < n && ok; ++i) ok = static_cast<const uint8_t*>(q->h_orig[i].src) == s0 + (size_t)i * stride;
    q->stage_src_strided = ok;
    q->src_stride = stride;
    if (ok) {   // a batch's rows share the pitch of its widest one
      need_in = 0;
      for (uint32_t b = 0, first = 0; first < n; ++b, first += q->per) {
        const uint32_t count = n - first < q->per ? n - first : q->per;
        uint32_t widest = 0;
        for (uint32_t i = first; i < first + count; ++i) widest = q->h_orig[i].src_len > widest ? q->h_orig[i].src_len : widest;
        q->batch_pitch[b] = (uint32_t)((mis0 + widest + 15u) & ~(size_t)15u);
        q->batch_in_off[b] = need_in;
        need_in += (size_t)count * q->batch_pitch[b] + 32u;
      }
    }
  }
  if (q->stage_src && q->stage_in_cap < need_in) {
    if (q->d_stage_in) cudaFree(q->d_stage_in);
    q->d_stage_in = nullptr; q->stage_in_cap = 0;
    CU_TRY(cudaMalloc((void**)&q->d_stage_in, need_in), BITAR_E_OUT_OF_MEMORY);
    q->stage_in_cap = need_in;
  }
  if (q->stage_dst && q->stage_out_cap < need_out) {
    if (q->d_stage_out) cudaFree(q->d_stage_out);
    q->d_stage_out = nullptr; q->stage_out_cap = 0;
    CU_TRY(cudaMalloc((void**)&q->d_stage_out, need_out), BITAR_E_OUT_OF_MEMORY);
    q->stage_out_cap = need_out;
  }
  if (q->stage_dst && !q->ev_fork) {
    for (uint32_t k = 0; k < kStageLanes; ++k) {
      // (highest priority: a batch's inflate kernel goes ahead of other queue pairs' compress launches, its copy-back
      // then runs beside them)
      int prio_lo = 0, prio_hi = 0;
      cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
      CU_TRY(cudaStreamCreateWithPriority(&q->lane[k], cudaStreamNonBlocking, prio_hi), BITAR_E_OUT_OF_MEMORY);
      CU_TRY(cudaEventCreateWithFlags(&q->ev_lane[k], cudaEventDisableTiming), BITAR_E_OUT_OF_MEMORY);
    }
    CU_TRY(cudaEventCreateWithFlags(&q->ev_fork, cudaEventDisableTiming), BITAR_E_OUT_OF_MEMORY);
  }
  // Decompress() lays segment i at out + i * S (src/memory.cc:482-493): then the staged output mirrors that
  // layout and a batch goes back with ONE copy-engine transfer (which, unlike a kernel, runs beside the inflate
  // kernels of the other batches and queue pairs); only the call's last segment, the one that may be short, is
  // copied by produced size.
  q->stage_dst_contig = q->stage_dst && all_dst_host && n > 1;
  for (uint32_t i = 1; i < n && q->stage_dst_contig; ++i)
    q->stage_dst_contig = q->h_orig[i].dst_cap == q->h_orig[0].dst_cap &&
                          static_cast<uint8_t*>(q->h_orig[i].dst) == static_cast<uint8_t*>(q->h_orig[0].dst) + (size_t)i * q->h_orig[0].dst_cap;
  size_t at_in = 0, at_out = 0;
  for (uint32_t i = 0; i < n; ++i) {
    bitar_chunk& c = q->h_ops[i];
    if (q->stage_src_strided) {
      const uint32_t b = i / q->per;
      c.src = q->d_stage_in + q->batch_in_off[b] + (size_t)(i - b * q->per) * q->batch_pitch[b] + mis0;
    } else if (q->stage_src && (all_src_host || (c.src != nullptr && c.src_len > 0 && is_host_memory(c.src)))) {
      const size_t mis = reinterpret_cast<uintptr_t>(c.src) & 15u;
      c.src = q->d_stage_in + at_in + mis;
      at_in += (((size_t)c.src_len + 15u) & ~(size_t)15u) + 32u;
    }
    if (q->stage_dst_contig) {
      c.dst = q->d_stage_out + (reinterpret_cast<uintptr_t>(q->h_orig[0].dst) & 15u) + (size_t)i * c.dst_cap;
    } else if (q->stage_dst && (all_dst_host || (c.dst != nullptr && c.dst_cap > 0 && is_host_memory(c.dst)))) {
      const size_t mis = reinterpret_cast<uintptr_t>(c.dst) & 15u;
      c.dst = q->d_stage_out + at_out + mis;
      at_out += (((size_t)c.dst_cap + 15u) & ~(size_t)15u) + 32u;
    }
  }
  return BITAR_OK;
}

// Compress() of a buffer in pinned / registered HOST memory: the kernel's own bulk loads over PCIe reach ~43 GB/s and
// slow to a crawl next to device-to-host traffic (read requests queue behind posted writes: measured 25 + 19 ms side by
// side -> 37 ms, tools/pcie_overlap_probe.py).  The segments of a Compress() call are consecutive pieces of one buffer
// (src/device.cc:168-170), so the call copies that range to a device stage with the copy engine, in pieces, and the
// kernel of a piece starts when the piece has arrived.  Rewrites q->h_ops to staged addresses.
constexpr size_t kDeflateStageMin = (size_t)4 << 20;
int deflate_prepare_staging(QueuePair* q, uint32_t n) {
  q->dstage = false;
  static const int off = getenv("BITAR_DEFLATE_STAGE") ? !atoi(getenv("BITAR_DEFLATE_STAGE")) : 0;
  if (off || n < 8 || q->h_ops[0].src == nullptr || !is_host_memory(q->h_ops[0].src)) return BITAR_OK;
  const uint8_t* base = static_cast<const uint8_t*>(q->h_ops[0].src);
  size_t total = 0;
  for (uint32_t i = 0; i < n; ++i) {
    if (static_cast<const uint8_t*>(q->h_ops[i].src) != base + total) return BITAR_OK;   // not one range: read in place
    total += q->h_ops[i].src_len;
  }
  if (total < kDeflateStageMin) return BITAR_OK;
  // few, long chunks (a CTA spends milliseconds on one): the pieces of a staged call run one after the other, each
  // with fewer chunks than the grid has CTAs -- reading in place, all chunks at once, is faster (1 MiB chunks: 8 -> 40 GB/s)
  if (total / n >= ((size_t)256 << 10) && n < 1184) return BITAR_OK;
  const size_t mis = reinterpret_cast<uintptr_t>(base) & 15u;
  if (q->stage_in_cap < total + 64) {
    if (q->d_stage_in) cudaFree(q->d_stage_in);
    q->d_stage_in = nullptr; q->stage_in_cap = 0;
    CU_TRY(cudaMalloc((void**)&q->d_stage_in, total + 64), BITAR_E_OUT_OF_MEMORY);
    q->stage_in_cap = total + 64;
  }
  if (!q->copy_stream) {
    CU_TRY(cudaStreamCreateWithFlags(&q->copy_stream, cudaStreamNonBlocking), BITAR_E_OUT_OF_MEMORY);
    for (uint32_t k = 0; k < kMaxStageBatches; ++k) {
      CU_TRY(cudaEventCreateWithFlags(&q->ev_copy[k], cudaEventDisableTiming), BITAR_E_OUT_OF_MEMORY);
    }
  }
  size_t at = 0;
  for (uint32_t i = 0; i < n; ++i) {
    q->h_ops[i].src = q->d_stage_in + mis + at;   // (same misalignment as the original: the bulk loads see the same layout)
    at += q->h_ops[i].src_len;
  }
  q->dstage = true;
  q->dstage_base = base;
  q->dstage_total = total;
  return BITAR_OK;
}

enum { kSubmitDeflate = 0, kSubmitInflate = 1, kSubmitInflateOneBatch = 2 };

// prepare(q, n): allocations for the whole call, before anything is enqueued;
// launch(q, first, count, counters, stream): the kernels of ops [first, first + count) on `stream`.
template <typename Prepare, typename Launch>
int qp_submit(bitar_dev* dev, uint16_t qp, const bitar_chunk* ops, uint32_t n, bitar_result* results, Prepare&& prepare,
              Launch&& launch, int mode) {
  const bool inflate = mode != kSubmitDeflate;
  if (!dev) return fail(BITAR_E_INVALID, "null device");
  if (qp >= dev->qps.size()) return fail(BITAR_E_INVALID, "queue_pair_id must be in the range of [0, %zu)", dev->qps.size());
  QueuePair* q = dev->qps[qp];
  if (q->busy.load(std::memory_order_acquire))
    return fail(BITAR_E_CANCELLED, "Queue pair %u of compress device %d is busy", (unsigned)qp, dev->id);
  if (n == 0) {
    q->last_status.store(BITAR_OK);
    return BITAR_OK;
  }
  if (!ops || !results) return fail(BITAR_E_INVALID, "null ops/results");
  if (!inflate)   // a chunk is at most one segment (src/device.cc:168-170 cuts the buffer); the kernels' block index is sized for it
    for (uint32_t i = 0; i < n; ++i)
      if (ops[i].src_len > BITAR_MAX_SEG_SIZE)
        return fail(BITAR_E_INVALID, "op %u: src_len %u above the largest segment (%u)", i, ops[i].src_len, BITAR_MAX_SEG_SIZE);
  CU_TRY(cudaSetDevice(dev->id), BITAR_E_INVALID);
  int rc = qp_reserve(dev, q, n);
  if (rc) return rc;
  memcpy(q->h_ops, ops, (size_t)n * sizeof(bitar_chunk));
  q->stage_src = q->stage_dst = q->stage_dst_contig = q->stage_src_strided = false;
  q->nb = 1;
  q->per = n;
  if (inflate) {
    rc = inflate_prepare_staging(q, n, mode == kSubmitInflate);
    if (rc) return rc;
  } else {
    rc = deflate_prepare_staging(q, n);
    if (rc) return rc;
  }
  q->user_out = results;
  q->pending_n = n;
  q->busy.store(1, std::memory_order_release);
  cudaError_t e = cudaEventRecord(q->ev_start, q->stream);
  if (e == cudaSuccess) e = words_copy(q->d_ops, q->h_ops, (size_t)n * sizeof(bitar_chunk), q->stream);
  if (e == cudaSuccess && (q->stage_src || q->stage_dst))
    e = words_copy(q->d_orig, q->h_orig, (size_t)n * sizeof(bitar_chunk), q->stream);
  // Staged inflate calls (host-resident buffers) run in batches: gather (PCIe host -> device), inflate and copy-back
  // (PCIe device -> host) of one batch in order on one of kStageLanes streams, the lanes side by side.  One batch
  // alone is latency-bound (a 2 KiB sub-range takes a lane ~0.8 ms whatever the batch size), hence several in flight.
  // Everything else is one batch on the queue pair's own stream.
  const uint32_t nb = q->nb;
  const bool lanes = nb > 1;
  if (e == cudaSuccess) e = prepare(q, n);     // buffers sized for the whole call: batches in flight share them
  if (e == cudaSuccess && lanes) e = cudaEventRecord(q->ev_fork, q->stream);
  if (e == cudaSuccess && lanes) e = cudaEventRecord(q->ev_k0, q->stream);
  const uint32_t per = q->per;
  for (uint32_t b = 0, first = 0; first < n && e == cudaSuccess; ++b, first += per) {
    const uint32_t count = n - first < per ? n - first : per;
    const bool last = first + count == n;
    cudaStream_t st = lanes ? q->lane[b % kStageLanes] : q->stream;
    if (lanes && b < kStageLanes) {
      e = cudaStreamWaitEvent(st, q->ev_fork, 0);
      if (e != cudaSuccess) continue;
    }
    if (q->stage_src_strided) {
      // one strided transfer for the batch's rows; the call's last op by its exact length (a row is read to the
      // batch's pitch, which stays inside the source range only while another row follows)
      const size_t mis0 = reinterpret_cast<uintptr_t>(q->h_orig[0].src) & 15u;
      const uint32_t rows = last ? count - 1 : count;
      uint8_t* const stage = q->d_stage_in + q->batch_in_off[b];
      if (rows)
        e = cudaMemcpy2DAsync(stage, q->batch_pitch[b], static_cast<const uint8_t*>(q->h_orig[first].src) - mis0, q->src_stride,
                              q->batch_pitch[b], rows, cudaMemcpyDefault, st);
      if (e == cudaSuccess && last)
        e = cudaMemcpyAsync(stage + (size_t)(count - 1) * q->batch_pitch[b] + mis0, q->h_orig[n - 1].src, q->h_orig[n - 1].src_len,
                            cudaMemcpyDefault, st);
    } else if (q->stage_src) {
      stage_copy_kernel<<<count, 256, 0, st>>>(q->d_orig + first, q->d_ops + first, nullptr, 0);
      e = cudaGetLastError();
      g_launches.fetch_add(1);
    }
    if (e == cudaSuccess && !lanes && b == 0) e = cudaEventRecord(q->ev_k0, st);
    unsigned int* counters = q->d_counter + kCountersPerBatch * b;
    if (e == cudaSuccess) {
      zero_counters_kernel<<<1, kCountersPerBatch, 0, st>>>(counters);
      e = cudaGetLastError();
    }
    // the contiguous copy-back below takes whole segments whatever their ops produced: a failed or short op must not hand
    // the caller bytes that an earlier call left in the stage
    if (e == cudaSuccess && q->stage_dst_contig)
      e = cudaMemsetAsync(q->h_ops[first].dst, 0, (size_t)count * q->h_orig[0].dst_cap, st);
    if (e == cudaSuccess) e = launch(q, first, count, counters, st);
    g_launches.fetch_add(2);   // the counter reset + the batch's last codec kernel (the others are counted where they launch)
    if (e == cudaSuccess && !lanes) e = cudaEventRecord(q->ev_k1, st);
    if (e != cudaSuccess || !q->stage_dst) continue;
    if (q->stage_dst_contig) {
      // one copy-engine transfer for the batch's full segments; the call's last segment, the one that may be
      // short, goes by its produced size
      const uint32_t full = last ? count - 1 : count;
      if (full)
        e = cudaMemcpyAsync(q->h_orig[first].dst, q->h_ops[first].dst, (size_t)full * q->h_orig[0].dst_cap, cudaMemcpyDefault, st);
      if (e == cudaSuccess && last) {
        stage_copy_kernel<<<1, 256, 0, st>>>(q->d_ops + (n - 1), q->d_orig + (n - 1), q->d_res + (n - 1), 1);
        e = cudaGetLastError();
        g_launches.fetch_add(1);
      }
    } else {
      stage_copy_kernel<<<count, 256, 0, st>>>(q->d_ops + first, q->d_orig + first, q->d_res + first, 1);
      e = cudaGetLastError();
      g_launches.fetch_add(1);
    }
  }
  if (lanes) {   // join
    for (uint32_t k = 0; k < kStageLanes && k < nb && e == cudaSuccess; ++k) {
      e = cudaEventRecord(q->ev_lane[k], q->lane[k]);
      if (e == cudaSuccess) e = cudaStreamWaitEvent(q->stream, q->ev_lane[k], 0);
    }
    if (e == cudaSuccess) e = cudaEventRecord(q->ev_k1, q->stream);
  }
  if (e == cudaSuccess) e = words_copy(q->h_res, q->d_res, (size_t)n * sizeof(bitar_result), q->stream);
  if (e == cudaSuccess) e = cudaEventRecord(q->ev_stop, q->stream);
  if (e == cudaSuccess) e = cudaLaunchHostFunc(q->stream, qp_finish, q);
  if (e != cudaSuccess) {
    q->busy.store(0);
    return fail(BITAR_E_IO_ERROR, "enqueue on queue pair %u of device %d failed: %s", (unsigned)qp, dev->id, cudaGetErrorString(e));
  }
  q->timed = true;
  return BITAR_OK;
}

bool spec_enabled() {   // BITAR_SPEC=0: streams without an index go straight to the whole-stream kernel (A/B runs)
  static const int on = getenv("BITAR_SPEC") ? atoi(getenv("BITAR_SPEC")) : 1;
  return on != 0;
}
std::atomic<int> g_spec_target{0};
uint32_t spec_target() {   // output bytes per speculative range (sp::range_bits)
  static const int env = getenv("BITAR_SPEC_TARGET") ? atoi(getenv("BITAR_SPEC_TARGET")) : 1024;
  const int o = g_spec_target.load(), t = o > 0 ? o : env;
  return (uint32_t)(t < 64 ? 64 : t > 1536 ? 1536 : t);
}
int inflate_variant() {
  int v = g_inflate_variant.load();
  if (v < 0) {
    const char* s = getenv("BITAR_INFLATE_VARIANT");
    v = s ? atoi(s) : 0;
    g_inflate_variant.store(v);
  }
  return v;
}

// (re)allocate a device buffer of the queue pair that only ever grows
template <typename T>
cudaError_t grow(T** p, size_t* cap, size_t need) {
  if (*cap >= need) return cudaSuccess;
  if (*p) cudaFree(*p);
  *p = nullptr;
  *cap = 0;
  cudaError_t e = cudaMalloc((void**)p, need * sizeof(T));
  if (e == cudaSuccess) *cap = need;
  return e;
}

}  // namespace

extern "C" {

int bitar_cuda_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    cudaGetLastError();
    fail(BITAR_E_INVALID, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    return 0;
  }
  return n;
}

int bitar_cuda_device_info(int device_id, bitar_dev_info* info) {
  if (!info) return fail(BITAR_E_INVALID, "null info");
  cudaDeviceProp p;
  CU_TRY(cudaGetDeviceProperties(&p, device_id), BITAR_E_INVALID);
  memset(info, 0, sizeof *info);
  info->device_id = device_id;
  info->cc_major = p.major;
  info->cc_minor = p.minor;
  info->sm_count = p.multiProcessorCount;
  info->total_mem = p.totalGlobalMem;
  info->max_queue_pairs = 64;
  info->window_min = 8;
  info->window_max = 15;
  info->supports_fixed = 1;
  info->supports_dynamic = 1;
  info->supports_crc32 = 1;
  info->supports_adler32 = 1;
  info->supports_sgl = 1;
  snprintf(info->name, sizeof info->name, "%s", p.name);
  return BITAR_OK;
}

uint32_t bitar_reference_compressed_seg_size(uint32_t seg) { return ref_compressed_seg_size(seg); }

uint32_t bitar_compressed_seg_size(uint32_t seg) {
  uint32_t ref = ref_compressed_seg_size(seg);
  uint32_t widened = (uint32_t)((double)seg * 1.1);
  uint32_t v = ref ? ref : widened;
  uint32_t need = stored_bound(seg);
  return v < need ? need : v;
}

int bitar_dev_open(int device_id, uint16_t n_qps, const bitar_cfg* cfg_in, bitar_dev** out) {
  if (!out) return fail(BITAR_E_INVALID, "null out");
  *out = nullptr;
  int count = bitar_cuda_device_count();
  if (device_id < 0 || device_id >= count) return fail(BITAR_E_INVALID, "device id %d not in [0, %d)", device_id, count);
  if (n_qps == 0) return fail(BITAR_E_INVALID, "a device needs at least one queue pair");
  bitar_dev_info info;
  int rc = bitar_cuda_device_info(device_id, &info);
  if (rc) return rc;
  if (info.cc_major < 10)
    return fail(BITAR_E_NOT_IMPLEMENTED, "Unsupported compress device %d (%s, sm_%d%d): kernels are built for sm_100a only",
                device_id, info.name, info.cc_major, info.cc_minor);
  if (n_qps > info.max_queue_pairs)
    return fail(BITAR_E_INVALID, "The requested number of queue pairs (%u) exceeds the maximum (%u) allowed for device %d",
                (unsigned)n_qps, info.max_queue_pairs, device_id);
  bitar_cfg cfg{};
  if (cfg_in) cfg = *cfg_in;
  // ValidateConfiguration, src/device.cc:352-415 (+ BlueField specifics 558-577)
  if (cfg.burst_size == 0) cfg.burst_size = 32;
  if (cfg.max_sgl_segs < 1) cfg.max_sgl_segs = 1;
  if (cfg.max_sgl_segs > BITAR_MAX_SGL_SEGS)
    return fail(BITAR_E_INVALID, "max_sgl_segs (%u) is not in the range of [1, %u]", (unsigned)cfg.max_sgl_segs, BITAR_MAX_SGL_SEGS);
  if (cfg.decompressed_seg_size == 0) cfg.decompressed_seg_size = 2048;
  if (cfg.decompressed_seg_size < BITAR_MIN_SEG_SIZE || cfg.decompressed_seg_size > BITAR_MAX_SEG_SIZE)
    return fail(BITAR_E_INVALID, "decompressed_seg_size is not in the range of [%u, %u]", BITAR_MIN_SEG_SIZE, BITAR_MAX_SEG_SIZE);
  if ((uint64_t)cfg.max_sgl_segs * cfg.decompressed_seg_size > BITAR_MAX_SEG_SIZE)
    return fail(BITAR_E_INVALID, "max_sgl_segs * decompressed_seg_size (%u * %u) is above the largest stream (%u bytes)",
                (unsigned)cfg.max_sgl_segs, cfg.decompressed_seg_size, BITAR_MAX_SEG_SIZE);
  if (cfg.window_size == 0) cfg.window_size = info.window_max;
  if (cfg.window_size < info.window_min || cfg.window_size > info.window_max)
    return fail(BITAR_E_INVALID, "window_size is not in the range of [%u, %u]", info.window_min, info.window_max);
  if (cfg.huffman_enc == BITAR_HUFFMAN_DEFAULT) cfg.huffman_enc = BITAR_HUFFMAN_DYNAMIC;
  if (cfg.huffman_enc != BITAR_HUFFMAN_FIXED && cfg.huffman_enc != BITAR_HUFFMAN_DYNAMIC)
    return fail(BITAR_E_INVALID, "unknown huffman_enc %u", cfg.huffman_enc);
  if (cfg.checksum_type > BITAR_CHECKSUM_CRC32_ADLER32) return fail(BITAR_E_INVALID, "unknown checksum_type %u", cfg.checksum_type);
  if (cfg.slot_mem_kind > BITAR_MEM_PINNED) return fail(BITAR_E_INVALID, "unknown slot_mem_kind %u", cfg.slot_mem_kind);
  if (cfg.max_preallocate_slots == 0) cfg.max_preallocate_slots = kRteMaxMemzone;
  if (cfg.max_preallocate_slots < BITAR_MIN_PREALLOCATE_SLOTS)
    return fail(BITAR_E_INVALID, "max_preallocate_memzones (%u) is not in the range of [%u, ...]", cfg.max_preallocate_slots,
                BITAR_MIN_PREALLOCATE_SLOTS);
  uint32_t min_slot = stored_bound(cfg.decompressed_seg_size);
  if (cfg.compressed_seg_size == 0) cfg.compressed_seg_size = bitar_compressed_seg_size(cfg.decompressed_seg_size);
  if (cfg.compressed_seg_size < min_slot) cfg.compressed_seg_size = min_slot;

  CU_TRY(cudaSetDevice(device_id), BITAR_E_INVALID);
  bitar_dev* dev = new (std::nothrow) bitar_dev();
  if (!dev) return fail(BITAR_E_OUT_OF_MEMORY, "out of host memory");
  dev->id = device_id;
  dev->sm_count = info.sm_count;
  dev->cfg = cfg;
  // chained segments (max_sgl_segs > 1): the slots of a stream are one contiguous range, so they lie back to back
  dev->slot_stride = cfg.max_sgl_segs > 1 ? cfg.compressed_seg_size : (cfg.compressed_seg_size + 255u) & ~255u;
  dev->group = cfg.max_sgl_segs > 1 ? cfg.max_sgl_segs : 1;
  {
    // keep freed device memory cached in the pool (allocation is off the timed path, as in
    // apps/demo_app.cc:517-522,587-590, but re-use must stay cheap)
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device_id) == cudaSuccess) {
      uint64_t thr = UINT64_MAX;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
  }
  // same carve-out for the small helper kernels as for the codec kernels (see inflate_kernel.cuh)
  cudaFuncSetAttribute(stage_copy_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  cudaFuncSetAttribute(zero_counters_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  cudaFuncSetAttribute(words_copy_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  cudaFuncSetAttribute(bitar::xk::inflate_plan_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  cudaFuncSetAttribute(bitar::xk::inflate_checksum_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  cudaError_t e = bitar::dk::deflate_grid(device_id, dev->sm_count, &dev->deflate_grid);
  if (e == cudaSuccess) e = bitar::dks::deflate_grid(device_id, dev->sm_count, &dev->deflate_grid_small);
  if (const char* g = getenv("BITAR_DEBUG_DEFLATE_GRID")) {   // tuning experiments only
    if (atoi(g) > 0) dev->deflate_grid_override = atoi(g);
  }
  if (e != cudaSuccess) {
    delete dev;
    return fail(BITAR_E_INVALID, "deflate kernel cannot be configured on device %d: %s", device_id, cudaGetErrorString(e));
  }
  for (uint16_t i = 0; i < n_qps; ++i) {
    QueuePair* q = new (std::nothrow) QueuePair();
    if (!q) {
      bitar_dev_close(dev);
      return fail(BITAR_E_OUT_OF_MEMORY, "out of host memory");
    }
    dev->qps.push_back(q);
    cudaError_t e2 = cudaStreamCreateWithFlags(&q->stream, cudaStreamNonBlocking);
    if (e2 == cudaSuccess) e2 = cudaEventCreate(&q->ev_start);
    if (e2 == cudaSuccess) e2 = cudaEventCreate(&q->ev_k0);
    if (e2 == cudaSuccess) e2 = cudaEventCreate(&q->ev_k1);
    if (e2 == cudaSuccess) e2 = cudaEventCreate(&q->ev_stop);
    if (e2 == cudaSuccess) e2 = cudaMalloc((void**)&q->d_counter, kCountersPerBatch * kMaxStageBatches * sizeof(unsigned int));
    if (e2 != cudaSuccess) {
      bitar_dev_close(dev);
      return fail(BITAR_E_INVALID, "Failed to setup queue pair %u for device %d: %s", (unsigned)i, device_id, cudaGetErrorString(e2));
    }
  }
  {
    std::lock_guard<std::mutex> lock(dev->mu);
    rc = pool_grow(dev, cfg.max_preallocate_slots);
  }
  if (rc) {
    bitar_dev_close(dev);
    return rc;
  }
  *out = dev;
  return BITAR_OK;
}

int bitar_dev_close(bitar_dev* dev) {
  if (!dev) return BITAR_OK;
  cudaSetDevice(dev->id);
  for (QueuePair* q : dev->qps) {
    if (q->stream) cudaStreamSynchronize(q->stream);
    if (q->h_ops) cudaFreeHost(q->h_ops);
    if (q->h_res) cudaFreeHost(q->h_res);
    if (q->d_ops) cudaFree(q->d_ops);
    if (q->d_res) cudaFree(q->d_res);
    if (q->d_counter) cudaFree(q->d_counter);
    if (q->d_far) cudaFree(q->d_far);   // (far candidates and tokens: one allocation)
    if (q->d_tasks) cudaFree(q->d_tasks);
    if (q->d_units) cudaFree(q->d_units);
    if (q->d_generic) cudaFree(q->d_generic);
    if (q->d_declined) cudaFree(q->d_declined);
    if (q->d_indexed) cudaFree(q->d_indexed);
    if (q->h_orig) cudaFreeHost(q->h_orig);
    if (q->d_orig) cudaFree(q->d_orig);
    if (q->d_stage_in) cudaFree(q->d_stage_in);
    if (q->d_stage_out) cudaFree(q->d_stage_out);
    if (q->ev_start) cudaEventDestroy(q->ev_start);
    if (q->ev_k0) cudaEventDestroy(q->ev_k0);
    if (q->ev_k1) cudaEventDestroy(q->ev_k1);
    if (q->ev_stop) cudaEventDestroy(q->ev_stop);
    for (uint32_t k = 0; k < kStageLanes; ++k) {
      if (q->ev_lane[k]) cudaEventDestroy(q->ev_lane[k]);
      if (q->lane[k]) cudaStreamDestroy(q->lane[k]);
    }
    if (q->ev_fork) cudaEventDestroy(q->ev_fork);
    for (uint32_t k = 0; k < kMaxStageBatches; ++k)
      if (q->ev_copy[k]) cudaEventDestroy(q->ev_copy[k]);
    if (q->copy_stream) cudaStreamDestroy(q->copy_stream);

    if (q->stream) cudaStreamDestroy(q->stream);
    delete q;
  }
  for (void* slab : dev->slabs) free_kind(dev->cfg.slot_mem_kind, dev->id, slab);
  delete dev;
  return BITAR_OK;
}

int bitar_dev_config(const bitar_dev* dev, bitar_cfg* cfg_out) {
  if (!dev || !cfg_out) return fail(BITAR_E_INVALID, "null argument");
  *cfg_out = dev->cfg;
  return BITAR_OK;
}

uint16_t bitar_dev_num_qps(const bitar_dev* dev) { return dev ? (uint16_t)dev->qps.size() : 0; }

int bitar_qp_deflate(bitar_dev* dev, uint16_t qp, const bitar_chunk* ops, uint32_t n, bitar_result* results) {
  return qp_submit(
      dev, qp, ops, n, results,
      [&](QueuePair* q, uint32_t) -> cudaError_t {
        if (q->d_tokens) return cudaSuccess;
        const size_t a = bitar::dk::deflate_scratch_bytes(dev->deflate_grid), b = bitar::dks::deflate_scratch_bytes(dev->deflate_grid_small);
        const size_t fa = bitar::dk::deflate_far_bytes(dev->deflate_grid), fb = bitar::dks::deflate_far_bytes(dev->deflate_grid_small);
        // one allocation: [far candidates | tokens].  Both are written and read once per chunk by the CTA that owns their
        // part and rewritten for its next chunk ~0.3 ms later.  BITAR_L2_PERSIST=1 marks them PERSISTING in L2 on the queue
        // pair's stream, so that the input streaming through does not push them out to HBM in between: measured, the
        // deflate kernel's DRAM traffic falls from 3.1 to 2.2 bytes per input byte at unchanged speed (it is not DRAM-bound),
        // but the set-aside is the DEVICE's and shrinks the L2 that the inflate kernel's token maps live in: inflate
        // 205 -> 154 GB/s.  Hence off by default.
        const size_t far_bytes = ((fa > fb ? fa : fb) + 255u) & ~(size_t)255u, tok_bytes = a > b ? a : b;
        cudaError_t e = cudaMalloc((void**)&q->d_far, far_bytes + tok_bytes);
        if (e != cudaSuccess) return e;
        q->d_tokens = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(q->d_far) + far_bytes);
        static const int persist = getenv("BITAR_L2_PERSIST") ? atoi(getenv("BITAR_L2_PERSIST")) : 0;   // off by default: see below
        int max_persist = 0, max_window = 0;
        cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev->id);
        cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev->id);
        if (persist && max_persist > 0 && max_window > 0) {
          size_t want = far_bytes + tok_bytes;
          if (want > (size_t)max_window) want = (size_t)max_window;
          const size_t carve = want < (size_t)max_persist ? want : (size_t)max_persist;
          if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve) == cudaSuccess) {
            cudaStreamAttrValue v{};
            v.accessPolicyWindow.base_ptr = q->d_far;
            v.accessPolicyWindow.num_bytes = want;
            v.accessPolicyWindow.hitRatio = (float)((double)carve / (double)want);
            v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            cudaStreamSetAttribute(q->stream, cudaStreamAttributeAccessPolicyWindow, &v);
          }
          cudaGetLastError();   // (the hint is optional: a refusal is not an error of the call)
        }
        return cudaSuccess;
      },
      [&](QueuePair* q, uint32_t first, uint32_t count, unsigned int* counters, cudaStream_t st) -> cudaError_t {
        // calls whose chunks all fit 16 KiB (8 sub-ranges: half of a 16-warp CTA would idle) go to the instance with
        // 4 warps per CTA and three to four times the CTAs per SM; same output
        uint32_t max_len = 0;
        for (uint32_t i = 0; i < count; ++i) max_len = q->h_ops[first + i].src_len > max_len ? q->h_ops[first + i].src_len : max_len;
        const int max_dist = 1 << dev->cfg.window_size, emit_index = dev->cfg.no_index ? 0 : 1;
        // host-resident input (read over PCIe by the kernel's bulk loads): the call yields the SMs a few times so that
        // other queue pairs' inflate batches -- whose copy-back uses the other PCIe direction -- run beside it
        static const int env_splits = getenv("BITAR_DEFLATE_SPLITS") ? atoi(getenv("BITAR_DEFLATE_SPLITS")) : 0;
        const uint32_t splits = env_splits > 0 ? (uint32_t)env_splits : 1u;
        static const int small_mode = getenv("BITAR_DEFLATE_SMALL") ? atoi(getenv("BITAR_DEFLATE_SMALL")) : 1;   // 0: off (A/B runs)
        const bool small = small_mode && max_len <= (uint32_t)bitar::dks::kBlockMax;
        auto launch_part = [&](uint32_t a, uint32_t cnt, unsigned int* ctr, uint32_t parts) -> cudaError_t {
          // calls whose chunks all fit 16 KiB (8 sub-ranges: half of a 16-warp CTA would idle) go to the instance with
          // 4 warps per CTA and three to four times the CTAs per SM; same output
          if (small)
            return bitar::dks::deflate_launch(q->d_ops + a, cnt, q->d_res + a, ctr, q->d_tokens, q->d_far, dev->id, dev->sm_count, 0, max_len,
                                              dev->cfg.huffman_enc, dev->cfg.checksum_type, max_dist, emit_index, g_deflate_prof.load(), st, parts);
          return bitar::dk::deflate_launch(q->d_ops + a, cnt, q->d_res + a, ctr, q->d_tokens, q->d_far, dev->id, dev->sm_count,
                                           dev->deflate_grid_override, (uint32_t)bitar::dk::kBlockMax, dev->cfg.huffman_enc,
                                           dev->cfg.checksum_type, max_dist, emit_index, g_deflate_prof.load(), st, parts);
        };
        if (!q->dstage) {
          g_launches.fetch_add(bitar::dk::deflate_splits(count, splits) - 1u);
          return launch_part(first, count, counters, splits);
        }
        // host-resident input: pieces of at least 8 MiB, at most 8; copy b runs beside kernel b - 1
        uint32_t pieces = (uint32_t)(q->dstage_total / ((size_t)8 << 20));
        pieces = pieces < 1 ? 1 : pieces > kMaxStageBatches ? kMaxStageBatches : pieces;
        const uint32_t per = (count + pieces - 1) / pieces;
        const size_t mis = reinterpret_cast<uintptr_t>(q->dstage_base) & 15u;
        size_t off = 0;
        uint32_t b = 0;
        for (uint32_t a = 0; a < count; a += per, ++b) {
          const uint32_t cnt = count - a < per ? count - a : per;
          size_t bytes = 0;
          for (uint32_t i = 0; i < cnt; ++i) bytes += q->h_ops[first + a + i].src_len;
          // (throttling the copies to two pieces ahead of the kernels measured slower: 43 against 41 ms per GiB end to end)
          cudaError_t e = cudaMemcpyAsync(q->d_stage_in + mis + off, q->dstage_base + off, bytes, cudaMemcpyDefault, q->copy_stream);
          if (e == cudaSuccess) e = cudaEventRecord(q->ev_copy[b], q->copy_stream);
          if (e == cudaSuccess) e = cudaStreamWaitEvent(st, q->ev_copy[b], 0);
          if (e == cudaSuccess) e = launch_part(first + a, cnt, counters + b, 1);
          if (e != cudaSuccess) return e;
          off += bytes;
        }
        g_launches.fetch_add(b - 1u);
        return cudaSuccess;
      },
      kSubmitDeflate);
}

namespace {
inline uint32_t op_blocks(const bitar_chunk& c) {   // 64 KiB blocks an op's output can span
  const uint32_t cap = c.dst_cap < BITAR_MAX_SEG_SIZE ? c.dst_cap : BITAR_MAX_SEG_SIZE;
  return (cap + 65535u) >> 16;
}
}  // namespace

using TokWide = bitar::xk::TokConfig<9, 864, 7, 256, 16, 32, 2>;    // a warp per 64 KiB block
using SpecWide = bitar::sk::SpecConfig<9, 864, 7, 256, 14, 2, 2048, 1024>;   // a warp per stream without an index: speculative lane-parallel decode (14 warps, 72 registers, 2 KiB rings)
using TokSmall = bitar::xk::TokConfig<9, 864, 7, 256, 2, 8, 7>;     // four blocks of at most 8 sub-ranges per warp; small CTAs: shared memory (3.9 KB per block) decides how many warps an SM holds (14)

int bitar_qp_inflate(bitar_dev* dev, uint16_t qp, const bitar_chunk* ops, uint32_t n, bitar_result* results) {
  const int variant = inflate_variant();
  const int ck = dev ? dev->cfg.checksum_type : 0, id = dev ? dev->id : 0, sms = dev ? dev->sm_count : 0;
  size_t total_blocks = 0;
  uint32_t subs = 32;          // unit slots per task: 32 sub-ranges of a 64 KiB block, fewer when every segment is small
  bool small_mode = false, spec_mode = false;
  return qp_submit(
      dev, qp, ops, n, results,
      [&](QueuePair* q, uint32_t n_all) -> cudaError_t {
        // buffers of the two-phase path, sized for the whole call (its batches run side by side, each on its own slice)
        using namespace bitar::xk;
        uint32_t max_cap = 0;
        for (uint32_t i = 0; i < n_all; ++i) {
          total_blocks += op_blocks(q->h_ops[i]);
          max_cap = q->h_ops[i].dst_cap > max_cap ? q->h_ops[i].dst_cap : max_cap;
        }
        if (variant == 5) return cudaSuccess;
        // segments of at most 8 sub-ranges (16 KiB) are decoded four to a warp, with as many unit slots as they need
        small_mode = max_cap <= kSmallSubs * bitar::dfl::kSub;
        if (small_mode) subs = max_cap ? (max_cap + bitar::dfl::kSub - 1u) >> bitar::dfl::kSubLog2 : 1u;
        // (staged calls run up to kStageLanes batches side by side: one scratch each)
        size_t one = small_mode ? TokSmall::scratch_bytes(id, sms, subs) : TokWide::scratch_bytes(id, sms, subs);
        if (one == 0) return cudaErrorLaunchOutOfResources;
        // streams without an index are decoded speculatively (inflate_spec_kernel.cuh) in the same scratch, after the
        // indexed ones; only calls that can hold such a stream pay for its size (a stream of < 64 bytes never is one)
        spec_mode = variant != 6 && spec_enabled();
        if (spec_mode) {
          bool any = false;
          for (uint32_t i = 0; i < n_all && !any; ++i) any = q->h_ops[i].src_len >= 64u;
          spec_mode = any;
        }
        if (spec_mode) {
          const size_t sp_raw = SpecWide::scratch_bytes(id, sms);
          const size_t sp_one = (sp_raw + 15u) / 16u * 16u;
          if (sp_one == 0) return cudaErrorLaunchOutOfResources;
          one = sp_one > one ? sp_one : one;
        }
        cudaError_t e = grow(&q->d_tasks, &q->tasks_cap, total_blocks);
        if (e == cudaSuccess) e = grow(&q->d_units, &q->units_cap, one * (q->nb > 1 ? kStageLanes : 1u));
        size_t cap2 = q->generic_cap, cap3 = q->generic_cap;
        if (e == cudaSuccess) e = grow(&q->d_declined, &cap3, (size_t)n_all);
        if (e == cudaSuccess) e = grow(&q->d_generic, &q->generic_cap, (size_t)n_all);
        if (e == cudaSuccess) e = grow(&q->d_indexed, &cap2, (size_t)n_all);
        return e;
      },
      [&](QueuePair* q, uint32_t first, uint32_t n, unsigned int* counters, cudaStream_t st) -> cudaError_t {
        using namespace bitar::ik;
        bitar_chunk* const d_ops = q->d_ops + first;       // this batch of the call (one batch unless staged)
        bitar_result* const d_res = q->d_res + first;
        const bitar_chunk* const h_ops = q->h_ops + first;
        if (variant == 5)   // tests: everything through the whole-stream kernel
          return InflateConfig<32, 10, 8, 1024, 4>::launch(d_ops, n, d_res, counters, ck, id, sms, st);
        // default: chunks carrying the parallel-inflate index take the two-phase path (a warp Huffman-decodes the 32
        // sub-ranges of a 64 KiB block into token units, a group of 8 lanes resolves the block's copies in order);
        // everything else (zlib streams, stored chunks) goes to the whole-stream kernel.
        using namespace bitar::xk;
        size_t before = 0, blocks = 0;
        for (uint32_t i = 0; i < first; ++i) before += op_blocks(q->h_ops[i]);
        for (uint32_t i = 0; i < n; ++i) blocks += op_blocks(h_ops[i]);
        Task* const tasks = q->d_tasks + before;
        // the batch's stream decides the scratch: batches on one stream run one after the other
        uint32_t lane_idx = 0;
        for (uint32_t k = 0; k < kStageLanes; ++k)
          if (st == q->lane[k]) lane_idx = k;
        uint8_t* const units = q->d_units + (q->nb > 1 ? lane_idx : 0u) * (q->units_cap / (q->nb > 1 ? kStageLanes : 1u) / 16u * 16u);
        uint32_t* const generic = q->d_generic + first;
        uint32_t* const indexed = q->d_indexed + first;
        Counters* pc = reinterpret_cast<Counters*>(counters);
        inflate_plan_kernel<<<(n + 127) / 128, 128, 0, st>>>(d_ops, n, d_res, tasks, indexed, generic, pc, 1);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        if (small_mode) e = TokSmall::launch(d_ops, d_res, tasks, pc, units, subs, (uint32_t)blocks, id, sms, st);
        else e = TokWide::launch(d_ops, d_res, tasks, pc, units, subs, (uint32_t)blocks, id, sms, st);
        if (e != cudaSuccess) return e;
        g_launches.fetch_add(2);
        if (ck != BITAR_CHECKSUM_NONE) {
          inflate_checksum_kernel<<<(n + 3) / 4 < (uint32_t)(8 * sms) ? (n + 3) / 4 : (uint32_t)(8 * sms), 128, 0, st>>>(d_ops, d_res, indexed, pc, ck);
          e = cudaGetLastError();
          if (e != cudaSuccess) return e;
          g_launches.fetch_add(1);
        }
        if (spec_mode) {
          // what the speculative kernel declines (stored blocks, damaged streams, ...) is left to the whole-stream kernel
          uint32_t* const declined = q->d_declined + first;
          e = SpecWide::launch(d_ops, d_res, generic, pc, declined, units, ck, spec_target(), n, id, sms, st);
          if (e != cudaSuccess) return e;
          g_launches.fetch_add(1);
          return InflateConfig<32, 10, 8, 1024, 4>::launch(d_ops, n, d_res, &pc->generic_next, ck, id, sms, st, declined, &pc->n_declined);
        }
        return InflateConfig<32, 10, 8, 1024, 4>::launch(d_ops, n, d_res, &pc->generic_next, ck, id, sms, st, generic, &pc->n_generic);
      },
      kSubmitInflate);
}

int bitar_qp_wait(bitar_dev* dev, uint16_t qp) {
  if (!dev || qp >= dev->qps.size()) return fail(BITAR_E_INVALID, "bad device/queue pair");
  QueuePair* q = dev->qps[qp];
  CU_TRY(cudaSetDevice(dev->id), BITAR_E_INVALID);
  cudaError_t e = cudaStreamSynchronize(q->stream);
  if (e != cudaSuccess) {
    q->busy.store(0);
    return fail(BITAR_E_IO_ERROR, "queue pair %u of device %d failed: %s", (unsigned)qp, dev->id, cudaGetErrorString(e));
  }
  int st = q->last_status.load(std::memory_order_acquire);
  if (st) return fail(st, "at least one operation on queue pair %u of device %d did not succeed", (unsigned)qp, dev->id);
  return BITAR_OK;
}

int bitar_qp_result(bitar_dev* dev, uint16_t qp) {
  if (!dev || qp >= dev->qps.size()) return fail(BITAR_E_INVALID, "bad device/queue pair");
  QueuePair* q = dev->qps[qp];
  if (q->busy.load(std::memory_order_acquire)) return fail(BITAR_E_CANCELLED, "queue pair %u of device %d is busy", (unsigned)qp, dev->id);
  int st = q->last_status.load(std::memory_order_acquire);
  if (st) return fail(st, "at least one operation on queue pair %u of device %d did not succeed", (unsigned)qp, dev->id);
  return BITAR_OK;
}

int bitar_qp_busy(bitar_dev* dev, uint16_t qp) {
  if (!dev || qp >= dev->qps.size()) return 0;
  return dev->qps[qp]->busy.load(std::memory_order_acquire);
}

int bitar_qp_on_complete(bitar_dev* dev, uint16_t qp, void (*fn)(void*), void* arg) {
  if (!dev || qp >= dev->qps.size() || !fn) return fail(BITAR_E_INVALID, "bad argument");
  CU_TRY(cudaSetDevice(dev->id), BITAR_E_INVALID);
  CU_TRY(cudaLaunchHostFunc(dev->qps[qp]->stream, (cudaHostFn_t)fn, arg), BITAR_E_IO_ERROR);
  return BITAR_OK;
}

int bitar_qp_last_ms(bitar_dev* dev, uint16_t qp, float* kernel_ms, float* total_ms) {
  if (!dev || qp >= dev->qps.size()) return fail(BITAR_E_INVALID, "bad device/queue pair");
  QueuePair* q = dev->qps[qp];
  if (!q->timed) return fail(BITAR_E_INVALID, "no completed call on this queue pair");
  CU_TRY(cudaSetDevice(dev->id), BITAR_E_INVALID);
  if (kernel_ms) CU_TRY(cudaEventElapsedTime(kernel_ms, q->ev_k0, q->ev_k1), BITAR_E_INVALID);
  if (total_ms) CU_TRY(cudaEventElapsedTime(total_ms, q->ev_start, q->ev_stop), BITAR_E_INVALID);
  return BITAR_OK;
}

void* bitar_qp_stream(bitar_dev* dev, uint16_t qp) {
  if (!dev || qp >= dev->qps.size()) return nullptr;
  return dev->qps[qp]->stream;
}

uint64_t bitar_kernel_launches(void) { return g_launches.load(); }

void* bitar_slot_take(bitar_dev* dev) {
  if (!dev) return nullptr;
  std::lock_guard<std::mutex> lock(dev->mu);
  if (dev->group > 1) {
    void* one = nullptr;
    return pool_take_groups(dev, 1, &one) == BITAR_OK ? one : nullptr;
  }
  if (dev->free_slots.empty()) {
    // growing on the critical path, as DeviceMemory::Take does with a warning (src/memory.cc:167-175)
    if (dev->grow_warned++ % 32 == 0)
      fprintf(stderr, "[bitar] WARNING: allocating output slots in the critical path (performance will be impacted)\n");
    uint32_t add = dev->cfg.burst_size ? dev->cfg.burst_size : 32;
    if (pool_grow(dev, add) != BITAR_OK) return nullptr;
  }
  void* p = dev->free_slots.back();
  dev->free_slots.pop_back();
  dev->occupied.insert(p);
  return p;
}

int bitar_slot_take_n(bitar_dev* dev, uint32_t n, void** slots) {
  if (!dev || (!slots && n)) return fail(BITAR_E_INVALID, "bad argument");
  std::lock_guard<std::mutex> lock(dev->mu);
  if (dev->group > 1) {
    int rc = pool_take_groups(dev, n, slots);
    return rc ? fail(BITAR_E_IO_ERROR, "output slot pool exhausted") : BITAR_OK;
  }
  if (dev->free_slots.size() < n) {
    uint32_t add = n - (uint32_t)dev->free_slots.size();
    if (dev->grow_warned++ % 32 == 0)
      fprintf(stderr, "[bitar] WARNING: allocating %u output slots in the critical path (performance will be impacted)\n", add);
    // new slots must come out before older free ones to keep runs contiguous: grow, then rotate
    std::vector<void*> old;
    old.swap(dev->free_slots);
    int rc = pool_grow(dev, add);
    if (rc) {
      dev->free_slots.swap(old);
      return fail(BITAR_E_IO_ERROR, "output slot pool exhausted");
    }
    std::vector<void*> fresh;
    fresh.swap(dev->free_slots);
    dev->free_slots = old;
    dev->free_slots.insert(dev->free_slots.begin(), fresh.begin(), fresh.end());
  }
  for (uint32_t i = 0; i < n; ++i) {
    void* p = dev->free_slots.back();
    dev->free_slots.pop_back();
    dev->occupied.insert(p);
    slots[i] = p;
  }
  return BITAR_OK;
}

int bitar_slot_put(bitar_dev* dev, const void* addr) {
  if (!dev || !addr) return 0;
  std::lock_guard<std::mutex> lock(dev->mu);
  if (dev->group > 1) return pool_put_grouped(dev, addr);
  if (dev->occupied.erase(addr) == 0) return 0;  // not a slot we handed out: ignore (src/memory.cc:201-205)
  dev->free_slots.push_back(const_cast<void*>(addr));
  return 1;
}

uint32_t bitar_slot_put_n(bitar_dev* dev, const void* const* addrs, uint32_t n) {
  if (!dev || (!addrs && n)) return 0;
  std::lock_guard<std::mutex> lock(dev->mu);
  uint32_t back = 0;
  for (uint32_t i = n; i-- > 0;) {   // in reverse, as CompressDevice::Recycle walks its BufferVector (src/device.cc:320-327)
    if (!addrs[i]) continue;
    if (dev->group > 1) {
      back += (uint32_t)pool_put_grouped(dev, addrs[i]);
      continue;
    }
    if (dev->occupied.erase(addrs[i]) == 0) continue;
    dev->free_slots.push_back(const_cast<void*>(addrs[i]));
    ++back;
  }
  return back;
}

uint32_t bitar_slot_size(const bitar_dev* dev) { return dev ? dev->cfg.compressed_seg_size : 0; }

uint32_t bitar_slots_free(bitar_dev* dev) {
  if (!dev) return 0;
  std::lock_guard<std::mutex> lock(dev->mu);
  return (uint32_t)dev->free_slots.size() * dev->group;
}

int bitar_mem_alloc(int kind, int device_id, size_t size, size_t alignment, void** out) {
  if (!out) return fail(BITAR_E_INVALID, "null out");
  if (alignment > 256) return fail(BITAR_E_INVALID, "alignment %zu above the 256-byte allocation granularity", alignment);
  return alloc_kind(kind, device_id, size, out);
}

int bitar_mem_free(int kind, int device_id, void* ptr) { return free_kind(kind, device_id, ptr); }

int bitar_host_register(void* ptr, size_t size) {
  cudaError_t e = cudaHostRegister(ptr, size, cudaHostRegisterPortable | cudaHostRegisterMapped);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(e == cudaErrorMemoryAllocation ? BITAR_E_OUT_OF_MEMORY : BITAR_E_INVALID, "cudaHostRegister(%p, %zu): %s", ptr, size,
                cudaGetErrorString(e));
  }
  return BITAR_OK;
}

int bitar_host_unregister(void* ptr) {
  CU_TRY(cudaHostUnregister(ptr), BITAR_E_INVALID);
  return BITAR_OK;
}

int bitar_ptr_kind(const void* ptr, int* device_id) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, ptr) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  if (device_id) *device_id = a.device;
  if (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged) return 1;
  if (a.type == cudaMemoryTypeHost) return 2;
  return 0;
}

int bitar_mem_copy(void* dst, const void* src, size_t n) {
  if (n == 0) return BITAR_OK;
  if (!dst || !src) return fail(BITAR_E_INVALID, "null pointer");
  CU_TRY(cudaMemcpy(dst, src, n, cudaMemcpyDefault), BITAR_E_IO_ERROR);
  return BITAR_OK;
}

int bitar_current_device(int* device_id) {
  if (!device_id) return fail(BITAR_E_INVALID, "null pointer");
  CU_TRY(cudaGetDevice(device_id), BITAR_E_INVALID);
  return BITAR_OK;
}

int bitar_set_device(int device_id) {
  CU_TRY(cudaSetDevice(device_id), BITAR_E_INVALID);
  return BITAR_OK;
}

int bitar_qp_memcpy(bitar_dev* dev, uint16_t qp, void* dst, const void* src, size_t n) {
  if (!dev || qp >= dev->qps.size()) return fail(BITAR_E_INVALID, "bad device/queue pair");
  CU_TRY(cudaSetDevice(dev->id), BITAR_E_INVALID);
  CU_TRY(cudaMemcpyAsync(dst, src, n, cudaMemcpyDefault, dev->qps[qp]->stream), BITAR_E_IO_ERROR);
  return BITAR_OK;
}

const char* bitar_last_error(void) { return g_err; }
const char* bitar_version(void) { return "bitar-b200 0.1.0 (sm_100a)"; }

// not part of the public header: per-phase cycle counters of the deflate kernel (debug/tuning).
// enable=1 allocates and zeroes the counters; out (8 x u64, host) receives the current totals.
BITAR_API int bitar_debug_deflate_profile(int enable, unsigned long long* out) {
  unsigned long long* p = g_deflate_prof.load();
  if (enable && !p) {
    if (cudaMalloc((void**)&p, 16 * sizeof(unsigned long long)) != cudaSuccess) return BITAR_E_OUT_OF_MEMORY;
    cudaMemset(p, 0, 16 * sizeof(unsigned long long));
    g_deflate_prof.store(p);
  }
  if (p && out) {
    cudaDeviceSynchronize();
    cudaMemcpy(out, p, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    cudaMemset(p, 0, 16 * sizeof(unsigned long long));
  }
  if (!enable && p) g_deflate_prof.store(nullptr);
  return BITAR_OK;
}

#if defined(BITAR_LANE_DEBUG)
BITAR_API int bitar_debug_lane(unsigned int* out16) {
  cudaDeviceSynchronize();
  return (int)cudaMemcpyFromSymbol(out16, bitar::infl::g_dbg, 16 * sizeof(unsigned int));
}
#endif

// not part of the public header: the 8 work counters of the first batch of the queue pair's last inflate call, after it
// has completed (xk::Counters: [1] ops without an index, [7] of those, declined by the speculative kernel)
BITAR_API int bitar_debug_inflate_counters(bitar_dev* dev, uint16_t qp, unsigned int* out8) {
  if (!dev || qp >= dev->qps.size() || !out8) return BITAR_E_INVALID;
  QueuePair* q = dev->qps[qp];
  if (cudaSetDevice(dev->id) != cudaSuccess || cudaStreamSynchronize(q->stream) != cudaSuccess) return BITAR_E_IO_ERROR;
  return cudaMemcpy(out8, q->d_counter, 8 * sizeof(unsigned int), cudaMemcpyDeviceToHost) == cudaSuccess ? BITAR_OK : BITAR_E_IO_ERROR;
}

// not part of the public header: output bytes per range of the speculative inflate kernel (0 = default)
BITAR_API void bitar_tune_spec_target(int bytes) { g_spec_target.store(bytes); }
// not part of the public header: selects the inflate kernel instantiation for tuning sweeps
BITAR_API void bitar_tune_inflate_variant(int v) { g_inflate_variant.store(v); }
// not part of the public header: least inflated bytes per batch of a staged (host-buffer) inflate call; 0 = default
BITAR_API void bitar_tune_stage_batch(unsigned long long bytes) { g_stage_batch_bytes.store((size_t)bytes); }
// not part of the public header: 0 = gather staged inputs with the kernel even when they lie at a constant stride
BITAR_API void bitar_tune_stage_strided(int on) { g_stage_strided.store(on); }

}  // extern "C"
