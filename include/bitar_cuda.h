/*
 * bitar_cuda.h -- C-ABI of the B200-native block DEFLATE engine (libbitar_cuda.so).
 *
 * This is the drop-in boundary: the symbols below are what a bitar back end binds in place of
 * DPDK's rte_compressdev.  Plain C types only (no C++, Arrow, torch or CUDA types in signatures);
 * pointers are device-accessible addresses: CUDA device memory, or pinned / registered host memory
 * (UVA), both of which the kernels read and write in place (zero-copy).
 *
 * Every entry point cites the reference interface (/root/reference, file:line) it replaces.
 * Return value: 0 on success or a NEGATIVE arrow::StatusCode, the convention of the reference's
 * internal int-returning functions (src/include/util.h:166-205, src/include/memory.h:113).
 */
#ifndef BITAR_CUDA_H_
#define BITAR_CUDA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define BITAR_API __attribute__((visibility("default")))
#else
#define BITAR_API
#endif

/* negative arrow::StatusCode values (arrow/status.h) */
#define BITAR_OK 0
#define BITAR_E_OUT_OF_MEMORY (-1)
#define BITAR_E_INVALID (-4)
#define BITAR_E_IO_ERROR (-5)
#define BITAR_E_CAPACITY (-6)
#define BITAR_E_CANCELLED (-8)
#define BITAR_E_UNKNOWN (-9)
#define BITAR_E_NOT_IMPLEMENTED (-10)

/* per-op status, the analogue of rte_comp_op::status checked at src/device.cc:512-520 */
#define BITAR_OP_OK 0u
#define BITAR_OP_OUT_OF_SPACE 1u /* RTE_COMP_OP_STATUS_OUT_OF_SPACE_TERMINATED */
#define BITAR_OP_DATA_ERROR 2u   /* RTE_COMP_OP_STATUS_ERROR: invalid DEFLATE stream */
#define BITAR_OP_TRUNCATED 3u    /* input ended inside a block */
#define BITAR_OP_NOT_RUN 0xFFFFFFFFu

/* rte_comp_huffman / rte_comp_checksum_type numbering, as stored by src/include/config.h:114-119,169-177 */
#define BITAR_HUFFMAN_DEFAULT 0
#define BITAR_HUFFMAN_FIXED 1
#define BITAR_HUFFMAN_DYNAMIC 2
#define BITAR_CHECKSUM_NONE 0
#define BITAR_CHECKSUM_CRC32 1
#define BITAR_CHECKSUM_ADLER32 2
#define BITAR_CHECKSUM_CRC32_ADLER32 3

/* memory kinds: replace MemoryPoolBackend::Rtemalloc / Rtememzone (src/include/memory_pool.h:65-71) */
#define BITAR_MEM_DEVICE 0 /* cudaMallocAsync from the device's pool */
#define BITAR_MEM_PINNED 1 /* cudaHostAlloc, portable + mapped (zero-copy over PCIe) */

/* segment-size limits: src/include/config.h:41-48 (kMinSegSize, kMaxSegSize), widened to 32 bit.
 * 59460 stays the reference-compatible default maximum; the engine accepts up to 1 MiB. */
#define BITAR_MIN_SEG_SIZE 8u
#define BITAR_REF_MAX_SEG_SIZE 59460u
#define BITAR_MAX_SEG_SIZE (1u << 20)
#define BITAR_MAX_SGL_SEGS 16u         /* segments chained into one stream; max_sgl_segs * S <= BITAR_MAX_SEG_SIZE */
#define BITAR_MIN_PREALLOCATE_SLOTS 20u /* kMinPreallocateMemzones, src/include/memory.h:51 */
#define BITAR_MAX_INFLIGHT_OPS 512u     /* kMaxInflightOps, src/include/memory.h:50 (informational) */

typedef struct bitar_dev bitar_dev; /* opaque: one CUDA device + its queue pairs + its slot pool */

/* rte_compressdev_info / capability as consumed by ValidateConfiguration, src/device.cc:352-415 */
typedef struct bitar_dev_info {
  int32_t device_id;
  int32_t cc_major, cc_minor;
  int32_t sm_count;
  uint64_t total_mem;
  uint32_t max_queue_pairs;
  uint8_t window_min, window_max; /* log2, 8..15 */
  uint8_t supports_fixed, supports_dynamic;
  uint8_t supports_crc32, supports_adler32;
  uint8_t supports_sgl; /* 1: max_sgl_segs up to BITAR_MAX_SGL_SEGS */
  uint8_t reserved;
  char name[64];
} bitar_dev_info;

/* Configuration<Class> + BlueFieldConfiguration fields, src/include/config.h:146-152,183 */
typedef struct bitar_cfg {
  uint32_t decompressed_seg_size;    /* S; 0 -> 2048 (kDefaultSegSize) */
  uint32_t compressed_seg_size;      /* slot bytes; 0 -> bitar_compressed_seg_size(S) */
  uint32_t max_preallocate_slots;    /* "max_preallocate_memzones"; 0 -> 2560 (RTE_MAX_MEMZONE) */
  uint16_t burst_size;               /* kept for API parity; the GPU path enqueues whole calls */
  uint16_t max_sgl_segs;             /* segments chained into one stream (src/include/config.h:90-96, src/memory.cc:394-398):
                                        1 .. 16, max_sgl_segs * S <= 1 MiB; the slots then lie back to back (a stream's
                                        output is one op whose dst spans its slots) */
  uint8_t window_size;               /* log2; 0 -> device max (15), src/device.cc:389-394; match distances never exceed 1 << window_size */
  uint8_t huffman_enc;               /* BITAR_HUFFMAN_*; DEFAULT -> DYNAMIC */
  uint8_t checksum_type;             /* BITAR_CHECKSUM_* */
  uint8_t slot_mem_kind;             /* BITAR_MEM_DEVICE or BITAR_MEM_PINNED */
  uint8_t no_index;                  /* 1: chunks are bare RFC 1951 streams (`produced` = stream length); 0: the parallel-
                                        inflate index follows the final block (deflate_common.h), which stock inflaters
                                        leave unread and which this engine's inflate needs for its fast path */
  uint8_t reserved[3];
} bitar_cfg;

/* One op: replaces an rte_comp_op with one src and one dst mbuf (src/memory.cc:350-430, 432-505).
 * Every op is stateless with flush FINAL (src/memory.cc:106-116): deflate emits one complete raw
 * DEFLATE stream per chunk, inflate consumes one. */
typedef struct bitar_chunk {
  const void* src;
  void* dst;
  uint32_t src_len;
  uint32_t dst_cap;
} bitar_chunk;

/* rte_comp_op result fields: produced, status, output_chksum (CRC-32 low word, Adler-32 high word) */
typedef struct bitar_result {
  uint32_t produced;
  uint32_t status;
  uint64_t checksum;
} bitar_result;

/* --- probe: replaces rte_compressdev_devices_get / rte_compressdev_info_get as used by
 *     CompressDriver::ListAvailableDeviceIds / GetDevices (src/driver.cc:173-223) --- */
BITAR_API int bitar_cuda_device_count(void);
BITAR_API int bitar_cuda_device_info(int device_id, bitar_dev_info* info);

/* Configuration::UpdateCompressedSegSize (src/config.cc:59-73), widened and made safe for
 * incompressible input: never below the stored-block bound S + 5*ceil(S/65535). */
BITAR_API uint32_t bitar_compressed_seg_size(uint32_t decompressed_seg_size);
/* The reference's own 16-bit formula, for parity checks. */
BITAR_API uint32_t bitar_reference_compressed_seg_size(uint32_t decompressed_seg_size);

/* --- device: replaces rte_compressdev_configure / queue_pair_setup / start and
 *     DeviceMemory / QueuePairMemory preallocation (src/device.cc:114-154, 429-441);
 *     close replaces rte_compressdev_stop / close (src/device.cc:329-343).
 *     A queue pair is a CUDA stream plus pinned/device descriptor and result rings. --- */
BITAR_API int bitar_dev_open(int device_id, uint16_t n_qps, const bitar_cfg* cfg, bitar_dev** out);
BITAR_API int bitar_dev_close(bitar_dev* dev);
BITAR_API int bitar_dev_config(const bitar_dev* dev, bitar_cfg* cfg_out); /* resolved values */
BITAR_API uint16_t bitar_dev_num_qps(const bitar_dev* dev);

/* --- data path: replace rte_compressdev_enqueue_burst / dequeue_burst (src/device.cc:464-535).
 *     Both calls ENQUEUE n ops on the queue pair's stream and return; `results` (host memory, n
 *     entries) is filled when the work completes and is valid after bitar_qp_wait() returns 0.
 *     Returns BITAR_E_CANCELLED when the queue pair still has pending ops (EntryGuard,
 *     src/device.cc:456-459).
 *     Buffers in pinned / registered host memory: deflate reads and writes them in place over PCIe;
 *     inflate stages them through device memory inside the call, in batches on up to three further
 *     streams of the queue pair that join its stream before the results are published (everything a
 *     caller orders after the call on bitar_qp_stream() still runs after it). --- */
BITAR_API int bitar_qp_deflate(bitar_dev* dev, uint16_t qp, const bitar_chunk* ops, uint32_t n,
                               bitar_result* results);
BITAR_API int bitar_qp_inflate(bitar_dev* dev, uint16_t qp, const bitar_chunk* ops, uint32_t n,
                               bitar_result* results);
/* Block until the queue pair drained (the busy-poll of src/device.cc:228-235); returns 0, or
 * BITAR_E_IO_ERROR when any op finished with status != BITAR_OP_OK (src/device.cc:512-520). */
BITAR_API int bitar_qp_wait(bitar_dev* dev, uint16_t qp);
/* Non-blocking bitar_qp_wait(): BITAR_E_CANCELLED while ops are pending, else what bitar_qp_wait() would return.
 * Makes no CUDA call, so it may be used inside a bitar_qp_on_complete() callback. */
BITAR_API int bitar_qp_result(bitar_dev* dev, uint16_t qp);
/* 1 while ops are pending (QueuePairMemory::has_pending_operations), else 0. */
BITAR_API int bitar_qp_busy(bitar_dev* dev, uint16_t qp);
/* Run fn(arg) on a driver thread once everything enqueued so far on the queue pair completed:
 * replaces the worker-lcore callback of LcoreCompressFunc (src/include/util.h:133-151). */
BITAR_API int bitar_qp_on_complete(bitar_dev* dev, uint16_t qp, void (*fn)(void*), void* arg);
/* Device time of the last deflate/inflate call on the queue pair: kernel only, and whole call
 * (descriptor upload + kernel + result download).  Valid after bitar_qp_wait(). */
BITAR_API int bitar_qp_last_ms(bitar_dev* dev, uint16_t qp, float* kernel_ms, float* total_ms);
/* The queue pair's cudaStream_t, for callers that order their own copies/events with the ops. */
BITAR_API void* bitar_qp_stream(bitar_dev* dev, uint16_t qp);
/* Number of kernel launches issued by this library since load (evidence counter for benchmarks). */
BITAR_API uint64_t bitar_kernel_launches(void);

/* --- output-slot pool: replaces DeviceMemory::Take / Put (src/memory.cc:160-209) and is what
 *     CompressDevice::Recycle returns buffers to (src/device.cc:320-327).  put returns 1 when the
 *     address was an occupied slot of this device, else 0. --- */
BITAR_API void* bitar_slot_take(bitar_dev* dev);
BITAR_API int bitar_slot_take_n(bitar_dev* dev, uint32_t n, void** slots);
BITAR_API int bitar_slot_put(bitar_dev* dev, const void* addr);
/* CompressDevice::Recycle(BufferVector) in one call (src/device.cc:320-327): puts the n addresses back in reverse order and
 * returns how many of them were occupied slots of this device. */
BITAR_API uint32_t bitar_slot_put_n(bitar_dev* dev, const void* const* addrs, uint32_t n);
BITAR_API uint32_t bitar_slot_size(const bitar_dev* dev);
BITAR_API uint32_t bitar_slots_free(bitar_dev* dev);

/* --- memory: replaces RtemallocAllocator / RtememzoneAllocator (src/memory_pool.cc:70-188) and
 *     rte_mem_virt2iova-style registration of caller memory (src/memory.cc:388,465,489). --- */
BITAR_API int bitar_mem_alloc(int kind, int device_id, size_t size, size_t alignment, void** out);
BITAR_API int bitar_mem_free(int kind, int device_id, void* ptr);
BITAR_API int bitar_host_register(void* ptr, size_t size);
BITAR_API int bitar_host_unregister(void* ptr);
/* What kind of memory an address is: 0 = pageable host (not device-accessible until registered),
 * 1 = device memory (*device_id receives the owning device), 2 = pinned / registered host memory.
 * The analogue of the rte_mem_virt2iova probe in src/memory.cc:388-391. */
BITAR_API int bitar_ptr_kind(const void* ptr, int* device_id);
/* Synchronous copy between any two device-accessible or host ranges (cudaMemcpyDefault): what the memory
 * pool's Reallocate uses in place of rte_memcpy (src/memory_pool.cc:151-174). */
BITAR_API int bitar_mem_copy(void* dst, const void* src, size_t n);
/* The calling thread's current CUDA device (device pools allocate there). */
BITAR_API int bitar_current_device(int* device_id);
/* Select the calling thread's current CUDA device (what the device pool allocates from next). */
BITAR_API int bitar_set_device(int device_id);
/* Plain copies on the queue pair's stream (host<->device staging for pageable buffers). */
BITAR_API int bitar_qp_memcpy(bitar_dev* dev, uint16_t qp, void* dst, const void* src, size_t n);

BITAR_API const char* bitar_last_error(void); /* thread-local, human readable */
BITAR_API const char* bitar_version(void);

#ifdef __cplusplus
}
#endif
#endif /* BITAR_CUDA_H_ */
