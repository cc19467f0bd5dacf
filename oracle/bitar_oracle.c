/*
 * bitar_oracle.c -- TEST INFRASTRUCTURE ONLY (parity oracle + CPU baseline).
 *
 * CPU restatement of the reference's hot path: bitar's chunking contract over
 * zlib raw DEFLATE, i.e. what `bitar demo_app` computes when DPDK's
 * `compress_zlib` software PMD sits behind rte_compressdev.
 *
 * Nothing in the product path (bitar_b200/, include/) may link, import or call
 * this file.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs use it, and only as the checker / reported baseline.
 *
 * PARITY UNPINNED: the reference ships no tests, fixtures or golden vectors
 * (/root/reference/test/CMakeLists.txt:23-26 is an empty scaffold), and its
 * sources cannot be compiled here (every file under src/ includes DPDK headers, e.g.
 * src/device.cc:31-37, src/memory.cc:33-43; DPDK, magic_enum, abseil and cxxopts
 * are absent and there is no network).  The codec arithmetic lives in a
 * third-party dependency that is NOT vendored under /root/reference:
 *   DPDK  >= 22.07#1 (vcpkg.json:45-52; minimum 21.11, CMakeLists.txt:120)
 *     -> drivers/compress/zlib (compress_zlib vdev) -> zlib (system, 1.3 here).
 * This oracle therefore calls the very same zlib the PMD would call, with the
 * parameters the reference resolves to, and is pinned by the known-answer facts
 * listed in SURVEY.md section 8(c) (tests/test_oracle.py) instead of by reference
 * fixtures.
 *
 * Reference call sites restated here:
 *   - segmenting / ordering of compress ops ... src/device.cc:156-238,
 *                                               src/memory.cc:350-430
 *   - decompress: segment i lands at i*S ...... src/device.cc:240-318,
 *                                               src/memory.cc:432-505
 *   - every op is stateless FLUSH_FINAL ....... src/memory.cc:106-116
 *   - xform: DEFLATE, level 1, window, huffman  src/config.cc:83-105
 *   - window 0 -> device max (15) ............. src/device.cc:389-394
 *   - compressed_seg_size formula ............. src/config.cc:59-73
 *   - seg limits 8 .. 59460 ................... src/include/config.h:41-48
 *
 * zlib mapping of the compress_zlib PMD [EXT, DPDK v22.07 zlib_pmd_ops.c]:
 *   deflateInit2(level, Z_DEFLATED, -window, 8, Z_DEFAULT_STRATEGY | Z_FIXED),
 *   deflate(Z_FINISH) per op, deflateReset between ops; inflateInit2(-window),
 *   inflate until Z_STREAM_END, inflateReset between ops.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <zlib.h>

#define ORACLE_API __attribute__((visibility("default")))

enum { ORACLE_HUFFMAN_FIXED = 1, ORACLE_HUFFMAN_DYNAMIC = 2 }; /* rte_comp_huffman order */

/* src/config.cc:59-73 -- Configuration::UpdateCompressedSegSize, 16-bit fields. */
ORACLE_API uint32_t oracle_compressed_seg_size(uint32_t decompressed_seg_size) {
  uint32_t lower_bound = decompressed_seg_size << 1;
  uint32_t num = 65536u;
  while ((num & lower_bound) == 0) {
    num >>= 1;
    if (num == 0) return 0;
  }
  if (num > 32768u) return (uint32_t)((double)decompressed_seg_size * 1.1);
  return num;
}

/* src/include/config.h:41-48 */
ORACLE_API uint32_t oracle_max_seg_size(void) { return (uint32_t)((65535 - 128) / 1.1); }
ORACLE_API uint32_t oracle_min_seg_size(void) { return 8; }

/* Worst-case size of a raw DEFLATE stream made only of stored blocks. */
ORACLE_API uint32_t oracle_stored_bound(uint32_t n) {
  uint32_t blocks = n == 0 ? 1 : (n + 65534u) / 65535u;
  return n + 5u * blocks;
}

/* One op = one complete raw DEFLATE stream (FLUSH_FINAL). Returns produced, or <0. */
static long deflate_one(z_stream* zs, const uint8_t* src, uint32_t n, uint8_t* dst,
                        uint32_t cap) {
  deflateReset(zs);
  zs->next_in = (Bytef*)src;
  zs->avail_in = n;
  zs->next_out = dst;
  zs->avail_out = cap;
  int rc = deflate(zs, Z_FINISH);
  if (rc != Z_STREAM_END) return -1; /* RTE_COMP_OP_STATUS_OUT_OF_SPACE_TERMINATED */
  return (long)(cap - zs->avail_out);
}

static long inflate_one(z_stream* zs, const uint8_t* src, uint32_t n, uint8_t* dst,
                        uint32_t cap) {
  inflateReset(zs);
  zs->next_in = (Bytef*)src;
  zs->avail_in = n;
  zs->next_out = dst;
  zs->avail_out = cap;
  int rc = inflate(zs, Z_FINISH);
  if (rc != Z_STREAM_END) return -1;
  return (long)(cap - zs->avail_out);
}

/*
 * Compress one chunk (single op).  level/window/huffman follow src/config.cc:83-91.
 * Returns produced bytes or -1.
 */
ORACLE_API long oracle_deflate_chunk(const uint8_t* src, uint32_t n, uint8_t* dst,
                                     uint32_t cap, int level, int window_log2,
                                     int huffman) {
  z_stream zs;
  memset(&zs, 0, sizeof zs);
  if (window_log2 == 0) window_log2 = 15; /* src/device.cc:389-394 */
  if (deflateInit2(&zs, level, Z_DEFLATED, -window_log2, 8,
                   huffman == ORACLE_HUFFMAN_FIXED ? Z_FIXED : Z_DEFAULT_STRATEGY) != Z_OK)
    return -2;
  long r = deflate_one(&zs, src, n, dst, cap);
  deflateEnd(&zs);
  return r;
}

ORACLE_API long oracle_inflate_chunk(const uint8_t* src, uint32_t n, uint8_t* dst,
                                     uint32_t cap, int window_log2) {
  z_stream zs;
  memset(&zs, 0, sizeof zs);
  if (window_log2 == 0) window_log2 = 15;
  if (inflateInit2(&zs, -window_log2) != Z_OK) return -2;
  long r = inflate_one(&zs, src, n, dst, cap);
  inflateEnd(&zs);
  return r;
}

ORACLE_API uint32_t oracle_crc32(const uint8_t* p, size_t n) {
  return (uint32_t)crc32(crc32(0L, Z_NULL, 0), p, (uInt)n);
}
ORACLE_API uint32_t oracle_adler32(const uint8_t* p, size_t n) {
  return (uint32_t)adler32(adler32(0L, Z_NULL, 0), p, (uInt)n);
}

/* ------------------------------------------------------------------------- */
/* Whole-buffer path: the n = ceil(bytes/S) op sequence of Compress/Decompress */
/* run by `threads` workers, each owning one z_stream that is reset between   */
/* ops ("one worker lcore per queue pair", src/driver.cc:198-220).            */
/* ------------------------------------------------------------------------- */

typedef struct {
  const uint8_t* src;
  uint64_t total;
  uint32_t seg;
  uint8_t* dst;        /* slot i at dst + i*slot */
  uint32_t slot;
  uint32_t* produced;  /* per chunk */
  int level, window_log2, huffman;
  uint64_t first, last; /* chunk range [first,last) */
  int rc;
} comp_job;

static void* comp_worker(void* arg) {
  comp_job* j = (comp_job*)arg;
  z_stream zs;
  memset(&zs, 0, sizeof zs);
  int w = j->window_log2 ? j->window_log2 : 15;
  if (deflateInit2(&zs, j->level, Z_DEFLATED, -w, 8,
                   j->huffman == ORACLE_HUFFMAN_FIXED ? Z_FIXED : Z_DEFAULT_STRATEGY) != Z_OK) {
    j->rc = -2;
    return NULL;
  }
  for (uint64_t i = j->first; i < j->last; ++i) {
    uint64_t off = i * (uint64_t)j->seg;
    uint32_t n = (uint32_t)((j->total - off) < j->seg ? (j->total - off) : j->seg);
    long r = deflate_one(&zs, j->src + off, n, j->dst + i * (uint64_t)j->slot, j->slot);
    if (r < 0) {
      j->rc = -1;
      break;
    }
    j->produced[i] = (uint32_t)r;
  }
  deflateEnd(&zs);
  return NULL;
}

/*
 * src/device.cc:156-238 restated.  dst must hold n_chunks*slot bytes; slot i receives the
 * stream of segment i (in input order, src/memory.cc:531-537).  Returns 0 or <0.
 */
ORACLE_API int oracle_compress_buffer(const uint8_t* src, uint64_t total, uint32_t seg,
                                      uint8_t* dst, uint32_t slot, uint32_t* produced,
                                      int level, int window_log2, int huffman, int threads) {
  if (total == 0) return 0; /* src/device.cc:161-164: empty input -> empty vector */
  uint64_t n = (total + seg - 1) / seg;
  if (threads < 1) threads = 1;
  if ((uint64_t)threads > n) threads = (int)n;
  pthread_t* th = (pthread_t*)calloc((size_t)threads, sizeof *th);
  comp_job* jobs = (comp_job*)calloc((size_t)threads, sizeof *jobs);
  /* contiguous near-equal ranges per worker, apps/demo_app.cc:579-596 */
  uint64_t per = n / (uint64_t)threads, rem = n % (uint64_t)threads, at = 0;
  for (int t = 0; t < threads; ++t) {
    uint64_t cnt = per + ((uint64_t)t < rem ? 1 : 0);
    comp_job j = {src, total, seg, dst, slot, produced, level, window_log2, huffman,
                  at,  at + cnt, 0};
    jobs[t] = j;
    at += cnt;
    pthread_create(&th[t], NULL, comp_worker, &jobs[t]);
  }
  int rc = 0;
  for (int t = 0; t < threads; ++t) {
    pthread_join(th[t], NULL);
    if (jobs[t].rc) rc = jobs[t].rc;
  }
  free(th);
  free(jobs);
  return rc;
}

typedef struct {
  const uint8_t* comp; /* slot i at comp + i*slot */
  uint32_t slot;
  const uint32_t* comp_len;
  uint8_t* out;
  uint32_t seg;
  uint32_t* produced;
  int window_log2;
  uint64_t first, last;
  int rc;
} decomp_job;

static void* decomp_worker(void* arg) {
  decomp_job* j = (decomp_job*)arg;
  z_stream zs;
  memset(&zs, 0, sizeof zs);
  int w = j->window_log2 ? j->window_log2 : 15;
  if (inflateInit2(&zs, -w) != Z_OK) {
    j->rc = -2;
    return NULL;
  }
  for (uint64_t i = j->first; i < j->last; ++i) {
    /* src/memory.cc:482-493: dst segment i = out + i*S, capacity S */
    long r = inflate_one(&zs, j->comp + i * (uint64_t)j->slot, j->comp_len[i],
                         j->out + i * (uint64_t)j->seg, j->seg);
    if (r < 0) {
      j->rc = -1;
      break;
    }
    j->produced[i] = (uint32_t)r;
  }
  inflateEnd(&zs);
  return NULL;
}

/* src/device.cc:240-318 restated.  out must hold n*seg bytes (CapacityError otherwise). */
ORACLE_API int oracle_decompress_buffer(const uint8_t* comp, uint32_t slot,
                                        const uint32_t* comp_len, uint64_t n, uint8_t* out,
                                        uint32_t seg, uint32_t* produced, int window_log2,
                                        int threads) {
  if (n == 0) return 0; /* src/device.cc:244-246 */
  if (threads < 1) threads = 1;
  if ((uint64_t)threads > n) threads = (int)n;
  pthread_t* th = (pthread_t*)calloc((size_t)threads, sizeof *th);
  decomp_job* jobs = (decomp_job*)calloc((size_t)threads, sizeof *jobs);
  uint64_t per = n / (uint64_t)threads, rem = n % (uint64_t)threads, at = 0;
  for (int t = 0; t < threads; ++t) {
    uint64_t cnt = per + ((uint64_t)t < rem ? 1 : 0);
    decomp_job j = {comp, slot, comp_len, out, seg, produced, window_log2, at, at + cnt, 0};
    jobs[t] = j;
    at += cnt;
    pthread_create(&th[t], NULL, decomp_worker, &jobs[t]);
  }
  int rc = 0;
  for (int t = 0; t < threads; ++t) {
    pthread_join(th[t], NULL);
    if (jobs[t].rc) rc = jobs[t].rc;
  }
  free(th);
  free(jobs);
  return rc;
}

ORACLE_API const char* oracle_zlib_version(void) { return zlibVersion(); }

ORACLE_API double oracle_now(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
