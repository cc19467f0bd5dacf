/* bitar_cuda_testing.h -- TEST / PROFILING hooks of libbitar_cuda.so.  Not part of the drop-in boundary
 * (include/bitar_cuda.h): nothing in bitar's API maps to them; tests/ and tools/ use them to force code paths that a
 * call would otherwise pick by itself, and to read the deflate kernel's phase timers. */
#ifndef BITAR_CUDA_TESTING_H
#define BITAR_CUDA_TESTING_H

#ifdef __cplusplus
extern "C" {
#endif

/* inflate path of later calls: 0 = default (two-phase kernel for indexed chunks, whole-stream kernel for the rest),
 * 5 = everything through the whole-stream kernel (tests/test_gpu_inflate.py),
 * 6 = no speculative kernel: chunks without an index go straight to the whole-stream kernel */
void bitar_tune_inflate_variant(int v);
/* output bytes each lane of the speculative inflate kernel aims at per round (64 .. 1536; 0 = default, 1024) */
void bitar_tune_spec_target(int bytes);
/* least inflated bytes per batch of a staged (host-buffer) inflate call; 0 = default (tests force many small batches) */
void bitar_tune_stage_batch(unsigned long long bytes);
/* 0 = gather staged inputs with the copy kernel even when they lie at a constant stride; 1 = default */
void bitar_tune_stage_strided(int on);
/* phase timers of the deflate kernel (thread-0 clock64 sums, tools/gpu_deflate_prof.py): enable = 1 starts, 0 stops and
 * copies 16 counters to out16 (may be NULL) */
int bitar_debug_deflate_profile(int enable, unsigned long long* out16);
/* the 8 work counters of the first batch of the queue pair's last inflate call, read after it completed: [0] 64 KiB tasks
 * of indexed chunks, [1] chunks without an index, [7] of those, declined by the speculative kernel (decoded by the
 * whole-stream kernel instead) */
struct bitar_dev;
int bitar_debug_inflate_counters(struct bitar_dev* dev, unsigned short qp, unsigned int* out8);

#ifdef __cplusplus
}
#endif
#endif
