/* bitar_cuda_testing.h -- TEST / PROFILING hooks of libbitar_cuda.so.  Not part of the drop-in boundary
 * (include/bitar_cuda.h): nothing in bitar's API maps to them; tests/ and tools/ use them to force code paths that a
 * call would otherwise pick by itself, and to read the deflate kernel's phase timers. */
#ifndef BITAR_CUDA_TESTING_H
#define BITAR_CUDA_TESTING_H

#ifdef __cplusplus
extern "C" {
#endif

/* inflate path of later calls: 0 = default (two-phase kernel for indexed chunks, whole-stream kernel for the rest),
 * 5 = everything through the whole-stream kernel (tests/test_gpu_inflate.py) */
void bitar_tune_inflate_variant(int v);
/* least inflated bytes per batch of a staged (host-buffer) inflate call; 0 = default (tests force many small batches) */
void bitar_tune_stage_batch(unsigned long long bytes);
/* 0 = gather staged inputs with the copy kernel even when they lie at a constant stride; 1 = default */
void bitar_tune_stage_strided(int on);
/* phase timers of the deflate kernel (thread-0 clock64 sums, tools/gpu_deflate_prof.py): enable = 1 starts, 0 stops and
 * copies 16 counters to out16 (may be NULL) */
int bitar_debug_deflate_profile(int enable, unsigned long long* out16);

#ifdef __cplusplus
}
#endif
#endif
