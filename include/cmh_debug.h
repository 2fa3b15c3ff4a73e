/* cmh_debug.h — tuning and trace hooks of libcmh.so.  NOT part of the product ABI (include/cmh.h): they set PROCESS-GLOBAL state
 * that every following cmh_gemm_bf16 call of every thread and stream reads, so they are for single-threaded tests, benchmarks
 * and profiling scripts only (tests/test_gpu_gemm.py, scripts/gemm_*.py).  The library's defaults are restored by passing 0
 * (force_tile, force_units), NULL (set_trace), 2 (mma_lookahead) and 1 (tail_slicing). */
#ifndef CMH_DEBUG_H_
#define CMH_DEBUG_H_

#ifdef __cplusplus
extern "C" {
#endif

/* Test / tuning hook: pin the tile width (128, 192, 256; 0 = automatic) and the CTA group (1 = one SM per tile,
 * 2 = tcgen05 cta_group::2 pairs on 256-row tiles; 0 = automatic) of every following cmh_gemm_bf16 call. */
int cmh_gemm_force_tile(int bn, int cta_group);
/* Debug timeline: when non-NULL, every following GEMM writes SM clock stamps into device_buffer[cta][64]
 * (0 entry, 1 set-up done, 2+4i.. per tile: accumulator wait / free / first operands landed / all MMAs issued,
 * 34+2i.. epilogue start / end of tile i, 63 exit). */
int cmh_gemm_set_trace(long long* device_buffer);
int cmh_gemm_mma_lookahead(int kblocks); /* tuning: k-blocks (4 MMAs each) the issuer may queue ahead, 1..8 (default 2) */
int cmh_gemm_tail_slicing(int on);   /* debug: 0 disables the column slicing of the last partial wave's tiles */
int cmh_gemm_force_units(int units); /* debug: cap the persistent grid at `units` CTAs (pairs for cta_group 2); 0 = all SMs */

#ifdef __cplusplus
}
#endif

#endif /* CMH_DEBUG_H_ */
