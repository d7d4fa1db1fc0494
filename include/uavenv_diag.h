/* uavenv_diag -- diagnostics of the observation stream's store mechanisms (libuavenv_diag.so).
 *
 * NOT part of the drop-in boundary (include/uavenv.h): these entry points have no counterpart in the reference.  They
 * are pure zero-fill kernels used by profiles/write_ceiling.py, ring_sweep.py and env_pattern_sweep.py to measure what
 * bounds the step kernel's dense-observation stream on a given box (profiles/r1/NOTES.md).
 */
#ifndef UAVENV_DIAG_H
#define UAVENV_DIAG_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { UAVENV_DIAG_OK = 0, UAVENV_DIAG_EINVAL = -1, UAVENV_DIAG_ECUDA = -2 };

/* Zero-fill `bytes` bytes (multiple of 16, 16-byte aligned device pointer) with
 * the store mechanism of the step kernel's observation stream -- mode 0: st.global.v4, mode 1: cp.async.bulk from a
 * zeroed shared-memory tile -- one CTA per bytes_per_cta. */
int uavenv_diag_fill(void *dst_dev, int64_t bytes, int64_t bytes_per_cta, int32_t mode, void *stream);

/* Diagnostic: the store-warp pattern of the step kernel in isolation -- `grid` persistent CTAs, one warp each, a ring
 * of `ring` shared-memory tiles of tile_bytes, chunk c (bytes_per_chunk) handled by CTA c % grid.  flags bit0: proxy
 * fence per tile, bit1: rotate the issuing lane. */
int uavenv_diag_fill_ring(void *dst_dev, int64_t bytes, int64_t bytes_per_chunk, int32_t grid, int32_t ring,
                          int32_t tile_bytes, int32_t flags, void *stream);

/* Diagnostic: the observation pattern of the step kernel's store warp in isolation -- per chunk: bulk copies of one
 * constant zero tile, then (flags bit0) n_red float REDs into the chunk whose copies have completed (the previous one,
 * or with flags bit1 the chunk itself). */
int uavenv_diag_fill_env(void *dst_dev, int64_t bytes, int64_t bytes_per_chunk, int32_t grid, int32_t tile_bytes,
                         int32_t flags, int32_t n_red, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* UAVENV_DIAG_H */
