/* flbgpu_internal.h -- seam between the host runtime (C) and the device back end.
 *
 * The product library links runtime.c + kernels.cu (CUDA, sm_100a) and nothing
 * else: there is no CPU implementation of bk_* in libflbgpu.so, so every entry
 * point fails loudly when no CUDA device is present.  tests/hostsim provides a
 * second implementation of the same functions for CPU-only CI of the host logic
 * and of the device algorithms; it is never linked into the product.
 */
#ifndef FLBGPU_INTERNAL_H
#define FLBGPU_INTERNAL_H

#include <stddef.h>
#include <stdint.h>
#include "flbgpu_prog.h"

#ifdef __cplusplus
extern "C" {
#endif

#define BK_INDEX_TILE   8192u     /* input bytes per index block */
#define BK_REC_BLOCK    256u      /* records per chain block */

struct bk_chain_args {
    const uint8_t *d_in;          /* input chunk (device) */
    uint32_t in_len;
    const uint8_t *d_blob;        /* chain program (device) */
    uint8_t *d_scr;               /* scratch or NULL */
    int32_t *d_capcache;          /* capture cache or NULL */
    uint32_t cap_stride;
    int64_t now;
    uint32_t assume;
    const uint32_t *d_off;        /* record index */
    const uint32_t *d_len;
    const uint8_t *d_kind;
    uint32_t n_rec;               /* records in the valid prefix */
    uint32_t *d_size;             /* [n_rec] output size per record */
    uint64_t *d_bsum;             /* [ceil(n_rec/BK_REC_BLOCK)] block sums -> exclusive offsets */
    uint32_t *d_flags;            /* [FLBGPU_MAX_FILTERS + 1]: CHF_* per filter, last = error word */
};

const char *bk_name(void);
int   bk_init(int device);                       /* 0 ok, -1 no usable device */
int   bk_device_count(void);
void *bk_alloc(size_t n);
void  bk_free(void *p);
void *bk_alloc_host(size_t n);                   /* pinned host memory */
void  bk_free_host(void *p);
int   bk_h2d(void *d, const void *h, size_t n);
int   bk_d2h(void *h, const void *d, size_t n);
int   bk_zero(void *d, size_t n);
int   bk_d2h_big(void *h_dst, const void *d_src, size_t n);   /* synchronous, pipelined through pinned slices */
int   bk_h2d_big(void *d_dst, const void *h_src, size_t n);   /* asynchronous on the library stream */
int   bk_sync(void);
void *bk_stream(void);
int   bk_kernel_ms(float out[3]);               /* CUDA-event ms of index / evaluate / emit in the last call */
const char *bk_last_error(void);

/* Record index (K1).  Pass 1 counts validated record candidates per tile and leaves
 * the exclusive tile offsets in d_tile; *n_cand gets the total (synchronises). */
int bk_index_count(const uint8_t *d_in, uint32_t len, uint32_t *d_tile, uint32_t n_tiles, uint32_t *n_cand);
/* Pass 2 writes (offset,length,kind) per candidate and checks that the candidates
 * tile [0,len) exactly.  *n_valid = records in the decodable prefix; *tiled = 1 when
 * the whole buffer is covered; returns -1 (FLBGPU_E_INDEX) when a candidate chain
 * breaks in the middle (nested record-shaped data), which this version refuses. */
int bk_index_fill(const uint8_t *d_in, uint32_t len, const uint32_t *d_tile, uint32_t n_tiles,
                  uint32_t n_cand, uint32_t *d_off, uint32_t *d_len, uint8_t *d_kind,
                  uint32_t *n_valid, int *tiled);

/* Chain evaluation pass: sizes + block sums + evidence.  h_flags receives
 * FLBGPU_MAX_FILTERS+1 words; *total the output size (synchronises). */
int bk_chain_size(const struct bk_chain_args *a, uint32_t *h_flags, uint64_t *total);
/* Chain emission pass into d_out (d_bsum holds exclusive block offsets). */
int bk_chain_emit(const struct bk_chain_args *a, uint8_t *d_out);

/* counters for bench.py's gpu_launches claim */
uint64_t bk_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif
