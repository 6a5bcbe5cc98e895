/* kernels.cu -- sm_100a kernels and the CUDA implementation of the bk_* seam.
 *
 *   k_index<FILL> ..... K1 record index: every 0x92 byte is a record candidate; a
 *                       candidate is kept when a complete, well-formed log event
 *                       ([[ts, meta], body] or [ts, body]) starts there
 *                       (reference framing rules: src/flb_log_event_decoder.c:214-297).
 *                       Two passes (count -> prefix sum -> fill) keep file order.
 *   k_index_check ..... the kept candidates must tile the chunk; the first gap ends
 *                       the decodable prefix (the reference decoder stops there too).
 *   k_scan_top ........ exclusive prefix sum over per-block totals (one CTA).
 *   k_chain<EMIT> ..... one lane = one record through the filter-chain interpreter
 *                       (dev_chain.cuh); EMIT=false sizes, EMIT=true writes.
 *
 * This is byte-stream work bounded by HBM traffic and instruction issue; there is no
 * contraction here, so no tensor-core path.  Loads of chunk bytes are 128-bit and
 * coalesced in k_index; k_chain lanes walk adjacent records (L1/L2 resident lines).
 */
#include <cuda_runtime.h>
#include <atomic>
#include <thread>
#include <string.h>
#include <sched.h>
#include <stdint.h>
#include <stdio.h>
#include "flbgpu_internal.h"
#include "dev_chain.cuh"

static char g_err[256];
static unsigned long long g_launches;
static cudaStream_t g_stream;
/* CUDA-event timing of the three kernel groups of the last call (index, evaluate, emit) */
static cudaEvent_t g_ev[6];
static int g_ev_used[3];
static int g_ev_ready;
static void ev_begin(int k) { if (g_ev_ready) { cudaEventRecord(g_ev[2 * k], g_stream); } }
static void ev_end(int k) { if (g_ev_ready) { cudaEventRecord(g_ev[2 * k + 1], g_stream); g_ev_used[k] = 1; } }

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    snprintf(g_err, sizeof(g_err), "%s: %s", #call, cudaGetErrorString(e_)); return -1; } } while (0)

/* ---------------------------------------------------------------- helpers */
/* exclusive scan of one value per thread over a 256-thread block */
template <typename T>
__device__ __forceinline__ T block_excl_scan_t(T v, T *total)
{
    __shared__ T wsum[8];
    __shared__ T tot;
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T y = __shfl_up_sync(0xffffffffu, x, d);
        if (lane >= (unsigned) d) x += y;
    }
    if (lane == 31) wsum[warp] = x;
    __syncthreads();
    if (warp == 0) {
        T s = lane < 8 ? wsum[lane] : 0;
#pragma unroll
        for (int d = 1; d < 8; d <<= 1) {
            T y = __shfl_up_sync(0xffffffffu, s, d);
            if (lane >= (unsigned) d) s += y;
        }
        if (lane < 8) wsum[lane] = s;
        if (lane == 7) tot = s;
    }
    __syncthreads();
    T base = warp ? wsum[warp - 1] : 0;
    *total = tot;
    __syncthreads();
    return base + x - v;
}
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t *total)
{
    return block_excl_scan_t<uint32_t>(v, total);
}

/* ------------------------------------------------------------------ index */
template <bool FILL>
__global__ void __launch_bounds__(256) k_index(const uint8_t *__restrict__ in, uint32_t len,
                                               uint32_t *__restrict__ tile, uint32_t *__restrict__ o_off,
                                               uint32_t *__restrict__ o_len, uint8_t *__restrict__ o_kind)
{
    __shared__ uint16_t cand[BK_INDEX_TILE];
    __shared__ uint32_t v2ok[BK_INDEX_TILE / 32];   /* bit per tile byte: a valid v2 frame starts here */
    __shared__ uint32_t s_ncand;
    const uint32_t base = blockIdx.x * BK_INDEX_TILE;
    uint32_t ncand = 0;

    /* phase 1: ordered list of candidate positions in this tile */
    for (uint32_t it = 0; it < BK_INDEX_TILE / (256 * 16); it++) {
        const uint32_t rel = (it * 256 + threadIdx.x) * 16;
        const uint32_t pos = base + rel;
        uint32_t mask = 0;
        if (pos + 16 <= len) {
            const uint4 v = *reinterpret_cast<const uint4 *>(in + pos);   /* 128-bit coalesced load */
            const uint32_t w[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
            for (int k = 0; k < 16; k++) if (((w[k >> 2] >> (8 * (k & 3))) & 0xff) == 0x92) mask |= 1u << k;
        }
        else {
            for (int k = 0; k < 16; k++) if (pos + k < len && in[pos + k] == 0x92) mask |= 1u << k;
        }
        uint32_t tot;
        uint32_t at = block_excl_scan(__popc(mask), &tot) + ncand;
        while (mask) {
            int k = __ffs(mask) - 1;
            mask &= mask - 1;
            cand[at++] = (uint16_t) (rel + k);
        }
        ncand += tot;
    }
    if (threadIdx.x == 0) s_ncand = ncand;
    v2ok[threadIdx.x] = 0;
    __syncthreads();
    ncand = s_ncand;

    /* phase 2: validate candidates, 256 at a time, keeping order.  A candidate whose
     * predecessor byte starts a valid v2 frame is that frame's header array, not a
     * record (rec_is_shadowed); within a tile the predecessor's verdict is read from
     * shared memory (it belongs to an earlier or the same round: positions ascend). */
    uint32_t kept = 0;
    const uint32_t tile_base = FILL ? tile[blockIdx.x] : 0;
    for (uint32_t r = 0; r < ncand; r += 256) {
        const uint32_t j = r + threadIdx.x;
        uint32_t ok = 0, rlen = 0;
        int kind = 0;
        uint32_t pos = 0, rel = 0;
        if (j < ncand) {
            rel = cand[j];
            pos = base + rel;
            const uint8_t *e = rec_frame(in + pos, in + len, &kind);
            if (e) {
                ok = 1; rlen = (uint32_t) (e - (in + pos));
                if (in[pos + 1] == 0x92) atomicOr(&v2ok[rel >> 5], 1u << (rel & 31));
            }
        }
        __syncthreads();
        if (ok && pos > 0 && in[pos - 1] == 0x92) {
            if (rel > 0) { if ((v2ok[(rel - 1) >> 5] >> ((rel - 1) & 31)) & 1) ok = 0; }
            else if (rec_is_shadowed(in, in + pos, in + len)) ok = 0;      /* predecessor lives in the previous tile */
        }
        uint32_t tot;
        uint32_t at = block_excl_scan(ok, &tot);
        if (FILL && ok) {
            const uint32_t o = tile_base + kept + at;
            o_off[o] = pos; o_len[o] = rlen; o_kind[o] = (uint8_t) kind;
        }
        kept += tot;
    }
    if (!FILL && threadIdx.x == 0) tile[blockIdx.x] = kept;
}

/* exclusive scan of a[0..n) in place, one CTA; total in *out_total */
template <typename T>
__global__ void __launch_bounds__(256) k_scan_top(T *a, uint32_t n, unsigned long long *out_total)
{
    unsigned long long carry = 0;
    for (uint32_t b = 0; b < n; b += 256) {
        const uint32_t i = b + threadIdx.x;
        unsigned long long v = i < n ? (unsigned long long) a[i] : 0, tot;
        unsigned long long ex = block_excl_scan_t<unsigned long long>(v, &tot);
        if (i < n) a[i] = (T) (carry + ex);
        carry += tot;
    }
    if (threadIdx.x == 0) *out_total = carry;
}

#define BK_MAX_BREAKS 8192u

/* Links of the candidate list: candidate i must end where candidate i+1 starts (the
 * last one at the end of the chunk).  Broken links are rare -- a false candidate is a
 * byte run inside a real record that happens to frame as an event, e.g. the timestamp
 * bytes `.. 92 ce 00 00 01 a6 | 80` read as the legacy event [422, {}] -- and are
 * collected for k_index_repair. */
__global__ void k_index_check(const uint32_t *__restrict__ off, const uint32_t *__restrict__ rlen, uint32_t n,
                              uint32_t total, uint32_t *n_breaks, uint32_t *breaks)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t next = (i + 1 < n) ? off[i + 1] : total;
    if (off[i] + rlen[i] != next) {
        const uint32_t at = atomicAdd(n_breaks, 1u);
        if (at < BK_MAX_BREAKS) breaks[at] = i;
    }
}

/* One CTA: sort the broken links, then follow the record chain across them.  At a
 * break the chain jumps to the candidate that starts exactly where the current record
 * ends; the candidates jumped over were false and get kind 2 (ignored by k_chain).
 * If nothing starts there the decodable prefix ends (the reference decoder stops at
 * the first undecodable byte too).  res[0] = records in the prefix, res[1] = tiled. */
__global__ void __launch_bounds__(1024) k_index_repair(const uint32_t *__restrict__ off, const uint32_t *__restrict__ rlen,
                                                        uint8_t *kind, uint32_t n, uint32_t total,
                                                        const uint32_t *n_breaks, const uint32_t *breaks, uint32_t *res)
{
    __shared__ uint32_t sorted[BK_MAX_BREAKS];
    const uint32_t m = *n_breaks;
    for (uint32_t i = threadIdx.x; i < m; i += blockDim.x) {          /* rank sort: values are distinct */
        const uint32_t v = breaks[i];
        uint32_t r = 0;
        for (uint32_t j = 0; j < m; j++) r += breaks[j] < v;
        sorted[r] = v;
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    uint32_t n_valid = n, tiled = 1, skip_until = 0;
    if (off[0] != 0) { res[0] = 0; res[1] = 0; return; }
    for (uint32_t q = 0; q < m; q++) {
        const uint32_t b = sorted[q];
        if (b < skip_until) continue;                                  /* a candidate already ruled out */
        const uint32_t target = off[b] + rlen[b];
        uint32_t k = b + 1;
        while (k < n && off[k] < target) { kind[k] = 2; k++; }
        if (k < n && off[k] == target) { skip_until = k; continue; }
        if (k == n && target == total) { skip_until = n; continue; }
        n_valid = b + 1; tiled = 0;                                    /* nothing decodable starts at `target` */
        break;
    }
    res[0] = n_valid; res[1] = tiled;
}

/* ------------------------------------------------------------------ chain */
struct k_chain_params {
    struct ch_env env;
    const uint32_t *off, *len;
    const uint8_t *kind;
    uint32_t n_rec;
    uint32_t *size;
    uint64_t *bsum;
    uint8_t *out;
};

template <bool EMIT>
__global__ void __launch_bounds__(BK_REC_BLOCK) k_chain(const k_chain_params p)
{
    const uint32_t i = blockIdx.x * BK_REC_BLOCK + threadIdx.x;
    uint32_t sz = 0;
    if (!EMIT) {
        if (i < p.n_rec && p.kind[i] == 0) sz = chain_record<false>(&p.env, i, p.off[i], p.len[i], 0);
        if (i < p.n_rec) p.size[i] = sz;
        uint32_t tot;
        block_excl_scan(sz, &tot);
        if (threadIdx.x == 0) p.bsum[blockIdx.x] = tot;
    }
    else {
        if (i < p.n_rec) sz = p.size[i];
        uint32_t tot;
        const uint32_t local = block_excl_scan(sz, &tot);
        if (sz) chain_record<true>(&p.env, i, p.off[i], p.len[i], p.out + p.bsum[blockIdx.x] + local);
    }
}

/* ------------------------------------------------------------ bk_* seam */
extern "C" {

const char *bk_name(void) { return "cuda-sm_100a"; }
const char *bk_last_error(void) { return g_err; }
uint64_t bk_launch_count(void) { return g_launches; }

int bk_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int bk_init(int device)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        snprintf(g_err, sizeof(g_err), "no CUDA device available (%s); libflbgpu has no CPU path",
                 e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return -1;
    }
    if (device < 0 || device >= n) { snprintf(g_err, sizeof(g_err), "device %d out of range (0..%d)", device, n - 1); return -1; }
    CK(cudaSetDevice(device));
    if (!g_stream) CK(cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking));
    if (!g_ev_ready) { for (int i = 0; i < 6; i++) CK(cudaEventCreate(&g_ev[i])); g_ev_ready = 1; }
    /* the interpreter keeps its field list and backtrack stack in local memory */
    CK(cudaFuncSetCacheConfig(k_chain<false>, cudaFuncCachePreferL1));
    CK(cudaFuncSetCacheConfig(k_chain<true>, cudaFuncCachePreferL1));
    return 0;
}

void *bk_alloc(size_t n) { void *p = 0; if (cudaMalloc(&p, n ? n : 16) != cudaSuccess) { snprintf(g_err, sizeof(g_err), "cudaMalloc(%zu) failed", n); return 0; } return p; }
void bk_free(void *p) { if (p) cudaFree(p); }
void *bk_alloc_host(size_t n) { void *p = 0; if (cudaMallocHost(&p, n ? n : 16) != cudaSuccess) return 0; return p; }
void bk_free_host(void *p) { if (p) cudaFreeHost(p); }
int bk_h2d(void *d, const void *h, size_t n) { CK(cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, g_stream)); return 0; }
int bk_d2h(void *h, const void *d, size_t n) { CK(cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, g_stream)); return 0; }
int bk_zero(void *d, size_t n) { CK(cudaMemsetAsync(d, 0, n, g_stream)); return 0; }
int bk_sync(void) { CK(cudaStreamSynchronize(g_stream)); return 0; }
void *bk_stream(void) { return (void *) g_stream; }

/* milliseconds of [index, evaluate (last pass), emit] of the last call; needs a prior bk_sync() */
int bk_kernel_ms(float out[3])
{
    for (int k = 0; k < 3; k++) {
        out[k] = 0.f;
        if (g_ev_used[k] && cudaEventElapsedTime(&out[k], g_ev[2 * k], g_ev[2 * k + 1]) != cudaSuccess) out[k] = -1.f;
    }
    return 0;
}


/* ---- large device->host results -------------------------------------------------
 * cb_filter must hand back a malloc()ed (pageable) buffer.  A plain cudaMemcpy into
 * pageable memory is staged by the driver on one thread and pays a page fault per
 * 4 KB of the fresh allocation.  Here the result is DMA'd in 16 MB slices into a ring
 * of pinned buffers on a copy stream while one host thread per ring slot moves the
 * finished slices into the destination, so the DMA and the (parallel, page-faulting)
 * host copies overlap. */
#define XF_SLOTS 8
#define XF_SLICE ((size_t) 16 << 20)
static uint8_t *xf_ring[XF_SLOTS];
static cudaEvent_t xf_ev[XF_SLOTS], xf_evc;
static cudaStream_t g_copy;
static int xf_ready;

static int xf_init(void)
{
    if (xf_ready) return 0;
    CK(cudaStreamCreateWithFlags(&g_copy, cudaStreamNonBlocking));
    for (int i = 0; i < XF_SLOTS; i++) {
        CK(cudaMallocHost((void **) &xf_ring[i], XF_SLICE));
        CK(cudaEventCreateWithFlags(&xf_ev[i], cudaEventDisableTiming));
    }
    CK(cudaEventCreateWithFlags(&xf_evc, cudaEventDisableTiming));
    xf_ready = 1;
    return 0;
}

int bk_d2h_big(void *h_dst, const void *d_src, size_t n)
{
    if (n < (4u << 20)) {
        CK(cudaMemcpyAsync(h_dst, d_src, n, cudaMemcpyDeviceToHost, g_stream));
        CK(cudaStreamSynchronize(g_stream));
        return 0;
    }
    if (xf_init()) return -1;
    CK(cudaEventRecord(xf_evc, g_stream));              /* the copy stream starts after the emission kernel */
    CK(cudaStreamWaitEvent(g_copy, xf_evc, 0));
    const size_t n_slices = (n + XF_SLICE - 1) / XF_SLICE;
    std::atomic<long> issued[XF_SLOTS], done[XF_SLOTS];
    for (int s = 0; s < XF_SLOTS; s++) { issued[s].store(-1); done[s].store(-1); }
    std::atomic<int> failed(0);
    std::thread workers[XF_SLOTS];
    const int nw = (int) (n_slices < XF_SLOTS ? n_slices : XF_SLOTS);
    for (int s = 0; s < nw; s++) {
        workers[s] = std::thread([&, s]() {
            for (size_t i = s; i < n_slices; i += XF_SLOTS) {
                while (issued[s].load(std::memory_order_acquire) < (long) i) { if (failed.load()) return; sched_yield(); }
                if (cudaEventSynchronize(xf_ev[s]) != cudaSuccess) { failed.store(1); return; }
                const size_t off = i * XF_SLICE, sz = (off + XF_SLICE <= n) ? XF_SLICE : n - off;
                memcpy((uint8_t *) h_dst + off, xf_ring[s], sz);
                done[s].store((long) i, std::memory_order_release);
            }
        });
    }
    int rc = 0;
    for (size_t i = 0; i < n_slices && !failed.load(); i++) {
        const int s = (int) (i % XF_SLOTS);
        if (i >= XF_SLOTS) while (done[s].load(std::memory_order_acquire) < (long) (i - XF_SLOTS)) { if (failed.load()) break; sched_yield(); }
        const size_t off = i * XF_SLICE, sz = (off + XF_SLICE <= n) ? XF_SLICE : n - off;
        if (cudaMemcpyAsync(xf_ring[s], (const uint8_t *) d_src + off, sz, cudaMemcpyDeviceToHost, g_copy) != cudaSuccess ||
            cudaEventRecord(xf_ev[s], g_copy) != cudaSuccess) { failed.store(1); break; }
        issued[s].store((long) i, std::memory_order_release);
    }
    for (int s = 0; s < nw; s++) workers[s].join();
    if (failed.load()) { snprintf(g_err, sizeof(g_err), "device->host transfer failed: %s", cudaGetErrorString(cudaGetLastError())); rc = -1; }
    return rc;
}

/* host->device: pinned (or registered) memory is DMA'd as is; pageable memory goes
 * through the driver's staging path */
int bk_h2d_big(void *d_dst, const void *h_src, size_t n)
{
    CK(cudaMemcpyAsync(d_dst, h_src, n, cudaMemcpyHostToDevice, g_stream));
    return 0;
}

static unsigned long long *g_dtotal;   /* device scratch for totals / first_break */
static int ensure_small(void)
{
    if (!g_dtotal) CK(cudaMalloc((void **) &g_dtotal, 64));
    return 0;
}

int bk_index_count(const uint8_t *d_in, uint32_t len, uint32_t *d_tile, uint32_t n_tiles, uint32_t *n_cand)
{
    unsigned long long tot = 0;
    if (ensure_small()) return -1;
    *n_cand = 0;
    g_ev_used[0] = g_ev_used[1] = g_ev_used[2] = 0;
    if (n_tiles == 0) return 0;
    ev_begin(0);
    k_index<false><<<n_tiles, 256, 0, g_stream>>>(d_in, len, d_tile, 0, 0, 0);
    k_scan_top<uint32_t><<<1, 256, 0, g_stream>>>(d_tile, n_tiles, g_dtotal);
    g_launches += 2;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(&tot, g_dtotal, sizeof(tot), cudaMemcpyDeviceToHost, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    *n_cand = (uint32_t) tot;
    return 0;
}

int bk_index_fill(const uint8_t *d_in, uint32_t len, const uint32_t *d_tile, uint32_t n_tiles, uint32_t n_cand,
                  uint32_t *d_off, uint32_t *d_len, uint8_t *d_kind, uint32_t *n_valid, int *tiled)
{
    static uint32_t *d_breaks;
    uint32_t h[4] = { 0, 0, 0, 0 };
    uint32_t *d_w = (uint32_t *) (g_dtotal + 1);          /* [0] n_breaks, [1] n_valid, [2] tiled */
    *n_valid = 0; *tiled = (len == 0);
    if (n_cand == 0) return 0;
    if (!d_breaks) CK(cudaMalloc((void **) &d_breaks, sizeof(uint32_t) * BK_MAX_BREAKS));
    k_index<true><<<n_tiles, 256, 0, g_stream>>>(d_in, len, (uint32_t *) d_tile, d_off, d_len, d_kind);
    CK(cudaMemsetAsync(d_w, 0, 16, g_stream));
    k_index_check<<<(n_cand + 255) / 256, 256, 0, g_stream>>>(d_off, d_len, n_cand, len, d_w, d_breaks);
    k_index_repair<<<1, 1024, 0, g_stream>>>(d_off, d_len, d_kind, n_cand, len, d_w, d_breaks, d_w + 1);
    ev_end(0);
    g_launches += 3;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(h, d_w, sizeof(h), cudaMemcpyDeviceToHost, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    if (h[0] > BK_MAX_BREAKS) {
        snprintf(g_err, sizeof(g_err), "record index: %u broken candidate links in one chunk (limit %u)", h[0], BK_MAX_BREAKS);
        return -1;
    }
    *n_valid = h[1];
    *tiled = (int) h[2];
    return 0;
}

static void fill_params(const struct bk_chain_args *a, k_chain_params *p, uint8_t *d_out)
{
    p->env.in = a->d_in; p->env.in_len = a->in_len; p->env.blob = a->d_blob; p->env.scr = a->d_scr;
    p->env.capcache = a->d_capcache; p->env.cap_stride = a->cap_stride; p->env.now = a->now;
    p->env.assume = a->assume; p->env.fl_flags = a->d_flags; p->env.err = a->d_flags + FLBGPU_MAX_FILTERS;
    p->off = a->d_off; p->len = a->d_len; p->kind = a->d_kind; p->n_rec = a->n_rec;
    p->size = a->d_size; p->bsum = a->d_bsum; p->out = d_out;
}

int bk_chain_size(const struct bk_chain_args *a, uint32_t *h_flags, uint64_t *total)
{
    k_chain_params p;
    unsigned long long tot = 0;
    const uint32_t nb = (a->n_rec + BK_REC_BLOCK - 1) / BK_REC_BLOCK;
    if (ensure_small()) return -1;
    *total = 0;
    CK(cudaMemsetAsync(a->d_flags, 0, sizeof(uint32_t) * (FLBGPU_MAX_FILTERS + 1), g_stream));
    if (nb) {
        fill_params(a, &p, 0);
        ev_begin(1);
        k_chain<false><<<nb, BK_REC_BLOCK, 0, g_stream>>>(p);
        ev_end(1);
        k_scan_top<uint64_t><<<1, 256, 0, g_stream>>>((uint64_t *) a->d_bsum, nb, g_dtotal);
        g_launches += 2;
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(&tot, g_dtotal, sizeof(tot), cudaMemcpyDeviceToHost, g_stream));
    }
    CK(cudaMemcpyAsync(h_flags, a->d_flags, sizeof(uint32_t) * (FLBGPU_MAX_FILTERS + 1), cudaMemcpyDeviceToHost, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    *total = tot;
    return 0;
}

int bk_chain_emit(const struct bk_chain_args *a, uint8_t *d_out)
{
    k_chain_params p;
    const uint32_t nb = (a->n_rec + BK_REC_BLOCK - 1) / BK_REC_BLOCK;
    if (!nb) return 0;
    fill_params(a, &p, d_out);
    ev_begin(2);
    k_chain<true><<<nb, BK_REC_BLOCK, 0, g_stream>>>(p);
    ev_end(2);
    g_launches += 1;
    CK(cudaGetLastError());
    return 0;
}

}
