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
#include <stdlib.h>
#include <sched.h>
#include <stdint.h>
#include <stdio.h>
#include "flbgpu_internal.h"
#include "dev_chain.cuh"

static char g_err[256];
static int g_nsurv;
static unsigned long long g_launches;
static cudaStream_t g_stream;
/* CUDA-event timing of the kernel groups of the last call: one event pair per launch
 * (0 = index, 1 = evaluate, 2 = emit); bk_kernel_ms() sums the pairs of a group. */
#define EV_MAX 1024
static cudaEvent_t g_evp[3][EV_MAX][2];
static int g_ev_made[3], g_ev_used[3];
static int g_ev_ready;
static cudaStream_t g_istream;
static void ev_begin_on(int k, cudaStream_t st)
{
    if (!g_ev_ready || g_ev_used[k] >= EV_MAX) return;
    if (g_ev_used[k] >= g_ev_made[k]) {
        cudaEventCreate(&g_evp[k][g_ev_made[k]][0]); cudaEventCreate(&g_evp[k][g_ev_made[k]][1]);
        g_ev_made[k]++;
    }
    cudaEventRecord(g_evp[k][g_ev_used[k]][0], st);
}
static void ev_end_on(int k, cudaStream_t st)
{
    if (!g_ev_ready || g_ev_used[k] >= EV_MAX) return;
    cudaEventRecord(g_evp[k][g_ev_used[k]][1], st);
    g_ev_used[k]++;
}
static void ev_begin(int k) { ev_begin_on(k, g_stream); }
static void ev_end(int k) { ev_end_on(k, g_stream); }

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    snprintf(g_err, sizeof(g_err), "%s: %s", #call, cudaGetErrorString(e_)); return -1; } } while (0)

/* 4 resident blocks per SM (64 registers per lane): measured best on B200 -- the interpreter is bound
 * by instruction fetch and load latency, which more resident warps hide better than fewer spills do
 * (profiles/r01_variants.txt) */
#ifndef BK_EVAL_MIN_BLOCKS
#define BK_EVAL_MIN_BLOCKS 4
#endif

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
__global__ void __launch_bounds__(256) k_index(const uint8_t *__restrict__ in, uint32_t len, uint32_t skip, uint32_t abs_base,
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
        /* `in` is the slice start rounded down to 16 bytes: bytes before `skip` belong to the previous slice */
        if (pos < skip) mask &= (pos + 16 <= skip) ? 0u : (0xffffu << (skip - pos));
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
            o_off[o] = abs_base + pos; o_len[o] = rlen; o_kind[o] = (uint8_t) kind;
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
                              uint32_t total /* absolute end of the slice */, uint32_t *n_breaks, uint32_t *breaks)
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
                                                        uint8_t *kind, uint32_t n, uint32_t base, uint32_t total,
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
    if (off[0] != base) { res[0] = 0; res[1] = 0; res[2] = base; return; }
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
    /* where the chain ends: the next slice starts here */
    {
        uint32_t last = n_valid;
        while (last > 0 && kind[last - 1] == 2) last--;
        res[2] = last ? off[last - 1] + rlen[last - 1] : base;
    }
}

/* ------------------------------------------------------------------ chain */
struct k_chain_params {
    struct ch_env env;
    const uint32_t *off, *len;
    const uint8_t *kind;
    uint32_t r0;                 /* first record / first block of this launch */
    uint32_t stage_bytes;        /* shared-memory slice per warp for input staging, 0 = none */
    uint32_t n_rec;
    uint32_t *size;
    uint64_t *bsum;
    uint8_t *out;
};

/* evaluation: record r0 + global thread id.  No barrier: a warp retires as soon as its
 * 32 records are done (block-level reductions happen in k_bsum). */

/* Optional staging (stage_bytes > 0, FLBGPU_STAGE_KB): the 32 records of a warp are adjacent in the
 * chunk, so the warp first copies their bytes to its own slice of shared memory with coalesced 128-bit
 * loads and the lanes then scan from there; `in` is rebased so that input offsets keep their meaning.
 * A warp whose records span more than its slice reads global memory as before. */
__global__ void __launch_bounds__(BK_REC_BLOCK, BK_EVAL_MIN_BLOCKS) k_chain_eval(const k_chain_params p)
{
    extern __shared__ __align__(16) uint8_t dsm[];
    const uint32_t i = p.r0 + blockIdx.x * BK_REC_BLOCK + threadIdx.x;
    const bool valid = i < p.n_rec;
    const uint32_t my_off = valid ? p.off[i] : 0, my_len = valid ? p.len[i] : 0;
    const bool live = valid && p.kind[i] == 0;
    const uint8_t *in = p.env.in;
    if (p.stage_bytes) {
        const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const uint32_t lo = __reduce_min_sync(0xffffffffu, live ? my_off : 0xffffffffu) & ~15u;
        const uint32_t hi = __reduce_max_sync(0xffffffffu, live ? my_off + my_len : 0u);
        if (hi > lo && hi - lo + 32 <= p.stage_bytes) {
            uint8_t *buf = dsm + (size_t) warp * p.stage_bytes;
            for (uint32_t o = lane * 16; o < hi - lo + 16; o += 512)
                *reinterpret_cast<uint4 *>(buf + o) = *reinterpret_cast<const uint4 *>(in + lo + o);
            __syncwarp();
            in = buf - lo;
        }
    }
    if (!valid) return;
    uint32_t sz = 0;
    if (live) {
        struct ch_env le = p.env;
        le.in = in;
        sz = chain_record<false>(&le, i, my_off, my_len, 0);
    }
    __stcs(&p.size[i], sz);
}

/* Chains with a log_to_metrics filter only: the events the decoder steps over (kind 1: group markers, negative
 * timestamps) are counted by that filter as long as nothing before it rewrote the chunk (chain_skipped_record).
 * A kernel of its own so that k_chain_eval stays what it is for every other chain. */
__global__ void __launch_bounds__(BK_REC_BLOCK) k_chain_skipped(const k_chain_params p)
{
    const uint32_t i = p.r0 + blockIdx.x * BK_REC_BLOCK + threadIdx.x;
    if (i >= p.n_rec || p.kind[i] != 1) return;
    chain_skipped_record(&p.env, i, p.off[i], p.len[i]);
}

/* per-block sums of the record sizes */
__global__ void __launch_bounds__(BK_REC_BLOCK) k_bsum(const uint32_t *__restrict__ size, uint32_t n, uint64_t *__restrict__ bsum)
{
    const uint32_t i = blockIdx.x * BK_REC_BLOCK + threadIdx.x;
    uint32_t tot;
    block_excl_scan(i < n ? size[i] : 0, &tot);
    if (threadIdx.x == 0) bsum[blockIdx.x] = tot;
}

/* ---- emission over a dense list of survivors ----
 * After a selective grep most records of a block are gone; packing the survivors of the whole range
 * (not just of one block) keeps every warp of the emission kernel full.
 *   k_surv_count: survivors per block            k_scan_top<uint32_t>: rank base per block
 *   k_surv_fill : (record, output offset) at its rank      k_chain_emit_list: one lane per rank */
__global__ void __launch_bounds__(BK_REC_BLOCK) k_surv_count(const uint32_t *__restrict__ size, uint32_t n, uint32_t *__restrict__ cnt)
{
    const uint32_t i = blockIdx.x * BK_REC_BLOCK + threadIdx.x;
    uint32_t tot;
    block_excl_scan((i < n && size[i]) ? 1u : 0u, &tot);
    if (threadIdx.x == 0) cnt[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(BK_REC_BLOCK) k_surv_fill(const uint32_t *__restrict__ size, uint32_t n, uint32_t rec0,
                                                            const uint32_t *__restrict__ base, const uint64_t *__restrict__ bsum,
                                                            uint32_t *__restrict__ l_rec, uint64_t *__restrict__ l_off)
{
    const uint32_t i = blockIdx.x * BK_REC_BLOCK + threadIdx.x;
    const uint32_t sz = i < n ? size[i] : 0;
    uint32_t tot, nsurv;
    const uint32_t local = block_excl_scan(sz, &tot);
    const uint32_t rank = block_excl_scan(sz ? 1u : 0u, &nsurv);
    if (sz) {
        l_rec[base[blockIdx.x] + rank] = rec0 + i;
        l_off[base[blockIdx.x] + rank] = bsum[blockIdx.x] + local;
    }
}

__global__ void __launch_bounds__(BK_REC_BLOCK, BK_EVAL_MIN_BLOCKS) k_chain_emit_list(const k_chain_params p, const uint32_t *__restrict__ l_rec,
                                                                                      const uint64_t *__restrict__ l_off,
                                                                                      const unsigned long long *__restrict__ n_list)
{
    const uint32_t t = blockIdx.x * BK_REC_BLOCK + threadIdx.x;
    if (t >= (uint32_t) *n_list) return;
    const uint32_t r = l_rec[t];
    chain_record<true>(&p.env, r, p.off[r], p.len[r], p.out + l_off[t]);
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

/* Keep this process's host side next to its GPU: the threads that stage, copy out and allocate the
 * pinned rings run on the CPUs local to the device's PCIe root (sysfs local_cpulist), so first-touch
 * places those buffers on that NUMA node.  Opt-in (FLBGPU_NUMA_BIND=1): on the bench box, with four
 * GPUs behind one socket, confining four processes to that socket lost 15 % end to end. */
static void bind_near_device(int device)
{
    char bus[32], path[128], line[1024];
    const char *e = getenv("FLBGPU_NUMA_BIND");
    FILE *f;
    cpu_set_t set;
    int any = 0;
    if (!(e && e[0] == '1')) return;
    if (cudaDeviceGetPCIBusId(bus, sizeof(bus), device) != cudaSuccess) { cudaGetLastError(); return; }
    for (char *c = bus; *c; c++) if (*c >= 'A' && *c <= 'F') *c = (char) (*c + 32);
    snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/local_cpulist", bus);
    f = fopen(path, "r");
    if (!f) return;
    if (!fgets(line, sizeof(line), f)) { fclose(f); return; }
    fclose(f);
    CPU_ZERO(&set);
    for (char *p = line; *p && *p != '\n'; ) {            /* "0-31,64-95" */
        char *end;
        long a = strtol(p, &end, 10), b = a;
        if (end == p) break;
        if (*end == '-') { p = end + 1; b = strtol(p, &end, 10); }
        for (long c = a; c <= b && c < CPU_SETSIZE; c++) { CPU_SET((int) c, &set); any = 1; }
        p = (*end == ',') ? end + 1 : end;
    }
    if (any) sched_setaffinity(0, sizeof(set), &set);
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
    bind_near_device(device);
    if (!g_stream) CK(cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking));
    g_ev_ready = 1;
    /* the interpreter keeps its field list and backtrack stack in local memory */
    CK(cudaFuncSetCacheConfig(k_chain_eval, cudaFuncCachePreferL1));
    CK(cudaFuncSetCacheConfig(k_chain_emit_list, cudaFuncCachePreferL1));
    {   /* and ask for the largest L1 the unified array can give (FLBGPU_MAX_L1=0: driver default) */
        const char *e = getenv("FLBGPU_MAX_L1");
        if (!(e && e[0] == '0')) {
            cudaFuncSetAttribute(k_chain_eval, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxL1);
            cudaFuncSetAttribute(k_chain_emit_list, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxL1);
            cudaGetLastError();
        }
    }
    return 0;
}

/* 64 bytes of slack: djf_scan_plain reads whole aligned 8-byte words */
void *bk_alloc(size_t n) { void *p = 0; if (cudaMalloc(&p, n + 64) != cudaSuccess) { snprintf(g_err, sizeof(g_err), "cudaMalloc(%zu) failed", n); return 0; } return p; }
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
        for (int i = 0; i < g_ev_used[k]; i++) {
            float ms = 0.f;
            if (cudaEventSynchronize(g_evp[k][i][1]) == cudaSuccess &&
                cudaEventElapsedTime(&ms, g_evp[k][i][0], g_evp[k][i][1]) == cudaSuccess) out[k] += ms;
        }
    }
    return 0;
}


static cudaStream_t g_h2d, g_copy;
static int g_streams_ready;
static int streams_init(void)
{
    if (g_streams_ready) return 0;
    {   /* the index kernels of the next slice are short and the host waits for their result: let their
         * blocks go ahead of the thousands of queued evaluation blocks of the previous slice */
        int lo = 0, hi = 0;
        CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CK(cudaStreamCreateWithPriority(&g_istream, cudaStreamNonBlocking, hi));
    }
    CK(cudaStreamCreateWithFlags(&g_h2d, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&g_copy, cudaStreamNonBlocking));
    g_streams_ready = 1;
    return 0;
}

/* ---- upload: pieces with events ---- */
#define UP_PIECE ((size_t) 32 << 20)
#define UP_MAX_EV 256
static cudaEvent_t up_ev[UP_MAX_EV];
static int up_ev_made;
static size_t up_total, up_piece;
static int up_active;

/* Pageable input (a chunk that lives in ordinary malloc()ed memory, as Fluent Bit's do): the driver
 * would stage every cudaMemcpyAsync itself, synchronously and single-threaded.  Instead UP_THREADS host
 * threads copy the pieces into pinned staging buffers in parallel and enqueue the H2D copies in piece
 * order; up_recorded[i] tells the indexing side when piece i's event exists. */
#define UP_THREADS 6
#define UP_STAGE_SLOTS 12
static uint8_t *up_stage[UP_STAGE_SLOTS];
static cudaEvent_t up_stage_ev[UP_STAGE_SLOTS];
static int up_stage_ready;
static std::atomic<int> up_recorded[UP_MAX_EV];
static std::atomic<long> up_next_issue;
static std::atomic<int> up_failed;
static std::thread up_threads[UP_THREADS];
static int up_threads_live, up_staged;

static void up_worker(int t, uint8_t *d_dst, const uint8_t *h_src, size_t n, size_t piece, size_t np)
{
    for (size_t i = (size_t) t; i < np; i += UP_THREADS) {
        const int slot = (int) (i % UP_STAGE_SLOTS);
        const size_t off = i * piece, sz = (off + piece <= n) ? piece : n - off;
        /* the slot was last used by piece i - UP_STAGE_SLOTS: its H2D must have left the buffer */
        if (i >= UP_STAGE_SLOTS) {
            while (!up_recorded[i - UP_STAGE_SLOTS].load(std::memory_order_acquire)) { if (up_failed.load()) return; sched_yield(); }
            if (cudaEventSynchronize(up_ev[i - UP_STAGE_SLOTS]) != cudaSuccess) { up_failed.store(1); return; }
        }
        memcpy(up_stage[slot], h_src + off, sz);
        while (up_next_issue.load(std::memory_order_acquire) != (long) i) { if (up_failed.load()) return; sched_yield(); }
        if (cudaMemcpyAsync(d_dst + off, up_stage[slot], sz, cudaMemcpyHostToDevice, g_h2d) != cudaSuccess ||
            cudaEventRecord(up_ev[i], g_h2d) != cudaSuccess) { up_failed.store(1); up_next_issue.store((long) i + 1); return; }
        up_recorded[i].store(1, std::memory_order_release);
        up_next_issue.store((long) i + 1, std::memory_order_release);
    }
}

void bk_upload_end(void)
{
    if (up_threads_live) {
        for (int t = 0; t < UP_THREADS; t++) up_threads[t].join();
        up_threads_live = 0;
    }
}

int bk_upload_start(void *d_dst, const void *h_src, size_t n)
{
    if (streams_init()) return -1;
    bk_upload_end();
    up_piece = UP_PIECE;
    while ((n + up_piece - 1) / up_piece > UP_MAX_EV) up_piece *= 2;
    const size_t np = (n + up_piece - 1) / up_piece;
    for (; up_ev_made < (int) np; up_ev_made++) CK(cudaEventCreateWithFlags(&up_ev[up_ev_made], cudaEventDisableTiming));
    up_total = n; up_active = 1; up_staged = 0;
    {
        cudaPointerAttributes at;
        const int pinned = cudaPointerGetAttributes(&at, h_src) == cudaSuccess && (at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged);
        cudaGetLastError();
        if (!pinned && up_piece == UP_PIECE && !getenv("FLBGPU_NO_STAGING")) {
            if (!up_stage_ready) {
                for (int i = 0; i < UP_STAGE_SLOTS; i++) CK(cudaMallocHost((void **) &up_stage[i], UP_PIECE));
                up_stage_ready = 1;
            }
            for (size_t i = 0; i < np; i++) up_recorded[i].store(0);
            up_next_issue.store(0); up_failed.store(0);
            up_staged = 1;
            for (int t = 0; t < UP_THREADS; t++)
                up_threads[t] = std::thread(up_worker, t, (uint8_t *) d_dst, (const uint8_t *) h_src, n, up_piece, np);
            up_threads_live = 1;
            return 0;
        }
    }
    for (size_t i = 0; i < np; i++) {
        const size_t off = i * up_piece, sz = (off + up_piece <= n) ? up_piece : n - off;
        CK(cudaMemcpyAsync((uint8_t *) d_dst + off, (const uint8_t *) h_src + off, sz, cudaMemcpyHostToDevice, g_h2d));
        CK(cudaEventRecord(up_ev[i], g_h2d));
    }
    return 0;
}
void bk_upload_none(void) { bk_upload_end(); up_active = 0; }
int bk_upload_wait_index(size_t upto)
{
    if (streams_init()) return -1;
    if (!up_active || upto == 0) return 0;
    if (upto > up_total) upto = up_total;
    const size_t last = (upto - 1) / up_piece;
    if (up_staged) {
        while (!up_recorded[last].load(std::memory_order_acquire)) {
            if (up_failed.load()) { snprintf(g_err, sizeof(g_err), "host->device staging failed"); return -1; }
            sched_yield();
        }
    }
    CK(cudaStreamWaitEvent(g_istream, up_ev[last], 0));
    return 0;
}

/* ---- download session: pinned ring + one host thread per slot ---- */
#define XF_MAX_SLOTS 32
/* ring geometry: FLBGPU_XF_SLOTS staging buffers (one copy-out thread each) of FLBGPU_XF_MB MiB */
static int XF_SLOTS = 16;           /* measured best on the bench box: 16 x 8 MiB (profiles/r01_variants.txt) */
static size_t XF_SLICE = (size_t) 8 << 20;
#define XF_MAX_RANGES 512
static cudaEvent_t xf_rev[XF_MAX_RANGES];
static int xf_rev_made;
static uint8_t *xf_ring[XF_MAX_SLOTS];
static cudaEvent_t xf_ev[XF_MAX_SLOTS], xf_evc;
static int xf_ready;
struct xf_session {
    uint8_t *h_dst; const uint8_t *d_src;
    std::atomic<long> issued[XF_MAX_SLOTS], done[XF_MAX_SLOTS];
    size_t s_off[XF_MAX_SLOTS], s_len[XF_MAX_SLOTS];
    std::atomic<long> n_issued;          /* slices issued so far */
    std::atomic<int> closed, failed;
    std::thread th[XF_MAX_SLOTS];
    int started;
    /* byte ranges handed over by bk_download_push(); an issuer thread turns them into ring slices so
     * that the caller (which also drives indexing and evaluation of the next slice) never blocks on a
     * full ring */
    size_t r_lo[XF_MAX_RANGES], r_hi[XF_MAX_RANGES];
    std::atomic<long> n_pushed;
    std::atomic<int> push_closed;
    std::thread issuer;
};
static xf_session *g_xf;

static int xf_init(void)
{
    if (xf_ready) return 0;
    if (streams_init()) return -1;
    {
        const char *es = getenv("FLBGPU_XF_SLOTS"), *em = getenv("FLBGPU_XF_MB");
        if (es && atoi(es) >= 2 && atoi(es) <= XF_MAX_SLOTS) XF_SLOTS = atoi(es);
        if (em && atoi(em) >= 1 && atoi(em) <= 256) XF_SLICE = (size_t) atoi(em) << 20;
    }
    for (int i = 0; i < XF_SLOTS; i++) {
        CK(cudaMallocHost((void **) &xf_ring[i], XF_SLICE));
        CK(cudaEventCreateWithFlags(&xf_ev[i], cudaEventDisableTiming));
    }
    CK(cudaEventCreateWithFlags(&xf_evc, cudaEventDisableTiming));
    xf_ready = 1;
    return 0;
}

static void xf_worker(xf_session *x, int s)
{
    for (long i = s;; i += XF_SLOTS) {
        while (x->issued[s].load(std::memory_order_acquire) < i) {
            if (x->failed.load()) return;
            if (x->closed.load() && x->n_issued.load() <= i) return;
            sched_yield();
        }
        if (cudaEventSynchronize(xf_ev[s]) != cudaSuccess) { x->failed.store(1); return; }
        flbgpu_stream_copy(x->h_dst + x->s_off[s], xf_ring[s], x->s_len[s]);
        x->done[s].store(i, std::memory_order_release);
    }
}

static void xf_issuer(xf_session *x)
{
    for (long r = 0;; r++) {
        while (x->n_pushed.load(std::memory_order_acquire) <= r) {
            if (x->failed.load() || x->push_closed.load()) {
                if (x->n_pushed.load(std::memory_order_acquire) > r) break;
                x->closed.store(1);
                return;
            }
            sched_yield();
        }
        /* bytes [lo,hi) exist once the emission recorded in xf_rev[r] is done */
        if (cudaStreamWaitEvent(g_copy, xf_rev[r], 0) != cudaSuccess) { x->failed.store(1); x->closed.store(1); return; }
        for (size_t off = x->r_lo[r]; off < x->r_hi[r] && !x->failed.load(); off += XF_SLICE) {
            const long i = x->n_issued.load();
            const int s = (int) (i % XF_SLOTS);
            const size_t sz = (off + XF_SLICE <= x->r_hi[r]) ? XF_SLICE : x->r_hi[r] - off;
            if (i >= XF_SLOTS) while (x->done[s].load(std::memory_order_acquire) < i - XF_SLOTS) { if (x->failed.load()) break; sched_yield(); }
            x->s_off[s] = off; x->s_len[s] = sz;
            if (cudaMemcpyAsync(xf_ring[s], x->d_src + off, sz, cudaMemcpyDeviceToHost, g_copy) != cudaSuccess ||
                cudaEventRecord(xf_ev[s], g_copy) != cudaSuccess) { x->failed.store(1); break; }
            x->issued[s].store(i, std::memory_order_release);
            x->n_issued.store(i + 1);
        }
    }
}

int bk_download_begin(void *h_dst, const void *d_out)
{
    if (xf_init()) return -1;
    xf_session *x = new xf_session();
    x->h_dst = (uint8_t *) h_dst; x->d_src = (const uint8_t *) d_out;
    for (int s = 0; s < XF_SLOTS; s++) { x->issued[s].store(-1); x->done[s].store(-1); }
    x->n_issued.store(0); x->closed.store(0); x->failed.store(0);
    x->n_pushed.store(0); x->push_closed.store(0);
    for (int s = 0; s < XF_SLOTS; s++) x->th[s] = std::thread(xf_worker, x, s);
    x->issuer = std::thread(xf_issuer, x);
    x->started = 1;
    g_xf = x;
    return 0;
}

int bk_download_push(size_t lo, size_t hi)
{
    xf_session *x = g_xf;
    if (!x) return -1;
    if (hi <= lo) return 0;
    const long r = x->n_pushed.load();
    if (r >= XF_MAX_RANGES) { snprintf(g_err, sizeof(g_err), "too many download ranges"); return -1; }
    for (; xf_rev_made <= (int) r; xf_rev_made++) CK(cudaEventCreateWithFlags(&xf_rev[xf_rev_made], cudaEventDisableTiming));
    CK(cudaEventRecord(xf_rev[r], g_stream));
    x->r_lo[r] = lo; x->r_hi[r] = hi;
    x->n_pushed.store(r + 1, std::memory_order_release);
    return x->failed.load() ? -1 : 0;
}

int bk_download_end(void)
{
    xf_session *x = g_xf;
    int rc = 0;
    if (!x) return -1;
    x->push_closed.store(1);
    x->issuer.join();
    for (int s = 0; s < XF_SLOTS; s++) x->th[s].join();
    if (x->failed.load()) { snprintf(g_err, sizeof(g_err), "device->host transfer failed: %s", cudaGetErrorString(cudaGetLastError())); rc = -1; }
    delete x;
    g_xf = 0;
    return rc;
}

int bk_d2d(void *dst, const void *src, size_t n)
{
    CK(cudaMemcpy(dst, src, n, cudaMemcpyDeviceToDevice));
    return 0;
}

static unsigned long long *g_dtotal;   /* device scratch for totals / repair results */
static uint32_t *g_dbreaks;
static int ensure_small(void)
{
    if (streams_init()) return -1;
    if (!g_dtotal) CK(cudaMalloc((void **) &g_dtotal, 128));
    if (!g_dbreaks) CK(cudaMalloc((void **) &g_dbreaks, sizeof(uint32_t) * BK_MAX_BREAKS));
    return 0;
}

int bk_index_count(const uint8_t *d_in, size_t slice_off, uint32_t slice_len, uint32_t *d_tile, uint32_t n_tiles,
                   uint32_t *n_cand)
{
    unsigned long long tot = 0;
    if (ensure_small()) return -1;
    *n_cand = 0;
    if (n_tiles == 0) return 0;
    ev_begin_on(0, g_istream);
    {
        const uint32_t skip = (uint32_t) (slice_off & 15);
        k_index<false><<<n_tiles, 256, 0, g_istream>>>(d_in + slice_off - skip, slice_len + skip, skip, (uint32_t) (slice_off - skip), d_tile, 0, 0, 0);
    }
    k_scan_top<uint32_t><<<1, 256, 0, g_istream>>>(d_tile, n_tiles, g_dtotal);
    ev_end_on(0, g_istream);
    g_launches += 2;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(&tot, g_dtotal, sizeof(tot), cudaMemcpyDeviceToHost, g_istream));
    CK(cudaStreamSynchronize(g_istream));
    *n_cand = (uint32_t) tot;
    return 0;
}

int bk_index_fill(const uint8_t *d_in, size_t slice_off, uint32_t slice_len, const uint32_t *d_tile, uint32_t n_tiles,
                  uint32_t n_cand, uint32_t *d_off, uint32_t *d_len, uint8_t *d_kind,
                  uint32_t *n_valid, uint64_t *end_off, int *tiled)
{
    uint32_t h[4] = { 0, 0, 0, 0 };
    uint32_t *d_w = (uint32_t *) (g_dtotal + 1);          /* [0] n_breaks, [1] n_valid, [2] tiled, [3] end offset */
    *n_valid = 0; *tiled = (slice_len == 0); *end_off = slice_off;
    if (n_cand == 0) return 0;
    ev_begin_on(0, g_istream);
    {
        const uint32_t skip = (uint32_t) (slice_off & 15);
        k_index<true><<<n_tiles, 256, 0, g_istream>>>(d_in + slice_off - skip, slice_len + skip, skip, (uint32_t) (slice_off - skip), (uint32_t *) d_tile, d_off, d_len, d_kind);
    }
    CK(cudaMemsetAsync(d_w, 0, 16, g_istream));
    k_index_check<<<(n_cand + 255) / 256, 256, 0, g_istream>>>(d_off, d_len, n_cand, (uint32_t) (slice_off + slice_len), d_w, g_dbreaks);
    k_index_repair<<<1, 1024, 0, g_istream>>>(d_off, d_len, d_kind, n_cand, (uint32_t) slice_off, (uint32_t) (slice_off + slice_len), d_w, g_dbreaks, d_w + 1);
    ev_end_on(0, g_istream);
    g_launches += 3;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(h, d_w, sizeof(h), cudaMemcpyDeviceToHost, g_istream));
    CK(cudaStreamSynchronize(g_istream));
    if (h[0] > BK_MAX_BREAKS) {
        snprintf(g_err, sizeof(g_err), "record index: %u broken candidate links in one slice (limit %u)", h[0], BK_MAX_BREAKS);
        return -1;
    }
    *n_valid = h[1];
    *tiled = (int) h[2];
    *end_off = h[3];
    return 0;
}

static void fill_params(const struct bk_chain_args *a, k_chain_params *p, uint8_t *d_out, uint32_t r0)
{
    p->env.in = a->d_in; p->env.in_len = a->in_len; p->env.blob = a->d_blob; p->env.scr = a->d_scr;
    p->env.capcache = a->d_capcache; p->env.cap_stride = a->cap_stride; p->env.now = a->now;
    p->env.assume = a->assume; p->env.fl_flags = a->d_flags; p->env.err = a->d_flags + FLBGPU_MAX_FILTERS;
    p->env.l2m = a->l2m;
    p->off = a->d_off; p->len = a->d_len; p->kind = a->d_kind; p->r0 = r0; p->n_rec = a->n_rec; p->stage_bytes = 0;
    p->size = a->d_size; p->bsum = a->d_bsum; p->out = d_out;
}

int bk_flags_clear(uint32_t *d_flags)
{
    if (streams_init()) return -1;
    CK(cudaMemsetAsync(d_flags, 0, sizeof(uint32_t) * (FLBGPU_MAX_FILTERS + 1), g_stream));
    g_ev_used[0] = g_ev_used[1] = g_ev_used[2] = 0;
    g_nsurv = 0;
    return 0;
}

/* The slice of input the next evaluation reads is touched once: mark it as streaming in L2 (access
 * policy window of the stream) so that it does not evict the lanes' local-memory lines (field lists,
 * regex stacks), which are re-used by every block.  Measured without effect (profiles/r01_variants.txt), so it
 * is opt-in: FLBGPU_L2_WINDOW=1. */
int bk_hint_streaming(const void *base, size_t bytes)
{
    static int max_win = -1, enabled = -1;
    cudaStreamAttrValue v;
    if (streams_init()) return -1;
    if (enabled < 0) { const char *e = getenv("FLBGPU_L2_WINDOW"); enabled = (e && e[0] == '1'); }   /* off: no measurable effect on B200 */
    if (!enabled) return 0;
    if (max_win < 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&max_win, cudaDevAttrMaxAccessPolicyWindowSize, dev) != cudaSuccess) max_win = 0;
    }
    if (max_win <= 0 || !base || !bytes) return 0;
    memset(&v, 0, sizeof(v));
    v.accessPolicyWindow.base_ptr = (void *) base;
    v.accessPolicyWindow.num_bytes = bytes < (size_t) max_win ? bytes : (size_t) max_win;
    v.accessPolicyWindow.hitRatio = 1.0f;
    v.accessPolicyWindow.hitProp = cudaAccessPropertyStreaming;
    v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    if (cudaStreamSetAttribute(g_stream, cudaStreamAttributeAccessPolicyWindow, &v) != cudaSuccess) cudaGetLastError();
    return 0;
}

int bk_chain_eval(const struct bk_chain_args *a, uint32_t r0, uint32_t r1)
{
    k_chain_params p;
    if (r1 <= r0) return 0;
    fill_params(a, &p, 0, r0);
    p.n_rec = r1;
    {
        static int stage_kb = -1;
        if (stage_kb < 0) {
            const char *e = getenv("FLBGPU_STAGE_KB");           /* KiB of shared memory per warp, 0 = off */
            stage_kb = e ? atoi(e) : 0;
            if (stage_kb < 0 || stage_kb > 24) stage_kb = 0;
            if (stage_kb) CK(cudaFuncSetAttribute(k_chain_eval, cudaFuncAttributeMaxDynamicSharedMemorySize, stage_kb * 1024 * (BK_REC_BLOCK / 32)));
        }
        p.stage_bytes = (uint32_t) stage_kb * 1024;
    }
    ev_begin(1);
    k_chain_eval<<<(r1 - r0 + BK_REC_BLOCK - 1) / BK_REC_BLOCK, BK_REC_BLOCK, (size_t) p.stage_bytes * (BK_REC_BLOCK / 32), g_stream>>>(p);
    if (p.env.l2m.hash) {
        k_chain_skipped<<<(r1 - r0 + BK_REC_BLOCK - 1) / BK_REC_BLOCK, BK_REC_BLOCK, 0, g_stream>>>(p);
        g_launches += 1;
    }
    ev_end(1);
    g_launches += 1;
    CK(cudaGetLastError());
    return 0;
}

int bk_flags_fetch(const uint32_t *d_flags, uint32_t *h_flags)
{
    CK(cudaMemcpyAsync(h_flags, d_flags, sizeof(uint32_t) * (FLBGPU_MAX_FILTERS + 1), cudaMemcpyDeviceToHost, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    return 0;
}

int bk_sizes_scan(const uint32_t *d_size, uint32_t n_rec, uint64_t *d_bsum, uint64_t *h_bsum)
{
    const uint32_t nb = (n_rec + BK_REC_BLOCK - 1) / BK_REC_BLOCK;
    unsigned long long tot = 0;
    if (ensure_small()) return -1;
    h_bsum[0] = 0;
    if (nb) {
        k_bsum<<<nb, BK_REC_BLOCK, 0, g_stream>>>(d_size, n_rec, d_bsum);
        k_scan_top<uint64_t><<<1, 256, 0, g_stream>>>(d_bsum, nb, g_dtotal);
        g_launches += 2;
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(&tot, g_dtotal, sizeof(tot), cudaMemcpyDeviceToHost, g_stream));
        CK(cudaMemcpyAsync(h_bsum, d_bsum, sizeof(uint64_t) * nb, cudaMemcpyDeviceToHost, g_stream));
    }
    CK(cudaStreamSynchronize(g_stream));
    h_bsum[nb] = tot;
    return 0;
}

/* Offsets of blocks [b0, b1) when everything before them is already placed: block sums, exclusive
 * scan continued from carry_in, host copy of the new entries; h_bsum[b1] = bytes placed so far. */
__global__ void __launch_bounds__(256) k_scan_carry(uint64_t *a, uint32_t n, unsigned long long *carry_io)
{
    unsigned long long carry = *carry_io;
    for (uint32_t b = 0; b < n; b += 256) {
        const uint32_t i = b + threadIdx.x;
        unsigned long long v = i < n ? (unsigned long long) a[i] : 0, tot;
        unsigned long long ex = block_excl_scan_t<unsigned long long>(v, &tot);
        if (i < n) a[i] = carry + ex;
        carry += tot;
    }
    __syncthreads();
    if (threadIdx.x == 0) *carry_io = carry;
}

int bk_sizes_scan_range(const uint32_t *d_size, uint32_t n_rec, uint32_t b0, uint32_t b1, uint64_t *d_bsum, uint64_t *h_bsum,
                        uint64_t carry_in)
{
    unsigned long long tot = carry_in;
    if (ensure_small()) return -1;
    h_bsum[b0] = carry_in;
    if (b1 > b0) {
        const uint32_t nb = b1 - b0;
        CK(cudaMemcpyAsync(g_dtotal, &tot, sizeof(tot), cudaMemcpyHostToDevice, g_stream));
        k_bsum<<<nb, BK_REC_BLOCK, 0, g_stream>>>(d_size + (size_t) b0 * BK_REC_BLOCK, n_rec - b0 * BK_REC_BLOCK, d_bsum + b0);
        k_scan_carry<<<1, 256, 0, g_stream>>>(d_bsum + b0, nb, g_dtotal);
        g_launches += 2;
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(&tot, g_dtotal, sizeof(tot), cudaMemcpyDeviceToHost, g_stream));
        CK(cudaMemcpyAsync(h_bsum + b0, d_bsum + b0, sizeof(uint64_t) * nb, cudaMemcpyDeviceToHost, g_stream));
    }
    CK(cudaStreamSynchronize(g_stream));
    h_bsum[b1] = tot;
    return 0;
}

#define BK_MAX_EMITS 1024
static unsigned long long *g_hsurv;

/* records emitted since the last bk_flags_clear(); synchronises the library stream */
int bk_records_out(uint64_t *n)
{
    unsigned long long t = 0;
    CK(cudaStreamSynchronize(g_stream));
    for (int i = 0; i < g_nsurv; i++) t += g_hsurv[i];
    *n = t;
    return 0;
}

int bk_chain_emit(const struct bk_chain_args *a, uint8_t *d_out, uint32_t b0, uint32_t b1)
{
    k_chain_params p;
    if (b1 <= b0) return 0;
    fill_params(a, &p, d_out, b0 * BK_REC_BLOCK);
    {
        const uint32_t nb = b1 - b0, rec0 = b0 * BK_REC_BLOCK;
        const uint32_t n = (a->n_rec > rec0) ? ((a->n_rec - rec0 < nb * BK_REC_BLOCK) ? a->n_rec - rec0 : nb * BK_REC_BLOCK) : 0;
        static uint32_t *d_cnt, *d_lrec; static uint64_t *d_loff; static unsigned long long *d_nlist; static size_t cap_b, cap_r;
        if (ensure_small()) return -1;
        if (cap_b < nb) { cudaFree(d_cnt); d_cnt = 0; cap_b = 0; CK(cudaMalloc((void **) &d_cnt, sizeof(uint32_t) * (nb + nb / 2 + 64))); cap_b = nb + nb / 2 + 64; }
        if (cap_r < (size_t) nb * BK_REC_BLOCK) {
            const size_t want = (size_t) (nb + nb / 2 + 64) * BK_REC_BLOCK;
            CK(cudaStreamSynchronize(g_stream));           /* an emission still running reads the old lists */
            cudaFree(d_lrec); cudaFree(d_loff); d_lrec = 0; d_loff = 0; cap_r = 0;
            CK(cudaMalloc((void **) &d_lrec, sizeof(uint32_t) * want));
            CK(cudaMalloc((void **) &d_loff, sizeof(uint64_t) * want));
            cap_r = want;
        }
        if (!d_nlist) CK(cudaMalloc((void **) &d_nlist, 64));
        ev_begin(2);
        k_surv_count<<<nb, BK_REC_BLOCK, 0, g_stream>>>(a->d_size + rec0, n, d_cnt);
        k_scan_top<uint32_t><<<1, 256, 0, g_stream>>>(d_cnt, nb, d_nlist);
        k_surv_fill<<<nb, BK_REC_BLOCK, 0, g_stream>>>(a->d_size + rec0, n, rec0, d_cnt, a->d_bsum + b0, d_lrec, d_loff);
        k_chain_emit_list<<<nb, BK_REC_BLOCK, 0, g_stream>>>(p, d_lrec, d_loff, d_nlist);
        ev_end(2);
        g_launches += 4;
        /* how many records this range emitted: read back after the call's last synchronisation */
        if (!g_hsurv) CK(cudaMallocHost((void **) &g_hsurv, sizeof(unsigned long long) * BK_MAX_EMITS));
        if (g_nsurv < BK_MAX_EMITS) CK(cudaMemcpyAsync(&g_hsurv[g_nsurv++], d_nlist, sizeof(unsigned long long), cudaMemcpyDeviceToHost, g_stream));
    }
    CK(cudaGetLastError());
    return 0;
}

}
