/* runtime_pool.h -- several devices inside one process (part of runtime.c, included there).
 *
 * The unit the path shards by without anything to agree on is the chunk: the filters decide MODIFIED / NOTOUCH per chunk
 * (record_modifier re-encodes every record of a chunk once one record of it lost a key, src/flb_filter.c:119-323 hands every
 * filter a whole chunk), so two halves of one chunk on two devices would have to exchange their evidence first, two chunks on
 * two devices exchange nothing.  A pool is n chains of the same configuration -- one per device, each with its own queue -- and a
 * worker thread per chain; flbgpu_pool_do() hands a batch of chunks out to whichever chain is free and returns the results in the
 * order of the chunks.  Filters that carry state from chunk to chunk (multiline: the rule a group is in) see the chunks of their
 * own chain only; log_to_metrics tables stay per chain (flbgpu_l2m_allreduce merges them). */
#include <stdatomic.h>

struct flbgpu_pool {
    int n;
    flbgpu_chain **chains;
    pthread_t *th;
    pthread_mutex_t m;
    pthread_cond_t go, done;
    long gen;
    int quit, running;
    /* the batch being worked on */
    int n_chunks;
    const void *const *data; const size_t *bytes;
    const char *tag; int tag_len;
    void **out_bufs; size_t *out_sizes; int *rets;
    atomic_int next;
};
struct pool_arg { struct flbgpu_pool *p; int k; };

static void *pool_worker(void *arg_)
{
    struct pool_arg *arg = arg_;
    struct flbgpu_pool *p = arg->p;
    const int k = arg->k;
    long seen = 0;
    free(arg);
    for (;;) {
        pthread_mutex_lock(&p->m);
        while (!p->quit && p->gen == seen) pthread_cond_wait(&p->go, &p->m);
        if (p->quit) { pthread_mutex_unlock(&p->m); return NULL; }
        seen = p->gen;
        pthread_mutex_unlock(&p->m);
        for (;;) {
            const int i = atomic_fetch_add(&p->next, 1);
            if (i >= p->n_chunks) break;
            p->rets[i] = flbgpu_chain_do(p->chains[k], p->data[i], p->bytes[i], p->tag, p->tag_len, &p->out_bufs[i], &p->out_sizes[i]);
        }
        pthread_mutex_lock(&p->m);
        if (--p->running == 0) pthread_cond_signal(&p->done);
        pthread_mutex_unlock(&p->m);
    }
}

flbgpu_pool *flbgpu_pool_new(flbgpu_chain *const *chains, int n)
{
    struct flbgpu_pool *p;
    int k;
    g_rt_err[0] = 0;
    if (!chains || n < 1) return NULL;
    for (k = 0; k < n; k++) if (!chains[k] || !chains[k]->inited) { set_err("flbgpu_pool_new: chain not initialised%s%s", NULL, NULL); return NULL; }
    p = calloc(1, sizeof(*p));
    if (!p) return NULL;
    p->n = n;
    p->chains = malloc(sizeof(*p->chains) * (size_t) n);
    p->th = calloc((size_t) n, sizeof(*p->th));
    if (!p->chains || !p->th) { free(p->chains); free(p->th); free(p); return NULL; }
    memcpy(p->chains, chains, sizeof(*p->chains) * (size_t) n);
    pthread_mutex_init(&p->m, NULL);
    pthread_cond_init(&p->go, NULL);
    pthread_cond_init(&p->done, NULL);
    for (k = 0; k < n; k++) {
        struct pool_arg *a = malloc(sizeof(*a));
        if (!a || (a->p = p, a->k = k, pthread_create(&p->th[k], NULL, pool_worker, a)) != 0) {
            free(a);
            p->n = k;
            flbgpu_pool_destroy(p);
            set_err("flbgpu_pool_new: cannot start a worker thread%s%s", NULL, NULL);
            return NULL;
        }
    }
    return p;
}

/* every chunk through one of the pool's chains; rets[i] is what flbgpu_chain_do() returned for chunk i (its result in
 * out_bufs[i] / out_sizes[i]).  Returns 0, or -1 when some call failed (the others are still valid). */
int flbgpu_pool_do(flbgpu_pool *p, int n_chunks, const void *const *data, const size_t *bytes, const char *tag, int tag_len,
                   void **out_bufs, size_t *out_sizes, int *rets)
{
    int i, bad = 0;
    if (!p || n_chunks < 0 || (n_chunks && (!data || !bytes || !out_bufs || !out_sizes || !rets))) return -1;
    if (n_chunks == 0) return 0;
    pthread_mutex_lock(&p->m);
    p->n_chunks = n_chunks; p->data = data; p->bytes = bytes; p->tag = tag; p->tag_len = tag_len;
    p->out_bufs = out_bufs; p->out_sizes = out_sizes; p->rets = rets;
    atomic_store(&p->next, 0);
    p->running = p->n;
    p->gen++;
    pthread_cond_broadcast(&p->go);
    while (p->running) pthread_cond_wait(&p->done, &p->m);
    pthread_mutex_unlock(&p->m);
    for (i = 0; i < n_chunks; i++) if (rets[i] < 0) bad = 1;
    return bad ? -1 : 0;
}

void flbgpu_pool_destroy(flbgpu_pool *p)
{
    int k;
    if (!p) return;
    pthread_mutex_lock(&p->m);
    p->quit = 1;
    pthread_cond_broadcast(&p->go);
    pthread_mutex_unlock(&p->m);
    for (k = 0; k < p->n; k++) if (p->th[k]) pthread_join(p->th[k], NULL);
    pthread_mutex_destroy(&p->m); pthread_cond_destroy(&p->go); pthread_cond_destroy(&p->done);
    free(p->chains); free(p->th); free(p);
}
