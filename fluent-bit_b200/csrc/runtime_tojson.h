/* runtime_tojson.h -- host side of the chunk -> JSON text conversion (part of runtime.c, included there).
 * flb_pack_msgpack_to_json_format(), src/flb_pack.c:1320-1602. */

#define TJ_MARKS_CAP 65536u      /* group markers per chunk */

static flbgpu_chain *tj_chain(flbgpu_ctx *ctx)
{
    flbgpu_chain *c = ctx->util;
    if (c) return c;
    c = flbgpu_chain_new(ctx);                         /* a queue and the growable buffers of a chain, no filters */
    if (!c) return NULL;
    c->d_flags = bk_alloc(c->q, sizeof(uint32_t) * (FLBGPU_MAX_FILTERS + 1));
    if (!c->d_flags) { flbgpu_chain_destroy(c); return NULL; }
    c->l2m_index = c->rtag_index = c->ml_index = -1;
    c->inited = 1;
    ctx->util = c;
    return c;
}

static int tj_run(flbgpu_chain *c, const uint8_t *h_in, size_t bytes, int json_format, int date_format, const char *date_key,
                  int escape_unicode, char **out, size_t *out_size, size_t *undefined_strings)
{
    struct tj_env e;
    const uint8_t *d_in;
    uint32_t n_rec = 0, nb, h_flags[FLBGPU_MAX_FILTERS + 1];
    size_t off = 0, S = slice_bytes(), at, need;
    unsigned long long undef = 0, cnt[2] = { 0, 0 };
    uint32_t *d_groups = NULL;
    uint64_t total;
    char *res;

    memset(&c->st, 0, sizeof(c->st));
    c->st.bytes_in = bytes;
    if (bytes >= 0xfff00000ull) { set_err("chunk larger than 4 GiB: split the append%s%s", NULL, NULL); return -1; }
    memset(&e, 0, sizeof(e));
    e.key_len = 0xffffffffu;
    if (date_key) {
        size_t kl = strlen(date_key);
        if (kl > sizeof(e.key)) { set_err("date key longer than %s bytes%s", "128", NULL); return -1; }
        memcpy(e.key, date_key, kl);
        e.key_len = (uint32_t) kl;
    }
    e.json_format = (uint32_t) json_format; e.date_format = (uint32_t) date_format; e.escape_unicode = escape_unicode ? 1u : 0u;
    if (bytes) {
        GROW(c->d_in, c->cap_in, bytes + 64, uint8_t);
        if (bk_upload_start(c->q, c->d_in, h_in, bytes)) return -1;
    }
    d_in = c->d_in;
    if (bk_flags_clear(c->q, c->d_flags)) return -1;
    while (off < bytes) {                              /* record index, slice by slice */
        size_t len = bytes - off < S ? bytes - off : S;
        uint32_t n_tiles = (uint32_t) ((len + ((uintptr_t) (d_in + off) & 15) + BK_INDEX_TILE - 1) / BK_INDEX_TILE), n_cand = 0, n_valid = 0;
        uint64_t end_off = off;
        int tiled = 0;
        if (bk_upload_wait_index(c->q, off + len)) return -1;
        GROW(c->d_tile, c->cap_tile, n_tiles + 1, uint32_t);
        if (bk_index_count(c->q, d_in, off, (uint32_t) len, c->d_tile, n_tiles, &n_cand)) return -1;
        if (ensure_rec_cap(c, (size_t) n_rec + n_cand, n_rec)) return -1;
        if (bk_index_fill(c->q, d_in, off, (uint32_t) len, c->d_tile, n_tiles, n_cand, c->d_off + n_rec, c->d_len + n_rec,
                          c->d_kind + n_rec, &n_valid, &end_off, &tiled)) {
            c->st.error_bits = FLBGPU_E_INDEX;
            return -1;
        }
        if (n_valid == 0) {
            if (off + len < bytes) { S *= 2; continue; }
            break;
        }
        n_rec += n_valid;
        off = (size_t) end_off;
    }
    REFUSE_WIDE_ARRAYS(h_in, d_in, return -1);
    c->st.records_in = n_rec;
    nb = (n_rec + BK_REC_BLOCK - 1) / BK_REC_BLOCK;
    total = 0;
    if (n_rec) {
        e.in = d_in; e.off = c->d_off; e.len = c->d_len; e.kind = c->d_kind; e.n_rec = n_rec;
        e.err = c->d_flags + FLBGPU_MAX_FILTERS;
        e.scr_pad = TJ_SCR_PAD(e.key_len == 0xffffffffu ? 0u : e.key_len);
        e.size = c->d_size;
        /* work memory: the counters (undefined strings, group markers), the marker lists, the packed lengths, the scratch slices */
        e.marks_cap = TJ_MARKS_CAP;
        need = 16 + 3 * sizeof(uint32_t) * (size_t) TJ_MARKS_CAP + sizeof(uint32_t) * (size_t) n_rec + 64;
        GROW(c->d_mlw, c->cap_mlw, need, uint8_t);
        at = 0;
        e.undefined = (unsigned long long *) (c->d_mlw + at); e.n_marks = e.undefined + 1; at += 16;
        e.marks = (uint32_t *) (c->d_mlw + at); at += 2 * sizeof(uint32_t) * (size_t) TJ_MARKS_CAP;
        d_groups = (uint32_t *) (c->d_mlw + at); at += sizeof(uint32_t) * (size_t) TJ_MARKS_CAP;
        e.plen = (uint32_t *) (c->d_mlw + at);
        GROW(c->d_scr, c->cap_scr, bytes + (size_t) n_rec * e.scr_pad + 64, uint8_t);
        e.scr = c->d_scr;
        if (bk_zero(c->q, e.undefined, 16)) return -1;
        GROW(c->d_bsum, c->cap_bsum, nb + 2, uint64_t);
        if (c->cap_hbsum < (size_t) nb + 2) {
            free(c->h_bsum);
            c->cap_hbsum = (size_t) nb + nb / 4 + 64;
            c->h_bsum = malloc(c->cap_hbsum * sizeof(uint64_t));
            if (!c->h_bsum) { c->cap_hbsum = 0; return -1; }
        }
        if (bk_tj_sizes(c->q, &e) || bk_sizes_scan(c->q, c->d_size, n_rec, c->d_bsum, c->h_bsum)) return -1;
        if (bk_flags_fetch(c->q, c->d_flags, h_flags)) return -1;
        if (h_flags[FLBGPU_MAX_FILTERS] & FLBGPU_E_JSONDATE) return 1;          /* the reference returns NULL for such a chunk */
        if (refused(c, h_flags[FLBGPU_MAX_FILTERS])) return -1;
        total = c->h_bsum[nb];
        if (total >= 0xfff00000ull) { set_err("result larger than 4 GiB%s%s", NULL, NULL); return -1; }
        if (bk_d2h(c->q, cnt, e.undefined, sizeof(cnt)) || bk_sync(c->q)) return -1;
        if (cnt[1]) {
            /* group markers in the chunk: the events behind a start carry its body as group_attributes -- the list in record
             * order, and the sizing pass again with it */
            uint32_t *h = malloc(2 * sizeof(uint32_t) * (size_t) cnt[1]), maxlen = 0;
            size_t a, b2;
            if (!h) return -1;
            if (bk_d2h(c->q, h, e.marks, 2 * sizeof(uint32_t) * (size_t) cnt[1]) || bk_sync(c->q)) { free(h); return -1; }
            for (a = 0; a < (size_t) cnt[1]; a++) { if (h[2 * a + 1] > maxlen) maxlen = h[2 * a + 1]; h[a] = h[2 * a]; }
            for (a = 1; a < (size_t) cnt[1]; a++) {      /* a handful: insertion sort */
                const uint32_t v = h[a];
                for (b2 = a; b2 > 0 && h[b2 - 1] > v; b2--) h[b2] = h[b2 - 1];
                h[b2] = v;
            }
            if (bk_h2d(c->q, d_groups, h, sizeof(uint32_t) * (size_t) cnt[1]) || bk_sync(c->q)) { free(h); return -1; }
            free(h);
            e.groups = d_groups; e.n_groups = (uint32_t) cnt[1];
            e.scr_pad += maxlen + 32;                   /* an event's packed map now holds a marker's body too */
            GROW(c->d_scr, c->cap_scr, bytes + (size_t) n_rec * e.scr_pad + 64, uint8_t);
            e.scr = c->d_scr;
            if (bk_zero(c->q, e.undefined, 8)) return -1;
            if (bk_tj_sizes(c->q, &e) || bk_sizes_scan(c->q, c->d_size, n_rec, c->d_bsum, c->h_bsum)) return -1;
            if (bk_flags_fetch(c->q, c->d_flags, h_flags)) return -1;
            if (h_flags[FLBGPU_MAX_FILTERS] & FLBGPU_E_JSONDATE) return 1;
            if (refused(c, h_flags[FLBGPU_MAX_FILTERS])) return -1;
            total = c->h_bsum[nb];
            if (total >= 0xfff00000ull) { set_err("result larger than 4 GiB%s%s", NULL, NULL); return -1; }
            if (bk_d2h(c->q, cnt, e.undefined, sizeof(cnt)) || bk_sync(c->q)) return -1;
        }
        undef = cnt[0];
    }
    bk_upload_end(c->q);
    c->st.kernel_launches = bk_launch_count();
    if (undefined_strings) *undefined_strings = (size_t) undef;
    /* json: "[" records joined by "," "]" -- also for no record at all; lines / stream: nothing converted -> NULL */
    if (json_format == (int) TJ_FORMAT_JSON) {
        res = malloc((size_t) total + 3);
        if (!res) return -1;
        if (total) {
            GROW(c->d_out, c->cap_out, total, uint8_t);
            if (bk_tj_emit(c->q, &e, c->d_bsum, c->d_out) || bk_d2h(c->q, res, c->d_out, (size_t) total) || bk_sync(c->q)) { free(res); return -1; }
            res[0] = '[';
            res[total] = ']'; res[total + 1] = 0;
            *out_size = (size_t) total + 1;
        }
        else { res[0] = '['; res[1] = ']'; res[2] = 0; *out_size = 2; }
        *out = res;
        return 0;
    }
    if (total == 0) return 1;
    res = malloc((size_t) total + 1);
    if (!res) return -1;
    GROW(c->d_out, c->cap_out, total, uint8_t);
    if (bk_tj_emit(c->q, &e, c->d_bsum, c->d_out) || bk_d2h(c->q, res, c->d_out, (size_t) total) || bk_sync(c->q)) { free(res); return -1; }
    res[total] = 0;
    *out = res; *out_size = (size_t) total;
    c->st.bytes_out = total;
    return 0;
}

/* 0: *out is the malloc()ed, NUL-terminated text; 1: the reference returns NULL for this input (nothing to convert);
 * -1: the call failed (flbgpu_last_error) */
int flbgpu_msgpack_to_json_format(flbgpu_ctx *ctx, const void *data, size_t bytes, int json_format, int date_format,
                                  const char *date_key, int escape_unicode, char **out, size_t *out_size, size_t *undefined_strings)
{
    flbgpu_chain *c;
    int r;
    g_rt_err[0] = 0;
    if (!ctx || !out || !out_size || (!data && bytes)) return -1;
    *out = NULL; *out_size = 0;
    if (undefined_strings) *undefined_strings = 0;
    if (json_format < 1 || json_format > 3) { set_err("json_format must be 1 (json), 2 (stream) or 3 (lines)%s%s", NULL, NULL); return -1; }
    if (date_format < 0 || date_format > 4) { set_err("unknown date format%s%s", NULL, NULL); return -1; }
    c = tj_chain(ctx);
    if (!c) return -1;

    pthread_mutex_lock(&c->lock);
    ctx->last_q = c->q;
    r = tj_run(c, data, bytes, json_format, date_format, date_key, escape_unicode, out, out_size, undefined_strings);
    if (r < 0) bk_upload_end(c->q);
    pthread_mutex_unlock(&c->lock);
    return r;
}
