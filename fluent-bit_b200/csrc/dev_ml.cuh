/* dev_ml.cuh -- filter_multiline (`buffer off`, parser mode) on the device.
 *
 * Reference: plugins/filter_multiline/ml.c:792-892 (cb_ml_filter: every event of the chunk through flb_ml_append_event(),
 * flb_ml_flush_pending_now() at the end, the flushed messages are the new chunk); src/multiline/flb_ml.c:207-345
 * (package_content), :398-496 (process_append), :758-860 (flb_ml_append_object: a line no parser takes flushes what is
 * pending and goes out on its own), :1585-1793 (flb_ml_flush_stream_group); src/multiline/flb_ml_rule.c:250-436
 * (try_flushing_buffer, try_start_state, flb_ml_rule_process); src/multiline/flb_ml_group.c:83-122 (flb_ml_group_cat).
 *
 * What the reference does line after line splits in two:
 *   per record, independent of everything else -- find key_content (first STR key with a STR value of that name,
 *     flb_ml.c:360-396), run every rule's regex over the content: ml_feat_record(), one lane per record;
 *   sequential -- which rule the group is in (rule_to_state), whether its buffer holds bytes, whether it holds a first-line
 *     context: a finite automaton (ml_step) with at most (rules + 1) * 4 states.  Its transition function over a block of
 *     ML_F1 records is a table of that many bytes (ml_up1); tables compose (ml_up2), a walk over the few top-level tables
 *     gives every block its incoming state (ml_top, ml_down2), and a second walk over the records with the state known
 *     writes what each record does (ml_apply): flush before, append, become the context, flush after.
 * Every flush that has something to say is one event of the result.  Lane j of the sizing / emission pass owns flush j: it
 * walks the records since the previous flush, adds up the buffer the reference would have built (separators as
 * flb_ml_rule_process adds them) and writes `[[time, {}], map]` with key_content replaced by the buffer.
 *
 * All functions take the index of the "thread" that does the piece of work, so that the kernels are one-line wrappers and the
 * CPU emulation of the tests (tests/hostsim) runs the very same code in loops.
 */
#ifndef FLBGPU_DEV_ML_CUH
#define FLBGPU_DEV_ML_CUH
#include "dev_chain.cuh"


FLB_HD const struct cf_ml *ml_cfg(const struct ml_env *e) { return (const struct cf_ml *) (e->blob + e->cfg_off); }
FLB_HD const struct cf_ml_rule *ml_rules(const struct ml_env *e) { return (const struct cf_ml_rule *) (e->blob + ml_cfg(e)->rules_off); }

/* timestamp, metadata and body map of the framed event at off (rec_frame() accepted it) */
struct ml_rec { int64_t sec, nsec; const uint8_t *meta, *meta_end, *body, *body_end; uint32_t n_kv; const uint8_t *kv; };
FLB_HD void ml_rec_open(const struct ml_env *e, uint32_t off, uint32_t len, struct ml_rec *r)
{
    const uint8_t *p = e->in + off, *end = p + len, *q = p + 1;
    struct mp_tok t;
    const int v2 = (*q == 0x92);
    if (v2) q++;
    mp_token(q, end, &t);
    r->nsec = 0;
    if (t.type == MPT_UINT || t.type == MPT_INT) r->sec = (int64_t) t.u;
    else if (t.type == MPT_F64) {
        union { uint64_t u; double d; } cv;
        cv.u = t.u;
        r->sec = (int64_t) cv.d;
        r->nsec = (int64_t) ((cv.d - (double) r->sec) * 1000000000.0);
    }
    else {
        r->sec = (int64_t) (int32_t) mp_be32(q + t.hdr);
        r->nsec = (int64_t) (int32_t) mp_be32(q + t.hdr + 4);
    }
    q += t.hdr + (t.type == MPT_EXT ? t.len : 0);
    r->meta = r->meta_end = 0;
    if (v2) { r->meta = q; q = mp_skip(q, end); r->meta_end = q; }
    r->body = q; r->body_end = end;
    mp_token(q, end, &t);
    r->n_kv = t.len;
    r->kv = q + t.hdr;
}

/* ---- the parallel pass: one record */
FLB_HDN void ml_feat_record(const struct ml_env *e, uint32_t i)
{
    const struct cf_ml *m = ml_cfg(e);
    struct ml_feat f;
    f.coff = f.clen = f.bits = 0;
    if (e->kind[i] == 0) {
        struct ml_rec r;
        uint32_t k;
        const uint8_t *q, *end;
        f.bits = MLF_LIVE;
        ml_rec_open(e, e->off[i], e->len[i], &r);
        q = r.kv; end = r.body_end;
        if (m->key_len != 0xffffffffu) {
            const uint8_t *key = e->blob + m->key_off;
            for (k = 0; k < r.n_kv; k++) {             /* get_key_id(): the first STR key of that name whose value is a STR */
                struct mp_tok tk, tv;
                const uint8_t *kp = q, *vp;
                mp_token(q, end, &tk);
                vp = mp_skip(q, end);
                mp_token(vp, end, &tv);
                q = mp_skip(vp, end);
                if (tk.type == MPT_STR && tv.type == MPT_STR && tk.len == m->key_len && bytes_eq(kp + tk.hdr, key, tk.len)) {
                    f.bits |= MLF_HAS;
                    f.coff = (uint32_t) (vp + tv.hdr - e->in);
                    f.clen = tv.len;
                    break;
                }
            }
        }
        if (f.bits & MLF_HAS) {
            const uint8_t *s = e->in + f.coff;
            if (m->type == ML_T_REGEX) {
                const struct cf_ml_rule *R = ml_rules(e);
                struct ch_env ce;                        /* rx_run() reports through the chain environment's words */
                struct ch_lane ln;
                int caps[2 * (RX_MAX_GROUPS + 1)];
                uint32_t stk[CH_RX_STACK], j;
                ce.blob = e->blob; ce.err = e->err;
                ln.scr = 0;
                for (j = 0; j < m->n_rules; j++)
                    if (rx_run(&ce, &ln, R[j].rx_off, s, f.clen, caps, stk)) f.bits |= 1u << j;
            }
            else {
                const uint8_t *ms = e->blob + m->match_off;
                int hit = 0;
                if (m->type == ML_T_ENDSWITH) {
                    if (m->match_len <= f.clen) { f.bits |= MLF_LENOK; hit = bytes_eq(s + f.clen - m->match_len, ms, m->match_len); }
                }
                else { f.bits |= MLF_LENOK; hit = f.clen == m->match_len && bytes_eq(s, ms, f.clen); }
                if (m->negate) hit = !hit;              /* match_negate(), flb_ml.c:36-56 */
                if (hit) f.bits |= 1u;
            }
        }
    }
    e->feat[i] = f;
}

/* ---- the automaton: state = rule code (0 none, 1 + rule) * 4 + (buffer holds bytes) * 2 + (first-line context held).
 * Returns the next state in bits 0..7 and the MLA_* actions in bits 8..15. */
FLB_HD uint32_t ml_step(const struct cf_ml *m, const struct cf_ml_rule *R, uint32_t s, uint32_t bits, uint32_t clen)
{
    uint32_t rc = s >> 2, b = (s >> 1) & 1u, c = s & 1u, act = 0;
    int processed = 0;
    if (!(bits & MLF_LIVE)) return s;
    if (bits & MLF_HAS) {
        if (m->type == ML_T_REGEX) {                    /* flb_ml_rule_process() */
            int rule = -1;
            uint32_t k;
            if (rc) {
                const struct cf_ml_rule *cur = &R[rc - 1];
                for (k = 0; k < cur->n_to; k++) {
                    if ((bits >> cur->to[k]) & 1u) {    /* a continuation */
                        act |= MLA_APP | MLA_SEP;
                        if (!clen) act |= MLA_NL;
                        b = 1;
                        rule = (int) cur->to[k];
                        break;
                    }
                }
            }
            if (rule < 0) {
                for (k = 0; k < m->n_rules; k++) {      /* try_start_state() */
                    if (R[k].start && ((bits >> k) & 1u)) {
                        if (b) { act |= MLA_FB; b = 0; c = 0; }
                        rule = (int) k;
                        if (clen) { act |= MLA_APP; b = 1; }
                        act |= MLA_CTXTIME;
                        if (!c) act |= MLA_CTXMAP;      /* flb_ml_register_context() appends to mp_sbuf: the first map is the one read back */
                        c = 1;
                        break;
                    }
                }
            }
            if (rule >= 0) {
                rc = (uint32_t) rule + 1;
                if (R[rule].next_start && b) { act |= MLA_FA; b = 0; c = 0; }    /* try_flushing_buffer() */
                processed = 1;
            }
        }
        else if (bits & MLF_LENOK) {                    /* package_content(): FLB_ML_ENDSWITH / FLB_ML_EQ */
            if (!c) { act |= MLA_CTXMAP | MLA_CTXTIME; c = 1; }
            if (clen) { act |= MLA_APP; b = 1; }
            if (bits & 1u) { act |= MLA_FA; b = 0; c = 0; }
            processed = 1;
        }
    }
    if (!processed) {
        /* flb_ml_append_object(): nothing took the line -- what is pending goes out, then the record on its own */
        if (b | c) act |= MLA_FB;
        act |= MLA_FA | MLA_CTXMAP | MLA_CTXTIME | MLA_ALONE;
        b = 0; c = 0;
    }
    return (rc << 2) | (b << 1) | c | (act << 8);
}

/* block t of ML_F1 records as a function on states: one thread per (block, state) -- S times the threads, each a chain of
 * ML_F1 steps instead of S * ML_F1 (the launch list showed the one-thread-per-block form at 2.7 ms per million records with
 * 123 blocks of threads on the whole GPU) */
FLB_HDN void ml_up1(const struct ml_env *e, uint32_t ts)
{
    const struct cf_ml *m = ml_cfg(e);
    const struct cf_ml_rule *R = ml_rules(e);
    const uint32_t t = ts / e->S, s0 = ts % e->S;
    const uint32_t lo = t * ML_F1, hi = lo + ML_F1 < e->n_rec ? lo + ML_F1 : e->n_rec;
    uint32_t s = s0, i;
    for (i = lo; i < hi; i++) s = ml_step(m, R, s, e->feat[i].bits, e->feat[i].clen) & 0xffu;
    e->T1[(size_t) t * e->S + s0] = (uint8_t) s;
}
FLB_HDN void ml_up2(const struct ml_env *e, uint32_t us)
{
    const uint32_t u = us / e->S, s0 = us % e->S;
    const uint32_t lo = u * ML_F2, hi = lo + ML_F2 < e->nt1 ? lo + ML_F2 : e->nt1;
    uint32_t s = s0, t;
    for (t = lo; t < hi; t++) s = e->T1[(size_t) t * e->S + s];
    e->T2[(size_t) u * e->S + s0] = (uint8_t) s;
}
FLB_HDN void ml_top(const struct ml_env *e)
{
    uint32_t s = e->state_in, u;
    for (u = 0; u < e->nt2; u++) { e->in2[u] = (uint8_t) s; s = e->T2[(size_t) u * e->S + s]; }
    e->res[1] = s;
}
FLB_HDN void ml_down2(const struct ml_env *e, uint32_t u)
{
    const uint32_t lo = u * ML_F2, hi = lo + ML_F2 < e->nt1 ? lo + ML_F2 : e->nt1;
    uint32_t s = e->in2[u], t;
    for (t = lo; t < hi; t++) { e->in1[t] = (uint8_t) s; s = e->T1[(size_t) t * e->S + s]; }
}
/* block t with its incoming state known: what every record does; events and last time registration of the block */
FLB_HDN void ml_apply(const struct ml_env *e, uint32_t t)
{
    const struct cf_ml *m = ml_cfg(e);
    const struct cf_ml_rule *R = ml_rules(e);
    const uint32_t lo = t * ML_F1, hi = lo + ML_F1 < e->n_rec ? lo + ML_F1 : e->n_rec;
    uint32_t s = e->in1[t], i, cnt = 0, lt = 0;
    for (i = lo; i < hi; i++) {
        const uint32_t r = ml_step(m, R, s, e->feat[i].bits, e->feat[i].clen), act = (r >> 8) & 0xffu;
        s = r & 0xffu;
        e->act[i] = (uint8_t) act;
        cnt += ((act & MLA_FB) ? 1u : 0u) + ((act & MLA_FA) ? 1u : 0u);
        if (act & MLA_CTXTIME) lt = i + 1;
    }
    e->cnt1[t] = cnt; e->lt1[t] = lt;
}
FLB_HDN void ml_cnt_up2(const struct ml_env *e, uint32_t u)
{
    const uint32_t lo = u * ML_F2, hi = lo + ML_F2 < e->nt1 ? lo + ML_F2 : e->nt1;
    uint32_t t, cnt = 0, lt = 0;
    for (t = lo; t < hi; t++) { cnt += e->cnt1[t]; if (e->lt1[t]) lt = e->lt1[t]; }
    e->cnt2[u] = cnt; e->lt2[u] = lt;
}
/* the time of the record with index lt - 1 (lt > 0) */
FLB_HD void ml_time_of(const struct ml_env *e, uint32_t lt, int64_t *sec, int64_t *nsec)
{
    struct ml_rec r;
    ml_rec_open(e, e->off[lt - 1], e->len[lt - 1], &r);
    *sec = r.sec; *nsec = r.nsec;
}
FLB_HDN void ml_cnt_top(const struct ml_env *e)
{
    uint32_t u, base = 0, tl = 0;
    const uint32_t sfin = (uint32_t) e->res[1];
    for (u = 0; u < e->nt2; u++) { e->base2[u] = base; e->tl2[u] = tl; base += e->cnt2[u]; if (e->lt2[u]) tl = e->lt2[u]; }
    if (sfin & 3u) e->ev_slot[base++] = 2u * e->n_rec;          /* flb_ml_flush_pending_now() finds something */
    e->res[0] = base;
    e->res[1] = sfin & ~3u;                                      /* flushed: the rule is what stays */
    if (tl) { int64_t s, ns; ml_time_of(e, tl, &s, &ns); e->res[2] = (unsigned long long) s; e->res[3] = (unsigned long long) ns; }
    else { e->res[2] = (unsigned long long) e->time_in[0]; e->res[3] = (unsigned long long) e->time_in[1]; }
}
FLB_HDN void ml_cnt_down2(const struct ml_env *e, uint32_t u)
{
    const uint32_t lo = u * ML_F2, hi = lo + ML_F2 < e->nt1 ? lo + ML_F2 : e->nt1;
    uint32_t t, base = e->base2[u], tl = e->tl2[u];
    for (t = lo; t < hi; t++) { e->base1[t] = base; e->tl1[t] = tl; base += e->cnt1[t]; if (e->lt1[t]) tl = e->lt1[t]; }
}
FLB_HDN void ml_fill(const struct ml_env *e, uint32_t t)
{
    const uint32_t lo = t * ML_F1, hi = lo + ML_F1 < e->n_rec ? lo + ML_F1 : e->n_rec;
    uint32_t i, at = e->base1[t], tl = e->tl1[t];
    for (i = lo; i < hi; i++) {
        const uint32_t act = e->act[i];
        if (act & MLA_FB) e->ev_slot[at++] = 2u * i;
        if (act & MLA_CTXTIME) tl = i + 1;
        e->tl[i] = tl;
        if (act & MLA_FA) e->ev_slot[at++] = 2u * i + 1u;
    }
}

/* ---- one flush = one event of the result.
 * Slot 2i is "before record i is looked at", 2i + 1 "after record i", 2n the end of the chunk.  The records whose pieces the
 * flush at slot e carries are those behind the previous flush: [first, last]. */
FLB_HD void ml_event_range(const struct ml_env *e, uint32_t j, uint32_t *first, uint32_t *last_plus1, uint32_t *tl)
{
    const uint32_t slot = e->ev_slot[j];
    uint32_t lo = 0, hi = (slot + 1u) >> 1;            /* slot 2i: records < i; slot 2i + 1: records <= i */
    if (j) { const uint32_t p = e->ev_slot[j - 1]; lo = (p + 1u) >> 1; }
    *first = lo; *last_plus1 = hi;
    *tl = hi ? e->tl[hi - 1] : 0;
}

/* ---- the metadata of a message: the members of its lines' metadata maps in the order the lines came, a member dropped when an
 * earlier one of the message has the same hash (flb_ml_flush_metadata_buffer(), flb_ml.c:1501-1583).  The hash
 * (flb_hash_msgpack_object_list, :1013-1135) is taken over key and value FLATTENED: scalars in document order, nil as eight
 * zero bytes, integers and reals as their eight bytes, strings / bins / ext bodies as their bytes, containers as nothing but
 * their members -- so {"a": "bc"} and {"ab": "c"} are the same member to it.  Bins have become lower-case hex strings before
 * (flb_ml_msgpack_object_deep_copy_convert(), :1286-1443).  Two independent 64-bit hashes over the same flattened bytes stand
 * in for cfl_hash_64bits here: equal bytes give equal hashes in both, different bytes collide in neither (2^-128). */
FLB_HD void ml_md_hash_bytes(uint64_t h[2], const uint8_t *p, uint32_t n)
{
    uint32_t i;
    for (i = 0; i < n; i++) {
        h[0] = (h[0] ^ p[i]) * 0x100000001b3ull;
        h[1] = (h[1] + p[i] + 1u) * 0x9e3779b97f4a7c15ull; h[1] ^= h[1] >> 29;
    }
}
FLB_HD void ml_md_hash(const uint8_t *p, const uint8_t *end, uint64_t h[2])      /* [p, end): a key and its value */
{
    struct mp_tok t;
    h[0] = 0xcbf29ce484222325ull; h[1] = 0x243f6a8885a308d3ull;
    while (p < end && mp_token(p, end, &t) == 0) {
        uint8_t w[8];
        uint64_t v = 0;
        uint32_t i;
        switch (t.type) {
        case MPT_NIL: for (i = 0; i < 8; i++) w[i] = 0; ml_md_hash_bytes(h, w, 8); break;
        case MPT_BOOL: w[0] = (uint8_t) t.u; ml_md_hash_bytes(h, w, 1); break;
        case MPT_UINT: case MPT_INT: case MPT_F64: case MPT_F32:
            v = t.u;
            if (t.type == MPT_F32) { union { uint32_t u; float f; } a; union { uint64_t u; double d; } c; a.u = (uint32_t) t.u; c.d = (double) a.f; v = c.u; }
            for (i = 0; i < 8; i++) w[i] = (uint8_t) (v >> (8 * i));
            ml_md_hash_bytes(h, w, 8);
            break;
        case MPT_STR: ml_md_hash_bytes(h, p + t.hdr, t.len); break;
        case MPT_BIN:
            for (i = 0; i < t.len; i++) { w[0] = "0123456789abcdef"[p[t.hdr + i] >> 4]; w[1] = "0123456789abcdef"[p[t.hdr + i] & 15]; ml_md_hash_bytes(h, w, 2); }
            break;
        case MPT_EXT: w[0] = (uint8_t) t.ext_type; ml_md_hash_bytes(h, w, 1); ml_md_hash_bytes(h, p + t.hdr, t.len); break;
        default: break;                                  /* array / map headers */
        }
        p += t.hdr;
        if (t.type == MPT_STR || t.type == MPT_BIN || t.type == MPT_EXT) p += t.len;
    }
}
/* msgpack_pack_object() of the converted copy: mp_canon() with bins as hex strings */
FLB_HD uint32_t ml_md_pack(const uint8_t *p, const uint8_t *end, uint8_t *o)
{
    struct mp_tok t;
    uint32_t n = 0, i;
    while (p < end && mp_token(p, end, &t) == 0) {
        if (t.type == MPT_BIN) {
            if (o) { mp_put_str_hdr(o + n, 2 * t.len); }
            n += mp_str_hdr_size(2 * t.len);
            for (i = 0; i < t.len; i++) {
                if (o) { o[n] = "0123456789abcdef"[p[t.hdr + i] >> 4]; o[n + 1] = "0123456789abcdef"[p[t.hdr + i] & 15]; }
                n += 2;
            }
            p += t.hdr + t.len;
        }
        else if (t.type == MPT_ARRAY || t.type == MPT_MAP) {
            if (o) { if (t.type == MPT_MAP) mp_put_map_hdr(o + n, t.len); else mp_put_array_hdr(o + n, t.len); }
            n += mp_cnt_hdr_size(t.len);
            p += t.hdr;
        }
        else {
            const uint8_t *nx = p + t.hdr + ((t.type == MPT_STR || t.type == MPT_EXT) ? t.len : 0);
            n += mp_canon(p, nx, o ? o + n : 0, 0);
            p = nx;
        }
    }
    return n;
}
/* the metadata map of the message flush j carries: `df` + count + the surviving members; returns its size */
FLB_HDN uint32_t ml_event_metadata(const struct ml_env *e, uint32_t j, uint32_t lo, uint32_t hi, uint8_t *o)
{
    const uint32_t slot = e->ev_slot[j];
    uint64_t hs[ML_MD_MAX][2];
    uint32_t n_seen = 0, kept = 0, n = 5, i, first, last1;
    const int alone = (slot & 1u) && slot < 2u * e->n_rec && (e->act[slot >> 1] & MLA_ALONE);
    if (alone) { first = slot >> 1; last1 = first + 1; }            /* everything older was purged when the line found no taker */
    else {
        /* the lines whose metadata was added since the previous flush: a line's goes in AFTER the rule ran on it, so the line a
         * flush-after belongs to hands its metadata to the NEXT message */
        const uint32_t p = j ? e->ev_slot[j - 1] : 0u;
        first = j ? p >> 1 : 0u;
        last1 = slot >> 1;                                           /* lines i with 2i + 1.5 < slot */
        if (last1 > e->n_rec) last1 = e->n_rec;
        (void) lo; (void) hi;
    }
    for (i = first; i < last1; i++) {
        struct ml_rec r;
        struct mp_tok t;
        const uint8_t *q;
        uint32_t k;
        if (!(e->feat[i].bits & MLF_LIVE)) continue;
        if (!alone && (e->act[i] & MLA_ALONE)) continue;             /* went out with its own */
        ml_rec_open(e, e->off[i], e->len[i], &r);
        if (!r.meta) continue;                                       /* a legacy event: the decoder's empty map */
        mp_token(r.meta, r.meta_end, &t);
        q = r.meta + t.hdr;
        for (k = 0; k < t.len; k++) {
            const uint8_t *vp = mp_skip(q, r.meta_end), *nx = mp_skip(vp, r.meta_end);
            uint64_t h[2];
            uint32_t x;
            int dup = 0;
            ml_md_hash(q, nx, h);
            for (x = 0; x < n_seen && !dup; x++) if (hs[x][0] == h[0] && hs[x][1] == h[1]) dup = 1;
            if (!dup) {
                if (n_seen >= ML_MD_MAX) { CH_ATOMIC_OR(e->err, FLBGPU_E_MLMETA); return 5; }
                hs[n_seen][0] = h[0]; hs[n_seen][1] = h[1]; n_seen++;
                n += ml_md_pack(q, nx, o ? o + n : 0);
                kept++;
            }
            q = nx;
        }
    }
    if (o) { o[0] = 0xdf; mp_put_be32(o + 1, kept); }
    return n;
}

/* size of the event (out == NULL) or its bytes.  The sizing pass leaves buffer length and context record for the emission. */
FLB_HDN uint32_t ml_event(const struct ml_env *e, uint32_t j, uint8_t *out)
{
    const struct cf_ml *m = ml_cfg(e);
    uint32_t lo, hi, tl, i, buflen = 0, ctx = 0, n = 0;
    int last = -1;                                       /* last byte of the buffer so far */
    int64_t sec, nsec;
    ml_event_range(e, j, &lo, &hi, &tl);
    if (!out) {
        for (i = lo; i < hi; i++) {
            const uint32_t act = e->act[i];
            if ((act & MLA_CTXMAP) && !ctx) ctx = i + 1;
            if (act & MLA_APP) {
                const struct ml_feat f = e->feat[i];
                if ((act & MLA_SEP) && buflen >= 1 && last != '\n') { buflen++; last = '\n'; }
                if (act & MLA_NL) { buflen++; last = '\n'; }
                else {
                    if (m->type == ML_T_REGEX && m->limit && (buflen >= m->limit || f.clen > m->limit - buflen))
                        CH_ATOMIC_OR(e->err, FLBGPU_E_MLLIMIT);            /* flb_ml_group_cat() would cut here */
                    buflen += f.clen;
                    last = e->in[f.coff + f.clen - 1];
                }
            }
        }
        e->ev_buflen[j] = buflen; e->ev_ctx[j] = ctx;
    }
    else { buflen = e->ev_buflen[j]; ctx = e->ev_ctx[j]; }
    if (buflen >= 0x0fffffffu) CH_ATOMIC_OR(e->err, FLBGPU_E_MLLIMIT);

    /* [[time, {}], body]: the encoder's forced map32 for the (empty) metadata, flb_log_event_encoder.c:195-218 */
    if (tl) ml_time_of(e, tl, &sec, &nsec);
    else { sec = e->time_in[0]; nsec = e->time_in[1]; }
    if (sec == 0 && nsec == 0) { sec = e->now[0]; nsec = e->now[1]; }       /* flb_time_get() in the reference */
    if (out) {
        out[0] = 0x92; out[1] = 0x92; out[2] = 0xd7; out[3] = 0x00;
        mp_put_be32(out + 4, (uint32_t) sec); mp_put_be32(out + 8, (uint32_t) nsec);
    }
    n = 12;
    n += ml_event_metadata(e, j, lo, hi, out ? out + n : 0);
    if (ctx) {
        struct ml_rec r;
        ml_rec_open(e, e->off[ctx - 1], e->len[ctx - 1], &r);
        if (buflen == 0) n += mp_canon(r.body, r.body_end, out ? out + n : 0, 0);   /* the original map from the context */
        else {
            /* the first line's keys, key_content's value replaced by the buffer.  `len` is the reference's variable of that name
             * (flb_ml.c:1655-1687): the key's length until the first replacement, the buffer's afterwards */
            const uint8_t *q = r.kv, *end = r.body_end, *key = e->blob + m->key_off;
            uint32_t k, cmp_len = m->key_len;
            if (out) mp_put_map_hdr(out + n, r.n_kv);
            n += mp_cnt_hdr_size(r.n_kv);
            for (k = 0; k < r.n_kv; k++) {
                struct mp_tok tk;
                const uint8_t *kp = q, *vp = mp_skip(q, end), *nx = mp_skip(vp, end);
                int is_content = 0;
                mp_token(kp, end, &tk);
                if (tk.type == MPT_STR && m->key_len != 0xffffffffu && tk.len == cmp_len) {
                    /* strncmp(k.ptr, key_content, len) == 0 over len bytes, stopping at a NUL of either */
                    uint32_t x;
                    is_content = 1;
                    for (x = 0; x < cmp_len; x++) {
                        const uint8_t a = kp[tk.hdr + x], bb = x < m->key_len ? key[x] : 0;
                        if (a != bb) { is_content = 0; break; }
                        if (a == 0) break;
                    }
                }
                n += mp_canon(kp, vp, out ? out + n : 0, 0);
                if (is_content) {
                    if (out) mp_put_str_hdr(out + n, buflen);
                    n += mp_str_hdr_size(buflen);
                    if (out) {
                        uint32_t w = 0;
                        last = -1;
                        for (i = lo; i < hi; i++) {
                            const uint32_t act = e->act[i];
                            if (act & MLA_APP) {
                                const struct ml_feat f = e->feat[i];
                                if ((act & MLA_SEP) && w >= 1 && last != '\n') { out[n + w++] = '\n'; last = '\n'; }
                                if (act & MLA_NL) { out[n + w++] = '\n'; last = '\n'; }
                                else { mp_copy(out + n + w, e->in + f.coff, f.clen); w += f.clen; last = e->in[f.coff + f.clen - 1]; }
                            }
                        }
                    }
                    n += buflen;
                    cmp_len = buflen;
                }
                else n += mp_canon(vp, nx, out ? out + n : 0, 0);
                q = nx;
            }
        }
    }
    else if (buflen) {
        /* no first line: the raw content under key_content ("log" when there is none) */
        const uint8_t *key = m->key_len != 0xffffffffu ? e->blob + m->key_off : (const uint8_t *) "log";
        const uint32_t klen = m->key_len != 0xffffffffu ? m->key_len : 3u;
        if (out) { out[n] = 0x81; mp_put_str_hdr(out + n + 1, klen); mp_copy(out + n + 1 + mp_str_hdr_size(klen), key, klen); }
        n += 1 + mp_str_hdr_size(klen) + klen;
        if (out) mp_put_str_hdr(out + n, buflen);
        n += mp_str_hdr_size(buflen);
        if (out) {
            uint32_t w = 0;
            last = -1;
            for (i = lo; i < hi; i++) {
                const uint32_t act = e->act[i];
                if (act & MLA_APP) {
                    const struct ml_feat f = e->feat[i];
                    if ((act & MLA_SEP) && w >= 1 && last != '\n') { out[n + w++] = '\n'; last = '\n'; }
                    if (act & MLA_NL) { out[n + w++] = '\n'; last = '\n'; }
                    else { mp_copy(out + n + w, e->in + f.coff, f.clen); w += f.clen; last = e->in[f.coff + f.clen - 1]; }
                }
            }
        }
        n += buflen;
    }
    else return 0;                                        /* cannot happen: a listed flush has a context or bytes */
    return n;
}

#endif
