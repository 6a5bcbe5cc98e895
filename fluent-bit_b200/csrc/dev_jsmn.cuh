/* dev_jsmn.cuh -- the streaming JSON packer of inputs such as in_tcp / in_lib / in_stdin:
 * flb_pack_json_state() (src/flb_pack.c:758-829) = jsmn tokeniser (lib/jsmn/jsmn.h, built with JSMN_STRICT and
 * JSMN_PARENT_LINKS: include/fluent-bit/flb_info.h:50-55) + tokens_to_msgpack() (src/flb_pack.c:512-592).
 *
 * One lane = one stream buffer (a connection's pending bytes): the tokeniser is inherently sequential within a
 * buffer, the parallelism is across buffers (flbgpu_pack_json_state_batch).  The token array lives in the
 * buffer's scratch region; a buffer that needs more tokens than it holds reports JM_NOMEM and the host grows it.
 *
 * Everything jsmn's strict mode lets through is reproduced, including its oddities (a string without a colon
 * becomes another child of the object, `tru` is true, `-` is the integer 0): the bytes that come out are the
 * bytes the reference packs.
 */
#ifndef FLBGPU_DEV_JSMN_CUH
#define FLBGPU_DEV_JSMN_CUH

#include <stdint.h>
#include "flbgpu_prog.h"
#include "dev_msgpack.cuh"
#include "dev_json.cuh"

/* struct jm_tok, struct jm_result and the JM_* codes are shared with the host: flbgpu_prog.h */

FLB_HD struct jm_tok *jm_alloc(struct jm_tok *tok, uint32_t cap, uint32_t *toknext)
{
    struct jm_tok *t;
    if (*toknext >= cap) return 0;
    t = &tok[(*toknext)++];
    t->start = t->end = -1; t->size = 0; t->parent = -1; t->type = JM_UNDEFINED;
    return t;
}

/* jsmn_parse() from a fresh parser (pos 0): >= 0 ok, JM_INVAL, JM_PART or JM_NOMEM; *toknext_out = tokens allocated */
FLB_HD int jm_tokenise(const uint8_t *js, uint32_t len, struct jm_tok *tok, uint32_t cap, uint32_t *toknext_out)
{
    uint32_t pos = 0, toknext = 0;
    int32_t toksuper = -1;
    int i;
    *toknext_out = 0;
    for (; pos < len && js[pos] != 0; pos++) {
        const uint32_t c = js[pos];
        struct jm_tok *t;
        if (c == '{' || c == '[') {
            t = jm_alloc(tok, cap, &toknext);
            if (!t) { *toknext_out = toknext; return JM_NOMEM; }
            if (toksuper != -1) {
                struct jm_tok *s = &tok[toksuper];
                if (s->type == JM_OBJECT) { *toknext_out = toknext; return JM_INVAL; }      /* an object or array cannot be a key */
                s->size++;
                t->parent = toksuper;
            }
            t->type = c == '{' ? JM_OBJECT : JM_ARRAY;
            t->start = (int32_t) pos;
            toksuper = (int32_t) toknext - 1;
        }
        else if (c == '}' || c == ']') {
            const int32_t type = c == '}' ? JM_OBJECT : JM_ARRAY;
            if (toknext < 1) { *toknext_out = toknext; return JM_INVAL; }
            t = &tok[toknext - 1];
            for (;;) {
                if (t->start != -1 && t->end == -1) {
                    if (t->type != type) { *toknext_out = toknext; return JM_INVAL; }
                    t->end = (int32_t) pos + 1;
                    toksuper = t->parent;
                    break;
                }
                if (t->parent == -1) {
                    if (t->type != type || toksuper == -1) { *toknext_out = toknext; return JM_INVAL; }
                    break;
                }
                t = &tok[t->parent];
            }
        }
        else if (c == '"') {
            const uint32_t start = pos;
            int closed = 0;
            pos++;
            for (; pos < len && js[pos] != 0; pos++) {
                const uint32_t d = js[pos];
                if (d == '"') { closed = 1; break; }
                if (d == '\\' && pos + 1 < len) {
                    pos++;
                    switch (js[pos]) {
                    case '"': case '/': case '\\': case 'b': case 'f': case 'r': case 'n': case 't': break;
                    case 'u':
                        pos++;
                        for (i = 0; i < 4 && pos < len && js[pos] != 0; i++) {
                            if (dj_hex(js[pos]) < 0) { *toknext_out = toknext; return JM_INVAL; }
                            pos++;
                        }
                        pos--;
                        break;
                    default: *toknext_out = toknext; return JM_INVAL;
                    }
                }
            }
            if (!closed) { *toknext_out = toknext; return JM_PART; }
            t = jm_alloc(tok, cap, &toknext);
            if (!t) { *toknext_out = toknext; return JM_NOMEM; }
            t->type = JM_STRING; t->start = (int32_t) start + 1; t->end = (int32_t) pos; t->size = 0;
            t->parent = toksuper;
            if (toksuper != -1) tok[toksuper].size++;
        }
        else if (c == '\t' || c == '\r' || c == '\n' || c == ' ') { }
        else if (c == ':') toksuper = (int32_t) toknext - 1;
        else if (c == ',') {
            if (toksuper != -1 && tok[toksuper].type != JM_ARRAY && tok[toksuper].type != JM_OBJECT) toksuper = tok[toksuper].parent;
        }
        else if (c == '-' || (c >= '0' && c <= '9') || c == 't' || c == 'f' || c == 'n') {
            const uint32_t start = pos;
            int found = 0;
            if (toksuper != -1) {                           /* a primitive cannot be a key */
                const struct jm_tok *s = &tok[toksuper];
                if (s->type == JM_OBJECT || (s->type == JM_STRING && s->size != 0)) { *toknext_out = toknext; return JM_INVAL; }
            }
            for (; pos < len && js[pos] != 0; pos++) {
                const uint32_t d = js[pos];
                if (d == '\t' || d == '\r' || d == '\n' || d == ' ' || d == ',' || d == ']' || d == '}') { found = 1; break; }
                if (d < 32 || d >= 127) { *toknext_out = toknext; return JM_INVAL; }
            }
            if (!found) { *toknext_out = toknext; return JM_PART; }       /* strict mode: a primitive must be followed by , ] } or space */
            t = jm_alloc(tok, cap, &toknext);
            if (!t) { *toknext_out = toknext; return JM_NOMEM; }
            t->type = JM_PRIMITIVE; t->start = (int32_t) start; t->end = (int32_t) pos; t->size = 0;
            t->parent = toksuper;
            pos--;
            if (toksuper != -1) tok[toksuper].size++;
        }
        else { *toknext_out = toknext; return JM_INVAL; }   /* unexpected character in strict mode */
    }
    *toknext_out = toknext;
    for (i = (int) toknext - 1; i >= 0; i--) if (tok[i].start != -1 && tok[i].end == -1) return JM_PART;
    return JM_OK;
}

/* is_float(), src/flb_pack.c:163-234 */
FLB_HD int jm_is_float(const uint8_t *p, uint32_t n)
{
    uint32_t i;
    for (i = 0; i < n; i++) {
        if (p[i] == '.') return 1;
        if ((p[i] == 'e' || p[i] == 'E') && i + 1 < n) {
            const uint32_t d = p[i + 1];
            if (d == '-' || d == '+' || (d >= '0' && d <= '9')) return 1;
        }
    }
    return 0;
}

/* pack_numeric_token(), src/flb_pack.c:236-271: size (o == NULL) or bytes; *refused when the text needs a strtod()
 * form that is not restated */
FLB_HD uint32_t jm_number(const uint8_t *p, uint32_t n, uint8_t *o, int *refused)
{
    uint32_t i = 0;
    if (jm_is_float(p, n)) {
        int ok = 1;
        const uint64_t u = dj_strtod(p, (int) n, &ok);
        if (!ok) *refused = 1;
        if (o) { o[0] = 0xcb; mp_put_be64(o + 1, u); }
        return 9;
    }
    if (p[0] == '-') {                                      /* strtoll(p, NULL, 10): ERANGE -> strtod */
        uint64_t v = 0;
        int ovf = 0;
        i = 1;
        for (; i < n && p[i] >= '0' && p[i] <= '9'; i++) {
            const uint64_t d = p[i] - '0';
            if (v > (9223372036854775808ull - d) / 10) ovf = 1; else v = v * 10 + d;
        }
        if (ovf) {
            int ok = 1;
            const uint64_t u = dj_strtod(p, (int) n, &ok);
            if (!ok) *refused = 1;
            if (o) { o[0] = 0xcb; mp_put_be64(o + 1, u); }
            return 9;
        }
        {
            const int64_t sv = (int64_t) (0 - v);
            if (o) mp_put_int(o, sv);
            return mp_int_size(sv);
        }
    }
    {                                                       /* strtoull(p, NULL, 10) */
        uint64_t v = 0;
        int ovf = 0;
        for (; i < n && p[i] >= '0' && p[i] <= '9'; i++) {
            const uint64_t d = p[i] - '0';
            if (v > (0xffffffffffffffffull - d) / 10) ovf = 1; else v = v * 10 + d;
        }
        if (ovf) {
            int ok = 1;
            const uint64_t u = dj_strtod(p, (int) n, &ok);
            if (!ok) *refused = 1;
            if (o) { o[0] = 0xcb; mp_put_be64(o + 1, u); }
            return 9;
        }
        if (v <= 9223372036854775807ull) { if (o) mp_put_int(o, (int64_t) v); return mp_int_size((int64_t) v); }
        if (o) mp_put_uint(o, v);
        return mp_uint_size(v);
    }
}

/* pack_string_token(), src/flb_pack.c:274-313.  tmp: at least n + 1 bytes of scratch for the decoded text. */
FLB_HD uint32_t jm_string(const uint8_t *s, uint32_t n, uint8_t *o, uint8_t *tmp)
{
    uint32_t i, bad = 0, dl;
    for (i = 0; i < n; i++) if (s[i] == '"' || s[i] == '\\' || s[i] < 0x20) { bad = 1; break; }
    if (!bad) {
        if (o) { const uint32_t h = mp_put_str_hdr(o, n); mp_copy(o + h, s, n); }
        return mp_str_hdr_size(n) + n;
    }
    dl = lf_unescape_raw(s, n, tmp);
    if (o) { const uint32_t h = mp_put_str_hdr(o, dl); mp_copy(o + h, tmp, dl); }
    return mp_str_hdr_size(dl) + dl;
}

/* flb_pack_json_state() after the tokeniser: which tokens count, then tokens_to_msgpack().  o == NULL measures.
 * tmp: len + 1 bytes of scratch. */
FLB_HD void jm_pack(const uint8_t *js, uint32_t len, const struct jm_tok *tok, uint32_t toknext, int tret, uint8_t *o, uint8_t *tmp,
                    struct jm_result *r)
{
    int32_t count = 0, last = 0, records = 0, i;
    uint32_t k = 0;
    int refused = 0;
    (void) len;
    r->toknext = toknext; r->status = JM_OK; r->last_byte = 0; r->tokens_count = 0; r->records = 0; r->out_size = 0;
    if (tret == JM_PART) {
        /* the last top-level token (complete or not) delimits what is whole: everything before it is packed */
        int found = 0, delim = 0;
        if (toknext == 0) { r->status = JM_PART; return; }
        for (i = (int32_t) toknext - 1; i >= 1; i--) {
            if (tok[i].parent == -1 && tok[i].end != 0) { found = 1; delim = i; break; }
        }
        if (!found) { r->status = JM_PART; return; }
        count = delim;
    }
    else if (tret != JM_OK) { r->status = tret; return; }
    else count = (int32_t) toknext;
    r->tokens_count = count;
    if (count == 0) { r->status = JM_INVAL; return; }
    for (i = 0; i < count; i++) {
        const struct jm_tok *t = &tok[i];
        const uint32_t flen = (uint32_t) (t->end - t->start);
        if (t->start < 0 || t->end <= 0) { r->status = JM_FAIL; return; }
        if (t->parent == -1) { last = t->end; records++; }
        switch (t->type) {
        case JM_OBJECT:
            if (o) mp_put_map_hdr(o + k, (uint32_t) t->size);
            k += mp_cnt_hdr_size((uint32_t) t->size);
            break;
        case JM_ARRAY:
            if (o) mp_put_array_hdr(o + k, (uint32_t) t->size);
            k += mp_cnt_hdr_size((uint32_t) t->size);
            break;
        case JM_STRING:
            k += jm_string(js + t->start, flen, o ? o + k : 0, tmp);
            break;
        case JM_PRIMITIVE: {
            const uint8_t *p = js + t->start;
            if (*p == 'f') { if (o) o[k] = 0xc2; k += 1; }
            else if (*p == 't') { if (o) o[k] = 0xc3; k += 1; }
            else if (*p == 'n') { if (o) o[k] = 0xc0; k += 1; }
            else k += jm_number(p, flen, o ? o + k : 0, &refused);
            break;
        }
        default: r->status = JM_FAIL; return;
        }
    }
    if (refused) { r->status = JM_REFUSED; return; }
    r->out_size = k; r->last_byte = last; r->records = records;
}

#endif
