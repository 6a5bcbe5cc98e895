/* oracle: flb_parser_create()/flb_parser_do() for the regex, json, ltsv and logfmt formats.
 * TEST INFRASTRUCTURE (see orc.h).
 *
 *   src/flb_parser.c:148-345     flb_parser_create (time format bookkeeping: year, tz, %L split)
 *   src/flb_parser.c:1159-1278   flb_parser_time_lookup, parse_subseconds (:1134)
 *   src/flb_parser.c:1280-1378   flb_parser_typecast
 *   src/flb_parser_regex.c:46-215
 *   src/flb_parser_json.c:29-250 (+ src/flb_pack.c JSON -> msgpack through yyjson)
 *   src/flb_parser_ltsv.c:82-197, src/flb_parser_logfmt.c:63-254
 * libc's atoll/strtoull/strtod/timegm are used exactly where the reference uses them. */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "orc_flb.h"

/* ---- creation ------------------------------------------------------------------------------ */
/* src/flb_parser.c:1063-1125 flb_parser_tzone_offset */
static int tzone_offset(const char *str, int len, int *tmdiff)
{
    int neg;
    long hour, min;
    const char *end, *p = str;
    if (*p == 'Z') { *tmdiff = 0; return 0; }
    if (*p != '+' && *p != '-') { *tmdiff = 0; return -1; }
    if (len < 4) { *tmdiff = 0; return -1; }
    neg = (*p++ == '-');
    end = str + len;
    hour = ((p[0] - '0') * 10) + (p[1] - '0');
    if (end - p == 5 && p[2] == ':') min = ((p[3] - '0') * 10) + (p[4] - '0');
    else min = ((p[2] - '0') * 10) + (p[3] - '0');
    if (hour < 0 || hour > 59 || min < 0 || min > 59) return -1;
    *tmdiff = (int) ((hour * 3600) + (min * 60));
    if (neg) *tmdiff = -*tmdiff;
    return 0;
}

struct orc_parser *orc_parser_create(struct orc_config *cfg, const char *name, const char *format, const char *regex,
                                     int skip_empty, const char *time_fmt, const char *time_key, const char *time_offset,
                                     int time_keep, int time_strict, int logfmt_no_bare_keys, const char *types_spec)
{
    struct orc_parser *p = calloc(1, sizeof(*p));
    snprintf(p->name, sizeof(p->name), "%s", name);
    if (!strcasecmp(format, "regex")) p->type = ORC_P_REGEX;
    else if (!strcasecmp(format, "json")) p->type = ORC_P_JSON;
    else if (!strcasecmp(format, "ltsv")) p->type = ORC_P_LTSV;
    else if (!strcasecmp(format, "logfmt")) p->type = ORC_P_LOGFMT;
    else { free(p); return NULL; }
    if (p->type == ORC_P_REGEX) {
        char err[128];
        if (!regex) { free(p); return NULL; }
        p->regex = orc_regex_create(regex, err, sizeof(err));
        if (!p->regex) { free(p); return NULL; }
    }
    p->skip_empty = skip_empty; p->time_keep = time_keep; p->time_strict = time_strict;
    p->logfmt_no_bare_keys = logfmt_no_bare_keys;
    if (time_fmt) {
        char *l;
        int is_epoch = 0;
        p->time_fmt = strdup(time_fmt);
        if (strstr(time_fmt, "%Y") || strstr(time_fmt, "%y")) p->with_year = 1;
        else if (strstr(time_fmt, "%s")) { is_epoch = 1; p->with_year = 1; }
        else {
            p->with_year = 0;
            p->time_fmt_year = malloc(strlen(time_fmt) + 4);
            memcpy(p->time_fmt_year, "%Y ", 3);
            strcpy(p->time_fmt_year + 3, time_fmt);
        }
        if (strstr(time_fmt, "%z") || strstr(time_fmt, "%Z") || strstr(time_fmt, "%SZ") || strstr(time_fmt, "%S.%LZ")) p->with_tz = 1;
        l = strstr((is_epoch || p->with_year) ? p->time_fmt : p->time_fmt_year, "%L");
        if (l) { l[0] = 0; l[1] = 0; p->time_frac_secs = l + 2; }
        if (time_offset) {
            int diff = 0;
            if (tzone_offset(time_offset, (int) strlen(time_offset), &diff) == -1) { free(p); return NULL; }
            p->time_offset = diff;
        }
    }
    if (time_key) p->time_key = strdup(time_key);
    if (types_spec) {              /* "key:type key:type" as in the parsers file (src/flb_parser.c:proc_types_str) */
        char *dup = strdup(types_spec), *tok, *save = NULL;
        for (tok = strtok_r(dup, " ", &save); tok && p->n_types < 32; tok = strtok_r(NULL, " ", &save)) {
            char *c = strchr(tok, ':');
            int t = ORC_T_STRING;
            if (!c) continue;
            *c++ = 0;
            if (!strcasecmp(c, "integer")) t = ORC_T_INT;
            else if (!strcasecmp(c, "bool")) t = ORC_T_BOOL;
            else if (!strcasecmp(c, "float")) t = ORC_T_FLOAT;
            else if (!strcasecmp(c, "hex")) t = ORC_T_HEX;
            p->types[p->n_types].key = strdup(tok);
            p->types[p->n_types].type = t;
            p->n_types++;
        }
        free(dup);
    }
    p->next = cfg->parsers; cfg->parsers = p;
    return p;
}

struct orc_parser *orc_parser_get(struct orc_config *cfg, const char *name)
{
    struct orc_parser *p;
    for (p = cfg->parsers; p; p = p->next) if (!strcmp(p->name, name)) return p;
    return NULL;
}

/* ---- time ----------------------------------------------------------------------------------- */
/* src/flb_parser.c:1134-1157 */
static int parse_subseconds(const char *str, int len, double *subsec)
{
    char buf[16], *end;
    int consumed, digits = 9;
    if (len < digits) digits = len;
    memcpy(buf, "0.", 2);
    memcpy(buf + 2, str, (size_t) digits);
    buf[digits + 2] = 0;
    *subsec = strtod(buf, &end);
    consumed = (int) (end - buf) - 2;
    if (consumed <= 0) return -1;
    return consumed;
}

/* src/flb_parser.c:1159-1278 */
int orc_time_lookup(const struct orc_parser *parser, const char *time_str, size_t tsize, struct orc_tm *tm, double *ns)
{
    char tmp[64];
    const char *time_ptr;
    int time_len = (int) tsize, used, ret;
    const char *p;
    *ns = 0;
    if (tsize > sizeof(tmp) - 1) return -1;
    if (!parser->with_year) {
        time_t now = orc_now ? orc_now : time(NULL);
        struct tm tmy;
        if (time_len + 6 >= (int) sizeof(tmp)) return -1;
        gmtime_r(&now, &tmy);
        tm->mon = tmy.tm_mon; tm->mday = tmy.tm_mday;
        snprintf(tmp, sizeof(tmp), "%04d ", tmy.tm_year + 1900);
        memcpy(tmp + 5, time_str, (size_t) time_len);
        tmp[5 + time_len] = 0;
        time_ptr = tmp; time_len = (int) strlen(tmp);
        used = orc_strptime(time_ptr, parser->time_fmt_year, tm, 0);
    }
    else {
        if (time_len >= (int) sizeof(tmp)) return -1;
        memcpy(tmp, time_str, (size_t) time_len);
        tmp[time_len] = 0;
        time_ptr = tmp; time_len = (int) strlen(tmp);
        used = orc_strptime(time_ptr, parser->time_fmt, tm, 0);
    }
    if (used < 0) return parser->time_strict ? -1 : 0;
    p = time_ptr + used;
    if (parser->time_frac_secs) {
        ret = parse_subseconds(p, time_len - (int) (p - time_ptr), ns);
        if (ret < 0) return parser->time_strict ? -1 : 0;
        p += ret;
        used = orc_strptime(p, parser->time_frac_secs, tm, 0);
        if (used < 0) return parser->time_strict ? -1 : 0;
    }
    if (!parser->with_tz) tm->gmtoff = parser->time_offset;
    return 0;
}

/* include/fluent-bit/flb_parser.h:78-92 (time_system_timezone is not part of the GPU path) */
static int64_t tm2time(const struct orc_tm *tm) { return orc_timegm(tm) - tm->gmtoff; }

/* ---- typecast -------------------------------------------------------------------------------- */
/* src/flb_parser.c:1280-1378 */
static void typecast(const struct orc_parser *p, struct orc_buf *b, const char *key, int key_len, const char *val, int val_len)
{
    int i;
    for (i = 0; i < p->n_types; i++) {
        if ((int) strlen(p->types[i].key) == key_len && !strncmp(key, p->types[i].key, (size_t) key_len)) {
            char *tmp = strndup(val, (size_t) val_len);
            int error = 0;
            ov_pack_str(b, key, (size_t) key_len);
            switch (p->types[i].type) {
            case ORC_T_INT: ov_pack_int(b, atoll(tmp)); break;
            case ORC_T_HEX: ov_pack_uint(b, strtoull(tmp, NULL, 16)); break;
            case ORC_T_FLOAT: ov_pack_double(b, atof(tmp)); break;
            case ORC_T_BOOL:
                if (val_len >= 4 && !strncasecmp(val, "true", 4)) orc_buf_u8(b, 0xc3);
                else if (val_len >= 5 && !strncasecmp(val, "false", 5)) orc_buf_u8(b, 0xc2);
                else error = 1;
                break;
            default: ov_pack_str(b, val, (size_t) val_len); break;
            }
            if (error) ov_pack_str(b, val, (size_t) val_len);
            free(tmp);
            return;
        }
    }
    ov_pack_str(b, key, (size_t) key_len);
    ov_pack_str(b, val, (size_t) val_len);
}

/* ---- regex parser ------------------------------------------------------------------------------ */
/* src/flb_parser_regex.c:113-215 with cb_results (:46-111) */
static int regex_do(const struct orc_parser *p, const char *buf, size_t length, struct orc_buf *out, int64_t *sec, int64_t *nsec)
{
    int ng = orc_regex_ngroups(p->regex), nn = orc_regex_nnames(p->regex), i, n, skipped = 0, last_pos = -1;
    int *region = malloc(sizeof(int) * 2 * (size_t) ng);
    int64_t lookup = 0;
    double tfrac = 0;
    size_t hdr_at = out->n;
    const char *time_key = p->time_key ? p->time_key : "time";

    if (orc_regex_search(p->regex, (const uint8_t *) buf, length, region, ng) != 1) { free(region); return -1; }
    n = ng - 1;                                        /* region->num_regs - 1 (src/flb_regex.c:223) */
    if (n <= 0) { free(region); return -1; }
    ov_pack_map_hdr(out, (uint32_t) n);                /* header style is chosen for n, patched below */
    for (i = 0; i < nn; i++) {                         /* onig_foreach_name: definition order */
        int g, beg, end;
        const char *name = orc_regex_name(p->regex, i, &g);
        size_t vlen;
        const char *val;
        beg = region[2 * g]; end = region[2 * g + 1];
        if (end >= 0) last_pos = end;
        vlen = (size_t) (end - beg);
        val = buf + beg;
        if (vlen == 0 && p->skip_empty) { skipped++; continue; }
        if (p->time_fmt && strcmp(name, time_key) == 0) {
            struct orc_tm tm;
            double frac = 0;
            memset(&tm, 0, sizeof(tm));
            if (orc_time_lookup(p, val, vlen, &tm, &frac) == -1) { skipped++; continue; }
            tfrac = frac;
            lookup = tm2time(&tm);
            if (!p->time_keep) { skipped++; continue; }
        }
        if (p->n_types) typecast(p, out, name, (int) strlen(name), val, (int) vlen);
        else { ov_pack_str(out, name, strlen(name)); ov_pack_str(out, val, vlen); }
    }
    free(region);
    if (last_pos == -1) { out->n = hdr_at; return -1; }
    if (skipped > 0) {                                 /* patch the count in place, keep the header type */
        uint32_t cnt = (uint32_t) (n - skipped);
        uint8_t *h = out->p + hdr_at;
        if ((h[0] >> 4) == 0x8) h[0] = (uint8_t) (0x80 | (cnt & 0x0f));
        else if (h[0] == 0xde) { h[1] = (uint8_t) (cnt >> 8); h[2] = (uint8_t) cnt; }
        else if (h[0] == 0xdf) { h[1] = (uint8_t) (cnt >> 24); h[2] = (uint8_t) (cnt >> 16); h[3] = (uint8_t) (cnt >> 8); h[4] = (uint8_t) cnt; }
    }
    *sec = lookup;
    *nsec = (int64_t) (tfrac * 1000000000);
    return last_pos;
}

/* ---- JSON -> msgpack (src/flb_pack.c through yyjson 0.10, default read flags) ------------------- */
struct jp { const uint8_t *s; size_t n, p; struct orc_buf *o; int depth; };

static void jws(struct jp *j) { while (j->p < j->n && (j->s[j->p] == ' ' || j->s[j->p] == '\t' || j->s[j->p] == '\n' || j->s[j->p] == '\r')) j->p++; }

static int jhex4(struct jp *j, uint32_t *v)
{
    int i;
    uint32_t r = 0;
    if (j->p + 4 > j->n) return -1;
    for (i = 0; i < 4; i++) {
        int c = j->s[j->p + i], d;
        if (c >= '0' && c <= '9') d = c - '0'; else if (c >= 'a' && c <= 'f') d = c - 'a' + 10; else if (c >= 'A' && c <= 'F') d = c - 'A' + 10; else return -1;
        r = r * 16 + (uint32_t) d;
    }
    j->p += 4; *v = r;
    return 0;
}

static int jstring(struct jp *j, struct orc_buf *tmp)
{
    tmp->n = 0;
    j->p++;
    for (;;) {
        uint32_t c;
        if (j->p >= j->n) return -1;
        c = j->s[j->p];
        if (c == '"') { j->p++; return 0; }
        if (c < 0x20) return -1;
        if (c == '\\') {
            if (++j->p >= j->n) return -1;
            c = j->s[j->p++];
            switch (c) {
            case '"': case '\\': case '/': orc_buf_u8(tmp, c); break;
            case 'b': orc_buf_u8(tmp, 8); break;
            case 'f': orc_buf_u8(tmp, 12); break;
            case 'n': orc_buf_u8(tmp, 10); break;
            case 'r': orc_buf_u8(tmp, 13); break;
            case 't': orc_buf_u8(tmp, 9); break;
            case 'u': {
                uint32_t cp, lo;
                if (jhex4(j, &cp)) return -1;
                if (cp >= 0xd800 && cp <= 0xdbff) {
                    if (j->p + 2 > j->n || j->s[j->p] != '\\' || j->s[j->p + 1] != 'u') return -1;
                    j->p += 2;
                    if (jhex4(j, &lo) || lo < 0xdc00 || lo > 0xdfff) return -1;
                    cp = 0x10000 + ((cp - 0xd800) << 10) + (lo - 0xdc00);
                }
                else if (cp >= 0xdc00 && cp <= 0xdfff) return -1;
                if (cp < 0x80) orc_buf_u8(tmp, cp);
                else if (cp < 0x800) { orc_buf_u8(tmp, 0xc0 | (cp >> 6)); orc_buf_u8(tmp, 0x80 | (cp & 63)); }
                else if (cp < 0x10000) { orc_buf_u8(tmp, 0xe0 | (cp >> 12)); orc_buf_u8(tmp, 0x80 | ((cp >> 6) & 63)); orc_buf_u8(tmp, 0x80 | (cp & 63)); }
                else { orc_buf_u8(tmp, 0xf0 | (cp >> 18)); orc_buf_u8(tmp, 0x80 | ((cp >> 12) & 63)); orc_buf_u8(tmp, 0x80 | ((cp >> 6) & 63)); orc_buf_u8(tmp, 0x80 | (cp & 63)); }
                break;
            }
            default: return -1;
            }
            continue;
        }
        if (c < 0x80) { orc_buf_u8(tmp, c); j->p++; continue; }
        {   /* UTF-8 must be well formed (RFC 3629: no overlongs, no surrogates, <= U+10FFFF) */
            int len = c >= 0xf0 ? 4 : c >= 0xe0 ? 3 : c >= 0xc2 ? 2 : 0, i;
            uint32_t b1;
            if (!len || c > 0xf4 || j->p + (size_t) len > j->n) return -1;
            b1 = j->s[j->p + 1];
            for (i = 1; i < len; i++) if ((j->s[j->p + i] & 0xc0) != 0x80) return -1;
            if (c == 0xe0 && b1 < 0xa0) return -1;
            if (c == 0xed && b1 > 0x9f) return -1;
            if (c == 0xf0 && b1 < 0x90) return -1;
            if (c == 0xf4 && b1 > 0x8f) return -1;
            orc_buf_put(tmp, j->s + j->p, (size_t) len);
            j->p += (size_t) len;
        }
    }
}

static int jnumber(struct jp *j)
{
    size_t st = j->p, q = j->p;
    int neg = 0, is_real = 0, nd = 0;
    if (q < j->n && j->s[q] == '-') { neg = 1; q++; }
    if (q >= j->n || j->s[q] < '0' || j->s[q] > '9') return -1;
    if (j->s[q] == '0') { q++; if (q < j->n && j->s[q] >= '0' && j->s[q] <= '9') return -1; }
    else while (q < j->n && j->s[q] >= '0' && j->s[q] <= '9') { q++; nd++; }
    if (q < j->n && j->s[q] == '.') {
        size_t f = ++q;
        while (q < j->n && j->s[q] >= '0' && j->s[q] <= '9') q++;
        if (q == f) return -1;
        is_real = 1;
    }
    if (q < j->n && (j->s[q] == 'e' || j->s[q] == 'E')) {
        size_t f;
        q++;
        if (q < j->n && (j->s[q] == '+' || j->s[q] == '-')) q++;
        f = q;
        while (q < j->n && j->s[q] >= '0' && j->s[q] <= '9') q++;
        if (q == f) return -1;
        is_real = 1;
    }
    {
        char tmp[512];
        size_t len = q - st;
        if (len >= sizeof(tmp)) return -1;
        memcpy(tmp, j->s + st, len); tmp[len] = 0;
        if (!is_real) {
            /* integers that fit: uint64 when positive, int64 when negative; otherwise a real */
            unsigned long long u = 0;
            const char *d = tmp + neg;
            int overflow = 0;
            for (; *d; d++) {
                unsigned dig = (unsigned) (*d - '0');
                if (u > (0xffffffffffffffffull - dig) / 10) { overflow = 1; break; }
                u = u * 10 + dig;
            }
            if (!overflow && !neg) { ov_pack_uint(j->o, u); j->p = q; return 0; }
            if (!overflow && neg && u <= 9223372036854775808ull) { ov_pack_int(j->o, (int64_t) (0 - u)); j->p = q; return 0; }
        }
        {
            double dv = strtod(tmp, NULL);
            if (dv - dv != 0) return -1;               /* overflow to infinity: yyjson rejects the document */
            ov_pack_double(j->o, dv);
        }
    }
    j->p = q;
    return 0;
}

static int jvalue(struct jp *j)
{
    jws(j);
    if (j->p >= j->n) return -1;
    switch (j->s[j->p]) {
    case '{': case '[': {
        int obj = j->s[j->p] == '{';
        struct orc_buf body = { 0, 0, 0 }, *save = j->o, tmp = { 0, 0, 0 };
        uint32_t cnt = 0;
        int rc = -1;
        if (++j->depth > 1024) return -1;
        j->p++;
        j->o = &body;
        jws(j);
        if (j->p < j->n && j->s[j->p] == (obj ? '}' : ']')) { j->p++; rc = 0; }
        else for (;;) {
            if (obj) {
                jws(j);
                if (j->p >= j->n || j->s[j->p] != '"' || jstring(j, &tmp)) break;
                ov_pack_str(&body, tmp.p, tmp.n);
                jws(j);
                if (j->p >= j->n || j->s[j->p] != ':') break;
                j->p++;
            }
            if (jvalue(j)) break;
            cnt++;
            jws(j);
            if (j->p >= j->n) break;
            if (j->s[j->p] == ',') { j->p++; continue; }
            if (j->s[j->p] == (obj ? '}' : ']')) { j->p++; rc = 0; }
            break;
        }
        j->o = save; j->depth--;
        if (rc == 0) {
            if (obj) ov_pack_map_hdr(j->o, cnt); else ov_pack_arr_hdr(j->o, cnt);
            orc_buf_put(j->o, body.p, body.n);
        }
        free(body.p); free(tmp.p);
        return rc;
    }
    case '"': {
        struct orc_buf tmp = { 0, 0, 0 };
        int rc = jstring(j, &tmp);
        if (rc == 0) ov_pack_str(j->o, tmp.p, tmp.n);
        free(tmp.p);
        return rc;
    }
    case 't': if (j->p + 4 <= j->n && !memcmp(j->s + j->p, "true", 4)) { j->p += 4; orc_buf_u8(j->o, 0xc3); return 0; } return -1;
    case 'f': if (j->p + 5 <= j->n && !memcmp(j->s + j->p, "false", 5)) { j->p += 5; orc_buf_u8(j->o, 0xc2); return 0; } return -1;
    case 'n': if (j->p + 4 <= j->n && !memcmp(j->s + j->p, "null", 4)) { j->p += 4; orc_buf_u8(j->o, 0xc0); return 0; } return -1;
    default: return jnumber(j);
    }
}

/* src/flb_parser_json.c:29-250 */
static int json_do(const struct orc_parser *p, const char *buf, size_t length, struct orc_buf *out, int64_t *sec, int64_t *nsec)
{
    struct orc_buf mp = { 0, 0, 0 }, scratch = { 0, 0, 0 };
    struct jp j = { (const uint8_t *) buf, length, 0, &mp, 0 };
    struct orc_arena arena = { 0 };
    struct ov map;
    size_t off = 0, consumed;
    uint32_t i, skip;
    const char *time_key = p->time_key ? p->time_key : "time";
    int ret = -1;

    if (jvalue(&j)) goto done;
    consumed = j.p;
    {   /* flb_pack_json_recs counts documents: a second complete one rejects the line */
        struct jp j2 = { (const uint8_t *) buf, length, j.p, &scratch, 0 };
        jws(&j2);
        if (j2.p < length && jvalue(&j2) == 0) goto done;
    }
    if (ov_unpack(&arena, mp.p, mp.n, &off, &map) != 0 || map.type != OV_MAP) goto done;
    ret = (int) consumed;
    if (!p->time_fmt) { orc_buf_put(out, mp.p, mp.n); goto done; }
    skip = map.n;
    for (i = 0; i < map.n; i++) {
        struct ov *k = &map.items[2 * i];
        if (k->len != strlen(time_key)) continue;      /* (compares via.str.size whatever the key type) */
        if (k->type != OV_STR && k->type != OV_BIN) continue;
        if (strncmp((const char *) k->p, time_key, k->len) == 0) break;
    }
    if (i >= map.n || map.items[2 * i + 1].type != OV_STR) { orc_buf_put(out, mp.p, mp.n); goto done; }
    {
        struct ov *v = &map.items[2 * i + 1];
        struct orc_tm tm;
        double frac = 0;
        int64_t lookup;
        uint32_t m;
        memset(&tm, 0, sizeof(tm));
        skip = p->time_keep ? 0xffffffffu : i;
        if (orc_time_lookup(p, (const char *) v->p, v->len, &tm, &frac) == -1) { lookup = 0; skip = map.n; }
        else lookup = tm2time(&tm);
        ov_pack_map_hdr(out, (!p->time_keep && skip < map.n) ? map.n - 1 : map.n);
        for (m = 0; m < map.n; m++) {
            if (m == skip) continue;
            ov_pack(out, &map.items[2 * m]); ov_pack(out, &map.items[2 * m + 1]);
        }
        *sec = lookup;
        *nsec = (int64_t) (frac * 1000000000);
    }
done:
    free(mp.p); free(scratch.p);
    orc_arena_free(&arena);
    return ret;
}

/* ---- LTSV (src/flb_parser_ltsv.c:82-197) --------------------------------------------------------- */
static int ltsv_label(int c) { return (c >= '0' && c <= '9') || (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || c == '_' || c == '.' || c == '-'; }
static int ltsv_field(int c) { return c != 0 && c != '\t' && c != '\n' && c != '\r'; }

static int pack_kv_time(const struct orc_parser *p, struct orc_buf *body, uint32_t *cnt, const char *k, size_t kl, const char *v, size_t vl,
                        int value_kind, int64_t *lookup, double *frac)
{
    /* shared tail of ltsv_parser / logfmt_parser: Time_Key lookup, Time_Keep, typecast.
     * value_kind: 0 string, 1 logfmt bare key (true), 2 empty quoted string */
    const char *time_key = p->time_key ? p->time_key : "time";
    int time_found = 0;
    if (p->time_fmt && kl == strlen(time_key) && vl > 0 && !strncmp(k, time_key, kl)) {
        struct orc_tm tm;
        memset(&tm, 0, sizeof(tm));
        if (orc_time_lookup(p, v, vl, &tm, frac) == -1) return -1;
        *lookup = tm2time(&tm);
        time_found = 1;
    }
    if (time_found && !p->time_keep) return 0;
    if (p->n_types) typecast(p, body, k, (int) kl, v, (int) vl);
    else {
        ov_pack_str(body, k, kl);
        if (value_kind == 1) orc_buf_u8(body, 0xc3); else ov_pack_str(body, v, vl);
    }
    (*cnt)++;
    return 0;
}

static int ltsv_do(const struct orc_parser *p, const char *s, size_t n, struct orc_buf *out, int64_t *sec, int64_t *nsec)
{
    struct orc_buf body = { 0, 0, 0 };
    size_t c = 0;
    uint32_t cnt = 0;
    int64_t lookup = 0;
    double frac = 0;
    while (c < n) {
        size_t label = c, label_len, field, field_len;
        while (c < n && ltsv_label((unsigned char) s[c])) c++;
        label_len = c - label;
        if (c == n || s[c] != ':') break;
        c++;
        field = c;
        while (c < n && ltsv_field((unsigned char) s[c])) c++;
        field_len = c - field;
        if (label_len > 0 && pack_kv_time(p, &body, &cnt, s + label, label_len, s + field, field_len, 0, &lookup, &frac)) { free(body.p); return -1; }
        if (c == n) break;
        if (s[c] == '\t') c++;
        if (c == n) break;
        if (s[c] == '\r' || s[c] == '\n') break;
    }
    if (cnt == 0) { free(body.p); return -1; }
    ov_pack_map_hdr(out, cnt);
    orc_buf_put(out, body.p, body.n);
    free(body.p);
    *sec = lookup; *nsec = (int64_t) (frac * 1000000000);
    return (int) c;
}

/* ---- logfmt (src/flb_parser_logfmt.c:63-254) ---------------------------------------------------- */
/* src/flb_unescape.c:186-271 flb_unescape_string_utf8 with u8_read_escape_sequence (:80-184) and
 * u8_wc_toutf8 (:40-66); the caller takes strlen() of the result (flb_parser_logfmt.c:207). */
static int hexval(int c)
{
    if (c >= '0' && c <= '9') return c - '0';
    if (c >= 'a' && c <= 'f') return c - 'a' + 10;
    if (c >= 'A' && c <= 'F') return c - 'A' + 10;
    return -1;
}

static size_t read_escape(const char *str, size_t size, uint32_t *dest)
{
    uint32_t ch = (uint32_t) (int32_t) (signed char) str[0], v = 0;
    size_t i = 1;
    int dno = 0;
    switch (str[0]) {
    case 'n': ch = 10; break;
    case 't': ch = 9; break;
    case 'r': ch = 13; break;
    case 'b': ch = 8; break;
    case 'f': ch = 12; break;
    case 'v': ch = 11; break;
    case 'a': ch = 7; break;
    case 'x':
        while (i < size && hexval((unsigned char) str[i]) >= 0 && dno < 2) { v = v * 16 + (uint32_t) hexval((unsigned char) str[i]); i++; dno++; }
        if (dno > 0) ch = v;
        break;
    case 'U':
        while (i < size && hexval((unsigned char) str[i]) >= 0 && dno < 8) { v = v * 16 + (uint32_t) hexval((unsigned char) str[i]); i++; dno++; }
        if (dno > 0) ch = v;
        break;
    case 'u':
        while (i < size && hexval((unsigned char) str[i]) >= 0 && dno < 4) { v = v * 16 + (uint32_t) hexval((unsigned char) str[i]); i++; dno++; }
        if (dno != 4 && dno > 0) { ch = 0xfffd; break; }
        ch = v;
        if (ch >= 0xdc00 && ch <= 0xdfff) ch = 0xfffd;
        else if (ch >= 0xd800 && ch <= 0xdbff) {
            uint32_t low = 0;
            if (!(i + 2 < size && str[i] == '\\' && str[i + 1] == 'u')) { ch = 0xfffd; break; }
            i += 2; dno = 0;
            while (i < size && hexval((unsigned char) str[i]) >= 0 && dno < 4) { low = low * 16 + (uint32_t) hexval((unsigned char) str[i]); i++; dno++; }
            if (dno != 4 && dno > 0) ch = 0xfffd;
            else if (low >= 0xdc00 && low <= 0xdfff) ch = 0x10000 + (((ch - 0xd800) << 10) | (low - 0xdc00));
            else ch = 0xfffd;
        }
        break;
    default:
        if (str[0] >= '0' && str[0] <= '7') {
            i = 0;
            do { v = v * 8 + (uint32_t) (str[i] - '0'); i++; dno++; } while (i < size && str[i] >= '0' && str[i] <= '7' && dno < 3);
            ch = v;
        }
    }
    *dest = ch;
    return i;
}

static size_t unescape_utf8(const char *in, size_t sz, char *out)
{
    size_t ci = 0, co = 0;
    while (ci < sz && in[ci]) {
        uint32_t ch;
        size_t used = 1, len;
        if (in[ci] == '\\' && ci + 1 < sz) {
            used = 2;
            switch (in[ci + 1]) {
            case '"': case '\'': case '\\': case '/': ch = (uint32_t) in[ci + 1]; break;
            case 'n': ch = 10; break;
            case 'b': ch = 8; break;
            case 't': ch = 9; break;
            case 'f': ch = 12; break;
            case 'r': ch = 13; break;
            default: used = read_escape(in + ci + 1, sz - ci - 1, &ch) + 1;
            }
        }
        else ch = (uint32_t) (int32_t) (signed char) in[ci];
        ci += used;
        len = ch < 0x80 ? 1 : ch < 0x800 ? 2 : ch < 0x10000 ? 3 : ch < 0x110000 ? 4 : 0;
        if (len > sz - co) break;
        if (len <= 1) out[co++] = (char) ch;
        else if (len == 2) { out[co++] = (char) ((ch >> 6) | 0xc0); out[co++] = (char) ((ch & 0x3f) | 0x80); }
        else if (len == 3) { out[co++] = (char) ((ch >> 12) | 0xe0); out[co++] = (char) (((ch >> 6) & 0x3f) | 0x80); out[co++] = (char) ((ch & 0x3f) | 0x80); }
        else { out[co++] = (char) ((ch >> 18) | 0xf0); out[co++] = (char) (((ch >> 12) & 0x3f) | 0x80); out[co++] = (char) (((ch >> 6) & 0x3f) | 0x80); out[co++] = (char) ((ch & 0x3f) | 0x80); }
    }
    out[co] = 0;
    return strlen(out);
}

static int logfmt_ident(int c) { return c > ' ' && c != '=' && c != '"'; }

static int logfmt_do(const struct orc_parser *p, const char *s, size_t n, struct orc_buf *out, int64_t *sec, int64_t *nsec)
{
    struct orc_buf body = { 0, 0, 0 };
    size_t c = 0;
    uint32_t cnt = 0;
    int64_t lookup = 0;
    double frac = 0;
    while (c < n) {
        size_t key, key_len, value = 0, value_len = 0;
        int value_set = 0, value_str = 0, value_escape = 0;
        while (c < n && !logfmt_ident((unsigned char) s[c])) c++;
        if (c == n) break;
        key = c;
        while (c < n && logfmt_ident((unsigned char) s[c])) c++;
        key_len = c - key;
        if (c < n && s[c] == '=') {
            value_set = 1;
            c++;
            if (c < n) {
                if (s[c] == '"') {
                    c++; value = c; value_str = 1;
                    while (c < n) {
                        if (s[c] != '\\' && s[c] != '"') c++;
                        else if (s[c] == '\\') { value_escape = 1; c++; if (c == n) break; c++; }
                        else break;
                    }
                    value_len = c - value;
                    if (c < n && s[c] == '"') c++;
                }
                else {
                    value = c;
                    while (c < n && logfmt_ident((unsigned char) s[c])) c++;
                    value_len = c - value;
                }
            }
        }
        if (key_len > 0) {
            if (p->logfmt_no_bare_keys && value_len == 0 && !value_set) { free(body.p); return -1; }
            if (value_escape && value_len > 0 && !p->n_types) {
                /* the Time_Key test sees the raw text; the packed value is the unescaped one (:159-214) */
                char *un = malloc(value_len + 1);
                size_t ul = unescape_utf8(s + value, value_len, un);
                const char *time_key = p->time_key ? p->time_key : "time";
                int is_time = p->time_fmt && key_len == strlen(time_key) && !strncmp(s + key, time_key, key_len);
                int rc = 0, time_found = 0;
                if (is_time) {
                    struct orc_tm tm;
                    memset(&tm, 0, sizeof(tm));
                    if (orc_time_lookup(p, s + value, value_len, &tm, &frac) == -1) rc = -1;
                    else { lookup = tm2time(&tm); time_found = 1; }
                }
                if (!rc && (!time_found || p->time_keep)) { ov_pack_str(&body, s + key, key_len); ov_pack_str(&body, un, ul); cnt++; }
                free(un);
                if (rc) { free(body.p); return -1; }
            }
            else if (pack_kv_time(p, &body, &cnt, s + key, key_len, s + value, value_len,
                             (value_len == 0 && !value_str && !p->n_types) ? 1 : 0, &lookup, &frac)) { free(body.p); return -1; }
        }
        if (c == n) break;
        if (s[c] == '\r' || s[c] == '\n') break;
    }
    if (cnt == 0) { free(body.p); return -1; }
    ov_pack_map_hdr(out, cnt);
    orc_buf_put(out, body.p, body.n);
    free(body.p);
    *sec = lookup; *nsec = (int64_t) (frac * 1000000000);
    return (int) c;
}

/* src/flb_parser.c:1044-1061 */
int orc_parser_do(const struct orc_parser *p, const char *buf, size_t length, struct orc_buf *out, int64_t *sec, int64_t *nsec)
{
    *sec = 0; *nsec = 0;
    switch (p->type) {
    case ORC_P_REGEX: return regex_do(p, buf, length, out, sec, nsec);
    case ORC_P_JSON: return json_do(p, buf, length, out, sec, nsec);
    case ORC_P_LTSV: return ltsv_do(p, buf, length, out, sec, nsec);
    case ORC_P_LOGFMT: return logfmt_do(p, buf, length, out, sec, nsec);
    }
    return -1;
}
