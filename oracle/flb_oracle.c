/* oracle: the filter callbacks of Fluent Bit's parse->filter path, restated on the CPU.
 * TEST INFRASTRUCTURE (see orc.h) -- never linked into, loaded by or called from the product.
 *
 *   src/flb_log_event_decoder.c:214-457   event framing, group markers, where decoding stops
 *   src/flb_log_event_encoder*.c          [[ext(0) sec32 nsec32, meta], body] output records
 *   src/flb_ra_key.c, src/flb_record_accessor.c   $key['sub'][0] lookups (last duplicate wins)
 *   plugins/filter_parser/filter_parser.c:174-442
 *   plugins/filter_grep/grep.c:56-392
 *   plugins/filter_modify/modify.c:141-1578
 *   plugins/filter_record_modifier/filter_modifier.c:37-486
 *   plugins/filter_log_to_metrics/log_to_metrics.c:247-1148 (counter, gauge, histogram)
 *   src/flb_filter.c:119-323             flb_filter_do: the chain, MODIFIED / NOTOUCH hand-over
 */
#define _GNU_SOURCE
#include <ctype.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "orc_flb.h"

time_t orc_now = 0;

#define ORC_MODIFIED 1
#define ORC_NOTOUCH 2

/* ---- events ------------------------------------------------------------------------------------ */
struct orc_event {
    struct ov root, *meta, *body;
    int64_t sec, nsec;
    size_t base, len;
};
static struct ov g_empty_map = { OV_MAP, 0, 0, 0, 0, 0, 0, 0, 0 };

/* flb_event_decoder_decode_object + decode_timestamp (src/flb_log_event_decoder.c:172-297).
 * 0 decoded, 1 the buffer ended (cleanly or inside an object), -1 anything undecodable. */
static int next_event(struct orc_arena *a, const uint8_t *buf, size_t len, size_t *off, struct orc_event *ev)
{
    for (;;) {
        size_t prev = *off;
        struct ov *ts;
        int r;
        if (len == 0 || *off >= len) return 1;
        r = ov_unpack(a, buf, len, off, &ev->root);
        if (r == 1) return 1;
        if (r != 0) return -1;
        if (ev->root.type != OV_ARR || ev->root.n != 2) return -1;
        if (ev->root.items[0].type == OV_ARR) {
            if (ev->root.items[0].n != 2) return -1;
            ts = &ev->root.items[0].items[0];
            ev->meta = &ev->root.items[0].items[1];
        }
        else { ts = &ev->root.items[0]; ev->meta = &g_empty_map; }
        if (ts->type != OV_UINT && ts->type != OV_F64 && ts->type != OV_F32 && ts->type != OV_EXT) return -1;
        if (ev->meta->type != OV_MAP) return -1;
        ev->body = &ev->root.items[1];
        if (ev->body->type != OV_MAP) return -1;
        ev->nsec = 0;
        if (ts->type == OV_UINT) ev->sec = (int64_t) ts->u;
        else if (ts->type == OV_F64) { ev->sec = (int64_t) ts->d; ev->nsec = (int64_t) ((ts->d - (double) ev->sec) * 1000000000); }
        else if (ts->type == OV_F32) return -1;        /* MSGPACK_OBJECT_FLOAT is the 64-bit type only */
        else {
            if (ts->ext != 0 || ts->len != 8) return -1;
            ev->sec = (int32_t) ((uint32_t) ts->p[0] << 24 | (uint32_t) ts->p[1] << 16 | (uint32_t) ts->p[2] << 8 | ts->p[3]);
            ev->nsec = (int32_t) ((uint32_t) ts->p[4] << 24 | (uint32_t) ts->p[5] << 16 | (uint32_t) ts->p[6] << 8 | ts->p[7]);
        }
        ev->base = prev; ev->len = *off - prev;
        if ((int32_t) ev->sec < 0) continue;           /* group markers and invalid negatives are skipped (:362-447) */
        return 0;
    }
}

/* one record as the log event encoder writes it (FLB_LOG_EVENT_FORMAT_FLUENT_BIT_V2) */
static void emit_header(struct orc_buf *o, int64_t sec, int64_t nsec, const struct ov *meta)
{
    uint8_t h[12] = { 0x92, 0x92, 0xd7, 0x00 };
    uint32_t s = (uint32_t) sec, n = (uint32_t) nsec;
    h[4] = (uint8_t) (s >> 24); h[5] = (uint8_t) (s >> 16); h[6] = (uint8_t) (s >> 8); h[7] = (uint8_t) s;
    h[8] = (uint8_t) (n >> 24); h[9] = (uint8_t) (n >> 16); h[10] = (uint8_t) (n >> 8); h[11] = (uint8_t) n;
    orc_buf_put(o, h, 12);
    ov_pack(o, meta);
}

/* ---- record accessor ----------------------------------------------------------------------------- */
struct orc_ra { char *key; int nsub; struct { int is_index; int index; char *str; } sub[32]; };

static struct orc_ra *ra_create(const char *s)
{
    struct orc_ra *ra = calloc(1, sizeof(*ra));
    const char *p;
    if (s[0] != '$') { ra->key = strdup(s); return ra; }
    p = s + 1;
    {
        const char *st = p;
        while (*p && *p != '[' && *p != '.' && *p != ' ' && *p != ',' && *p != '"') p++;
        ra->key = strndup(st, (size_t) (p - st));
    }
    while (*p == '[' && ra->nsub < 32) {
        p++;
        if (*p == '\'') {
            const char *st = ++p;
            while (*p && *p != '\'') p++;
            ra->sub[ra->nsub].str = strndup(st, (size_t) (p - st));
            if (*p) p++;
        }
        else { ra->sub[ra->nsub].is_index = 1; ra->sub[ra->nsub].index = atoi(p); while (isdigit((unsigned char) *p)) p++; }
        if (*p == ']') p++;
        ra->nsub++;
    }
    return ra;
}

/* src/flb_ra_key.c: ra_key_val_id (last matching STR key), subkey_to_object */
static const struct ov *map_find_last(const struct ov *map, const char *key, size_t kl)
{
    const struct ov *hit = NULL;
    uint32_t i;
    if (map->type != OV_MAP) return NULL;
    for (i = 0; i < map->n; i++) if (ov_str_eq(&map->items[2 * i], key, kl)) hit = &map->items[2 * i + 1];
    return hit;
}

static const struct ov *ra_get(const struct orc_ra *ra, const struct ov *map)
{
    const struct ov *v = map_find_last(map, ra->key, strlen(ra->key));
    int i;
    if (!v) return NULL;
    if (ra->nsub && v->type != OV_MAP && v->type != OV_ARR) return v;   /* flb_ra_key_to_value: subkeys only walk containers */
    for (i = 0; i < ra->nsub; i++) {
        if (ra->sub[i].is_index) {
            if (v->type != OV_ARR || (uint32_t) ra->sub[i].index >= v->n) return NULL;
            v = &v->items[ra->sub[i].index];
        }
        else {
            if (v->type != OV_MAP) return NULL;
            v = map_find_last(v, ra->sub[i].str, strlen(ra->sub[i].str));
            if (!v) return NULL;
        }
    }
    return v;
}

/* flb_ra_regex_match (src/flb_record_accessor.c:753): the value must be a string */
static int ra_regex_match(const struct orc_ra *ra, const struct ov *map, const struct orc_regex *rx)
{
    const struct ov *v = ra_get(ra, map);
    int region[2];
    if (!v || v->type != OV_STR) return 0;
    return orc_regex_search(rx, v->p, v->len, region, 1) == 1;
}

/* ---- filters ---------------------------------------------------------------------------------------- */
enum { F_PARSER = 1, F_GREP, F_MODIFY, F_RECMOD, F_L2M };
enum { GREP_REGEX = 1, GREP_EXCLUDE };
enum { OP_LEGACY, OP_OR, OP_AND };
enum { C_KEY_EXISTS, C_KEY_DOES_NOT_EXIST, C_A_KEY_MATCHES, C_NO_KEY_MATCHES, C_KEY_VALUE_EQUALS, C_KEY_VALUE_DOES_NOT_EQUAL,
       C_KEY_VALUE_MATCHES, C_KEY_VALUE_DOES_NOT_MATCH, C_MATCHING_KEYS_HAVE_MATCHING_VALUES,
       C_MATCHING_KEYS_DO_NOT_HAVE_MATCHING_VALUES };
enum { R_RENAME, R_HARD_RENAME, R_ADD, R_SET, R_REMOVE, R_REMOVE_WILDCARD, R_REMOVE_REGEX, R_COPY, R_HARD_COPY,
       R_MOVE_TO_START, R_MOVE_TO_END };

struct kvp { char *k, *v; struct kvp *next; };
struct grep_rule { int type; struct orc_ra *ra; struct orc_regex *rx; };
struct mod_cond { int type; struct orc_ra *ra; struct orc_regex *a_rx, *b_rx; char *b; };
struct mod_rule { int type; char *key, *val; struct orc_regex *key_rx; };
struct rm_key { char *key; size_t len; int dynamic; };
struct l2m_set { char **labels; uint64_t count; double sum; uint64_t *buckets; };

struct orc_filter {
    struct orc_filter *next;
    int kind;
    struct kvp *props, *props_tail;
    /* parser */
    char *key_name; struct orc_ra *ra_key; struct orc_parser *parsers[16]; int n_parsers, reserve_data, preserve_key;
    /* grep (also the gate of log_to_metrics) */
    struct grep_rule rules[64]; int n_rules, op;
    /* modify */
    struct mod_cond conds[64]; int n_conds; struct mod_rule mrules[256]; int n_mrules;
    /* record_modifier */
    struct rm_key remove[64], allow[64]; int n_remove, n_allow; struct { char *k, *v; } records[64]; int n_records;
    /* log_to_metrics */
    int mode, discard, n_labels, n_buckets;
    char *label_keys[16]; struct orc_ra *label_ras[16], *value_ra;
    double buckets[64];
    char *ns, *subsystem, *mname;
    struct l2m_set *sets; int n_sets;
};

static int split_tokens(const char *line, int max_split, char **out, int max_out)
{
    /* flb_utils_split / split_quoted (src/flb_utils.c): max_split tokens, then the rest verbatim */
    int n = 0;
    const char *p = line;
    while (*p && n < max_out) {
        while (*p == ' ') p++;
        if (!*p) break;
        if (max_split > 0 && n >= max_split) { out[n++] = strdup(p); break; }
        if (*p == '"' || *p == '\'') {
            char q = *p++;
            const char *st = p;
            while (*p && *p != q) p++;
            out[n++] = strndup(st, (size_t) (p - st));
            if (*p) p++;
        }
        else {
            const char *st = p;
            while (*p && *p != ' ') p++;
            out[n++] = strndup(st, (size_t) (p - st));
        }
        if (*p == ' ') p++;
    }
    return n;
}

static int parse_bool(const char *v)
{
    return !strcasecmp(v, "true") || !strcasecmp(v, "on") || !strcasecmp(v, "yes");
}

struct orc_filter *orc_filter_create(struct orc_config *cfg, const char *plugin)
{
    struct orc_filter *f = calloc(1, sizeof(*f));
    if (!strcasecmp(plugin, "parser")) f->kind = F_PARSER;
    else if (!strcasecmp(plugin, "grep")) f->kind = F_GREP;
    else if (!strcasecmp(plugin, "modify")) f->kind = F_MODIFY;
    else if (!strcasecmp(plugin, "record_modifier")) f->kind = F_RECMOD;
    else if (!strcasecmp(plugin, "log_to_metrics")) f->kind = F_L2M;
    else { free(f); return NULL; }
    if (cfg->filters_tail) cfg->filters_tail->next = f; else cfg->filters = f;
    cfg->filters_tail = f;
    return f;
}

int orc_filter_set(struct orc_filter *f, const char *k, const char *v)
{
    struct kvp *n = calloc(1, sizeof(*n));
    n->k = strdup(k); n->v = strdup(v);
    if (f->props_tail) f->props_tail->next = n; else f->props = n;
    f->props_tail = n;
    return 0;
}

static int add_grep_rule(struct orc_filter *f, int type, const char *val, int dollar)
{
    char *tok[3], field[512], err[64];
    int nt = split_tokens(val, 1, tok, 3);
    struct grep_rule *r = &f->rules[f->n_rules];
    if (nt != 2 || f->n_rules >= 64) return -1;
    /* grep.c:109-117 prepends '$' to a bare key; log_to_metrics.c:318 does not */
    if (dollar && tok[0][0] != '$') snprintf(field, sizeof(field), "$%s", tok[0]);
    else snprintf(field, sizeof(field), "%s", tok[0]);
    r->type = type;
    r->ra = ra_create(field);
    r->rx = orc_regex_create(tok[1], err, sizeof(err));
    if (!r->rx) return -1;
    f->n_rules++;
    return 0;
}

static int cmp_double(const void *a, const void *b) { double x = *(const double *) a, y = *(const double *) b; return x < y ? -1 : x > y; }

int orc_filter_init(struct orc_config *cfg, struct orc_filter *f)
{
    struct kvp *p;
    char err[64];
    if (f->kind == F_PARSER) {
        for (p = f->props; p; p = p->next) {
            if (!strcasecmp(p->k, "key_name")) { if (p->v[0] == '$') f->ra_key = ra_create(p->v); else f->key_name = strdup(p->v); }
            else if (!strcasecmp(p->k, "parser")) {
                struct orc_parser *ps = orc_parser_get(cfg, p->v);
                if (!ps || f->n_parsers >= 16) return -1;
                f->parsers[f->n_parsers++] = ps;
            }
            else if (!strcasecmp(p->k, "reserve_data")) f->reserve_data = parse_bool(p->v);
            else if (!strcasecmp(p->k, "preserve_key")) f->preserve_key = parse_bool(p->v);
        }
        return (f->n_parsers && (f->key_name || f->ra_key)) ? 0 : -1;
    }
    if (f->kind == F_GREP) {
        int first = 0;
        f->op = OP_LEGACY;
        for (p = f->props; p; p = p->next)
            if (!strcasecmp(p->k, "logical_op")) f->op = !strcasecmp(p->v, "AND") ? OP_AND : !strcasecmp(p->v, "OR") ? OP_OR : OP_LEGACY;
        for (p = f->props; p; p = p->next) {
            int type = !strcasecmp(p->k, "regex") ? GREP_REGEX : !strcasecmp(p->k, "exclude") ? GREP_EXCLUDE : 0;
            if (!type) {
                if (!strcasecmp(p->k, "logical_op")) continue;
                return -1;               /* flb_filter_config_map_set(): a key outside the plugin's config map stops flb_start() */
            }
            if (f->op != OP_LEGACY && first && first != type) return -1;
            first = type;
            if (add_grep_rule(f, type, p->v, 1)) return -1;
        }
        return 0;
    }
    if (f->kind == F_MODIFY) {
        static const char *cn[] = { "key_exists", "key_does_not_exist", "a_key_matches", "no_key_matches", "key_value_equals",
                                    "key_value_does_not_equal", "key_value_matches", "key_value_does_not_match",
                                    "matching_keys_have_matching_values", "matching_keys_do_not_have_matching_values" };
        static const char *rn[] = { "rename", "hard_rename", "add", "set", "remove", "remove_wildcard", "remove_regex", "copy",
                                    "hard_copy", "move_to_start", "move_to_end" };
        for (p = f->props; p; p = p->next) {
            char *tok[4];
            int nt = split_tokens(p->v, 3, tok, 4), i;
            if (!strcasecmp(p->k, "condition")) {
                struct mod_cond *c = &f->conds[f->n_conds];
                int t = -1;
                if (nt < 2 || f->n_conds >= 64) return -1;
                for (i = 0; i < 10; i++) if (!strcasecmp(tok[0], cn[i])) t = i;
                if (t < 0) return -1;
                c->type = t;
                if (t == C_A_KEY_MATCHES || t == C_NO_KEY_MATCHES || t >= C_MATCHING_KEYS_HAVE_MATCHING_VALUES) {
                    if (!(c->a_rx = orc_regex_create(tok[1], err, sizeof(err)))) return -1;
                }
                else c->ra = ra_create(tok[1]);
                if (t == C_KEY_VALUE_EQUALS || t == C_KEY_VALUE_DOES_NOT_EQUAL) { if (nt < 3) return -1; c->b = strdup(tok[2]); }
                if (t == C_KEY_VALUE_MATCHES || t == C_KEY_VALUE_DOES_NOT_MATCH || t >= C_MATCHING_KEYS_HAVE_MATCHING_VALUES) {
                    if (nt < 3 || !(c->b_rx = orc_regex_create(tok[2], err, sizeof(err)))) return -1;
                }
                f->n_conds++;
            }
            else {
                struct mod_rule *r = &f->mrules[f->n_mrules];
                int t = -1;
                for (i = 0; i < 11; i++) if (!strcasecmp(p->k, rn[i])) t = i;
                if (t < 0 || nt < 1 || nt > 3 || f->n_mrules >= 256) return -1;
                {   /* modify.c:412-466: one word names the removal / move rules, two words the others; three words leave
                     * the calloc()ed type 0 = RENAME in place, with the first and the last word */
                    int one = t == R_REMOVE || t == R_REMOVE_WILDCARD || t == R_REMOVE_REGEX || t == R_MOVE_TO_START || t == R_MOVE_TO_END;
                    if (nt == 3) t = R_RENAME;
                    else if ((nt == 1) != one) return -1;
                }
                r->type = t;
                r->key = strdup(tok[0]);
                r->val = nt > 1 ? strdup(tok[nt - 1]) : strdup("");
                if (t == R_REMOVE_REGEX && !(r->key_rx = orc_regex_create(tok[0], err, sizeof(err)))) return -1;
                {   /* modify.c:468-507: key and value text of every rule must be regexes Onigmo accepts */
                    struct orc_regex *probe;
                    if (!(probe = orc_regex_create(tok[0], err, sizeof(err))) && !strstr(err, "not supported")) return -1;
                    if (!(probe = orc_regex_create(tok[nt - 1], err, sizeof(err))) && !strstr(err, "not supported")) return -1;
                    (void) probe;
                }
                f->n_mrules++;
            }
        }
        return 0;
    }
    if (f->kind == F_RECMOD) {
        for (p = f->props; p; p = p->next) {
            if (!strcasecmp(p->k, "record")) {
                char *tok[3];
                if (split_tokens(p->v, 1, tok, 3) != 2 || f->n_records >= 64) return -1;
                f->records[f->n_records].k = tok[0]; f->records[f->n_records].v = tok[1]; f->n_records++;
            }
            else if (!strcasecmp(p->k, "remove_key") || !strcasecmp(p->k, "allowlist_key") || !strcasecmp(p->k, "whitelist_key")) {
                int rm = !strcasecmp(p->k, "remove_key");
                struct rm_key *k = rm ? &f->remove[f->n_remove] : &f->allow[f->n_allow];
                k->key = strdup(p->v); k->len = strlen(p->v);
                if (k->len && k->key[k->len - 1] == '*') { k->dynamic = 1; k->len--; }
                if (rm) f->n_remove++; else f->n_allow++;
            }
        }
        if (f->n_remove > 0 && f->n_allow > 0) return -1;      /* filter_modifier.c: "remove_keys and allowlist_keys are exclusive" */
        return 0;
    }
    {   /* log_to_metrics.c:649-962 */
        const char *mode = "counter", *value_field = NULL, *desc = NULL, *tag = NULL;
        static const double defb[11] = { 0.005, 0.01, 0.025, 0.05, 0.1, 0.25, 0.5, 1.0, 2.5, 5.0, 10.0 };
        f->ns = strdup("log_metric"); f->mname = strdup("a");
        for (p = f->props; p; p = p->next) {     /* kubernetes_mode: five fixed labels first (:43-50, 148-156) */
            if (!strcasecmp(p->k, "kubernetes_mode") && parse_bool(p->v) && f->n_labels == 0) {
                static const char *k8s[5] = { "namespace_name", "pod_name", "container_name", "docker_id", "pod_id" };
                int q;
                for (q = 0; q < 5; q++) {
                    char acc[64];
                    snprintf(acc, sizeof(acc), "$kubernetes['%s']", k8s[q]);
                    f->label_keys[f->n_labels] = strdup(k8s[q]); f->label_ras[f->n_labels] = ra_create(acc); f->n_labels++;
                }
            }
        }
        for (p = f->props; p; p = p->next) {
            if (!strcasecmp(p->k, "regex")) { if (add_grep_rule(f, GREP_REGEX, p->v, 0)) return -1; }
            else if (!strcasecmp(p->k, "exclude")) { if (add_grep_rule(f, GREP_EXCLUDE, p->v, 0)) return -1; }
            else if (!strcasecmp(p->k, "metric_mode")) mode = p->v;
            else if (!strcasecmp(p->k, "value_field")) value_field = p->v;
            else if (!strcasecmp(p->k, "metric_name")) f->mname = strdup(p->v);
            else if (!strcasecmp(p->k, "metric_namespace")) f->ns = strdup(p->v);
            else if (!strcasecmp(p->k, "metric_subsystem")) f->subsystem = strdup(p->v);
            else if (!strcasecmp(p->k, "metric_description")) desc = p->v;
            else if (!strcasecmp(p->k, "tag")) tag = p->v;
            else if (!strcasecmp(p->k, "discard_logs")) f->discard = parse_bool(p->v);
            else if (!strcasecmp(p->k, "bucket")) { if (f->n_buckets < 64) f->buckets[f->n_buckets++] = strtod(p->v, NULL); }
            else if (!strcasecmp(p->k, "label_field") && f->n_labels < 16) {
                f->label_keys[f->n_labels] = strdup(p->v); f->label_ras[f->n_labels] = ra_create(p->v); f->n_labels++;
            }
            else if (!strcasecmp(p->k, "add_label") && f->n_labels < 16) {
                char *tok[3];
                if (split_tokens(p->v, 1, tok, 3) != 2) return -1;
                f->label_keys[f->n_labels] = tok[0]; f->label_ras[f->n_labels] = ra_create(tok[1]); f->n_labels++;
            }
        }
        if (!tag || !*tag || !desc || !*desc) return -1;
        if (!strcasecmp(mode, "counter")) f->mode = 0;
        else if (!strcasecmp(mode, "gauge")) f->mode = 1;
        else if (!strcasecmp(mode, "histogram")) f->mode = 2;
        else return -1;
        if (f->mode != 0) {
            if (!value_field || !*value_field) return -1;
            f->value_ra = ra_create(value_field);
        }
        if (f->mode == 2) {
            if (f->n_buckets == 0) { memcpy(f->buckets, defb, sizeof(defb)); f->n_buckets = 11; }
            else qsort(f->buckets, (size_t) f->n_buckets, sizeof(double), cmp_double);
        }
        if (!f->subsystem || !*f->subsystem) f->subsystem = strdup(mode);   /* :762-770: an empty subsystem falls back to the mode */
        return 0;
    }
}


/* msgpack-c's streaming parser (lib/msgpack-c/include/msgpack/unpack_template.h: template_execute) on the bytes
 * behind the last whole event: it eats a header byte, then that header's fixed-size part (length field or scalar)
 * only if all of it is there, then a payload only if all of it is there.  When the buffer runs out exactly at one
 * of those points, inside the first unfinished object, msgpack_unpack_next() reports CONTINUE with the offset at
 * the end of the buffer -- which flb_log_event_decoder_next() turns into INSUFFICIENT_DATA and the filters
 * (grep.c:357-360, modify.c) accept as a clean end when `offset == bytes`.  1 = that is the case for b[0..n). */
static int tail_runs_out_cleanly(const uint8_t *b, size_t n)
{
    size_t p = 0;
    uint64_t open[64];                  /* elements still owed per open container */
    int depth = 0;
    if (n == 0) return 1;
    for (;;) {
        uint32_t c, fixed = 0, is_len = 0, items = 0, is_container = 0;
        uint64_t payload = 0;
        if (p == n) return 1;                                   /* ran out between two objects of an open container */
        c = b[p++];
        if (c <= 0x7f || c >= 0xe0 || c == 0xc0 || c == 0xc2 || c == 0xc3) { }
        else if (c >= 0xa0 && c <= 0xbf) payload = c & 31;
        else if (c >= 0x90 && c <= 0x9f) { is_container = 1; items = c & 15; }
        else if (c >= 0x80 && c <= 0x8f) { is_container = 1; items = 2 * (c & 15); }
        else switch (c) {
        case 0xcc: case 0xd0: fixed = 1; break;
        case 0xcd: case 0xd1: fixed = 2; break;
        case 0xce: case 0xd2: case 0xca: fixed = 4; break;
        case 0xcf: case 0xd3: case 0xcb: fixed = 8; break;
        case 0xd4: fixed = 2; break; case 0xd5: fixed = 3; break; case 0xd6: fixed = 5; break;
        case 0xd7: fixed = 9; break; case 0xd8: fixed = 17; break;
        case 0xd9: case 0xc4: fixed = 1; is_len = 1; break;
        case 0xda: case 0xc5: fixed = 2; is_len = 1; break;
        case 0xdb: case 0xc6: fixed = 4; is_len = 1; break;
        case 0xc7: fixed = 1; is_len = 2; break;
        case 0xc8: fixed = 2; is_len = 2; break;
        case 0xc9: fixed = 4; is_len = 2; break;
        case 0xdc: fixed = 2; is_len = 3; break;
        case 0xdd: fixed = 4; is_len = 3; break;
        case 0xde: fixed = 2; is_len = 4; break;
        case 0xdf: fixed = 4; is_len = 4; break;
        default: return 0;                                      /* 0xc1: a parse error, not a shortage */
        }
        if (fixed) {
            uint64_t v = 0;
            uint32_t i;
            if (n - p < fixed) return p == n;                   /* stops behind the header byte */
            for (i = 0; i < fixed; i++) v = (v << 8) | b[p + i];
            p += fixed;
            if (is_len == 1) payload = v;
            else if (is_len == 2) payload = v + 1;              /* ext: type byte + data */
            else if (is_len == 3) { is_container = 1; items = (uint32_t) v; if (v > 0x7fffffffu) return 0; }
            else if (is_len == 4) { is_container = 1; if (v > 0x3fffffffu) return 0; items = (uint32_t) (2 * v); }
        }
        if (payload) {
            if (n - p < payload) return p == n;                 /* stops where the payload begins */
            p += (size_t) payload;
        }
        if (is_container && items) {
            if (depth >= 64) return 0;
            open[depth++] = items;
            continue;
        }
        /* one object done: pay it to the containers it closes */
        for (;;) {
            if (depth == 0) return 0;                           /* a whole top-level object fits: not a shortage */
            if (--open[depth - 1]) break;
            depth--;
        }
    }
}

/* grep.c:167-194 (legacy) and :250-284 (AND / OR) */
static int grep_keep(const struct orc_filter *f, const struct ov *map)
{
    int i, found = 0;
    if (f->op == OP_LEGACY) {
        for (i = 0; i < f->n_rules; i++) {
            int m = ra_regex_match(f->rules[i].ra, map, f->rules[i].rx);
            if (!m) { if (f->rules[i].type == GREP_REGEX) return 0; }
            else return f->rules[i].type == GREP_EXCLUDE ? 0 : 1;
        }
        return 1;
    }
    if (f->n_rules == 0) return 1;
    for (i = 0; i < f->n_rules; i++) {
        found = ra_regex_match(f->rules[i].ra, map, f->rules[i].rx);
        if (f->op == OP_OR && found) break;
        if (f->op == OP_AND && !found) break;
    }
    if (i == f->n_rules) i = f->n_rules - 1;
    return f->rules[i].type == GREP_REGEX ? found : !found;
}

/* ---- per-filter callbacks.  in/out are whole chunks. ------------------------------------------------ */
static int cb_grep(struct orc_filter *f, const uint8_t *in, size_t len, struct orc_buf *out)
{
    struct orc_arena a = { 0 };
    struct orc_event ev;
    size_t off = 0;
    int old = 0, kept = 0, r;
    while ((r = next_event(&a, in, len, &off, &ev)) == 0) {
        old++;
        if (grep_keep(f, ev.body)) { orc_buf_put(out, in + ev.base, ev.len); kept++; }
    }
    orc_arena_free(&a);
    if (old == kept) return ORC_NOTOUCH;
    if (!(r == 1 && (off == len || tail_runs_out_cleanly(in + off, len - off)))) return ORC_NOTOUCH;     /* decoder stopped early: "Log event encoder error" */
    return ORC_MODIFIED;
}

static int str_or_bin(const struct ov *v) { return v->type == OV_STR || v->type == OV_BIN; }

static int cb_parser(struct orc_filter *f, const uint8_t *in, size_t len, struct orc_buf *out)
{
    struct orc_arena a = { 0 };
    struct orc_event ev;
    size_t off = 0;
    while (next_event(&a, in, len, &off, &ev) == 0) {
        struct orc_buf parsed = { 0, 0, 0 };
        const struct ov *map = ev.body;
        uint8_t *keep = calloc(map->n + 1, 1);
        int have_arr = f->reserve_data || f->preserve_key, preserved = -1, parse_ok = 0, pi;
        int64_t sec = ev.sec, nsec = ev.nsec;
        uint32_t i;
        if (f->reserve_data) memset(keep, 1, map->n);
        if (f->ra_key) {
            const struct ov *v = ra_get(f->ra_key, map);
            if (v && str_or_bin(v)) {
                for (pi = 0; pi < f->n_parsers; pi++) {
                    int64_t ps = 0, pns = 0;
                    parsed.n = 0;
                    if (orc_parser_do(f->parsers[pi], (const char *) v->p, v->len, &parsed, &ps, &pns) >= 0) {
                        if ((uint64_t) ps * 1000000000ull + (uint64_t) pns != 0) { sec = ps; nsec = pns; }
                        parse_ok = 1;
                        break;
                    }
                }
            }
        }
        else {
            for (i = 0; i < map->n; i++) {
                const struct ov *k = &map->items[2 * i], *v = &map->items[2 * i + 1];
                if (!str_or_bin(k) || k->len != strlen(f->key_name) || strncmp((const char *) k->p, f->key_name, k->len)) continue;
                if (!str_or_bin(v)) continue;
                parse_ok = 0;                                  /* parse_ret of the LAST attempt decides (filter_parser.c:355) */
                for (pi = 0; pi < f->n_parsers; pi++) {
                    int64_t ps = 0, pns = 0;
                    parsed.n = 0;
                    if (orc_parser_do(f->parsers[pi], (const char *) v->p, v->len, &parsed, &ps, &pns) >= 0) {
                        if ((uint64_t) ps * 1000000000ull + (uint64_t) pns != 0) { sec = ps; nsec = pns; }
                        parse_ok = 1;
                        if (have_arr) { if (!f->preserve_key) keep[i] = 0; else if (!f->reserve_data) preserved = (int) i; }
                        break;
                    }
                }
            }
        }
        emit_header(out, sec, nsec, ev.meta);
        if (parse_ok) {
            uint32_t extra = 0;
            if (f->reserve_data) { for (i = 0; i < map->n; i++) extra += keep[i]; }
            else if (preserved >= 0) extra = 1;
            if (extra == 0) orc_buf_put(out, parsed.p, parsed.n);
            else {                                             /* flb_msgpack_expand_map (src/flb_pack.c): full re-pack */
                struct ov pm;
                size_t po = 0;
                uint32_t m;
                ov_unpack(&a, parsed.p, parsed.n, &po, &pm);
                ov_pack_map_hdr(out, pm.n + extra);
                for (m = 0; m < 2 * pm.n; m++) ov_pack(out, &pm.items[m]);
                if (f->reserve_data) { for (i = 0; i < map->n; i++) if (keep[i]) { ov_pack(out, &map->items[2 * i]); ov_pack(out, &map->items[2 * i + 1]); } }
                else { ov_pack(out, &map->items[2 * preserved]); ov_pack(out, &map->items[2 * preserved + 1]); }
            }
        }
        else ov_pack(out, map);
        free(keep);
        free(parsed.p);
        /* `parsed` must outlive `pm` above: both are consumed before the free */
    }
    orc_arena_free(&a);
    return out->n > 0 ? ORC_MODIFIED : ORC_NOTOUCH;
}

/* a record as an editable list of pairs */
struct pairs { struct ov *k, *v; int n, cap; };
static void pairs_push(struct pairs *p, struct ov k, struct ov v)
{
    if (p->n == p->cap) { p->cap = p->cap ? p->cap * 2 : 32; p->k = realloc(p->k, sizeof(struct ov) * (size_t) p->cap); p->v = realloc(p->v, sizeof(struct ov) * (size_t) p->cap); }
    p->k[p->n] = k; p->v[p->n] = v; p->n++;
}
static int key_is(const struct ov *k, const char *s) { return str_or_bin(k) && k->len == strlen(s) && !memcmp(k->p, s, k->len); }
static int key_prefix(const struct ov *k, const char *s) { return str_or_bin(k) && k->len >= strlen(s) && !memcmp(k->p, s, strlen(s)); }
/* modify.c:523-552: STR, or BOOLEAN as "true"/"false" */
static int obj_rx(const struct ov *o, const struct orc_regex *rx)
{
    int region[2];
    if (o->type == OV_STR) return orc_regex_search(rx, o->p, o->len, region, 1) == 1;
    if (o->type == OV_BOOL) return orc_regex_search(rx, (const uint8_t *) (o->u ? "true" : "false"), o->u ? 4 : 5, region, 1) == 1;
    return 0;
}
static int count_keys(const struct pairs *p, const char *s) { int i, c = 0; for (i = 0; i < p->n; i++) c += key_is(&p->k[i], s); return c; }
static void pairs_remove(struct pairs *p, const uint8_t *del)
{
    int i, j = 0;
    for (i = 0; i < p->n; i++) if (!del[i]) { p->k[j] = p->k[i]; p->v[j] = p->v[i]; j++; }
    p->n = j;
}

/* modify.c:746-953 */
static int mod_conditions(const struct orc_filter *f, const struct ov *map, const struct pairs *p)
{
    int ci, ok = 1, i;
    for (ci = 0; ci < f->n_conds; ci++) {
        const struct mod_cond *c = &f->conds[ci];
        const struct ov *v = c->ra ? ra_get(c->ra, map) : NULL;
        int r = 0, cnt = 0;
        switch (c->type) {
        case C_KEY_EXISTS: r = v != NULL; break;
        case C_KEY_DOES_NOT_EXIST: r = v == NULL; break;
        case C_A_KEY_MATCHES: case C_NO_KEY_MATCHES:
            for (i = 0; i < p->n; i++) cnt += obj_rx(&p->k[i], c->a_rx);
            r = c->type == C_A_KEY_MATCHES ? cnt > 0 : cnt == 0;
            break;
        case C_KEY_VALUE_EQUALS: r = v && key_is(v, c->b); break;
        case C_KEY_VALUE_DOES_NOT_EQUAL: r = v && !key_is(v, c->b); break;
        case C_KEY_VALUE_MATCHES: r = v && obj_rx(v, c->b_rx); break;
        case C_KEY_VALUE_DOES_NOT_MATCH: r = v && !obj_rx(v, c->b_rx); break;
        default:
            r = 1;
            for (i = 0; i < p->n; i++) if (obj_rx(&p->k[i], c->a_rx) && !obj_rx(&p->v[i], c->b_rx)) { r = 0; break; }
            if (c->type == C_MATCHING_KEYS_DO_NOT_HAVE_MATCHING_VALUES) r = !r;
        }
        if (!r) ok = 0;
    }
    return ok;
}

/* modify.c:955-1302; 1 = the rule re-packed the map */
static int mod_rule(const struct mod_rule *r, struct pairs *p)
{
    uint8_t *del = calloc((size_t) p->n + 1, 1);
    int i, match, conflict, ret = 0;
    struct ov kk = ov_str(r->key, strlen(r->key)), vv = ov_str(r->val, strlen(r->val));
    switch (r->type) {
    case R_RENAME: case R_HARD_RENAME:
        match = count_keys(p, r->key); conflict = count_keys(p, r->val);
        if (match == 0 || (r->type == R_RENAME && conflict > 0)) break;
        for (i = 0; i < p->n; i++) del[i] = (uint8_t) (conflict > 0 && key_is(&p->k[i], r->val));
        for (i = 0; i < p->n; i++) if (!del[i] && key_is(&p->k[i], r->key)) p->k[i] = vv;
        pairs_remove(p, del);
        ret = 1;
        break;
    case R_COPY: case R_HARD_COPY:
        match = count_keys(p, r->key); conflict = count_keys(p, r->val);
        if (match != 1 || (r->type == R_COPY && conflict > 0) || conflict > 1) break;
        if (conflict == 1) { for (i = 0; i < p->n; i++) del[i] = (uint8_t) key_is(&p->k[i], r->val); pairs_remove(p, del); }
        for (i = 0; i < p->n; i++) if (key_is(&p->k[i], r->key)) break;
        if (i < p->n) {
            int j;
            pairs_push(p, vv, p->v[i]);
            for (j = p->n - 1; j > i + 1; j--) { p->k[j] = p->k[j - 1]; p->v[j] = p->v[j - 1]; }
            p->k[i + 1] = vv; p->v[i + 1] = p->v[i];
        }
        ret = 1;
        break;
    case R_ADD:
        if (count_keys(p, r->key) != 0) break;
        pairs_push(p, kk, vv);
        ret = 1;
        break;
    case R_SET:
        for (i = 0; i < p->n; i++) del[i] = (uint8_t) key_is(&p->k[i], r->key);
        pairs_remove(p, del);
        pairs_push(p, kk, vv);
        ret = 1;
        break;
    case R_REMOVE: case R_REMOVE_WILDCARD: case R_REMOVE_REGEX:
        match = 0;
        for (i = 0; i < p->n; i++) {
            del[i] = (uint8_t) (r->type == R_REMOVE ? key_is(&p->k[i], r->key) : r->type == R_REMOVE_WILDCARD ? key_prefix(&p->k[i], r->key)
                                                                               : obj_rx(&p->k[i], r->key_rx));
            match += del[i];
        }
        if (match) { pairs_remove(p, del); ret = 1; }
        break;
    case R_MOVE_TO_START: case R_MOVE_TO_END: {
        struct pairs q = { 0, 0, 0, 0 };
        int first = r->type == R_MOVE_TO_START;
        match = 0;
        for (i = 0; i < p->n; i++) { del[i] = (uint8_t) key_prefix(&p->k[i], r->key); match += del[i]; }
        if (!match) break;
        for (i = 0; i < p->n; i++) if (del[i] == first) pairs_push(&q, p->k[i], p->v[i]);
        for (i = 0; i < p->n; i++) if (del[i] != first) pairs_push(&q, p->k[i], p->v[i]);
        for (i = 0; i < p->n; i++) { p->k[i] = q.k[i]; p->v[i] = q.v[i]; }
        free(q.k); free(q.v);
        ret = 1;
        break;
    }
    }
    free(del);
    return ret;
}

static int cb_modify(struct orc_filter *f, const uint8_t *in, size_t len, struct orc_buf *out)
{
    struct orc_arena a = { 0 };
    struct orc_event ev;
    size_t off = 0;
    int total = 0, r;
    while ((r = next_event(&a, in, len, &off, &ev)) == 0) {
        struct pairs p = { 0, 0, 0, 0 };
        int modified = 0, i;
        uint32_t m;
        for (m = 0; m < ev.body->n; m++) pairs_push(&p, ev.body->items[2 * m], ev.body->items[2 * m + 1]);
        if (mod_conditions(f, ev.body, &p)) {
            for (i = 0; i < f->n_mrules; i++) modified |= mod_rule(&f->mrules[i], &p);
        }
        if (modified) {
            emit_header(out, ev.sec, ev.nsec, ev.meta);
            ov_pack_map_hdr(out, (uint32_t) p.n);
            for (i = 0; i < p.n; i++) { ov_pack(out, &p.k[i]); ov_pack(out, &p.v[i]); }
            total++;
        }
        else orc_buf_put(out, in + ev.base, ev.len);
        free(p.k); free(p.v);
    }
    orc_arena_free(&a);
    if (total == 0) return ORC_NOTOUCH;
    if (!(r == 1 && (off == len || tail_runs_out_cleanly(in + off, len - off)))) return ORC_NOTOUCH;
    return ORC_MODIFIED;
}

/* filter_modifier.c:213-279, 298-486 */
static int cb_recmod(struct orc_filter *f, const uint8_t *in, size_t len, struct orc_buf *out)
{
    struct orc_arena a = { 0 };
    struct orc_event ev;
    size_t off = 0;
    int is_modified = 0;
    while (next_event(&a, in, len, &off, &ev) == 0) {
        const struct ov *map = ev.body;
        uint8_t *del = calloc(map->n + 1, 1);
        const struct rm_key *keys = f->n_remove ? f->remove : f->n_allow ? f->allow : NULL;
        int nk = f->n_remove ? f->n_remove : f->n_allow, is_delete = f->n_remove > 0, remaining = (int) map->n, q;
        uint32_t i;
        for (i = 0; keys && i < map->n; i++) {
            const struct ov *k = &map->items[2 * i];
            int result = 0;
            for (q = 0; q < nk && str_or_bin(k); q++) {
                if (!keys[q].dynamic && k->len != keys[q].len) continue;
                if (keys[q].dynamic && k->len < keys[q].len) continue;
                if (!strncasecmp((const char *) k->p, keys[q].key, keys[q].len)) { result = 1; break; }
            }
            if (result == is_delete) { del[i] = 1; remaining--; }
        }
        if (remaining != (int) map->n) is_modified = 1;
        if (remaining + f->n_records > 0) {
            uint32_t total = (uint32_t) (remaining + f->n_records);
            uint8_t h[5] = { 0xdf, (uint8_t) (total >> 24), (uint8_t) (total >> 16), (uint8_t) (total >> 8), (uint8_t) total };
            emit_header(out, ev.sec, ev.nsec, ev.meta);
            orc_buf_put(out, h, 5);                            /* the dynamic body of the encoder is always a map32 */
            for (i = 0; i < map->n; i++) if (!del[i]) { ov_pack(out, &map->items[2 * i]); ov_pack(out, &map->items[2 * i + 1]); }
            if (f->n_records > 0) is_modified = 1;
            for (q = 0; q < f->n_records; q++) { ov_pack_str(out, f->records[q].k, strlen(f->records[q].k)); ov_pack_str(out, f->records[q].v, strlen(f->records[q].v)); }
        }
        free(del);
    }
    orc_arena_free(&a);
    return (is_modified && out->n > 0) ? ORC_MODIFIED : ORC_NOTOUCH;
}

/* log_to_metrics.c:964-1148.  Walks the chunk with msgpack_unpack_next, not the event decoder. */
static int cb_l2m(struct orc_filter *f, const uint8_t *in, size_t len, struct orc_buf *out)
{
    struct orc_arena a = { 0 };
    size_t off = 0;
    struct ov root;
    double val = 0;                    /* :983-984: lives across records, so a text sscanf() cannot convert leaves the previous value */
    (void) out;
    while (ov_unpack(&a, in, len, &off, &root) == 0) {
        const struct ov *map;
        char labels[16][256];
        int i, s;
        if (root.type != OV_ARR || root.n < 2) continue;
        map = &root.items[1];
        if (!grep_keep(f, map)) continue;
        for (i = 0; i < f->n_labels; i++) {
            const struct ov *v = ra_get(f->label_ras[i], map);
            labels[i][0] = 0;
            if (!v) continue;
            if (v->type == OV_STR) { char *t = strndup((const char *) v->p, v->len); snprintf(labels[i], 253 - 1, "%s", t); free(t); }
            else if (v->type == OV_F64 || v->type == OV_F32) snprintf(labels[i], 253 - 1, "%f", v->d);
            else if (v->type == OV_UINT) snprintf(labels[i], 253 - 1, "%ld", (long) v->u);
            else if (v->type == OV_INT) snprintf(labels[i], 253 - 1, "%ld", (long) v->i);
        }
        if (f->mode != 0) {                                    /* gauge :1052-1080, histogram :1082-1110 */
            const struct ov *v = ra_get(f->value_ra, map);
            if (!v) continue;
            if (v->type == OV_STR) { char *t = strndup((const char *) v->p, v->len); sscanf(t, "%lf", &val); free(t); }
            else if (v->type == OV_F64 || v->type == OV_F32) val = v->d;
            else if (v->type == OV_UINT) val = (double) (int64_t) v->u;
            else if (v->type == OV_INT) val = (double) v->i;
            else continue;
        }
        {   /* cmt_map_metric_get (cmt_map.c:208-224): the metric is found by the hash of the label values run together,
             * so tuples that concatenate to the same text are one metric */
            char cat[16 * 256], have[16 * 256];
            cat[0] = 0;
            for (i = 0; i < f->n_labels; i++) strcat(cat, labels[i]);
            for (s = 0; s < f->n_sets; s++) {
                have[0] = 0;
                for (i = 0; i < f->n_labels; i++) strcat(have, f->sets[s].labels[i]);
                if (!strcmp(have, cat)) break;
            }
        }
        if (s == f->n_sets) {                                  /* cmetrics appends a new metric to the map (cmt_map.c:209-243) */
            f->sets = realloc(f->sets, sizeof(*f->sets) * (size_t) (f->n_sets + 1));
            f->sets[s].labels = calloc(16, sizeof(char *));
            for (i = 0; i < f->n_labels; i++) f->sets[s].labels[i] = strdup(labels[i]);
            f->sets[s].count = 0; f->sets[s].sum = 0;
            f->sets[s].buckets = calloc((size_t) f->n_buckets + 1, sizeof(uint64_t));
            f->n_sets++;
        }
        f->sets[s].count++;
        if (f->mode == 1) f->sets[s].sum = val;                /* cmt_gauge_set: the latest record's value stays */
        if (f->mode == 2) {                                    /* cmt_histogram_observe */
            for (i = f->n_buckets - 1; i >= 0; i--) { if (val > f->buckets[i]) break; f->sets[s].buckets[i]++; }
            f->sets[s].buckets[f->n_buckets]++;
            f->sets[s].sum += val;
        }
    }
    orc_arena_free(&a);
    return f->discard ? ORC_MODIFIED : ORC_NOTOUCH;
}

/* text of the metric as cmt_encode_text_create() prints it, without the timestamp column */
char *orc_l2m_text(struct orc_filter *f)
{
    struct orc_buf b = { 0, 0, 0 };
    char tmp[512];
    int s, i;
    if (f->n_sets == 0 && f->n_labels == 0 && f->mode != 2) {      /* cmetrics' static metric: there, at 0, from creation */
        int n = snprintf(tmp, sizeof(tmp), "%s%s%s%s%s = 0\n", f->ns, *f->ns ? "_" : "", f->subsystem, *f->subsystem ? "_" : "", f->mname);
        orc_buf_put(&b, tmp, (size_t) n);
    }
    for (s = 0; s < f->n_sets; s++) {
        int n = snprintf(tmp, sizeof(tmp), "%s%s%s%s%s", f->ns, *f->ns ? "_" : "", f->subsystem, *f->subsystem ? "_" : "", f->mname);   /* cmt_opts_init: empty parts left out */
        orc_buf_put(&b, tmp, (size_t) n);
        for (i = 0; i < f->n_labels; i++) {
            n = snprintf(tmp, sizeof(tmp), "%s%s=\"%s\"", i ? "," : "{", f->label_keys[i], f->sets[s].labels[i]);
            orc_buf_put(&b, tmp, (size_t) n);
        }
        if (f->n_labels) orc_buf_put(&b, "}", 1);
        if (f->mode == 0) { n = snprintf(tmp, sizeof(tmp), " = %.17g\n", (double) f->sets[s].count); orc_buf_put(&b, tmp, (size_t) n); }
        else if (f->mode == 1) { n = snprintf(tmp, sizeof(tmp), " = %.17g\n", f->sets[s].sum); orc_buf_put(&b, tmp, (size_t) n); }
        else {
            orc_buf_put(&b, " = { buckets = { ", 17);
            for (i = 0; i < f->n_buckets; i++) {
                n = snprintf(tmp, sizeof(tmp), "%g=%llu, ", f->buckets[i], (unsigned long long) f->sets[s].buckets[i]);
                orc_buf_put(&b, tmp, (size_t) n);
            }
            n = snprintf(tmp, sizeof(tmp), "+Inf=%llu }, sum=%g, count=%llu }\n", (unsigned long long) f->sets[s].buckets[f->n_buckets],
                         f->sets[s].sum, (unsigned long long) f->sets[s].count);
            orc_buf_put(&b, tmp, (size_t) n);
        }
    }
    orc_buf_u8(&b, 0);
    return (char *) b.p;
}

/* ---- ctypes doors ------------------------------------------------------------------------------------- */
struct orc_config *orc_config_create(void) { return calloc(1, sizeof(struct orc_config)); }
void orc_set_now(long t) { orc_now = (time_t) t; }
void orc_free(void *p) { free(p); }

int orc_filter_cb(struct orc_filter *f, const void *data, size_t len, void **out, size_t *out_len)
{
    struct orc_buf b = { 0, 0, 0 };
    int r;
    switch (f->kind) {
    case F_PARSER: r = cb_parser(f, data, len, &b); break;
    case F_GREP: r = cb_grep(f, data, len, &b); break;
    case F_MODIFY: r = cb_modify(f, data, len, &b); break;
    case F_RECMOD: r = cb_recmod(f, data, len, &b); break;
    default: r = cb_l2m(f, data, len, &b); if (r == ORC_MODIFIED) b.n = 0; break;
    }
    if (r == ORC_MODIFIED) { *out = b.p ? b.p : malloc(1); *out_len = b.n; }
    else { free(b.p); *out = NULL; *out_len = 0; }
    return r;
}

/* flb_filter_do (src/flb_filter.c:119-323): every filter in creation order; a MODIFIED result
 * replaces the working buffer, an empty one ends the chain. */
int orc_chain_do(struct orc_config *cfg, const void *data, size_t len, void **out, size_t *out_len)
{
    struct orc_filter *f;
    const uint8_t *cur = data;
    uint8_t *owned = NULL;
    size_t cur_len = len;
    int modified = 0;
    for (f = cfg->filters; f; f = f->next) {
        void *o = NULL;
        size_t ol = 0;
        int r = orc_filter_cb(f, cur, cur_len, &o, &ol);
        if (r != ORC_MODIFIED) continue;
        free(owned);
        owned = o; cur = o; cur_len = ol; modified = 1;
        if (ol == 0) break;
    }
    if (!modified) { *out = NULL; *out_len = 0; return ORC_NOTOUCH; }
    *out = owned; *out_len = cur_len;
    return ORC_MODIFIED;
}

int orc_parser_do_line(struct orc_parser *p, const char *line, size_t len, void **out, size_t *out_len, int64_t *sec, int64_t *nsec)
{
    struct orc_buf b = { 0, 0, 0 };
    int r = orc_parser_do(p, line, len, &b, sec, nsec);
    if (r >= 0) { *out = b.p; *out_len = b.n; } else { free(b.p); *out = NULL; *out_len = 0; }
    return r;
}

int orc_time_lookup_str(struct orc_parser *p, const char *s, size_t n, int64_t *sec, int64_t *nsec)
{
    struct orc_tm tm;
    double frac = 0;
    memset(&tm, 0, sizeof(tm));
    if (orc_time_lookup(p, s, n, &tm, &frac) == -1) return -1;
    *sec = orc_timegm(&tm) - tm.gmtoff;
    *nsec = (int64_t) (frac * 1000000000);
    return 0;
}
