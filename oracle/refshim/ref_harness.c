/* oracle/_ref harness: a plain-C, ctypes-friendly door onto the UNMODIFIED
 * reference functions of the parse->filter path:
 *   flb_parser_create / flb_parser_do            (src/flb_parser.c:148,1044)
 *   flb_filter_new / _set_property / _init / flb_filter_do  (src/flb_filter.c:426,325,550,119)
 *   plugin cb_filter callbacks                   (plugins/filter_*)
 *   flb_regex_create / flb_regex_do              (src/flb_regex.c:160,182)
 *   flb_pack_json                                (src/flb_pack.c)
 *   flb_parser_time_lookup                       (src/flb_parser.c:1159)
 * TEST INFRASTRUCTURE ONLY: nothing under fluent-bit_b200/ may link or call this. */
#include <stdio.h>
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>
#include <fluent-bit/flb_info.h>
#include <fluent-bit/flb_config.h>
#include <fluent-bit/flb_env.h>
#include <fluent-bit/flb_mem.h>
#include <fluent-bit/flb_parser.h>
#include <fluent-bit/flb_regex.h>
#include <onigmo.h>
#include <fluent-bit/flb_pack.h>
#include <fluent-bit/flb_time.h>
#include <fluent-bit/flb_filter.h>
#include <fluent-bit/flb_input.h>
#include <fluent-bit/flb_input_chunk.h>
#include <fluent-bit/flb_mp.h>
#include <fluent-bit/flb_sds.h>
#include <cmetrics/cmetrics.h>
#include <cmetrics/cmt_encode_text.h>
#include <cmetrics/cmt_encode_msgpack.h>
#include "filter_log_to_metrics/log_to_metrics.h"
#include <fluent-bit/multiline/flb_ml.h>
#include <fluent-bit/multiline/flb_ml_parser.h>
#include <fluent-bit/multiline/flb_ml_rule.h>

extern struct flb_filter_plugin filter_parser_plugin;
extern struct flb_filter_plugin filter_grep_plugin;
extern struct flb_filter_plugin filter_modify_plugin;
extern struct flb_filter_plugin filter_record_modifier_plugin;
extern struct flb_filter_plugin filter_log_to_metrics_plugin;
extern struct flb_filter_plugin filter_rewrite_tag_plugin;
extern struct flb_filter_plugin filter_multiline_plugin;

struct flbref_cfg {
    struct flb_config *config;
    struct flb_input_instance in;
    struct flb_filter_plugin plugins[7];
};

void *flbref_config_create(void)
{
    struct flbref_cfg *c = calloc(1, sizeof(*c));
    struct flb_config *config = calloc(1, sizeof(struct flb_config));
    int i;

    c->config = config;
    mk_list_init(&config->parsers);
    mk_list_init(&config->filters);
    mk_list_init(&config->filter_plugins);
    mk_list_init(&config->inputs);
    mk_list_init(&config->cf_parsers_list);
    config->env = flb_env_create();
    /* the dlopen path memcpy's the plugin struct the same way (src/flb_plugin.c:274) */
    c->plugins[0] = filter_parser_plugin;
    c->plugins[1] = filter_grep_plugin;
    c->plugins[2] = filter_modify_plugin;
    c->plugins[3] = filter_record_modifier_plugin;
    c->plugins[4] = filter_log_to_metrics_plugin;
    c->plugins[5] = filter_rewrite_tag_plugin;
    c->plugins[6] = filter_multiline_plugin;
    /* what flb_config_init() does for the multiline core (src/flb_config.c:424-455): the parser list, the buffer limit,
     * the built-in multiline parsers */
    mk_list_init(&config->multiline_parsers);
    config->multiline_buffer_limit = flb_strdup(FLB_ML_BUFFER_LIMIT_DEFAULT_STR);
    flb_ml_parser_builtin_create(config);
    for (i = 0; i < 7; i++) {
        mk_list_add(&c->plugins[i]._head, &config->filter_plugins);
    }
    mk_list_init(&c->in.properties);
    c->in.config = config;
    return c;
}

/* what flb_plugin_load() does for a dynamic filter plugin (src/flb_plugin.c:200-324): dlopen, look the registration
 * struct up by name, memcpy it, link it to config->filter_plugins.  0 ok, -1 not. */
int flbref_plugin_load(void *cfg, const char *path, const char *struct_name)
{
    struct flbref_cfg *c = cfg;
    void *dso = dlopen(path, RTLD_LAZY);
    struct flb_filter_plugin *sym, *copy;
    if (!dso) { fprintf(stderr, "flbref_plugin_load: %s\n", dlerror()); return -1; }
    sym = dlsym(dso, struct_name);
    if (!sym) { fprintf(stderr, "flbref_plugin_load: %s lacks %s\n", path, struct_name); return -1; }
    copy = flb_malloc(sizeof(*copy));
    memcpy(copy, sym, sizeof(*copy));
    mk_list_add(&copy->_head, &c->config->filter_plugins);
    return 0;
}

/* parsers of the reference's own configuration text ([PARSER] sections, src/flb_parser.c flb_parser_conf_file) are
 * not needed here: tests create them with flbref_parser_create() */

void *flbref_config_raw(void *cfg) { return ((struct flbref_cfg *) cfg)->config; }

static int type_from_name(const char *s, int n)
{
    if (n == 7 && !strncasecmp(s, "integer", 7)) return FLB_PARSER_TYPE_INT;
    if (n == 4 && !strncasecmp(s, "bool", 4)) return FLB_PARSER_TYPE_BOOL;
    if (n == 5 && !strncasecmp(s, "float", 5)) return FLB_PARSER_TYPE_FLOAT;
    if (n == 3 && !strncasecmp(s, "hex", 3)) return FLB_PARSER_TYPE_HEX;
    return FLB_PARSER_TYPE_STRING;
}

/* types_spec: "key:integer key2:float" (same text as the Types option,
 * src/flb_parser.c proc_types_str) or NULL. */
void *flbref_parser_create(void *cfg, const char *name, const char *format, const char *regex,
                           int skip_empty, const char *time_fmt, const char *time_key,
                           const char *time_offset, int time_keep, int time_strict,
                           int logfmt_no_bare_keys, const char *types_spec)
{
    struct flbref_cfg *c = cfg;
    struct flb_parser_types *types = NULL;
    int types_len = 0;

    if (types_spec && *types_spec) {
        const char *p = types_spec;
        types = calloc(64, sizeof(*types));
        while (*p && types_len < 63) {
            const char *e, *colon;
            while (*p == ' ') p++;
            if (!*p) break;
            e = p;
            while (*e && *e != ' ') e++;
            colon = memchr(p, ':', e - p);
            if (colon) {
                types[types_len].key = strndup(p, colon - p);
                types[types_len].key_len = colon - p;
                types[types_len].type = type_from_name(colon + 1, e - colon - 1);
                types_len++;
            }
            p = e;
        }
    }
    return flb_parser_create(name, format, regex, skip_empty, time_fmt, time_key, time_offset,
                             time_keep, time_strict, FLB_FALSE, logfmt_no_bare_keys,
                             types, types_len, NULL, c->config);
}

/* the same with field decoders: `decoders` holds "property\tvalue\n" lines (Decode_Field / Decode_Field_As entries of a
 * [PARSER] section), handed to the reference's own flb_parser_decoder_list_create() through a section made of them */
#include <fluent-bit/flb_config_format.h>
#include <fluent-bit/flb_parser_decoder.h>
#include <cfl/cfl_kvlist.h>
void *flbref_parser_create_dec(void *cfg, const char *name, const char *format, const char *regex,
                               int skip_empty, const char *time_fmt, const char *time_key,
                               const char *time_offset, int time_keep, int time_strict,
                               int logfmt_no_bare_keys, const char *types_spec, const char *decoders)
{
    struct flbref_cfg *c = cfg;
    struct flb_parser_types *types = NULL;
    struct mk_list *dec_list = NULL;
    int types_len = 0;

    if (types_spec && *types_spec) {
        const char *p = types_spec;
        types = calloc(64, sizeof(*types));
        while (*p && types_len < 63) {
            const char *e, *colon;
            while (*p == ' ') p++;
            if (!*p) break;
            e = p;
            while (*e && *e != ' ') e++;
            colon = memchr(p, ':', e - p);
            if (colon) {
                types[types_len].key = strndup(p, colon - p);
                types[types_len].key_len = colon - p;
                types[types_len].type = type_from_name(colon + 1, e - colon - 1);
                types_len++;
            }
            p = e;
        }
    }
    if (decoders && *decoders) {
        struct flb_cf_section sec;
        const char *p = decoders;
        memset(&sec, 0, sizeof(sec));
        sec.properties = cfl_kvlist_create();
        while (*p) {
            const char *tab = strchr(p, '\t'), *nl = strchr(p, '\n');
            char *k, *v;
            if (!tab || !nl || tab > nl) break;
            k = strndup(p, tab - p); v = strndup(tab + 1, nl - tab - 1);
            cfl_kvlist_insert_string(sec.properties, k, v);
            free(k); free(v);
            p = nl + 1;
        }
        dec_list = flb_parser_decoder_list_create(&sec);
        cfl_kvlist_destroy(sec.properties);
        if (!dec_list) return NULL;
    }
    return flb_parser_create(name, format, regex, skip_empty, time_fmt, time_key, time_offset,
                             time_keep, time_strict, FLB_FALSE, logfmt_no_bare_keys,
                             types, types_len, dec_list, c->config);
}

int flbref_parser_do(void *parser, const char *buf, size_t len, void **out_buf, size_t *out_size,
                     long long *sec, long long *nsec)
{
    struct flb_time t;
    int ret;

    flb_time_zero(&t);
    *out_buf = NULL;
    *out_size = 0;
    ret = flb_parser_do(parser, buf, len, out_buf, out_size, &t);
    *sec = (long long) t.tm.tv_sec;
    *nsec = (long long) t.tm.tv_nsec;
    return ret;
}

int flbref_time_lookup(void *parser, const char *str, size_t len, long long now,
                       long long *sec, double *ns)
{
    struct flb_tm tm;
    int ret;

    memset(&tm, 0, sizeof(tm));
    *ns = 0;
    ret = flb_parser_time_lookup(str, len, (time_t) now, parser, &tm, ns);
    if (ret == 0) {
        *sec = (long long) flb_parser_tm2time(&tm, FLB_FALSE);
    }
    return ret;
}

void flbref_free(void *p) { flb_free(p); }

/* ---- regex ---- */
void *flbref_regex_create(const char *pattern) { return flb_regex_create(pattern); }
void flbref_regex_destroy(void *re) { flb_regex_destroy(re); }

struct names_ctx { char *buf; size_t cap; size_t len; int n; };
static void names_cb(const char *name, const char *value, size_t vlen, void *data)
{
    struct names_ctx *nc = data;
    size_t l = strlen(name);
    if (nc->len + l + 1 < nc->cap) {
        memcpy(nc->buf + nc->len, name, l);
        nc->len += l;
        nc->buf[nc->len++] = '\n';
    }
    nc->n++;
}

/* Leftmost search over [str,str+len).  Returns flb_regex_do()'s value (number of
 * named groups, or -1).  beg/end get the byte offsets of group 0..maxregs-1
 * (-1 when unset); *nregs the region size. */
int flbref_regex_search(void *re, const char *str, size_t len, int *beg, int *end, int maxregs,
                        int *nregs)
{
    struct flb_regex_search res;
    int ret, i;

    memset(&res, 0, sizeof(res));
    *nregs = 0;
    ret = flb_regex_do(re, str, len, &res);
    if (ret <= 0) {            /* 0: matched, but no groups -> region already freed */
        return ret;
    }
    *nregs = ((OnigRegion *) res.region)->num_regs;
    for (i = 0; i < ((OnigRegion *) res.region)->num_regs && i < maxregs; i++) {
        beg[i] = ((OnigRegion *) res.region)->beg[i];
        end[i] = ((OnigRegion *) res.region)->end[i];
    }
    /* flb_regex_parse frees the region */
    {
        char tmp[4]; struct names_ctx nc = { tmp, 0, 0, 0 };
        flb_regex_parse(re, &res, names_cb, &nc);
    }
    return ret;
}

/* names of the named groups in the order flb_regex_parse reports them, '\n' separated */
int flbref_regex_names(void *re, const char *str, size_t len, char *out, size_t cap)
{
    struct flb_regex_search res;
    struct names_ctx nc = { out, cap, 0, 0 };
    int ret;

    memset(&res, 0, sizeof(res));
    ret = flb_regex_do(re, str, len, &res);
    if (ret <= 0) return ret;
    flb_regex_parse(re, &res, names_cb, &nc);
    if (nc.len < cap) out[nc.len] = '\0';
    return nc.n;
}

int flbref_regex_match(void *re, const char *str, size_t len)
{
    return flb_regex_match(re, (unsigned char *) str, len);
}

/* ---- JSON -> msgpack ---- */
int flbref_pack_json(const char *js, size_t len, void **out, size_t *out_size)
{
    int root_type;
    char *buf = NULL;
    int ret;
    size_t consumed = 0;

    ret = flb_pack_json(js, len, &buf, out_size, &root_type, &consumed);
    *out = buf;
    return ret;
}

/* flb_pack_json_state() on a fresh state (src/flb_pack.c:758): returns its value; *last_byte / *tokens_count from the state */
int flbref_pack_json_state(const char *js, size_t len, void **out, int *out_size, int *last_byte, int *tokens_count)
{
    struct flb_pack_state st;
    char *buf = NULL;
    int ret, size = 0;

    *out = NULL; *out_size = 0;
    flb_pack_state_init(&st);
    ret = flb_pack_json_state(js, len, &buf, &size, &st);
    *last_byte = st.last_byte;
    *tokens_count = st.tokens_count;
    if (ret == 0) { *out = buf; *out_size = size; }
    flb_pack_state_reset(&st);
    return ret;
}

/* ---- filters ---- */
void *flbref_filter_create(void *cfg, const char *plugin)
{
    struct flbref_cfg *c = cfg;
    struct flb_filter_instance *ins = flb_filter_new(c->config, plugin, NULL);
    if (ins) {
        flb_filter_set_property(ins, "match", "*");
    }
    return ins;
}

int flbref_filter_set(void *filter, const char *k, const char *v)
{
    return flb_filter_set_property(filter, k, v);
}

int flbref_filter_init(void *cfg, void *filter)
{
    struct flbref_cfg *c = cfg;
    return flb_filter_init(c->config, filter);
}

/* one plugin callback; returns FLB_FILTER_MODIFIED(1) / FLB_FILTER_NOTOUCH(2) */
int flbref_filter_cb(void *cfg, void *filter, const void *data, size_t bytes, const char *tag,
                     void **out, size_t *out_size)
{
    struct flbref_cfg *c = cfg;
    struct flb_filter_instance *ins = filter;

    *out = NULL;
    *out_size = 0;
    return ins->p->cb_filter(data, bytes, tag, (int) strlen(tag), out, out_size, ins, &c->in,
                             ins->context, c->config);
}

/* the whole configured chain through flb_filter_do() (src/flb_filter.c:119).
 * returns 0 when the result is the caller's buffer (untouched), 1 when *out is a
 * new heap buffer (free with flbref_free), and *out_size==0 when all dropped. */
int flbref_filter_do(void *cfg, const void *data, size_t bytes, int records, const char *tag,
                     void **out, size_t *out_size)
{
    struct flbref_cfg *c = cfg;
    struct flb_input_chunk ic;

    memset(&ic, 0, sizeof(ic));
    ic.in = &c->in;
    ic.added_records = records;
    ic.total_records = records;
    flb_filter_do(&ic, data, bytes, out, out_size, tag, (int) strlen(tag), c->config);
    if (*out == data) {
        return 0;
    }
    return 1;
}

int flbref_count_records(const void *buf, size_t size)
{
    return flb_mp_count_log_records(buf, size);
}

/* text dump of a filter instance's framework counters (src/flb_filter.c:574-616) */
char *flbref_filter_cmt_text(void *filter)
{
    struct flb_filter_instance *ins = filter;
    cfl_sds_t t = cmt_encode_text_create(ins->cmt);
    char *r = strdup(t ? t : "");
    if (t) cmt_encode_text_destroy(t);
    return r;
}

/* text dump of filter_log_to_metrics' own cmetrics context */
char *flbref_l2m_cmt_text(void *filter)
{
    struct flb_filter_instance *ins = filter;
    struct log_to_metrics_ctx *ctx = ins->context;
    cfl_sds_t t = cmt_encode_text_create(ctx->cmt);
    char *r = strdup(t ? t : "");
    if (t) cmt_encode_text_destroy(t);
    return r;
}

int flbref_l2m_cmt_msgpack(void *filter, void **out, size_t *out_size)
{
    struct flb_filter_instance *ins = filter;
    struct log_to_metrics_ctx *ctx = ins->context;
    char *b = NULL;
    int ret = cmt_encode_msgpack_create(ctx->cmt, &b, out_size);
    if (ret == 0) {
        *out = malloc(*out_size);
        memcpy(*out, b, *out_size);
        cmt_encode_msgpack_destroy(b);
    }
    return ret;
}

void flbref_cfree(void *p) { free(p); }

/* ---- output side: flb_pack_msgpack_to_json_format() (src/flb_pack.c:1320-1602), what out_stdout / out_http / out_file ...
 * call on a chunk.  Returns a malloc()ed copy of the sds (free with flbref_cfree) or NULL. */
char *flbref_to_json_format(const void *data, size_t bytes, int json_format, int date_format, const char *date_key,
                            int escape_unicode, size_t *out_len)
{
    flb_sds_t key = date_key ? flb_sds_create(date_key) : NULL;
    flb_sds_t js = flb_pack_msgpack_to_json_format(data, bytes, json_format, date_format, key, escape_unicode);
    char *r = NULL;
    *out_len = 0;
    if (js) {
        *out_len = flb_sds_len(js);
        r = malloc(*out_len + 1);
        memcpy(r, js, *out_len);
        r[*out_len] = 0;
        flb_sds_destroy(js);
    }
    if (key) flb_sds_destroy(key);
    return r;
}

/* ---- ingest side: the line loop of in_tail (plugins/in_tail/tail_file.c process_content() :629-700 + go_next, and
 * flb_tail_file_pack_line() :338-391) without a file behind it.  The loop -- cut at '\n', the skip_empty_lines rule, the
 * trailing '\r' of lines of two bytes and more -- is RESTATED here line for line (the original is a static function over
 * struct flb_tail_file); every event is encoded by the reference's own flb_log_event_encoder calls, the ones pack_line makes,
 * with the timestamp given instead of "now".  *consumed = processed_bytes.  Returns a malloc()ed chunk (flbref_cfree). */
#include <fluent-bit/flb_log_event_encoder.h>
char *flbref_lines_to_events(const char *text, size_t bytes, const char *key, int skip_empty_lines, long long sec, long long nsec,
                             const char *path_key, const char *path, const char *offset_key, unsigned long long stream_offset,
                             size_t *out_len, size_t *consumed, size_t *lines)
{
    struct flb_log_event_encoder *enc = flb_log_event_encoder_create(FLB_LOG_EVENT_FORMAT_DEFAULT);
    const char *data = text, *end = text + bytes, *p;
    size_t processed_bytes = 0, n_lines = 0;
    struct flb_time tm;
    char *out;

    tm.tm.tv_sec = (time_t) sec; tm.tm.tv_nsec = (long) nsec;
    while (data < end && (p = memchr(data, '\n', end - data))) {
        size_t len = (size_t) (p - data);
        int crlf = 0, result;
        if (skip_empty_lines) {
            if (len == 0) { data++; processed_bytes++; continue; }
            else if (len == 1 && data[0] == '\r') { data += 2; processed_bytes += 2; continue; }
        }
        if (len >= 2) crlf = (data[len - 1] == '\r');
        result = flb_log_event_encoder_begin_record(enc);
        if (result == FLB_EVENT_ENCODER_SUCCESS) result = flb_log_event_encoder_set_timestamp(enc, &tm);
        if (path_key && result == FLB_EVENT_ENCODER_SUCCESS)
            result = flb_log_event_encoder_append_body_values(enc, FLB_LOG_EVENT_CSTRING_VALUE(path_key),
                                                              FLB_LOG_EVENT_STRING_VALUE(path, strlen(path)));
        if (offset_key && result == FLB_EVENT_ENCODER_SUCCESS)
            result = flb_log_event_encoder_append_body_values(enc, FLB_LOG_EVENT_CSTRING_VALUE(offset_key),
                                                              FLB_LOG_EVENT_UINT64_VALUE(stream_offset + processed_bytes));
        if (result == FLB_EVENT_ENCODER_SUCCESS)
            result = flb_log_event_encoder_append_body_values(enc, FLB_LOG_EVENT_CSTRING_VALUE(key),
                                                              FLB_LOG_EVENT_STRING_VALUE(data, len - crlf));
        if (result == FLB_EVENT_ENCODER_SUCCESS) result = flb_log_event_encoder_commit_record(enc);
        n_lines++;
        data += len + 1;
        processed_bytes += len + 1;
    }
    *out_len = enc->output_length;
    out = malloc(enc->output_length + 1);
    memcpy(out, enc->output_buffer, enc->output_length);
    *consumed = processed_bytes; *lines = n_lines;
    flb_log_event_encoder_destroy(enc);
    return out;
}

/* ---- multiline parser definitions: what a [MULTILINE_PARSER] section becomes (src/flb_parser.c:815-935):
 * flb_ml_parser_create(), one flb_ml_rule_create() per `rule`, flb_ml_parser_init() ---- */
void *flbref_ml_parser_create(void *cfg, const char *name, const char *type, const char *match_string, int negate,
                              int flush_ms, const char *key_content, const char *key_group, const char *key_pattern,
                              const char *parser_name)
{
    struct flbref_cfg *c = cfg;
    struct flb_parser *pctx = parser_name ? flb_parser_get(parser_name, c->config) : NULL;
    int t = flb_ml_type_lookup((char *) type);
    if (t == -1) return NULL;
    return flb_ml_parser_create(c->config, (char *) name, t, (char *) match_string, negate, flush_ms, (char *) key_content,
                                (char *) key_group, (char *) key_pattern, pctx, (char *) parser_name);
}
int flbref_ml_parser_rule(void *mlp, const char *from_states, const char *regex, const char *to_state)
{
    return flb_ml_rule_create(mlp, (char *) from_states, (char *) regex, (char *) to_state, NULL);
}
int flbref_ml_parser_init(void *mlp) { return flb_ml_parser_init(mlp); }
void flbref_set_ml_buffer_limit(void *cfg, const char *limit)
{
    struct flbref_cfg *c = cfg;
    flb_free(c->config->multiline_buffer_limit);
    c->config->multiline_buffer_limit = flb_strdup(limit);
}

/* ---- filter_rewrite_tag's emitter: in_emitter_add_record() (plugins/in_emitter/emitter.c:124) is the one function of the
 * emitter input the filter calls per re-tagged record.  Here it appends (tag, record bytes) to a log the tests read:
 * u32 tag_len, u32 size, tag, bytes -- one entry per call, in call order. ---- */
static char *g_emit_log;
static size_t g_emit_len, g_emit_cap;
static int g_emit_fail_after = -1;       /* >= 0: calls from this one on return -1 (a paused / full emitter) */
static int g_emit_calls;
int in_emitter_add_record(const char *tag, int tag_len, const char *buf_data, size_t buf_size,
                          struct flb_input_instance *in, struct flb_input_instance *i_ins)
{
    uint32_t h[2];
    size_t need = g_emit_len + 8 + (size_t) tag_len + buf_size;
    if (g_emit_fail_after >= 0 && g_emit_calls++ >= g_emit_fail_after) return -1;
    if (need > g_emit_cap) { g_emit_cap = need * 2 + 4096; g_emit_log = realloc(g_emit_log, g_emit_cap); }
    h[0] = (uint32_t) tag_len; h[1] = (uint32_t) buf_size;
    memcpy(g_emit_log + g_emit_len, h, 8);
    memcpy(g_emit_log + g_emit_len + 8, tag, tag_len);
    memcpy(g_emit_log + g_emit_len + 8 + tag_len, buf_data, buf_size);
    g_emit_len = need;
    return 0;
}
int in_emitter_get_collector_id(struct flb_input_instance *in) { return 0; }
void flbref_emit_log(const void **log, size_t *len) { *log = g_emit_log; *len = g_emit_len; }
void flbref_emit_reset(int fail_after) { g_emit_len = 0; g_emit_calls = 0; g_emit_fail_after = fail_after; }
