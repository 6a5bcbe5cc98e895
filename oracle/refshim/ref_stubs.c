/* Stubs for the engine/control-plane symbols the parse->filter sources of the
 * reference pull in but never need on this path (logging sink, worker TLS,
 * scheduler timers, config-file loader, multiline, hidden emitter input).
 * Test infrastructure only (part of oracle/_ref); nothing here is product code. */
#include <stdio.h>
#include <stdlib.h>
#include <stdarg.h>
#include <string.h>
#include <errno.h>
#include <fluent-bit/flb_info.h>
#include <fluent-bit/flb_config.h>
#include <fluent-bit/flb_input.h>
#include <fluent-bit/flb_filter.h>
#include <fluent-bit/flb_worker.h>
#include <fluent-bit/flb_log.h>
#include <fluent-bit/flb_metrics.h>
#include <fluent-bit/flb_scheduler.h>

FLB_TLS_DEFINE(struct flb_worker, flb_worker_ctx);

static int ref_verbose(void)
{
    static int v = -1;
    if (v < 0) v = getenv("FLBREF_VERBOSE") ? 1 : 0;
    return v;
}

void flb_log_print(int type, const char *file, int line, const char *fmt, ...)
{
    va_list ap;
    if (!ref_verbose()) return;
    va_start(ap, fmt);
    fprintf(stderr, "[ref:%d] ", type);
    vfprintf(stderr, fmt, ap);
    fputc('\n', stderr);
    va_end(ap);
}
int flb_errno_print(int errnum, const char *file, int line)
{
    if (ref_verbose()) fprintf(stderr, "[ref] errno=%d at %s:%d\n", errnum, file ? file : "?", line);
    return 0;
}
int flb_worker_log_level(struct flb_worker *worker) { return 3; }
struct flb_worker *flb_worker_get(void) { return NULL; }
int flb_log_cache_check_suppress(struct flb_log_cache *cache, char *msg_buf, size_t msg_size) { return 0; }
int flb_log_get_level_str(char *str) { return -1; }

/* legacy metrics API: counters only, kept so flb_filter_init()/flb_filter_do() run */
struct flb_metrics *flb_metrics_create(const char *title)
{
    struct flb_metrics *m = calloc(1, sizeof(*m));
    if (m) mk_list_init(&m->list);
    return m;
}
int flb_metrics_add(int id, const char *title, struct flb_metrics *metrics) { return id; }
int flb_metrics_sum(int id, size_t val, struct flb_metrics *metrics) { return 0; }
int flb_metrics_destroy(struct flb_metrics *metrics) { free(metrics); return 0; }

/* tag routing: the harness always passes a matching tag */


/* --- log_to_metrics' hidden emitter input: recorded, not run --- */
struct flb_input_instance *flb_input_new(struct flb_config *config, const char *input, void *data, int public_only)
{
    struct flb_input_instance *ins = calloc(1, sizeof(*ins));
    if (ins) { mk_list_init(&ins->properties); ins->config = config; }
    return ins;
}
int flb_input_name_exists(const char *name, struct flb_config *config) { return 0; }
int flb_input_set_property(struct flb_input_instance *ins, const char *k, const char *v) { return 0; }
int flb_input_instance_init(struct flb_input_instance *ins, struct flb_config *config) { return 0; }
int flb_storage_input_create(void *cio, struct flb_input_instance *in) { return 0; }
static unsigned long g_metrics_appends;
int flb_input_metrics_append(struct flb_input_instance *ins, const char *tag, size_t tag_len, struct cmt *cmt)
{
    g_metrics_appends++;
    return 0;
}
unsigned long ref_metrics_appends(void) { return g_metrics_appends; }
struct flb_sched *flb_sched_ctx_get(void) { return NULL; }
int flb_sched_timer_cb_create(struct flb_sched *sched, int type, int ms, void (*cb)(struct flb_config *, void *),
                              void *data, struct flb_sched_timer **out_timer) { if (out_timer) *out_timer = NULL; return 0; }
int flb_sched_timer_cb_disable(struct flb_sched_timer *timer) { return 0; }
int flb_sched_timer_destroy(struct flb_sched_timer *timer) { return 0; }
