/* oracle: shared declarations of the fluent-bit level restatement.  TEST INFRASTRUCTURE (see orc.h). */
#ifndef ORC_FLB_H
#define ORC_FLB_H

#include <strings.h>
#include <time.h>
#include "orc.h"

enum { ORC_P_REGEX = 1, ORC_P_JSON, ORC_P_LTSV, ORC_P_LOGFMT };
enum { ORC_T_STRING = 0, ORC_T_INT, ORC_T_FLOAT, ORC_T_BOOL, ORC_T_HEX };

struct orc_parser {
    struct orc_parser *next;
    char name[64];
    int type;
    struct orc_regex *regex;
    int skip_empty, time_keep, time_strict, logfmt_no_bare_keys;
    char *time_fmt, *time_fmt_year, *time_key;
    const char *time_frac_secs;
    int with_year, with_tz, time_offset;
    struct { char *key; int type; } types[32];
    int n_types;
};

struct orc_config { struct orc_parser *parsers; struct orc_filter *filters, *filters_tail; };

extern time_t orc_now;       /* 0 = wall clock; tests pin it for formats without a year */

struct orc_parser *orc_parser_create(struct orc_config *cfg, const char *name, const char *format, const char *regex,
                                     int skip_empty, const char *time_fmt, const char *time_key, const char *time_offset,
                                     int time_keep, int time_strict, int logfmt_no_bare_keys, const char *types_spec);
struct orc_parser *orc_parser_get(struct orc_config *cfg, const char *name);
/* >= 0 parsed (msgpack map appended to out), -1 not parsed */
int orc_parser_do(const struct orc_parser *p, const char *buf, size_t length, struct orc_buf *out, int64_t *sec, int64_t *nsec);
int orc_time_lookup(const struct orc_parser *parser, const char *time_str, size_t tsize, struct orc_tm *tm, double *ns);

#endif
