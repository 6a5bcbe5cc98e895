/*
 * oracle/ -- CPU restatement of Fluent Bit's parse->filter path.  TEST INFRASTRUCTURE ONLY.
 *
 * Plain C, written from the reference's behaviour (every function cites the reference file:line it
 * follows; paths are relative to the fluent-bit source tree).  It is an independent second
 * implementation: it shares no code with the product under fluent-bit_b200/csrc (different regex
 * engine design -- recursive backtracking over a split/jump program --, tree-based msgpack, scalar
 * loops) so that agreement between the two means something.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load liboracle.so.  The product never does and has no CPU fallback.
 *
 * Parity of this restatement is PINNED: tests/test_oracle.py checks it against the committed golden
 * vectors (tests/golden/ JSON files, produced by the unmodified reference build oracle/_ref) and, when
 * oracle/_ref is present, live against the reference on the same seeded inputs.
 */
#ifndef ORC_H
#define ORC_H

#include <stddef.h>
#include <stdint.h>

/* ---- arena ------------------------------------------------------------------------ */
struct orc_arena { struct orc_chunk *head; };
void *orc_alloc(struct orc_arena *a, size_t n);
void  orc_arena_free(struct orc_arena *a);

/* ---- growing byte buffer ------------------------------------------------------------ */
struct orc_buf { uint8_t *p; size_t n, cap; };
void orc_buf_put(struct orc_buf *b, const void *p, size_t n);
void orc_buf_u8(struct orc_buf *b, unsigned v);

/* ---- msgpack values (lib/msgpack-c/include/msgpack/object.h) ------------------------ */
enum { OV_NIL, OV_BOOL, OV_UINT, OV_INT, OV_F32, OV_F64, OV_STR, OV_BIN, OV_EXT, OV_ARR, OV_MAP };
struct ov {
    int type;
    uint64_t u;              /* OV_UINT; OV_BOOL 0/1 */
    int64_t i;               /* OV_INT (always negative, as msgpack-c classifies) */
    double d;                /* OV_F32 / OV_F64 */
    const uint8_t *p;        /* STR/BIN/EXT payload */
    uint32_t len;
    int8_t ext;
    struct ov *items;        /* ARR: n items; MAP: 2n items (k0 v0 k1 v1 ...) */
    uint32_t n;
};
/* 0 = one object decoded (*off advanced), 1 = buffer ends inside the object, -1 = invalid */
int  ov_unpack(struct orc_arena *a, const uint8_t *buf, size_t len, size_t *off, struct ov *out);
void ov_pack(struct orc_buf *b, const struct ov *v);           /* msgpack_pack_object() */
void ov_pack_str(struct orc_buf *b, const void *p, size_t n);
void ov_pack_map_hdr(struct orc_buf *b, uint32_t n);
void ov_pack_arr_hdr(struct orc_buf *b, uint32_t n);
void ov_pack_uint(struct orc_buf *b, uint64_t v);
void ov_pack_int(struct orc_buf *b, int64_t v);
void ov_pack_double(struct orc_buf *b, double d);
struct ov ov_str(const void *p, size_t n);
int ov_str_eq(const struct ov *v, const char *s, size_t n);

/* ---- regex (lib/onigmo, Ruby syntax, UTF-8) ------------------------------------------ */
struct orc_regex;
struct orc_regex *orc_regex_create(const char *pattern, char *err, size_t errlen);
void orc_regex_destroy(struct orc_regex *r);
/* leftmost match: 1 + region[2*g], region[2*g+1] byte offsets (-1 unset); 0 no match */
int orc_regex_search(const struct orc_regex *r, const uint8_t *s, size_t n, int *region, int max_groups);
int orc_regex_ngroups(const struct orc_regex *r);              /* capture groups that report (incl. 0) */
int orc_regex_nnames(const struct orc_regex *r);
const char *orc_regex_name(const struct orc_regex *r, int i, int *group);   /* definition order */

/* ---- time (src/flb_strptime.c, src/flb_parser.c:1159-1278) --------------------------- */
struct orc_tm { int sec, min, hour, mday, mon, year, wday, yday, isdst; long gmtoff; };
/* returns consumed length or -1 */
int orc_strptime(const char *buf, const char *fmt, struct orc_tm *tm, int64_t now_year);
int64_t orc_timegm(const struct orc_tm *tm);

#endif
