/* oracle: msgpack object tree.  TEST INFRASTRUCTURE (see orc.h).
 * Follows lib/msgpack-c: unpack_template.h (type classification, 32-level container stack,
 * 0xc1 is invalid) and src/objectc.c:msgpack_pack_object (canonical, smallest encodings). */
#include <stdlib.h>
#include <string.h>
#include "orc.h"

struct orc_chunk { struct orc_chunk *next; size_t used, cap; unsigned char mem[]; };

void *orc_alloc(struct orc_arena *a, size_t n)
{
    struct orc_chunk *c = a->head;
    void *p;
    n = (n + 15) & ~(size_t) 15;
    if (!c || c->used + n > c->cap) {
        size_t cap = n > 65536 ? n : 65536;
        c = malloc(sizeof(*c) + cap);
        c->next = a->head; c->used = 0; c->cap = cap;
        a->head = c;
    }
    p = c->mem + c->used;
    c->used += n;
    return p;
}

void orc_arena_free(struct orc_arena *a)
{
    while (a->head) { struct orc_chunk *n = a->head->next; free(a->head); a->head = n; }
}

void orc_buf_put(struct orc_buf *b, const void *p, size_t n)
{
    if (b->n + n > b->cap) {
        while (b->n + n > b->cap) b->cap = b->cap ? b->cap * 2 : 4096;
        b->p = realloc(b->p, b->cap);
    }
    if (n) memcpy(b->p + b->n, p, n);
    b->n += n;
}

void orc_buf_u8(struct orc_buf *b, unsigned v) { uint8_t c = (uint8_t) v; orc_buf_put(b, &c, 1); }

static uint64_t be(const uint8_t *p, int n) { uint64_t v = 0; while (n--) v = (v << 8) | *p++; return v; }

static void put_be(struct orc_buf *b, uint64_t v, int n)
{
    uint8_t t[8];
    int i;
    for (i = 0; i < n; i++) t[i] = (uint8_t) (v >> (8 * (n - 1 - i)));
    orc_buf_put(b, t, (size_t) n);
}

/* lib/msgpack-c/include/msgpack/unpack_template.h: one object at buf[*off] */
static int unpack_at(struct orc_arena *a, const uint8_t *buf, size_t len, size_t *off, struct ov *o, int depth)
{
    size_t i = *off;
    unsigned c;
    uint32_t n = 0, k;
    int hdr = 0, is_map = 0;
    memset(o, 0, sizeof(*o));
    if (i >= len) return 1;
    c = buf[i];
#define NEED(x) do { if (i + (x) > len) return 1; } while (0)
    if (c <= 0x7f) { o->type = OV_UINT; o->u = c; *off = i + 1; return 0; }
    if (c >= 0xe0) { o->type = OV_INT; o->i = (int8_t) c; *off = i + 1; return 0; }
    if (c >= 0xa0 && c <= 0xbf) { n = c & 31; hdr = 1; goto str; }
    if (c >= 0x90 && c <= 0x9f) { n = c & 15; hdr = 1; goto container; }
    if (c >= 0x80 && c <= 0x8f) { n = c & 15; hdr = 1; is_map = 1; goto container; }
    switch (c) {
    case 0xc0: o->type = OV_NIL; *off = i + 1; return 0;
    case 0xc2: o->type = OV_BOOL; o->u = 0; *off = i + 1; return 0;
    case 0xc3: o->type = OV_BOOL; o->u = 1; *off = i + 1; return 0;
    case 0xcc: NEED(2); o->type = OV_UINT; o->u = buf[i + 1]; *off = i + 2; return 0;
    case 0xcd: NEED(3); o->type = OV_UINT; o->u = be(buf + i + 1, 2); *off = i + 3; return 0;
    case 0xce: NEED(5); o->type = OV_UINT; o->u = be(buf + i + 1, 4); *off = i + 5; return 0;
    case 0xcf: NEED(9); o->type = OV_UINT; o->u = be(buf + i + 1, 8); *off = i + 9; return 0;
    case 0xd0: case 0xd1: case 0xd2: case 0xd3: {
        int w = 1 << (c - 0xd0);
        int64_t v;
        NEED(1 + w);
        v = (int64_t) be(buf + i + 1, w);
        if (w < 8) { int sh = 64 - 8 * w; v = (int64_t) ((uint64_t) v << sh) >> sh; }
        /* a non-negative value in a signed encoding is a POSITIVE_INTEGER */
        if (v >= 0) { o->type = OV_UINT; o->u = (uint64_t) v; } else { o->type = OV_INT; o->i = v; }
        *off = i + 1 + w; return 0;
    }
    case 0xca: { union { uint32_t u; float f; } cv; NEED(5); cv.u = (uint32_t) be(buf + i + 1, 4); o->type = OV_F32; o->d = cv.f; *off = i + 5; return 0; }
    case 0xcb: { union { uint64_t u; double f; } cv; NEED(9); cv.u = be(buf + i + 1, 8); o->type = OV_F64; o->d = cv.f; *off = i + 9; return 0; }
    case 0xd9: NEED(2); n = buf[i + 1]; hdr = 2; goto str;
    case 0xda: NEED(3); n = (uint32_t) be(buf + i + 1, 2); hdr = 3; goto str;
    case 0xdb: NEED(5); n = (uint32_t) be(buf + i + 1, 4); hdr = 5; goto str;
    case 0xc4: NEED(2); n = buf[i + 1]; hdr = 2; goto bin;
    case 0xc5: NEED(3); n = (uint32_t) be(buf + i + 1, 2); hdr = 3; goto bin;
    case 0xc6: NEED(5); n = (uint32_t) be(buf + i + 1, 4); hdr = 5; goto bin;
    case 0xd4: case 0xd5: case 0xd6: case 0xd7: case 0xd8:
        n = 1u << (c - 0xd4); NEED(2 + (size_t) n);
        o->type = OV_EXT; o->ext = (int8_t) buf[i + 1]; o->p = buf + i + 2; o->len = n; *off = i + 2 + n; return 0;
    case 0xc7: NEED(3); n = buf[i + 1]; hdr = 2; goto ext;
    case 0xc8: NEED(4); n = (uint32_t) be(buf + i + 1, 2); hdr = 3; goto ext;
    case 0xc9: NEED(6); n = (uint32_t) be(buf + i + 1, 4); hdr = 5; goto ext;
    case 0xdc: NEED(3); n = (uint32_t) be(buf + i + 1, 2); hdr = 3; goto container;
    case 0xdd: NEED(5); n = (uint32_t) be(buf + i + 1, 4); hdr = 5; goto container;
    case 0xde: NEED(3); n = (uint32_t) be(buf + i + 1, 2); hdr = 3; is_map = 1; goto container;
    case 0xdf: NEED(5); n = (uint32_t) be(buf + i + 1, 4); hdr = 5; is_map = 1; goto container;
    default: return -1;                                  /* 0xc1 */
    }
str:
    NEED((size_t) hdr + n);
    o->type = OV_STR; o->p = buf + i + hdr; o->len = n; *off = i + hdr + n; return 0;
bin:
    NEED((size_t) hdr + n);
    o->type = OV_BIN; o->p = buf + i + hdr; o->len = n; *off = i + hdr + n; return 0;
ext:
    NEED((size_t) hdr + 1 + n);
    o->type = OV_EXT; o->ext = (int8_t) buf[i + hdr]; o->p = buf + i + hdr + 1; o->len = n; *off = i + hdr + 1 + n; return 0;
container:
    o->type = is_map ? OV_MAP : OV_ARR;
    o->n = n;
    i += (size_t) hdr;
    if (n == 0) { *off = i; return 0; }
    if (depth >= 32) return -1;                          /* MSGPACK_EMBED_STACK_SIZE */
    {
        uint64_t total = (uint64_t) n * (is_map ? 2 : 1);
        if (total > len - i) return 1;                   /* every element takes >= 1 byte */
        o->items = orc_alloc(a, sizeof(struct ov) * (size_t) total);
        for (k = 0; k < total; k++) {
            int r = unpack_at(a, buf, len, &i, &o->items[k], depth + 1);
            if (r) return r;
        }
    }
    *off = i;
    return 0;
#undef NEED
}

int ov_unpack(struct orc_arena *a, const uint8_t *buf, size_t len, size_t *off, struct ov *out)
{
    size_t o = *off;
    int r = unpack_at(a, buf, len, &o, out, 0);
    if (r == 0) *off = o;
    return r;
}

/* lib/msgpack-c/include/msgpack/pack_template.h: msgpack_pack_uint64 / int64 */
void ov_pack_uint(struct orc_buf *b, uint64_t v)
{
    if (v < 128) orc_buf_u8(b, (unsigned) v);
    else if (v < 256) { orc_buf_u8(b, 0xcc); put_be(b, v, 1); }
    else if (v < 65536) { orc_buf_u8(b, 0xcd); put_be(b, v, 2); }
    else if (v < 4294967296ull) { orc_buf_u8(b, 0xce); put_be(b, v, 4); }
    else { orc_buf_u8(b, 0xcf); put_be(b, v, 8); }
}

void ov_pack_int(struct orc_buf *b, int64_t v)
{
    if (v >= 0) { ov_pack_uint(b, (uint64_t) v); return; }
    if (v >= -32) orc_buf_u8(b, (unsigned) (uint8_t) v);
    else if (v >= -128) { orc_buf_u8(b, 0xd0); put_be(b, (uint64_t) v, 1); }
    else if (v >= -32768) { orc_buf_u8(b, 0xd1); put_be(b, (uint64_t) v, 2); }
    else if (v >= -2147483648ll) { orc_buf_u8(b, 0xd2); put_be(b, (uint64_t) v, 4); }
    else { orc_buf_u8(b, 0xd3); put_be(b, (uint64_t) v, 8); }
}

void ov_pack_double(struct orc_buf *b, double d)
{
    union { double f; uint64_t u; } cv;
    cv.f = d;
    orc_buf_u8(b, 0xcb); put_be(b, cv.u, 8);
}

void ov_pack_str(struct orc_buf *b, const void *p, size_t n)
{
    if (n < 32) orc_buf_u8(b, 0xa0 | (unsigned) n);
    else if (n < 256) { orc_buf_u8(b, 0xd9); put_be(b, n, 1); }
    else if (n < 65536) { orc_buf_u8(b, 0xda); put_be(b, n, 2); }
    else { orc_buf_u8(b, 0xdb); put_be(b, n, 4); }
    orc_buf_put(b, p, n);
}

void ov_pack_map_hdr(struct orc_buf *b, uint32_t n)
{
    if (n < 16) orc_buf_u8(b, 0x80 | n);
    else if (n < 65536) { orc_buf_u8(b, 0xde); put_be(b, n, 2); }
    else { orc_buf_u8(b, 0xdf); put_be(b, n, 4); }
}

void ov_pack_arr_hdr(struct orc_buf *b, uint32_t n)
{
    if (n < 16) orc_buf_u8(b, 0x90 | n);
    else if (n < 65536) { orc_buf_u8(b, 0xdc); put_be(b, n, 2); }
    else { orc_buf_u8(b, 0xdd); put_be(b, n, 4); }
}

/* lib/msgpack-c/src/objectc.c: msgpack_pack_object */
void ov_pack(struct orc_buf *b, const struct ov *v)
{
    uint32_t k;
    switch (v->type) {
    case OV_NIL: orc_buf_u8(b, 0xc0); break;
    case OV_BOOL: orc_buf_u8(b, v->u ? 0xc3 : 0xc2); break;
    case OV_UINT: ov_pack_uint(b, v->u); break;
    case OV_INT: ov_pack_int(b, v->i); break;
    case OV_F32: { union { float f; uint32_t u; } cv; cv.f = (float) v->d; orc_buf_u8(b, 0xca); put_be(b, cv.u, 4); break; }
    case OV_F64: ov_pack_double(b, v->d); break;
    case OV_STR: ov_pack_str(b, v->p, v->len); break;
    case OV_BIN:
        if (v->len < 256) { orc_buf_u8(b, 0xc4); put_be(b, v->len, 1); }
        else if (v->len < 65536) { orc_buf_u8(b, 0xc5); put_be(b, v->len, 2); }
        else { orc_buf_u8(b, 0xc6); put_be(b, v->len, 4); }
        orc_buf_put(b, v->p, v->len);
        break;
    case OV_EXT:
        if (v->len == 1) orc_buf_u8(b, 0xd4);
        else if (v->len == 2) orc_buf_u8(b, 0xd5);
        else if (v->len == 4) orc_buf_u8(b, 0xd6);
        else if (v->len == 8) orc_buf_u8(b, 0xd7);
        else if (v->len == 16) orc_buf_u8(b, 0xd8);
        else if (v->len < 256) { orc_buf_u8(b, 0xc7); put_be(b, v->len, 1); }
        else if (v->len < 65536) { orc_buf_u8(b, 0xc8); put_be(b, v->len, 2); }
        else { orc_buf_u8(b, 0xc9); put_be(b, v->len, 4); }
        orc_buf_u8(b, (unsigned) (uint8_t) v->ext);
        orc_buf_put(b, v->p, v->len);
        break;
    case OV_ARR:
        ov_pack_arr_hdr(b, v->n);
        for (k = 0; k < v->n; k++) ov_pack(b, &v->items[k]);
        break;
    case OV_MAP:
        ov_pack_map_hdr(b, v->n);
        for (k = 0; k < 2 * v->n; k++) ov_pack(b, &v->items[k]);
        break;
    }
}

struct ov ov_str(const void *p, size_t n)
{
    struct ov v;
    memset(&v, 0, sizeof(v));
    v.type = OV_STR; v.p = p; v.len = (uint32_t) n;
    return v;
}

int ov_str_eq(const struct ov *v, const char *s, size_t n)
{
    return v->type == OV_STR && v->len == n && memcmp(v->p, s, n) == 0;
}
