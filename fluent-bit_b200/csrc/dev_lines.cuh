/* dev_lines.cuh -- the ingest side of the path: raw text to log events (row f4, input side).
 *
 * Reference: the line loop of in_tail, plugins/in_tail/tail_file.c process_content() :629-700 (cut at '\n'; with
 * skip_empty_lines an empty line and a lone "\r" are stepped over; a line of two bytes or more loses its trailing '\r') and
 * go_next :776-785 (processed_bytes), and flb_tail_file_pack_line() :338-391: one event per line --
 *     92 92 d7 00 <sec> <nsec> df 00000000 df <n> [path_key: path] [offset_key: stream_offset + processed_bytes] key: line
 * (the encoder's forced map32 headers for metadata and body, flb_mp.c:591-640).  Parsers, docker mode and the tail's own
 * multiline modes are other code paths of that function and not this one.
 *
 *   ln_count .... one lane per tile of LN_TILE bytes: line feeds in the tile
 *   ln_fill ..... the positions of the line feeds, at the tile's offset in the list
 *   ln_size ..... one lane per line: where it starts, what it keeps, the bytes of its event (0: stepped over)
 *   ln_emit ..... one lane per line: the event
 * Functions of a "thread index" like the rest, so that the CPU emulation of the tests runs the same code.
 */
#ifndef FLBGPU_DEV_LINES_CUH
#define FLBGPU_DEV_LINES_CUH
#include "dev_msgpack.cuh"

/* 0x80 in every byte of x that is a line feed (exact: no borrow between bytes) */
FLB_HD uint64_t ln_nl_mask(uint64_t x)
{
    const uint64_t m = x ^ 0x0a0a0a0a0a0a0a0aull, k = 0x7f7f7f7f7f7f7f7full;
    return ~(((m & k) + k) | m | k);
}
FLB_HD int ln_popc64(uint64_t v)
{
#ifdef __CUDA_ARCH__
    return __popcll(v);
#else
    return __builtin_popcountll(v);
#endif
}
FLB_HD int ln_ctz64(uint64_t v)
{
#ifdef __CUDA_ARCH__
    return __ffsll((long long) v) - 1;
#else
    return __builtin_ctzll(v);
#endif
}
/* a whole tile is read as 64-bit words (the text starts at an allocation, tiles at multiples of LN_TILE); the last one by bytes */
FLB_HD uint32_t ln_count(const struct ln_env *e, uint32_t t)
{
    const size_t lo = (size_t) t * LN_TILE, hi = lo + LN_TILE < e->bytes ? lo + LN_TILE : e->bytes;
    uint32_t n = 0;
    size_t i;
    if (hi - lo == LN_TILE && (((uintptr_t) (e->text + lo)) & 7u) == 0) {
        const uint64_t *w = (const uint64_t *) (e->text + lo);
        for (i = 0; i < LN_TILE / 8; i++) n += (uint32_t) ln_popc64(ln_nl_mask(w[i]));
        return n;
    }
    for (i = lo; i < hi; i++) n += e->text[i] == '\n';
    return n;
}
FLB_HD void ln_fill(const struct ln_env *e, uint32_t t, uint64_t at)
{
    const size_t lo = (size_t) t * LN_TILE, hi = lo + LN_TILE < e->bytes ? lo + LN_TILE : e->bytes;
    size_t i;
    if (hi - lo == LN_TILE && (((uintptr_t) (e->text + lo)) & 7u) == 0) {
        const uint64_t *w = (const uint64_t *) (e->text + lo);
        for (i = 0; i < LN_TILE / 8; i++) {
            uint64_t m = ln_nl_mask(w[i]);                /* little-endian: byte j of the word is bits 8j .. 8j + 7 */
            while (m) { e->nl[at++] = (uint32_t) (lo + 8 * i + (size_t) (ln_ctz64(m) >> 3)); m &= m - 1; }
        }
        return;
    }
    for (i = lo; i < hi; i++) if (e->text[i] == '\n') e->nl[at++] = (uint32_t) i;
}
/* line k: [start, start + keep) is what the event carries; returns the event's size, 0 when the line is stepped over */
FLB_HD uint32_t ln_line(const struct ln_env *e, uint32_t k, uint32_t *start, uint32_t *keep)
{
    const uint32_t s = k ? e->nl[k - 1] + 1u : 0u, len = e->nl[k] - s;
    uint32_t crlf = 0, n;
    *start = s;
    if (e->skip_empty_lines && (len == 0 || (len == 1 && e->text[s] == '\r'))) return 0;
    if (len >= 2) crlf = e->text[s + len - 1] == '\r';
    *keep = len - crlf;
    n = 12 + 5 + 5;
    if (e->path_key_len != 0xffffffffu) n += mp_str_hdr_size(e->path_key_len) + e->path_key_len + mp_str_hdr_size(e->path_len) + e->path_len;
    if (e->offset_key_len != 0xffffffffu) n += mp_str_hdr_size(e->offset_key_len) + e->offset_key_len + mp_uint_size(e->stream_offset + s);
    n += mp_str_hdr_size(e->key_len) + e->key_len + mp_str_hdr_size(*keep) + *keep;
    return n;
}
FLB_HD uint32_t ln_put(uint8_t *o, const uint8_t *s, uint32_t n) { uint32_t i; for (i = 0; i < n; i++) o[i] = s[i]; return n; }
FLB_HD void ln_emit(const struct ln_env *e, uint32_t k, uint8_t *o)
{
    uint32_t start, keep = 0, n = 0, entries = 1;
    if (!ln_line(e, k, &start, &keep)) return;
    o[0] = 0x92; o[1] = 0x92; o[2] = 0xd7; o[3] = 0x00;
    mp_put_be32(o + 4, (uint32_t) e->sec); mp_put_be32(o + 8, (uint32_t) e->nsec);
    o[12] = 0xdf; o[13] = o[14] = o[15] = o[16] = 0;
    n = 17 + 5;
    if (e->path_key_len != 0xffffffffu) {
        n += mp_put_str_hdr(o + n, e->path_key_len); n += ln_put(o + n, e->strs + e->path_key_off, e->path_key_len);
        n += mp_put_str_hdr(o + n, e->path_len); n += ln_put(o + n, e->strs + e->path_off, e->path_len);
        entries++;
    }
    if (e->offset_key_len != 0xffffffffu) {
        n += mp_put_str_hdr(o + n, e->offset_key_len); n += ln_put(o + n, e->strs + e->offset_key_off, e->offset_key_len);
        n += mp_put_uint(o + n, e->stream_offset + start);
        entries++;
    }
    n += mp_put_str_hdr(o + n, e->key_len); n += ln_put(o + n, e->strs + e->key_off, e->key_len);
    n += mp_put_str_hdr(o + n, keep);
    mp_copy(o + n, e->text + start, keep);               /* (the text buffer is padded like every device buffer) */
    o[17] = 0xdf; mp_put_be32(o + 18, entries);
}

#endif
