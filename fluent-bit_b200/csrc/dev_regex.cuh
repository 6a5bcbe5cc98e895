/* dev_regex.cuh -- device-side regex matcher (one lane = one subject string).
 *
 * Executes struct rx_prog (rx_compile.c) with the semantics of Onigmo's
 * backtracking matcher as Fluent Bit drives it:
 *   search loop ......... lib/onigmo/regexec.c:3793 onig_search (forward, start
 *                         positions are character boundaries, one extra attempt at
 *                         the end of the subject)
 *   one attempt ......... regexec.c:1431 match_at (priority-ordered alternatives,
 *                         captures restored on backtrack)
 *   character length .... enc/utf_8.c mbc_enc_len via regenc.c onigenc_mbclen
 *                         (valid sequence -> its length, truncated -> rest, invalid -> 1)
 * Compiled for the GPU by nvcc; the same text compiles with g++ for the CPU-only
 * device-code emulation used by `-m "not gpu"` tests (tests/hostsim) -- never by
 * the product library.
 */
#ifndef FLBGPU_DEV_REGEX_CUH
#define FLBGPU_DEV_REGEX_CUH

#include <stdint.h>
#include "flbgpu_prog.h"

#ifndef FLB_HD
#ifdef __CUDACC__
#define FLB_HD __host__ __device__ __forceinline__
#ifdef FLB_INLINE_ALL
#define FLB_HDN __host__ __device__ __forceinline__
#else
#define FLB_HDN __host__ __device__ __noinline__
#endif
#else
#define FLB_HD static inline
#define FLB_HDN static
#endif
#endif

/* Lanes of a warp run the same program over different subjects.  Left alone they drift
 * apart (one lane is still stepping a lazy loop while its neighbours are three groups
 * further) and the interpreter's switch serialises them.  On the device every iteration
 * therefore serves only the lanes at the SMALLEST pc among the lanes currently inside the
 * interpreter: laggards catch up, and once aligned the whole warp executes each op -- in
 * particular each character-class run -- together. */
#ifdef __CUDA_ARCH__
#define RX_ALIGN_PC(pc) { const unsigned m_ = __activemask(); \
        if ((unsigned) (pc) != __reduce_min_sync(m_, (unsigned) (pc))) continue; }
#else
#define RX_ALIGN_PC(pc)
#endif

#define RXT_ALT      0u
#define RXT_RESTORE  1u
#define RXT_BACKOFF  2u
#define RXT_LAZY     3u
#define RXT_NULLF    4u
#define RXT_MARK     5u
#define RXT_VOID2    6u
#define RXT_VOID3    7u

/* length of the character starting at s[pos] (pos < len) */
FLB_HD int rx_u8len(const uint8_t *s, int pos, int len)
{
    uint32_t b0 = s[pos], b1;
    int rem = len - pos;
    if (b0 < 0x80) return 1;
    if (b0 < 0xC2 || b0 > 0xF4) return 1;
    if (rem < 2) return rem;                         /* truncated: NEEDMORE -> rest */
    b1 = s[pos + 1];
    if (b0 < 0xE0) return (b1 & 0xC0) == 0x80 ? 2 : 1;
    if (b0 < 0xF0) {
        uint32_t lo = (b0 == 0xE0) ? 0xA0 : 0x80, hi = (b0 == 0xED) ? 0x9F : 0xBF;
        if (b1 < lo || b1 > hi) return 1;
        if (rem < 3) return rem;
        return (s[pos + 2] & 0xC0) == 0x80 ? 3 : 1;
    }
    {
        uint32_t lo = (b0 == 0xF0) ? 0x90 : 0x80, hi = (b0 == 0xF4) ? 0x8F : 0xBF;
        if (b1 < lo || b1 > hi) return 1;
        if (rem < 3) return rem;
        if ((s[pos + 2] & 0xC0) != 0x80) return 1;
        if (rem < 4) return rem;
        return (s[pos + 3] & 0xC0) == 0x80 ? 4 : 1;
    }
}

FLB_HD uint32_t rx_u8code(const uint8_t *s, int pos, int n)
{
    uint32_t b0 = s[pos];
    if (n == 2) return ((b0 & 0x1f) << 6) | (s[pos + 1] & 0x3f);
    if (n == 3) return ((b0 & 0x0f) << 12) | ((uint32_t) (s[pos + 1] & 0x3f) << 6) | (s[pos + 2] & 0x3f);
    if (n == 4) return ((b0 & 0x07) << 18) | ((uint32_t) (s[pos + 1] & 0x3f) << 12) |
                       ((uint32_t) (s[pos + 2] & 0x3f) << 6) | (s[pos + 3] & 0x3f);
    return b0;
}

/* does the character at s[pos] belong to the class?  *n receives its length */
FLB_HD int rx_class_match(const struct rx_prog *pg, const struct rx_class *cl, const uint8_t *s,
                          int pos, int len, int *n)
{
    uint32_t b = s[pos];
    if (b < 0x80) { *n = 1; return (cl->bits[b >> 5] >> (b & 31)) & 1; }
    *n = rx_u8len(s, pos, len);
    if (*n == 1) return (cl->bits[b >> 5] >> (b & 31)) & 1;
    if (cl->mb_mode == RX_MB_ALL) return 1;
    if (cl->mb_mode == RX_MB_NONE) return 0;
    {
        uint32_t cp = rx_u8code(s, pos, *n);
        const uint32_t *r = (const uint32_t *) ((const char *) pg + cl->ranges_off);
        uint32_t i;
        for (i = 0; i < cl->n_ranges; i++) {
            if (cp < r[2 * i]) break;
            if (cp <= r[2 * i + 1]) return 1;
        }
        return 0;
    }
}

/* first position in [pos, len) holding one of k (1..4) ASCII bytes packed in sb (unused slots repeat
 * the first); len if none.  Eight bytes per step on aligned words (buffers are padded, see bk_alloc). */
FLB_HD int rx_find_first_of4(const uint8_t *s, int pos, int len, uint32_t sb, uint32_t k)
{
    const uint64_t ones = 0x0101010101010101ull;
    const uint64_t r0 = ones * (sb & 0xff), r1 = ones * ((sb >> 8) & 0xff), r2 = ones * ((sb >> 16) & 0xff), r3 = ones * (sb >> 24);
    while (pos < len) {
        const uintptr_t a = (uintptr_t) (s + pos);
        const unsigned sh = (unsigned) (a & 7) * 8;
        uint64_t w = *(const uint64_t *) (a & ~(uintptr_t) 7), t, hit;
        w >>= sh;
        if (sh) w |= 0x8080808080808080ull << (64 - sh);           /* filler that is never a stop byte */
        t = w ^ r0; hit = (t - ones) & ~t;
        if (k > 1) {
            t = w ^ r1; hit |= (t - ones) & ~t;
            t = w ^ r2; hit |= (t - ones) & ~t;
            t = w ^ r3; hit |= (t - ones) & ~t;
        }
        hit &= 0x8080808080808080ull;
        if (hit) {
#ifdef __CUDA_ARCH__
            pos += (__ffsll((long long) hit) - 1) >> 3;
#else
            pos += __builtin_ctzll(hit) >> 3;
#endif
            return pos < len ? pos : len;
        }
        pos += 8 - (int) (sh >> 3);
    }
    return len;
}

/* cls*: advance over the longest run of class members starting at pos.  The four ASCII
 * bitmap words live in registers for the whole run (the hot loop of every log pattern). */
FLB_HD int rx_class_run(const struct rx_prog *pg, const struct rx_class *cl, const uint8_t *s, int pos, int len)
{
    const uint32_t w0 = cl->bits[0], w1 = cl->bits[1], w2 = cl->bits[2], w3 = cl->bits[3];
    int n;
    if (cl->pad & 0xff) return rx_find_first_of4(s, pos, len, cl->ranges_off, cl->pad & 0xff);
    while (pos < len) {
        const uint32_t b = s[pos];
        if (b < 0x80) {
            const uint32_t w = (b < 64) ? ((b < 32) ? w0 : w1) : ((b < 96) ? w2 : w3);
            if (!((w >> (b & 31)) & 1)) break;
            pos++;
        }
        else {
            if (!rx_class_match(pg, cl, s, pos, len, &n)) break;
            pos += n;
        }
    }
    return pos;
}

FLB_HD int rx_is_word_at(const uint8_t *s, int pos, int len)
{
    uint32_t b;
    if (pos < 0 || pos >= len) return 0;
    b = s[pos];
    if (b >= 0x80) return 1;   /* ONIG_OPTION_WORD_BOUND_ALL_RANGE: non-ASCII letters are word
                                  characters; non-ASCII punctuation is approximated */
    return (b >= '0' && b <= '9') || (b >= 'A' && b <= 'Z') || (b >= 'a' && b <= 'z') || b == '_';
}

/* start of the character that ends just before `pos` (pos > lo) */
FLB_HD int rx_prev_char(const uint8_t *s, int lo, int pos, int len)
{
    int q = pos - 1;
    if (s[q] < 0x80) return q;
    while (q > lo && (s[q] & 0xC0) == 0x80 && pos - q < 4) q--;
    if (q + rx_u8len(s, q, len) == pos) return q;
    return pos - 1;
}

/* One anchored attempt at `st`.  caps[] must hold 2*(n_groups+1) ints, all -1. */
FLB_HD int rx_match_at(const struct rx_prog *pg, const uint8_t *s, int len, int st, int *caps,
                       uint32_t *stk, int stk_cap, uint32_t *budget)
{
    const uint32_t *code = (const uint32_t *) ((const char *) pg + pg->code_off);
    const struct rx_class *cls = (const struct rx_class *) ((const char *) pg + pg->class_off);
    int pc = 0, pos = st, sp = 0, n, unrestored = 0;
    uint32_t steps = *budget;

#define RX_PUSH2(a, tagword) do { if (sp + 2 > stk_cap) { *budget = steps; return RX_R_ESTACK; } \
        stk[sp] = (uint32_t) (a); stk[sp + 1] = (uint32_t) (tagword); sp += 2; } while (0)

    for (;;) {
        uint32_t w, op, arg;
        RX_ALIGN_PC(pc)
        if (steps-- == 0) { *budget = 0; return RX_R_EBUDGET; }
        w = code[pc]; op = RX_OP(w); arg = RX_ARG(w);
        switch (op) {
        case RX_MATCH:
            *budget = steps;
            return RX_R_MATCH;
        case RX_FAIL:
            goto fail;
        case RX_BYTE:
            if (pos < len && s[pos] == arg) { pos++; pc++; continue; }
            goto fail;
        case RX_STR: {
            int k, nb = (int) arg;
            if (pos + nb > len) goto fail;
            for (k = 0; k < nb; k++) {
                if (s[pos + k] != ((code[pc + 1 + (k >> 2)] >> (8 * (k & 3))) & 0xff)) goto fail;
            }
            pos += nb;
            pc += 1 + ((nb + 3) >> 2);
            continue;
        }
        case RX_CLASS:
            if (pos < len && rx_class_match(pg, &cls[arg], s, pos, len, &n)) { pos += n; pc++; continue; }
            goto fail;
        case RX_JMP:
            pc = (int) arg;
            continue;
        case RX_SPLIT:
            RX_PUSH2(pos, (arg << 3) | RXT_ALT);
            pc++;
            continue;
        case RX_SPLIT_LAZY:
            RX_PUSH2(pos, ((uint32_t) (pc + 1) << 3) | RXT_ALT);
            pc = (int) arg;
            continue;
        case RX_SAVE:
            /* the old value only matters to whoever backtracks past this point: with nothing on the
             * stack that can only be the end of this attempt, which then resets every capture */
            if (sp > 0) RX_PUSH2(caps[arg], (arg << 3) | RXT_RESTORE);
            else unrestored = 1;
            caps[arg] = pos;
            pc++;
            continue;
        case RX_CSTAR_POSS:
            pos = rx_class_run(pg, &cls[arg], s, pos, len);
            pc++;
            continue;
        case RX_CSTAR_BT: {
            int start = pos;
            pos = rx_class_run(pg, &cls[arg], s, pos, len);
            if (pos > start) {
                if (sp + 3 > stk_cap) { *budget = steps; return RX_R_ESTACK; }
                stk[sp] = (uint32_t) start; stk[sp + 1] = (uint32_t) pos;
                stk[sp + 2] = ((uint32_t) pc << 3) | RXT_BACKOFF;
                sp += 3;
            }
            pc++;
            continue;
        }
        case RX_CSTAR_LAZY: {
            /* word 2: 1 + index of the "continuation can start here" byte class (0 = unknown).
             * Positions whose next byte cannot start the continuation would fail at once, so
             * the lazy loop steps over them right away. */
            const uint32_t stop = code[pc + 1];
            if (stop) {
                const struct rx_class *cl = &cls[arg], *sc = &cls[stop - 1];
                const uint32_t nneg = cl->pad & 0xff, npos = (sc->pad >> 9) & 7;
                if (nneg && (sc->pad & 0x100) && nneg + npos <= 4) {
                    /* "every byte but a few" stepping lazily towards "one of a few bytes": the first
                     * byte of either set is where the loop below would stop */
                    uint32_t sb = 0, q = 0, z;
                    for (z = 0; z < nneg; z++) sb |= ((cl->ranges_off >> (8 * z)) & 0xff) << (8 * q++);
                    for (z = 0; z < npos; z++) sb |= ((sc->n_ranges >> (8 * z)) & 0xff) << (8 * q++);
                    for (z = q; z < 4; z++) sb |= (sb & 0xff) << (8 * z);
                    pos = rx_find_first_of4(s, pos, len, sb, q);
                }
                else while (pos < len && !((sc->bits[s[pos] >> 5] >> (s[pos] & 31)) & 1) && rx_class_match(pg, cl, s, pos, len, &n)) pos += n;
            }
            RX_PUSH2(pos, ((uint32_t) pc << 3) | RXT_LAZY);
            pc += 2;
            continue;
        }
        case RX_BOL:
            if (pos == 0 || (s[pos - 1] == '\n' && pos != len)) { pc++; continue; }   /* regexec.c OP_BEGIN_LINE: !ON_STR_END */
            goto fail;
        case RX_EOL:
            if (pos == len || s[pos] == '\n') { pc++; continue; }
            goto fail;
        case RX_BEGIN_BUF:
            if (pos == 0) { pc++; continue; }
            goto fail;
        case RX_END_BUF:
            if (pos == len) { pc++; continue; }
            goto fail;
        case RX_SEMI_END_BUF:
            if (pos == len || (pos == len - 1 && s[pos] == '\n')) { pc++; continue; }
            goto fail;
        case RX_WORD_B:
        case RX_NOT_WORD_B: {
            int a = (pos > 0) ? rx_is_word_at(s, rx_prev_char(s, 0, pos, len), len) : 0;
            int b = rx_is_word_at(s, pos, len);
            if ((a != b) == (op == RX_WORD_B)) { pc++; continue; }
            goto fail;
        }
        case RX_NULL_START:
            RX_PUSH2(pos, (arg << 3) | RXT_NULLF);
            pc++;
            continue;
        case RX_NULL_END: {
            int k = sp, isnull = 0;
            while (k > 0) {
                uint32_t t = stk[k - 1] & 7u;
                if (t == RXT_NULLF && (stk[k - 1] >> 3) == arg) { isnull = ((int) stk[k - 2] == pos); break; }
                k -= (t == RXT_BACKOFF || t == RXT_VOID3) ? 3 : 2;
            }
            pc += isnull ? 2 : 1;
            continue;
        }
        case RX_MARK:
            RX_PUSH2(pos, (arg << 3) | RXT_MARK);
            pc++;
            continue;
        case RX_CUT_POS:
        case RX_CUT_ATOMIC: {
            int k = sp;
            while (k > 0) {
                uint32_t t = stk[k - 1] & 7u;
                if (t == RXT_MARK) {
                    if (op == RX_CUT_POS) pos = (int) stk[k - 2];
                    stk[k - 1] = RXT_VOID2;
                    break;
                }
                if (t == RXT_BACKOFF) { stk[k - 1] = RXT_VOID3; k -= 3; continue; }
                if (t == RXT_VOID3) { k -= 3; continue; }
                if (t != RXT_RESTORE) stk[k - 1] = RXT_VOID2;
                k -= 2;
            }
            pc++;
            continue;
        }
        case RX_CUT_NEG: {
            /* the body of a negative look-ahead matched: unwind to its mark, then
             * discard the "look-ahead failed, carry on" alternative under it */
            while (sp > 0) {
                uint32_t t = stk[sp - 1] & 7u;
                if (t == RXT_MARK) { sp -= 2; break; }
                if (t == RXT_RESTORE) { caps[stk[sp - 1] >> 3] = (int) stk[sp - 2]; sp -= 2; continue; }
                sp -= (t == RXT_BACKOFF || t == RXT_VOID3) ? 3 : 2;
            }
            if (sp >= 2) sp -= 2;
            goto fail;
        }
        case RX_BACKREF: {
            int b = caps[2 * arg], e = caps[2 * arg + 1], k;
            if (b < 0 || e < 0) goto fail;
            if (pos + (e - b) > len) goto fail;
            for (k = 0; k < e - b; k++) if (s[pos + k] != s[b + k]) goto fail;
            pos += e - b;
            pc++;
            continue;
        }
        default:
            goto fail;
        }
fail:
        for (;;) {
            uint32_t top, t;
            if (sp == 0) {
                /* captures written without a restore record (see RX_SAVE) go back to "unset" here */
                if (unrestored) { const int nc = 2 * ((int) pg->n_groups + 1); int z; for (z = 0; z < nc; z++) caps[z] = -1; }
                *budget = steps;
                return RX_R_NOMATCH;
            }
            top = stk[sp - 1]; t = top & 7u;
            if (t == RXT_ALT) { pc = (int) (top >> 3); pos = (int) stk[sp - 2]; sp -= 2; break; }
            if (t == RXT_RESTORE) { caps[top >> 3] = (int) stk[sp - 2]; sp -= 2; continue; }
            if (t == RXT_BACKOFF) {
                int start = (int) stk[sp - 3], cur = (int) stk[sp - 2];
                cur = rx_prev_char(s, start, cur, len);
                pos = cur;
                pc = (int) (top >> 3) + 1;
                if (cur > start) stk[sp - 2] = (uint32_t) cur; else sp -= 3;
                break;
            }
            if (t == RXT_LAZY) {
                int p0 = (int) stk[sp - 2], opc = (int) (top >> 3);
                if (p0 < len && rx_class_match(pg, &cls[RX_ARG(code[opc])], s, p0, len, &n)) {
                    const uint32_t stop = code[opc + 1];
                    pos = p0 + n;
                    if (stop) {
                        const struct rx_class *cl = &cls[RX_ARG(code[opc])], *sc = &cls[stop - 1];
                        while (pos < len && !((sc->bits[s[pos] >> 5] >> (s[pos] & 31)) & 1) && rx_class_match(pg, cl, s, pos, len, &n)) pos += n;
                    }
                    stk[sp - 2] = (uint32_t) pos;
                    pc = opc + 2;
                    break;
                }
                sp -= 2;
                continue;
            }
            sp -= (t == RXT_VOID3) ? 3 : 2;     /* NULLF, MARK, VOID2, VOID3 */
        }
    }
#undef RX_PUSH2
}

/* Leftmost search over s[0,len).  Returns RX_R_*; on RX_R_MATCH caps[0..2*n_groups+1]
 * hold begin/end byte offsets (-1 = group did not participate). */
FLB_HD int rx_search(const struct rx_prog *pg, const uint8_t *s, int len, int *caps,
                     uint32_t *stk, int stk_cap, uint32_t *budget)
{
    int st = 0, ncap = 2 * ((int) pg->n_groups + 1), i, r;
    for (i = 0; i < ncap; i++) caps[i] = -1;
    if (pg->flags & RX_F_ASCII_ONLY) { for (i = 0; i < len; i++) if (s[i] >= 0x80) return RX_R_EUNICODE; }
    for (;;) {
        int ok = 1;
        if ((pg->flags & RX_F_ANCHOR_BUF) && st > 0) return RX_R_NOMATCH;
        if ((pg->flags & RX_F_ANCHOR_BOL) && st > 0 && (s[st - 1] != '\n' || st == len)) ok = 0;
        if (ok && (pg->flags & RX_F_HAS_FIRSTSET)) {
            if (st >= len) return RX_R_NOMATCH;
            ok = (pg->first[s[st] >> 5] >> (s[st] & 31)) & 1;
        }
        if (ok && (pg->flags & RX_F_HAS_SECONDSET) && st < len && s[st] < 0x80) {
            /* the first character is this one byte; something has to follow it */
            ok = st + 1 < len && ((pg->second[s[st + 1] >> 5] >> (s[st + 1] & 31)) & 1);
        }
        if (ok) {
            r = rx_match_at(pg, s, len, st, caps, stk, stk_cap, budget);
            if (r != RX_R_NOMATCH) return r;
        }
        if (st >= len) return RX_R_NOMATCH;
        if (pg->flags & RX_F_ANCHOR_BOL) {
            /* every match starts at a line start: the next candidate is the position behind the next line feed (never a
             * continuation byte, so also a character boundary) -- a line without one is done after its first position */
            st = rx_find_first_of4(s, st, len, 0x0a0a0a0au, 1);
            if (st >= len) return RX_R_NOMATCH;
            st++;
        }
        else st += rx_u8len(s, st, len);
    }
}

#endif
