/* oracle: regular expressions with the semantics fluent-bit gets from Onigmo.
 * TEST INFRASTRUCTURE (see orc.h).
 *
 * What is restated: src/flb_regex.c:60-150 (the /pattern/imx wrapper), and from lib/onigmo the
 * behaviour of ONIG_SYNTAX_RUBY + ONIG_ENCODING_UTF8 as fluent-bit uses it:
 *   - regparse.c: Ruby syntax; named groups switch plain (...) to non-capturing (regparse.c:66
 *     ONIG_OPTION_CAPTURE_GROUP off); \d \s \w \h and are ASCII-only (ONIG_OPTION_ASCII_RANGE,
 *     regsyntax.c OnigSyntaxRuby); a fixed interval followed by '?' is an optional, not lazy
 *   - regexec.c:match_at (1431): backtracking with alternatives in priority order, greedy / lazy /
 *     possessive repeats, an iteration that consumed nothing ends its loop (NULL_CHECK), captures
 *     restored on backtrack; OP_BEGIN_LINE does not match at the very end of the subject
 *   - regexec.c:onig_search (3793): leftmost match, start positions on character boundaries
 *   - enc/utf_8.c: character length from the lead byte, clipped to the subject end
 * Design: AST + recursive matcher with explicit continuations -- deliberately unlike the
 * product's bytecode VM.  Approximations shared with the product's documentation: POSIX brackets
 * and case folding are ASCII; a non-ASCII character counts as a word character for \b. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "orc.h"

enum { N_EMPTY, N_BYTES, N_ANY, N_CLASS, N_CAT, N_ALT, N_REP, N_GROUP, N_ATOMIC, N_LOOK, N_NLOOK, N_BACKREF,
       N_BOL, N_EOL, N_BBUF, N_EBUF, N_SEBUF, N_WB, N_NWB };

struct crange { uint32_t lo, hi; };
struct cclass { uint8_t ascii[16]; struct crange *r; int nr; int neg; };

struct node {
    int type;
    struct node *a, *b;
    uint8_t *bytes; int nbytes; int icase;       /* N_BYTES */
    int dotall;                                  /* N_ANY */
    struct cclass *cc;                           /* N_CLASS */
    int min, max, mode;                          /* N_REP: mode 0 greedy 1 lazy 2 possessive; max -1 = inf */
    int grp;                                     /* N_GROUP / N_BACKREF */
};

struct orc_regex {
    struct orc_arena arena;
    struct node *root;
    int ngroups;                                 /* including group 0 */
    int nnames;
    char *names[256]; int name_grp[256];
};

struct parser {
    struct orc_regex *re;
    const uint8_t *p, *end;
    int icase, dotall, extend;
    int has_named, ngroups, found_named;
    const char *err;
};

static struct node *mk(struct parser *P, int type)
{
    struct node *n = orc_alloc(&P->re->arena, sizeof(*n));
    memset(n, 0, sizeof(*n));
    n->type = type;
    return n;
}

/* ---- character classes ---------------------------------------------------------------- */
static void cc_set(struct cclass *c, unsigned b) { c->ascii[b >> 3] |= (uint8_t) (1u << (b & 7)); }
static int cc_get(const struct cclass *c, unsigned b) { return (c->ascii[b >> 3] >> (b & 7)) & 1; }

static void cc_add_range(struct parser *P, struct cclass *c, uint32_t lo, uint32_t hi)
{
    uint32_t v;
    for (v = lo; v <= hi && v < 128; v++) cc_set(c, v);
    if (hi >= 128) {
        struct crange *nr = orc_alloc(&P->re->arena, sizeof(*nr) * (size_t) (c->nr + 1));
        if (c->nr) memcpy(nr, c->r, sizeof(*nr) * (size_t) c->nr);
        nr[c->nr].lo = lo < 128 ? 128 : lo; nr[c->nr].hi = hi;
        c->r = nr; c->nr++;
    }
}

static int cc_has(const struct cclass *c, uint32_t cp)
{
    int i, in = 0;
    if (cp < 128) in = cc_get(c, cp);
    else for (i = 0; i < c->nr; i++) if (cp >= c->r[i].lo && cp <= c->r[i].hi) { in = 1; break; }
    return in != c->neg;
}

/* union of `src` (honouring its own negation) into `dst` */
static void cc_union(struct parser *P, struct cclass *dst, const struct cclass *src)
{
    uint32_t v;
    int i;
    if (!src->neg) {
        for (v = 0; v < 16; v++) dst->ascii[v] |= src->ascii[v];
        for (i = 0; i < src->nr; i++) cc_add_range(P, dst, src->r[i].lo, src->r[i].hi);
        return;
    }
    for (v = 0; v < 16; v++) dst->ascii[v] |= (uint8_t) ~src->ascii[v];
    {   /* complement of the non-ASCII ranges over [0x80, 0x7fffffff] */
        uint32_t cur = 128;
        for (;;) {
            uint32_t best_lo = 0xffffffffu, best_hi = 0;
            for (i = 0; i < src->nr; i++)
                if (src->r[i].hi >= cur && src->r[i].lo < best_lo) { best_lo = src->r[i].lo; best_hi = src->r[i].hi; }
            if (best_lo == 0xffffffffu) { cc_add_range(P, dst, cur, 0x7fffffffu); break; }
            if (best_lo > cur) cc_add_range(P, dst, cur, best_lo - 1);
            if (best_hi >= 0x7fffffffu) break;
            cur = best_hi + 1 > cur ? best_hi + 1 : cur;
        }
    }
}

/* \d \w \s \h (lower case) into c; upper case = complement, which includes every non-ASCII char */
static int cc_shorthand(struct parser *P, struct cclass *dst, int ch)
{
    struct cclass t;
    int lower = ch | 0x20, v;
    memset(&t, 0, sizeof(t));
    for (v = 0; v < 128; v++) {
        int in = 0;
        switch (lower) {
        case 'd': in = v >= '0' && v <= '9'; break;
        case 'w': in = (v >= '0' && v <= '9') || (v >= 'a' && v <= 'z') || (v >= 'A' && v <= 'Z') || v == '_'; break;
        case 's': in = v == ' ' || (v >= 9 && v <= 13); break;
        case 'h': in = (v >= '0' && v <= '9') || (v >= 'a' && v <= 'f') || (v >= 'A' && v <= 'F'); break;
        default: return -1;
        }
        if (in) cc_set(&t, (unsigned) v);
    }
    t.neg = (ch != lower);
    cc_union(P, dst, &t);
    return 0;
}

static int cc_posix(struct parser *P, struct cclass *dst, const char *name, size_t n, int neg)
{
    struct cclass t;
    int v;
    static const char *names[] = { "alnum", "alpha", "ascii", "blank", "cntrl", "digit", "graph", "lower", "print",
                                   "punct", "space", "upper", "xdigit", "word" };
    int which = -1;
    for (v = 0; v < 14; v++) if (strlen(names[v]) == n && !memcmp(names[v], name, n)) which = v;
    if (which < 0) return -1;
    memset(&t, 0, sizeof(t));
    for (v = 0; v < 128; v++) {
        int up = v >= 'A' && v <= 'Z', lo = v >= 'a' && v <= 'z', dg = v >= '0' && v <= '9', in = 0;
        switch (which) {
        case 0: in = up || lo || dg; break;
        case 1: in = up || lo; break;
        case 2: in = 1; break;
        case 3: in = v == ' ' || v == '\t'; break;
        case 4: in = v < 32 || v == 127; break;
        case 5: in = dg; break;
        case 6: in = v > 32 && v < 127; break;
        case 7: in = lo; break;
        case 8: in = v >= 32 && v < 127; break;
        case 9: in = v > 32 && v < 127 && !(up || lo || dg); break;
        case 10: in = v == ' ' || (v >= 9 && v <= 13); break;
        case 11: in = up; break;
        case 12: in = dg || (v >= 'a' && v <= 'f') || (v >= 'A' && v <= 'F'); break;
        case 13: in = up || lo || dg || v == '_'; break;
        }
        if (in) cc_set(&t, (unsigned) v);
    }
    t.neg = neg;
    cc_union(P, dst, &t);
    return 0;
}

/* ---- UTF-8 (lib/onigmo/enc/utf_8.c) ------------------------------------------------------ */
static int u8len(unsigned b)
{
    if (b < 0xc2) return 1;
    if (b < 0xe0) return 2;
    if (b < 0xf0) return 3;
    if (b < 0xf5) return 4;
    return 1;
}

static uint32_t u8decode(const uint8_t *s, int pos, int len, int *l)
{
    int n = u8len(s[pos]), i;
    uint32_t c = s[pos];
    if (pos + n > len) n = len - pos;
    *l = n;
    if (n == 1) return c;
    c &= (0xffu >> (n + 1));
    for (i = 1; i < n; i++) c = (c << 6) | (s[pos + i] & 0x3f);
    return c;
}

static int u8encode(uint32_t c, uint8_t *o)
{
    if (c < 0x80) { o[0] = (uint8_t) c; return 1; }
    if (c < 0x800) { o[0] = (uint8_t) (0xc0 | (c >> 6)); o[1] = (uint8_t) (0x80 | (c & 63)); return 2; }
    if (c < 0x10000) { o[0] = (uint8_t) (0xe0 | (c >> 12)); o[1] = (uint8_t) (0x80 | ((c >> 6) & 63)); o[2] = (uint8_t) (0x80 | (c & 63)); return 3; }
    o[0] = (uint8_t) (0xf0 | (c >> 18)); o[1] = (uint8_t) (0x80 | ((c >> 12) & 63)); o[2] = (uint8_t) (0x80 | ((c >> 6) & 63));
    o[3] = (uint8_t) (0x80 | (c & 63));
    return 4;
}

/* ---- parser (lib/onigmo/regparse.c, Ruby syntax) ----------------------------------------- */
static struct node *parse_alt(struct parser *P, int depth);

static int hexv(int c)
{
    if (c >= '0' && c <= '9') return c - '0';
    if (c >= 'a' && c <= 'f') return c - 'a' + 10;
    if (c >= 'A' && c <= 'F') return c - 'A' + 10;
    return -1;
}

/* escape that denotes a single code point; returns -1 when it is something else */
static int64_t parse_char_escape(struct parser *P, int in_class)
{
    int c = *P->p;
    switch (c) {
    case 't': P->p++; return '\t';
    case 'n': P->p++; return '\n';
    case 'r': P->p++; return '\r';
    case 'f': P->p++; return '\f';
    case 'v': P->p++; return '\v';
    case 'a': P->p++; return 7;
    case 'e': P->p++; return 27;
    case 'b': if (in_class) { P->p++; return 8; } return -1;
    case 'x': {
        int64_t v = 0; int nd = 0;
        P->p++;
        if (P->p < P->end && *P->p == '{') {
            P->p++;
            while (P->p < P->end && hexv(*P->p) >= 0 && nd < 8) { v = v * 16 + hexv(*P->p); P->p++; nd++; }
            if (P->p >= P->end || *P->p != '}' || !nd) { P->err = "bad \\x{}"; return 0; }
            P->p++;
            return v;
        }
        while (P->p < P->end && hexv(*P->p) >= 0 && nd < 2) { v = v * 16 + hexv(*P->p); P->p++; nd++; }
        if (v >= 0x80) { P->err = "raw high byte escape unsupported"; return 0; }
        return v;
    }
    case 'u': {
        int64_t v = 0; int nd = 0;
        P->p++;
        while (P->p < P->end && hexv(*P->p) >= 0 && nd < 4) { v = v * 16 + hexv(*P->p); P->p++; nd++; }
        if (nd != 4) { P->err = "bad \\u"; return 0; }
        return v;
    }
    case '0': {
        int64_t v = 0; int nd = 0;
        P->p++;
        while (P->p < P->end && *P->p >= '0' && *P->p <= '7' && nd < 2) { v = v * 8 + (*P->p - '0'); P->p++; nd++; }
        return v;
    }
    default:
        break;
    }
    if ((c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || (c >= '0' && c <= '9')) return -1;
    {   /* escaped punctuation or multibyte: the character itself */
        int l;
        uint32_t cp = u8decode(P->p, 0, (int) (P->end - P->p), &l);
        P->p += l;
        return cp;
    }
}

static struct cclass *parse_class(struct parser *P, int depth)
{
    struct cclass *c = orc_alloc(&P->re->arena, sizeof(*c));
    int first = 1;
    memset(c, 0, sizeof(*c));
    if (depth > 16) { P->err = "class nesting"; return c; }
    if (P->p < P->end && *P->p == '^') { c->neg = 1; P->p++; }
    for (;;) {
        int64_t lo;
        int have_lo = 0;
        if (P->p >= P->end) { P->err = "premature end of char-class"; return c; }
        if (*P->p == ']' && !first) { P->p++; break; }
        first = 0;
        if (*P->p == '[') {
            if (P->p + 1 < P->end && P->p[1] == ':') {
                const uint8_t *q = P->p + 2, *st;
                int neg = 0;
                if (q < P->end && *q == '^') { neg = 1; q++; }
                st = q;
                while (q < P->end && *q >= 'a' && *q <= 'z') q++;
                if (q + 1 < P->end && q[0] == ':' && q[1] == ']' && cc_posix(P, c, (const char *) st, (size_t) (q - st), neg) == 0) {
                    P->p = q + 2;
                    continue;
                }
            }
            P->p++;
            { struct cclass *sub = parse_class(P, depth + 1); if (P->err) return c; cc_union(P, c, sub); }
            continue;
        }
        if (*P->p == '&' && P->p + 1 < P->end && P->p[1] == '&') { P->err = "class intersection unsupported"; return c; }
        if (*P->p == '\\') {
            P->p++;
            if (P->p >= P->end) { P->err = "end pattern at escape"; return c; }
            if (cc_shorthand(P, c, *P->p) == 0) { P->p++; continue; }
            lo = parse_char_escape(P, 1);
            if (P->err) return c;
            if (lo < 0) { P->err = "unsupported escape in class"; return c; }
            have_lo = 1;
        }
        if (!have_lo) { int l; lo = u8decode(P->p, 0, (int) (P->end - P->p), &l); P->p += l; }
        if (P->p + 1 < P->end && P->p[0] == '-' && P->p[1] != ']') {
            int64_t hi;
            const uint8_t *save = P->p;
            P->p++;
            if (*P->p == '[') { P->p = save; cc_add_range(P, c, (uint32_t) lo, (uint32_t) lo); continue; }
            if (*P->p == '\\') {
                P->p++;
                if (P->p >= P->end) { P->err = "end pattern at escape"; return c; }
                if (cc_shorthand(P, c, *P->p) == 0) {      /* [a-\d]: '-' literal */
                    P->p++; cc_add_range(P, c, (uint32_t) lo, (uint32_t) lo); cc_add_range(P, c, '-', '-'); continue;
                }
                hi = parse_char_escape(P, 1);
                if (P->err) return c;
                if (hi < 0) { P->err = "unsupported escape in class"; return c; }
            }
            else { int l; hi = u8decode(P->p, 0, (int) (P->end - P->p), &l); P->p += l; }
            if (hi < lo) { P->err = "empty range in char class"; return c; }
            cc_add_range(P, c, (uint32_t) lo, (uint32_t) hi);
            continue;
        }
        cc_add_range(P, c, (uint32_t) lo, (uint32_t) lo);
    }
    if (P->icase) {
        int v;
        for (v = 'a'; v <= 'z'; v++) {
            if (cc_get(c, (unsigned) v) || cc_get(c, (unsigned) (v - 32))) { cc_set(c, (unsigned) v); cc_set(c, (unsigned) (v - 32)); }
        }
    }
    return c;
}

static struct node *mk_class_short(struct parser *P, int ch)
{
    struct node *n = mk(P, N_CLASS);
    n->cc = orc_alloc(&P->re->arena, sizeof(*n->cc));
    memset(n->cc, 0, sizeof(*n->cc));
    cc_shorthand(P, n->cc, ch);
    return n;
}

static struct node *mk_cp(struct parser *P, uint32_t cp)
{
    struct node *n = mk(P, N_BYTES);
    n->bytes = orc_alloc(&P->re->arena, 8);
    n->nbytes = u8encode(cp, n->bytes);
    n->icase = P->icase;
    return n;
}

static int parse_int(struct parser *P)
{
    int v = -1;
    while (P->p < P->end && *P->p >= '0' && *P->p <= '9') { if (v < 0) v = 0; if (v < 100000) v = v * 10 + (*P->p - '0'); P->p++; }
    return v;
}

static struct node *parse_atom(struct parser *P, int depth)
{
    int c;
    if (depth > 200) { P->err = "too deep"; return mk(P, N_EMPTY); }
    c = *P->p;
    if (c == '(') {
        struct node *n, *g;
        int si = P->icase, sd = P->dotall, sx = P->extend;
        P->p++;
        if (P->p < P->end && *P->p == '?') {
            P->p++;
            if (P->p >= P->end) { P->err = "end pattern in group"; return mk(P, N_EMPTY); }
            c = *P->p;
            if (c == '#') { while (P->p < P->end && *P->p != ')') P->p++; if (P->p < P->end) P->p++; return mk(P, N_EMPTY); }
            if (c == ':' || c == '>' || c == '=' || c == '!') {
                P->p++;
                n = parse_alt(P, depth + 1);
                if (P->p >= P->end || *P->p != ')') { if (!P->err) P->err = "end pattern with unmatched parenthesis"; return n; }
                P->p++;
                P->icase = si; P->dotall = sd; P->extend = sx;
                if (c == ':') return n;
                g = mk(P, c == '>' ? N_ATOMIC : c == '=' ? N_LOOK : N_NLOOK);
                g->a = n;
                return g;
            }
            if (c == '<' || c == '\'') {
                int close = c == '<' ? '>' : '\'';
                const uint8_t *st;
                if (c == '<' && P->p + 1 < P->end && (P->p[1] == '=' || P->p[1] == '!')) { P->err = "look-behind unsupported"; return mk(P, N_EMPTY); }
                P->p++;
                st = P->p;
                while (P->p < P->end && *P->p != close) P->p++;
                if (P->p >= P->end || P->p == st) { P->err = "group name is empty"; return mk(P, N_EMPTY); }
                P->found_named = 1;
                g = mk(P, N_GROUP);
                g->grp = P->ngroups++;
                if (P->re->nnames < 256 && P->has_named) {
                    char *nm = orc_alloc(&P->re->arena, (size_t) (P->p - st) + 1);
                    memcpy(nm, st, (size_t) (P->p - st)); nm[P->p - st] = 0;
                    P->re->names[P->re->nnames] = nm; P->re->name_grp[P->re->nnames] = g->grp; P->re->nnames++;
                }
                P->p++;
                g->a = parse_alt(P, depth + 1);
                if (P->p >= P->end || *P->p != ')') { if (!P->err) P->err = "end pattern with unmatched parenthesis"; return g; }
                P->p++;
                P->icase = si; P->dotall = sd; P->extend = sx;
                return g;
            }
            {   /* (?imx-imx) and (?imx-imx:...) */
                int on = 1, ni = P->icase, nd = P->dotall, nx = P->extend;
                for (;; P->p++) {
                    if (P->p >= P->end) { P->err = "end pattern in group"; return mk(P, N_EMPTY); }
                    c = *P->p;
                    if (c == '-') on = 0;
                    else if (c == 'i') ni = on;
                    else if (c == 'm') nd = on;
                    else if (c == 'x') nx = on;
                    else if (c == ')' || c == ':') break;
                    else { P->err = "undefined group option"; return mk(P, N_EMPTY); }
                }
                P->p++;
                P->icase = ni; P->dotall = nd; P->extend = nx;
                if (c == ')') {              /* applies to the rest of the enclosing group */
                    n = parse_alt(P, depth + 1);
                    P->icase = si; P->dotall = sd; P->extend = sx;
                    return n;
                }
                n = parse_alt(P, depth + 1);
                if (P->p >= P->end || *P->p != ')') { if (!P->err) P->err = "end pattern with unmatched parenthesis"; return n; }
                P->p++;
                P->icase = si; P->dotall = sd; P->extend = sx;
                return n;
            }
        }
        /* plain group: captures only when the pattern has no named group */
        g = NULL;
        if (!P->has_named) { g = mk(P, N_GROUP); g->grp = P->ngroups++; }
        n = parse_alt(P, depth + 1);
        if (P->p >= P->end || *P->p != ')') { if (!P->err) P->err = "end pattern with unmatched parenthesis"; return n; }
        P->p++;
        P->icase = si; P->dotall = sd; P->extend = sx;
        if (g) { g->a = n; return g; }
        return n;
    }
    if (c == '[') { struct node *n = mk(P, N_CLASS); P->p++; n->cc = parse_class(P, 0); return n; }
    if (c == '.') { struct node *n = mk(P, N_ANY); n->dotall = P->dotall; P->p++; return n; }
    if (c == '^') { P->p++; return mk(P, N_BOL); }
    if (c == '$') { P->p++; return mk(P, N_EOL); }
    if (c == '\\') {
        int64_t cp;
        P->p++;
        if (P->p >= P->end) { P->err = "end pattern at escape"; return mk(P, N_EMPTY); }
        c = *P->p;
        switch (c) {
        case 'd': case 'D': case 'w': case 'W': case 's': case 'S': case 'h': case 'H': P->p++; return mk_class_short(P, c);
        case 'A': P->p++; return mk(P, N_BBUF);
        case 'z': P->p++; return mk(P, N_EBUF);
        case 'Z': P->p++; return mk(P, N_SEBUF);
        case 'b': P->p++; return mk(P, N_WB);
        case 'B': P->p++; return mk(P, N_NWB);
        case 'k': {
            const uint8_t *st;
            struct node *n = mk(P, N_BACKREF);
            int i;
            P->p++;
            if (P->p >= P->end || *P->p != '<') { P->err = "invalid backref"; return n; }
            st = ++P->p;
            while (P->p < P->end && *P->p != '>') P->p++;
            if (P->p >= P->end) { P->err = "invalid backref"; return n; }
            n->grp = -1; n->icase = P->icase;
            for (i = 0; i < P->re->nnames; i++)
                if (strlen(P->re->names[i]) == (size_t) (P->p - st) && !memcmp(P->re->names[i], st, (size_t) (P->p - st))) n->grp = P->re->name_grp[i];
            if (n->grp < 0) {
                int v = 0, ok = P->p > st;
                const uint8_t *q;
                for (q = st; q < P->p; q++) { if (*q < '0' || *q > '9') ok = 0; else v = v * 10 + (*q - '0'); }
                if (ok) n->grp = v; else if (P->has_named || !P->found_named) P->err = "undefined name reference";
            }
            P->p++;
            return n;
        }
        default: break;
        }
        if (c >= '1' && c <= '9') {
            struct node *n = mk(P, N_BACKREF);
            n->grp = parse_int(P); n->icase = P->icase;
            return n;
        }
        cp = parse_char_escape(P, 0);
        if (P->err) return mk(P, N_EMPTY);
        if (cp < 0) { P->err = "unsupported escape"; return mk(P, N_EMPTY); }
        return mk_cp(P, (uint32_t) cp);
    }
    {   /* literal character */
        int l;
        uint32_t cp = u8decode(P->p, 0, (int) (P->end - P->p), &l);
        struct node *n;
        if (l > 1 || cp < 0x80) {
            n = mk(P, N_BYTES);
            n->bytes = orc_alloc(&P->re->arena, 8);
            memcpy(n->bytes, P->p, (size_t) l); n->nbytes = l; n->icase = P->icase;
        }
        else { n = mk(P, N_BYTES); n->bytes = orc_alloc(&P->re->arena, 8); n->bytes[0] = *P->p; n->nbytes = 1; }
        P->p += l;
        return n;
    }
}

static void skip_extended(struct parser *P)
{
    while (P->extend && P->p < P->end) {
        if (*P->p == ' ' || (*P->p >= 9 && *P->p <= 13)) P->p++;
        else if (*P->p == '#') { while (P->p < P->end && *P->p != '\n') P->p++; }
        else break;
    }
}

static struct node *parse_piece(struct parser *P, int depth)
{
    struct node *a = parse_atom(P, depth);
    for (;;) {
        int min, max, interval = 0;
        const uint8_t *save;
        struct node *r;
        if (P->err) return a;
        skip_extended(P);
        if (P->p >= P->end) return a;
        save = P->p;
        if (*P->p == '*') { min = 0; max = -1; P->p++; }
        else if (*P->p == '+') { min = 1; max = -1; P->p++; }
        else if (*P->p == '?') { min = 0; max = 1; P->p++; }
        else if (*P->p == '{') {
            P->p++;
            min = parse_int(P);
            if (P->p < P->end && *P->p == ',') {
                P->p++;
                max = parse_int(P);
                if (min < 0 && max < 0) { P->p = save; return a; }
                if (min < 0) min = 0;
            }
            else { if (min < 0) { P->p = save; return a; } max = min; }
            if (P->p >= P->end || *P->p != '}') { P->p = save; return a; }
            P->p++;
            interval = 1;
            if (max >= 0 && max < min) { P->err = "too big wide range"; return a; }
        }
        else return a;
        if (a->type == N_BOL || a->type == N_EOL || a->type == N_BBUF || a->type == N_EBUF || a->type == N_SEBUF ||
            a->type == N_WB || a->type == N_NWB || a->type == N_LOOK || a->type == N_NLOOK) {
            if (a->type != N_LOOK && a->type != N_NLOOK) { P->err = "target of repeat operator is invalid"; return a; }
        }
        r = mk(P, N_REP);
        r->a = a; r->min = min; r->max = max; r->mode = 0;
        if (P->p < P->end && *P->p == '?') {
            if (interval && min == max) {       /* ONIG_SYN_FIXED_INTERVAL_IS_GREEDY_ONLY: {n}? = optional */
                struct node *o = mk(P, N_REP);
                P->p++;
                o->a = r; o->min = 0; o->max = 1;
                a = o;
                continue;
            }
            r->mode = 1; P->p++;
        }
        else if (!interval && P->p < P->end && *P->p == '+') { r->mode = 2; P->p++; }
        a = r;
    }
}

static struct node *parse_cat(struct parser *P, int depth)
{
    struct node *head = NULL;
    for (;;) {
        struct node *piece, *c;
        skip_extended(P);
        if (P->err || P->p >= P->end || *P->p == '|' || *P->p == ')') break;
        /* regparse.c parse_exp, TK_OP_REPEAT with nothing before it: Ruby syntax has CONTEXT_INVALID_REPEAT_OPS */
        if (*P->p == '*' || *P->p == '+' || *P->p == '?') { P->err = "target of repeat operator is not specified"; break; }
        piece = parse_piece(P, depth);
        if (!head) head = piece;
        else { c = mk(P, N_CAT); c->a = head; c->b = piece; head = c; }
    }
    return head ? head : mk(P, N_EMPTY);
}

static struct node *parse_alt(struct parser *P, int depth)
{
    struct node *a = parse_cat(P, depth);
    while (!P->err && P->p < P->end && *P->p == '|') {
        struct node *n = mk(P, N_ALT);
        P->p++;
        n->a = a; n->b = parse_cat(P, depth);
        a = n;
    }
    return a;
}

struct orc_regex *orc_regex_create(const char *pattern, char *err, size_t errlen)
{
    struct orc_regex *re = calloc(1, sizeof(*re));
    struct parser P;
    size_t len = strlen(pattern);
    const char *start = pattern, *end = pattern + len;
    int icase = 0, dotall = 0, extend = 0, pass;

    /* src/flb_regex.c:60-150 */
    if (len && pattern[0] == '/') {
        const char *last = strrchr(pattern, '/');
        const char *new_end = NULL;
        if (last && last != pattern && last != end) {
            const char *q;
            int any = 0, bad = 0;
            for (q = last + 1; *q; q++) {
                if (*q == 'm') { dotall = 1; any = 1; }
                else if (*q == 'i') { icase = 1; any = 1; }
                else if (*q == 'x') { extend = 1; any = 1; }
                else if (*q == 'o') { }
                else bad = 1;
            }
            if (bad || !any) { icase = dotall = extend = 0; }
            else new_end = last;
        }
        if (pattern[len - 1] == '/') { start++; end--; }
        if (new_end) { start = pattern + 1; end = new_end; }
    }
    for (pass = 0; pass < 2; pass++) {
        memset(&P, 0, sizeof(P));
        orc_arena_free(&re->arena);
        re->nnames = 0;
        P.re = re; P.p = (const uint8_t *) start; P.end = (const uint8_t *) end;
        P.icase = icase; P.dotall = dotall; P.extend = extend;
        P.has_named = pass; P.ngroups = 1;
        re->root = parse_alt(&P, 0);
        if (!P.err && P.p < P.end) P.err = "unmatched close parenthesis";
        if (P.err) {
            if (err) snprintf(err, errlen, "%s", P.err);
            orc_regex_destroy(re);
            return NULL;
        }
        re->ngroups = P.ngroups;
        if (!P.found_named) break;             /* pass 0 was right: no named group */
    }
    return re;
}

void orc_regex_destroy(struct orc_regex *r)
{
    if (!r) return;
    orc_arena_free(&r->arena);
    free(r);
}

int orc_regex_ngroups(const struct orc_regex *r) { return r->ngroups; }
int orc_regex_nnames(const struct orc_regex *r) { return r->nnames; }
const char *orc_regex_name(const struct orc_regex *r, int i, int *group) { *group = r->name_grp[i]; return r->names[i]; }

/* ---- matcher (lib/onigmo/regexec.c:match_at) ------------------------------------------------ */
enum { K_NODE, K_REP, K_CLOSE, K_STOP };
struct cont {
    int kind;
    const struct node *n;
    const struct cont *next;
    int count, start;            /* K_REP: iterations done, position where this iteration began */
    int gstart;                  /* K_CLOSE */
    int *stop_pos;               /* K_STOP */
};

struct mctx {
    const uint8_t *s;
    int len;
    int *cap;                    /* 2 * ngroups */
    int ngroups;
    long steps;
    int end;
};

#define ORC_STEP_LIMIT 50000000L

static int run(struct mctx *M, const struct cont *k, int pos);
static int m(struct mctx *M, const struct node *n, int pos, const struct cont *k);

static int lower(int c) { return (c >= 'A' && c <= 'Z') ? c + 32 : c; }

static int is_word_at(const struct mctx *M, int pos)
{
    unsigned c;
    if (pos < 0 || pos >= M->len) return 0;
    c = M->s[pos];
    if (c >= 0x80) return 1;
    return (c >= '0' && c <= '9') || (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || c == '_';
}

static int prev_char(const struct mctx *M, int pos)
{
    int p = pos - 1;
    while (p > 0 && (M->s[p] & 0xc0) == 0x80 && pos - p < 4) p--;
    /* onigenc left_adjust_char_head: step back over continuation bytes */
    if (p + u8len(M->s[p]) < pos) return pos - 1;
    return p;
}

/* one character matched by a simple node: returns its length or -1 */
static int single(struct mctx *M, const struct node *n, int pos)
{
    int l, i;
    if (n->type == N_BYTES) {
        if (pos + n->nbytes > M->len) return -1;
        for (i = 0; i < n->nbytes; i++) {
            int a = M->s[pos + i], b = n->bytes[i];
            if (n->icase && n->nbytes == 1) { a = lower(a); b = lower(b); }
            if (a != b) return -1;
        }
        return n->nbytes;
    }
    if (pos >= M->len) return -1;
    if (n->type == N_ANY) {
        if (M->s[pos] == '\n' && !n->dotall) return -1;
        u8decode(M->s, pos, M->len, &l);
        return l;
    }
    {
        uint32_t cp = u8decode(M->s, pos, M->len, &l);
        return cc_has(n->cc, cp) ? l : -1;
    }
}

static int rep(struct mctx *M, const struct node *n, int count, int pos, const struct cont *k)
{
    struct cont c;
    int can_more = n->max < 0 || count < n->max;
    if (++M->steps > ORC_STEP_LIMIT) return -2;
    c.kind = K_REP; c.n = n; c.next = k; c.count = count + 1; c.start = pos;
    if (n->mode == 1) {                              /* lazy: leave first */
        if (count >= n->min) { int r = run(M, k, pos); if (r) return r; }
        if (can_more) return m(M, n->a, pos, &c);
        return 0;
    }
    if (can_more) { int r = m(M, n->a, pos, &c); if (r) return r; }
    if (count >= n->min) return run(M, k, pos);
    return 0;
}

static int run(struct mctx *M, const struct cont *k, int pos)
{
    if (!k) { M->end = pos; return 1; }
    switch (k->kind) {
    case K_NODE: return m(M, k->n, pos, k->next);
    case K_REP:
        /* an iteration that consumed nothing ends the loop (regexec.c NULL_CHECK_END) */
        if (pos == k->start && k->count > k->n->min) return run(M, k->next, pos);
        if (pos == k->start && k->count >= k->n->min) return run(M, k->next, pos);
        return rep(M, k->n, k->count, pos, k->next);
    case K_CLOSE: {
        int g = k->n->grp, os = M->cap[2 * g], oe = M->cap[2 * g + 1], r;
        M->cap[2 * g] = k->gstart; M->cap[2 * g + 1] = pos;
        r = run(M, k->next, pos);
        if (r <= 0) { M->cap[2 * g] = os; M->cap[2 * g + 1] = oe; }
        return r;
    }
    case K_STOP: *k->stop_pos = pos; return 1;
    }
    return 0;
}

static int m(struct mctx *M, const struct node *n, int pos, const struct cont *k)
{
    struct cont c;
    int r, l;
    if (++M->steps > ORC_STEP_LIMIT) return -2;
    switch (n->type) {
    case N_EMPTY: return run(M, k, pos);
    case N_BYTES: case N_ANY: case N_CLASS:
        l = single(M, n, pos);
        if (l < 0) return 0;
        return run(M, k, pos + l);
    case N_CAT:
        c.kind = K_NODE; c.n = n->b; c.next = k;
        return m(M, n->a, pos, &c);
    case N_ALT:
        r = m(M, n->a, pos, k);
        if (r) return r;
        return m(M, n->b, pos, k);
    case N_REP:
        if (n->mode == 2) {                          /* possessive = atomic(greedy) */
            struct node g = *n;
            struct cont stop;
            int endp = -1, *saved = alloca(sizeof(int) * 2 * (size_t) M->ngroups);
            g.mode = 0;
            memcpy(saved, M->cap, sizeof(int) * 2 * (size_t) M->ngroups);
            stop.kind = K_STOP; stop.stop_pos = &endp; stop.next = NULL; stop.n = NULL;
            r = rep(M, &g, 0, pos, &stop);
            if (r <= 0) return r;
            r = run(M, k, endp);
            if (r <= 0) memcpy(M->cap, saved, sizeof(int) * 2 * (size_t) M->ngroups);
            return r;
        }
        if ((n->a->type == N_BYTES || n->a->type == N_ANY || n->a->type == N_CLASS) && n->mode == 0) {
            /* greedy single-character loop without recursion per character */
            int cnt = 0, p = pos, stackn = 0, cap = 64, *ends = malloc(sizeof(int) * 64);
            while (n->max < 0 || cnt < n->max) {
                l = single(M, n->a, p);
                if (l < 0) break;
                if (stackn == cap) { cap *= 2; ends = realloc(ends, sizeof(int) * (size_t) cap); }
                ends[stackn++] = p;
                p += l; cnt++;
            }
            for (;;) {
                if (cnt >= n->min) { r = run(M, k, p); if (r) { free(ends); return r; } }
                if (cnt == 0 || cnt <= n->min) break;
                p = ends[--stackn]; cnt--;
                if (++M->steps > ORC_STEP_LIMIT) { free(ends); return -2; }
            }
            free(ends);
            return 0;
        }
        return rep(M, n, 0, pos, k);
    case N_GROUP:
        c.kind = K_CLOSE; c.n = n; c.next = k; c.gstart = pos;
        return m(M, n->a, pos, &c);
    case N_ATOMIC: case N_LOOK: case N_NLOOK: {
        struct cont stop;
        int endp = -1, *saved = alloca(sizeof(int) * 2 * (size_t) M->ngroups);
        memcpy(saved, M->cap, sizeof(int) * 2 * (size_t) M->ngroups);
        stop.kind = K_STOP; stop.stop_pos = &endp; stop.next = NULL; stop.n = NULL;
        r = m(M, n->a, pos, &stop);
        if (r < 0) return r;
        if (n->type == N_NLOOK) {
            memcpy(M->cap, saved, sizeof(int) * 2 * (size_t) M->ngroups);
            return r ? 0 : run(M, k, pos);
        }
        if (!r) return 0;
        r = run(M, k, n->type == N_LOOK ? pos : endp);
        if (r <= 0) memcpy(M->cap, saved, sizeof(int) * 2 * (size_t) M->ngroups);
        return r;
    }
    case N_BACKREF: {
        int g = n->grp, s, e, i;
        if (g <= 0 || g >= M->ngroups) return 0;
        s = M->cap[2 * g]; e = M->cap[2 * g + 1];
        if (s < 0 || e < 0) return 0;
        if (pos + (e - s) > M->len) return 0;
        for (i = 0; i < e - s; i++) {
            int a = M->s[s + i], b = M->s[pos + i];
            if (n->icase) { a = lower(a); b = lower(b); }
            if (a != b) return 0;
        }
        return run(M, k, pos + (e - s));
    }
    case N_BOL: if (pos == 0 || (M->s[pos - 1] == '\n' && pos != M->len)) return run(M, k, pos); return 0;
    case N_EOL: if (pos == M->len || M->s[pos] == '\n') return run(M, k, pos); return 0;
    case N_BBUF: if (pos == 0) return run(M, k, pos); return 0;
    case N_EBUF: if (pos == M->len) return run(M, k, pos); return 0;
    case N_SEBUF: if (pos == M->len || (pos == M->len - 1 && M->s[pos] == '\n')) return run(M, k, pos); return 0;
    case N_WB: case N_NWB: {
        int a = pos > 0 ? is_word_at(M, prev_char(M, pos)) : 0, b = is_word_at(M, pos);
        if ((a != b) == (n->type == N_WB)) return run(M, k, pos);
        return 0;
    }
    }
    return 0;
}

int orc_regex_search(const struct orc_regex *re, const uint8_t *s, size_t n, int *region, int max_groups)
{
    struct mctx M;
    int start, i, r, *cap = malloc(sizeof(int) * 2 * (size_t) re->ngroups);
    M.s = s; M.len = (int) n; M.cap = cap; M.ngroups = re->ngroups; M.steps = 0; M.end = -1;
    for (start = 0; start <= (int) n; ) {
        for (i = 0; i < 2 * re->ngroups; i++) cap[i] = -1;
        r = m(&M, re->root, start, NULL);
        if (r < 0) { free(cap); return -2; }
        if (r) {
            cap[0] = start; cap[1] = M.end;
            for (i = 0; i < re->ngroups && i < max_groups; i++) { region[2 * i] = cap[2 * i]; region[2 * i + 1] = cap[2 * i + 1]; }
            free(cap);
            return 1;
        }
        if (start == (int) n) break;
        { int l; u8decode(s, start, (int) n, &l); start += l; }
    }
    free(cap);
    return 0;
}
