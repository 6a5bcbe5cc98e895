/* rx_compile.c -- Ruby-syntax regex -> struct rx_prog (see rx_compile.h).
 *
 * Semantics follow Onigmo 6.2.0 as configured by Fluent Bit
 * (src/flb_regex.c:143-146: ONIG_ENCODING_UTF8 + ONIG_SYNTAX_RUBY):
 *   - only named groups capture when at least one is present
 *     (ONIG_SYN_CAPTURE_ONLY_NAMED_GROUP, lib/onigmo/regparse.c:66)
 *   - \d \s \w \h are ASCII-only (ONIG_OPTION_ASCII_RANGE, regparse.c:72)
 *   - ^ and $ are line anchors, '.' excludes '\n' unless (?m)
 *   - {n}? is "(?:x{n})?" (ONIG_SYN_FIXED_INTERVAL_IS_GREEDY_ONLY)
 *   - an invalid interval is a literal '{' (ONIG_SYN_ALLOW_INVALID_INTERVAL)
 *   - x*+ x++ x?+ are possessive, x{n,m}+ is not
 * The code generator additionally proves some greedy single-character loops
 * "possessive-safe" (no shorter repetition can ever let the continuation match)
 * and emits them without back-off state; this never changes the result.
 *
 * Host code, plain C.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdarg.h>
#include <setjmp.h>
#include "rx_compile.h"

#define RX_MAX_CODE   (1u << 16)
#define RX_REP_INF    (-1)
#define RX_MAX_REPEAT 100000   /* ONIG_MAX_REPEAT_NUM */

/* ------------------------------------------------------------ char sets */
struct cset {
    uint32_t ascii[4];
    int      hi;          /* invalid single high bytes are members */
    int      nr;          /* ranges over [0x80, 0x10FFFF], sorted, disjoint */
    int      cap;
    uint32_t *r;          /* pairs lo,hi */
};

static void cset_init(struct cset *s) { memset(s, 0, sizeof(*s)); }
static void cset_free(struct cset *s) { free(s->r); s->r = NULL; s->nr = s->cap = 0; }

static void cset_norm(struct cset *s)
{
    int i, j, n = s->nr;
    /* insertion sort by lo, then merge */
    for (i = 1; i < n; i++) {
        uint32_t lo = s->r[2 * i], hi = s->r[2 * i + 1];
        for (j = i - 1; j >= 0 && s->r[2 * j] > lo; j--) {
            s->r[2 * j + 2] = s->r[2 * j];
            s->r[2 * j + 3] = s->r[2 * j + 1];
        }
        s->r[2 * j + 2] = lo;
        s->r[2 * j + 3] = hi;
    }
    j = 0;
    for (i = 0; i < n; i++) {
        if (j > 0 && s->r[2 * i] <= s->r[2 * j - 1] + 1) {
            if (s->r[2 * i + 1] > s->r[2 * j - 1]) s->r[2 * j - 1] = s->r[2 * i + 1];
        }
        else {
            s->r[2 * j] = s->r[2 * i];
            s->r[2 * j + 1] = s->r[2 * i + 1];
            j++;
        }
    }
    s->nr = j;
}

static void cset_add_range(struct cset *s, uint32_t lo, uint32_t hi)
{
    uint32_t c;
    if (lo > hi) return;
    for (c = lo; c <= hi && c < 0x80; c++) s->ascii[c >> 5] |= 1u << (c & 31);
    if (hi < 0x80) return;
    if (lo < 0x80) lo = 0x80;
    if (hi > 0x10FFFF) hi = 0x10FFFF;
    if (s->nr == s->cap) {
        s->cap = s->cap ? s->cap * 2 : 8;
        s->r = realloc(s->r, sizeof(uint32_t) * 2 * s->cap);
    }
    s->r[2 * s->nr] = lo;
    s->r[2 * s->nr + 1] = hi;
    s->nr++;
    cset_norm(s);
}

static void cset_union(struct cset *d, const struct cset *a)
{
    int i;
    for (i = 0; i < 4; i++) d->ascii[i] |= a->ascii[i];
    d->hi |= a->hi;
    for (i = 0; i < a->nr; i++) cset_add_range(d, a->r[2 * i], a->r[2 * i + 1]);
}

static void cset_complement(struct cset *s)
{
    struct cset t;
    uint32_t next = 0x80;
    int i;
    cset_init(&t);
    for (i = 0; i < 4; i++) t.ascii[i] = ~s->ascii[i];
    t.hi = !s->hi;
    for (i = 0; i < s->nr; i++) {
        if (s->r[2 * i] > next) cset_add_range(&t, next, s->r[2 * i] - 1);
        next = s->r[2 * i + 1] + 1;
    }
    if (next <= 0x10FFFF) cset_add_range(&t, next, 0x10FFFF);
    cset_free(s);
    *s = t;
}

static void cset_intersect(struct cset *d, const struct cset *a)
{
    struct cset na;
    cset_init(&na);
    cset_union(&na, a);
    cset_complement(&na);
    cset_complement(d);
    cset_union(d, &na);
    cset_complement(d);
    cset_free(&na);
}

static void cset_copy(struct cset *d, const struct cset *a) { cset_init(d); cset_union(d, a); }

/* ------------------------------------------------------------------ AST */
enum { N_EMPTY, N_LIT, N_CLASS, N_CAT, N_ALT, N_REP, N_GROUP, N_ANCHOR, N_LOOK, N_ATOMIC, N_BACKREF };

struct node {
    int type;
    uint32_t cp;            /* N_LIT */
    int icase;              /* N_LIT / N_BACKREF */
    int cls;                /* N_CLASS: index into comp->sets */
    int min, max, greedy, poss;   /* N_REP */
    int group;              /* N_GROUP capture number (>=1) or 0 ; N_BACKREF group */
    int anchor;             /* N_ANCHOR: RX_BOL... ; N_LOOK: 1 positive, 0 negative */
    int n;                  /* children */
    struct node **kid;
};

struct comp {
    const unsigned char *p, *end;
    int icase, dotall, extend;
    int ascii_only;         /* a construct whose non-ASCII behaviour is approximated was used */
    int has_named;
    int n_groups;
    struct rx_compiled *out;
    struct cset *sets;      /* final (already negated) character sets */
    int n_sets, cap_sets;
    int set_any, set_any_nl;
    /* arena */
    struct node **all;
    int n_all, cap_all;
    /* code */
    uint32_t *code;
    int n_code, cap_code;
    int n_null;
    int depth;
    struct { int idx; uint32_t b[8]; } *raw_sets;   /* classes given as raw byte bitmaps */
    int n_raw;
    jmp_buf jb;
};

static void fail(struct comp *c, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(c->out->err, sizeof(c->out->err), fmt, ap);
    va_end(ap);
    longjmp(c->jb, 1);
}

static struct node *mk(struct comp *c, int type)
{
    struct node *n = calloc(1, sizeof(*n));
    if (c->n_all == c->cap_all) {
        c->cap_all = c->cap_all ? c->cap_all * 2 : 64;
        c->all = realloc(c->all, sizeof(*c->all) * c->cap_all);
    }
    c->all[c->n_all++] = n;
    n->type = type;
    return n;
}

static void add_kid(struct node *n, struct node *k)
{
    n->kid = realloc(n->kid, sizeof(*n->kid) * (n->n + 1));
    n->kid[n->n++] = k;
}

static int add_set(struct comp *c, struct cset *s)   /* takes ownership */
{
    if (c->n_sets == c->cap_sets) {
        c->cap_sets = c->cap_sets ? c->cap_sets * 2 : 16;
        c->sets = realloc(c->sets, sizeof(*c->sets) * c->cap_sets);
    }
    c->sets[c->n_sets] = *s;
    return c->n_sets++;
}

/* ---------------------------------------------------------------- lexing */
static int peek(struct comp *c) { return c->p < c->end ? *c->p : -1; }

static uint32_t utf8_next(struct comp *c)
{
    uint32_t b = *c->p++, cp;
    int n, i;
    if (b < 0x80) return b;
    if (b >= 0xC2 && b <= 0xDF) { n = 1; cp = b & 0x1f; }
    else if (b >= 0xE0 && b <= 0xEF) { n = 2; cp = b & 0x0f; }
    else if (b >= 0xF0 && b <= 0xF4) { n = 3; cp = b & 0x07; }
    else { fail(c, "invalid UTF-8 in pattern"); return 0; }
    for (i = 0; i < n; i++) {
        if (c->p >= c->end || (*c->p & 0xC0) != 0x80) fail(c, "invalid UTF-8 in pattern");
        cp = (cp << 6) | (*c->p++ & 0x3f);
    }
    return cp;
}

static int hexval(int ch)
{
    if (ch >= '0' && ch <= '9') return ch - '0';
    if (ch >= 'a' && ch <= 'f') return ch - 'a' + 10;
    if (ch >= 'A' && ch <= 'F') return ch - 'A' + 10;
    return -1;
}

/* ctype sets, ASCII range (ONIG_OPTION_ASCII_RANGE) */
static void set_ctype(struct cset *s, int kind)
{
    switch (kind) {
    case 'd': cset_add_range(s, '0', '9'); break;
    case 'w': cset_add_range(s, '0', '9'); cset_add_range(s, 'A', 'Z');
              cset_add_range(s, 'a', 'z'); cset_add_range(s, '_', '_'); break;
    case 's': cset_add_range(s, 9, 13); cset_add_range(s, ' ', ' '); break;
    case 'h': cset_add_range(s, '0', '9'); cset_add_range(s, 'A', 'F');
              cset_add_range(s, 'a', 'f'); break;
    }
}

/* add a (possibly negated) ctype escape to a set under construction */
static void add_ctype(struct cset *dst, int esc)
{
    struct cset t;
    int lower = esc | 0x20;
    cset_init(&t);
    set_ctype(&t, lower);
    if (esc != lower) {       /* \D \W \S \H: everything else, incl. all multibyte */
        cset_complement(&t);
    }
    cset_union(dst, &t);
    cset_free(&t);
}

static int posix_bracket(struct comp *c, struct cset *dst)
{
    c->ascii_only = 1;
    /* at "[:" ; returns 1 if consumed */
    static const char *names[] = { "alnum", "alpha", "ascii", "blank", "cntrl", "digit", "graph",
                                   "lower", "print", "punct", "space", "upper", "xdigit", "word", NULL };
    const unsigned char *q = c->p + 2;
    int neg = 0, i, ch;
    struct cset t;
    if (q < c->end && *q == '^') { neg = 1; q++; }
    for (i = 0; names[i]; i++) {
        size_t l = strlen(names[i]);
        if ((size_t) (c->end - q) >= l + 2 && !memcmp(q, names[i], l) && q[l] == ':' && q[l + 1] == ']') {
            cset_init(&t);
            for (ch = 0; ch < 128; ch++) {
                int in = 0;
                switch (i) {
                case 0: in = (ch >= '0' && ch <= '9') || (ch >= 'A' && ch <= 'Z') || (ch >= 'a' && ch <= 'z'); break;
                case 1: in = (ch >= 'A' && ch <= 'Z') || (ch >= 'a' && ch <= 'z'); break;
                case 2: in = 1; break;
                case 3: in = ch == ' ' || ch == '\t'; break;
                case 4: in = ch < 32 || ch == 127; break;
                case 5: in = ch >= '0' && ch <= '9'; break;
                case 6: in = ch > 32 && ch < 127; break;
                case 7: in = ch >= 'a' && ch <= 'z'; break;
                case 8: in = ch >= 32 && ch < 127; break;
                case 9: in = (ch > 32 && ch < 127) && !((ch >= '0' && ch <= '9') || (ch >= 'A' && ch <= 'Z') || (ch >= 'a' && ch <= 'z')); break;
                case 10: in = (ch >= 9 && ch <= 13) || ch == ' '; break;
                case 11: in = ch >= 'A' && ch <= 'Z'; break;
                case 12: in = hexval(ch) >= 0; break;
                case 13: in = (ch >= '0' && ch <= '9') || (ch >= 'A' && ch <= 'Z') || (ch >= 'a' && ch <= 'z') || ch == '_'; break;
                }
                if (in) t.ascii[ch >> 5] |= 1u << (ch & 31);
            }
            /* ONIG_OPTION_POSIX_BRACKET_ALL_RANGE would also admit non-ASCII members;
             * this build keeps brackets ASCII-only (documented limitation). */
            if (neg) cset_complement(&t);
            cset_union(dst, &t);
            cset_free(&t);
            c->p = q + l + 2;
            return 1;
        }
    }
    return 0;
}

/* parse one escape after the backslash when a single code point is expected.
 * returns code point, or -1 and sets *ctype when it is a class escape */
static long parse_escape_cp(struct comp *c, int in_class, int *ctype)
{
    int ch, v, i;
    uint32_t cp;
    *ctype = 0;
    if (c->p >= c->end) fail(c, "end pattern at escape");
    ch = *c->p;
    switch (ch) {
    case 't': c->p++; return '\t';
    case 'n': c->p++; return '\n';
    case 'r': c->p++; return '\r';
    case 'f': c->p++; return '\f';
    case 'v': c->p++; return '\v';
    case 'a': c->p++; return 7;
    case 'b': if (in_class) { c->p++; return 8; } break;
    case 'e': c->p++; return 27;
    case 'd': case 'D': case 'w': case 'W': case 's': case 'S': case 'h': case 'H':
        c->p++; *ctype = ch; return -1;
    case 'x':
        c->p++;
        if (peek(c) == '{') {
            c->p++; cp = 0; i = 0;
            while (c->p < c->end && (v = hexval(*c->p)) >= 0) { cp = cp * 16 + v; c->p++; if (++i > 8) fail(c, "too long wide-char value"); }
            if (peek(c) != '}' || i == 0) fail(c, "invalid \\x{} escape");
            c->p++;
            if (cp > 0x10FFFF) fail(c, "too big wide-char value");
            return cp;
        }
        cp = 0;
        for (i = 0; i < 2 && c->p < c->end && (v = hexval(*c->p)) >= 0; i++) { cp = cp * 16 + v; c->p++; }
        if (cp >= 0x80) fail(c, "\\x%02X: raw high bytes in a UTF-8 pattern are not supported", cp);
        return cp;
    case 'u':
        c->p++; cp = 0;
        for (i = 0; i < 4; i++) {
            if (c->p >= c->end || (v = hexval(*c->p)) < 0) fail(c, "too short escape sequence");
            cp = cp * 16 + v; c->p++;
        }
        return cp;
    case '0': case '1': case '2': case '3': case '4': case '5': case '6': case '7':
        if (!in_class && ch != '0') {
            /* \1..\9: numbered back reference */
            return -2;
        }
        cp = 0;
        for (i = 0; i < 3 && c->p < c->end && *c->p >= '0' && *c->p <= '7'; i++) cp = cp * 8 + (*c->p++ - '0');
        if (cp >= 0x80) fail(c, "octal escape above 0x7f is not supported");
        return cp;
    case '8': case '9':
        if (!in_class) return -2;
        c->p++; return ch;
    case 'c': case 'C': case 'M':
        fail(c, "control/meta escapes (\\c \\C- \\M-) are not supported");
    case 'p': case 'P':
        fail(c, "\\p{...} character properties are not supported");
    default:
        break;
    }
    /* any other escaped character is itself */
    return utf8_next(c);
}

static void add_cp_icase(struct comp *c, struct cset *s, uint32_t lo, uint32_t hi)
{
    uint32_t ch;
    cset_add_range(s, lo, hi);
    if (!c->icase) return;
    c->ascii_only = 1;            /* Onigmo folds through its Unicode tables (multi-character folds included): ASCII subjects only */
    for (ch = lo; ch <= hi && ch < 128; ch++) {
        if (ch >= 'a' && ch <= 'z') cset_add_range(s, ch - 32, ch - 32);
        if (ch >= 'A' && ch <= 'Z') cset_add_range(s, ch + 32, ch + 32);
        /* Unicode simple folds into ASCII: U+212A KELVIN SIGN ~ k, U+017F LONG S ~ s */
        if (ch == 'k' || ch == 'K') cset_add_range(s, 0x212A, 0x212A);
        if (ch == 's' || ch == 'S') cset_add_range(s, 0x017F, 0x017F);
    }
    if (lo <= 0x212A && hi >= 0x212A) { cset_add_range(s, 'k', 'k'); cset_add_range(s, 'K', 'K'); }
    if (lo <= 0x017F && hi >= 0x017F) { cset_add_range(s, 's', 's'); cset_add_range(s, 'S', 'S'); }
}

/* parse "[...]" (c->p just after '['); result is the final member set */
static void parse_class(struct comp *c, struct cset *out)
{
    struct cset acc, cur;
    int neg = 0, first = 1, have_and = 0, ctype;
    long lo, hi;

    cset_init(&acc);
    cset_init(&cur);
    if (++c->depth > 64) fail(c, "too deep nesting");
    if (peek(c) == '^') { neg = 1; c->p++; }
    for (;;) {
        int ch;
        if (c->p >= c->end) fail(c, "premature end of char-class");
        ch = *c->p;
        if (ch == ']' && !first) { c->p++; break; }
        first = 0;
        if (ch == '[') {
            if (c->p + 1 < c->end && c->p[1] == ':' && posix_bracket(c, &cur)) continue;
            {
                struct cset sub;
                c->p++;
                cset_init(&sub);
                parse_class(c, &sub);
                cset_union(&cur, &sub);
                cset_free(&sub);
            }
            continue;
        }
        if (ch == '&' && c->p + 1 < c->end && c->p[1] == '&') {
            c->p += 2;
            if (have_and) cset_intersect(&acc, &cur);
            else { cset_free(&acc); cset_copy(&acc, &cur); }
            have_and = 1;
            cset_free(&cur);
            cset_init(&cur);
            continue;
        }
        if (ch == '\\') {
            c->p++;
            lo = parse_escape_cp(c, 1, &ctype);
            if (lo == -1) { add_ctype(&cur, ctype); continue; }
        }
        else {
            lo = utf8_next(c);
        }
        hi = lo;
        if (peek(c) == '-' && c->p + 1 < c->end && c->p[1] != ']') {
            const unsigned char *save = c->p;
            c->p++;
            if (peek(c) == '[' || (peek(c) == '&' && c->p + 1 < c->end && c->p[1] == '&')) {
                c->p = save;              /* "a-[" : '-' is a literal */
            }
            else if (peek(c) == '\\') {
                c->p++;
                hi = parse_escape_cp(c, 1, &ctype);
                if (hi == -1) {           /* [a-\d] : literal 'a', '-', then the ctype */
                    add_cp_icase(c, &cur, lo, lo);
                    add_cp_icase(c, &cur, '-', '-');
                    add_ctype(&cur, ctype);
                    continue;
                }
                if (hi < lo) fail(c, "empty range in char class");
            }
            else {
                hi = utf8_next(c);
                if (hi < lo) fail(c, "empty range in char class");
            }
        }
        add_cp_icase(c, &cur, (uint32_t) lo, (uint32_t) hi);
    }
    if (have_and) { cset_intersect(&acc, &cur); cset_free(&cur); cur = acc; cset_init(&acc); }
    if (neg) cset_complement(&cur);
    cset_free(&acc);
    c->depth--;
    *out = cur;
}

/* --------------------------------------------------------------- parsing */
static struct node *parse_alt(struct comp *c);

static struct node *lit_node(struct comp *c, uint32_t cp)
{
    struct node *n;
    if (c->icase && ((cp >= 'a' && cp <= 'z') || (cp >= 'A' && cp <= 'Z') || cp == 0x212A || cp == 0x017F)) {
        struct cset s;
        cset_init(&s);
        add_cp_icase(c, &s, cp, cp);
        n = mk(c, N_CLASS);
        n->cls = add_set(c, &s);
        return n;
    }
    if (c->icase && cp >= 0x80) c->ascii_only = 1;      /* its case partners are not restated */
    n = mk(c, N_LIT);
    n->cp = cp;
    return n;
}

static struct node *class_node_from_ctype(struct comp *c, int ctype)
{
    struct cset s;
    struct node *n = mk(c, N_CLASS);
    cset_init(&s);
    add_ctype(&s, ctype);
    n->cls = add_set(c, &s);
    return n;
}

static int name_char(int ch)
{
    return (ch >= 'a' && ch <= 'z') || (ch >= 'A' && ch <= 'Z') || (ch >= '0' && ch <= '9') || ch == '_' || ch >= 0x80;
}

static void register_name(struct comp *c, const unsigned char *s, int len, int group)
{
    struct rx_compiled *o = c->out;
    int i;
    for (i = 0; i < o->n_names; i++) {
        if ((int) strlen(o->names[i].name) == len && !memcmp(o->names[i].name, s, len)) {
            if (o->names[i].n_groups >= 8) fail(c, "too many definitions of one group name");
            o->names[i].groups[o->names[i].n_groups++] = group;
            return;
        }
    }
    o->names = realloc(o->names, sizeof(*o->names) * (o->n_names + 1));
    memset(&o->names[o->n_names], 0, sizeof(*o->names));
    o->names[o->n_names].name = malloc(len + 1);
    memcpy(o->names[o->n_names].name, s, len);
    o->names[o->n_names].name[len] = 0;
    o->names[o->n_names].groups[0] = group;
    o->names[o->n_names].n_groups = 1;
    o->n_names++;
}

static int lookup_name(struct comp *c, const unsigned char *s, int len)
{
    struct rx_compiled *o = c->out;
    int i;
    for (i = 0; i < o->n_names; i++)
        if ((int) strlen(o->names[i].name) == len && !memcmp(o->names[i].name, s, len)) return i;
    return -1;
}

/* (?imx-imx) / (?imx-imx:...) ; c->p just after "(?" at the first option letter */
static struct node *parse_option_group(struct comp *c)
{
    int on = 1, icase = c->icase, dotall = c->dotall, extend = c->extend;
    for (;;) {
        int ch = peek(c);
        if (ch < 0) fail(c, "end pattern in group");
        c->p++;
        if (ch == '-') { on = 0; continue; }
        if (ch == 'i') { icase = on; continue; }
        if (ch == 'm') { dotall = on; continue; }
        if (ch == 'x') { extend = on; continue; }
        if (ch == 'a') { continue; }                  /* ASCII range: already the default */
        if (ch == ')') {                              /* applies to the rest of the enclosing group */
            struct node *n;
            c->icase = icase; c->dotall = dotall; c->extend = extend;
            n = parse_alt(c);
            return n;                                 /* caller sees ')' of the enclosing group / end */
        }
        if (ch == ':') {
            int si = c->icase, sd = c->dotall, sx = c->extend;
            struct node *n;
            c->icase = icase; c->dotall = dotall; c->extend = extend;
            n = parse_alt(c);
            if (peek(c) != ')') fail(c, "end pattern with unmatched parenthesis");
            c->p++;
            c->icase = si; c->dotall = sd; c->extend = sx;
            return n;
        }
        fail(c, "undefined group option");
    }
}

static struct node *parse_group(struct comp *c)
{
    /* c->p just after '(' */
    struct node *n, *sub;
    int si = c->icase, sd = c->dotall, sx = c->extend;

    if (++c->depth > 64) fail(c, "too deep nesting");
    if (peek(c) == '?') {
        int ch;
        c->p++;
        ch = peek(c);
        if (ch == ':') {
            c->p++;
            sub = parse_alt(c);
            n = sub;
        }
        else if (ch == '=' || ch == '!') {
            c->p++;
            n = mk(c, N_LOOK);
            n->anchor = (ch == '=');
            add_kid(n, parse_alt(c));
        }
        else if (ch == '>') {
            c->p++;
            n = mk(c, N_ATOMIC);
            add_kid(n, parse_alt(c));
        }
        else if (ch == '#') {
            while (c->p < c->end && *c->p != ')') { if (*c->p == '\\') c->p++; c->p++; }
            if (c->p >= c->end) fail(c, "end pattern in group");
            c->p++;
            c->depth--;
            return mk(c, N_EMPTY);
        }
        else if (ch == '<' || ch == '\'') {
            int close = ch == '<' ? '>' : '\'';
            const unsigned char *ns;
            c->p++;
            if (ch == '<' && (peek(c) == '=' || peek(c) == '!')) fail(c, "look-behind is not supported");
            ns = c->p;
            while (c->p < c->end && *c->p != close) {
                if (!name_char(*c->p)) fail(c, "invalid group name");
                c->p++;
            }
            if (c->p >= c->end || c->p == ns) fail(c, "group name is empty");
            if (*ns >= '0' && *ns <= '9') fail(c, "invalid group name");
            n = mk(c, N_GROUP);
            n->group = ++c->n_groups;
            if (n->group > RX_MAX_GROUPS) fail(c, "too many capture groups (max %d)", RX_MAX_GROUPS);
            register_name(c, ns, (int) (c->p - ns), n->group);
            c->p++;
            add_kid(n, parse_alt(c));
        }
        else if (ch == '~') fail(c, "absent operator (?~...) is not supported");
        else if (ch == '(') fail(c, "conditional groups (?(cond)...) are not supported");
        else if (ch == '^' ) fail(c, "undefined group option");
        else {
            /* "(?i:...)" consumed its own ')' and restored the options; "(?i)" parsed
             * the rest of the ENCLOSING group body (Onigmo: option applies up to the
             * enclosing group's end, alternations included) and left its ')' alone */
            n = parse_option_group(c);
            c->depth--;
            return n;
        }
    }
    else {
        if (c->has_named) {
            n = parse_alt(c);                 /* plain group: not captured */
        }
        else {
            n = mk(c, N_GROUP);
            n->group = ++c->n_groups;
            if (n->group > RX_MAX_GROUPS) fail(c, "too many capture groups (max %d)", RX_MAX_GROUPS);
            add_kid(n, parse_alt(c));
        }
    }
    if (peek(c) != ')') fail(c, "end pattern with unmatched parenthesis");
    c->p++;
    c->icase = si; c->dotall = sd; c->extend = sx;
    c->depth--;
    return n;
}

static void skip_extended(struct comp *c)
{
    while (c->extend && c->p < c->end) {
        int ch = *c->p;
        if (ch == ' ' || ch == '\t' || ch == '\n' || ch == '\r' || ch == '\f' || ch == '\v') c->p++;
        else if (ch == '#') { while (c->p < c->end && *c->p != '\n') c->p++; }
        else break;
    }
}

static struct node *parse_atom(struct comp *c)
{
    int ch, ctype;
    long cp;
    struct node *n;

    ch = peek(c);
    switch (ch) {
    case '(':
        c->p++;
        return parse_group(c);
    case '[': {
        struct cset s;
        c->p++;
        parse_class(c, &s);
        n = mk(c, N_CLASS);
        n->cls = add_set(c, &s);
        return n;
    }
    case '.':
        c->p++;
        n = mk(c, N_CLASS);
        n->cls = c->dotall ? c->set_any_nl : c->set_any;
        return n;
    case '^': c->p++; n = mk(c, N_ANCHOR); n->anchor = RX_BOL; return n;
    case '$': c->p++; n = mk(c, N_ANCHOR); n->anchor = RX_EOL; return n;
    case '\\':
        c->p++;
        if (c->p >= c->end) fail(c, "end pattern at escape");
        switch (*c->p) {
        case 'A': c->p++; n = mk(c, N_ANCHOR); n->anchor = RX_BEGIN_BUF; return n;
        case 'z': c->p++; n = mk(c, N_ANCHOR); n->anchor = RX_END_BUF; return n;
        case 'Z': c->p++; n = mk(c, N_ANCHOR); n->anchor = RX_SEMI_END_BUF; return n;
        case 'b': c->p++; n = mk(c, N_ANCHOR); n->anchor = RX_WORD_B; c->ascii_only = 1; return n;
        case 'B': c->p++; n = mk(c, N_ANCHOR); n->anchor = RX_NOT_WORD_B; c->ascii_only = 1; return n;
        case 'G': fail(c, "\\G is not supported");
        case 'K': fail(c, "\\K is not supported");
        case 'R': fail(c, "\\R is not supported");
        case 'X': fail(c, "\\X is not supported");
        case 'g': fail(c, "subexpression calls \\g<...> are not supported");
        case 'k': {
            const unsigned char *ns;
            int close, idx;
            c->p++;
            if (peek(c) != '<' && peek(c) != '\'') fail(c, "invalid backref");
            close = *c->p == '<' ? '>' : '\'';
            c->p++;
            ns = c->p;
            while (c->p < c->end && *c->p != close) c->p++;
            if (c->p >= c->end) fail(c, "invalid backref");
            idx = lookup_name(c, ns, (int) (c->p - ns));
            if (idx < 0) fail(c, "undefined name reference");
            if (c->out->names[idx].n_groups != 1) fail(c, "back reference to a multiply defined name is not supported");
            if (c->icase) fail(c, "case-insensitive back references are not supported");
            c->p++;
            n = mk(c, N_BACKREF);
            n->group = c->out->names[idx].groups[0];
            return n;
        }
        default:
            break;
        }
        cp = parse_escape_cp(c, 0, &ctype);
        if (cp == -1) return class_node_from_ctype(c, ctype);
        if (cp == -2) {
            int g = 0;
            if (c->has_named) fail(c, "numbered backref/call is not allowed. (use name)");
            while (c->p < c->end && *c->p >= '0' && *c->p <= '9') g = g * 10 + (*c->p++ - '0');
            if (g < 1 || g > c->n_groups) fail(c, "invalid backref number");
            if (c->icase) fail(c, "case-insensitive back references are not supported");
            n = mk(c, N_BACKREF);
            n->group = g;
            return n;
        }
        return lit_node(c, (uint32_t) cp);
    default:
        return lit_node(c, utf8_next(c));
    }
}

/* try to read {n}, {n,}, {n,m}, {,m}; returns 1 and advances on success */
static int parse_interval(struct comp *c, int *min, int *max, int *fixed)
{
    const unsigned char *q = c->p + 1;
    long lo = 0, hi = 0;
    int have_lo = 0, have_hi = 0, comma = 0;
    while (q < c->end && *q >= '0' && *q <= '9') { lo = lo * 10 + (*q++ - '0'); have_lo = 1; if (lo > RX_MAX_REPEAT) fail(c, "too big number for repeat range"); }
    if (q < c->end && *q == ',') {
        comma = 1; q++;
        while (q < c->end && *q >= '0' && *q <= '9') { hi = hi * 10 + (*q++ - '0'); have_hi = 1; if (hi > RX_MAX_REPEAT) fail(c, "too big number for repeat range"); }
    }
    if (q >= c->end || *q != '}') return 0;
    if (!have_lo && !have_hi) return 0;
    if (!comma) { *min = (int) lo; *max = (int) lo; *fixed = 1; }
    else {
        *fixed = 0;
        *min = have_lo ? (int) lo : 0;
        *max = have_hi ? (int) hi : RX_REP_INF;
        if (have_hi && have_lo && hi < lo) fail(c, "too big wide range");  /* Onigmo swaps for some syntaxes; Ruby errors */
    }
    c->p = q + 1;
    return 1;
}

static struct node *parse_quantified(struct comp *c)
{
    struct node *atom = parse_atom(c);
    for (;;) {
        int ch, min, max, fixed = 0, greedy = 1, poss = 0, is_interval = 0;
        struct node *r;
        skip_extended(c);
        ch = peek(c);
        if (ch == '*') { min = 0; max = RX_REP_INF; c->p++; }
        else if (ch == '+') { min = 1; max = RX_REP_INF; c->p++; }
        else if (ch == '?') { min = 0; max = 1; c->p++; }
        else if (ch == '{') {
            if (!parse_interval(c, &min, &max, &fixed)) return atom;   /* literal '{' next */
            is_interval = 1;
        }
        else return atom;
        if (atom->type == N_ANCHOR || atom->type == N_LOOK)
            fail(c, "target of repeat operator is invalid");
        if (peek(c) == '?' && !(is_interval && fixed)) { greedy = 0; c->p++; }
        else if (peek(c) == '+' && !is_interval) { poss = 1; c->p++; }
        r = mk(c, N_REP);
        r->min = min; r->max = max; r->greedy = greedy; r->poss = poss;
        add_kid(r, atom);
        atom = r;
    }
}

static struct node *parse_cat(struct comp *c)
{
    struct node *cat = mk(c, N_CAT);
    for (;;) {
        int ch;
        skip_extended(c);
        ch = peek(c);
        if (ch < 0 || ch == '|' || ch == ')') break;
        if (ch == '*' || ch == '+' || ch == '?') fail(c, "target of repeat operator is not specified");
        add_kid(cat, parse_quantified(c));
    }
    return cat;
}

static struct node *parse_alt(struct comp *c)
{
    struct node *first = parse_cat(c), *alt;
    if (peek(c) != '|') return first;
    alt = mk(c, N_ALT);
    add_kid(alt, first);
    while (peek(c) == '|') {
        c->p++;
        add_kid(alt, parse_cat(c));
    }
    return alt;
}

/* ------------------------------------------------------------- analysis */
struct fset { uint32_t b[8]; int wild; };

static void fs_or(struct fset *d, const struct fset *s)
{
    int i;
    for (i = 0; i < 8; i++) d->b[i] |= s->b[i];
    d->wild |= s->wild;
}

static void cls_first_bytes(const struct cset *s, struct fset *f)
{
    int i;
    for (i = 0; i < 4; i++) f->b[i] |= s->ascii[i];
    if (s->hi || s->nr) for (i = 4; i < 8; i++) f->b[i] = 0xffffffffu;
}

static int utf8_enc(uint32_t cp, unsigned char *o)
{
    if (cp < 0x80) { o[0] = (unsigned char) cp; return 1; }
    if (cp < 0x800) { o[0] = 0xC0 | (cp >> 6); o[1] = 0x80 | (cp & 0x3f); return 2; }
    if (cp < 0x10000) { o[0] = 0xE0 | (cp >> 12); o[1] = 0x80 | ((cp >> 6) & 0x3f); o[2] = 0x80 | (cp & 0x3f); return 3; }
    o[0] = 0xF0 | (cp >> 18); o[1] = 0x80 | ((cp >> 12) & 0x3f); o[2] = 0x80 | ((cp >> 6) & 0x3f); o[3] = 0x80 | (cp & 0x3f);
    return 4;
}

/* Returns: bit0 = may pass with zero width (possibly conditionally),
 *          bit1 = passes with zero width unconditionally.
 * look!=0: zero-width assertions contribute the bytes they may inspect. */
static int first_of(struct comp *c, struct node *n, struct fset *f, int look)
{
    int i, r, acc;
    unsigned char u[4];
    switch (n->type) {
    case N_EMPTY: return 3;
    case N_LIT:
        utf8_enc(n->cp, u);
        f->b[u[0] >> 5] |= 1u << (u[0] & 31);
        return 0;
    case N_CLASS:
        cls_first_bytes(&c->sets[n->cls], f);
        return 0;
    case N_CAT:
        for (i = 0, acc = 3; i < n->n; i++) {
            r = first_of(c, n->kid[i], f, look);
            acc &= r;
            if (!(r & 1)) return 0;
        }
        return acc;
    case N_ALT:
        for (i = 0, acc = 0; i < n->n; i++) acc |= first_of(c, n->kid[i], f, look);
        return acc;
    case N_REP:
        r = first_of(c, n->kid[0], f, look);
        if (n->min == 0) return 3;
        return r;
    case N_GROUP:
    case N_ATOMIC:
        return first_of(c, n->kid[0], f, look);
    case N_ANCHOR:
        if (look) {
            if (n->anchor == RX_EOL || n->anchor == RX_SEMI_END_BUF) f->b[0] |= 1u << '\n';
            else if (n->anchor == RX_END_BUF) { /* can only hold at the very end */ }
            else f->wild = 1;
        }
        return 1;
    case N_LOOK:
        if (look) {
            if (n->anchor) { r = first_of(c, n->kid[0], f, 1); if (r & 1) f->wild = 1; }
            else f->wild = 1;
        }
        return 1;
    case N_BACKREF:
        f->wild = 1;
        return 1;
    }
    return 0;
}

/* the pattern as "exactly one character, then something that consumes at least one more byte": the bytes that something can
 * begin with (rx_prog.second).  Returns 1 and fills f, or 0 when the pattern does not have that shape. */
static int second_of(struct comp *c, struct node *root, struct fset *f)
{
    struct node *n = root, *e0;
    int i, r, acc = 3;
    while ((n->type == N_GROUP || n->type == N_ATOMIC) && n->n == 1) n = n->kid[0];
    if (n->type != N_CAT || n->n < 2) return 0;
    e0 = n->kid[0];
    while ((e0->type == N_GROUP || e0->type == N_ATOMIC) && e0->n == 1) e0 = e0->kid[0];
    if (e0->type != N_CLASS && e0->type != N_LIT) return 0;
    memset(f, 0, sizeof(*f));
    for (i = 1; i < n->n; i++) {
        r = first_of(c, n->kid[i], f, 0);
        acc &= r;
        if (!(r & 1)) break;
    }
    if (i == n->n && (acc & 1)) return 0;          /* the rest can be empty: a match may end right behind the first character */
    if (f->wild) return 0;
    return 1;
}

static int contains_capture(struct node *n)
{
    int i;
    if (n->type == N_GROUP && n->group > 0) return 1;
    for (i = 0; i < n->n; i++) if (contains_capture(n->kid[i])) return 1;
    return 0;
}

static int leading_anchor(struct node *n)
{
    int i, a;
    switch (n->type) {
    case N_ANCHOR: return n->anchor;
    case N_CAT: return n->n ? leading_anchor(n->kid[0]) : 0;
    case N_GROUP: case N_ATOMIC: return leading_anchor(n->kid[0]);
    case N_REP: return n->min >= 1 ? leading_anchor(n->kid[0]) : 0;
    case N_ALT:
        a = leading_anchor(n->kid[0]);
        for (i = 1; i < n->n; i++) if (leading_anchor(n->kid[i]) != a) return 0;
        return a;
    }
    return 0;
}

/* ----------------------------------------------------------------- emit */
struct follow { struct fset f; int sure; };   /* sure: continuation reaches MATCH unconditionally with zero width */

static int emit_word(struct comp *c, uint32_t w)
{
    if (c->n_code >= (int) RX_MAX_CODE) fail(c, "regex program too large");
    if (c->n_code == c->cap_code) {
        c->cap_code = c->cap_code ? c->cap_code * 2 : 256;
        c->code = realloc(c->code, sizeof(uint32_t) * c->cap_code);
    }
    c->code[c->n_code] = w;
    return c->n_code++;
}

static void patch(struct comp *c, int at, int op, int target) { c->code[at] = RX_MK(op, target); }

static void emit_node(struct comp *c, struct node *n, const struct follow *fo);

static int single_class(struct comp *c, struct node *n)
{
    if (n->type == N_CLASS) return n->cls;
    if (n->type == N_LIT) {
        struct cset s;
        cset_init(&s);
        cset_add_range(&s, n->cp, n->cp);
        return add_set(c, &s);
    }
    return -1;
}

static void follow_of_seq(struct comp *c, struct node **kids, int nk, const struct follow *fo, struct follow *out)
{
    int i, r = 3;
    memset(out, 0, sizeof(*out));
    for (i = 0; i < nk; i++) {
        r = first_of(c, kids[i], &out->f, 1);
        if (!(r & 1)) return;          /* something must be consumed: stop */
        if (!(r & 2)) {                /* conditional pass: continue, but no longer sure */
            struct follow rest;
            follow_of_seq(c, kids + i + 1, nk - i - 1, fo, &rest);
            fs_or(&out->f, &rest.f);
            out->sure = 0;
            return;
        }
    }
    fs_or(&out->f, &fo->f);
    out->sure = fo->sure;
}

static void emit_cat(struct comp *c, struct node *n, const struct follow *fo)
{
    int i = 0;
    while (i < n->n) {
        struct follow sub;
        /* merge literal runs */
        if (n->kid[i]->type == N_LIT) {
            unsigned char buf[256];
            int len = 0, j = i;
            while (j < n->n && n->kid[j]->type == N_LIT && len + 4 <= (int) sizeof(buf))
                len += utf8_enc(n->kid[j++]->cp, buf + len);
            if (len == 1) emit_word(c, RX_MK(RX_BYTE, buf[0]));
            else {
                int k;
                emit_word(c, RX_MK(RX_STR, len));
                for (k = 0; k < len; k += 4) {
                    uint32_t w = 0;
                    int m;
                    for (m = 0; m < 4 && k + m < len; m++) w |= (uint32_t) buf[k + m] << (8 * m);
                    emit_word(c, w);
                }
            }
            i = j;
            continue;
        }
        follow_of_seq(c, n->kid + i + 1, n->n - i - 1, fo, &sub);
        emit_node(c, n->kid[i], &sub);
        i++;
    }
}

static void emit_rep(struct comp *c, struct node *n, const struct follow *fo)
{
    struct node *kid = n->kid[0];
    int cls = single_class(c, kid);
    int i, min = n->min, max = n->max;
    struct fset kf;
    int knull;
    struct follow again;      /* after a body: maybe another body, maybe the exit */
    struct follow must;       /* after a body that MUST be followed by another body */

    memset(&kf, 0, sizeof(kf));
    knull = first_of(c, kid, &kf, 1);

    if (n->poss) {
        /* x*+ == (?>x*) */
        struct node a, r, *ak[1];
        r = *n; r.poss = 0;
        memset(&a, 0, sizeof(a));
        a.type = N_ATOMIC; a.n = 1; ak[0] = &r; a.kid = ak;
        emit_node(c, &a, fo);
        return;
    }
    if ((knull & 1) && contains_capture(kid) && (max == RX_REP_INF || max > 1))
        fail(c, "a repeated group that can match the empty string and contains captures is not supported");
    if ((long) min + (max == RX_REP_INF ? 1 : (long) max - min) > 1024)
        fail(c, "repeat count too large for the device matcher");

    again = *fo;
    fs_or(&again.f, &kf);
    must = again;
    must.sure = 0;

    /* mandatory copies */
    for (i = 0; i < min; i++) {
        const struct follow *f2;
        if (i + 1 < min) f2 = &must;
        else if (max == RX_REP_INF || max > min) f2 = &again;
        else f2 = fo;
        emit_node(c, kid, f2);
    }
    if (max == RX_REP_INF) {
        int l1, id = -1;
        if (cls >= 0) {
            struct fset cf;
            int k, disjoint = 1;
            if (!n->greedy) {
                /* lazy cls*?: second word names the bytes that can start the continuation
                 * (byte-level class, bits only) so the matcher can skip hopeless positions */
                uint32_t stop = 0;
                if (!fo->sure && !fo->f.wild) {
                    struct cset st;
                    int sidx;
                    cset_init(&st);
                    sidx = add_set(c, &st);
                    c->raw_sets = realloc(c->raw_sets, sizeof(*c->raw_sets) * (c->n_raw + 1));
                    c->raw_sets[c->n_raw].idx = sidx;
                    memcpy(c->raw_sets[c->n_raw].b, fo->f.b, sizeof(fo->f.b));
                    c->n_raw++;
                    stop = (uint32_t) sidx + 1;
                }
                emit_word(c, RX_MK(RX_CSTAR_LAZY, cls));
                emit_word(c, stop);
                return;
            }
            memset(&cf, 0, sizeof(cf));
            cls_first_bytes(&c->sets[cls], &cf);
            for (k = 0; k < 8; k++) if (cf.b[k] & fo->f.b[k]) disjoint = 0;
            if (fo->sure || (disjoint && !fo->f.wild)) emit_word(c, RX_MK(RX_CSTAR_POSS, cls));
            else emit_word(c, RX_MK(RX_CSTAR_BT, cls));
            return;
        }
        /* L1: SPLIT/SPLIT_LAZY Lend ; [NULL_START] body [NULL_END] ; JMP L1 ; Lend: */
        l1 = emit_word(c, 0);
        if (knull & 1) { id = c->n_null++; emit_word(c, RX_MK(RX_NULL_START, id)); }
        emit_node(c, kid, &again);
        if (id >= 0) emit_word(c, RX_MK(RX_NULL_END, id));
        emit_word(c, RX_MK(RX_JMP, l1));
        patch(c, l1, n->greedy ? RX_SPLIT : RX_SPLIT_LAZY, c->n_code);
        return;
    }
    /* bounded optional copies: (x(x(x)?)?)? with one common exit */
    if (max > min) {
        int cnt = max - min, *fix = malloc(sizeof(int) * cnt), k;
        for (k = 0; k < cnt; k++) {
            fix[k] = emit_word(c, 0);
            emit_node(c, kid, (k + 1 < cnt) ? &again : fo);
        }
        for (k = 0; k < cnt; k++) patch(c, fix[k], n->greedy ? RX_SPLIT : RX_SPLIT_LAZY, c->n_code);
        free(fix);
    }
}

static void emit_node(struct comp *c, struct node *n, const struct follow *fo)
{
    int i;
    switch (n->type) {
    case N_EMPTY: return;
    case N_LIT: {
        unsigned char u[4];
        int len = utf8_enc(n->cp, u), k;
        if (len == 1) { emit_word(c, RX_MK(RX_BYTE, u[0])); return; }
        emit_word(c, RX_MK(RX_STR, len));
        {
            uint32_t w = 0;
            for (k = 0; k < len; k++) w |= (uint32_t) u[k] << (8 * k);
            emit_word(c, w);
        }
        return;
    }
    case N_CLASS: emit_word(c, RX_MK(RX_CLASS, n->cls)); return;
    case N_CAT: emit_cat(c, n, fo); return;
    case N_ALT: {
        int *jmps = malloc(sizeof(int) * n->n);
        for (i = 0; i < n->n; i++) {
            int sp = -1;
            if (i + 1 < n->n) sp = emit_word(c, 0);
            emit_node(c, n->kid[i], fo);
            if (i + 1 < n->n) {
                jmps[i] = emit_word(c, 0);
                patch(c, sp, RX_SPLIT, c->n_code);
            }
        }
        for (i = 0; i + 1 < n->n; i++) patch(c, jmps[i], RX_JMP, c->n_code);
        free(jmps);
        return;
    }
    case N_REP: emit_rep(c, n, fo); return;
    case N_GROUP:
        if (n->group > 0) emit_word(c, RX_MK(RX_SAVE, 2 * n->group));
        emit_node(c, n->kid[0], fo);
        if (n->group > 0) emit_word(c, RX_MK(RX_SAVE, 2 * n->group + 1));
        return;
    case N_ANCHOR: emit_word(c, RX_MK(n->anchor, 0)); return;
    case N_LOOK: {
        struct follow inner;
        memset(&inner, 0, sizeof(inner));
        inner.sure = 1;
        if (n->anchor) {
            emit_word(c, RX_MK(RX_MARK, 0));
            emit_node(c, n->kid[0], &inner);
            emit_word(c, RX_MK(RX_CUT_POS, 0));
        }
        else {
            int sp = emit_word(c, 0);
            emit_word(c, RX_MK(RX_MARK, 1));
            emit_node(c, n->kid[0], &inner);
            emit_word(c, RX_MK(RX_CUT_NEG, 0));
            patch(c, sp, RX_SPLIT, c->n_code);
        }
        return;
    }
    case N_ATOMIC: {
        struct follow inner;
        memset(&inner, 0, sizeof(inner));
        inner.sure = 1;
        emit_word(c, RX_MK(RX_MARK, 2));
        emit_node(c, n->kid[0], &inner);
        emit_word(c, RX_MK(RX_CUT_ATOMIC, 0));
        return;
    }
    case N_BACKREF: emit_word(c, RX_MK(RX_BACKREF, n->group)); return;
    }
}

/* ----------------------------------------------------------- front door */
static int prescan_named(const unsigned char *p, const unsigned char *end)
{
    int in_class = 0;
    for (; p < end; p++) {
        if (*p == '\\') { p++; continue; }
        if (in_class) { if (*p == ']') in_class--; else if (*p == '[') in_class++; continue; }
        if (*p == '[') { in_class = 1; if (p + 1 < end && p[1] == '^') p++; if (p + 1 < end && p[1] == ']') p++; continue; }
        if (*p == '(' && p + 2 < end && p[1] == '?') {
            if (p[2] == '\'') return 1;
            if (p[2] == '<' && p + 3 < end && p[3] != '=' && p[3] != '!') return 1;
        }
    }
    return 0;
}

void rx_compiled_free(struct rx_compiled *c)
{
    int i;
    if (!c) return;
    free(c->prog);
    for (i = 0; i < c->n_names; i++) free(c->names[i].name);
    free(c->names);
    c->prog = NULL; c->names = NULL; c->n_names = 0;
}

int rx_compile(const char *pattern, struct rx_compiled *out)
{
    struct comp c;
    struct node *root = NULL;
    size_t len = strlen(pattern);
    const char *start = pattern, *end = pattern + len;
    int i, ok = 0;
    volatile int opt_i = 0, opt_m = 0, opt_x = 0;

    memset(out, 0, sizeof(*out));
    memset(&c, 0, sizeof(c));
    c.out = out;

    /* /pat/imx handling, mirroring check_option()+str_to_regex() (src/flb_regex.c:60-152) */
    {
        const char *new_end = NULL;
        if (len > 0 && start[0] == '/') {
            const char *chr = strrchr(start, '/');
            if (chr && chr != start && chr != end) {
                const char *q = chr + 1;
                int oi = 0, om = 0, ox = 0, bad = 0, any = 0;
                for (; q != end && *q; q++) {
                    if (*q == 'm') { om = 1; any = 1; }
                    else if (*q == 'i') { oi = 1; any = 1; }
                    else if (*q == 'x') { ox = 1; any = 1; }
                    else if (*q == 'o') { /* accepted, ignored */ }
                    else { bad = 1; break; }
                }
                if (!bad && any) { new_end = chr; opt_i = oi; opt_m = om; opt_x = ox; }
            }
        }
        if (len > 0 && pattern[0] == '/' && pattern[len - 1] == '/') { start++; end--; }
        if (new_end != NULL) { start = pattern + 1; end = new_end; }
        if (end < start) end = start;
    }

    if (setjmp(c.jb)) goto done;

    c.p = (const unsigned char *) start;
    c.end = (const unsigned char *) end;
    c.icase = opt_i; c.dotall = opt_m; c.extend = opt_x;
    c.has_named = prescan_named(c.p, c.end);
    {
        struct cset s;
        cset_init(&s); cset_add_range(&s, 0, 0x10FFFF); s.hi = 1;
        s.ascii[0] &= ~(1u << '\n');
        c.set_any = add_set(&c, &s);
        cset_init(&s); cset_add_range(&s, 0, 0x10FFFF); s.hi = 1;
        c.set_any_nl = add_set(&c, &s);
    }
    root = parse_alt(&c);
    if (c.p < c.end) {
        if (*c.p == ')') fail(&c, "unmatched close parenthesis");
        fail(&c, "syntax error");
    }
    {
        struct follow top;
        struct fset ff;
        int r, la;
        memset(&top, 0, sizeof(top));
        top.sure = 1;
        emit_word(&c, RX_MK(RX_SAVE, 0));
        emit_node(&c, root, &top);
        emit_word(&c, RX_MK(RX_SAVE, 1));
        emit_word(&c, RX_MK(RX_MATCH, 0));

        /* build the blob */
        {
            size_t ranges = 0, off;
            struct rx_prog *pg;
            struct rx_class *cl;
            uint32_t *rg;
            for (i = 0; i < c.n_sets; i++) ranges += (size_t) c.sets[i].nr * 2;
            off = sizeof(struct rx_prog);
            pg = calloc(1, off + sizeof(uint32_t) * c.n_code + sizeof(struct rx_class) * c.n_sets + sizeof(uint32_t) * ranges + 16);
            pg->n_code = c.n_code;
            pg->code_off = (uint32_t) off;
            memcpy((char *) pg + off, c.code, sizeof(uint32_t) * c.n_code);
            off += sizeof(uint32_t) * c.n_code;
            pg->n_classes = c.n_sets;
            pg->class_off = (uint32_t) off;
            cl = (struct rx_class *) ((char *) pg + off);
            off += sizeof(struct rx_class) * c.n_sets;
            for (i = 0; i < c.n_sets; i++) {
                struct cset *s = &c.sets[i];
                int k;
                for (k = 0; k < 4; k++) cl[i].bits[k] = s->ascii[k];
                for (k = 4; k < 8; k++) cl[i].bits[k] = s->hi ? 0xffffffffu : 0;
                for (k = 0; k < c.n_raw; k++) if (c.raw_sets[k].idx == i) memcpy(cl[i].bits, c.raw_sets[k].b, sizeof(cl[i].bits));
                if (s->nr == 0) cl[i].mb_mode = RX_MB_NONE;
                else if (s->nr == 1 && s->r[0] == 0x80 && s->r[1] == 0x10FFFF) cl[i].mb_mode = RX_MB_ALL;
                else {
                    cl[i].mb_mode = RX_MB_RANGES;
                    cl[i].n_ranges = s->nr;
                    cl[i].ranges_off = (uint32_t) off;
                    rg = (uint32_t *) ((char *) pg + off);
                    memcpy(rg, s->r, sizeof(uint32_t) * 2 * s->nr);
                    off += sizeof(uint32_t) * 2 * s->nr;
                }
                /* "everything but a few ASCII bytes" ([^ ], [^"], [^\]] ...): the run is a search
                 * for the first stop byte, which the device does eight bytes at a time.  pad = number
                 * of stop bytes (1..4), ranges_off = the stop bytes (unused by RX_MB_ALL otherwise). */
                /* a plain set of <= 4 ASCII bytes (typically "what the continuation can start with"):
                 * pad bit 8, count in bits 9..11, the bytes in n_ranges (unused by RX_MB_NONE) */
                if (cl[i].mb_mode == RX_MB_NONE) {
                    uint32_t mem = 0, n = 0, b;
                    int okc = 1;
                    for (b = 0; b < 256 && okc; b++) {
                        if ((cl[i].bits[b >> 5] >> (b & 31)) & 1) {
                            if (b >= 0x80 || n >= 4) okc = 0;
                            else mem |= b << (8 * n++);
                        }
                    }
                    if (okc && n >= 1) { cl[i].pad = 0x100 | (n << 9); cl[i].n_ranges = mem; }
                }
                if (cl[i].mb_mode == RX_MB_ALL) {
                    uint32_t stops = 0, n = 0, b;
                    int okc = 1;
                    for (b = 0; b < 256 && okc; b++) {
                        if (!((cl[i].bits[b >> 5] >> (b & 31)) & 1)) {
                            if (b >= 0x80 || n >= 4) okc = 0;
                            else stops |= b << (8 * n++);
                        }
                    }
                    if (okc && n >= 1) {
                        for (b = n; b < 4; b++) stops |= (stops & 0xff) << (8 * b);
                        cl[i].pad = n; cl[i].ranges_off = stops;
                    }
                }
            }
            pg->total_bytes = (uint32_t) ((off + 15) & ~(size_t) 15);
            pg->n_groups = c.n_groups;
            pg->n_null = c.n_null;
            memset(&ff, 0, sizeof(ff));
            r = first_of(&c, root, &ff, 0);
            if (r & 1) pg->flags |= RX_F_NULLABLE;
            else { pg->flags |= RX_F_HAS_FIRSTSET; memcpy(pg->first, ff.b, sizeof(ff.b)); }
            if (c.ascii_only) pg->flags |= RX_F_ASCII_ONLY;
            /* (not for patterns with case folding: a folded first character may stand for two subject bytes) */
            if (!c.ascii_only && second_of(&c, root, &ff)) { pg->flags |= RX_F_HAS_SECONDSET; memcpy(pg->second, ff.b, sizeof(ff.b)); }
            la = leading_anchor(root);
            if (la == RX_BOL) pg->flags |= RX_F_ANCHOR_BOL;
            if (la == RX_BEGIN_BUF) pg->flags |= RX_F_ANCHOR_BUF;
            out->prog = pg;
        }
    }
    ok = 1;
done:
    for (i = 0; i < c.n_all; i++) { free(c.all[i]->kid); free(c.all[i]); }
    free(c.all);
    for (i = 0; i < c.n_sets; i++) cset_free(&c.sets[i]);
    free(c.sets);
    free(c.raw_sets);
    free(c.code);
    if (!ok) { rx_compiled_free(out); return -1; }
    return 0;
}
