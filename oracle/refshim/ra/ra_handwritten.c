/* Hand-written recursive-descent replacement for the flex/bison pair
 * src/record_accessor/ra.l:53-67 and ra.y:59-99 of the reference:
 *
 *   record_accessor := '$' IDENTIFIER ( '[' STRING ']' | '[' INTEGER ']' )*
 *   IDENTIFIER := [_A-Za-z][A-Za-z0-9_.\-/]*      STRING := '\'' ([^']|'')* '\''
 *   INTEGER := [1-9][0-9]*|0                        whitespace ignored
 *
 * It drives the reference's own flb_ra_parser_key_add / _subentry_add_* so the
 * rest of the record accessor is the unmodified reference code.
 * Test infrastructure only (part of oracle/_ref). */
#include <stdlib.h>
#include <string.h>
#include <ctype.h>
#include <fluent-bit/flb_info.h>
#include <fluent-bit/flb_mem.h>
#include <fluent-bit/flb_log.h>
#include <fluent-bit/record_accessor/flb_ra_parser.h>
#include "ra_parser.h"
#include "ra_lex.h"

struct ora_buf { const char *s; size_t pos; };
struct ora_scanner { struct ora_buf *cur; };

int flb_ra_lex_init(yyscan_t *scanner)
{
    struct ora_scanner *sc = calloc(1, sizeof(*sc));
    if (!sc) return -1;
    *scanner = sc;
    return 0;
}
int flb_ra_lex_destroy(yyscan_t scanner) { free(scanner); return 0; }
YY_BUFFER_STATE flb_ra__scan_string(const char *str, yyscan_t scanner)
{
    struct ora_scanner *sc = scanner;
    struct ora_buf *b = calloc(1, sizeof(*b));
    b->s = str; b->pos = 0; sc->cur = b;
    return b;
}
void flb_ra__delete_buffer(YY_BUFFER_STATE b, yyscan_t scanner) { (void) scanner; free(b); }

static void skip_ws(struct ora_buf *b)
{
    while (b->s[b->pos] == ' ' || b->s[b->pos] == '\t' || b->s[b->pos] == '\n') b->pos++;
}
static int is_id0(int c) { return c == '_' || isalpha(c); }
static int is_idn(int c) { return c == '_' || c == '.' || c == '-' || c == '/' || isalnum(c); }

int flb_ra_parse(struct flb_ra_parser *rp, const char *query, void *scanner)
{
    struct ora_scanner *sc = scanner;
    struct ora_buf *b = sc->cur;
    size_t st, n, i, j;
    char *tmp;
    void *key;
    int subkeys = 0;
    (void) query;

    skip_ws(b);
    if (b->s[b->pos] != '$') goto syntax;
    b->pos++;
    skip_ws(b);
    if (!is_id0((unsigned char) b->s[b->pos])) goto syntax;
    st = b->pos;
    while (is_idn((unsigned char) b->s[b->pos])) b->pos++;
    tmp = flb_malloc(b->pos - st + 1);
    memcpy(tmp, b->s + st, b->pos - st);
    tmp[b->pos - st] = '\0';

    /* bison reduces the subkeys before the record_key action: collect them first */
    struct { int is_str; char *s; int id; } sk[64];
    for (;;) {
        skip_ws(b);
        if (b->s[b->pos] != '[') break;
        b->pos++;
        skip_ws(b);
        if (subkeys >= 64) { flb_free(tmp); goto syntax; }
        if (b->s[b->pos] == '\'') {
            st = ++b->pos;
            for (;;) {
                if (b->s[b->pos] == '\0') { flb_free(tmp); goto syntax; }
                if (b->s[b->pos] == '\'') {
                    if (b->s[b->pos + 1] == '\'') { b->pos += 2; continue; }
                    break;
                }
                b->pos++;
            }
            n = b->pos - st;
            char *str = flb_malloc(n + 1);
            for (i = 0, j = 0; i < n; i++, j++) {
                str[j] = b->s[st + i];
                if (b->s[st + i] == '\'') i++;
            }
            str[j] = '\0';
            b->pos++;
            sk[subkeys].is_str = 1; sk[subkeys].s = str; subkeys++;
        }
        else if (isdigit((unsigned char) b->s[b->pos])) {
            st = b->pos;
            if (b->s[b->pos] == '0') b->pos++;
            else while (isdigit((unsigned char) b->s[b->pos])) b->pos++;
            sk[subkeys].is_str = 0; sk[subkeys].s = NULL;
            sk[subkeys].id = atoi(b->s + st); subkeys++;
        }
        else { flb_free(tmp); goto syntax; }
        skip_ws(b);
        if (b->s[b->pos] != ']') { flb_free(tmp); goto syntax; }
        b->pos++;
    }
    for (i = 0; i < (size_t) subkeys; i++) {
        if (sk[i].is_str) { flb_ra_parser_subentry_add_string(rp, sk[i].s); flb_free(sk[i].s); }
        else flb_ra_parser_subentry_add_array_id(rp, sk[i].id);
    }
    rp->type = FLB_RA_PARSER_KEYMAP;
    key = flb_ra_parser_key_add(rp, tmp);
    if (key) rp->key = key;
    flb_free(tmp);
    skip_ws(b);
    if (b->s[b->pos] != '\0') goto syntax;
    return 0;
syntax:
    flb_error("[record accessor] syntax error at '%s'", b->s);
    return 1;
}
