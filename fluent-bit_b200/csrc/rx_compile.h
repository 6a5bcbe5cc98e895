/* rx_compile.h -- host-side regex compiler: Ruby-syntax (Onigmo-compatible subset)
 * pattern text -> relocatable struct rx_prog for the device matcher.
 *
 * Replaces, for the GPU path, what flb_regex_create() obtains from onig_new()
 * (reference src/flb_regex.c:60-180, lib/onigmo/regcomp.c:5876).  Constructs the
 * device matcher cannot reproduce exactly are REJECTED here, at create time
 * (there is no CPU fallback): look-behind, \G \K \R \X, \p{...}, (?~...),
 * (?(cond)...), \g<...> calls, numbered back references next to named groups.
 */
#ifndef FLBGPU_RX_COMPILE_H
#define FLBGPU_RX_COMPILE_H

#include "flbgpu_prog.h"

#ifdef __cplusplus
extern "C" {
#endif

struct rx_name {
    char *name;          /* group name */
    int   n_groups;      /* >1 when the name is defined more than once */
    int   groups[8];     /* capture group numbers, definition order */
};

struct rx_compiled {
    struct rx_prog *prog;    /* malloc'd; prog->total_bytes long */
    int n_names;             /* named groups, in first-definition order -- the order
                                onig_foreach_name() visits them (regparse.c:582-597) */
    struct rx_name *names;
    char err[160];
};

/* pattern: the text a Fluent Bit config gives, including the optional /.../imx
 * wrapper (src/flb_regex.c:60-152).  Returns 0 or -1 (out->err says why). */
int  rx_compile(const char *pattern, struct rx_compiled *out);
void rx_compiled_free(struct rx_compiled *c);

#ifdef __cplusplus
}
#endif
#endif
